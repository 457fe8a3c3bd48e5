export SN_CTA2_VERBOSE=1
for L in conv5 conv7 up2_conv1; do
echo -n "CTA2: "; SN_CTA2=2 python tools/profile_layer.py $L 64 2>&1 | tail -n 2
echo -n "single: "; SN_CTA2=0 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
done
