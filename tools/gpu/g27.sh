export SN_BUILD_KNOBS=1
for d in 0 7; do
  echo -n "CTA2 conv5 dbg=$d: "; SN_CTA2=2 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
echo -n "CTA2 conv5 sa=2 (sb 12): "; SN_CTA2=2 SN_CTA2_SA=2 python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
for L in conv7 up2_conv1 up1_conv1 conv4; do
echo -n "CTA2 $L: "; SN_CTA2=2 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
echo -n "single $L: "; SN_CTA2=0 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
done
