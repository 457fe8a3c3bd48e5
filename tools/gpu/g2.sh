set -x
for L in conv1 up4_conv1; do
python tools/profile_layer.py $L 64 > gpurun_out/r02_layer_${L}_kwc.txt 2>&1
SN_KWC=0 python tools/profile_layer.py $L 64 > gpurun_out/r02_layer_${L}_nokwc.txt 2>&1
done
cat gpurun_out/r02_layer_*.txt
ncu --set full --clock-control none --import-source on -k regex:conv_moments_halo -s 2 -c 1 -f -o gpurun_out/r02_conv1_kwc python tools/profile_layer.py conv1 64 > gpurun_out/ncu_a.log 2>&1
SN_KWC=0 ncu --set full --clock-control none --import-source on -k regex:conv_moments_halo -s 2 -c 1 -f -o gpurun_out/r02_conv1_nokwc python tools/profile_layer.py conv1 64 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
