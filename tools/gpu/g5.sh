set -x
python tools/profile_layer.py conv1 64 > gpurun_out/r02_layer_conv1_kwc2.txt 2>&1
cat gpurun_out/r02_layer_conv1_kwc2.txt
ncu --set full --clock-control none --import-source on -k regex:conv_moments_halo -s 2 -c 1 -f -o gpurun_out/r02_conv1_kwc2 python tools/profile_layer.py conv1 64 > gpurun_out/ncu_a.log 2>&1
tail -n 3 gpurun_out/ncu_a.log
