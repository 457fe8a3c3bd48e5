timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r02_t7.log 2>&1; tail -n 5 gpurun_out/r02_t7.log
for L in conv1 up4_conv2 up4_conv1; do
echo -n "tma: "; python tools/profile_layer.py $L 64 2>&1 | tail -n 1
echo -n "direct: "; SN_TMA_STORE=0 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
done
for d in 1 2 4 8; do
  echo -n "KWC tma dbg=$d: "; SN_HL_DBG=$d python tools/profile_layer.py conv1 64 2>&1 | tail -n 1
done
