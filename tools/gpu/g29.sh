export SN_BUILD_KNOBS=1
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -x -q -k "cta_pair or cta2" > gpurun_out/r02_t29.log 2>&1; tail -n 3 gpurun_out/r02_t29.log
for d in 0 7; do
  echo -n "CTA2 conv5 dbg=$d: "; SN_CTA2=2 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
for L in conv7 up2_conv1; do
echo -n "CTA2 $L: "; SN_CTA2=2 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
echo -n "single $L: "; SN_CTA2=0 python tools/profile_layer.py $L 64 2>&1 | tail -n 1
done
