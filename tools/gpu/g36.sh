set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r02_t36.log 2>&1; tail -n 3 gpurun_out/r02_t36.log
for l in conv3 up3_conv1 conv3 up3_conv1; do python tools/profile_layer.py $l 64; done > gpurun_out/r02_layers36.txt 2>&1; cat gpurun_out/r02_layers36.txt
