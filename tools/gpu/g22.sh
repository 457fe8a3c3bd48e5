timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "cta_pair" > gpurun_out/r02_t22.log 2>&1; tail -n 15 gpurun_out/r02_t22.log
