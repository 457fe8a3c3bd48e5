timeout 1500 python -m pytest tests/test_gpu_layers_fast.py -x -q -s > gpurun_out/r02_t11a.log 2>&1; tail -n 25 gpurun_out/r02_t11a.log
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r02_t11.log 2>&1; tail -n 25 gpurun_out/r02_t11.log
