set -x
timeout 900 python -m pytest tests/test_gpu_layers_fast.py tests/test_gpu_drivers.py -q -x 2>&1 | tail -n 4
python __graft_entry__.py smoke 2>&1 | tail -n 1
