set -x
SN_FIRST_WS=1 timeout 900 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -n 1
SN_FUSE_HEAD=0 SN_CTA2_64=0 timeout 900 python -m pytest tests/ -q -m gpu -x --deselect tests/test_gpu_tc.py::test_fused_head_engine_is_bit_identical_to_two_kernels --deselect tests/test_gpu_layers_fast.py::test_layerwise_fast_forward_equals_the_engine 2>&1 | tail -n 1
SN_CTA2=0 SN_KWC=0 timeout 900 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -n 1
