set -x
for l in conv3 up3_conv1; do SN_CTA2=0 python tools/profile_layer.py $l 64; python tools/profile_layer.py $l 64; done
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -k "test_conv_tc_f32_dst or packed_window" 2>&1 | tail -n 3
