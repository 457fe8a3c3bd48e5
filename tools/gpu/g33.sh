set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x -k "head" > gpurun_out/r02_t33a.log 2>&1; tail -n 5 gpurun_out/r02_t33a.log
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r02_t33.log 2>&1; tail -n 5 gpurun_out/r02_t33.log
for l in conv1 up4_conv1 up4_conv2 conv3 up3_conv1; do python tools/profile_layer.py $l 64; done > gpurun_out/r02_layers33.txt 2>&1; cat gpurun_out/r02_layers33.txt
for rep in 1 2; do
for m in 1 0; do
SN_FUSE_HEAD=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b33_$m.json 2> gpurun_out/r02_b33.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b33_$m.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("FUSE=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], k, "clk", d["clocks"]["sm_mhz"])
PY
done
done
