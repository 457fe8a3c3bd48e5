set -x
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t43.log 2>&1; tail -n 3 gpurun_out/r02_t43.log
for l in conv1 conv3 up3_conv1 up4_conv1 up4_conv2 conv5 up2_conv1; do python tools/profile_layer.py $l 64; done
for m in 1 2; do
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b43.json 2> gpurun_out/r02_b43.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b43.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("rep $m value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], k, "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"], d["roofline"]["per_launch_bound"])
PY
done
