set -x
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -q -x -k "cta_pair" 2>&1 | tail -n 5
for l in conv3 conv2; do SN_CTA2=0 python tools/profile_layer.py $l 64; python tools/profile_layer.py $l 64; SN_CTA2_64=2 python tools/profile_layer.py $l 64; done
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t42.log 2>&1; tail -n 3 gpurun_out/r02_t42.log
for m in 1 2; do
SN_CTA2_64=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b42.json 2> gpurun_out/r02_b42.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b42.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("CTA2_64=$m value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:v for n,v in k.items() if n in ("conv2","conv3","up3_conv1","up3_conv2")}, "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"])
PY
done
