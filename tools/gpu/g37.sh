set -x
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t37.log 2>&1; tail -n 3 gpurun_out/r02_t37.log
for l in conv2 conv2; do python tools/profile_layer.py $l 64; done
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b37.json 2> gpurun_out/r02_b37.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b37.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"], "train", d["aux"]["train"])
PY
python tools/bench_fgsm_fast.py --train --batch 64 > gpurun_out/r02_train37.json 2> gpurun_out/r02_train37.err; tail -c 3000 gpurun_out/r02_train37.json
