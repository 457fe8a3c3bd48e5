set -x
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t45.log 2>&1; tail -n 3 gpurun_out/r02_t45.log
python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b45.json 2> gpurun_out/r02_b45.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b45.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"], "train", d["aux"]["train"]["slices_per_s"])
PY
