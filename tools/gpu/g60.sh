set -x
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_n1_rep2.json 2> gpurun_out/r02_bench_n1_rep2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1_rep2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "clk", d["clocks"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"])
PY
