python tools/grad_error_split.py 2025 2026 > gpurun_out/r02_grad_split.md 2>&1; cat gpurun_out/r02_grad_split.md
SN_KWC=0 python tools/grad_error_split.py 2025 > gpurun_out/r02_grad_split_nokwc.md 2>&1; cat gpurun_out/r02_grad_split_nokwc.md
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench9.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"])
    print([(k["name"],k["ms"]) for k in d["kernels"]])
PY
