timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_t8.log 2>&1; tail -n 12 gpurun_out/r02_t8.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench8.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("aux"))
        print([(k["name"],k["ms"]) for k in d["kernels"]])
    except Exception as e:
        print(f, "ERR", e)
PY
