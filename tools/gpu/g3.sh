set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r02_t4.log 2>&1; tail -3 gpurun_out/r02_t3.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"
SN_KWC=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r02_bench4_nokwc.json 2> gpurun_out/r02_bench4_nokwc.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench4.json","gpurun_out/r02_bench4_nokwc.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"])
        print([(k["name"],k["ms"]) for k in d["kernels"]])
    except Exception as e:
        print(f, "ERR", e)
PY
