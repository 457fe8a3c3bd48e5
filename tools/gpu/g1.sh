set -x
tools/bin/probe_mma_rate > gpurun_out/r02_probe_mma_rate.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -x -q > gpurun_out/r02_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t1.log
tail -15 gpurun_out/r02_t1.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
SN_KWC=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r02_bench1_nokwc.json 2> gpurun_out/r02_bench1_nokwc.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench1.json","gpurun_out/r02_bench1_nokwc.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("aux"))
        print([(k["name"],k["ms"]) for k in d["kernels"]])
    except Exception as e:
        print(f, "ERR", e)
PY
