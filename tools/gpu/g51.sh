set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "first_conv" 2>&1 | tail -n 5
for m in "1 1" "1 0" "0 0"; do
set -- $m
SN_FIRST_WS=$1 SN_FIRST_BULK=$2 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b51.json 2> gpurun_out/r02_b51.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b51.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("WS=$1 BULK=$2 value", d["value"], "e2e", d["e2e"]["value"], {n:v for n,v in k.items() if n in ("conv_input","conv1")}, "clk", d["clocks"]["sm_mhz"])
PY
done
