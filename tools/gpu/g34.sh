set -x
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t34.log 2>&1; tail -n 5 gpurun_out/r02_t34.log
for rep in 1 2; do
for m in "1 1" "1 0" "0 1"; do
set -- $m
SN_FUSE_HEAD=$1 SN_FIRST_V8=$2 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b34.json 2> gpurun_out/r02_b34.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b34.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("FUSE=$1 V8=$2 rep $rep value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:v for n,v in k.items() if n in ("conv_input","conv1","up4_conv1","up4_conv2","conv_final","up4_conv2+conv_final")}, "clk", d["clocks"]["sm_mhz"])
PY
done
done
