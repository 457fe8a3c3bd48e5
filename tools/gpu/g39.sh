set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "streaming or pipeline or inplace" 2>&1 | tail -n 2
for rep in 1 2 3; do
for m in 1 0; do
SN_PIPE_SERIAL=$m timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b39.json 2> gpurun_out/r02_b39.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b39.json").read().strip().splitlines()[-1])
print("SERIAL=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], "ratio", round(d["e2e"]["value"]/d["value"],4), "clk", d["clocks"]["sm_mhz"])
PY
done
done
