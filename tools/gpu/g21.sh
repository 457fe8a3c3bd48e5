for rep in 1 2; do
for d in 2 3 4; do
SN_PIPE_DEPTH=$d timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b21_$d.json 2> gpurun_out/r02_b21.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b21_$d.json").read().strip().splitlines()[-1])
print("depth=$d rep $rep value", d["value"], "e2e", d["e2e"]["value"], "ratio", round(d["e2e"]["value"]/d["value"],3), "clk", d["clocks"]["sm_mhz"])
PY
done
done
