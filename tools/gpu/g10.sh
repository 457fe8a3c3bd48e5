for rep in 1 2; do
for m in 0 1 2; do
SN_KWC=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b10_$m.json 2> gpurun_out/r02_b10.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b10_$m.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("KWC=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], "conv1", k["conv1"], "up4_conv1", k["up4_conv1"], "up4_conv2", k["up4_conv2"], "conv5", k["conv5"], "clk", d["clocks"]["sm_mhz"])
PY
done
done
