timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r02_t17.log 2>&1; tail -n 3 gpurun_out/r02_t17.log
for rep in 1 2; do
for m in 1 2; do
SN_DUAL=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b17_$m.json 2> gpurun_out/r02_b17.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b17_$m.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("DUAL=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], {n:k[n] for n in ("conv3","up3_conv1","up3_conv2","conv5","conv7","up1_conv1","up2_conv1","up2_conv2")}, "clk", d["clocks"]["sm_mhz"])
PY
done
done
