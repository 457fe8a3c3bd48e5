timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -x -q > gpurun_out/r02_t23.log 2>&1; tail -n 4 gpurun_out/r02_t23.log
for rep in 1 2; do
for m in 0 1 2; do
SN_CTA2=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b23_$m.json 2> gpurun_out/r02_b23.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b23_$m.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("CTA2=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:k[n] for n in ("conv4","conv5","conv6","conv7","conv8","conv9","up1_conv2x2","up1_conv1","up1_conv2","up2_conv2x2","up2_conv1","up2_conv2","up3_conv2x2","up4_conv2x2")}, "clk", d["clocks"]["sm_mhz"])
PY
done
done
