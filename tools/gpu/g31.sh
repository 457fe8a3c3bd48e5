for rep in 1 2; do
for m in 0 1; do
SN_CTA2=$m timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-aux > gpurun_out/r02_b31_$m.json 2> gpurun_out/r02_b31.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b31_$m.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("CTA2=$m rep $rep value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:k[n] for n in ("conv3","conv4","conv5","conv6","conv7","conv8","conv9","up1_conv2x2","up1_conv1","up1_conv2","up2_conv2x2","up2_conv1","up2_conv2","up3_conv2x2","up3_conv1","up3_conv2","up4_conv2x2")}, "clk", d["clocks"]["sm_mhz"])
PY
done
done
