set -x
for l in conv1 up4_conv2; do python tools/profile_layer.py $l 64; SN_KWC=2 python tools/profile_layer.py $l 64; done
python tools/profile_layer.py conv9 64; python tools/profile_layer.py conv9 64 im2col
python tools/profile_layer.py conv7 64; python tools/profile_layer.py conv7 64 im2col
SN_KWC=2 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b40.json 2> gpurun_out/r02_b40.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b40.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("KWC=2 value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:v for n,v in k.items() if n in ("conv_input","conv1","up4_conv1","up4_conv2","conv_final","up4_conv2+conv_final")}, "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"])
PY
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_b40.json 2> gpurun_out/r02_b40.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_b40.json").read().strip().splitlines()[-1])
k={r["name"]:r["ms"] for r in d["kernels"]}
print("default value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], {n:v for n,v in k.items() if n in ("conv_input","conv1","up4_conv1","up4_conv2","conv_final","up4_conv2+conv_final")}, "clk", d["clocks"]["sm_mhz"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"])
PY
