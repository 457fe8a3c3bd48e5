set -x
for l in conv1 up4_conv2; do python tools/profile_layer.py $l 64; python tools/profile_layer.py $l 64 rsum; done
for i in 1 2; do python tools/bench_fgsm_fast.py --train --batch 64 > gpurun_out/r02_train38_$i.json 2> gpurun_out/r02_train38.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_train38_$i.json").read().strip().splitlines()[0])
print(d["fgsm_ms_per_step"], {r["name"]:r["ms"] for r in d["kernels"] if r["name"] in ("conv1","up4_conv1","up4_conv2","conv_final","conv2")})
PY
done
