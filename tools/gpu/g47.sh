set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], "clk", d["clocks"], "fgsm", d["aux"]["fgsm"], "train", d["aux"]["train"])
PY
