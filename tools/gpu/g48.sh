set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"]["value"], "clk", d["clocks"], "fgsm", d["aux"]["fgsm"], "train", d["aux"]["train"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
print("N=4 value", d["value"], "e2e", d["e2e"]["value"], "clk", d["clocks"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"], d["aux"]["train"].get("replicas_in_sync"))
PY
