set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/probe_host_dma.py > gpurun_out/r02_probe_dma_n8.json 2> gpurun_out/r02_probe_dma_n8.err; cat gpurun_out/r02_probe_dma_n8.json
python tools/probe_host_dma.py > gpurun_out/r02_probe_dma_n1.json 2>&1; cat gpurun_out/r02_probe_dma_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"], "aux", d["aux"])
PY
