set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29558 tools/bench_sweep.py > gpurun_out/r02_sweep_n4.jsonl 2> gpurun_out/r02_sweep_n4.err; echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29559 tools/bench_sweep.py > gpurun_out/r02_sweep_n2.jsonl 2> gpurun_out/r02_sweep_n2.err; echo "rc=$?"
cat gpurun_out/r02_sweep_n4.jsonl gpurun_out/r02_sweep_n2.jsonl
