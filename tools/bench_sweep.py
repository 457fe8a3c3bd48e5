"""BASELINE.json configs[4]: throughput sweep, GLOBAL batch 1-512 BraTS slices over N GPUs (per-GPU batch =
ceil(global / N)), full mean + variance output, FAST mode, CUDA-graph replay, batch resident; CUDA events
between barriers, max over ranks.  Ranks whose shard is empty idle (global < N).
usage: [torchrun --nproc-per-node N] tools/bench_sweep.py > profiles/rNN_sweep_nN.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import supernet_b200 as S
from supernet_b200 import dp
from supernet_b200.engine import InferenceEngine
from oracle import supernet_oracle as O

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model = S.Density_prop_with_pad_UNET(32, 4, variant="brats", mode="fast")
model.load_weight_dict(O.make_weights("brats", 32, 4, 4), device=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for gb in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    a, b = dp.shard_bounds(gb, world, rank)
    n = b - a
    eng = None
    if n > 0:
        eng = InferenceEngine(model, n, 204, 204, 4, dev, graph=True, keep_presoftmax=False)
        eng.x_in.copy_(O.make_input("brats", n, seed=2025 + rank, alpha=O.BRATS_ALPHA))
        for _ in range(5):
            eng.forward_resident()
    barrier()
    steps = 100 if gb <= 64 else 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if eng is not None:
        for _ in range(steps):
            eng.forward_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "global_batch": gb, "per_gpu_batch": -(-gb // world), "ms_per_step": round(ms, 4),
                          "slices_per_s": round(gb / ms * 1e3, 1)}), flush=True)
    del eng
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
