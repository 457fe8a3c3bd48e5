"""Raw pinned-memory DMA probe for the host-to-host path: every rank copies the bench's per-step volumes (42.6 MB in,
70.9 MB out at batch 64) H2D and D2H on two streams, concurrently, with NO kernels in between.  Settles whether the
aggregate the e2e path reaches at 8 ranks (118 GB/s) is the box or the pipeline.
usage: torchrun --nproc-per-node N tools/probe_host_dma.py   (or plain python for one GPU)"""
import json
import os
import sys

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
IN_B, OUT_B = 64 * 204 * 204 * 4 * 4, 2 * 64 * 186 * 186 * 4 * 4
h_in = torch.empty(IN_B, dtype=torch.uint8).pin_memory()
h_out = torch.empty(OUT_B, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN_B, dtype=torch.uint8, device=dev)
d_out = torch.empty(OUT_B, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    def step():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    a.record()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    n = 50
    for _ in range(n):
        step()
    cur.wait_stream(s1)
    cur.wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    byts = (IN_B if mode != "d2h" else 0) + (OUT_B if mode != "h2d" else 0)
    res[mode] = {"ms_per_step": round(ms / n, 3), "per_rank_GBs": round(byts * n / ms / 1e6, 1),
                 "aggregate_GBs": round(world * byts * n / ms / 1e6, 1),
                 "steps_per_s_bound": round(n / ms * 1e3, 1), "slices_per_s_bound": round(world * 64 * n / ms * 1e3)}
if rank == 0:
    print(json.dumps({"ranks": world, "h2d_bytes": IN_B, "d2h_bytes": OUT_B, **res}))
if world > 1:
    dist.destroy_process_group()
