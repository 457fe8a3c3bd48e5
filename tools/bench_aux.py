"""Secondary measurements (BASELINE.json configs 3 and 4; not the bench.py headline): FGSM input-gradient steps/s and
ELBO training steps/s in FP32 mode (CUDA-core kernels + autograd) or FAST mode (tcgen05 forward / dgrad / wgrad), single GPU or batch-sharded under torchrun with the NCCL gradient all-reduce.
usage: [torchrun ...] python tools/bench_aux.py [--variant brats|hippocampus] [--batch B] [--steps K] [--mode fp32|fast]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import supernet_b200 as S
from supernet_b200 import dp
from oracle import supernet_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="brats")
ap.add_argument("--batch", type=int, default=8, help="slices per GPU")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--mode", default="fp32", choices=["fp32", "fast"])
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
C, in_ch = (4, 4) if args.variant == "brats" else (3, 1)
hw = O.output_hw(args.variant)
w = O.make_weights(args.variant, 32, C, in_ch)
model = S.Density_prop_with_pad_UNET(32, C, variant=args.variant, mode=args.mode).load_weight_dict(w, device=dev)
B = args.batch
alpha = O.BRATS_ALPHA if args.variant == "brats" else 1.0
x = O.make_input(args.variant, B, seed=2025 + rank, alpha=alpha).to(dev)
y = O.make_labels(B, hw * hw, C, seed=7 + rank).to(dev)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


fgsm_ms = timed(lambda: S.create_adversarial_pattern(model, x, y))
trainer = dp.DataParallelTrainer(model, lr=1e-4, kl_factor=1e-5)
train_ms = timed(lambda: trainer.step(x, y, global_batch=B * world))
# replicas must stay bit-identical after sharded training steps (same all-reduced gradient on every rank)
chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum()
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool(lo == hi)
else:
    in_sync = True
if rank == 0:
    print(json.dumps({"variant": args.variant, "n_gpus": world, "batch_per_gpu": B, "mode": args.mode,
                      "fgsm_ms_per_step": round(fgsm_ms, 3), "fgsm_slices_per_s": round(B * world / fgsm_ms * 1e3, 1),
                      "train_ms_per_step": round(train_ms, 3),
                      "train_slices_per_s": round(B * world / train_ms * 1e3, 1), "replicas_in_sync": in_sync}))
if world > 1:
    dist.destroy_process_group()
