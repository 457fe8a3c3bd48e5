"""FAST-mode FGSM step (BASELINE.json configs[2]): forward + 0.5 NLL + input-gradient chain on the tensor cores
(engine.GradientEngine), resident batch, CUDA-graph replay; per-launch CUDA-event table; optional error study
against the FP32-mode gradient.  usage: python tools/bench_fgsm_fast.py [--batch B] [--steps K] [--study]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import supernet_b200 as S
from supernet_b200.engine import GradientEngine
from oracle import supernet_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="brats")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--study", action="store_true")
ap.add_argument("--train", action="store_true", help="time the training chain (weight gradients) instead")
args = ap.parse_args()
C, in_ch, hw_in = (4, 4, 204) if args.variant == "brats" else (3, 1, 64)
hw = O.output_hw(args.variant)
alpha = O.BRATS_ALPHA if args.variant == "brats" else 1.0
w = O.make_weights(args.variant, 32, C, in_ch)
model = S.Density_prop_with_pad_UNET(32, C, variant=args.variant, mode="fast").load_weight_dict(w, device="cuda")
B = args.batch
eng = GradientEngine(model, B, hw_in, hw_in, in_ch, "cuda", graph=True, train=args.train)
if args.train:
    eng._loss_scale, eng._clip = 1.0, (1e-12, 1e3)
eng.x_in.copy_(O.make_input(args.variant, B, alpha=alpha))
eng.y_in.copy_(O.make_labels(B, hw * hw, C))
for _ in range(3):
    eng.loss_and_input_gradient_resident()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    eng.loss_and_input_gradient_resident()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
# per-launch table (eager launches, CUDA events around each)
rows = []
names = list(eng.step_names) + ["nll_fwd"] + list(eng.bwd_step_names)
fns = list(eng._steps) + [lambda: S.fastops.nll_gaussian_fwd(eng.y_in, eng.p, eng.v, eng._clip, eng.nll_acc,
                                                            eng.nll_loss)] + list(eng._bwd_steps)
for name, fn in zip(names, fns):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    rows.append({"name": name, "ms": round(e0.elapsed_time(e1) / 5, 4)})
fwd_ms = sum(r["ms"] for r in rows[:len(eng.step_names)])
out = {"variant": args.variant, "batch": B, "chain": "train (fwd + dgrad + wgrad)" if args.train else "fgsm (fwd + dgrad)", "fgsm_ms_per_step": round(ms, 4),
       "fgsm_slices_per_s": round(B / ms * 1e3, 1), "launches": len(names) + 1,
       "forward_ms_sum": round(fwd_ms, 4), "backward_ms_sum": round(sum(r["ms"] for r in rows) - fwd_ms, 4),
       "kernels": rows}
print(json.dumps(out))
if args.study:
    Bs = 1
    x = O.make_input(args.variant, Bs, alpha=alpha).cuda()
    y = O.make_labels(Bs, hw * hw, C).cuda()
    slow = S.Density_prop_with_pad_UNET(32, C, variant=args.variant, mode="fp32").load_weight_dict(w, device="cuda")
    _, g32 = S.create_adversarial_pattern(slow, x, y)
    _, gf = S.create_adversarial_pattern(model, x, y)
    oracle = O.UNetOracle(args.variant, 32, C, in_ch, torch.float64)
    g_ref, _ = oracle.fgsm_gradient(x.cpu(), y.cpu().double())
    print(json.dumps({"study": "input gradient, batch 1", "fp32_vs_oracle": O.rel_l2(g32.cpu(), g_ref),
                      "fast_vs_oracle": O.rel_l2(gf.cpu(), g_ref), "fast_vs_fp32": O.rel_l2(gf.cpu(), g32.cpu())}))
