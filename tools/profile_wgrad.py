"""Run one FAST-mode weight-gradient layer (BraTS shapes) alone; prints CUDA-event times.
usage: profile_wgrad.py <layer> <batch> [im2col]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S  # noqa: F401
from supernet_b200 import fastops as F

LAYERS = {  # name: (H, cin0, cin1, cout)
    "conv1": (202, 32, 0, 32), "conv2": (100, 32, 0, 64), "conv3": (98, 64, 0, 64), "conv5": (46, 128, 0, 128),
    "conv7": (20, 256, 0, 256), "up2_conv1": (48, 128, 128, 128), "up3_conv1": (90, 64, 64, 64),
    "up4_conv1": (186, 32, 32, 32), "up4_conv2": (188, 32, 0, 32),
}
name = sys.argv[1]
B = int(sys.argv[2])
im2col = len(sys.argv) > 3 and sys.argv[3] == "im2col"
H, c0, c1, cout = LAYERS[name]
k = 3
g = torch.Generator(device="cuda").manual_seed(0)
src0 = torch.randn((B, H, H, 3, c0), device="cuda", generator=g).bfloat16().abs()
src1 = torch.randn((B, H, H, 3, max(c1, 32)), device="cuda", generator=g).bfloat16().abs()
gout = torch.randn((B, H - 2, H - 2, 3, cout), device="cuda", generator=g).bfloat16()
w = torch.randn((k, k, c0 + c1, cout), device="cuda", generator=g) * 0.1
ws = torch.full((cout,), -5.0, device="cuda")
rsum = torch.rand((B, H - 2, H - 2), device="cuda")
gw, gws = torch.empty_like(w), torch.empty_like(ws)
work = F.wgrad_workspace(k, c0 + c1, cout, "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def run():
    F.conv_moments_bwd_weight_tc(F.PackedView(gout), B, H, H, k, cout, F.PackedView(src0), c0, rsum, w, ws, work, gw, gws,
                                 in1=F.PackedView(src1) if c1 else None, c1=c1, im2col=im2col)


run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
flops = 2 * 2 * B * (H - 2) ** 2 * 9 * (c0 + c1) * cout
print(f"{name} B={B} {'im2col' if im2col else 'rows'}: min {min(ts):.4f} ms (GEMM + dsigma + finalize), "
      f"{flops / min(ts) / 1e9:.1f} TFLOP/s algorithmic")
