"""Run one FAST-mode conv layer (BraTS shapes) alone, L2 flushed before each launch; prints CUDA-event times.
usage: profile_layer.py <layer> <batch> [im2col|rsum]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from supernet_b200 import fastops as F

LAYERS = {  # name: (H, cin0, cin1, cout, k, upconv)
    "conv1": (202, 32, 0, 32, 3, False), "conv2": (100, 32, 0, 64, 3, False), "conv3": (98, 64, 0, 64, 3, False),
    "conv5": (46, 128, 0, 128, 3, False), "conv7": (20, 256, 0, 256, 3, False), "conv9": (8, 512, 0, 512, 3, False),
    "up2_conv1": (42, 128, 128, 128, 3, False), "up3_conv1": (90, 64, 64, 64, 3, False),
    "up4_conv1": (186, 32, 32, 32, 3, False), "up4_conv2": (188, 32, 0, 32, 3, False),
    "up4_conv2x2": (90, 64, 0, 32, 2, True), "up3_conv2x2": (42, 128, 0, 64, 2, True),
}
name = sys.argv[1]
B = int(sys.argv[2])
im2col = len(sys.argv) > 3 and sys.argv[3] == "im2col"
want_rsum = len(sys.argv) > 3 and sys.argv[3] == "rsum"        # training forward: also emit the rank-1 statistic
H, c0, c1, cout, k, up = LAYERS[name]
g = torch.Generator(device="cuda").manual_seed(0)
src0 = torch.randn((B, H, H, 3, c0), device="cuda", generator=g).bfloat16().abs()
src1 = torch.randn((B, H, H, 3, max(c1, 32)), device="cuda", generator=g).bfloat16().abs()
w = torch.randn((k, k, c0 + c1, cout), device="cuda", generator=g) * 0.1
ws = torch.full((cout,), -5.0, device="cuda")
wp, s = F.prepare_weights(w, ws, upconv=up)
Ho = 2 * H if up else H - k + 1
out = F.packed_empty(B, Ho, Ho, cout, "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
rs = torch.empty((B, H, H) if up else (B, Ho, Ho), device="cuda") if want_rsum else None


def run():
    F.conv_moments_tc(F.PackedView(src0), c0, B, H, H, k, cout, wp, s, dst=F.PackedView(out), relu=not up, upconv=up,
                      src1=F.PackedView(src1) if c1 else None, c1=c1, im2col=im2col, rsum_out=rs)


run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    if not os.environ.get('NO_FLUSH'):
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
inb = B * H * H * (c0 + c1) * 6
outb = B * Ho * Ho * cout * 6
t = sorted(ts)[len(ts) // 2] * 1e-3
print(f"{name} B={B} {'im2col' if im2col else 'halo'}: {t*1e3:.4f} ms  {(inb+outb)/t/1e9:.0f} GB/s  in {inb/1e6:.0f} MB out {outb/1e6:.0f} MB")
