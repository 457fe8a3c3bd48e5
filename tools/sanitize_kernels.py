"""Tiny invocations of every warp-specialised kernel, to be run under compute-sanitizer:
    compute-sanitizer --tool {memcheck,racecheck,synccheck} python tools/sanitize_kernels.py
Shapes are small (the sanitizer serialises and instruments every access) but cover each template family: halo forward
(tap-shift NT 32 / 64 / 128, resident and streamed weights, G = 2 pairs, up-conv, kw-concatenated with direct and TMA
stores), halo data gradient (both variants, up-conv), both weight-gradient kernels, first conv (tensor-core), pooling,
head forward / backward.  Results are also compared with nothing: correctness is the parity tests' job."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S

F = S.fastops
g = torch.Generator(device="cuda").manual_seed(0)


def packed(B, H, W, c):
    return F.PackedView((torch.randn((B, H, W, 3, c), device="cuda", generator=g) * 0.5).bfloat16().abs().contiguous())


def layer(cin, cout, k, upconv=False):
    w = torch.randn((k, k, cin, cout), device="cuda", generator=g) * 0.1
    ws = torch.full((cout,), -4.0, device="cuda")
    return w, ws, F.prepare_weights(w, ws, upconv=upconv), F.prepare_weights_bwd(w, upconv=upconv)


def fwd(B, H, W, cin, cout, k, upconv=False, kwc=None, c1=0):
    w, ws, (wp, s), _ = layer(cin + c1, cout, k, upconv)
    Ho, Wo = (2 * H, 2 * W) if upconv else (H - k + 1, W - k + 1)
    out = F.PackedView(F.packed_empty(B, Ho, Wo, cout, "cuda"))
    F.conv_moments_tc(packed(B, H, W, cin), cin, B, H, W, k, cout, wp, s, dst=out, relu=not upconv, upconv=upconv,
                      src1=packed(B, H, W, c1) if c1 else None, c1=c1, kwc=kwc)


def bwd(B, H, W, cin, cout, k, upconv=False, kwc=None):
    w, ws, (wp, s), wt = layer(cin, cout, k, upconv)
    Ho, Wo = (2 * H, 2 * W) if upconv else (H - k + 1, W - k + 1)
    g_out, saved = packed(B, Ho, Wo, cout), packed(B, H, W, cin)
    g_in = F.PackedView(F.packed_empty(B, H, W, cin, "cuda"))
    F.conv_moments_bwd_data_tc(g_out, B, H, W, k, cout, wt, s, saved, g_in, cin, True, upconv=upconv, kwc=kwc)
    for im2col in ((None,) if upconv or k != 3 else (True, False)):
        rsum = torch.rand((B, H, W) if upconv else (B, Ho, Wo), device="cuda", generator=g)
        work = F.wgrad_workspace(k, cin, cout, "cuda")
        gw, gws = torch.empty_like(w), torch.empty_like(ws)
        F.conv_moments_bwd_weight_tc(g_out, B, H, W, k, cout, saved, cin, rsum, w, ws, work, gw, gws, upconv=upconv,
                                     im2col=im2col)


fwd(2, 10, 12, 32, 32, 3)                      # NT 32, resident
fwd(2, 10, 12, 64, 64, 3)                      # NT 64, resident or streamed
fwd(40, 20, 20, 64, 64, 3)                     # NT 64 streamed, G = 2 pairs
fwd(1, 12, 12, 128, 128, 3)                    # NT 128 streamed
fwd(2, 6, 6, 64, 32, 2, upconv=True)           # up-conv
fwd(2, 9, 36, 32, 32, 3, kwc=True)             # kw-concatenated, TMA-store epilogue
fwd(2, 9, 36, 32, 32, 3, kwc=True, c1=32)      # kw-concatenated, two sources, direct stores
fwd(2, 9, 40, 128, 32, 3, kwc=True)            # kw-concatenated, streamed weights
bwd(2, 10, 12, 32, 32, 3)
bwd(2, 10, 12, 64, 64, 3)
bwd(1, 12, 12, 128, 128, 3)
bwd(2, 6, 6, 64, 32, 2, upconv=True)
bwd(2, 9, 34, 32, 64, 3, kwc=True)             # kw-concatenated data gradient (N = cin = 32)
# first conv (tensor-core), pooling, head
x = torch.rand((2, 12, 14, 4), device="cuda", generator=g)
w0 = torch.randn((3, 3, 4, 32), device="cuda", generator=g) * 0.1
ws0 = torch.full((32,), -4.0, device="cuda")
a0 = F.PackedView(F.packed_empty(2, 10, 12, 32, "cuda"))
F.first_conv_packed(x, w0, ws0, a0, relu=True)
p0 = F.PackedView(F.packed_empty(2, 5, 6, 32, "cuda"))
F.maxpool2_packed(a0, 2, 10, 12, 32, p0)
wf = torch.randn((1, 1, 32, 4), device="cuda", generator=g) * 0.1
wsf = torch.full((4,), -3.0, device="cuda")
pr = torch.empty((2, 120, 4), device="cuda")
vr = torch.empty_like(pr)
F.final_conv_softmax_packed(a0, 2, 10, 12, 32, wf, wsf, pr, vr)
torch.cuda.synchronize()
print("sanitize_kernels: all launches completed")
