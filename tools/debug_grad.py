"""Debug helper: lock-step hippocampus forward through S.ops (GPU) and the oracle (CPU fp64), then compare
the gradient of the adversarial loss at every intermediate tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from oracle import supernet_oracle as O

variant, C = "hippocampus", 3
W = O.make_weights(variant, 32, C, 1)
x = O.make_input(variant, 2)
y = O.make_labels(2, 54 * 54, C, dtype=torch.float64)


def run(gpu):
    taps = []
    if gpu:
        w = {k: (a.cuda(), b.cuda()) for k, (a, b) in W.items()}
        conv = lambda n, m, s, relu: S.ops.conv_moments(m, s, w[n][0], w[n][1], relu)
        pool = S.ops.maxpool2_moments
        ups = lambda m, s: (S.ops.unpool(m), S.ops.unpool(s))
        pad = lambda m, s, a, b: (S.ops.pad_hw(m, a, b, 0.0), S.ops.pad_hw(s, a, b, 0.02))
        conc = lambda m, s, me, se: (S.ops.crop_concat(m, me), S.ops.crop_concat(s, se))
        xin = x.cuda().requires_grad_(True)
    else:
        w = {k: (a.double(), b.double()) for k, (a, b) in W.items()}
        def conv(n, m, s, relu):
            m2, s2 = (O.conv_input_conv_form(m, *w[n]) if s is None else O.conv_intermediate_conv_form(m, s, *w[n]))
            return O.relu(m2, s2) if relu else (m2, s2)
        pool = O.maxpooling
        ups = O.upsampling
        pad = lambda m, s, a, b: O.padding(m, s, (a, b), 0.02)
        conc = O.conc
        xin = x.double().requires_grad_(True)

    def tap(name, m, s):
        m.retain_grad(); s.retain_grad(); taps.append((name, m, s))
    m, s = conv("conv_input", xin, None, True); tap("conv_input", m, s)
    m, s = conv("conv1", m, s, True); tap("conv1", m, s)
    skips = [(m, s)]
    ci = 2
    for lvl in (1, 2):
        m, s = pool(m, s); tap(f"pool{lvl}", m, s)
        for _ in range(2):
            m, s = conv(f"conv{ci}", m, s, True); tap(f"conv{ci}", m, s); ci += 1
        if lvl < 2:
            skips.append((m, s))
    for d in (1, 2):
        me, se = skips[2 - d]
        m, s = ups(m, s); tap(f"ups{d}", m, s)
        m, s = conv(f"up{d}_conv2x2", m, s, False); tap(f"up{d}_conv2x2", m, s)
        m, s = pad(m, s, 3, 3); tap(f"pad3_{d}", m, s)
        m, s = conc(m, s, me, se); tap(f"conc{d}", m, s)
        m, s = conv(f"up{d}_conv1", m, s, True); tap(f"up{d}_conv1", m, s)
        m, s = pad(m, s, 2, 2); tap(f"pad2_{d}", m, s)
        m, s = conv(f"up{d}_conv2", m, s, True); tap(f"up{d}_conv2", m, s)
    m, s = conv("conv_final", m, s, False); tap("conv_final", m, s)
    if gpu:
        p, v = S.ops.softmax_moments(m, s)
        p, v = p.reshape(2, -1, C), v.reshape(2, -1, C)
        p.retain_grad(); v.retain_grad(); taps.append(("softmax", p, v))
        loss = 0.5 * S.nll_gaussian(y.float().cuda(), p, v, clip=(-1e4, 1e3))
    else:
        p, v = O.softmax_closed_form(m, s)
        p.retain_grad(); v.retain_grad(); taps.append(("softmax", p, v))
        loss = 0.5 * O.nll_gaussian(y, p, torch.clamp(v, -1e4, 1e3))
    loss.backward()
    return xin, taps, loss


xg, tg, lg = run(True)
xc, tc, lc = run(False)
print("loss", float(lg), float(lc))
print("x grad", O.rel_l2(xg.grad.cpu(), xc.grad))
for (n, m1, s1), (_, m2, s2) in zip(tg, tc):
    print(f"{n:14s} fwd mu {O.rel_l2(m1.cpu(), m2):.2e} var {O.rel_l2(s1.cpu(), s2):.2e} | "
          f"grad mu {O.rel_l2(m1.grad.cpu(), m2.grad):.2e} var {O.rel_l2(s1.grad.cpu(), s2.grad):.2e}")

# ---- pool2 mismatch census
for (n, m1, s1), (_, m2, s2) in zip(tg, tc):
    if n.startswith("pool"):
        d = (s1.detach().cpu().double() - s2.detach()).abs()
        bad = d > 1e-4 * s2.detach().abs().clamp_min(1e-30)
        print(n, "var mismatches", int(bad.sum()), "of", bad.numel())
        idx = bad.nonzero()[:5]
        prev = [t for t in tc if t[0] == ("conv1" if n == "pool1" else "conv3")][0]
        prevg = [t for t in tg if t[0] == ("conv1" if n == "pool1" else "conv3")][0]
        for b, yy, xx, c in idx.tolist():
            win64 = prev[1].detach()[b, 2*yy:2*yy+2, 2*xx:2*xx+2, c].flatten()
            win32 = prevg[1].detach().cpu()[b, 2*yy:2*yy+2, 2*xx:2*xx+2, c].flatten()
            v64 = prev[2].detach()[b, 2*yy:2*yy+2, 2*xx:2*xx+2, c].flatten()
            print("  ", (b, yy, xx, c), "mu64", win64.tolist(), "mu32", win32.tolist(), "var64", v64.tolist(),
                  "got", float(s1[b, yy, xx, c]), "want", float(s2[b, yy, xx, c]))
