import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import supernet_b200 as S
from oracle import supernet_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_tc import rand_layer, dev
F = S.fastops
B, H, W = 3, 37, 70
k, cout = 3, 32
mu_e, var_e, _, _ = rand_layer(B, H + 4, W + 4, 32, cout, k, seed=16)
_, _, w, ws = rand_layer(B, H, W, 32, cout, k, seed=17)
m_in, v_in = mu_e[:, 2:2 + H, 2:2 + W], var_e[:, 2:2 + H, 2:2 + W]
m_ref, v_ref = O.relu(*O.conv_intermediate_conv_form(m_in, v_in, w, ws))
ebuf = F.pack_moments(dev(mu_e), dev(var_e))
wp, s = F.prepare_weights(dev(w), dev(ws))
for kwc in (True, False):
    out = F.packed_empty(B, H - 2, W - 2, cout, "cuda")
    out.zero_()
    F.conv_moments_tc(F.PackedView(ebuf, 2, 2, 0), 32, B, H, W, k, cout, wp, s, dst=F.PackedView(out), relu=True, kwc=kwc)
    m, v = F.unpack_moments(out)
    m, v = m.cpu().double(), v.cpu().double()
    print("kwc", kwc, "rel m", O.rel_l2(m, m_ref), "rel v", O.rel_l2(v, v_ref))
    err = (v - v_ref).abs() / (v_ref.abs() + 1e-6 * v_ref.abs().max())
    bad = err > 0.02
    print(" bad fraction", float(bad.double().mean()), "max err", float(err.max()))
    idx = bad.nonzero()
    print(" first bad idx", idx[:10].tolist())
    if len(idx):
        ys = idx[:, 1].unique().tolist(); xs = idx[:, 2].unique().tolist(); cs = idx[:, 3].unique().tolist()
        print(" bad rows", ys[:40], "\n bad cols", xs[:80], "\n bad ch", cs[:40])
        i = tuple(idx[0].tolist())
        print(" value", float(v[i]), "ref", float(v_ref[i]), "mean", float(m[i]), float(m_ref[i]))
