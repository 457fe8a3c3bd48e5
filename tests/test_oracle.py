"""CPU tests that pin the oracle: the reference ships no tests or golden
vectors ("parity unpinned"), so the restatement is validated by independent formulations,
a Monte-Carlo check, structural identities, hand-computed cases and committed golden fixtures."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import supernet_oracle as O

D = torch.float64
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand_layer(B=2, H=7, W=6, cin=3, cout=4, k=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(B, H, W, cin, generator=g, dtype=D)
    var = torch.rand(B, H, W, cin, generator=g, dtype=D)
    w = torch.randn(k, k, cin, cout, generator=g, dtype=D) * 0.1
    ws = torch.empty(cout, dtype=D).uniform_(-6, -2, generator=g)
    return mu, var, w, ws


@pytest.mark.parametrize("k", [1, 2, 3])
def test_conv_forms_agree(k):
    mu, var, w, ws = _rand_layer(k=k)
    m1, v1 = O.conv_intermediate_as_written(mu, var, w, ws)
    m2, v2 = O.conv_intermediate_conv_form(mu, var, w, ws)
    assert torch.allclose(m1, m2, atol=1e-13) and torch.allclose(v1, v2, atol=1e-13)
    m1, v1 = O.conv_input_as_written(mu, w, ws)
    m2, v2 = O.conv_input_conv_form(mu, w, ws)
    assert torch.allclose(m1, m2, atol=1e-13) and torch.allclose(v1, v2, atol=1e-13)


def test_conv_against_naive_loops():
    """Pure-python loops on a tiny case: K ordering (kh,kw,ci) and VALID geometry."""
    mu, var, w, ws = _rand_layer(B=1, H=4, W=5, cin=2, cout=3, k=3, seed=3)
    s = torch.log1p(torch.exp(ws))
    m_ref = torch.zeros(1, 2, 3, 3, dtype=D)
    v_ref = torch.zeros(1, 2, 3, 3, dtype=D)
    for i in range(2):
        for j in range(3):
            for n in range(3):
                for kh in range(3):
                    for kw in range(3):
                        for c in range(2):
                            x, v, ww = mu[0, i + kh, j + kw, c], var[0, i + kh, j + kw, c], w[kh, kw, c, n]
                            m_ref[0, i, j, n] += x * ww
                            v_ref[0, i, j, n] += v * ww * ww + s[n] * (x * x + v)
    m, v = O.conv_intermediate_as_written(mu, var, w, ws)
    assert torch.allclose(m, m_ref, atol=1e-13) and torch.allclose(v, v_ref, atol=1e-13)


def test_conv_variance_monte_carlo():
    """Var[sum W x] for independent W~N(w_mu, s), x~N(mu, var) equals the propagated variance."""
    g = torch.Generator().manual_seed(5)
    mu, var, w, ws = _rand_layer(B=1, H=3, W=3, cin=2, cout=2, k=3, seed=11)
    ws = ws + 3.0                      # larger weight variance so all three terms matter
    s = torch.log1p(torch.exp(ws))
    _, v = O.conv_intermediate_conv_form(mu, var, w, ws)
    n = 200000
    xs = mu.flatten() + var.flatten().sqrt() * torch.randn(n, mu.numel(), generator=g, dtype=D)
    wk = w.reshape(-1, 2)                                             # [(kh,kw,ci), n]
    wsamp = wk + s.sqrt() * torch.randn(n, *wk.shape, generator=g, dtype=D)
    y = torch.einsum("bk,bkn->bn", xs, wsamp)
    emp = y.var(0)
    assert torch.allclose(emp, v.flatten(), rtol=3e-2)


def test_unpool_identity_and_upconv_as_transposed_conv():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 4, 5, generator=g, dtype=D)
    u = O.unpool(x)
    assert u.shape == (2, 7, 9, 5)
    ref = torch.zeros_like(u)
    ref[:, 1::2, 1::2, :] = x
    assert torch.equal(u, ref)
    # unpool + 2x2 VALID conv == stride-2 transposed conv with the flipped kernel == 4 parity GEMMs
    w = torch.randn(2, 2, 5, 3, generator=g, dtype=D)
    y = O.conv2d_valid(u, w)
    par = torch.zeros(2, 6, 8, 3, dtype=D)
    for a in range(2):
        for b in range(2):
            par[:, a::2, b::2, :] = torch.einsum("bhwc,cn->bhwn", x, w[1 - a, 1 - b])
    assert torch.allclose(y, par, atol=1e-13)
    wt = w.flip(0, 1).permute(2, 3, 0, 1)                             # [Cin, Cout, kh, kw]
    yt = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt, stride=2).permute(0, 2, 3, 1)
    assert torch.allclose(y, yt, atol=1e-13)


def test_padding_pool_relu_concat_hand_cases():
    mu = torch.tensor([[[[1.0], [-2.0]], [[3.0], [0.5]]]], dtype=D)       # [1,2,2,1]
    var = torch.tensor([[[[0.1], [0.2]], [[0.3], [0.4]]]], dtype=D)
    m, v = O.padding(mu, var, (1, 0), 0.1)
    assert m.shape == (1, 3, 3, 1) and m[0, 0, 0, 0] == 0 and v[0, 0, 0, 0] == 0.1 and v[0, 1, 1, 0] == 0.1 * 1
    assert m[0, 1, 1, 0] == 1.0 and v[0, 2, 2, 0] == 0.4
    mp, vp = O.maxpooling(mu, var)
    assert mp.flatten().tolist() == [3.0] and vp.flatten().tolist() == [0.3]
    mr, vr = O.relu(mu, var)
    assert mr.flatten().tolist() == [1.0, 0.0, 3.0, 0.5] and vr.flatten().tolist() == [0.1, 0.0, 0.3, 0.4]
    enc = torch.arange(16, dtype=D).reshape(1, 4, 4, 1)
    mc, vc = O.conc(mu, var, enc, enc * 2)
    assert mc.shape == (1, 2, 2, 2)
    assert mc[0, :, :, 1].flatten().tolist() == [5.0, 6.0, 9.0, 10.0]
    assert mc[0, :, :, 0].flatten().tolist() == [1.0, -2.0, 3.0, 0.5]
    assert vc[0, :, :, 1].flatten().tolist() == [10.0, 12.0, 18.0, 20.0]


def test_relu_gate_is_strict_and_constant_for_autodiff():
    mu = torch.tensor([[[[0.0, 1.0, -1.0]]]], dtype=D, requires_grad=True)
    var = torch.ones(1, 1, 1, 3, dtype=D, requires_grad=True)
    m, v = O.relu(mu, var)
    assert v.flatten().tolist() == [0.0, 1.0, 0.0]
    (m.sum() + (v * torch.tensor([1.0, 2.0, 3.0], dtype=D)).sum()).backward()
    assert mu.grad.flatten().tolist() == [0.0, 1.0, 0.0]
    assert var.grad.flatten().tolist() == [0.0, 2.0, 0.0]


def test_maxpool_odd_size_same_padding():
    mu = -torch.arange(1, 10, dtype=D).reshape(1, 3, 3, 1)             # all negative, padded cells must not win
    var = torch.arange(1, 10, dtype=D).reshape(1, 3, 3, 1)
    m, v = O.maxpooling(mu, var)
    assert m.shape == (1, 2, 2, 1)
    assert m.flatten().tolist() == [-1.0, -3.0, -7.0, -9.0] and v.flatten().tolist() == [1.0, 3.0, 7.0, 9.0]


@pytest.mark.parametrize("C", [3, 4, 5])
def test_softmax_forms_agree(C):
    g = torch.Generator().manual_seed(C)
    mu = torch.randn(2, 3, 3, C, generator=g, dtype=D) * 2
    var = torch.rand(2, 3, 3, C, generator=g, dtype=D)
    p1, v1 = O.softmax_as_written(mu, var)
    p2, v2 = O.softmax_closed_form(mu, var)
    assert p1.shape == (2, 9, C)
    assert torch.allclose(p1, p2, atol=1e-14) and torch.allclose(v1, v2, atol=1e-14)
    # first-order delta method == J diag(var) J^T diagonal via autograd Jacobian
    J = torch.autograd.functional.jacobian(lambda z: torch.softmax(z, -1), mu[0, 0, 0])
    assert torch.allclose((J.square() @ var[0, 0, 0]), v1[0, 0], atol=1e-13)


def test_nll_and_regularizer_hand_values():
    y = torch.tensor([[[1.0, 0.0]]], dtype=D)
    p = torch.tensor([[[0.75, 0.25]]], dtype=D)
    v = torch.tensor([[[0.5, 0.25]]], dtype=D)
    q = 0.0625 / 0.501 + 0.0625 / 0.251
    l = np.log(0.501 * 0.251)
    assert abs(float(O.nll_gaussian(y, p, v)) - 0.5 * (q + l)) < 1e-14
    ws = torch.tensor([-3.0, 0.5], dtype=D)
    s = np.log1p(np.exp(ws.numpy()))
    assert abs(float(O.sigma_regularizer(ws, 9.0)) - (-9.0 * np.mean(1 + np.log(s) - s))) < 1e-14
    assert float(O.l2_regularizer(torch.tensor([1.0, -2.0], dtype=D))) == 5.0


def test_backward_formulas_match_autograd():
    """SURVEY.md A.3 closed-form conv backward vs autograd of the as-written forward."""
    mu, var, w, ws = _rand_layer(B=2, H=6, W=5, cin=3, cout=4, k=3, seed=7)
    for t in (mu, var, w, ws):
        t.requires_grad_(True)
    m, v = O.conv_intermediate_as_written(mu, var, w, ws)
    g = torch.Generator().manual_seed(9)
    gm = torch.randn(m.shape, generator=g, dtype=D)
    gv = torch.randn(v.shape, generator=g, dtype=D)
    a_mu, a_var, a_w, a_ws = torch.autograd.grad((m * gm).sum() + (v * gv).sum(), (mu, var, w, ws))
    s = O.softplus(ws).detach()
    wd, mud, vard = w.detach(), mu.detach(), var.detach()
    k = 3
    t = (gv * s).sum(-1)                                              # [B,Ho,Wo]
    boxT = F.conv_transpose2d(t.unsqueeze(1), torch.ones(1, 1, k, k, dtype=D)).squeeze(1)
    def convT(gt, wt):                                                # full correlation with flipped W
        return F.conv_transpose2d(gt.permute(0, 3, 1, 2), wt.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    g_mu = convT(gm, wd) + 2 * mud * boxT.unsqueeze(-1)
    g_var = convT(gv, wd.square()) + boxT.unsqueeze(-1)
    pm = O.extract_patches(mud, k).reshape(-1, k * k * 3)
    pv = O.extract_patches(vard, k).reshape(-1, k * k * 3)
    g_w = (pm.T @ gm.reshape(-1, 4) + 2 * wd.reshape(-1, 4) * (pv.T @ gv.reshape(-1, 4))).reshape(wd.shape)
    r = (pm.square() + pv).sum(-1)
    g_ws = (gv.reshape(-1, 4) * r.unsqueeze(-1)).sum(0) * torch.sigmoid(ws.detach())
    for a, b in ((a_mu, g_mu), (a_var, g_var), (a_w, g_w), (a_ws, g_ws)):
        assert torch.allclose(a, b, atol=1e-11)


def test_gradcheck_small_layers():
    mu, var, w, ws = _rand_layer(B=1, H=4, W=4, cin=2, cout=2, k=2, seed=13)
    for t in (mu, var, w, ws):
        t.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda *a: O.conv_intermediate_conv_form(*a), (mu, var, w, ws), atol=1e-7)
    z = torch.randn(1, 2, 2, 3, dtype=D, requires_grad=True)
    zv = torch.rand(1, 2, 2, 3, dtype=D, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: O.softmax_closed_form(a, b), (z, zv), atol=1e-7)


@pytest.mark.parametrize("variant,C,cin", [("hippocampus", 3, 1), ("brats", 4, 4)])
def test_model_geometry_and_param_count(variant, C, cin):
    m = O.UNetOracle(variant, 32, C, cin, torch.float32)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == (466019 if variant == "hippocampus" else 7760484)   # SURVEY.md 6 / BASELINE.md 2
    if variant == "hippocampus":
        p, v = m(O.make_input(variant, 1))
        assert p.shape == (1, 54 * 54, 3) and v.shape == p.shape
        assert bool((v >= 0).all()) and torch.allclose(p.sum(-1), torch.ones(1, 2916), atol=1e-5)


def test_model_forms_agree_hippocampus():
    x = O.make_input("hippocampus", 1)
    a = O.UNetOracle("hippocampus", 32, 3, 1, D, form="as_written")(x, True)
    b = O.UNetOracle("hippocampus", 32, 3, 1, D, form="conv")(x, True)
    for ta, tb in zip(a, b):
        assert O.rel_l2(ta, tb) < 1e-12


def test_truncated_normal_and_weight_determinism():
    w1 = O.make_weights("hippocampus", 32, 3, 1)
    w2 = O.make_weights("hippocampus", 32, 3, 1)
    for k in w1:
        assert torch.equal(w1[k][0], w2[k][0]) and torch.equal(w1[k][1], w2[k][1])
        assert float(w1[k][0].abs().max()) <= 0.2
    lo, hi = w1["conv_final"][1].min(), w1["conv_final"][1].max()
    assert -4.6 <= lo and hi <= -2.2
    assert w1["conv1"][1].min() >= -12 and w1["conv1"][1].max() <= -4.6


def test_golden_fixtures_match_oracle():
    """The committed fixtures (tests/golden/make_golden.py) must still be reproduced."""
    path = os.path.join(GOLD, "hippocampus_b2_fp64.npz")
    z = np.load(path)
    x = O.make_input("hippocampus", 2)
    p, v, mf, sf = O.UNetOracle("hippocampus", 32, 3, 1, D)(x, True)
    idx = z["idx"]
    for name, t in (("p", p), ("v", v), ("mf", mf), ("sf", sf)):
        got = t.detach().flatten()[idx].numpy()
        assert np.allclose(got, z[name], rtol=1e-9, atol=0), name
    assert abs(float(z["p_sum"]) - float(p.sum())) < 1e-6
    assert np.isclose(float(z["v_sum"]), float(v.sum()), rtol=1e-9)
    lay = np.load(os.path.join(GOLD, "layers_fp64.npz"))
    mu, var, w, ws = (torch.from_numpy(lay[k]) for k in ("mu", "var", "w", "ws"))
    m, vv = O.conv_intermediate_conv_form(mu, var, w, ws)
    assert np.allclose(m.numpy(), lay["m_out"], rtol=1e-12) and np.allclose(vv.numpy(), lay["v_out"], rtol=1e-12)


def test_forward_on_own_trajectory_is_the_identity():
    """UNetOracle.forward(trajectory=...) with the oracle's own layer outputs reproduces outputs and gradients; with a
    trajectory whose gates differ, the gates of the TRAJECTORY decide (the straight-through construction the FAST-mode
    gradient tests rely on)."""
    orc = O.UNetOracle("hippocampus", 32, 3, 1, torch.float64)
    x = O.make_input("hippocampus", 1)
    y = O.make_labels(1, 54 * 54, 3, dtype=torch.float64)
    taps = {}
    p0, v0 = orc.forward(x, taps=taps)
    p1, v1 = orc.forward(x, trajectory=taps)
    assert torch.allclose(p0, p1, atol=1e-14) and torch.allclose(v0, v1, atol=1e-14)
    g0, l0 = orc.fgsm_gradient(x, y)
    g1, l1 = orc.fgsm_gradient(x, y, trajectory=taps)
    assert O.rel_l2(g1, g0) < 1e-12 and abs(float(l0) - float(l1)) < 1e-14
    # close one open gate of conv1 in the trajectory: that unit's output and gradient path must vanish
    m, s = taps["conv1"]
    idx = (m > 0).nonzero()[0]
    m2, s2 = m.clone(), s.clone()
    m2[tuple(idx)] = 0.0
    s2[tuple(idx)] = 0.0
    t2 = dict(taps)
    t2["conv1"] = (m2, s2)
    taps2 = {}
    orc.forward(x, taps=taps2, trajectory=t2)
    assert float(taps2["conv1"][0][tuple(idx)]) == 0.0
    g2, _ = orc.fgsm_gradient(x, y, trajectory=t2)
    assert O.rel_l2(g2, g0) > 0


# ------------------------------------------------------------------------------------------------------------
# the torch oracle against the independent NumPy restatement (oracle/numpy_check.py) and its committed output
# ------------------------------------------------------------------------------------------------------------
def _golden_np(name):
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


@pytest.mark.parametrize("form", ["conv", "as_written"])
@pytest.mark.parametrize("variant,C,in_ch,B,alpha,fname", [
    ("hippocampus", 3, 1, 2, 1.0, "hippocampus_b2_numpy_fp64.npz"),
    ("brats", 4, 4, 1, O.BRATS_ALPHA, "brats_b1_numpy_fp64.npz")])
def test_torch_oracle_reproduces_the_numpy_restatement(variant, C, in_ch, B, alpha, fname, form):
    """The fixtures were produced by NumPy index arithmetic that shares nothing with the torch oracle (different conv:
    patches x matrix; pooling with TF's batch-inclusive flat arg-max index + flat gather; unpool op by op).  Both forms
    of the torch oracle must land on them to 1e-12."""
    if variant == "brats" and form == "as_written":
        pytest.skip("BraTS as-written in fp64 needs ~2 GB of patch matrices; the conv form covers the graph")
    z = _golden_np(fname)
    p, v, mf, sf = O.UNetOracle(variant, 32, C, in_ch, torch.float64, form=form)(O.make_input(variant, B, alpha=alpha), True)
    idx = torch.from_numpy(z["idx"])
    for got, key in ((p, "p"), (v, "v"), (mf, "mf"), (sf, "sf")):
        assert O.rel_l2(got.flatten()[idx], torch.from_numpy(z[key])) < 1e-12, key
    assert abs(float(p.sum()) - float(z["p_sum"])) < 1e-9 * float(z["p_sum"])
    assert abs(float(v.sum()) - float(z["v_sum"])) < 1e-9 * float(z["v_sum"])
    assert abs(float(mf.abs().sum()) - float(z["mf_abs_sum"])) < 1e-9 * float(z["mf_abs_sum"])
    assert abs(float(sf.sum()) - float(z["sf_sum"])) < 1e-9 * float(z["sf_sum"])


def test_numpy_restatement_layers_against_torch_oracle_and_hand_cases():
    """Layer by layer, including what the whole-network fixtures exercise only lightly: odd sizes and ties in the
    pooling, the flat arg-max index formula, a hand-computed first conv."""
    from oracle import numpy_check as N
    g = np.random.default_rng(3)
    mu, var = g.standard_normal((2, 7, 9, 5)), g.random((2, 7, 9, 5))
    w, ws = g.standard_normal((3, 3, 5, 6)) * 0.1, g.uniform(-8, -2, 6)
    a, b = N.conv_intermediate(mu, var, w, ws)
    c, d = O.conv_intermediate_conv_form(*(torch.from_numpy(t) for t in (mu, var, w, ws)))
    assert np.allclose(a, c.numpy(), rtol=0, atol=1e-13) and np.allclose(b, d.numpy(), rtol=0, atol=1e-13)
    m = np.maximum(mu, 0)                                    # post-ReLU tensors have tied zeros
    pm, pv = N.maxpooling(m, var)
    qm, qv = O.maxpooling(torch.from_numpy(m), torch.from_numpy(var))
    assert np.array_equal(pm, qm.numpy()) and np.array_equal(pv, qv.numpy())
    _, arg = N.maxpool_with_argmax(m)
    assert np.array_equal(m.reshape(-1)[arg], pm)           # include_batch_in_index=True: index into the whole batch
    assert arg[1].min() >= 7 * 9 * 5                        # second image's indices lie beyond the first image
    u = N.unpool(mu)
    assert u.shape == (2, 15, 19, 5) and np.array_equal(u[:, 1::2, 1::2], mu) and float(np.abs(u).sum()) == float(np.abs(mu).sum())
    # hand case: 1 pixel of output, x = ones(3x3x1): mean = sum(w), var = softplus(w_sigma) * 9
    w1 = np.arange(9, dtype=np.float64).reshape(3, 3, 1, 1)
    mo, vo = N.conv_input(np.ones((1, 3, 3, 1)), w1, np.array([0.0]))
    assert float(mo) == 36.0 and abs(float(vo) - 9 * np.log(2.0)) < 1e-15
    p, s = N.softmax_moments(np.zeros((1, 1, 1, 2)), np.array([1.0, 3.0]).reshape(1, 1, 1, 2))
    assert np.allclose(p, 0.5) and np.allclose(s, 0.0625 * 4.0)        # J = [[.25,-.25],[-.25,.25]]: J^2 @ [1,3] = .25
