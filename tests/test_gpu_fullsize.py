"""Size-independent properties of the FAST-mode path at BASELINE.json's full benchmark size (BraTS 204x204x4, batch 64
per GPU), where the fp64 oracle would take minutes per slice:

  * slices are independent (that is what makes the path shard without a collective): slice k of a
    batch-64 run equals the batch-1 run on that slice, bit for bit -- tile packing and the persistent tile walk differ,
    the per-pixel arithmetic must not;
  * the mean path has no bias, so it is degree-1 homogeneous: doubling the input doubles every
    pre-softmax mean EXACTLY (a power-of-two scale commutes with every bf16 / fp32 rounding on the way);
  * every variance is finite and non-negative;
  * the input gradient of slice k does not depend on the other slices either (up to the exact 1/B of the NLL mean).
The batch-1 runs themselves are tied to the oracle by tests/test_gpu_tc.py and tests/test_gpu_tc_bwd.py.
"""
import pytest
import torch

from oracle import supernet_oracle as O

pytestmark = pytest.mark.gpu

B_FULL = 64


@pytest.fixture(scope="module")
def S():
    import supernet_b200 as S_
    assert S_._lib.load().sn_device_check() == 0
    return S_


@pytest.fixture(scope="module")
def model(S):
    w32 = O.make_weights("brats", 32, 4, 4)
    return S.Density_prop_with_pad_UNET(32, 4, variant="brats", mode="fast").load_weight_dict(w32, device="cuda")


def test_brats_full_batch_slices_are_independent(S, model):
    from supernet_b200.engine import InferenceEngine
    x = O.make_input("brats", B_FULL, alpha=O.BRATS_ALPHA).cuda()
    big = InferenceEngine(model, B_FULL, 204, 204, 4, "cuda", graph=True, keep_presoftmax=True)
    p, v, pm, pv = big.run(x, return_presoftmax=True)
    assert bool(torch.isfinite(p).all()) and bool(torch.isfinite(v).all())
    assert float(v.min()) >= 0.0 and float(pv.min()) >= 0.0
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5
    one = InferenceEngine(model, 1, 204, 204, 4, "cuda", graph=False, keep_presoftmax=True)
    for k in (0, 17, B_FULL - 1):
        p1, v1, pm1, pv1 = one.run(x[k:k + 1], return_presoftmax=True)
        assert torch.equal(pm1[0], pm[k]) and torch.equal(pv1[0], pv[k])
        assert torch.equal(p1[0], p[k]) and torch.equal(v1[0], v[k])


def test_brats_full_batch_mean_path_is_homogeneous(S, model):
    from supernet_b200.engine import InferenceEngine
    x = O.make_input("brats", B_FULL, alpha=O.BRATS_ALPHA).cuda()
    eng = InferenceEngine(model, B_FULL, 204, 204, 4, "cuda", graph=True, keep_presoftmax=True)
    _, _, pm, _ = eng.run(x, return_presoftmax=True)
    _, _, pm2, _ = eng.run(2.0 * x, return_presoftmax=True)
    assert torch.equal(pm2, 2.0 * pm)


def test_brats_full_batch_input_gradient_is_per_slice(S, model):
    from supernet_b200.engine import GradientEngine
    x = O.make_input("brats", B_FULL, alpha=O.BRATS_ALPHA).cuda()
    y = O.make_labels(B_FULL, 186 * 186, 4).cuda()
    big = GradientEngine(model, B_FULL, 204, 204, 4, "cuda", graph=True)
    loss, g = big.input_gradient(x, y)
    assert bool(torch.isfinite(g).all()) and bool(torch.isfinite(loss).all())
    one = GradientEngine(model, 1, 204, 204, 4, "cuda", graph=False)
    losses = []
    for k in (0, 40):
        l1, g1 = one.input_gradient(x[k:k + 1], y[k:k + 1])
        losses.append(float(l1))
        # the NLL is a mean over B*HW pixels: the batch-64 gradient of slice k is the batch-1 gradient / 64 (exact)
        assert O.rel_l2(g[k].cpu() * B_FULL, g1[0].cpu()) < 1e-6
    assert abs(float(loss) - sum(losses) / len(losses)) < 0.2 * abs(float(loss))     # same order: mean of slice losses
