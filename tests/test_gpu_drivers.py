"""GPU tests of the callers either side of the hot path: FGSM / PGD / saliency through the moment
layers' data-gradient chain, the noise sweep of testing(), and one ELBO training step (Adam + per-variable
clipnorm) -- each against the CPU oracle's autograd on identical weights and inputs."""
import pytest
import torch

from oracle import supernet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supernet_b200 as S_
    assert S_._lib.load().sn_device_check() == 0
    return S_


def _pair(S, mode="fp32", C=3):
    oracle = O.UNetOracle("hippocampus", 32, C, 1, torch.float64)
    w32 = O.make_weights("hippocampus", 32, C, 1)
    model = S.Density_prop_with_pad_UNET(32, C, variant="hippocampus", mode=mode).load_weight_dict(w32, device="cuda")
    return oracle, model


def test_fgsm_and_pgd(S):
    from supernet_b200 import robustness as R
    oracle, model = _pair(S)
    x = O.make_input("hippocampus", 2)
    g = torch.Generator().manual_seed(7)
    labels = torch.randint(0, 3, (2, 54, 54), generator=g)
    y = R.one_hot_flat(labels, 3)
    g_ref, _ = oracle.fgsm_gradient(x, y.double())
    eps = 1e-2
    adv, sign = R.fgsm_untargeted(model, x.cuda(), labels.cuda(), eps)
    big = g_ref.abs() > 1e-3 * g_ref.abs().max()
    assert float((torch.sign(g_ref)[big] == sign.cpu().double()[big]).double().mean()) >= 0.999
    assert float((adv.cpu() - x).abs().max()) <= eps + 1e-7
    assert float(adv.min()) >= float(x.min()) and float(adv.max()) <= float(x.max())
    # the attack must not lower the adversarial loss (0.5 NLL, Brats.py:587-590)
    with torch.no_grad():
        l0 = float(model.adversarial_loss(x.cuda(), y.cuda()))
        l1 = float(model.adversarial_loss(adv, y.cuda()))
    assert l1 >= l0 - 1e-6
    advp = R.pgd_targeted(model, x.cuda(), labels.cuda(), source_class=2, target_class=1, epsilon=eps, steps=3,
                          step_size=eps / 2)
    assert float((advp.cpu() - x).abs().max()) <= eps + 1e-7


def test_saliency_map(S):
    from supernet_b200 import robustness as R
    oracle, model = _pair(S)
    x = O.make_input("hippocampus", 1)
    grad, relu_grad, pred = R.create_saliency_map(model, x.cuda(), target_class=1, class_only=True)
    xr = x.double().requires_grad_(True)
    p, _ = oracle(xr)
    mask = (p.argmax(-1) == 1).double()
    (g_ref,) = torch.autograd.grad((p[..., 1] * mask).sum(), xr)
    assert O.rel_l2(grad.cpu(), g_ref) < 1e-2           # arg-max near-tie caveat of test_gpu_fp32.py applies
    assert torch.equal(relu_grad, torch.relu(grad))
    assert O.rel_l2(pred.cpu(), p.detach()) < 1e-5
    # the reference's own formula sums every class of the selected pixels == their count: zero up to rounding
    g0, _, _ = R.create_saliency_map(model, x.cuda(), target_class=1)
    assert float(g0.abs().max()) < 1e-4 * float(grad.abs().max())


def test_noise_sweep_fast_mode(S):
    """testing()'s call pattern (Hippocampus.py:1580: std in {0.05, 0.1}) through mode='fast'."""
    from supernet_b200 import robustness as R
    oracle, model = _pair(S, mode="fast")
    x = O.make_input("hippocampus", 4)
    g = torch.Generator().manual_seed(11)
    labels = torch.randint(0, 3, (4, 64, 64), generator=g)
    var_means = []
    for std, where in ((0.0, "all"), (0.05, "O"), (0.1, "all")):
        xn = R.apply_noise(x, labels, R.make_noise(x, "gaussian", std, g), where) if std > 0 else x
        p_ref, v_ref = oracle(xn)
        with torch.no_grad():
            p, v = model(xn.cuda())
        assert O.rel_l2(p.cpu(), p_ref) < 1e-3 and O.rel_l2(v.cpu(), v_ref) < 1e-2
        assert O.argmax_agreement(p.cpu(), p_ref) >= 0.999
        var_means.append(R.mean_predicted_class_variance(p, v))
        assert abs(var_means[-1] - R.mean_predicted_class_variance(p_ref, v_ref)) < 1e-2 * var_means[-1]
        if std > 0:
            assert R.snr_db(x, xn) > 0


def test_training_step_matches_oracle_adam(S):
    """Three train_on_batch steps (Brats.py:569-580; lr 1e-4, kl_factor 1e-3 as in Hippocampus.py:425) track the
    fp64 oracle running the same Keras-style Adam (per-variable clipnorm 1.0, eps 1e-7)."""
    from supernet_b200 import dp
    oracle, model = _pair(S)
    oracle.requires_grad_(True)
    x = O.make_input("hippocampus", 2)
    y = O.make_labels(2, 54 * 54, 3)
    kl, lr = 1e-3, 1e-4
    params = oracle.parameters()
    opt_ref = dp.make_adam(params, lr=lr)
    trainer = dp.DataParallelTrainer(model, lr=lr, kl_factor=kl)
    before = [p.detach().clone() for c in model.convs() for p in c.weights()]
    ref0 = [p.detach().clone() for p in params]
    for step in range(3):
        loss_ref = oracle.elbo_loss(x, y.double(), kl)
        grads = torch.autograd.grad(loss_ref, params)
        for p, g_ in zip(params, grads):
            p.grad = g_.clone()
        dp.clip_by_norm_per_variable_(params, 1.0)
        opt_ref.step()
        loss = trainer.step(x.cuda(), y.cuda(), global_batch=2)
        assert abs(float(loss) - float(loss_ref)) < 2e-3 * abs(float(loss_ref)), (step, float(loss), float(loss_ref))
        if step == 0:
            # Adam's first step moves every weight by ~lr * sign(grad): compare the UPDATE, not the weight
            after = [p for c in model.convs() for p in c.weights()]
            worst = max(O.rel_l2(a.detach().cpu().double() - b.cpu().double(), r.detach() - r0)
                        for b, a, r, r0 in zip(before, after, params, ref0))
            assert worst < 5e-2, worst


def test_metrics_and_batch_feeding_stay_on_the_device(S):
    """Dice / sensitivity / precision / specificity and the shard -> batch conversion run on CUDA
    tensors (no .numpy() round trip per step, Brats.py:688-705) and agree with the CPU evaluation."""
    import numpy as np
    from supernet_b200 import dataio, metrics
    g = np.random.default_rng(11)
    x = g.random((4, 1, 64, 64)).astype(np.float32)                 # shard layout [B,C,H,W]
    y = g.integers(0, 3, size=(4, 64, 64))
    xt, yc, onehot = dataio.batch_from_shard(x, y, out_size=54, n_labels=3, device="cuda")
    assert xt.is_cuda and xt.shape == (4, 64, 64, 1) and onehot.shape == (4, 54 * 54, 3)
    _, model = _pair(S, mode="fast")
    with torch.no_grad():
        p, _ = model(xt)
    pred = metrics.predictions_to_labels(p, 54, 54)
    assert pred.is_cuda
    rep_gpu = metrics.region_report(yc, pred)
    rep_cpu = metrics.region_report(yc.cpu(), pred.cpu())
    for region in rep_gpu:
        for k, v in rep_gpu[region].items():
            a, b = v, rep_cpu[region][k]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-12
