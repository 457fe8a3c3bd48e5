"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/supernet.h declares,
argument validation answers without touching a device, and the data-parallel host logic is exercised with
two gloo ranks."""
import ctypes as C
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _S():
    import supernet_b200 as S
    return S


def test_library_exports_every_declared_symbol():
    S = _S()
    lib = S._lib.load()
    names = S._lib.declared_symbols()
    assert len(names) >= 28 and "sn_conv_moments_fwd_tc" in names and "sn_nll_gaussian_bwd" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.sn_version() == 2          # SN_ABI_VERSION: 2 added the FAST-mode backward entry points


def test_struct_layouts_match_header():
    S = _S()
    L = S._lib
    assert C.sizeof(L.sn_conv_desc) == 32 and C.sizeof(L.sn_window) == 72
    assert C.sizeof(L.sn_packed_view) == 40
    assert C.sizeof(L.sn_tc_conv_desc) == 2 * 40 + 8 + 24 + 8 + 8 + 40 + 8 + 8 + 8          # + rsum_out (ABI v2)
    # backward descriptors (ABI v2): g_out + in[2] + g_in[2] views, 4 + 6 ints, 2 pointers / g_out + in[2], 2 + 6 ints, 6 pointers
    assert C.sizeof(L.sn_tc_dgrad_desc) == 5 * 40 + 4 * 4 + 6 * 4 + 2 * 8
    assert C.sizeof(L.sn_tc_wgrad_desc) == 3 * 40 + 2 * 4 + 6 * 4 + 6 * 8
    assert C.sizeof(L.sn_tc_head_desc) == 8 + 6 * 8                                          # n_labels (+ pad), 6 pointers


def test_argument_validation_needs_no_device():
    """Bad arguments are rejected with a negative status and a message before any CUDA call."""
    S = _S()
    lib = S._lib.load()
    d = S._lib.sn_conv_desc(1, 2, 2, 4, 4, 3, 0, 0)          # input smaller than the kernel
    buf = (C.c_float * 64)()
    rc = lib.sn_conv_moments_fwd(C.byref(d), buf, None, buf, buf, buf, buf, None, None)
    assert rc == -1 and b"smaller" in lib.sn_last_error()
    assert lib.sn_conv_moments_fwd(None, buf, None, buf, buf, buf, buf, None, None) == -1
    assert lib.sn_softmax_moments_fwd(C.c_size_t(4), 9, buf, buf, buf, buf, None) == -2      # > 8 classes
    w = S._lib.sn_window(1, 2, 2, 1, 2, 2, 1, 1, 0, 0, 2, 2, 1, 0, 0, 0, 1, 1)               # src window OOB
    assert lib.sn_window_copy(C.byref(w), buf, buf, None) == -1
    t = S._lib.sn_tc_conv_desc()
    t.batch, t.in_h, t.in_w, t.ksize, t.cout = 1, 8, 8, 3, 32
    t.src_c[0] = 16                                                                          # not a multiple of 32
    assert lib.sn_conv_moments_fwd_tc(C.byref(t), None) == -2
    # fused head: shape query and argument checks (no launch)
    t.src_c[0], t.flags = 32, S._lib.SN_TC_RELU
    assert lib.sn_tc_head_fusable(C.byref(t), 4) == 1 and lib.sn_tc_head_fusable(C.byref(t), 6) == 0
    assert lib.sn_tc_head_fusable(None, 4) == 0
    t.cout = 64
    assert lib.sn_tc_head_fusable(C.byref(t), 4) == 0
    hd = S._lib.sn_tc_head_desc()
    hd.n_labels = 4
    assert lib.sn_conv_moments_fwd_tc_head(None, C.byref(hd), None) == -1
    g = S._lib.sn_tc_dgrad_desc()
    g.batch, g.in_h, g.in_w, g.ksize, g.cout = 1, 8, 8, 3, 32
    g.in_c[0] = 48                                                                           # not a multiple of 32
    assert lib.sn_conv_moments_bwd_data_tc(C.byref(g), None) == -2
    wg = S._lib.sn_tc_wgrad_desc()
    wg.batch, wg.in_h, wg.in_w, wg.ksize, wg.cout = 1, 8, 8, 4, 32                            # kernel size 4
    assert lib.sn_conv_moments_bwd_weight_tc(C.byref(wg), None) == -2
    assert lib.sn_conv_moments_bwd_weight_tc(None, None) == -1
    assert lib.sn_wgrad_workspace_bytes(3, 64, 32) == (2 * 9 * 64 * 32 + 32) * 4
    assert lib.sn_packed_bytes(2, 3, 4, 32) == 2 * 3 * 4 * 3 * 32 * 2
    assert lib.sn_prepared_weight_bytes(3, 64, 32) == 3 * 9 * 64 * 32 * 2


def test_ops_refuse_cpu_tensors():
    S = _S()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.ops.relu_moments(torch.zeros(4), torch.zeros(4))
    m = S.Density_prop_with_pad_UNET(32, 3, variant="hippocampus")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 64, 64, 1))


def test_model_structure_matches_reference():
    """Layer names/order, parameter count and sigma-init ranges of Brats.py:331-367 / Hippocampus.py:343-363."""
    S = _S()
    m = S.Density_prop_with_pad_UNET(32, 5)
    m.build_with_input(4, None)
    assert sum(p.numel() for p in m.parameters()) == 7_760_517
    assert m.conv_names[0] == "conv_input" and m.conv_names[-1] == "conv_final" and len(m.conv_names) == 23
    assert m.up1_conv2x2.sigma_min == -4.6 and m.up3_conv2x2.sigma_min == -12 and m.conv_final.sigma_max == -2.2
    assert tuple(m.conv_input.w_mu1.shape) == (3, 3, 4, 32) and tuple(m.up1_conv1.w_mu.shape) == (3, 3, 512, 256)
    assert float(m.conv1.w_mu.abs().max()) <= 0.2 + 1e-6          # truncated normal at 2 sigma
    assert -12 <= float(m.conv1.w_sigma.min()) and float(m.conv1.w_sigma.max()) <= -4.6
    h = S.Density_prop_with_pad_UNET(32, 3, variant="hippocampus")
    h.build_with_input(1, None)
    assert sum(p.numel() for p in h.parameters()) == 466_019 and len(h.conv_names) == 13
    assert h.sigma_fill == 0.02 and m.sigma_fill == 0.1


def test_shard_bounds():
    from supernet_b200 import dp
    assert [dp.shard_bounds(20, 8, r) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 18),
                                                             (18, 20), (20, 20)]
    assert dp.shard_bounds(1, 4, 3) == (1, 1) and dp.shard_bounds(512, 8, 7) == (448, 512)
    with pytest.raises(ValueError):
        dp.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from supernet_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # a tiny stand-in model with the same loss structure: mean-over-samples data term + weight-only regulariser
    w1 = torch.nn.Parameter(torch.randn(6, 5))
    w2 = torch.nn.Parameter(torch.randn(5, 3))
    params = [w1, w2]
    x = torch.randn(7, 6)                       # global batch 7 -> shards of 4 and 3 (unequal on purpose)
    y = torch.randn(7, 3)

    def loss_fn(xs, ys):
        return ((torch.relu(xs @ w1) @ w2 - ys) ** 2).mean() + 1e-2 * (w1.square().sum() + w2.square().sum())

    a, b = dp.shard_bounds(7, world, rank)
    loss_fn(x[a:b], y[a:b]).backward()
    dp.allreduce_gradients(params, b - a, 7, bucket_bytes=64)       # tiny buckets: exercises the multi-bucket path
    got = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    loss_fn(x, y).backward()
    ok = all(torch.allclose(g, p.grad, atol=1e-6) for g, p in zip(got, params))
    # per-variable clipnorm
    for p in params:
        p.grad.mul_(100.0)
    dp.clip_by_norm_per_variable_(params, 1.0)
    ok = ok and all(abs(float(p.grad.norm()) - 1.0) < 1e-5 for p in params)
    # the FAST-mode trainer's path: ONE flat gradient buffer (per-variable gradients are views of it), reduced in place
    for p in params:
        p.grad = None
    flat = torch.zeros(sum(p.numel() for p in params))
    views, off = [], 0
    for p in params:
        views.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    for v, g in zip(views, torch.autograd.grad(loss_fn(x[a:b], y[a:b]), params)):
        v.copy_(g)
    dp.allreduce_flat_(flat, b - a, 7, bucket_bytes=32)             # several slices of the flat buffer
    full = torch.autograd.grad(loss_fn(x, y), params)
    ok = ok and all(torch.allclose(v, g, atol=1e-6) for v, g in zip(views, full))
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_two_gloo_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


class _StubConv:
    def __init__(self, w, ws):
        self.w, self.ws = w, ws

    def weights(self):
        return self.w, self.ws


class _StubFastModel(torch.nn.Module):
    """mode='fast' stand-in with the model surface DataParallelTrainer touches (conv_names, layer.weights())."""
    mode = "fast"
    conv_names = ["a", "b"]

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.wa = torch.nn.Parameter(torch.randn(6, 5, generator=g))
        self.sa = torch.nn.Parameter(torch.randn(5, generator=g))
        self.wb = torch.nn.Parameter(torch.randn(5, 3, generator=g))
        self.sb = torch.nn.Parameter(torch.randn(3, generator=g))
        self.a, self.b = _StubConv(self.wa, self.sa), _StubConv(self.wb, self.sb)

    def loss(self, x, y):
        h = torch.relu(x @ self.wa) * torch.nn.functional.softplus(self.sa)
        return ((h @ self.wb + self.sb - y) ** 2).mean()


def _dp_empty_shard_worker(rank, world, port, out):
    """global batch 1 on 2 ranks -> rank 1's shard is empty.  Both ranks must issue the SAME collective
    sequence in fast mode (fixed slices of one flat buffer); gloo fails or hangs on a mismatch."""
    sys.path.insert(0, ROOT)
    import types
    from supernet_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=__import__("datetime").timedelta(seconds=60))
    model = _StubFastModel()
    g = torch.Generator().manual_seed(9)
    x, y = torch.randn(1, 6, generator=g), torch.randn(1, 3, generator=g)
    trainer = dp.DataParallelTrainer(model, lr=1e-2, kl_factor=0.0)
    calls = []

    def fake_fast_backward(xs, ys):           # what GradientEngine does: gradients are views of ONE flat buffer
        params = [model.wa, model.sa, model.wb, model.sb]
        flat = torch.zeros(sum(p.numel() for p in params))
        off = 0
        loss = model.loss(xs, ys)
        for p, g_ in zip(params, torch.autograd.grad(loss, params)):
            flat[off:off + p.numel()].copy_(g_.reshape(-1))
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        trainer._engine = types.SimpleNamespace(flat_grad=flat)
        calls.append(xs.shape[0])
        return loss.detach()

    trainer._fast_backward = fake_fast_backward
    ref = _StubFastModel()
    ref_opt = dp.make_adam(ref.parameters(), lr=1e-2)
    ok = True
    for _ in range(2):
        a, b = dp.shard_bounds(1, world, rank)
        trainer.step(x[a:b], y[a:b], global_batch=1)
        ref_opt.zero_grad()
        ref.loss(x, y).backward()
        dp.clip_by_norm_per_variable_(list(ref.parameters()), 1.0)
        ref_opt.step()
        ok = ok and all(torch.allclose(p, q, atol=1e-6) for p, q in zip(model.parameters(), ref.parameters()))
    ok = ok and len(calls) == (2 if rank == 0 else 0)
    out[rank] = ok
    dist.destroy_process_group()


def test_fast_mode_empty_shard_issues_the_same_collectives():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_dp_empty_shard_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_noise_drivers_match_reference_semantics():
    """apply_noise / snr / salt_and_pepper against a numpy restatement of Brats.py:1247-1283."""
    import numpy as np
    from supernet_b200 import robustness as R
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 8, 8, 4, generator=g)
    labels = torch.randint(0, 4, (2, 8, 8), generator=g)
    noise = torch.randn(2, 8, 8, 4, generator=g) * 0.3
    xn, ln, nn_ = x.numpy(), labels.numpy(), noise.numpy()
    y_exp = np.broadcast_to(ln[..., None], xn.shape)
    want_o = np.clip(xn + np.ma.filled(np.ma.masked_where(y_exp == 0, nn_), fill_value=0), xn.min(), xn.max())
    want_b = np.clip(xn + np.ma.filled(np.ma.masked_where(y_exp > 0, nn_), fill_value=0), xn.min(), xn.max())
    assert np.allclose(R.apply_noise(x, labels, noise, "O").numpy(), want_o)
    assert np.allclose(R.apply_noise(x, labels, noise, "B").numpy(), want_b)
    assert np.allclose(R.apply_noise(x, labels, noise, "all").numpy(), np.clip(xn + nn_, xn.min(), xn.max()))
    noisy = R.apply_noise(x, labels, noise, "all")
    num = np.sum(np.square(xn)) / 2 * 4
    den = np.sum(np.square(noisy.numpy() - xn)) / 2 * 4
    assert abs(R.snr_db(x, noisy) - 10 * np.log10(num / den)) < 1e-4
    sp = R.salt_and_pepper(x, 0.2, generator=g)
    assert set(sp.unique().tolist()) <= {0.0, 1.0} and 0.03 < float((sp == 1).float().mean()) < 0.2
    sp2 = R.salt_and_pepper(x - 0.5, 0.5, generator=g)
    assert set(sp2.unique().tolist()) <= {-1.0, 0.0, 1.0}
    assert R.make_noise(x, "speckle", 0.1, g).shape == x.shape
    assert R.center_crop(x, 4).shape == (2, 4, 4, 4)
    assert R.one_hot_flat(labels, 4).shape == (2, 64, 4)


def test_build_stamp_survives_relocation(tmp_path):
    """The GPU box runs from a scratch copy of the tree: the shipped library must be accepted there without a rebuild
    (a path-dependent stamp once made eight torchrun ranks rebuild at the same time and corrupt the link)."""
    import importlib.util
    import shutil
    S = _S()
    assert not S.build.needs_build()
    pkg = os.path.dirname(S.build.__file__)
    root = tmp_path / "elsewhere"
    shutil.copytree(pkg, root / os.path.basename(pkg), ignore=shutil.ignore_patterns("build", "__pycache__"))
    shutil.copytree(os.path.join(ROOT, "include"), root / "include")
    spec = importlib.util.spec_from_file_location("sn_build_copy", str(root / os.path.basename(pkg) / "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ROOT == str(root) and not mod.needs_build()
