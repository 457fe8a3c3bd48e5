"""PyTorch custom ops over the C-ABI (FP32 mode): every forward and backward is one or a few calls into
libsupernet_b200.so on the current CUDA stream.  Tensors are NHWC contiguous float32 on a CUDA device;
"sigma"/"var" is the VARIANCE, as in the reference.

There is no CPU path here: a CPU tensor raises.  Reference citations are file:line into /root/reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import SN_CONV_RELU, check, ptr, sn_conv_desc, sn_window, stream_ptr

Tensor = torch.Tensor


def _req(t: Optional[Tensor], name: str) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the moment path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    _lib.check_device(t)
    return t.contiguous()


def _conv_desc(B, H, W, cin, cout, k, flags=0) -> sn_conv_desc:
    return sn_conv_desc(B, H, W, cin, cout, k, flags, 0)


# ---------------------------------------------------------------------------------------------------
# moment convolution (Brats.py:65-76, 118-137)
# ---------------------------------------------------------------------------------------------------
class _ConvMoments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, var, w_mu, w_sigma, relu: bool):
        lib = _lib.load()
        mu = _req(mu, "mu")
        var = _req(var, "var")
        w_mu = _req(w_mu, "w_mu")
        w_sigma = _req(w_sigma, "w_sigma")
        B, H, W, cin = mu.shape
        k, k2, wcin, cout = w_mu.shape
        if k != k2 or wcin != cin or w_sigma.shape != (cout,):
            raise RuntimeError(f"conv_moments: weight {tuple(w_mu.shape)} / sigma {tuple(w_sigma.shape)} "
                               f"do not match input channels {cin}")
        if var is not None and var.shape != mu.shape:
            raise RuntimeError("conv_moments: mean and variance shapes differ")
        Ho, Wo = H - k + 1, W - k + 1
        mu_out = torch.empty((B, Ho, Wo, cout), device=mu.device, dtype=torch.float32)
        var_out = torch.empty_like(mu_out)
        rsum = torch.empty((B, Ho, Wo), device=mu.device, dtype=torch.float32)
        d = _conv_desc(B, H, W, cin, cout, k, SN_CONV_RELU if relu else 0)
        check(lib.sn_conv_moments_fwd(C.byref(d), ptr(mu), ptr(var), ptr(w_mu), ptr(w_sigma), ptr(mu_out),
                                      ptr(var_out), ptr(rsum), stream_ptr()), "conv_moments_fwd")
        ctx.relu = relu
        ctx.has_var = var is not None
        ctx.save_for_backward(mu, var, w_mu, w_sigma, rsum, mu_out if relu else None)
        return mu_out, var_out

    @staticmethod
    def backward(ctx, g_mu_out, g_var_out):
        lib = _lib.load()
        mu, var, w_mu, w_sigma, rsum, mu_out = ctx.saved_tensors
        B, H, W, cin = mu.shape
        k, _, _, cout = w_mu.shape
        st = stream_ptr()
        g_mu_out = g_mu_out.contiguous()
        g_var_out = g_var_out.contiguous()
        if ctx.relu:
            gm = torch.empty_like(g_mu_out)
            gv = torch.empty_like(g_var_out)
            check(lib.sn_relu_moments_bwd(C.c_size_t(g_mu_out.numel()), ptr(mu_out), ptr(g_mu_out), ptr(g_var_out),
                                          ptr(gm), ptr(gv), st), "relu_moments_bwd")
            g_mu_out, g_var_out = gm, gv
        d = _conv_desc(B, H, W, cin, cout, k, 0)
        g_mu = g_var = g_w = g_ws = None
        if ctx.needs_input_grad[0] or (ctx.has_var and ctx.needs_input_grad[1]):
            g_mu = torch.empty_like(mu)
            g_var = torch.empty_like(mu) if ctx.has_var else None
            check(lib.sn_conv_moments_bwd_data(C.byref(d), ptr(g_mu_out), ptr(g_var_out), ptr(mu), ptr(w_mu),
                                               ptr(w_sigma), ptr(g_mu), ptr(g_var), st), "conv_moments_bwd_data")
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            g_w = torch.empty_like(w_mu)
            g_ws = torch.empty_like(w_sigma)
            check(lib.sn_conv_moments_bwd_weight(C.byref(d), ptr(mu), ptr(var), ptr(g_mu_out), ptr(g_var_out),
                                                 ptr(rsum), ptr(w_mu), ptr(w_sigma), ptr(g_w), ptr(g_ws), st),
                  "conv_moments_bwd_weight")
        return g_mu, g_var, g_w, g_ws, None


def conv_bwd_weight_raw(B, H, W, cin, cout, k, mu, var, g_mu_out, g_var_out, rsum, w_mu, w_sigma, g_w, g_ws) -> None:
    """sn_conv_moments_bwd_weight on caller-owned tensors (no autograd, no allocation): used by the FAST-mode
    training engine for the two layers too thin for the tensor cores."""
    d = _conv_desc(B, H, W, cin, cout, k, 0)
    check(_lib.load().sn_conv_moments_bwd_weight(C.byref(d), ptr(mu), ptr(var), ptr(g_mu_out), ptr(g_var_out),
                                                 ptr(rsum), ptr(w_mu), ptr(w_sigma), ptr(g_w), ptr(g_ws),
                                                 stream_ptr()), "conv_moments_bwd_weight")


def conv_moments(mu: Tensor, var: Optional[Tensor], w_mu: Tensor, w_sigma: Tensor, relu: bool = False):
    """myConv_input.call (var=None, Brats.py:65-76) / myConv_intermediate.call (Brats.py:118-137), optionally
    fused with myReLU (Brats.py:233-238)."""
    return _ConvMoments.apply(mu, var, w_mu, w_sigma, relu)


# ---------------------------------------------------------------------------------------------------
# ReLU moment gate (Brats.py:220-238)
# ---------------------------------------------------------------------------------------------------
class _ReluMoments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, var):
        lib = _lib.load()
        mu = _req(mu, "mu")
        var = _req(var, "var")
        mu_out = torch.empty_like(mu)
        var_out = torch.empty_like(var)
        check(lib.sn_relu_moments_fwd(C.c_size_t(mu.numel()), ptr(mu), ptr(var), ptr(mu_out), ptr(var_out),
                                      stream_ptr()), "relu_moments_fwd")
        ctx.save_for_backward(mu_out)
        return mu_out, var_out

    @staticmethod
    def backward(ctx, g_mu_out, g_var_out):
        lib = _lib.load()
        (mu_out,) = ctx.saved_tensors
        g_mu_out = g_mu_out.contiguous()
        g_var_out = g_var_out.contiguous()
        g_mu = torch.empty_like(g_mu_out)
        g_var = torch.empty_like(g_var_out)
        check(lib.sn_relu_moments_bwd(C.c_size_t(mu_out.numel()), ptr(mu_out), ptr(g_mu_out), ptr(g_var_out),
                                      ptr(g_mu), ptr(g_var), stream_ptr()), "relu_moments_bwd")
        return g_mu, g_var


def relu_moments(mu: Tensor, var: Tensor):
    return _ReluMoments.apply(mu, var)


# ---------------------------------------------------------------------------------------------------
# arg-max pooling (Brats.py:171-174, 206-216)
# ---------------------------------------------------------------------------------------------------
class _MaxPoolMoments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, var):
        lib = _lib.load()
        mu = _req(mu, "mu")
        var = _req(var, "var")
        B, H, W, Cc = mu.shape
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        mu_out = torch.empty((B, Ho, Wo, Cc), device=mu.device, dtype=torch.float32)
        var_out = torch.empty_like(mu_out)
        amax = torch.empty((B, Ho, Wo, Cc), device=mu.device, dtype=torch.uint8)
        check(lib.sn_maxpool2_moments_fwd(B, H, W, Cc, ptr(mu), ptr(var), ptr(mu_out), ptr(var_out), ptr(amax),
                                          stream_ptr()), "maxpool2_moments_fwd")
        ctx.shape = (B, H, W, Cc)
        ctx.save_for_backward(amax)
        return mu_out, var_out

    @staticmethod
    def backward(ctx, g_mu_out, g_var_out):
        lib = _lib.load()
        (amax,) = ctx.saved_tensors
        B, H, W, Cc = ctx.shape
        g_mu_out = g_mu_out.contiguous()
        g_var_out = g_var_out.contiguous()
        g_mu = torch.empty((B, H, W, Cc), device=g_mu_out.device, dtype=torch.float32)
        g_var = torch.empty_like(g_mu)
        check(lib.sn_maxpool2_moments_bwd(B, H, W, Cc, ptr(amax), ptr(g_mu_out), ptr(g_var_out), ptr(g_mu),
                                          ptr(g_var), stream_ptr()), "maxpool2_moments_bwd")
        return g_mu, g_var


def maxpool2_moments(mu: Tensor, var: Tensor):
    return _MaxPoolMoments.apply(mu, var)


# ---------------------------------------------------------------------------------------------------
# window copies: unpool / pad / crop+concat and their adjoints
# ---------------------------------------------------------------------------------------------------
def _window_copy(src: Tensor, dst: Tensor, h: int, w: int, c: int, src_org=(0, 0, 0), dst_org=(0, 0, 0),
                 dst_step: int = 1, src_step: int = 1) -> None:
    lib = _lib.load()
    B, sh, sw, sc = src.shape
    _, dh, dw, dc = dst.shape
    win = sn_window(B, h, w, c, sh, sw, sc, src_org[0], src_org[1], src_org[2], dh, dw, dc, dst_org[0], dst_org[1],
                    dst_org[2], dst_step, src_step)
    check(lib.sn_window_copy(C.byref(win), ptr(src), ptr(dst), stream_ptr()), "window_copy")


def _fill(t: Tensor, value: float) -> None:
    check(_lib.load().sn_fill(ptr(t), C.c_size_t(t.numel()), C.c_float(value), stream_ptr()), "fill")


class _Unpool(torch.autograd.Function):
    """unpool (Brats.py:178-203): out[b,2y+1,2x+1,c] = in[b,y,x,c], zeros elsewhere, size 2H+1."""

    @staticmethod
    def forward(ctx, x):
        x = _req(x, "x")
        B, H, W, Cc = x.shape
        out = torch.empty((B, 2 * H + 1, 2 * W + 1, Cc), device=x.device, dtype=torch.float32)
        _fill(out, 0.0)
        _window_copy(x, out, H, W, Cc, dst_org=(1, 1, 0), dst_step=2)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, H2, W2, Cc = g.shape
        H, W = (H2 - 1) // 2, (W2 - 1) // 2
        gi = torch.empty((B, H, W, Cc), device=g.device, dtype=torch.float32)
        _window_copy(g, gi, H, W, Cc, src_org=(1, 1, 0), src_step=2)
        return gi


def unpool(x: Tensor) -> Tensor:
    return _Unpool.apply(x)


class _Pad(torch.autograd.Function):
    """tf.pad CONSTANT on H and W by [a, b] (Brats.py:160-162)."""

    @staticmethod
    def forward(ctx, x, a: int, b: int, value: float):
        x = _req(x, "x")
        B, H, W, Cc = x.shape
        out = torch.empty((B, H + a + b, W + a + b, Cc), device=x.device, dtype=torch.float32)
        _fill(out, float(value))
        _window_copy(x, out, H, W, Cc, dst_org=(a, a, 0))
        ctx.geom = (a, H, W, Cc)
        return out

    @staticmethod
    def backward(ctx, g):
        a, H, W, Cc = ctx.geom
        g = g.contiguous()
        gi = torch.empty((g.shape[0], H, W, Cc), device=g.device, dtype=torch.float32)
        _window_copy(g, gi, H, W, Cc, src_org=(a, a, 0))
        return gi, None, None, None


def pad_hw(x: Tensor, a: int, b: int, value: float) -> Tensor:
    return _Pad.apply(x, int(a), int(b), float(value))


class _CropConcat(torch.autograd.Function):
    """concat([dec, centre_crop(enc)], C) (Brats.py:247-261, Brats_functions.py:518-526)."""

    @staticmethod
    def forward(ctx, dec, enc):
        dec = _req(dec, "dec")
        enc = _req(enc, "enc")
        B, H, W, Cd = dec.shape
        _, He, We, Ce = enc.shape
        if He < H or We < W or enc.shape[0] != B:
            raise RuntimeError("crop_concat: encoder tensor smaller than decoder tensor")
        oy, ox = (He - H) // 2, (We - W) // 2
        out = torch.empty((B, H, W, Cd + Ce), device=dec.device, dtype=torch.float32)
        _window_copy(dec, out, H, W, Cd)
        _window_copy(enc, out, H, W, Ce, src_org=(oy, ox, 0), dst_org=(0, 0, Cd))
        ctx.geom = (H, W, Cd, He, We, Ce, oy, ox)
        return out

    @staticmethod
    def backward(ctx, g):
        H, W, Cd, He, We, Ce, oy, ox = ctx.geom
        g = g.contiguous()
        B = g.shape[0]
        gd = torch.empty((B, H, W, Cd), device=g.device, dtype=torch.float32)
        ge = torch.empty((B, He, We, Ce), device=g.device, dtype=torch.float32)
        _window_copy(g, gd, H, W, Cd)
        if He != H or We != W:
            _fill(ge, 0.0)
        _window_copy(g, ge, H, W, Ce, src_org=(0, 0, Cd), dst_org=(oy, ox, 0))
        return gd, ge


def crop_concat(dec: Tensor, enc: Tensor) -> Tensor:
    return _CropConcat.apply(dec, enc)


# ---------------------------------------------------------------------------------------------------
# softmax with Jacobian-propagated variance (Brats.py:269-283)
# ---------------------------------------------------------------------------------------------------
class _SoftmaxMoments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, var):
        lib = _lib.load()
        mu = _req(mu, "mu")
        var = _req(var, "var")
        Cc = mu.shape[-1]
        rows = mu.numel() // Cc
        p = torch.empty_like(mu)
        vo = torch.empty_like(var)
        check(lib.sn_softmax_moments_fwd(C.c_size_t(rows), Cc, ptr(mu), ptr(var), ptr(p), ptr(vo), stream_ptr()),
              "softmax_moments_fwd")
        ctx.save_for_backward(p, var)
        return p, vo

    @staticmethod
    def backward(ctx, g_p, g_vo):
        lib = _lib.load()
        p, var = ctx.saved_tensors
        Cc = p.shape[-1]
        rows = p.numel() // Cc
        g_p = g_p.contiguous()
        g_vo = g_vo.contiguous()
        g_mu = torch.empty_like(p)
        g_var = torch.empty_like(p)
        check(lib.sn_softmax_moments_bwd(C.c_size_t(rows), Cc, ptr(p), ptr(var), ptr(g_p), ptr(g_vo), ptr(g_mu),
                                         ptr(g_var), stream_ptr()), "softmax_moments_bwd")
        return g_mu, g_var


def softmax_moments(mu: Tensor, var: Tensor):
    """Softmax over the last axis with var_out_i = sum_j (p_i (delta_ij - p_j))^2 var_j; shapes are kept."""
    return _SoftmaxMoments.apply(mu, var)


# ---------------------------------------------------------------------------------------------------
# Gaussian NLL (Brats.py:293-311) on clip(var, lo, hi) (Brats.py:573-574, 588-589)
# ---------------------------------------------------------------------------------------------------
class _NllGaussian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, p, var, lo: float, hi: float):
        lib = _lib.load()
        y = _req(y, "y")
        p = _req(p, "p")
        var = _req(var, "var")
        if y.shape != p.shape or var.shape != p.shape:
            raise RuntimeError("nll_gaussian: y, mean and variance shapes differ")
        Cc = p.shape[-1]
        rows = p.numel() // Cc
        acc = torch.empty(2, device=p.device, dtype=torch.float64)
        loss = torch.empty((), device=p.device, dtype=torch.float32)
        check(lib.sn_nll_gaussian_fwd(C.c_size_t(rows), Cc, ptr(y), ptr(p), ptr(var), C.c_float(lo), C.c_float(hi),
                                      ptr(acc), ptr(loss), stream_ptr()), "nll_gaussian_fwd")
        ctx.clip = (lo, hi)
        ctx.save_for_backward(y, p, var, acc)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        y, p, var, acc = ctx.saved_tensors
        lo, hi = ctx.clip
        Cc = p.shape[-1]
        rows = p.numel() // Cc
        g_loss = g_loss.contiguous().to(torch.float32)
        g_p = torch.empty_like(p)
        g_var = torch.empty_like(p)
        check(lib.sn_nll_gaussian_bwd(C.c_size_t(rows), Cc, ptr(y), ptr(p), ptr(var), C.c_float(lo), C.c_float(hi),
                                      ptr(acc), ptr(g_loss), ptr(g_p), ptr(g_var), stream_ptr()), "nll_gaussian_bwd")
        return None, g_p, g_var, None, None


def nll_gaussian_clipped(y: Tensor, p: Tensor, var: Tensor, lo: float, hi: float) -> Tensor:
    return _NllGaussian.apply(y, p, var, float(lo), float(hi))


# ---------------------------------------------------------------------------------------------------
# KL regulariser: sum over convs of l2(1.)(w_mu) + sigma_regularizer(k*k)(w_sigma) (Brats.py:56,314-320,575)
# ---------------------------------------------------------------------------------------------------
class _KlRegularizer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *params):
        lib = _lib.load()
        assert len(params) % 2 == 0
        acc = torch.zeros(1, device=params[0].device, dtype=torch.float64)
        st = stream_ptr()
        ps = []
        for i in range(0, len(params), 2):
            w = _req(params[i], "w_mu")
            ws = _req(params[i + 1], "w_sigma")
            ps += [w, ws]
            check(lib.sn_kl_regularizer_fwd(ptr(w), C.c_size_t(w.numel()), ptr(ws), ws.numel(), w.shape[0], ptr(acc),
                                            st), "kl_regularizer_fwd")
        ctx.save_for_backward(*ps)
        return acc[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        ps = ctx.saved_tensors
        scale = float(g)  # scalar coefficient (kl_factor * 0.5 * upstream); one host read per backward
        st = stream_ptr()
        grads = []
        for i in range(0, len(ps), 2):
            w, ws = ps[i], ps[i + 1]
            gw = torch.zeros_like(w)
            gws = torch.zeros_like(ws)
            check(lib.sn_kl_regularizer_bwd(ptr(w), C.c_size_t(w.numel()), ptr(ws), ws.numel(), w.shape[0],
                                            C.c_float(scale), ptr(gw), ptr(gws), st), "kl_regularizer_bwd")
            grads += [gw, gws]
        return tuple(grads)


def kl_regularizer(params: Sequence[Tuple[Tensor, Tensor]]) -> Tensor:
    flat = []
    for w, ws in params:
        flat += [w, ws]
    return _KlRegularizer.apply(*flat)
