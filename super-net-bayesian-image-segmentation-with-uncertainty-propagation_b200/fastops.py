"""FAST-mode (tcgen05) entry points of the C-ABI as thin Python functions over torch tensors.

A packed moment tensor is a torch.bfloat16 tensor of shape [n, h, w, 3, c] (planes mean_hi, mean_lo, variance
per pixel; include/supernet.h).  `PackedView` names a window of one.  No autograd tape: the backward entry points
(data gradients on packed gradient tensors, same layout) are called explicitly by engine.GradientEngine.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (SN_TC_CTA2, SN_TC_DST_F32, SN_TC_EXACT, SN_TC_IM2COL, SN_TC_KWC, SN_TC_NO_CTA2, SN_TC_NO_KWC, SN_TC_RELU,
                   SN_TC_ROWS,
                   SN_TC_UPCONV, check, ptr, sn_packed_view, sn_tc_conv_desc,
                   sn_tc_dgrad_desc, sn_tc_head_desc, sn_tc_wgrad_desc, stream_ptr)

Tensor = torch.Tensor


@dataclass
class PackedView:
    buf: Tensor                 # [n, h, w, 3, c] bfloat16, contiguous
    y0: int = 0
    x0: int = 0
    c0: int = 0

    def c_view(self) -> sn_packed_view:
        n, h, w, three, c = self.buf.shape
        assert three == 3 and self.buf.dtype == torch.bfloat16 and self.buf.is_contiguous()
        _lib.check_device(self.buf)
        return sn_packed_view(self.buf.data_ptr(), n, h, w, c, self.y0, self.x0, self.c0, 0)


def packed_empty(n: int, h: int, w: int, c: int, device) -> Tensor:
    return torch.empty((n, h, w, 3, c), device=device, dtype=torch.bfloat16)


def packed_fill(buf: Tensor, var_fill: float) -> None:
    n, h, w, _, c = buf.shape
    check(_lib.load().sn_packed_fill(ptr(buf), C.c_size_t(n * h * w), c, C.c_float(var_fill), stream_ptr()),
          "packed_fill")


def pack_moments(mu: Tensor, var: Optional[Tensor]) -> Tensor:
    """fp32 NHWC (mu, var) -> packed [n,h,w,3,c]."""
    n, h, w, c = mu.shape
    out = packed_empty(n, h, w, c, mu.device)
    check(_lib.load().sn_pack_moments(C.c_size_t(n * h * w), c, ptr(mu.contiguous()),
                                      ptr(var.contiguous() if var is not None else None), ptr(out), stream_ptr()),
          "pack_moments")
    return out


def unpack_moments(buf: Tensor, out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    n, h, w, _, c = buf.shape
    if out is not None:
        mu, var = out
    else:
        mu = torch.empty((n, h, w, c), device=buf.device, dtype=torch.float32)
        var = torch.empty_like(mu)
    check(_lib.load().sn_unpack_moments(C.c_size_t(n * h * w), c, ptr(buf), ptr(mu), ptr(var), stream_ptr()),
          "unpack_moments")
    return mu, var


def prepare_weights(w_mu: Tensor, w_sigma: Tensor, upconv: bool = False,
                    out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """HWIO fp32 -> ([3, taps, cout, cin] bf16 operands, softplus(w_sigma) [cout])."""
    k, _, cin, cout = w_mu.shape
    if out is not None:
        wp, s = out
    else:
        wp = torch.empty((3, k * k, cout, cin), device=w_mu.device, dtype=torch.bfloat16)
        s = torch.empty(cout, device=w_mu.device, dtype=torch.float32)
    check(_lib.load().sn_prepare_weights(ptr(w_mu.detach().contiguous()), ptr(w_sigma.detach().contiguous()), k, cin,
                                         cout, 1 if upconv else 0, ptr(wp), ptr(s), stream_ptr()), "prepare_weights")
    return wp, s


def conv_moments_tc(src0: PackedView, c0: int, batch: int, in_h: int, in_w: int, ksize: int, cout: int,
                    w_packed: Tensor, s: Tensor, dst: Optional[PackedView] = None, relu: bool = False,
                    upconv: bool = False, src1: Optional[PackedView] = None, c1: int = 0,
                    dst_f32: Optional[Tuple[Tensor, Tensor]] = None, im2col: bool = False,
                    rsum_out: Optional[Tensor] = None, kwc: Optional[bool] = None, cta2: Optional[bool] = None) -> None:
    """kwc / cta2: None = the library's choice, True / False force / forbid the kw-concatenated halo kernel
    (SN_TC_KWC) / the CTA-pair variant (SN_TC_CTA2)."""
    d = sn_tc_conv_desc()
    d.src[0] = src0.c_view()
    d.src[1] = (src1 if src1 is not None else src0).c_view()
    d.src_c[0], d.src_c[1] = c0, c1
    d.batch, d.in_h, d.in_w, d.ksize, d.cout = batch, in_h, in_w, ksize, cout
    d.flags = ((SN_TC_RELU if relu else 0) | (SN_TC_UPCONV if upconv else 0) | (SN_TC_DST_F32 if dst_f32 else 0) |
               (SN_TC_IM2COL if im2col else 0) | (SN_TC_KWC if kwc else (SN_TC_NO_KWC if kwc is False else 0)) |
               (SN_TC_CTA2 if cta2 else (SN_TC_NO_CTA2 if cta2 is False else 0)))
    d.w_packed = w_packed.data_ptr()
    d.s = s.data_ptr()
    if rsum_out is not None:
        d.rsum_out = rsum_out.data_ptr()
    if dst_f32 is not None:
        d.dst_mu, d.dst_var = dst_f32[0].data_ptr(), dst_f32[1].data_ptr()
    else:
        d.dst = dst.c_view()
    check(_lib.load().sn_conv_moments_fwd_tc(C.byref(d), stream_ptr()), "conv_moments_fwd_tc")


def _head_conv_desc(src: PackedView, batch: int, in_h: int, in_w: int, w_packed: Tensor, s: Tensor,
                    dst: Optional[PackedView]) -> sn_tc_conv_desc:
    d = sn_tc_conv_desc()
    d.src[0] = src.c_view()
    d.src[1] = d.src[0]
    d.src_c[0], d.src_c[1] = 32, 0
    d.batch, d.in_h, d.in_w, d.ksize, d.cout = batch, in_h, in_w, 3, 32
    d.flags = SN_TC_RELU
    d.w_packed = w_packed.data_ptr()
    d.s = s.data_ptr()
    if dst is not None:
        d.dst = dst.c_view()
    return d


def tc_head_fusable(cin: int, cin1: int, cout: int, ksize: int, relu: bool, n_labels: int) -> bool:
    """Can a conv of this shape end in the fused conv_final + softmax head (sn_tc_head_fusable)?"""
    d = sn_tc_conv_desc()
    d.src_c[0], d.src_c[1] = cin, cin1
    d.batch, d.in_h, d.in_w, d.ksize, d.cout = 1, ksize, ksize, ksize, cout
    d.flags = SN_TC_RELU if relu else 0
    return bool(_lib.load().sn_tc_head_fusable(C.byref(d), n_labels))


def conv_moments_tc_head(src: PackedView, batch: int, in_h: int, in_w: int, w_packed: Tensor, s: Tensor,
                         w_final: Tensor, ws_final: Tensor, p_out: Tensor, var_out: Tensor,
                         pre_mu: Optional[Tensor] = None, pre_var: Optional[Tensor] = None,
                         dst: Optional[PackedView] = None) -> None:
    """The last 3x3 conv (32 -> 32, ReLU) + conv_final + mysoftmax in one launch (sn_conv_moments_fwd_tc_head);
    dst=None: the 32-channel tensor is not written."""
    d = _head_conv_desc(src, batch, in_h, in_w, w_packed, s, dst)
    h = sn_tc_head_desc()
    h.n_labels = w_final.shape[-1]
    h.w_mu, h.w_sigma = w_final.data_ptr(), ws_final.data_ptr()
    h.p_out, h.var_out = p_out.data_ptr(), var_out.data_ptr()
    if pre_mu is not None:
        h.presoftmax_mu, h.presoftmax_var = pre_mu.data_ptr(), pre_var.data_ptr()
    for t in (w_final, ws_final, p_out, var_out, pre_mu, pre_var):
        _lib.check_device(t)
    check(_lib.load().sn_conv_moments_fwd_tc_head(C.byref(d), C.byref(h), stream_ptr()), "conv_moments_fwd_tc_head")


def first_conv_packed(x: Tensor, w_mu: Tensor, w_sigma: Tensor, dst: PackedView, relu: bool = True,
                      exact: bool = False, gen1: bool = False, ws: bool = False) -> None:
    """exact=True: the fp32 CUDA-core kernel instead of the tensor-core one (see SN_TC_EXACT in supernet.h);
    ws=True: the warp-specialised tensor-core kernel (SN_TC_ROWS; builder / UMMA / epilogue roles of one persistent CTA,
    TMA-store epilogue), gen1=True: force the default one-role-per-CTA kernel (SN_TC_IM2COL).  Bit-identical results,
    same measured time; both are kept for A/B measurements."""
    B, H, W, cin = x.shape
    k, _, _, cout = w_mu.shape
    v = dst.c_view()
    flags = ((SN_TC_RELU if relu else 0) | (SN_TC_EXACT if exact else 0) | (SN_TC_IM2COL if gen1 else 0) |
             (SN_TC_ROWS if ws else 0))
    check(_lib.load().sn_first_conv_fwd_packed(B, H, W, cin, cout, k, ptr(x), ptr(w_mu), ptr(w_sigma), C.byref(v),
                                               flags, stream_ptr()), "first_conv_fwd_packed")


def maxpool2_packed(src: PackedView, batch: int, in_h: int, in_w: int, c: int, dst: PackedView) -> None:
    a, b = src.c_view(), dst.c_view()
    check(_lib.load().sn_maxpool2_packed(C.byref(a), batch, in_h, in_w, c, C.byref(b), stream_ptr()),
          "maxpool2_packed")


def final_conv_softmax_packed(src: PackedView, batch: int, in_h: int, in_w: int, cin: int, w_mu: Tensor,
                              w_sigma: Tensor, p_out: Tensor, var_out: Tensor, pre_mu: Optional[Tensor] = None,
                              pre_var: Optional[Tensor] = None) -> None:
    v = src.c_view()
    n_labels = w_mu.shape[-1]
    check(_lib.load().sn_final_conv_softmax_packed(C.byref(v), batch, in_h, in_w, cin, n_labels, ptr(w_mu),
                                                   ptr(w_sigma), ptr(p_out), ptr(var_out), ptr(pre_mu), ptr(pre_var),
                                                   stream_ptr()), "final_conv_softmax_packed")


# ---- backward (data gradient) ----------------------------------------------------------------------------
def prepare_weights_bwd(w_mu: Tensor, upconv: bool = False, out: Optional[Tensor] = None) -> Tensor:
    """HWIO fp32 -> the data-gradient operands [3, taps, cin, K] bf16 (flipped / transposed filter; up-conv: K =
    (parity, cout), taps = 1)."""
    k, _, cin, cout = w_mu.shape
    shape = (3, 1, cin, 4 * cout) if upconv else (3, k * k, cin, cout)
    wt = out if out is not None else torch.empty(shape, device=w_mu.device, dtype=torch.bfloat16)
    check(_lib.load().sn_prepare_weights_bwd(ptr(w_mu.detach().contiguous()), k, cin, cout, 1 if upconv else 0,
                                             ptr(wt), stream_ptr()), "prepare_weights_bwd")
    return wt


def conv_moments_bwd_data_tc(g_out: PackedView, batch: int, in_h: int, in_w: int, ksize: int, cout: int,
                             wt_packed: Tensor, s: Tensor, in0: PackedView, g_in0: PackedView, c0: int, gate0: bool,
                             in1: Optional[PackedView] = None, g_in1: Optional[PackedView] = None, c1: int = 0,
                             gate1: bool = False, upconv: bool = False, kwc: Optional[bool] = None,
                             cta2: Optional[bool] = None) -> None:
    d = sn_tc_dgrad_desc()
    d.g_out = g_out.c_view()
    d.in_[0], d.g_in[0] = in0.c_view(), g_in0.c_view()
    d.in_[1] = (in1 if in1 is not None else in0).c_view()
    d.g_in[1] = (g_in1 if g_in1 is not None else g_in0).c_view()
    d.in_c[0], d.in_c[1] = c0, c1
    d.gate[0], d.gate[1] = int(gate0), int(gate1)
    d.batch, d.in_h, d.in_w, d.ksize, d.cout = batch, in_h, in_w, ksize, cout
    d.flags = ((SN_TC_UPCONV if upconv else 0) | (SN_TC_KWC if kwc else (SN_TC_NO_KWC if kwc is False else 0)) |
               (SN_TC_CTA2 if cta2 else (SN_TC_NO_CTA2 if cta2 is False else 0)))
    d.wt_packed = wt_packed.data_ptr()
    d.s = s.data_ptr()
    check(_lib.load().sn_conv_moments_bwd_data_tc(C.byref(d), stream_ptr()), "conv_moments_bwd_data_tc")


def maxpool2_bwd_packed(inp: PackedView, batch: int, in_h: int, in_w: int, c: int, g_out: PackedView,
                        g_in: PackedView, keep: Tuple[int, int, int, int] = (0, 0, 0, 0)) -> None:
    a, b, g = inp.c_view(), g_out.c_view(), g_in.c_view()
    check(_lib.load().sn_maxpool2_bwd_packed(C.byref(a), batch, in_h, in_w, c, C.byref(b), C.byref(g), keep[0],
                                             keep[1], keep[2], keep[3], stream_ptr()), "maxpool2_bwd_packed")


def head_bwd_packed(inp: PackedView, batch: int, in_h: int, in_w: int, cin: int, w_mu: Tensor, w_sigma: Tensor,
                    y: Tensor, clip: Tuple[float, float], acc: Tensor, loss_scale: float, g_in: PackedView,
                    logit_grads: Optional[Tuple[Tensor, Tensor, Tensor]] = None) -> None:
    a, g = inp.c_view(), g_in.c_view()
    n_labels = w_mu.shape[-1]
    lg = logit_grads if logit_grads is not None else (None, None, None)
    check(_lib.load().sn_head_bwd_packed(C.byref(a), batch, in_h, in_w, cin, n_labels, ptr(w_mu), ptr(w_sigma),
                                         ptr(y), C.c_float(clip[0]), C.c_float(clip[1]), ptr(acc),
                                         C.c_float(loss_scale), C.byref(g), ptr(lg[0]), ptr(lg[1]), ptr(lg[2]),
                                         stream_ptr()), "head_bwd_packed")


def first_conv_bwd_data_packed(x: Tensor, w_mu: Tensor, w_sigma: Tensor, g_out: PackedView, g_x: Tensor) -> None:
    B, H, W, cin = x.shape
    k, _, _, cout = w_mu.shape
    g = g_out.c_view()
    check(_lib.load().sn_first_conv_bwd_data_packed(B, H, W, cin, cout, k, ptr(x), ptr(w_mu), ptr(w_sigma),
                                                    C.byref(g), ptr(g_x), stream_ptr()), "first_conv_bwd_data_packed")


def nll_gaussian_fwd(y: Tensor, p: Tensor, var: Tensor, clip: Tuple[float, float], acc: Tensor, loss: Tensor) -> None:
    """sn_nll_gaussian_fwd on preallocated workspaces (acc: 2 doubles, loss: 1 float): graph-capturable."""
    rows, c = p.numel() // p.shape[-1], p.shape[-1]
    check(_lib.load().sn_nll_gaussian_fwd(C.c_size_t(rows), c, ptr(y), ptr(p), ptr(var), C.c_float(clip[0]),
                                          C.c_float(clip[1]), ptr(acc), ptr(loss), stream_ptr()), "nll_gaussian_fwd")


# ---- backward (weight gradient) ---------------------------------------------------------------------------
def wgrad_workspace(ksize: int, cin: int, cout: int, device) -> Tensor:
    n = _lib.load().sn_wgrad_workspace_bytes(ksize, cin, cout)
    return torch.empty(n // 4, device=device, dtype=torch.float32)


def conv_moments_bwd_weight_tc(g_out: PackedView, batch: int, in_h: int, in_w: int, ksize: int, cout: int,
                               in0: PackedView, c0: int, rsum: Tensor, w_mu: Tensor, w_sigma: Tensor,
                               workspace: Tensor, g_w_mu: Tensor, g_w_sigma: Tensor,
                               in1: Optional[PackedView] = None, c1: int = 0, upconv: bool = False,
                               im2col: Optional[bool] = None) -> None:
    """im2col: None = the library's choice, True / False force the general / the row-halo kernel."""
    d = sn_tc_wgrad_desc()
    d.g_out = g_out.c_view()
    d.in_[0] = in0.c_view()
    d.in_[1] = (in1 if in1 is not None else in0).c_view()
    d.in_c[0], d.in_c[1] = c0, c1
    d.batch, d.in_h, d.in_w, d.ksize, d.cout = batch, in_h, in_w, ksize, cout
    d.flags = (SN_TC_UPCONV if upconv else 0) | (SN_TC_IM2COL if im2col else (SN_TC_ROWS if im2col is False else 0))
    d.rsum, d.w_mu, d.w_sigma = rsum.data_ptr(), w_mu.data_ptr(), w_sigma.data_ptr()
    d.workspace, d.g_w_mu, d.g_w_sigma = workspace.data_ptr(), g_w_mu.data_ptr(), g_w_sigma.data_ptr()
    check(_lib.load().sn_conv_moments_bwd_weight_tc(C.byref(d), stream_ptr()), "conv_moments_bwd_weight_tc")


def first_conv_rsum(x: Tensor, ksize: int, rsum: Tensor) -> None:
    B, H, W, cin = x.shape
    check(_lib.load().sn_first_conv_rsum(B, H, W, cin, ksize, ptr(x), ptr(rsum), stream_ptr()), "first_conv_rsum")


def first_conv_bwd_weight_packed(x: Tensor, w_mu: Tensor, w_sigma: Tensor, g_out: PackedView, workspace: Tensor,
                                 g_w_mu: Tensor, g_w_sigma: Tensor) -> None:
    B, H, W, cin = x.shape
    k, _, _, cout = w_mu.shape
    g = g_out.c_view()
    check(_lib.load().sn_first_conv_bwd_weight_packed(B, H, W, cin, cout, k, ptr(x), ptr(w_sigma), C.byref(g),
                                                      ptr(workspace), ptr(g_w_mu), ptr(g_w_sigma), stream_ptr()),
          "first_conv_bwd_weight_packed")


def final_conv_bwd_weight_packed(inp: PackedView, batch: int, in_h: int, in_w: int, cin: int, w_mu: Tensor,
                                 w_sigma: Tensor, logit_grads: Tuple[Tensor, Tensor, Tensor], workspace: Tensor,
                                 g_w_mu: Tensor, g_w_sigma: Tensor) -> None:
    a = inp.c_view()
    check(_lib.load().sn_final_conv_bwd_weight_packed(C.byref(a), batch, in_h, in_w, cin, w_mu.shape[-1], ptr(w_mu),
                                                      ptr(w_sigma), ptr(logit_grads[0]), ptr(logit_grads[1]),
                                                      ptr(logit_grads[2]), ptr(workspace), ptr(g_w_mu),
                                                      ptr(g_w_sigma), stream_ptr()), "final_conv_bwd_weight_packed")


def head_bwd_upstream_packed(inp: PackedView, batch: int, in_h: int, in_w: int, cin: int, w_mu: Tensor,
                             w_sigma: Tensor, g_p: Tensor, g_var_out: Optional[Tensor], g_in: PackedView) -> None:
    a, g = inp.c_view(), g_in.c_view()
    check(_lib.load().sn_head_bwd_upstream_packed(C.byref(a), batch, in_h, in_w, cin, w_mu.shape[-1], ptr(w_mu),
                                                  ptr(w_sigma), ptr(g_p), ptr(g_var_out), C.byref(g), stream_ptr()),
          "head_bwd_upstream_packed")
