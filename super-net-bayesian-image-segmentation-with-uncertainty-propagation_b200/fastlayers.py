"""FAST mode behind the reference's per-layer call surface.

The reference composes its network one layer call at a time, every call taking and returning the pair
``(mean, sigma)`` (Brats.py:379-455).  In FAST mode the pair travels as ONE handle, `PackedMoments`, in the packed
bf16 layout the tensor-core kernels read (include/supernet.h: planes mean_hi, mean_lo, variance per pixel): a layer
returns ``(h, h)`` so user code written as ``m, s = layer(m, s)`` runs unchanged, and the next layer recognises the
handle.  Layer-by-layer code -- the two built-in graphs, a deeper or shallower U-Net, any other wiring of these
layers -- therefore runs `sn_conv_moments_fwd_tc` (tcgen05 / TMEM / TMA), not the CUDA-core FP32 kernels.

A handle is lazy: what the reference does with separate ops stays address arithmetic / epilogue flags here, exactly
like in engine.InferenceEngine, but decided at run time from how the handle is consumed:
  myReLU after a conv            -> the conv is (re)issued with its ReLU epilogue flag       (Brats.py:233-238)
  mypadding after conv / pool    -> the producer writes the interior of a pre-filled buffer   (Brats.py:159-163)
  myupsampling + 2x2 conv        -> four parity GEMMs scattered to (2y+a, 2x+b)              (Brats.py:178-203,414-415)
  myConc                         -> the consuming conv reads two source windows             (Brats.py:247-261)
  1x1 conv to n_labels + mysoftmax -> one kernel                                             (Brats.py:454-455)
  3x3 conv 32 -> 32 + myReLU + that head -> ONE kernel (sn_conv_moments_fwd_tc_head)          (Brats.py:451-455)
Anything outside those patterns still works through general packed kernels (sn_relu_packed, window copies).
Inference only: the handles carry no autograd history (gradients: mode='fp32', or engine.GradientEngine).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Tuple

import torch

from . import _lib
from . import fastops as F
from .fastops import PackedView

Tensor = torch.Tensor


def _relu_copy(src: PackedView, batch: int, h: int, w: int, c: int, dst: PackedView, gate: bool) -> None:
    a, b = src.c_view(), dst.c_view()
    _lib.check(_lib.load().sn_relu_packed(C.byref(a), batch, h, w, c, C.byref(b), 1 if gate else 0, _lib.stream_ptr()),
               "relu_packed")


class PackedMoments:
    """(mean, variance) of one activation tensor, logical shape [B, h, w, c], in the packed layout.

    Either materialised (`view` names its window of a packed buffer) or pending (`_op(dst, relu)` writes it into a
    window when somebody needs it).  Pending handles are immutable descriptions: myReLU / mypadding return NEW handles
    that wrap the same pending op with another flag / destination, so a handle kept aside (a skip connection) is never
    changed by what happens to its successors."""

    is_packed_moments = True

    def __init__(self, shape: Tuple[int, int, int, int], device, view: Optional[PackedView] = None,
                 op: Optional[Callable[[PackedView, bool], None]] = None, relu: bool = False, relu_fusable: bool = False):
        self.shape = tuple(int(v) for v in shape)
        self.device = device
        self._view = view
        self._op = op
        self._relu = relu
        self._relu_fusable = relu_fusable
        self._conv_info = None          # pending single-source conv: (layer, source handle, h, w) -- what the fused head needs

    # ---- materialisation -------------------------------------------------------------------------------------
    def materialize(self, dst: Optional[PackedView] = None) -> PackedView:
        """The window holding this tensor; computes it (into `dst` when given) on first use."""
        B, h, w, c = self.shape
        if self._view is None:
            target = dst if dst is not None else PackedView(F.packed_empty(B, h, w, c, self.device))
            self._op(target, self._relu)
            self._view, self._op = target, None
            return target
        if dst is not None:                          # already in memory elsewhere: window copy
            _relu_copy(self._view, B, h, w, c, dst, False)
            return dst
        return self._view

    def unpack(self) -> Tuple[Tensor, Tensor]:
        """fp32 NHWC (mean, variance) -- what the reference layer would have returned here."""
        B, h, w, c = self.shape
        v = self.materialize()
        buf = v.buf
        if (v.y0, v.x0, v.c0) != (0, 0, 0) or tuple(buf.shape) != (B, h, w, 3, c):
            buf = buf[:B, v.y0:v.y0 + h, v.x0:v.x0 + w, :, v.c0:v.c0 + c].contiguous()
        return F.unpack_moments(buf)

    # tensor-like conveniences so user code that inspects shapes keeps working
    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    def __repr__(self) -> str:
        state = "materialised" if self._view is not None else "pending"
        return f"PackedMoments(shape={self.shape}, {state})"


class Unpooled:
    """myupsampling's output (zero-stuffed to 2H+1, Brats.py:178-203) as a marker around its source: the 2x2 conv that
    follows consumes the source directly as four parity GEMMs."""

    is_packed_moments = True

    def __init__(self, src: PackedMoments):
        self.src = src
        B, h, w, c = src.shape
        self.shape = (B, 2 * h + 1, 2 * w + 1, c)
        self.device = src.device

    def dense(self) -> PackedMoments:
        """The zero-stuffed tensor itself (only when something other than a 2x2 conv consumes it)."""
        B, h, w, c = self.src.shape
        v = self.src.materialize()
        buf = F.packed_empty(B, 2 * h + 1, 2 * w + 1, c, self.device)
        buf.zero_()                                                      # mean 0, variance 0: "certain zeros"
        buf[:, 1::2, 1::2] = v.buf[:B, v.y0:v.y0 + h, v.x0:v.x0 + w, :, v.c0:v.c0 + c]
        return PackedMoments(self.shape, self.device, view=PackedView(buf))


class Concat:
    """myConc's output (Brats.py:247-261): [decoder, centre-cropped encoder] along channels, as two source windows."""

    is_packed_moments = True

    def __init__(self, dec: PackedMoments, enc: PackedMoments):
        B, h, w, cd = dec.shape
        _, he, we, ce = enc.shape
        if he < h or we < w:
            raise RuntimeError("myConc: the encoder tensor is smaller than the decoder tensor")
        self.dec, self.enc = dec, enc
        self.oy, self.ox = (he - h) // 2, (we - w) // 2                  # crop_tensor, Brats_functions.py:518-526
        self.shape = (B, h, w, cd + ce)
        self.device = dec.device


class PendingHead:
    """A 1x1 moment conv to n_labels < 32 channels (conv_final, Brats.py:454): too thin for the tensor cores; it runs
    fused with the softmax that follows it (sn_final_conv_softmax_packed)."""

    is_packed_moments = True

    def __init__(self, src: PackedMoments, layer):
        self.src, self.layer = src, layer
        B, h, w, _ = src.shape
        self.shape = (B, h, w, layer.kernel_num)
        self.device = src.device

    def run(self, want_presoftmax: bool):
        B, h, w, cin = self.src.shape
        Cn = self.layer.kernel_num
        wm, ws = self.layer.weights()
        p = torch.empty((B, h * w, Cn), device=self.device, dtype=torch.float32)
        v = torch.empty_like(p)
        pre = (torch.empty_like(p), torch.empty_like(p)) if want_presoftmax else (None, None)
        info = self.src._conv_info if isinstance(self.src, PackedMoments) else None
        if (info is not None and self.src._view is None and self.src._relu and os.environ.get("SN_FUSE_HEAD", "1") != "0"
                and F.tc_head_fusable(info[1].shape[3], 0, cin, info[0].kernel_size, True, Cn)):
            # the producer is a still-pending 3x3 conv 32 -> 32 with its ReLU: run it WITH the head in its epilogue (the
            # engines' last launch); the 32-channel tensor is not written -- the handle stays pending and can still
            # materialise itself if somebody else asks for it.  Bit-identical to the two-kernel path.
            layer, src, hin, win = info
            wp, s = _prepared(layer, False)
            sv = src.materialize()
            F.conv_moments_tc_head(sv, B, hin, win, wp, s, wm.detach(), ws.detach(), p, v, pre[0], pre[1])
            return p, v, pre
        F.final_conv_softmax_packed(self.src.materialize(), B, h, w, cin, wm.detach(), ws.detach(), p, v, pre[0], pre[1])
        return p, v, pre

    def unpack(self) -> Tuple[Tensor, Tensor]:
        B, h, w, Cn = self.shape
        _, _, pre = self.run(True)
        return pre[0].reshape(B, h, w, Cn), pre[1].reshape(B, h, w, Cn)


def is_handle(x) -> bool:
    return getattr(x, "is_packed_moments", False)


def _same_pair(mu, sigma) -> None:
    if sigma is not mu:
        raise RuntimeError("FAST mode: pass the (mean, sigma) pair a layer returned, unchanged, to the next layer")


# -------------------------------------------------------------------------------------------------------------
# the layers
# -------------------------------------------------------------------------------------------------------------
def _prepared(layer, upconv: bool):
    """bf16 operands of a conv layer, re-derived when its parameters were written in place."""
    wm, ws = layer.weights()
    stamp = (wm._version, ws._version, wm.data_ptr(), upconv)
    cache = getattr(layer, "_fast_prepared", None)
    if cache is None or cache[0] != stamp:
        out = cache[1] if cache is not None and cache[0][2:] == stamp[2:] else None
        cache = (stamp, F.prepare_weights(wm, ws, upconv=upconv, out=out))
        layer._fast_prepared = cache
    return cache[1]


def conv_input(layer, x: Tensor) -> Tuple[PackedMoments, PackedMoments]:
    """myConv_input.call (Brats.py:65-76) on the fp32 NHWC image."""
    if not x.is_cuda:
        raise RuntimeError("FAST mode needs CUDA tensors (there is no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    B, H, W, _ = x.shape
    k = layer.kernel_size
    wm, ws = layer.weights()

    def op(dst: PackedView, relu: bool) -> None:
        F.first_conv_packed(x, wm.detach(), ws.detach(), dst, relu=relu)

    h = PackedMoments((B, H - k + 1, W - k + 1, layer.kernel_num), x.device, op=op, relu_fusable=True)
    return h, h


def conv_intermediate(layer, mu, sigma) -> Tuple[object, object]:
    """myConv_intermediate.call (Brats.py:118-137) on a handle (plain, up-sampled or concatenated)."""
    _same_pair(mu, sigma)
    k, cout = layer.kernel_size, layer.kernel_num
    upconv = isinstance(mu, Unpooled) and k == 2
    if isinstance(mu, Unpooled) and not upconv:
        mu = mu.dense()
    if isinstance(mu, PendingHead):
        raise RuntimeError("a conv cannot consume an un-materialised n_labels-channel head")
    if isinstance(mu, Concat):
        srcs = [(mu.dec, 0, 0), (mu.enc, mu.oy, mu.ox)]
    else:
        srcs = [((mu.src if upconv else mu), 0, 0)]
    B, h, w, _ = (mu.src.shape if upconv else mu.shape)
    cin = sum(s.shape[3] for s, _, _ in srcs)
    if not layer.built:
        layer._build(cin, mu.device)
    if any(s.shape[3] % 32 for s, _, _ in srcs):
        raise RuntimeError(f"FAST mode: input channel counts {[s.shape[3] for s, _, _ in srcs]} must be multiples of 32")
    if cout % 32 != 0:
        if k == 1 and cout <= 8 and len(srcs) == 1 and not upconv:
            hd = PendingHead(srcs[0][0], layer)
            return hd, hd
        raise RuntimeError(f"FAST mode: kernel_num {cout} must be a multiple of 32 (or a 1x1 head of <= 8 classes)")
    out_shape = (B, 2 * h, 2 * w, cout) if upconv else (B, h - k + 1, w - k + 1, cout)

    def op(dst: PackedView, relu: bool) -> None:
        wp, s = _prepared(layer, upconv)
        views = []
        for src, oy, ox in srcs:
            v = src.materialize()
            views.append((PackedView(v.buf, v.y0 + oy, v.x0 + ox, v.c0), src.shape[3]))
        F.conv_moments_tc(views[0][0], views[0][1], B, h, w, k, cout, wp, s, dst=dst, relu=relu, upconv=upconv,
                          src1=views[1][0] if len(views) > 1 else None, c1=views[1][1] if len(views) > 1 else 0)

    hdl = PackedMoments(out_shape, mu.device, op=op, relu=bool(layer.fuse_relu), relu_fusable=not layer.fuse_relu)
    if len(srcs) == 1 and not upconv:
        hdl._conv_info = (layer, srcs[0][0], h, w)
    return hdl, hdl


def relu(mu, sigma):
    """myReLU.call (Brats.py:233-238)."""
    _same_pair(mu, sigma)
    if isinstance(mu, (Unpooled, Concat, PendingHead)):
        raise RuntimeError("FAST mode: myReLU directly after myupsampling / myConc / the n_labels head is not supported")
    if mu._view is None and mu._relu_fusable and not mu._relu:
        out = PackedMoments(mu.shape, mu.device, op=mu._op, relu=True)          # epilogue flag of the pending conv
        out._conv_info = mu._conv_info
        return out, out
    if mu._relu:                                                                # already gated: idempotent
        return mu, mu
    B, h, w, c = mu.shape

    def op(dst: PackedView, _relu: bool) -> None:
        _relu_copy(mu.materialize(), B, h, w, c, dst, True)

    out = PackedMoments(mu.shape, mu.device, op=op)
    return out, out


def padding(layer, mu, sigma):
    """mypadding.call (Brats.py:159-163): the producer writes the interior of a buffer pre-filled with mean 0 /
    variance sigma_fill."""
    _same_pair(mu, sigma)
    if isinstance(mu, Unpooled):
        mu = mu.dense()
    if isinstance(mu, (Concat, PendingHead)):
        raise RuntimeError("FAST mode: mypadding directly after myConc / the n_labels head is not supported")
    a, b = layer.pad_size
    B, h, w, c = mu.shape
    fill = layer.sigma_fill

    def op(dst: PackedView, _relu: bool) -> None:
        if (dst.y0, dst.x0, dst.c0) != (0, 0, 0) or tuple(dst.buf.shape) != (B, h + a + b, w + a + b, 3, c):
            own = PackedView(F.packed_empty(B, h + a + b, w + a + b, c, mu.device))
            op(own, False)
            _relu_copy(own, B, h + a + b, w + a + b, c, dst, False)
            return
        F.packed_fill(dst.buf, fill)
        mu.materialize(PackedView(dst.buf, a, a, 0))

    out = PackedMoments((B, h + a + b, w + a + b, c), mu.device, op=op)
    return out, out


def maxpooling(mu, sigma):
    """mymaxpooling.call (Brats.py:171-174, 206-216): 2x2/2, variance at the arg-max."""
    _same_pair(mu, sigma)
    if isinstance(mu, Unpooled):
        mu = mu.dense()
    if isinstance(mu, (Concat, PendingHead)):
        raise RuntimeError("FAST mode: mymaxpooling directly after myConc / the n_labels head is not supported")
    B, h, w, c = mu.shape

    def op(dst: PackedView, _relu: bool) -> None:
        F.maxpool2_packed(mu.materialize(), B, h, w, c, dst)

    out = PackedMoments((B, (h + 1) // 2, (w + 1) // 2, c), mu.device, op=op)
    return out, out


def upsampling(mu, sigma):
    """myupsampling.call (Brats.py:145-148)."""
    _same_pair(mu, sigma)
    if not isinstance(mu, PackedMoments):
        raise RuntimeError("FAST mode: myupsampling expects a plain moment tensor")
    out = Unpooled(mu)
    return out, out


def conc(muD, SigmaD, muE, SigmaE):
    """myConc.call (Brats.py:247-261)."""
    _same_pair(muD, SigmaD)
    _same_pair(muE, SigmaE)
    if not isinstance(muD, PackedMoments) or not isinstance(muE, PackedMoments):
        raise RuntimeError("FAST mode: myConc expects two plain moment tensors")
    out = Concat(muD, muE)
    return out, out


def softmax(mu, sigma):
    """mysoftmax.call (Brats.py:269-283) -> fp32 ([B, HW, C], [B, HW, C])."""
    _same_pair(mu, sigma)
    if isinstance(mu, PendingHead):
        p, v, _ = mu.run(False)
        return p, v
    from . import ops
    m, s = (mu.dense() if isinstance(mu, Unpooled) else mu).unpack()
    B, Cc = m.shape[0], m.shape[-1]
    p, v = ops.softmax_moments(m, s)
    return p.reshape(B, -1, Cc), v.reshape(B, -1, Cc)
