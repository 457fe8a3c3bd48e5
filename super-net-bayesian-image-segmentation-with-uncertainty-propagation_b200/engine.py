"""FAST-mode inference engine: the whole Density_prop_with_pad_UNET forward (Brats.py:377-457 /
Hippocampus.py:373-421) as a fixed sequence of C-ABI calls over preallocated packed buffers.

What the reference does with separate ops is address arithmetic here:
  * myReLU (Brats.py:233-238)             -> conv epilogue flag
  * mypadding (Brats.py:159-163)          -> the producer writes into the interior of a buffer whose border
                                            was filled once (mean 0, variance sigma_fill)
  * myConc + crop_tensor (Brats.py:247-261) -> the conv reads two source windows (decoder, cropped encoder)
  * myupsampling + 2x2 conv (Brats.py:414-415) -> four parity GEMMs scattered to (2y+a, 2x+b)
  * conv_final + mysoftmax (Brats.py:454-455) -> one kernel
Nothing is allocated and nothing synchronises inside run(), so the sequence is CUDA-graph capturable; the
engine captures it on first use (graph=True) to remove ~40 launches of host latency per forward.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import os

import torch

from . import fastops as F
from .fastops import PackedView

Tensor = torch.Tensor


def _on_device(fn):
    """Run an engine method with the engine's GPU as the current device: the C-ABI launches on the current device's
    stream, so a model / input on cuda:1 must not be driven while cuda:0 is current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


def _capture_graph(device, launch: Callable[[], None]) -> "torch.cuda.CUDAGraph":
    """Warm `launch` up on a side stream (lazy module loading, shared-memory attributes), then capture it.
    Garbage is collected BEFORE and the collector is paused DURING the capture: freeing an older engine (its CUDA
    graph, its private memory pool) in the middle of a capture is an operation CUDA does not permit and it
    invalidates the capture (seen in the test suite, where earlier engines die at arbitrary times)."""
    import gc
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        launch()
    torch.cuda.current_stream().wait_stream(side)
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            launch()
    finally:
        if was_enabled:
            gc.enable()
    return g


class InferenceEngine:
    def __init__(self, model, batch: int, in_h: int, in_w: int, in_c: int, device, graph: bool = True,
                 keep_presoftmax: bool = True, fuse_head: Optional[bool] = None):
        """fuse_head: end the forward inside the last 3x3 conv's epilogue (sn_conv_moments_fwd_tc_head: conv_final +
        softmax there, the last 32-channel tensor never written; bit-identical outputs).  None = whenever the shape
        allows and nothing needs that tensor (a GradientEngine does: it passes False); SN_FUSE_HEAD=0 turns it off."""
        self.model = model
        self.fuse_head = fuse_head
        self.keep_presoftmax = keep_presoftmax
        self.shape = (batch, in_h, in_w, in_c)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # "cuda" -> the current device, so that
            self.device = torch.device("cuda", torch.cuda.current_device())   # matches() can compare with x.device
        self.use_graph = graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._steps: List[Callable[[], None]] = []
        self.step_names: List[str] = []
        self.records: List[dict] = []        # one entry per forward launch (what GradientEngine walks backwards)
        self._weights_version = None
        self._build()

    def matches(self, x: Tensor) -> bool:
        return tuple(x.shape) == self.shape and x.device == self.device

    # ------------------------------------------------------------------------------------------------
    def _prepare_weights(self) -> None:
        """softplus / W^2 / bf16 operand split, once per weight update (sn_prepare_weights)."""
        m = self.model
        self.prepared = {}
        for name in m.conv_names:
            if name in ("conv_input", "conv_final"):
                continue
            w, ws = getattr(m, name).weights()
            self.prepared[name] = F.prepare_weights(w, ws, upconv=name.endswith("conv2x2"))

    @_on_device
    def refresh_weights(self) -> None:
        """After a weight update: re-derive the bf16 operands IN PLACE (a captured graph keeps its pointers)."""
        m = self.model
        for name in m.conv_names:
            if name in ("conv_input", "conv_final"):
                continue
            w, ws = getattr(m, name).weights()
            F.prepare_weights(w, ws, upconv=name.endswith("conv2x2"), out=self.prepared[name])
        self._weights_version = self._weights_stamp()

    def _weights_stamp(self):
        """Changes whenever any conv weight is written in place -- optimizer.step(), load_state_dict(),
        layer.set_weights(), load_weight_dict(): torch bumps Tensor._version on every in-place write, so the stamp
        is derived from the parameters themselves, not from callers remembering to bump a counter."""
        m = self.model
        return (getattr(m, "_weights_version", 0),) + tuple(p._version for c in m.convs() for p in c.weights())

    def sync_weights(self) -> None:
        """Refresh the derived operands (bf16 splits, W^2, softplus) if any weight changed since they were made."""
        if self._weights_version != self._weights_stamp():
            self.refresh_weights()

    @_on_device
    def _build(self) -> None:
        m = self.model
        B, H, W, Cin = self.shape
        dev = self.device
        n = m.n_kernels
        L = m.levels
        fill = m.sigma_fill
        if not all(c.built for c in m.convs()):
            m.build_with_input(Cin, dev)
        self._prepare_weights()
        self._weights_version = self._weights_stamp()
        steps = self._steps
        self.skip_window = {}
        self.x_in = torch.empty(self.shape, device=dev, dtype=torch.float32)

        def new(h, w, c, prefill=None):
            t = F.packed_empty(B, h, w, c, dev)
            if prefill is not None:
                F.packed_fill(t, prefill)
            return t

        def conv(name, src: PackedView, c0, h, w, k, dst: PackedView, relu, src1=None, c1=0, upconv=False):
            wp, s = self.prepared[name]
            cout = getattr(m, name).kernel_num
            self.step_names.append(name)
            rs = None
            if getattr(self, "want_rsum", False):          # training: keep box_k(sum_c mu^2 + var) for d/d w_sigma
                rs = torch.empty((B, h, w) if upconv else (B, h - k + 1, w - k + 1), device=dev, dtype=torch.float32)
            self.records.append(dict(kind="conv", name=name, src0=src, c0=c0, src1=src1, c1=c1, h=h, w=w, k=k,
                                     cout=cout, dst=dst, relu=relu, upconv=upconv, rsum=rs))
            steps.append(lambda: F.conv_moments_tc(src, c0, B, h, w, k, cout, wp, s, dst=dst, relu=relu,
                                                   upconv=upconv, src1=src1, c1=c1, rsum_out=rs))

        # ---- encoder --------------------------------------------------------------------------------
        h, w = H - 2, W - 2
        a0 = new(h, w, n)
        w_in, ws_in = m.conv_input.weights()
        self.step_names.append("conv_input")
        self.records.append(dict(kind="first", dst=PackedView(a0)))
        exact0 = bool(getattr(self, "exact_first_conv", False))     # GradientEngine: fp32-exact gates in the first layer
        steps.append(lambda: F.first_conv_packed(self.x_in, w_in, ws_in, PackedView(a0), relu=True, exact=exact0))
        skip = new(h - 2, w - 2, n)
        conv("conv1", PackedView(a0), n, h, w, 3, PackedView(skip), True)
        h, w = h - 2, w - 2
        skips: List[Tuple[Tensor, int, int, int]] = [(skip, h, w, n)]
        cur, c = skip, n
        ci = 2
        for lvl in range(1, L + 1):
            ph, pw = (h + 1) // 2, (w + 1) // 2
            if m.variant == "brats" and lvl == L:
                pooled = new(ph + 1, pw + 1, c, prefill=fill)          # mypad1 [1,0] (Brats.py:407)
                pview = PackedView(pooled, 1, 1, 0)
                ph, pw = ph + 1, pw + 1
            else:
                pooled = new(ph, pw, c)
                pview = PackedView(pooled)
            self.step_names.append(f"pool{lvl}")
            self.records.append(dict(kind="pool", src=PackedView(cur), h=h, w=w, c=c, dst=pview))
            steps.append(lambda s=PackedView(cur), hh=h, ww=w, cc=c, d=pview: F.maxpool2_packed(s, B, hh, ww, cc, d))
            h, w = ph, pw
            cur = pooled
            for j in range(2):
                name = f"conv{ci}"
                cout = getattr(m, name).kernel_num
                out = new(h - 2, w - 2, cout)
                conv(name, PackedView(cur), c, h, w, 3, PackedView(out), True)
                cur, c, h, w = out, cout, h - 2, w - 2
                ci += 1
            if lvl < L:
                skips.append((cur, h, w, c))
        # ---- decoder --------------------------------------------------------------------------------
        for d in range(1, L + 1):
            enc, eh, ew, ec = skips[L - d]
            cu = getattr(m, f"up{d}_conv2x2").kernel_num
            uh, uw = 2 * h, 2 * w                                        # unpool (2h+1) then 2x2 VALID -> 2h
            up = new(uh + 6, uw + 6, cu, prefill=fill)                   # mypad_up6 [3,3] (Brats.py:416)
            conv(f"up{d}_conv2x2", PackedView(cur), c, h, w, 2, PackedView(up, 3, 3, 0), False, upconv=True)
            h, w = uh + 6, uw + 6
            oy, ox = (eh - h) // 2, (ew - w) // 2                        # crop_tensor (Brats_functions.py:518-526)
            if oy < 0 or ox < 0:
                raise RuntimeError("input too small: the skip tensor is smaller than the decoder tensor")
            c1n = getattr(m, f"up{d}_conv1").kernel_num
            mid = new(h - 2 + 4, w - 2 + 4, c1n, prefill=fill)           # mypad [2,2] (Brats.py:420)
            conv(f"up{d}_conv1", PackedView(up), cu, h, w, 3, PackedView(mid, 2, 2, 0), True,
                 src1=PackedView(enc, oy, ox, 0), c1=ec)
            self.skip_window[enc.data_ptr()] = (oy, ox, h, w)            # where the decoder reads the skip tensor
            h, w = h + 2, w + 2
            c2n = getattr(m, f"up{d}_conv2").kernel_num
            if d == L and self._want_fused_head(c1n, c2n):
                break                                                    # the head step below runs this conv as well
            out = new(h - 2, w - 2, c2n)
            conv(f"up{d}_conv2", PackedView(mid), c1n, h, w, 3, PackedView(out), True)
            cur, c, h, w = out, c2n, h - 2, w - 2
        # ---- head -----------------------------------------------------------------------------------
        C = m.n_labels
        wf, wsf = m.conv_final.weights()
        pre = lambda: (self.pre_m, self.pre_v) if self.keep_presoftmax else (None, None)
        if self.head_fused:
            # up{L}_conv2 + ReLU + conv_final + softmax in ONE launch; its 32-channel output is never written
            name = f"up{L}_conv2"
            wp, s = self.prepared[name]
            oh, ow = h - 2, w - 2
            self._alloc_outputs(B, oh, ow, C, dev)
            self.step_names.append(name + "+conv_final")
            dstv = None
            if getattr(self, "keep_last", False):
                # a gradient engine needs the 32-channel tensor (the head's backward recomputes conv_final from it):
                # the fused launch stores it as well, and the records read as if the two layers had run separately
                dstv = PackedView(new(oh, ow, c2n))
                self.records.append(dict(kind="conv", name=name, src0=PackedView(mid), c0=c1n, src1=None, c1=0, h=h, w=w,
                                         k=3, cout=c2n, dst=dstv, relu=True, upconv=False, rsum=None))
                self.records.append(dict(kind="head", src=dstv, h=oh, w=ow, c=c2n))
            else:
                self.records.append(dict(kind="conv+head", name=name, src0=PackedView(mid), c0=c1n, h=h, w=w, k=3,
                                         cout=c2n))
            steps.append(lambda src=PackedView(mid), hh=h, ww=w, dstv=dstv: F.conv_moments_tc_head(
                src, B, hh, ww, wp, s, wf, wsf, self.p, self.v, *pre(), dst=dstv))
            h, w = oh, ow
        else:
            self._alloc_outputs(B, h, w, C, dev)
            last, lc = cur, c
            self.step_names.append("conv_final")
            self.records.append(dict(kind="head", src=PackedView(last), h=h, w=w, c=lc))
            steps.append(lambda: F.final_conv_softmax_packed(PackedView(last), B, h, w, lc, wf, wsf, self.p, self.v,
                                                             *pre()))
        self.out_hw = (h, w)
        self.n_launches = len(steps)

    def _want_fused_head(self, cin: int, cout: int) -> bool:
        want = self.fuse_head
        if want is None:
            want = os.environ.get("SN_FUSE_HEAD", "1") != "0" and not getattr(self, "want_rsum", False)
        self.head_fused = bool(want) and F.tc_head_fusable(cin, 0, cout, 3, True, self.model.n_labels)
        if self.fuse_head and not self.head_fused:
            raise RuntimeError("fuse_head=True, but the last conv of this model cannot end in the fused head "
                               "(needs 32 -> 32 channels, k = 3, 2..5 labels)")
        return self.head_fused

    def _alloc_outputs(self, B: int, h: int, w: int, C: int, dev) -> None:
        # both output maps live in ONE buffer (p = pv[0], v = pv[1]): the host-facing pipeline moves them with a single
        # device-to-host copy per batch
        self.pv = torch.empty((2, B, h * w, C), device=dev, dtype=torch.float32)
        self.p, self.v = self.pv[0], self.pv[1]
        self.pre_m = torch.empty_like(self.p)
        self.pre_v = torch.empty_like(self.p)

    # ------------------------------------------------------------------------------------------------
    def _launch_all(self) -> None:
        for s in self._steps:
            s()

    @_on_device
    def forward_resident(self) -> Tuple[Tensor, Tensor]:
        """Run the forward on whatever is in self.x_in; returns the engine-owned output tensors."""
        if not self.use_graph:
            self._launch_all()
        else:
            if self._graph is None:
                self._graph = _capture_graph(self.device, self._launch_all)
            self._graph.replay()
        return self.p, self.v

    @_on_device
    def saved_activations(self):
        """layer name -> (mean, variance) fp32 NHWC of every conv output as this engine last computed it (post-ReLU
        where the layer has one; the interior of padded buffers).  What the layer-by-layer reference API would have
        returned at that point; used by the gradient tests to tell forward-decision errors from backward errors."""
        out = {}
        for r in self.records:
            if r["kind"] == "first":
                name, v = "conv_input", r["dst"]
                oh, ow, c = v.buf.shape[1], v.buf.shape[2], v.buf.shape[4]
            elif r["kind"] == "conv":
                name, v, c = r["name"], r["dst"], r["cout"]
                oh, ow = (2 * r["h"], 2 * r["w"]) if r["upconv"] else (r["h"] - r["k"] + 1, r["w"] - r["k"] + 1)
            else:
                continue
            t = v.buf[:, v.y0:v.y0 + oh, v.x0:v.x0 + ow, :, v.c0:v.c0 + c].float()
            out[name] = (t[..., 0, :] + t[..., 1, :], t[..., 2, :])
        return out

    @_on_device
    def run(self, x: Tensor, return_presoftmax: bool = False):
        if not self.matches(x):
            raise RuntimeError(f"engine built for input {self.shape}, got {tuple(x.shape)}")
        self.sync_weights()
        self.x_in.copy_(x, non_blocking=True)
        p, v = self.forward_resident()
        if return_presoftmax:
            if not self.keep_presoftmax:
                raise RuntimeError("engine was built with keep_presoftmax=False")
            return p.clone(), v.clone(), self.pre_m.clone(), self.pre_v.clone()
        return p.clone(), v.clone()


class GradientEngine(InferenceEngine):
    """FAST-mode forward + input-gradient chain: what create_adversarial_pattern (Brats.py:582-596) needs, on the
    tensor cores.  The forward is InferenceEngine's launch sequence (every activation buffer is kept, nothing is
    aliased); the backward walks its records in reverse:

      head      sn_head_bwd_packed            NLL -> softmax Jacobian -> conv_final -> ReLU gate
      conv      sn_conv_moments_bwd_data_tc   halo GEMM over the output gradient, flipped/transposed weights; the
                                              adjoints of pad / crop / concat / unpool are the forward's windows
      pool      sn_maxpool2_bwd_packed        arg-max routing recomputed from the saved pool input, summed with the
                                              gradient the decoder's concat already left in the skip tensor
      first     sn_first_conv_bwd_data_packed g_x (fp32 NHWC)

    Gradient tensors use the packed layout of the activation they belong to (g_mean hi/lo, g_variance).  The whole
    forward + loss + backward sequence is allocation-free and is captured in one CUDA graph."""

    def __init__(self, model, batch: int, in_h: int, in_w: int, in_c: int, device, graph: bool = True,
                 train: bool = False):
        self.want_rsum = train
        self.train = train
        self.exact_first_conv = True
        self.keep_last = True
        # FGSM / saliency chains end the forward in the fused head too (it also stores the last 32-channel tensor);
        # training needs that layer's rank-1 statistic (rsum_out), which the fused launch does not emit
        super().__init__(model, batch, in_h, in_w, in_c, device, graph=False, keep_presoftmax=False,
                         fuse_head=False if train else None)
        self.use_graph_bwd = graph
        self._graph_bwd: Optional[torch.cuda.CUDAGraph] = None
        self._bwd_steps: List[Callable[[], None]] = []
        self.bwd_step_names: List[str] = []
        self._clip = (-1e4, 1e3)
        self._loss_scale = 0.5
        self._build_backward()

    def _prepare_weights(self) -> None:
        super()._prepare_weights()
        m = self.model
        self.prepared_bwd = {}
        for name in m.conv_names:
            if name in ("conv_input", "conv_final"):
                continue
            w, _ = getattr(m, name).weights()
            self.prepared_bwd[name] = F.prepare_weights_bwd(w, upconv=name.endswith("conv2x2"))

    @_on_device
    def _build_backward(self) -> None:
        m = self.model
        B, H, W, Cin = self.shape
        dev = self.device
        C = m.n_labels
        oh, ow = self.out_hw
        self.y_in = torch.zeros((B, oh * ow, C), device=dev, dtype=torch.float32)
        self.g_x = torch.empty(self.shape, device=dev, dtype=torch.float32)
        self.nll_acc = torch.zeros(2, device=dev, dtype=torch.float64)
        self.nll_loss = torch.zeros(1, device=dev, dtype=torch.float32)
        gbuf = {}                 # activation buffer -> gradient buffer of the same shape
        gated = {}                # activation buffer -> its producer applied the ReLU gate (or pooled such a tensor)

        def g_of(view: PackedView) -> PackedView:
            key = view.buf.data_ptr()
            if key not in gbuf:
                gbuf[key] = torch.empty_like(view.buf)
            return PackedView(gbuf[key], view.y0, view.x0, view.c0)

        for r in self.records:
            if r["kind"] == "first":
                gated[r["dst"].buf.data_ptr()] = True
            elif r["kind"] == "conv":
                gated[r["dst"].buf.data_ptr()] = bool(r["relu"])
            elif r["kind"] == "pool":
                gated[r["dst"].buf.data_ptr()] = gated[r["src"].buf.data_ptr()]
        steps, names = self._bwd_steps, self.bwd_step_names
        wf, wsf = m.conv_final.weights()
        w_in, ws_in = m.conv_input.weights()
        self.grads = {}           # training: layer name -> (g_w_mu, g_w_sigma), overwritten by every step
        logit_grads = None
        if self.train:
            # one flat fp32 buffer in parameter order (w_mu, w_sigma per layer): the data-parallel all-reduce runs on
            # it directly, the per-layer gradients are views
            total = sum(w_.numel() + ws_.numel() for w_, ws_ in (getattr(m, n_).weights() for n_ in m.conv_names))
            self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
            off = 0
            for name in m.conv_names:
                w_, ws_ = getattr(m, name).weights()
                gw_ = self.flat_grad[off:off + w_.numel()].view_as(w_)
                off += w_.numel()
                gws_ = self.flat_grad[off:off + ws_.numel()].view_as(ws_)
                off += ws_.numel()
                self.grads[name] = (gw_, gws_)
            self.reg_acc = torch.zeros(1, device=dev, dtype=torch.float64)
            self._kl_scale = 0.0
            n_max = max(w_.numel() for w_, _ in (getattr(m, n).weights() for n in m.conv_names))
            self._wg_work = torch.empty(2 * n_max + 1024, device=dev, dtype=torch.float32)
            rows = B * oh * ow
            logit_grads = (torch.empty((rows, C), device=dev), torch.empty((rows, C), device=dev),
                           torch.empty(rows, device=dev))
        for r in reversed(self.records):
            kind = r["kind"]
            if kind == "head":
                src = r["src"]
                self._head_record = r
                if not gated[src.buf.data_ptr()]:
                    raise RuntimeError("conv_final must read a post-ReLU tensor")
                names.append("head_bwd")
                steps.append(lambda src=src, r=r: F.head_bwd_packed(src, B, r["h"], r["w"], r["c"], wf, wsf, self.y_in,
                                                                    self._clip, self.nll_acc, self._loss_scale,
                                                                    g_of(src), logit_grads))
                g_of(src)
                if self.train:
                    # conv_final (32 -> n_labels, k = 1) and conv_input (Cin <= 8) are too thin for the tensor
                    # cores: their weight gradients run in the FP32-mode kernel on unpacked operands
                    names.append("conv_final_wgrad")
                    if r["c"] == 32:
                        gwf, gwsf = self.grads["conv_final"]
                        steps.append(lambda src=src, r=r: F.final_conv_bwd_weight_packed(
                            src, B, r["h"], r["w"], r["c"], wf, wsf, logit_grads, self._wg_work, gwf, gwsf))
                    else:                               # other widths: the general FP32-mode kernel on unpacked operands
                        mu32 = torch.empty((B, r["h"], r["w"], r["c"]), device=dev)
                        var32 = torch.empty_like(mu32)
                        steps.append(lambda src=src, r=r, mu32=mu32, var32=var32: self._wgrad_f32(
                            "conv_final", src.buf, mu32, var32, logit_grads[0], logit_grads[1], logit_grads[2], wf,
                            wsf, r["h"], r["w"], r["c"], C, 1))
            elif kind == "conv":
                name = r["name"]
                wt = self.prepared_bwd[name]
                s = self.prepared[name][1]
                src0, src1 = r["src0"], r["src1"]
                args = dict(g_out=g_of(r["dst"]), batch=B, in_h=r["h"], in_w=r["w"], ksize=r["k"], cout=r["cout"],
                            wt_packed=wt, s=s, in0=src0, g_in0=g_of(src0), c0=r["c0"],
                            gate0=gated[src0.buf.data_ptr()], upconv=r["upconv"])
                if src1 is not None:
                    args.update(in1=src1, g_in1=g_of(src1), c1=r["c1"], gate1=gated[src1.buf.data_ptr()])
                if self.train:
                    w_, ws_ = getattr(m, name).weights()
                    gw, gws = self.grads[name]
                    wargs = dict(g_out=g_of(r["dst"]), batch=B, in_h=r["h"], in_w=r["w"], ksize=r["k"],
                                 cout=r["cout"], in0=src0, c0=r["c0"], rsum=r["rsum"], w_mu=w_, w_sigma=ws_,
                                 workspace=self._wg_work, g_w_mu=gw, g_w_sigma=gws, upconv=r["upconv"])
                    if src1 is not None:
                        wargs.update(in1=src1, c1=r["c1"])
                    names.append(name + "_wgrad")
                    steps.append(lambda a=wargs: F.conv_moments_bwd_weight_tc(**a))
                names.append(name + "_dgrad")
                steps.append(lambda a=args: F.conv_moments_bwd_data_tc(**a))
            elif kind == "pool":
                src, dst = r["src"], r["dst"]
                keep = self.skip_window.get(src.buf.data_ptr(), (0, 0, 0, 0))
                names.append("pool_bwd")
                steps.append(lambda src=src, dst=dst, r=r, keep=keep: F.maxpool2_bwd_packed(
                    src, B, r["h"], r["w"], r["c"], g_of(dst), g_of(src), keep))
                g_of(dst), g_of(src)
            elif kind == "first":
                if self.train:
                    a0 = r["dst"].buf
                    _, ah, aw, _, ac = a0.shape
                    k0 = w_in.shape[0]
                    names.append("conv_input_wgrad")
                    if k0 == 3 and ac == 32 and Cin in (1, 4):
                        gw0, gws0 = self.grads["conv_input"]
                        steps.append(lambda r=r: F.first_conv_bwd_weight_packed(
                            self.x_in, w_in, ws_in, g_of(r["dst"]), self._wg_work, gw0, gws0))
                    else:
                        gm32 = torch.empty((B, ah, aw, ac), device=dev)
                        gv32 = torch.empty_like(gm32)
                        rs0 = torch.empty((B, ah, aw), device=dev)
                        steps.append(lambda r=r, gm32=gm32, gv32=gv32, rs0=rs0: self._first_wgrad(
                            g_of(r["dst"]).buf, gm32, gv32, rs0, w_in, ws_in, k0))
                else:                                    # the input gradient is what FGSM needs; training stops here
                    names.append("conv_input_dgrad")
                    steps.append(lambda r=r: F.first_conv_bwd_data_packed(self.x_in, w_in, ws_in, g_of(r["dst"]),
                                                                          self.g_x))
                g_of(r["dst"])
        self.grad_buffers = gbuf
        self.n_launches_bwd = len(steps) + 2          # + the two NLL forward kernels

    @_on_device
    def refresh_weights(self) -> None:
        """After an optimiser step: re-derive forward AND data-gradient operands in place."""
        super().refresh_weights()
        m = self.model
        for name in m.conv_names:
            if name in ("conv_input", "conv_final"):
                continue
            w, _ = getattr(m, name).weights()
            F.prepare_weights_bwd(w, upconv=name.endswith("conv2x2"), out=self.prepared_bwd[name])

    def _regulariser(self) -> None:
        """add_n(model.losses) (Brats.py:575) and its gradient, kl_scale = kl_factor * 0.5 (Brats.py:576): the value
        accumulates into reg_acc, the gradient into the weight-gradient buffers (after the data terms were written)."""
        import ctypes as C
        from ._lib import check, load, ptr, stream_ptr
        lib = load()
        self.reg_acc.zero_()
        for name in self.model.conv_names:
            w, ws = getattr(self.model, name).weights()
            gw, gws = self.grads[name]
            check(lib.sn_kl_regularizer_fwd(ptr(w), C.c_size_t(w.numel()), ptr(ws), ws.numel(), w.shape[0],
                                            ptr(self.reg_acc), stream_ptr()), "kl_regularizer_fwd")
            check(lib.sn_kl_regularizer_bwd(ptr(w), C.c_size_t(w.numel()), ptr(ws), ws.numel(), w.shape[0],
                                            C.c_float(self._kl_scale), ptr(gw), ptr(gws), stream_ptr()),
                  "kl_regularizer_bwd")

    # FP32-mode weight gradient on unpacked operands (thin layers only)
    def _wgrad_f32(self, name, packed_in, mu32, var32, g_mu, g_var, rsum, w, ws, h, wd, cin, cout, k) -> None:
        from . import ops
        F.unpack_moments(packed_in, out=(mu32, var32))
        gw, gws = self.grads[name]
        ops.conv_bwd_weight_raw(self.shape[0], h, wd, cin, cout, k, mu32, var32, g_mu, g_var, rsum, w, ws, gw, gws)

    def _first_wgrad(self, g_packed, gm32, gv32, rs0, w, ws, k) -> None:
        from . import ops
        F.unpack_moments(g_packed, out=(gm32, gv32))
        F.first_conv_rsum(self.x_in, k, rs0)
        B, H, W, cin = self.shape
        gw, gws = self.grads["conv_input"]
        ops.conv_bwd_weight_raw(B, H, W, cin, w.shape[-1], k, self.x_in, None, gm32, gv32, rs0, w, ws, gw, gws)

    @_on_device
    def loss_and_weight_gradients(self, x: Tensor, y_onehot: Tensor, clip: Tuple[float, float] = (1e-12, 1e3),
                                  kl_factor: float = 0.0):
        """train_on_batch's loss and gradients (Brats.py:572-578): loss = NLL(clip(var)) + kl_factor * 0.5 *
        add_n(model.losses).  Returns (loss [1] device tensor, {layer: (d/dw_mu, d/dw_sigma)}); the gradients are views
        of self.flat_grad, engine-owned and overwritten by the next call.  One CUDA-graph replay: operand refresh,
        forward, NLL, data and weight gradients, regularisers."""
        if not self.train:
            raise RuntimeError("engine was built with train=False")
        if not self.matches(x):
            raise RuntimeError(f"engine built for input {self.shape}, got {tuple(x.shape)}")
        want = (1.0, (float(clip[0]), float(clip[1])), float(kl_factor) * 0.5)
        if want != (self._loss_scale, self._clip, self._kl_scale):
            self._loss_scale, self._clip, self._kl_scale = want
            self._graph_bwd = None                       # scalars are baked into the captured launches
        self.x_in.copy_(x, non_blocking=True)
        self.y_in.copy_(y_onehot.reshape(self.y_in.shape), non_blocking=True)
        nll, _ = self.loss_and_input_gradient_resident()
        loss = nll + (self._kl_scale * self.reg_acc).to(torch.float32)
        return loss, self.grads

    def _launch_fwd_bwd(self) -> None:
        if self.train:
            self.refresh_weights()       # inside the captured sequence: the operands follow every optimiser step
        self._launch_all()
        F.nll_gaussian_fwd(self.y_in, self.p, self.v, self._clip, self.nll_acc, self.nll_loss)
        for s in self._bwd_steps:
            s()
        if self.train:
            self._regulariser()

    @_on_device
    def loss_and_input_gradient_resident(self) -> Tuple[Tensor, Tensor]:
        """Forward, loss_scale * NLL and d(loss)/dx on whatever is in self.x_in / self.y_in; returns engine-owned
        (loss [1] = the un-scaled NLL, g_x)."""
        if not self.use_graph_bwd:
            self._launch_fwd_bwd()
        else:
            if self._graph_bwd is None:
                self._graph_bwd = _capture_graph(self.device, self._launch_fwd_bwd)
            self._graph_bwd.replay()
        return self.nll_loss, self.g_x

    @_on_device
    def forward_then_input_gradient(self, x: Tensor, upstream) -> Tuple[Tensor, Tensor, Tensor]:
        """Forward, then d <g_p, p> + <g_v, v> / dx for upstream gradients chosen AFTER looking at the outputs:
        `upstream(p, v) -> (g_p, g_v or None)`, both [B, HW, C] device tensors.  This is create_saliency_map's shape
        (Brats.py:598-609): the mask depends on the prediction.  Eager launches (the callback sits in the middle).
        Returns (g_x, p, v), engine-owned."""
        if self.train:
            raise RuntimeError("build the engine with train=False for input gradients")
        if not self.matches(x):
            raise RuntimeError(f"engine built for input {self.shape}, got {tuple(x.shape)}")
        self.sync_weights()
        self.x_in.copy_(x, non_blocking=True)
        self._launch_all()
        g_p, g_v = upstream(self.p, self.v)
        g_p = g_p.to(torch.float32).contiguous()
        g_v = g_v.to(torch.float32).contiguous() if g_v is not None else None
        head = self._head_record
        src = head["src"]
        gkey = src.buf.data_ptr()
        B = self.shape[0]
        wf, wsf = self.model.conv_final.weights()
        F.head_bwd_upstream_packed(src, B, head["h"], head["w"], head["c"], wf, wsf, g_p, g_v,
                                   PackedView(self.grad_buffers[gkey], src.y0, src.x0, src.c0))
        for s_ in self._bwd_steps[1:]:
            s_()
        return self.g_x, self.p, self.v

    @_on_device
    def input_gradient(self, x: Tensor, y_onehot: Tensor, loss_scale: float = 0.5,
                       clip: Tuple[float, float] = (-1e4, 1e3)) -> Tuple[Tensor, Tensor]:
        """(loss, d loss / d x) with loss = loss_scale * nll_gaussian(y, p, clip(var)): the defaults are
        create_adversarial_pattern's loss (Brats.py:587-590)."""
        if not self.matches(x):
            raise RuntimeError(f"engine built for input {self.shape}, got {tuple(x.shape)}")
        if (loss_scale, tuple(clip)) != (self._loss_scale, self._clip):
            self._loss_scale, self._clip = float(loss_scale), (float(clip[0]), float(clip[1]))
            self._graph_bwd = None                       # scalars are baked into the captured launches
        self.sync_weights()
        self.x_in.copy_(x, non_blocking=True)
        self.y_in.copy_(y_onehot.reshape(self.y_in.shape), non_blocking=True)
        loss, g = self.loss_and_input_gradient_resident()
        return loss_scale * loss.clone(), g.clone()


class StreamingPipeline:
    """Host-facing inference: pinned host batch -> H2D -> forward -> D2H of both maps, software-pipelined.

    `depth` engines (each with its own buffers, CUDA graph and stream) are used round-robin, so the H2D copy of
    batch i+1 and the D2H copies of batch i-1 overlap the kernels of batch i (B200 has separate copy engines per
    direction).  This is the e2e path bench.py times; testing()'s loop (Brats.py:1197-1298) maps onto
    submit()/result()."""

    def __init__(self, model, batch: int, in_h: int, in_w: int, in_c: int, device, depth: int = 2):
        self.device = torch.device(device)
        self.engines = [InferenceEngine(model, batch, in_h, in_w, in_c, device, graph=True, keep_presoftmax=False)
                        for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        # One forward at a time: slot i's graph waits for the graph submitted before it.  Every conv kernel is persistent
        # and owns all 148 SMs, so two graphs on independent streams only interleave kernel by kernel -- which breaks the
        # programmatic-dependent-launch chaining inside a graph and halves the L2 reuse between a layer and the next --
        # while the copies (separate DMA engines) are what the extra streams are for.  SN_PIPE_SERIAL=0: old behaviour.
        self.serial = os.environ.get("SN_PIPE_SERIAL", "1") != "0"
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self._last: Optional[int] = None
        e0 = self.engines[0]
        self.pv_host = [torch.empty(e0.pv.shape, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.p_host = [t[0] for t in self.pv_host]
        self.v_host = [t[1] for t in self.pv_host]
        self._next = 0
        self._busy = [False] * depth
        for eng, st in zip(self.engines, self.streams):          # capture each graph on its own stream
            st.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(st):
                eng.forward_resident()
        torch.cuda.synchronize(self.device)

    @property
    def n_launches(self) -> int:
        return self.engines[0].n_launches

    def submit(self, x_host: Tensor) -> int:
        """Enqueue one batch (a pinned host tensor); returns the slot to pass to result()."""
        slot = self._next
        self._next = (slot + 1) % len(self.engines)
        if self._busy[slot]:
            self.done[slot].synchronize()        # the slot's previous outputs must have landed before reuse
        eng, st = self.engines[slot], self.streams[slot]
        with torch.cuda.stream(st):
            eng.sync_weights()                   # operands follow in-place weight updates (every engine has its own)
            eng.x_in.copy_(x_host, non_blocking=True)
            if self.serial and self._last is not None and self._last != slot:
                st.wait_event(self.computed[self._last])
            p, v = eng.forward_resident()
            self.computed[slot].record(st)
            self.pv_host[slot].copy_(eng.pv, non_blocking=True)      # both maps, one DMA
            self.done[slot].record(st)
        self._busy[slot] = True
        self._last = slot
        return slot

    def result(self, slot: int) -> Tuple[Tensor, Tensor]:
        self.done[slot].synchronize()
        return self.p_host[slot], self.v_host[slot]

    def join(self) -> None:
        """Make the current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)
