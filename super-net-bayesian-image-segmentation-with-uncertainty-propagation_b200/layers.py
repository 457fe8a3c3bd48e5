"""Host-side mirror of the reference's Keras layers (same names, constructor arguments, call order and
tuple returns) as torch.nn.Modules whose arithmetic runs in libsupernet_b200.so.

Reference (file:line into /root/reference): myConv_input Brats.py:34-76, myConv_intermediate :80-137,
myupsampling :140-148, mypadding :151-163, mymaxpooling :166-174, myReLU :227-238, myConc :241-261,
mysoftmax :264-283, nll_gaussian :293-311, sigma_regularizer :314-320, Density_prop_with_pad_UNET
:323-457 (BraTS, 5 levels) and Hippocampus.py:335-421 (3 levels).

Conventions kept from the reference: activations NHWC float32; "sigma" is the VARIANCE; weights `w_mu`
HWIO [k,k,Cin,Cout] and raw (pre-softplus) `w_sigma` [Cout]; no bias.  Deliberate deviation: mysoftmax keeps
[B,HW,C] for B == 1 (the reference's bare tf.squeeze would drop the batch axis).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

import contextlib

from . import fastlayers as FL
from . import ops

Tensor = torch.Tensor

# ---- execution mode of the per-layer API -------------------------------------------------------------------
# 'fp32': CUDA-core fp32 kernels with autograd (1e-5 parity, training through torch autograd).
# 'fast': tcgen05 tensor-core kernels on packed moments (inference; fastlayers.py).  The mode is chosen where the
# pair (mean, sigma) is BORN -- myConv_input -- and every later layer follows the kind of pair it is handed.
_DEFAULT_MODE = "fp32"


def set_default_mode(mode: str) -> None:
    """Mode of myConv_input layers constructed without an explicit `mode=`."""
    global _DEFAULT_MODE
    if mode not in ("fp32", "fast"):
        raise ValueError("mode must be 'fp32' or 'fast'")
    _DEFAULT_MODE = mode


def get_default_mode() -> str:
    return _DEFAULT_MODE


@contextlib.contextmanager
def fast_mode(enabled: bool = True):
    """with fast_mode(): ...  -- layer-by-layer code (Brats.py:379-455 style) runs on the tensor cores."""
    prev = _DEFAULT_MODE
    set_default_mode("fast" if enabled else "fp32")
    try:
        yield
    finally:
        set_default_mode(prev)


def _truncated_normal_(t: Tensor, mean: float, std: float, gen: Optional[torch.Generator] = None) -> Tensor:
    """tf.keras.initializers.TruncatedNormal: values beyond mean +- 2 std are redrawn (Brats.py:52)."""
    with torch.no_grad():
        t.normal_(mean, std, generator=gen)
        while True:
            bad = (t - mean).abs() > 2 * std
            n = int(bad.sum())
            if n == 0:
                return t
            t[bad] = torch.empty(n, dtype=t.dtype, device=t.device).normal_(mean, std, generator=gen)


class _MomentConv(nn.Module):
    """Shared parameter handling of the two moment convolutions (weights are created at the first call
    when `in_channels` is not given, like Keras' build())."""

    def __init__(self, kernel_num, kernel_size, kernel_stride, padding, mean_mu, mean_sigma, sigma_min, sigma_max,
                 in_channels, mu_name, sigma_name):
        super().__init__()
        if kernel_stride != 1 or str(padding).upper() != "VALID":
            raise ValueError("only stride 1 / VALID is supported (the only mode the reference uses, Brats.py:37,89)")
        self.kernel_num = int(kernel_num)
        self.kernel_size = int(kernel_size)
        self.kernel_stride = 1
        self.padding = "VALID"
        self.mean_mu, self.mean_sigma = float(mean_mu), float(mean_sigma)
        self.sigma_min, self.sigma_max = float(sigma_min), float(sigma_max)
        self._mu_name, self._sigma_name = mu_name, sigma_name
        self.fuse_relu = False        # set by the model graph to fold the following myReLU into the conv epilogue
        if in_channels is not None:
            self._build(int(in_channels), None)

    def _build(self, cin: int, device) -> None:
        k, n = self.kernel_size, self.kernel_num
        w = _truncated_normal_(torch.empty(k, k, cin, n, dtype=torch.float32), self.mean_mu, self.mean_sigma)
        s = torch.empty(n, dtype=torch.float32).uniform_(self.sigma_min, self.sigma_max)
        self.register_parameter(self._mu_name, nn.Parameter(w.to(device) if device is not None else w))
        self.register_parameter(self._sigma_name, nn.Parameter(s.to(device) if device is not None else s))

    @property
    def built(self) -> bool:
        return self._mu_name in self._parameters

    def weights(self) -> Tuple[Tensor, Tensor]:
        return self._parameters[self._mu_name], self._parameters[self._sigma_name]

    def set_weights(self, w_mu: Tensor, w_sigma: Tensor) -> None:
        if not self.built:
            self._build(w_mu.shape[2], w_mu.device)
        wm, ws = self.weights()
        with torch.no_grad():
            wm.copy_(w_mu.to(wm.dtype))
            ws.copy_(w_sigma.to(ws.dtype))


class myConv_input(_MomentConv):
    """First convolution: deterministic input, random weights (Brats.py:34-76).
    mu = x (*) w_mu;  sigma[b,i,j,n] = softplus(w_sigma[n]) * sum_{patch} x^2."""

    def __init__(self, kernel_num=128, kernel_size=3, kernel_stride=1, padding="VALID", mean_mu=0, mean_sigma=0.1,
                 sigma_min=-12, sigma_max=-4.6, in_channels: Optional[int] = None, mode: Optional[str] = None):
        super().__init__(kernel_num, kernel_size, kernel_stride, padding, mean_mu, mean_sigma, sigma_min, sigma_max,
                         in_channels, "w_mu1", "w_sigma1")
        if mode not in (None, "fp32", "fast"):
            raise ValueError("mode must be None, 'fp32' or 'fast'")
        self.mode = mode            # None: layers.get_default_mode() at call time

    def forward(self, inputs: Tensor):
        if not self.built:
            self._build(inputs.shape[-1], inputs.device)
        if (self.mode or _DEFAULT_MODE) == "fast":
            m, s = FL.conv_input(self, inputs)
            return FL.relu(m, s) if self.fuse_relu else (m, s)
        return ops.conv_moments(inputs, None, self.w_mu1, self.w_sigma1, self.fuse_relu)


class myConv_intermediate(_MomentConv):
    """Intermediate convolution: random input and weights (Brats.py:80-137).
    mu = mu_in (*) w_mu;  sigma = sigma_in (*) w_mu^2 + softplus(w_sigma)[n] * sum_{patch}(mu_in^2 + sigma_in)."""

    def __init__(self, kernel_num=64, kernel_size=3, kernel_stride=1, padding="VALID", mean_mu=0, mean_sigma=0.1,
                 sigma_min=-12, sigma_max=-4.6, in_channels: Optional[int] = None):
        super().__init__(kernel_num, kernel_size, kernel_stride, padding, mean_mu, mean_sigma, sigma_min, sigma_max,
                         in_channels, "w_mu", "w_sigma")

    def forward(self, inputs: Tensor, sigma_input: Tensor):
        if FL.is_handle(inputs):                     # FAST mode: the pair is a packed-moments handle
            return FL.conv_intermediate(self, inputs, sigma_input)
        if not self.built:
            self._build(inputs.shape[-1], inputs.device)
        return ops.conv_moments(inputs, sigma_input, self.w_mu, self.w_sigma, self.fuse_relu)


class myupsampling(nn.Module):
    """Zero-stuffing up-sampling of both moments to 2H+1 (Brats.py:140-148, unpool :178-203)."""

    def forward(self, mu_in: Tensor, sigma_in: Tensor):
        if FL.is_handle(mu_in):
            return FL.upsampling(mu_in, sigma_in)
        return ops.unpool(mu_in), ops.unpool(sigma_in)


class mypadding(nn.Module):
    """Zero-pad the mean, constant-pad the variance with `sigma_fill` (Brats.py:151-163)."""

    def __init__(self, pad_size: Sequence[int] = (2, 2), sigma_fill=0, mode="CONSTANT"):
        super().__init__()
        if str(mode).upper() != "CONSTANT":
            raise ValueError("only CONSTANT padding is supported (Brats.py:154)")
        self.pad_size = [int(pad_size[0]), int(pad_size[1])]
        self.sigma_fill = float(sigma_fill)
        self.mode = "CONSTANT"

    def forward(self, mu_in: Tensor, sigma_in: Tensor):
        if FL.is_handle(mu_in):
            return FL.padding(self, mu_in, sigma_in)
        a, b = self.pad_size
        return ops.pad_hw(mu_in, a, b, 0.0), ops.pad_hw(sigma_in, a, b, self.sigma_fill)


class mymaxpooling(nn.Module):
    """2x2/2 max-pool of the mean; the variance is taken at the arg-max (Brats.py:166-174, 206-216)."""

    def forward(self, mu_in: Tensor, sigma_in: Tensor):
        if FL.is_handle(mu_in):
            return FL.maxpooling(mu_in, sigma_in)
        return ops.maxpool2_moments(mu_in, sigma_in)


class myReLU(nn.Module):
    """mu -> relu(mu), sigma -> sigma * 1[mu > 0] (Brats.py:227-238)."""

    def forward(self, mu_in: Tensor, Sigma_in: Tensor):
        if FL.is_handle(mu_in):
            return FL.relu(mu_in, Sigma_in)
        return ops.relu_moments(mu_in, Sigma_in)


class myConc(nn.Module):
    """Centre-crop the encoder moments to the decoder size and concat [decoder, encoder] (Brats.py:241-261)."""

    def forward(self, muD: Tensor, SigmaD: Tensor, muE: Tensor, SigmaE: Tensor):
        if FL.is_handle(muD):
            return FL.conc(muD, SigmaD, muE, SigmaE)
        return ops.crop_concat(muD, muE), ops.crop_concat(SigmaD, SigmaE)


class mysoftmax(nn.Module):
    """Softmax over classes with Jacobian-propagated variance, flattened to [B, H*W, C] (Brats.py:264-283)."""

    def forward(self, mu_in: Tensor, sigma_in: Tensor):
        if FL.is_handle(mu_in):
            return FL.softmax(mu_in, sigma_in)
        B, Cc = mu_in.shape[0], mu_in.shape[-1]
        p, v = ops.softmax_moments(mu_in, sigma_in)
        return p.reshape(B, -1, Cc), v.reshape(B, -1, Cc)


def nll_gaussian(y_test: Tensor, y_pred_mean: Tensor, y_pred_sd: Tensor, clip: Tuple[float, float] = None) -> Tensor:
    """nll_gaussian (Brats.py:293-311).  The reference clips the variance before the call (Brats.py:573-574
    [1e-12, 1e3]; :588-589 [-1e4, 1e3]); pass `clip=(lo, hi)` to fuse that clip (and its gradient mask)."""
    lo, hi = clip if clip is not None else (-3.0e38, 3.0e38)
    return ops.nll_gaussian_clipped(y_test, y_pred_mean, y_pred_sd, lo, hi)


class sigma_regularizer:
    """sigma_regularizer (Brats.py:314-320): -strength * mean(1 + log softplus(x) - softplus(x)).
    Host-side scalar utility for API parity; the model's summed regulariser runs in ops.kl_regularizer."""

    def __init__(self, strength):
        self.strength = strength

    def __call__(self, x: Tensor) -> Tensor:
        f_s = torch.nn.functional.softplus(x)
        return -self.strength * torch.mean(1.0 + torch.log(f_s) - f_s, dim=-1)


# -------------------------------------------------------------------------------------------------------
# model graph
# -------------------------------------------------------------------------------------------------------
_HI = dict(sigma_min=-4.6, sigma_max=-2.2)


def conv_layer_specs(variant: str, n: int, n_labels: int) -> List[Tuple[str, int, int, dict]]:
    """(name, kernel_num, kernel_size, sigma-init kwargs) in __init__ order (Brats.py:331-367,
    Hippocampus.py:343-363)."""
    if variant == "brats":
        return [("conv_input", n, 3, {}), ("conv1", n, 3, {}), ("conv2", 2 * n, 3, {}), ("conv3", 2 * n, 3, {}),
                ("conv4", 4 * n, 3, {}), ("conv5", 4 * n, 3, {}), ("conv6", 8 * n, 3, {}), ("conv7", 8 * n, 3, {}),
                ("conv8", 16 * n, 3, {}), ("conv9", 16 * n, 3, {}),
                ("up1_conv2x2", 8 * n, 2, _HI), ("up1_conv1", 8 * n, 3, {}), ("up1_conv2", 8 * n, 3, {}),
                ("up2_conv2x2", 4 * n, 2, _HI), ("up2_conv1", 4 * n, 3, {}), ("up2_conv2", 4 * n, 3, {}),
                ("up3_conv2x2", 2 * n, 2, {}), ("up3_conv1", 2 * n, 3, {}), ("up3_conv2", 2 * n, 3, {}),
                ("up4_conv2x2", n, 2, {}), ("up4_conv1", n, 3, {}), ("up4_conv2", n, 3, {}),
                ("conv_final", n_labels, 1, _HI)]
    if variant == "hippocampus":
        return [("conv_input", n, 3, {}), ("conv1", n, 3, {}), ("conv2", 2 * n, 3, {}), ("conv3", 2 * n, 3, {}),
                ("conv4", 4 * n, 3, {}), ("conv5", 4 * n, 3, {}),
                ("up1_conv2x2", 2 * n, 2, _HI), ("up1_conv1", 2 * n, 3, {}), ("up1_conv2", 2 * n, 3, {}),
                ("up2_conv2x2", n, 2, _HI), ("up2_conv1", n, 3, {}), ("up2_conv2", n, 3, {}),
                ("conv_final", n_labels, 1, _HI)]
    raise ValueError(f"unknown variant {variant!r} (expected 'brats' or 'hippocampus')")


class Density_prop_with_pad_UNET(nn.Module):
    """The moment-propagation U-Net (Brats.py:323-457; `variant='hippocampus'` gives Hippocampus.py:335-421).

    call(inputs[B,H,W,Cin], training=True) -> (outputs, Sigma), both [B, H_out*W_out, n_labels]; `outputs` are
    softmax probabilities (the reference calls them "logits").

    `mode='fp32'` runs every layer through the FP32-mode kernels with autograd (training, FGSM);
    `mode='fast'` (inference only) runs the fused tcgen05 pipeline of engine.InferenceEngine.
    """

    def __init__(self, n_kernels, n_labels, name=None, variant: str = "brats", in_channels: Optional[int] = None,
                 mode: str = "fp32"):
        super().__init__()
        self.n_kernels, self.n_labels, self.variant = int(n_kernels), int(n_labels), variant
        self.model_name = name
        if mode not in ("fp32", "fast"):
            raise ValueError("mode must be 'fp32' or 'fast'")
        self.mode = mode
        self.levels = 4 if variant == "brats" else 2
        self.sigma_fill = 0.1 if variant == "brats" else 0.02     # Brats.py:370-372 / Hippocampus.py:366-368
        specs = conv_layer_specs(variant, self.n_kernels, self.n_labels)
        self.conv_names = [s[0] for s in specs]
        cin = in_channels
        for name_, kn, ks, kw in specs:
            if name_ == "conv_input":
                layer = myConv_input(kernel_num=kn, kernel_size=ks, in_channels=cin, **kw)
            else:
                layer = myConv_intermediate(kernel_num=kn, kernel_size=ks, **kw)
            setattr(self, name_, layer)
        self.maxp = mymaxpooling()
        self.myups = myupsampling()
        self.myrelu = myReLU()
        self.myconc = myConc()
        self.mysoft = mysoftmax()
        self.mypad = mypadding(pad_size=[2, 2], sigma_fill=self.sigma_fill)
        self.mypad_up6 = mypadding(pad_size=[3, 3], sigma_fill=self.sigma_fill)
        self.mypad1 = mypadding(pad_size=[1, 0], sigma_fill=self.sigma_fill)
        # fold every myReLU that directly follows a conv into that conv's epilogue
        for name_ in self.conv_names:
            if not name_.endswith("conv2x2") and name_ != "conv_final":
                getattr(self, name_).fuse_relu = True
        self._engine = None
        self._grad_engine = None

    # -- Keras-like accessors ---------------------------------------------------------------------------
    def convs(self) -> List[_MomentConv]:
        return [getattr(self, n) for n in self.conv_names]

    @property
    def trainable_weights(self) -> List[Tensor]:
        return [p for p in self.parameters() if p.requires_grad]

    def regularization(self) -> Tensor:
        """add_n(model.losses) (Brats.py:575): sum over convs of l2(1.)(w_mu) + sigma_regularizer(k*k)(w_sigma)."""
        return ops.kl_regularizer([c.weights() for c in self.convs()])

    @property
    def losses(self) -> List[Tensor]:
        return [self.regularization()]

    def load_weight_dict(self, weights: Dict[str, Tuple[Tensor, Tensor]], device=None) -> "Density_prop_with_pad_UNET":
        """weights: layer name -> (w_mu HWIO, raw w_sigma), e.g. oracle.make_weights(...)."""
        for n in self.conv_names:
            w, s = weights[n]
            if device is not None:
                w, s = w.to(device), s.to(device)
            getattr(self, n).set_weights(w, s)
        self._engine = None
        self._grad_engine = None
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return self

    def build_with_input(self, in_channels: int, device) -> None:
        """Create every weight without a dummy forward (the reference runs one, Brats.py:617-620)."""
        cin = in_channels
        chans = {}
        n = self.n_kernels
        for name_ in self.conv_names:
            layer = getattr(self, name_)
            if name_ == "conv_input":
                c = cin
            elif name_.endswith("_conv1") and name_.startswith("up"):
                c = 2 * layer.kernel_num
            elif name_.endswith("conv2x2"):
                c = 2 * layer.kernel_num
            elif name_ == "conv_final":
                c = n
            else:
                c = chans["prev"]
            if not layer.built:
                layer._build(c, device)
            chans["prev"] = layer.kernel_num

    # -- forward ------------------------------------------------------------------------------------------
    def forward(self, inputs: Tensor, training: bool = True, return_presoftmax: bool = False):
        if self.mode == "fast":        # inference only: the outputs carry no autograd history
            return self._forward_fast(inputs, return_presoftmax)
        return self._forward_fp32(inputs, return_presoftmax)

    def _cr(self, name: str, m: Tensor, s: Optional[Tensor]):
        """conv followed by myrelu (fused into the conv epilogue when fuse_relu is set)."""
        layer = getattr(self, name)
        m, s = layer(m) if s is None else layer(m, s)
        if not layer.fuse_relu:
            m, s = self.myrelu(m, s)
        return m, s

    def _forward_fp32(self, x: Tensor, return_presoftmax: bool = False):
        L = self.levels
        m, s = self._cr("conv_input", x, None)                       # Brats.py:379-380
        m, s = self._cr("conv1", m, s)                               # :381-382
        skips = [(m, s)]
        ci = 2
        for lvl in range(1, L + 1):
            m, s = self.maxp(m, s)                                   # :383,390,397,404
            if self.variant == "brats" and lvl == L:
                m, s = self.mypad1(m, s)                             # :407
            for _ in range(2):
                m, s = self._cr(f"conv{ci}", m, s)
                ci += 1
            if lvl < L:
                skips.append((m, s))
        for d in range(1, L + 1):
            me, se = skips[L - d]
            m, s = self.myups(m, s)                                  # :414
            m, s = getattr(self, f"up{d}_conv2x2")(m, s)             # :415 (no ReLU)
            m, s = self.mypad_up6(m, s)                              # :416
            m, s = self.myconc(m, s, me, se)                         # :417
            m, s = self._cr(f"up{d}_conv1", m, s)                    # :418-419
            m, s = self.mypad(m, s)                                  # :420
            m, s = self._cr(f"up{d}_conv2", m, s)                    # :421-422
        mf, sf = self.conv_final(m, s)                               # :454
        if FL.is_handle(mf) and return_presoftmax:                   # FAST handles: one fused launch gives all four
            p, v, pre = mf.run(True)
            return p, v, pre[0], pre[1]
        outputs, Sigma = self.mysoft(mf, sf)                         # :455
        if return_presoftmax:
            return outputs, Sigma, mf, sf
        return outputs, Sigma

    def forward_layerwise_fast(self, x: Tensor, return_presoftmax: bool = False):
        """The same graph as _forward_fp32 -- one reference-style layer call after the other -- with the pair travelling
        as a packed-moments handle (fastlayers.py): every conv is a tcgen05 launch, eager, no engine, no CUDA graph."""
        prev = self.conv_input.mode
        self.conv_input.mode = "fast"
        try:
            with torch.no_grad():
                return self._forward_fp32(x, return_presoftmax)
        finally:
            self.conv_input.mode = prev

    def _forward_fast(self, x: Tensor, return_presoftmax: bool = False):
        from .engine import InferenceEngine
        if self._engine is None or not self._engine.matches(x):
            self._engine = InferenceEngine(self, x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
        return self._engine.run(x, return_presoftmax)

    def grad_engine_for(self, x: Tensor):
        """The FAST-mode gradient engine (engine.GradientEngine) for inputs shaped like x, built on first use."""
        from .engine import GradientEngine
        if self._grad_engine is None or not self._grad_engine.matches(x):
            self._grad_engine = GradientEngine(self, x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
        return self._grad_engine

    def input_gradient_fast(self, x: Tensor, y_onehot: Tensor, loss_scale: float = 0.5,
                            clip: Tuple[float, float] = (-1e4, 1e3)):
        """(loss, d loss/dx), loss = loss_scale * NLL(clip(var)), through the FAST-mode gradient engine."""
        return self.grad_engine_for(x).input_gradient(x, y_onehot, loss_scale, clip)

    # -- losses of the reference's step functions ---------------------------------------------------------
    def elbo_loss(self, x: Tensor, y_onehot: Tensor, kl_factor: float = 1e-5) -> Tensor:
        """train_on_batch loss (Brats.py:572-576): NLL(clip(var,1e-12,1e3)) + kl_factor * 0.5 * regularisers."""
        p, v = self._forward_fp32(x)
        return nll_gaussian(y_onehot, p, v, clip=(1e-12, 1e3)) + (kl_factor * 0.5) * self.regularization()

    def adversarial_loss(self, x: Tensor, y_onehot: Tensor) -> Tensor:
        """create_adversarial_pattern loss (Brats.py:587-590): 0.5 * NLL with clip [-1e4, 1e3]."""
        p, v = self._forward_fp32(x)
        return 0.5 * nll_gaussian(y_onehot, p, v, clip=(-1e4, 1e3))


def create_adversarial_pattern(model: Density_prop_with_pad_UNET, input_image: Tensor, input_label: Tensor):
    """create_adversarial_pattern (Brats.py:582-596): sign of d(0.5 NLL)/dx.  Returns (signed_grad, gradient)."""
    if model.mode == "fast":
        # tensor-core forward + data-gradient chain (engine.GradientEngine); no autograd tape involved
        _, g = model.input_gradient_fast(input_image.detach(), input_label)
        return torch.sign(g), g
    x = input_image.detach().clone().requires_grad_(True)
    loss = model.adversarial_loss(x, input_label)
    (g,) = torch.autograd.grad(loss, x)
    return torch.sign(g), g
