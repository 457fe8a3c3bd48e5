"""CPU oracle for the SUPER-Net moment-propagation hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.
The product path (the CUDA kernels behind ``include/supernet.h``) never routes
through it.

PARITY UNPINNED: the reference (`/root/reference`, TensorFlow/Keras) ships no tests,
golden vectors or saved outputs, and TensorFlow is not installable in this image, so
the reference itself cannot be executed here.  This file is a line-by-line
restatement of the reference *formulas*; it is pinned only by (i) two independent
formulations that must agree (the "as written" patch-matmul form and the conv form),
(ii) a Monte-Carlo check of the variance formula, (iii) structural identities,
(iv) hand-computed tiny cases and (v) an INDEPENDENT NumPy restatement of the as-written
forward (oracle/numpy_check.py: explicit index arithmetic, no torch) whose committed
whole-network outputs (tests/golden/*_numpy_fp64.npz) both forms of this oracle must
reproduce to 1e-12 -- see tests/test_oracle.py.  That is as close to pinned as an image
without TensorFlow allows; it is still not a run of the reference.

Everything is NHWC, weights HWIO ``[k,k,Cin,Cout]``, "sigma" means VARIANCE (as in the
reference).  All functions are differentiable torch-CPU code so autograd provides the
reference gradients (the reference obtains its gradients from tf.GradientTape).

Reference citations are ``file:line`` into /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def softplus(x: Tensor) -> Tensor:
    """tf.math.softplus (Brats.py:67,120,319): log(1+exp(x)), overflow-safe."""
    return F.softplus(x, beta=1.0, threshold=1e9)


def _nhwc_to_nchw(x: Tensor) -> Tensor:
    return x.permute(0, 3, 1, 2)


def _nchw_to_nhwc(x: Tensor) -> Tensor:
    return x.permute(0, 2, 3, 1)


def extract_patches(x: Tensor, k: int) -> Tensor:
    """tf.image.extract_patches(sizes=[1,k,k,1], strides 1, rates 1, VALID) (Brats.py:69,122,124).

    Input NHWC ``[B,H,W,C]`` -> ``[B,Ho,Wo,k*k*C]`` with depth order (row, col, channel),
    the order that matches the ``[k,k,Cin,Cout] -> [k*k*Cin,Cout]`` reshape at Brats.py:130-131.
    """
    B, H, W, C = x.shape
    Ho, Wo = H - k + 1, W - k + 1
    cols = F.unfold(_nhwc_to_nchw(x), kernel_size=k)            # [B, C*k*k, Ho*Wo] order (c,kh,kw)
    cols = cols.reshape(B, C, k, k, Ho, Wo).permute(0, 4, 5, 2, 3, 1)  # [B,Ho,Wo,kh,kw,c]
    return cols.reshape(B, Ho, Wo, k * k * C)


def conv2d_valid(x: Tensor, w: Tensor) -> Tensor:
    """tf.nn.conv2d(x, w, strides=1, padding='VALID') on NHWC/HWIO (Brats.py:66,119)."""
    y = F.conv2d(_nhwc_to_nchw(x), w.permute(3, 2, 0, 1))
    return _nchw_to_nhwc(y)


# --------------------------------------------------------------------------------------
# L0 moment layers -- "as written" (mirrors the reference's op sequence) and conv form
# --------------------------------------------------------------------------------------
def conv_input_as_written(x: Tensor, w_mu: Tensor, w_sigma: Tensor) -> Tuple[Tensor, Tensor]:
    """myConv_input.call, Brats.py:65-76: deterministic input, random weights."""
    k, _, cin, cout = w_mu.shape
    mu_out = conv2d_valid(x, w_mu)                                        # :66
    ws = softplus(w_sigma)                                                # :67
    vect_sigma = ws.expand(k * k * cin, cout)                             # :68
    x_matrix = extract_patches(x, k).reshape(x.shape[0], -1, k * k * cin)  # :69-71
    sigma = torch.matmul(x_matrix.square(), vect_sigma)                   # :73
    return mu_out, sigma.reshape(mu_out.shape)                            # :75-76


def conv_intermediate_as_written(mu: Tensor, var: Tensor, w_mu: Tensor, w_sigma: Tensor) -> Tuple[Tensor, Tensor]:
    """myConv_intermediate.call, Brats.py:118-137: both input and weights random."""
    k, _, cin, cout = w_mu.shape
    B = mu.shape[0]
    mu_out = conv2d_valid(mu, w_mu)                                       # :119
    ws = softplus(w_sigma)                                                # :120
    vect_sigma = ws.expand(k * k * cin, cout)                             # :121
    x_matrix = extract_patches(mu, k).reshape(B, -1, k * k * cin)         # :122,125
    sigma_matrix = extract_patches(var, k).reshape(B, -1, k * k * cin)    # :124,127
    sigma1 = torch.matmul(x_matrix.square(), vect_sigma)                  # :128
    w_mean = w_mu.reshape(-1, cin, cout).reshape(-1, cout)                # :130-131
    sigma2 = torch.matmul(sigma_matrix, w_mean.square())                  # :132
    sigma3 = torch.matmul(sigma_matrix, vect_sigma)                       # :133
    sigma = sigma1 + sigma2 + sigma3                                      # :134
    return mu_out, sigma.reshape(mu_out.shape)                            # :135-137


def _box_sum(q: Tensor, k: int) -> Tensor:
    """k x k VALID box-sum of a ``[B,H,W]`` map."""
    ones = torch.ones(1, 1, k, k, dtype=q.dtype)
    return F.conv2d(q.unsqueeze(1), ones).squeeze(1)


def conv_input_conv_form(x: Tensor, w_mu: Tensor, w_sigma: Tensor) -> Tuple[Tensor, Tensor]:
    """var[b,i,j,n] = softplus(w_sigma[n]) * box_k(sum_c x^2)."""
    k = w_mu.shape[0]
    mu_out = conv2d_valid(x, w_mu)
    r = _box_sum(x.square().sum(-1), k)
    return mu_out, r.unsqueeze(-1) * softplus(w_sigma)


def conv_intermediate_conv_form(mu: Tensor, var: Tensor, w_mu: Tensor, w_sigma: Tensor) -> Tuple[Tensor, Tensor]:
    """var' = var (*) W^2 + s_n * box_k(sum_c (mu^2 + var))."""
    k = w_mu.shape[0]
    mu_out = conv2d_valid(mu, w_mu)
    r = _box_sum((mu.square() + var).sum(-1), k)
    var_out = conv2d_valid(var, w_mu.square()) + r.unsqueeze(-1) * softplus(w_sigma)
    return mu_out, var_out


def unpool(value: Tensor) -> Tensor:
    """unpool, Brats.py:178-203: ``out[b,2x+1,2y+1,c] = in[b,x,y,c]``, zeros elsewhere, size 2H+1.

    Written exactly as the reference does it: concat zeros along each spatial axis,
    reshape, then pad one row/column in front.
    """
    sh = list(value.shape)
    dim = len(sh[1:-1])
    out = value.reshape([-1] + sh[-dim:])                                 # :196
    for i in range(dim, 0, -1):
        out = torch.cat([out, torch.zeros_like(out)], i)                  # :197-198
    out_size = [-1] + [s * 2 for s in sh[1:-1]] + [sh[-1]]
    out = out.reshape(out_size)                                           # :199-200
    return F.pad(out, (0, 0, 1, 0, 1, 0))                                 # :201-202 pad W,H by [1,0]


def upsampling(mu: Tensor, var: Tensor) -> Tuple[Tensor, Tensor]:
    """myupsampling.call, Brats.py:145-148."""
    return unpool(mu), unpool(var)


def padding(mu: Tensor, var: Tensor, pad_size: Sequence[int] = (2, 2), sigma_fill: float = 0.0) -> Tuple[Tensor, Tensor]:
    """mypadding.call, Brats.py:159-163: zero-pad the mean, constant-pad the variance."""
    a, b = int(pad_size[0]), int(pad_size[1])
    mu_out = F.pad(mu, (0, 0, a, b, a, b), value=0.0)
    var_out = F.pad(var, (0, 0, a, b, a, b), value=float(sigma_fill))
    return mu_out, var_out


def maxpooling(mu: Tensor, var: Tensor) -> Tuple[Tensor, Tensor]:
    """mymaxpooling.call + get_pooled, Brats.py:171-174,206-216.

    2x2/2 max-pool of the mean (SAME; every size on the path is even so SAME == VALID; for odd
    sizes SAME pads at the bottom/right and padded cells never win), variance gathered at the
    arg-max.  Ties: the first maximum in row-major window order wins, which is what
    max_pool_with_argmax does on CPU and what torch's max_pool2d does.
    """
    B, H, W, C = mu.shape
    m = _nhwc_to_nchw(mu)
    v = _nhwc_to_nchw(var)
    ph, pw = H % 2, W % 2
    if ph or pw:
        m = F.pad(m, (0, pw, 0, ph), value=float("-inf"))
        v = F.pad(v, (0, pw, 0, ph), value=0.0)
    mo, idx = F.max_pool2d(m, 2, 2, return_indices=True)
    vo = v.flatten(2).gather(2, idx.flatten(2)).reshape(mo.shape)
    return _nchw_to_nhwc(mo), _nchw_to_nhwc(vo)


def relu(mu: Tensor, var: Tensor) -> Tuple[Tensor, Tensor]:
    """myReLU.call + grad_ReLU, Brats.py:220-238.

    The gate comes from an inner GradientTape (ReluGrad: ``features > 0``) and is a constant for
    the outer autodiff, hence the detach.
    """
    gate = (mu > 0).to(mu.dtype).detach()
    return F.relu(mu), gate.square() * var


def crop_tensor(x1: Tensor, x2: Tensor) -> Tensor:
    """crop_tensor, Brats_functions.py:518-526: centre-crop x1 (encoder) to x2's H,W."""
    oh = (x1.shape[1] - x2.shape[1]) // 2
    ow = (x1.shape[2] - x2.shape[2]) // 2
    return x1[:, oh:oh + x2.shape[1], ow:ow + x2.shape[2], :]


def conc(mu_d: Tensor, var_d: Tensor, mu_e: Tensor, var_e: Tensor) -> Tuple[Tensor, Tensor]:
    """myConc.call, Brats.py:247-261: decoder first, cropped encoder second."""
    return (torch.cat([mu_d, crop_tensor(mu_e, mu_d)], -1),
            torch.cat([var_d, crop_tensor(var_e, var_d)], -1))


def softmax_as_written(mu: Tensor, var: Tensor) -> Tuple[Tensor, Tensor]:
    """mysoftmax.call, Brats.py:269-283 with the Jacobian materialised.

    Deviation: the output keeps ``[B,HW,C]`` even for B == 1 (the reference's
    ``tf.squeeze`` would drop the batch axis).
    """
    B, C = mu.shape[0], mu.shape[3]
    mu_r = mu.reshape(B, -1, C)
    var_r = var.reshape(B, -1, C)
    p = torch.softmax(mu_r, -1)                                           # :272
    ppT = p.unsqueeze(3) @ p.unsqueeze(2)                                 # :273-275
    grad = torch.diag_embed(p) - ppT                                      # :276-277
    sigma = grad.square() @ var_r.unsqueeze(3)                            # :278-280
    return p, sigma.squeeze(3)


def softmax_closed_form(mu: Tensor, var: Tensor) -> Tuple[Tensor, Tensor]:
    """Identity: var_i = p_i^2 [ (1-2 p_i) v_i + sum_j p_j^2 v_j ]."""
    B, C = mu.shape[0], mu.shape[3]
    p = torch.softmax(mu.reshape(B, -1, C), -1)
    v = var.reshape(B, -1, C)
    s = (p.square() * v).sum(-1, keepdim=True)
    return p, p.square() * ((1 - 2 * p) * v + s)


# --------------------------------------------------------------------------------------
# loss (Brats.py:293-320, 569-596)
# --------------------------------------------------------------------------------------
def nll_gaussian(y_test: Tensor, y_pred_mean: Tensor, y_pred_sd: Tensor) -> Tensor:
    """nll_gaussian, Brats.py:293-311."""
    eps = 1e-3
    inv = 1.0 / (y_pred_sd + eps)                                         # :296
    mu_square = (y_pred_mean - y_test).square()                           # :297
    loss1 = (mu_square * inv).sum(-1)                                     # :298-300 [1,C]x[C,1] matmul
    loss = loss1.mean(0).mean()                                           # :301-302
    if not torch.isfinite(loss):                                          # :304-305 NaN/Inf -> 0
        loss = torch.zeros_like(loss)
    loss2 = torch.log(torch.prod(y_pred_sd + eps, dim=-1)).mean()         # :307-309
    return 0.5 * (loss + loss2)                                           # :310


def sigma_regularizer(w_sigma: Tensor, strength: float) -> Tensor:
    """sigma_regularizer.__call__, Brats.py:314-320 (strength = k*k, Brats.py:53,61)."""
    f_s = softplus(w_sigma)
    return -strength * torch.mean(1.0 + torch.log(f_s) - f_s, dim=-1)


def l2_regularizer(w_mu: Tensor, tau: float = 1.0) -> Tensor:
    """tf.keras.regularizers.l2(tau): tau * sum(w^2) (Brats.py:56,109)."""
    return tau * w_mu.square().sum()


# --------------------------------------------------------------------------------------
# model graph (Brats.py:323-457, Hippocampus.py:335-421)
# --------------------------------------------------------------------------------------
@dataclass
class ConvSpec:
    name: str
    cin: int
    cout: int
    k: int
    sigma_min: float = -12.0
    sigma_max: float = -4.6


def unet_conv_specs(variant: str, n_kernels: int, n_labels: int, in_ch: int) -> List[ConvSpec]:
    """Conv layers in ``__init__`` order (Brats.py:331-367 / Hippocampus.py:343-363)."""
    n = n_kernels
    hi = dict(sigma_min=-4.6, sigma_max=-2.2)
    if variant == "brats":
        return [
            ConvSpec("conv_input", in_ch, n, 3), ConvSpec("conv1", n, n, 3),
            ConvSpec("conv2", n, 2 * n, 3), ConvSpec("conv3", 2 * n, 2 * n, 3),
            ConvSpec("conv4", 2 * n, 4 * n, 3), ConvSpec("conv5", 4 * n, 4 * n, 3),
            ConvSpec("conv6", 4 * n, 8 * n, 3), ConvSpec("conv7", 8 * n, 8 * n, 3),
            ConvSpec("conv8", 8 * n, 16 * n, 3), ConvSpec("conv9", 16 * n, 16 * n, 3),
            ConvSpec("up1_conv2x2", 16 * n, 8 * n, 2, **hi),                 # Brats.py:349
            ConvSpec("up1_conv1", 16 * n, 8 * n, 3), ConvSpec("up1_conv2", 8 * n, 8 * n, 3),
            ConvSpec("up2_conv2x2", 8 * n, 4 * n, 2, **hi),                  # :353
            ConvSpec("up2_conv1", 8 * n, 4 * n, 3), ConvSpec("up2_conv2", 4 * n, 4 * n, 3),
            ConvSpec("up3_conv2x2", 4 * n, 2 * n, 2),                        # :357 default sigma range
            ConvSpec("up3_conv1", 4 * n, 2 * n, 3), ConvSpec("up3_conv2", 2 * n, 2 * n, 3),
            ConvSpec("up4_conv2x2", 2 * n, n, 2),                            # :362
            ConvSpec("up4_conv1", 2 * n, n, 3), ConvSpec("up4_conv2", n, n, 3),
            ConvSpec("conv_final", n, n_labels, 1, **hi),                    # :367
        ]
    if variant == "hippocampus":
        return [
            ConvSpec("conv_input", in_ch, n, 3), ConvSpec("conv1", n, n, 3),
            ConvSpec("conv2", n, 2 * n, 3), ConvSpec("conv3", 2 * n, 2 * n, 3),
            ConvSpec("conv4", 2 * n, 4 * n, 3), ConvSpec("conv5", 4 * n, 4 * n, 3),
            ConvSpec("up1_conv2x2", 4 * n, 2 * n, 2, **hi),                  # Hippocampus.py:354
            ConvSpec("up1_conv1", 4 * n, 2 * n, 3), ConvSpec("up1_conv2", 2 * n, 2 * n, 3),
            ConvSpec("up2_conv2x2", 2 * n, n, 2, **hi),                      # :358
            ConvSpec("up2_conv1", 2 * n, n, 3), ConvSpec("up2_conv2", n, n, 3),
            ConvSpec("conv_final", n, n_labels, 1, **hi),                    # :363
        ]
    raise ValueError(f"unknown variant {variant!r}")


def truncated_normal_(t: Tensor, mean: float, std: float, gen: torch.Generator) -> Tensor:
    """tf.keras.initializers.TruncatedNormal: resample values beyond mean +- 2 std (Brats.py:52)."""
    t.normal_(mean, std, generator=gen)
    while True:
        bad = (t - mean).abs() > 2 * std
        n = int(bad.sum())
        if n == 0:
            return t
        t[bad] = torch.empty(n, dtype=t.dtype).normal_(mean, std, generator=gen)


def make_weights(variant: str, n_kernels: int = 32, n_labels: int = 4, in_ch: int = 4,
                 seed: int = 1000, dtype: torch.dtype = torch.float32) -> Dict[str, Tuple[Tensor, Tensor]]:
    """Fixed-seed random-init weights shared by the oracle and the CUDA path.

    Per conv: ``w_mu ~ TruncNormal(0, 0.1)`` HWIO fp32, raw ``w_sigma ~ U[sigma_min, sigma_max]``;
    layer i uses ``torch.Generator().manual_seed(seed + i)``.  Values are generated in fp32 and
    then cast, so fp32 and fp64 models share bit-identical parameters.
    """
    out: Dict[str, Tuple[Tensor, Tensor]] = {}
    for i, s in enumerate(unet_conv_specs(variant, n_kernels, n_labels, in_ch)):
        g = torch.Generator(device="cpu").manual_seed(seed + i)
        w_mu = truncated_normal_(torch.empty(s.k, s.k, s.cin, s.cout, dtype=torch.float32), 0.0, 0.1, g)
        w_sigma = torch.empty(s.cout, dtype=torch.float32).uniform_(s.sigma_min, s.sigma_max, generator=g)
        out[s.name] = (w_mu.to(dtype), w_sigma.to(dtype))
    return out


def make_input(variant: str, batch: int, seed: int = 2025, alpha: float = 1.0,
               dtype: torch.dtype = torch.float32, in_ch: Optional[int] = None) -> Tensor:
    """Synthetic slices ``alpha * U[0,1)`` NHWC: BraTS 204x204x4, Hippocampus 64x64x1."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if variant == "brats":
        shape = (batch, 204, 204, 4 if in_ch is None else in_ch)
    else:
        shape = (batch, 64, 64, 1 if in_ch is None else in_ch)
    return (torch.rand(shape, generator=g, dtype=torch.float32) * alpha).to(dtype)


def make_labels(batch: int, hw: int, n_labels: int, seed: int = 7, dtype: torch.dtype = torch.float32) -> Tensor:
    """One-hot labels ``[B, HW, C]`` already at the output resolution (Brats.py:680-683)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    y = torch.randint(0, n_labels, (batch, hw), generator=g)
    return F.one_hot(y, n_labels).to(dtype)


#: input scale that brings the BraTS final-logit std to O(1) at random init (SURVEY.md 8d, E;
#: legitimate because the mean path is degree-1 homogeneous -- there is no bias).
BRATS_ALPHA = 1.0 / 8192.0


@dataclass
class UNetOracle:
    """Density_prop_with_pad_UNET (Brats.py:323-457 / Hippocampus.py:335-421) on the CPU."""
    variant: str = "brats"
    n_kernels: int = 32
    n_labels: int = 4
    in_ch: int = 4
    dtype: torch.dtype = torch.float64
    seed: int = 1000
    form: str = "conv"            # "conv" | "as_written"
    weights: Dict[str, Tuple[Tensor, Tensor]] = field(default_factory=dict)

    def __post_init__(self):
        if not self.weights:
            self.weights = make_weights(self.variant, self.n_kernels, self.n_labels, self.in_ch,
                                        self.seed, self.dtype)
        self.sigma_fill = 0.1 if self.variant == "brats" else 0.02   # Brats.py:370-372 / Hippocampus.py:366-368
        self.specs = unet_conv_specs(self.variant, self.n_kernels, self.n_labels, self.in_ch)

    # -- parameters ------------------------------------------------------------------
    def parameters(self) -> List[Tensor]:
        ps: List[Tensor] = []
        for s in self.specs:
            ps.extend(self.weights[s.name])
        return ps

    def requires_grad_(self, flag: bool = True) -> "UNetOracle":
        self.weights = {k: (a.detach().clone().requires_grad_(flag), b.detach().clone().requires_grad_(flag))
                        for k, (a, b) in self.weights.items()}
        return self

    def regularization(self) -> Tensor:
        """add_n(model.losses), Brats.py:575: per conv l2(w_mu) + sigma_regularizer(k*k)(w_sigma)."""
        tot = torch.zeros((), dtype=self.dtype)
        for s in self.specs:
            w_mu, w_sigma = self.weights[s.name]
            tot = tot + l2_regularizer(w_mu) + sigma_regularizer(w_sigma, float(s.k * s.k))
        return tot

    # -- layers ----------------------------------------------------------------------
    def _first(self, x):
        f = conv_input_as_written if self.form == "as_written" else conv_input_conv_form
        return f(x, *self.weights["conv_input"])

    def _conv(self, name, mu, var):
        f = conv_intermediate_as_written if self.form == "as_written" else conv_intermediate_conv_form
        return f(mu, var, *self.weights[name])

    def forward(self, x: Tensor, return_presoftmax: bool = False, taps: Optional[dict] = None,
                trajectory: Optional[Dict[str, Tuple[Tensor, Tensor]]] = None):
        """call(), Brats.py:377-457 / Hippocampus.py:373-421.

        ``trajectory`` (layer name -> post-activation (mean, variance) another implementation computed for the same
        input) re-runs the graph ON THAT FORWARD TRAJECTORY: every conv output takes the given value (straight-through:
        the autograd formulas stay this function's) and every ReLU gate / pooling arg-max is decided by the given
        values.  Autograd through it is the exact gradient arithmetic given those decisions, which separates "the
        other forward flipped a gate" from "the other backward is inaccurate" in a gradient comparison."""
        x = x.to(self.dtype)
        fill = self.sigma_fill
        levels = 4 if self.variant == "brats" else 2

        def post(name, m, s, has_relu):
            if trajectory is None or name not in trajectory:
                m, s = relu(m, s) if has_relu else (m, s)
            else:
                mf, sf = (t.to(self.dtype) for t in trajectory[name])
                if has_relu:
                    gate = (mf > 0).to(self.dtype)
                    m, s = m * gate, s * gate
                m, s = m + (mf - m).detach(), s + (sf - s).detach()
            if taps is not None:
                taps[name] = (m.detach(), s.detach())
            return m, s

        m, s = post("conv_input", *self._first(x), True)
        m, s = post("conv1", *self._conv("conv1", m, s), True)
        skips = [(m, s)]
        ci = 2
        for lvl in range(1, levels + 1):
            m, s = maxpooling(m, s)
            if self.variant == "brats" and lvl == levels:
                m, s = padding(m, s, (1, 0), fill)                         # mypad1, Brats.py:407
            for _ in range(2):
                name = f"conv{ci}"
                m, s = post(name, *self._conv(name, m, s), True)
                ci += 1
            if lvl < levels:
                skips.append((m, s))
        for d in range(1, levels + 1):
            me, se = skips[levels - d]
            m, s = upsampling(m, s)
            m, s = post(f"up{d}_conv2x2", *self._conv(f"up{d}_conv2x2", m, s), False)
            m, s = padding(m, s, (3, 3), fill)                             # mypad_up6
            m, s = conc(m, s, me, se)
            m, s = post(f"up{d}_conv1", *self._conv(f"up{d}_conv1", m, s), True)
            m, s = padding(m, s, (2, 2), fill)                             # mypad
            m, s = post(f"up{d}_conv2", *self._conv(f"up{d}_conv2", m, s), True)
        mf, sf = post("conv_final", *self._conv("conv_final", m, s), False)
        sm = softmax_as_written if self.form == "as_written" else softmax_closed_form
        p, v = sm(mf, sf)
        if return_presoftmax:
            return p, v, mf, sf
        return p, v

    __call__ = forward

    # -- losses (Brats.py:569-596) ------------------------------------------------------
    def elbo_loss(self, x: Tensor, y_onehot: Tensor, kl_factor: float = 1e-5, trajectory=None) -> Tensor:
        """train_on_batch loss, Brats.py:572-576."""
        p, v = self.forward(x, trajectory=trajectory)
        nll = nll_gaussian(y_onehot.to(self.dtype), p, torch.clamp(v, 1e-12, 1e3))
        return nll + kl_factor * 0.5 * self.regularization()

    def adversarial_loss(self, x: Tensor, y_onehot: Tensor, trajectory=None) -> Tensor:
        """create_adversarial_pattern loss, Brats.py:587-590: 0.5 * NLL with clip [-1e4, 1e3]."""
        p, v = self.forward(x, trajectory=trajectory)
        return 0.5 * nll_gaussian(y_onehot.to(self.dtype), p, torch.clamp(v, -1e4, 1e3))

    def fgsm_gradient(self, x: Tensor, y_onehot: Tensor, trajectory=None) -> Tuple[Tensor, Tensor]:
        """Brats.py:583-596: returns (d loss / d x, loss)."""
        x = x.detach().clone().to(self.dtype).requires_grad_(True)
        loss = self.adversarial_loss(x, y_onehot, trajectory=trajectory)
        (g,) = torch.autograd.grad(loss, x)
        return g, loss.detach()


def output_hw(variant: str) -> int:
    return 186 if variant == "brats" else 54


def rel_l2(a: Tensor, b: Tensor) -> float:
    """||a-b||_2 / ||b||_2 over the whole tensor (SURVEY.md 8d parity metric)."""
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def argmax_agreement(p_a: Tensor, p_b: Tensor) -> float:
    return float((p_a.argmax(-1) == p_b.argmax(-1)).double().mean())
