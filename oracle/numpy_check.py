"""TEST INFRASTRUCTURE -- an independent check of the oracle: the reference's forward restated a second time in
plain NumPy (float64), sharing no code and no kernels with oracle/supernet_oracle.py (which runs on torch).

Why: the reference ships no golden vectors and TensorFlow cannot run in this image ("parity unpinned"), so the
torch oracle is the thing every CUDA result is compared with.  This file re-derives the same numbers from the
reference's text by a different route -- every op written out the way TensorFlow documents it, as explicit index
arithmetic: extract_patches as strided windows, conv2d as patches x reshaped filter, max_pool_with_argmax with its
batch-inclusive flat index followed by the flat gather, unpool as reshape / concat-zeros / reshape / pad, the softmax
Jacobian materialised.  tests/golden/hippocampus_b2_numpy_fp64.npz is ITS output (tests/golden/make_golden.py);
tests/test_oracle.py requires the torch oracle (both of its forms) to reproduce it to 1e-12.  This is as close to
pinned as a TensorFlow-less image allows; it does not replace a run of the reference itself.

Each function cites the reference lines it follows (Brats.py / Hippocampus.py of /root/reference).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

Arr = np.ndarray


def softplus(x: Arr) -> Arr:
    """tf.math.softplus = log(1 + exp(x)) (Brats.py:67,120)."""
    return np.logaddexp(0.0, x)


def extract_patches(x: Arr, k: int) -> Arr:
    """tf.image.extract_patches(sizes=[1,k,k,1], strides 1, rates 1, VALID) (Brats.py:69,122,124):
    [B,H,W,C] -> [B,Ho,Wo,k*k*C], depth ordered (patch row, patch column, channel)."""
    B, H, W, C = x.shape
    Ho, Wo = H - k + 1, W - k + 1
    out = np.empty((B, Ho, Wo, k * k * C), dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            out[..., (i * k + j) * C:(i * k + j + 1) * C] = x[:, i:i + Ho, j:j + Wo, :]
    return out


def conv2d_valid(x: Arr, w: Arr) -> Arr:
    """tf.nn.conv2d(x, w, strides=1, padding='VALID') with HWIO filters (Brats.py:66,119): cross-correlation,
    out[b,i,j,n] = sum_{di,dj,c} x[b,i+di,j+dj,c] w[di,dj,c,n], here as patches x filter-matrix."""
    k, _, C, N = w.shape
    P = extract_patches(x, k)
    B, Ho, Wo, K = P.shape
    return (P.reshape(-1, K) @ w.reshape(K, N)).reshape(B, Ho, Wo, N)


def conv_input(x: Arr, w_mu: Arr, w_sigma: Arr) -> Tuple[Arr, Arr]:
    """myConv_input.call (Brats.py:65-76)."""
    k, _, C, N = w_mu.shape
    mu_out = conv2d_valid(x, w_mu)                                           # :66
    vect_sigma = np.broadcast_to(softplus(w_sigma), (k * k * C, N))          # :67-68
    x_matrix = extract_patches(x, k).reshape(x.shape[0], -1, k * k * C)      # :69-71
    sigma = np.square(x_matrix) @ vect_sigma                                 # :73
    return mu_out, sigma.reshape(mu_out.shape)                               # :75-76


def conv_intermediate(mu: Arr, var: Arr, w_mu: Arr, w_sigma: Arr) -> Tuple[Arr, Arr]:
    """myConv_intermediate.call (Brats.py:118-137): sigma1 + sigma2 + sigma3."""
    k, _, C, N = w_mu.shape
    B = mu.shape[0]
    mu_out = conv2d_valid(mu, w_mu)                                          # :119
    vect_sigma = np.broadcast_to(softplus(w_sigma), (k * k * C, N))          # :120-121
    x_matrix = extract_patches(mu, k).reshape(B, -1, k * k * C)              # :122,125
    sigma_matrix = extract_patches(var, k).reshape(B, -1, k * k * C)         # :124,127
    sigma1 = np.square(x_matrix) @ vect_sigma                                # :128
    w_mean = w_mu.reshape(-1, C, N).reshape(-1, N)                           # :130-131
    sigma2 = sigma_matrix @ np.square(w_mean)                                # :132
    sigma3 = sigma_matrix @ vect_sigma                                       # :133
    return mu_out, (sigma1 + sigma2 + sigma3).reshape(mu_out.shape)          # :134-137


def relu(mu: Arr, var: Arr) -> Tuple[Arr, Arr]:
    """myReLU.call + grad_ReLU (Brats.py:220-238): ReluGrad is 1 where the feature is > 0."""
    grad = (mu > 0).astype(mu.dtype)
    return np.maximum(mu, 0.0), np.square(grad) * var                        # :234-237


def maxpool_with_argmax(mu: Arr) -> Tuple[Arr, Arr]:
    """tf.nn.max_pool_with_argmax(ksize 2, strides 2, SAME, include_batch_in_index=True) (Brats.py:172): the flat
    index of element [b, y, x, c] is ((b * H + y) * W + x) * C + c; SAME pads at the bottom / right and padding never
    wins; the first maximum in row-major window order wins a tie."""
    B, H, W, C = mu.shape
    Ho, Wo = -(-H // 2), -(-W // 2)
    padded = np.full((B, 2 * Ho, 2 * Wo, C), -np.inf, dtype=mu.dtype)
    padded[:, :H, :W, :] = mu
    out = np.full((B, Ho, Wo, C), -np.inf, dtype=mu.dtype)
    arg = np.zeros((B, Ho, Wo, C), dtype=np.int64)
    b_idx = np.arange(B).reshape(B, 1, 1, 1)
    c_idx = np.arange(C).reshape(1, 1, 1, C)
    for dy in range(2):
        for dx in range(2):
            cand = padded[:, dy::2, dx::2, :]
            ys = (np.arange(Ho) * 2 + dy).reshape(1, Ho, 1, 1)
            xs = (np.arange(Wo) * 2 + dx).reshape(1, 1, Wo, 1)
            flat = ((b_idx * H + ys) * W + xs) * C + c_idx                    # index into the UNPADDED tensor
            better = cand > out                                               # strict: the earlier window cell keeps a tie
            out = np.where(better, cand, out)
            arg = np.where(better, flat, arg)
    return out, arg


def maxpooling(mu: Arr, var: Arr) -> Tuple[Arr, Arr]:
    """mymaxpooling.call + get_pooled (Brats.py:171-174, 206-216)."""
    mu_out, argmax = maxpool_with_argmax(mu)
    return mu_out, var.reshape(-1)[argmax]                                   # :215 tf.gather on the flattened tensor


def unpool(value: Arr) -> Arr:
    """unpool (Brats.py:178-203), op by op: reshape, concat zeros along each spatial axis (last first), reshape to the
    doubled size, pad one row / column in FRONT."""
    sh = list(value.shape)
    dim = len(sh[1:-1])
    out = value.reshape([-1] + sh[-dim:])                                    # :196
    for i in range(dim, 0, -1):
        out = np.concatenate([out, np.zeros_like(out)], axis=i)              # :197-198
    out = out.reshape([-1] + [s * 2 for s in sh[1:-1]] + [sh[-1]])           # :199-200
    return np.pad(out, [[0, 0], [1, 0], [1, 0], [0, 0]])                     # :201-202


def padding(mu: Arr, var: Arr, pad: Tuple[int, int], sigma_fill: float) -> Tuple[Arr, Arr]:
    """mypadding.call (Brats.py:159-163)."""
    p = [[0, 0], list(pad), list(pad), [0, 0]]
    return np.pad(mu, p), np.pad(var, p, constant_values=sigma_fill)


def crop_tensor(x1: Arr, x2: Arr) -> Arr:
    """crop_tensor (Brats_functions.py:518-526): tf.slice of x1 at offset (H1 - H2) // 2 to x2's H, W."""
    oh, ow = (x1.shape[1] - x2.shape[1]) // 2, (x1.shape[2] - x2.shape[2]) // 2
    return x1[:, oh:oh + x2.shape[1], ow:ow + x2.shape[2], :]


def conc(mu_d: Arr, var_d: Arr, mu_e: Arr, var_e: Arr) -> Tuple[Arr, Arr]:
    """myConc.call (Brats.py:247-261): decoder first."""
    return (np.concatenate([mu_d, crop_tensor(mu_e, mu_d)], axis=-1),
            np.concatenate([var_d, crop_tensor(var_e, var_d)], axis=-1))


def softmax_moments(mu: Arr, var: Arr) -> Tuple[Arr, Arr]:
    """mysoftmax.call (Brats.py:269-283) with the [C, C] Jacobian per pixel materialised; the batch axis is kept for
    B == 1 (SURVEY.md 8b deviation from the bare tf.squeeze)."""
    B, C = mu.shape[0], mu.shape[3]
    m = mu.reshape(B, -1, C)
    s = var.reshape(B, -1, C)
    e = np.exp(m - m.max(axis=-1, keepdims=True))
    p = e / e.sum(axis=-1, keepdims=True)                                    # :272
    ppT = p[..., :, None] * p[..., None, :]                                  # :273-275
    grad = p[..., :, None] * np.eye(C) - ppT                                 # :276-277 diag(p) - p p^T
    sigma = (np.square(grad) @ s[..., None])[..., 0]                         # :278-281
    return p, sigma


def unet_forward(x: Arr, W: Dict[str, Tuple[Arr, Arr]], variant: str):
    """Density_prop_with_pad_UNET.call: Brats.py:377-457 (four levels, mypad1 [1,0] before the bottleneck, sigma_fill
    0.1, :370-372) or Hippocampus.py:373-421 (two levels, sigma_fill 0.02, :366-368; mypad1 constructed but unused)."""
    levels, fill = (4, 0.1) if variant == "brats" else (2, 0.02)
    m, s = relu(*conv_input(x, *W["conv_input"]))
    m, s = relu(*conv_intermediate(m, s, *W["conv1"]))
    skips = [(m, s)]
    ci = 2
    for lvl in range(1, levels + 1):
        m, s = maxpooling(m, s)
        if variant == "brats" and lvl == levels:
            m, s = padding(m, s, (1, 0), fill)                                # Brats.py:407
        for _ in range(2):
            m, s = relu(*conv_intermediate(m, s, *W[f"conv{ci}"]))
            ci += 1
        if lvl < levels:
            skips.append((m, s))
    for d in range(1, levels + 1):
        me, se = skips[levels - d]
        m, s = unpool(m), unpool(s)
        m, s = conv_intermediate(m, s, *W[f"up{d}_conv2x2"])                  # no ReLU after the up-conv
        m, s = padding(m, s, (3, 3), fill)
        m, s = conc(m, s, me, se)
        m, s = relu(*conv_intermediate(m, s, *W[f"up{d}_conv1"]))
        m, s = padding(m, s, (2, 2), fill)
        m, s = relu(*conv_intermediate(m, s, *W[f"up{d}_conv2"]))
    mf, sf = conv_intermediate(m, s, *W["conv_final"])
    p, v = softmax_moments(mf, sf)
    return p, v, mf, sf
