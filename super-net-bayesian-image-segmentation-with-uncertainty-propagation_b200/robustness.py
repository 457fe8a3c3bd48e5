"""The callers of the hot path in the reference's robustness evaluation, restated on torch tensors:
noise injection of testing() (Brats.py:1247-1283), the FGSM / targeted-PGD loop of main_function
(Brats.py:969-991) and the saliency map (Brats.py:598-609, Brats_functions.py:131-140).

These are host-side drivers: the arithmetic of the network itself runs in the CUDA kernels through `model`.
The reference's `for ... else` bug is not reproduced: untargeted FGSM and targeted PGD are two
separate functions with the intended semantics.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


def salt_and_pepper(image: Tensor, p: float, q: float = 0.5, generator: Optional[torch.Generator] = None) -> Tensor:
    """salt_and_pepper (Brats_functions.py:565-582): a NOISE image that is 1 where salted, low_clip (-1 for signed
    inputs, else 0) where peppered, 0 elsewhere; `p` = flipped fraction, `q` = salt share."""
    low_clip = -1.0 if float(image.min()) < 0 else 0.0
    u = torch.rand(image.shape, generator=generator, device=image.device)
    s = torch.rand(image.shape, generator=generator, device=image.device)
    flipped, salted = u < p, s < q
    out = torch.zeros_like(image)
    out[flipped & salted] = 1.0
    out[flipped & ~salted] = low_clip
    return out


def make_noise(x: Tensor, kind: str, level: float, generator: Optional[torch.Generator] = None) -> Tensor:
    """Random_noise / Speckle / S_and_P branches of testing() (Brats.py:1247-1255)."""
    if kind == "gaussian":
        return torch.randn(x.shape, generator=generator, device=x.device, dtype=x.dtype) * level
    if kind == "speckle":
        return x * (torch.randn(x.shape, generator=generator, device=x.device, dtype=x.dtype) * level)
    if kind == "salt_and_pepper":
        return salt_and_pepper(x, level, generator=generator).to(x.dtype)
    raise ValueError(f"unknown noise kind {kind!r}")


def apply_noise(x: Tensor, labels: Tensor, noise: Tensor, noise_on: str = "all") -> Tensor:
    """Mask the noise to the object ('O': labels > 0), the background ('B': labels == 0) or everywhere, add it and
    clip to the clean batch's [min, max] (Brats.py:1257-1276).  x: [B,H,W,C]; labels: [B,H,W] at input size."""
    lo, hi = x.min(), x.max()
    if noise_on == "O":
        noise = noise * (labels > 0).unsqueeze(-1).to(noise.dtype)
    elif noise_on == "B":
        noise = noise * (labels == 0).unsqueeze(-1).to(noise.dtype)
    elif noise_on != "all":
        raise ValueError("noise_on must be 'O', 'B' or 'all'")
    return torch.minimum(torch.maximum(x + noise, lo), hi)


def snr_db(clean: Tensor, noisy: Tensor) -> float:
    """10 log10(sum clean^2 / sum (noisy - clean)^2) (Brats.py:1279-1283; the reference's batch/channel factors
    cancel in the ratio)."""
    num = clean.double().square().sum()
    den = (noisy.double() - clean.double()).square().sum()
    return float(10.0 * torch.log10(num / den))


def center_crop(x: Tensor, size: int) -> Tensor:
    """crop_numpy_image (Brats_functions.py:528-546) on [B,H,W,...]."""
    oh, ow = (x.shape[1] - size) // 2, (x.shape[2] - size) // 2
    return x[:, oh:oh + size, ow:ow + size]


def one_hot_flat(labels: Tensor, n_labels: int) -> Tensor:
    """labels [B,h,w] int -> [B, h*w, C] float32 (Brats.py:680-683)."""
    return torch.nn.functional.one_hot(labels.long(), n_labels).to(torch.float32).reshape(labels.shape[0], -1, n_labels)


def fgsm_untargeted(model, x: Tensor, labels_out: Tensor, epsilon: float) -> Tuple[Tensor, Tensor]:
    """One signed-gradient step (the `else` block of Brats.py:984-991 with its intended meaning): returns
    (adv_x, signed_grad).  labels_out: [B,h,w] labels already cropped to the output size."""
    from .layers import create_adversarial_pattern
    y = one_hot_flat(labels_out, model.n_labels)
    sign, _ = create_adversarial_pattern(model, x, y)
    lo, hi = x.min(), x.max()
    adv = torch.clamp(x + sign, x - epsilon, x + epsilon)
    return torch.minimum(torch.maximum(adv, lo), hi), sign


def pgd_targeted(model, x: Tensor, labels_out: Tensor, source_class: int, target_class: int, epsilon: float,
                 steps: int = 20, step_size: float = 1.0) -> Tensor:
    """Targeted iterative attack (Brats.py:969-983): relabel `source_class` pixels as `target_class`, follow the
    reference's update adv += step * sign(grad), project to the eps-ball and the clean range each step."""
    from .layers import create_adversarial_pattern
    masked = torch.where(labels_out == source_class, torch.full_like(labels_out, target_class), labels_out)
    y = one_hot_flat(masked, model.n_labels)
    lo, hi = x.min(), x.max()
    adv = x.clone()
    for _ in range(steps):
        sign, _ = create_adversarial_pattern(model, adv, y)
        adv = torch.clamp(adv + step_size * sign, x - epsilon, x + epsilon)
        adv = torch.minimum(torch.maximum(adv, lo), hi)
    return adv


def create_saliency_map(model, x: Tensor, target_class: int, tumor_structure: bool = False,
                        class_only: bool = False):
    """create_saliency_map + get_mask (Brats.py:598-609, Brats_functions.py:131-140): gradient w.r.t. the input of
    the summed probabilities over the pixels predicted as `target_class` (or any non-background class).
    Returns (gradient, relu(gradient), prediction).

    As written, the reference sums ALL class probabilities of the selected pixels (tf.boolean_mask keeps whole
    class vectors), i.e. the number of selected pixels: its saliency is zero up to rounding.  `class_only=True`
    sums only the target class' probability, the quantity the name suggests; the default keeps the reference's
    formula."""
    if getattr(model, "mode", "fp32") == "fast":
        # tensor-core forward + data-gradient chain; the mask is built on the device from the engine's own prediction
        def upstream(p, v):
            label = p.argmax(-1)
            mask = (label > 0) if tumor_structure else (label == target_class)
            g_p = torch.zeros_like(p)
            if class_only:
                g_p[..., target_class] = mask.to(p.dtype)
            else:
                g_p += mask.unsqueeze(-1).to(p.dtype)
            return g_p, None
        g, p, _ = model.grad_engine_for(x).forward_then_input_gradient(x.detach(), upstream)
        g = g.clone()
        return g, torch.relu(g), p.clone()
    xi = x.detach().clone().requires_grad_(True)
    p, _ = model._forward_fp32(xi)
    label = p.argmax(-1)
    mask = (label > 0) if tumor_structure else (label == target_class)
    if class_only:
        mask_sum = (p[..., target_class] * mask.to(p.dtype)).sum()
    else:
        mask_sum = (p * mask.unsqueeze(-1).to(p.dtype)).sum()
    (g,) = torch.autograd.grad(mask_sum, xi)
    return g, torch.relu(g), p.detach()


def mean_predicted_class_variance(p: Tensor, var: Tensor) -> float:
    """Mean output variance at the predicted class (Brats.py:1350-1351)."""
    idx = p.argmax(-1, keepdim=True)
    return float(var.gather(-1, idx).mean())
