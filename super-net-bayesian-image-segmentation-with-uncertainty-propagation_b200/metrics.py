"""Segmentation metrics of the reference's drivers on torch tensors (any device), so the per-step `.numpy()` round
trip of Brats.py:688-705 can stay on the GPU.  Same definitions, same NaN handling:

  sensitivity / precision / specificity   Brats_functions.py:372-397,426-441  (per-image ratio, NaN ratios dropped,
                                                                              mean over the rest)
  dice                                    Brats_functions.py:400-414          (2|A.B| / (|A|+|B|), invalid entries masked)
  mask_tumor / mask_core / mask_enh       Brats_functions.py:443-484          (BraTS label regions: >0; {1,3,4}; ==4)
  compute_H (Hausdorff)                   Brats_functions.py:416-423          stays on the CPU (scipy), like the reference

Inputs are label maps [B,H,W] (integer or float); binary masks are float tensors of 0/1.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor


def _nanmean(x: Tensor) -> Tensor:
    keep = ~torch.isnan(x)
    return x[keep].mean() if bool(keep.any()) else torch.full((), float("nan"), device=x.device, dtype=x.dtype)


def sensitivity(y_true: Tensor, y_pred: Tensor) -> Tensor:
    """TP / (TP + FN) per image, NaN (empty ground truth) dropped (Brats_functions.py:372-385)."""
    t, p = y_true.double(), y_pred.double()
    return _nanmean((t * p).sum((1, 2)) / t.sum((1, 2)))


def precision(y_true: Tensor, y_pred: Tensor) -> Tensor:
    """TP / (TP + FP) per image, NaN (empty prediction) dropped (Brats_functions.py:387-397)."""
    t, p = y_true.double(), y_pred.double()
    return _nanmean((t * p).sum((1, 2)) / p.sum((1, 2)))


def specificity(y_true: Tensor, y_pred: Tensor) -> Tensor:
    """TN / (TN + FP) per image (Brats_functions.py:426-441)."""
    tn = ((y_true == 0) & (y_pred == 0)).double().sum((1, 2))
    den = (y_true == 0).double().sum((1, 2))
    return _nanmean(tn / den)


def dice(y_true: Tensor, y_pred: Tensor) -> Tuple[Tensor, Tensor]:
    """(mean over valid images, per-image values with NaN where |A|+|B| == 0) (Brats_functions.py:400-414)."""
    t, p = y_true.double(), y_pred.double()
    c = 2.0 * (t * p).sum((1, 2)) / (t.sum((1, 2)) + p.sum((1, 2)))
    c = torch.where(torch.isfinite(c), c, torch.full_like(c, float("nan")))
    return _nanmean(c), c


def hausdorff(mask_a: Tensor, mask_b: Tensor) -> float:
    """compute_H (Brats_functions.py:416-423): symmetric directed Hausdorff distance between the two mask IMAGES
    treated as point sets of rows (scipy.spatial.distance.directed_hausdorff on the 2-D arrays, as the reference
    calls it), averaged over the batch.  CPU + scipy, like the reference."""
    from scipy.spatial.distance import directed_hausdorff
    a, b = mask_a.detach().cpu().double().numpy(), mask_b.detach().cpu().double().numpy()
    h = 0.0
    for i in range(a.shape[0]):
        h += max(directed_hausdorff(b[i], a[i])[0], directed_hausdorff(a[i], b[i])[0])
    return h / a.shape[0]


def region_masks(labels: Tensor) -> Dict[str, Tensor]:
    """BraTS evaluation regions as 0/1 float masks: whole tumour (label > 0, mask_tumor), tumour core (labels other
    than 0 and 2 = edema, mask_core), enhancing tumour (label == 4, mask_enh)."""
    lab = labels
    return {"tumor": (lab > 0).to(torch.float32),
            "core": ((lab > 0) & (lab != 2)).to(torch.float32),
            "enh": (lab == 4).to(torch.float32)}


def region_report(y_true: Tensor, y_pred: Tensor, with_hausdorff: bool = False) -> Dict[str, Dict[str, float]]:
    """What mask_tumor / mask_core / mask_enh return (Brats_functions.py:443-484), for all three regions at once."""
    out = {}
    mt, mp = region_masks(y_true), region_masks(y_pred)
    for name in ("tumor", "core", "enh"):
        a, b = mt[name], mp[name]
        d, _ = dice(a, b)
        row = {"dice": float(d), "sensitivity": float(sensitivity(a, b)), "precision": float(precision(a, b)),
               "specificity": float(specificity(a, b))}
        if with_hausdorff:
            row["hausdorff"] = hausdorff(a, b)
        out[name] = row
    return out


def predictions_to_labels(probs: Tensor, out_h: int, out_w: int) -> Tensor:
    """argmax over classes of the model's [B, HW, C] output, back to [B, h, w] (Brats.py:688-690)."""
    return probs.argmax(-1).reshape(probs.shape[0], out_h, out_w)
