"""The data formats on either side of the path, restated on torch:

  * batch feeding: a reference pickle shard is (x [B,C,H,W], y [B,H,W]) (Brats_functions.py:549-562); the model wants
    NHWC float32 and the loss wants centre-cropped one-hot labels flattened to [B, h*w, C] (Brats.py:677-683);
  * checkpoints: the reference saves Keras weights per layer as `w_mu1`/`w_sigma1` (first conv) or `w_mu`/`w_sigma`
    in HWIO / raw pre-softplus form (Brats.py:54-63,107-116,732).  `save_weights` / `load_weights` keep exactly those
    names, layouts and the __init__ layer order in a NumPy .npz archive; `from_keras_h5` reads a Keras-3
    `.weights.h5` when h5py is installed (it is not in this image: the function then raises).
"""
from __future__ import annotations

import pickle
from typing import Dict, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def batch_from_shard(x_nchw, y, out_size: int, n_labels: int, device=None) -> Tuple[Tensor, Tensor, Tensor]:
    """(x [B,C,H,W], y [B,H,W]) -> (x NHWC float32, labels centre-cropped to out_size [B,o,o] int64, one-hot
    [B, o*o, n_labels] float32) -- load_pickle + the label handling of the training loop (Brats_functions.py:549-556,
    crop_numpy_image :500-516, Brats.py:680-683).  With `device` the tensors are moved first, so the transpose,
    crop and one-hot run on the GPU."""
    x = torch.as_tensor(np.asarray(x_nchw))
    y = torch.as_tensor(np.asarray(y))
    if device is not None:
        x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
    x = x.to(torch.float32).permute(0, 2, 3, 1).contiguous()
    start = int((y.shape[1] - out_size) / 2)
    end = y.shape[1] - start
    yc = y[:, start:end, start:end].to(torch.int64)
    onehot = torch.nn.functional.one_hot(yc, n_labels).to(torch.float32).reshape(yc.shape[0], -1, n_labels)
    return x, yc, onehot


def load_shard(path: str):
    """One reference pickle shard: returns the raw (x, y) pair (Brats_functions.py:552-553)."""
    with open(path, "rb") as f:
        x, y = pickle.load(f)
    return x, y


def weight_names(layer_name: str) -> Tuple[str, str]:
    """Variable names the reference gives a layer's weights (Brats.py:54,59 vs 107,112)."""
    return ("w_mu1", "w_sigma1") if layer_name == "conv_input" else ("w_mu", "w_sigma")


def save_weights(model, path: str) -> None:
    """Layer order of Density_prop_with_pad_UNET.__init__, HWIO w_mu, raw (pre-softplus) w_sigma."""
    arrays = {}
    for name in model.conv_names:
        w, s = getattr(model, name).weights()
        mu_n, sg_n = weight_names(name)
        arrays[f"{name}/{mu_n}"] = w.detach().cpu().numpy()
        arrays[f"{name}/{sg_n}"] = s.detach().cpu().numpy()
    arrays["__layers__"] = np.array(model.conv_names)
    np.savez(path, **arrays)


def read_weights(path: str) -> Dict[str, Tuple[Tensor, Tensor]]:
    z = np.load(path, allow_pickle=False)
    out = {}
    for name in [str(n) for n in z["__layers__"]]:
        mu_n, sg_n = weight_names(name)
        out[name] = (torch.from_numpy(z[f"{name}/{mu_n}"]), torch.from_numpy(z[f"{name}/{sg_n}"]))
    return out


def load_weights(model, path: str, device=None):
    return model.load_weight_dict(read_weights(path), device=device)


def from_keras_h5(path: str, conv_names) -> Dict[str, Tuple[Tensor, Tensor]]:
    """Keras-3 `vdp_UNET_model.weights.h5` (Brats.py:732) -> {layer: (w_mu, w_sigma)}.  Keras stores each layer's
    variables in creation order under `layers/<layer>/vars/<i>`: w_mu first, w_sigma second (Brats.py:54-63)."""
    try:
        import h5py
    except ImportError as e:          # not installed in this image; no fallback parser is attempted
        raise RuntimeError("reading a Keras .weights.h5 needs h5py") from e
    out = {}
    with h5py.File(path, "r") as f:
        layers = f["layers"] if "layers" in f else f
        keys = sorted(layers.keys())
        if len(keys) < len(conv_names):
            raise RuntimeError(f"{path}: {len(keys)} layer groups, expected at least {len(conv_names)}")
        groups = [k for k in keys if "vars" in layers[k] and len(layers[k]["vars"]) == 2]
        for name, g in zip(conv_names, groups):
            v = layers[g]["vars"]
            out[name] = (torch.from_numpy(np.asarray(v["0"])), torch.from_numpy(np.asarray(v["1"])))
    return out
