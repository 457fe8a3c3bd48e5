"""The data formats on either side of the path, restated on torch:

  * batch feeding: a reference pickle shard is (x [B,C,H,W], y [B,H,W]) (Brats_functions.py:549-562); the model wants
    NHWC float32 and the loss wants centre-cropped one-hot labels flattened to [B, h*w, C] (Brats.py:677-683);
  * checkpoints: the reference saves Keras weights per layer as `w_mu1`/`w_sigma1` (first conv) or `w_mu`/`w_sigma`
    in HWIO / raw pre-softplus form (Brats.py:54-63,107-116,732).  `save_weights` / `load_weights` keep exactly those
    names, layouts and the __init__ layer order in a NumPy .npz archive; `from_keras_h5` / `to_keras_h5` read and
    write the Keras-3 `.weights.h5` layout through the in-repo HDF5 parser (h5min.py; no h5py in this image);
  * results: the `[logits_, sigma_, x, y]` pickle of testing() and the adversarial test (Brats.py:1375,1427).
"""
from __future__ import annotations

import pickle
import re
from typing import Dict, Optional, Tuple

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


_AUTO_NAME = re.compile(r"^(.*?)(?:_(\d+))?$")


def _natural_key(name: str):
    """Keras auto-names count per class: my_conv_intermediate, my_conv_intermediate_1, ..., _21.  Sorted as text,
    _10 comes before _2; the creation order is the NUMERIC order of the suffix."""
    m = _AUTO_NAME.match(name)
    return (m.group(1), int(m.group(2)) if m.group(2) is not None else 0)


def from_keras_h5(path: str, conv_names, expected_shapes: Optional[Dict[str, Tuple[int, ...]]] = None
                  ) -> Dict[str, Tuple[Tensor, Tensor]]:
    """Keras-3 `vdp_UNET_model.weights.h5` (Brats.py:732; loaded at :611-622,933,1195) -> {layer: (w_mu, w_sigma)}.

    Keras' H5IOStore writes each layer's variables in creation order as `<layer path>/vars/<i>`: w_mu is variable 0,
    w_sigma variable 1 (add_weight order, Brats.py:54-63,107-116).  The layer path of a subclassed model ends in the
    ATTRIBUTE name (conv1 ... conv_final, conv_input, up1_conv2x2 ...; Brats.py:331-367) -- then groups are looked up
    by name -- or, in files that list layers by auto-name (my_conv_input, my_conv_intermediate, my_conv_intermediate_1,
    ...), the numeric order of the suffix is the construction order of __init__.  Every pair is checked: w_mu is 4-D
    HWIO, w_sigma is [Cout] and, with `expected_shapes` (layer -> w_mu shape), the exact shape.
    Read with the in-repo HDF5 parser (h5min.py): no h5py in this image."""
    from . import h5min
    groups: Dict[str, Dict[int, np.ndarray]] = {}
    for p, arr in h5min.read_h5(path).items():
        parts = p.strip("/").split("/")
        if len(parts) >= 3 and parts[-2] == "vars" and parts[-1].isdigit():
            groups.setdefault("/".join(parts[:-2]), {})[int(parts[-1])] = arr
    pairs = {g: v for g, v in groups.items() if set(v) == {0, 1}}
    leaf = {}
    for g, v in pairs.items():
        leaf.setdefault(g.split("/")[-1], []).append(v)
    conv_names = list(conv_names)
    if all(len(leaf.get(n, [])) == 1 for n in conv_names):
        chosen = {n: leaf[n][0] for n in conv_names}
    else:
        autos = sorted((k for k in leaf if k.startswith("my_conv")), key=_natural_key)
        firsts = [k for k in autos if k.startswith("my_conv_input")]
        inter = [k for k in autos if k.startswith("my_conv_intermediate")]
        if len(firsts) != 1 or len(inter) != len(conv_names) - 1:
            raise RuntimeError(f"{path}: cannot match the layer groups {sorted(leaf)} to {conv_names}")
        order = iter(inter)
        chosen = {n: (leaf[firsts[0]][0] if n == "conv_input" else leaf[next(order)][0]) for n in conv_names}
    out = {}
    for n in conv_names:
        w, s = chosen[n][0], chosen[n][1]
        if w.ndim != 4 or s.ndim != 1 or s.shape[0] != w.shape[-1] or w.shape[0] != w.shape[1]:
            raise RuntimeError(f"{path}: layer {n}: variables {w.shape}, {s.shape} are not (HWIO w_mu, [Cout] w_sigma)")
        if expected_shapes is not None and tuple(w.shape) != tuple(expected_shapes[n]):
            raise RuntimeError(f"{path}: layer {n}: w_mu {tuple(w.shape)}, expected {tuple(expected_shapes[n])}")
        out[n] = (torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)),
                  torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)))
    return out


def to_keras_h5(path: str, weights: Dict[str, Tuple[Tensor, Tensor]], conv_names) -> None:
    """Write {layer: (w_mu, w_sigma)} in the Keras-3 `.weights.h5` layout of the reference model (attribute-named
    layer groups, `vars/0` = w_mu HWIO, `vars/1` = raw w_sigma), so a checkpoint trained here loads with the
    reference's `UNET_model.load_weights` (Brats.py:622)."""
    from . import h5min
    ds = {}
    for n in conv_names:
        w, s = weights[n]
        ds[f"/{n}/vars/0"] = w.detach().cpu().numpy().astype(np.float32)
        ds[f"/{n}/vars/1"] = s.detach().cpu().numpy().astype(np.float32)
    h5min.write_h5(path, ds)


def model_weight_shapes(model) -> Dict[str, Tuple[int, ...]]:
    return {n: tuple(getattr(model, n).weights()[0].shape) for n in model.conv_names}


# ---- result pickle of testing() / the adversarial test ------------------------------------------------------
def save_result_pickle(path: str, probs: Tensor, sigma: Tensor, x: Tensor, y: Tensor, out_hw: Tuple[int, int]) -> None:
    """`pickle.dump([logits_, sigma_, x, y], pf)` (Brats.py:1375,1427): the reference stores the softmax maps and the
    variance maps reshaped to images [N, h, w, C] (Brats.py:1290-1298), then the (noisy) inputs and the labels, as
    NumPy arrays in a plain list."""
    h, w = out_hw
    n, _, c = probs.shape
    arrs = [probs.detach().cpu().numpy().reshape(n, h, w, c), sigma.detach().cpu().numpy().reshape(n, h, w, c),
            np.asarray(x.detach().cpu() if torch.is_tensor(x) else x), np.asarray(y.detach().cpu() if torch.is_tensor(y) else y)]
    with open(path, "wb") as pf:
        pickle.dump(arrs, pf)


def load_result_pickle(path: str):
    """-> [logits_, sigma_, x, y] as the reference's plotting helpers read them back (Brats_functions.py:23-129)."""
    with open(path, "rb") as pf:
        logits_, sigma_, x, y = pickle.load(pf)
    return logits_, sigma_, x, y
