"""CPU tests of the host-side pieces next to the path: segmentation metrics restated from
Brats_functions.py:372-484 against a direct NumPy transcription of the same formulas, batch feeding
(Brats_functions.py:549-556, Brats.py:680-683) and the weight archive (names / layouts of Brats.py:54-63,107-116)."""
import os

import numpy as np
import pytest
import torch

import supernet_b200 as S
from supernet_b200 import dataio, metrics


def _np_ratio(num, den):
    with np.errstate(divide="ignore", invalid="ignore"):
        x = np.divide(num, den)
    return np.mean(x[np.logical_not(np.isnan(x))])


def _np_sensitivity(t, p):          # Brats_functions.py:372-385
    return _np_ratio(np.sum(t * p, axis=(1, 2)), np.sum(t, axis=(1, 2)))


def _np_precision(t, p):            # Brats_functions.py:387-397
    return _np_ratio(np.sum(t * p, axis=(1, 2)), np.sum(p, axis=(1, 2)))


def _np_specificity(t, p):          # Brats_functions.py:426-441
    tn = np.where((t == 0) & (p == 0), 1.0, 0.0).sum(axis=(1, 2))
    den = np.where(t == 0, 1.0, 0.0).sum(axis=(1, 2))
    return _np_ratio(tn, den)


def _np_dice(t, p):                 # Brats_functions.py:400-414
    with np.errstate(divide="ignore", invalid="ignore"):
        c = 2.0 * np.sum(t * p, axis=(1, 2)) / (np.sum(t, axis=(1, 2)) + np.sum(p, axis=(1, 2)))
    cm = np.ma.masked_invalid(c)
    return float(np.mean(cm)), cm


def _labels(seed, B=5, hw=24):
    g = np.random.default_rng(seed)
    y = g.choice([0, 1, 2, 4], size=(B, hw, hw), p=[0.7, 0.1, 0.1, 0.1])
    y[0] = 0                                   # an image with no tumour: NaN ratios must be dropped
    return y


def test_metrics_match_the_reference_formulas():
    yt, yp = _labels(1), _labels(2)
    yp[0, 3:6, 3:6] = 4
    regions_t, regions_p = metrics.region_masks(torch.tensor(yt)), metrics.region_masks(torch.tensor(yp))
    np_regions = {"tumor": lambda a: (a > 0).astype(np.float64),
                  "core": lambda a: np.where((a > 0) & (a != 2), 1.0, 0.0),      # mask_core :455-469
                  "enh": lambda a: (a == 4).astype(np.float64)}                  # mask_enh :470-484
    for name, fn in np_regions.items():
        a, b = fn(yt), fn(yp)
        ta, tb = regions_t[name], regions_p[name]
        assert np.array_equal(ta.numpy(), a) and np.array_equal(tb.numpy(), b)
        assert float(metrics.sensitivity(ta, tb)) == pytest.approx(_np_sensitivity(a, b), rel=1e-12)
        assert float(metrics.precision(ta, tb)) == pytest.approx(_np_precision(a, b), rel=1e-12)
        assert float(metrics.specificity(ta, tb)) == pytest.approx(_np_specificity(a, b), rel=1e-12)
        d, per = metrics.dice(ta, tb)
        d_ref, per_ref = _np_dice(a, b)
        assert float(d) == pytest.approx(d_ref, rel=1e-12)
        assert np.array_equal(np.isnan(per.numpy()), np.ma.getmaskarray(per_ref))
    rep = metrics.region_report(torch.tensor(yt), torch.tensor(yp), with_hausdorff=True)
    assert set(rep) == {"tumor", "core", "enh"} and all(np.isfinite(r["hausdorff"]) for r in rep.values())


def test_hausdorff_matches_scipy_call_pattern():
    from scipy.spatial.distance import directed_hausdorff
    a = (np.random.default_rng(3).random((2, 10, 10)) > 0.6).astype(np.float64)
    b = (np.random.default_rng(4).random((2, 10, 10)) > 0.6).astype(np.float64)
    ref = np.mean([max(directed_hausdorff(b[i], a[i])[0], directed_hausdorff(a[i], b[i])[0]) for i in range(2)])
    assert metrics.hausdorff(torch.tensor(a), torch.tensor(b)) == pytest.approx(ref)


def test_predictions_to_labels_and_batch_feeding(tmp_path):
    g = np.random.default_rng(5)
    x = g.random((3, 4, 20, 20)).astype(np.float64)          # shard layout: [B,C,H,W]
    y = g.integers(0, 5, size=(3, 20, 20))
    xt, yc, onehot = dataio.batch_from_shard(x, y, out_size=14, n_labels=5)
    assert xt.dtype == torch.float32 and xt.shape == (3, 20, 20, 4)
    assert np.array_equal(xt.numpy(), x.transpose(0, 2, 3, 1).astype("float32"))      # load_pickle :554-555
    assert np.array_equal(yc.numpy(), y[:, 3:17, 3:17])                               # crop_numpy_image :500-516
    assert onehot.shape == (3, 14 * 14, 5) and float(onehot.sum()) == 3 * 14 * 14
    assert np.array_equal(onehot.argmax(-1).reshape(3, 14, 14).numpy(), yc.numpy())
    probs = onehot * 0.9 + 0.02
    assert torch.equal(metrics.predictions_to_labels(probs, 14, 14), yc)
    import pickle
    with open(tmp_path / "shard.pkl", "wb") as f:
        pickle.dump((x, y), f)
    x2, y2 = dataio.load_shard(str(tmp_path / "shard.pkl"))
    assert np.array_equal(x2, x) and np.array_equal(y2, y)


def test_weight_archive_roundtrip(tmp_path):
    m = S.Density_prop_with_pad_UNET(8, 3, variant="hippocampus", in_channels=1)
    m.build_with_input(1, None)
    path = str(tmp_path / "vdp_UNET_model.npz")
    dataio.save_weights(m, path)
    z = np.load(path)
    assert "conv_input/w_mu1" in z and "conv_input/w_sigma1" in z and "conv1/w_mu" in z     # Brats.py:54,59,107,112
    assert z["conv1/w_mu"].shape == (3, 3, 8, 8) and z["conv1/w_sigma"].shape == (8,)        # HWIO, one sigma per filter
    m2 = S.Density_prop_with_pad_UNET(8, 3, variant="hippocampus", in_channels=1)
    dataio.load_weights(m2, path)
    for n in m.conv_names:
        for a, b in zip(getattr(m, n).weights(), getattr(m2, n).weights()):
            assert torch.equal(a, b)
    with pytest.raises((RuntimeError, OSError)):      # no h5py in this image (or, with h5py, no such file)
        dataio.from_keras_h5(os.path.join(str(tmp_path), "missing.weights.h5"), m.conv_names)


# ------------------------------------------------------------------------------------------------------------
# Keras-3 `.weights.h5` interchange (Brats.py:732,611-622,933,1195) and the result pickle (Brats.py:1375,1427)
# ------------------------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _hip_names(n=8):
    from oracle import supernet_oracle as O
    return O, [s.name for s in O.unet_conv_specs("hippocampus", n, 3, 1)]


def test_keras_h5_fixture_loads_into_the_model():
    """The committed checkpoint (tests/golden/make_golden.py) -> from_keras_h5 -> load_weight_dict, every tensor equal
    to the generator's weights; HWIO w_mu / raw w_sigma (Brats.py:54-63,107-116)."""
    O, names = _hip_names()
    path = os.path.join(GOLDEN, "keras3_hippocampus_n8.weights.h5")
    with open(path, "rb") as f:
        assert f.read(8) == b"\x89HDF\r\n\x1a\n"
    model = S.Density_prop_with_pad_UNET(8, 3, variant="hippocampus")
    model.build_with_input(1, "cpu")
    w = dataio.from_keras_h5(path, model.conv_names, dataio.model_weight_shapes(model))
    ref = O.make_weights("hippocampus", 8, 3, 1)
    assert list(w) == names
    for n in names:
        assert torch.equal(w[n][0], ref[n][0]) and torch.equal(w[n][1], ref[n][1])
    model.load_weight_dict(w)
    assert torch.equal(model.up1_conv2x2.w_mu.detach(), ref["up1_conv2x2"][0])
    assert torch.equal(model.conv_input.w_sigma1.detach(), ref["conv_input"][1])


def test_keras_h5_auto_named_layers_use_numeric_suffix_order(tmp_path):
    """groups named my_conv_intermediate, _1, _10, _11, _2 ... sort as text in the wrong order (conv2 / conv3
    have the same shape and would swap silently).  The numeric suffix is the construction order of __init__."""
    from supernet_b200 import h5min
    O, names = _hip_names()
    ref = O.make_weights("hippocampus", 8, 3, 1)
    ds = {}
    k = 0
    for n in names:
        if n == "conv_input":
            g = "my_conv_input"
        else:
            g = "my_conv_intermediate" + (f"_{k}" if k else "")
            k += 1
        ds[f"/layers/{g}/vars/0"] = ref[n][0].numpy()
        ds[f"/layers/{g}/vars/1"] = ref[n][1].numpy()
    ds["/layers/my_max_pooling/vars/0"] = np.zeros(3, dtype=np.float32)       # a layer without a weight pair
    ds["/vars/0"] = np.arange(4, dtype=np.int32)
    path = str(tmp_path / "auto.weights.h5")
    h5min.write_h5(path, ds)
    assert sorted(f"my_conv_intermediate_{i}" for i in (1, 2, 10, 11))[1] == "my_conv_intermediate_10"   # the trap
    w = dataio.from_keras_h5(path, names)
    for n in names:
        assert torch.equal(w[n][0], ref[n][0]) and torch.equal(w[n][1], ref[n][1]), n
    # a shape that does not fit the model is an error, not a silent mis-load
    bad = dict(ds)
    bad["/layers/my_conv_intermediate_3/vars/0"] = np.zeros((3, 3, 8, 8), dtype=np.float32)
    h5min.write_h5(str(tmp_path / "bad.weights.h5"), bad)
    model = S.Density_prop_with_pad_UNET(8, 3, variant="hippocampus")
    model.build_with_input(1, "cpu")
    with pytest.raises(RuntimeError):
        dataio.from_keras_h5(str(tmp_path / "bad.weights.h5"), names, dataio.model_weight_shapes(model))
    with pytest.raises(RuntimeError):                      # a layer group is missing
        dataio.from_keras_h5(path, names + ["conv_extra"])


def test_h5min_round_trips_groups_dtypes_and_rejects_what_it_does_not_parse(tmp_path):
    from supernet_b200 import h5min
    g = np.random.default_rng(5)
    ds = {f"/g{i:02d}/deep/er/x": g.random((i + 1, 3)).astype(np.float32) for i in range(21)}    # > 8 and > 16 links
    ds["/f64"] = g.random((2, 3, 4))
    ds["/i64"] = g.integers(-5, 5, size=7)
    ds["/u8"] = g.integers(0, 255, size=(3, 3)).astype(np.uint8)
    ds["/scalar"] = np.float32(2.5)
    ds["/empty"] = np.zeros((0, 4), dtype=np.float32)
    path = str(tmp_path / "t.h5")
    h5min.write_h5(path, ds)
    back = h5min.read_h5(path)
    assert sorted(back) == sorted(ds)
    for k_, v in ds.items():
        assert back[k_].dtype == np.asarray(v).dtype and np.array_equal(back[k_], np.asarray(v)), k_
    raw = bytearray(open(path, "rb").read())
    with pytest.raises(h5min.H5FormatError):
        h5min.H5Reader(bytes(raw[:4]) + b"XXXX" + bytes(raw[8:]))             # signature
    v2 = bytearray(raw)
    v2[8] = 2
    with pytest.raises(h5min.H5FormatError):
        h5min.H5Reader(bytes(v2))                                             # superblock version 2
    # flip the layout class of one dataset to "chunked": the reader must refuse, not return garbage
    i = raw.find(b"\x08\x00\x18\x00")                                      # a layout message header (type 8, 24 bytes)
    assert i > 0 and raw[i + 8] == 3 and raw[i + 9] == 1
    raw[i + 9] = 2
    with pytest.raises(h5min.H5FormatError):
        h5min.H5Reader(bytes(raw)).datasets()


def test_result_pickle_has_the_reference_structure(tmp_path):
    """pickle.dump([logits_, sigma_, x, y]) with image-shaped maps (Brats.py:1290-1298,1375,1427)."""
    import pickle
    p = torch.rand(3, 6 * 5, 4)
    v = torch.rand(3, 6 * 5, 4)
    x = torch.rand(3, 10, 9, 2)
    y = torch.randint(0, 4, (3, 6, 5))
    path = str(tmp_path / "uncertainty_info.pkl")
    dataio.save_result_pickle(path, p, v, x, y, (6, 5))
    with open(path, "rb") as f:
        obj = pickle.load(f)
    assert isinstance(obj, list) and len(obj) == 4 and all(isinstance(a, np.ndarray) for a in obj)
    logits_, sigma_, xb, yb = dataio.load_result_pickle(path)
    assert logits_.shape == (3, 6, 5, 4) and sigma_.shape == (3, 6, 5, 4)
    assert np.array_equal(logits_.reshape(3, -1, 4), p.numpy()) and np.array_equal(xb, x.numpy())
    assert np.array_equal(yb, y.numpy())
