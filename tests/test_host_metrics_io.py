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
