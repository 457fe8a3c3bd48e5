"""Regenerates the golden fixtures (the TensorFlow reference cannot run in this image -- "parity unpinned", see
oracle/supernet_oracle.py): two from the fp64 torch oracle, two from the independent NumPy restatement
(oracle/numpy_check.py) that the torch oracle is checked against, and one Keras-3 checkpoint.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import supernet_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    x = O.make_input("hippocampus", 2)
    p, v, mf, sf = O.UNetOracle("hippocampus", 32, 3, 1, torch.float64)(x, True)
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(p.numel(), 512, replace=False))
    np.savez_compressed(
        os.path.join(HERE, "hippocampus_b2_fp64.npz"), idx=idx,
        p=p.flatten()[idx].numpy(), v=v.flatten()[idx].numpy(),
        mf=mf.flatten()[idx].numpy(), sf=sf.flatten()[idx].numpy(),
        p_sum=float(p.sum()), v_sum=float(v.sum()))
    g = torch.Generator().manual_seed(42)
    mu = torch.randn(1, 6, 6, 4, generator=g, dtype=torch.float64)
    var = torch.rand(1, 6, 6, 4, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 4, 8, generator=g, dtype=torch.float64) * 0.1
    ws = torch.empty(8, dtype=torch.float64).uniform_(-8, -2, generator=g)
    m, vv = O.conv_intermediate_as_written(mu, var, w, ws)
    np.savez_compressed(os.path.join(HERE, "layers_fp64.npz"), mu=mu.numpy(), var=var.numpy(), w=w.numpy(),
                        ws=ws.numpy(), m_out=m.numpy(), v_out=vv.numpy())
    # the SAME networks through the independent NumPy restatement (oracle/numpy_check.py: no torch arithmetic): these
    # two files are what pins the torch oracle (tests/test_oracle.py)
    from oracle import numpy_check as N
    for variant, C, in_ch, alpha, fname in (("hippocampus", 3, 1, 1.0, "hippocampus_b2_numpy_fp64.npz"),
                                            ("brats", 4, 4, O.BRATS_ALPHA, "brats_b1_numpy_fp64.npz")):
        Wn = {k: (a.double().numpy(), b.double().numpy()) for k, (a, b) in O.make_weights(variant, 32, C, in_ch).items()}
        xn = O.make_input(variant, 2 if variant == "hippocampus" else 1, alpha=alpha).double().numpy()
        pn, vn, mfn, sfn = N.unet_forward(xn, Wn, variant)
        idx = np.sort(np.random.default_rng(1).choice(pn.size, 2048, replace=False))
        np.savez_compressed(os.path.join(HERE, fname), idx=idx, p=pn.reshape(-1)[idx], v=vn.reshape(-1)[idx],
                            mf=mfn.reshape(-1)[idx], sf=sfn.reshape(-1)[idx], p_sum=float(pn.sum()),
                            v_sum=float(vn.sum()), mf_abs_sum=float(np.abs(mfn).sum()), sf_sum=float(sfn.sum()))
    # a Keras-3 `.weights.h5` checkpoint in the reference model's layout (Brats.py:732): Hippocampus graph at
    # n_kernels = 8 (147 KB), attribute-named layer groups, written by the in-repo HDF5 writer (no h5py here)
    import supernet_b200  # noqa: F401
    from supernet_b200 import dataio
    names = [s.name for s in O.unet_conv_specs("hippocampus", 8, 3, 1)]
    dataio.to_keras_h5(os.path.join(HERE, "keras3_hippocampus_n8.weights.h5"), O.make_weights("hippocampus", 8, 3, 1),
                       names)


if __name__ == "__main__":
    main()
