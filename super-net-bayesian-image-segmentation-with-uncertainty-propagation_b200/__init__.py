"""supernet_b200: the SUPER-Net moment-propagation hot path on B200 (sm_100a).

Host side = the reference's Keras layer surface as torch modules (layers.py) over PyTorch custom ops
(ops.py) that call hand-written CUDA kernels through the C-ABI of include/supernet.h (_lib.py).
The directory name carries the reference's name and is not a valid identifier; import it through the
`supernet_b200` shim at the repo root.
"""
from . import _lib, build, ops  # noqa: F401
from .fastlayers import PackedMoments  # noqa: F401
from .layers import (Density_prop_with_pad_UNET, create_adversarial_pattern, fast_mode, myConc,  # noqa: F401
                     myConv_input, myConv_intermediate, mymaxpooling, mypadding, myReLU, mysoftmax, myupsampling,
                     nll_gaussian, set_default_mode, sigma_regularizer)



def available_modes():
    """Execution modes this build of the library provides: 'fp32' (CUDA-core, autograd) and, when the
    tcgen05 entry points are exported, 'fast' (tensor-core inference pipeline)."""
    lib = _lib.load()
    return ["fp32"] + (["fast"] if hasattr(lib, "sn_conv_moments_fwd_tc") else [])


__all__ = ["available_modes", "fast_mode", "set_default_mode", "PackedMoments", "Density_prop_with_pad_UNET",
           "create_adversarial_pattern", "myConc", "myConv_input",
           "myConv_intermediate", "mymaxpooling", "mypadding", "myReLU", "mysoftmax", "myupsampling",
           "nll_gaussian", "sigma_regularizer", "ops", "build"]
