"""ctypes binding of libsupernet_b200.so (include/supernet.h).

This is the only way the Python host reaches the kernels: raw device pointers, sizes and the
current CUDA stream cross the boundary; no torch types do.  There is no CPU fallback: if the
library is missing, or a call returns a negative status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List, Optional

from . import build as _build

_LIB: Optional[C.CDLL] = None


class sn_conv_desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "in_h", "in_w", "cin", "cout", "ksize", "flags", "reserved")]


class sn_window(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "h", "w", "c",
        "src_h", "src_w", "src_c", "src_y0", "src_x0", "src_c0",
        "dst_h", "dst_w", "dst_c", "dst_y0", "dst_x0", "dst_c0",
        "dst_step", "src_step")]


class sn_packed_view(C.Structure):
    _fields_ = [("base", C.c_void_p),
                ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("y0", C.c_int32), ("x0", C.c_int32), ("c0", C.c_int32), ("reserved", C.c_int32)]


class sn_tc_conv_desc(C.Structure):
    _fields_ = [("src", sn_packed_view * 2), ("src_c", C.c_int32 * 2),
                ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("ksize", C.c_int32),
                ("cout", C.c_int32), ("flags", C.c_int32),
                ("w_packed", C.c_void_p), ("s", C.c_void_p),
                ("dst", sn_packed_view), ("dst_mu", C.c_void_p), ("dst_var", C.c_void_p), ("rsum_out", C.c_void_p)]


class sn_tc_head_desc(C.Structure):
    _fields_ = [("n_labels", C.c_int32), ("w_mu", C.c_void_p), ("w_sigma", C.c_void_p), ("p_out", C.c_void_p),
                ("var_out", C.c_void_p), ("presoftmax_mu", C.c_void_p), ("presoftmax_var", C.c_void_p)]


class sn_tc_dgrad_desc(C.Structure):
    _fields_ = [("g_out", sn_packed_view), ("in_", sn_packed_view * 2), ("g_in", sn_packed_view * 2),
                ("in_c", C.c_int32 * 2), ("gate", C.c_int32 * 2),
                ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("ksize", C.c_int32),
                ("cout", C.c_int32), ("flags", C.c_int32),
                ("wt_packed", C.c_void_p), ("s", C.c_void_p)]


class sn_tc_wgrad_desc(C.Structure):
    _fields_ = [("g_out", sn_packed_view), ("in_", sn_packed_view * 2), ("in_c", C.c_int32 * 2),
                ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("ksize", C.c_int32),
                ("cout", C.c_int32), ("flags", C.c_int32),
                ("rsum", C.c_void_p), ("w_mu", C.c_void_p), ("w_sigma", C.c_void_p), ("workspace", C.c_void_p),
                ("g_w_mu", C.c_void_p), ("g_w_sigma", C.c_void_p)]


SN_CONV_RELU = 1
SN_TC_RELU, SN_TC_UPCONV, SN_TC_DST_F32, SN_TC_IM2COL, SN_TC_ROWS, SN_TC_EXACT = 1, 2, 4, 8, 16, 32
SN_TC_KWC, SN_TC_NO_KWC, SN_TC_CTA2, SN_TC_NO_CTA2 = 64, 128, 256, 512


def header_path() -> str:
    return os.path.join(_build.INCLUDE, "supernet.h")


def declared_symbols() -> List[str]:
    """Every function include/supernet.h declares (used by the CPU-side export test)."""
    with open(header_path()) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", text)))


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first when the sources changed and nvcc is present) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB_PATH):
        raise RuntimeError(f"{_build.LIB_PATH} is missing: the CUDA extension is required (no CPU fallback)")
    lib = C.CDLL(_build.LIB_PATH)
    lib.sn_last_error.restype = C.c_char_p
    lib.sn_version.restype = C.c_int
    for name in declared_symbols():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        if name not in ("sn_last_error", "sn_packed_bytes", "sn_prepared_weight_bytes", "sn_conv_workspace_bytes",
                        "sn_wgrad_workspace_bytes"):
            fn.restype = C.c_int
    for name in ("sn_packed_bytes", "sn_prepared_weight_bytes", "sn_wgrad_workspace_bytes"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_size_t
    _LIB = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().sn_last_error()
        raise RuntimeError(f"libsupernet_b200 {what} failed ({status}): {msg.decode() if msg else ''}")


def ptr(t) -> C.c_void_p:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def check_device(t) -> None:
    """Every launch goes to torch.cuda.current_stream() of the CURRENT device: a tensor that lives on another GPU would
    be dereferenced there (illegal address, or silent peer access).  The engines enter torch.cuda.device(self.device)
    themselves; direct callers of the ops must do the same."""
    import torch
    if t is not None and t.is_cuda and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"tensor on {t.device}, but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           f"wrap the call in `with torch.cuda.device(tensor.device):`")


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
