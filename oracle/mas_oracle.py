"""ctypes wrapper over oracle/mas_oracle.c plus the host wrapper's semantics -- TEST INFRASTRUCTURE ONLY.

`maximum_path(value, mask)` restates /root/reference/model/monotonic_align/__init__.py:8-23:
value*mask, float32 copy, lengths from mask sums, call the C kernel, cast path to value.dtype.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "_build", "libmas_oracle.so")


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            from . import build_oracle
            build_oracle.build_c_oracle()
        _LIB = ctypes.CDLL(lib_path())
        _LIB.mas_oracle_maximum_path_c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_float]
        _LIB.mas_oracle_maximum_path_c.restype = None
    return _LIB


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9):
    """Same contract as core.pyx:40 (numpy int32/float32 C-contiguous arrays, values mutated)."""
    assert paths.dtype == np.int32 and values.dtype == np.float32
    assert paths.flags.c_contiguous and values.flags.c_contiguous
    t_xs = np.ascontiguousarray(t_xs, dtype=np.int32)
    t_ys = np.ascontiguousarray(t_ys, dtype=np.int32)
    b, tx, ty = values.shape
    _lib().mas_oracle_maximum_path_c(paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data,
                                     t_ys.ctypes.data, b, tx, ty, max_neg_val)


def maximum_path_np(value, mask):
    """numpy in / numpy out version of monotonic_align.maximum_path (__init__.py:13-22)."""
    value = (value * mask).astype(np.float32)
    value = np.ascontiguousarray(value)
    path = np.zeros_like(value).astype(np.int32)
    t_x_max = mask.sum(1)[:, 0].astype(np.int32)
    t_y_max = mask.sum(2)[:, 0].astype(np.int32)
    maximum_path_c(path, value, t_x_max, t_y_max)
    return path


def maximum_path(value, mask):
    """torch in / torch out, CPU (__init__.py:8-23)."""
    import torch
    path = maximum_path_np(value.detach().cpu().numpy(), mask.detach().cpu().numpy())
    return torch.from_numpy(path).to(device=value.device, dtype=value.dtype)
