"""monotonic_align.maximum_path on device (replaces reference model/monotonic_align/__init__.py:8-23).

The reference multiplies value*mask, copies both to the host, derives the lengths from mask sums, runs a
serial Cython DP per utterance and copies the path back.  Here the whole thing is one kernel launch per
batch (csrc/mas.cu) on the tensors where they already live; the result is bit-identical.
"""
import ctypes

import torch

from .. import utils  # noqa: F401  (keeps `model.utils` importable next to this module, like the reference)
from ... import _lib, ops  # noqa: F401  (ops registers torch.ops.gradtts_b200.*)


def maximum_path(value, mask, check=True):
    """value: [b, t_x, t_y], mask: [b, t_x, t_y] -> path of value.dtype on value.device, entries {0,1}.

    `check=True` reads back one status word (a stream sync, which the reference's `.cpu()` also implies)
    and raises if any item has t_x > t_y, the case the reference leaves undefined.
    """
    _lib.require_cuda_tensor(value, "value")
    _lib.require_cuda_tensor(mask, "mask")
    if value.dim() != 3 or mask.shape != value.shape:
        raise ValueError(f"maximum_path expects value and mask of the same [b, t_x, t_y] shape, got "
                         f"{tuple(value.shape)} and {tuple(mask.shape)}")
    dtype = value.dtype
    v = value.detach().to(torch.float32).contiguous()
    m = mask.detach().to(torch.float32).contiguous()
    b, tx, ty = v.shape
    if b == 0 or tx == 0 or ty == 0:
        return torch.zeros_like(value)
    path, status = torch.ops.gradtts_b200.maximum_path(v, m)
    if check and int(status.item()) != 0:
        raise RuntimeError("maximum_path: an item has t_x > t_y (more text rows than mel frames); "
                           "the reference's result is undefined for that input")
    return path.to(dtype)


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9):
    """Device twin of the Cython entry point (core.pyx:40): int32 `paths` is filled in place.

    Unlike the reference, `values` is not modified (the DP table never leaves shared memory)."""
    for t, n in ((paths, "paths"), (values, "values"), (t_xs, "t_xs"), (t_ys, "t_ys")):
        _lib.require_cuda_tensor(t, n)
    if paths.dtype != torch.int32 or values.dtype != torch.float32 or t_xs.dtype != torch.int32 \
            or t_ys.dtype != torch.int32:
        raise TypeError("maximum_path_c expects int32 paths/t_xs/t_ys and float32 values")
    if not (paths.is_contiguous() and values.is_contiguous() and t_xs.is_contiguous() and t_ys.is_contiguous()):
        raise ValueError("maximum_path_c expects C-contiguous tensors (int[:,:,::1] in the reference)")
    lib = _lib.load()
    b, tx, ty = values.shape
    status = torch.empty(1, dtype=torch.int32, device=values.device)
    ws_bytes = lib.gtts_mas_workspace_bytes(b, tx, ty)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=values.device) if ws_bytes else None
    with torch.cuda.device(values.device):
        stream = torch.cuda.current_stream(values.device).cuda_stream
        rc = lib.gtts_mas_maximum_path_c(paths.data_ptr(), values.data_ptr(), t_xs.data_ptr(), t_ys.data_ptr(),
                                         b, tx, ty, float(max_neg_val),
                                         ws.data_ptr() if ws is not None else None, ws_bytes,
                                         status.data_ptr(), ctypes.c_void_p(stream))
    _lib.check(rc, "maximum_path_c")
    return status
