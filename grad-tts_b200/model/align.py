"""Alignment stage around MAS on the device (reference model/tts.py:139-185).

`log_prior` is the tensor `monotonic_align.maximum_path` maximises (tts.py:143-149); `logw_from_path` and `mu_y_from_path` are
the two tensors the training / scoring code derives from the path (tts.py:155, 184-185).  CUDA fp32 tensors in, fresh CUDA
tensors out; CPU tensors raise (no fallback).
"""
import ctypes

import torch

from .. import _lib, ops  # noqa: F401  (ops registers torch.ops.gradtts_b200.*)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def log_prior(mu_x, y):
    """mu_x: (B, n_feats, t_x), y: (B, n_feats, t_y) -> (B, t_x, t_y)"""
    _lib.require_cuda_tensor(mu_x, "mu_x")
    _lib.require_cuda_tensor(y, "y")
    if mu_x.dim() != 3 or y.dim() != 3 or mu_x.shape[:2] != y.shape[:2]:
        raise ValueError(f"log_prior expects (B, n_feats, t_x) and (B, n_feats, t_y), got {tuple(mu_x.shape)} and {tuple(y.shape)}")
    m = mu_x.detach().to(torch.float32).contiguous()
    v = y.detach().to(torch.float32).contiguous()
    return torch.ops.gradtts_b200.log_prior(m, v).to(mu_x.dtype)


def _outputs(attn, mu_x, x_mask, want_logw, want_mu_y):
    _lib.require_cuda_tensor(attn, "attn")
    _lib.require_cuda_tensor(mu_x, "mu_x")
    a = attn.detach().to(torch.float32).contiguous()
    m = mu_x.detach().to(torch.float32).contiguous()
    if a.dim() == 4:
        a = a.squeeze(1)
    B, tx, ty = a.shape
    C = m.shape[1]
    if m.shape[0] != B or m.shape[2] != tx:
        raise ValueError(f"attn {tuple(attn.shape)} and mu_x {tuple(mu_x.shape)} do not match")
    xm = None
    if want_logw:
        _lib.require_cuda_tensor(x_mask, "x_mask")
        xm = x_mask.detach().to(torch.float32).reshape(B, tx).contiguous()
    logw, mu_y = torch.ops.gradtts_b200.align_outputs(a, m, xm, bool(want_mu_y))
    return (logw if want_logw else None), (mu_y if want_mu_y else None)


def logw_from_path(attn, x_mask):
    """attn: (B, t_x, t_y), x_mask: (B, 1, t_x) -> log(1e-8 + durations) * x_mask, (B, 1, t_x)   (tts.py:155)"""
    return _outputs(attn, torch.empty(attn.shape[0], 1, attn.shape[-2], device=attn.device), x_mask, True, False)[0]


def mu_y_from_path(attn, mu_x):
    """attn: (B, t_x, t_y), mu_x: (B, n_feats, t_x) -> (B, n_feats, t_y)   (tts.py:184-185)"""
    return _outputs(attn, mu_x, None, False, True)[1].to(mu_x.dtype)
