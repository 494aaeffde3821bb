"""`torch.library` custom-op registration of the hot path: namespace `gradtts_b200`.

SURVEY 8(b) / BASELINE.json north star: "Python/PyTorch host code calls hand-written sm_100a CUDA kernels through a thin C-ABI
torch custom-op extension".  Every op below is a thin wrapper that hands raw device pointers, sizes and the current CUDA stream to
one `extern "C"` entry point of libgradtts_b200.so (include/gradtts_b200.h).  Registering them makes the path visible to the
dispatcher: FakeTensor / `torch.compile` shape propagation (the `register_fake` bodies), a CUDA-only kernel table (calling an op
with CPU tensors fails in the dispatcher -- there is no CPU implementation to fall back to), and an autograd hook point
(`estimator_vjp` is what `GradLogPEstimator2d.forward`'s backward calls).

    torch.ops.gradtts_b200.reverse_diffusion(handle, z, mask, mu, spk, noise, n_timesteps, flags) -> xt
    torch.ops.gradtts_b200.estimator(handle, x, mask, mu, t, spk, flags)                          -> score   (autograd w.r.t. x)
    torch.ops.gradtts_b200.estimator_vjp(handle, x, mask, mu, t, spk, v, flags)                   -> (score, J^T v)
    torch.ops.gradtts_b200.maximum_path(value, mask)                                               -> (path, status)
    torch.ops.gradtts_b200.log_prior(mu_x, y)                                                      -> log_prior
    torch.ops.gradtts_b200.align_outputs(attn, mu_x, x_mask)                                       -> (logw_, mu_y)
    torch.ops.gradtts_b200.forward_diffusion(x0, mask, mu, t, noise, beta_min, beta_max)           -> (xt, z_masked)
    torch.ops.gradtts_b200.score_loss(est, z_masked, mask, t, beta_min, beta_max)                  -> loss
    torch.ops.gradtts_b200.text_encoder(handle, tokens, lengths, spk, n_feats)                     -> (mu, logw, x_mask)
    torch.ops.gradtts_b200.vocoder(handle, mel, hop, flags)                                        -> audio

`handle` is the integer value of the `gtts_decoder*` owned by the calling module (model/diffusion.py).  All tensors are fp32,
contiguous, on the handle's sm_100 device; the Python modules do the casting and validation before they get here.
"""
import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

NS = "gradtts_b200"


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return t.data_ptr() if t is not None else None


# ---------------------------------------------------------------------------------------------------------------- decoder
@torch.library.custom_op(f"{NS}::reverse_diffusion", mutates_args=(), device_types="cuda")
def reverse_diffusion(handle: int, z: Tensor, mask: Tensor, mu: Tensor, spk: Optional[Tensor], noise: Optional[Tensor],
                      n_timesteps: int, flags: int) -> Tensor:
    """Diffusion.reverse_diffusion (reference model/diffusion.py:254-268) -> gtts_decoder_reverse_diffusion."""
    B, _, T = z.shape
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        rc = _lib.load().gtts_decoder_reverse_diffusion(ctypes.c_void_p(handle), z.data_ptr(), mask.data_ptr(), mu.data_ptr(),
                                                        _ptr(spk), out.data_ptr(), B, T, n_timesteps, flags, _ptr(noise), _stream(z))
    _lib.check(rc, "reverse_diffusion")
    return out


@reverse_diffusion.register_fake
def _(handle, z, mask, mu, spk, noise, n_timesteps, flags):
    return torch.empty_like(z)


@torch.library.custom_op(f"{NS}::estimator", mutates_args=(), device_types="cuda")
def estimator(handle: int, x: Tensor, mask: Tensor, mu: Tensor, t: Tensor, spk: Optional[Tensor], flags: int) -> Tensor:
    """GradLogPEstimator2d.forward (reference model/diffusion.py:174-216) -> gtts_decoder_estimator."""
    B, _, T = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().gtts_decoder_estimator(ctypes.c_void_p(handle), x.data_ptr(), mask.data_ptr(), mu.data_ptr(), t.data_ptr(),
                                                _ptr(spk), out.data_ptr(), B, T, flags, _stream(x))
    _lib.check(rc, "estimator")
    return out


@estimator.register_fake
def _(handle, x, mask, mu, t, spk, flags):
    return torch.empty_like(x)


@torch.library.custom_op(f"{NS}::estimator_vjp", mutates_args=(), device_types="cuda")
def estimator_vjp(handle: int, x: Tensor, mask: Tensor, mu: Tensor, t: Tensor, spk: Optional[Tensor], v: Tensor,
                  flags: int) -> Tuple[Tensor, Tensor]:
    """(score, (d score / d x)^T v) in one pass -> gtts_decoder_estimator_vjp.  The gradient torch.autograd.grad takes through the
    score network in the Hutchinson divergence (reference n_best/likelihood/likelihood.py:27-38)."""
    B, _, T = x.shape
    score, gx = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().gtts_decoder_estimator_vjp(ctypes.c_void_p(handle), x.data_ptr(), mask.data_ptr(), mu.data_ptr(), t.data_ptr(),
                                                    _ptr(spk), v.data_ptr(), score.data_ptr(), gx.data_ptr(), B, T, flags, _stream(x))
    _lib.check(rc, "estimator_vjp")
    return score, gx


@estimator_vjp.register_fake
def _(handle, x, mask, mu, t, spk, v, flags):
    return torch.empty_like(x), torch.empty_like(x)


# autograd for `estimator` w.r.t. x only (the reference differentiates the score model w.r.t. its input in the likelihood code; the
# parameter gradients of training are not built).  The backward re-runs the forward inside gtts_decoder_estimator_vjp: nothing is
# kept alive between the two calls.
def _estimator_setup(ctx, inputs, output):
    handle, x, mask, mu, t, spk, flags = inputs
    ctx.handle, ctx.flags = handle, flags
    ctx.save_for_backward(x, mask, mu, t, *([spk] if spk is not None else []))
    ctx.has_spk = spk is not None


def _estimator_backward(ctx, grad_out):
    saved = ctx.saved_tensors
    x, mask, mu, t = saved[:4]
    spk = saved[4] if ctx.has_spk else None
    _, gx = torch.ops.gradtts_b200.estimator_vjp(ctx.handle, x, mask, mu, t, spk, grad_out.contiguous().to(torch.float32), ctx.flags)
    return None, gx, None, None, None, None, None


estimator.register_autograd(_estimator_backward, setup_context=_estimator_setup)


# ---------------------------------------------------------------------------------------------------------------- MAS + alignment
@torch.library.custom_op(f"{NS}::maximum_path", mutates_args=(), device_types="cuda")
def maximum_path(value: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor]:
    """monotonic_align.maximum_path (reference model/monotonic_align/__init__.py:8-23) -> gtts_mas_maximum_path.
    Returns (path fp32 {0,1}, status int32[1]: 1 if an item had t_x > t_y, the reference's undefined case)."""
    lib = _lib.load()
    b, tx, ty = value.shape
    path = torch.empty_like(value)
    status = torch.zeros(1, dtype=torch.int32, device=value.device)
    if b == 0 or tx == 0 or ty == 0:
        return path.zero_(), status
    ws_bytes = lib.gtts_mas_workspace_bytes(b, tx, ty)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=value.device) if ws_bytes else None
    with torch.cuda.device(value.device):
        rc = lib.gtts_mas_maximum_path(value.data_ptr(), mask.data_ptr(), path.data_ptr(), b, tx, ty, _ptr(ws), ws_bytes,
                                       status.data_ptr(), _stream(value))
    _lib.check(rc, "maximum_path")
    return path, status


@maximum_path.register_fake
def _(value, mask):
    return torch.empty_like(value), value.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op(f"{NS}::log_prior", mutates_args=(), device_types="cuda")
def log_prior(mu_x: Tensor, y: Tensor) -> Tensor:
    """log N(y; mu_x, I) for every (text row, mel frame) pair (reference model/tts.py:143-149) -> gtts_align_log_prior."""
    B, C, tx = mu_x.shape
    ty = y.shape[2]
    out = torch.empty(B, tx, ty, dtype=torch.float32, device=mu_x.device)
    with torch.cuda.device(mu_x.device):
        rc = _lib.load().gtts_align_log_prior(mu_x.data_ptr(), y.data_ptr(), out.data_ptr(), B, C, tx, ty, _stream(mu_x))
    _lib.check(rc, "log_prior")
    return out


@log_prior.register_fake
def _(mu_x, y):
    return mu_x.new_empty((mu_x.shape[0], mu_x.shape[2], y.shape[2]))


@torch.library.custom_op(f"{NS}::align_outputs", mutates_args=(), device_types="cuda")
def align_outputs(attn: Tensor, mu_x: Tensor, x_mask: Optional[Tensor], want_mu_y: bool) -> Tuple[Tensor, Tensor]:
    """logw_ = log(1e-8 + sum_j attn) * x_mask and mu_y = attn^T mu_x (reference model/tts.py:155,184-185) -> gtts_align_outputs.
    logw_ is empty (0 elements) when x_mask is None, mu_y when want_mu_y is False."""
    B, tx, ty = attn.shape
    C = mu_x.shape[1]
    logw = torch.empty((B, 1, tx) if x_mask is not None else (0,), dtype=torch.float32, device=attn.device)
    mu_y = torch.empty((B, C, ty) if want_mu_y else (0,), dtype=torch.float32, device=attn.device)
    with torch.cuda.device(attn.device):
        rc = _lib.load().gtts_align_outputs(attn.data_ptr(), mu_x.data_ptr(), _ptr(x_mask), logw.data_ptr() if x_mask is not None else None,
                                            mu_y.data_ptr() if want_mu_y else None, B, C, tx, ty, _stream(attn))
    _lib.check(rc, "align_outputs")
    return logw, mu_y


@align_outputs.register_fake
def _(attn, mu_x, x_mask, want_mu_y):
    B, tx, ty = attn.shape
    return (attn.new_empty((B, 1, tx) if x_mask is not None else (0,)), attn.new_empty((B, mu_x.shape[1], ty) if want_mu_y else (0,)))


# ---------------------------------------------------------------------------------------------------------------- training objective
@torch.library.custom_op(f"{NS}::forward_diffusion", mutates_args=(), device_types="cuda")
def forward_diffusion(x0: Tensor, mask: Tensor, mu: Tensor, t: Tensor, noise: Tensor, beta_min: float,
                      beta_max: float) -> Tuple[Tensor, Tensor]:
    """Diffusion.forward_diffusion with the caller's N(0,1) draw (reference model/diffusion.py:244-252) -> gtts_forward_diffusion."""
    B, C, T = x0.shape
    xt, zm = torch.empty_like(x0), torch.empty_like(x0)
    with torch.cuda.device(x0.device):
        rc = _lib.load().gtts_forward_diffusion(x0.data_ptr(), mask.data_ptr(), mu.data_ptr(), t.data_ptr(), noise.data_ptr(),
                                                xt.data_ptr(), zm.data_ptr(), B, C, T, beta_min, beta_max, _stream(x0))
    _lib.check(rc, "forward_diffusion")
    return xt, zm


@forward_diffusion.register_fake
def _(x0, mask, mu, t, noise, beta_min, beta_max):
    return torch.empty_like(x0), torch.empty_like(x0)


@torch.library.custom_op(f"{NS}::score_loss", mutates_args=(), device_types="cuda")
def score_loss(est: Tensor, z_masked: Tensor, mask: Tensor, t: Tensor, beta_min: float, beta_max: float) -> Tensor:
    """The scalar of Diffusion.loss_t (reference model/diffusion.py:276-280) -> gtts_score_loss."""
    lib = _lib.load()
    B, C, T = est.shape
    ws = torch.empty(int(lib.gtts_score_loss_workspace_bytes()), dtype=torch.uint8, device=est.device)
    loss = torch.empty((), dtype=torch.float32, device=est.device)
    with torch.cuda.device(est.device):
        rc = lib.gtts_score_loss(est.data_ptr(), z_masked.data_ptr(), mask.data_ptr(), t.data_ptr(), ws.data_ptr(), ws.numel(),
                                 loss.data_ptr(), B, C, T, beta_min, beta_max, _stream(est))
    _lib.check(rc, "score_loss")
    return loss


@score_loss.register_fake
def _(est, z_masked, mask, t, beta_min, beta_max):
    return est.new_empty(())


# ---------------------------------------------------------------------------------------------------------------- text encoder, vocoder
@torch.library.custom_op(f"{NS}::text_encoder", mutates_args=(), device_types="cuda")
def text_encoder(handle: int, tokens: Tensor, lengths: Tensor, spk: Optional[Tensor], n_feats: int) -> Tuple[Tensor, Tensor, Tensor]:
    """TextEncoder.forward (reference model/text_encoder.py:321-335) -> gtts_encoder_forward.  tokens (B, T) int64, lengths (B)
    int64 -> mu (B, n_feats, T), logw (B, 1, T), x_mask (B, 1, T)."""
    B, T = tokens.shape
    mu = torch.empty(B, n_feats, T, dtype=torch.float32, device=tokens.device)
    logw = torch.empty(B, 1, T, dtype=torch.float32, device=tokens.device)
    x_mask = torch.empty(B, 1, T, dtype=torch.float32, device=tokens.device)
    lib = _lib.load()
    with torch.cuda.device(tokens.device):
        rc = lib.gtts_encoder_forward(ctypes.c_void_p(handle), tokens.data_ptr(), lengths.data_ptr(), _ptr(spk), mu.data_ptr(),
                                      logw.data_ptr(), x_mask.data_ptr(), B, T, _stream(tokens))
        _lib.check(rc, "encoder_forward")
        _lib.check(lib.gtts_encoder_check_tokens(ctypes.c_void_p(handle), _stream(tokens)), "encoder_forward")
    return mu, logw, x_mask


@text_encoder.register_fake
def _(handle, tokens, lengths, spk, n_feats):
    B, T = tokens.shape
    f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=tokens.device)   # noqa: E731
    return f(B, n_feats, T), f(B, 1, T), f(B, 1, T)


@torch.library.custom_op(f"{NS}::vocoder", mutates_args=(), device_types="cuda")
def vocoder(handle: int, mel: Tensor, hop: int, flags: int) -> Tensor:
    """HiFi-GAN Generator.forward (reference hifi-gan/models.py:101-118) -> gtts_vocoder_forward.  mel (B, 80, T) -> (B, 1, T * hop)."""
    B, _, T = mel.shape
    out = torch.empty(B, 1, T * hop, dtype=torch.float32, device=mel.device)
    with torch.cuda.device(mel.device):
        rc = _lib.load().gtts_vocoder_forward(ctypes.c_void_p(handle), mel.data_ptr(), out.data_ptr(), B, T, flags, _stream(mel))
    _lib.check(rc, "vocoder_forward")
    return out


@vocoder.register_fake
def _(handle, mel, hop, flags):
    B, _, T = mel.shape
    return torch.empty(B, 1, T * hop, dtype=torch.float32, device=mel.device)


OPS = ("reverse_diffusion", "estimator", "estimator_vjp", "maximum_path", "log_prior", "align_outputs", "forward_diffusion", "score_loss",
       "text_encoder", "vocoder")
