"""Drop-in `Diffusion` / `GradLogPEstimator2d` backed by the sm_100a kernels (reference model/diffusion.py).

The modules keep the reference's constructor arguments, attribute names and `state_dict()` keys
(e.g. `estimator.downs.0.0.block1.block.0.weight`), so a reference checkpoint loads unchanged; the forward
passes hand raw device pointers to the C ABI (include/gradtts_b200.h).  Nothing here computes on the CPU
or through ATen: if the parameters are not on an sm_100 CUDA device the call raises.
"""
import ctypes
import math

import torch

from .. import _lib, ops, synth  # noqa: F401  (ops registers torch.ops.gradtts_b200.*)
from .base import BaseModule


class _Node(torch.nn.Module):
    """Anonymous container used to reproduce the reference's nested parameter names."""


def _build_param_tree(root, names_shapes, init):
    for name, shape in names_shapes:
        parts = name.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        node.register_parameter(parts[-1], torch.nn.Parameter(init(name, shape)))


def _default_init(name, shape):
    """PyTorch-default-like init from the global RNG: U(+-1/sqrt(fan_in)); GroupNorm 1/0; Rezero g = 0."""
    if name.endswith(".fn.g"):
        return torch.zeros(shape)                                     # model/diffusion.py:43
    if ".block.1.weight" in name:
        return torch.ones(shape)
    if ".block.1.bias" in name:
        return torch.zeros(shape)
    if len(shape) > 1:
        fan_in = math.prod(shape[1:])
        if ".3.conv.weight" in name and name.startswith("ups"):
            fan_in = shape[1] * shape[2] * shape[3]
    else:
        fan_in = _default_init.last_fan_in
    _default_init.last_fan_in = fan_in
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape) * 2.0 - 1.0) * bound


_default_init.last_fan_in = 1


_RESNET_ORDER = ("downs.0.0", "downs.0.1", "downs.1.0", "downs.1.1", "downs.2.0", "downs.2.1",
                 "mid_block1", "mid_block2", "ups.0.0", "ups.0.1", "ups.1.0", "ups.1.1")      # rows of the 1792 time biases


def _mish(x):
    return x * torch.tanh(torch.nn.functional.softplus(x))


def _is_host_side(name):
    """Parameters whose gradients PyTorch computes on the host side of the op (tiny MLPs): time MLP, per-block Linear, spk_mlp."""
    return ".mlp." in name or name.startswith("mlp.") or name.startswith("spk_mlp.")


class _EstimatorTrainFn(torch.autograd.Function):
    """Training-mode estimator call: forward = gtts_decoder_estimator, backward = gtts_decoder_estimator_backward.

    Inputs after the module: x, mask, mu, t, spk (raw), tb (B,1792) and splane (B,80)|None -- the per-block time biases and the
    speaker plane recomputed with PyTorch ops ONLY to tie the tiny time / speaker MLPs into the autograd graph (the device code
    computes its own copies for the forward value) -- then every convolution / GroupNorm / attention parameter."""

    @staticmethod
    def forward(ctx, est, x, mask, mu, t, spk, tb, splane, *params):
        h = est._get_handle()
        ctx.est, ctx.flags = est, est._flags()
        ctx.save_for_backward(x, mask, mu, t, *([spk] if spk is not None else []))
        ctx.has_spk, ctx.has_splane = spk is not None, splane is not None
        return torch.ops.gradtts_b200.estimator(int(h.value), x, mask, mu, t, spk, ctx.flags)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        est = ctx.est
        saved = ctx.saved_tensors
        x, mask, mu, t = saved[:4]
        spk = saved[4] if ctx.has_spk else None
        h = est._get_handle()
        lib = _lib.load()
        B, _, T = x.shape
        g = g.contiguous().to(torch.float32)
        gx, gmu, gs = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        gtb = torch.empty(B, 1792, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            rc = lib.gtts_decoder_estimator_backward(h, x.data_ptr(), mask.data_ptr(), mu.data_ptr(), t.data_ptr(),
                                                     spk.data_ptr() if spk is not None else None, g.data_ptr(), None, gx.data_ptr(),
                                                     gmu.data_ptr(), gs.data_ptr(), gtb.data_ptr(), B, T, ctx.flags, stream)
            _lib.check(rc, "estimator_backward")
            # every parameter gradient in one device copy; the per-parameter tensors are views of it (172 small copies cost
            # ~0.5 ms of a 11 ms step)
            slots = est._grad_slots(h)
            flat = torch.empty(slots["total"], dtype=torch.float32, device=x.device)
            _lib.check(lib.gtts_decoder_get_param_grads_flat(h, flat.data_ptr(), flat.numel(), stream), "get_param_grads_flat")
            pgrads = []
            for name, p in est._device_params():
                off, n = slots[name]
                pgrads.append(flat[off:off + n].view(p.shape).to(p.dtype))
        gsplane = gs.sum(-1) if ctx.has_splane else None
        return (None, gx, None, gmu, None, None, gtb, gsplane, *pgrads)


class GradLogPEstimator2d(BaseModule):
    """Score U-Net (reference model/diffusion.py:128-216). Parameters only; the maths lives in csrc/."""

    def __init__(self, dim, dim_mults=(1, 2, 4), groups=8, n_spks=None, spk_emb_dim=64, n_feats=80, pe_scale=1000):
        super().__init__()
        if dim != 64 or tuple(dim_mults) != (1, 2, 4) or groups != 8 or spk_emb_dim != 64 or n_feats != 80:
            raise NotImplementedError("the sm_100a kernels are specialised for the reference configuration "
                                      "dim=64, dim_mults=(1,2,4), groups=8, spk_emb_dim=64, n_feats=80")
        self.dim = dim
        self.dim_mults = dim_mults
        self.groups = groups
        self.n_spks = n_spks if n_spks is not None else 1
        self.spk_emb_dim = spk_emb_dim
        self.n_feats = n_feats
        self.pe_scale = pe_scale
        self.beta_min, self.beta_max = 0.05, 20.0          # overwritten by the owning Diffusion
        self.precision = "bf16"                            # "bf16" (tcgen05 convs) or "fp32" (true-fp32 arithmetic)
        self.max_chunk = 64                                # samples per workspace chunk
        _build_param_tree(self, synth.decoder_param_shapes(self.n_spks, pfx=""), _default_init)
        self._handle = None
        self._handle_key = None
        self._uploaded = None
        self._plist = None
        self._opts = {}

    # ---- handle management -------------------------------------------------------------------------
    def _grad_slots(self, h):
        """(offset, numel) of every device-side parameter gradient in the handle's flat buffer (fixed per handle)"""
        if getattr(self, "_slots", None) is None or self._slots.get("handle") != h.value:
            lib = _lib.load()
            off, n = ctypes.c_size_t(), ctypes.c_size_t()
            _lib.check(lib.gtts_decoder_param_grad_slot(h, None, ctypes.byref(off), ctypes.byref(n)), "param_grad_slot")
            slots = {"handle": h.value, "total": int(n.value)}
            for name, _ in self._device_params():
                _lib.check(lib.gtts_decoder_param_grad_slot(h, ("estimator." + name).encode(), ctypes.byref(off), ctypes.byref(n)),
                           f"param_grad_slot({name})")
                slots[name] = (int(off.value), int(n.value))
            self._slots = slots
        return self._slots

    def _param_signature(self):
        # flat list cached: walking the module tree (176 tensors) on every forward cost more than the C call at batch 1
        if self._plist is None:
            self._plist = list(self.parameters())
        return tuple((p.data_ptr(), p._version) for p in self._plist)

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .float() may replace the Parameter objects
        self._plist = None
        return super()._apply(fn, *a, **k)

    def _get_handle(self):
        lib = _lib.load()
        p0 = next(self.parameters())
        _lib.require_cuda_tensor(p0, "decoder parameters")
        dev = p0.device.index if p0.device.index is not None else torch.cuda.current_device()
        key = (dev, float(self.beta_min), float(self.beta_max), float(self.pe_scale))
        if self._handle is None or self._handle_key != key:
            self._release()
            h = ctypes.c_void_p()
            rc = lib.gtts_decoder_create(ctypes.byref(h), int(self.n_spks), int(self.n_feats), int(self.dim),
                                         float(self.beta_min), float(self.beta_max), float(self.pe_scale), dev)
            _lib.check(rc, "decoder_create")
            self._handle, self._handle_key, self._uploaded, self._opts = h, key, None, {}
        sig = self._param_signature()
        if self._uploaded != sig:
            # the parameters may have been produced on a side stream (optimizer step, .to() under torch.cuda.stream):
            # the upload below is ordered on the library's legacy stream only, so finish the producer first
            torch.cuda.current_stream(p0.device).synchronize()
            for name, p in self.named_parameters():
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(torch.float32).contiguous()
                rc = lib.gtts_decoder_set_param(self._handle, ("estimator." + name).encode(), t.data_ptr(), t.numel())
                _lib.check(rc, f"decoder_set_param({name})")
            self._uploaded = sig
        if self._opts.get("max_chunk") != int(self.max_chunk):
            _lib.check(lib.gtts_decoder_set_option(self._handle, b"max_chunk", int(self.max_chunk)), "set_option")
            self._opts["max_chunk"] = int(self.max_chunk)
        return self._handle

    def fp32_conv_impl(self):
        """What precision='fp32' runs its convolutions on (library option fp32_tc)."""
        return "ffma" if self._opts.get("fp32_tc", 1) == 0 else "tcgen05 bf16x3 split (six partial products, fp32 TMEM accumulation)"

    def set_option(self, key, value):
        """Library option on this module's native handle (include/gradtts_b200.h: gtts_decoder_set_option)."""
        h = self._get_handle()
        _lib.check(_lib.load().gtts_decoder_set_option(h, key.encode(), int(value)), f"set_option({key})")
        self._opts[key] = int(value)

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().gtts_decoder_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:                  # interpreter shutdown: torch's module machinery may already be gone
            pass

    def __getstate__(self):
        d = self.__dict__.copy()          # the native handle is per-process: never pickled / deep-copied
        d["_handle"], d["_handle_key"], d["_uploaded"], d["_plist"], d["_opts"] = None, None, None, None, {}
        return d

    def _flags(self):
        if self.precision == "fp32":
            return _lib.FLAG_FP32
        if self.precision == "bf16":
            return 0
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")

    @staticmethod
    def _prep(t, name, device, keep_graph=False):
        _lib.require_cuda_tensor(t, name)
        if t.device != device:
            raise RuntimeError(f"{name} is on {t.device} but the decoder parameters are on {device}")
        if not keep_graph:
            t = t.detach()
        return t.to(torch.float32).contiguous()

    def _check_shapes(self, x, mask, mu, spk, need_spk=True):
        if x.dim() != 3 or x.shape[1] != self.n_feats or mu.shape != x.shape:
            raise ValueError(f"expected x and mu of shape (B, {self.n_feats}, T), got {tuple(x.shape)}, {tuple(mu.shape)}")
        B, _, T = x.shape
        if tuple(mask.shape) != (B, 1, T):
            raise ValueError(f"expected mask of shape (B, 1, T) = {(B, 1, T)}, got {tuple(mask.shape)}")
        if T % 4 != 0:
            raise ValueError("T must be a multiple of 4 (model/utils.py fix_len_compatibility)")
        if self.n_spks > 1 and need_spk:
            if spk is None or tuple(spk.shape) != (B, self.spk_emb_dim):
                raise ValueError(f"n_spks > 1: spk of shape (B, {self.spk_emb_dim}) is required")
        return B, T

    # ---- reference API -----------------------------------------------------------------------------
    def _device_params(self):
        """(name, parameter) of everything the device-side backward produces a gradient for, in a fixed order."""
        return [(n, p) for n, p in self.named_parameters() if not _is_host_side(n)]

    def _time_biases_torch(self, t):
        """The 1792 per-block time biases with PyTorch ops (model/diffusion.py:113-125, 143-144, 64-65): only to give the
        training backward an autograd path into the time MLP and the twelve Linear(64, C) layers."""
        P = dict(self.named_parameters())
        half = 32
        emb = torch.exp(torch.arange(half, device=t.device).float() * -(math.log(10000) / (half - 1)))
        emb = self.pe_scale * t.unsqueeze(1) * emb.unsqueeze(0)
        temb = torch.cat((emb.sin(), emb.cos()), dim=-1)
        temb = torch.nn.functional.linear(temb, P["mlp.0.weight"], P["mlp.0.bias"])
        temb = torch.nn.functional.linear(_mish(temb), P["mlp.2.weight"], P["mlp.2.bias"])
        a = _mish(temb)
        return torch.cat([torch.nn.functional.linear(a, P[r + ".mlp.1.weight"], P[r + ".mlp.1.bias"]) for r in _RESNET_ORDER], dim=1)

    def _spk_plane_torch(self, spk):
        P = dict(self.named_parameters())                 # model/diffusion.py:139-141, 176
        s = torch.nn.functional.linear(spk, P["spk_mlp.0.weight"], P["spk_mlp.0.bias"])
        return torch.nn.functional.linear(_mish(s), P["spk_mlp.2.weight"], P["spk_mlp.2.bias"])

    def forward(self, x, mask, mu, t, spk=None):
        """Score estimate, reference model/diffusion.py:174-216. x, mu: (B,80,T); mask: (B,1,T); t: (B,).

        Autograd: in training mode (`module.train()`, gradients enabled) the call is differentiable w.r.t. x, mu, spk and EVERY
        parameter (gtts_decoder_estimator_backward: data gradients on the tensor cores, weight gradients as fp32 reductions; the tiny
        time / speaker MLPs through PyTorch).  In eval mode it is differentiable w.r.t. `x` only
        (torch.autograd.grad(sum(est(x) * eps), x) as in the reference's likelihood code, n_best/likelihood/likelihood.py:30-34)."""
        h = self._get_handle()
        dev = next(self.parameters()).device
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            x_ = self._prep(x, "x", dev, keep_graph=True)
            mu_ = self._prep(mu, "mu", dev, keep_graph=True)
            mask_ = self._prep(mask, "mask", dev)
            B, T = self._check_shapes(x_, mask_, mu_, spk)
            t_ = self._prep(t, "t", dev).reshape(-1)
            use_spk = spk is not None and self.n_spks > 1
            spk_ = self._prep(spk, "spk", dev, keep_graph=True) if use_spk else None
            tb = self._time_biases_torch(t_)
            splane = self._spk_plane_torch(spk_) if use_spk else None
            params = [p for _, p in self._device_params()]
            return _EstimatorTrainFn.apply(self, x_, mask_, mu_, t_, spk_.detach() if use_spk else None, tb, splane, *params).to(x.dtype)
        track = torch.is_grad_enabled() and isinstance(x, torch.Tensor) and x.requires_grad
        x_ = self._prep(x, "x", dev, keep_graph=track)
        mask_, mu_ = (self._prep(v, n, dev) for v, n in ((mask, "mask"), (mu, "mu")))
        B, T = self._check_shapes(x_, mask_, mu_, spk)
        t_ = self._prep(t, "t", dev).reshape(-1)
        if t_.numel() != B:
            raise ValueError("t must have one entry per sample")
        spk_ = self._prep(spk, "spk", dev) if (spk is not None and self.n_spks > 1) else None
        out = torch.ops.gradtts_b200.estimator(int(h.value), x_, mask_, mu_, t_, spk_, self._flags())
        return out.to(x.dtype)

    @torch.no_grad()
    def vjp(self, x, mask, mu, t, v, spk=None):
        """(score, J^T v): the estimator and its vector-Jacobian product w.r.t. x in ONE pass (forward + backward share the
        intermediates on the device).  Equals (est(x), torch.autograd.grad(sum(est(x) * v), x)[0]) of the reference."""
        h = self._get_handle()
        dev = next(self.parameters()).device
        x_, mask_, mu_, v_ = (self._prep(a, n, dev) for a, n in ((x, "x"), (mask, "mask"), (mu, "mu"), (v, "v")))
        B, T = self._check_shapes(x_, mask_, mu_, spk)
        if v_.shape != x_.shape:
            raise ValueError("v must have the shape of x")
        t_ = self._prep(t, "t", dev).reshape(-1)
        if t_.numel() != B:
            raise ValueError("t must have one entry per sample")
        spk_ = self._prep(spk, "spk", dev) if (spk is not None and self.n_spks > 1) else None
        score, gx = torch.ops.gradtts_b200.estimator_vjp(int(h.value), x_, mask_, mu_, t_, spk_, v_, self._flags())
        return score.to(x.dtype), gx.to(x.dtype)

    def cache_info(self):
        """Workspace bookkeeping of the native handle (gtts_decoder_cache_info)."""
        if self._handle is None:
            return {}
        a = (ctypes.c_longlong * 5)()
        _lib.check(_lib.load().gtts_decoder_cache_info(self._handle, a, 5), "cache_info")
        return dict(plans_cached=int(a[0]), plans_created=int(a[1]), pool_bytes=int(a[2]), plan_workspace_bytes=int(a[3]),
                    plan_workspace_bytes_without_reuse=int(a[4]))

    def launches_last_call(self):
        return int(_lib.load().gtts_decoder_launches_last_call(self._handle)) if self._handle is not None else 0


def get_noise(t, beta_init, beta_term, cumulative=False):
    # model/diffusion.py:219-224 (kept for callers that import it)
    if cumulative:
        return beta_init * t + 0.5 * (beta_term - beta_init) * (t ** 2)
    return beta_init + (beta_term - beta_init) * t


class Diffusion(BaseModule):
    """Reverse-diffusion sampler (reference model/diffusion.py:227-272)."""

    def __init__(self, n_feats, dim, n_spks=1, spk_emb_dim=64, beta_min=0.05, beta_max=20, pe_scale=1000):
        super().__init__()
        self.n_feats = n_feats
        self.dim = dim
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        self.beta_min = beta_min
        self.beta_max = beta_max
        self.pe_scale = pe_scale
        self.estimator = GradLogPEstimator2d(dim, n_spks=n_spks, spk_emb_dim=spk_emb_dim, n_feats=n_feats,
                                             pe_scale=pe_scale)
        self.estimator.beta_min, self.estimator.beta_max = beta_min, beta_max

    @property
    def precision(self):
        return self.estimator.precision

    @precision.setter
    def precision(self, v):
        self.estimator.precision = v

    @torch.no_grad()
    def reverse_diffusion(self, z, mask, mu, n_timesteps, stoc=False, spk=None, sde_noise=None):
        """N Euler steps of the reverse ODE (model/diffusion.py:254-268).

        `stoc` is accepted and ignored, exactly like the reference fork (it never reads the flag).
        `sde_noise` (n_timesteps, B, 80, T) is an extension that switches on the upstream (huawei-noah Grad-TTS) stochastic
        branch this fork deleted, with caller-supplied noise in place of its torch.randn draw:
        x <- (x - ((0.5*(mu-x) - score)*beta*h + noise*sqrt(beta*h))) * mask   (the noise is subtracted, as upstream does;
        BASELINE.json's "+ sqrt(beta*h)*z" form is the same update with -noise).
        """
        est = self.estimator
        est.beta_min, est.beta_max, est.pe_scale = self.beta_min, self.beta_max, self.pe_scale
        h = est._get_handle()
        dev = next(self.parameters()).device
        z_, mask_, mu_ = (est._prep(v, n, dev) for v, n in ((z, "z"), (mask, "mask"), (mu, "mu")))
        B, T = est._check_shapes(z_, mask_, mu_, spk)
        n_timesteps = int(n_timesteps)
        spk_ = est._prep(spk, "spk", dev) if (spk is not None and self.n_spks > 1) else None
        flags = est._flags()
        noise_ = None
        if sde_noise is not None:
            noise_ = est._prep(sde_noise, "sde_noise", dev)
            if tuple(noise_.shape) != (n_timesteps, B, self.n_feats, T):
                raise ValueError("sde_noise must have shape (n_timesteps, B, 80, T)")
            flags |= _lib.FLAG_SDE
        out = torch.ops.gradtts_b200.reverse_diffusion(int(h.value), z_, mask_, mu_, spk_, noise_, n_timesteps, flags)
        return out.to(z.dtype)

    @torch.no_grad()
    def forward(self, z, mask, mu, n_timesteps, stoc=False, spk=None):
        # model/diffusion.py:270-272
        return self.reverse_diffusion(z, mask, mu, n_timesteps, stoc, spk)

    def reverse_diffusion_host(self, z, mask, mu, n_timesteps, spk=None, out=None):
        """Same computation through the host-buffer C entry point: CPU (ideally pinned) tensors in, CPU tensor
        out; the H2D/D2H copies happen inside the call (this is what bench.py's `e2e` leg times)."""
        est = self.estimator
        est.beta_min, est.beta_max, est.pe_scale = self.beta_min, self.beta_max, self.pe_scale
        h = est._get_handle()
        for t, n in ((z, "z"), (mask, "mask"), (mu, "mu")) + (((spk, "spk"),) if spk is not None else ()):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous float32 CPU tensor")
        B, T = est._check_shapes(z, mask, mu, spk)
        if self.n_spks <= 1:
            spk = None
        if out is None:
            out = torch.empty_like(z)
        elif out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.shape != z.shape:
            raise ValueError("out must be a contiguous float32 CPU tensor of the shape of z")
        # ordering against earlier calls on other streams is done inside the library (one event per handle)
        rc = _lib.load().gtts_decoder_reverse_diffusion_host(
            h, z.data_ptr(), mask.data_ptr(), mu.data_ptr(), spk.data_ptr() if spk is not None else None,
            out.data_ptr(), B, T, int(n_timesteps), est._flags())
        _lib.check(rc, "reverse_diffusion_host")
        return out

    # ---- training-side members of the reference class: forward VALUE only (no estimator backward yet) ----
    def _fd_args(self, x0, mask, mu, t):
        est = self.estimator
        dev = next(self.parameters()).device
        x_, mask_, mu_ = (est._prep(v, n, dev) for v, n in ((x0, "x0"), (mask, "mask"), (mu, "mu")))
        B, T = est._check_shapes(x_, mask_, mu_, None, need_spk=False)
        t_ = est._prep(t, "t", dev).reshape(-1)
        if t_.numel() != B:
            raise ValueError("t must have one entry per sample")
        return dev, x_, mask_, mu_, t_, B, T

    @torch.no_grad()
    def forward_diffusion(self, x0, mask, mu, t, noise=None):
        """model/diffusion.py:244-252 as one kernel.  `noise` (B,80,T) replaces the reference's internal torch.randn draw
        (default: drawn here with torch.randn on the device, like the reference).  Returns (xt * mask, z * mask)."""
        dev, x_, mask_, mu_, t_, B, T = self._fd_args(x0, mask, mu, t)
        if noise is None:
            noise = torch.randn(x_.shape, dtype=torch.float32, device=dev)
        z_ = self.estimator._prep(noise, "noise", dev)
        if z_.shape != x_.shape:
            raise ValueError("noise must have the shape of x0")
        xt, zm = torch.ops.gradtts_b200.forward_diffusion(x_, mask_, mu_, t_, z_, float(self.beta_min), float(self.beta_max))
        return xt.to(x0.dtype), zm.to(x0.dtype)

    def loss_t(self, x0, mask, mu, t, spk=None, noise=None):
        """model/diffusion.py:274-281 -> (loss, xt).  In training mode (`train()`, gradients enabled) the loss carries the autograd
        graph: loss.backward() fills .grad of every estimator parameter and flows back into mu / spk.  Otherwise the forward value
        is computed entirely on the device (forward diffusion, estimator, squared-error reduction kernels)."""
        if torch.is_grad_enabled() and self.estimator.training:
            # training: the elementwise parts (model/diffusion.py:244-252, 276-280) stay PyTorch ops so that autograd carries the loss
            # back to mu (the text encoder) and into the estimator call, whose backward is gtts_decoder_estimator_backward
            time = t.unsqueeze(-1).unsqueeze(-1)
            cum_noise = get_noise(time, self.beta_min, self.beta_max, cumulative=True)
            mean = x0 * torch.exp(-0.5 * cum_noise) + mu * (1.0 - torch.exp(-0.5 * cum_noise))
            variance = 1.0 - torch.exp(-cum_noise)
            z = noise if noise is not None else torch.randn(x0.shape, dtype=x0.dtype, device=x0.device, requires_grad=False)
            xt = (mean + z * torch.sqrt(variance)) * mask
            z = z * mask
            noise_estimation = self.estimator(xt, mask, mu, t, spk) * torch.sqrt(1.0 - torch.exp(-cum_noise))
            loss = torch.sum((noise_estimation + z) ** 2) / (torch.sum(mask) * self.n_feats)
            return loss, xt
        dev, x_, mask_, mu_, t_, B, T = self._fd_args(x0, mask, mu, t)
        xt, zm = self.forward_diffusion(x_, mask_, mu_, t_, noise)
        est = self.estimator(xt, mask_, mu_, t_, spk)
        loss = torch.ops.gradtts_b200.score_loss(est.to(torch.float32).contiguous(), zm, mask_, t_, float(self.beta_min),
                                                 float(self.beta_max))
        return loss.to(x0.dtype), xt.to(x0.dtype)

    def compute_loss(self, x0, mask, mu, spk=None, offset=1e-5):
        # model/diffusion.py:283-287
        t = torch.rand(x0.shape[0], dtype=x0.dtype, device=x0.device, requires_grad=False)
        t = torch.clamp(t, offset, 1.0 - offset)
        return self.loss_t(x0, mask, mu, t, spk)
