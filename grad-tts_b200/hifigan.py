"""HiFi-GAN generator on the B200: host-side mirror of the reference's `hifi-gan/models.py::Generator`.

Same constructor (`Generator(h)` with the attribute-dict of checkpts/hifigan-config.json), same `state_dict` keys as the
weight-normed reference module (`conv_pre.weight_g / weight_v / bias`, `ups.N.*`, `resblocks.N.convs1.M.*`, `conv_post.*`), so
`vocoder.load_state_dict(torch.load(ckpt)['generator'])`, `.cuda().eval()`, `.remove_weight_norm()` and
`vocoder.forward(mel)` (reference inference.py:73-76,97) work unchanged.  The module holds parameters only: the arithmetic runs
in csrc/vocoder.cu behind `gtts_vocoder_*` (include/gradtts_b200.h).  No CPU path.
"""
import ctypes
import math

import torch
from torch import nn

from . import _lib
from . import ops as _ops  # noqa: F401  (registers torch.ops.gradtts_b200.*)


class AttrDict(dict):
    """hifi-gan/env.py:8-11."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


V1_CONFIG = AttrDict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
                     resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


class _WNConv(nn.Module):
    """Parameters of one weight-normed Conv1d / ConvTranspose1d: `weight_g`, `weight_v`, `bias` (what
    torch.nn.utils.weight_norm(conv) registers, hifi-gan/models.py:17-23,81,88-91), or `weight`, `bias` after
    remove_weight_norm()."""

    def __init__(self, w_shape, n_bias, fan_in):
        super().__init__()
        v = torch.randn(*w_shape) * 0.01                                    # init_weights, hifi-gan/xutils.py:25-28
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).reshape(-1, 1, 1))
        self.weight_v = nn.Parameter(v)
        bound = 1.0 / math.sqrt(fan_in)
        self.bias = nn.Parameter(torch.empty(n_bias).uniform_(-bound, bound))

    def effective_weight(self):
        if "weight" in self._parameters:
            return self.weight
        v = self.weight_v
        norm = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)                   # weight_norm(dim=0): one norm per slice of dim 0
        return v * (self.weight_g / norm)

    def remove_weight_norm(self):
        if "weight" in self._parameters:
            return
        w = self.effective_weight().detach().clone()
        del self.weight_g
        del self.weight_v
        self.weight = nn.Parameter(w)


class _ResBlock(nn.Module):
    def __init__(self, kind, channels, kernel_size, dilation):
        super().__init__()
        mk = lambda: _WNConv((channels, channels, kernel_size), channels, channels * kernel_size)   # noqa: E731
        if kind == "1":
            self.convs1 = nn.ModuleList([mk() for _ in dilation])
            self.convs2 = nn.ModuleList([mk() for _ in dilation])
        else:
            self.convs = nn.ModuleList([mk() for _ in dilation])

    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.remove_weight_norm()


class Generator(nn.Module):
    """hifi-gan/models.py:77-126."""

    def __init__(self, h):
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.num_mels = int(getattr(h, "num_mels", 80)) if not isinstance(h, dict) else int(h.get("num_mels", 80))
        if self.num_mels != 80:
            # the reference hard-codes Conv1d(80, ...) (models.py:81)
            raise ValueError("the reference generator takes 80 mel channels")
        c0 = h.upsample_initial_channel
        self.conv_pre = _WNConv((c0, 80, 7), c0, 80 * 7)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
            self.ups.append(_WNConv((cin, cout, k), cout, cin * k))          # ConvTranspose1d weight: (in, out, k)
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(_ResBlock(str(h.resblock), ch, k, d))
        self.conv_post = _WNConv((1, ch, 7), 1, ch * 7)
        self.precision = "bf16"            # "bf16": tcgen05 convs, bf16 activations; "fp32": fp32 activations, FFMA convs
        self.max_chunk = 32
        self._handle = None
        self._handle_dev = None
        self._uploaded = None
        self._plist = None
        self._opts = {}

    # ------------------------------------------------------------------------------------------------ reference API
    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.remove_weight_norm()
        self._plist = None

    def forward(self, x):
        """mel (B, 80, T) -> waveform (B, 1, T * prod(upsample_rates)), models.py:101-118."""
        _lib.require_cuda_tensor(x, "mel")
        if x.dim() != 3 or x.shape[1] != 80:
            raise ValueError("mel must be (B, 80, T)")
        mel = x.detach()
        if mel.dtype != torch.float32 or not mel.is_contiguous():
            mel = mel.to(torch.float32).contiguous()
        B, _, T = mel.shape
        if B == 0 or T == 0:
            return torch.empty(B, 1, T * self.hop, dtype=x.dtype, device=mel.device)
        h = self._get_handle(mel.device)
        out = torch.ops.gradtts_b200.vocoder(int(h.value), mel, self.hop, self._flags())
        return out.to(x.dtype)

    # ------------------------------------------------------------------------------------------------ plumbing
    @property
    def hop(self):
        u = 1
        for r in self.h.upsample_rates:
            u *= int(r)
        return u

    def _flags(self):
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        return _lib.FLAG_FP32 if self.precision == "fp32" else 0

    def _convs(self):
        for name, m in self.named_modules():
            if isinstance(m, _WNConv):
                yield name, m

    def _signature(self):
        # flat list cached (the module-tree walk over 150 tensors costs more than the call at one utterance); weight norm on / off
        # and .to() / .cuda() change the Parameter objects and reset it
        if self._plist is None:
            self._plist = list(self.parameters())
        return tuple((id(p), p.data_ptr(), p._version) for p in self._plist)

    def _apply(self, fn, *a, **k):
        self._plist = None
        return super()._apply(fn, *a, **k)

    def _release(self):
        if self._handle is not None:
            _lib.load().gtts_vocoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def set_option(self, name, value):
        h = self._get_handle(next(self.parameters()).device)
        _lib.check(_lib.load().gtts_vocoder_set_option(h, name.encode(), int(value)), f"vocoder_set_option({name})")

    def cache_info(self):
        """plans cached, bytes of the shared workspace pool, peak live workspace of the most recent plan"""
        if self._handle is None:
            return {"plans": 0, "pool_bytes": 0, "plan_peak_bytes": 0}
        out = (ctypes.c_longlong * 3)()
        _lib.check(_lib.load().gtts_vocoder_cache_info(self._handle, out, 3), "vocoder_cache_info")
        return {"plans": int(out[0]), "pool_bytes": int(out[1]), "plan_peak_bytes": int(out[2])}

    def launches_last_call(self):
        return int(_lib.load().gtts_vocoder_launches_last_call(self._handle)) if self._handle is not None else 0

    def profile(self, B, T):
        h = self._get_handle(next(self.parameters()).device)
        buf = ctypes.create_string_buffer(1 << 16)
        dev = next(self.parameters()).device
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().gtts_vocoder_profile(h, int(B), int(T), self._flags(), buf, len(buf), stream), "vocoder_profile")
        return buf.value.decode()

    def _get_handle(self, device):
        lib = _lib.load()
        p0 = next(self.parameters())
        _lib.require_cuda_tensor(p0, "vocoder parameters")
        if p0.device != device:
            raise ValueError("mel and vocoder parameters are on different devices")
        dev = p0.device.index if p0.device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_dev != dev:
            self._release()
            hcfg = self.h
            rates = [int(r) for r in hcfg.upsample_rates]
            upk = [int(k) for k in hcfg.upsample_kernel_sizes]
            rbk = [int(k) for k in hcfg.resblock_kernel_sizes]
            dil = [[int(d) for d in row] for row in hcfg.resblock_dilation_sizes]
            n_dil = len(dil[0])
            if any(len(row) != n_dil for row in dil) or len(dil) != len(rbk):
                raise ValueError("resblock_dilation_sizes must be a rectangular list, one row per resblock kernel")
            arr = lambda xs: (ctypes.c_int * len(xs))(*xs)                   # noqa: E731
            flat = [d for row in dil for d in row]
            handle = ctypes.c_void_p()
            rc = lib.gtts_vocoder_create(ctypes.byref(handle), int(str(hcfg.resblock)), len(rates), arr(rates), arr(upk),
                                         int(hcfg.upsample_initial_channel), len(rbk), arr(rbk), arr(flat), n_dil, 80, dev)
            _lib.check(rc, "vocoder_create")
            self._handle, self._handle_dev, self._uploaded, self._opts = handle, dev, None, {}
        sig = self._signature()
        if self._uploaded != sig:
            torch.cuda.current_stream(p0.device).synchronize()
            with torch.no_grad():
                for name, m in self._convs():
                    w = m.effective_weight().detach().to(torch.float32).contiguous()
                    b = m.bias.detach().to(torch.float32).contiguous()
                    torch.cuda.current_stream(p0.device).synchronize()
                    _lib.check(lib.gtts_vocoder_set_param(self._handle, (name + ".weight").encode(), w.data_ptr(), w.numel()),
                               f"vocoder_set_param({name}.weight)")
                    _lib.check(lib.gtts_vocoder_set_param(self._handle, (name + ".bias").encode(), b.data_ptr(), b.numel()),
                               f"vocoder_set_param({name}.bias)")
            self._uploaded = sig
        if self._opts.get("max_chunk") != int(self.max_chunk):
            _lib.check(lib.gtts_vocoder_set_option(self._handle, b"max_chunk", int(self.max_chunk)), "vocoder_set_option")
            self._opts["max_chunk"] = int(self.max_chunk)
        return self._handle
