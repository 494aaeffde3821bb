"""Text encoder on the B200: host-side mirror of the reference's `model/text_encoder.py::TextEncoder`.

Same constructor, same `state_dict` keys and shapes (so a reference checkpoint loads with strict=True), same
`forward(x, x_lengths, spk=None) -> (mu, logw, x_mask)`.  The modules below hold parameters only; the arithmetic runs in
csrc/text_encoder.cu behind `gtts_encoder_*` (include/gradtts_b200.h), fp32 on the CUDA cores.  Inference only: dropout is the
identity and no autograd graph is built (for training inject the reference's PyTorch encoder: `GradTTS(..., encoder=...)`).
"""
import ctypes

import torch
from torch import nn

from .. import _lib
from .. import ops as _ops  # noqa: F401  (registers torch.ops.gradtts_b200.*)
from .base import BaseModule


class _Norm(nn.Module):                      # parameters of text_encoder.py:11-18
    def __init__(self, channels):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


class _Prenet(nn.Module):                    # parameters of ConvReluNorm, text_encoder.py:30-51
    def __init__(self, channels, kernel_size=5, n_layers=3):
        super().__init__()
        self.conv_layers = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, padding=kernel_size // 2) for _ in range(n_layers)])
        self.norm_layers = nn.ModuleList([_Norm(channels) for _ in range(n_layers)])
        self.proj = nn.Conv1d(channels, channels, 1)
        self.proj.weight.data.zero_()
        self.proj.bias.data.zero_()


class _Attention(nn.Module):                 # parameters of MultiHeadAttention, text_encoder.py:96-138
    def __init__(self, channels, n_heads, window_size):
        super().__init__()
        kc = channels // n_heads
        self.conv_q = nn.Conv1d(channels, channels, 1)
        self.conv_k = nn.Conv1d(channels, channels, 1)
        self.conv_v = nn.Conv1d(channels, channels, 1)
        if window_size is not None:
            self.emb_rel_k = nn.Parameter(torch.randn(1, window_size * 2 + 1, kc) * kc ** -0.5)
            self.emb_rel_v = nn.Parameter(torch.randn(1, window_size * 2 + 1, kc) * kc ** -0.5)
        self.conv_o = nn.Conv1d(channels, channels, 1)
        for c in (self.conv_q, self.conv_k, self.conv_v):
            nn.init.xavier_uniform_(c.weight)


class _FFN(nn.Module):                       # text_encoder.py:219-233
    def __init__(self, channels, filter_channels, kernel_size):
        super().__init__()
        self.conv_1 = nn.Conv1d(channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.conv_2 = nn.Conv1d(filter_channels, channels, kernel_size, padding=kernel_size // 2)


class _Encoder(nn.Module):                   # text_encoder.py:244-269
    def __init__(self, channels, filter_channels, n_heads, n_layers, kernel_size, window_size):
        super().__init__()
        self.attn_layers = nn.ModuleList([_Attention(channels, n_heads, window_size) for _ in range(n_layers)])
        self.norm_layers_1 = nn.ModuleList([_Norm(channels) for _ in range(n_layers)])
        self.ffn_layers = nn.ModuleList([_FFN(channels, filter_channels, kernel_size) for _ in range(n_layers)])
        self.norm_layers_2 = nn.ModuleList([_Norm(channels) for _ in range(n_layers)])


class _DurationPredictor(nn.Module):         # text_encoder.py:67-81
    def __init__(self, channels, filter_channels, kernel_size):
        super().__init__()
        self.conv_1 = nn.Conv1d(channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_1 = _Norm(filter_channels)
        self.conv_2 = nn.Conv1d(filter_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_2 = _Norm(filter_channels)
        self.proj = nn.Conv1d(filter_channels, 1, 1)


class TextEncoder(BaseModule):
    """reference model/text_encoder.py:285-335."""

    def __init__(self, n_vocab, n_feats, n_channels, filter_channels, filter_channels_dp, n_heads, n_layers, kernel_size,
                 p_dropout, window_size=None, spk_emb_dim=64, n_spks=1):
        super().__init__()
        self.n_vocab = n_vocab
        self.n_feats = n_feats
        self.n_channels = n_channels
        self.filter_channels = filter_channels
        self.filter_channels_dp = filter_channels_dp
        self.n_heads = n_heads
        self.n_layers = n_layers
        self.kernel_size = kernel_size
        self.p_dropout = p_dropout
        self.window_size = window_size
        self.spk_emb_dim = spk_emb_dim
        self.n_spks = n_spks
        width = n_channels + (spk_emb_dim if n_spks > 1 else 0)
        self.emb = nn.Embedding(n_vocab, n_channels)
        nn.init.normal_(self.emb.weight, 0.0, n_channels ** -0.5)
        self.prenet = _Prenet(n_channels)
        self.encoder = _Encoder(width, filter_channels, n_heads, n_layers, kernel_size, window_size)
        self.proj_m = nn.Conv1d(width, n_feats, 1)
        self.proj_w = _DurationPredictor(width, filter_channels_dp, kernel_size)
        self._handle = None
        self._handle_dev = None
        self._uploaded = None
        self._plist = None

    def _apply(self, fn, *a, **k):          # .to() / .cuda() may replace the Parameter objects
        self._plist = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x, x_lengths, spk=None):
        """tokens (B, T) int64, lengths (B) -> mu (B, n_feats, T), logw (B, 1, T), x_mask (B, 1, T); text_encoder.py:321-335."""
        if self.training and self.p_dropout > 0:
            raise RuntimeError("the native TextEncoder is inference-only (eval mode): dropout and the backward pass are not implemented; "
                               "inject the reference's PyTorch encoder for training (GradTTS(..., encoder=...))")
        _lib.require_cuda_tensor(x, "x")
        if x.dim() != 2:
            raise ValueError("x must be (B, T) token ids")
        B, T = x.shape
        dev = x.device
        tokens = x.to(torch.int64).contiguous()
        lengths = x_lengths.to(device=dev, dtype=torch.int64).contiguous()
        if lengths.shape != (B,):
            raise ValueError("x_lengths must be (B,)")
        if self.n_spks > 1:
            if spk is None:
                raise ValueError("this encoder was built with n_spks > 1: spk is required")
            spk = spk.detach().to(device=dev, dtype=torch.float32).contiguous()
            if spk.shape != (B, self.spk_emb_dim):
                raise ValueError("spk must be (B, spk_emb_dim)")
        if B == 0 or T == 0:
            f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)      # noqa: E731
            return f(B, self.n_feats, T), f(B, 1, T), f(B, 1, T)
        h = self._get_handle(dev)
        return torch.ops.gradtts_b200.text_encoder(int(h.value), tokens, lengths, spk if self.n_spks > 1 else None, int(self.n_feats))

    # ------------------------------------------------------------------------------------------------ plumbing
    def launches_last_call(self):
        return int(_lib.load().gtts_encoder_launches_last_call(self._handle)) if self._handle is not None else 0

    def _release(self):
        if self._handle is not None:
            _lib.load().gtts_encoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _get_handle(self, device):
        lib = _lib.load()
        p0 = next(self.parameters())
        _lib.require_cuda_tensor(p0, "text encoder parameters")
        if p0.device != device:
            raise ValueError("tokens and text encoder parameters are on different devices")
        dev = p0.device.index if p0.device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_dev != dev:
            self._release()
            h = ctypes.c_void_p()
            rc = lib.gtts_encoder_create(ctypes.byref(h), int(self.n_vocab), int(self.n_feats), int(self.n_channels),
                                         int(self.filter_channels), int(self.filter_channels_dp), int(self.n_heads), int(self.n_layers),
                                         int(self.kernel_size), -1 if self.window_size is None else int(self.window_size),
                                         int(self.spk_emb_dim), int(self.n_spks), dev)
            _lib.check(rc, "encoder_create")
            self._handle, self._handle_dev, self._uploaded = h, dev, None
        if self._plist is None:            # flat list cached: the module-tree walk costs more than the call at one utterance
            self._plist = list(self.parameters())
        sig = tuple((id(p), p.data_ptr(), p._version) for p in self._plist)
        if self._uploaded != sig:
            torch.cuda.current_stream(p0.device).synchronize()
            for name, p in self.named_parameters():
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(torch.float32).contiguous()
                    torch.cuda.current_stream(p0.device).synchronize()
                _lib.check(lib.gtts_encoder_set_param(self._handle, name.encode(), t.data_ptr(), t.numel()), f"encoder_set_param({name})")
            self._uploaded = sig
        return self._handle
