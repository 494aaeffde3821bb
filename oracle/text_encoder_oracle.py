"""TEST INFRASTRUCTURE -- restatement of the reference's TextEncoder forward (eval mode) on a state dict.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file; the product never does.
Follows /root/reference/model/text_encoder.py: TextEncoder.forward :321-335, ConvReluNorm.forward :53-63, Encoder.forward
:271-282, MultiHeadAttention.forward/attention :140-182 (relative-position logits and values :185-216 written here as an explicit
gather over the offset j - i instead of the pad/reshape trick), FFN.forward :235-241, DurationPredictor.forward :83-93,
LayerNorm.forward :20-27.  Pinned against the reference module by tests/golden/enc_*.npz (tests/golden/make_golden.py enc).
"""
import math

import torch
import torch.nn.functional as F


def layer_norm(sd, name, x, eps=1e-4):
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    x = (x - mean) * torch.rsqrt(var + eps)
    return x * sd[name + ".gamma"].view(1, -1, 1) + sd[name + ".beta"].view(1, -1, 1)


def conv(sd, name, x):
    w = sd[name + ".weight"]
    return F.conv1d(x, w, sd[name + ".bias"], padding=w.shape[-1] // 2)


def attention(sd, name, x, attn_mask, n_heads, window):
    B, C, T = x.shape
    kc = C // n_heads
    q = conv(sd, name + ".conv_q", x).view(B, n_heads, kc, T).transpose(2, 3)        # (B, H, T, kc)
    k = conv(sd, name + ".conv_k", x).view(B, n_heads, kc, T).transpose(2, 3)
    v = conv(sd, name + ".conv_v", x).view(B, n_heads, kc, T).transpose(2, 3)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(kc)
    if window is not None:
        idx = torch.arange(T, device=x.device)
        d = idx[None, :] - idx[:, None]                                             # d[i, j] = j - i
        inside = (d.abs() <= window)
        ek = sd[name + ".emb_rel_k"][0]                                             # (2w+1, kc), shared by the heads
        ev = sd[name + ".emb_rel_v"][0]
        ek_ij = ek[(d.clamp(-window, window) + window)] * inside[..., None]         # (T, T, kc), zero outside the window
        scores = scores + torch.einsum("bhic,ijc->bhij", q, ek_ij) / math.sqrt(kc)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p = torch.softmax(scores, dim=-1)
    out = torch.matmul(p, v)
    if window is not None:
        ev_ij = ev[(d.clamp(-window, window) + window)] * inside[..., None]
        out = out + torch.einsum("bhij,ijc->bhic", p, ev_ij)
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return conv(sd, name + ".conv_o", out)


def text_encoder_forward(sd, cfg, x, x_lengths, spk=None):
    """tokens (B, T) -> mu (B, n_feats, T), logw (B, 1, T), x_mask (B, 1, T)."""
    C0 = cfg["n_channels"]
    T = x.shape[1]
    h = F.embedding(x, sd["emb.weight"]) * math.sqrt(C0)
    h = h.transpose(1, -1)
    x_mask = (torch.arange(T, device=x.device)[None, :] < x_lengths[:, None]).unsqueeze(1).to(h.dtype)
    # prenet
    org = h
    for i in range(3):
        h = conv(sd, f"prenet.conv_layers.{i}", h * x_mask)
        h = torch.relu(layer_norm(sd, f"prenet.norm_layers.{i}", h))
    h = (org + conv(sd, "prenet.proj", h)) * x_mask
    if cfg["n_spks"] > 1:
        h = torch.cat([h, spk.unsqueeze(-1).repeat(1, 1, T)], dim=1)
    # encoder
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    for i in range(cfg["n_layers"]):
        h = h * x_mask
        y = attention(sd, f"encoder.attn_layers.{i}", h, attn_mask, cfg["n_heads"], cfg["window_size"])
        h = layer_norm(sd, f"encoder.norm_layers_1.{i}", h + y)
        y = torch.relu(conv(sd, f"encoder.ffn_layers.{i}.conv_1", h * x_mask))
        y = conv(sd, f"encoder.ffn_layers.{i}.conv_2", y * x_mask) * x_mask
        h = layer_norm(sd, f"encoder.norm_layers_2.{i}", h + y)
    h = h * x_mask
    mu = conv(sd, "proj_m", h) * x_mask
    # duration predictor
    d = torch.relu(conv(sd, "proj_w.conv_1", h * x_mask))
    d = layer_norm(sd, "proj_w.norm_1", d)
    d = torch.relu(conv(sd, "proj_w.conv_2", d * x_mask))
    d = layer_norm(sd, "proj_w.norm_2", d)
    logw = conv(sd, "proj_w.proj", d * x_mask) * x_mask
    return mu, logw, x_mask
