"""TEST INFRASTRUCTURE -- CPU (or any-device) restatement of the reference's HiFi-GAN generator forward.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file; the product never does.
Follows /root/reference/hifi-gan/models.py: Generator.forward :101-118, ResBlock1.forward :38-45, ResBlock2.forward :66-70,
get_padding hifi-gan/xutils.py:37-38, weight norm = torch.nn.utils.weight_norm(dim=0) as applied at models.py:17-35,81,88-99.
Pinned against the reference module itself by tests/golden/voc_*.npz (tests/golden/make_golden.py voc).
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1


def effective_weight(sd, name):
    """`name.weight` after remove_weight_norm: v * (g / ||v||), the norm over every dim but 0."""
    if name + ".weight" in sd:
        return sd[name + ".weight"]
    v, g = sd[name + ".weight_v"], sd[name + ".weight_g"]
    return v * (g / v.flatten(1).norm(dim=1).reshape(-1, 1, 1))


def _pad(k, d=1):
    return int((k * d - d) / 2)


def _conv(sd, name, x, k, d=1):
    return F.conv1d(x, effective_weight(sd, name), sd[name + ".bias"], padding=_pad(k, d), dilation=d)


def resblock(sd, base, kind, x, k, dilations):
    for m, d in enumerate(dilations):
        if str(kind) == "1":
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = _conv(sd, f"{base}.convs1.{m}", xt, k, d)
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = _conv(sd, f"{base}.convs2.{m}", xt, k, 1)
        else:
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = _conv(sd, f"{base}.convs.{m}", xt, k, d)
        x = xt + x
    return x


def generator_forward(sd, cfg, mel):
    """mel (B, 80, T) -> (B, 1, T * prod(upsample_rates))."""
    nk = len(cfg["resblock_kernel_sizes"])
    x = _conv(sd, "conv_pre", mel, 7)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, effective_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            y = resblock(sd, f"resblocks.{i * nk + j}", cfg["resblock"], x, rk, rd)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                       # PyTorch default slope 0.01 (models.py:114)
    x = _conv(sd, "conv_post", x, 7)
    return torch.tanh(x)
