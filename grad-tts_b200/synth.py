"""Seeded synthetic weights and inputs for the decoder (no checkpoints ship with the reference).

`decoder_param_shapes` lists every tensor of `Diffusion.state_dict()` in the reference's own
order and naming (model/diffusion.py:128-172, 227-242); `make_decoder_state_dict` fills them
from a CPU `torch.Generator`, so the same seed gives the same weights in the build container
(where the golden fixtures are made with the real reference) and on the GPU box.
"""
import math

import torch

DIM = 64
DIM_MULTS = (1, 2, 4)
N_FEATS = 80
SPK_EMB_DIM = 64


def _resnet_shapes(p, cin, cout, tdim=DIM):
    s = [(p + ".mlp.1.weight", (cout, tdim)), (p + ".mlp.1.bias", (cout,))]
    for blk, ci in (("block1", cin), ("block2", cout)):
        s += [(f"{p}.{blk}.block.0.weight", (cout, ci, 3, 3)), (f"{p}.{blk}.block.0.bias", (cout,)),
              (f"{p}.{blk}.block.1.weight", (cout,)), (f"{p}.{blk}.block.1.bias", (cout,))]
    if cin != cout:
        s += [(p + ".res_conv.weight", (cout, cin, 1, 1)), (p + ".res_conv.bias", (cout,))]
    return s


def _attn_shapes(p, c):
    return [(p + ".fn.g", (1,)), (p + ".fn.fn.to_qkv.weight", (384, c, 1, 1)),
            (p + ".fn.fn.to_out.weight", (c, 128, 1, 1)), (p + ".fn.fn.to_out.bias", (c,))]


def decoder_param_shapes(n_spks=1, pfx="estimator."):
    """(name, shape) for every tensor in the reference `Diffusion.state_dict()`, in order."""
    s = []
    if n_spks > 1 or n_spks == -1:
        s += [("spk_mlp.0.weight", (SPK_EMB_DIM * 4, SPK_EMB_DIM)), ("spk_mlp.0.bias", (SPK_EMB_DIM * 4,)),
              ("spk_mlp.2.weight", (N_FEATS, SPK_EMB_DIM * 4)), ("spk_mlp.2.bias", (N_FEATS,))]
    s += [("mlp.0.weight", (DIM * 4, DIM)), ("mlp.0.bias", (DIM * 4,)),
          ("mlp.2.weight", (DIM, DIM * 4)), ("mlp.2.bias", (DIM,))]
    dims = [2 + (1 if n_spks > 1 else 0)] + [DIM * m for m in DIM_MULTS]
    in_out = list(zip(dims[:-1], dims[1:]))
    for i, (ci, co) in enumerate(in_out):
        s += _resnet_shapes(f"downs.{i}.0", ci, co)
        s += _resnet_shapes(f"downs.{i}.1", co, co)
        s += _attn_shapes(f"downs.{i}.2", co)
        if i < len(in_out) - 1:
            s += [(f"downs.{i}.3.conv.weight", (co, co, 3, 3)), (f"downs.{i}.3.conv.bias", (co,))]
    ups = []
    for i, (ci, co) in enumerate(reversed(in_out[1:])):
        ups += _resnet_shapes(f"ups.{i}.0", co * 2, ci)
        ups += _resnet_shapes(f"ups.{i}.1", ci, ci)
        ups += _attn_shapes(f"ups.{i}.2", ci)
        ups += [(f"ups.{i}.3.conv.weight", (ci, ci, 4, 4)), (f"ups.{i}.3.conv.bias", (ci,))]
    s += ups
    mid = dims[-1]
    s += _resnet_shapes("mid_block1", mid, mid)
    s += _attn_shapes("mid_attn", mid)
    s += _resnet_shapes("mid_block2", mid, mid)
    s += _resnet_shapes_block("final_block", DIM, DIM)
    s += [("final_conv.weight", (1, DIM, 1, 1)), ("final_conv.bias", (1,))]
    return [(pfx + n, sh) for n, sh in s]


def _resnet_shapes_block(p, cin, cout):
    return [(f"{p}.block.0.weight", (cout, cin, 3, 3)), (f"{p}.block.0.bias", (cout,)),
            (f"{p}.block.1.weight", (cout,)), (f"{p}.block.1.bias", (cout,))]


def make_decoder_state_dict(n_spks=1, seed=0, g=0.05, pfx="estimator."):
    """Random weights at PyTorch-default-like scale; GroupNorm affine and Rezero `g` perturbed so
    that every term of the maths is exercised (default init has g=0, gamma=1, beta=0)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    sd = {}
    for name, shape in decoder_param_shapes(n_spks, pfx):
        if name.endswith(".fn.g"):
            v = torch.full(shape, float(g))
        elif ".block.1.weight" in name:                       # GroupNorm gamma
            v = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif ".block.1.bias" in name:                         # GroupNorm beta
            v = 0.1 * torch.randn(shape, generator=gen)
        else:
            if len(shape) == 1:
                # bias of the layer just created: fan_in of the matching weight
                fan_in = sd[name.replace(".bias", ".weight")][0].numel()
                if ".3.conv." in name and name.startswith(pfx + "ups"):
                    w = sd[name.replace(".bias", ".weight")]   # ConvTranspose2d: fan_in = Cout*kh*kw
                    fan_in = w.shape[1] * w.shape[2] * w.shape[3]
            else:
                fan_in = math.prod(shape[1:])
            bound = 1.0 / math.sqrt(fan_in)
            v = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
        sd[name] = v.float().contiguous()
    return sd


def make_inputs(B, T, n_spks=1, seed=1, ragged=True, temperature=1.5):
    """mu ~ N(0,1), z = mu + N(0,1)/temperature, prefix mask with lengths U[0.6T, T] (item 0 full)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    mu = torch.randn(B, N_FEATS, T, generator=gen)
    z = mu + torch.randn(B, N_FEATS, T, generator=gen) / temperature
    if ragged:
        lengths = torch.randint(int(0.6 * T), T + 1, (B,), generator=gen)
        lengths[0] = T
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lengths[:, None]).float().unsqueeze(1)   # (B,1,T)
    spk = torch.randn(B, SPK_EMB_DIM, generator=gen) if (n_spks > 1) else None
    return z.contiguous(), mask.contiguous(), mu.contiguous(), spk, lengths


def make_mas_inputs(B, t_x, t_y, seed=1234, ragged=True):
    """value = 5*N(0,1) - 40 (log-prior-like), prefix masks with t_x~U[t_x/2,t_x], t_y~U[0.6t_y,t_y]."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    value = 5.0 * torch.randn(B, t_x, t_y, generator=gen) - 40.0
    if ragged:
        tx = torch.randint(max(1, t_x // 2), t_x + 1, (B,), generator=gen)
        ty = torch.randint(max(1, int(0.6 * t_y)), t_y + 1, (B,), generator=gen)
        tx[0], ty[0] = t_x, t_y
        tx = torch.minimum(tx, ty)
    else:
        tx = torch.full((B,), t_x, dtype=torch.long)
        ty = torch.full((B,), t_y, dtype=torch.long)
    xm = (torch.arange(t_x)[None, :] < tx[:, None]).float()
    ym = (torch.arange(t_y)[None, :] < ty[:, None]).float()
    mask = xm[:, :, None] * ym[:, None, :]
    return value.contiguous(), mask.contiguous(), tx, ty


# ---------------------------------------------------------------------------------------------------------------- vocoder
VOCODER_CONFIGS = {
    # checkpts/hifigan-config.json (HiFi-GAN V1)
    "v1": dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
               resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]]),
    # a ResBlock2 stack (the shape family of the upstream V3 config): two dilations per block, stride-4 upsampling with kernel 8
    "rb2": dict(resblock="2", upsample_rates=[8, 4, 4], upsample_kernel_sizes=[16, 8, 8], upsample_initial_channel=256,
                resblock_kernel_sizes=[3, 5, 7], resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]]),
}


def vocoder_param_shapes(cfg):
    """(module name, weight shape, transposed) for every conv of hifi-gan/models.py::Generator(h), in state_dict order."""
    c0 = cfg["upsample_initial_channel"]
    s = [("conv_pre", (c0, 80, 7), False)]
    for i, k in enumerate(cfg["upsample_kernel_sizes"]):
        s.append((f"ups.{i}", (c0 // 2 ** i, c0 // 2 ** (i + 1), k), True))
    n = 0
    ch = c0
    for i in range(len(cfg["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for k, d in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
            groups = ("convs1", "convs2") if str(cfg["resblock"]) == "1" else ("convs",)
            for grp in groups:
                for m in range(len(d)):
                    s.append((f"resblocks.{n}.{grp}.{m}", (ch, ch, k), False))
            n += 1
    s.append(("conv_post", (1, ch, 7), False))
    return s


def make_vocoder_state_dict(cfg, seed=0):
    """Weight-normed state dict (`*.weight_g`, `*.weight_v`, `*.bias`) with gains that keep the activations O(1) through the
    stack (the reference's own init, N(0, 0.01), gives a near-constant output -- useless as a test signal)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    rates = cfg["upsample_rates"]
    for name, shape, transposed in vocoder_param_shapes(cfg):
        v = torch.randn(*shape, generator=gen)
        if transposed:
            u = rates[int(name.split(".")[1])]
            gain = math.sqrt(u / 2.0)
        elif name.startswith("resblocks"):
            gain = 0.6
        else:
            gain = 1.2
        g = gain * (1.0 + 0.2 * torch.randn(shape[0], 1, 1, generator=gen))
        nb = shape[1] if transposed else shape[0]
        sd[name + ".weight_g"] = g
        sd[name + ".weight_v"] = v
        sd[name + ".bias"] = 0.1 * torch.randn(nb, generator=gen)
    return sd


def make_mel(B, T, seed=1):
    """A mel-like input: smooth in time, roughly the range of log-mel features."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N_FEATS, T + 8, generator=gen)
    x = torch.nn.functional.avg_pool1d(x, 5, stride=1, padding=2)[..., 4:4 + T]
    return (2.0 * x - 1.0).contiguous()


# ---------------------------------------------------------------------------------------------------------------- text encoder
TEXT_ENCODER_CONFIGS = {
    # params.py:31-38 as GradTTS passes them (model/tts.py:49-51: without spk_emb_dim / n_spks, so the encoder ignores the speaker)
    "ref": dict(n_vocab=149, n_feats=80, n_channels=192, filter_channels=768, filter_channels_dp=256, n_heads=2, n_layers=6,
                kernel_size=3, p_dropout=0.1, window_size=4, spk_emb_dim=64, n_spks=1),
    # the class's own multi-speaker mode (text_encoder.py:310,327-328): speaker embedding concatenated to the prenet output
    "spk": dict(n_vocab=60, n_feats=80, n_channels=192, filter_channels=256, filter_channels_dp=128, n_heads=2, n_layers=2,
                kernel_size=3, p_dropout=0.1, window_size=4, spk_emb_dim=64, n_spks=4),
}


def text_encoder_param_shapes(cfg):
    """(name, shape) for every tensor of the reference `TextEncoder.state_dict()` (model/text_encoder.py:285-319)."""
    C0, F, Fdp, k = cfg["n_channels"], cfg["filter_channels"], cfg["filter_channels_dp"], cfg["kernel_size"]
    C = C0 + (cfg["spk_emb_dim"] if cfg["n_spks"] > 1 else 0)
    kc, w = C // cfg["n_heads"], cfg["window_size"]
    s = [("emb.weight", (cfg["n_vocab"], C0))]
    for i in range(3):
        s += [(f"prenet.conv_layers.{i}.weight", (C0, C0, 5)), (f"prenet.conv_layers.{i}.bias", (C0,))]
    for i in range(3):
        s += [(f"prenet.norm_layers.{i}.gamma", (C0,)), (f"prenet.norm_layers.{i}.beta", (C0,))]
    s += [("prenet.proj.weight", (C0, C0, 1)), ("prenet.proj.bias", (C0,))]
    for i in range(cfg["n_layers"]):
        p = f"encoder.attn_layers.{i}"
        if w is not None:
            s += [(p + ".emb_rel_k", (1, 2 * w + 1, kc)), (p + ".emb_rel_v", (1, 2 * w + 1, kc))]
        for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
            s += [(f"{p}.{c}.weight", (C, C, 1)), (f"{p}.{c}.bias", (C,))]
    for i in range(cfg["n_layers"]):
        s += [(f"encoder.norm_layers_1.{i}.gamma", (C,)), (f"encoder.norm_layers_1.{i}.beta", (C,))]
    for i in range(cfg["n_layers"]):
        s += [(f"encoder.ffn_layers.{i}.conv_1.weight", (F, C, k)), (f"encoder.ffn_layers.{i}.conv_1.bias", (F,)),
              (f"encoder.ffn_layers.{i}.conv_2.weight", (C, F, k)), (f"encoder.ffn_layers.{i}.conv_2.bias", (C,))]
    for i in range(cfg["n_layers"]):
        s += [(f"encoder.norm_layers_2.{i}.gamma", (C,)), (f"encoder.norm_layers_2.{i}.beta", (C,))]
    s += [("proj_m.weight", (cfg["n_feats"], C, 1)), ("proj_m.bias", (cfg["n_feats"],))]
    s += [("proj_w.conv_1.weight", (Fdp, C, k)), ("proj_w.conv_1.bias", (Fdp,)), ("proj_w.norm_1.gamma", (Fdp,)), ("proj_w.norm_1.beta", (Fdp,)),
          ("proj_w.conv_2.weight", (Fdp, Fdp, k)), ("proj_w.conv_2.bias", (Fdp,)), ("proj_w.norm_2.gamma", (Fdp,)), ("proj_w.norm_2.beta", (Fdp,)),
          ("proj_w.proj.weight", (1, Fdp, 1)), ("proj_w.proj.bias", (1,))]
    return s


def make_text_encoder_state_dict(cfg, seed=0):
    """Seeded weights at trained-model scale: fan-in scaled convs, norm gains around 1, duration head biased to a few frames."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in text_encoder_param_shapes(cfg):
        if name == "emb.weight":
            t = torch.randn(*shape, generator=gen) * cfg["n_channels"] ** -0.5
        elif name.endswith(".gamma"):
            t = 1.0 + 0.1 * torch.randn(*shape, generator=gen)
        elif name.endswith(".beta") or name.endswith(".bias"):
            t = 0.1 * torch.randn(*shape, generator=gen)
            if name == "proj_w.proj.bias":
                t = t + 1.0
        elif "emb_rel" in name:
            t = torch.randn(*shape, generator=gen) * shape[-1] ** -0.5
        else:
            fan_in = shape[1] * (shape[2] if len(shape) > 2 else 1)
            t = torch.randn(*shape, generator=gen) * (1.0 / math.sqrt(fan_in))
        sd[name] = t
    return sd


def make_text_inputs(cfg, B, T, seed=1, ragged=True):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randint(0, cfg["n_vocab"], (B, T), generator=gen)
    if ragged and B > 1:
        lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=gen)
        lengths[0] = T
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
    spk = torch.randn(B, cfg["spk_emb_dim"], generator=gen) if cfg["n_spks"] > 1 else None
    return x, lengths, spk
