"""CPU oracle for the Grad-TTS reverse-diffusion decoder (TEST INFRASTRUCTURE ONLY).

This file is a plain-PyTorch, CPU, functional restatement of the reference algorithm in
`/root/reference/model/diffusion.py`.  It exists so that the CUDA path can be checked against
something that travels to the GPU box (the reference itself does not).  It is pinned against
the reference by `tests/golden/make_golden.py`, which imports the real reference in the build
container, runs it on seeded inputs and commits the outputs under `tests/golden/`;
`tests/test_oracle_golden.py` replays those fixtures through this file.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product (`grad-tts_b200/`) never does.

Every function cites the reference lines it follows.  Weights come in as a flat `state_dict`
with the reference's own key names relative to `Diffusion` (i.e. `estimator.downs.0.0...`).
"""
import math

import torch
import torch.nn.functional as F


def mish(x):
    # model/diffusion.py:16-18  x * tanh(softplus(x)), torch softplus(beta=1, threshold=20)
    return x * torch.tanh(F.softplus(x))


def _block(sd, p, x, mask, groups=8):
    # model/diffusion.py:49-58  Conv3x3(x*mask) -> GroupNorm(8) -> Mish -> *mask
    h = F.conv2d(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    h = F.group_norm(h, groups, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], eps=1e-5)
    return mish(h) * mask


def _resnet(sd, p, x, mask, temb):
    # model/diffusion.py:61-79
    h = _block(sd, p + ".block1", x, mask)
    tb = F.linear(mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
    h = h + tb[:, :, None, None]
    h = _block(sd, p + ".block2", h, mask)
    if (p + ".res_conv.weight") in sd:
        r = F.conv2d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    else:
        r = x * mask
    return h + r


def _linear_attention(sd, p, x, heads=4, dim_head=32):
    # model/diffusion.py:82-100 (einops rearranges restated with view/permute)
    b, c, h, w = x.shape
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"])                      # (b, 3*heads*dh, h, w)
    qkv = qkv.view(b, 3, heads, dim_head, h * w)                     # 'b (qkv heads c) h w'
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b, heads * dim_head, h, w)
    return F.conv2d(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def _attn_residual(sd, p, x):
    # Residual(Rezero(LinearAttention)):  fn(x) * g + x   (model/diffusion.py:39-46,103-110)
    return _linear_attention(sd, p + ".fn.fn", x) * sd[p + ".fn.g"] + x


def sinusoidal_pos_emb(t, dim=64, scale=1000.0):
    # model/diffusion.py:113-125
    half = dim // 2
    c = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device).to(t.dtype) * -c)
    emb = scale * t.unsqueeze(1) * emb.unsqueeze(0)
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def estimator_forward(sd, x, mask, mu, t, spk=None, n_spks=1, pe_scale=1000.0, pfx="estimator."):
    """GradLogPEstimator2d.forward (model/diffusion.py:174-216). x, mu: (B,80,T); mask (B,1,T); t (B,)."""
    sd = {k[len(pfx):]: v for k, v in sd.items() if k.startswith(pfx)}
    if spk is not None:
        s = F.linear(spk, sd["spk_mlp.0.weight"], sd["spk_mlp.0.bias"])
        s = F.linear(mish(s), sd["spk_mlp.2.weight"], sd["spk_mlp.2.bias"])
    temb = sinusoidal_pos_emb(t, 64, pe_scale)
    temb = F.linear(temb, sd["mlp.0.weight"], sd["mlp.0.bias"])
    temb = F.linear(mish(temb), sd["mlp.2.weight"], sd["mlp.2.bias"])

    if n_spks < 2:                                                  # :180-184 channel order [mu, x, (s)]
        h = torch.stack([mu, x], 1)
    else:
        s = s.unsqueeze(-1).repeat(1, 1, x.shape[-1])
        h = torch.stack([mu, x, s], 1)
    mask = mask.unsqueeze(1)

    hiddens = []
    masks = [mask]
    for lvl in range(3):                                            # :189-196
        m = masks[-1]
        h = _resnet(sd, f"downs.{lvl}.0", h, m, temb)
        h = _resnet(sd, f"downs.{lvl}.1", h, m, temb)
        h = _attn_residual(sd, f"downs.{lvl}.2", h)
        hiddens.append(h)
        h = h * m
        if lvl < 2:
            h = F.conv2d(h, sd[f"downs.{lvl}.3.conv.weight"], sd[f"downs.{lvl}.3.conv.bias"],
                         stride=2, padding=1)
        masks.append(m[:, :, :, ::2])

    masks = masks[:-1]                                              # :199-203
    m = masks[-1]
    h = _resnet(sd, "mid_block1", h, m, temb)
    h = _attn_residual(sd, "mid_attn", h)
    h = _resnet(sd, "mid_block2", h, m, temb)

    for u in range(2):                                              # :205-211
        m = masks.pop()
        h = torch.cat((h, hiddens.pop()), dim=1)
        h = _resnet(sd, f"ups.{u}.0", h, m, temb)
        h = _resnet(sd, f"ups.{u}.1", h, m, temb)
        h = _attn_residual(sd, f"ups.{u}.2", h)
        h = F.conv_transpose2d(h * m, sd[f"ups.{u}.3.conv.weight"], sd[f"ups.{u}.3.conv.bias"],
                               stride=2, padding=1)
    h = _block(sd, "final_block", h, mask)                          # :212-216
    out = F.conv2d(h * mask, sd["final_conv.weight"], sd["final_conv.bias"])
    return (out * mask).squeeze(1)


def reverse_diffusion(sd, z, mask, mu, n_timesteps, stoc=False, spk=None, n_spks=1,
                      beta_min=0.05, beta_max=20.0, pe_scale=1000.0, sde_noise=None):
    """Diffusion.reverse_diffusion (model/diffusion.py:254-268).

    `stoc` is accepted and ignored exactly like the reference fork (it never reads the flag).
    `sde_noise` (n_timesteps,B,80,T) switches on the *upstream* (huawei-noah Grad-TTS) stochastic
    branch, which this fork deleted: dxt_det = (0.5*(mu-xt) - est)*noise_t*h, dxt_stoc =
    z*sqrt(noise_t*h), xt = (xt - (dxt_det + dxt_stoc))*mask -- the injected noise is SUBTRACTED.
    BASELINE.json's north-star line writes "+ sqrt(beta*h)*z", which is the same update with -z.
    This branch is NOT pinned by the reference (it has no such code): parity unpinned for it only.
    """
    h = 1.0 / n_timesteps
    xt = z * mask
    for i in range(n_timesteps):
        t = (1.0 - (i + 0.5) * h) * torch.ones(z.shape[0], dtype=z.dtype, device=z.device)
        time = t.unsqueeze(-1).unsqueeze(-1)
        noise_t = beta_min + (beta_max - beta_min) * time           # get_noise, :219-224
        est = estimator_forward(sd, xt, mask, mu, t, spk, n_spks, pe_scale)
        if sde_noise is None:
            dxt = 0.5 * (mu - xt - est)
            dxt = dxt * noise_t * h
        else:
            dxt_det = (0.5 * (mu - xt) - est) * noise_t * h
            dxt_stoc = sde_noise[i] * torch.sqrt(noise_t * h)
            dxt = dxt_det + dxt_stoc
        xt = (xt - dxt) * mask
    return xt


def cast_state_dict(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}
