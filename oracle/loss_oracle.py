"""CPU oracle for the forward value of the training objective (TEST INFRASTRUCTURE ONLY).

Restates `Diffusion.forward_diffusion` (/root/reference/model/diffusion.py:244-252) with the N(0,1) draw supplied by the
caller, and the scalar of `Diffusion.loss_t` (:274-281).  Pinned against the reference by `tests/golden/make_golden.py loss`
(vectors captured from the reference's own `loss_t`: the noise it drew, `xt`, the estimator output and the loss) and replayed by
`tests/test_oracle_golden.py`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this module.
"""
import torch

from . import decoder_oracle


def cum_noise(t, beta_min=0.05, beta_max=20.0):
    # get_noise(..., cumulative=True), model/diffusion.py:219-222
    return beta_min * t + 0.5 * (beta_max - beta_min) * (t ** 2)


def forward_diffusion(x0, mask, mu, t, noise, beta_min=0.05, beta_max=20.0):
    """model/diffusion.py:244-252 with `z = noise` instead of torch.randn.  Returns (xt * mask, z * mask)."""
    c = cum_noise(t.unsqueeze(-1).unsqueeze(-1), beta_min, beta_max)
    mean = x0 * torch.exp(-0.5 * c) + mu * (1.0 - torch.exp(-0.5 * c))
    variance = 1.0 - torch.exp(-c)
    xt = mean + noise * torch.sqrt(variance)
    return xt * mask, noise * mask


def score_loss(noise_estimation, z_masked, mask, t, n_feats=80, beta_min=0.05, beta_max=20.0):
    """model/diffusion.py:276-280 given the estimator output."""
    c = cum_noise(t.unsqueeze(-1).unsqueeze(-1), beta_min, beta_max)
    e = noise_estimation * torch.sqrt(1.0 - torch.exp(-c))
    return torch.sum((e + z_masked) ** 2) / (torch.sum(mask) * n_feats)


def loss_t(sd, x0, mask, mu, t, noise, spk=None, n_spks=1, beta_min=0.05, beta_max=20.0, pe_scale=1000.0):
    """Diffusion.loss_t (model/diffusion.py:274-281) -> (loss, xt)."""
    xt, zm = forward_diffusion(x0, mask, mu, t, noise, beta_min, beta_max)
    est = decoder_oracle.estimator_forward(sd, xt, mask, mu, t, spk, n_spks, pe_scale)
    return score_loss(est, zm, mask, t, x0.shape[1], beta_min, beta_max), xt


def loss_t_grads(sd, x0, mask, mu, t, noise, spk=None, n_spks=1):
    """loss_t with torch.autograd: (loss, {name: d loss / d parameter}, d loss / d mu, d loss / d spk) -- what loss.backward() leaves in
    .grad in the reference's training step (train.py -> GradTTS.compute_loss -> Diffusion.compute_loss -> loss_t)."""
    with torch.enable_grad():
        sdg = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        mug = mu.detach().clone().requires_grad_(True)
        spkg = spk.detach().clone().requires_grad_(True) if spk is not None else None
        loss, _ = loss_t(sdg, x0, mask, mug, t, noise, spkg, n_spks)
        names = [k for k in sdg]
        grads = torch.autograd.grad(loss, [sdg[k] for k in names] + [mug] + ([spkg] if spkg is not None else []), allow_unused=True)
    out = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads[:len(names)])}
    return loss.detach(), out, grads[len(names)], (grads[len(names) + 1] if spkg is not None else None)
