"""CPU restatement of the probability-flow likelihood and of the score network's input gradient (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/n_best/likelihood/likelihood.py:27-38 (Hutchinson divergence through torch.autograd), :62-66 (drift of the
probability-flow ODE), :99-107,113-131 (fixed-step loop, prior, bpd) and /root/reference/n_best/likelihood/sde_lib.py:93-100
(reverse SDE drift), :278-297 (SPEECHSDE.sde / prior_logp), on top of oracle/decoder_oracle.py.  Pinned against vectors produced
by the real reference code (tests/golden/make_golden.py vjp -> vjp_*.npz, lik_*.npz; tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import math

import torch

from . import decoder_oracle


def estimator_vjp(sd, x, mask, mu, t, v, spk=None, n_spks=1):
    """(score, d sum(score * v) / dx) with autograd, likelihood.py:30-34."""
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        score = decoder_oracle.estimator_forward(sd, xr, mask, mu, t, spk, n_spks)
        gx = torch.autograd.grad(torch.sum(score * v), xr)[0]
    return score.detach(), gx


def likelihood(sd, data, mask, mu, n_euler, eps, spk=None, n_spks=1, beta_min=0.05, beta_max=20.0):
    """likelihood_fn(model, data) of get_likelihood_fn(sde, ..., euler=n_euler) with the Hutchinson noise `eps` given
    -> (bpd, prior_logp, delta_logp, z)."""
    def drift_fn(x, t):                                              # likelihood.py:62-66 + sde_lib.py:93-100, 278-282
        x = x * mask
        beta_t = beta_min + t * (beta_max - beta_min)
        drift = 0.5 * beta_t[:, None, None] * (mu - x)
        score = decoder_oracle.estimator_forward(sd, x, mask, mu, t, spk, n_spks)
        drift = drift - torch.sqrt(beta_t)[:, None, None] ** 2 * score * 0.5
        return drift * mask

    def div_fn(x, t):                                                # likelihood.py:27-38
        with torch.enable_grad():
            xr = x.detach().clone().requires_grad_(True)
            fn_eps = torch.sum(drift_fn(xr, t) * eps)
            g = torch.autograd.grad(fn_eps, xr)[0]
        return torch.sum(g * eps, dim=(1, 2))

    x = data * mask                                                  # likelihood.py:110
    B = x.shape[0]
    logp = torch.zeros(B, device=x.device)
    h = 1.0 / n_euler
    for i in range(n_euler):                                         # likelihood.py:99-107
        t = torch.ones(B, device=x.device) * ((i + 0.5) * h)
        with torch.no_grad():
            d = drift_fn(x, t)
        dv = div_fn(x, t)
        x = x + d * h
        logp = logp + dv * h
    n = math.prod(x.shape[1:])
    prior_logp = -n / 2.0 * math.log(2 * math.pi) - torch.sum((x - mu) ** 2, dim=(1, 2)) / 2.0     # sde_lib.py:293-297
    return -(prior_logp + logp), prior_logp, logp, x
