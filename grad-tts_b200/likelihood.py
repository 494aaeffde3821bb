"""Probability-flow log-likelihood of a mel under the Grad-TTS score model, on the device
(reference n_best/likelihood/likelihood.py:41-131 with n_best/likelihood/sde_lib.py:256-297).

The reference integrates d[x, logp]/dt = [drift(x, t), div drift(x, t)] with a fixed-step loop (`euler > 0`) and a Hutchinson
estimate of the divergence; every evaluation there makes TWO estimator forwards (drift, then the same forward again under autograd)
plus an autograd backward, and flattens the state to numpy and back (likelihood.py:92-97).  Here one fused call per step returns the
score AND its vector-Jacobian product (`GradLogPEstimator2d.vjp` -> gtts_decoder_estimator_vjp) and the state never leaves the
GPU.  Same formulas, same step rule, same RNG call for the Hutchinson noise (torch.randint_like / randn_like on the data's device).
"""
import math

import torch


class SPEECHSDE:
    """The Grad-TTS forward SDE of one text/speaker pair (reference sde_lib.py:256-297): the pieces the likelihood needs."""

    def __init__(self, beta_min, beta_max, N, mu, spk, mask):
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.mu, self.speaker, self.mask = mu, spk, mask

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)          # sde_lib.py:278-282
        return 0.5 * beta_t[:, None, None] * (self.mu - x), torch.sqrt(beta_t)

    def prior_logp(self, z):
        n = math.prod(z.shape[1:])                                      # sde_lib.py:293-297
        return -n / 2.0 * math.log(2 * math.pi) - torch.sum((z - self.mu) ** 2, dim=(1, 2)) / 2.0


def _drift_and_div(sde, model, x, t, eps):
    """drift(x, t) * mask of the probability-flow ODE (likelihood.py:62-66 with sde_lib.py:93-100) and the Hutchinson estimate
    eps . d(sum(drift * eps))/dx (likelihood.py:27-38), from ONE fused estimator call."""
    mask = sde.mask
    xm = x * mask
    beta_t = sde.beta_0 + t * (sde.beta_1 - sde.beta_0)
    b = beta_t[:, None, None]
    em = eps * mask
    est = getattr(model, "estimator", None)
    if est is not None and hasattr(est, "vjp"):
        score, jtv = est.vjp(xm, model.y_mask, model.mu_y, t, em, getattr(model, "spk", None))
    else:                                                               # any other score model: two passes through autograd
        with torch.enable_grad():
            xr = xm.detach().requires_grad_(True)
            score = model(xr, t)
            jtv = torch.autograd.grad(torch.sum(score * em), xr)[0]
        score = score.detach()
    drift = (0.5 * b * (sde.mu - xm) - 0.5 * b * score) * mask
    grad_fn_eps = mask * (-0.5 * b * em - 0.5 * b * jtv)                # d sum(drift * eps) / dx
    div = torch.sum(grad_fn_eps * eps, dim=tuple(range(1, x.dim())))
    return drift, div


def get_likelihood_fn(sde, inverse_scaler=None, hutchinson_type="Rademacher", rtol=1e-5, atol=1e-5, method="RK45", eps=1e-5,
                      euler=0):
    """Same signature and return values as the reference (likelihood.py:41): likelihood_fn(model, data) ->
    (bpd, prior_logp, delta_logp, z).  `euler > 0`: that many fixed steps, entirely on the device.  `euler == 0`: scipy's
    black-box solver drives the same device functions (state crosses to the host once per evaluation, as in the reference)."""

    @torch.no_grad()
    def likelihood_fn(model, data, epsilon=None):
        if epsilon is None:
            if hutchinson_type == "Gaussian":
                epsilon = torch.randn_like(data)
            elif hutchinson_type == "Rademacher":
                epsilon = torch.randint_like(data, low=0, high=2).float() * 2 - 1.0
            else:
                raise NotImplementedError(f"Hutchinson type {hutchinson_type} unknown.")
        x = (data * sde.mask).to(torch.float32)
        B = x.shape[0]
        logp = torch.zeros(B, dtype=torch.float32, device=x.device)
        if euler > 0:
            h = 1.0 / euler                                             # likelihood.py:99-107
            for i in range(euler):
                t = torch.full((B,), (i + 0.5) * h, dtype=torch.float32, device=x.device)
                drift, div = _drift_and_div(sde, model, x, t, epsilon)
                x = x + drift * h
                logp = logp + div * h
        else:
            import numpy as np
            from scipy import integrate
            shape = x.shape

            def ode_func(tt, y):
                xs = torch.from_numpy(y[:-B].reshape(shape)).to(x.device, torch.float32)
                t = torch.full((B,), float(tt), dtype=torch.float32, device=x.device)
                drift, div = _drift_and_div(sde, model, xs, t, epsilon)
                return np.concatenate([drift.reshape(-1).cpu().numpy(), div.cpu().numpy()], axis=0)

            init = np.concatenate([x.reshape(-1).cpu().numpy(), np.zeros((B,))], axis=0)
            sol = integrate.solve_ivp(ode_func, (eps, sde.T), init, rtol=rtol, atol=atol, method=method)
            zp = sol.y[:, -1]
            x = torch.from_numpy(zp[:-B].reshape(shape)).to(x.device, torch.float32)
            logp = torch.from_numpy(zp[-B:]).to(x.device, torch.float32)
        prior_logp = sde.prior_logp(x)
        bpd = -(prior_logp + logp)
        return bpd, prior_logp, logp, x

    return likelihood_fn
