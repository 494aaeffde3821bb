"""GPU: the score network's gradient w.r.t. its input (vector-Jacobian product) and the probability-flow likelihood built on it,
against vectors produced by the REAL reference (torch.autograd through model/diffusion.py; n_best/likelihood/likelihood.py with
SPEECHSDE) and against the CPU oracle.

Tolerances: fp32 mode max-abs 1e-4 on |gx|max ~ 1.2-1.5 (single call, like the forward's 1e-4); bf16 mode rel-rms 5e-2 (bf16
activations AND bf16 gradients through ~50 layers).  Likelihood (3 Euler steps, fp32): delta_logp / bpd within 1e-4 relative.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import likelihood_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _module(pkg, synth, n_spks, wseed, precision):
    sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV).eval()                                          # the likelihood code runs the model in eval mode (train mode: test_gpu_train.py)
    dec.precision = precision
    return dec, sd


def _g(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return g, (lambda k: torch.from_numpy(g[k]).to(DEV) if k in g else None)


@pytest.mark.parametrize("name", ["vjp_spk1_b2_t48", "vjp_spk247_b2_t40"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vjp_matches_reference_autograd(name, precision, pkg, synth):
    g, t = _g(name)
    dec, _ = _module(pkg, synth, int(g["n_spks"]), int(g["wseed"]), precision)
    score, gx = dec.estimator.vjp(t("x"), t("mask"), t("mu"), t("t"), t("v"), t("spk"))
    ref_s, ref_g = t("score"), t("gx")
    assert torch.isfinite(gx).all()
    assert float((gx * (1 - t("mask"))).abs().max()) == 0.0           # no gradient flows into padded frames
    if precision == "fp32":
        assert float((score - ref_s).abs().max()) <= 1e-4
        assert float((gx - ref_g).abs().max()) <= 1e-4, float((gx - ref_g).abs().max())
    else:
        rr = lambda a, b: float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())
        assert rr(score, ref_s) <= 2e-2 and rr(gx, ref_g) <= 5e-2, (rr(score, ref_s), rr(gx, ref_g))
    # the same score as the plain forward, bit for bit (the VJP plan runs the same forward kernels)
    if precision == "fp32":
        assert torch.equal(score, dec.estimator(t("x"), t("mask"), t("mu"), t("t"), t("spk")))


def test_autograd_through_the_module_equals_vjp(pkg, synth):
    """torch.autograd.grad(sum(estimator(x) * eps), x) -- the reference's own call (likelihood.py:30-34) -- works on the drop-in module
    (registered torch.library autograd) and gives exactly what the fused vjp call gives."""
    g, t = _g("vjp_spk1_b2_t48")
    dec, _ = _module(pkg, synth, 1, int(g["wseed"]), "fp32")
    x = t("x").clone().requires_grad_(True)
    with torch.enable_grad():
        fn_eps = torch.sum(dec.estimator(x, t("mask"), t("mu"), t("t")) * t("v"))
        gx = torch.autograd.grad(fn_eps, x)[0]
    _, gx2 = dec.estimator.vjp(t("x"), t("mask"), t("mu"), t("t"), t("v"))
    assert torch.equal(gx, gx2)
    assert float((gx - t("gx")).abs().max()) <= 1e-4
    y = dec.estimator(t("x"), t("mask"), t("mu"), t("t"))             # no grad requested: no graph
    assert not y.requires_grad


def test_vjp_chunks_and_edge_masks_against_oracle(pkg, synth):
    """More samples than one VJP workspace chunk (16), an empty and a one-frame utterance, per-sample t: against the CPU oracle's
    autograd; chunked == unchunked bit for bit."""
    B, T = 19, 24
    dec, sd = _module(pkg, synth, 1, 0, "fp32")
    gen = torch.Generator().manual_seed(5)
    x, mu, v = (torch.randn(B, 80, T, generator=gen) for _ in range(3))
    lengths = torch.randint(8, T + 1, (B,), generator=gen)
    lengths[0], lengths[1], lengths[2] = T, 0, 1
    mask = (torch.arange(T)[None] < lengths[:, None]).float().unsqueeze(1)
    tt = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
    torch.set_num_threads(8)
    ref_s, ref_g = likelihood_oracle.estimator_vjp(sd, x * mask, mask, mu, tt, v)
    a = [u.to(DEV) for u in (x * mask, mask, mu, tt, v)]
    score, gx = dec.estimator.vjp(*a)
    assert float((score.cpu() - ref_s).abs().max()) <= 1e-4 and float((gx.cpu() - ref_g).abs().max()) <= 1e-4
    one = dec.estimator.vjp(*(u[17:18] for u in a))
    assert torch.equal(one[1], gx[17:18])


class _ScoreModel(torch.nn.Module):                                   # what GradTTS.get_score_model returns (model/tts.py:237-250)
    def __init__(self, estimator, y_mask, mu_y, spk=None):
        super().__init__()
        self.estimator, self.y_mask, self.mu_y, self.spk = estimator, y_mask, mu_y, spk

    def forward(self, x, t):
        return self.estimator(x=x, mask=self.y_mask, mu=self.mu_y, t=t, spk=self.spk)


def test_likelihood_matches_reference(pkg, synth):
    """get_likelihood_fn(SPEECHSDE, euler=3) on the device against the values the reference's own likelihood code produced."""
    g, t = _g("lik_spk1_b2_t48_e3")
    dec, _ = _module(pkg, synth, 1, int(g["wseed"]), "fp32")
    lik = pkg.likelihood
    sde = lik.SPEECHSDE(beta_min=0.05, beta_max=20.0, N=1000, mu=t("mu"), spk=None, mask=t("mask"))
    fn = lik.get_likelihood_fn(sde, lambda x: x, rtol=1e-3, atol=1e-3, euler=int(g["n_euler"]))
    model = _ScoreModel(dec.estimator, t("mask"), t("mu"))
    bpd, prior, dlogp, z = fn(model, t("y"), epsilon=t("eps"))
    assert float((z - t("z")).abs().max()) <= 1e-3
    for got, key in ((bpd, "bpd"), (prior, "prior_logp"), (dlogp, "delta_logp")):
        assert torch.allclose(got, t(key), rtol=1e-4, atol=5e-2), (key, got.tolist(), t(key).tolist())

    # a score model that is just a callable (no .estimator): the generic path differentiates it with torch.autograd, like the reference
    class Plain(torch.nn.Module):
        def forward(self, x, tt):
            return dec.estimator(x, t("mask"), t("mu"), tt)
    bpd2, _, dlogp2, z2 = fn(Plain(), t("y"), epsilon=t("eps"))
    assert torch.allclose(dlogp2, dlogp, rtol=1e-6, atol=1e-3) and torch.equal(z2, z)
    # the Hutchinson noise is drawn like the reference does when none is supplied
    torch.manual_seed(3)
    out = fn(model, t("y"))
    assert all(torch.isfinite(o).all() for o in out)
