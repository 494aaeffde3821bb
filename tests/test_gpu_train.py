"""GPU: training backward of the decoder -- loss.backward() through Diffusion.loss_t on the drop-in modules against the gradients the
REAL reference produced (tests/golden/grad_*.npz: torch.autograd through model/diffusion.py in train mode) and against the CPU oracle's
autograd for every one of the 172 / 176 parameter tensors.

Tolerances, per tensor, relative to the tensor's own max-abs gradient: fp32 mode 2e-3 (fp32 reductions over up to 10^5 pixels in a
different order than ATen), bf16 mode rel-rms 8e-2 on tensors with more than 256 entries (bf16 activations and bf16 gradients).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import loss_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _module(pkg, synth, n_spks, wseed, precision):
    sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV)
    dec.precision = precision
    return dec, sd


def _run(pkg, synth, name, precision):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = lambda k: torch.from_numpy(g[k]).to(DEV) if k in g.files else None
    n_spks = int(g["n_spks"])
    dec, sd = _module(pkg, synth, n_spks, int(g["wseed"]), precision)
    dec.train()
    mu = t("mu").clone().requires_grad_(True)
    spk = t("spk").clone().requires_grad_(True) if "spk" in g.files else None
    loss, xt = dec.loss_t(t("x0"), t("mask"), mu, t("t"), spk, noise=t("zm"))
    loss.backward()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in dec.named_parameters()}
    return g, sd, float(loss), grads, mu.grad.cpu(), (spk.grad.cpu() if spk is not None else None)


@pytest.mark.parametrize("name", ["grad_spk1_b2_t48", "grad_spk247_b2_t40"])
def test_training_backward_fp32_matches_reference(name, pkg, synth):
    g, sd, loss, grads, gmu, gspk = _run(pkg, synth, name, "fp32")
    assert abs(loss - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"])))
    ref_gmu = torch.from_numpy(g["gmu"])
    assert float((gmu - ref_gmu).abs().max()) <= 2e-3 * float(ref_gmu.abs().max())
    if gspk is not None:
        ref = torch.from_numpy(g["gspk"])
        assert float((gspk - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    # every parameter: the reference's digest (sum, abs-sum, first 16 entries)
    dig = torch.from_numpy(g["grad_digest"])
    for i, k in enumerate([str(s) for s in g["grad_names"]]):
        got = grads[k]
        assert got is not None, f"{k}: no gradient"
        flat = got.reshape(-1)
        scale = max(float(flat.abs().max()), 1e-12)
        n = min(16, flat.numel())
        assert float((flat[:n] - dig[i, 2:2 + n]).abs().max()) <= 2e-3 * scale + 1e-9, k
        assert abs(float(flat.double().sum()) - float(dig[i, 0])) <= 2e-3 * max(float(dig[i, 1]), 1e-9), k
    for k in g.files:
        if k.startswith("full:"):
            ref = torch.from_numpy(g[k])
            assert float((grads[k[5:]] - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), k
    # ... and the CPU oracle's autograd, all entries of all tensors
    torch.set_num_threads(8)
    t = lambda kk: torch.from_numpy(g[kk])
    _, ograds, _, _ = loss_oracle.loss_t_grads(sd, t("x0"), t("mask"), t("mu"), t("t"), t("zm"), t("spk") if "spk" in g.files else None, int(g["n_spks"]))
    worst = ("", 0.0)
    for k, ref in ograds.items():
        err = float((grads[k] - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] <= 2e-3, worst


def test_training_backward_bf16_close_to_reference(pkg, synth):
    g, sd, loss, grads, gmu, _ = _run(pkg, synth, "grad_spk1_b2_t48", "bf16")
    assert abs(loss - float(g["loss"])) <= 2e-2 * abs(float(g["loss"]))
    torch.set_num_threads(8)
    t = lambda kk: torch.from_numpy(g[kk])
    _, ograds, ogmu, _ = loss_oracle.loss_t_grads(sd, t("x0"), t("mask"), t("mu"), t("t"), t("zm"), None, 1)
    rr = lambda a, b: float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-20))
    assert rr(gmu, ogmu) <= 8e-2
    bad = [(k, rr(grads[k], ref)) for k, ref in ograds.items() if ref.numel() > 256 and rr(grads[k], ref) > 8e-2]
    assert not bad, bad[:5]


def test_training_step_runs_and_likelihood_path_unaffected(pkg, synth):
    """One optimizer step in train mode changes the parameters and the next forward sees them; in eval mode the same module still
    gives the x-only autograd of the likelihood code."""
    dec, _ = _module(pkg, synth, 1, 0, "bf16")
    dec.train()
    opt = torch.optim.SGD(dec.parameters(), lr=1e-3)
    x0, mask, mu, _, _ = synth.make_inputs(18, 24, 1, seed=5)             # 18 > one backward workspace chunk of 16
    a = [v.to(DEV) for v in (x0, mask, mu)]
    tt = torch.rand(18, generator=torch.Generator().manual_seed(1)).clamp(1e-5, 1 - 1e-5).to(DEV)
    z = torch.randn(18, 80, 24, generator=torch.Generator().manual_seed(2)).to(DEV)
    l0, _ = dec.loss_t(a[0], a[1], a[2], tt, noise=z)
    opt.zero_grad()
    l0.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dec.parameters())
    opt.step()
    l1, _ = dec.loss_t(a[0], a[1], a[2], tt, noise=z)
    assert torch.isfinite(l1) and float(l1) != float(l0)
    dec.eval()
    x = a[0].clone().requires_grad_(True)
    gx = torch.autograd.grad(dec.estimator(x, a[1], a[2], tt).sum(), x)[0]
    assert torch.isfinite(gx).all() and all(p.grad is not None for p in dec.parameters())
