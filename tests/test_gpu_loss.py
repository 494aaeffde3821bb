"""GPU: forward value of the training objective -- forward-diffusion kernel, estimator, loss reduction --
against vectors captured from the reference's own Diffusion.loss_t and against the oracle restatement.

Tolerances: forward diffusion fp32 elementwise (expf/sqrtf vs torch) 2e-6 * max|xt|; loss reduction given the reference's
estimator output 1e-6 relative (double accumulation vs torch's fp32 sum); whole loss_t 1e-4 relative in fp32 mode and 2e-2
relative in bf16 mode (the estimator's tolerance)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import loss_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["loss_spk1_b2_t48", "loss_spk247_b3_t40"]


def _module(pkg, synth, n_spks, wseed, precision):
    sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(DEV)
    dec.precision = precision
    return dec, sd


def _score_loss(pkg, est, zm, mask, t):
    lib = importlib.import_module("grad-tts_b200._lib").load()
    ws = torch.empty(int(lib.gtts_score_loss_workspace_bytes()), dtype=torch.uint8, device=DEV)
    out = torch.empty((), dtype=torch.float32, device=DEV)
    B, C, T = est.shape
    rc = lib.gtts_score_loss(est.data_ptr(), zm.data_ptr(), mask.data_ptr(), t.data_ptr(), ws.data_ptr(), ws.numel(),
                             out.data_ptr(), B, C, T, 0.05, 20.0, None)
    assert rc == 0
    torch.cuda.synchronize()
    return float(out)


@pytest.mark.parametrize("name", CASES)
def test_forward_diffusion_and_loss_match_reference_capture(pkg, synth, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n_spks = int(g["n_spks"])
    x0, mask, mu, t, zm, est = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "mask", "mu", "t", "zm", "est"))
    spk = torch.from_numpy(g["spk"]).to(DEV) if "spk" in g else None
    ref_xt, ref_loss = torch.from_numpy(g["xt"]), float(g["loss"])
    dec, _ = _module(pkg, synth, n_spks, int(g["wseed"]), "fp32")
    x0c, muc = x0.clone(), mu.clone()
    xt, zm2 = dec.forward_diffusion(x0, mask, mu, t, noise=zm)
    assert torch.equal(x0, x0c) and torch.equal(mu, muc)                                  # inputs untouched
    assert float((xt.cpu() - ref_xt).abs().max()) <= 2e-6 * float(ref_xt.abs().max())
    assert torch.equal(zm2, zm)                                                            # z * mask is exact
    assert float((xt * (1 - mask)).abs().max()) == 0.0
    # the reduction alone, fed with the estimator output the reference saw
    got = _score_loss(pkg, est, zm, mask.reshape(mask.shape[0], -1).contiguous(), t)
    assert abs(got - ref_loss) <= 1e-6 * ref_loss
    assert got == _score_loss(pkg, est, zm, mask.reshape(mask.shape[0], -1).contiguous(), t)   # fixed-shape reduction: same bits
    # the whole loss_t
    with torch.no_grad():
        loss, xt2 = dec.loss_t(x0, mask, mu, t, spk, noise=zm)
        assert torch.equal(xt2, xt)
        assert abs(float(loss) - ref_loss) <= 1e-4 * ref_loss, (float(loss), ref_loss)
        dec.precision = "bf16"
        loss_bf, _ = dec.loss_t(x0, mask, mu, t, spk, noise=zm)
        assert abs(float(loss_bf) - ref_loss) <= 2e-2 * ref_loss, (float(loss_bf), ref_loss)


def test_loss_is_differentiable_and_draws_its_own_noise(pkg, synth):
    dec, sd = _module(pkg, synth, 1, 0, "fp32")
    x0, mask, mu, _, _ = synth.make_inputs(2, 64, 1, seed=5)
    x0, mask, mu = x0.to(DEV), mask.to(DEV), mu.to(DEV)
    t = torch.tensor([0.3, 0.9], device=DEV)
    dec.train()
    loss, _ = dec.loss_t(x0, mask, mu, t)                 # gradients enabled, train mode: the loss carries the autograd graph
    assert loss.requires_grad
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dec.parameters())
    dec.eval()
    with torch.no_grad():
        torch.manual_seed(11)
        a, xt_a = dec.loss_t(x0, mask, mu, t)
        torch.manual_seed(11)
        noise = torch.randn(x0.shape, dtype=torch.float32, device=DEV)
        b, xt_b = dec.loss_t(x0, mask, mu, t, noise=noise)
        assert torch.equal(xt_a, xt_b) and float(a) == float(b)      # the default draw is torch.randn on the device (:249)
        ref, _ = loss_oracle.loss_t(sd, x0.cpu(), mask.cpu(), mu.cpu(), t.cpu(), noise.cpu())
        assert abs(float(a) - float(ref)) <= 1e-4 * float(ref)
        torch.manual_seed(12)
        c, _ = dec.compute_loss(x0, mask, mu)                         # :283-287: t ~ U(offset, 1 - offset), then loss_t
        assert np.isfinite(float(c)) and float(c) > 0


def test_large_ragged_loss_against_oracle(pkg):
    """Reduction at a size where every CTA has work (64 x 80 x 1000) and a ragged mask: vs the oracle formula in float64."""
    gen = torch.Generator().manual_seed(3)
    B, C, T = 64, 80, 1000
    est, z = torch.randn(B, C, T, generator=gen), torch.randn(B, C, T, generator=gen)
    lengths = torch.randint(600, T + 1, (B,), generator=gen)
    mask = (torch.arange(T)[None] < lengths[:, None]).float()
    t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
    zm = z * mask[:, None]
    ref = loss_oracle.score_loss((est * mask[:, None]).double(), zm.double(), mask[:, None].double(), t.double())
    got = _score_loss(pkg, (est * mask[:, None]).to(DEV), zm.to(DEV), mask.to(DEV), t.to(DEV))
    assert abs(got - float(ref)) <= 2e-6 * float(ref)


def test_gradtts_compute_loss_forward_values(pkg, synth):
    """GradTTS.compute_loss (tts.py:110-194) forward values with a stub encoder replaying the reference capture: duration loss and
    prior loss against the torch formulas on the captured logw_/mu_y, diffusion loss against the oracle with the same draws."""
    g = np.load(os.path.join(GOLDEN, "align_b2_50x200.npz"))
    mu_x, x_mask, y = (torch.from_numpy(g[k]) for k in ("mu_x", "x_mask", "y"))
    x_len, y_len = torch.from_numpy(g["x_len"]), torch.from_numpy(g["y_len"])
    B, _, tx = mu_x.shape
    logw = torch.randn(B, 1, tx, generator=torch.Generator().manual_seed(1)) * x_mask

    class Enc(torch.nn.Module):
        def forward(self, x, x_lengths, spk=None):
            return mu_x.to(DEV), logw.to(DEV), x_mask.to(DEV)

    net = pkg.GradTTS(60, 1, 64, 192, 768, 256, 2, 2, 3, 0.0, 4, 80, 64, 0.05, 20.0, 1000, encoder=Enc())
    sd = synth.make_decoder_state_dict(1, seed=0, g=0.05)
    net.decoder.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    net.decoder.precision = "fp32"
    x = torch.zeros(B, tx, dtype=torch.long)
    net.train()
    _, _, diff_t = net.compute_loss(x, x_len, y, y_len)                  # train mode: differentiable down to the decoder parameters
    assert diff_t.requires_grad
    diff_t.backward()
    assert all(p.grad is not None for p in net.decoder.parameters())
    net.eval()
    with torch.no_grad():
        torch.manual_seed(21)
        dur, prior, diff = net.compute_loss(x.to(DEV), x_len.to(DEV), y.to(DEV), y_len.to(DEV))
        torch.manual_seed(21)                                             # replay the two draws: t (diffusion.py:284), z (:249)
        t = torch.rand(B, dtype=torch.float32, device=DEV).clamp(1e-5, 1 - 1e-5).cpu()
        noise = torch.randn(B, 80, y.shape[-1], dtype=torch.float32, device=DEV).cpu()
    y_mask = (torch.arange(y.shape[-1])[None] < y_len[:, None]).float().unsqueeze(1)
    logw_, mu_y = torch.from_numpy(g["logw_"]), torch.from_numpy(g["mu_y"])
    ref_dur = torch.sum((logw - logw_) ** 2) / torch.sum(x_len)                        # model/utils.py duration_loss
    ref_prior = torch.sum(0.5 * ((y - mu_y) ** 2 + np.log(2 * np.pi)) * y_mask) / (torch.sum(y_mask) * 80)
    with torch.no_grad():
        ref_diff, _ = loss_oracle.loss_t(sd, y, y_mask, mu_y, t, noise)
    assert abs(float(dur) - float(ref_dur)) <= 1e-5 * float(ref_dur)
    assert abs(float(prior) - float(ref_prior)) <= 1e-5 * float(ref_prior)
    assert abs(float(diff) - float(ref_diff)) <= 1e-4 * float(ref_diff), (float(diff), float(ref_diff))
