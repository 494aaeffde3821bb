"""GPU: alignment stage around MAS -- log-prior kernel, MAS, durations and aligned means against the
vectors captured from the reference's own GradTTS.compute_loss and against the oracle restatement."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import align_oracle, mas_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("grad-tts_b200")


@pytest.mark.parametrize("name", ["align_b3_17x61", "align_b2_50x200"])
def test_alignment_stage_matches_reference_capture(pkg, name):
    al = importlib.import_module("grad-tts_b200.model.align")
    ma = importlib.import_module("grad-tts_b200.model.monotonic_align")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    mu_x, y = torch.from_numpy(g["mu_x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    x_mask, mask = torch.from_numpy(g["x_mask"]).to(DEV), torch.from_numpy(g["mask"]).to(DEV)
    ref_lp = torch.from_numpy(g["log_prior"])
    lp = al.log_prior(mu_x, y)
    # fp32, 80-term sums of squares up to ~170: summation order differs from MKL -> 1e-5 relative to the tensor scale
    assert float((lp.cpu() - ref_lp).abs().max()) <= 1e-5 * float(ref_lp.abs().max())
    attn = ma.maximum_path(lp, mask)
    assert torch.equal(attn.cpu().to(torch.int8), torch.from_numpy(g["attn"]))     # same path as the reference
    logw = al.logw_from_path(attn, x_mask)
    assert float((logw.cpu() - torch.from_numpy(g["logw_"])).abs().max()) <= 2e-6  # logf vs torch.log
    mu_y = al.mu_y_from_path(attn, mu_x)
    assert torch.equal(mu_y.cpu(), torch.from_numpy(g["mu_y"]))                    # 0/1 path: a gather, exact


def test_log_prior_large_and_ragged_against_oracle(pkg):
    """BASELINE C2 shape (64 x 200 x 1000) and an odd shape: kernel vs torch-CPU restatement, and linearity in the constant."""
    al = importlib.import_module("grad-tts_b200.model.align")
    for B, tx, ty, seed in [(64, 200, 1000, 1), (3, 65, 129, 2), (1, 1, 4, 3)]:
        gen = torch.Generator().manual_seed(seed)
        mu_x, y = torch.randn(B, 80, tx, generator=gen), torch.randn(B, 80, ty, generator=gen) * 1.5
        ref = align_oracle.log_prior(mu_x, y, 80)
        got = al.log_prior(mu_x.to(DEV), y.to(DEV)).cpu()
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # property: log N(y; mu, I) is maximal (= const) when y == mu
    m = torch.randn(2, 80, 33, generator=torch.Generator().manual_seed(5))
    d = al.log_prior(m.to(DEV), m.to(DEV)).cpu()
    const = -0.5 * np.log(2 * np.pi) * 80
    assert float((torch.diagonal(d, dim1=1, dim2=2) - const).abs().max()) <= 1e-3
    assert float(d.max()) <= const + 1e-3


def test_gradtts_align_uses_device_stage(pkg):
    """GradTTS.align / align_outputs end to end on random encoder outputs vs the oracle chain."""
    net = pkg.GradTTS(30, 1, 64, 192, 768, 256, 2, 2, 3, 0.0, 4, 80, 64, 0.05, 20.0, 1000, encoder=torch.nn.Identity())
    gen = torch.Generator().manual_seed(9)
    B, tx, ty = 4, 23, 90
    mu_x, y = torch.randn(B, 80, tx, generator=gen), torch.randn(B, 80, ty, generator=gen)
    xl, yl = torch.tensor([23, 11, 17, 5]), torch.tensor([90, 40, 77, 64])
    x_mask = (torch.arange(tx)[None] < xl[:, None]).float().unsqueeze(1)
    y_mask = (torch.arange(ty)[None] < yl[:, None]).float().unsqueeze(1)
    attn = net.align(mu_x.to(DEV), x_mask.to(DEV), y.to(DEV), y_mask.to(DEV))
    lp = align_oracle.log_prior(mu_x, y, 80)
    ref = mas_oracle.maximum_path(lp, (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1))
    assert torch.equal(attn.cpu(), ref)
    logw, mu_y = net.align_outputs(attn, mu_x.to(DEV), x_mask.to(DEV))
    assert float((logw.cpu() - align_oracle.logw_from_path(ref, x_mask)).abs().max()) <= 2e-6
    assert torch.equal(mu_y.cpu(), align_oracle.mu_y_from_path(ref, mu_x))
    with pytest.raises(RuntimeError):
        importlib.import_module("grad-tts_b200.model.align").log_prior(mu_x, y)      # CPU tensors: no fallback
