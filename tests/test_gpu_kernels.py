"""GPU: every convolution variant of the U-Net, CUDA-core and tcgen05 implementations, against torch CPU."""
import pytest
import torch

import gpu_util

pytestmark = pytest.mark.gpu


def _gu():
    return gpu_util


@pytest.mark.parametrize("name,kw", gpu_util.CONV_CASES)
@pytest.mark.parametrize("impl,act", [(0, 0), (0, 1), (1, 1)])
def test_conv_variants(name, kw, impl, act):
    gu = _gu()
    c = gu.conv_case(seed=hash(name) % 1000, **kw)
    stats_ok = c["kind"] in (0, 1) and c["r"] is None and c["m"] is None and not c["per_sample"]
    out, st = gu.run_conv(c, impl, act, want_stats=stats_ok)
    ref, raw = gu.conv_reference(c, round_bf16=bool(act))
    assert not torch.isnan(out).any(), "output has unwritten (NaN) entries"
    err = float((out - ref).abs().max())
    # fp32: FFMA vs MKLDNN summation order; bf16: output rounding 2^-9 relative on |y| ~ 4
    tol = 2e-4 if act == 0 else 4e-2
    assert err <= tol, f"{name} impl={impl} act={act}: max-abs err {err}"
    if stats_ok:
        sref = gu.gn_stats_reference(raw)
        serr = float(((st - sref).abs() / (sref.abs() + 1.0)).max())
        assert serr <= (1e-4 if act == 0 else 2e-3), f"{name}: GN stats err {serr}"


HALO_CASES = [(n, kw) for (n, kw) in gpu_util.CONV_CASES if kw["kind"] == 0] + [
    ("3x3_128_128_pass", dict(kind=0, B=3, H=40, W=52, Cin0=128, Cin1=0, Cout=128)),      # streamed weights, odd tile count
    ("3x3_256_64_cat", dict(kind=0, B=2, H=40, W=30, Cin0=128, Cin1=128, Cout=64)),
    ("3x3_256_256_many", dict(kind=0, B=8, H=20, W=216, Cin0=256, Cin1=0, Cout=256)),     # > 1 pass per CTA
]


@pytest.mark.parametrize("name,kw", HALO_CASES)
@pytest.mark.parametrize("impl", [2, 3])
def test_conv_halo(name, kw, impl, monkeypatch):
    """Halo-box tcgen05 kernel (resident or streamed weights, one or two tiles per weight pass) vs torch CPU."""
    monkeypatch.setenv("GTTS_HALO256", "1")          # Cout = 256 is opt-in (the per-tap kernel is faster there)
    gu = _gu()
    c = gu.conv_case(seed=hash(name) % 1000, **kw)
    out, st = gu.run_conv(c, impl, 1, want_stats=True)
    ref, raw = gu.conv_reference(c, round_bf16=True)
    assert not torch.isnan(out).any(), "output has unwritten (NaN) entries"
    err = float((out - ref).abs().max())
    assert err <= 4e-2, f"{name} impl={impl}: max-abs err {err}"
    sref = gu.gn_stats_reference(raw)
    serr = float(((st - sref).abs() / (sref.abs() + 1.0)).max())
    assert serr <= 2e-3, f"{name}: GN stats err {serr}"


def test_tc_matches_ffma_bitwise_inputs():
    """Same bf16 operands through both implementations: only accumulation order differs."""
    gu = _gu()
    c = gu.conv_case(0, 2, 40, 36, 128, 0, 128, seed=5)
    a, _ = gu.run_conv(c, 0, 1)
    b, _ = gu.run_conv(c, 1, 1)
    assert float((a - b).abs().max()) <= 4e-2
