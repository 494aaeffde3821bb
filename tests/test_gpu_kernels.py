"""GPU: every convolution variant of the U-Net, CUDA-core and tcgen05 implementations, against torch CPU."""
import pytest
import torch

import gpu_util

pytestmark = pytest.mark.gpu


def _gu():
    return gpu_util


def _seed(name):
    """Stable per-case seed (Python's hash() of a str is salted per process, which made failures unrepeatable)."""
    import zlib
    return zlib.crc32(name.encode()) % 1000


@pytest.mark.parametrize("name,kw", gpu_util.CONV_CASES)
@pytest.mark.parametrize("impl,act", [(0, 0), (0, 1), (1, 1)])
def test_conv_variants(name, kw, impl, act):
    gu = _gu()
    c = gu.conv_case(seed=_seed(name), **kw)
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
    c = gu.conv_case(seed=_seed(name), **kw)
    out, st = gu.run_conv(c, impl, 1, want_stats=True)
    ref, raw = gu.conv_reference(c, round_bf16=True)
    assert not torch.isnan(out).any(), "output has unwritten (NaN) entries"
    err = float((out - ref).abs().max())
    assert err <= 4e-2, f"{name} impl={impl}: max-abs err {err}"
    sref = gu.gn_stats_reference(raw)
    serr = float(((st - sref).abs() / (sref.abs() + 1.0)).max())
    assert serr <= 2e-3, f"{name}: GN stats err {serr}"


CONVT_HALO_CASES = [
    ("convT_64_h40", dict(kind=3, B=2, H=40, W=36, Cin0=64, Cin1=0, Cout=64, mask=True)),            # resident weights
    ("convT_128_h20", dict(kind=3, B=3, H=20, W=44, Cin0=128, Cin1=0, Cout=128, mask=True)),         # streamed weights, 2 chunks
    ("convT_64_odd", dict(kind=3, B=1, H=16, W=18, Cin0=64, Cin1=0, Cout=64, mask=True)),            # odd tile count (dummy tile)
    ("convT_128_many", dict(kind=3, B=7, H=24, W=130, Cin0=128, Cin1=0, Cout=128, mask=True)),       # several tiles per CTA pair
]


@pytest.mark.parametrize("name,kw", CONVT_HALO_CASES)
def test_convT_halo(name, kw):
    """Transposed conv on the CTA-pair halo kernel (4 phases x 4 taps as views of one halo box) vs torch CPU and vs the
    per-tap kernel."""
    gu = _gu()
    c = gu.conv_case(seed=_seed(name), **kw)
    out, _ = gu.run_conv(c, 3, 1)
    ref, _ = gu.conv_reference(c, round_bf16=True)
    assert not torch.isnan(out).any(), "output has unwritten (NaN) entries"
    err = float((out - ref).abs().max())
    assert err <= 4e-2, f"{name}: max-abs err {err}"
    tap, _ = gu.run_conv(c, 1, 1)
    assert float((out - tap).abs().max()) <= 4e-2


def test_tc_matches_ffma_bitwise_inputs():
    """Same bf16 operands through both implementations: only accumulation order differs."""
    gu = _gu()
    c = gu.conv_case(0, 2, 40, 36, 128, 0, 128, seed=5)
    a, _ = gu.run_conv(c, 0, 1)
    b, _ = gu.run_conv(c, 1, 1)
    assert float((a - b).abs().max()) <= 4e-2


def _attn_reference(x, wkv):
    """softmax over positions of k = Wk x per (head, dim); ctx[h, d, e] = sum_n p[h, d, n] v[h, e, n]  (diffusion.py:93-97)."""
    B, n, C = x.shape
    k = torch.einsum("bnc,dc->bdn", x, wkv[:128]).view(B, 4, 32, n)
    v = torch.einsum("bnc,ec->ben", x, wkv[128:]).view(B, 4, 32, n)
    p = torch.softmax(k, dim=-1)
    return torch.einsum("bhdn,bhen->bhde", p, v)


def _merge_partials(part):
    m, l, ctx = part[..., :32], part[..., 32:64], part[..., 64:].reshape(*part.shape[:-1], 32, 32)
    M = m.max(dim=2, keepdim=True).values
    w = torch.exp(m - M)                                    # [B, 4, chunks, 32]
    num = (w.unsqueeze(-1) * ctx).sum(dim=2)
    den = (w * l).sum(dim=2)
    return num / den.unsqueeze(-1)


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("B,n,chunks,chunk_len,scale", [(2, 1000, 3, 384, 1.0), (1, 4096, 4, 1024, 1.0), (3, 700, 1, 768, 6.0),
                                                        (5, 3000, 40, 128, 1.0),
                                                        # activations of a diverging random-weight run (|x| ~ 1e16, |k| ~ 1e17): the
                                                        # exponent must be formed as (k - m) * log2(e), not k*log2(e) - m*log2(e)
                                                        (2, 1000, 3, 384, 1e16)])
def test_attn_xk_matches_reference(use_tc, C, B, n, chunks, chunk_len, scale, monkeypatch):
    """Fused k-projection + online softmax + context (tcgen05 and mma.sync kernels) against the plain formula.  scale = 6
    makes the running maximum jump between tiles (exercises the lazy rescale of the TMEM accumulator); GTTS_ATTN_TAU=0 forces a
    rescale on every increase."""
    import ctypes
    import importlib
    pkg = importlib.import_module("grad-tts_b200")
    if scale != 1.0:
        monkeypatch.setenv("GTTS_ATTN_TAU", "0")
    g = torch.Generator().manual_seed(n + B)
    x = (torch.randn(B, n, C, generator=g) * scale).to(torch.bfloat16)
    ramp = torch.linspace(0.2, 1.5, n).view(1, n, 1)        # later pixels are larger: the maximum keeps growing
    x = (x.float() * ramp).to(torch.bfloat16)
    wkv = (torch.randn(256, C, generator=g) / C ** 0.5).to(torch.bfloat16)
    xd, wd = x.cuda(), wkv.cuda()
    part = torch.full((B, 4, chunks, 1088), float("nan"), device="cuda")
    lib = pkg._lib.load()
    rc = lib.gtts_test_attn_xk(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(wd.data_ptr()), ctypes.c_void_p(part.data_ptr()),
                               B, n, C, chunks, chunk_len, use_tc, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg._lib.check(rc, "gtts_test_attn_xk")
    torch.cuda.synchronize()
    got = _merge_partials(part.cpu())
    ref = _attn_reference(x.float(), wkv.float())
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    assert err <= 2e-2, f"use_tc={use_tc}: relative max error {err}"          # bf16 P and bf16 S


@pytest.mark.parametrize("C", [64, 128, 256])
@pytest.mark.parametrize("B", [1, 3, 19])
@pytest.mark.parametrize("out_bf16", [0, 1])
def test_attn_fold_variants_agree_bitwise(C, B, out_bf16):
    """Folded attention weights M_b = g * Wout . blockdiag(ctx_b^T) . Wq (model/diffusion.py:95-104 applied to a per-sample matrix):
    both kernels (16-row tiles for small batches, full-row tiles for large ones) against the fp64 formula (fp32 tolerance 1e-5
    relative to the largest entry), and against each other BIT FOR BIT -- the product path picks one by batch size, so a sample's
    result must not depend on which."""
    import ctypes
    import importlib
    pkg = importlib.import_module("grad-tts_b200")
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(1000 * C + 10 * B + out_bf16)
    ctx = torch.randn(B, 4, 32, 32, generator=g) / 6.0                 # [b][h][d][e]
    wout = torch.randn(C, 128, generator=g) / 128 ** 0.5               # [co][h*32 + e]
    wq = torch.randn(128, C, generator=g) / C ** 0.5                   # [h*32 + d][ci]
    gain = 0.37
    P = torch.einsum("che,bhde->bchd", wout.double().view(C, 4, 32), ctx.double()).reshape(B, C, 128)
    ref = gain * P @ wq.double()
    outs = []
    for variant in (0, 1, -1):
        out = torch.full((B, C, C), float("nan"), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
        a = [t.cuda().contiguous() for t in (ctx, wout, wq)]
        rc = lib.gtts_test_attn_fold(ctypes.c_void_p(a[0].data_ptr()), ctypes.c_void_p(a[1].data_ptr()), ctypes.c_void_p(a[2].data_ptr()),
                                     ctypes.c_float(gain), ctypes.c_void_p(out.data_ptr()), B, C, out_bf16, variant,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        pkg._lib.check(rc, "gtts_test_attn_fold")
        torch.cuda.synchronize()
        outs.append(out.cpu())
    tol = (8e-3 if out_bf16 else 1e-5) * float(ref.abs().max())
    assert float((outs[0].double() - ref).abs().max()) <= tol
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


APPLY_CASES = [
    # name, B, H, W, Cin0, Cin1, Cout, residual (else time bias), per-sample time bias
    ("l1_128_128_tb", 3, 40, 52, 128, 0, 128, False, False),        # 3 x 21 tiles (odd count per sample): runs straddle CTA pairs
    ("l1_128_128_res", 2, 40, 64, 128, 0, 128, True, False),
    ("l2_256_256_tb_persample", 5, 20, 44, 256, 0, 256, False, True),
    ("l2_256_256_res", 4, 20, 108, 256, 0, 256, True, False),
    ("l2_cat_512_128_tb", 2, 20, 30, 256, 256, 128, False, False),
    ("l1_256_64_cat_tb", 2, 40, 30, 128, 128, 64, False, False),
    ("l1_64_64_res", 3, 40, 36, 64, 0, 64, True, False),
    ("l0_64_64_tb_small", 2, 80, 24, 64, 0, 64, False, False),
    ("l1_many_tiles", 12, 40, 216, 128, 0, 128, True, False),        # 12 x 135 tiles over 148 CTAs: up to two runs in flight per CTA
    ("l2_many_tiles", 40, 20, 216, 256, 0, 256, False, False),       # 40 x 42 tiles, two TMEM buffers: tile-level pipelining
    ("l1_runs_of_two", 2, 40, 860, 128, 0, 128, False, False),       # 270 tiles per sample on 148 CTAs: two tiles of a sample wait in TMEM
    ("l0_long_runs", 1, 80, 1200, 64, 0, 64, True, False),           # 750 tiles of ONE sample: up to six accumulators per CTA wait for the barrier
]


@pytest.mark.parametrize("variant", ["tmem", "async"])
@pytest.mark.parametrize("case", APPLY_CASES, ids=[c[0] for c in APPLY_CASES])
def test_conv_apply_epilogue(case, variant):
    """Block conv with the GroupNorm-apply epilogue (accumulators wait in TMEM for the per-sample grid barrier) vs torch CPU:
    out = (mish(group_norm(conv(x) + b)) + time bias | + residual) * mask, bf16 operands, fp32 everything else."""
    import ctypes
    import torch.nn.functional as F
    name, B, H, W, Cin0, Cin1, Cout, use_res, per_sample = case
    gu = _gu()
    lib = gu._lib.load()
    g = torch.Generator().manual_seed(_seed(name))
    Cin = Cin0 + Cin1
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    gamma = 1.0 + 0.1 * torch.randn(Cout, generator=g)
    beta = 0.1 * torch.randn(Cout, generator=g)
    mask = (torch.rand(B, W, generator=g) > 0.25).float()
    x = x * mask[:, None, None, :]                                   # the decoder's inputs are stored masked
    tb = torch.randn(B if per_sample else 1, Cout, generator=g) if not use_res else None
    res = torch.randn(B, Cout, H, W, generator=g) if use_res else None
    rb = lambda t: t.to(torch.bfloat16).float()
    y = F.conv2d(rb(x), rb(w), b, padding=1)
    sref = gu.gn_stats_reference(y)
    y = F.group_norm(y, 8, gamma, beta, eps=1e-5)
    y = y * torch.tanh(F.softplus(y))
    if tb is not None:
        y = y + (tb if per_sample else tb.expand(B, -1))[:, :, None, None]
    if res is not None:
        y = y + rb(res)
    ref = y * mask[:, None, None, :]
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    x0 = gu.nhwc(x[:, :Cin0], dt).to(dev)
    x1 = gu.nhwc(x[:, Cin0:], dt).to(dev) if Cin1 else None
    out = torch.full((B, H, W, Cout), float("nan"), dtype=dt, device=dev)
    stats = torch.zeros(B, 8, 2, dtype=torch.float32, device=dev)
    d = lambda t: t.float().contiguous().to(dev) if t is not None else None
    wd, bd, gd, bed, md, tbd = d(w), d(b), d(gamma), d(beta), d(mask), d(tb)
    rd = gu.nhwc(res, dt).to(dev) if res is not None else None
    p = lambda t: t.data_ptr() if t is not None else None
    rc = lib.gtts_test_conv_apply(B, H, W, Cin0, Cin1, Cout, p(x0), p(x1), p(wd), p(bd), p(gd), p(bed), p(tbd),
                                  Cout if per_sample else 0, p(rd), p(md), p(out), p(stats), 3 if variant == "tmem" else -3,
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    gu._lib.check(rc, "gtts_test_conv_apply")
    torch.cuda.synchronize()
    got = gu.nchw(out.float().cpu())
    live = mask[:, None, None, :].expand_as(ref) > 0
    assert not torch.isnan(got).any(), "output has unwritten (NaN) entries"
    err = float((got - ref).abs().max())
    # bf16 output (half an ulp is 0.031 at |y| ~ 9) on top of the bf16 rounding of conv + bias that the kernel applies on purpose
    # (bitwise equality with the unfused plan)
    tol = max(4e-2, 6e-3 * float(ref.abs().max()))
    assert err <= tol, f"{name}: max-abs err {err} (|ref|max {float(ref.abs().max())})"
    assert float(got[~live].abs().max() if (~live).any() else 0.0) == 0.0
    serr = float(((stats.cpu() - sref).abs() / (sref.abs() + 1.0)).max())
    assert serr <= 2e-3, f"{name}: GN stats err {serr}"


@pytest.mark.parametrize("name,kw", [(n, kw) for (n, kw) in gpu_util.CONV_CASES if not kw.get("per_sample")] +
                         [("3x3_128_128_pass", dict(kind=0, B=3, H=40, W=52, Cin0=128, Cin1=0, Cout=128)),
                          ("3x3_256_256_many", dict(kind=0, B=8, H=20, W=216, Cin0=256, Cin1=0, Cout=256))])
def test_conv_fp32_on_tensor_cores(name, kw):
    """fp32 mode on the tensor cores: fp32 NHWC in, [hi|mid|lo] bf16 planes x [wh wh wh wm wm wl] weights (six partial products
    accumulated in fp32 TMEM), fp32 out -- every conv variant against torch CPU fp32 with the FFMA kernel's tolerance."""
    gu = _gu()
    c = gu.conv_case(seed=_seed(name), **kw)
    stats_ok = c["kind"] in (0, 1) and c["r"] is None and c["m"] is None
    out, st = gu.run_conv(c, 4, 0, want_stats=stats_ok)
    ref, raw = gu.conv_reference(c, round_bf16=False)
    assert not torch.isnan(out).any(), "output has unwritten (NaN) entries"
    err = float((out - ref).abs().max())
    assert err <= 2e-4, f"{name}: max-abs err {err} (|ref|max {float(ref.abs().max())})"
    ffma, _ = gu.run_conv(c, 0, 0)
    # the two fp32 implementations agree to accumulation noise (the tensor core truncates its fp32 accumulation: ~K/16 * 2^-24)
    assert float((out - ffma).abs().max()) <= 6e-5, float((out - ffma).abs().max())
    if stats_ok:
        sref = gu.gn_stats_reference(raw)
        assert float(((st - sref).abs() / (sref.abs() + 1.0)).max()) <= 1e-4
