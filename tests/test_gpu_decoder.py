"""GPU: estimator / sampler / GradTTS through the drop-in modules, against golden fixtures and the oracle.

Tolerances, all relative to the reference output's max-abs `s`:
  fp32 mode : estimator 1e-5*max(1,s)  ... stated per test; 10-step decoder max-abs <= 1e-3 at |x|max ~ 140
  bf16 mode : estimator max-abs <= 5e-2 on |score|max ~ 1.6; decoder rel-rms <= 2e-2
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, decoder_case_names, load_decoder_case
from oracle import decoder_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# est_*/dec_*: small fixtures; c1*/c3*: the shapes BASELINE.json names (C1 = 1 x 400 frames x 10 steps, the bench's own weights and
# inputs; C3 shape = n_spks 247, 800 frames, ragged)
DEC = decoder_case_names(("est_", "dec_", "c1", "c3"))


def _module(pkg, synth, n_spks, wseed, precision):
    sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    missing = dec.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    dec = dec.to(DEV)
    dec.precision = precision
    return dec, sd


def _run_golden(pkg, synth, name, precision):
    g = load_decoder_case(name)
    n_spks, n_steps = g["n_spks"], g["n_steps"]
    dec, _ = _module(pkg, synth, n_spks, g["wseed"], precision)
    z, mask, mu = (g[k].to(DEV) for k in ("z", "mask", "mu"))
    spk = g["spk"].to(DEV) if g["spk"] is not None else None
    if n_steps == 0:
        y = dec.estimator(z * mask, mask, mu, g["t"].to(DEV), spk)
    else:
        y = dec(z, mask, mu, n_steps, True, spk)            # stoc=True: ignored like the reference
    return y.cpu(), g["y"], n_steps


@pytest.mark.parametrize("name", DEC)
def test_golden_fp32(name, pkg, synth):
    y, ref, n_steps = _run_golden(pkg, synth, name, "fp32")
    s = float(ref.abs().max())
    err = float((y - ref).abs().max())
    # single call: 1e-4 absolute on |score| ~ 1.6 ; sampler: the north-star bound max-abs 1e-3 (|x|max up to 141)
    tol = 1e-4 if n_steps == 0 else 1e-3
    assert err <= tol, f"{name}: max-abs {err} (|y|max {s})"


@pytest.mark.parametrize("name", DEC)
def test_golden_bf16(name, pkg, synth):
    y, ref, n_steps = _run_golden(pkg, synth, name, "bf16")
    s = float(ref.abs().max())
    err = float((y - ref).abs().max())
    relrms = float(((y - ref).pow(2).mean().sqrt()) / ref.pow(2).mean().sqrt())
    if n_steps == 0:
        assert err <= 5e-2, f"{name}: max-abs {err} (|y|max {s})"
    else:
        assert relrms <= 2e-2, f"{name}: rel-rms {relrms}, max-abs {err} (|y|max {s})"


def test_masked_region_is_zero_and_inputs_untouched(pkg, synth):
    dec, _ = _module(pkg, synth, 1, 0, "bf16")
    z, mask, mu, _, lengths = synth.make_inputs(3, 48, 1, seed=3)
    zc, mc, muc = z.to(DEV), mask.to(DEV), mu.to(DEV)
    z0, mu0 = zc.clone(), muc.clone()
    y = dec(zc, mc, muc, 3)
    assert torch.equal(zc, z0) and torch.equal(muc, mu0)
    assert float((y * (1 - mc)).abs().max()) == 0.0


def test_batch_chunking_is_bitwise_invariant(pkg, synth):
    """Per-sample maths: the same sample gives the same bits alone, in a batch, and across chunk sizes."""
    dec, _ = _module(pkg, synth, 247, 3, "bf16")
    z, mask, mu, spk, _ = synth.make_inputs(5, 40, 247, seed=4, ragged=False)
    a = [t.to(DEV) for t in (z, mask, mu, spk)]
    dec.estimator.max_chunk = 8
    y_all = dec(a[0], a[1], a[2], 2, False, a[3])
    dec.estimator.max_chunk = 2
    y_chunk = dec(a[0], a[1], a[2], 2, False, a[3])
    y_one = dec(a[0][3:4], a[1][3:4], a[2][3:4], 2, False, a[3][3:4])
    assert torch.equal(y_all, y_chunk)
    assert torch.equal(y_all[3:4], y_one)


def test_fused_gn_input_is_bitwise_equal_to_separate_pass(pkg, synth):
    """block2 convs that apply block1's GroupNorm+Mish(+time bias, mask) on their operand tiles (option fuse_gn=1) must give
    exactly the bits of the separate gn_apply pass (fuse_gn=0, the default): same formulas, same bf16 rounding point."""
    import ctypes
    outs = []
    for fuse in (1, 0):
        dec, _ = _module(pkg, synth, 247, 3, "bf16")
        z, mask, mu, spk, _ = synth.make_inputs(3, 88, 247, seed=21)
        h = dec.estimator._get_handle()
        dec.estimator.set_option("fuse_epi", 0)                # the operand-side variant only exists without the epilogue fusions
        dec.estimator.set_option("fuse_async", 0)
        pkg._lib.check(pkg._lib.load().gtts_decoder_set_option(h, b"fuse_gn", fuse), "set_option")
        outs.append(dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 3, False, spk.to(DEV)))
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("n_spks,B,T", [(1, 2, 88), (247, 3, 88), (1, 1, 4), (247, 2, 1032)])
def test_first_conv_on_tensor_cores_matches_ffma_kernel(pkg, synth, n_spks, B, T):
    """bf16 mode, first conv of the U-Net (2 | 3 fp32 planes -> 64 channels, model/diffusion.py:181-184, 52): the mma.sync kernel
    (bf16 weights, inputs split into bf16 hi + lo so they stay fp32-accurate; option first_conv_mma=1, the default) against the
    FFMA kernel with fp32 weights (first_conv_mma=0).  Two bf16 runs that round differently anywhere differ by about sqrt(2) x the
    bf16 error of either (measured 1.4e-2 rel-rms between them), so the yardstick is the fp32 mode of the same weights (within 1e-4
    of the reference): the tensor-core variant must be as close to it as the FFMA variant is (<= 1.5x its rel-rms error, both
    <= 2.5e-2), and close to the FFMA variant (<= 3e-2).  T = 88 / 1032 put 16-pixel groups across row ends (a wrong border
    tap would show as an error of order 1), T = 4 is the smallest legal width."""
    z, mask, mu, spk, _ = synth.make_inputs(B, T, n_spks, seed=33)
    t = torch.linspace(0.2, 0.9, B)
    args = ((z * mask).to(DEV), mask.to(DEV), mu.to(DEV), t.to(DEV), spk.to(DEV) if spk is not None else None)
    dec32, _ = _module(pkg, synth, n_spks, 5, "fp32")
    ref = dec32.estimator(*args).cpu()
    outs = []
    for mma in (1, 0):
        dec, _ = _module(pkg, synth, n_spks, 5, "bf16")
        dec.estimator.set_option("first_conv_mma", mma)
        outs.append(dec.estimator(*args).cpu())
    assert torch.isfinite(outs[0]).all()
    rms = lambda a: float(a.pow(2).mean().sqrt())
    e_mma, e_ffma, diff = rms(outs[0] - ref) / rms(ref), rms(outs[1] - ref) / rms(ref), rms(outs[0] - outs[1]) / rms(ref)
    msg = f"rel-rms vs fp32 mode: mma {e_mma:.3e}, ffma {e_ffma:.3e}; between them {diff:.3e}"
    print(msg)
    assert e_mma <= 2.5e-2 and e_ffma <= 2.5e-2 and e_mma <= 1.5 * e_ffma + 1e-3 and diff <= 3e-2, msg
    assert float((outs[0] - ref).abs().max()) <= 6e-2 * max(1.0, float(ref.abs().max())), msg


def test_apply_epilogue_is_bitwise_equal_to_separate_gn_pass(pkg, synth):
    """Plan whose Block convs finish GroupNorm+Mish(+time bias / residual) in their own epilogue (fuse_epi=2: always; the default
    uses it for small batches) against the plan with the separate gn_apply pass (fuse_epi=0): the same bits -- the fused epilogue
    rounds conv+bias to bf16 like the stored raw tensor and uses gn_apply's formulas -- with 22 fewer launches per Euler step."""
    outs, launches = [], []
    z, mask, mu, spk, _ = synth.make_inputs(3, 88, 247, seed=21)
    for fuse_epi, fuse_async in ((2, 0), (0, 1), (0, 0)):           # TMEM-resident apply / asynchronous apply warps / separate pass
        dec, sd = _module(pkg, synth, 247, 3, "bf16")
        dec.estimator.set_option("fuse_epi", fuse_epi)
        dec.estimator.set_option("fuse_async", fuse_async)
        outs.append(dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 3, False, spk.to(DEV)).cpu())
        launches.append(dec.estimator.launches_last_call())
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[2])
    assert launches[0] == launches[2] - 3 * 22 and launches[1] == launches[2] - 3 * 22, launches


def test_fp32_mode_tensor_core_and_ffma_convs_agree(pkg, synth):
    """precision='fp32' runs its convolutions on the tensor cores by default (bf16 x 3 split, six partial products, fp32 accumulate);
    option fp32_tc=0 selects the CUDA-core FFMA kernels.  Both meet the fp32 tolerance; they agree to accumulation-order noise."""
    z, mask, mu, spk, _ = synth.make_inputs(2, 72, 247, seed=33)
    outs = []
    for tc in (1, 0):
        dec, sd = _module(pkg, synth, 247, 3, "fp32")
        dec.estimator.set_option("fp32_tc", tc)
        outs.append(dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 3, False, spk.to(DEV)).cpu())
    with torch.no_grad():
        ref = decoder_oracle.reverse_diffusion(sd, z, mask, mu, 3, False, spk, 247)
    assert float((outs[0] - ref).abs().max()) <= 1e-3 and float((outs[1] - ref).abs().max()) <= 1e-3
    assert float((outs[0] - outs[1]).abs().max()) <= 5e-4, float((outs[0] - outs[1]).abs().max())


def test_long_random_weight_run_stays_finite(pkg, synth, monkeypatch):
    """100 Euler steps with random weights drive internal activations to ~1e16 (the bench workload does this): the output must
    stay finite on both attention kernels (regression: the tcgen05 softmax once formed k*log2e - m*log2e with two roundings)."""
    z, mask, mu, _, _ = synth.make_inputs(2, 344, 1, seed=7, ragged=False)
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("GTTS_ATTN_TC", tc)
        dec, _ = _module(pkg, synth, 1, 0, "bf16")
        outs[tc] = dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 100)
        assert torch.isfinite(outs[tc]).all(), f"GTTS_ATTN_TC={tc}: non-finite output"
    assert float(outs["1"].abs().max()) > 0.0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sde_extension_matches_restatement(pkg, synth, precision):
    """Upstream stochastic branch with injected noise (extension; unpinned: the fork has no such code) against the CPU restatement.
    Runs through the captured graph (noise base and stride live in device memory) and, with max_chunk=1, through per-chunk offsets
    into the caller's noise tensor; chunked and unchunked runs must agree bit for bit."""
    n_spks, B, T, n = 1, 3, 40, 3
    dec, sd = _module(pkg, synth, n_spks, 0, precision)
    z, mask, mu, _, _ = synth.make_inputs(B, T, n_spks, seed=12)
    noise = torch.randn(n, B, 80, T, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = decoder_oracle.reverse_diffusion(sd, z, mask, mu, n, True, None, n_spks, sde_noise=noise)
        ode = decoder_oracle.reverse_diffusion(sd, z, mask, mu, n, True, None, n_spks)
    assert float((ref - ode).abs().max()) > 0.1                       # the noise term really acts
    a = [t.to(DEV) for t in (z, mask, mu, noise)]
    y = dec.reverse_diffusion(a[0], a[1], a[2], n, True, None, sde_noise=a[3])
    dec.estimator.max_chunk = 1
    y1 = dec.reverse_diffusion(a[0], a[1], a[2], n, True, None, sde_noise=a[3])
    assert torch.equal(y, y1)
    y = y.cpu()
    if precision == "fp32":
        assert float((y - ref).abs().max()) <= 1e-3
    else:
        assert float((y - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) <= 2e-2
    # the literal north-star form  x - (...)*beta*h + sqrt(beta*h)*z  is this update with -z
    with torch.no_grad():
        ref_m = decoder_oracle.reverse_diffusion(sd, z, mask, mu, 1, True, None, n_spks, sde_noise=-noise[:1])
        est = decoder_oracle.estimator_forward(sd, z * mask, mask, mu, torch.full((B,), 0.5), None, n_spks)
    beta = 0.05 + (20.0 - 0.05) * 0.5
    lit = ((z * mask) - (0.5 * (mu - z * mask) - est) * beta * 1.0 + (beta * 1.0) ** 0.5 * noise[0]) * mask
    assert float((lit - ref_m).abs().max()) <= 1e-4


def test_plan_cache_pool_and_stream_order(pkg, synth):
    """Workspace hygiene: (1) plans of different (B, T) share one pooled workspace and an LRU-bounded plan cache, and a shape
    evicted and rebuilt gives the same bits; (2) liveness-based reuse shrinks a plan's workspace at least 3x; (3) two calls on
    different streams without any host synchronisation in between are ordered by the library (they use the same plan buffers)."""
    dec, _ = _module(pkg, synth, 1, 0, "bf16")
    dec.estimator.set_option("max_plans", 2)
    runs = {}
    for T in (40, 48, 56, 40, 48):
        z, mask, mu, _, _ = synth.make_inputs(2, T, 1, seed=T)
        y = dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 2)
        if T in runs:
            assert torch.equal(runs[T], y), T
        runs[T] = y
    info = dec.estimator.cache_info()
    assert info["plans_cached"] == 2 and info["plans_created"] == 5, info
    assert info["plan_workspace_bytes"] * 3 <= info["plan_workspace_bytes_without_reuse"], info
    fresh, _ = _module(pkg, synth, 1, 0, "bf16")
    z, mask, mu, _, _ = synth.make_inputs(2, 56, 1, seed=56)
    assert torch.equal(fresh(z.to(DEV), mask.to(DEV), mu.to(DEV), 2), runs[56])
    # two streams, same plan, no host sync between the calls
    z2, mask2, mu2, _, _ = synth.make_inputs(2, 56, 1, seed=57)
    a = [t.to(DEV) for t in (z, mask, mu)]
    b = [t.to(DEV) for t in (z2, mask2, mu2)]
    ref_b = fresh(b[0], b[1], b[2], 2)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        ya = dec(a[0], a[1], a[2], 2)
    with torch.cuda.stream(s2):
        yb = dec(b[0], b[1], b[2], 2)
    torch.cuda.synchronize()
    assert torch.equal(ya, runs[56]) and torch.equal(yb, ref_b)
    dec.estimator.set_option("trim", 1)
    assert dec.estimator.cache_info()["pool_bytes"] == 0
    assert torch.equal(dec(a[0], a[1], a[2], 2), runs[56])


def test_errors_are_loud(pkg, synth):
    dec, _ = _module(pkg, synth, 1, 0, "bf16")
    z, mask, mu, _, _ = synth.make_inputs(1, 40, 1, seed=3)
    with pytest.raises(RuntimeError):
        dec(z, mask, mu, 2)                                   # CPU inputs: no fallback
    with pytest.raises(ValueError):
        dec(z[:, :, :38].to(DEV), mask[:, :, :38].to(DEV), mu[:, :, :38].to(DEV), 2)      # T % 4 != 0
    dec2 = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000)     # parameters on CPU
    with pytest.raises(RuntimeError):
        dec2(z.to(DEV), mask.to(DEV), mu.to(DEV), 2)


class _StubEncoder(torch.nn.Module):
    """Deterministic stand-in with the reference TextEncoder contract (mu_x, logw, x_mask)."""

    def __init__(self, n_vocab, n_feats):
        super().__init__()
        self.emb = torch.nn.Embedding(n_vocab, n_feats)
        self.dur = torch.nn.Embedding(n_vocab, 1)

    def forward(self, x, x_lengths, spk=None):
        x_mask = (torch.arange(x.shape[1], device=x.device)[None, :] < x_lengths[:, None]).unsqueeze(1).float()
        mu_x = self.emb(x).transpose(1, 2) * x_mask
        logw = (self.dur(x).transpose(1, 2).tanh() + 0.7) * x_mask
        return mu_x, logw, x_mask


def test_gradtts_forward_dropin(pkg, synth):
    """Same seed => same z as the reference glue; return shapes incl. the attn slicing quirk (tts.py:108)."""
    from oracle import decoder_oracle as do
    torch.manual_seed(0)
    enc = _StubEncoder(20, 80)
    m = pkg.GradTTS(20, 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000, encoder=enc)
    sd = synth.make_decoder_state_dict(1, seed=0, g=0.05)
    m.decoder.load_state_dict(sd)
    m = m.to(DEV).eval()
    m.decoder.precision = "fp32"
    x = torch.randint(0, 20, (2, 13))
    x_lengths = torch.tensor([13, 9])
    torch.manual_seed(123)
    enc_out, dec_out, attn = m(x, x_lengths, n_timesteps=3, temperature=1.5, stoc=False, length_scale=1.0)
    # restate the glue on CPU with the same seed and compare against the oracle decoder
    enc_cpu = _StubEncoder(20, 80)
    enc_cpu.load_state_dict({k: v.cpu() for k, v in m.encoder.state_dict().items()})
    mu_x, logw, x_mask = enc_cpu(x, x_lengths)
    w_ceil = torch.ceil(torch.exp(logw) * x_mask)
    y_lengths = torch.clamp_min(w_ceil.sum([1, 2]), 1).long()
    y_max = int(y_lengths.max())
    T = (y_max + 3) // 4 * 4
    assert enc_out.shape == (2, 80, y_max) and dec_out.shape == (2, 80, y_max) and attn.shape == (2, 1, 13, T)
    y_mask = (torch.arange(T)[None, :] < y_lengths[:, None]).unsqueeze(1).float()
    utils = __import__("importlib").import_module("grad-tts_b200.model.utils")
    attn_ref = utils.generate_path(w_ceil.squeeze(1), (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1))
    mu_y = torch.matmul(attn_ref.transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
    torch.manual_seed(123)
    z = mu_y.to(DEV) + torch.randn_like(mu_y.to(DEV)) / 1.5          # the reference draws on the model device
    with torch.no_grad():
        ref = do.reverse_diffusion(sd, z.cpu(), y_mask, mu_y, 3)
    assert float((enc_out.cpu() - mu_y[:, :, :y_max]).abs().max()) <= 1e-5
    assert float((dec_out.cpu() - ref[:, :, :y_max]).abs().max()) <= 1e-3


def test_c_abi_writes_stay_inside_the_caller_buffers(pkg, synth):
    """Own bounds check (no sanitizer on the pool): inputs and outputs handed to the C ABI live inside larger buffers filled with a
    canary; after sampler, estimator, MAS and alignment calls at sizes that leave ragged tile tails, every canary is intact and
    the outputs equal those of the ordinary (unguarded) call bit for bit."""
    import ctypes
    lib = pkg._lib.load()
    G, CAN = 4096, -777.25

    def guarded(src=None, numel=None):
        n = src.numel() if src is not None else numel
        buf = torch.full((n + 2 * G,), CAN, dtype=torch.float32, device=DEV)
        if src is not None:
            buf[G:G + n] = src.reshape(-1).to(DEV)
        return buf, buf[G:G + n]

    def intact(buf):
        return bool((buf[:G] == CAN).all()) and bool((buf[-G:] == CAN).all())

    for n_spks, wseed, B, T in [(1, 0, 3, 44), (247, 3, 2, 36)]:
        dec, _ = _module(pkg, synth, n_spks, wseed, "bf16")
        z, mask, mu, spk, _ = synth.make_inputs(B, T, n_spks, seed=17)
        ref = dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 3, False, spk.to(DEV) if spk is not None else None)
        t = torch.tensor([0.2, 0.7, 0.95][:B])
        ref_e = dec.estimator((z * mask).to(DEV), mask.to(DEV), mu.to(DEV), t.to(DEV), spk.to(DEV) if spk is not None else None)
        bufs = [guarded(v) for v in (z, mask, mu)] + [guarded(numel=z.numel())]
        sb = guarded(spk) if spk is not None else (None, None)
        h = dec.estimator._get_handle()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ptr = lambda v: v.data_ptr() if v is not None else None
        rc = lib.gtts_decoder_reverse_diffusion(h, ptr(bufs[0][1]), ptr(bufs[1][1]), ptr(bufs[2][1]), ptr(sb[1]), ptr(bufs[3][1]),
                                                B, T, 3, dec.estimator._flags(), None, stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert all(intact(b) for b, _ in bufs) and (sb[0] is None or intact(sb[0]))
        assert torch.equal(bufs[3][1].reshape(B, 80, T), ref)
        assert torch.equal(bufs[0][1].cpu(), z.reshape(-1))                    # inputs untouched
        zb, tb, ob = guarded(z * mask), guarded(t), guarded(numel=z.numel())
        rc = lib.gtts_decoder_estimator(h, ptr(zb[1]), ptr(bufs[1][1]), ptr(bufs[2][1]), ptr(tb[1]), ptr(sb[1]), ptr(ob[1]), B, T,
                                        dec.estimator._flags(), stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert intact(zb[0]) and intact(tb[0]) and intact(ob[0]) and all(intact(b) for b, _ in bufs)
        assert torch.equal(ob[1].reshape(B, 80, T), ref_e)

    # MAS + alignment stage, odd sizes
    B, tx, ty = 3, 37, 131
    value, mmask, _, _ = synth.make_mas_inputs(B, tx, ty, seed=5, ragged=True)
    vb, mb, pb = guarded(value), guarded(mmask), guarded(numel=value.numel())
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws_bytes = int(lib.gtts_mas_workspace_bytes(B, tx, ty))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=DEV)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.gtts_mas_maximum_path(vb[1].data_ptr(), mb[1].data_ptr(), pb[1].data_ptr(), B, tx, ty, ws.data_ptr(), ws_bytes,
                                   status.data_ptr(), stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert intact(vb[0]) and intact(mb[0]) and intact(pb[0]) and int(status) == 0
    ma = __import__("importlib").import_module("grad-tts_b200.model.monotonic_align")
    assert torch.equal(pb[1].reshape(B, tx, ty), ma.maximum_path(value.to(DEV), mmask.to(DEV)))
    gen = torch.Generator().manual_seed(6)
    mu_x, y = torch.randn(B, 80, tx, generator=gen), torch.randn(B, 80, ty, generator=gen)
    xb, yb, lb = guarded(mu_x), guarded(y), guarded(numel=B * tx * ty)
    assert lib.gtts_align_log_prior(xb[1].data_ptr(), yb[1].data_ptr(), lb[1].data_ptr(), B, 80, tx, ty, stream) == 0
    xm = guarded(torch.ones(B, tx))
    wb, ub = guarded(numel=B * tx), guarded(numel=B * 80 * ty)
    assert lib.gtts_align_outputs(pb[1].data_ptr(), xb[1].data_ptr(), xm[1].data_ptr(), wb[1].data_ptr(), ub[1].data_ptr(), B, 80,
                                  tx, ty, stream) == 0
    torch.cuda.synchronize()
    assert all(intact(b) for b, _ in (xb, yb, lb, xm, wb, ub, pb))
    assert not bool((lb[1] == CAN).any()) and not bool((ub[1] == CAN).any()) and not bool((wb[1] == CAN).any())   # fully written


def _oracle_est(sd, x, mask, mu, t, spk=None, n_spks=1):
    torch.set_num_threads(8)
    with torch.no_grad():
        return decoder_oracle.estimator_forward(sd, x, mask, mu, t, spk, n_spks)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 5e-2)])
def test_edge_shapes_and_masks_against_oracle(pkg, synth, precision, tol):
    """Edge cases of the estimator (the reference has no tests; the oracle is the checker): the smallest legal T (level 2 is one
    frame wide), a batch holding an empty (length 0) and a one-frame utterance, and an arbitrary non-prefix binary mask, which the
    reference API accepts (model/diffusion.py:174)."""
    dec, sd = _module(pkg, synth, 1, 0, precision)
    gen = torch.Generator().manual_seed(77)
    cases = []
    for B, T, kind in [(1, 4, "full"), (2, 8, "full"), (3, 24, "empty+one"), (2, 40, "random")]:
        x, mu = torch.randn(B, 80, T, generator=gen), torch.randn(B, 80, T, generator=gen)
        if kind == "full":
            mask = torch.ones(B, 1, T)
        elif kind == "empty+one":
            lengths = torch.tensor([0, 1, T])
            mask = (torch.arange(T)[None] < lengths[:, None]).float().unsqueeze(1)
        else:
            mask = (torch.rand(B, 1, T, generator=gen) > 0.35).float()
        t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
        cases.append((x * mask, mask, mu, t))
    for x, mask, mu, t in cases:
        ref = _oracle_est(sd, x, mask, mu, t)
        got = dec.estimator(x.to(DEV), mask.to(DEV), mu.to(DEV), t.to(DEV)).cpu()
        assert torch.isfinite(got).all()
        assert float((got * (1 - mask)).abs().max()) == 0.0
        err = float((got - ref).abs().max())
        assert err <= tol, f"T={x.shape[-1]} B={x.shape[0]}: max-abs {err} (|ref|max {float(ref.abs().max())})"


def test_long_utterance_against_oracle(pkg, synth):
    """One utterance of 6000 frames (480 000 attention positions at level 0, 70 s of audio): index arithmetic, attention chunking and
    tile tails far from the fixture sizes.  bf16 mode, same bounds as the fixtures (max-abs 5e-2 on a score of O(1); rel-rms 2e-2,
    measured 1.2e-2 / 2.5e-2 max-abs)."""
    dec, sd = _module(pkg, synth, 1, 0, "bf16")
    B, T = 1, 6000
    x, mask, mu, _, _ = synth.make_inputs(B, T, 1, seed=91, ragged=False)
    mask[:, :, 5555:] = 0
    t = torch.tensor([0.37])
    ref = _oracle_est(sd, x * mask, mask, mu, t)
    got = dec.estimator((x * mask).to(DEV), mask.to(DEV), mu.to(DEV), t.to(DEV)).cpu()
    err = float((got - ref).abs().max())
    rel = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert err <= 5e-2 and rel <= 2e-2, (err, rel, float(ref.abs().max()))


def test_batch_larger_than_one_workspace_chunk(pkg, synth):
    """65 utterances = one chunk of 64 and one of 1 (max_chunk is 64): every sample equals its single-sample run bit for bit."""
    dec, _ = _module(pkg, synth, 1, 0, "bf16")
    z, mask, mu, _, _ = synth.make_inputs(65, 8, 1, seed=13)
    y = dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 2)
    for b in (0, 31, 63, 64):
        yb = dec(z[b:b + 1].to(DEV), mask[b:b + 1].to(DEV), mu[b:b + 1].to(DEV), 2)
        assert torch.equal(y[b:b + 1], yb), b


@pytest.mark.parametrize("precision,tol_abs,tol_rel", [("bf16", 5e-2, 2e-2), ("fp32", 1e-4, 2e-5)])
def test_bench_chunk_shape_against_oracle(pkg, synth, precision, tol_abs, tol_rel):
    """The bench workload's own chunk shape (BASELINE config 5: 64 utterances x 1720 frames per workspace chunk, the bench's
    weights and inputs): one estimator call and one Euler step on the full chunk; samples 0, 31 and 63 are compared with the CPU
    oracle run on those samples alone (the maths is per-sample, so the oracle need not run all 64)."""
    B, T = 64, 1720
    dec, sd = _module(pkg, synth, 1, 0, precision)
    z, mask, mu, _, _ = synth.make_inputs(B, T, 1, seed=1, ragged=False)
    mask[31, :, 1500:] = 0                                   # one padded utterance inside the chunk
    pick = [0, 31, 63]
    t = torch.full((B,), 0.995)                              # t of the first of 100 Euler steps
    zd, md, mud = z.to(DEV), mask.to(DEV), mu.to(DEV)
    got_e = dec.estimator(zd * md, md, mud, t.to(DEV))[pick].cpu()
    got_s = dec(zd, md, mud, 1)[pick].cpu()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref_e = decoder_oracle.estimator_forward(sd, (z * mask)[pick], mask[pick], mu[pick], t[pick], None, 1)
        ref_s = decoder_oracle.reverse_diffusion(sd, z[pick], mask[pick], mu[pick], 1, False, None, 1)
    err_e = float((got_e - ref_e).abs().max())
    rel_e = float((got_e - ref_e).pow(2).mean().sqrt() / ref_e.pow(2).mean().sqrt())
    assert err_e <= tol_abs and rel_e <= tol_rel, (err_e, rel_e, float(ref_e.abs().max()))
    # one Euler step with n_timesteps=1: h = 1, beta(0.5) = 10.0 -> the update is 5x the score error
    err_s = float((got_s - ref_s).abs().max())
    assert err_s <= 5.5 * tol_abs, (err_s, float(ref_s.abs().max()))
    assert float((got_s * (1 - mask[pick])).abs().max()) == 0.0
