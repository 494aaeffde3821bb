"""GPU: the HiFi-GAN generator (csrc/vocoder.cu behind gtts_vocoder_*) against the golden vectors made by the reference's
own hifi-gan/models.py::Generator, and against the oracle at larger sizes."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

VOC = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "voc_*.npz")))

# tolerances.  fp32 mode: fp32 FFMA convs against the reference's fp32 CPU convs, output in [-1, 1]: max-abs 2e-4.
# bf16 mode: bf16 operands and activations, fp32 accumulation, ~50 convs deep with a residual stream: rel-rms 3e-2.
FP32_MAX_ABS = 2e-4
BF16_REL_RMS = 3e-2


def _make(pkg, synth, cfg_name, wseed, weight_norm=True):
    cfg = synth.VOCODER_CONFIGS[cfg_name]
    sd = synth.make_vocoder_state_dict(cfg, seed=wseed)
    gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
    gen.load_state_dict(sd, strict=True)
    gen = gen.cuda().eval()
    if not weight_norm:
        gen.remove_weight_norm()
    return gen, cfg, sd


def _rel_rms(a, b):
    return float(((a - b).pow(2).mean() / b.pow(2).mean()).sqrt())


@pytest.mark.parametrize("name", VOC)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vocoder_matches_reference_golden(name, precision, pkg, synth):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    gen, cfg, _ = _make(pkg, synth, str(g["cfg"]), int(g["wseed"]), weight_norm=False)
    gen.precision = precision
    y = gen(torch.from_numpy(g["mel"]).cuda()).cpu()
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape and torch.isfinite(y).all()
    if precision == "fp32":
        assert float((y - ref).abs().max()) <= FP32_MAX_ABS
    else:
        assert _rel_rms(y, ref) <= BF16_REL_RMS
        assert gen.launches_last_call() > 0


def test_vocoder_weight_norm_and_removed_weight_norm_agree(pkg, synth):
    """inference.py:74-76 calls remove_weight_norm() after loading; forward before and after must give the same audio."""
    gen, cfg, _ = _make(pkg, synth, "v1", 31, weight_norm=True)
    mel = synth.make_mel(2, 12, seed=3).cuda()
    y0 = gen(mel)
    gen.remove_weight_norm()
    y1 = gen(mel)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vocoder_chunking_is_bitwise_invariant(precision, pkg, synth):
    gen, cfg, _ = _make(pkg, synth, "v1", 32)
    gen.precision = precision
    mel = synth.make_mel(3, 20, seed=4).cuda()
    y_all = gen(mel)
    gen.max_chunk = 2                                           # 2 + 1
    y_chunks = gen(mel)
    y_single = torch.cat([gen(mel[i:i + 1]) for i in range(3)])
    assert torch.equal(y_all, y_chunks)
    assert torch.equal(y_all, y_single)


@pytest.mark.parametrize("T", [1, 3, 129])
def test_vocoder_edge_lengths_against_oracle(T, pkg, synth):
    """one frame, a length below every tile size, a length that straddles a 128-position tile at the first stage"""
    from oracle import vocoder_oracle
    gen, cfg, sd = _make(pkg, synth, "v1", 33)
    mel = synth.make_mel(2, T, seed=5)
    torch.set_num_threads(8)
    with torch.no_grad():
        ref = vocoder_oracle.generator_forward(sd, cfg, mel)
    gen.precision = "fp32"
    y = gen(mel.cuda()).cpu()
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) <= FP32_MAX_ABS
    gen.precision = "bf16"
    y = gen(mel.cuda()).cpu()
    assert _rel_rms(y, ref) <= BF16_REL_RMS


def test_vocoder_tensor_core_and_ffma_bf16_paths_agree(pkg, synth):
    """same bf16 operands through the tcgen05 kernels and through the CUDA-core kernel: only the accumulation order differs"""
    gen, cfg, _ = _make(pkg, synth, "rb2", 34)
    mel = synth.make_mel(2, 40, seed=6).cuda()
    y_tc = gen(mel)
    gen.set_option("force_ffma", 1)
    y_ffma = gen(mel)
    gen.set_option("force_ffma", 0)
    assert _rel_rms(y_tc, y_ffma) <= 1e-2


def test_vocoder_host_entry_point(pkg, synth):
    gen, cfg, _ = _make(pkg, synth, "v1", 35)
    mel = synth.make_mel(2, 16, seed=7)
    y_dev = gen(mel.cuda()).cpu()
    lib = pkg._lib.load()
    out = torch.empty(2, 1, 16 * gen.hop)
    rc = lib.gtts_vocoder_forward_host(gen._handle, mel.data_ptr(), out.data_ptr(), 2, 16, 0)
    pkg._lib.check(rc, "vocoder_forward_host")
    assert torch.equal(out, y_dev)


def test_vocoder_rejects_bad_input(pkg, synth):
    gen, cfg, _ = _make(pkg, synth, "v1", 36)
    with pytest.raises(ValueError):
        gen(torch.zeros(1, 79, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        gen(torch.zeros(1, 80, 8))
    lib = pkg._lib.load()
    h = ctypes.c_void_p()
    arr = lambda xs: (ctypes.c_int * len(xs))(*xs)               # noqa: E731
    # upsample kernel smaller than its rate
    assert lib.gtts_vocoder_create(ctypes.byref(h), 1, 1, arr([8]), arr([4]), 64, 1, arr([3]), arr([1]), 1, 80, 0) != 0


def test_synthesize_text_to_waveform(pkg, synth):
    """inference.py:84-97 in one call: tokens -> int16 samples; equals the manual composition of its parts."""
    ecfg = synth.TEXT_ENCODER_CONFIGS["ref"]
    net = pkg.GradTTS(ecfg["n_vocab"], 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000)
    net.encoder.load_state_dict(synth.make_text_encoder_state_dict(ecfg, seed=61), strict=True)
    net.decoder.load_state_dict(synth.make_decoder_state_dict(1, seed=0, g=0.05), strict=True)
    net = net.cuda().eval()
    voc, _, _ = _make(pkg, synth, "v1", 62)
    x, lengths, _ = synth.make_text_inputs(ecfg, 2, 12, seed=63)
    torch.manual_seed(3)
    audio, y_dec, attn = pkg.inference.synthesize(net, voc, x.cuda(), lengths.cuda(), n_timesteps=2)
    assert audio.dtype == torch.int16 and audio.device.type == "cpu" and audio.shape == (2, y_dec.shape[-1] * 256)
    torch.manual_seed(3)
    _, y2, _ = net(x.cuda(), lengths.cuda(), n_timesteps=2, temperature=1.5)
    assert torch.equal(y2, y_dec)
    want = (voc(y2).squeeze(1).clamp(-1, 1) * 32768).to(torch.int16).cpu()
    assert torch.equal(audio, want)
    assert int(audio.abs().max()) > 0


def test_vocoder_varying_lengths_share_one_workspace_pool(pkg, synth):
    """A different mel length on every call (the serving pattern): plans are rebuilt, the workspace is not -- the pool stops growing
    once the longest shape has been seen, the plan cache is LRU-bounded, and results do not depend on what ran before."""
    gen, cfg, _ = _make(pkg, synth, "v1", 37)
    gen.set_option("max_plans", 3)
    mels = {T: synth.make_mel(2, T, seed=100 + T).cuda() for T in (40, 24, 33, 17, 28)}
    first = {T: gen(m) for T, m in mels.items()}
    info1 = gen.cache_info()
    assert info1["plans"] <= 3 and info1["pool_bytes"] > 0
    again = {T: gen(m) for T, m in sorted(mels.items())}
    info2 = gen.cache_info()
    assert info2["pool_bytes"] == info1["pool_bytes"]                # every shape fits the blocks the longest one allocated
    assert all(torch.equal(first[T], again[T]) for T in mels)
    # two streams on one handle are ordered by the library
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        a = gen(mels[40])
    with torch.cuda.stream(s2):
        b = gen(mels[24])
    torch.cuda.synchronize()
    assert torch.equal(a, first[40]) and torch.equal(b, first[24])


def test_vocoder_full_length_utterance_against_oracle(pkg, synth):
    """BASELINE-length mel (1720 frames -> 440 320 samples): every stage runs thousands of tiles per launch and the last stage is
    3 440 position tiles per utterance; against the CPU oracle, both precisions."""
    from oracle import vocoder_oracle
    gen, cfg, sd = _make(pkg, synth, "v1", 38)
    mel = synth.make_mel(1, 1720, seed=8)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = vocoder_oracle.generator_forward(sd, cfg, mel)
    gen.precision = "bf16"
    y = gen(mel.cuda()).cpu()
    assert y.shape == ref.shape == (1, 1, 1720 * 256) and torch.isfinite(y).all()
    assert _rel_rms(y, ref) <= BF16_REL_RMS
    gen.precision = "fp32"
    y = gen(mel.cuda()).cpu()
    assert float((y - ref).abs().max()) <= FP32_MAX_ABS


def test_vocoder_c_abi_writes_only_its_output(pkg, synth):
    """gtts_vocoder_forward through ctypes with guard bands around the output and an untouched-input check"""
    gen, cfg, _ = _make(pkg, synth, "v1", 39)
    B, T = 2, 21
    mel = synth.make_mel(B, T, seed=9).cuda()
    want = gen(mel)
    mel_copy = mel.clone()
    n, guard = B * T * gen.hop, 4096
    buf = torch.full((n + 2 * guard,), -123.0, device="cuda")
    lib = pkg._lib.load()
    rc = lib.gtts_vocoder_forward(gen._handle, mel.data_ptr(), buf[guard:].data_ptr(), B, T, 0,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg._lib.check(rc, "vocoder_forward")
    torch.cuda.synchronize()
    assert torch.equal(buf[guard:guard + n].view(B, 1, -1), want)
    assert bool((buf[:guard] == -123.0).all()) and bool((buf[guard + n:] == -123.0).all())
    assert torch.equal(mel, mel_copy)
