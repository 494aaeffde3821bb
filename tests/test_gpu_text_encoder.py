"""GPU: the text encoder (csrc/text_encoder.cu behind gtts_encoder_*) against golden vectors made by the reference's own
model/text_encoder.py::TextEncoder, and against the oracle at other sizes."""
import glob
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

ENC = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "enc_*.npz")))

# fp32 FFMA against the reference's fp32 CPU run: the values are O(1) after 6 layers of LayerNorm; 1e-4 max-abs
MAX_ABS = 1e-4


def _make(synth, cfg_name, wseed):
    te = importlib.import_module("grad-tts_b200.model.text_encoder")
    cfg = synth.TEXT_ENCODER_CONFIGS[cfg_name]
    sd = synth.make_text_encoder_state_dict(cfg, seed=wseed)
    enc = te.TextEncoder(**cfg)
    enc.load_state_dict(sd, strict=True)
    return enc.cuda().eval(), cfg, sd


@pytest.mark.parametrize("name", ENC)
def test_text_encoder_matches_reference_golden(name, synth):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    enc, cfg, _ = _make(synth, str(g["cfg"]), int(g["wseed"]))
    spk = torch.from_numpy(g["spk"]).cuda() if "spk" in g else None
    mu, logw, x_mask = enc(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["lengths"]).cuda(), spk)
    assert torch.equal(x_mask.cpu(), torch.from_numpy(g["x_mask"]))
    assert float((mu.cpu() - torch.from_numpy(g["mu"])).abs().max()) <= MAX_ABS
    assert float((logw.cpu() - torch.from_numpy(g["logw"])).abs().max()) <= MAX_ABS
    # the integer durations the sampler derives from logw (model/tts.py:86-88)
    w_ref = torch.ceil(torch.exp(torch.from_numpy(g["logw"])) * torch.from_numpy(g["x_mask"]))
    w = torch.ceil(torch.exp(logw.cpu()) * x_mask.cpu())
    assert torch.equal(w, w_ref)
    assert enc.launches_last_call() > 0


@pytest.mark.parametrize("B,T", [(1, 1), (3, 2), (2, 130), (4, 300)])
def test_text_encoder_against_oracle(B, T, synth):
    """one token; lengths below the relative window; a length past one 64-position conv tile and the 32-position projection tile;
    a long ragged batch"""
    from oracle import text_encoder_oracle
    enc, cfg, sd = _make(synth, "ref", 51)
    x, lengths, _ = synth.make_text_inputs(cfg, B, T, seed=52 + T)
    if B > 1:
        lengths[-1] = 1
    torch.set_num_threads(8)
    with torch.no_grad():
        mu_r, logw_r, mask_r = text_encoder_oracle.text_encoder_forward(sd, cfg, x, lengths)
    mu, logw, x_mask = enc(x.cuda(), lengths.cuda())
    assert torch.equal(x_mask.cpu(), mask_r)
    assert float((mu.cpu() - mu_r).abs().max()) <= MAX_ABS
    assert float((logw.cpu() - logw_r).abs().max()) <= MAX_ABS
    # padded positions are exactly zero, as in the reference
    pad = (mask_r == 0).expand_as(mu_r)
    assert float(mu.cpu()[pad].abs().max() if pad.any() else 0.0) == 0.0


def test_text_encoder_batch_invariance(synth):
    """an utterance gives the same bits alone and inside a longer, padded batch"""
    enc, cfg, _ = _make(synth, "ref", 53)
    x, lengths, _ = synth.make_text_inputs(cfg, 3, 40, seed=54)
    lengths[1] = 17
    mu, logw, _ = enc(x.cuda(), lengths.cuda())
    mu1, logw1, _ = enc(x[1:2, :17].cuda(), lengths[1:2].cuda())
    assert torch.equal(mu[1, :, :17], mu1[0]) and torch.equal(logw[1, :, :17], logw1[0])


def test_text_encoder_rejects_bad_input(synth):
    enc, cfg, _ = _make(synth, "ref", 55)
    with pytest.raises(RuntimeError):
        enc(torch.full((1, 4), cfg["n_vocab"], dtype=torch.long, device="cuda"), torch.tensor([4], device="cuda"))   # id out of range
    enc.train()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 4, dtype=torch.long, device="cuda"), torch.tensor([4], device="cuda"))
    enc2, cfg2, _ = _make(synth, "spk", 56)
    with pytest.raises(ValueError):
        enc2(torch.zeros(1, 4, dtype=torch.long, device="cuda"), torch.tensor([4], device="cuda"))                    # spk missing


def test_gradtts_forward_with_native_encoder_matches_injected_oracle_encoder(pkg, synth):
    """GradTTS.forward (model/tts.py:54-108) text -> mel with the native encoder vs the same model with a PyTorch encoder built on
    the oracle: same durations, same alignment, same z draw, decoder output within the bf16 tolerance of two identical runs."""
    from oracle import text_encoder_oracle
    cfg = synth.TEXT_ENCODER_CONFIGS["ref"]
    sd_enc = synth.make_text_encoder_state_dict(cfg, seed=57)
    sd_dec = synth.make_decoder_state_dict(1, seed=0, g=0.05)

    class OracleEncoder(torch.nn.Module):
        def forward(self, x, x_lengths, spk=None):
            # on the CPU in fp32: cuDNN's default TF32 convolutions move logw by ~1e-3, enough to flip a ceil()
            out = text_encoder_oracle.text_encoder_forward(sd_enc, cfg, x.cpu(), x_lengths.cpu(), spk)
            return tuple(t.to(x.device) for t in out)

    args = (cfg["n_vocab"], 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000)
    outs = []
    x, lengths, _ = synth.make_text_inputs(cfg, 2, 24, seed=58)
    for native in (True, False):
        net = pkg.GradTTS(*args) if native else pkg.GradTTS(*args, encoder=OracleEncoder())
        if native:
            net.encoder.load_state_dict(sd_enc, strict=True)
        net.decoder.load_state_dict(sd_dec, strict=True)
        net = net.cuda().eval()
        torch.manual_seed(7)
        outs.append(net(x.cuda(), lengths.cuda(), n_timesteps=3, temperature=1.5))
    (enc_a, dec_a, attn_a), (enc_b, dec_b, attn_b) = outs
    assert attn_a.shape == attn_b.shape and torch.equal(attn_a, attn_b)
    assert float((enc_a - enc_b).abs().max()) <= MAX_ABS
    rel = float(((dec_a - dec_b).pow(2).mean() / dec_b.pow(2).mean()).sqrt())
    assert rel <= 2e-2, rel


def test_text_encoder_c_abi_writes_only_its_outputs(pkg, synth):
    """gtts_encoder_forward through ctypes with guard bands around the three outputs"""
    import ctypes
    enc, cfg, _ = _make(synth, "ref", 59)
    B, T = 3, 19
    x, lengths, _ = synth.make_text_inputs(cfg, B, T, seed=60)
    xd, ld = x.cuda(), lengths.cuda()
    mu, logw, x_mask = enc(xd, ld)
    guard = 1024
    sizes = (B * 80 * T, B * T, B * T)
    bufs = [torch.full((n + 2 * guard,), -77.0, device="cuda") for n in sizes]
    lib = pkg._lib.load()
    rc = lib.gtts_encoder_forward(enc._handle, xd.data_ptr(), ld.data_ptr(), None, bufs[0][guard:].data_ptr(), bufs[1][guard:].data_ptr(),
                                  bufs[2][guard:].data_ptr(), B, T, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg._lib.check(rc, "encoder_forward")
    torch.cuda.synchronize()
    for buf, n, ref in zip(bufs, sizes, (mu, logw, x_mask)):
        assert torch.equal(buf[guard:guard + n], ref.reshape(-1))
        assert bool((buf[:guard] == -77.0).all()) and bool((buf[guard + n:] == -77.0).all())
    assert torch.equal(xd.cpu(), x)
