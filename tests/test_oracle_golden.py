"""CPU: the oracle restatements reproduce the golden fixtures made by the real reference."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, decoder_case_names, load_decoder_case
from oracle import build_oracle, decoder_oracle, mas_oracle


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


DEC = decoder_case_names(("est_", "dec_", "c1", "c3"))      # c1*/c3*: BASELINE.json config shapes (C1: 1 x 400 x 10 steps, C3: n_spks=247, 800 frames)
MAS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "mas_*.npz")))


def golden_path(g, B, tx, ty):
    idx = g["idx"].astype(np.int64)
    path = np.zeros((B, tx, ty), dtype=np.int32)
    b, y = np.nonzero(idx >= 0)
    path[b, idx[b, y], y] = 1
    return path


@pytest.mark.parametrize("name", DEC)
def test_decoder_oracle_matches_reference(name, synth):
    torch.set_num_threads(8)
    g = load_decoder_case(name)
    n_spks, n_steps = g["n_spks"], g["n_steps"]
    sd = synth.make_decoder_state_dict(n_spks, seed=g["wseed"], g=0.05)
    z, mask, mu, spk = g["z"], g["mask"], g["mu"], g["spk"]
    with torch.no_grad():
        if n_steps == 0:
            y = decoder_oracle.estimator_forward(sd, z * mask, mask, mu, g["t"], spk, n_spks)
        else:
            y = decoder_oracle.reverse_diffusion(sd, z, mask, mu, n_steps, True, spk, n_spks)
    ref = g["y"]
    # same ATen kernels, same thread count: tolerance covers only op-order noise (|y|max up to ~140)
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    assert float((y - ref).abs().max()) <= tol


def test_synth_weights_are_reproducible(synth):
    import hashlib
    g = _load("dec_spk1_b1_t64_n10")
    sd = synth.make_decoder_state_dict(1, seed=int(g["wseed"]), g=0.05)
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == str(g["sd_sha256"])


@pytest.mark.parametrize("name", MAS)
def test_mas_oracle_matches_reference_golden(name, synth):
    g = _load(name)
    B, tx, ty = (int(v) for v in g["shape"])
    value, mask, txs, tys = synth.make_mas_inputs(B, tx, ty, seed=int(g["seed"]), ragged=bool(g["ragged"]))
    assert abs(float(value.double().sum()) - float(g["value_sum"])) < 1e-6 * abs(float(g["value_sum"]))
    path = mas_oracle.maximum_path(value, mask)
    assert path.dtype == value.dtype
    assert np.array_equal(path.numpy().astype(np.int32), golden_path(g, B, tx, ty))


def test_mas_oracle_matches_compiled_reference(synth):
    """oracle/_ref = the reference's own core.pyx compiled here; skipped where it is absent."""
    build_oracle.build_ref()
    ref = build_oracle.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(0)
    for B, tx, ty in [(3, 7, 7), (4, 1, 5), (6, 17, 40), (2, 50, 300)]:
        value = (5 * rng.standard_normal((B, tx, ty)) - 40).astype(np.float32)
        txs = rng.integers(1, tx + 1, B).astype(np.int32)
        tys = np.maximum(rng.integers(1, ty + 1, B), txs).astype(np.int32)
        v1, v2 = value.copy(), value.copy()
        p1 = np.zeros_like(value, dtype=np.int32)
        p2 = np.zeros_like(value, dtype=np.int32)
        ref.maximum_path_c(p1, v1, txs, tys)
        mas_oracle.maximum_path_c(p2, v2, txs, tys)
        assert np.array_equal(p1, p2)
        assert np.array_equal(v1, v2)          # the in-place DP table is bit-identical too


# ---------------------------------------------------------------------------------------------------------------------
# Alignment stage: oracle restatement vs vectors captured from the reference's GradTTS.compute_loss
@pytest.mark.parametrize("name", ["align_b3_17x61", "align_b2_50x200"])
def test_align_oracle_matches_reference_capture(name):
    import numpy as np
    import torch
    from oracle import align_oracle, mas_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    mu_x, y, x_mask = torch.from_numpy(g["mu_x"]), torch.from_numpy(g["y"]), torch.from_numpy(g["x_mask"])
    lp = align_oracle.log_prior(mu_x, y, 80)
    assert torch.equal(lp, torch.from_numpy(g["log_prior"]))                 # same ops on the same CPU: bit-identical
    attn = mas_oracle.maximum_path(torch.from_numpy(g["log_prior"]), torch.from_numpy(g["mask"]))
    assert torch.equal(attn.to(torch.int8), torch.from_numpy(g["attn"]))
    assert torch.equal(align_oracle.logw_from_path(attn, x_mask), torch.from_numpy(g["logw_"]))
    assert torch.equal(align_oracle.mu_y_from_path(attn, mu_x), torch.from_numpy(g["mu_y"]))


# ---------------------------------------------------------------------------------------------------------------------
# Forward value of the training objective: oracle vs vectors captured from the reference's Diffusion.loss_t
@pytest.mark.parametrize("name", ["loss_spk1_b2_t48", "loss_spk247_b3_t40"])
def test_loss_oracle_matches_reference_capture(name, synth):
    from oracle import loss_oracle
    torch.set_num_threads(8)
    g = _load(name)
    n_spks = int(g["n_spks"])
    x0, mask, mu, t, zm = (torch.from_numpy(g[k]) for k in ("x0", "mask", "mu", "t", "zm"))
    spk = torch.from_numpy(g["spk"]) if "spk" in g else None
    xt, zm2 = loss_oracle.forward_diffusion(x0, mask, mu, t, zm)            # the masked draw is enough: xt is masked anyway
    assert torch.equal(xt, torch.from_numpy(g["xt"])) and torch.equal(zm2, zm)   # same ops, same CPU: bit-identical
    loss = loss_oracle.score_loss(torch.from_numpy(g["est"]), zm, mask, t)
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * float(g["loss"])
    sd = synth.make_decoder_state_dict(n_spks, seed=int(g["wseed"]), g=0.05)
    with torch.no_grad():
        full, _ = loss_oracle.loss_t(sd, x0, mask, mu, t, zm, spk, n_spks)
    assert abs(float(full) - float(g["loss"])) <= 1e-5 * float(g["loss"])


@pytest.mark.parametrize("name", ["vjp_spk1_b2_t48", "vjp_spk247_b2_t40"])
def test_vjp_oracle_matches_reference_autograd(name, synth):
    """oracle/likelihood_oracle.estimator_vjp against the gradient the REAL reference modules gave under torch.autograd."""
    from oracle import likelihood_oracle
    torch.set_num_threads(8)
    g = _load(name)
    n_spks = int(g["n_spks"])
    sd = synth.make_decoder_state_dict(n_spks, seed=int(g["wseed"]), g=0.05)
    t = lambda k: torch.from_numpy(g[k])
    spk = t("spk") if "spk" in g else None
    score, gx = likelihood_oracle.estimator_vjp(sd, t("x"), t("mask"), t("mu"), t("t"), t("v"), spk, n_spks)
    assert float((score - t("score")).abs().max()) <= 2e-5
    assert float((gx - t("gx")).abs().max()) <= 2e-5


def test_likelihood_oracle_matches_reference(synth):
    """oracle/likelihood_oracle.likelihood against the reference's own get_likelihood_fn(..., euler=3) (bpd, prior, delta_logp, z)."""
    from oracle import likelihood_oracle
    torch.set_num_threads(8)
    g = _load("lik_spk1_b2_t48_e3")
    sd = synth.make_decoder_state_dict(1, seed=int(g["wseed"]), g=0.05)
    t = lambda k: torch.from_numpy(g[k])
    bpd, prior, dlogp, z = likelihood_oracle.likelihood(sd, t("y"), t("mask"), t("mu"), int(g["n_euler"]), t("eps"))
    assert float((z - t("z")).abs().max()) <= 1e-4
    assert torch.allclose(dlogp, t("delta_logp"), rtol=1e-4, atol=1e-2)
    assert torch.allclose(prior, t("prior_logp"), rtol=1e-4, atol=1e-2)
    assert torch.allclose(bpd, t("bpd"), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("name", ["grad_spk1_b2_t48", "grad_spk247_b2_t40"])
def test_training_gradient_oracle_matches_reference_backward(name, synth):
    """oracle/loss_oracle.loss_t_grads (torch.autograd through the functional restatement) against loss.backward() through the REAL
    reference modules: loss, d loss / d mu, a digest of every parameter gradient and the full gradient of six tensors."""
    from oracle import loss_oracle
    torch.set_num_threads(8)
    g = _load(name)
    n_spks = int(g["n_spks"])
    sd = synth.make_decoder_state_dict(n_spks, seed=int(g["wseed"]), g=0.05)
    t = lambda k: torch.from_numpy(g[k])
    spk = t("spk") if "spk" in g else None
    loss, grads, gmu, gspk = loss_oracle.loss_t_grads(sd, t("x0"), t("mask"), t("mu"), t("t"), t("zm"), spk, n_spks)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5
    assert float((gmu - t("gmu")).abs().max()) <= 1e-6
    if spk is not None:
        assert float((gspk - t("gspk")).abs().max()) <= 1e-6
    dig = t("grad_digest")
    for i, k in enumerate([str(s) for s in g["grad_names"]]):
        flat = grads[k].reshape(-1)
        assert abs(float(flat.double().sum()) - float(dig[i, 0])) <= 1e-4 * max(1.0, float(dig[i, 1])), k
        n = min(16, flat.numel())
        assert float((flat[:n] - dig[i, 2:2 + n]).abs().max()) <= 2e-5 * max(1.0, float(flat.abs().max())), k
    for k in g.files:
        if k.startswith("full:"):
            ref = t(k)
            assert float((grads[k[5:]] - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max())), k


VOC = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "voc_*.npz")))


@pytest.mark.parametrize("name", VOC)
def test_vocoder_oracle_matches_reference_generator(name, synth):
    """oracle/vocoder_oracle.py vs the reference's hifi-gan Generator (weight-normed state dict loaded, remove_weight_norm, forward)."""
    from oracle import vocoder_oracle
    g = _load(name)
    cfg = synth.VOCODER_CONFIGS[str(g["cfg"])]
    sd = synth.make_vocoder_state_dict(cfg, seed=int(g["wseed"]))
    import hashlib
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == str(g["sd_digest"]), "synthetic vocoder weights differ from the ones the fixture was made with"
    torch.set_num_threads(8)
    with torch.no_grad():
        y = vocoder_oracle.generator_forward(sd, cfg, torch.from_numpy(g["mel"]))
    assert y.shape == g["y"].shape
    err = float((y - torch.from_numpy(g["y"])).abs().max())
    assert err <= 2e-6, err                                     # tolerance: fp32 reassociation inside conv1d only


def test_vocoder_module_has_the_reference_state_dict_layout(pkg, synth):
    """Generator(h).state_dict() carries the reference's weight-norm keys and shapes (strict load of a reference-shaped dict),
    and remove_weight_norm() leaves `weight` = g * v / ||v||."""
    from oracle import vocoder_oracle
    for cfg_name, cfg in synth.VOCODER_CONFIGS.items():
        sd = synth.make_vocoder_state_dict(cfg, seed=5)
        gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
        assert list(gen.state_dict().keys()) == list(sd.keys()), cfg_name
        gen.load_state_dict(sd, strict=True)
        gen.remove_weight_norm()
        sd2 = gen.state_dict()
        for name, shape, _ in synth.vocoder_param_shapes(cfg):
            assert tuple(sd2[name + ".weight"].shape) == tuple(shape)
            assert torch.allclose(sd2[name + ".weight"], vocoder_oracle.effective_weight(sd, name), atol=1e-7)
    with pytest.raises(RuntimeError):
        gen.forward(torch.zeros(1, 80, 8))                       # CPU tensors: no fallback


ENC = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "enc_*.npz")))


def _enc_case(name, synth):
    import hashlib
    g = _load(name)
    cfg = synth.TEXT_ENCODER_CONFIGS[str(g["cfg"])]
    sd = synth.make_text_encoder_state_dict(cfg, seed=int(g["wseed"]))
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == str(g["sd_digest"]), "synthetic text-encoder weights differ from the ones the fixture was made with"
    spk = torch.from_numpy(g["spk"]) if "spk" in g else None
    return g, cfg, sd, torch.from_numpy(g["x"]), torch.from_numpy(g["lengths"]), spk


@pytest.mark.parametrize("name", ENC)
def test_text_encoder_oracle_matches_reference_module(name, synth):
    """oracle/text_encoder_oracle.py vs the reference's TextEncoder (eval mode), incl. T = 3 < window + 1 and the multi-speaker mode."""
    from oracle import text_encoder_oracle
    g, cfg, sd, x, lengths, spk = _enc_case(name, synth)
    torch.set_num_threads(8)
    with torch.no_grad():
        mu, logw, x_mask = text_encoder_oracle.text_encoder_forward(sd, cfg, x, lengths, spk)
    assert torch.equal(x_mask, torch.from_numpy(g["x_mask"]))
    assert float((mu - torch.from_numpy(g["mu"])).abs().max()) <= 2e-5          # fp32 reassociation (matmul vs gather-einsum)
    assert float((logw - torch.from_numpy(g["logw"])).abs().max()) <= 2e-5


def test_text_encoder_module_has_the_reference_state_dict_layout(pkg, synth):
    import importlib
    te = importlib.import_module("grad-tts_b200.model.text_encoder")
    for cfg in synth.TEXT_ENCODER_CONFIGS.values():
        enc = te.TextEncoder(**cfg)
        want = synth.text_encoder_param_shapes(cfg)                                # checked against the reference in make_golden.py
        got = [(k, tuple(v.shape)) for k, v in enc.state_dict().items()]
        assert got == [(k, tuple(s)) for k, s in want]
        enc.eval()
        with pytest.raises(RuntimeError):
            enc(torch.zeros(1, 5, dtype=torch.long), torch.tensor([5]))           # CPU tensors: no fallback
    net = pkg.GradTTS(149, 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000)
    assert isinstance(net.encoder, te.TextEncoder) and net.nparams == 14835032    # the reference's parameter count


@pytest.mark.parametrize("C", [64, 256])
def test_attention_fold_identity_against_the_oracle(C):
    """What `attn_fold` + the per-sample 1x1 conv compute (csrc/attention.cu, include/gradtts_b200.h: gtts_test_attn_fold):
    Residual(Rezero(LinearAttention))(x) = (g * Wout . blockdiag(ctx_b^T) . Wq) x + g * bias + x, with ctx_b the softmax-over-positions
    context of sample b (model/diffusion.py:90-104, 39-46).  Checked in fp64 against the oracle's restatement of the reference module,
    with the layouts the C ABI documents (ctx [b][h][d][e], Wout [co][h*32 + e], Wq [h*32 + d][ci])."""
    g0 = torch.Generator().manual_seed(C)
    b, h, w, heads, dh = 2, 5, 7, 4, 32
    p = "attn"
    sd = {p + ".fn.fn.to_qkv.weight": torch.randn(3 * heads * dh, C, 1, 1, generator=g0, dtype=torch.float64) / C ** 0.5,
          p + ".fn.fn.to_out.weight": torch.randn(C, heads * dh, 1, 1, generator=g0, dtype=torch.float64) / 128 ** 0.5,
          p + ".fn.fn.to_out.bias": torch.randn(C, generator=g0, dtype=torch.float64),
          p + ".fn.g": torch.tensor([0.37], dtype=torch.float64)}
    x = torch.randn(b, C, h, w, generator=g0, dtype=torch.float64)
    ref = decoder_oracle._attn_residual(sd, p, x)

    wqkv = sd[p + ".fn.fn.to_qkv.weight"].view(3, heads * dh, C)
    wq, wk, wv = wqkv[0], wqkv[1], wqkv[2]                              # rows h*32 + d
    xf = x.view(b, C, h * w)
    k = torch.einsum("kc,bcn->bkn", wk, xf).view(b, heads, dh, -1).softmax(dim=-1)
    v = torch.einsum("kc,bcn->bkn", wv, xf).view(b, heads, dh, -1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)                         # [b][h][d][e]
    wout = sd[p + ".fn.fn.to_out.weight"].view(C, heads, dh)            # [co][h][e]
    gain = sd[p + ".fn.g"]
    P = torch.einsum("che,bhde->bchd", wout, ctx).reshape(b, C, heads * dh)
    M = gain * P @ wq                                                   # [b][co][ci]
    got = torch.einsum("boc,bcn->bon", M, xf) + (gain * sd[p + ".fn.fn.to_out.bias"]).view(1, C, 1) + xf
    assert float((got.view_as(ref) - ref).abs().max()) <= 1e-10 * float(ref.abs().max())


@pytest.mark.parametrize("cin", [2, 3])
def test_first_conv_split_arithmetic(cin):
    """The arithmetic of the bf16-mode first conv (csrc/pointwise.cu: first_conv_mma_kernel; reference op model/diffusion.py:52 on
    stack([mu, x, (s)]) * mask, :181-184): every fp32 input is split into bf16 hi + lo (x = hi + lo to ~2^-17 relative), the weights
    are rounded to bf16 like those of every other conv of the bf16 mode, products are exact and summed in fp32.  Emulated on the CPU:
    the split alone is fp32-accurate (<= 1e-4 of the output scale); with bf16 weights the error against the fp32 conv is the
    weight rounding (<= 2^-8 relative to sum |w| |x| per output), far below the bf16 activations that follow."""
    import torch.nn.functional as F
    g0 = torch.Generator().manual_seed(7 + cin)
    x = torch.randn(2, cin, 80, 24, generator=g0) * 3.0
    w = torch.randn(64, cin, 3, 3, generator=g0) / (9 * cin) ** 0.5
    bias = torch.randn(64, generator=g0)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    hi = x.bfloat16().float()
    lo = (x - hi).bfloat16().float()
    assert float((hi + lo - x).abs().max()) <= 2.0 ** -16 * float(x.abs().max())
    split_only = F.conv2d(hi.double(), w.double(), bias.double(), padding=1) + F.conv2d(lo.double(), w.double(), None, padding=1)
    assert float((split_only - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    wb = w.bfloat16().double()
    got = F.conv2d(hi.double(), wb, bias.double(), padding=1) + F.conv2d(lo.double(), wb, None, padding=1)
    bound = 2.0 ** -8 * F.conv2d(x.abs().double(), w.abs().double(), None, padding=1) + 1e-4 * float(ref.abs().max())
    assert bool(((got - ref).abs() <= bound).all())
