"""Generate the golden fixtures in tests/golden/ by running the REAL reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py          # everything
    python tests/golden/make_golden.py loss     # only the loss_t vectors
    python tests/golden/make_golden.py NAME...  # only the named decoder cases
It copies /root/reference/model to a temp dir, builds the reference's Cython extension with the
reference's own setup.py (README.md:31), imports `model`, loads seeded synthetic weights
(grad-tts_b200/synth.py) with strict=True into the reference modules, runs them in fp32 on CPU
(torch threads fixed to 8) and stores inputs + outputs as small .npz files.
"""
import hashlib
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("grad-tts_b200.synth")


def import_reference():
    tmp = tempfile.mkdtemp(prefix="gtts_ref_")
    shutil.copytree("/root/reference/model", os.path.join(tmp, "model"))
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"],
                          cwd=os.path.join(tmp, "model", "monotonic_align"),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sys.path.insert(0, tmp)
    import model  # noqa: F401
    return tmp


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def decoder_cases():
    # name, n_spks, B, T, n_steps (0 = single estimator call), weight seed, input seed
    return [
        ("est_spk1_b2_t48", 1, 2, 48, 0, 0, 1),
        ("est_spk247_b2_t40", 247, 2, 40, 0, 3, 4),
        ("dec_spk1_b1_t64_n10", 1, 1, 64, 10, 0, 5),
        ("dec_spk1_b3_t56_n4", 1, 3, 56, 4, 7, 8),
        ("dec_spk247_b2_t40_n3", 247, 2, 40, 3, 3, 6),
        # third mode (params_tedlium.py:22-23): n_spks = -1, spk_mlp exists and is evaluated, the U-Net takes two channels
        ("dec_spkm1_b2_t32_n3", -1, 2, 32, 3, 9, 10),
    ]


def baseline_shape_cases():
    """Fixtures at the shapes BASELINE.json names.  Only the reference OUTPUT is stored (plus a digest of the inputs): the
    inputs are regenerated from the seed by grad-tts_b200/synth.py (CPU torch.Generator), like the MAS fixtures.
    name, n_spks, B, T, n_steps, weight seed, input seed, ragged"""
    return [
        # C1: LJSpeech single speaker, batch 1 x 400 frames, 10 Euler steps -- the exact weights/inputs bench.py uses
        ("c1_spk1_b1_t400_n10", 1, 1, 400, 10, 0, 1, False),
        # C1 shape with a second, shorter utterance in the batch (padding participates in GroupNorm / softmax)
        ("c1r_spk1_b2_t400_n10", 1, 2, 400, 10, 0, 21, True),
        # C3 shape: Libri-TTS multispeaker (n_spks=247, speaker channel), 800 frames, stoc=True passed (a no-op in the fork)
        ("c3_spk247_b2_t800_n3", 247, 2, 800, 3, 0, 22, True),
    ]


def tensors_digest(*ts):
    h = hashlib.sha256()
    for t in ts:
        if t is not None:
            h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def make_baseline_shape_vectors(only=()):
    from model.diffusion import Diffusion
    for name, n_spks, B, T, n_steps, wseed, iseed, ragged in baseline_shape_cases():
        if only and name not in only:
            continue
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        z, mask, mu, spk, lengths = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=ragged)
        with torch.no_grad():
            y = dec(z, mask, mu, n_steps, True, spk)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), y=y.numpy(), n_spks=np.int64(n_spks), n_steps=np.int64(n_steps),
                            wseed=np.int64(wseed), iseed=np.int64(iseed), ragged=np.bool_(ragged), shape=np.array([B, T]),
                            lengths=lengths.numpy(), sd_sha256=np.array(sd_digest(sd)),
                            in_sha256=np.array(tensors_digest(z, mask, mu, spk)))
        print(name, "y absmax", float(y.abs().max()))


def loss_cases():
    # name, n_spks, B, T, weight seed, input seed
    return [("loss_spk1_b2_t48", 1, 2, 48, 0, 41), ("loss_spk247_b3_t40", 247, 3, 40, 3, 42)]


def make_loss_vectors():
    """Diffusion.loss_t (model/diffusion.py:274-281) on seeded inputs: the noise the reference drew (captured from
    forward_diffusion's second return value), xt, the estimator output it saw and the loss."""
    from model.diffusion import Diffusion
    for name, n_spks, B, T, wseed, iseed in loss_cases():
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        x0, mask, mu, spk, lengths = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)
        gen = torch.Generator().manual_seed(iseed + 100)
        t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
        cap = {}
        orig_fd = dec.forward_diffusion

        def fd(*a):
            xt, zm = orig_fd(*a)
            cap["xt"], cap["zm"] = xt.clone(), zm.clone()
            return xt, zm

        dec.forward_diffusion = fd
        hook = dec.estimator.register_forward_hook(lambda m, i, o: cap.__setitem__("est", o.clone()))
        torch.manual_seed(iseed + 200)
        with torch.no_grad():
            loss, xt = dec.loss_t(x0, mask, mu, t, spk)
        hook.remove()
        assert torch.equal(xt, cap["xt"])
        out = dict(x0=x0.numpy(), mask=mask.numpy(), mu=mu.numpy(), t=t.numpy(), zm=cap["zm"].numpy(), xt=xt.numpy(),
                   est=cap["est"].numpy(), loss=np.float32(loss.item()), n_spks=np.int64(n_spks), wseed=np.int64(wseed),
                   sd_sha256=np.array(sd_digest(sd)))
        if spk is not None:
            out["spk"] = spk.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "loss", float(loss))


def vjp_cases():
    # name, n_spks, B, T, weight seed, input seed
    return [("vjp_spk1_b2_t48", 1, 2, 48, 0, 51), ("vjp_spk247_b2_t40", 247, 2, 40, 3, 52)]


def make_vjp_vectors():
    """The gradient the reference's likelihood code takes through the score network (n_best/likelihood/likelihood.py:30-34):
    torch.autograd.grad(sum(estimator(x, mask, mu, t, spk) * v), x) on the real reference modules, seeded inputs."""
    from model.diffusion import Diffusion
    for name, n_spks, B, T, wseed, iseed in vjp_cases():
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        z, mask, mu, spk, _ = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)
        gen = torch.Generator().manual_seed(iseed + 100)
        t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
        v = torch.randn(B, 80, T, generator=gen)
        x = (z * mask).clone().requires_grad_(True)
        score = dec.estimator(x, mask, mu, t, spk)
        gx = torch.autograd.grad(torch.sum(score * v), x)[0]
        out = dict(x=x.detach().numpy(), mask=mask.numpy(), mu=mu.numpy(), t=t.numpy(), v=v.numpy(), score=score.detach().numpy(),
                   gx=gx.numpy(), n_spks=np.int64(n_spks), wseed=np.int64(wseed), sd_sha256=np.array(sd_digest(sd)))
        if spk is not None:
            out["spk"] = spk.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "gx absmax", float(gx.abs().max()), "score absmax", float(score.abs().max()))


def make_train_grad_vectors():
    """loss.backward() through the REAL reference Diffusion.loss_t (model/diffusion.py:274-281) in train mode: the loss, a digest of
    every parameter gradient (sum, sum of absolute values, the first 16 entries) and the full gradient of a few tensors, d loss / d mu."""
    from model.diffusion import Diffusion
    for name, n_spks, B, T, wseed, iseed in [("grad_spk1_b2_t48", 1, 2, 48, 0, 71), ("grad_spk247_b2_t40", 247, 2, 40, 3, 72)]:
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).train()
        dec.load_state_dict(sd, strict=True)
        x0, mask, mu, spk, _ = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)
        gen = torch.Generator().manual_seed(iseed + 100)
        t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
        cap = {}
        orig_fd = dec.forward_diffusion

        def fd(*a):
            xt, zm = orig_fd(*a)
            cap["zm"] = zm.clone()
            return xt, zm

        dec.forward_diffusion = fd
        mu_g = mu.clone().requires_grad_(True)
        spk_g = spk.clone().requires_grad_(True) if spk is not None else None
        torch.manual_seed(iseed + 200)
        loss, _ = dec.loss_t(x0, mask, mu_g, t, spk_g)
        loss.backward()
        out = dict(x0=x0.numpy(), mask=mask.numpy(), mu=mu.numpy(), t=t.numpy(), zm=cap["zm"].detach().numpy(), loss=np.float32(loss.item()),
                   gmu=mu_g.grad.numpy(), n_spks=np.int64(n_spks), wseed=np.int64(wseed), sd_sha256=np.array(sd_digest(sd)))
        if spk is not None:
            out["spk"] = spk.numpy()
            out["gspk"] = spk_g.grad.numpy()
        names, dig = [], []
        for k, p in dec.state_dict(keep_vars=True).items():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            names.append(k)
            flat = g.reshape(-1)
            head = torch.zeros(16)
            head[: min(16, flat.numel())] = flat[:16]
            dig.append(torch.cat([flat.double().sum().float().reshape(1), flat.abs().double().sum().float().reshape(1), head]))
        out["grad_names"] = np.array(names)
        out["grad_digest"] = torch.stack(dig).numpy()
        for k in ("estimator.downs.1.0.block1.block.0.weight", "estimator.mid_attn.fn.fn.to_qkv.weight", "estimator.ups.1.3.conv.weight",
                  "estimator.downs.0.0.block1.block.0.weight", "estimator.mlp.0.weight", "estimator.final_conv.weight"):
            out["full:" + k] = dict(dec.named_parameters())[k].grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "loss", float(loss), "params", len(names))


def make_likelihood_vectors():
    """The reference's probability-flow likelihood (n_best/likelihood/likelihood.py get_likelihood_fn with euler > 0, SPEECHSDE from
    sde_lib.py) driven exactly as n_best/get_score_parallel.py:77-83 does, on the real reference estimator with seeded weights."""
    import types
    for m in ("matplotlib", "matplotlib.pyplot"):                 # likelihood.py imports pyplot at module level and never uses it
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, "/root/reference/n_best")
    from likelihood import likelihood as ref_lik, sde_lib as ref_sde
    from model.diffusion import Diffusion
    for name, n_spks, B, T, n_euler, wseed, iseed in [("lik_spk1_b2_t48_e3", 1, 2, 48, 3, 0, 61)]:
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        y, mask, mu, spk, _ = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)

        class ScoreModel(torch.nn.Module):                         # model/tts.py:237-250
            def forward(self, x, t):
                return dec.estimator(x=x, mask=mask, mu=mu, t=t, spk=spk)

        sde = ref_sde.SPEECHSDE(beta_min=0.05, beta_max=20.0, N=1000, mu=mu, spk=spk, mask=mask)
        fn = ref_lik.get_likelihood_fn(sde, lambda x: x, rtol=1e-3, atol=1e-3, euler=n_euler)
        torch.manual_seed(iseed + 7)
        eps_draw = torch.randint_like(y, low=0, high=2).float() * 2 - 1.0        # the draw likelihood_fn makes first (likelihood.py:87)
        torch.manual_seed(iseed + 7)
        bpd, prior_logp, delta_logp, z = fn(ScoreModel(), y)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), y=y.numpy(), mask=mask.numpy(), mu=mu.numpy(), eps=eps_draw.numpy(),
                            bpd=bpd.numpy(), prior_logp=prior_logp.numpy(), delta_logp=delta_logp.numpy(), z=z.numpy(),
                            n_spks=np.int64(n_spks), n_euler=np.int64(n_euler), wseed=np.int64(wseed), sd_sha256=np.array(sd_digest(sd)))
        print(name, "bpd", bpd.tolist(), "delta_logp", delta_logp.tolist())


def vocoder_cases():
    # name, config, B, T, weight seed, input seed
    return [("voc_v1_b2_t24", "v1", 2, 24, 21, 22), ("voc_v1_b1_t100", "v1", 1, 100, 23, 24), ("voc_rb2_b2_t16", "rb2", 2, 16, 25, 26)]


def make_vocoder_vectors():
    """hifi-gan/models.py::Generator run as inference.py:73-76,97 does: load_state_dict, eval, remove_weight_norm, forward."""
    import types
    for name in ("matplotlib", "matplotlib.pylab"):             # hifi-gan/xutils.py imports matplotlib for its plot helper only
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    sys.path.insert(0, "/root/reference/hifi-gan")
    sys.dont_write_bytecode = True
    from models import Generator
    from env import AttrDict
    for name, cfg_name, B, T, wseed, iseed in vocoder_cases():
        cfg = synth.VOCODER_CONFIGS[cfg_name]
        sd = synth.make_vocoder_state_dict(cfg, seed=wseed)
        gen = Generator(AttrDict(cfg))
        gen.load_state_dict(sd, strict=True)
        gen.eval()
        gen.remove_weight_norm()
        mel = synth.make_mel(B, T, seed=iseed)
        with torch.no_grad():
            y = gen(mel)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mel=mel.numpy(), y=y.numpy(), cfg=cfg_name,
                            wseed=wseed, sd_digest=sd_digest(sd))
        print(name, tuple(y.shape), float(y.abs().mean()), float(y.abs().max()))


def text_encoder_cases():
    # name, config, B, T, weight seed, input seed
    return [("enc_ref_b2_t37", "ref", 2, 37, 41, 42), ("enc_ref_b1_t3", "ref", 1, 3, 41, 43), ("enc_spk_b3_t20", "spk", 3, 20, 44, 45)]


def make_text_encoder_vectors():
    """model/text_encoder.py::TextEncoder in eval mode, constructed as model/tts.py:49-51 does ("ref") and with its own
    multi-speaker arguments ("spk")."""
    import contextlib
    import io
    from model.text_encoder import TextEncoder
    for name, cfg_name, B, T, wseed, iseed in text_encoder_cases():
        cfg = synth.TEXT_ENCODER_CONFIGS[cfg_name]
        sd = synth.make_text_encoder_state_dict(cfg, seed=wseed)
        with contextlib.redirect_stdout(io.StringIO()):          # the reference constructor prints its channel counts
            enc = TextEncoder(cfg["n_vocab"], cfg["n_feats"], cfg["n_channels"], cfg["filter_channels"], cfg["filter_channels_dp"],
                              cfg["n_heads"], cfg["n_layers"], cfg["kernel_size"], cfg["p_dropout"], cfg["window_size"],
                              cfg["spk_emb_dim"], cfg["n_spks"])
        assert list(enc.state_dict().keys()) == list(sd.keys()), "synth.text_encoder_param_shapes is out of step with the reference"
        enc.load_state_dict(sd, strict=True)
        enc.eval()
        x, lengths, spk = synth.make_text_inputs(cfg, B, T, seed=iseed)
        with torch.no_grad():
            mu, logw, x_mask = enc(x, lengths, spk)
        out = dict(x=x.numpy(), lengths=lengths.numpy(), mu=mu.numpy(), logw=logw.numpy(), x_mask=x_mask.numpy(), cfg=cfg_name,
                   wseed=wseed, sd_digest=sd_digest(sd))
        if spk is not None:
            out["spk"] = spk.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, tuple(mu.shape), float(mu.abs().mean()), float(logw.abs().mean()), float(torch.exp(logw).max()))


def main():
    torch.set_num_threads(8)
    if sys.argv[1:] == ["voc"]:
        return make_vocoder_vectors()
    if sys.argv[1:] == ["enc"]:
        import_reference()
        return make_text_encoder_vectors()
    import_reference()
    if sys.argv[1:] == ["loss"]:
        return make_loss_vectors()
    if sys.argv[1:] == ["vjp"]:
        make_vjp_vectors()
        return make_likelihood_vectors()
    if sys.argv[1:] == ["grad"]:
        return make_train_grad_vectors()
    only = set(sys.argv[1:])                                   # optional: names of decoder cases to (re)generate
    make_baseline_shape_vectors(only)
    if not only:
        make_vocoder_vectors()
        make_text_encoder_vectors()
        make_loss_vectors()
        make_vjp_vectors()
        make_likelihood_vectors()
        make_train_grad_vectors()
    from model.diffusion import Diffusion
    from model.monotonic_align import maximum_path

    for name, n_spks, B, T, n_steps, wseed, iseed in decoder_cases():
        if only and name not in only:
            continue
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        z, mask, mu, spk, lengths = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)
        if n_spks == -1:
            spk = torch.randn(B, 64, generator=torch.Generator().manual_seed(iseed + 300))
        out = {}
        with torch.no_grad():
            if n_steps == 0:
                gen = torch.Generator().manual_seed(iseed + 100)
                t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
                y = dec.estimator(z * mask, mask, mu, t, spk)
                out["t"] = t.numpy()
            else:
                y = dec(z, mask, mu, n_steps, False, spk)
                y_stoc = dec(z, mask, mu, n_steps, True, spk)
                assert torch.equal(y, y_stoc), "reference fork: stoc flag must be a no-op"
        out.update(z=z.numpy(), mask=mask.numpy(), mu=mu.numpy(), y=y.numpy(),
                   n_spks=np.int64(n_spks), n_steps=np.int64(n_steps), wseed=np.int64(wseed),
                   sd_sha256=np.array(sd_digest(sd)))
        if spk is not None:
            out["spk"] = spk.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "y absmax", float(y.abs().max()))
    if only:
        return

    # Alignment stage (tts.py:139-185): vectors captured from the reference's own GradTTS.compute_loss -- mu_x from the encoder,
    # the log_prior / mask handed to maximum_path, the path it returned, logw_ (argument of duration_loss) and mu_y (argument of
    # decoder.compute_loss).  Random-init reference model, seeded inputs.
    import model.tts as ref_tts
    for name, B, n_vocab, tx, ty, seed in [("align_b3_17x61", 3, 40, 17, 61, 31), ("align_b2_50x200", 2, 60, 50, 200, 32)]:
        torch.manual_seed(seed)
        net = ref_tts.GradTTS(n_vocab, 1, 64, 192, 768, 256, 2, 2, 3, 0.0, 4, 80, 64, 0.05, 20.0, 1000).eval()
        g = torch.Generator().manual_seed(seed + 1)
        x_len = torch.randint(max(1, tx // 2), tx + 1, (B,), generator=g); x_len[0] = tx
        y_len = torch.randint(max(tx, ty // 2), ty + 1, (B,), generator=g); y_len[0] = ty
        x = torch.randint(0, n_vocab, (B, tx), generator=g)
        y = torch.randn(B, 80, ty, generator=g)
        cap = {}
        orig_mp, orig_dl, orig_cl = ref_tts.monotonic_align.maximum_path, ref_tts.duration_loss, net.decoder.compute_loss

        def mp(value, mask):
            cap["log_prior"], cap["mask"] = value.clone(), mask.clone()
            cap["attn"] = orig_mp(value, mask)
            return cap["attn"]

        def dl(logw, logw_, lengths):
            cap["logw_"] = logw_.clone()
            return orig_dl(logw, logw_, lengths)

        def cl(y_, y_mask, mu_y, spk=None):
            cap["mu_y"] = mu_y.clone()
            return torch.zeros(()), y_

        hook = net.encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc", [t.clone() for t in o]))
        ref_tts.monotonic_align.maximum_path, ref_tts.duration_loss, net.decoder.compute_loss = mp, dl, cl
        try:
            with torch.no_grad():
                net.compute_loss(x, x_len, y, y_len, spk=None, out_size=None)
        finally:
            ref_tts.monotonic_align.maximum_path, ref_tts.duration_loss, net.decoder.compute_loss = orig_mp, orig_dl, orig_cl
            hook.remove()
        mu_x, _, x_mask = cap["enc"]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), mu_x=mu_x.numpy(), x_mask=x_mask.numpy(), y=y.numpy(),
                            y_len=y_len.numpy(), x_len=x_len.numpy(), log_prior=cap["log_prior"].numpy(),
                            mask=cap["mask"].numpy(), attn=cap["attn"].numpy().astype(np.int8), logw_=cap["logw_"].numpy(),
                            mu_y=cap["mu_y"].numpy())
        print(name, "log_prior absmax", float(cap["log_prior"].abs().max()), "path cells", int(cap["attn"].sum()))

    # MAS: value/mask regenerated from the seed in the tests; only the int8 path is stored.
    mas_cases = [("mas_b4_20x50", 4, 20, 50, 11, True), ("mas_b3_33x33", 3, 33, 33, 12, False),
                 ("mas_b5_1x9", 5, 1, 9, 13, True), ("mas_b2_64x257", 2, 64, 257, 14, True),
                 ("mas_b64_200x1000", 64, 200, 1000, 1234, True)]
    for name, B, tx, ty, seed, ragged in mas_cases:
        value, mask, txs, tys = synth.make_mas_inputs(B, tx, ty, seed=seed, ragged=ragged)
        path = maximum_path(value, mask)
        assert path.dtype == value.dtype
        p8 = path.numpy().astype(np.int8)
        # store the path compactly: for every (b, y) the row index of the single 1 (or -1)
        idx = np.where(p8.sum(1) > 0, p8.argmax(1), -1).astype(np.int16)      # (B, t_y)
        assert (p8.sum(1) <= 1).all()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), idx=idx, tx=txs.numpy(), ty=tys.numpy(),
                            seed=np.int64(seed), ragged=np.bool_(ragged),
                            shape=np.array([B, tx, ty]), value_sum=np.float64(value.double().sum()))
        print(name, "ones", int(p8.sum()))


if __name__ == "__main__":
    main()
