"""First-contact diagnostics on the GPU box: every stage prints numbers instead of asserting, so one run tells
what works.  Usage: python tools/gpu_diag.py <stage>   (stages: mas conv_ffma conv_tc dec_fp32 dec_bf16_ffma dec_bf16_tc perf)"""
import importlib
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

pkg = importlib.import_module("grad-tts_b200")
from oracle import decoder_oracle, mas_oracle  # noqa: E402

DEV = "cuda:0"


def stage_mas():
    for B, tx, ty in [(2, 7, 9), (4, 50, 120), (64, 200, 1000)]:
        value, mask, _, _ = pkg.synth.make_mas_inputs(B, tx, ty, seed=5)
        ref = mas_oracle.maximum_path(value, mask)
        v, m = value.to(DEV), mask.to(DEV)
        got = pkg.maximum_path(v, m).cpu()
        print(f"mas {B}x{tx}x{ty}: equal={torch.equal(ref, got)} ones ref/got {int(ref.sum())}/{int(got.sum())}")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            pkg.maximum_path(v, m, check=False)
        torch.cuda.synchronize()
        print(f"   wrapper time {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")


def stage_conv(impl, act):
    import gpu_util as gu
    for name, kw in gu.CONV_CASES:
        try:
            c = gu.conv_case(seed=hash(name) % 1000, **kw)
            stats_ok = c["kind"] in (0, 1) and c["r"] is None and c["m"] is None and not c["per_sample"]
            out, st = gu.run_conv(c, impl, act, want_stats=stats_ok)
            ref, raw = gu.conv_reference(c, round_bf16=bool(act))
            err = float((out - ref).abs().max())
            nan = int(torch.isnan(out).sum())
            serr = -1.0
            if stats_ok:
                sref = gu.gn_stats_reference(raw)
                serr = float(((st - sref).abs() / (sref.abs() + 1.0)).max())
            print(f"conv impl={impl} act={act} {name}: max-abs err {err:.3e} nan {nan} stats-err {serr:.3e} "
                  f"|ref|max {float(ref.abs().max()):.2f}")
            if err > 0.1 or nan:
                d = (out - ref).abs()
                d[torch.isnan(d)] = 1e9
                bad = (d > 0.1)
                idx = bad.nonzero()
                print(f"     bad {int(bad.sum())} of {bad.numel()}; first {idx[:6].tolist()} last {idx[-3:].tolist()}")
                print("     per-channel-bad", bad.sum((0, 2, 3))[:16].tolist(), "per-row-bad", bad.sum((0, 1, 3))[:24].tolist())
                print("     per-col-bad", bad.sum((0, 1, 2))[:40].tolist())
        except Exception:
            print(f"conv impl={impl} act={act} {name}: EXCEPTION")
            traceback.print_exc()


def _decoder(n_spks, wseed, precision, conv_impl=None):
    sd = pkg.synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd)
    dec = dec.to(DEV)
    dec.precision = precision
    if conv_impl is not None:
        h = dec.estimator._get_handle()
        pkg._lib.check(pkg._lib.load().gtts_decoder_set_option(h, b"conv_impl_bf16", conv_impl), "opt")
    return dec, sd


def stage_dec(precision, conv_impl=None):
    torch.set_num_threads(os.cpu_count() or 8)
    for n_spks, B, T, n in [(1, 2, 48, 0), (247, 2, 40, 0), (1, 1, 64, 10), (247, 2, 40, 3)]:
        try:
            dec, sd = _decoder(n_spks, 0, precision, conv_impl)
            z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=7)
            d = lambda t: None if t is None else t.to(DEV)
            with torch.no_grad():
                if n == 0:
                    t = torch.tensor([0.3, 0.9][:B])
                    ref = decoder_oracle.estimator_forward(sd, z * mask, mask, mu, t, spk, n_spks)
                    got = dec.estimator(d(z * mask), d(mask), d(mu), d(t), d(spk)).cpu()
                else:
                    ref = decoder_oracle.reverse_diffusion(sd, z, mask, mu, n, False, spk, n_spks)
                    got = dec(d(z), d(mask), d(mu), n, False, d(spk)).cpu()
            err = float((got - ref).abs().max())
            rr = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
            print(f"dec {precision} impl={conv_impl} n_spks={n_spks} B={B} T={T} steps={n}: max-abs {err:.3e} "
                  f"rel-rms {rr:.3e} |ref|max {float(ref.abs().max()):.2f} nan {int(torch.isnan(got).sum())} "
                  f"launches {dec.estimator.launches_last_call()}")
        except Exception:
            print(f"dec {precision} impl={conv_impl} n_spks={n_spks}: EXCEPTION")
            traceback.print_exc()


def stage_perf():
    cfgs = [(1, 1, 400, 10, 8), (1, 8, 800, 4, 8), (247, 32, 800, 2, 8), (1, 16, 1720, 2, 4), (1, 16, 1720, 2, 8),
            (1, 16, 1720, 2, 16)]
    if os.environ.get("GTTS_PERF") == "chunks":
        cfgs = [(1, 64, 1720, 2, 16), (1, 64, 1720, 2, 32), (1, 100, 400, 10, 25), (1, 100, 400, 10, 32), (247, 32, 800, 4, 16), (247, 32, 800, 4, 32)]
    for (n_spks, B, T, n, chunk) in cfgs:
        try:
            dec, _ = _decoder(n_spks, 0, "bf16")
            dec.estimator.max_chunk = chunk
            z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=7, ragged=False)
            d = lambda t: None if t is None else t.to(DEV)
            a = (d(z), d(mask), d(mu))
            s = d(spk)
            dec(*a, n, False, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec(*a, n, False, s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            fs = B * T * n / (ms * 1e-3)
            print(f"perf n_spks={n_spks} B={B} T={T} steps={n} chunk={chunk}: {ms:.2f} ms  {fs / 1e6:.3f} M frame-steps/s  "
                  f"{fs * 134.154e6 / 1e12:.1f} TFLOP/s  mem {torch.cuda.memory_allocated() / 2**30:.1f} GiB torch")
            del dec
        except Exception:
            print(f"perf {n_spks} {B} {T}: EXCEPTION")
            traceback.print_exc()


def stage_halo():
    import gpu_util as gu
    cases = [("3x3_64_64_l0", dict(kind=0, B=2, H=80, W=24, Cin0=64, Cin1=0, Cout=64)),
             ("3x3_64_128_odd", dict(kind=0, B=1, H=40, W=22, Cin0=64, Cin1=0, Cout=128)),
             ("3x3_128_128", dict(kind=0, B=2, H=40, W=36, Cin0=128, Cin1=0, Cout=128)),
             ("3x3_cat_256_64", dict(kind=0, B=1, H=40, W=20, Cin0=128, Cin1=128, Cout=64)),
             ("3x3_cat_512_128", dict(kind=0, B=1, H=20, W=12, Cin0=256, Cin1=256, Cout=128))]
    for impl in (2, 3):
        for name, kw in cases:
            try:
                c = gu.conv_case(seed=hash(name) % 1000, **kw)
                out, st = gu.run_conv(c, impl, 1, want_stats=True)
                ref, raw = gu.conv_reference(c, round_bf16=True)
                err = float((out - ref).abs().max())
                sref = gu.gn_stats_reference(raw)
                serr = float(((st - sref).abs() / (sref.abs() + 1.0)).max())
                d = (out - ref).abs()
                bad = d > 0.1
                print(f"halo impl={impl} {name}: max-abs err {err:.3e} nan {int(torch.isnan(out).sum())} stats-err {serr:.3e} "
                      f"bad {int(bad.sum())}/{bad.numel()}")
                if int(bad.sum()):
                    print("     per-col-bad", bad.sum((0, 1, 2))[:24].tolist(), "per-row-bad", bad.sum((0, 1, 3))[:40].tolist())
            except Exception:
                print(f"halo impl={impl} {name}: EXCEPTION")
                traceback.print_exc()


def stage_convdbg():
    """Times one large conv through the test hook (GTTS_CONV_REPS) under the experiment switches GTTS_CONV_DBG."""
    import gpu_util as gu
    cin, cout, H, W, B = [int(v) for v in os.environ.get("GTTS_SHAPE", "64,64,80,1720,8").split(",")]
    kind = int(os.environ.get("GTTS_KIND", "0"))
    if kind != 0:       # 1x1 (optionally per-sample weights + residual + mask: GTTS_KIND=1 GTTS_RES=1), stride-2, ConvT
        res = bool(int(os.environ.get("GTTS_RES", "0")))
        c = gu.conv_case(kind, B, H, W, cin, 0, cout, seed=1, residual=res, mask=res or kind in (2, 3), per_sample=res)
        os.environ["GTTS_CONV_REPS"] = "5"
        for dbg in (0, 1, 2, 3):
            os.environ["GTTS_CONV_DBG"] = str(dbg)
            print(f"impl=1 kind={kind} dbg={dbg}", flush=True)
            gu.run_conv(c, 1, 1)
        os.environ["GTTS_CONV_DBG"] = "0"
        return
    c = gu.conv_case(0, B, H, W, cin, 0, cout, seed=1)
    os.environ["GTTS_CONV_REPS"] = "5"
    if os.environ.get("GTTS_CONV_TIMING"):
        os.environ["GTTS_CONV_DBG"] = os.environ.get("GTTS_CONV_DBG_T", "0")
        gu.run_conv(c, 3, 1, want_stats=True)
        return
    for impl in (1, 3):
        for dbg in [int(v) for v in os.environ.get("GTTS_DBG_LIST", "0,1,2,3").split(",")]:
            os.environ["GTTS_CONV_DBG"] = str(dbg)
            for mc in ((0, 1) if impl == 1 else (0,)):
                os.environ["GTTS_MC"] = str(mc)
                print(f"impl={impl} dbg={dbg} mc={mc}", flush=True)
                gu.run_conv(c, impl, 1, want_stats=True)
    os.environ["GTTS_CONV_DBG"] = "0"


def stage_mbench():
    """tcgen05 issue-path micro-benchmark: what do commits cost relative to MMAs?"""
    import ctypes
    lib = pkg._lib.load()
    a, b = ctypes.c_double(), ctypes.c_double()
    iters = 200
    for N in (64, 128, 256):
        for (nm, nc, we) in [(0, 1, 0), (0, 2, 0), (0, 1, 1), (36, 0, 0), (36, 1, 0), (36, 2, 0), (36, 4, 0), (36, 1, 1), (36, 2, 1),
                             (4, 0, 0), (4, 1, 0), (4, 1, 1), (8, 1, 0), (16, 1, 0), (72, 2, 0)]:
            pkg._lib.check(lib.gtts_test_issue_microbench(N, nm, nc, iters, we, 148, ctypes.byref(a), ctypes.byref(b)), "mb")
            print(f"N={N:3d} mma={nm:2d} commit={nc} wait_each={we}: issue {a.value / iters:8.1f}  total {b.value / iters:8.1f} cycles/round"
                  + (f"  ({b.value / iters / nm:6.1f}/mma)" if nm else ""), flush=True)


def stage_align():
    """Alignment stage at BASELINE C2 (64 x 200 x 1000): log-prior + MAS + durations/mu_y on the device vs the CPU restatement."""
    import time
    from oracle import align_oracle, mas_oracle
    al = importlib.import_module("grad-tts_b200.model.align")
    ma = importlib.import_module("grad-tts_b200.model.monotonic_align")
    B, tx, ty = 64, 200, 1000
    gen = torch.Generator().manual_seed(1234)
    mu_x, y = torch.randn(B, 80, tx, generator=gen), torch.randn(B, 80, ty, generator=gen)
    xl = torch.randint(100, 201, (B,), generator=gen); yl = torch.randint(600, 1001, (B,), generator=gen); xl[0], yl[0] = tx, ty
    x_mask = (torch.arange(tx)[None] < xl[:, None]).float().unsqueeze(1)
    y_mask = (torch.arange(ty)[None] < yl[:, None]).float().unsqueeze(1)
    mask = (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1)
    d = [t.to(DEV) for t in (mu_x, y, x_mask, mask)]

    def dev_chain():
        lp = al.log_prior(d[0], d[1])
        attn = ma.maximum_path(lp, d[3], check=False)
        return lp, attn, al.logw_from_path(attn, d[2]), al.mu_y_from_path(attn, d[0])

    for _ in range(3):
        out = dev_chain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = dev_chain()
    e1.record(); torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 10
    t0 = time.perf_counter()
    lp = align_oracle.log_prior(mu_x, y, 80)
    attn = mas_oracle.maximum_path(lp, mask)
    lw, my = align_oracle.logw_from_path(attn, x_mask), align_oracle.mu_y_from_path(attn, mu_x)
    t_cpu = (time.perf_counter() - t0) * 1e3
    same = bool(torch.equal(out[1].cpu(), attn))
    print(f"align C2 (64x200x1000): device {t_dev:.3f} ms (log_prior + MAS + logw + mu_y), CPU restatement {t_cpu:.1f} ms "
          f"({torch.get_num_threads()} threads), path identical: {same}, "
          f"log_prior max err {float((out[0].cpu() - lp).abs().max()):.2e}", flush=True)


def stage_profile_vjp():
    """per-op times of the forward + backward (VJP) plan, both precisions"""
    import ctypes, json
    train = os.environ.get("GTTS_PROFILE_TRAIN") is not None      # the training plan (parameter gradients) at train.py's shape
    for prec, flags in ((("bf16", 12),) if train else (("bf16", 4), ("fp32", 5))):
        dec, _ = _decoder(1, 0, prec)
        B, T = (16, 172) if train else (16, 400)
        buf = ctypes.create_string_buffer(1 << 18)
        h = dec.estimator._get_handle()
        rc = pkg._lib.load().gtts_decoder_profile_step(h, B, T, flags, 3, buf, len(buf), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        pkg._lib.check(rc, "profile")
        rep = json.loads(buf.value.decode())
        tot = sum(o["ms"] for o in rep["ops"])
        print(f"--- VJP profile {prec} B={B} T={T}: total {tot:.3f} ms over {len(rep['ops'])} ops")
        agg = {}
        for o in rep["ops"]:
            import re
            k = re.sub(r"(_\d+)*(_h\d+)?$", "", o["name"])
            agg[k] = agg.get(k, 0.0) + o["ms"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            print(f"   {k:28s} {v * 1e3:9.1f} us  {100 * v / tot:5.1f}%")
        for o in rep["ops"]:
            if o["name"].startswith("bwd_"):
                tf = o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] > 0 else 0
                print(f"      {o['name']:32s} {o['ms'] * 1e3:9.1f} us  {tf:8.1f} TFLOP/s")


def stage_profile():
    import ctypes, json
    for (n_spks, B, T) in [(1, 16, 1720), (1, 1, 400)]:
        dec, _ = _decoder(n_spks, 0, "bf16")
        dec.estimator.max_chunk = B
        hm = int(os.environ.get("GTTS_HALO", "2"))
        pkg._lib.check(pkg._lib.load().gtts_decoder_set_option(dec.estimator._get_handle(), b"halo_mode", hm), "opt")
        pkg._lib.check(pkg._lib.load().gtts_decoder_set_option(dec.estimator._get_handle(), b"fused_attn",
                                                               int(os.environ.get("GTTS_FUSED_ATTN", "1"))), "opt")
        z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=7, ragged=False)
        dec(z.to(DEV), mask.to(DEV), mu.to(DEV), 2)
        torch.cuda.synchronize()
        buf = ctypes.create_string_buffer(1 << 17)
        h = dec.estimator._get_handle()
        rc = pkg._lib.load().gtts_decoder_profile_step(h, B, T, 0, 5, buf, len(buf),
                                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        pkg._lib.check(rc, "profile")
        rep = json.loads(buf.value.decode())
        tot = sum(o["ms"] for o in rep["ops"])
        print(f"--- profile B={B} T={T}: total {tot:.3f} ms over {len(rep['ops'])} launches")
        for o in rep["ops"]:
            tf = o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] > 0 else 0
            gb = o["bytes"] / (o["ms"] * 1e-3) / 1e9 if o["ms"] > 0 else 0
            print(f"   {o['name']:28s} {o['ms'] * 1e3:9.1f} us  {tf:8.1f} TFLOP/s  {gb:8.0f} GB/s(alg)")


def stage_vocoder():
    """HiFi-GAN generator: error against the goldens in both precisions, per-launch profile and whole-forward time"""
    import glob
    import numpy as np
    synth = pkg.synth
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "voc_*.npz"))):
        g = np.load(path)
        cfg = synth.VOCODER_CONFIGS[str(g["cfg"])]
        gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
        gen.load_state_dict(synth.make_vocoder_state_dict(cfg, seed=int(g["wseed"])))
        gen = gen.to(DEV).eval()
        ref = torch.from_numpy(g["y"])
        for prec in ("fp32", "bf16"):
            gen.precision = prec
            y = gen(torch.from_numpy(g["mel"]).to(DEV)).cpu()
            print(f"{os.path.basename(path)} {prec}: max-abs {float((y - ref).abs().max()):.3e} rel-rms "
                  f"{float(((y - ref).pow(2).mean() / ref.pow(2).mean()).sqrt()):.3e} finite {bool(torch.isfinite(y).all())}", flush=True)
    cfg = synth.VOCODER_CONFIGS["v1"]
    gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
    gen.load_state_dict(synth.make_vocoder_state_dict(cfg, seed=1))
    gen = gen.to(DEV).eval()
    for (B, T) in [(16, 1720), (1, 400)]:
        gen.max_chunk = B
        mel = synth.make_mel(B, T, seed=2).to(DEV)
        for _ in range(2):
            y = gen(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = gen(mel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"--- vocoder forward B={B} T={T}: {ms:.3f} ms  ({B * T / ms * 1e3:.0f} frames/s, {gen.launches_last_call()} launches, "
              f"finite {bool(torch.isfinite(y).all())})", flush=True)
        print(gen.profile(B, T), flush=True)


def stage_encoder():
    """text encoder: forward time at a few shapes, with the latency-shaped conv kernel on (default threshold) and off"""
    import importlib
    te = importlib.import_module("grad-tts_b200.model.text_encoder")
    from oracle import text_encoder_oracle
    synth = pkg.synth
    cfg = synth.TEXT_ENCODER_CONFIGS["ref"]
    sd = synth.make_text_encoder_state_dict(cfg, seed=1)
    sd_d = {k: v.to(DEV) for k, v in sd.items()}
    for small_m in ("1024", "0", "100000"):
        os.environ["GTTS_ENC_SMALL_M"] = small_m
        enc = te.TextEncoder(**cfg)
        enc.load_state_dict(sd)
        enc = enc.to(DEV).eval()
        for (B, T) in [(1, 100), (4, 150), (16, 200), (128, 200)]:
            x, lengths, _ = synth.make_text_inputs(cfg, B, T, seed=3, ragged=False)
            xd, ld = x.to(DEV), lengths.to(DEV)
            mu, logw, _ = enc(xd, ld)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                enc(xd, ld)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            with torch.no_grad():
                mu_r, logw_r, _ = text_encoder_oracle.text_encoder_forward(sd, cfg, x, lengths)
            print(f"small_m={small_m:>6s} B={B:3d} T={T:3d}: {ms:7.3f} ms  max-abs mu {float((mu.cpu() - mu_r).abs().max()):.2e} "
                  f"logw {float((logw.cpu() - logw_r).abs().max()):.2e}", flush=True)
        del enc
    for (B, T) in [(1, 100), (128, 200)]:
        x, lengths, _ = synth.make_text_inputs(cfg, B, T, seed=3, ragged=False)
        xd, ld = x.to(DEV), lengths.to(DEV)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    text_encoder_oracle.text_encoder_forward(sd_d, cfg, xd, ld)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    text_encoder_oracle.text_encoder_forward(sd_d, cfg, xd, ld)
                e1.record()
                torch.cuda.synchronize()
            print(f"eager tf32={tf32} B={B} T={T}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)


if __name__ == "__main__":
    st = sys.argv[1]
    print(f"===== stage {st} on {torch.cuda.get_device_name(0)}", flush=True)
    {"profile_vjp": stage_profile_vjp, "mas": stage_mas, "conv_ffma": lambda: (stage_conv(0, 0), stage_conv(0, 1)), "conv_tc": lambda: stage_conv(1, 1),
     "dec_fp32": lambda: stage_dec("fp32"), "dec_bf16_ffma": lambda: stage_dec("bf16", 0),
     "dec_bf16_tc": lambda: stage_dec("bf16", 1), "perf": stage_perf, "profile": stage_profile, "halo": stage_halo, "convdbg": stage_convdbg, "mbench": stage_mbench, "align": stage_align, "vocoder": stage_vocoder, "encoder": stage_encoder}[st]()
    torch.cuda.synchronize()
    print(f"===== stage {st} done", flush=True)
