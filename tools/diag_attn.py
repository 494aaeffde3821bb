import importlib, os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
pkg = importlib.import_module("grad-tts_b200")
DEV = "cuda:0"
sd = pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05)
def est(env, x, mask, mu, t):
    for k in ("GTTS_ATTN_TC", "GTTS_ATTN_TC_ONLY"): os.environ.pop(k, None)
    os.environ.update(env)
    dec = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000).to(DEV)
    dec.load_state_dict(sd)
    pkg._lib.check(pkg._lib.load().gtts_decoder_set_option(dec.estimator._get_handle(), b"use_graph", 0), "opt")
    y = dec.estimator(x, mask, mu, t, None)
    torch.cuda.synchronize()
    return y
B, T = 2, 1720
z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, 1, seed=7, ragged=False)
z, mask, mu = z.to(DEV), mask.to(DEV), mu.to(DEV)
t = torch.full((B,), 0.5, device=DEV)
n0, n1, n2 = 80 * T, 40 * (T // 2), 20 * (T // 4)
for name, env in [("all tc", {"GTTS_ATTN_CHECK": "1"}), ("none", {"GTTS_ATTN_TC": "0"})]:
    a = est(env, z * 100.0, mask, mu, t)
    print(f"{name}: finite {bool(torch.isfinite(a).all())} absmax {float(a[torch.isfinite(a)].abs().max()) if torch.isfinite(a).any() else float('nan'):.4g}", flush=True)
