"""weight gradients of the bf16 training plan: tcgen05 kernel vs mma.sync kernel, parameter by parameter"""
import importlib
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("grad-tts_b200")
B, T = int(os.environ.get("WG_B", 3)), int(os.environ.get("WG_T", 172))
sd = pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05)
x0, mask, mu, _, _ = pkg.synth.make_inputs(B, T, 1, seed=11, ragged=True)
a = [v.cuda() for v in (x0, mask, mu)]
tt = torch.rand(B, generator=torch.Generator().manual_seed(1)).clamp(1e-5, 1 - 1e-5).cuda()
z = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(2)).cuda()
grads = {}
for mode in (0, 1):
    dec = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd)
    dec = dec.cuda().train()
    dec.precision = "bf16"
    dec.estimator.set_option("wgrad_tc", mode)
    loss, _ = dec.loss_t(a[0], a[1], a[2], tt, noise=z)
    loss.backward()
    torch.cuda.synchronize()
    grads[mode] = {n: p.grad.detach().float().cpu() for n, p in dec.named_parameters()}
    print("mode", mode, "loss", float(loss), flush=True)
bad = 0
for n in grads[0]:
    g0, g1 = grads[0][n], grads[1][n]
    if g0.dim() != 4:
        continue
    rel = float((g1 - g0).pow(2).mean().sqrt() / g0.pow(2).mean().sqrt().clamp_min(1e-30))
    flag = "" if rel < 2e-3 else "   <-- MISMATCH"
    bad += rel >= 2e-3
    print(f"{n:50s} {tuple(g0.shape)!s:20s} rel {rel:.2e}{flag}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
