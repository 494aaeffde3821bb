"""one text-encoder forward at B=1, T=100 and at B=128, T=200 (ncu target)"""
import importlib
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("grad-tts_b200")
te = importlib.import_module("grad-tts_b200.model.text_encoder")
cfg = pkg.synth.TEXT_ENCODER_CONFIGS["ref"]
enc = te.TextEncoder(**cfg)
enc.load_state_dict(pkg.synth.make_text_encoder_state_dict(cfg, 1))
enc = enc.cuda().eval()
shapes = [(1, 100), (128, 200)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split("x"))]
for B, T in shapes:
    x, l, _ = pkg.synth.make_text_inputs(cfg, B, T, seed=3, ragged=False)
    for _ in range(2):
        enc(x.cuda(), l.cuda())
    torch.cuda.synchronize()
print("ok")
