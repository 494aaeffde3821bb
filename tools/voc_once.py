"""two vocoder forwards at 16 x 1720 frames (ncu target)"""
import importlib
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("grad-tts_b200")
cfg = pkg.synth.VOCODER_CONFIGS["v1"]
gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
gen.load_state_dict(pkg.synth.make_vocoder_state_dict(cfg, seed=1))
gen = gen.cuda().eval()
gen.set_option("use_graph", 0)          # eager launches: ncu sees every kernel of the second forward in order
gen.max_chunk = 16
mel = pkg.synth.make_mel(16, 1720, seed=2).cuda()
for _ in range(2):
    y = gen(mel)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(y).all()))
