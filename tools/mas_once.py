"""a few MAS calls at the C2 shape (ncu target)"""
import importlib
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("grad-tts_b200")
value, mask, _, _ = pkg.synth.make_mas_inputs(64, 200, 1000, seed=1234)
v, m = value.cuda(), mask.cuda()
for _ in range(3):
    pkg.maximum_path(v, m, check=False)
torch.cuda.synchronize()
print("ok")
