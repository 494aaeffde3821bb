"""Timing of one Block conv with the GroupNorm-apply variants at a real chunk-16 shape (prints us/launch from the C hook)."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
pkg = importlib.import_module("grad-tts_b200")
lib = pkg._lib.load()
dev = torch.device("cuda:0")
B, H, W, C = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 40, 860, 128))]
variant = sys.argv[5] if len(sys.argv) > 5 else "async"
use_res = len(sys.argv) > 6 and sys.argv[6] == "res"
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(dev)
b = torch.randn(C, generator=g).to(dev); ga = torch.ones(C).to(dev); be = torch.zeros(C).to(dev)
mask = torch.ones(B, W).to(dev)
tb = None if use_res else torch.randn(1, C, generator=g).to(dev)
res = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev) if use_res else None
out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev); st = torch.zeros(B, 8, 2, device=dev)
p = lambda t: t.data_ptr() if t is not None else None
reps = 20 if variant == "tmem" else -20
rc = lib.gtts_test_conv_apply(B, H, W, C, 0, C, p(x), None, p(w), p(b), p(ga), p(be), p(tb), 0, p(res), p(mask), p(out), p(st), reps,
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
pkg._lib.check(rc, "apply timing")
torch.cuda.synchronize()
