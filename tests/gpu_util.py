"""Helpers shared by the GPU tests and tools/gpu_diag.py: call the kernel-level C-ABI test hooks."""
import ctypes
import importlib

import torch
import torch.nn.functional as F

pkg = importlib.import_module("grad-tts_b200")
_lib = pkg._lib


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous().float()


def conv_case(kind, B, H, W, Cin0, Cin1, Cout, seed=0, residual=False, mask=False, bias=True, per_sample=False):
    """Random inputs for one conv variant (CPU fp32 tensors, NCHW)."""
    g = torch.Generator().manual_seed(seed)
    Cin = Cin0 + Cin1
    x = torch.randn(B, Cin, H, W, generator=g)
    if kind in (0, 2):
        w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    elif kind == 1:
        rows = B * Cout if per_sample else Cout
        w = torch.randn(rows, Cin, 1, 1, generator=g) / Cin ** 0.5
    else:
        w = torch.randn(Cin, Cout, 4, 4, generator=g) / (Cin * 4) ** 0.5
    b = torch.randn(Cout, generator=g) if bias else None
    Ho, Wo = (H // 2, W // 2) if kind == 2 else ((2 * H, 2 * W) if kind == 3 else (H, W))
    r = torch.randn(B, Cout, Ho, Wo, generator=g) if residual else None
    m = (torch.rand(B, Wo, generator=g) > 0.3).float() if mask else None
    return dict(kind=kind, B=B, H=H, W=W, Cin0=Cin0, Cin1=Cin1, Cout=Cout, x=x, w=w, b=b, r=r, m=m,
                per_sample=per_sample, Ho=Ho, Wo=Wo)


def conv_reference(c, round_bf16):
    """torch CPU reference in fp32 on (optionally bf16-rounded) operands; returns (out NCHW, gn stats or None)."""
    rb = (lambda t: t.to(torch.bfloat16).float()) if round_bf16 else (lambda t: t)
    x, w = rb(c["x"]), rb(c["w"])
    k = c["kind"]
    if k == 0:
        y = F.conv2d(x, w, c["b"], padding=1)
    elif k == 2:
        y = F.conv2d(x, w, c["b"], stride=2, padding=1)
    elif k == 1:
        if c["per_sample"]:
            B, Cout = c["B"], c["Cout"]
            ws = w.view(B, Cout, -1)
            y = torch.einsum("boc,bchw->bohw", ws, x)
            if c["b"] is not None:
                y = y + c["b"][None, :, None, None]
        else:
            y = F.conv2d(x, w, c["b"])
    else:
        y = F.conv_transpose2d(x, w, c["b"], stride=2, padding=1)
    raw = y
    if c["r"] is not None:
        y = y + rb(c["r"])
    if c["m"] is not None:
        y = y * c["m"][:, None, None, :]
    return y, raw


def gn_stats_reference(raw):
    B, C = raw.shape[:2]
    g = raw.double().view(B, 8, -1)
    mean = g.mean(-1)
    var = g.var(-1, unbiased=False)
    return torch.stack([mean, 1.0 / torch.sqrt(var + 1e-5)], -1).float()      # (B, 8, 2)


def run_conv(c, impl, act, want_stats=False, device="cuda:0"):
    """impl 0 = FFMA, 1 = tcgen05; act 0 = fp32, 1 = bf16.  Returns (out NCHW fp32 cpu, stats (B,8,2) or None)."""
    lib = _lib.load()
    dev = torch.device(device)
    dt = torch.bfloat16 if act else torch.float32
    x = c["x"]
    x0 = nhwc(x[:, : c["Cin0"]], dt).to(dev)
    x1 = nhwc(x[:, c["Cin0"]:], dt).to(dev) if c["Cin1"] else None
    w = c["w"].float().contiguous().to(dev)
    b = c["b"].float().contiguous().to(dev) if c["b"] is not None else None
    r = nhwc(c["r"], dt).to(dev) if c["r"] is not None else None
    m = c["m"].float().contiguous().to(dev) if c["m"] is not None else None
    out = torch.full((c["B"], c["Ho"], c["Wo"], c["Cout"]), float("nan"), dtype=dt, device=dev)
    stats = torch.zeros(c["B"], 8, 2, dtype=torch.float32, device=dev) if want_stats else None
    p = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        rc = lib.gtts_test_conv(impl, act, c["kind"], c["B"], c["H"], c["W"], c["Cin0"], c["Cin1"], c["Cout"],
                                p(x0), p(x1), p(w), p(b), p(r), p(m), p(out), p(stats), int(c["per_sample"]),
                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "gtts_test_conv")
    torch.cuda.synchronize(dev)
    return nchw(out.float().cpu()), (stats.cpu() if stats is not None else None)


CONV_CASES = [
    # name, kwargs
    ("3x3_64_64_l0", dict(kind=0, B=2, H=80, W=24, Cin0=64, Cin1=0, Cout=64)),
    ("3x3_64_128_odd", dict(kind=0, B=1, H=40, W=22, Cin0=64, Cin1=0, Cout=128)),
    ("3x3_256_256_l2", dict(kind=0, B=2, H=20, W=10, Cin0=256, Cin1=0, Cout=256)),
    ("3x3_cat_512_128", dict(kind=0, B=1, H=20, W=12, Cin0=256, Cin1=256, Cout=128)),
    ("3x3_128_256_oddtiles", dict(kind=0, B=1, H=20, W=30, Cin0=128, Cin1=0, Cout=256)),
    ("3x3_256_256_big", dict(kind=0, B=3, H=20, W=108, Cin0=256, Cin1=0, Cout=256)),
    ("1x1_kv_64_256", dict(kind=1, B=2, H=80, W=16, Cin0=64, Cin1=0, Cout=256, bias=False)),
    ("1x1_cat_256_64", dict(kind=1, B=1, H=40, W=12, Cin0=128, Cin1=128, Cout=64)),
    ("1x1_persample_res_mask", dict(kind=1, B=3, H=20, W=12, Cin0=128, Cin1=0, Cout=128, residual=True, mask=True,
                                    per_sample=True)),
    ("3x3s2_64", dict(kind=2, B=2, H=80, W=24, Cin0=64, Cin1=0, Cout=64, mask=True)),
    ("3x3s2_128", dict(kind=2, B=1, H=40, W=20, Cin0=128, Cin1=0, Cout=128, mask=True)),
    ("convT_128", dict(kind=3, B=2, H=20, W=10, Cin0=128, Cin1=0, Cout=128, mask=True)),
    ("convT_64", dict(kind=3, B=1, H=40, W=14, Cin0=64, Cin1=0, Cout=64, mask=True)),
]
