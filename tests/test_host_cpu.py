"""CPU: host-side logic, C-ABI surface, state_dict compatibility, multi-process sharding (gloo)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "gradtts_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gtts_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(pkg._lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert pkg._lib.load().gtts_version() == 100


def test_no_gpu_means_loud_failure(pkg, synth):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = pkg._lib.load()
    assert lib.gtts_sm100_device_count() == 0
    h = ctypes.c_void_p()
    rc = lib.gtts_decoder_create(ctypes.byref(h), 1, 80, 64, 0.05, 20.0, 1000.0, 0)
    assert rc != 0 and len(pkg._lib.last_error()) > 0
    dec = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000)
    z, mask, mu, _, _ = synth.make_inputs(1, 40, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(z, mask, mu, 2)
    value, m, _, _ = synth.make_mas_inputs(1, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.maximum_path(value, m)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grad-tts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in src and "import oracle" not in src and "oracle." not in src, f


@pytest.mark.parametrize("n_spks", [1, 247, -1])
def test_state_dict_keys_match_reference_layout(pkg, synth, n_spks):
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    keys = [(k, tuple(v.shape)) for k, v in dec.state_dict().items()]
    assert keys == synth.decoder_param_shapes(n_spks)
    sd = synth.make_decoder_state_dict(n_spks, seed=1)
    res = dec.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert dec.nparams == sum(v.numel() for v in sd.values())
    assert float(dec.estimator.downs._modules["0"]._modules["2"].fn.g.detach()) == pytest.approx(0.05)


def test_default_init_matches_reference_defaults(pkg):
    torch.manual_seed(0)
    dec = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000)
    sd = dec.state_dict()
    assert float(sd["estimator.mid_attn.fn.g"]) == 0.0                      # Rezero g = 0 (diffusion.py:43)
    assert torch.all(sd["estimator.final_block.block.1.weight"] == 1)
    assert torch.all(sd["estimator.final_block.block.1.bias"] == 0)
    w = sd["estimator.downs.1.0.block2.block.0.weight"]
    assert float(w.abs().max()) <= 1.0 / (128 * 9) ** 0.5 + 1e-6


def test_utils_match_reference_semantics(pkg):
    u = importlib.import_module("grad-tts_b200.model.utils")
    assert [u.fix_len_compatibility(n) for n in (1, 4, 5, 82, 400)] == [4, 4, 8, 84, 400]
    m = u.sequence_mask(torch.tensor([3, 1, 0]), 4)
    assert m.tolist() == [[True, True, True, False], [True, False, False, False], [False] * 4]
    dur = torch.tensor([[2.0, 1.0, 3.0], [1.0, 1.0, 0.0]])
    mask = torch.ones(2, 3, 6)
    mask[1, 2:, :] = 0
    mask[1, :, 2:] = 0
    path = u.generate_path(dur, mask)
    assert path[0].tolist() == [[1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 1, 1]]
    assert path[1].tolist() == [[1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]]


def test_shard_bounds(pkg):
    d = pkg.dist
    assert d.shard_counts(100, 8) == [13, 13, 13, 13, 12, 12, 12, 12]          # BASELINE config 4
    assert d.shard_counts(128, 8) == [16] * 8
    for n, w in [(100, 8), (7, 3), (5, 5), (128, 4)]:
        b = [d.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


_WORKER = r"""
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
pkg = importlib.import_module("grad-tts_b200")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n = 7
g = torch.Generator().manual_seed(0)
z = torch.randn(n, 80, 8, generator=g); mu = torch.randn(n, 80, 8, generator=g)
fake = lambda z_, mu_, spk_: (z_ * 2 + mu_) if spk_ is None else None          # stands in for the per-sample decoder
out = pkg.dist.sharded_call(fake, [z, mu, None], n)
assert out.shape == z.shape and torch.equal(out, z * 2 + mu), "gathered result differs from the 1-rank result"
lo, hi = pkg.dist.shard_bounds(n, 2, dist.get_rank())
assert (hi - lo) == (4 if dist.get_rank() == 0 else 3)
# even split: shards land straight in the output tensor
out8 = pkg.dist.sharded_call(fake, [z[:6], mu[:6], None], 6)
assert torch.equal(out8, z[:6] * 2 + mu[:6])
# fewer items than ranks: refused on EVERY rank before any collective (a rank-local raise would hang the others)
try:
    pkg.dist.sharded_call(fake, [z[:1], mu[:1], None], 1)
    raise SystemExit("expected a RuntimeError")
except RuntimeError as e:
    assert "cannot shard" in str(e)
dist.barrier()
dist.destroy_process_group()
print("ok")
"""


def test_sharded_call_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU leg the driver runs beside ours) prints one JSON line with the agreed keys.  C1 workload
    (1 x 400 frames) so the whole run is a few seconds; no GPU is touched."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "decoder_mel_frames_per_sec" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "oracle/decoder_oracle.py" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("C1")


def test_torch_library_ops_are_registered(pkg):
    """The hot path is exposed as torch.library custom ops over the C ABI: every op exists in the dispatcher, propagates
    shapes under FakeTensorMode without touching a device, and has NO CPU kernel (CPU tensors fail in the dispatcher)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    ops = importlib.import_module("grad-tts_b200.ops")
    for name in ops.OPS:
        assert hasattr(torch.ops.gradtts_b200, name), name
    with FakeTensorMode():
        z, m = torch.empty(3, 80, 44, device="cuda"), torch.empty(3, 1, 44, device="cuda")
        t = torch.empty(3, device="cuda")
        assert torch.ops.gradtts_b200.reverse_diffusion(0, z, m, z, None, None, 10, 0).shape == z.shape
        assert torch.ops.gradtts_b200.estimator(0, z, m, z, t, None, 0).shape == z.shape
        v = torch.empty(2, 7, 19, device="cuda")
        path, status = torch.ops.gradtts_b200.maximum_path(v, v)
        assert path.shape == v.shape and status.shape == (1,) and status.dtype == torch.int32
        assert torch.ops.gradtts_b200.log_prior(torch.empty(2, 80, 7, device="cuda"), torch.empty(2, 80, 19, device="cuda")).shape == (2, 7, 19)
        logw, mu_y = torch.ops.gradtts_b200.align_outputs(v, torch.empty(2, 80, 7, device="cuda"), torch.empty(2, 7, device="cuda"), True)
        assert logw.shape == (2, 1, 7) and mu_y.shape == (2, 80, 19)
        xt, zm = torch.ops.gradtts_b200.forward_diffusion(z, m, z, t, z, 0.05, 20.0)
        assert xt.shape == z.shape and zm.shape == z.shape
        assert torch.ops.gradtts_b200.score_loss(z, z, m, t, 0.05, 20.0).shape == ()
        tok = torch.empty(2, 9, dtype=torch.long, device="cuda")
        mu, lw, xm = torch.ops.gradtts_b200.text_encoder(0, tok, torch.empty(2, dtype=torch.long, device="cuda"), None, 80)
        assert mu.shape == (2, 80, 9) and lw.shape == (2, 1, 9) and xm.shape == (2, 1, 9)
        assert torch.ops.gradtts_b200.vocoder(0, torch.empty(2, 80, 5, device="cuda"), 256, 0).shape == (2, 1, 1280)
    with pytest.raises(NotImplementedError):
        torch.ops.gradtts_b200.log_prior(torch.zeros(1, 80, 3), torch.zeros(1, 80, 5))
