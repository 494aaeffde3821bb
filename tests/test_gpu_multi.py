"""GPU, >= 2 devices: batch sharding over ranks gives bitwise the single-GPU result."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r"""
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
pkg = importlib.import_module("grad-tts_b200")
rank, world = int(sys.argv[3]), int(sys.argv[4])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=world, device_id=dev)
n_spks, B, T, n = 247, 5, 40, 3
sd = pkg.synth.make_decoder_state_dict(n_spks, seed=3, g=0.05)
dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
dec.load_state_dict(sd); dec = dec.to(dev)
z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=4)
full = [t.to(dev) for t in (z, mask, mu, spk)]
ref = dec(full[0], full[1], full[2], n, False, full[3])                      # whole batch on this GPU
out = pkg.dist.sharded_call(lambda z_, m_, mu_, s_: dec(z_, m_, mu_, n, False, s_), full, B)
assert out.shape == ref.shape and torch.equal(out, ref), "sharded result differs from the single-GPU result"
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_decoder_bitwise_equal_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
