"""GPU: maximum_path through the C ABI, bit-exact against the golden fixtures and the C oracle."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mas_oracle

pytestmark = pytest.mark.gpu
MAS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "mas_*.npz")))
DEV = "cuda:0"


@pytest.mark.parametrize("name", MAS)
def test_mas_golden(name, pkg, synth):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, tx, ty = (int(v) for v in g["shape"])
    value, mask, _, _ = synth.make_mas_inputs(B, tx, ty, seed=int(g["seed"]), ragged=bool(g["ragged"]))
    path = pkg.maximum_path(value.to(DEV), mask.to(DEV)).cpu()
    assert path.dtype == value.dtype and path.shape == value.shape
    idx = g["idx"].astype(np.int64)
    ref = np.zeros((B, tx, ty), dtype=np.float32)
    b, y = np.nonzero(idx >= 0)
    ref[b, idx[b, y], y] = 1
    assert np.array_equal(path.numpy(), ref)


@pytest.mark.parametrize("B,tx,ty", [(1, 1, 1), (3, 1, 40), (2, 7, 7), (4, 31, 33), (2, 33, 64), (3, 100, 257),
                                     (2, 300, 1203), (1, 1024, 1100), (1, 1500, 1600),
                                     (2, 224, 300), (2, 225, 300), (2, 416, 500), (2, 417, 500), (1, 200, 4400), (1, 200, 4500)])   # edges of the warp-serial variants
def test_mas_random_vs_oracle(B, tx, ty, pkg, synth):
    value, mask, _, _ = synth.make_mas_inputs(B, tx, ty, seed=B * 1000 + tx + ty)
    ref = mas_oracle.maximum_path(value, mask)
    got = pkg.maximum_path(value.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(ref, got)


def test_mas_large_negative_and_ties(pkg):
    g = torch.Generator().manual_seed(7)
    value = torch.randint(-3, 3, (4, 24, 60), generator=g).float()          # many exact ties
    value[1] = -1e9
    value[2, :, ::3] = -3e8
    mask = torch.ones_like(value)
    ref = mas_oracle.maximum_path(value, mask)
    got = pkg.maximum_path(value.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(ref, got)


def test_mas_general_binary_mask_multiplies_values(pkg):
    """The wrapper computes value*mask (reference __init__.py:13) also for non-prefix masks."""
    g = torch.Generator().manual_seed(11)
    value = 5 * torch.randn(2, 16, 40, generator=g) - 40
    mask = torch.ones_like(value)
    mask[:, :, 30:] = 0
    mask[:, 12:, :] = 0
    mask[0, 3, 5] = 0            # a hole: value there becomes 0 (> all negative neighbours)
    ref = mas_oracle.maximum_path(value, mask)
    got = pkg.maximum_path(value.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(ref, got)


def test_mas_dtype_and_errors(pkg, synth):
    value, mask, _, _ = synth.make_mas_inputs(2, 10, 20, seed=1)
    got = pkg.maximum_path(value.double().to(DEV), mask.double().to(DEV))
    assert got.dtype == torch.float64
    with pytest.raises(RuntimeError):
        pkg.maximum_path(value, mask)                        # CPU tensors: no fallback
    bad = torch.ones(1, 8, 4)                                # t_x > t_y
    with pytest.raises(RuntimeError):
        pkg.maximum_path(bad.to(DEV), bad.to(DEV))


def test_mas_c_entry_point_int32(pkg, synth):
    from importlib import import_module
    ma = import_module("grad-tts_b200.model.monotonic_align")
    value, mask, tx, ty = synth.make_mas_inputs(5, 40, 90, seed=21)
    v = (value * mask).contiguous()
    paths = torch.full(v.shape, 7, dtype=torch.int32, device=DEV)
    ma.maximum_path_c(paths, v.to(DEV), tx.int().to(DEV), ty.int().to(DEV))
    ref = mas_oracle.maximum_path(value, mask).int()
    assert torch.equal(paths.cpu(), ref)


def test_mas_host_entry_point(pkg, synth):
    lib = pkg._lib.load()
    value, mask, _, _ = synth.make_mas_inputs(3, 20, 50, seed=31)
    out = torch.empty_like(value)
    status = ctypes.c_int32(-1)
    rc = lib.gtts_mas_maximum_path_host(value.data_ptr(), mask.data_ptr(), out.data_ptr(), 3, 20, 50,
                                        ctypes.addressof(status), 0)
    pkg._lib.check(rc, "maximum_path_host")
    assert status.value == 0
    assert torch.equal(out, mas_oracle.maximum_path(value, mask))


def test_mas_full_size_properties(pkg, synth):
    """BASELINE config 2 (B=64, 200x1000) + size-independent properties of a monotonic path."""
    value, mask, tx, ty = synth.make_mas_inputs(64, 200, 1000, seed=1234)
    path = pkg.maximum_path(value.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(path, mas_oracle.maximum_path(value, mask))
    assert torch.equal(path.sum(1).sum(1).long(), ty)                    # one cell per valid frame
    assert float((path * (1 - mask)).abs().sum()) == 0.0                 # nothing outside the mask
    rows = path.argmax(1)                                                # (B, t_y) row index per frame
    for b in range(64):
        r = rows[b, : int(ty[b])]
        d = r[1:] - r[:-1]
        assert int(r[0]) == 0 and int(r[-1]) == int(tx[b]) - 1 and bool(((d == 0) | (d == 1)).all())
