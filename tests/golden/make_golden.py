"""Generate the golden fixtures in tests/golden/ by running the REAL reference on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
It copies /root/reference/model to a temp dir, builds the reference's Cython extension with the
reference's own setup.py (README.md:31), imports `model`, loads seeded synthetic weights
(grad-tts_b200/synth.py) with strict=True into the reference modules, runs them in fp32 on CPU
(torch threads fixed to 8) and stores inputs + outputs as small .npz files.
"""
import hashlib
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("grad-tts_b200.synth")


def import_reference():
    tmp = tempfile.mkdtemp(prefix="gtts_ref_")
    shutil.copytree("/root/reference/model", os.path.join(tmp, "model"))
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"],
                          cwd=os.path.join(tmp, "model", "monotonic_align"),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sys.path.insert(0, tmp)
    import model  # noqa: F401
    return tmp


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def decoder_cases():
    # name, n_spks, B, T, n_steps (0 = single estimator call), weight seed, input seed
    return [
        ("est_spk1_b2_t48", 1, 2, 48, 0, 0, 1),
        ("est_spk247_b2_t40", 247, 2, 40, 0, 3, 4),
        ("dec_spk1_b1_t64_n10", 1, 1, 64, 10, 0, 5),
        ("dec_spk1_b3_t56_n4", 1, 3, 56, 4, 7, 8),
        ("dec_spk247_b2_t40_n3", 247, 2, 40, 3, 3, 6),
    ]


def main():
    torch.set_num_threads(8)
    import_reference()
    from model.diffusion import Diffusion
    from model.monotonic_align import maximum_path

    for name, n_spks, B, T, n_steps, wseed, iseed in decoder_cases():
        sd = synth.make_decoder_state_dict(n_spks, seed=wseed, g=0.05)
        dec = Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000).eval()
        dec.load_state_dict(sd, strict=True)
        z, mask, mu, spk, lengths = synth.make_inputs(B, T, n_spks, seed=iseed, ragged=True)
        out = {}
        with torch.no_grad():
            if n_steps == 0:
                gen = torch.Generator().manual_seed(iseed + 100)
                t = torch.rand(B, generator=gen).clamp(1e-5, 1 - 1e-5)
                y = dec.estimator(z * mask, mask, mu, t, spk)
                out["t"] = t.numpy()
            else:
                y = dec(z, mask, mu, n_steps, False, spk)
                y_stoc = dec(z, mask, mu, n_steps, True, spk)
                assert torch.equal(y, y_stoc), "reference fork: stoc flag must be a no-op"
        out.update(z=z.numpy(), mask=mask.numpy(), mu=mu.numpy(), y=y.numpy(),
                   n_spks=np.int64(n_spks), n_steps=np.int64(n_steps), wseed=np.int64(wseed),
                   sd_sha256=np.array(sd_digest(sd)))
        if spk is not None:
            out["spk"] = spk.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "y absmax", float(y.abs().max()))

    # MAS: value/mask regenerated from the seed in the tests; only the int8 path is stored.
    mas_cases = [("mas_b4_20x50", 4, 20, 50, 11, True), ("mas_b3_33x33", 3, 33, 33, 12, False),
                 ("mas_b5_1x9", 5, 1, 9, 13, True), ("mas_b2_64x257", 2, 64, 257, 14, True),
                 ("mas_b64_200x1000", 64, 200, 1000, 1234, True)]
    for name, B, tx, ty, seed, ragged in mas_cases:
        value, mask, txs, tys = synth.make_mas_inputs(B, tx, ty, seed=seed, ragged=ragged)
        path = maximum_path(value, mask)
        assert path.dtype == value.dtype
        p8 = path.numpy().astype(np.int8)
        # store the path compactly: for every (b, y) the row index of the single 1 (or -1)
        idx = np.where(p8.sum(1) > 0, p8.argmax(1), -1).astype(np.int16)      # (B, t_y)
        assert (p8.sum(1) <= 1).all()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), idx=idx, tx=txs.numpy(), ty=tys.numpy(),
                            seed=np.int64(seed), ragged=np.bool_(ragged),
                            shape=np.array([B, tx, ty]), value_sum=np.float64(value.double().sum()))
        print(name, "ones", int(p8.sum()))


if __name__ == "__main__":
    main()
