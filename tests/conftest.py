import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("grad-tts_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("grad-tts_b200.synth")


def load_decoder_case(name):
    """Inputs and reference output of a decoder fixture.  Full fixtures (est_*/dec_*) store their inputs; the BASELINE-shape
    fixtures (c1*/c3*) store only the reference output and regenerate the inputs from the seed (digest-checked)."""
    import hashlib
    import numpy as np
    import torch
    synth_ = importlib.import_module("grad-tts_b200.synth")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n_spks, n_steps = int(g["n_spks"]), int(g["n_steps"])
    case = dict(n_spks=n_spks, n_steps=n_steps, wseed=int(g["wseed"]), y=torch.from_numpy(g["y"]), t=None)
    if "z" in g:
        case.update(z=torch.from_numpy(g["z"]), mask=torch.from_numpy(g["mask"]), mu=torch.from_numpy(g["mu"]),
                    spk=torch.from_numpy(g["spk"]) if "spk" in g else None)
        if "t" in g:
            case["t"] = torch.from_numpy(g["t"])
    else:
        B, T = (int(v) for v in g["shape"])
        z, mask, mu, spk, _ = synth_.make_inputs(B, T, n_spks, seed=int(g["iseed"]), ragged=bool(g["ragged"]))
        h = hashlib.sha256()
        for t in (z, mask, mu, spk):
            if t is not None:
                h.update(t.contiguous().numpy().tobytes())
        assert h.hexdigest() == str(g["in_sha256"]), f"{name}: regenerated inputs differ from the ones the fixture was made with"
        case.update(z=z, mask=mask, mu=mu, spk=spk)
    return case


def decoder_case_names(prefixes):
    import glob
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if os.path.basename(p).startswith(tuple(prefixes)))
