import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SAMPLES = "/root/reference/samples"
# full captures staged for the GPU box by __graft_entry__.build() (git-ignored, travels with gpurun)
STAGED_SAMPLES = os.path.join(GOLDEN, "_samples")

FILES = {  # name: (fs, fc)
    "headset3": (2e6, 2476e6),
    "headset1": (8e6, 2476.5e6),
    "keyboard1": (8e6, 2476.5e6),
    "headset2": (4e6, 2476e6),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def kats():
    return json.load(open(os.path.join(GOLDEN, "ref_kats.json")))


def load_excerpt(name, mode):
    """-> dict(iq complex64, fs, fc, nslots, stdout, energy, noise, nsym, bits (list per window), heavy)."""
    z = np.load(os.path.join(GOLDEN, "%s_%s.npz" % (name, mode)), allow_pickle=False)
    zc = np.load(os.path.join(GOLDEN, "%s_chained.npz" % name), allow_pickle=False)
    xi = zc["iq_i16"].astype(np.float32)
    d = dict(iq=(xi[0::2] + 1j * xi[1::2]).astype(np.complex64), fs=float(z["fs"]), fc=float(z["fc"]),
             nslots=int(z["nslots"]), stdout=str(z["stdout"]), energy=z["energy"], noise=z["noise"],
             nsym=z["nsym"], bits_packed=z["bits_packed"])
    d["heavy"] = {k: z[k] for k in z.files if k.split("_")[0] in ("ddc", "soft", "demod", "mu")}
    return d


def golden_bits(ex, call, chi):
    n = int(ex["nsym"][call, chi])
    return np.unpackbits(ex["bits_packed"][call, chi])[:n]


def full_capture(name):
    """Full bundled capture if reachable (container: /root/reference; GPU box: staged copy)."""
    p = os.path.join(REF_SAMPLES, name + ".cfile")
    if os.path.exists(p):
        return np.fromfile(p, dtype=np.complex64)
    p = os.path.join(STAGED_SAMPLES, name + ".i16.npy")
    if os.path.exists(p):
        xi = np.load(p).astype(np.float32)
        return (xi[0::2] + 1j * xi[1::2]).astype(np.complex64)
    return None
