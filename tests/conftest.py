import ast
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CUDA library is git-ignored: build it in-tree (nvcc cross-compiles sm_100a without a GPU, ~2.5 min) if a
    fresh checkout does not have it yet, so that neither the CPU nor the GPU suite silently skips the product."""
    import dmel_b200
    dmel_b200.build(force=False)


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    rec = {k: z[k] for k in z.files if k != "meta"}
    rec["meta"] = ast.literal_eval(str(z["meta"]))
    rec["name"] = name
    return rec


def oracle_kwargs(meta):
    """meta (scripts/gen_golden.py CASES) -> keyword arguments of the oracle / dmel op."""
    kw = dict(hop=meta["hop"], optimized=meta["opt"], norm=meta["norm"])
    if meta["kind"] == "mel":
        kw.update(n_mels=meta["M"], sample_rate=meta["sr"])
    return kw


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
