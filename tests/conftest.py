import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_golden(name):
    """Load tests/golden/<name>.npz; keys 'x__float16'/'x__bfloat16' are uint16 bit patterns."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        if "__" in k:
            base, dt = k.split("__")
            t = torch.from_numpy(a.view(np.int16).copy()).view(getattr(torch, dt))
            out[base] = t
        elif a.shape == ():
            out[k] = a.item()
        else:
            out[k] = torch.from_numpy(a.copy())
    return out


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def cos_sim(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
