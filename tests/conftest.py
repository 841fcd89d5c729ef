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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rel=1e-4, name=""):
    """|a-b|_inf <= rel * |b|_inf  (the north-star's 1e-4 rel fp32 bar), NaN-pattern equal."""
    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    b = b.detach().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(torch.isnan(a), torch.isnan(b)), f"{name}: NaN pattern differs"
    a, b = torch.nan_to_num(a.double()), torch.nan_to_num(b.double())
    scale = float(b.abs().max().clamp_min(1e-30))
    err = float((a - b).abs().max())
    assert err <= rel * scale, f"{name}: max abs err {err:.3e} > {rel:g} * {scale:.3e}"
