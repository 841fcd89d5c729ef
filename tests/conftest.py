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


def assert_grad_close(a, b, name="", rel=1e-4, sparse_switching=False):
    """Gradient comparison.  `sparse_switching=True` is for gradients that are sums of terms which switch on and
    off with the sample positions (hash-table rows picked by floor/ceil of a position, clip(w - w_outer, 0) of the
    interlevel loss): a 1e-5 shift of a sample legitimately changes individual entries by far more than 1e-4, so
    those are compared by direction (cosine) and norm; everything else entry-wise at `rel` of the max-norm."""
    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    b = b.detach().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    if not sparse_switching:
        return assert_close(a, b, rel, name)
    x, y = a.double().flatten(), b.double().flatten()
    cos = float((x @ y) / (x.norm() * y.norm()).clamp_min(1e-300))
    ratio = float(x.norm() / y.norm().clamp_min(1e-300))
    # measured on B200 (scripts/composed_errors.py): cosine >= 0.999999, norm ratio within 5.1e-4, entries up to 1.6e-3
    assert cos > 0.99999 and abs(ratio - 1) < 2e-3, f"{name}: cosine {cos:.7f}, norm ratio {ratio:.6f}"
