"""CPU: the C-ABI library builds/loads and exports every symbol include/b200nerf.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from nerfstudio_b200.build import build

    return build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200nerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2n_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in b200nerf.h but not exported"


def test_python_binding_covers_header(libpath):
    from nerfstudio_b200 import lib

    assert set(lib.EXPORTED_SYMBOLS) == set(header_symbols())
    lib.load()
    assert b"sm_100a" in lib.load().b2n_version()


def test_argument_errors_without_gpu(libpath):
    """Argument validation happens before any CUDA call, so it is checkable on a CPU-only box."""
    from nerfstudio_b200 import lib

    L = lib.load()
    assert L.b2n_sh_fwd(None, 4, 4, 0, None, None) == -1
    assert b"null" in L.b2n_last_error()
    assert L.b2n_adam_step(None, None, None, None, 4, 1, 0.1, 0.9, 0.999, 1e-8, 1.0, None) == -1
    with pytest.raises(ValueError):
        lib.call("b2n_pack_info", None, 5, 3, None, None)


def test_no_cpu_fallback_for_cpu_tensors(libpath):
    import torch

    from nerfstudio_b200 import functional as F

    with pytest.raises(RuntimeError, match="no CPU path"):
        F.sh_encode(torch.zeros(4, 3), 4)


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, smoke() and bench.py's CPU-baseline legs may import it.  Static
    check over the package sources (an import anywhere else would put a CPU path behind the product API)."""
    import ast
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerfstudio_b200")
    offenders = []
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            path = os.path.join(dirpath, fn)
            rel = os.path.relpath(path, root)
            tree = ast.parse(open(path).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                if any(n == "oracle" or n.startswith("oracle.") for n in names) and rel != "smoke.py":
                    offenders.append(rel)
    assert offenders == [], offenders
