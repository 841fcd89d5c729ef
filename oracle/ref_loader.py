"""Import the UNMODIFIED reference package (TEST INFRASTRUCTURE; only tests/ and bench.py's reference arm use this).

Search order: $NERFSTUDIO_REFERENCE, /root/reference (build container), oracle/_ref (GPU box; made by oracle/make_ref.py).
The reference imports three packages this image does not have; they are stubbed because nothing on the hot path calls
into them: `viser` (OrientedBox.from_params only), `matplotlib` (utils/colormaps, imported by models/*), `torchmetrics`
(PSNR/SSIM/LPIPS objects built in populate_modules, used by get_image_metrics_and_images only).  `nerfacc` /
`tinycudann` are NOT stubbed here: `nerfstudio_b200.integration.install()` provides them, or the caller registers the
shim first (the reference's torch path needs `nerfacc` importable at module import time, ray_samplers.py:24)."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    for cand in (os.environ.get("NERFSTUDIO_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "nerfstudio", "models")):
            return cand
    return None


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def to(self, *a, **k):
        return self


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    m = types.ModuleType(name)

    def _getattr(k):  # any public attribute is a class that accepts anything
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything

    m.__getattr__ = _getattr
    sys.modules[name] = m


def load(with_nerfacc_shim: bool = True):
    """Make `import nerfstudio` resolve to the reference.  Returns its root directory or None if unavailable."""
    root = reference_root()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("viser", "viser.transforms", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "torchmetrics",
                 "torchmetrics.functional", "torchmetrics.image", "torchmetrics.image.lpip"):
        try:
            __import__(name)
        except Exception:  # noqa: BLE001
            _stub(name)
    if with_nerfacc_shim and "nerfacc" not in sys.modules:
        try:
            __import__("nerfacc")
        except Exception:  # noqa: BLE001
            import importlib

            sys.modules["nerfacc"] = importlib.import_module("nerfstudio_b200.shims.nerfacc")
    return root
