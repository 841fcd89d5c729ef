"""Drop the B200 core in behind an importable nerfstudio — models run unmodified.

    import nerfstudio_b200.integration as b200
    b200.install()            # before nerfstudio models are constructed
    # ns-train nerfacto ... / instant-ngp ... now build their fields, samplers and renderers from this package

What `install()` does (each step is the binding a nerfstudio maintainer would otherwise write by hand):

1. registers `nerfstudio_b200.shims.nerfacc` / `.tinycudann` as the modules `nerfacc` / `tinycudann` when the real
   packages are not importable — nerfstudio imports `nerfacc` at module import time
   (model_components/ray_samplers.py:24, renderers.py:34) and gates its tcnn path on `import tinycudann`
   (utils/external.py:38-58);
2. replaces the hot-path classes inside the reference's own modules, and re-binds the names the model files
   imported from them (`from nerfstudio.fields.nerfacto_field import NerfactoField` binds at import time):
   HashEncoding / SHEncoding / NeRFEncoding, MLP / MLPWithHashEncoding, trunc_exp, NerfactoField,
   HashMLPDensityField, the samplers, RGB / accumulation / depth renderers, interlevel / distortion losses;
3. patches `RaySamples.get_weights` (cameras/rays.py:129-152) to the warp-scan kernel;
4. patches `CameraOptimizer.apply_to_raybundle` (cameras/camera_optimizers.py:148-153) to the fused SO3xR3 kernel
   (the module, its `pose_adjustment` parameter and every other method stay the reference's);
5. patches `Cameras.generate_rays` (cameras/cameras.py:321-503) to the ray-generation kernel for perspective cameras on
   the GPU — same signature (camera_indices, coords, camera_opt_to_camera, distortion_params_delta, keep_shape,
   disable_distortion, aabb_box), the reference's own code for every other camera model.

`install_fused_trainer(trainer)` additionally routes a reference `Trainer`'s `train_iteration` through the captured
step (`pipeline.FusedTrainStep`) when its model is a nerfacto model.

Everything replaced keeps the reference's constructor signature, attributes and state_dict keys, so configs and
checkpoints are untouched.  `uninstall()` restores the originals.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

_ORIGINALS: List[Tuple[object, str, object]] = []

# reference module -> {attribute: (our module, our attribute)}
_REPLACEMENTS: Dict[str, Dict[str, Tuple[str, str]]] = {
    "nerfstudio.field_components.encodings": {
        "HashEncoding": ("nerfstudio_b200.field_components.encodings", "HashEncoding"),
        "SHEncoding": ("nerfstudio_b200.field_components.encodings", "SHEncoding"),
        "NeRFEncoding": ("nerfstudio_b200.field_components.encodings", "NeRFEncoding"),
    },
    "nerfstudio.field_components.mlp": {
        "MLP": ("nerfstudio_b200.field_components.mlp", "MLP"),
        "MLPWithHashEncoding": ("nerfstudio_b200.field_components.mlp", "MLPWithHashEncoding"),
    },
    "nerfstudio.field_components.activations": {
        "trunc_exp": ("nerfstudio_b200.field_components.activations", "trunc_exp"),
    },
    "nerfstudio.fields.nerfacto_field": {"NerfactoField": ("nerfstudio_b200.fields.nerfacto_field", "NerfactoField")},
    "nerfstudio.fields.density_fields": {
        "HashMLPDensityField": ("nerfstudio_b200.fields.density_fields", "HashMLPDensityField")},
    "nerfstudio.model_components.ray_samplers": {
        k: ("nerfstudio_b200.model_components.ray_samplers", k)
        for k in ("UniformSampler", "LinearDisparitySampler", "SqrtSampler", "LogSampler",
                  "UniformLinDispPiecewiseSampler", "PDFSampler", "ProposalNetworkSampler", "VolumetricSampler")},
    "nerfstudio.model_components.renderers": {
        k: ("nerfstudio_b200.model_components.renderers", k)
        for k in ("RGBRenderer", "AccumulationRenderer", "DepthRenderer")},
    "nerfstudio.model_components.losses": {
        k: ("nerfstudio_b200.model_components.losses", k) for k in ("interlevel_loss", "distortion_loss")},
}
# modules that did `from <reference module> import <name>` and therefore hold their own binding
_REBIND_IN = ("nerfstudio.models.nerfacto", "nerfstudio.models.instant_ngp", "nerfstudio.models.vanilla_nerf",
              "nerfstudio.models.depth_nerfacto", "nerfstudio.fields.nerfacto_field", "nerfstudio.fields.density_fields",
              "nerfstudio.fields.vanilla_nerf_field", "nerfstudio.field_components.mlp")


def _importable(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:  # noqa: BLE001
        return False


def _set(obj, attr: str, value) -> None:
    _ORIGINALS.append((obj, attr, getattr(obj, attr, None)))
    setattr(obj, attr, value)


def install(shim_third_party: bool = True, patch_get_weights: bool = True) -> List[str]:
    """Returns the list of `module.attribute` names that were replaced."""
    from . import lib

    lib.load()  # fail loudly if the CUDA library is not built: there is nothing to fall back to
    done: List[str] = []
    if shim_third_party:
        for name, ours in (("nerfacc", "nerfstudio_b200.shims.nerfacc"), ("tinycudann", "nerfstudio_b200.shims.tinycudann")):
            if name not in sys.modules and not _importable(name):
                sys.modules[name] = importlib.import_module(ours)
                done.append(f"sys.modules[{name!r}]")
        # gsplat (models/splatfacto.py:26-31 imports gsplat.rendering.rasterization and gsplat.strategy at import time)
        if "gsplat" not in sys.modules and not _importable("gsplat"):
            import types

            shim = importlib.import_module("nerfstudio_b200.shims.gsplat")
            pkg = types.ModuleType("gsplat")
            pkg.__path__ = []
            rendering = types.ModuleType("gsplat.rendering")
            rendering.rasterization = shim.rasterization
            strategy = types.ModuleType("gsplat.strategy")
            strategy.DefaultStrategy, strategy.MCMCStrategy = shim.DefaultStrategy, shim.MCMCStrategy
            pkg.rendering, pkg.strategy = rendering, strategy
            for nm, mod in (("gsplat", pkg), ("gsplat.rendering", rendering), ("gsplat.strategy", strategy)):
                sys.modules[nm] = mod
                done.append(f"sys.modules[{nm!r}]")
    for ref_mod_name, attrs in _REPLACEMENTS.items():
        ref_mod = importlib.import_module(ref_mod_name)
        for attr, (our_mod_name, our_attr) in attrs.items():
            ours = getattr(importlib.import_module(our_mod_name), our_attr)
            original = getattr(ref_mod, attr)
            _set(ref_mod, attr, ours)
            done.append(f"{ref_mod_name}.{attr}")
            for holder_name in _REBIND_IN:
                holder = sys.modules.get(holder_name)
                if holder is None and _importable(holder_name):
                    holder = sys.modules.get(holder_name)
                if holder is not None and holder is not ref_mod and getattr(holder, attr, None) is original:
                    _set(holder, attr, ours)
                    done.append(f"{holder_name}.{attr}")
    # our fields key their output dictionaries with FieldHeadNames; inside nerfstudio that must be the REFERENCE's enum
    # object (models index field_outputs[FieldHeadNames.DENSITY], models/nerfacto.py:308).  When this package was imported
    # before nerfstudio became importable it defined its own enum: re-bind the name everywhere it is held.
    ref_heads = importlib.import_module("nerfstudio.field_components.field_heads")
    for mod_name in ("nerfstudio_b200.field_components.field_heads", "nerfstudio_b200.fields.base_field",
                     "nerfstudio_b200.fields.nerfacto_field", "nerfstudio_b200.fields.vanilla_nerf_field",
                     "nerfstudio_b200.nerfacto", "nerfstudio_b200.instant_ngp"):
        mod = importlib.import_module(mod_name)
        if getattr(mod, "FieldHeadNames", None) is not ref_heads.FieldHeadNames:
            _set(mod, "FieldHeadNames", ref_heads.FieldHeadNames)
            done.append(f"{mod_name}.FieldHeadNames")
    if patch_get_weights:
        from .cameras.rays import RaySamples as OurSamples

        rays = importlib.import_module("nerfstudio.cameras.rays")
        _set(rays.RaySamples, "get_weights", OurSamples.get_weights)
        done.append("nerfstudio.cameras.rays.RaySamples.get_weights")
    # 4. the camera optimiser's per-ray pose correction (cameras/camera_optimizers.py:148-153): one fused kernel for
    #    mode SO3xR3, the reference's own code for anything else
    cam_opt = importlib.import_module("nerfstudio.cameras.camera_optimizers")
    from .cameras.camera_optimizers import fused_apply_to_raybundle

    original_apply = cam_opt.CameraOptimizer.apply_to_raybundle

    def apply_to_raybundle(self, raybundle) -> None:
        if self.config.mode == "SO3xR3" and raybundle.origins.is_cuda:
            fused_apply_to_raybundle(self, raybundle)
        else:
            original_apply(self, raybundle)

    _set(cam_opt.CameraOptimizer, "apply_to_raybundle", apply_to_raybundle)
    done.append("nerfstudio.cameras.camera_optimizers.CameraOptimizer.apply_to_raybundle")
    # 5. Cameras.generate_rays (cameras/cameras.py:321-503): perspective cameras on the GPU -> one ray-generation kernel
    #    behind the reference's own signature (training batches with coords, whole images with keep_shape); every other
    #    camera model, multi-dimensional camera batches, oriented boxes and CPU cameras keep the reference's code
    cams_mod = importlib.import_module("nerfstudio.cameras.cameras")
    rays_mod = importlib.import_module("nerfstudio.cameras.rays")
    from .cameras.cameras import fused_generate_rays

    original_generate = cams_mod.Cameras.generate_rays

    def generate_rays(self, camera_indices, coords=None, camera_opt_to_camera=None, distortion_params_delta=None,
                      keep_shape=None, disable_distortion=False, aabb_box=None, obb_box=None):
        cams = self if self.shape else self.reshape((1,))
        out = None
        if cams.camera_to_worlds.is_cuda and len(cams.shape) == 1 and obb_box is None:
            out = fused_generate_rays(cams, camera_indices, coords, camera_opt_to_camera, distortion_params_delta, keep_shape,
                                      disable_distortion, aabb_box, None, bundle_cls=rays_mod.RayBundle)
        if out is None:
            return original_generate(self, camera_indices, coords, camera_opt_to_camera, distortion_params_delta,
                                     keep_shape, disable_distortion, aabb_box, obb_box)
        return out

    _set(cams_mod.Cameras, "generate_rays", generate_rays)
    done.append("nerfstudio.cameras.cameras.Cameras.generate_rays")
    return done


def uninstall() -> None:
    while _ORIGINALS:
        obj, attr, value = _ORIGINALS.pop()
        if value is None:
            try:
                delattr(obj, attr)
            except AttributeError:
                pass
        else:
            setattr(obj, attr, value)
    for name in ("nerfacc", "tinycudann"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("nerfstudio_b200.shims"):
            del sys.modules[name]
    rend = sys.modules.get("gsplat.rendering")
    if rend is not None and getattr(getattr(rend, "rasterization", None), "__module__", "").startswith("nerfstudio_b200"):
        for name in ("gsplat.strategy", "gsplat.rendering", "gsplat"):
            sys.modules.pop(name, None)


def install_fused_trainer(trainer, **engine_kwargs):
    """Bind the captured step to a reference `Trainer` (engine/trainer.py): `trainer.train_iteration(step)` keeps its
    signature and return tuple `(loss, loss_dict, metrics_dict)` but runs forward + losses + backward + optimiser as one
    CUDA-graph replay; rays still come from `trainer.pipeline.datamanager.next_train(step)`.  The optimiser settings are
    read from the trainer's own config (Adam lr / eps of the "fields" group, configs/method_configs.py:106-119).
    Returns the `FusedTrainStep`; `trainer._b200_original_train_iteration` keeps the reference method."""
    from .pipeline import FusedTrainStep

    pipeline = trainer.pipeline
    model = getattr(pipeline, "model", None) or pipeline._model
    dm = pipeline.datamanager
    n_rays = int(getattr(getattr(dm, "config", None), "train_num_rays_per_batch", 4096))
    lr, eps = 1e-2, 1e-15
    try:
        oc = trainer.config.optimizers["fields"]["optimizer"]
        lr, eps = float(oc.lr), float(oc.eps)
    except Exception:  # noqa: BLE001
        pass
    fused = FusedTrainStep(model, dm, n_rays, lr=lr, eps=eps, **engine_kwargs)
    trainer._b200_original_train_iteration = trainer.train_iteration
    trainer.train_iteration = fused.train_iteration
    pipeline.get_train_loss_dict = fused.get_train_loss_dict
    # multi-GPU sharded update: the parameter all-gather of the last step may still be in flight when the trainer turns to
    # evaluation or checkpointing (engine/trainer.py:532-580, 435-480) — make those wait for it
    for name in ("eval_iteration", "save_checkpoint"):
        orig = getattr(trainer, name, None)
        if callable(orig):
            def flushed(*a, _orig=orig, **k):
                fused.flush()
                return _orig(*a, **k)

            setattr(trainer, name, flushed)
    return fused
