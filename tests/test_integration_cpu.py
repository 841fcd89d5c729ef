"""CPU: `integration.install()` against the real reference package when it is importable (build container);
skipped on the GPU box, where /root/reference does not exist.  Structural only — no kernels are launched."""
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("NERFSTUDIO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nerfstudio")), reason="reference not present")


@pytest.fixture()
def installed():
    sys.path.insert(0, REF)
    stub = types.ModuleType("viser")
    stub.transforms = types.ModuleType("viser.transforms")
    sys.modules.setdefault("viser", stub)
    sys.modules.setdefault("viser.transforms", stub.transforms)
    from nerfstudio_b200 import integration

    done = integration.install()
    yield done
    integration.uninstall()
    sys.path.remove(REF)


def test_install_replaces_hot_path_classes(installed):
    import nerfstudio.field_components.encodings as enc
    import nerfstudio.fields.density_fields as dens
    import nerfstudio.fields.nerfacto_field as nf
    import nerfstudio.model_components.ray_samplers as rs

    assert "nerfstudio.fields.nerfacto_field.NerfactoField" in installed
    assert enc.HashEncoding.__module__.startswith("nerfstudio_b200")
    assert nf.NerfactoField.__module__.startswith("nerfstudio_b200")
    assert dens.HashMLPDensityField.__module__.startswith("nerfstudio_b200")
    assert rs.ProposalNetworkSampler.__module__.startswith("nerfstudio_b200")
    assert sys.modules["nerfacc"].__name__ == "nerfstudio_b200.shims.nerfacc"  # the import-time dependency is satisfied
    import tinycudann  # noqa: F401  -> TCNN_EXISTS becomes true for nerfstudio.utils.external


def test_state_dict_layout_identical_to_reference(installed):
    """Same constructor call, same keys and shapes as the reference's own torch-path classes (checkpoints round-trip)."""
    from nerfstudio_b200 import integration

    import nerfstudio.fields.density_fields as dens
    import nerfstudio.fields.nerfacto_field as nf

    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    kw = dict(num_images=7, num_levels=4, max_res=128, log2_hashmap_size=10, implementation="torch")
    ours_f = nf.NerfactoField(aabb, **kw)
    ours_d = dens.HashMLPDensityField(aabb, hidden_dim=16, num_levels=5, max_res=128, log2_hashmap_size=10,
                                      implementation="torch")
    integration.uninstall()
    import importlib

    ref_nf = importlib.reload(importlib.import_module("nerfstudio.fields.nerfacto_field"))
    ref_dens = importlib.reload(importlib.import_module("nerfstudio.fields.density_fields"))
    ref_f = ref_nf.NerfactoField(aabb, **kw)
    ref_d = ref_dens.HashMLPDensityField(aabb, hidden_dim=16, num_levels=5, max_res=128, log2_hashmap_size=10,
                                         implementation="torch")
    for ours, ref in ((ours_f, ref_f), (ours_d, ref_d)):
        a, b = ours.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape, k
        ref.load_state_dict(a)   # our checkpoint loads into the reference ...
        ours.load_state_dict(b)  # ... and the reference's into ours
    assert torch.equal(ours_f.mlp_base.model[0].scalings, ref_f.mlp_base.model[0].scalings)


def test_camera_optimizer_patch_keeps_reference_module(installed):
    """install() swaps only CameraOptimizer.apply_to_raybundle; parameters / state_dict stay the reference's, our mirror
    has the same key, and CPU bundles (no kernel available) still take the reference's own code path."""
    import numpy as np

    from nerfstudio.cameras.camera_optimizers import CameraOptimizer as RefOpt
    from nerfstudio.cameras.camera_optimizers import CameraOptimizerConfig as RefCfg
    from nerfstudio.cameras.rays import RayBundle as RefBundle
    from nerfstudio_b200.cameras.camera_optimizers import CameraOptimizer, CameraOptimizerConfig

    assert "nerfstudio.cameras.camera_optimizers.CameraOptimizer.apply_to_raybundle" in installed
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_opt.npz"))
    C = g["pose"].shape[0]
    ref = RefOpt(RefCfg(mode="SO3xR3"), num_cameras=C, device="cpu")
    ours = CameraOptimizer(CameraOptimizerConfig(mode="SO3xR3"), num_cameras=C, device="cpu")
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys()) == ["pose_adjustment"]
    ours.load_state_dict(ref.state_dict())
    with torch.no_grad():
        ref.pose_adjustment.copy_(torch.from_numpy(g["pose"]))
    rb = RefBundle(origins=torch.from_numpy(g["origins"]), directions=torch.from_numpy(g["directions"]),
                   pixel_area=torch.ones(g["origins"].shape[0], 1), camera_indices=torch.from_numpy(g["cams"]))
    ref.apply_to_raybundle(rb)  # patched method, CPU tensors -> the reference's original code
    assert torch.allclose(rb.directions, torch.from_numpy(g["out_directions"]), atol=1e-6)
    assert torch.allclose(rb.origins, torch.from_numpy(g["out_origins"]), atol=1e-6)
