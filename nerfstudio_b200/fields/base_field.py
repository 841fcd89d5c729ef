"""Field base class (mirror of nerfstudio/fields/base_field.py:40-142): `get_density`, `get_outputs`, `forward`,
`density_fn`.  Helpers shared by the concrete fields turn a RaySamples into the unit-cube points the grid
kernels take, using the fused position kernel (frustum centre -> L-inf contraction -> normalise -> selector)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import functional as F
from ..cameras.rays import Frustums, RaySamples, ray_form
from ..field_components.field_heads import FieldHeadNames
from ..field_components.spatial_distortions import SceneContraction


def get_normalized_directions(directions: Tensor) -> Tensor:
    """Map unit directions to [0,1] for the SH encoding (reference: fields/base_field.py:136-142)."""
    return (directions + 1.0) / 2.0


def unit_cube_points(ray_samples, spatial_distortion, aabb: Tensor) -> Tuple[Tensor, Tensor, int, int]:
    """-> (x [N,3] in the unit cube with out-of-range points zeroed, selector uint8 [N], R, S)."""
    o, d, iv = ray_form(ray_samples)
    if spatial_distortion is None or (isinstance(spatial_distortion, SceneContraction) and spatial_distortion.is_linf) \
            or _is_reference_linf(spatial_distortion):
        box = aabb.flatten().tolist() if spatial_distortion is None else None
        if torch.is_grad_enabled() and (o.requires_grad or d.requires_grad):
            # rays that carry a gradient (behind the camera optimiser's pose correction): positions stay differentiable
            x, sel = F.positions_to_unit_cube_diff(o, d, iv, spatial_distortion is not None, box)
        else:
            x, sel = F.positions_to_unit_cube(o, d, iv, spatial_distortion is not None, box)
        return x, sel, iv.R, iv.S
    # any other distortion module: evaluate it with its own torch ops, then normalise/select in the kernel
    pos = o[:, None, :] + d[:, None, :] * ((iv.starts() + iv.ends()) / 2)[..., None]
    pos = (spatial_distortion(pos) + 2.0) / 4.0
    x, sel = F.positions_to_unit_cube(pos.reshape(-1, 3), None, None, False, [0.0, 0.0, 0.0, 1.0, 1.0, 1.0])
    return x, sel, iv.R, iv.S


def is_linf_contraction(mod) -> bool:
    return (isinstance(mod, SceneContraction) and mod.is_linf) or _is_reference_linf(mod)


def _is_reference_linf(mod) -> bool:
    """nerfstudio's own SceneContraction(order=inf) instance (when our fields run inside unmodified models)."""
    return type(mod).__name__ == "SceneContraction" and getattr(mod, "order", None) is not None \
        and float(mod.order) == float("inf")


class Field(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self._sample_locations = None
        self._density_before_activation = None
        self._want_position_grad = False

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
        """Density at raw positions [..., 3] -> [..., 1] (occupancy grid / proposal sampler hook)."""
        del times
        flat = positions.reshape(-1, 3)
        zeros = torch.zeros_like(flat[:, :1])
        rs = RaySamples(frustums=Frustums(origins=flat, directions=torch.ones_like(flat), starts=zeros, ends=zeros,
                                          pixel_area=torch.ones_like(zeros)))
        density, _ = self.get_density(rs)
        return density.view(*positions.shape[:-1], 1)

    def get_density(self, ray_samples) -> Tuple[Tensor, Optional[Tensor]]:
        raise NotImplementedError

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> Dict:
        raise NotImplementedError

    def get_normals(self) -> Tensor:
        """-normalize(d density_before_activation / d sample_locations) (fields/base_field.py:80-102).  The position
        gradient runs through the kernels' own backward: `mlp_bwd` (d input) and `hashgrid_bwd` (d x)."""
        assert self._sample_locations is not None, "Sample locations must be set before calling get_normals."
        assert self._density_before_activation is not None, "Density must be set before calling get_normals."
        assert self._sample_locations.shape[:-1] == self._density_before_activation.shape[:-1], (
            "Sample locations and density must have the same shape besides the last dimension.")
        normals = torch.autograd.grad(self._density_before_activation, self._sample_locations,
                                      grad_outputs=torch.ones_like(self._density_before_activation), retain_graph=True)[0]
        return -torch.nn.functional.normalize(normals, dim=-1)

    def forward(self, ray_samples, compute_normals: bool = False) -> Dict:
        # sample locations only become a differentiable leaf when normals are asked for: the reference marks them
        # unconditionally (nerfacto_field.py:215-217) and pays for an unused position gradient in every training step
        self._want_position_grad = bool(compute_normals)
        if compute_normals:
            with torch.enable_grad():
                density, density_embedding = self.get_density(ray_samples)
        else:
            density, density_embedding = self.get_density(ray_samples)
        outputs = self.get_outputs(ray_samples, density_embedding=density_embedding)
        outputs[FieldHeadNames.DENSITY] = density
        if compute_normals:
            with torch.enable_grad():
                normals = self.get_normals()
            outputs[FieldHeadNames.NORMALS] = normals.view(*density.shape[:-1], 3)
        self._want_position_grad = False
        return outputs
