"""Proposal density field: hash grid (L=5) + 10->16->1 MLP + trunc_exp.

Mirror of nerfstudio/fields/density_fields.py:33-120 — same constructor, same buffers (`aabb`, `max_res`,
`num_levels`, `log2_hashmap_size`) and the same double registration of the hash table (`encoding.hash_table` and
`mlp_base.0.hash_table` are one Parameter), so checkpoints round-trip.  The 1.44 M points per 4096-ray step that
go through this class are the largest point count on the path."""
from __future__ import annotations

from typing import Literal, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import functional as F
from ..field_components.encodings import HashEncoding
from ..field_components.mlp import MLP
from ..cameras.rays import ray_form
from .base_field import Field, is_linf_contraction, unit_cube_points


class HashMLPDensityField(Field):
    aabb: Tensor

    def __init__(self, aabb: Tensor, num_layers: int = 2, hidden_dim: int = 64, spatial_distortion=None,
                 use_linear: bool = False, num_levels: int = 8, max_res: int = 1024, base_res: int = 16,
                 log2_hashmap_size: int = 18, features_per_level: int = 2, average_init_density: float = 1.0,
                 implementation: Literal["tcnn", "torch"] = "tcnn") -> None:
        super().__init__()
        self.register_buffer("aabb", aabb)
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.average_init_density = average_init_density
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.encoding = HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res,
                                     log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level,
                                     implementation=implementation)
        if not use_linear:
            network = MLP(in_dim=self.encoding.get_out_dim(), num_layers=num_layers, layer_width=hidden_dim, out_dim=1,
                          activation=nn.ReLU(), out_activation=None, implementation=implementation)
            self.mlp_base = torch.nn.Sequential(self.encoding, network)
        else:
            self.linear = torch.nn.Linear(self.encoding.get_out_dim(), 1)

    def _fused_ok(self) -> bool:
        if self.use_linear or self.encoding.tcnn_encoding is not None:
            return False
        net = self.mlp_base[1]
        linf = self.spatial_distortion is None or is_linf_contraction(self.spatial_distortion)
        return linf and getattr(net, "_fused", False) and F.density_field_supported(self.encoding.grid, net.spec)

    def get_density(self, ray_samples) -> Tuple[Tensor, None]:
        o, d, iv = ray_form(ray_samples)
        rays_need_grad = torch.is_grad_enabled() and (o.requires_grad or d.requires_grad)
        # the fused kernel has no position gradient: rays behind a trainable camera optimiser take the unfused kernels,
        # whose backward carries d x (hashgrid_bwd, mlp_bwd) to positions_bwd and on to the pose corrections
        if self._fused_ok() and not rays_need_grad:  # one launch: samples -> unit cube -> grid -> 10->16->1 MLP -> trunc_exp
            net = self.mlp_base[1]
            contraction = self.spatial_distortion is not None
            density = F.density_field(self.encoding.grid, net.spec, self.encoding.hash_table,
                                      [l.weight for l in net.layers], [l.bias for l in net.layers], o, d, iv, contraction,
                                      None if contraction else self.aabb.flatten().tolist(), self.average_init_density)
            return density.view(*ray_samples.frustums.shape, 1), None
        x, sel, R, S = unit_cube_points(ray_samples, self.spatial_distortion, self.aabb)
        if not self.use_linear:
            h = self.mlp_base(x)
        else:
            h = self.linear(self.encoding(x))
        density = F.density_activation(h.float().reshape(-1), sel, self.average_init_density)
        return density.view(*ray_samples.frustums.shape, 1), None

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> dict:
        return {}
