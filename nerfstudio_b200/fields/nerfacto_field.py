"""Nerfacto field: hash grid (L=16) + base MLP -> density & 15 geo features; SH(dir) ‖ geo ‖ appearance -> rgb.

Mirror of nerfstudio/fields/nerfacto_field.py:49-310 for the default heads (density + rgb).  Constructor
arguments, buffers and sub-module names (`mlp_base`, `mlp_head`, `embedding_appearance`, `direction_encoding`,
`position_encoding`) match the reference so that state_dicts are interchangeable.  Transient / semantic /
predicted-normal heads are other model families' options (off by default) and are not built.
"""
from __future__ import annotations

from typing import Dict, Literal, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import functional as F
from ..cameras.rays import ray_form, sample_camera_indices
from ..field_components.embedding import Embedding
from ..field_components.encodings import NeRFEncoding, SHEncoding
from ..field_components.field_heads import FieldHeadNames
from ..field_components.mlp import MLP, MLPWithHashEncoding
from .base_field import Field, unit_cube_points


class NerfactoField(Field):
    aabb: Tensor

    def __init__(self, aabb: Tensor, num_images: int, num_layers: int = 2, hidden_dim: int = 64,
                 geo_feat_dim: int = 15, num_levels: int = 16, base_res: int = 16, max_res: int = 2048,
                 log2_hashmap_size: int = 19, num_layers_color: int = 3, num_layers_transient: int = 2,
                 features_per_level: int = 2, hidden_dim_color: int = 64, hidden_dim_transient: int = 64,
                 appearance_embedding_dim: int = 32, transient_embedding_dim: int = 16,
                 use_transient_embedding: bool = False, use_semantics: bool = False, num_semantic_classes: int = 100,
                 pass_semantic_gradients: bool = False, use_pred_normals: bool = False,
                 use_average_appearance_embedding: bool = False, spatial_distortion=None,
                 average_init_density: float = 1.0, implementation: Literal["tcnn", "torch"] = "tcnn") -> None:
        super().__init__()
        if use_transient_embedding or use_semantics or use_pred_normals:
            raise NotImplementedError("transient / semantic / predicted-normal heads are outside the BASELINE hot path")
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.embedding_appearance = Embedding(num_images, appearance_embedding_dim) if appearance_embedding_dim > 0 else None
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.use_transient_embedding, self.use_semantics, self.use_pred_normals = False, False, False
        self.pass_semantic_gradients = pass_semantic_gradients
        self.base_res = base_res
        self.average_init_density = average_init_density
        self.step = 0
        self.direction_encoding = SHEncoding(levels=4, implementation=implementation)
        self.position_encoding = NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=2 - 1,
                                              implementation=implementation)
        self.mlp_base = MLPWithHashEncoding(
            num_levels=num_levels, min_res=base_res, max_res=max_res, log2_hashmap_size=log2_hashmap_size,
            features_per_level=features_per_level, num_layers=num_layers, layer_width=hidden_dim,
            out_dim=1 + geo_feat_dim, activation=nn.ReLU(), out_activation=None, implementation=implementation)
        self.mlp_head = MLP(
            in_dim=self.direction_encoding.get_out_dim() + geo_feat_dim + appearance_embedding_dim,
            num_layers=num_layers_color, layer_width=hidden_dim_color, out_dim=3, activation=nn.ReLU(),
            out_activation=nn.Sigmoid(), implementation=implementation)

    def get_density(self, ray_samples) -> Tuple[Tensor, Tensor]:
        x, sel, R, S = unit_cube_points(ray_samples, self.spatial_distortion, self.aabb)
        assert x.numel() > 0, "positions is empty."
        shape = ray_samples.frustums.shape
        if getattr(self, "_want_position_grad", False) and not x.requires_grad:
            x.requires_grad_(True)  # analytic normals: d(density_pre)/d(x) through mlp_bwd (dx) and hashgrid_bwd (dx)
        h = self.mlp_base(x).float()
        density_pre, base_mlp_out = torch.split(h, [1, self.geo_feat_dim], dim=-1)
        self._sample_locations, self._density_before_activation = x, density_pre
        density = F.density_activation(density_pre.reshape(-1), sel, self.average_init_density)
        return density.view(*shape, 1), base_mlp_out.reshape(*shape, self.geo_feat_dim)

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> Dict:
        assert density_embedding is not None
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        o, d, iv = ray_form(ray_samples)
        R, S = iv.R, iv.S
        shape = ray_samples.frustums.shape
        # SH of the [0,1]-mapped direction, once per RAY (the reference evaluates the identical value S times)
        sh = F.sh_encode(d, self.direction_encoding.levels, remap01=True) if self.direction_encoding.tcnn_encoding is None \
            else self.direction_encoding((d + 1.0) / 2.0).float()
        parts = [sh[:, None, :].expand(R, S, -1).reshape(R * S, -1), density_embedding.reshape(R * S, self.geo_feat_dim)]
        if self.embedding_appearance is not None:
            if self.training:
                cam = sample_camera_indices(ray_samples, R, S)
                emb = self.embedding_appearance(cam)  # [R,E] when the index is per ray, [R*S,E] when per sample
                if emb.shape[0] == R:
                    emb = emb[:, None, :].expand(R, S, -1)
                parts.append(emb.reshape(R * S, -1))
            elif self.use_average_appearance_embedding:
                parts.append(self.embedding_appearance.mean(dim=0)[None, :].expand(R * S, -1))
            else:
                parts.append(torch.zeros(R * S, self.appearance_embedding_dim, device=d.device))
        rgb = self.mlp_head(torch.cat(parts, dim=-1)).float().view(*shape, 3)
        return {FieldHeadNames.RGB: rgb}
