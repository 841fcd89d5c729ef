"""Vanilla NeRF field (mirror of nerfstudio/fields/vanilla_nerf_field.py:31-107): frequency encodings (kernel),
8x256 skip MLP and 2x128 head (library GEMMs — wider than the fused kernel), softplus density / sigmoid rgb heads.
This is BASELINE config 1, the reference's CPU-runnable case: a parity row, not a bench line."""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Type

import torch
from torch import Tensor, nn

from ..cameras.rays import ray_form
from ..field_components.encodings import Encoding, Identity
from ..field_components.field_heads import DensityFieldHead, FieldHead, FieldHeadNames, RGBFieldHead
from ..field_components.mlp import MLP
from .base_field import Field


class NeRFField(Field):
    def __init__(self, position_encoding: Encoding = Identity(in_dim=3), direction_encoding: Encoding = Identity(in_dim=3),
                 base_mlp_num_layers: int = 8, base_mlp_layer_width: int = 256, head_mlp_num_layers: int = 2,
                 head_mlp_layer_width: int = 128, skip_connections: Tuple[int] = (4,),
                 field_heads: Optional[Tuple[Type[FieldHead]]] = (RGBFieldHead,), use_integrated_encoding: bool = False,
                 spatial_distortion=None) -> None:
        super().__init__()
        if use_integrated_encoding:
            raise NotImplementedError("mip-NeRF integrated encodings are outside the BASELINE hot path")
        self.position_encoding, self.direction_encoding = position_encoding, direction_encoding
        self.use_integrated_encoding = False
        self.spatial_distortion = spatial_distortion
        self.mlp_base = MLP(in_dim=position_encoding.get_out_dim(), num_layers=base_mlp_num_layers,
                            layer_width=base_mlp_layer_width, skip_connections=skip_connections,
                            out_activation=nn.ReLU())
        self.field_output_density = DensityFieldHead(in_dim=self.mlp_base.get_out_dim())
        if field_heads:
            self.mlp_head = MLP(in_dim=self.mlp_base.get_out_dim() + direction_encoding.get_out_dim(),
                                num_layers=head_mlp_num_layers, layer_width=head_mlp_layer_width,
                                out_activation=nn.ReLU())
        self.field_heads = nn.ModuleList([h() for h in field_heads] if field_heads else [])
        for h in self.field_heads:
            h.set_in_dim(self.mlp_head.get_out_dim())

    def _positions(self, ray_samples) -> Tensor:
        o, d, iv = ray_form(ray_samples)
        return o[:, None, :] + d[:, None, :] * ((iv.starts() + iv.ends()) / 2)[..., None]

    def get_density(self, ray_samples) -> Tuple[Tensor, Tensor]:
        pos = self._positions(ray_samples)
        if self.spatial_distortion is not None:
            pos = self.spatial_distortion(pos)
        base = self.mlp_base(self.position_encoding(pos))
        return self.field_output_density(base), base

    def get_outputs(self, ray_samples, density_embedding: Optional[Tensor] = None) -> Dict:
        o, d, iv = ray_form(ray_samples)
        enc_dir = self.direction_encoding(d)[:, None, :].expand(iv.R, iv.S, -1)
        out = {}
        for head in self.field_heads:
            out[head.field_head_name] = head(self.mlp_head(torch.cat([enc_dir, density_embedding], dim=-1)))
        return out
