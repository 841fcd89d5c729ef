"""Scene contraction (mirror of nerfstudio/field_components/spatial_distortions.py:30-100).

The L-inf contraction nerfacto/instant-ngp use is normally fused into the position kernel
(`functional.positions_to_unit_cube`); calling the module directly gives the same values.  Other norms are not on
the BASELINE path and are evaluated with torch ops."""
from typing import Optional, Union

import torch
from torch import nn


class SpatialDistortion(nn.Module):
    def forward(self, positions):
        raise NotImplementedError


class SceneContraction(SpatialDistortion):
    def __init__(self, order: Optional[Union[float, int]] = None) -> None:
        super().__init__()
        self.order = order

    @property
    def is_linf(self) -> bool:
        return self.order is not None and float(self.order) == float("inf")

    def forward(self, positions):
        mag = torch.linalg.norm(positions, ord=self.order, dim=-1)[..., None]
        return torch.where(mag < 1, positions, (2 - (1 / mag)) * (positions / mag))
