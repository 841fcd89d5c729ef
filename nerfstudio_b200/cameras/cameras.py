"""Perspective cameras -> rays, one kernel (mirror of the perspective/OpenCV-distortion path of
nerfstudio/cameras/cameras.py:321-929 and model_components/ray_generators.py:29-56).

Exotic camera models (fisheye624, equirectangular, ODS/VR180, orthographic) are outside the BASELINE configs."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from .. import functional as F
from .rays import RayBundle


@dataclass
class Cameras:
    camera_to_worlds: Tensor            # [C,3,4]
    fx: Tensor                          # [C] or [C,1]
    fy: Tensor
    cx: Tensor
    cy: Tensor
    width: Optional[Tensor] = None
    height: Optional[Tensor] = None
    distortion_params: Optional[Tensor] = None  # [C,6] (k1,k2,k3,k4,p1,p2)

    def __post_init__(self):
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        self.fx, self.fy, self.cx, self.cy = f(self.fx), f(self.fy), f(self.cx), f(self.cy)
        self.camera_to_worlds = self.camera_to_worlds.float().reshape(-1, 3, 4)

    @property
    def device(self):
        return self.camera_to_worlds.device

    def to(self, device) -> "Cameras":
        mv = lambda t: None if t is None else t.to(device)
        return Cameras(mv(self.camera_to_worlds), mv(self.fx), mv(self.fy), mv(self.cx), mv(self.cy), mv(self.width),
                       mv(self.height), mv(self.distortion_params))

    def intrinsics(self) -> Tensor:
        return torch.stack([self.fx, self.fy, self.cx, self.cy], dim=-1).contiguous()

    def generate_rays_from_indices(self, ray_indices: Tensor, disable_distortion: bool = False) -> RayBundle:
        """ray_indices int64 [R,3] = (camera, row, col); pixel centres (+0.5) as get_image_coords does."""
        dist = None if disable_distortion else self.distortion_params
        if dist is not None and not bool((dist != 0).any()):
            dist = None
        r = F.generate_rays(self.camera_to_worlds, self.intrinsics(), dist, ray_indices)
        return RayBundle(origins=r["origins"], directions=r["directions"], pixel_area=r["pixel_area"],
                         camera_indices=r["camera_indices"], metadata={"directions_norm": r["directions_norm"]})


class RayGenerator(nn.Module):
    """(camera,row,col) -> RayBundle (mirror of model_components/ray_generators.py:29-56)."""

    def __init__(self, cameras: Cameras) -> None:
        super().__init__()
        self.cameras = cameras

    def forward(self, ray_indices: Tensor) -> RayBundle:
        return self.cameras.generate_rays_from_indices(ray_indices)
