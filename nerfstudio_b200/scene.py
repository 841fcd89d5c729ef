"""Synthetic ray batches of the BASELINE shapes: origins on a radius-1 ring looking
inward with N(0,0.3) jitter, unit directions, random camera indices and random target colours."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .cameras.rays import RayBundle


def synthetic_rays(n_rays: int, num_images: int = 200, seed: int = 0, device="cpu") -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """-> ({origins, directions, pixel_area, camera_indices}, target rgb [R,3]) as plain tensors on `device`."""
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(n_rays, generator=g) * 2 * math.pi
    o = torch.stack([torch.cos(ang), torch.sin(ang), torch.zeros(n_rays)], -1) + 0.3 * torch.randn(n_rays, 3, generator=g)
    target = 0.3 * torch.randn(n_rays, 3, generator=g)
    d = torch.nn.functional.normalize(target - o, dim=-1)
    rays = dict(origins=o, directions=d, pixel_area=torch.full((n_rays, 1), 1e-6),
                camera_indices=torch.randint(0, num_images, (n_rays, 1), generator=g))
    rgb = torch.rand(n_rays, 3, generator=g)
    return {k: v.to(device) for k, v in rays.items()}, rgb.to(device)


def bundle_from(rays: Dict[str, torch.Tensor]) -> RayBundle:
    return RayBundle(origins=rays["origins"], directions=rays["directions"], pixel_area=rays["pixel_area"],
                     camera_indices=rays["camera_indices"])
