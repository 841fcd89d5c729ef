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


def sphere_scene_rays(n_rays: int, seed: int = 0, radius: float = 0.5, cam_dist: float = 4.0, num_images: int = 100,
                      device="cpu") -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """BASELINE configs[1] stand-in: Blender-Lego camera geometry (cameras on a sphere of radius 4
    looking at the origin, ~40 degree field of view) around an analytic opaque sphere of radius 0.5 whose colour is
    0.5 + 0.5 * normal, black elsewhere.  Deterministic targets, so the occupancy grid converges to the sphere and the
    packed sample count M is reproducible."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n_rays, generator=g) * 2 * math.pi
    v = torch.acos(torch.rand(n_rays, generator=g) * 1.2 - 0.2)  # mostly upper hemisphere, like the Blender scenes
    o = cam_dist * torch.stack([torch.sin(v) * torch.cos(u), torch.sin(v) * torch.sin(u), torch.cos(v)], -1)
    fwd = torch.nn.functional.normalize(-o, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(n_rays, 3)
    right = torch.nn.functional.normalize(torch.cross(fwd, up, dim=-1), dim=-1)
    upv = torch.cross(right, fwd, dim=-1)
    half = math.tan(0.6911112070083618 / 2)
    px = (torch.rand(n_rays, 2, generator=g) * 2 - 1) * half
    d = torch.nn.functional.normalize(fwd + px[:, :1] * right + px[:, 1:] * upv, dim=-1)
    b = (o * d).sum(-1)
    disc = b * b - ((o * o).sum(-1) - radius * radius)
    hit = disc > 0
    t = -b - torch.sqrt(disc.clamp_min(0))
    n = torch.nn.functional.normalize(o + d * t[:, None], dim=-1)
    rgb = torch.where(hit[:, None], 0.5 + 0.5 * n, torch.zeros(n_rays, 3))
    rays = dict(origins=o, directions=d, pixel_area=torch.full((n_rays, 1), 1e-6),
                camera_indices=torch.randint(0, num_images, (n_rays, 1), generator=g))
    return {k: v.to(device) for k, v in rays.items()}, rgb.to(device)
