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

    @property
    def shape(self):
        return (self.camera_to_worlds.shape[0],)

    def generate_rays(self, camera_indices, coords: Optional[Tensor] = None, camera_opt_to_camera: Optional[Tensor] = None,
                      distortion_params_delta: Optional[Tensor] = None, keep_shape: Optional[bool] = None,
                      disable_distortion: bool = False, aabb_box=None, obb_box=None) -> RayBundle:
        """The reference's signature (cameras/cameras.py:321-331): camera_indices int | [*num_rays, 1], coords
        [*num_rays, 2] = (y, x) pixel-centre coordinates or None for whole images ([H, W, *num_rays] rays)."""
        if obb_box is not None:
            raise NotImplementedError("oriented boxes are outside the BASELINE hot path")
        out = fused_generate_rays(self, camera_indices, coords, camera_opt_to_camera, distortion_params_delta, keep_shape,
                                  disable_distortion, aabb_box, None)
        if out is None:
            raise NotImplementedError("only perspective cameras (1-D batch, CUDA) are implemented")
        return out

    def generate_rays_from_indices(self, ray_indices: Tensor, disable_distortion: bool = False) -> RayBundle:
        """ray_indices int64 [R,3] = (camera, row, col); pixel centres (+0.5) as get_image_coords does."""
        dist = None if disable_distortion else self.distortion_params
        if dist is not None and not bool((dist != 0).any()):
            dist = None
        r = F.generate_rays(self.camera_to_worlds, self.intrinsics(), dist, ray_indices)
        return RayBundle(origins=r["origins"], directions=r["directions"], pixel_area=r["pixel_area"],
                         camera_indices=r["camera_indices"], metadata={"directions_norm": r["directions_norm"]})


def fused_generate_rays(cameras, camera_indices, coords=None, camera_opt_to_camera=None, distortion_params_delta=None,
                        keep_shape=None, disable_distortion: bool = False, aabb_box=None, obb_box=None,
                        bundle_cls=RayBundle):
    """`Cameras.generate_rays` (cameras/cameras.py:321-503) for perspective, non-jagged, 1-D camera batches on ONE
    kernel: the argument standardisation of the reference (its four camera_indices x coords cases) is reproduced on the
    host, the per-ray arithmetic of `_generate_rays_from_coords` (:505-929) runs in raygen.cu.  `cameras` is this
    package's `Cameras` or, duck-typed, the reference's (install() patches the reference class to call this).  Returns
    None when the configuration is outside the kernel's scope (caller falls back to / raises as it sees fit)."""
    c2w = cameras.camera_to_worlds
    if c2w.dim() != 3 or not c2w.is_cuda or obb_box is not None:
        return None
    ctype = getattr(cameras, "camera_type", None)
    if ctype is not None and bool((ctype != 1).any()):  # CameraType.PERSPECTIVE.value == 1
        return None
    if getattr(cameras, "is_jagged", False) and coords is None:
        return None
    dev = c2w.device
    if isinstance(camera_indices, int):
        camera_indices = torch.tensor([camera_indices], device=dev)
    camera_indices = camera_indices.to(dev)
    if camera_indices.shape[-1] != 1:
        return None  # multi-dimensional camera batches: the reference's own code
    n_cam = c2w.shape[0]
    flat = lambda t: t.reshape(n_cam, -1)[:, 0] if t.numel() != n_cam else t.reshape(-1)
    intr = torch.stack([flat(cameras.fx), flat(cameras.fy), flat(cameras.cx), flat(cameras.cy)], -1).float().contiguous()
    dist = None if disable_distortion else getattr(cameras, "distortion_params", None)
    if disable_distortion:
        distortion_params_delta = None
    if coords is not None:
        shape = tuple(coords.shape[:-1])
        ci = camera_indices.broadcast_to(shape + (1,)).reshape(-1)
        r = F.generate_rays_coords(c2w, intr, dist, ci, coords.to(dev), cam_opt=camera_opt_to_camera,
                                   dist_delta=distortion_params_delta)
    else:
        lead = tuple(camera_indices.shape[:-1])
        first = int(camera_indices.reshape(-1)[0])
        H, W = int(cameras.height.reshape(-1)[first]), int(cameras.width.reshape(-1)[first])
        if keep_shape is True or not getattr(cameras, "is_jagged", False):
            hh, ww = cameras.height.reshape(-1)[camera_indices.reshape(-1)], cameras.width.reshape(-1)[camera_indices.reshape(-1)]
            assert bool((hh == H).all()) and bool((ww == W).all()), "Can only keep shape if all cameras have the same height and width"
        shape = (H, W) + lead
        r = F.generate_rays_coords(c2w, intr, dist, camera_indices.reshape(-1), None, H, W, cam_opt=camera_opt_to_camera,
                                   dist_delta=distortion_params_delta)
    rs = lambda t: t.reshape(shape + (t.shape[-1],))
    meta = {"directions_norm": rs(r["directions_norm"])}
    cam_meta = getattr(cameras, "metadata", None)
    cams_out = rs(r["camera_indices"])
    if cam_meta:
        for k, v in cam_meta.items():
            meta[k] = v[cams_out[..., 0]]
    times = getattr(cameras, "times", None)
    bundle = bundle_cls(origins=rs(r["origins"]), directions=rs(r["directions"]), pixel_area=rs(r["pixel_area"]),
                        camera_indices=cams_out, metadata=meta)
    if times is not None:
        bundle.times = times[cams_out[..., 0], 0][..., None] if times.dim() > 1 else times[cams_out[..., 0]][..., None]
    if keep_shape is False:
        bundle = bundle.flatten() if hasattr(bundle, "flatten") else bundle
    if aabb_box is not None:
        o, d = bundle.origins.reshape(-1, 3), bundle.directions.reshape(-1, 3)
        box = aabb_box.aabb.flatten().to(dev) if hasattr(aabb_box, "aabb") else torch.as_tensor(aabb_box, device=dev).flatten()
        tmin, tmax, _ = F.ray_aabb_intersect(o, d, box.reshape(1, 6), 0.0, 1e10, 1e10)  # utils/math.py:138-175 semantics
        lead_shape = tuple(bundle.origins.shape[:-1])
        bundle.nears, bundle.fars = tmin.reshape(lead_shape + (1,)), tmax.reshape(lead_shape + (1,))
    return bundle


class RayGenerator(nn.Module):
    """(camera,row,col) -> RayBundle (mirror of model_components/ray_generators.py:29-56)."""

    def __init__(self, cameras: Cameras) -> None:
        super().__init__()
        self.cameras = cameras

    def forward(self, ray_indices: Tensor) -> RayBundle:
        return self.cameras.generate_rays_from_indices(ray_indices)
