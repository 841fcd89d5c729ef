"""Device-side training-ray pipeline.

The reference's `ParallelDataManager` (data/datamanagers/parallel_datamanager.py:168-246) runs pixel sampling and ray
generation in CPU worker processes and ships every batch to the GPU; at B200 step times (~1.3 ms per 4096-ray step) four
DataLoader workers cannot keep up.  `DeviceRayPipeline` keeps the image cache as uint8 in HBM and produces each batch —
pixel sampling (`PixelSampler.sample_method`, data/pixel_samplers.py:137-174), colour lookup + float conversion
(`collate_image_dataset_batch` :265-318, `get_image_float32`), ray generation (`RayGenerator`,
model_components/ray_generators.py:41-56) — with ONE kernel launch, optionally straight into the captured step's input
buffers.  It offers `DataManager.next_train(step) -> (RayBundle, batch)`, so `pipeline.FusedTrainStep` (or a reference
pipeline) can use it in place of the reference's data manager.

Random draws are the reference's `torch.rand((num_rays, 3))` (never seeded per op upstream); tests pass them in.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from ..cameras.cameras import Cameras
from ..cameras.rays import RayBundle
from ..lib import call, ptr, stream

NULL = C.c_void_p(0)


class DeviceRayPipeline:
    def __init__(self, cameras: Cameras, images_uint8: Tensor, num_rays_per_batch: int = 4096,
                 image_idx: Optional[Tensor] = None) -> None:
        if images_uint8.dtype != torch.uint8 or images_uint8.dim() != 4 or images_uint8.shape[-1] < 3:
            raise ValueError("images must be uint8 [num_images, height, width, channels >= 3]")
        if not images_uint8.is_cuda:
            raise RuntimeError("the image cache must live on the GPU (there is no CPU path)")
        self.cameras = cameras
        self.images = images_uint8.contiguous()
        self.R = int(num_rays_per_batch)
        self.image_idx = None if image_idx is None else image_idx.to(self.images.device, torch.int64).contiguous()
        dist = cameras.distortion_params
        self._dist = None if dist is None or not bool((dist != 0).any()) else dist.float().contiguous()
        self._c2w, self._intr = cameras.camera_to_worlds.contiguous(), cameras.intrinsics()

    def _launch(self, u: Tensor, ray_indices, origins, directions, pixel_area, dnorm, cams, rgb) -> None:
        n, h, w, ch = self.images.shape
        call("b2n_pixel_sample_raygen", ptr(self._c2w), ptr(self._intr), ptr(self._dist), ptr(self.images, torch.uint8), n, h,
             w, ch, ptr(self.image_idx, torch.int64), ptr(u), u.shape[0], ptr(ray_indices, torch.int64), ptr(origins),
             ptr(directions), ptr(pixel_area), ptr(dnorm), ptr(cams, torch.int64), ptr(rgb), stream())

    def sample(self, u: Optional[Tensor] = None) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """One training batch on the device.  u: optional recorded U(0,1) draws [R,3]."""
        dev = self.images.device
        if u is None:
            u = torch.rand(self.R, 3, device=dev)
        u = u.float().contiguous()
        R = u.shape[0]
        idx = torch.empty(R, 3, device=dev, dtype=torch.int64)
        o, d = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev)
        area, nrm = torch.empty(R, 1, device=dev), torch.empty(R, 1, device=dev)
        cams = torch.empty(R, 1, device=dev, dtype=torch.int64)
        rgb = torch.empty(R, 3, device=dev)
        self._launch(u, idx, o, d, area, nrm, cams, rgb)
        bundle = RayBundle(origins=o, directions=d, pixel_area=area, camera_indices=cams, metadata={"directions_norm": nrm})
        return bundle, {"image": rgb, "indices": idx}

    def next_train(self, step: int) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """`DataManager.next_train` (data/datamanagers/base_datamanager.py): (RayBundle, batch) for one iteration."""
        return self.sample()

    def fill_engine(self, engine, u: Optional[Tensor] = None) -> None:
        """Write the batch straight into a `NerfactoStep`'s static input buffers: no host round trip, no copies."""
        if u is None:
            u = torch.rand(engine.R, 3, device=self.images.device)
        self._launch(u.float().contiguous(), None, engine.origins_in, engine.directions_in, None, None, engine.cams, engine.gt)
