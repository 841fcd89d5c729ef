"""Near/far assignment (mirror of nerfstudio/model_components/scene_colliders.py:28-191)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import functional as F


class SceneCollider(nn.Module):
    def __init__(self, **kwargs) -> None:
        self.kwargs = kwargs
        super().__init__()

    def set_nears_and_fars(self, ray_bundle):
        raise NotImplementedError

    def forward(self, ray_bundle):
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        return self.set_nears_and_fars(ray_bundle)


class AABBBoxCollider(SceneCollider):
    """Slab test against the scene box; `scene_box` is anything with an `.aabb` [2,3] tensor."""

    def __init__(self, scene_box, near_plane: float = 0.0, **kwargs) -> None:
        super().__init__(**kwargs)
        self.scene_box, self.near_plane = scene_box, near_plane

    def set_nears_and_fars(self, ray_bundle):
        aabb = self.scene_box.aabb if hasattr(self.scene_box, "aabb") else self.scene_box
        near_plane = self.near_plane if self.training else 0
        n, f = F.aabb_collide(ray_bundle.origins, ray_bundle.directions, torch.as_tensor(aabb).flatten().tolist(), near_plane)
        ray_bundle.nears, ray_bundle.fars = n, f
        return ray_bundle


class NearFarCollider(SceneCollider):
    def __init__(self, near_plane: float, far_plane: float, reset_near_plane: bool = True, **kwargs) -> None:
        self.near_plane, self.far_plane, self.reset_near_plane = near_plane, far_plane, reset_near_plane
        super().__init__(**kwargs)

    def set_nears_and_fars(self, ray_bundle):
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        near_plane = self.near_plane if (self.training or not self.reset_near_plane) else 0
        ray_bundle.nears, ray_bundle.fars = ones * near_plane, ones * self.far_plane
        return ray_bundle
