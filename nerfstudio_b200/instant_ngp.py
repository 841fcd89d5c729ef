"""The instant-ngp (packed) path assembled from the B200 components.

Mirrors the wiring of nerfstudio/models/instant_ngp.py (`populate_modules` :88-147, `update_occupancy_grid`
:149-164, `get_outputs` :173-218, losses :220-246): an occupancy grid marched into PACKED samples, the nerfacto
field on those samples, packed transmittance weights and per-ray accumulation.  nerfacc's role is played by
`nerfstudio_b200.shims.nerfacc` (parity with nerfacc itself unpinned — DESIGN.md §2).  Like nerfacto.py this file
exists so that the path can be driven without nerfstudio installed; inside nerfstudio the unmodified NGPModel runs on
the same components via `integration.install()`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Literal, Optional

import torch
from torch import Tensor, nn

from .cameras.rays import RayBundle
from .field_components.field_heads import FieldHeadNames
from .field_components.spatial_distortions import SceneContraction
from .fields.nerfacto_field import NerfactoField
from .model_components.ray_samplers import VolumetricSampler
from .model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
from .shims import nerfacc


@dataclass
class InstantNGPModelConfig:
    """Hot-path fields of nerfstudio's InstantNGPModelConfig (models/instant_ngp.py:42-81), same names / defaults."""

    grid_resolution: int = 128
    grid_levels: int = 4
    max_res: int = 2048
    log2_hashmap_size: int = 19
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    render_step_size: Optional[float] = None
    near_plane: float = 0.05
    far_plane: float = 1e3
    use_appearance_embedding: bool = False
    background_color: Literal["random", "black", "white"] = "random"
    disable_scene_contraction: bool = False
    implementation: Literal["tcnn", "torch"] = "tcnn"


class NGPModel(nn.Module):
    def __init__(self, config: InstantNGPModelConfig, aabb: Tensor, num_train_data: int) -> None:
        super().__init__()
        self.config = config
        c = config
        contraction = None if c.disable_scene_contraction else SceneContraction(order=float("inf"))
        # NB the reference passes `0 if use_appearance_embedding else 32` (models/instant_ngp.py:107): kept as is
        self.field = NerfactoField(aabb=aabb, appearance_embedding_dim=0 if c.use_appearance_embedding else 32,
                                   num_images=num_train_data, log2_hashmap_size=c.log2_hashmap_size, max_res=c.max_res,
                                   spatial_distortion=contraction, implementation=c.implementation)
        self.scene_aabb = nn.Parameter(aabb.flatten(), requires_grad=False)
        if c.render_step_size is None:  # ~1000 samples across the box diagonal
            c.render_step_size = float(((self.scene_aabb[3:] - self.scene_aabb[:3]) ** 2).sum().sqrt().item() / 1000)
        self.occupancy_grid = nerfacc.OccGridEstimator(roi_aabb=self.scene_aabb, resolution=c.grid_resolution,
                                                       levels=c.grid_levels)
        self.sampler = VolumetricSampler(occupancy_grid=self.occupancy_grid, density_fn=self.field.density_fn)
        self.renderer_rgb = RGBRenderer(background_color=c.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="expected")
        self.rgb_loss = nn.MSELoss()

    def update_occupancy_grid(self, step: int) -> None:
        """BEFORE_TRAIN_ITERATION callback (models/instant_ngp.py:152-164)."""
        self.occupancy_grid.update_every_n_steps(
            step=step, occ_eval_fn=lambda x: self.field.density_fn(x) * self.config.render_step_size)

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        return {"fields": list(self.field.parameters())}

    def get_outputs(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        num_rays = len(ray_bundle)
        c = self.config
        with torch.no_grad():
            ray_samples, ray_indices = self.sampler(ray_bundle=ray_bundle, near_plane=c.near_plane, far_plane=c.far_plane,
                                                    render_step_size=c.render_step_size, alpha_thre=c.alpha_thre,
                                                    cone_angle=c.cone_angle)
        field_outputs = self.field(ray_samples)
        packed_info = nerfacc.pack_info(ray_indices, num_rays)
        weights = nerfacc.render_weight_from_density(
            t_starts=ray_samples.frustums.starts[..., 0], t_ends=ray_samples.frustums.ends[..., 0],
            sigmas=field_outputs[FieldHeadNames.DENSITY][..., 0], packed_info=packed_info)[0][..., None]
        rgb = self.renderer_rgb(rgb=field_outputs[FieldHeadNames.RGB], weights=weights, ray_indices=ray_indices,
                                num_rays=num_rays)
        depth = self.renderer_depth(weights=weights, ray_samples=ray_samples, ray_indices=ray_indices, num_rays=num_rays)
        accumulation = self.renderer_accumulation(weights=weights, ray_indices=ray_indices, num_rays=num_rays)
        return {"rgb": rgb, "accumulation": accumulation, "depth": depth, "num_samples_per_ray": packed_info[:, 1]}

    def forward(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        return self.get_outputs(ray_bundle)

    def get_loss_dict(self, outputs, batch) -> Dict[str, Tensor]:
        pred, gt = self.renderer_rgb.blend_background_for_loss_computation(
            pred_image=outputs["rgb"], pred_accumulation=outputs["accumulation"], gt_image=batch["image"])
        return {"rgb_loss": self.rgb_loss(gt, pred)}


class NGPTrainer:
    """One instant-ngp optimisation step as engine/trainer.py:486-530 runs it: occupancy callback
    (BEFORE_TRAIN_ITERATION, models/instant_ngp.py:149-164), forward on packed samples, MSE loss, backward, (gradient
    all-reduce), fused Adam over one flat buffer (configs/method_configs.py:263-270: Adam lr 1e-2 eps 1e-15)."""

    def __init__(self, model: NGPModel, lr: float = 1e-2, eps: float = 1e-15, allreduce=None) -> None:
        from .optim import FlatAdam

        self.model, self.allreduce, self.step = model, allreduce, 0
        self.optim = FlatAdam(model, lr=lr, eps=eps)
        self.last_num_samples = 0

    def train_iteration(self, ray_bundle: RayBundle, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        m = self.model
        m.train()
        m.update_occupancy_grid(self.step)
        self.optim.zero_grad()
        out = m(ray_bundle)
        loss_dict = m.get_loss_dict(out, batch)
        loss = loss_dict["rgb_loss"]
        loss.backward()
        scale = self.allreduce(self.optim.flat_grad) if self.allreduce is not None else 1.0
        self.optim.step(grad_scale=scale)
        self.step += 1
        self.num_samples = out["num_samples_per_ray"]  # device tensor; summed by the caller when it wants M
        return {"loss": loss.detach(), "rgb_loss": loss.detach()}
