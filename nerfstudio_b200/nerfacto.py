"""The nerfacto hot path assembled from the B200 components — the unit `bench.py` times.

Mirrors the wiring of nerfstudio/models/nerfacto.py (`populate_modules` :139-252, `get_outputs` :298-348,
`get_metrics_dict` :350-361, `get_loss_dict` :363-391) with the same configuration field names and defaults, so
that numbers are quoted on the reference's own model.  When nerfstudio itself is installed the unmodified
`NerfactoModel` runs on these components through `nerfstudio_b200.integration.install()`; this class exists so
the path can be driven (tests, smoke, bench) on a box that has only this repository.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Literal, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from .cameras.camera_optimizers import CameraOptimizer, CameraOptimizerConfig
from .cameras.rays import RayBundle
from .field_components.spatial_distortions import SceneContraction
from .fields.density_fields import HashMLPDensityField
from .fields.nerfacto_field import NerfactoField
from .field_components.field_heads import FieldHeadNames
from .model_components.losses import distortion_loss, interlevel_loss
from .model_components.ray_samplers import ProposalNetworkSampler, UniformSampler
from .model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
from .model_components.scene_colliders import NearFarCollider


@dataclass
class NerfactoModelConfig:
    """Hot-path fields of nerfstudio's NerfactoModelConfig (models/nerfacto.py:50-136), same names and defaults."""

    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: Literal["random", "last_sample", "black", "white"] = "last_sample"
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_update_every: int = 5
    proposal_warmup: int = 5000
    num_proposal_iterations: int = 2
    use_same_proposal_network: bool = False
    proposal_net_args_list: List[Dict] = field(default_factory=lambda: [
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
    ])
    proposal_initial_sampler: Literal["piecewise", "uniform"] = "piecewise"
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    use_proposal_weight_anneal: bool = True
    use_appearance_embedding: bool = True
    use_average_appearance_embedding: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    use_single_jitter: bool = True
    disable_scene_contraction: bool = False
    implementation: Literal["tcnn", "torch"] = "torch"
    appearance_embed_dim: int = 32
    average_init_density: float = 1.0
    eval_num_rays_per_chunk: int = 1 << 15
    camera_optimizer: CameraOptimizerConfig = field(default_factory=lambda: CameraOptimizerConfig(mode="SO3xR3"))


class NerfactoModel(nn.Module):
    def __init__(self, config: NerfactoModelConfig, aabb: Tensor, num_train_data: int) -> None:
        super().__init__()
        self.config = config
        self.num_train_data = num_train_data
        c = config
        contraction = None if c.disable_scene_contraction else SceneContraction(order=float("inf"))
        self.field = NerfactoField(
            aabb, hidden_dim=c.hidden_dim, num_levels=c.num_levels, max_res=c.max_res, base_res=c.base_res,
            features_per_level=c.features_per_level, log2_hashmap_size=c.log2_hashmap_size,
            hidden_dim_color=c.hidden_dim_color, spatial_distortion=contraction, num_images=num_train_data,
            use_average_appearance_embedding=c.use_average_appearance_embedding,
            appearance_embedding_dim=c.appearance_embed_dim if c.use_appearance_embedding else 0,
            average_init_density=c.average_init_density, implementation=c.implementation)
        # models/nerfacto.py:176-178: built between the field and the proposal networks (parameter order matters for the
        # flat gradient buffer: [field | camera_opt | proposal_networks])
        self.camera_optimizer = CameraOptimizer(c.camera_optimizer, num_cameras=num_train_data, device="cpu")
        self.proposal_networks = nn.ModuleList()
        n_nets = 1 if c.use_same_proposal_network else c.num_proposal_iterations
        for i in range(n_nets):
            args = c.proposal_net_args_list[min(i, len(c.proposal_net_args_list) - 1)]
            self.proposal_networks.append(HashMLPDensityField(
                aabb, spatial_distortion=contraction, **args, average_init_density=c.average_init_density,
                implementation=c.implementation))
        nets = [self.proposal_networks[0 if c.use_same_proposal_network else i] for i in range(c.num_proposal_iterations)]
        self.density_fns = [n.density_fn for n in nets]

        def update_schedule(step):
            return np.clip(np.interp(step, [0, c.proposal_warmup], [0, c.proposal_update_every]), 1, c.proposal_update_every)

        initial = UniformSampler(single_jitter=c.use_single_jitter) if c.proposal_initial_sampler == "uniform" else None
        self.proposal_sampler = ProposalNetworkSampler(
            num_nerf_samples_per_ray=c.num_nerf_samples_per_ray,
            num_proposal_samples_per_ray=c.num_proposal_samples_per_ray,
            num_proposal_network_iterations=c.num_proposal_iterations, single_jitter=c.use_single_jitter,
            update_sched=update_schedule, initial_sampler=initial)
        self.collider = NearFarCollider(near_plane=c.near_plane, far_plane=c.far_plane)
        self.renderer_rgb = RGBRenderer(background_color=c.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="median")
        self.renderer_expected_depth = DepthRenderer(method="expected")
        self.rgb_loss = nn.MSELoss()
        self.step = 0

    # ---- training-schedule hooks (models/nerfacto.py:262-296) ----
    def before_train_iteration(self, step: int) -> None:
        c = self.config
        self.step = step
        if c.use_proposal_weight_anneal:
            n = c.proposal_weights_anneal_max_num_iters
            train_frac = np.clip(step / n, 0, 1)
            bias = lambda x, b: b * x / ((b - 1) * x + 1)
            self.proposal_sampler.set_anneal(float(bias(train_frac, c.proposal_weights_anneal_slope)))

    def after_train_iteration(self, step: int) -> None:
        self.proposal_sampler.step_cb(step)

    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        groups = {"proposal_networks": list(self.proposal_networks.parameters()), "fields": list(self.field.parameters())}
        self.camera_optimizer.get_param_groups(param_groups=groups)  # adds "camera_opt" unless the mode is "off"
        return groups

    # ---- forward (models/base_model.py:132-143, models/nerfacto.py:298-348) ----
    def forward(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    def get_outputs(self, ray_bundle: RayBundle) -> Dict:
        if self.training:  # models/nerfacto.py:300-301
            self.camera_optimizer.apply_to_raybundle(ray_bundle)
        ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
        field_outputs = self.field.forward(ray_samples)
        weights = ray_samples.get_weights(field_outputs[FieldHeadNames.DENSITY])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        rgb = self.renderer_rgb(rgb=field_outputs[FieldHeadNames.RGB], weights=weights)
        with torch.no_grad():
            depth = self.renderer_depth(weights=weights, ray_samples=ray_samples)
        expected_depth = self.renderer_expected_depth(weights=weights, ray_samples=ray_samples)
        accumulation = self.renderer_accumulation(weights=weights)
        outputs = {"rgb": rgb, "accumulation": accumulation, "depth": depth, "expected_depth": expected_depth}
        if self.training:
            outputs["weights_list"] = weights_list
            outputs["ray_samples_list"] = ray_samples_list
        for i in range(self.config.num_proposal_iterations):
            with torch.no_grad():
                outputs[f"prop_depth_{i}"] = self.renderer_depth(weights=weights_list[i], ray_samples=ray_samples_list[i])
        return outputs

    def get_metrics_dict(self, outputs, batch) -> Dict[str, Tensor]:
        gt = self.renderer_rgb.blend_background(batch["image"])
        metrics = {"psnr": -10.0 * torch.log10(torch.mean((outputs["rgb"].detach() - gt) ** 2))}
        if self.training:
            metrics["distortion"] = distortion_loss(outputs["weights_list"], outputs["ray_samples_list"])
        self.camera_optimizer.get_metrics_dict(metrics)
        return metrics

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, Tensor]:
        pred, gt = self.renderer_rgb.blend_background_for_loss_computation(
            pred_image=outputs["rgb"], pred_accumulation=outputs["accumulation"], gt_image=batch["image"])
        loss = {"rgb_loss": self.rgb_loss(gt, pred)}
        if self.training:
            loss["interlevel_loss"] = self.config.interlevel_loss_mult * interlevel_loss(
                outputs["weights_list"], outputs["ray_samples_list"])
            assert metrics_dict is not None and "distortion" in metrics_dict
            loss["distortion_loss"] = self.config.distortion_loss_mult * metrics_dict["distortion"]
            self.camera_optimizer.get_loss_dict(loss)  # models/nerfacto.py:389-390
        return loss

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, ray_bundle: RayBundle) -> Dict[str, Tensor]:
        """Chunked inference over a flat ray bundle (models/base_model.py:177-205)."""
        chunk = self.config.eval_num_rays_per_chunk
        outs: Dict[str, List[Tensor]] = {}
        for i in range(0, len(ray_bundle), chunk):
            o = self.forward(ray_bundle[i: i + chunk])
            for k, v in o.items():
                if isinstance(v, Tensor):
                    outs.setdefault(k, []).append(v)
        return {k: torch.cat(v) for k, v in outs.items()}


class Trainer:
    """One optimisation step = forward + losses + backward + Adam, as engine/trainer.py:486-530 runs it with the
    nerfacto optimiser settings (configs/method_configs.py:106-119: Adam lr 1e-2 eps 1e-15 for both groups).
    Gradient averaging across ranks (the DDP allreduce of pipelines/base_pipeline.py:280-281) is done by
    `nerfstudio_b200.distributed.FlatGradAllReduce` when world_size > 1."""

    def __init__(self, model: NerfactoModel, lr: float = 1e-2, eps: float = 1e-15, allreduce=None,
                 camera_lr: float = 1e-3) -> None:
        from .optim import FlatAdam

        self.model = model
        self.optim = FlatAdam(model, lr=lr, eps=eps, group_lr={"camera_opt": camera_lr})
        self.allreduce = allreduce
        self.step = 0

    def train_iteration(self, ray_bundle: RayBundle, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        m = self.model
        m.train()
        m.before_train_iteration(self.step)
        self.optim.zero_grad()
        outputs = m(ray_bundle)
        metrics = m.get_metrics_dict(outputs, batch)
        loss_dict = m.get_loss_dict(outputs, batch, metrics)
        loss = sum(loss_dict.values())
        loss.backward()
        scale = 1.0
        if self.allreduce is not None:
            scale = self.allreduce(self.optim.flat_grad)
        # frozen proposal networks are not stepped (their grads are None in the reference: engine/optimizers.py:155)
        active = None if m.proposal_sampler.last_updated else [g for g in self.optim.group_steps if g != "proposal_networks"]
        self.optim.step(grad_scale=scale, active=active)
        m.after_train_iteration(self.step)
        self.step += 1
        loss_dict["loss"] = loss.detach()
        loss_dict["psnr"] = metrics["psnr"]
        return loss_dict
