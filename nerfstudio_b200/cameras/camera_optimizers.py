"""Mirror of nerfstudio/cameras/camera_optimizers.py (CameraOptimizerConfig :41-85, CameraOptimizer :87-215).

Same constructor, parameter name (`pose_adjustment` [num_cameras, 6] = translation | so(3) log-rotation), methods and
`state_dict` key, so a reference checkpoint loads unchanged.  `apply_to_raybundle` — the per-ray call on the training
hot path — runs as one fused kernel (exponential map evaluated per ray, gradients accumulated into
`pose_adjustment`); the per-camera helpers (`forward`, `apply_to_camera`, regulariser, metrics) are [C,6]-sized and
stay in torch.  Mode "SE3" is not provided (the reference recommends SO3xR3).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Literal, Optional, Union

import torch
from torch import Tensor, nn

from nerfstudio_b200 import functional as F


@dataclass
class CameraOptimizerConfig:
    mode: Literal["off", "SO3xR3", "SE3"] = "off"
    trans_l2_penalty: float = 1e-2
    rot_l2_penalty: float = 1e-3

    def setup(self, **kwargs) -> "CameraOptimizer":
        return CameraOptimizer(self, **kwargs)


def exp_map_SO3xR3(tangent_vector: Tensor) -> Tensor:
    """[B,6] -> [B,3,4] = [R | t]  (cameras/lie_groups.py:25-58), closed form: R = I + f1 K + f2 K^2."""
    t, w = tangent_vector[:, :3], tangent_vector[:, 3:]
    theta = torch.clamp((w * w).sum(1), 1e-4).sqrt()
    inv = 1.0 / theta
    f1, f2 = inv * theta.sin(), inv * inv * (1.0 - theta.cos())
    zero = torch.zeros_like(w[:, 0])
    K = torch.stack([zero, -w[:, 2], w[:, 1], w[:, 2], zero, -w[:, 0], -w[:, 1], w[:, 0], zero], dim=1).view(-1, 3, 3)
    R = f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + torch.eye(3, dtype=w.dtype, device=w.device)[None]
    return torch.cat([R, t[:, :, None]], dim=2)


def _frozen_mask(opt) -> Optional[Tensor]:
    """uint8 [num_cameras] mask of the non-trainable cameras (cached on the module), or None."""
    idx = getattr(opt, "non_trainable_camera_indices", None)
    if idx is None:
        return None
    dev = opt.pose_adjustment.device
    cached = getattr(opt, "_b2n_frozen", None)
    if cached is None or cached.device != dev:
        cached = torch.zeros(opt.num_cameras, dtype=torch.uint8, device=dev)
        cached[idx.to(dev)] = 1
        opt._b2n_frozen = cached
    return cached


def fused_apply_to_raybundle(opt, raybundle) -> None:
    """The SO3xR3 ray correction as one kernel; works on this module's CameraOptimizer and — through
    integration.install() — on the reference's (same attributes: config.mode, pose_adjustment,
    non_trainable_camera_indices)."""
    if raybundle.camera_indices is None:
        raise AttributeError("Camera indices are not provided.")
    raybundle.origins, raybundle.directions = F.pose_apply(
        opt.pose_adjustment, raybundle.camera_indices, raybundle.origins, raybundle.directions, _frozen_mask(opt))


class CameraOptimizer(nn.Module):
    config: CameraOptimizerConfig

    def __init__(self, config: CameraOptimizerConfig, num_cameras: int, device: Union[torch.device, str],
                 non_trainable_camera_indices: Optional[Tensor] = None, **kwargs) -> None:
        super().__init__()
        self.config, self.num_cameras, self.device = config, num_cameras, device
        self.non_trainable_camera_indices = non_trainable_camera_indices
        if config.mode == "off":
            pass
        elif config.mode == "SO3xR3":
            self.pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6), device=device))
        elif config.mode == "SE3":
            raise NotImplementedError("CameraOptimizer mode 'SE3' is not provided; use 'SO3xR3'")
        else:
            raise ValueError(f"unknown camera optimizer mode {config.mode!r}")

    def forward(self, indices: Tensor) -> Tensor:
        """[len(indices), 3, 4] correction matrices (identity rows for mode "off" / non-trainable cameras)."""
        if self.config.mode == "off":
            return torch.eye(4, device=self.device)[None, :3, :4].tile(indices.shape[0], 1, 1)
        out = exp_map_SO3xR3(self.pose_adjustment[indices, :])
        if self.non_trainable_camera_indices is not None:
            frozen = _frozen_mask(self)[indices].bool()
            eye = torch.eye(4, device=out.device)[:3, :4].expand_as(out)
            out = torch.where(frozen[:, None, None], eye, out)
        return out

    def apply_to_raybundle(self, raybundle) -> None:
        """origins += t[cam], directions = R[cam] @ directions (in place on the bundle, like the reference)."""
        if self.config.mode == "off":
            return
        fused_apply_to_raybundle(self, raybundle)

    def apply_to_camera(self, camera) -> Tensor:
        if self.config.mode == "off":
            return camera.camera_to_worlds
        if camera.metadata is None or "cam_idx" not in camera.metadata:
            return camera.camera_to_worlds
        adj = self(torch.tensor([camera.metadata["cam_idx"]], dtype=torch.long, device=self.pose_adjustment.device))
        adj = adj.to(camera.camera_to_worlds.device)
        c2w = camera.camera_to_worlds
        return torch.cat([torch.bmm(adj[..., :3, :3], c2w[..., :3, :3]), c2w[..., :3, 3:] + adj[..., :3, 3:]], dim=-1)

    def get_loss_dict(self, loss_dict: dict) -> None:
        if self.config.mode != "off":
            loss_dict["camera_opt_regularizer"] = (
                self.pose_adjustment[:, :3].norm(dim=-1).mean() * self.config.trans_l2_penalty
                + self.pose_adjustment[:, 3:].norm(dim=-1).mean() * self.config.rot_l2_penalty)

    def get_correction_matrices(self) -> Tensor:
        dev = self.pose_adjustment.device if self.config.mode != "off" else self.device
        return self(torch.arange(0, self.num_cameras, device=dev).long())

    def get_metrics_dict(self, metrics_dict: dict) -> None:
        if self.config.mode != "off":
            trans = self.pose_adjustment[:, :3].detach().norm(dim=-1)
            rot = self.pose_adjustment[:, 3:].detach().norm(dim=-1)
            metrics_dict["camera_opt_translation_max"] = trans.max()
            metrics_dict["camera_opt_translation_mean"] = trans.mean()
            metrics_dict["camera_opt_rotation_mean"] = math.degrees(float(rot.mean()))
            metrics_dict["camera_opt_rotation_max"] = math.degrees(float(rot.max()))

    def get_param_groups(self, param_groups: dict) -> None:
        params = list(self.parameters())
        if self.config.mode != "off":
            assert len(params) > 0
            param_groups["camera_opt"] = params
        else:
            assert len(params) == 0
