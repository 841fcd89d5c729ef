"""Volume renderers behind the reference's interface (mirror of nerfstudio/model_components/renderers.py:60-449).

RGB / accumulation / depth on dense `[R,S]` samples run the warp-per-ray compositing kernel; packed samples
(instant-ngp) run the per-ray segmented accumulation kernels that stand in for nerfacc.accumulate_along_rays."""
from __future__ import annotations

from typing import Literal, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from .. import functional as F
from ..cameras.rays import intervals_of

BackgroundColor = Union[Literal["random", "last_sample", "black", "white"], Tensor]
_NAMED = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
          "blue": (0.0, 0.0, 1.0)}
BACKGROUND_COLOR_OVERRIDE: Optional[Tensor] = None


def _bg(background_color):
    if isinstance(background_color, str):
        if background_color in ("random", "last_sample"):
            return background_color
        if background_color not in _NAMED:
            raise ValueError(f"unknown background colour {background_color}")
        return _NAMED[background_color]
    t = torch.as_tensor(background_color).flatten()
    if t.numel() != 3:
        raise NotImplementedError("per-ray background colours: blend outside the kernel")
    return tuple(float(v) for v in t)


class RGBRenderer(nn.Module):
    def __init__(self, background_color: BackgroundColor = "random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb: Tensor, weights: Tensor, background_color: BackgroundColor = "random",
                    ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None,
                    eval_mode: bool = False) -> Tensor:
        if BACKGROUND_COLOR_OVERRIDE is not None:
            background_color = BACKGROUND_COLOR_OVERRIDE
        if ray_indices is not None and num_rays is not None:
            if isinstance(background_color, str) and background_color == "last_sample":
                raise NotImplementedError("Background color 'last_sample' not implemented for packed samples.")
            info = F.pack_info(ray_indices, num_rays)
            comp = F.packed_accumulate(weights[..., 0], rgb, info)
            bg = _bg(background_color)
            if bg == "random":
                return comp
            acc = F.packed_accumulate(weights[..., 0], None, info)
            return comp + torch.tensor(bg, device=comp.device) * (1.0 - acc)
        shape = weights.shape[:-2]
        S = weights.shape[-2]
        w = weights.reshape(-1, S)
        comp, _, _ = F.composite(rgb.reshape(-1, S, 3), w, _unit_edges(w), _bg(background_color), eval_mode)
        return comp.view(*shape, 3)

    @classmethod
    def get_background_color(cls, background_color: BackgroundColor, shape: Tuple[int, ...], device) -> Tensor:
        assert background_color not in {"last_sample", "random"}
        c = torch.tensor(_bg(background_color), device=device) if isinstance(background_color, str) \
            else background_color.to(device)
        return c.expand(shape)

    def blend_background(self, image: Tensor, background_color: Optional[BackgroundColor] = None) -> Tensor:
        if image.size(-1) < 4:
            return image
        rgb, opacity = image[..., :3], image[..., 3:]
        if background_color is None:
            background_color = self.background_color
            if background_color in {"last_sample", "random"}:
                background_color = "black"
        bg = self.get_background_color(background_color, shape=rgb.shape, device=rgb.device)
        return rgb * opacity + bg.to(rgb.device) * (1 - opacity)

    def blend_background_for_loss_computation(self, pred_image: Tensor, pred_accumulation: Tensor, gt_image: Tensor):
        background_color = self.background_color
        if isinstance(background_color, str) and background_color == "last_sample":
            background_color = "black"
        elif isinstance(background_color, str) and background_color == "random":
            background_color = torch.rand_like(pred_image)
            pred_image = pred_image + background_color * (1.0 - pred_accumulation)
        return pred_image, self.blend_background(gt_image, background_color=background_color)

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None,
                background_color: Optional[BackgroundColor] = None) -> Tensor:
        if background_color is None:
            background_color = self.background_color
        eval_mode = not self.training
        if ray_indices is not None and eval_mode:
            rgb = torch.nan_to_num(rgb)
        out = self.combine_rgb(rgb, weights, background_color=background_color, ray_indices=ray_indices,
                               num_rays=num_rays, eval_mode=eval_mode)
        if ray_indices is not None and eval_mode:
            out = out.clamp(0.0, 1.0)
        return out


def _unit_edges(w: Tensor) -> Tensor:
    """Placeholder interval edges for compositing calls that do not need depth."""
    R, S = w.shape
    return torch.zeros(R, S + 1, device=w.device, dtype=torch.float32)


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights: Tensor, ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None) -> Tensor:
        if ray_indices is not None and num_rays is not None:
            return F.packed_accumulate(weights[..., 0], None, F.pack_info(ray_indices, num_rays))
        return torch.sum(weights, dim=-2)


class DepthRenderer(nn.Module):
    def __init__(self, method: Literal["median", "expected"] = "median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights: Tensor, ray_samples, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None) -> Tensor:
        packed = ray_indices is not None and num_rays is not None
        if self.method == "median":
            if packed:
                raise NotImplementedError("Median depth calculation is not implemented for packed samples.")
            iv = intervals_of(ray_samples)
            depth, _ = F.median_depth(weights.reshape(iv.R, iv.S), iv)
            return depth.view(*weights.shape[:-2], 1)
        if self.method == "expected":
            eps = 1e-10
            fr = ray_samples.frustums
            if packed:
                steps = (fr.starts + fr.ends) / 2
                info = F.pack_info(ray_indices, num_rays)
                depth = F.packed_accumulate(weights[..., 0], steps, info)
                acc = F.packed_accumulate(weights[..., 0], None, info)
                depth = depth / (acc + eps)
                return torch.clip(depth, steps.min(), steps.max())
            iv = intervals_of(ray_samples)
            w = weights.reshape(iv.R, iv.S)
            lo = ((iv.starts() + iv.ends()) / 2)
            _, _, depth = F.composite(torch.zeros(iv.R, iv.S, 3, device=w.device), w, iv, "random")
            return torch.clip(depth, lo.min(), lo.max()).view(*weights.shape[:-2], 1)
        raise NotImplementedError(f"Method {self.method} not implemented")
