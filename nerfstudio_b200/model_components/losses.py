"""Proposal-network losses (mirror of nerfstudio/model_components/losses.py:53-155): interlevel (histogram
envelope) and distortion, each one kernel per level with the gradient produced in the same pass."""
from __future__ import annotations

from typing import List

import torch
from torch import Tensor, nn

from .. import functional as F

MSELoss = nn.MSELoss
L1Loss = nn.L1Loss
EPS = 1.0e-7


def ray_samples_to_sdist(ray_samples) -> Tensor:
    """Spacing-domain bin edges [R,S+1] of a RaySamples."""
    sb = getattr(ray_samples, "spacing_bins", None)
    if sb is not None:
        return sb
    return torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)


def interlevel_loss(weights_list: List[Tensor], ray_samples_list) -> Tensor:
    assert len(ray_samples_list) > 0
    return F.interlevel_loss([w[..., 0] for w in weights_list], [ray_samples_to_sdist(r) for r in ray_samples_list])


def distortion_loss(weights_list: List[Tensor], ray_samples_list) -> Tensor:
    return F.distortion_loss(weights_list[-1][..., 0], ray_samples_to_sdist(ray_samples_list[-1]))


class _DistortionRowsFn(torch.autograd.Function):
    """Per-ray distortion values [R] (losses.py:135-146), differentiable w.r.t. w."""

    @staticmethod
    def forward(ctx, t: Tensor, w: Tensor) -> Tensor:
        import ctypes as C

        from ..lib import call, ptr, stream

        t, w = t.contiguous().float(), w.contiguous().float()
        rows = torch.empty(w.shape[0], device=w.device)
        d_w = torch.empty_like(w)
        call("b2n_distortion_fwd_bwd", ptr(t), ptr(w), w.shape[0], w.shape[1], 1.0, ptr(rows), ptr(d_w), stream())
        ctx.save_for_backward(d_w)
        return rows

    @staticmethod
    def backward(ctx, g: Tensor):
        (d_w,) = ctx.saved_tensors
        return None, d_w * g[:, None]


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    """Per-ray distortion values (no mean): t [R,S+1] (or [S+1]), w [R,S] (or [S]) -> [R] (or a scalar)."""
    if w.dim() == 1:
        return _DistortionRowsFn.apply(t.detach()[None], w[None])[0]
    return _DistortionRowsFn.apply(t.detach(), w)
