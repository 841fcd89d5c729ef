"""`gsplat`-compatible surface (the subset nerfstudio's splatfacto calls) on the B200 splatting kernels (csrc/splat.cu).

nerfstudio imports `from gsplat.rendering import rasterization` and `from gsplat.strategy import DefaultStrategy,
MCMCStrategy` at module import time (models/splatfacto.py:26-31) and calls `rasterization(...)` once per step
(:555-581).  gsplat 1.4.0's sources are not available here: semantics follow its published algorithm, restated in
oracle/splat_oracle.py — parity with gsplat itself is UNPINNED, parity with the oracle (values and gradients) is tested.
The densification strategies (gsplat.strategy) are control plane and are not provided.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .. import functional as F
from ..lib import call, host_floats, ptr, stream

NULL = C.c_void_p(0)


def _cam_args(viewmat: Tensor, K: Tensor):
    vm = host_floats(viewmat.detach().float().cpu().reshape(-1).tolist())
    kk = host_floats(K.detach().float().cpu().reshape(-1).tolist())
    return vm, kk


class _Rasterize(torch.autograd.Function):
    """One camera: (means, quats, scales, opacities, colors | sh) -> (render [H,W,3|4], alpha [H,W,1])."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, colors, viewmat, K, width, height, near, far, sh_degree, eps2d,
                radius_clip, with_depth):
        means, quats, scales = means.contiguous().float(), quats.contiguous().float(), scales.contiguous().float()
        opacities, colors = opacities.contiguous().float(), colors.contiguous().float()
        N, dev = means.shape[0], means.device
        vm, kk = _cam_args(viewmat, K)
        use_sh = sh_degree is not None
        means2d, depths = torch.empty(N, 2, device=dev), torch.empty(N, device=dev)
        conics = torch.empty(N, 3, device=dev)
        radii, touched = torch.empty(N, device=dev, dtype=torch.int32), torch.empty(N, device=dev, dtype=torch.int32)
        rgb = torch.empty(N, 3, device=dev) if use_sh else colors
        call("b2n_gs_project_fwd", ptr(means), ptr(quats), ptr(scales), ptr(colors) if use_sh else NULL,
             colors.shape[1] if use_sh else 0, int(sh_degree) if use_sh else 0, N, C.cast(vm, C.c_void_p),
             C.cast(kk, C.c_void_p), width, height, float(near), float(far), float(eps2d), float(radius_clip), ptr(means2d),
             ptr(depths), ptr(conics), ptr(radii, torch.int32), ptr(touched, torch.int32), ptr(rgb) if use_sh else NULL,
             stream())
        offsets, M = F.scan_counts(touched)
        keys = torch.empty(M, device=dev, dtype=torch.int64)
        ids = torch.empty(M, device=dev, dtype=torch.int32)
        tiles = ((width + 15) // 16) * ((height + 15) // 16)
        tile_lo = torch.zeros(tiles, device=dev, dtype=torch.int32)
        tile_hi = torch.zeros(tiles, device=dev, dtype=torch.int32)
        if M > 0:
            call("b2n_gs_emit", ptr(means2d), ptr(radii, torch.int32), ptr(depths), ptr(offsets, torch.int64), N, width, height,
                 ptr(keys, torch.int64), ptr(ids, torch.int32), stream())
            keys, perm = torch.sort(keys, stable=True)  # 64-bit radix sort (library primitive, like gsplat's cub call)
            ids = ids[perm].contiguous()
            call("b2n_gs_tile_ranges", ptr(keys, torch.int64), M, ptr(tile_lo, torch.int32), ptr(tile_hi, torch.int32), stream())
        ch = 4 if with_depth else 3
        out = torch.empty(height, width, ch, device=dev)
        alpha = torch.empty(height, width, device=dev)
        last = torch.empty(height, width, device=dev, dtype=torch.int32)
        call("b2n_gs_rasterize_fwd", width, height, ptr(tile_lo, torch.int32), ptr(tile_hi, torch.int32), ptr(ids, torch.int32),
             ptr(means2d), ptr(conics), ptr(opacities), ptr(rgb), ptr(depths) if with_depth else NULL, ptr(out), ptr(alpha),
             ptr(last, torch.int32), stream())
        ctx.save_for_backward(means, quats, scales, opacities, colors, rgb, means2d, depths, conics, radii, ids, tile_lo, tile_hi,
                              alpha, last)
        ctx.cfg = (vm, kk, width, height, near, far, sh_degree, eps2d, radius_clip, with_depth)
        ctx.info = dict(means2d=means2d, radii=radii, depths=depths, conics=conics, tiles_per_gauss=touched, isect_ids=keys,
                        flatten_ids=ids, width=width, height=height, n_cameras=1)
        ctx.mark_non_differentiable(radii)
        return out, alpha[..., None], means2d, radii

    @staticmethod
    def backward(ctx, v_out, v_alpha, v_means2d_user, _v_radii):
        (means, quats, scales, opacities, colors, rgb, means2d, depths, conics, radii, ids, tile_lo, tile_hi, alpha,
         last) = ctx.saved_tensors
        vm, kk, width, height, near, far, sh_degree, eps2d, radius_clip, with_depth = ctx.cfg
        N, dev = means.shape[0], means.device
        use_sh = sh_degree is not None
        v_m2, v_con = torch.zeros(N, 2, device=dev), torch.zeros(N, 3, device=dev)
        v_op, v_rgb = torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
        v_dep = torch.zeros(N, device=dev) if with_depth else None
        call("b2n_gs_rasterize_bwd", width, height, ptr(tile_lo, torch.int32), ptr(tile_hi, torch.int32), ptr(ids, torch.int32),
             ptr(means2d), ptr(conics), ptr(opacities), ptr(rgb), ptr(depths) if with_depth else NULL, ptr(alpha),
             ptr(last, torch.int32), ptr(v_out.contiguous().float()),
             ptr(v_alpha.contiguous().float().reshape(height, width)) if v_alpha is not None else NULL, ptr(v_m2), ptr(v_con),
             ptr(v_op), ptr(v_rgb), ptr(v_dep), stream())
        if v_means2d_user is not None:
            v_m2 = v_m2 + v_means2d_user
        v_means, v_quats, v_scales = torch.empty_like(means), torch.empty_like(quats), torch.empty_like(scales)
        v_sh = torch.empty_like(colors) if use_sh else None
        call("b2n_gs_project_bwd", ptr(means), ptr(quats), ptr(scales), ptr(colors) if use_sh else NULL,
             colors.shape[1] if use_sh else 0, int(sh_degree) if use_sh else 0, N, C.cast(vm, C.c_void_p),
             C.cast(kk, C.c_void_p), width, height, float(near), float(far), float(eps2d), float(radius_clip),
             ptr(radii, torch.int32), ptr(conics), ptr(rgb), ptr(v_m2), ptr(v_dep), ptr(v_con), ptr(v_rgb) if use_sh else NULL,
             ptr(v_means), ptr(v_quats), ptr(v_scales), ptr(v_sh), stream())
        ctx.v_means2d = v_m2
        v_colors = v_sh if use_sh else v_rgb
        return (v_means, v_quats, v_scales, v_op, v_colors) + (None,) * 10


def rasterization(means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor, viewmats: Tensor,
                  Ks: Tensor, width: int, height: int, near_plane: float = 0.01, far_plane: float = 1e10,
                  radius_clip: float = 0.0, eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = False,
                  tile_size: int = 16, backgrounds: Optional[Tensor] = None, render_mode: str = "RGB",
                  sparse_grad: bool = False, absgrad: bool = False, rasterize_mode: str = "classic",
                  **_unused) -> Tuple[Tensor, Tensor, Dict]:
    """gsplat.rendering.rasterization (1.4.0 signature subset): -> (render [C,H,W,3(+1)], alphas [C,H,W,1], info)."""
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16")
    if rasterize_mode != "classic":
        raise NotImplementedError("rasterize_mode 'antialiased' is not implemented")
    if render_mode not in ("RGB", "D", "ED", "RGB+D", "RGB+ED"):
        raise ValueError(f"unknown render_mode {render_mode!r}")
    if packed or sparse_grad:
        raise NotImplementedError("packed / sparse_grad are not implemented (splatfacto passes False)")
    with_depth = render_mode != "RGB"
    if means.shape[0] == 0:  # nothing to draw (splatfacto's empty crop): zeros, as get_empty_outputs expects
        C_ = viewmats.shape[0]
        ch = {"RGB": 3, "D": 1, "ED": 1}.get(render_mode, 4)
        z = means.new_zeros(C_, height, width, ch)
        if backgrounds is not None and ch >= 3:
            z = torch.cat([z[..., :3] + backgrounds[:, None, None, :], z[..., 3:]], -1)
        return z, means.new_zeros(C_, height, width, 1), dict(means2d=means.new_zeros(C_, 0, 2), width=width, height=height,
                                                               radii=torch.zeros(C_, 0, dtype=torch.int32, device=means.device),
                                                               n_cameras=C_, tile_size=16)
    renders, alphas, infos = [], [], []
    for c in range(viewmats.shape[0]):
        out, alpha, means2d, radii = _Rasterize.apply(means, quats, scales, opacities, colors, viewmats[c], Ks[c], int(width),
                                                      int(height), near_plane, far_plane, sh_degree, eps2d, radius_clip,
                                                      with_depth)
        if with_depth and render_mode in ("ED", "RGB+ED"):  # expected depth: accumulated depth / alpha
            out = torch.cat([out[..., :3], out[..., 3:4] / alpha.clamp(min=1e-10)], -1)
        if render_mode in ("D", "ED"):
            out = out[..., 3:4]
        if backgrounds is not None:
            out = torch.cat([out[..., :3] + (1.0 - alpha) * backgrounds[c], out[..., 3:]], -1)
        renders.append(out), alphas.append(alpha), infos.append((means2d, radii))
    info = dict(means2d=torch.stack([m for m, _ in infos]), radii=torch.stack([r for _, r in infos]), width=width,
                height=height, n_cameras=viewmats.shape[0], tile_size=16)
    return torch.stack(renders), torch.stack(alphas), info


class _NoDensification:
    """Stand-in for gsplat.strategy.{DefaultStrategy, MCMCStrategy}: splatfacto constructs one and calls its hooks every
    step (models/splatfacto.py:294-330,583-586).  Densification / pruning is control plane (out of scope): the hooks
    keep the call protocol and change nothing, so a splatfacto model trains with a fixed set of Gaussians."""

    def __init__(self, *args, **kwargs) -> None:
        self.absgrad = bool(kwargs.get("absgrad", False))
        self.kwargs = kwargs

    def check_sanity(self, params, optimizers) -> None:
        return None

    def initialize_state(self, *args, **kwargs) -> dict:
        return {}

    def step_pre_backward(self, params, optimizers, state, step, info) -> None:
        if isinstance(info, dict) and "means2d" in info and info["means2d"].requires_grad:
            info["means2d"].retain_grad()

    def step_post_backward(self, *args, **kwargs) -> None:
        return None


class DefaultStrategy(_NoDensification):
    pass


class MCMCStrategy(_NoDensification):
    pass
