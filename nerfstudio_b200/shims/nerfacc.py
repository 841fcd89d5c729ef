"""`nerfacc`-compatible surface (the subset nerfstudio calls) on the B200 kernels.

nerfstudio imports `nerfacc` at module import time (model_components/ray_samplers.py:24, renderers.py:34) and
calls `OccGridEstimator(...).sampling / .update_every_n_steps`, `pack_info`, `render_weight_from_density`,
`accumulate_along_rays`, `ray_aabb_intersect` (models/instant_ngp.py:120-198).  nerfacc 0.5.2's
sources are not available here; semantics follow its published behaviour and are restated in
oracle/nerf_oracle.py — parity with nerfacc itself is UNPINNED, parity with the oracle is bit-exact for
ray indices / counts.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from .. import functional as F


def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    if n_rays is None:
        n_rays = int(ray_indices.max().item()) + 1 if ray_indices.numel() else 0
    return F.pack_info(ray_indices, n_rays)


def render_weight_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Optional[Tensor] = None,
                               ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                               prefix_trans: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    if prefix_trans is not None:
        raise NotImplementedError("prefix_trans is not used by nerfstudio")
    if packed_info is None:
        assert ray_indices is not None and n_rays is not None, "packed_info or (ray_indices, n_rays) required"
        packed_info = F.pack_info(ray_indices, n_rays)
    return F.packed_weights(t_starts, t_ends, sigmas, packed_info)


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                          n_rays: Optional[int] = None) -> Tensor:
    assert ray_indices is not None and n_rays is not None
    return F.packed_accumulate(weights, values, F.pack_info(ray_indices, n_rays))


def ray_aabb_intersect(rays_o: Tensor, rays_d: Tensor, aabbs: Tensor, near_plane: float = -float("inf"),
                       far_plane: float = float("inf"), miss_value: float = float("inf")):
    """-> (t_mins [n,K], t_maxs [n,K], hits [n,K]) — small slab test, torch ops (not on the per-step path)."""
    o, d = rays_o[:, None, :], rays_d[:, None, :]
    inv = 1.0 / d
    t1 = (aabbs[None, :, :3] - o) * inv
    t2 = (aabbs[None, :, 3:] - o) * inv
    tmin = torch.minimum(t1, t2).amax(-1).clamp(min=near_plane)
    tmax = torch.maximum(t1, t2).amin(-1).clamp(max=far_plane)
    hits = tmax > tmin
    return torch.where(hits, tmin, torch.full_like(tmin, miss_value)), \
        torch.where(hits, tmax, torch.full_like(tmax, miss_value)), hits


class OccGridEstimator(nn.Module):
    """Multi-level occupancy grid: marching (K7), visibility pruning (K8) and EMA update (K9)."""

    def __init__(self, roi_aabb: Union[Tensor, list], resolution: Union[int, list, Tensor] = 128, levels: int = 1) -> None:
        super().__init__()
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        res = torch.as_tensor(resolution, dtype=torch.int32)
        assert int(res[0]) == int(res[1]) == int(res[2]), "cubic grids only"
        self.levels = levels
        self.cells_per_lvl = int(res.prod().item())
        centre, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
        aabbs = torch.stack([torch.cat([centre - half * 2 ** i, centre + half * 2 ** i]) for i in range(levels)])
        self.register_buffer("resolution", res)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + res.tolist(), dtype=torch.bool))
        r = int(res[0])
        coords = torch.stack(torch.meshgrid([torch.arange(r)] * 3, indexing="ij"), dim=-1).reshape(-1, 3)
        self.register_buffer("grid_coords", coords, persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None,
                 alpha_fn: Optional[Callable] = None, near_plane: float = 0.0, far_plane: float = 1e10,
                 t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None, render_step_size: float = 1e-3,
                 early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0) -> Tuple[Tensor, Tensor, Tensor]:
        jitter = torch.rand(rays_o.shape[0], device=rays_o.device) if stratified else None
        ri, ts, te = F.occgrid_march(rays_o, rays_d, self.binaries, self.aabbs[0].tolist(), render_step_size,
                                     near_plane, far_plane, cone_angle, jitter, t_min, t_max)
        if (sigma_fn is not None or alpha_fn is not None) and ri.numel() > 0:
            alpha_thre = min(alpha_thre, float(self.occs.mean().item()))
            n_rays = rays_o.shape[0]
            info = F.pack_info(ri, n_rays)
            if sigma_fn is not None:
                sigmas = sigma_fn(ts, te, ri).reshape(-1).float()
                _, trans, alphas = F.packed_weights(ts, te, sigmas, info)
            else:
                alphas = alpha_fn(ts, te, ri).reshape(-1).float()
                # T = exclusive product of (1 - alpha): reuse the density kernel with sigma*dt = -log(1-alpha)
                sd = -torch.log1p(-alphas.clamp(max=1 - 1e-7))
                _, trans, _ = F.packed_weights(torch.zeros_like(sd), torch.ones_like(sd), sd, info)
            keep = (trans >= early_stop_eps) & (alphas >= alpha_thre)
            ri, ts, te = ri[keep], ts[keep], te[keep]
        return ri, ts, te

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        if not self.training or step % n != 0:
            return
        r = int(self.resolution[0])
        dev = self.occs.device
        for lvl in range(self.levels):
            if step < warmup_steps:
                idx = self.grid_indices
            else:
                n_s = self.cells_per_lvl // 4
                uniform = torch.randint(self.cells_per_lvl, (n_s,), device=dev)
                occupied = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
                if occupied.numel() > n_s:
                    occupied = occupied[torch.randint(occupied.numel(), (n_s,), device=dev)]
                idx = torch.cat([uniform, occupied])
            coords = self.grid_coords[idx]
            x = (coords + torch.rand(coords.shape, device=dev)) / r
            lo, hi = self.aabbs[lvl, :3], self.aabbs[lvl, 3:]
            occ = occ_eval_fn(lo + x * (hi - lo)).reshape(-1).float()
            cell = idx + lvl * self.cells_per_lvl
            self.occs[cell] = torch.maximum(self.occs[cell] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
