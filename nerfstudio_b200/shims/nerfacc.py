"""`nerfacc`-compatible surface (the subset nerfstudio calls) on the B200 kernels.

nerfstudio imports `nerfacc` at module import time (model_components/ray_samplers.py:24, renderers.py:34) and
calls `OccGridEstimator(...).sampling / .update_every_n_steps`, `pack_info`, `render_weight_from_density`,
`accumulate_along_rays`, `ray_aabb_intersect` (models/instant_ngp.py:120-198).  nerfacc 0.5.2's
sources are not available here; semantics follow its published behaviour and are restated in
oracle/nerf_oracle.py — parity with nerfacc itself is UNPINNED, parity with the oracle is bit-exact for
ray indices / counts / occupancy bits.  Every arithmetic step runs in csrc/packed.cu; torch supplies buffers and
random draws only.
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
    """-> (t_mins [n,K], t_maxs [n,K], hits [n,K]): slab test kernel (packed.cu:ray_aabb_kernel)."""
    return F.ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane, far_plane, miss_value)


class OccGridEstimator(nn.Module):
    """Multi-level occupancy grid: marching (K7), visibility pruning (K8) and EMA update (K9) on the kernels of
    csrc/packed.cu.  torch is used for the buffers (state_dict-visible, as nerfacc's) and for the random draws
    (stratified offsets, cell sampling, jitter) — the reference never seeds these per op, so tests pass them in."""

    def __init__(self, roi_aabb: Union[Tensor, list], resolution: Union[int, list, Tensor] = 128, levels: int = 1) -> None:
        super().__init__()
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        res = torch.as_tensor(resolution, dtype=torch.int32)
        assert int(res[0]) == int(res[1]) == int(res[2]), "cubic grids only"
        self.levels = levels
        self.cells_per_lvl = int(res.prod().item())
        self._res = int(res[0])
        centre, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
        aabbs = torch.stack([torch.cat([centre - half * 2 ** i, centre + half * 2 ** i]) for i in range(levels)])
        self._roi = roi.tolist()
        self._level_boxes = [aabbs[i].tolist() for i in range(levels)]
        self.register_buffer("resolution", res)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + res.tolist(), dtype=torch.bool))
        self._stats: Optional[Tensor] = None  # device [threshold, mean(occs)] of the current occs (None = stale)

    def _load_from_state_dict(self, *a, **k):
        self._stats = None
        return super()._load_from_state_dict(*a, **k)

    def _occ_stats(self) -> Tensor:
        if self._stats is None or self._stats.device != self.occs.device:
            self._stats = F.occgrid_binarize(self.occs, 1e-2, None)
        return self._stats

    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None,
                 alpha_fn: Optional[Callable] = None, near_plane: float = 0.0, far_plane: float = 1e10,
                 t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None, render_step_size: float = 1e-3,
                 early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0, jitter: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """`jitter` (extension, tests): the per-ray stratified offsets in [0,1) instead of a fresh torch.rand draw."""
        if jitter is None and stratified:
            jitter = torch.rand(rays_o.shape[0], device=rays_o.device)
        ri, ts, te = F.occgrid_march(rays_o, rays_d, self.binaries, self._roi, render_step_size,
                                     near_plane, far_plane, cone_angle, jitter, t_min, t_max)
        if (sigma_fn is not None or alpha_fn is not None) and ri.numel() > 0:
            info = F.pack_info(ri, rays_o.shape[0])
            if sigma_fn is not None:
                sigmas = sigma_fn(ts, te, ri).reshape(-1).float()
                _, trans, alphas = F.packed_weights(ts, te, sigmas, info)
            else:
                alphas = alpha_fn(ts, te, ri).reshape(-1).float()
                # T = exclusive product of (1 - alpha): reuse the density kernel with sigma*dt = -log(1-alpha)
                sd = -torch.log1p(-alphas.clamp(max=1 - 1e-7))
                _, trans, _ = F.packed_weights(torch.zeros_like(sd), torch.ones_like(sd), sd, info)
            # alpha_thre = min(alpha_thre, occs.mean()) — the mean stays on the device (stats[1])
            ri, ts, te = F.packed_prune(ri, ts, te, trans, alphas, info, early_stop_eps, alpha_thre,
                                        alpha_cap=self._occ_stats()[1:2])
        return ri, ts, te

    def _sample_cells(self, lvl: int, n: int) -> Tensor:
        """nerfacc `_sample_uniform_and_occupied_cells`: n uniform cells + (up to) n currently occupied ones."""
        dev = self.occs.device
        uniform = torch.randint(self.cells_per_lvl, (n,), device=dev)
        occupied = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
        if n < occupied.numel():
            occupied = occupied[torch.randint(occupied.numel(), (n,), device=dev)]
        return torch.cat([uniform, occupied])

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16, cells: Optional[list] = None,
                             jitters: Optional[list] = None) -> None:
        """`cells` / `jitters` (extension, tests): per-level recorded cell ids (None = all cells) and [n,3] jitter
        instead of fresh draws."""
        if not self.training or step % n != 0:
            return
        dev = self.occs.device
        for lvl in range(self.levels):
            if cells is not None:
                idx = cells[lvl]
            elif step < warmup_steps:
                idx = None  # every cell of the level, in order
            else:
                idx = self._sample_cells(lvl, self.cells_per_lvl // 4)
            count = self.cells_per_lvl if idx is None else idx.numel()
            jit = jitters[lvl] if jitters is not None else torch.rand(count, 3, device=dev)
            x = F.occgrid_points(idx, jit, self._res, self._level_boxes[lvl])
            occ = occ_eval_fn(x).reshape(-1).float()
            F.occgrid_ema(self.occs, idx, occ, lvl * self.cells_per_lvl, ema_decay)
        self._stats = F.occgrid_binarize(self.occs, occ_thre, self.binaries)
