"""Ray containers of the hot path, B200 layout.

Mirror of nerfstudio/cameras/rays.py:33-295 (`Frustums`, `RaySamples`, `RayBundle`) with the same attribute
names and the same methods the hot path calls (`get_positions`, `get_weights`, `get_ray_samples`), but laid out
the way the kernels consume them: per-RAY origins/directions [R,3] plus per-sample bin edges [R,S+1], instead
of stride-0 broadcast views over [R,S,3].  The broadcast views of the reference are still available
(`frustums.origins` etc. are expanded views) so code written against the reference keeps working, and
`ray_form()` also accepts the reference's own `RaySamples` objects (duck-typed), recognising their broadcast
layout without materialising it.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from .. import functional as F


@dataclass
class Frustums:
    origins: Tensor      # [..., 3]
    directions: Tensor   # [..., 3]
    starts: Tensor       # [..., 1]
    ends: Tensor         # [..., 1]
    pixel_area: Tensor   # [..., 1]
    offsets: Optional[Tensor] = None

    @property
    def shape(self):
        return torch.broadcast_shapes(self.origins.shape[:-1], self.starts.shape[:-1])

    def get_positions(self) -> Tensor:
        """o + d * (start + end) / 2   (reference: cameras/rays.py:50-59)."""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        return pos if self.offsets is None else pos + self.offsets

    def get_start_positions(self) -> Tensor:
        return self.origins + self.directions * self.starts


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None
    deltas: Optional[Tensor] = None
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None
    # dense per-ray form (set by our samplers; None for hand-built / packed samples)
    ray_origins: Optional[Tensor] = None      # [R,3]
    ray_directions: Optional[Tensor] = None   # [R,3]
    euclidean_bins: Optional[Tensor] = None   # [R,S+1]
    spacing_bins: Optional[Tensor] = None     # [R,S+1]
    ray_camera_indices: Optional[Tensor] = None  # [R]

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights(self, densities: Tensor) -> Tensor:
        """alpha-compositing weights, [..., S, 1] -> [..., S, 1]   (reference: cameras/rays.py:129-152).
        Warp-shuffle transmittance scan kernel; differentiable w.r.t. densities."""
        iv = intervals_of(self)
        w = F.render_weights(iv, densities.reshape(iv.R, iv.S))
        return w.view(densities.shape)


@dataclass
class RayBundle:
    origins: Tensor
    directions: Tensor
    pixel_area: Tensor
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self) -> int:
        return self.origins.shape[0]

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def to(self, device) -> "RayBundle":
        mv = lambda t: None if t is None else t.to(device)
        return replace(self, origins=mv(self.origins), directions=mv(self.directions), pixel_area=mv(self.pixel_area),
                       camera_indices=mv(self.camera_indices), nears=mv(self.nears), fars=mv(self.fars),
                       metadata={k: mv(v) for k, v in self.metadata.items()}, times=mv(self.times))

    def __getitem__(self, idx) -> "RayBundle":
        ix = lambda t: None if t is None else t[idx]
        return replace(self, origins=ix(self.origins), directions=ix(self.directions), pixel_area=ix(self.pixel_area),
                       camera_indices=ix(self.camera_indices), nears=ix(self.nears), fars=ix(self.fars),
                       metadata={k: ix(v) for k, v in self.metadata.items()}, times=ix(self.times))

    def reshape(self, shape) -> "RayBundle":
        """Batch dims -> `shape` (utils/tensor_dataclass.py:reshape): trailing feature dims are kept."""
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        rs = lambda t: None if t is None else t.reshape(shape + t.shape[len(self.shape):])
        return replace(self, origins=rs(self.origins), directions=rs(self.directions), pixel_area=rs(self.pixel_area),
                       camera_indices=rs(self.camera_indices), nears=rs(self.nears), fars=rs(self.fars),
                       metadata={k: rs(v) for k, v in self.metadata.items()}, times=rs(self.times))

    def flatten(self) -> "RayBundle":
        return self.reshape((-1,))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        """cameras/rays.py:235-249: rays [start, end) of the flattened (row-major) bundle."""
        return self.flatten()[start_idx:end_idx]

    def get_ray_samples(self, bin_starts: Tensor, bin_ends: Tensor, spacing_starts: Optional[Tensor] = None,
                        spacing_ends: Optional[Tensor] = None,
                        spacing_to_euclidean_fn: Optional[Callable] = None) -> RaySamples:
        """Same contract as the reference (cameras/rays.py:251-295): bins are [R,S,1]."""
        frustums = Frustums(origins=self.origins[..., None, :], directions=self.directions[..., None, :],
                            starts=bin_starts, ends=bin_ends, pixel_area=self.pixel_area[..., None, :])
        return RaySamples(
            frustums=frustums,
            camera_indices=None if self.camera_indices is None else self.camera_indices[..., None, :],
            deltas=bin_ends - bin_starts, spacing_starts=spacing_starts, spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn, metadata=self.metadata,
            times=None if self.times is None else self.times[..., None, :],
            ray_origins=self.origins, ray_directions=self.directions,
            ray_camera_indices=None if self.camera_indices is None else self.camera_indices.reshape(-1),
        )

    def samples_from_bins(self, euclidean_bins: Tensor, spacing_bins: Optional[Tensor],
                          spacing_to_euclidean_fn: Optional[Callable]) -> RaySamples:
        """Build RaySamples from [R,S+1] edge arrays; starts/ends are views of the edge array (no copies)."""
        e = euclidean_bins
        rs = self.get_ray_samples(
            bin_starts=e[:, :-1, None], bin_ends=e[:, 1:, None],
            spacing_starts=None if spacing_bins is None else spacing_bins[:, :-1, None],
            spacing_ends=None if spacing_bins is None else spacing_bins[:, 1:, None],
            spacing_to_euclidean_fn=spacing_to_euclidean_fn)
        rs.euclidean_bins, rs.spacing_bins = e, spacing_bins
        return rs


# ------------------------------------------------------------------------------------------------
# layout recognition (works for our RaySamples and, duck-typed, for nerfstudio's TensorDataclass ones)
# ------------------------------------------------------------------------------------------------
def _edges_view(starts: Tensor, ends: Tensor) -> Optional[Tensor]:
    """If starts/ends [R,S,1] are the [:-1]/[1:] views of one contiguous [R,S+1] edge array, return that array."""
    if starts.dim() != 3 or starts.shape != ends.shape or starts.shape[-1] != 1 or starts.dtype != torch.float32:
        return None
    R, S = starts.shape[0], starts.shape[1]
    st = starts.stride()
    if st[:2] != ends.stride()[:2] or st[1] != 1 or (R > 1 and st[0] != S + 1):
        return None
    if ends.data_ptr() != starts.data_ptr() + 4 or starts.untyped_storage().data_ptr() != ends.untyped_storage().data_ptr():
        return None
    return starts.as_strided((R, S + 1), (S + 1, 1))


def intervals_of(rs) -> F.Intervals:
    """Sample intervals of a RaySamples in kernel form (no copy when built by a sampler)."""
    e = getattr(rs, "euclidean_bins", None)
    fr = rs.frustums
    if e is None and fr.starts.dim() == 3:
        e = _edges_view(fr.starts, fr.ends)
    if e is not None:
        return F.Intervals.from_edges(e)
    s, en = fr.starts[..., 0], fr.ends[..., 0]
    if s.dim() == 1:  # packed samples [M,1]: M rays of one sample
        s, en = s[:, None], en[:, None]
    return F.Intervals.from_pairs(s.reshape(-1, s.shape[-1]), en.reshape(-1, en.shape[-1]))


def ray_form(rs) -> Tuple[Tensor, Tensor, F.Intervals]:
    """(origins [R,3], directions [R,3], intervals) of a RaySamples without materialising broadcast views."""
    fr = rs.frustums
    iv = intervals_of(rs)
    o = getattr(rs, "ray_origins", None)
    d = getattr(rs, "ray_directions", None)
    if o is None or d is None:
        o, d = fr.origins, fr.directions
        if o.dim() == 3:
            if o.shape[1] == 1 or o.stride(1) == 0:  # [R,1,3] or a stride-0 broadcast over samples
                o, d = o[:, 0, :], d[:, 0, :]
            else:  # genuinely per-sample origins: treat every sample as its own ray
                o, d = o.reshape(-1, 3), d.reshape(-1, 3)
                iv = F.Intervals.from_pairs(iv.starts().reshape(-1, 1), iv.ends().reshape(-1, 1))
    return o, d, iv


def sample_camera_indices(rs, n_rays: int, n_samples: int) -> Optional[Tensor]:
    """Per-ray camera index [R] (int64) when it is constant along the ray, else per-sample [R*S]."""
    ci = getattr(rs, "ray_camera_indices", None)
    if ci is not None:
        return ci
    ci = rs.camera_indices
    if ci is None:
        return None
    ci = ci.squeeze(-1) if ci.shape[-1] == 1 else ci
    if ci.dim() == 2 and (ci.shape[1] == 1 or ci.stride(1) == 0):
        return ci[:, 0]
    return ci.reshape(-1)
