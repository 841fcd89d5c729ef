"""Ray samplers behind the reference's `Sampler` interface (mirror of nerfstudio/model_components/ray_samplers.py).

`SpacedSampler` and its subclasses (:53-248), `PDFSampler` (:251-372), `ProposalNetworkSampler` (:522-617) and
`VolumetricSampler` (:375-519) keep the reference's constructor arguments and return types, but each
`generate_ray_samples` is one kernel launch (bins + spacing->euclidean map; fp64-scanned cdf + searchsorted +
lerp) instead of ~25 elementwise torch kernels.  Stratified jitter is drawn with torch's generator exactly where
the reference draws it (`torch.rand((R,1))` / `((R,S+1))`), so seeded runs consume the same random stream.
"""
from __future__ import annotations

import contextlib

from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import functional as F
from ..cameras.rays import Frustums, RayBundle, RaySamples, intervals_of


class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def generate_ray_samples(self, *args, **kwargs) -> Any:
        raise NotImplementedError

    def forward(self, *args, **kwargs) -> Any:
        return self.generate_ray_samples(*args, **kwargs)


class _SpacingMap:
    """spacing-domain bins -> euclidean distances for one ray bundle (the `spacing_to_euclidean_fn` closure)."""

    _TORCH = {
        "uniform": (lambda x: x, lambda x: x),
        "piecewise": (lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x)),
                      lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))),
        "lindisp": (lambda x: 1 / x, lambda x: 1 / x),
        "sqrt": (torch.sqrt, lambda x: x ** 2),
        "log": (torch.log, torch.exp),
    }

    def __init__(self, kind: str, nears: Tensor, fars: Tensor):
        self.kind, self.nears, self.fars = kind, nears, fars

    def __call__(self, x: Tensor) -> Tensor:
        fn, inv = self._TORCH[self.kind]
        s_near, s_far = fn(self.nears), fn(self.fars)
        return inv(x * s_far + (1 - x) * s_near)


class SpacedSampler(Sampler):
    """Bins spaced by `spacing` in [near, far]; `spacing` is one of the named maps the kernel implements."""

    spacing = "uniform"

    def __init__(self, spacing_fn: Optional[Callable] = None, spacing_fn_inv: Optional[Callable] = None,
                 num_samples: Optional[int] = None, train_stratified=True, single_jitter=False,
                 spacing: Optional[str] = None) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified, self.single_jitter = train_stratified, single_jitter
        if spacing is not None:
            self.spacing = spacing
        if spacing_fn is not None and spacing is None and type(self) is SpacedSampler:
            raise NotImplementedError("arbitrary python spacing functions cannot run in the sampling kernel; "
                                      "use one of uniform / piecewise / lindisp / sqrt / log")

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        R = ray_bundle.origins.shape[0]
        dev = ray_bundle.origins.device
        jitter = None
        if self.train_stratified and self.training:
            jitter = torch.rand((R, 1) if self.single_jitter else (R, num_samples + 1), dtype=torch.float32, device=dev)
        sb, eb = F.spaced_sample(ray_bundle.nears, ray_bundle.fars, num_samples, self.spacing, jitter)
        return _samples(ray_bundle, eb, sb, _SpacingMap(self.spacing, ray_bundle.nears, ray_bundle.fars))


def _samples(ray_bundle, ebins: Tensor, sbins: Tensor, to_euclid) -> RaySamples:
    if isinstance(ray_bundle, RayBundle):
        return ray_bundle.samples_from_bins(ebins, sbins, to_euclid)
    # nerfstudio's own RayBundle (TensorDataclass): build its RaySamples through its own method
    return ray_bundle.get_ray_samples(bin_starts=ebins[..., :-1, None], bin_ends=ebins[..., 1:, None],
                                      spacing_starts=sbins[..., :-1, None], spacing_ends=sbins[..., 1:, None],
                                      spacing_to_euclidean_fn=to_euclid)


class UniformSampler(SpacedSampler):
    spacing = "uniform"

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class LinearDisparitySampler(SpacedSampler):
    spacing = "lindisp"

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class SqrtSampler(SpacedSampler):
    spacing = "sqrt"

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class LogSampler(SpacedSampler):
    spacing = "log"

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class UniformLinDispPiecewiseSampler(SpacedSampler):
    """First half uniform, second half linear in disparity (nerfacto's initial sampler)."""

    spacing = "piecewise"

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


def _spacing_bins(ray_samples) -> Tensor:
    sb = getattr(ray_samples, "spacing_bins", None)
    if sb is not None:
        return sb
    assert ray_samples.spacing_starts is not None and ray_samples.spacing_ends is not None, \
        "ray_sample spacing_starts and spacing_ends must be provided"
    return torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)


class PDFSampler(Sampler):
    """Inverse-CDF resampling of a weight histogram."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified, self.include_original = train_stratified, include_original
        self.histogram_padding, self.single_jitter = histogram_padding, single_jitter

    def generate_ray_samples(self, ray_bundle=None, ray_samples=None, weights: Optional[Tensor] = None,
                             num_samples: Optional[int] = None, eps: float = 1e-5, anneal: float = 1.0) -> RaySamples:
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None, "weights must be provided"
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        to_euclid = ray_samples.spacing_to_euclidean_fn
        assert to_euclid is not None, "ray_samples.spacing_to_euclidean_fn must be provided"
        sb = _spacing_bins(ray_samples)
        R = sb.shape[0]
        w = weights[..., 0] if weights.dim() == 3 else weights
        jitter = None
        if self.train_stratified and self.training:
            jitter = torch.rand((R, 1) if self.single_jitter else (R, num_samples + 1), device=sb.device)
        if isinstance(to_euclid, _SpacingMap):
            kind, nears, fars = to_euclid.kind, to_euclid.nears, to_euclid.fars
        else:  # a python closure from the reference's sampler: sample in the spacing domain, map with the closure
            kind, nears, fars = "uniform", torch.zeros(R, 1, device=sb.device), torch.ones(R, 1, device=sb.device)
        new_sb, new_eb = F.pdf_sample(sb, w, num_samples, jitter, nears, fars, kind, anneal=anneal,
                                      histogram_padding=self.histogram_padding, eps=eps)
        if self.include_original:
            new_sb, _ = torch.sort(torch.cat([sb, new_sb], -1), -1)
            new_eb = to_euclid(new_sb)
        elif not isinstance(to_euclid, _SpacingMap):
            new_eb = to_euclid(new_sb)
        return _samples(ray_bundle, new_eb, new_sb, to_euclid)


class ProposalNetworkSampler(Sampler):
    """Coarse-to-fine sampling driven by proposal density networks (mip-NeRF 360 / nerfacto)."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = False,
                 update_sched: Callable = lambda x: 1, initial_sampler: Optional[Sampler] = None,
                 pdf_sampler: Optional[PDFSampler] = None) -> None:
        super().__init__()
        if num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        self.initial_sampler = initial_sampler or UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        self.pdf_sampler = pdf_sampler or PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0
        self.last_updated = True

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def generate_ray_samples(self, ray_bundle=None, density_fns: Optional[List[Callable]] = None):
        assert ray_bundle is not None and density_fns is not None
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights, ray_samples = None, None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples)
            else:
                assert weights is not None
                # the anneal exponent (w ** anneal) is applied inside the PDF kernel
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples,
                                               anneal=self._anneal)
            if is_prop:
                fn = density_fns[i_level]
                field = getattr(fn, "__self__", None)
                # the reference adds no_grad on frozen steps and otherwise inherits the caller's grad mode (:604-609)
                with contextlib.nullcontext() if updated else torch.no_grad():
                    if field is not None and hasattr(field, "get_density") and getattr(fn, "__name__", "") == "density_fn":
                        density, _ = field.get_density(ray_samples)  # ray form: no [R,S,3] positions materialised
                    else:
                        density = fn(ray_samples.frustums.get_positions())
                    weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        self.last_updated = bool(updated)  # the trainer skips the optimiser step of frozen proposal networks
        if updated:
            self._steps_since_update = 0
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list


class VolumetricSampler(Sampler):
    """Occupancy-grid marching (instant-ngp); returns packed samples + their ray indices."""

    def __init__(self, occupancy_grid, density_fn: Optional[Callable] = None):
        super().__init__()
        assert occupancy_grid is not None
        self.density_fn = density_fn
        self.occupancy_grid = occupancy_grid

    def get_sigma_fn(self, origins, directions, times=None) -> Optional[Callable]:
        if self.density_fn is None or not self.training:
            return None
        density_fn = self.density_fn

        def sigma_fn(t_starts, t_ends, ray_indices):
            positions = F.packed_positions(origins, directions, ray_indices, t_starts, t_ends)  # o + d (ts+te)/2
            if times is None:
                return density_fn(positions).squeeze(-1)
            return density_fn(positions, times[ray_indices]).squeeze(-1)

        return sigma_fn

    def generate_ray_samples(self) -> RaySamples:
        raise RuntimeError("The VolumetricSampler fuses sample generation and density check together. "
                           "Please call forward() directly.")

    def forward(self, ray_bundle, render_step_size: float, near_plane: float = 0.0, far_plane: Optional[float] = None,
                alpha_thre: float = 0.01, cone_angle: float = 0.0):
        rays_o, rays_d = ray_bundle.origins.contiguous(), ray_bundle.directions.contiguous()
        t_min = t_max = None
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            t_min, t_max = ray_bundle.nears.contiguous().reshape(-1), ray_bundle.fars.contiguous().reshape(-1)
        far_plane = 1e10 if far_plane is None else far_plane
        camera_indices = None if ray_bundle.camera_indices is None else ray_bundle.camera_indices.contiguous()
        ray_indices, starts, ends = self.occupancy_grid.sampling(
            rays_o=rays_o, rays_d=rays_d, t_min=t_min, t_max=t_max,
            sigma_fn=self.get_sigma_fn(rays_o, rays_d, ray_bundle.times), render_step_size=render_step_size,
            near_plane=near_plane, far_plane=far_plane, stratified=self.training, cone_angle=cone_angle,
            alpha_thre=alpha_thre)
        if starts.shape[0] == 0:  # one fake sample so downstream shapes stay valid (reference :494-500)
            ray_indices = torch.zeros((1,), dtype=torch.long, device=rays_o.device)
            starts = torch.ones((1,), dtype=starts.dtype, device=rays_o.device)
            ends = torch.ones((1,), dtype=ends.dtype, device=rays_o.device)
        ray_samples = RaySamples(
            frustums=Frustums(origins=rays_o[ray_indices], directions=rays_d[ray_indices], starts=starts[..., None],
                              ends=ends[..., None], pixel_area=ray_bundle.pixel_area[ray_indices]),
            camera_indices=None if camera_indices is None else camera_indices[ray_indices])
        if ray_bundle.times is not None:
            ray_samples.times = ray_bundle.times[ray_indices]
        return ray_samples, ray_indices
