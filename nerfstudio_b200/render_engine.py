"""Full-image / eval rendering: `Model.get_outputs_for_camera_ray_bundle` (models/base_model.py:166-205)
and `get_outputs_for_camera` (:160-164, rays from `Cameras.generate_rays(keep_shape=True)`, cameras/cameras.py:438-453)
without the per-chunk Python work of the reference.

The reference walks the image in `eval_num_rays_per_chunk` slices and, for each, runs the whole module graph (hundreds of
framework kernels and a few allocations per slice).  Here the forward of one chunk — both proposal levels, the main field,
eval-mode compositing, median depths — is the forward half of `engine.NerfactoStep` in eval mode, captured ONCE as a CUDA
graph on fixed-size buffers; an image is a loop of [device-to-device slice copy -> graph replay -> device-to-device copy
of the outputs], all asynchronous on the stream: no host round trip, no allocation, no synchronisation per chunk.  Rays of
a whole camera come from the ray-generation kernel directly on the device.  With `torch.distributed` initialised the
chunks are dealt round-robin to the ranks and the image planes are summed over NVLink (every pixel is written by exactly
one rank).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .engine import NerfactoStep


class NerfactoRender:
    OUTPUTS = ("rgb", "accumulation", "depth", "expected_depth", "prop_depth_0", "prop_depth_1")

    def __init__(self, model, chunk_rays: Optional[int] = None, use_graph: bool = True, **engine_kwargs) -> None:
        self.model = model
        self.chunk = int(chunk_rays or model.config.eval_num_rays_per_chunk)
        self.step = NerfactoStep(model, self.chunk, eval_mode=True, use_graph=False, **engine_kwargs)
        self.use_graph = use_graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    # ------------------------------------------------------------------------------------------------
    def _refresh(self) -> None:
        """Per render call: the proposal-weight anneal exponent in force and the mean appearance embedding."""
        e = self.step
        e.hyper[3] = float(getattr(self.model.proposal_sampler, "_anneal", 1.0))
        if e.emb is not None and self.model.field.use_average_appearance_embedding:
            e.emb_mean.copy_(e.emb.detach().mean(dim=0, keepdim=True))

    def _run_chunk(self) -> None:
        e = self.step
        if not self.use_graph:
            e._forward()
            return
        if self._graph is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):  # warm-up outside capture: module loading, shared-memory attributes
                e._forward()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                e._forward()
        self._graph.replay()

    def _outputs_of_chunk(self, n: int) -> Dict[str, Tensor]:
        e = self.step
        eb2 = e.eb[2]
        # DepthRenderer("expected") clips to the [min, max] of the chunk's sample mid-points (renderers.py:379-381)
        lo = ((eb2[:n, 0] + eb2[:n, 1]) / 2).amin()
        hi = ((eb2[:n, -2] + eb2[:n, -1]) / 2).amax()
        return {"rgb": e.rgb_out[:n], "accumulation": e.acc[:n, None], "depth": e.depth_med[:n, None],
                "expected_depth": torch.clip(e.depth_exp[:n, None], lo, hi), "prop_depth_0": e.prop_depth[0][:n, None],
                "prop_depth_1": e.prop_depth[1][:n, None]}

    @torch.no_grad()
    def render_rays(self, origins: Tensor, directions: Tensor, camera_indices: Optional[Tensor] = None,
                    shard: bool = False) -> Dict[str, Tensor]:
        """origins / directions [N,3] (device) -> {name: [N, C]} for the six eval outputs of nerfacto."""
        e = self.step
        N = origins.shape[0]
        dev = e.dev
        o, d = origins.reshape(-1, 3).float(), directions.reshape(-1, 3).float()
        cams = None if camera_indices is None else camera_indices.reshape(-1).to(torch.int64)
        out = {k: torch.zeros(N, 3 if k == "rgb" else 1, device=dev) for k in self.OUTPUTS}
        world = dist.get_world_size() if (shard and dist.is_initialized()) else 1
        rank = dist.get_rank() if world > 1 else 0
        self._refresh()
        for ci, i in enumerate(range(0, N, self.chunk)):
            if ci % world != rank:
                continue
            n = min(self.chunk, N - i)
            e.origins_in[:n].copy_(o[i: i + n], non_blocking=True)
            e.directions_in[:n].copy_(d[i: i + n], non_blocking=True)
            if cams is not None:
                e.cams[:n].copy_(cams[i: i + n], non_blocking=True)
            self._run_chunk()  # rows >= n of a ragged last chunk hold the previous chunk's rays: computed, never read
            for k, v in self._outputs_of_chunk(n).items():
                out[k][i: i + n].copy_(v, non_blocking=True)
        if world > 1:
            for k in self.OUTPUTS:  # every pixel was produced by exactly one rank: a sum is a gather
                dist.all_reduce(out[k], op=dist.ReduceOp.SUM)
        return out

    @torch.no_grad()
    def render_camera(self, cameras, camera_index: int, shard: bool = False) -> Dict[str, Tensor]:
        """`Model.get_outputs_for_camera`: whole-image rays (keep_shape) on the device, outputs reshaped to [H, W, C]."""
        bundle = cameras.generate_rays(camera_indices=int(camera_index), keep_shape=True)
        H, W = bundle.origins.shape[:2]
        flat = self.render_rays(bundle.origins.reshape(-1, 3), bundle.directions.reshape(-1, 3),
                                bundle.camera_indices.reshape(-1), shard=shard)
        return {k: v.view(H, W, -1) for k, v in flat.items()}


def chunk_owner(chunk_index: int, world: int) -> int:
    """Rank that renders chunk `chunk_index` (round-robin) — host logic shared with the gloo test."""
    return chunk_index % max(world, 1)
