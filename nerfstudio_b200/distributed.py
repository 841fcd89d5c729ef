"""Ray-batch data parallelism: every rank renders its own rays; ONE collective per step averages the flat
gradient buffer (replaces DDP's bucketed allreduce + `find_unused_parameters` graph walk of the reference:
nerfstudio/pipelines/base_pipeline.py:279-282, scripts/train.py:139-157).

The sum-allreduce runs through torch.distributed (NCCL over NVLink/NVSwitch on a B200 box; gloo in CPU tests);
the 1/world_size scaling is folded into the fused Adam kernel's `grad_scale`, so no extra pass over the buffer."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl") -> tuple:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the process group when W > 1."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


class FlatGradAllReduce:
    """callable(flat_grad) -> grad_scale.  Sums the flat gradient over ranks in place, returns 1/world_size."""

    def __init__(self, group=None) -> None:
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, flat_grad: torch.Tensor) -> float:
        if self.world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / self.world

    def start(self, segment: torch.Tensor):
        """Launch the sum of one contiguous segment asynchronously (NCCL runs it on its own stream, ordered after
        the work already enqueued on the current stream) so that it overlaps whatever is enqueued next."""
        if self.world <= 1:
            return None
        return dist.all_reduce(segment, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    # ---- sharded update of a large segment (the 67 MB hash table): reduce-scatter -> every rank runs Adam on its 1/W slice
    # -> all-gather of the updated parameters.  Same bytes on the wire as the all-reduce it replaces, but the second half
    # moves AFTER the optimiser and overlaps the next step's proposal sampling, and Adam touches 1/W of the segment.
    def shard_chunk(self, n: int) -> int:
        """Elements per rank (multiple of 4: 16-byte aligned slices) for a segment of n elements; the tail
        n - world * chunk (< 4 * world elements) is left to the caller's all-reduce."""
        return (n // self.world) // 4 * 4

    def start_reduce_scatter(self, segment: torch.Tensor):
        """Sum `segment` (numel = world * chunk) over ranks; afterwards rank r's slice [r*chunk, (r+1)*chunk) holds the
        sums (the other slices are unspecified).  In place, asynchronous."""
        if self.world <= 1:
            return None
        chunk = segment.numel() // self.world
        assert chunk * self.world == segment.numel()
        rank = dist.get_rank(self.group)
        mine = segment[rank * chunk: (rank + 1) * chunk]
        if dist.get_backend(self.group) == "gloo":  # CPU tests: gloo has no reduce-scatter
            return dist.all_reduce(segment, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return dist.reduce_scatter_tensor(mine, segment, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def start_all_gather(self, segment: torch.Tensor):
        """Every rank contributes its slice [r*chunk, (r+1)*chunk) of `segment`; afterwards all of `segment` is identical
        on every rank.  In place, asynchronous."""
        if self.world <= 1:
            return None
        chunk = segment.numel() // self.world
        assert chunk * self.world == segment.numel()
        rank = dist.get_rank(self.group)
        mine = segment[rank * chunk: (rank + 1) * chunk]
        if dist.get_backend(self.group) == "gloo":  # in-place all_gather_into_tensor is NCCL-only
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)
            for r, p_ in enumerate(parts):
                segment[r * chunk: (r + 1) * chunk].copy_(p_)
            return None
        return dist.all_gather_into_tensor(segment, mine, group=self.group, async_op=True)

    @staticmethod
    def finish(*handles) -> None:
        """Make the current stream wait for the collectives started with `start`."""
        for h in handles:
            if h is not None:
                h.wait()


def broadcast_parameters(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s weights (what DDP's constructor does)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
