"""The captured step behind the reference's TRAINING plugin surface.

The reference trains through two calls (engine/trainer.py:486-530, pipelines/base_pipeline.py:289-304):

    loss, loss_dict, metrics_dict = trainer.train_iteration(step)
        -> model_outputs, loss_dict, metrics_dict = pipeline.get_train_loss_dict(step)
               ray_bundle, batch = datamanager.next_train(step)
               model_outputs = model(ray_bundle); metrics / losses from model.get_*_dict
        -> loss.backward(); optimizers.optimizer_scaler_step_all(); schedulers

`FusedTrainStep` offers exactly those two methods (same arguments, same return tuples) and runs the whole iteration —
forward, losses, hand-written backward, gradient all-reduce, fused Adam — as ONE CUDA-graph replay of
`engine.NerfactoStep`.  It accepts either this package's `NerfactoModel` or the reference's own, unmodified
`nerfstudio.models.nerfacto.NerfactoModel` built after `integration.install()` (the attributes it reads — `config`,
`field`, `proposal_networks`, `proposal_sampler.update_sched` — are the reference's names).  `integration.
install_fused_trainer(trainer)` binds it to a reference `Trainer`.

The ray batch is taken from `datamanager.next_train(step)` as the reference's pipeline does: a `RayBundle` plus
`batch["image"]`.  Host (pinned) bundles are copied to the device inside the call; when the four tensors are views of one
pinned blob laid out by `NerfactoStep.pack_batch` (see `HostRayQueue`) that is a single asynchronous H2D copy.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from .cameras.rays import RayBundle
from .engine import NerfactoStep


class HostRayQueue:
    """Stand-in for `DataManager.next_train(step)` (data/datamanagers/base_datamanager.py): yields (RayBundle,
    batch) whose tensors live in PINNED host memory, each batch one contiguous blob — what a loader process hands over."""

    def __init__(self, engine: NerfactoStep, batches: List[Tuple[Dict[str, Tensor], Tensor]]) -> None:
        self.items = []
        R = engine.R
        for rays, gt in batches:
            blob = engine.pack_batch(rays["origins"], rays["directions"], rays["camera_indices"], gt)
            bundle = RayBundle(origins=blob[2 * R: 5 * R].view(R, 3), directions=blob[5 * R: 8 * R].view(R, 3),
                               pixel_area=rays.get("pixel_area"), camera_indices=blob[: 2 * R].view(torch.int64).view(R, 1))
            self.items.append((bundle, {"image": blob[8 * R: 11 * R].view(R, 3)}, blob))
        self.bytes_per_batch = self.items[0][2].numel() * 4 if self.items else 0

    def next_train(self, step: int):
        bundle, batch, _blob = self.items[step % len(self.items)]
        return bundle, batch


class FusedTrainStep:
    def __init__(self, model, datamanager, n_rays: int, lr: float = 1e-2, eps: float = 1e-15,
                 lr_schedule: Optional[Callable[[int], float]] = None, allreduce=None, **engine_kwargs) -> None:
        self.model, self.datamanager = model, datamanager
        self.engine = NerfactoStep(model, n_rays, lr=lr, eps=eps, lr_schedule=lr_schedule, allreduce=allreduce,
                                   **engine_kwargs)
        self.optim = self.engine.optim

    # ---- pipelines/base_pipeline.py:289-304 ----
    def _load(self, ray_bundle, batch) -> None:
        e, R = self.engine, self.engine.R
        o, d, c, g = ray_bundle.origins, ray_bundle.directions, ray_bundle.camera_indices, batch["image"]
        if len(ray_bundle) != R:
            raise ValueError(f"the captured step was built for {R} rays per batch, got {len(ray_bundle)}")
        base = c.data_ptr()
        packed = (not o.is_cuda and c.dtype == torch.int64 and o.data_ptr() == base + 8 * R
                  and d.data_ptr() == base + 20 * R and g.data_ptr() == base + 32 * R
                  and all(t.is_contiguous() for t in (o, d, c, g)))
        if packed:  # one pinned blob in the engine's input layout: a single H2D copy
            blob = torch.as_strided(c.view(torch.float32), (11 * R,), (1,))
            e.set_batch_packed(blob)
        else:
            e.set_batch(o, d, c, g)

    def get_train_loss_dict(self, step: int):
        if hasattr(self.datamanager, "fill_engine"):  # data.device_pipeline.DeviceRayPipeline: batch made on the device,
            self.datamanager.fill_engine(self.engine)  # written straight into the step's input buffers
        else:
            ray_bundle, batch = self.datamanager.next_train(step)
            self._load(ray_bundle, batch)
        e = self.engine
        losses = e.step()
        c = self.model.config
        loss_dict = {"rgb_loss": losses[0], "interlevel_loss": losses[1], "distortion_loss": losses[2]}
        model_outputs = {"rgb": e.rgb_out, "accumulation": e.acc[:, None], "depth": e.depth_med[:, None],
                         "expected_depth": e.depth_exp[:, None]}
        metrics_dict = {"distortion": losses[2] / c.distortion_loss_mult if c.distortion_loss_mult else losses[2]}
        return model_outputs, loss_dict, metrics_dict

    def flush(self) -> None:
        """Multi-GPU sharded update: wait for the in-flight parameter all-gather before anything but the next training
        step reads the model (evaluation, checkpoints)."""
        self.engine.flush()

    # ---- engine/trainer.py:486-530 ----
    def train_iteration(self, step: int):
        _, loss_dict, metrics_dict = self.get_train_loss_dict(step)
        return self.engine.losses[3], loss_dict, metrics_dict
