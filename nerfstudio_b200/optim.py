"""Flat-buffer Adam: every parameter of the model is re-homed into ONE contiguous fp32 buffer (the Parameters
become views, state_dict keys/shapes unchanged), gradients accumulate into ONE contiguous buffer, and the
optimiser step is a single fused kernel over it (b2n_adam_step).  The same flat gradient buffer is what the
multi-GPU path all-reduces with one NCCL call.

Semantics = torch.optim.Adam as the reference configures it (engine/optimizers.py:51-58;
configs/method_configs.py:106-119): lr 1e-2, betas (0.9, 0.999), eps 1e-15, no weight decay, every element
updated on every step IN WHICH ITS GROUP RECEIVED GRADIENTS.  The reference's
`Optimizers.optimizer_scaler_step_all` (engine/optimizers.py:142-159) steps a parameter group only when some
`p.grad is not None`; nerfacto's proposal networks run under `no_grad` on most steps after the first ten
(model_components/ray_samplers.py:590,604-609), so on those steps their parameters, moments and bias-correction step
count stay untouched.  `step(active=...)` reproduces that: groups come from the module's `get_param_groups()`, each
keeps its own step counter, and contiguous active segments with equal counters share one kernel launch."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from . import functional as F


class FlatAdam:
    ALIGN = 4  # elements (16 B): keeps float2/float4 table rows and the vectorised Adam kernel aligned

    def __init__(self, module: nn.Module, lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-15,
                 lr_schedule: Optional[Callable[[int], float]] = None, group_lr: Optional[dict] = None) -> None:
        seen, params = set(), []
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        for p, off in zip(params, offs):
            n = p.numel()
            self.flat[off: off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off: off + n].view(p.shape)
            p.grad = self.flat_grad[off: off + n].view(p.shape)
        self.params, self.offsets = params, offs
        self.lr, self.betas, self.eps, self.lr_schedule = lr, betas, eps, lr_schedule
        self.steps = 0
        # parameter groups (reference: Model.get_param_groups) -> contiguous [start, end) segments of the flat buffer
        group_of = {}
        if hasattr(module, "get_param_groups"):
            for name, plist in module.get_param_groups().items():
                for p in plist:
                    group_of[id(p)] = name
        self.segments = []  # [name, start, end]
        for p, off in zip(params, offs):
            name = group_of.get(id(p), "default")
            end = off + (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            if self.segments and self.segments[-1][0] == name and self.segments[-1][2] == off:
                self.segments[-1][2] = end
            else:
                self.segments.append([name, off, end])
        self.group_steps = {name: 0 for name, _, _ in self.segments}
        self.group_lr = dict(group_lr or {})  # per-group learning rate overrides (e.g. camera_opt: 1e-3)

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):  # re-attach views if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off: off + p.numel()].view(p.shape)

    def segment_of(self, group: str):
        """(start, end) of a group that occupies ONE contiguous range of the flat buffer, else None."""
        segs = [(a, b) for name, a, b in self.segments if name == group]
        return segs[0] if len(segs) == 1 else None

    def step(self, grad_scale: float = 1.0, active=None) -> None:
        """active: iterable of group names that received gradients this step (None = all)."""
        self.steps += 1
        lr = self.lr_schedule(self.steps - 1) if self.lr_schedule is not None else self.lr
        names = set(self.group_steps) if active is None else set(active) & set(self.group_steps)
        for name in names:
            self.group_steps[name] += 1
        runs = []  # merge neighbouring active segments whose step counts agree: one launch in the common case
        for name, a, b in self.segments:
            if name not in names:
                continue
            t, glr = self.group_steps[name], self.group_lr.get(name, lr)
            if runs and runs[-1][1] == a and runs[-1][2] == t and runs[-1][3] == glr:
                runs[-1][1] = b
            else:
                runs.append([a, b, t, glr])
        for a, b, t, glr in runs:
            F.adam_step(self.flat[a:b], self.flat_grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], t, glr, self.betas,
                        self.eps, grad_scale)
