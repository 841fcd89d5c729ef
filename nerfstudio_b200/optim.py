"""Flat-buffer Adam: every parameter of the model is re-homed into ONE contiguous fp32 buffer (the Parameters
become views, state_dict keys/shapes unchanged), gradients accumulate into ONE contiguous buffer, and the
optimiser step is a single fused kernel over it (b2n_adam_step).  The same flat gradient buffer is what the
multi-GPU path all-reduces with one NCCL call.

Semantics = torch.optim.Adam as the reference configures it (engine/optimizers.py:51-58;
configs/method_configs.py:106-119): lr 1e-2, betas (0.9, 0.999), eps 1e-15, no weight decay, every element
updated every step (moments decay even where the gradient is zero)."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from . import functional as F


class FlatAdam:
    ALIGN = 4  # elements (16 B): keeps float2/float4 table rows and the vectorised Adam kernel aligned

    def __init__(self, module: nn.Module, lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-15,
                 lr_schedule: Optional[Callable[[int], float]] = None) -> None:
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

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):  # re-attach views if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off: off + p.numel()].view(p.shape)

    def step(self, grad_scale: float = 1.0) -> None:
        self.steps += 1
        lr = self.lr_schedule(self.steps - 1) if self.lr_schedule is not None else self.lr
        F.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.steps, lr, self.betas, self.eps,
                    grad_scale)
