"""``LpLoss`` — host-side mirror of fourierflow/modules/loss.py (relative L2 used by the routines).

``rel`` with p=2 runs the per-sample reduction in libffno_b200 (ffno_rel_l2); the mean over the batch —
the one value a sharded run all-gathers — is a single tiny torch op.
"""
from __future__ import annotations

import torch

from .. import _ops


class LpLoss:
    def __init__(self, d=2, p=2, size_average=True, reduction=True):
        assert d > 0 and p > 0
        self.d, self.p = d, p
        self.reduction = reduction
        self.size_average = size_average

    def rel_per_sample(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.p != 2:
            raise RuntimeError("fourierflow_b200.LpLoss: only p=2 has a CUDA kernel")
        if torch.is_grad_enabled() and x.requires_grad:      # training: CUDA backward (ffno_rel_l2_bwd)
            return _ops.rel_l2_differentiable(x, y)
        return _ops.rel_l2(x, y)

    def rel(self, x, y):
        r = self.rel_per_sample(x, y)
        if self.reduction:
            return torch.mean(r) if self.size_average else torch.sum(r)
        return r

    def __call__(self, x, y):
        return self.rel(x, y)
