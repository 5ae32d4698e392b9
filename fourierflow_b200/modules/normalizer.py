"""``Normalizer`` — host-side mirror of fourierflow/modules/normalizer.py:6-77 (same buffers, same API).

Statistics accumulation happens once per dataset in training mode (cold path, plain torch on whatever
device the buffers live on).  In the rollout hot path the normalise / de-normalise arithmetic is fused
into libffno_b200's rollout kernels (ffno_rollout_fwd), which take ``mean``/``std`` from this module.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Normalizer(nn.Module):
    def __init__(self, size, max_accumulations=10**6, std_epsilon=1e-8):
        super().__init__()
        self.max_accumulations = max_accumulations
        self.register_buffer("count", torch.tensor(0.0))
        self.register_buffer("n_accumulations", torch.tensor(0.0))
        self.register_buffer("sum", torch.full(size, 0.0))
        self.register_buffer("sum_squared", torch.full(size, 0.0))
        self.register_buffer("one", torch.tensor(1.0))
        self.register_buffer("std_epsilon", torch.full(size, std_epsilon))

    def accumulate(self, x: torch.Tensor) -> None:
        """Fold ``x[..., H]`` into the running sums (normalizer.py:18-26)."""
        flat = x.reshape(-1, x.shape[-1])
        self.sum += flat.sum(dim=0)
        self.sum_squared += (flat ** 2).sum(dim=0)
        self.count += flat.shape[0]
        self.n_accumulations += 1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and self.n_accumulations < self.max_accumulations:
            self.accumulate(x)
        return (x - self.mean) / self.std

    def inverse(self, x: torch.Tensor, channel=None) -> torch.Tensor:
        if channel is None:
            return x * self.std + self.mean
        return x * self.std[channel] + self.mean[channel]

    @property
    def mean(self):
        safe_count = torch.maximum(self.count, self.one)
        return self.sum / safe_count

    @property
    def std(self):
        safe_count = torch.maximum(self.count, self.one)
        std = torch.sqrt(self.sum_squared / safe_count - self.mean ** 2)
        return torch.maximum(std, self.std_epsilon)
