"""``StructuredMeshExperiment`` — host-side mirror of fourierflow/routines/structured_mesh.py:8-51
(training / validation / test steps): ``out = self.model(x)``; relative-L2 against ``y``."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..modules.loss import LpLoss
from .base import RoutineMixin


class StructuredMeshExperiment(RoutineMixin, nn.Module):
    def __init__(self, model: nn.Module, loss_scale: float = 1.0, **kwargs):
        super().__init__()
        self.model = model
        self.l2_loss = LpLoss(size_average=True)
        self.loss_scale = loss_scale

    def training_step(self, batch, batch_idx: int = 0, optimizer=None, scheduler=None, clip_val=None,
                      world_size: int = 1):
        """structured_mesh.py:22-32: loss = LpLoss(model(x), y); the scaled loss goes through the manual optimisation
        (differentiable through ffno_block_bwd); returns the scaled loss like the reference."""
        x, y = batch['x'], batch['y']
        B = x.shape[0]
        out = self.model(x)
        loss = self.l2_loss(out.reshape(B, -1), y.reshape(B, -1)) * self.loss_scale
        self.optimize_manually(loss, batch_idx, optimizer, scheduler, clip_val, world_size)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        x, y = batch['x'], batch['y']
        B = x.shape[0]
        out = self.model(x)
        return self.l2_loss(out.reshape(B, -1), y.reshape(B, -1))

    test_step = validation_step

    def forward(self, batch):
        return self.model(batch['x'] if isinstance(batch, dict) else batch)
