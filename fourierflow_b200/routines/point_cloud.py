"""``PointCloudExperiment`` — host-side mirror of fourierflow/routines/point_cloud.py:9-65 (elasticity: the geo-F-FNO on
the mesh points ``xy`` with the geometry code ``rr`` and the learned deformation ``iphi``), without pytorch_lightning.
Inference / evaluation only: the geo operator has no CUDA backward."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..modules.loss import LpLoss
from .base import RoutineMixin


class PointCloudExperiment(RoutineMixin, nn.Module):
    def __init__(self, model: nn.Module, iphi: nn.Module, N: int, **kwargs):
        super().__init__()
        self.model = model
        self.iphi = iphi
        self.N = N
        self.l2_loss = LpLoss(size_average=True)

    def forward(self, batch):
        """point_cloud.py:21-27: the predicted stress at every mesh point."""
        xy, rr = batch['xy'].cuda(), batch['rr'].cuda()
        return self.model(xy, code=rr, iphi=self.iphi)

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        """point_cloud.py:46-54: relative L2 of the prediction against ``sigma``."""
        sigma = batch['sigma'].cuda()
        B = sigma.shape[0]
        out = self.forward(batch)
        return self.l2_loss(out.reshape(B, -1), sigma.reshape(B, -1))

    test_step = validation_step

    def training_step(self, batch, batch_idx=0, **kwargs):
        raise RuntimeError("PointCloudExperiment (B200 backend): the geo operator is inference-only (no CUDA backward for the "
                           "point-cloud end layers); train with the reference, evaluate / predict here")
