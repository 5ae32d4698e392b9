"""``PointCloudExperiment`` — host-side mirror of fourierflow/routines/point_cloud.py:9-65 (elasticity: the geo-F-FNO on
the mesh points ``xy`` with the geometry code ``rr`` and the learned deformation ``iphi``), without pytorch_lightning.
Training goes through torch autograd for the point-cloud end layers / ``iphi`` and through ffno_layers_bwd for the
interior layer loop."""
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

    def training_step(self, batch, batch_idx: int = 0, optimizer=None, scheduler=None, clip_val=None, world_size: int = 1):
        """point_cloud.py:29-43: data loss + 0 x the deformation regulariser on N random points, then the manual
        optimisation (routines/base.py:27-52)."""
        xy, rr, sigma = batch['xy'].cuda(), batch['rr'].cuda(), batch['sigma'].cuda()
        B = rr.shape[0]
        out = self.model(xy, code=rr, iphi=self.iphi)
        loss_data = self.l2_loss(out.reshape(B, -1), sigma.reshape(B, -1))
        samples_x = torch.rand(B, self.N, 2, device=xy.device) * 3 - 1
        samples_xi = self.iphi(samples_x, code=rr)
        loss_reg = self.l2_loss(samples_xi.reshape(B, -1), samples_x.reshape(B, -1))
        loss = loss_data + 0 * loss_reg
        self.optimize_manually(loss, batch_idx, optimizer, scheduler, clip_val, world_size)
        return loss
