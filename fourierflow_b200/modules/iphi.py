"""``IPhi`` — the learned coordinate deformation of the geo-operators (mirror of fourierflow/modules/iphi.py:6-58).

Torch glue OUTSIDE the hot path: a small tanh MLP applied to the mesh points before the Fourier sums of the
end layers of ``FNOFactorizedPointCloud2D``.  It is here so that the ``experiments/elasticity/ffno*`` configs load and run on
this backend; the parameter names and their creation order follow the reference (checkpoints, seeded initialisation).
Unlike the reference (iphi.py:21-23 builds its constants with ``device="cuda"`` in the constructor) the module can be
constructed without a GPU; the constants follow the input's device."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class IPhi(nn.Module):
    def __init__(self, width=32):
        super().__init__()
        self.width = width
        self.fc0 = nn.Linear(4, self.width)
        self.fc_code = nn.Linear(42, self.width)
        self.fc_no_code = nn.Linear(3 * self.width, 4 * self.width)
        self.fc1 = nn.Linear(4 * self.width, 4 * self.width)
        self.fc2 = nn.Linear(4 * self.width, 4 * self.width)
        self.fc3 = nn.Linear(4 * self.width, 4 * self.width)
        self.fc4 = nn.Linear(4 * self.width, 2)
        self.activation = torch.tanh

    def forward(self, x, code=None):
        """x:[B, N, 2] mesh points, code:[B, 42] geometry features -> deformed points x + x * mlp(features) (iphi.py:26-58).
        Features per point: (x0, x1, angle, radius) about the centre (1e-4, 1e-4), their image under ``fc0``, and
        sin / cos of the four raw features at the frequencies pi * 2^j, j < width / 4."""
        b, n = x.shape[0], x.shape[1]
        rel = x - 0.0001
        raw = torch.stack([x[..., 0], x[..., 1], torch.atan2(rel[..., 1], rel[..., 0]), torch.norm(rel, dim=-1, p=2)], dim=-1)
        freqs = math.pi * torch.pow(2, torch.arange(0, self.width // 4, dtype=torch.float, device=x.device))
        ang = (raw.unsqueeze(-1) * freqs).reshape(b, n, 4 * (self.width // 4))
        feats = torch.cat([self.fc0(raw), torch.sin(ang), torch.cos(ang)], dim=-1).reshape(b, n, 3 * self.width)
        if code is not None:
            feats = torch.cat([self.fc_code(code).unsqueeze(1).expand(b, n, self.width), feats], dim=-1)
        else:
            feats = self.fc_no_code(feats)
        h = self.activation(self.fc1(feats))
        h = self.activation(self.fc2(h))
        h = self.activation(self.fc3(h))
        return x + x * self.fc4(h)
