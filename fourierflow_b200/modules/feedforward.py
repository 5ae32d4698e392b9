"""``FeedForward`` — host-side mirror of fourierflow/modules/feedforward.py:6-24.

Parameter schema identical to the reference (``layers.{i}.0.{bias,weight_g,weight_v}``, optional
``layers.{n-1}.3.{weight,bias}`` LayerNorm).  Inside a block the whole FF runs fused in libffno_b200
(ffno_block_fwd); called on its own it runs the same CUDA kernels through ffno_linear_fwd.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _ops
from .linear import WNLinear


class FeedForward(nn.Module):
    def __init__(self, dim, factor, ff_weight_norm, n_layers, layer_norm, dropout):
        super().__init__()
        self.dim, self.factor, self.n_layers = dim, factor, n_layers
        self.layer_norm = bool(layer_norm)
        self.dropout = dropout
        self.layers = nn.ModuleList([])
        for i in range(n_layers):
            in_dim = dim if i == 0 else dim * factor
            out_dim = dim if i == n_layers - 1 else dim * factor
            self.layers.append(nn.Sequential(
                WNLinear(in_dim, out_dim, wnorm=ff_weight_norm),
                nn.Dropout(dropout),
                nn.ReLU(inplace=True) if i < n_layers - 1 else nn.Identity(),
                nn.LayerNorm(out_dim) if layer_norm and i == n_layers - 1 else nn.Identity(),
            ))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _ops.require_inference(self, x)
        for i, layer in enumerate(self.layers):
            x = _ops.linear_forward(layer[0], x, relu=i < self.n_layers - 1)
            if isinstance(layer[3], nn.LayerNorm):
                x = _ops.layernorm_forward(layer[3], x)
        return x
