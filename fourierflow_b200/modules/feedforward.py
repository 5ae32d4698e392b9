"""``FeedForward`` — host-side mirror of fourierflow/modules/feedforward.py:6-24.

The module tree is what the reference builds (so ``state_dict`` keys match: ``layers.{i}.0.{bias,weight_g,weight_v}``
and, with ``layer_norm``, ``layers.{n-1}.3.{weight,bias}``): ``n_layers`` stages, each a 4-slot ``nn.Sequential``
``[WNLinear, Dropout, activation, norm]``; widths ``dim -> dim*factor -> ... -> dim``; ReLU after every stage but the
last; LayerNorm only on the last stage.

Inside a block the whole FF runs fused in libffno_b200 (``ffno_block_fwd``: tcgen05 kernel ``ff_ts_kernel``); called on
its own it runs through the stateless C-ABI ops ``ffno_linear_fwd`` / ``ffno_layernorm_fwd`` (CUDA, no torch math).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _ops
from .linear import WNLinear


def _stage(width_in: int, width_out: int, weight_norm: bool, p_drop: float, is_last: bool, with_norm: bool) -> nn.Sequential:
    slots = [
        WNLinear(width_in, width_out, wnorm=weight_norm),                     # slot 0: the only slot with FF weights
        nn.Dropout(p_drop),                                                   # slot 1: p = 0 in every shipped config
        nn.Identity() if is_last else nn.ReLU(inplace=True),                  # slot 2
        nn.LayerNorm(width_out) if (with_norm and is_last) else nn.Identity(),  # slot 3
    ]
    return nn.Sequential(*slots)


class FeedForward(nn.Module):
    def __init__(self, dim, factor, ff_weight_norm, n_layers, layer_norm, dropout):
        super().__init__()
        self.dim, self.factor, self.n_layers = dim, factor, n_layers
        self.layer_norm = bool(layer_norm)
        self.dropout = dropout
        hidden = dim * factor
        widths = [dim] + [hidden] * (n_layers - 1) + [dim]
        self.layers = nn.ModuleList(
            _stage(widths[i], widths[i + 1], ff_weight_norm, dropout, i == n_layers - 1, self.layer_norm)
            for i in range(n_layers))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _ops.require_inference(self, x)
        last = self.n_layers - 1
        for i, stage in enumerate(self.layers):
            x = _ops.linear_forward(stage[0], x, relu=i < last)
            if isinstance(stage[3], nn.LayerNorm):
                x = _ops.layernorm_forward(stage[3], x)
        return x
