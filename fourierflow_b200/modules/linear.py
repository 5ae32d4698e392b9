"""``WNLinear`` — host-side mirror of fourierflow/modules/linear.py:41-79.

Same constructor, same parameter names (``bias, weight_g, weight_v`` with ``wnorm=True``; ``weight, bias``
otherwise) and the same seeded initialisation as the reference, so reference checkpoints load unchanged.
Unlike the reference there is no ``torch.nn.utils.weight_norm`` forward pre-hook: the fold
``w = g * v / ||v||_row`` runs inside libffno_b200 (csrc/generic_kernels.cu: weight_fold_transpose_kernel)
when a plan loads parameters, which also makes ``copy.deepcopy`` (SWA) work without the reference's
``_fix_weight_norm_deepcopy`` patch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _ops


class WNLinear(nn.Linear):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None,
                 wnorm: bool = False):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias, device=device,
                         dtype=dtype)
        self.wnorm = bool(wnorm)
        if wnorm:
            # torch weight_norm(dim=0): g = ||w||_2 over every dim but 0 (shape [out, 1]), v = w;
            # `weight` is removed and g, v appended — same registration order as the reference.
            w = self._parameters.pop("weight").detach()
            self.weight_g = nn.Parameter(w.norm(dim=1, keepdim=True))
            self.weight_v = nn.Parameter(w.clone())

    @property
    def weight(self):                       # noqa: D401 — nn.Linear API compatibility
        """Effective weight.  With ``wnorm`` this is a fresh tensor each access (torch ops, host-side
        convenience for inspection/export such as commands/infer.py:95-122 — not used by forward)."""
        w = self._parameters.get("weight")
        if w is not None:
            return w
        v, g = self.weight_v, self.weight_g
        return v * (g / v.norm(dim=1, keepdim=True))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _ops.linear_forward(self, x, relu=False)

    def extra_repr(self) -> str:
        return super().extra_repr() + f", wnorm={self.wnorm}"
