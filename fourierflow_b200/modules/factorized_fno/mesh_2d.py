"""2-D structured-mesh F-FNO — host-side mirror of fourierflow/modules/factorized_fno/mesh_2d.py.

Per-axis mode counts, linspace grid features appended to the input, +8 zero padding on the high side
of both axes before the layer stack and a crop before the head (mesh_2d.py:149-165) — all folded into
libffno_b200's lift / head kernels.  ``fourier_weight[a]`` acts on tensor axis ``a`` (0→X, 1→Y).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _ops
from ..feedforward import FeedForward
from ..linear import WNLinear
from ._base import PlanCacheMixin, StackFunction, check_input, check_trainable, default_path


class SpectralConv2d(PlanCacheMixin, nn.Module):
    _transform, _weight_tail = "rfft", (2,)      # the factorized_cno mirrors override these (DCT, real weights)

    def __init__(self, in_dim, out_dim, modes_x, modes_y, forecast_ff, backcast_ff, fourier_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode):
        super().__init__()
        if in_dim != out_dim:
            raise RuntimeError("SpectralConv2d: the B200 backend needs in_dim == out_dim")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.modes_x, self.modes_y = modes_x, modes_y
        self.mode, self.use_fork = mode, use_fork
        self.factor, self.n_ff_layers, self.layer_norm = factor, n_ff_layers, layer_norm

        self.fourier_weight = fourier_weight
        if not self.fourier_weight:
            self.fourier_weight = nn.ParameterList([])
            for n_modes in [modes_x, modes_y]:
                param = nn.Parameter(torch.empty(in_dim, out_dim, n_modes, *self._weight_tail))
                nn.init.xavier_normal_(param)
                self.fourier_weight.append(param)

        if use_fork:
            self.forecast_ff = forecast_ff
            if not self.forecast_ff:
                self.forecast_ff = FeedForward(out_dim, factor, ff_weight_norm, n_ff_layers, layer_norm, dropout)

        self.backcast_ff = backcast_ff
        if not self.backcast_ff:
            self.backcast_ff = FeedForward(out_dim, factor, ff_weight_norm, n_ff_layers, layer_norm, dropout)

    def layer_spec(self) -> _ops.LayerSpec:
        return _ops.LayerSpec([self.fourier_weight[0], self.fourier_weight[1]], self.backcast_ff,
                              self.forecast_ff if self.use_fork else None)

    def _plan(self, x: torch.Tensor) -> _ops.StackPlan:
        plan = self._get_plan(
            x.device, x.shape[1:3], pad=(0, 0), modes=(self.modes_x, self.modes_y), width=self.in_dim,
            in_features=1, append_grid=False, out_features=1, head_hidden=1, n_layers=1,
            ff_factor=self.factor, n_ff_layers=self.n_ff_layers, layer_norm=self.layer_norm,
            use_fork=self.use_fork, mode=self.mode, path=default_path(),
            transform=self._transform)
        plan.sync_params(self._flat_params(), None, None, [self.layer_spec()])
        return plan

    def forward(self, x):
        x = check_input(x, 2, self.in_dim, "SpectralConv2d.forward")
        _ops.require_inference(self, x)
        plan = self._plan(x)
        s = plan.spectral_forward(0, x) if self.mode != "no-fourier" else x
        b = plan.ff_forward(0, 0, s, None)
        f = plan.ff_forward(0, 1, s, None) if self.use_fork else None
        return b, f

    def forward_fourier(self, x):
        """mesh_2d.py:56-104."""
        x = check_input(x, 2, self.in_dim, "SpectralConv2d.forward_fourier")
        _ops.require_inference(self, x)
        return self._plan(x).spectral_forward(0, x)


class FNOFactorizedMesh2D(PlanCacheMixin, nn.Module):
    _transform, _weight_tail, _layer_cls = "rfft", (2,), SpectralConv2d

    def __init__(self, modes_x, modes_y, width, input_dim, n_layers, share_weight, factor, ff_weight_norm,
                 n_ff_layers, layer_norm):
        super().__init__()
        self.padding = 8  # pad the domain if input is non-periodic
        self.modes_x, self.modes_y = modes_x, modes_y
        self.width, self.input_dim = width, input_dim
        self.in_proj = WNLinear(input_dim, self.width, wnorm=ff_weight_norm)
        self.n_layers = n_layers
        self.factor, self.n_ff_layers, self.layer_norm = factor, n_ff_layers, layer_norm

        self.fourier_weight = None
        if share_weight:
            self.fourier_weight = nn.ParameterList([])
            for n_modes in [modes_x, modes_y]:
                param = nn.Parameter(torch.empty(width, width, n_modes, *self._weight_tail))
                nn.init.xavier_normal_(param)
                self.fourier_weight.append(param)

        self.spectral_layers = nn.ModuleList([])
        for _ in range(n_layers):
            self.spectral_layers.append(self._layer_cls(
                in_dim=width, out_dim=width, modes_x=modes_x, modes_y=modes_y, forecast_ff=None,
                backcast_ff=None, fourier_weight=self.fourier_weight, factor=factor,
                ff_weight_norm=ff_weight_norm, n_ff_layers=n_ff_layers, layer_norm=layer_norm,
                use_fork=False, dropout=0.0, mode='full'))

        self.out = nn.Sequential(WNLinear(self.width, 128, wnorm=ff_weight_norm),
                                 WNLinear(128, 1, wnorm=ff_weight_norm))

    def plan_for(self, device, size, path: str = None) -> _ops.StackPlan:
        plan = self._get_plan(
            device, size, pad=(self.padding,) * 2, modes=(self.modes_x, self.modes_y), width=self.width,
            in_features=self.input_dim - 2, append_grid=True, out_features=1, head_hidden=128,
            n_layers=self.n_layers, ff_factor=self.factor, n_ff_layers=self.n_ff_layers,
            layer_norm=self.layer_norm, use_fork=False, mode='full',
            path=path or default_path(), transform=self._transform)
        params = self._flat_params()                   # validates the caches first (may drop a stale _spec_cache)
        specs = self.__dict__.get("_spec_cache")
        if specs is None:
            specs = [layer.layer_spec() for layer in self.spectral_layers]
            self.__dict__["_spec_cache"] = specs
        plan.sync_params(params, self.in_proj, self.out, specs)
        return plan

    def forward(self, x):
        """x:[B, X, Y, input_dim-2] → [B, X, Y, 1] (mesh_2d.py:149-165)."""
        x = check_input(x, 2, self.input_dim - 2, "FNOFactorizedMesh2D.forward")
        if _ops.needs_grad(self, x):      # training (routines/structured_mesh.py): differentiable through ffno_block_bwd
            check_trainable(self)
            return StackFunction.apply(self, x, *self._flat_params())
        out, _ = self.plan_for(x.device, x.shape[1:3]).block_forward(x)
        return out
