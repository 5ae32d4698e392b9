"""3-D structured-mesh F-FNO — host-side mirror of fourierflow/modules/factorized_fno/mesh_3d.py.

The three-axis spectral layer is (as in the reference) a class named ``SpectralConv2d``;
``fourier_weight[a]`` acts on tensor axis ``a`` (0→X, 1→Y, 2→Z) (mesh_3d.py:68-71, 83-86, 98-101).
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

    def __init__(self, in_dim, out_dim, modes_x, modes_y, modes_z, forecast_ff, backcast_ff, fourier_weight,
                 factor, ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout):
        super().__init__()
        if in_dim != out_dim:
            raise RuntimeError("SpectralConv2d: the B200 backend needs in_dim == out_dim")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.modes_x, self.modes_y, self.modes_z = modes_x, modes_y, modes_z
        self.use_fork = use_fork
        self.factor, self.n_ff_layers, self.layer_norm = factor, n_ff_layers, layer_norm

        self.fourier_weight = fourier_weight
        if not self.fourier_weight:
            self.fourier_weight = nn.ParameterList([])
            for n_modes in [modes_x, modes_y, modes_z]:
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
        return _ops.LayerSpec([self.fourier_weight[0], self.fourier_weight[1], self.fourier_weight[2]],
                              self.backcast_ff, self.forecast_ff if self.use_fork else None)

    def _plan(self, x: torch.Tensor) -> _ops.StackPlan:
        plan = self._get_plan(
            x.device, x.shape[1:4], pad=(0, 0, 0), modes=(self.modes_x, self.modes_y, self.modes_z),
            width=self.in_dim, in_features=1, append_grid=False, out_features=1, head_hidden=1, n_layers=1,
            ff_factor=self.factor, n_ff_layers=self.n_ff_layers, layer_norm=self.layer_norm,
            use_fork=self.use_fork, mode='full', path=default_path(), transform=self._transform)
        plan.sync_params(self._flat_params(), None, None, [self.layer_spec()])
        return plan

    def forward(self, x):
        x = check_input(x, 3, self.in_dim, "SpectralConv2d.forward")
        _ops.require_inference(self, x)
        plan = self._plan(x)
        s = plan.spectral_forward(0, x)
        b = plan.ff_forward(0, 0, s, None)
        f = plan.ff_forward(0, 1, s, None) if self.use_fork else None
        return b, f

    def forward_fourier(self, x):
        """mesh_3d.py:55-112."""
        x = check_input(x, 3, self.in_dim, "SpectralConv2d.forward_fourier")
        _ops.require_inference(self, x)
        return self._plan(x).spectral_forward(0, x)


class FNOFactorizedMesh3D(PlanCacheMixin, nn.Module):
    _transform, _weight_tail, _layer_cls = "rfft", (2,), SpectralConv2d

    def __init__(self, modes_x, modes_y, modes_z, width, input_dim, output_dim, n_layers, share_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm):
        super().__init__()
        self.padding = 8  # pad the domain if input is non-periodic
        self.modes_x, self.modes_y, self.modes_z = modes_x, modes_y, modes_z
        self.width, self.input_dim, self.output_dim = width, input_dim, output_dim
        self.in_proj = WNLinear(input_dim, self.width, wnorm=ff_weight_norm)
        self.n_layers = n_layers
        self.factor, self.n_ff_layers, self.layer_norm = factor, n_ff_layers, layer_norm

        self.fourier_weight = None
        if share_weight:
            self.fourier_weight = nn.ParameterList([])
            for n_modes in [modes_x, modes_y, modes_z]:
                param = nn.Parameter(torch.empty(width, width, n_modes, *self._weight_tail))
                nn.init.xavier_normal_(param)
                self.fourier_weight.append(param)

        self.spectral_layers = nn.ModuleList([])
        for _ in range(n_layers):
            self.spectral_layers.append(self._layer_cls(
                in_dim=width, out_dim=width, modes_x=modes_x, modes_y=modes_y, modes_z=modes_z,
                forecast_ff=None, backcast_ff=None, fourier_weight=self.fourier_weight, factor=factor,
                ff_weight_norm=ff_weight_norm, n_ff_layers=n_ff_layers, layer_norm=layer_norm,
                use_fork=False, dropout=0.0))

        self.out = nn.Sequential(WNLinear(self.width, 128, wnorm=ff_weight_norm),
                                 WNLinear(128, output_dim, wnorm=ff_weight_norm))

    def plan_for(self, device, size, path: str = None) -> _ops.StackPlan:
        plan = self._get_plan(
            device, size, pad=(self.padding,) * 3, modes=(self.modes_x, self.modes_y, self.modes_z),
            width=self.width, in_features=self.input_dim - 3, append_grid=True, out_features=self.output_dim,
            head_hidden=128, n_layers=self.n_layers, ff_factor=self.factor, n_ff_layers=self.n_ff_layers,
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
        """x:[B, S1, S2, S3, input_dim-3] → [B, S1, S2, S3, output_dim] (mesh_3d.py:160-176)."""
        x = check_input(x, 3, self.input_dim - 3, "FNOFactorizedMesh3D.forward")
        if _ops.needs_grad(self, x):      # training (routines/structured_mesh.py): differentiable through ffno_block_bwd
            check_trainable(self)
            return StackFunction.apply(self, x, *self._flat_params())
        out, _ = self.plan_for(x.device, x.shape[1:4]).block_forward(x)
        return out
