"""2-D periodic-grid F-FNO — host-side mirror of fourierflow/modules/factorized_fno/grid_2d.py.

``SpectralConv2d`` and ``FNOFactorized2DBlock`` keep the reference's constructor arguments, parameter
names/initialisation and return types; ``forward`` hands device pointers to libffno_b200
(ffno_block_fwd / ffno_spectral_fwd / ffno_ff_fwd) instead of running eager torch ops.

Axis convention (grid_2d.py:65-68, :83-86): ``fourier_weight[0]`` mixes the modes of the LAST spatial
axis (N, "Y"), ``fourier_weight[1]`` those of M ("X").  The C ABI takes weights in tensor-axis order, so
this file passes ``[fourier_weight[1], fourier_weight[0]]``.
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

    def __init__(self, in_dim, out_dim, n_modes, forecast_ff, backcast_ff, fourier_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode):
        super().__init__()
        if in_dim != out_dim:
            raise RuntimeError("SpectralConv2d: the B200 backend needs in_dim == out_dim (as every config has)")
        self.in_dim, self.out_dim, self.n_modes = in_dim, out_dim, n_modes
        self.mode, self.use_fork = mode, use_fork
        self.factor, self.n_ff_layers, self.layer_norm = factor, n_ff_layers, layer_norm

        self.fourier_weight = fourier_weight
        if not self.fourier_weight:
            self.fourier_weight = nn.ParameterList([])
            for _ in range(2):
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

    # -- plumbing -------------------------------------------------------------------------------
    def layer_spec(self) -> _ops.LayerSpec:
        fw = [self.fourier_weight[1], self.fourier_weight[0]]      # tensor-axis order (X, Y)
        return _ops.LayerSpec(fw, self.backcast_ff, self.forecast_ff if self.use_fork else None)

    def _plan(self, x: torch.Tensor) -> _ops.StackPlan:
        plan = self._get_plan(
            x.device, x.shape[1:3], pad=(0, 0), modes=(self.n_modes, self.n_modes), width=self.in_dim,
            in_features=1, append_grid=False, out_features=1, head_hidden=1, n_layers=1,
            ff_factor=self.factor, n_ff_layers=self.n_ff_layers, layer_norm=self.layer_norm,
            use_fork=self.use_fork, mode=self.mode, path=default_path(),
            transform=self._transform)
        plan.sync_params(self._flat_params(), None, None, [self.layer_spec()])
        return plan

    # -- reference API ----------------------------------------------------------------------------
    def forward(self, x):
        """-> (b, f): backcast / forecast FeedForward of the spectral output (grid_2d.py:42-49)."""
        x = check_input(x, 2, self.in_dim, "SpectralConv2d.forward")
        _ops.require_inference(self, x)
        plan = self._plan(x)
        s = plan.spectral_forward(0, x) if self.mode != "no-fourier" else x
        b = plan.ff_forward(0, 0, s, None)
        f = plan.ff_forward(0, 1, s, None) if self.use_fork else None
        return b, f

    def forward_fourier(self, x):
        """Per-axis rfft → low-mode truncation → complex channel mix → irfft, summed over axes
        (grid_2d.py:51-99), x:[B, M, N, C] → [B, M, N, C]."""
        x = check_input(x, 2, self.in_dim, "SpectralConv2d.forward_fourier")
        _ops.require_inference(self, x)
        return self._plan(x).spectral_forward(0, x)


class FNOFactorized2DBlock(PlanCacheMixin, nn.Module):
    _transform, _weight_tail, _layer_cls = "rfft", (2,), SpectralConv2d

    def __init__(self, modes, width, input_dim=12, dropout=0.0, in_dropout=0.0, n_layers=4,
                 share_weight: bool = False, share_fork=False, factor=2, ff_weight_norm=False, n_ff_layers=2,
                 gain=1, layer_norm=False, use_fork=False, mode='full'):
        super().__init__()
        if dropout != 0.0 or in_dropout != 0.0:
            # every shipped config uses 0.0; the forward-only CUDA path has no RNG stream for dropout
            raise RuntimeError("FNOFactorized2DBlock: dropout > 0 is not supported by the B200 backend")
        if mode not in ("full", "low-pass", "no-fourier"):
            raise ValueError(f"unknown mode {mode!r}")
        self.modes, self.width, self.input_dim = modes, width, input_dim
        self.in_proj = WNLinear(input_dim, self.width, wnorm=ff_weight_norm)
        self.drop = nn.Dropout(in_dropout)
        self.n_layers, self.use_fork = n_layers, use_fork
        self.factor, self.n_ff_layers, self.layer_norm, self.mode = factor, n_ff_layers, layer_norm, mode

        self.forecast_ff = self.backcast_ff = None
        if share_fork:
            if use_fork:
                self.forecast_ff = FeedForward(width, factor, ff_weight_norm, n_ff_layers, layer_norm, dropout)
            self.backcast_ff = FeedForward(width, factor, ff_weight_norm, n_ff_layers, layer_norm, dropout)

        self.fourier_weight = None
        if share_weight:
            self.fourier_weight = nn.ParameterList([])
            for _ in range(2):
                param = nn.Parameter(torch.empty(width, width, modes, *self._weight_tail))
                nn.init.xavier_normal_(param, gain=gain)
                self.fourier_weight.append(param)

        self.spectral_layers = nn.ModuleList([])
        for _ in range(n_layers):
            self.spectral_layers.append(self._layer_cls(
                in_dim=width, out_dim=width, n_modes=modes, forecast_ff=self.forecast_ff,
                backcast_ff=self.backcast_ff, fourier_weight=self.fourier_weight, factor=factor,
                ff_weight_norm=ff_weight_norm, n_ff_layers=n_ff_layers, layer_norm=layer_norm,
                use_fork=use_fork, dropout=dropout, mode=mode))

        self.out = nn.Sequential(WNLinear(self.width, 128, wnorm=ff_weight_norm),
                                 WNLinear(128, 1, wnorm=ff_weight_norm))

    def plan_for(self, device: torch.device, size, path: str = None) -> _ops.StackPlan:
        plan = self._get_plan(
            device, size, pad=(0, 0), modes=(self.modes, self.modes), width=self.width,
            in_features=self.input_dim, append_grid=False, out_features=1, head_hidden=128,
            n_layers=self.n_layers, ff_factor=self.factor, n_ff_layers=self.n_ff_layers,
            layer_norm=self.layer_norm, use_fork=self.use_fork, mode=self.mode,
            path=path or default_path(), transform=self._transform)
        params = self._flat_params()                   # validates the caches first (may drop a stale _spec_cache)
        specs = self.__dict__.get("_spec_cache")
        if specs is None:
            specs = [layer.layer_spec() for layer in self.spectral_layers]
            self.__dict__["_spec_cache"] = specs
        plan.sync_params(params, self.in_proj, self.out, specs)
        return plan

    def forward(self, x, **kwargs):
        """x:[B, M, N, input_dim] → {'forecast': [B, M, N, 1], 'forecast_list': [...]} (grid_2d.py:154-177)."""
        x = check_input(x, 2, self.input_dim, "FNOFactorized2DBlock.forward")
        if _ops.needs_grad(self, x):
            # training (routines/grid_2d_markov.py:172-193): differentiable through ffno_block_bwd
            check_trainable(self)
            return {'forecast': StackFunction.apply(self, x, *self._flat_params()), 'forecast_list': []}
        plan = self.plan_for(x.device, x.shape[1:3])
        forecast, taps = plan.block_forward(x, want_forecast_list=self.use_fork)
        return {'forecast': forecast, 'forecast_list': taps.get("forecast_list", [])}

    def forward_with_taps(self, x):
        """forecast + per-layer intermediates (lift, x after every residual, spectral outputs, last b)
        for per-layer parity checks."""
        x = check_input(x, 2, self.input_dim, "FNOFactorized2DBlock.forward_with_taps")
        plan = self.plan_for(x.device, x.shape[1:3])
        return plan.block_forward(x, want_taps=True, want_forecast_list=self.use_fork)
