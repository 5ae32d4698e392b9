"""Un-factorized 2-D Fourier block with the F-FNO feed-forward / residual structure — mirror of
fourierflow/modules/zongyi_fno/grid_plus_2d.py (the ``no_factorization`` ablation of torus_li).

The spectral layer (grid_plus_2d.py:52-83) is ``rfft2`` -> the two ``n_modes x n_modes`` corner blocks (rows ``:n_modes``
with ``fourier_weight[0]``, rows ``-n_modes:`` with ``fourier_weight[1]``, both ``[in, out, n_modes, n_modes, 2]``) ->
``irfft2``.  On the device it is the same plan with ``transform='rfft2'`` (FFNO_TRANSFORM_RFFT2): the column transform
is the F-FNO table kernel, the row transform a complex one, the mix runs over ``2 * n_modes**2`` modes; FP32 kernels.
Everything around the spectral layer (lift, FeedForward, residual, head, fork options) is the F-FNO block's.
"""
from __future__ import annotations

from ... import _ops
from ..factorized_fno import grid_2d as _fno


class SpectralConv2d(_fno.SpectralConv2d):
    _transform = "rfft2"

    def __init__(self, in_dim, out_dim, n_modes, forecast_ff, backcast_ff, fourier_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode):
        if mode == 'low-pass':
            raise ValueError("FNOPlus2DBlock has no low-pass mode (a bare `raise` at zongyi_fno/grid_plus_2d.py:74-75)")
        object.__setattr__(self, "_weight_tail", (n_modes, 2))       # [in, out, n_modes, n_modes, 2]
        super().__init__(in_dim, out_dim, n_modes, forecast_ff, backcast_ff, fourier_weight, factor,
                         ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode)

    def layer_spec(self) -> _ops.LayerSpec:
        # index 0 = low row block, 1 = high row block (not per-axis weights: no reordering)
        return _ops.LayerSpec([self.fourier_weight[0], self.fourier_weight[1]], self.backcast_ff,
                              self.forecast_ff if self.use_fork else None)


class FNOPlus2DBlock(_fno.FNOFactorized2DBlock):
    """zongyi_fno/grid_plus_2d.py:86-161 (the constructor of FNOFactorized2DBlock with 5-D spectral weights)."""
    _transform, _layer_cls = "rfft2", SpectralConv2d

    def __init__(self, modes, width, input_dim=12, dropout=0.0, in_dropout=0.0, n_layers=4,
                 share_weight: bool = False, share_fork=False, factor=2, ff_weight_norm=False, n_ff_layers=2,
                 gain=1, layer_norm=False, use_fork=False, mode='full'):
        if mode == 'low-pass':
            raise ValueError("FNOPlus2DBlock has no low-pass mode (zongyi_fno/grid_plus_2d.py:74-75)")
        object.__setattr__(self, "_weight_tail", (modes, 2))
        super().__init__(modes, width, input_dim, dropout, in_dropout, n_layers, share_weight, share_fork, factor,
                         ff_weight_norm, n_ff_layers, gain, layer_norm, use_fork, mode)
