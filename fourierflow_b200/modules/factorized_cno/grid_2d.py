"""2-D periodic-grid factorized cosine operator — mirror of fourierflow/modules/factorized_cno/grid_2d.py.

The reference file is the F-FNO one with three changes (grid_2d.py:27, :58-69, :73-86): the per-axis transform is
the ortho DCT-II (modules/dct.py:16-45) with the DCT-III as its inverse, ``fourier_weight`` is REAL
``[in, out, modes]``, and there is no ``mode`` switch.  On the device that is the same plan with
``transform='dct'`` (ffno_desc.transform = FFNO_TRANSFORM_DCT): cosine tables instead of the rfft tables and a
block-diagonal mix, every kernel unchanged — so the classes here only change the weight shape and the plan flag.
"""
from __future__ import annotations

from ..factorized_fno import grid_2d as _fno


class SpectralConv2d(_fno.SpectralConv2d):
    _transform, _weight_tail = "dct", ()

    def __init__(self, in_dim, out_dim, n_modes, forecast_ff, backcast_ff, fourier_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode='full'):
        if mode != 'full':
            raise ValueError("the factorized cosine operator has no low-pass / no-fourier mode (factorized_cno/grid_2d.py:44-49)")
        super().__init__(in_dim, out_dim, n_modes, forecast_ff, backcast_ff, fourier_weight, factor,
                         ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, 'full')


class CNOFactorized2DBlock(_fno.FNOFactorized2DBlock):
    """factorized_cno/grid_2d.py:98-172 (same constructor as FNOFactorized2DBlock; ``mode`` is accepted and unused there)."""
    _transform, _weight_tail, _layer_cls = "dct", (), SpectralConv2d

    def __init__(self, modes, width, input_dim=12, dropout=0.0, in_dropout=0.0, n_layers=4,
                 share_weight: bool = False, share_fork=False, factor=2, ff_weight_norm=False, n_ff_layers=2,
                 gain=1, layer_norm=False, use_fork=False, mode='full'):
        super().__init__(modes, width, input_dim, dropout, in_dropout, n_layers, share_weight, share_fork, factor,
                         ff_weight_norm, n_ff_layers, gain, layer_norm, use_fork, 'full')
