"""2-D structured-mesh factorized cosine operator — mirror of fourierflow/modules/factorized_cno/mesh_2d.py
(the F-FNO mesh_2d with DCT-II / DCT-III per axis and real ``[in, out, modes_a]`` weights, mesh_2d.py:32, :63-91)."""
from __future__ import annotations

from ..factorized_fno import mesh_2d as _fno


class SpectralConv2d(_fno.SpectralConv2d):
    _transform, _weight_tail = "dct", ()

    def __init__(self, in_dim, out_dim, modes_x, modes_y, forecast_ff, backcast_ff, fourier_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, mode='full'):
        super().__init__(in_dim, out_dim, modes_x, modes_y, forecast_ff, backcast_ff, fourier_weight, factor,
                         ff_weight_norm, n_ff_layers, layer_norm, use_fork, dropout, 'full')


class CNOFactorizedMesh2D(_fno.FNOFactorizedMesh2D):
    """factorized_cno/mesh_2d.py:103-170.  ``share_weight=True`` builds 4-D ``[width, width, modes, 2]`` parameters in
    the reference (mesh_2d.py:118-124) that its own 3-index einsum (:69-72) rejects; the same call raises here."""
    _transform, _weight_tail, _layer_cls = "dct", (), SpectralConv2d

    def __init__(self, modes_x, modes_y, width, input_dim, n_layers, share_weight, factor, ff_weight_norm,
                 n_ff_layers, layer_norm):
        if share_weight:
            raise RuntimeError("CNOFactorizedMesh2D: share_weight=True cannot run in the reference either "
                               "(4-D shared weights against a 3-index einsum, factorized_cno/mesh_2d.py:118-124, :69-72)")
        super().__init__(modes_x, modes_y, width, input_dim, n_layers, share_weight, factor, ff_weight_norm,
                         n_ff_layers, layer_norm)
