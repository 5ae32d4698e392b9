"""3-D structured-mesh factorized cosine operator — mirror of fourierflow/modules/factorized_cno/mesh_3d.py
(the F-FNO mesh_3d with DCT-II / DCT-III per axis and real ``[in, out, modes_a]`` weights, mesh_3d.py:33, :64-108).

``share_weight=True`` builds 4-D ``[width, width, modes, 2]`` parameters in the reference (mesh_3d.py:138-144) that its
own 3-index einsum then rejects; the same constructor call raises here."""
from __future__ import annotations

from ..factorized_fno import mesh_3d as _fno


class SpectralConv2d(_fno.SpectralConv2d):
    _transform, _weight_tail = "dct", ()


class CNOFactorizedMesh3D(_fno.FNOFactorizedMesh3D):
    """factorized_cno/mesh_3d.py:120-194."""
    _transform, _weight_tail, _layer_cls = "dct", (), SpectralConv2d

    def __init__(self, modes_x, modes_y, modes_z, width, input_dim, output_dim, n_layers, share_weight, factor,
                 ff_weight_norm, n_ff_layers, layer_norm):
        if share_weight:
            raise RuntimeError("CNOFactorizedMesh3D: share_weight=True cannot run in the reference either "
                               "(4-D shared weights against a 3-index einsum, factorized_cno/mesh_3d.py:138-144, :68-71)")
        super().__init__(modes_x, modes_y, modes_z, width, input_dim, output_dim, n_layers, share_weight, factor,
                         ff_weight_norm, n_ff_layers, layer_norm)
