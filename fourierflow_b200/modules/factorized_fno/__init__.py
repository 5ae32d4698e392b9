from .grid_2d import FNOFactorized2DBlock
from .mesh_2d import FNOFactorizedMesh2D
from .mesh_3d import FNOFactorizedMesh3D
