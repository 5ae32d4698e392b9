from .grid_2d import CNOFactorized2DBlock
from .mesh_2d import CNOFactorizedMesh2D
from .mesh_3d import CNOFactorizedMesh3D
