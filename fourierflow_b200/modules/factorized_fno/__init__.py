from .grid_2d import FNOFactorized2DBlock
from .mesh_2d import FNOFactorizedMesh2D
from .mesh_3d import FNOFactorizedMesh3D
from .point_cloud_2d import FNOFactorizedPointCloud2D
