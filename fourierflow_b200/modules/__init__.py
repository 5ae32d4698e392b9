"""Drop-in mirrors of the F-FNO operator modules exported by fourierflow/modules/__init__.py:1-11
(the subset on the hot path: SURVEY.md §8(a))."""
from .factorized_cno import CNOFactorized2DBlock, CNOFactorizedMesh2D, CNOFactorizedMesh3D
from .factorized_fno import FNOFactorized2DBlock, FNOFactorizedMesh2D, FNOFactorizedMesh3D, FNOFactorizedPointCloud2D
from .feedforward import FeedForward
from .iphi import IPhi
from .linear import WNLinear
from .loss import LpLoss
from .normalizer import Normalizer
from .zongyi_fno import FNOPlus2DBlock
