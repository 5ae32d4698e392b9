"""In-kernel timelines (FFNO_TIMELINE build) of the LAST layer's kernels inside a full 24-layer C2 forward."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FFNO_B200_GRAPH"] = "0"
from fourierflow_b200 import _lib  # noqa: E402
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
x = torch.randn(32, 64, 64, 3, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    lib.ffno_debug_timeline(1, None)
    m(x)
    torch.cuda.synchronize()
    buf = np.zeros(1024, dtype=np.int64)
    lib.ffno_debug_timeline(0, buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(8, 16, 8)
for grp, roles in (("FF", {0: "epi1", 1: "store", 2: "mma", 3: "loader"}), ("AXIS", {5: "epi", 6: "mma", 7: "loader"})):
    sel = t[list(roles)]
    t0 = sel[sel > 0].min() if (sel > 0).any() else 0
    print("==", grp)
    for r, nm in roles.items():
        print(nm)
        for n in range(8):
            print("   ", n, [int(v - t0) if v > 0 else -1 for v in t[r][n]])
