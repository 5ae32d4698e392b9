"""Dump the in-kernel clock64 timeline of axis_pipe_kernel block (0,0) for the last spectral op launched."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200 import _lib  # noqa: E402
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=1, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
layer = m.spectral_layers[0]
x = torch.randn(32, 64, 64, 64, device="cuda")
with torch.no_grad():
    plan = layer._plan(x)
    for _ in range(3):
        plan.spectral_forward(0, x)
    torch.cuda.synchronize()
    lib.ffno_debug_timeline(1, None)
    m2 = FNOFactorized2DBlock(modes=16, width=64, n_layers=1, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                              gain=0.1).cuda().eval()
    xin = torch.randn(32, 64, 64, 3, device="cuda")
    os.environ["FFNO_B200_GRAPH"] = "0"
    m2(xin)
    torch.cuda.synchronize()
    lib.ffno_debug_timeline(1, None)
    m2(xin)                              # layer path: fwd (both axes), mix, inv (both axes): the inverse launch stamps last
    torch.cuda.synchronize()
    buf = np.zeros(1024, dtype=np.int64)
    lib.ffno_debug_timeline(0, buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(8, 16, 8)
for r, nm in {5: "epi", 6: "mma", 7: "loader"}.items():
    tt = t[r]
    t0 = tt[tt > 0].min() if (tt > 0).any() else 0
    print(nm)
    for n in range(8):
        print("  ", n, [int(v - t0) if v > 0 else -1 for v in tt[n][:4]])
