"""Dump the in-kernel clock64 timeline of ff_pipe_kernel block 0 (diagnostics; run on the GPU box)."""
import ctypes as C
import json
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
s = torch.randn(32, 64, 64, 64, device="cuda")
x = torch.randn(32, 64, 64, 64, device="cuda")
with torch.no_grad():
    plan = layer._plan(s)
    for _ in range(3):
        plan.ff_forward(0, 0, s, x)
    torch.cuda.synchronize()
    lib.ffno_debug_timeline(1, None)
    plan.ff_forward(0, 0, s, x)
    torch.cuda.synchronize()
    buf = np.zeros(1024, dtype=np.int64)
    lib.ffno_debug_timeline(0, buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(8, 16, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
names = {0: "epi_team0", 5: "epi_team1", 1: "store", 2: "g1", 3: "loader", 4: "g2"}
out = {}
for r, nm in names.items():
    out[nm] = rel[r, :8].tolist()
    print(nm)
    for n in range(8):
        print("  tile", n, rel[r, n].tolist())
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ff_timeline.json"), "w"))
