"""In-kernel clock64 timelines (FFNO_TIMELINE build: FFNO_TIMELINE=1 python -m fourierflow_b200.build --force, then
FFNO_B200_LIB=fourierflow_b200/lib/libffno_b200_timeline.so) of block 0 of each pipeline kernel inside a 4-layer C2
forward at batch 32: per role and tile, the cycle at which each pipeline event happened.  One forward per kernel
(g_timeline_on selects the kernel that records: 1 FF, 2 forward transforms, 3 inverse transforms, 4 mode mix)."""
import ctypes as C
import json
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
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=4, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
x = torch.randn(int(os.environ.get("PROFILE_BATCH", "32")), 64, 64, 3, device="cuda")
ROLES = {1: {0: "chunk-epi team0", 5: "chunk-epi team1", 1: "store", 2: "G1 issue", 4: "G2 issue", 3: "loader"},
         2: {5: "epilogue", 6: "mma", 7: "converter"}, 3: {5: "epilogue", 6: "mma", 7: "converter"},
         4: {5: "epilogue team0", 6: "epilogue team1", 4: "mma", 7: "loader"}}
NAMES = {1: "ff_ts_kernel", 2: "axis_pipe_kernel forward", 3: "axis_pipe_kernel inverse", 4: "mix_pipe_kernel"}
out = {}
with torch.no_grad():
    for _ in range(3):
        m(x)
    for sel in (2, 4, 3, 1):
        torch.cuda.synchronize()
        lib.ffno_debug_timeline(sel, None)
        m(x)
        torch.cuda.synchronize()
        buf = np.zeros(1024, dtype=np.int64)
        lib.ffno_debug_timeline(0, buf.ctypes.data_as(C.c_void_p))
        t = buf.reshape(8, 16, 8)
        t0 = t[t > 0].min() if (t > 0).any() else 0
        print("==", NAMES[sel], "(last layer's launch, block 0; cycles since its first stamp)")
        out[NAMES[sel]] = {}
        for r, nm in ROLES[sel].items():
            rows = [[int(v - t0) if v > 0 else -1 for v in t[r][n]] for n in range(16) if (t[r][n] > 0).any()]
            out[NAMES[sel]][nm] = rows
            print(" ", nm)
            for n, row in enumerate(rows):
                print("    ", n, [v for v in row if v >= 0])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_timelines.json"), "w"))
