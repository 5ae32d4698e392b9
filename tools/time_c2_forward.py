"""Time of the C2 forward (24 layers, 64 x 64, batch 32) with whatever library FFNO_B200_LIB names — no parity gate:
for knock-out / experiment builds whose results are wrong by construction (tools/experiments/README.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
x = torch.randn(32, 64, 64, 3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(5):
        m(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get("STEPS", "30"))):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m(x)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
ts.sort()
print(os.environ.get("FFNO_B200_LIB", "product"), os.environ.get("FFNO_B200_AXIS_TS", ""), f"median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}")
