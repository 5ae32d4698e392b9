"""Run a few C2-shape layer operators (for ncu captures): spectral + FF through the C ABI."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=2, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
layer = m.spectral_layers[0]
B = int(os.environ.get("PROFILE_BATCH", "32"))
x = torch.randn(B, 64, 64, 64, device="cuda")
with torch.no_grad():
    plan = layer._plan(x)
    for _ in range(int(os.environ.get("PROFILE_ITERS", "3"))):
        s = plan.spectral_forward(0, x)
        y = plan.ff_forward(0, 0, s, x)
torch.cuda.synchronize()
print("done", float(y.abs().max()))
