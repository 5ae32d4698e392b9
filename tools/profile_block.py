"""A short C2 forward (4 layers, batch 32, eager launches) for ncu captures of the pipeline kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FFNO_B200_GRAPH"] = "0"
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=4, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
x = torch.randn(int(os.environ.get("PROFILE_BATCH", "32")), 64, 64, 3, device="cuda")
with torch.no_grad():
    for _ in range(int(os.environ.get("PROFILE_ITERS", "2"))):
        y = m(x)["forecast"]
torch.cuda.synchronize()
print("done", float(y.abs().max()))
