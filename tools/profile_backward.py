"""Two C2 training forwards + backwards (24 layers, batch 32) for an ncu launch list of the backward pass."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fourierflow_b200.modules import FNOFactorized2DBlock, LpLoss  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("TRAIN_BATCH", "32"))
m = FNOFactorized2DBlock(**bench.C2).cuda().train()
x = torch.randn(B, 64, 64, 3, device="cuda")
y = torch.randn(B, 64, 64, 1, device="cuda")
for _ in range(2):
    m.zero_grad()
    LpLoss()(m(x)["forecast"].reshape(B, -1), y.reshape(B, -1)).backward()
torch.cuda.synchronize()
print("done")
