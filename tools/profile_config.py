"""A few forwards of another BASELINE configuration for ncu captures: C4 (Kochkov 256x256, modes 64, batch 2, 12 layers)
or C5 (Mesh3D 32^3 -> 40^3, modes (12,12,8), batch 8, 4 layers — the per-layer launches repeat).  PROFILE_CONFIG=c4|c5."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FFNO_B200_GRAPH"] = "0"
import fourierflow_b200.modules as M  # noqa: E402

cfg = os.environ.get("PROFILE_CONFIG", "c4")
torch.manual_seed(0)
if cfg == "c4":
    m = M.FNOFactorized2DBlock(modes=64, width=64, n_layers=4, input_dim=5, share_weight=True, factor=4,
                               ff_weight_norm=True, gain=0.1).cuda().eval()
    x = torch.randn(2, 256, 256, 5, device="cuda")
    run = lambda: m(x)["forecast"]
else:
    m = M.FNOFactorizedMesh3D(modes_x=12, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=4,
                              share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False).cuda().eval()
    x = torch.randn(8, 32, 32, 32, 1, device="cuda")
    run = lambda: m(x)
with torch.no_grad():
    for _ in range(int(os.environ.get("PROFILE_ITERS", "2"))):
        y = run()
torch.cuda.synchronize()
print("done", cfg, float(y.abs().max()))
