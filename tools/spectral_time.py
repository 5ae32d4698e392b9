"""Time the spectral operator of one C2 layer (forward transforms, mode mix, inverse transforms: 3 launches) alone,
cold L2, CUDA events — the quantity bench.py reports as roofline_spectral."""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from fourierflow_b200.modules import FNOFactorized2DBlock
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=1, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1).cuda().eval()
layer = m.spectral_layers[0]
xs = torch.randn(32, 64, 64, 64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    lplan = layer._plan(xs)
    bufs = [torch.empty_like(xs), torch.empty_like(xs)]
    for _ in range(5): lplan.spectral_split_forward(0, xs, bufs)
    torch.cuda.synchronize()
    tot = 0
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lplan.spectral_split_forward(0, xs, bufs); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
print("spectral operator (3 launches, cold L2) us:", round(tot / 20 * 1000, 1))
