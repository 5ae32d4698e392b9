"""Informational: the reference's op sequence (torch.fft.rfft / einsum / F.linear — the oracle port of
fourierflow.modules) run EAGERLY on the B200 through torch's own CUDA kernels (cuFFT / cuBLAS), C2 shape.
This is the "reference PyTorch-eager forward on 1xB200" denominator of the north star (>= 10x), measured
with CUDA events; it is not part of the product or of bench.py.  It lives under tests/ because it executes oracle/
(test infrastructure): only tests/, smoke() and bench.py's CPU baseline may do that.  Not a pytest module."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ffno_oracle as O  # noqa: E402
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                         ff_weight_norm=True, gain=0.1).eval()
sd = {k: v.detach().cuda() for k, v in m.state_dict().items()}
x = torch.randn(32, 64, 64, 3, device="cuda")
with torch.no_grad():
    for _ in range(5):
        O.block_grid2d_forward(sd, x, modes=16, n_layers=24)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 20
    for _ in range(n):
        O.block_grid2d_forward(sd, x, modes=16, n_layers=24)
    b.record()
    b.synchronize()
ms = a.elapsed_time(b) / n
res = {"torch_eager_oracle_ms_per_forward": ms, "samples_per_s": 32 / (ms * 1e-3), "batch": 32}
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "eager_gpu_baseline.json"), "w"))
