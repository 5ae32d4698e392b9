"""Forward time of the other BASELINE configurations (informational; bench.py's headline is configs[1]):
C3 on one GPU (batch 256), C4 (Kochkov 256x256, modes 64, 12 and 24 layers, batch 2), C5 (Mesh3D 32^3 -> 40^3, batch 8)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock, FNOFactorizedMesh3D  # noqa: E402


def timed(fn, n=10, warm=4):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


out = {}
torch.manual_seed(0)
with torch.no_grad():
    m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                             ff_weight_norm=True, gain=0.1).cuda().eval()
    for B in (32, 256):
        x = torch.randn(B, 64, 64, 3, device="cuda")
        ms = timed(lambda: m(x))
        out[f"C2/C3 24 layers 64x64 batch {B}"] = {"ms": ms, "samples_per_s": B / ms * 1e3,
                                                 "tcgen05": bool(m.plan_for(x.device, (64, 64)).uses_umma)}
    del m
    for L in (12, 24):
        m = FNOFactorized2DBlock(modes=64, width=64, n_layers=L, input_dim=5, share_weight=True, factor=4,
                                 ff_weight_norm=True, gain=0.1).cuda().eval()
        x = torch.randn(2, 256, 256, 5, device="cuda")
        ms = timed(lambda: m(x))
        out[f"C4 {L} layers 256x256 modes 64 batch 2"] = {"ms": ms, "samples_per_s": 2 / ms * 1e3,
                                                         "tcgen05": bool(m.plan_for(x.device, (256, 256)).uses_umma)}
        del m
    m = FNOFactorizedMesh3D(modes_x=12, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=24,
                            share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False).cuda().eval()
    x = torch.randn(8, 32, 32, 32, 1, device="cuda")
    ms = timed(lambda: m(x), n=5, warm=3)
    out["C5 Mesh3D 24 layers 32^3 (40^3 padded) batch 8"] = {"ms": ms, "samples_per_s": 8 / ms * 1e3}
# 10-step rollouts through the routine mirror (one C-ABI call each): torus_li (3 features) and torus_kochkov
# (use_velocity: 5 features, stream-function velocities recomputed every step)
from fourierflow_b200.routines import Grid2DMarkovExperiment  # noqa: E402
with torch.no_grad():
    for name, kw, B, G, vel in (("torus_li 24 layers 64x64 batch 32", dict(modes=16, input_dim=3), 32, 64, False),
                                ("torus_kochkov 24 layers 256x256 batch 2 (use_velocity)", dict(modes=64, input_dim=5), 2, 256, True)):
        conv = FNOFactorized2DBlock(width=64, n_layers=24, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1, **kw)
        exp = Grid2DMarkovExperiment(conv, n_steps=10, use_velocity=vel).cuda().eval()
        data = torch.randn(B, G, G, 12, device="cuda")
        exp.accumulate_statistics(data)
        ms = timed(lambda: exp.predict(data), n=5, warm=3)
        out[f"rollout 10 steps: {name}"] = {"ms": ms, "ms_per_step": ms / 10, "sample_steps_per_s": B * 10 / ms * 1e3}
for k, v in out.items():
    print(k, json.dumps(v))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_times.json"), "w"), indent=1)
