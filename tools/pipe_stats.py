"""Per-stage wait statistics of the stage-pipelined C2 forward (FFNO_B200_PIPE_DEBUG): which stage starves, which one
is the bottleneck, and when each stage's CTAs start and stop.  Usage: python tools/pipe_stats.py [f,m,i,ff]"""
import ctypes as C
import json
import os
import sys

os.environ["FFNO_B200_PIPE_DEBUG"] = "1"
os.environ["FFNO_B200_GRAPH"] = "0"
os.environ["FFNO_B200_PERSIST"] = "1"
if len(sys.argv) > 1:
    os.environ["FFNO_B200_PIPE_SMS"] = sys.argv[1]
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

B = int(os.environ.get("SWEEP_BATCH", "32"))
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True,
                         gain=0.1).cuda().eval()
x = torch.randn(B, 64, 64, 3, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m(x)
    b.record()
    b.synchronize()
plan = m.plan_for(x.device, (64, 64))
buf = (C.c_uint64 * (8 + 4 * 256 + 256))()
n = plan.lib.ffno_debug_pipe_stats(plan._plan, buf, len(buf))
assert n > 0, n
g = [int(buf[i]) for i in range(4)]
off, t_first = 8, None
rows = []
for name, cnt in zip(("fwd", "mix", "inv", "ff"), g):
    recs = [[int(buf[off + 4 * i + j]) for j in range(4)] for i in range(cnt)]
    off += 4 * cnt
    rows.append((name, recs))
    starts = [r[2] for r in recs if r[2]]
    if starts:
        t_first = min(starts) if t_first is None else min(t_first, min(starts))
out = {"ms": a.elapsed_time(b), "stages": {}}
for name, recs in rows:
    recs = [r for r in recs if r[1]]
    blocked = sum(r[0] for r in recs) / max(1, sum(r[1] for r in recs))
    out["stages"][name] = {"ctas": len(recs), "blocked_frac": round(blocked, 3),
                           "kernel_us": round(sum(r[1] for r in recs) / max(1, len(recs)) / 1965.0, 1),
                           "start_us": round((min(r[2] for r in recs) - t_first) / 1e3, 1),
                           "end_us": round((max(r[3] for r in recs) - t_first) / 1e3, 1)}
print(json.dumps(out))
ts0 = 8 + 4 * 256
base = min(int(buf[ts0 + i]) for i in range(256) if buf[ts0 + i])
for si, name in enumerate(("fwd", "mix", "inv", "ff")):
    for layer in range(3):
        print(name, "layer", layer, " ".join(f"{(int(buf[ts0 + 64 * si + 16 * layer + u]) - base) / 1e3:6.1f}" for u in range(16)))
