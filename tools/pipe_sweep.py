"""Time the C2 forward (24 layers, batch 32, 64x64, L2 flushed between timed forwards) on the stage-pipelined path for
several SM partitions (FFNO_B200_PIPE_SMS = CTAs of the forward-transform, mix, inverse-transform and FF stage) and on
the launch-per-layer path (FFNO_B200_PERSIST=0, with and without CUDA-graph replay).  Diagnostics; bench.py is the headline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

B = int(os.environ.get("SWEEP_BATCH", "32"))
CONFIGS = sys.argv[1:] or ["auto", "30,32,30,56", "32,32,32,52", "34,32,34,48", "36,32,36,44", "38,32,38,40", "28,32,36,52",
                           "36,32,28,52", "persist0", "persist0_nograph"]
x = torch.randn(B, 64, 64, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref, out = None, []
for cfg in CONFIGS:
    for k in ("FFNO_B200_PIPE_SMS", "FFNO_B200_PERSIST", "FFNO_B200_GRAPH"):
        os.environ.pop(k, None)
    if cfg.startswith("persist0"):
        os.environ["FFNO_B200_PERSIST"] = "0"
        if cfg.endswith("nograph"):
            os.environ["FFNO_B200_GRAPH"] = "0"
    else:
        os.environ["FFNO_B200_PERSIST"] = "1"
        if cfg != "auto":
            os.environ["FFNO_B200_PIPE_SMS"] = cfg
    torch.manual_seed(0)
    m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                             ff_weight_norm=True, gain=0.1).cuda().eval()
    with torch.no_grad():
        for _ in range(4):
            y = m(x)["forecast"]
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = m(x)["forecast"]
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
    plan = m.plan_for(x.device, (64, 64))
    if ref is None:
        ref = y.clone()
    ts.sort()
    rec = {"config": cfg, "ms_median": ts[len(ts) // 2], "ms_min": ts[0], "unit": plan.pipeline_unit(B),
           "graph": plan.graph_active, "launches": plan.last_launch_count, "identical": bool(torch.equal(y, ref))}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del m
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pipe_sweep.json"), "w"), indent=1)
