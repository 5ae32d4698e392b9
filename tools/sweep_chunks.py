"""Time the C2 forward (24 layers, batch 32, 64x64) under several batch-chunking configurations of the plan
(FFNO_B200_CHUNK = samples per chunk, FFNO_B200_STREAMS = chunks running side by side) and check that every
configuration produces the same forecast as the unchunked run.  Diagnostics only; bench.py is the headline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

CONFIGS = [c.split(":") for c in (sys.argv[1:] or ["0:1", "16:1", "8:1", "0:2", "8:2", "4:2", "0:3", "0:4"])]
B = int(os.environ.get("SWEEP_BATCH", "32"))
x = torch.randn(B, 64, 64, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
out = []
for chunk, streams in CONFIGS:
    os.environ["FFNO_B200_CHUNK"] = chunk
    os.environ["FFNO_B200_STREAMS"] = streams
    torch.manual_seed(0)
    m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                             ff_weight_norm=True, gain=0.1).cuda().eval()
    with torch.no_grad():
        for _ in range(4):
            y = m(x)["forecast"]
        torch.cuda.synchronize()
        tot = 0.0
        n = 20
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = m(x)["forecast"]
            b.record()
            b.synchronize()
            tot += a.elapsed_time(b)
    if ref is None:
        ref = y.clone()
    err = float((y - ref).abs().max() / ref.abs().max())
    rec = {"chunk": int(chunk), "streams": int(streams), "ms": tot / n, "samples_per_s": B / (tot / n * 1e-3),
           "max_rel_diff_vs_first": err}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del m
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_chunks.json"), "w"), indent=1)
