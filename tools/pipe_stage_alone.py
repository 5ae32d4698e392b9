"""Raw throughput of ONE stage of the stage-pipelined C2 forward (24 layers, batch 32) running alone with its
dependencies pre-satisfied (FFNO_B200_PIPE_ONLY; outputs are garbage, only the time means something), for several
CTA counts.  Tells what each stage costs per tile when nothing starves it.  Usage: python tools/pipe_stage_alone.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fourierflow_b200.modules import FNOFactorized2DBlock  # noqa: E402

TILES = {0: 24 * 2048, 1: 24 * 512, 2: 24 * 2048, 3: 24 * 1024}
NAMES = {0: "fwd", 1: "mix", 2: "inv", 3: "ff"}
os.environ["FFNO_B200_GRAPH"] = "0"
os.environ["FFNO_B200_PERSIST"] = "1"
x = torch.randn(32, 64, 64, 3, device="cuda")
for stage, counts in ((0, (16, 26, 36, 52)), (2, (16, 26, 36, 52)), (3, (24, 40, 52, 64)), (1, (32,))):
    for n in counts:
        sms = [2, 32, 2, 1]
        sms[stage] = n
        os.environ["FFNO_B200_PIPE_SMS"] = ",".join(map(str, sms))
        os.environ["FFNO_B200_PIPE_ONLY"] = str(stage)
        torch.manual_seed(0)
        m = FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                                 ff_weight_norm=True, gain=0.1).cuda().eval()
        with torch.no_grad():
            for _ in range(2):
                m(x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m(x)
            b.record()
            b.synchronize()
        ms = a.elapsed_time(b)
        ctas = n if stage != 1 else 32
        print(json.dumps({"stage": NAMES[stage], "ctas": ctas, "ms": round(ms, 3),
                          "us_per_tile_per_cta": round(ms * 1e3 * ctas / TILES[stage], 3)}), flush=True)
        del m
