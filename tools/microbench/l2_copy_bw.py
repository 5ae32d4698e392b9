"""Copy bandwidth (read + write bytes) of torch's elementwise copy kernel for working sets from L2-resident to
DRAM-sized: tells whether a kernel that streams at the measured HBM copy rate would go faster out of the 126 MB L2."""
import json
import torch

out = []
for mb in (4, 8, 16, 32, 48, 64, 128, 512, 2048):
    n = mb * (1 << 20) // 4
    a = torch.randn(n, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    torch.cuda.synchronize()
    reps = max(10, min(400, 8192 // mb))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        b.copy_(a)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / reps
    out.append({"buffer_MB": mb, "working_set_MB": 2 * mb, "us": round(ms * 1e3, 2), "GBps": round(2 * mb * (1 << 20) / (ms * 1e-3) / 1e9, 1)})
    print(json.dumps(out[-1]), flush=True)
