"""Training-step time of the C2 model (24 layers, 64 x 64, batch 32; routines/grid_2d_markov.py:172-193 loss, AdamW): this
backend (forward = tcgen05 kernels, backward = ffno_block_bwd FP32 adjoints) next to the reference modules
(oracle/_ref, eager PyTorch autograd, allow_tf32=False) on the same GPU, plus the gradient agreement between the two.
Diagnostics for the backward row; bench.py stays the forward headline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fourierflow_b200.modules import FNOFactorized2DBlock, LpLoss  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B = int(os.environ.get("TRAIN_BATCH", "32"))
torch.manual_seed(0)
m = FNOFactorized2DBlock(**bench.C2).cuda().train()
x = torch.randn(B, 64, 64, 3, device="cuda")
y = torch.randn(B, 64, 64, 1, device="cuda")
l2 = LpLoss(size_average=True)


def timed(step, n=10, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


opt = torch.optim.AdamW(m.parameters(), lr=1e-4)


def ours(update=True):
    opt.zero_grad()
    loss = l2(m(x)["forecast"].reshape(B, -1), y.reshape(B, -1))
    loss.backward()
    if update:
        opt.step()
    return loss


def fwd_only():
    with torch.no_grad():
        return m(x)["forecast"]


out = {"batch": B, "ours_fwd_ms": timed(fwd_only), "ours_train_step_ms": timed(ours),
       "ours_fwd_bwd_ms": timed(lambda: ours(update=False))}
out["ours_backward_launches"] = m.plan_for(x.device, (64, 64)).last_launch_count
ref = bench.reference_model({k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
if ref is not None:
    ref = ref.cuda().train()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-4)

    def theirs(update=True):
        ropt.zero_grad()
        f = ref(x)["forecast"]
        loss = (torch.linalg.vector_norm((f - y).reshape(B, -1), dim=1) / torch.linalg.vector_norm(y.reshape(B, -1), dim=1)).mean()
        loss.backward()
        if update:
            ropt.step()
        return loss

    # gradient agreement on identical weights (before any update diverges them)
    ref.load_state_dict({k: v.detach().clone() for k, v in m.state_dict().items()})
    lo, lr_ = ours(update=False), theirs(update=False)
    worst, worst_name = 0.0, ""
    rp = dict(ref.named_parameters())
    for k, p in m.named_parameters():
        g, r = p.grad.double(), rp[k].grad.double()
        e = ((g - r).abs().max() / r.abs().max().clamp(min=1e-30)).item()
        if e > worst:
            worst, worst_name = e, f"{k} (max|g| = {r.abs().max().item():.2e})"
    out["loss_ours"], out["loss_reference"], out["worst_grad_rel_err_vs_reference"] = lo.item(), lr_.item(), worst
    out["worst_grad_tensor"] = worst_name
    out["reference_eager_train_step_ms"] = timed(theirs)
    with torch.no_grad():
        out["reference_eager_fwd_ms"] = timed(lambda: ref(x)["forecast"])
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_step_time.json"), "w"), indent=1)
