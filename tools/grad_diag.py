import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from fourierflow_b200.modules import FNOFactorized2DBlock, LpLoss
from oracle import ffno_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
B = 4
torch.manual_seed(0)
m = FNOFactorized2DBlock(**bench.C2).cuda().train()
x = torch.randn(B, 64, 64, 3, device="cuda"); y = torch.randn(B, 64, 64, 1, device="cuda")
l2 = LpLoss(size_average=True)
loss = l2(m(x)["forecast"].reshape(B, -1), y.reshape(B, -1)); loss.backward()
ref = bench.reference_model({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}).cuda().train()
f = ref(x)["forecast"]
lr = (torch.linalg.vector_norm((f - y).reshape(B, -1), dim=1) / torch.linalg.vector_norm(y.reshape(B, -1), dim=1)).mean(); lr.backward()
# float64 oracle
sdk = m.state_dict(keep_vars=True)
leaves = {}
for k, v in sdk.items():
    leaves.setdefault(id(v), v.detach().cpu().double().clone().requires_grad_(True))
p = {k: leaves[id(v)] for k, v in sdk.items()}
out = O.block_grid2d_forward(p, x.cpu().double(), modes=16, n_layers=24)
lo = O.lp_loss_rel(out["forecast"].reshape(B, -1), y.cpu().double().reshape(B, -1)); lo.backward()
print("loss", loss.item(), lr.item(), lo.item())
rp = dict(ref.named_parameters())
rows = []
for k, v in m.named_parameters():
    t = p[k].grad
    sc = t.abs().max().clamp(min=1e-300)
    e_ours = ((v.grad.double().cpu() - t).abs().max() / sc).item()
    e_ref = ((rp[k].grad.double().cpu() - t).abs().max() / sc).item()
    rows.append((e_ours, e_ref, k, sc.item()))
rows.sort(reverse=True)
for r in rows[:12]: print("%.2e ours  %.2e ref32  %-60s max|g|=%.2e" % r)
print("median ours %.2e ref %.2e" % (sorted(r[0] for r in rows)[len(rows)//2], sorted(r[1] for r in rows)[len(rows)//2]))
