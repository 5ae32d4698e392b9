import torch, os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
g.smoke()
# backward + force/mu rollout under the sanitizer too (small shapes)
from fourierflow_b200.modules import FNOFactorized2DBlock, LpLoss
from fourierflow_b200.routines import Grid2DMarkovExperiment
torch.manual_seed(0)
m = FNOFactorized2DBlock(modes=8, width=64, n_layers=2, input_dim=5, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1).cuda().train()
x = torch.randn(2, 64, 64, 5, device="cuda"); y = torch.randn(2, 64, 64, 1, device="cuda")
LpLoss()(m(x)["forecast"].reshape(2, -1), y.reshape(2, -1)).backward()
exp = Grid2DMarkovExperiment(m, n_steps=2, append_force=True, append_mu=True).cuda().eval()
data = torch.randn(2, 64, 64, 4, device="cuda")
exp.accumulate_statistics(data, force=torch.randn(2, 64, 64, device="cuda"), mu=torch.rand(2, device="cuda"))
with torch.no_grad():
    for _ in range(3):
        out = exp({"data": data, "f": torch.randn(2, 64, 64, device="cuda"), "mu": torch.rand(2, device="cuda")})
torch.cuda.synchronize()
print("sanitizer workload done", float(out[0]))
