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
# sibling operators: DCT tables / pair mix on both kernel families, rfft2 passes + tcgen05 FF, geo interior
from fourierflow_b200.modules import CNOFactorized2DBlock, CNOFactorizedMesh3D, FNOPlus2DBlock, FNOFactorizedPointCloud2D
with torch.no_grad():
    for width in (32, 64):
        c = CNOFactorized2DBlock(modes=7, width=width, n_layers=2, input_dim=3, share_weight=True, factor=4,
                                 ff_weight_norm=True).cuda().eval()
        c(torch.randn(2, 24, 20, 3, device="cuda"))
        p2 = FNOPlus2DBlock(modes=5, width=width, n_layers=2, input_dim=3, share_weight=False, factor=4,
                            ff_weight_norm=True).cuda().eval()
        p2(torch.randn(2, 12, 18, 3, device="cuda"))
    c3 = CNOFactorizedMesh3D(5, 4, 3, 64, 4, 4, 2, False, 4, True, 2, False).cuda().eval()
    c3(torch.randn(1, 9, 7, 6, 1, device="cuda"))
    geo = FNOFactorizedPointCloud2D(modes1=4, modes2=4, width=32, in_channels=2, out_channels=1, n_layers=3, s1=12, s2=10).cuda().eval()
    geo(torch.rand(2, 30, 2, device="cuda"))
# ... and the sibling backward passes (DCT pair layout, rfft2 adjoint with c2c_expand / two-tensor weight gradient)
for mod in (CNOFactorized2DBlock(modes=5, width=64, n_layers=2, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True),
            FNOPlus2DBlock(modes=3, width=32, n_layers=2, input_dim=3, share_weight=False, factor=4, ff_weight_norm=True)):
    mod = mod.cuda().train()
    xs = torch.randn(2, 12, 10, 3, device="cuda", requires_grad=True)
    LpLoss()(mod(xs)["forecast"].reshape(2, -1), torch.randn(2, 120, device="cuda")).backward()
torch.cuda.synchronize()
print("sanitizer workload done", float(out[0]))
