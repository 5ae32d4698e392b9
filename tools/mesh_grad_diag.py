import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from golden_util import load, rel_err
import fourierflow_b200.modules as M
from fourierflow_b200.modules import LpLoss
for name, cls in [("grad_mesh3d_w64", "FNOFactorizedMesh3D"), ("grad_mesh2d_w32", "FNOFactorizedMesh2D")]:
    kw, sd, a = load(name)
    m = getattr(M, cls)(**kw); m.load_state_dict(sd, strict=True); m = m.cuda().train()
    x = a["x"].cuda().requires_grad_(True)
    B = x.shape[0]
    out = m(x)
    loss = LpLoss(size_average=True)(out.reshape(B, -1), a["y"].cuda().reshape(B, -1)); loss.backward()
    print(name, os.environ.get("FFNO_B200_BWD"), "out", rel_err(out, a["out"]), "dx", rel_err(x.grad, a["grad::x"]))
    params = dict(m.named_parameters())
    rows = sorted(((rel_err(params[k[6:]].grad, ref), k, ref.abs().max().item()) for k, ref in a.items() if k.startswith("grad::") and k != "grad::x"), reverse=True)
    for r in rows[:6]: print("   %.2e %-55s max|g| %.2e" % r)
