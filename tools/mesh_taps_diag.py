"""Per-layer taps of the 3-D mesh stack: tcgen05 path against the FP32 path of this library (diagnostics)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import load, rel_err
import fourierflow_b200.modules as M
kw, sd, a = load("grad_mesh3d_w64")
res = {}
for path in ("generic", "auto"):
    os.environ["FFNO_B200_PATH"] = path
    m = M.FNOFactorizedMesh3D(**kw); m.load_state_dict(sd, strict=True); m = m.cuda().eval()
    x = a["x"].cuda()
    with torch.no_grad():
        plan = m.plan_for(x.device, x.shape[1:4])
        out, taps = plan.block_forward(x, want_taps=True)
        out2, _ = plan.block_forward(x)
    res[path] = (out, taps, out2, plan.uses_umma)
g, u = res["generic"], res["auto"]
print("uses_umma", u[3], "out(taps) vs generic", rel_err(u[0], g[0]), "out(no taps) vs generic", rel_err(u[2], g[0]))
print("lift", rel_err(u[1]["lift"], g[1]["lift"]), "b_last", rel_err(u[1]["b_last"], g[1]["b_last"]))
for l in range(kw["n_layers"]):
    print(l, "s", rel_err(u[1]["s"][l], g[1]["s"][l]), "x", rel_err(u[1]["x"][l], g[1]["x"][l]))
