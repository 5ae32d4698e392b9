"""Backward row: ffno_block_bwd / ffno_rel_l2_bwd through the C ABI and torch.autograd.Function,
against (i) the gradients the EXECUTED reference computed for its one-step training loss
(routines/grid_2d_markov.py:172-193; fixtures tests/golden/grad_*.npz made by oracle/make_golden.py) and (ii)
torch.autograd through the CPU oracle at the C2 layer shape.  Tolerance: max|g - ref| / max|ref| <= 1e-4 per tensor
(FP32 kernels, atomics over the point reduction; measured ~1e-6)."""
import numpy as np
import pytest
import torch

from golden_util import load, rel_err
from oracle import ffno_oracle as O


def build(cls, kw, sd):
    import fourierflow_b200.modules as M
    m = getattr(M, cls)(**kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda()

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _loss(m, x, y):
    from fourierflow_b200.modules import LpLoss
    B = x.shape[0]
    forecast = m(x)["forecast"]
    return forecast, LpLoss(size_average=True)(forecast.reshape(B, -1), y.reshape(B, -1))


@pytest.mark.parametrize("name", ["grad_c2arch_16", "grad_unshared_w32", "grad_cno_grid2d_w64", "grad_plus2d_w32",
                                  "grad_plus2d_w64"])
@pytest.mark.parametrize("path", ["auto", "generic"])
def test_gradients_match_the_executed_reference(name, path, monkeypatch):
    """Input gradient + every parameter gradient (weight-norm g / v, shared spectral weights summed over layers, plain
    linears, biases) of the reference's training loss."""
    monkeypatch.setenv("FFNO_B200_PATH", path)
    kw, sd, a = load(name)
    cls = "CNOFactorized2DBlock" if "cno" in name else "FNOPlus2DBlock" if "plus2d" in name else "FNOFactorized2DBlock"
    m = build(cls, kw, sd).train()                                    # cno: DCT sibling, plus2d: un-factorized sibling
    x = a["x"].cuda().requires_grad_(True)
    forecast, loss = _loss(m, x, a["y"].cuda())
    loss.backward()
    assert rel_err(forecast, a["forecast"]) < TOL
    assert abs(loss.item() - a["loss"].item()) < 1e-5 * abs(a["loss"].item())
    e = rel_err(x.grad, a["grad::x"])
    print(name, path, f"dx {e:.2e}")
    assert e < TOL
    params = dict(m.named_parameters())
    checked, worst = 0, 0.0
    for k, ref in a.items():
        if not k.startswith("grad::") or k == "grad::x":
            continue
        g = params[k[6:]].grad
        assert g is not None, k
        e = rel_err(g, ref)
        worst = max(worst, e)
        assert e < TOL, (k, e)
        checked += 1
    print(name, path, f"{checked} parameter gradients, worst {worst:.2e}")
    assert checked >= 8


def test_gradients_accumulate_and_match_oracle_autograd_at_the_c2_layer_shape():
    """64 x 64, width 64, modes 16, 2 layers, batch 2 (the C2 layer shape): CUDA backward vs torch.autograd through the
    oracle; a second backward accumulates into .grad like torch does."""
    torch.manual_seed(3)
    from fourierflow_b200.modules import FNOFactorized2DBlock
    m = FNOFactorized2DBlock(modes=16, width=64, n_layers=2, input_dim=3, share_weight=True, factor=4,
                             ff_weight_norm=True, gain=0.1).cuda().train()
    x = torch.randn(2, 64, 64, 3, device="cuda")
    y = torch.randn(2, 64, 64, 1, device="cuda")
    _, loss = _loss(m, x, y)
    loss.backward()
    g1 = {id(p): p.grad.clone() for p in m.parameters()}
    _, loss = _loss(m, x, y)
    loss.backward()
    for k, p in m.named_parameters():
        assert rel_err(p.grad, 2 * g1[id(p)]) < 1e-5, k
    # oracle autograd on the CPU
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    leaves = {}
    for k, v in m.state_dict(keep_vars=True).items():
        leaves.setdefault(id(v), sd[k].clone().requires_grad_(True))
    p = {k: leaves[id(v)] for k, v in m.state_dict(keep_vars=True).items()}
    out = O.block_grid2d_forward(p, x.cpu(), modes=16, n_layers=2)
    lo = O.lp_loss_rel(out["forecast"].reshape(2, -1), y.cpu().reshape(2, -1))
    lo.backward()
    assert abs(lo.item() - loss.item()) < 1e-5 * abs(lo.item())
    for k, v in m.state_dict(keep_vars=True).items():
        if v.requires_grad:
            assert rel_err(g1[id(v)], p[k].grad) < TOL, k


def test_rel_l2_backward_matches_autograd():
    from fourierflow_b200 import _ops
    x = torch.randn(5, 777, device="cuda", requires_grad=True)
    y = torch.randn(5, 777, device="cuda")
    w = torch.rand(5, device="cuda")
    (_ops.rel_l2_differentiable(x, y) * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    ((torch.linalg.vector_norm(xr - y, dim=1) / torch.linalg.vector_norm(y, dim=1)) * w).sum().backward()
    assert rel_err(x.grad, xr.grad) < 1e-5


def test_one_optimizer_step_reduces_the_training_loss():
    """routines/base.py:27-52 applies the optimizer to these gradients: an SGD step along them must lower the loss."""
    torch.manual_seed(0)
    from fourierflow_b200.modules import FNOFactorized2DBlock
    m = FNOFactorized2DBlock(modes=8, width=64, n_layers=3, input_dim=3, share_weight=True, factor=4,
                             ff_weight_norm=True, gain=0.1).cuda().train()
    x = torch.randn(4, 32, 32, 3, device="cuda")
    y = torch.randn(4, 32, 32, 1, device="cuda")
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        _, loss = _loss(m, x, y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", losses)
    assert losses[-1] < losses[0]


def test_routine_training_step_matches_the_reference_loop():
    """Grid2DMarkovExperiment.training_step (grid_2d_markov.py:374-390, routines/base.py:27-52): epoch 0 accumulates the
    normaliser, later epochs take an optimizer step; the loss of the step equals the oracle's
    features -> normalise -> stack -> de-normalise -> LpLoss on the same batch, and the loss goes down."""
    torch.manual_seed(1)
    from fourierflow_b200.modules import FNOFactorized2DBlock
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    conv = FNOFactorized2DBlock(modes=8, width=64, n_layers=2, input_dim=3, share_weight=True, factor=4,
                                ff_weight_norm=True, gain=0.1)
    exp = Grid2DMarkovExperiment(conv, n_steps=2).cuda().train()
    B, X = 4, 32
    batch = {"x": torch.randn(B, X, X, 1, device="cuda"), "y": torch.randn(B, X, X, 1, device="cuda")}
    assert exp.training_step(batch, current_epoch=0) is None and exp.normalizer.count.item() == B * X * X
    exp.normalizer.eval()                                   # statistics frozen (max_accumulations reached in a real run)
    # oracle loss on the same features
    mean, std = exp.normalizer.mean.cpu(), exp.normalizer.std.cpu()
    pos = O.position_features((X, X), 0.0, 1.0, torch.float32).unsqueeze(0).expand(B, X, X, 2)
    feats = (torch.cat([batch["x"].cpu(), pos], dim=-1) - mean) / std
    sd = {k: v.detach().cpu() for k, v in conv.state_dict().items()}
    im = O.block_grid2d_forward(sd, feats, modes=8, n_layers=2)["forecast"] * std[0] + mean[0]
    ref_loss = O.lp_loss_rel(im.reshape(B, -1), batch["y"].cpu().reshape(B, -1)).item()
    opt = torch.optim.AdamW(exp.parameters(), lr=2.5e-3, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    losses = [exp.training_step(batch, batch_idx=i, optimizer=opt, scheduler=sch, current_epoch=1).item() for i in range(5)]
    print("routine losses", losses, "oracle first", ref_loss)
    assert abs(losses[0] - ref_loss) < 1e-4 * abs(ref_loss)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("mode", ["default", "fast", "fp32"])
def test_full_depth_c2_gradients_vs_fp64_oracle(mode, monkeypatch):
    """The benchmark's own model (24 layers, 64 x 64, modes 16, shared weights, weight-norm): every parameter gradient and
    the input gradient against torch.autograd through the oracle in FLOAT64.
    default (FP32 forward recompute, spectral adjoint on the tcgen05 kernels) and fp32 (everything FP32): every tensor
    within 1e-4 of the largest gradient of its kind and the whole gradient within 1e-4 in relative L2; fp32 additionally
    within 1e-4 of each tensor's OWN max|ref| (measured 1e-6: the level of the reference's own FP32 autograd).
    fast (FFNO_B200_BWD=fast, recompute on the tcgen05 kernels): activations carry the forward's ~1e-5 noise, hidden units
    that close to the ReLU kink flip their mask — same bounds here (few flips at this initialisation), see
    test_fast_mode_mask_flips for a case where they show."""
    if mode != "default":
        monkeypatch.setenv("FFNO_B200_BWD", mode)
    torch.manual_seed(0)
    from fourierflow_b200.modules import FNOFactorized2DBlock
    kw = dict(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1)
    m = FNOFactorized2DBlock(**kw).cuda().train()
    B = 2
    x = torch.randn(B, 64, 64, 3, device="cuda", requires_grad=True)
    y = torch.randn(B, 64, 64, 1, device="cuda")
    _, loss = _loss(m, x, y)
    loss.backward()
    sdk = m.state_dict(keep_vars=True)
    leaves = {}
    for k, v in sdk.items():
        leaves.setdefault(id(v), v.detach().cpu().double().clone().requires_grad_(True))
    p = {k: leaves[id(v)] for k, v in sdk.items()}
    xd = x.detach().cpu().double().requires_grad_(True)
    out = O.block_grid2d_forward(p, xd, modes=16, n_layers=24)
    lo = O.lp_loss_rel(out["forecast"].reshape(B, -1), y.cpu().double().reshape(B, -1))
    lo.backward()
    assert abs(lo.item() - loss.item()) < 1e-5 * abs(lo.item())
    assert rel_err(x.grad, xd.grad) < TOL
    kind_scale = {}
    for k, v in m.named_parameters():
        kind = k.split(".")[-1]
        kind_scale[kind] = max(kind_scale.get(kind, 0.0), p[k].grad.abs().max().item())
    worst_own, worst_kind, num, den = 0.0, 0.0, 0.0, 0.0
    for k, v in m.named_parameters():
        ref = p[k].grad
        d = (v.grad.double().cpu() - ref)
        worst_own = max(worst_own, (d.abs().max() / ref.abs().max()).item())
        worst_kind = max(worst_kind, (d.abs().max() / kind_scale[k.split(".")[-1]]).item())
        num += (d ** 2).sum().item()
        den += (ref ** 2).sum().item()
    l2 = (num / den) ** 0.5
    print(f"full depth [{mode}]: worst vs own max {worst_own:.2e}, vs kind max {worst_kind:.2e}, whole-gradient L2 {l2:.2e}")
    assert l2 < TOL and worst_kind < TOL
    if mode == "fp32":
        assert worst_own < TOL


@pytest.mark.parametrize("name,cls", [("grad_mesh3d_w64", "FNOFactorizedMesh3D"), ("grad_mesh2d_w32", "FNOFactorizedMesh2D"),
                                      ("grad_cno_mesh2d_w32", "CNOFactorizedMesh2D")])
def test_mesh_gradients_match_the_executed_reference(name, cls):
    """Mesh variants (plasticity / airfoil training, routines/structured_mesh.py): the linspace grid append, the zero
    padding after the lift and the crop before the head are part of the backward (mesh_3d.py:160-176, mesh_2d.py:149-165).
    3-D width 64 runs the tcgen05 kernels (3 axes), 2-D width 32 the FP32 ones."""
    kw, sd, a = load(name)
    m = build(cls, kw, sd).train()
    x = a["x"].cuda().requires_grad_(True)
    from fourierflow_b200.modules import LpLoss
    B = x.shape[0]
    out = m(x)
    loss = LpLoss(size_average=True)(out.reshape(B, -1), a["y"].cuda().reshape(B, -1))
    loss.backward()
    assert rel_err(out, a["out"]) < TOL
    assert abs(loss.item() - a["loss"].item()) < 1e-5 * abs(a["loss"].item())
    e = rel_err(x.grad, a["grad::x"])
    params = dict(m.named_parameters())
    checked, worst = 0, 0.0
    for k, ref in a.items():
        if not k.startswith("grad::") or k == "grad::x":
            continue
        g = params[k[6:]].grad
        assert g is not None, k
        worst = max(worst, rel_err(g, ref))
        assert rel_err(g, ref) < TOL, (k, rel_err(g, ref))
        checked += 1
    print(name, f"dx {e:.2e}, {checked} parameter gradients, worst {worst:.2e}")
    assert e < TOL and checked >= 10


def test_fast_mode_mask_flips(monkeypatch):
    """FFNO_B200_BWD=fast on the perturbed-weight 3-D fixture (8 k points): the gradient is that of the tcgen05-computed
    forward — hidden units within ~1e-5 of the ReLU kink take the other branch, each a discrete change of single terms.
    The whole gradient stays within 2e-2 of the reference's in relative L2 (measured ~5e-3); the default mode holds 1e-4
    per tensor on the same fixture (test_mesh_gradients_match_the_executed_reference)."""
    monkeypatch.setenv("FFNO_B200_BWD", "fast")
    kw, sd, a = load("grad_mesh3d_w64")
    m = build("FNOFactorizedMesh3D", kw, sd).train()
    x = a["x"].cuda().requires_grad_(True)
    from fourierflow_b200.modules import LpLoss
    B = x.shape[0]
    LpLoss(size_average=True)(m(x).reshape(B, -1), a["y"].cuda().reshape(B, -1)).backward()
    params = dict(m.named_parameters())
    num = den = 0.0
    for k, ref in a.items():
        if k.startswith("grad::") and k != "grad::x":
            d = params[k[6:]].grad.double().cpu() - ref.double()
            num += float((d * d).sum())
            den += float((ref.double() ** 2).sum())
    l2 = (num / den) ** 0.5
    print(f"fast mode, mesh3d fixture: whole-gradient relative L2 {l2:.2e}, dx {rel_err(x.grad, a['grad::x']):.2e}")
    assert l2 < 2e-2


def test_structured_mesh_training_step():
    """StructuredMeshExperiment.training_step (structured_mesh.py:22-32) on a plasticity-like 3-D stack: losses go down."""
    torch.manual_seed(2)
    from fourierflow_b200.modules import FNOFactorizedMesh3D
    from fourierflow_b200.routines import StructuredMeshExperiment
    model = FNOFactorizedMesh3D(modes_x=6, modes_y=6, modes_z=4, width=64, input_dim=4, output_dim=4, n_layers=2,
                                share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False)
    exp = StructuredMeshExperiment(model, loss_scale=2.0).cuda().train()
    batch = {"x": torch.randn(2, 12, 10, 8, 1, device="cuda"), "y": torch.randn(2, 12, 10, 8, 4, device="cuda")}
    opt = torch.optim.AdamW(exp.parameters(), lr=2e-3)
    losses = [exp.training_step(batch, i, optimizer=opt).item() for i in range(5)]
    print("mesh losses", losses)
    assert losses[-1] < losses[0]
    with torch.no_grad():
        assert abs(exp.validation_step(batch).item() * 2.0 - losses[-1]) < 0.05 * losses[-1]


def test_geo_ffno_gradients_match_the_executed_reference():
    """geo-F-FNO training (experiments/elasticity/ffno): (i) ffno_layers_bwd alone — gradients of <r, interior(uc, bias)>
    w.r.t. the latent grid, the grid bias and the interior parameters; (ii) the whole training loss through the torch end
    layers + the interior autograd node, against the reference's autograd (fixture grad_geo_pointcloud)."""
    from golden_util import load_geo
    import fourierflow_b200.modules as M
    kw, sd, a = load_geo("grad_geo_pointcloud")
    m = M.FNOFactorizedPointCloud2D(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    uc, bias = a["uc_in"].cuda().requires_grad_(True), a["grid_bias"].cuda().requires_grad_(True)
    out = m.interior_forward(uc, bias)
    (out * a["r"].cuda()).sum().backward()
    errs = {"uc_out": rel_err(out, a["uc_out"]), "d_uc": rel_err(uc.grad, a["igrad::uc"]), "d_bias": rel_err(bias.grad, a["igrad::bias"])}
    params = dict(m.named_parameters())
    n_int = 0
    for k, ref in a.items():
        if k.startswith("igrad::convs."):
            errs[k] = rel_err(params[k[7:]].grad, ref)
            n_int += 1
    print("geo interior", {k: f"{v:.1e}" for k, v in list(errs.items())[:3]}, f"{n_int} parameter gradients, worst {max(errs.values()):.2e}")
    assert n_int >= 18 and max(errs.values()) < TOL, errs
    m.zero_grad()
    u = a["u"].cuda()
    B = u.shape[0]
    outp = m(u)
    loss = M.LpLoss(size_average=True)(outp.reshape(B, -1), a["y"].cuda().reshape(B, -1))
    loss.backward()
    assert rel_err(outp, a["out"]) < TOL and abs(loss.item() - a["loss"].item()) < 1e-5 * abs(a["loss"].item())
    worst, n = 0.0, 0
    for k, ref in a.items():
        if k.startswith("grad::"):
            g = params[k[6:]].grad
            assert g is not None, k
            g = torch.view_as_real(g) if g.is_complex() else g
            e = rel_err(g, ref)
            worst, n = max(worst, e), n + 1
            assert e < TOL, (k, e)
    print(f"geo training loss: {n} parameter gradients, worst {worst:.2e}")
    assert n >= 25
