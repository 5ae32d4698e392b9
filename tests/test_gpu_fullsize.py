"""GPU (B200): parity of the CUDA path against the CPU oracle, run on the box, AT THE SIZES THAT ARE TIMED —
the full BASELINE configurations, not scaled-down stand-ins ("What's weak" 1a/1c):

  C2  torus_li/markov/24_layers: B=32, 64x64, 24 layers — per-layer x_l taps, last b, forecast
      and the 10-step Markov rollout of the same model (per step)
  C4  torus_kochkov 256x256, modes 64, input_dim 5, B=2: 12 layers and the 24 the config ships
  C5  Mesh3D 32^3 -> 40^3 padded, modes (12,12,8), B=8, 24 layers, unshared weights
  Mesh2D width 64 on the tcgen05 path: airfoil shape 221x51 -> 229x59, modes 32/16
  Mesh3D at the real plasticity shape 101x31x20 -> 109x39x28, modes (32,12,8), B=2

Tolerance: the north star's rtol 1e-4, measured as max|y-ref| / max|ref| per tap and per rollout step
(SURVEY.md §8c); the oracle is fp32 torch-CPU (its own distance to float64 is ~5e-7 over 24 layers).
Reference lines: grid_2d.py:154-177, mesh_2d.py:149-165, mesh_3d.py:160-176, routines/grid_2d_markov.py:263-321.
"""
import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.set_num_threads(16)          # the oracle legs run on the host cores


def M():
    import fourierflow_b200.modules as m
    return m


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def c2_model(seed=0, n_layers=24):
    torch.manual_seed(seed)
    return M().FNOFactorized2DBlock(modes=16, width=64, n_layers=n_layers, input_dim=3, share_weight=True, factor=4,
                                    ff_weight_norm=True, gain=0.1).eval()


def test_c2_full_size_per_layer_vs_oracle():
    """B=32, 64x64, 24 layers: lift, every x_l, the last b and the forecast against the oracle; the plain forward
    (the launch sequence bench.py times, fused head) against the same oracle forecast."""
    from oracle import ffno_oracle as O
    m = c2_model()
    x = torch.randn(32, 64, 64, 3, generator=torch.Generator().manual_seed(1))     # bench.py's rank-0 input
    taps = {}
    with torch.no_grad():
        ref = O.block_grid2d_forward(sd_of(m), x, modes=16, n_layers=24, taps=taps)["forecast"]
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())["forecast"]
        y2 = mc(x.cuda())["forecast"]
        y3 = mc(x.cuda())["forecast"]            # third call: CUDA-graph replay
        fc, t = mc.forward_with_taps(x.cuda())
    plan = mc.plan_for(y.device, (64, 64))
    assert plan.uses_umma and plan.graph_active
    errs = {"forecast": rel_err(y, ref), "forecast_replay": rel_err(y3, ref), "forecast_tapped": rel_err(fc, ref),
            "lift": rel_err(t["lift"], taps["lift"]), "b_last": rel_err(t["b_last"], taps["b23"])}
    for l in range(24):
        errs[f"x{l}"] = rel_err(t["x"][l], taps[f"x{l}"])
        errs[f"s{l}"] = rel_err(t["s"][l], taps[f"s{l}"])
    print("C2 full size", {k: f"{v:.1e}" for k, v in errs.items() if k in ("forecast", "x0", "x11", "x23", "s23", "b_last")},
          "worst", max(errs, key=errs.get), f"{max(errs.values()):.2e}")
    assert torch.equal(y2, y3)
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


def test_c2_full_size_10_step_rollout_vs_oracle():
    """The 10-step rollout of the 24-layer model at 64x64, B=32 (routines/grid_2d_markov.py:263-321), per step."""
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    from oracle import ffno_oracle as O
    conv = c2_model(seed=2)
    sd = sd_of(conv)
    g = torch.Generator().manual_seed(22)
    data = torch.randn(32, 64, 64, 1, generator=g) + 0.2 * torch.cumsum(torch.randn(32, 64, 64, 12, generator=g), dim=-1)
    exp = Grid2DMarkovExperiment(conv, n_steps=10).cuda().eval()
    exp.accumulate_statistics(data.cuda())
    frames = data[..., :-1].unsqueeze(-1)
    pos = O.position_features((64, 64), 0.0, 1.0, data.dtype)[None, :, :, None, :].expand(32, 64, 64, 11, 2)
    stats = O.normalizer_stats(torch.cat([frames, pos], dim=-1))
    with torch.no_grad():
        ref = O.markov_rollout(sd, data, stats, modes=16, n_layers=24, n_steps=10)
        loss, step_losses, preds, layer_list = exp({"data": data.cuda()})
        exp({"data": data.cuda()})
        loss3, _, preds3, _ = exp({"data": data.cuda()})        # graph replay
    per_step = [rel_err(preds[..., t], ref["preds"][..., t]) for t in range(10)]
    print("rollout full size per step", [f"{e:.1e}" for e in per_step], f"loss {loss.item():.6f} vs {ref['loss'].item():.6f}")
    assert max(per_step) < TOL
    assert torch.equal(preds, preds3)
    assert abs(loss.item() - ref["loss"].item()) < 1e-4 * abs(ref["loss"].item())
    assert rel_err(torch.stack(step_losses), ref["step_losses"]) < 1e-4
    assert len(layer_list) == 10 and all(entry == [] for entry in layer_list)
    assert conv.plan_for(preds.device, (64, 64)).graph_active


@pytest.mark.parametrize("n_layers", [12, 24])
def test_c4_full_size_vs_oracle(n_layers):
    """BASELINE configs[3]: Kochkov 256x256, modes 64, input_dim 5, batch 2; 12 layers and the 24 the config ships."""
    from oracle import ffno_oracle as O
    torch.manual_seed(31)
    m = M().FNOFactorized2DBlock(modes=64, width=64, n_layers=n_layers, input_dim=5, share_weight=True, factor=4,
                                 ff_weight_norm=True, gain=0.1).eval()
    x = torch.randn(2, 256, 256, 5, generator=torch.Generator().manual_seed(32))
    taps = {}
    with torch.no_grad():
        ref = O.block_grid2d_forward(sd_of(m), x, modes=64, n_layers=n_layers, taps=taps)["forecast"]
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())["forecast"]
        fc, t = mc.forward_with_taps(x.cuda())
    assert mc.plan_for(y.device, (256, 256)).uses_umma
    errs = {"forecast": rel_err(y, ref), "forecast_tapped": rel_err(fc, ref)}
    for l in range(n_layers):
        errs[f"x{l}"] = rel_err(t["x"][l], taps[f"x{l}"])
    print("C4", n_layers, "worst", max(errs, key=errs.get), f"{max(errs.values()):.2e}", f"forecast {errs['forecast']:.2e}")
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


def test_c5_full_size_vs_oracle():
    """BASELINE configs[4]: FNOFactorizedMesh3D, 32^3 cube (40^3 padded), modes (12,12,8), batch 8, 24 layers."""
    from oracle import ffno_oracle as O
    torch.manual_seed(41)
    m = M().FNOFactorizedMesh3D(modes_x=12, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=24,
                                share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False).eval()
    x = torch.randn(8, 32, 32, 32, 1, generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        ref = O.block_mesh_forward(sd_of(m), x, modes=(12, 12, 8), n_layers=24)
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())
    assert mc.plan_for(y.device, (32, 32, 32)).uses_umma
    e = rel_err(y, ref)
    print("C5 full size", f"{e:.2e}")
    assert y.shape == ref.shape and e < TOL


def test_mesh2d_width64_airfoil_shape_vs_oracle():
    """FNOFactorizedMesh2D on the tcgen05 path (width 64): the airfoil shape 221x51 (229x59 padded: odd, non-power-of-two),
    modes 32/16 (experiments/airfoil/ffno), 4 layers, unshared weights."""
    from oracle import ffno_oracle as O
    torch.manual_seed(51)
    m = M().FNOFactorizedMesh2D(modes_x=32, modes_y=16, width=64, input_dim=4, n_layers=4, share_weight=False, factor=4,
                                ff_weight_norm=True, n_ff_layers=2, layer_norm=False).eval()
    x = torch.randn(3, 221, 51, 2, generator=torch.Generator().manual_seed(52))
    with torch.no_grad():
        ref = O.block_mesh_forward(sd_of(m), x, modes=(32, 16), n_layers=4)
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())
    assert mc.plan_for(y.device, (221, 51)).uses_umma
    e = rel_err(y, ref)
    print("mesh2d w64 airfoil", f"{e:.2e}")
    assert y.shape == ref.shape and e < TOL


def test_mesh3d_real_plasticity_shape_vs_oracle():
    """The real plasticity shape (SURVEY §8d stretch): 101x31x20 -> 109 (prime) x 39 x 28, modes (32,12,8), batch 2."""
    from oracle import ffno_oracle as O
    torch.manual_seed(61)
    m = M().FNOFactorizedMesh3D(modes_x=32, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=4,
                                share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False).eval()
    x = torch.randn(2, 101, 31, 20, 1, generator=torch.Generator().manual_seed(62))
    with torch.no_grad():
        ref = O.block_mesh_forward(sd_of(m), x, modes=(32, 12, 8), n_layers=4)
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())
    assert mc.plan_for(y.device, (101, 31, 20)).uses_umma
    e = rel_err(y, ref)
    print("mesh3d plasticity shape", f"{e:.2e}")
    assert y.shape == ref.shape and e < TOL


def test_axis_beyond_the_pipelined_transform_falls_back_per_axis():
    """An axis whose inverse table does not fit the pipelined tcgen05 transform (length 264 > 256 output rows) runs that
    transform on the FP32 table kernel while the other axis, every mix and the FF stay on tcgen05."""
    from oracle import ffno_oracle as O
    torch.manual_seed(71)
    m = M().FNOFactorized2DBlock(modes=16, width=64, n_layers=2, input_dim=3, share_weight=True, factor=4,
                                 ff_weight_norm=True, gain=0.1).eval()
    x = torch.randn(2, 264, 64, 3, generator=torch.Generator().manual_seed(72))
    with torch.no_grad():
        ref = O.block_grid2d_forward(sd_of(m), x, modes=16, n_layers=2)["forecast"]
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())["forecast"]
    assert mc.plan_for(y.device, (264, 64)).uses_umma
    e = rel_err(y, ref)
    print("fallback axis", f"{e:.2e}")
    assert e < TOL


def test_host_side_contracts_on_the_device():
    """stale parameter caches, .data edits, wrong trailing dimensions, rollout length."""
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    m = c2_model(n_layers=1).cuda()
    x = torch.randn(1, 64, 64, 3, device="cuda")
    with torch.no_grad():
        y0 = m(x)["forecast"].clone()
        sd = {k: (v.clone() * 1.25 if k.endswith("weight_g") else v.clone()) for k, v in m.state_dict().items()}
        m.load_state_dict(sd, assign=True)                       # replaces every Parameter object
        y1 = m(x)["forecast"].clone()
        assert rel_err(y1, y0) > 1e-3
        m.spectral_layers[0].backcast_ff.layers[1][0].weight_g.data.mul_(2.0)    # no version bump
        m.invalidate_plans()
        assert rel_err(m(x)["forecast"], y1) > 1e-3
        lin = M().WNLinear(3, 8, wnorm=True).cuda()
        with pytest.raises(RuntimeError, match="in_features"):
            lin(torch.randn(4, 6, device="cuda"))                # 6 = 2 x 3 would silently become 8 rows
        exp = Grid2DMarkovExperiment(m, n_steps=6).cuda().eval()
        with pytest.raises(RuntimeError, match="needs 7 frames"):
            exp.predict(torch.randn(1, 64, 64, 5, device="cuda"))


def test_plans_on_two_devices_in_one_process():
    """The dynamic shared-memory opt-in is per device: a second GPU used from the same process must get its own."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run by bench.py --selftest-multidev on the scaling box)")
    m0 = c2_model(n_layers=2).to("cuda:0")
    m1 = c2_model(n_layers=2).to("cuda:1")
    x = torch.randn(2, 64, 64, 3)
    with torch.no_grad():
        y0 = m0(x.to("cuda:0"))["forecast"]
        y1 = m1(x.to("cuda:1"))["forecast"]
    assert rel_err(y1.cpu(), y0.cpu()) < 1e-6


@pytest.mark.parametrize("batch", [8, 32, 256])
def test_stage_pipelined_forward_is_bit_identical_to_one_launch_per_layer(batch, monkeypatch):
    """The stage-pipelined forward (four persistent, flag-synchronised launches per forward) runs the same tile
    arithmetic as the launch-per-stage-and-layer path: outputs must be identical bit for bit, call after call."""
    x = torch.randn(batch, 64, 64, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(81))
    monkeypatch.setenv("FFNO_B200_PERSIST", "1")
    m = c2_model().cuda()
    with torch.no_grad():
        ys = [m(x)["forecast"].clone() for _ in range(4)]          # eager, eager, captured, replayed
    plan = m.plan_for(x.device, (64, 64))
    assert plan.pipeline_unit(batch) == 2 and plan.pipeline_unit(6) == 0 and plan.pipeline_unit(2) == 0
    assert plan.last_launch_count <= 8, plan.last_launch_count       # lift + 4 stage kernels (+ probes), not 97
    monkeypatch.setenv("FFNO_B200_PERSIST", "0")
    m2 = c2_model().cuda()
    with torch.no_grad():
        y2 = m2(x)["forecast"]
    plan2 = m2.plan_for(x.device, (64, 64))
    assert plan2.pipeline_unit(batch) == 0 and plan2.last_launch_count > 90
    for y in ys:
        assert torch.equal(y, y2)
    assert torch.isfinite(y2).all()


def test_stage_pipelined_rollout_matches_the_launch_per_layer_rollout(monkeypatch):
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    g = torch.Generator(device="cuda").manual_seed(82)
    data = torch.randn(8, 64, 64, 1, device="cuda", generator=g) + 0.2 * torch.cumsum(
        torch.randn(8, 64, 64, 8, device="cuda", generator=g), dim=-1)
    outs = []
    for persist in ("1", "0"):
        monkeypatch.setenv("FFNO_B200_PERSIST", persist)
        exp = Grid2DMarkovExperiment(c2_model(n_layers=6).cuda(), n_steps=5).cuda().eval()
        exp.accumulate_statistics(data)
        with torch.no_grad():
            for _ in range(3):
                loss, _, preds, _ = exp({"data": data})
        outs.append((loss.item(), preds.clone(), exp.conv.plan_for(data.device, (64, 64)).pipeline_unit(8)))
    assert outs[0][2] == 2 and outs[1][2] == 0
    assert torch.equal(outs[0][1], outs[1][1]) and outs[0][0] == outs[1][0]


# ---- sibling operators at the shapes of their shipped configs -----------------------------------------

def test_cno_plasticity_shape_vs_oracle():
    """experiments/plasticity/fcno: CNOFactorizedMesh3D on the real mesh 101x31x20 -> 109x39x28, coefficient counts
    (32, 12, 8) -> pairs (16, 6, 4) on the tcgen05 path, batch 2, 4 layers (factorized_cno/mesh_3d.py:160-176)."""
    from oracle import ffno_oracle as O
    torch.manual_seed(81)
    m = M().CNOFactorizedMesh3D(modes_x=32, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=4,
                                share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False).eval()
    x = torch.randn(2, 101, 31, 20, 1, generator=torch.Generator().manual_seed(82))
    with torch.no_grad():
        ref = O.block_mesh_forward(sd_of(m), x, modes=(32, 12, 8), n_layers=4)
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())
    assert mc.plan_for(y.device, (101, 31, 20)).uses_umma
    e = rel_err(y, ref)
    print("cno plasticity shape", f"{e:.2e}")
    assert y.shape == ref.shape and e < TOL


def test_cno_odd_coefficient_counts_on_tcgen05_vs_oracle():
    """An odd number of kept DCT coefficients leaves the Im row of the last pair empty (zero table column, zero weight
    block): 64x48 grid, 15 coefficients, 24 layers, batch 4 (experiments/torus_kochkov/fcno shape family)."""
    from oracle import ffno_oracle as O
    torch.manual_seed(83)
    m = M().CNOFactorized2DBlock(modes=15, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                                 ff_weight_norm=True, gain=0.1).eval()
    x = torch.randn(4, 64, 48, 3, generator=torch.Generator().manual_seed(84))
    taps = {}
    with torch.no_grad():
        ref = O.block_grid2d_forward(sd_of(m), x, modes=15, n_layers=24, taps=taps)["forecast"]
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())["forecast"]
        _, t = mc.forward_with_taps(x.cuda())
    assert mc.plan_for(y.device, (64, 48)).uses_umma
    errs = {"forecast": rel_err(y, ref)}
    for l in (0, 11, 23):
        errs[f"x{l}"] = rel_err(t["x"][l], taps[f"x{l}"])
    print("cno 15 coefficients", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


def test_fnoplus_c2_shape_vs_oracle():
    """experiments/torus_li/ablation/no_factorization: FNOPlus2DBlock at 64x64, modes 16, batch 8, 4 layers, unshared
    weights (zongyi_fno/grid_plus_2d.py:143-161): FP32 rfft2 passes + tcgen05 FeedForward."""
    from oracle import ffno_oracle as O
    torch.manual_seed(85)
    m = M().FNOPlus2DBlock(modes=16, width=64, n_layers=4, input_dim=3, share_weight=False, factor=4,
                           ff_weight_norm=True, gain=0.1).eval()
    x = torch.randn(8, 64, 64, 3, generator=torch.Generator().manual_seed(86))
    taps = {}
    with torch.no_grad():
        ref = O.block_grid2d_forward(sd_of(m), x, modes=16, n_layers=4, taps=taps)["forecast"]
    mc = m.cuda()
    with torch.no_grad():
        y = mc(x.cuda())["forecast"]
        _, t = mc.forward_with_taps(x.cuda())
    errs = {"forecast": rel_err(y, ref), "s0": rel_err(t["s"][0], taps["s0"]), "x3": rel_err(t["x"][3], taps["x3"])}
    print("fnoplus 64x64", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad


def test_geo_interior_elasticity_shape_vs_oracle():
    """experiments/elasticity/ffno: the interior of FNOFactorizedPointCloud2D on the 64x64 latent grid, width 64, modes 16,
    8 layers (7 interior), batch 4 (factorized_fno/point_cloud_2d.py:198-210)."""
    from oracle import ffno_oracle as O
    torch.manual_seed(87)
    m = M().FNOFactorizedPointCloud2D(modes1=16, modes2=16, width=64, in_channels=2, out_channels=1, n_layers=8,
                                      s1=64, s2=64).eval()
    g = torch.Generator().manual_seed(88)
    uc = torch.randn(4, 64, 64, 64, generator=g)
    bias = 0.1 * torch.randn(64, 64, 64, generator=g)
    with torch.no_grad():
        ref = O.geo_interior_forward(sd_of(m), uc, bias, modes=16, n_layers=8)
    mc = m.cuda()
    with torch.no_grad():
        y = mc.interior_forward(uc.cuda(), bias.cuda())
    e = rel_err(y, ref)
    print("geo interior 64x64", f"{e:.2e}")
    assert e < 1e-5
