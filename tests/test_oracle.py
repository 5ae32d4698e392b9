"""CPU: pin the oracle restatement (oracle/ffno_oracle.py) to the executed reference's outputs."""
import numpy as np
import pytest
import torch

from golden_util import load, load_geo, rel_err
from oracle import ffno_oracle as O

TOL = 2e-6   # same torch CPU kernels in (almost) the same order; reference fp32 self-noise ~5e-7


def _grid2d(name):
    kw, sd, a = load(name)
    taps = {}
    out = O.block_grid2d_forward(sd, a["x"], modes=kw["modes"], n_layers=kw["n_layers"],
                                 n_ff_layers=kw.get("n_ff_layers", 2),
                                 layer_norm=kw.get("layer_norm", False),
                                 use_fork=kw.get("use_fork", False), mode=kw.get("mode", "full"),
                                 taps=taps)
    return kw, a, out, taps


@pytest.mark.parametrize("name", ["grid2d_c2arch_32", "grid2d_c2_24layers_32", "grid2d_gain1_unshared",
                                  "grid2d_ln_w32", "grid2d_fork", "grid2d_lowpass",
                                  "grid2d_nofourier", "grid2d_nyquist", "cno_grid2d_w32", "cno_grid2d_w64",
                                  "plus2d_w32", "plus2d_w64", "plus2d_shared_fork"])
def test_grid2d_block(name):
    kw, a, out, taps = _grid2d(name)
    assert rel_err(out["forecast"], a["forecast"]) < TOL
    for i, f in enumerate(out["forecast_list"]):
        assert rel_err(f, a[f"forecast_list{i}"]) < TOL
    if "tap_lift" in a:
        assert rel_err(taps["lift"], a["tap_lift"]) < TOL
        for l in range(kw["n_layers"]):
            assert rel_err(taps[f"x{l}"], a[f"tap_x{l}"]) < TOL
        assert rel_err(taps[f"b{kw['n_layers'] - 1}"], a["tap_b_last"]) < TOL
    if "tap_s0" in a:
        assert rel_err(taps["s0"], a["tap_s0"]) < TOL


@pytest.mark.parametrize("name", ["mesh2d_small", "mesh3d_small", "mesh3d_w64", "cno_mesh2d_small", "cno_mesh3d_w64"])
def test_mesh_block(name):
    kw, sd, a = load(name)
    modes = [kw["modes_x"], kw["modes_y"]] + ([kw["modes_z"]] if "modes_z" in kw else [])
    out = O.block_mesh_forward(sd, a["x"], modes=modes, n_layers=kw["n_layers"],
                               n_ff_layers=kw["n_ff_layers"], layer_norm=kw["layer_norm"])
    assert out.shape == a["out"].shape
    assert rel_err(out, a["out"]) < TOL


@pytest.mark.parametrize("name", ["geo_pointcloud_w32", "geo_pointcloud_shared"])
def test_geo_interior(name):
    """Interior layers of the geo-F-FNO (point_cloud_2d.py:198-210) against the executed reference's latent grids."""
    kw, sd, a = load_geo(name)
    uc = O.geo_interior_forward(sd, a["uc_in"], a["grid_bias"], modes=kw["modes1"], n_layers=kw["n_layers"])
    assert rel_err(uc, a["uc_out"]) < TOL


def test_spectral_layer_c2_shape():
    kw, sd, a = load("spectral_c2_layer")
    s = O.forward_fourier_grid2d(a["x"], sd["fourier_weight.0"], sd["fourier_weight.1"], kw["K"])
    assert rel_err(s, a["s"]) < TOL


def test_rollout():
    kw, sd, a = load("rollout_c2arch_16")
    stats = {"sum": a["norm_sum"], "sum_squared": a["norm_sum_squared"], "count": a["norm_count"]}
    mean, std = O.normalizer_mean_std(stats)
    assert torch.allclose(mean, a["norm_mean"]) and torch.allclose(std, a["norm_std"])
    r = O.markov_rollout(sd, a["data"], stats, modes=kw["modes"], n_layers=kw["n_layers"],
                         n_steps=kw["n_steps"])
    assert rel_err(r["preds"], a["preds"]) < 1e-5
    assert abs(r["loss"].item() - a["loss"].item()) < 1e-5 * abs(a["loss"].item())
    assert rel_err(r["step_losses"], a["step_losses"]) < 1e-5


@pytest.mark.parametrize("name", ["rollout_nopos_16", "rollout_shuffle_16", "rollout_difference_16"])
def test_rollout_ablation_switches(name):
    """use_position=False / shuffle_grid / learn_difference (routines/grid_2d_markov.py:286-318, five shipped ablation
    configs) against fixtures driven through the reference's own conv, Normalizer and LpLoss."""
    kw, sd, a = load(name)
    stats = {"sum": a["norm_sum"], "sum_squared": a["norm_sum_squared"], "count": a["norm_count"]}
    r = O.markov_rollout(sd, a["data"], stats, modes=kw["modes"], n_layers=kw["n_layers"], n_steps=kw["n_steps"],
                         use_position=kw["use_position"], learn_difference=kw["learn_difference"],
                         shuffle=(a["x_idx"], a["y_idx"]) if kw["shuffle_grid"] else None)
    assert rel_err(r["preds"], a["preds"]) < 1e-5
    assert rel_err(r["step_losses"], a["step_losses"]) < 1e-5


@pytest.mark.parametrize("name", ["rollout_force_mu_16", "rollout_force_static_16"])
def test_rollout_force_mu(name):
    """append_force (static / time-varying forcing) and append_mu feature sets of the torus_vis configs
    (routines/grid_2d_markov.py:246-260, :288-291) against the executed reference."""
    kw, sd, a = load(name)
    stats = {"sum": a["norm_sum"], "sum_squared": a["norm_sum_squared"], "count": a["norm_count"]}
    r = O.markov_rollout(sd, a["data"], stats, modes=kw["modes"], n_layers=kw["n_layers"], n_steps=kw["n_steps"],
                         force=a["force"], mu=a.get("mu"))
    assert rel_err(r["preds"], a["preds"]) < 1e-5
    assert rel_err(r["step_losses"], a["step_losses"]) < 1e-5


def test_normalizer_stats_restatement():
    x = torch.randn(7, 5, 4, 3)
    st = O.normalizer_stats(x)
    mean, std = O.normalizer_mean_std(st)
    flat = x.reshape(-1, 3)
    assert torch.allclose(mean, flat.mean(0), atol=1e-6)
    assert torch.allclose(std, flat.std(0, unbiased=False), atol=1e-5)
    empty = {"sum": torch.zeros(3), "sum_squared": torch.zeros(3), "count": torch.tensor(0.0)}
    assert torch.equal(O.normalizer_mean_std(empty)[1], torch.full((3,), 1e-8))


@pytest.mark.parametrize("L,K", [(64, 16), (40, 12), (109, 32), (8, 5), (9, 5), (256, 64), (39, 12)])
def test_dft_matrices_match_torch_fft(L, K):
    """The explicit truncated real-DFT matrices (what the CUDA library builds) reproduce
    rfft(ortho)[:K] and irfft(zero-padded, n=L, ortho), including the C2R traps."""
    g = torch.Generator().manual_seed(L * 1000 + K)
    x = torch.randn(L, 3, dtype=torch.float64, generator=g)
    D = torch.from_numpy(O.dft_forward_matrix(L, K))
    F = torch.fft.rfft(x, dim=0, norm="ortho")[:K]
    got = D @ x
    assert torch.allclose(got[0::2], F.real, atol=1e-12) and torch.allclose(got[1::2], F.imag, atol=1e-12)
    R = torch.randn(K, 3, dtype=torch.complex128, generator=g)       # complex DC / Nyquist on purpose
    pad = torch.zeros(L // 2 + 1, 3, dtype=torch.complex128)
    pad[:K] = R
    y = torch.fft.irfft(pad, n=L, dim=0, norm="ortho")
    E = torch.from_numpy(O.dft_inverse_matrix(L, K))
    r = torch.empty(2 * K, 3, dtype=torch.float64)
    r[0::2], r[1::2] = R.real, R.imag
    assert torch.allclose(E @ r, y, atol=1e-12)


@pytest.mark.parametrize("shape", [(64, 64), (48, 40), (33, 17)])
def test_velocity_features_restatement(shape):
    """The use_velocity features (routines/grid_2d_markov.py:82-93, :206-220): for a stream function psi and
    w = -laplace(psi) the oracle must return q = psi_y, v = -psi_x (its kx/ky/lap buffers restate jax-cfd's
    rfft_mesh, which cannot be executed here: this analytic identity is what pins them)."""
    import math
    X, Y = shape
    a, b = 3, 2
    x = torch.arange(X, dtype=torch.float64) * 2 * math.pi / X
    y = torch.arange(Y, dtype=torch.float64) * 2 * math.pi / Y
    xx, yy = torch.meshgrid(x, y, indexing="ij")
    psi = torch.sin(a * xx) * torch.cos(b * yy) + 0.5 * torch.cos(xx + 2 * yy)
    w = (a * a + b * b) * torch.sin(a * xx) * torch.cos(b * yy) + 0.5 * 5 * torch.cos(xx + 2 * yy)
    q, v = O.velocity_features(w[None, :, :, None])
    psi_y = -b * torch.sin(a * xx) * torch.sin(b * yy) - 0.5 * 2 * torch.sin(xx + 2 * yy)
    psi_x = a * torch.cos(a * xx) * torch.cos(b * yy) - 0.5 * torch.sin(xx + 2 * yy)
    assert (q[0, ..., 0] - psi_y).abs().max() < 1e-12
    assert (v[0, ..., 0] + psi_x).abs().max() < 1e-12
    # a constant vorticity has no velocity (lap[0, 0] = 1 only avoids the division by zero), fp32 path included
    q32, v32 = O.velocity_features(torch.ones(2, X, Y, 1))
    assert q32.abs().max() < 1e-6 and v32.abs().max() < 1e-6
    # domain scaling: on a box twice as long the same index-space field has velocities twice as large
    q2, _ = O.velocity_features(w[None, :, :, None], domain=((0, 4 * math.pi), (0, 4 * math.pi)))
    assert (q2 - 2 * q).abs().max() < 1e-10


def test_three_pass_bf16_split_error_budget():
    """The numerics claim behind the tcgen05 path: a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with
    a_hi = bf16(a), a_lo = bf16(a - a_hi) and FP32 accumulation leaves errors of order 2^-16 |a||b| — two orders
    below the rtol 1e-4 budget — while a single BF16 pass does not fit in it.  Emulated here on the FeedForward of
    the C2 layer (64 -> 256 -> 64, ReLU) against float64."""
    kw, sd, a = load("grid2d_c2arch_32")
    p = "spectral_layers.0.backcast_ff."
    w1 = O.weight_norm_fold(sd[p + "layers.0.0.weight_g"], sd[p + "layers.0.0.weight_v"])
    w2 = O.weight_norm_fold(sd[p + "layers.1.0.weight_g"], sd[p + "layers.1.0.weight_v"])
    b1, b2 = sd[p + "layers.0.0.bias"], sd[p + "layers.1.0.bias"]
    x = torch.randn(4096, 64, generator=torch.Generator().manual_seed(0))

    def split(t):
        hi = t.to(torch.bfloat16).float()
        return hi, (t - hi).to(torch.bfloat16).float()

    def mm3(u, w):                       # u @ w.T in three BF16 x BF16 -> FP32 passes
        uh, ul = split(u)
        wh, wl = split(w)
        return uh @ wh.T + uh @ wl.T + ul @ wh.T

    ref = torch.relu(x.double() @ w1.double().T + b1.double()) @ w2.double().T + b2.double()
    y3 = mm3(torch.relu(mm3(x, w1) + b1), w2) + b2
    y1 = (torch.relu(split(x)[0] @ split(w1)[0].T + b1).to(torch.bfloat16).float() @ split(w2)[0].T) + b2
    scale = ref.abs().max()
    e3 = ((y3.double() - ref).abs().max() / scale).item()
    e1 = ((y1.double() - ref).abs().max() / scale).item()
    assert e3 < 2e-5, e3                 # ~8e-6 here; measured on the GPU path: <= 7e-6 per layer tap
    assert e1 > 1e-4, e1                 # single-pass BF16 breaks the budget on one FeedForward already


@pytest.mark.parametrize("name", ["grad_c2arch_16", "grad_unshared_w32", "grad_cno_grid2d_w64", "grad_plus2d_w32", "grad_plus2d_w64"])
def test_oracle_is_differentiable_and_matches_reference_gradients(name):
    """Backward row groundwork: autograd through the oracle restatement reproduces the gradients the
    executed reference computes for its one-step training loss (routines/grid_2d_markov.py:172-193) w.r.t. the input
    and every parameter — weight-norm (g, v) pairs, spectral weights shared over layers, rfft/einsum/irfft adjoints.
    A CUDA backward can then be checked against ``torch.autograd`` of the oracle at any size."""
    kw, sd, a = load(name)
    leaves = {}                         # keep the aliasing of a shared ParameterList: one leaf per distinct tensor
    for k, v in sd.items():
        if id(v) not in leaves:
            leaves[id(v)] = v.clone().requires_grad_(True)
    p = {k: leaves[id(v)] for k, v in sd.items()}
    x = a["x"].clone().requires_grad_(True)
    out = O.block_grid2d_forward(p, x, modes=kw["modes"], n_layers=kw["n_layers"],
                                 n_ff_layers=kw.get("n_ff_layers", 2), layer_norm=kw.get("layer_norm", False))
    B = x.shape[0]
    loss = O.lp_loss_rel(out["forecast"].reshape(B, -1), a["y"].reshape(B, -1))
    loss.backward()
    assert rel_err(out["forecast"], a["forecast"]) < TOL
    assert abs(loss.item() - a["loss"].item()) < 1e-6 * abs(a["loss"].item())
    assert rel_err(x.grad, a["grad::x"]) < 2e-5
    checked = 0
    for k, ref in a.items():
        if not k.startswith("grad::") or k == "grad::x":
            continue
        g = p[k[6:]].grad
        assert g is not None, k
        assert rel_err(g, ref) < 2e-5, k
        checked += 1
    assert checked >= 8


@pytest.mark.parametrize("name,modes", [("grad_mesh3d_w64", ("modes_x", "modes_y", "modes_z")),
                                        ("grad_mesh2d_w32", ("modes_x", "modes_y")),
                                        ("grad_cno_mesh2d_w32", ("modes_x", "modes_y"))])
def test_oracle_mesh_gradients_match_reference(name, modes):
    """Mesh variants of the backward row: grid append, zero padding and crop are inside the graph."""
    kw, sd, a = load(name)
    leaves = {}
    for k, v in sd.items():
        if id(v) not in leaves:
            leaves[id(v)] = v.clone().requires_grad_(True)
    p = {k: leaves[id(v)] for k, v in sd.items()}
    x = a["x"].clone().requires_grad_(True)
    out = O.block_mesh_forward(p, x, modes=[kw[m] for m in modes], n_layers=kw["n_layers"])
    B = x.shape[0]
    loss = O.lp_loss_rel(out.reshape(B, -1), a["y"].reshape(B, -1))
    loss.backward()
    assert rel_err(out, a["out"]) < TOL
    assert rel_err(x.grad, a["grad::x"]) < 2e-5
    checked = 0
    for k, ref in a.items():
        if k.startswith("grad::") and k != "grad::x":
            assert rel_err(p[k[6:]].grad, ref) < 2e-5, k
            checked += 1
    assert checked >= 10


def test_oracle_geo_interior_gradients_match_reference():
    """Autograd through the oracle's interior loop reproduces the reference's gradients of <r, interior(uc, bias)> w.r.t.
    the latent grid, the grid bias and the interior layers' parameters (the pin of ffno_layers_bwd)."""
    kw, sd, a = load_geo("grad_geo_pointcloud")
    p = {k: (v.clone().requires_grad_(True) if k.startswith("convs.") and not v.is_complex() else v) for k, v in sd.items()}
    uc, bias = a["uc_in"].clone().requires_grad_(True), a["grid_bias"].clone().requires_grad_(True)
    out = O.geo_interior_forward(p, uc, bias, modes=kw["modes1"], n_layers=kw["n_layers"])
    (out * a["r"]).sum().backward()
    assert rel_err(out, a["uc_out"]) < TOL
    assert rel_err(uc.grad, a["igrad::uc"]) < 2e-5 and rel_err(bias.grad, a["igrad::bias"]) < 2e-5
    checked = 0
    for k, ref in a.items():
        if k.startswith("igrad::convs."):
            assert rel_err(p[k[7:]].grad, ref) < 2e-5, k
            checked += 1
    assert checked >= 18
