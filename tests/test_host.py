"""CPU: host-side logic and the C-ABI boundary (no GPU compute): symbols, struct layout, state_dict
schema / seeded init parity against the executed reference, loud failure without a device."""
import copy
import ctypes as C
import json
import os
import re
import subprocess
import tempfile

import pytest
import torch

from golden_util import GOLDEN, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffno_b200.h")


@pytest.fixture(scope="module")
def lib():
    from fourierflow_b200 import build, _lib
    build.build()                      # no-op when up to date
    return _lib.load()


def _header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"FFNO_API\s+[\w\s\*]+?\b(ffno_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from fourierflow_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 19
    assert sorted(_lib.EXPORTS) == declared            # binding covers the header exactly
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ffno_\w+)", nm))
    assert set(declared) <= exported
    assert lib.ffno_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_the_c_header():
    """sizeof/offsetof of every ABI struct as gcc sees the header == the ctypes mirror."""
    from fourierflow_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ffno_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ffno_desc), sizeof(ffno_linear_params), sizeof(ffno_ff_params),
         sizeof(ffno_layer_params), sizeof(ffno_block_params), sizeof(ffno_taps));
  printf("%zu %zu %zu %zu\n", offsetof(ffno_desc, width), offsetof(ffno_desc, path),
         offsetof(ffno_layer_params, forecast_ff), offsetof(ffno_block_params, layers));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(ffno_desc, transform), sizeof(ffno_rollout_extras), sizeof(ffno_linear_grads),
         sizeof(ffno_layer_grads), sizeof(ffno_block_grads), offsetof(ffno_block_grads, layers));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [C.sizeof(t) for t in (_lib.Desc, _lib.LinearParams, _lib.FFParams, _lib.LayerParams,
                                   _lib.BlockParams, _lib.Taps)]
    offs = [_lib.Desc.width.offset, _lib.Desc.path.offset, _lib.LayerParams.forecast_ff.offset,
            _lib.BlockParams.layers.offset]
    more = [_lib.Desc.transform.offset, C.sizeof(_lib.RolloutExtras), C.sizeof(_lib.LinearGrads), C.sizeof(_lib.LayerGrads),
            C.sizeof(_lib.BlockGrads), _lib.BlockGrads.layers.offset]
    assert [int(v) for v in out] == sizes + offs + more


def test_plan_create_fails_loudly_without_a_device(lib):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fourierflow_b200 import _lib
    d = _lib.Desc()
    d.abi_version, d.ndim = _lib.ABI_VERSION, 2
    d.size[0] = d.size[1] = 16
    d.modes[0] = d.modes[1] = 4
    d.width, d.in_features, d.out_features, d.head_hidden = 32, 3, 1, 128
    d.n_layers, d.ff_factor, d.n_ff_layers = 1, 4, 2
    plan = C.c_void_p()
    st = lib.ffno_plan_create(C.byref(d), C.byref(plan))
    assert st == -3 and not plan.value                         # FFNO_ERR_CUDA, never a CPU fallback
    assert b"cuda" in lib.ffno_last_error().lower()
    assert lib.ffno_device_ok() == 0


def test_plan_create_rejects_bad_descriptors(lib):
    from fourierflow_b200 import _lib
    d = _lib.Desc()
    plan = C.c_void_p()
    assert lib.ffno_plan_create(C.byref(d), C.byref(plan)) == -1          # wrong ABI version
    d.abi_version, d.ndim = _lib.ABI_VERSION, 2
    d.size[0], d.size[1], d.modes[0], d.modes[1] = 40, 40, 32, 4
    d.width, d.in_features, d.out_features, d.head_hidden = 64, 3, 1, 128
    d.n_layers, d.ff_factor, d.n_ff_layers = 1, 4, 2
    # modes 32 on a 40-long axis has only 21 rfft bins: the reference raises too
    assert lib.ffno_plan_create(C.byref(d), C.byref(plan)) == -1
    assert b"rfft bins" in lib.ffno_last_error()


def _cls(name):
    import fourierflow_b200.modules as M
    return getattr(M, name)


def test_seeded_init_matches_reference():
    cases = json.load(open(os.path.join(GOLDEN, "init_parity.json")))
    for c in cases:
        torch.manual_seed(c["seed"])
        m = _cls(c["cls"])(**c["kwargs"])
        sd = m.state_dict()
        assert list(sd.keys()) == list(c["keys"].keys()), c["cls"]
        for k, (shape, s, a) in c["keys"].items():
            v = sd[k].double()
            assert list(v.shape) == shape, k
            assert abs(v.sum().item() - s) <= 1e-6 * max(1.0, a), k
            assert abs(v.abs().sum().item() - a) <= 1e-6 * max(1.0, a), k


@pytest.mark.parametrize("name,cls", [("grid2d_c2arch_32", "FNOFactorized2DBlock"),
                                      ("grid2d_fork", "FNOFactorized2DBlock"),
                                      ("grid2d_ln_w32", "FNOFactorized2DBlock"),
                                      ("mesh2d_small", "FNOFactorizedMesh2D"),
                                      ("mesh3d_w64", "FNOFactorizedMesh3D"),
                                      ("cno_grid2d_w64", "CNOFactorized2DBlock"),
                                      ("plus2d_w32", "FNOPlus2DBlock"),
                                      ("plus2d_shared_fork", "FNOPlus2DBlock"),
                                      ("cno_mesh2d_small", "CNOFactorizedMesh2D"),
                                      ("cno_mesh3d_w64", "CNOFactorizedMesh3D")])
def test_reference_checkpoints_load_strict(name, cls):
    kw, sd, _ = load(name)
    m = _cls(cls)(**kw)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    m2 = copy.deepcopy(m)                                      # SWA deep-copies the routine (linear.py:53-79)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2) and v1.data_ptr() != v2.data_ptr()
    if "fourier_weight.0" in sd:                               # shared ParameterList stays shared
        assert m.spectral_layers[0].fourier_weight[0] is m.fourier_weight[0]
        assert m2.spectral_layers[1].fourier_weight[0] is m2.fourier_weight[0]


def test_weight_property_folds_weight_norm():
    from fourierflow_b200.modules import WNLinear
    from oracle.ffno_oracle import weight_norm_fold
    lin = WNLinear(8, 5, wnorm=True)
    with torch.no_grad():
        lin.weight_g.mul_(1.7)
    assert torch.allclose(lin.weight, weight_norm_fold(lin.weight_g, lin.weight_v))
    assert list(lin.state_dict().keys()) == ["bias", "weight_g", "weight_v"]
    assert list(WNLinear(8, 5).state_dict().keys()) == ["weight", "bias"]


def test_forward_refuses_cpu_tensors_and_autograd():
    m = _cls("FNOFactorized2DBlock")(modes=4, width=32, n_layers=1, input_dim=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m(torch.randn(1, 8, 8, 3))
    with pytest.raises(RuntimeError, match="dropout"):
        _cls("FNOFactorized2DBlock")(modes=4, width=32, dropout=0.1)


def test_normalizer_mirror_matches_oracle():
    from fourierflow_b200.modules import Normalizer
    from oracle import ffno_oracle as O
    n = Normalizer([3], 1e6)
    x = torch.randn(5, 4, 4, 3)
    n.train()
    y = n(x)
    mean, std = O.normalizer_mean_std(O.normalizer_stats(x))
    assert torch.allclose(n.mean, mean) and torch.allclose(n.std, std)
    assert torch.allclose(y, (x - mean) / std)
    n.eval()
    assert torch.allclose(n.inverse(y[..., :1], channel=0), x[..., :1], atol=1e-5)
    assert set(n.state_dict()) == {"count", "n_accumulations", "sum", "sum_squared", "one", "std_epsilon"}


def test_routine_loads_a_reference_style_lightning_checkpoint(tmp_path):
    """routines/base.py:79-102: checkpoint['state_dict'] with the reference's key names; the kx/ky/lap buffers of a
    use_velocity routine are dropped (this backend computes them in ffno_velocity_fwd) and strict is relaxed only then."""
    from fourierflow_b200.modules import FNOFactorized2DBlock
    from fourierflow_b200.routines import Grid2DMarkovExperiment

    def make(seed):
        torch.manual_seed(seed)
        conv = FNOFactorized2DBlock(modes=8, width=32, n_layers=2, input_dim=5, share_weight=True, factor=4,
                                    ff_weight_norm=True, gain=0.1)
        return Grid2DMarkovExperiment(conv, n_steps=3, use_velocity=True)

    src, dst = make(0), make(1)
    src.normalizer.sum += 3.0
    src.normalizer.count += 7.0
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    assert "conv.spectral_layers.1.backcast_ff.layers.0.0.weight_g" in sd and "normalizer.sum" in sd and "_float" in sd
    sd.update({"kx_64": torch.zeros(64, 33), "ky_64": torch.zeros(64, 33), "lap_64": torch.zeros(64, 33, dtype=torch.complex64)})
    path = tmp_path / "epoch=3.ckpt"
    torch.save({"state_dict": sd, "epoch": 3}, path)
    dst.load_lightning_model_state(str(path))
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    # without removable keys the load is strict: a missing key must raise
    bad = {k: v for k, v in src.state_dict().items() if k != "normalizer.sum"}
    with pytest.raises(RuntimeError, match="normalizer.sum"):
        make(2).load_lightning_model_state({"state_dict": bad})
    assert make(3).infer.__func__ is src.infer.__func__ and src.warmup() is None


def test_parameter_cache_follows_replaced_parameters_and_modules():
    """The plan's view of the parameters must follow load_state_dict(assign=True), parameter re-registration and
    swapped / added sub-modules (the reference re-reads the live parameters on every forward, linear.py:49)."""
    torch.manual_seed(0)
    m = _cls("FNOFactorized2DBlock")(modes=4, width=32, n_layers=2, input_dim=3, share_weight=True, factor=4,
                                     ff_weight_norm=True)
    p0 = m._flat_params()
    assert len(p0) == len(list(m.parameters())) and m._flat_params() is p0       # cached while nothing changed
    m.__dict__["_spec_cache"] = "stale"
    sd = {k: v.clone() + 1.0 for k, v in m.state_dict().items()}
    m.load_state_dict(sd, assign=True)                                          # every Parameter object is replaced
    p1 = m._flat_params()
    assert p1 is not p0 and "_spec_cache" not in m.__dict__
    assert {id(p) for p in p1} == {id(p) for p in m.parameters()}
    assert not ({id(p) for p in p1} & {id(p) for p in p0})
    m.spectral_layers[1] = copy.deepcopy(m.spectral_layers[0])                   # swapped sub-module
    p2 = m._flat_params()
    assert {id(p) for p in p2} == {id(p) for p in m.parameters()}
    m.spectral_layers.append(copy.deepcopy(m.spectral_layers[0]))                # added layer
    assert {id(p) for p in m._flat_params()} == {id(p) for p in m.parameters()}
    m.invalidate_plans()
    assert "_param_cache" not in m.__dict__


def test_rollout_contract_checks():
    """n_steps must leave a start frame (a negative index would wrap silently); dropout / unsupported options raise."""
    from fourierflow_b200.modules import FNOFactorized2DBlock
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    conv = FNOFactorized2DBlock(modes=4, width=32, n_layers=1, input_dim=3)
    exp = Grid2DMarkovExperiment(conv, n_steps=6)
    with pytest.raises(RuntimeError):           # CPU tensor: refused before anything else
        exp.predict(torch.randn(1, 8, 8, 4))


def test_rollout_feature_set_options():
    """append_force / append_mu / use_position are part of the feature count; use_fourier_position raises (it cannot
    run in the reference either), the other ablation switches select the host step loop."""
    from fourierflow_b200.modules import FNOFactorized2DBlock
    from fourierflow_b200.routines import Grid2DMarkovExperiment
    conv5 = FNOFactorized2DBlock(modes=4, width=32, n_layers=1, input_dim=5)
    exp = Grid2DMarkovExperiment(conv5, n_steps=2, append_force=True, append_mu=True)
    assert exp.append_force and exp.append_mu and not exp.use_velocity
    with pytest.raises(RuntimeError, match="input features"):      # 3 + force = 4 features, conv takes 5
        Grid2DMarkovExperiment(conv5, n_steps=2, append_force=True)
    conv3 = FNOFactorized2DBlock(modes=4, width=32, n_layers=1, input_dim=3)
    with pytest.raises(RuntimeError, match="k_max"):
        Grid2DMarkovExperiment(conv3, use_fourier_position=True)
    assert Grid2DMarkovExperiment(conv3)._fused
    for flag in ("shuffle_grid", "learn_difference"):
        assert not Grid2DMarkovExperiment(conv3, grid_size=[8], **{flag: True})._fused
    assert not Grid2DMarkovExperiment(FNOFactorized2DBlock(modes=4, width=32, n_layers=1, input_dim=1), use_position=False)._fused
    with pytest.raises(RuntimeError, match="input features"):      # no position features: 1 input feature, conv takes 3
        Grid2DMarkovExperiment(conv3, use_position=False)
    # statistics: frame t of a time-varying forcing accompanies input frame t, mu is broadcast
    B, X, T = 2, 8, 5
    data, f, mu = torch.randn(B, X, X, T), torch.randn(B, X, X, T), torch.rand(B)
    exp.accumulate_statistics(data, force=f, mu=mu)
    mean = exp.normalizer.mean
    assert torch.allclose(mean[0], data[..., :-1].mean(), atol=1e-5)
    assert torch.allclose(mean[3], f[..., :-1].mean(), atol=1e-5) and torch.allclose(mean[4], mu.mean(), atol=1e-6)
    with pytest.raises(RuntimeError, match="batch\\['f'\\]"):
        exp.accumulate_statistics(data, mu=mu)
    with pytest.raises(RuntimeError, match="must be \\[B\\]"):
        exp.accumulate_statistics(data, force=f, mu=torch.rand(B, 1))


def test_cno_mirrors_keep_the_reference_contract():
    """factorized_cno: real [in, out, modes] weights; shared weights of the mesh variants are 4-D in the reference and
    cannot run there (mesh_2d.py:118-124 vs :69-72) — the mirrors refuse them at construction."""
    import fourierflow_b200.modules as M
    b = M.CNOFactorized2DBlock(modes=5, width=32, input_dim=3, n_layers=2, share_weight=True, factor=4, ff_weight_norm=True)
    assert [tuple(p.shape) for p in b.fourier_weight] == [(32, 32, 5)] * 2
    assert b.spectral_layers[1].fourier_weight[0] is b.fourier_weight[0]
    with pytest.raises(RuntimeError, match="share_weight"):
        M.CNOFactorizedMesh2D(4, 4, 32, 4, 2, True, 4, True, 2, False)
    with pytest.raises(RuntimeError, match="share_weight"):
        M.CNOFactorizedMesh3D(4, 4, 4, 32, 4, 4, 2, True, 4, True, 2, False)
    from fourierflow_b200.modules.factorized_fno._base import check_trainable
    check_trainable(b)                                        # the DCT and the un-factorized stacks are trainable
    check_trainable(M.FNOPlus2DBlock(modes=4, width=32, input_dim=3, n_layers=1, factor=4, ff_weight_norm=True))


@pytest.mark.parametrize("name", ["geo_pointcloud_w32", "geo_pointcloud_shared"])
def test_geo_mirror_loads_reference_checkpoints_strict(name):
    from golden_util import load_geo
    import fourierflow_b200.modules as M
    kw, sd, _ = load_geo(name)
    m = M.FNOFactorizedPointCloud2D(**kw)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert m.convs[m.n_layers].weights1.dtype == torch.cfloat and not hasattr(m.convs[0], "weights1")
    if kw["share_weight"]:
        assert m.convs[1].fourier_weight[0] is m.fourier_weight[0]
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(1, 8, 2))


def test_iphi_mirror_keeps_the_reference_parameters():
    """modules/iphi.py:14-20: parameter names, shapes and creation order (the reference module itself cannot be built
    without a GPU: it creates its constants with device="cuda"); forward on CPU runs (torch glue, no CUDA library)."""
    import fourierflow_b200.modules as M
    m = M.IPhi(width=32)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [
        ("fc0.weight", (32, 4)), ("fc0.bias", (32,)), ("fc_code.weight", (32, 42)), ("fc_code.bias", (32,)),
        ("fc_no_code.weight", (128, 96)), ("fc_no_code.bias", (128,)), ("fc1.weight", (128, 128)), ("fc1.bias", (128,)),
        ("fc2.weight", (128, 128)), ("fc2.bias", (128,)), ("fc3.weight", (128, 128)), ("fc3.bias", (128,)),
        ("fc4.weight", (2, 128)), ("fc4.bias", (2,))]
    x = torch.rand(2, 7, 2)
    assert m(x, code=torch.rand(2, 42)).shape == x.shape and m(x).shape == x.shape
