"""CPU: the hydra/omegaconf-free config loader (fourierflow_b200/config.py).

The YAML below is written for this test in the schema of the reference's experiment files
(experiments/torus_li/markov/24_layers/config.yaml, experiments/plasticity/ffno/24_layers/config.yaml): same block
names, `_target_` paths, resolvers and hyper-parameters.  When the reference tree is mounted (this container only),
every torus_li / plasticity F-FNO config in it is loaded as well.
"""
import functools
import glob
import os

import pytest
import torch

from fourierflow_b200 import config as C
from fourierflow_b200.modules import FNOFactorized2DBlock, FNOFactorizedMesh3D
from fourierflow_b200.routines import Grid2DMarkovExperiment, StructuredMeshExperiment

MARKOV_YAML = """
wandb:
  project: torus_li
  group: markov/24_layers
builder:
  _target_: fourierflow.builders.NSMarkovBuilder
  data_path: ${oc.env:DATA_ROOT}/zongyi/NavierStokes_V1e-5_N1200_T20.mat
  train_size: 1000
  test_size: 200
  batch_size: 19
routine:
  _target_: fourierflow.routines.Grid2DMarkovExperiment
  conv:
    _target_: fourierflow.modules.FNOFactorized2DBlock
    modes: 16
    width: 64
    n_layers: 24
    input_dim: 3
    share_weight: true
    factor: 4
    ff_weight_norm: true
    gain: 0.1
    dropout: 0.0
    in_dropout: 0.0
  n_steps: 10
  max_accumulations: 1000
  noise_std: 0.01
  optimizer:
    _target_: functools.partial
    _args_: ["${get_method: torch.optim.AdamW}"]
    lr: 0.0025
    weight_decay: 0.0001
  scheduler:
    scheduler:
      _target_: functools.partial
      _args_: ["${get_method: fourierflow.schedulers.CosineWithWarmupScheduler}"]
      num_warmup_steps: 500
      num_training_steps: 100000
      num_cycles: 0.5
    name: learning_rate
trainer:
  accelerator: gpu
  devices: 1
  max_epochs: ${eval:1 + 100}
callbacks:
  - _target_: fourierflow.callbacks.CustomModelCheckpoint
    save_top_k: 1
  - _target_: pytorch_lightning.callbacks.ModelSummary
    max_depth: 4
"""

MESH_YAML = """
routine:
  _target_: fourierflow.routines.StructuredMeshExperiment
  model:
    _target_: fourierflow.modules.FNOFactorizedMesh3D
    modes_x: 12
    modes_y: 12
    modes_z: 8
    width: 64
    input_dim: 4
    output_dim: 4
    n_layers: 4
    share_weight: false
    factor: 4
    ff_weight_norm: true
    n_ff_layers: 2
    layer_norm: false
  optimizer:
    _target_: functools.partial
    _args_: ["${get_method: torch.optim.AdamW}"]
    lr: ${lr}
lr: 0.001
"""


def test_markov_routine_block_instantiates_unchanged():
    torch.manual_seed(0)
    routine, cfg = C.load_routine(MARKOV_YAML)
    assert isinstance(routine, Grid2DMarkovExperiment)
    assert isinstance(routine.conv, FNOFactorized2DBlock)
    assert routine.n_steps == 10 and routine.conv.input_dim == 3
    sd = routine.conv.state_dict()
    assert len(sd) == 203                                   # 203 keys for the C2 model
    assert sum(p.numel() for p in routine.conv.parameters()) == 1072834
    assert sd["spectral_layers.3.backcast_ff.layers.0.0.weight_v"].shape == (256, 64)
    # the rest of the file is carried along: resolvers evaluated, missing env var left as written
    assert cfg["trainer"]["max_epochs"] == 101
    assert cfg["builder"]["data_path"].startswith("${oc.env:DATA_ROOT}")


def test_partials_and_out_of_scope_targets():
    cfg = C.load_config(MARKOV_YAML)
    opt = C.instantiate(cfg["routine"]["optimizer"])
    assert isinstance(opt, functools.partial) and opt.func is torch.optim.AdamW and opt.keywords["lr"] == 0.0025
    sched = C.instantiate(cfg["routine"]["scheduler"]["scheduler"])
    assert isinstance(sched.func, C.MissingTarget)          # training harness: loads, raises only when used
    with pytest.raises(RuntimeError, match="not available in fourierflow_b200"):
        sched(None)
    with pytest.raises(RuntimeError, match="fourierflow.callbacks.CustomModelCheckpoint"):
        C.instantiate(cfg["callbacks"][0])


def test_env_resolver_overrides_and_references(monkeypatch):
    monkeypatch.setenv("DATA_ROOT", "/data")
    cfg = C.load_config(MARKOV_YAML, overrides=["routine.conv.n_layers=4", "routine.n_steps=3", "+extra.flag=true"])
    assert cfg["builder"]["data_path"] == "/data/zongyi/NavierStokes_V1e-5_N1200_T20.mat"
    assert cfg["routine"]["conv"]["n_layers"] == 4 and cfg["routine"]["n_steps"] == 3 and cfg["extra"]["flag"] is True
    routine, _ = C.load_routine(MESH_YAML)
    assert isinstance(routine, StructuredMeshExperiment) and isinstance(routine.model, FNOFactorizedMesh3D)
    assert C.instantiate(C.load_config(MESH_YAML)["routine"]["optimizer"]).keywords["lr"] == 0.001   # ${lr}
    monkeypatch.delenv("DATA_ROOT")
    with pytest.raises(KeyError, match="DATA_ROOT"):
        C.load_config(MARKOV_YAML, strict=True)
    assert C.load_config("v: ${oc.env:NOPE_NOT_SET,fallback}")["v"] == "fallback"


def test_unknown_symbols_fail_loudly():
    with pytest.raises(ImportError):
        C.locate("no_such_package.thing")
    # a reference operator outside the hot path (e.g. the geo-FNO baselines) is named in the error, not silently faked
    with pytest.raises(RuntimeError, match="fourierflow.modules.FNOZongyi2DBlock.*not available"):
        C.instantiate({"_target_": "fourierflow.modules.FNOZongyi2DBlock", "modes1": 12, "modes2": 12, "width": 20})


REF_EXPERIMENTS = "/root/reference/experiments"


@pytest.mark.skipif(not os.path.isdir(REF_EXPERIMENTS), reason="reference tree not mounted (GPU box)")
def test_reference_ffno_configs_load_from_the_mounted_tree():
    """Every shipped torus_li F-FNO (markov) config and the plasticity F-FNO configs: routine block -> our modules."""
    paths = sorted(glob.glob(os.path.join(REF_EXPERIMENTS, "torus_li", "markov", "*", "config.yaml")) +
                   glob.glob(os.path.join(REF_EXPERIMENTS, "plasticity", "ffno*", "*", "config.yaml")) +
                   glob.glob(os.path.join(REF_EXPERIMENTS, "plasticity", "fcno", "*", "config.yaml")) +      # DCT siblings
                   glob.glob(os.path.join(REF_EXPERIMENTS, "torus_kochkov", "fcno", "*", "*", "config.yaml")))
    assert len(paths) >= 4
    seen = 0
    for p in paths:
        raw = C.load_config(p, resolve=False)
        tgt = raw["routine"].get("conv", raw["routine"].get("model", {})).get("_target_", "")
        if not tgt.startswith(("fourierflow.modules.FNOFactorized", "fourierflow.modules.CNOFactorized")):
            continue
        routine, _ = C.load_routine(p)
        op = getattr(routine, "conv", None) or routine.model
        assert type(op).__module__.startswith("fourierflow_b200.modules")
        seen += 1
    assert seen >= 4


KOCHKOV_YAML = """
routine:
  _target_: fourierflow.routines.Grid2DMarkovExperiment
  conv:
    _target_: fourierflow.modules.FNOFactorized2DBlock
    modes: 64
    width: 64
    n_layers: 2
    input_dim: 5
    share_weight: true
    factor: 4
    ff_weight_norm: true
    gain: 0.1
  step_size: 0.28049934407051724
  max_accumulations: 38736
  noise_std: 0.01
  use_velocity: true
  grid_size: [256]
  domain:
    - [0, '${eval:2 * ${import:numpy.pi}}']
    - [0, '${eval:2 * ${import:numpy.pi}}']
"""


def test_kochkov_routine_with_velocity_features_and_nested_interpolation():
    """experiments/torus_kochkov/ffno/grid_sizes/256/config.yaml schema: nested ${eval:.. ${import:..}} keeps its
    float type and use_velocity switches the routine to the 5-feature input [w, q, v, gx, gy]."""
    import math
    routine, cfg = C.load_routine(KOCHKOV_YAML)
    assert cfg["routine"]["domain"][0][1] == pytest.approx(2 * math.pi) and isinstance(cfg["routine"]["domain"][0][1], float)
    assert routine.use_velocity and routine.conv.input_dim == 5
    assert routine.domain_lengths == (pytest.approx(2 * math.pi), pytest.approx(2 * math.pi))
    assert routine.normalizer.sum.shape == (5,)
    with pytest.raises(RuntimeError, match="input_dim"):
        C.load_routine(KOCHKOV_YAML, overrides=["routine.use_velocity=false"])


def test_predict_command_parses_a_config_and_refuses_to_run_without_a_gpu(tmp_path, capsys):
    """python -m fourierflow_b200.predict mirrors commands/predict.py:87-105; without a CUDA device it must stop loudly
    after building the routine (there is no CPU path to time)."""
    import torch
    from fourierflow_b200 import predict
    cfg = tmp_path / "config.yaml"
    cfg.write_text(MARKOV_YAML)
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu-marked test in test_gpu_parity.py")
    with pytest.raises(SystemExit, match="no CUDA device"):
        predict.main([str(cfg), "routine.conv.n_layers=2", "--samples", "4"])


@pytest.mark.skipif(not os.path.isdir(REF_EXPERIMENTS), reason="reference tree not mounted (GPU box)")
def test_reference_torus_vis_configs_load_with_force_and_mu_channels():
    """experiments/torus_vis*/: Grid2DMarkovExperiment with append_force / append_mu (routines/grid_2d_markov.py:146-152)
    — the feature count the routine builds must equal the conv's input_dim in every shipped F-FNO config."""
    paths = sorted(glob.glob(os.path.join(REF_EXPERIMENTS, "torus_vis*", "*", "config.yaml")))
    seen = 0
    for p in paths:
        raw = C.load_config(p, resolve=False)
        r = raw["routine"]
        if not r.get("conv", {}).get("_target_", "").startswith("fourierflow.modules.FNOFactorized2DBlock"):
            continue
        if r["conv"].get("use_fork") or r["conv"].get("dropout", 0) or r["conv"].get("in_dropout", 0):
            continue
        routine, cfg = C.load_routine(p)
        want = 3 + int(bool(r.get("append_force"))) + int(bool(r.get("append_mu")))
        assert routine.conv.input_dim == want and routine.normalizer.sum.shape == (want,)
        assert routine.append_force == bool(r.get("append_force")) and routine.append_mu == bool(r.get("append_mu"))
        seen += 1
    assert seen >= 4


@pytest.mark.skipif(not os.path.isdir(REF_EXPERIMENTS), reason="reference tree not mounted (GPU box)")
def test_every_shipped_config_of_a_mirrored_operator_loads(monkeypatch):
    """All 256 experiment configs: every one whose routine and operator are mirrored here must build on this backend
    (routine block -> fourierflow_b200 classes); the rest must fail with the name of what is out of scope — the
    non-factorized geo-FNO / FNO baselines, the JAX experiments, and two ablation configs that cannot run in the
    reference either (use_fourier_position: undefined k_max; no_velocity_positional: 1 feature into input_dim 2)."""
    monkeypatch.setenv("DATA_ROOT", "/tmp/data")
    allowed = ("FNOMesh2D", "FNOMesh3D", "FNOPointCloud2D", "FNOZongyi2DBlock", "optax", "jax_cfd",
               "k_max", "input features")      # the last two: configs the reference itself cannot run
    loaded, refused = {}, {}
    for p in sorted(glob.glob(os.path.join(REF_EXPERIMENTS, "**", "config.yaml"), recursive=True)):
        grp = os.path.relpath(p, REF_EXPERIMENTS).split(os.sep)[0]
        try:
            routine, _ = C.load_routine(p)
        except (RuntimeError, ImportError) as e:
            assert any(a in str(e) for a in allowed), (p, str(e))
            refused[grp] = refused.get(grp, 0) + 1
            continue
        op = getattr(routine, "conv", None) or routine.model
        assert type(op).__module__.startswith("fourierflow_b200.modules"), p
        loaded[grp] = loaded.get(grp, 0) + 1
    assert loaded == {"airfoil": 24, "elasticity": 18, "pipe": 12, "plasticity": 24, "torus_kochkov": 52, "torus_li": 46,
                      "torus_vis": 3, "torus_vis_force": 4}, loaded
    assert sum(refused.values()) == 73, refused
