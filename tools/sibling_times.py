"""Forward time of the sibling operators at the shapes their shipped configs use, next to the reference's
own modules (oracle/_ref, eager PyTorch, allow_tf32 off) on the same GPU with the same weights and input; informational.

  plasticity/fcno   : CNOFactorizedMesh3D, 101 x 31 x 20 mesh -> padded 109 x 39 x 28, modes 32/12/8 capped to the DCT count, batch 4
  kochkov/fcno 64   : CNOFactorized2DBlock, 64 x 64, modes 16, 24 layers, batch 32
  no_factorization  : FNOPlus2DBlock, 64 x 64, modes 16, 4 layers, batch 32 (the ablation's depth range is 4..24)
  elasticity/ffno   : FNOFactorizedPointCloud2D, 972 points, 64 x 64 latent grid, width 64, 8 layers, batch 20 (iphi = None)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import fourierflow_b200.modules as OURS  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def timed(fn, n=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


def ref_class(path, name):
    import importlib
    return getattr(importlib.import_module(path), name)


def first(y):
    return y["forecast"] if isinstance(y, dict) else y


CASES = [
    ("plasticity/fcno CNOFactorizedMesh3D 12 layers batch 4", "CNOFactorizedMesh3D", "fourierflow.modules.factorized_cno.mesh_3d",
     dict(modes_x=32, modes_y=12, modes_z=8, width=64, input_dim=4, output_dim=4, n_layers=12, share_weight=False, factor=4,
          ff_weight_norm=True, n_ff_layers=2, layer_norm=False), (4, 101, 31, 20, 1)),
    ("torus_kochkov/fcno CNOFactorized2DBlock 24 layers 64x64 batch 32", "CNOFactorized2DBlock",
     "fourierflow.modules.factorized_cno.grid_2d",
     dict(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1), (32, 64, 64, 3)),
    ("torus_li/no_factorization FNOPlus2DBlock 4 layers 64x64 batch 32", "FNOPlus2DBlock",
     "fourierflow.modules.zongyi_fno.grid_plus_2d",
     dict(modes=16, width=64, n_layers=4, input_dim=3, share_weight=False, factor=4, ff_weight_norm=True, gain=0.1), (32, 64, 64, 3)),
    ("elasticity/ffno FNOFactorizedPointCloud2D 8 layers 972 points batch 20", "FNOFactorizedPointCloud2D",
     "fourierflow.modules.factorized_fno.point_cloud_2d",
     dict(modes1=16, modes2=16, width=64, in_channels=2, out_channels=1, n_layers=8, s1=64, s2=64), (20, 972, 2)),
]

out = {}
with torch.no_grad():
    for title, cls, ref_path, kw, shape in CASES:
        torch.manual_seed(0)
        ref = ref_class(ref_path, cls)(**kw).eval()
        ours = getattr(OURS, cls)(**kw)
        ours.load_state_dict(ref.state_dict(), strict=True)
        ref, ours = ref.cuda(), ours.cuda().eval()
        x = torch.rand(*shape, device="cuda") if "PointCloud" in cls else torch.randn(*shape, device="cuda")
        y, y_ref = first(ours(x)), first(ref(x))
        err = float((y.double() - y_ref.double()).abs().max() / y_ref.double().abs().max())
        ms, ms_ref = timed(lambda: ours(x)), timed(lambda: ref(x), n=5, warm=2)
        out[title] = {"ms": ms, "reference_eager_ms": ms_ref, "speedup": ms_ref / ms, "max_rel_err_vs_reference": err}
        print(title, json.dumps(out[title]), flush=True)
        del ref, ours
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sibling_times.json"), "w"), indent=1)
