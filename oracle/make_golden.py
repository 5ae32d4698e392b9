"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (test infrastructure).

Run in the build container only (needs /root/reference):  ``python oracle/make_golden.py``

The reference's ``fourierflow/__init__.py`` imports hydra/xarray/jax/pytorch_lightning (absent
offline), so the operator sub-package is imported through a stub parent package.
``fourierflow.routines`` cannot be imported at all; the 10-step Markov rollout fixture is therefore
produced by driving the reference's own ``FNOFactorized2DBlock`` + ``Normalizer`` + ``LpLoss``
objects with the loop of routines/grid_2d_markov.py:195-326 written out below.

Each fixture stores: the constructor kwargs (JSON), the full reference ``state_dict``, the input,
and the reference outputs (final + per-layer taps).  Nothing here is imported by the product.
"""
from __future__ import annotations

import json
import os
import re
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REF = os.environ.get("FFNO_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    pkg = types.ModuleType("fourierflow")
    pkg.__path__ = [os.path.join(REF, "fourierflow")]
    sys.modules["fourierflow"] = pkg
    import fourierflow.modules as M  # noqa
    from fourierflow.modules.loss import LpLoss  # noqa
    return M, LpLoss


def sd_np(module):
    """state_dict → arrays.  With share_weight=True the reference registers the SAME ParameterList on
    the block and on every layer (grid_2d.py:125-147), so ``spectral_layers.{l}.fourier_weight.{a}``
    alias ``fourier_weight.{a}``; the aliases are dropped here and re-created by the loader
    (tests/golden_util.py) to keep the fixtures small."""
    sd = module.state_dict()
    shared = "fourier_weight.0" in sd
    out = {}
    for k, v in sd.items():
        if shared and re.match(r"spectral_layers\.\d+\.fourier_weight\.\d+$", k):
            continue
        out["sd::" + k] = v.detach().cpu().numpy()
    return out


def save(name, kwargs, arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = dict(arrays)
    arrays["kwargs_json"] = np.frombuffer(json.dumps(kwargs).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def perturb_(module, seed, scale=0.5):
    """Move biases / weight_g / LayerNorm params off their trivial init so parity sees them."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in module.named_parameters():
            if k.endswith("bias") or k.endswith("weight_g") or ".3.weight" in k:
                p.add_(scale * torch.randn(p.shape, generator=g) * p.abs().mean().clamp(min=0.1))


def grid2d_case(M, name, kwargs, shape, seed, taps=True, x_scale=1.0, cls="FNOFactorized2DBlock"):
    torch.manual_seed(seed)
    m = getattr(M, cls)(**kwargs).eval()
    perturb_(m, seed + 100)
    x = x_scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed + 1))
    arrays = sd_np(m)
    arrays["x"] = x.numpy()
    with torch.no_grad():
        out = m(x)
        arrays["forecast"] = out["forecast"].numpy()
        for i, f in enumerate(out["forecast_list"]):
            arrays[f"forecast_list{i}"] = f.numpy()
        if taps:
            h = m.drop(m.in_proj(x))
            arrays["tap_lift"] = h.numpy()
            for l, layer in enumerate(m.spectral_layers):
                if l == 0 and kwargs.get("mode", "full") != "no-fourier":
                    arrays["tap_s0"] = layer.forward_fourier(h).numpy()
                b, _ = layer(h)
                h = h + b
                arrays[f"tap_x{l}"] = h.numpy()
            arrays["tap_b_last"] = b.numpy()
    save(name, kwargs, arrays)


def mesh_case(M, cls, name, kwargs, shape, seed):
    torch.manual_seed(seed)
    m = getattr(M, cls)(**kwargs).eval()
    perturb_(m, seed + 100)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed + 1))
    arrays = sd_np(m)
    arrays["x"] = x.numpy()
    with torch.no_grad():
        arrays["out"] = m(x).numpy()
    save(name, kwargs, arrays)


def spectral_case(M, name, C, K, shape, seed):
    """One bare ``SpectralConv2d.forward_fourier`` at the C2 layer shape (N=64, K=16, C=64)."""
    from fourierflow.modules.factorized_fno.grid_2d import SpectralConv2d
    torch.manual_seed(seed)
    layer = SpectralConv2d(in_dim=C, out_dim=C, n_modes=K, forecast_ff=None, backcast_ff=None,
                           fourier_weight=None, factor=4, ff_weight_norm=True, n_ff_layers=2,
                           layer_norm=False, use_fork=False, dropout=0.0, mode="full").eval()
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        s = layer.forward_fourier(x)
    arrays = {"sd::" + k: v.numpy() for k, v in layer.state_dict().items() if "fourier_weight" in k}
    arrays.update(x=x.numpy(), s=s.numpy())
    save(name, {"C": C, "K": K}, arrays)


def rollout_case(M, LpLoss, name, kwargs, B, X, T, n_steps, seed, force_dims=0, with_mu=False):
    """routines/grid_2d_markov.py:195-326 driven with the reference's own module objects.
    force_dims = 3 / 4: append_force with a static [B,X,Y] / time-varying [B,X,Y,T] forcing (:246-255, :288-289);
    with_mu: append_mu (:257-260, :290-291)."""
    torch.manual_seed(seed)
    conv = M.FNOFactorized2DBlock(**kwargs).eval()
    perturb_(conv, seed + 100)
    normalizer = M.Normalizer([conv.input_dim], 1e6)
    l2 = LpLoss(size_average=True)
    g = torch.Generator().manual_seed(seed + 1)
    # smooth-ish trajectories so the rel-L2 denominators are well away from zero
    base = torch.randn(B, X, X, 1, generator=g)
    data = base + 0.3 * torch.cumsum(torch.randn(B, X, X, T, generator=g), dim=-1)

    grids = [torch.linspace(0, 1, steps=X) for _ in range(2)]                    # :101-106, low=0 high=1
    pos = torch.stack(torch.meshgrid(*grids, indexing="ij"), dim=-1)
    pos = pos.unsqueeze(0).repeat(B, 1, 1, 1)

    force = mu = None
    if force_dims == 3:
        force = torch.randn(B, X, X, generator=g)
    elif force_dims == 4:
        force = torch.randn(B, X, X, T, generator=g)
    if with_mu:
        mu = torch.rand(B, generator=g) * 1e-3 + 1e-4

    def extras(t_index, n_frames):
        """force / mu channels of `n_frames` consecutive frames starting at frame t_index: [B,X,Y,n_frames,E]"""
        cols = []
        if force is not None:
            f = force.unsqueeze(-1).repeat(1, 1, 1, n_frames) if force.dim() == 3 else force[..., t_index:t_index + n_frames]
            cols.append(f.unsqueeze(-1))
        if mu is not None:
            cols.append(mu.reshape(B, 1, 1, 1, 1).repeat(1, X, X, n_frames, 1))
        return cols

    # one training-mode pass accumulates the running statistics (:376-378 → _build_features)
    normalizer.train()
    feats = torch.cat([data[..., :-1].unsqueeze(-1),
                       pos.unsqueeze(-2).repeat(1, 1, 1, T - 1, 1)] + extras(0, T - 1), dim=-1)    # [B,X,Y,T-1,F]
    feats = feats.permute(0, 3, 1, 2, 4).reshape(B * (T - 1), X, X, feats.shape[-1])
    normalizer(feats)
    normalizer.eval()

    yy = data[..., -n_steps:]
    preds, step_losses, loss = None, [], 0
    with torch.no_grad():
        for t in range(n_steps):
            if t == 0:
                x = torch.cat([data[..., T - n_steps - 1].unsqueeze(-1), pos], dim=-1)
            else:
                x = torch.cat([im, pos], dim=-1)
            if force is not None:      # :246-255: a 4-D force contributes its LAST n_steps frames, frame t at step t
                f_t = force if force.dim() == 3 else force[..., -n_steps:][..., t]
                x = torch.cat([x, f_t.unsqueeze(-1)], dim=-1)
            if mu is not None:
                x = torch.cat([x, mu.reshape(B, 1, 1, 1).repeat(1, X, X, 1)], dim=-1)
            x = normalizer(x)
            im = conv(x)["forecast"]
            im = normalizer.inverse(im, channel=0)
            l = l2(im.reshape(B, -1), yy[..., t].reshape(B, -1))
            step_losses.append(l)
            loss = loss + l
            preds = im if t == 0 else torch.cat((preds, im), dim=-1)
    arrays = sd_np(conv)
    arrays.update(data=data.numpy(), preds=preds.numpy(), loss=np.asarray(loss.item()),
                  step_losses=torch.stack(step_losses).numpy(),
                  norm_sum=normalizer.sum.numpy(), norm_sum_squared=normalizer.sum_squared.numpy(),
                  norm_count=normalizer.count.numpy(),
                  norm_mean=normalizer.mean.numpy(), norm_std=normalizer.std.numpy())
    if force is not None:
        arrays["force"] = force.numpy()
    if mu is not None:
        arrays["mu"] = mu.numpy()
    save(name, dict(kwargs, n_steps=n_steps), arrays)


def ablation_rollout_case(M, LpLoss, name, kwargs, B, X, T, n_steps, seed, use_position=True, shuffle=False,
                          learn_difference=False):
    """The ablation switches of `_valid_step` (grid_2d_markov.py:263-321) that five shipped configs use: no position
    features (use_position=False), a fixed random permutation of the grid rows / columns around the operator
    (shuffle_grid, :297-304), increments instead of frames (learn_difference, :309-318 — including the reference's
    `yy[..., t-1]` at t = 0, which wraps to the last frame).  Driven through the reference's own conv / Normalizer / LpLoss."""
    torch.manual_seed(seed)
    conv = M.FNOFactorized2DBlock(**kwargs).eval()
    perturb_(conv, seed + 100)
    normalizer = M.Normalizer([conv.input_dim], 1e6)
    l2 = LpLoss(size_average=True)
    g = torch.Generator().manual_seed(seed + 1)
    data = torch.randn(B, X, X, 1, generator=g) + 0.3 * torch.cumsum(torch.randn(B, X, X, T, generator=g), dim=-1)
    pos = torch.stack(torch.meshgrid(torch.linspace(0, 1, steps=X), torch.linspace(0, 1, steps=X), indexing="ij"), dim=-1)
    pos = pos.unsqueeze(0).repeat(B, 1, 1, 1)
    x_idx, y_idx = torch.randperm(X, generator=g), torch.randperm(X, generator=g)
    x_inv, y_inv = torch.argsort(x_idx), torch.argsort(y_idx)
    normalizer.train()
    cols = [data[..., :-1].unsqueeze(-1)] + ([pos.unsqueeze(-2).repeat(1, 1, 1, T - 1, 1)] if use_position else [])
    feats = torch.cat(cols, dim=-1).permute(0, 3, 1, 2, 4).reshape(B * (T - 1), X, X, -1)
    normalizer(feats)
    normalizer.eval()
    yy = data[..., -n_steps:]
    preds, step_losses, loss = None, [], 0
    with torch.no_grad():
        for t in range(n_steps):
            if t == 0:
                im = data[..., T - n_steps - 1].unsqueeze(-1)
                prev_im = im
            x = torch.cat([im, pos], dim=-1) if use_position else im
            x = normalizer(x)
            if shuffle:
                x = x[:, x_idx][:, :, y_idx]
            im = conv(x)["forecast"]
            if shuffle:
                im = im[:, :, y_inv][:, x_inv]
            im = normalizer.inverse(im, channel=0)
            y = yy[..., t] - yy[..., t - 1] if learn_difference else yy[..., t]
            l = l2(im.reshape(B, -1), y.reshape(B, -1))
            step_losses.append(l)
            loss = loss + l
            if learn_difference:
                im = prev_im + im
                prev_im = im
            preds = im if t == 0 else torch.cat((preds, im), dim=-1)
    arrays = sd_np(conv)
    arrays.update(data=data.numpy(), preds=preds.numpy(), loss=np.asarray(loss.item()),
                  step_losses=torch.stack(step_losses).numpy(), x_idx=x_idx.numpy(), y_idx=y_idx.numpy(),
                  norm_sum=normalizer.sum.numpy(), norm_sum_squared=normalizer.sum_squared.numpy(),
                  norm_count=normalizer.count.numpy())
    save(name, dict(kwargs, n_steps=n_steps, use_position=use_position, shuffle_grid=shuffle,
                    learn_difference=learn_difference), arrays)


def ablation_rollout_cases(M, LpLoss, c2):
    small = dict(c2, n_layers=3, modes=8)
    ablation_rollout_case(M, LpLoss, "rollout_nopos_16", dict(small, input_dim=1), B=2, X=16, T=8, n_steps=5, seed=60,
                          use_position=False)
    ablation_rollout_case(M, LpLoss, "rollout_shuffle_16", small, B=2, X=16, T=8, n_steps=5, seed=61, shuffle=True)
    ablation_rollout_case(M, LpLoss, "rollout_difference_16", small, B=2, X=16, T=8, n_steps=5, seed=62,
                          learn_difference=True)


def mesh_grad_case(M, LpLoss, cls, name, kwargs, shape, out_dim, seed):
    """The same pin for the mesh variants (routines/structured_mesh.py: LpLoss of the model output against the target):
    grid-coordinate append, zero padding before the layers and crop before the head are inside the autograd graph."""
    torch.manual_seed(seed)
    m = getattr(M, cls)(**kwargs).train()
    perturb_(m, seed + 100)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(*shape, generator=g).requires_grad_(True)
    y = torch.randn(*shape[:-1], out_dim, generator=g)
    B = shape[0]
    arrays = sd_np(m)
    arrays["x"] = x.detach().numpy()
    arrays["y"] = y.numpy()
    out = m(x)
    loss = LpLoss(size_average=True)(out.reshape(B, -1), y.reshape(B, -1))
    loss.backward()
    arrays["out"] = out.detach().numpy()
    arrays["loss"] = loss.detach().numpy()
    arrays["grad::x"] = x.grad.numpy()
    seen = set()
    for k, p_ in m.named_parameters():
        if id(p_) in seen:
            continue
        seen.add(id(p_))
        arrays["grad::" + k] = p_.grad.detach().numpy()
    save(name, kwargs, arrays)


def grad_case(M, LpLoss, name, kwargs, shape, seed, cls="FNOFactorized2DBlock"):
    """Gradients of the reference's one-step training loss (routines/grid_2d_markov.py:172-193: forecast ->
    LpLoss against the next frame) w.r.t. the input and every parameter, by the reference's own autograd graph —
    the pin for the backward row: weight-norm (g, v), shared spectral weights accumulated over
    layers, complex einsum / rfft / irfft adjoints."""
    torch.manual_seed(seed)
    m = getattr(M, cls)(**kwargs).train()
    perturb_(m, seed + 100)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(*shape, generator=g).requires_grad_(True)
    y = torch.randn(*shape[:-1], 1, generator=g)
    B = shape[0]
    arrays = sd_np(m)
    arrays["x"] = x.detach().numpy()
    arrays["y"] = y.numpy()
    forecast = m(x)["forecast"]
    loss = LpLoss(size_average=True)(forecast.reshape(B, -1), y.reshape(B, -1))
    loss.backward()
    arrays["forecast"] = forecast.detach().numpy()
    arrays["loss"] = loss.detach().numpy()
    arrays["grad::x"] = x.grad.numpy()
    seen = set()
    for k, p_ in m.named_parameters():            # named_parameters de-duplicates the shared ParameterList
        if id(p_) in seen:
            continue
        seen.add(id(p_))
        arrays["grad::" + k] = p_.grad.detach().numpy()
    save(name, kwargs, arrays)


def cno_cases():
    """The DCT siblings (fourierflow/modules/factorized_cno/*, not exported by modules/__init__.py): same fixtures as
    the F-FNO ones — generic-path shapes with odd coefficient counts, and width-64 shapes for the tcgen05 path."""
    import fourierflow.modules.factorized_cno as CNO
    grid2d_case(CNO, "cno_grid2d_w32", dict(modes=5, width=32, n_layers=2, input_dim=3, share_weight=False, factor=4,
                ff_weight_norm=True, gain=1), (2, 12, 10, 3), seed=30, cls="CNOFactorized2DBlock")
    grid2d_case(CNO, "cno_grid2d_w64", dict(modes=16, width=64, n_layers=3, input_dim=3, share_weight=True, factor=4,
                ff_weight_norm=True, gain=0.5), (1, 32, 32, 3), seed=31, cls="CNOFactorized2DBlock")
    mesh_case(CNO, "CNOFactorizedMesh2D", "cno_mesh2d_small", dict(modes_x=7, modes_y=4, width=32, input_dim=4,
              n_layers=2, share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False),
              (2, 11, 9, 2), seed=32)
    mesh_case(CNO, "CNOFactorizedMesh3D", "cno_mesh3d_w64", dict(modes_x=6, modes_y=5, modes_z=4, width=64,
              input_dim=4, output_dim=4, n_layers=2, share_weight=False, factor=4, ff_weight_norm=True,
              n_ff_layers=2, layer_norm=False), (1, 8, 8, 8, 1), seed=33)


def geo_case(M, name, kwargs, B, N, seed):
    """FNOFactorizedPointCloud2D on a random point cloud, iphi=None: the full forward, and the latent grid before /
    after the interior layers (point_cloud_2d.py:198-210) replayed with the module's own sub-modules."""
    from einops import rearrange
    torch.manual_seed(seed)
    m = M.FNOFactorizedPointCloud2D(**kwargs).eval()
    perturb_(m, seed + 100)
    g = torch.Generator().manual_seed(seed + 1)
    u = torch.rand(B, N, 2, generator=g)
    arrays = {"sd::" + k: (torch.view_as_real(v) if v.is_complex() else v).detach().numpy()
              for k, v in m.state_dict().items()
              if not (kwargs.get("share_weight") and re.match(r"convs\.\d+\.fourier_weight\.\d+$", k))}
    with torch.no_grad():
        arrays["u"] = u.numpy()
        arrays["out"] = m(u).numpy()
        grid = m.get_grid([B, m.s1, m.s2], u.device).permute(0, 3, 1, 2)
        v = m.fc0(u).permute(0, 2, 1)
        uc = m.convs[0](v, x_in=u, iphi=None, code=None, transform=False) + m.bs[0](grid)
        arrays["uc_in"] = rearrange(uc, "b c h w -> b h w c").contiguous().numpy()
        arrays["grid_bias"] = m.bs[0](grid)[0].permute(1, 2, 0).contiguous().numpy()
        for i in range(1, m.n_layers):
            uc1 = rearrange(m.convs[i](rearrange(uc, "b c h w -> b h w c"))[0], "b h w c -> b c h w")
            uc = uc + uc1 + m.bs[0](grid)
        arrays["uc_out"] = rearrange(uc, "b c h w -> b h w c").contiguous().numpy()
    save(name, kwargs, arrays)


def geo_grad_case(M, LpLoss, name, kwargs, B, N, seed):
    """Gradients through the geo-F-FNO: (i) of the training loss LpLoss(model(u), y) w.r.t. every parameter the forward
    uses (routines/point_cloud.py:29-43, iphi = None), and (ii) of <r, interior(uc, bias)> w.r.t. uc, the grid bias and
    the interior layers' parameters — the pin of ffno_layers_bwd — with the interior replayed from the module's own
    sub-modules (point_cloud_2d.py:198-210)."""
    from einops import rearrange
    torch.manual_seed(seed)
    m = M.FNOFactorizedPointCloud2D(**kwargs).train()
    perturb_(m, seed + 100)
    g = torch.Generator().manual_seed(seed + 1)
    u = torch.rand(B, N, 2, generator=g)
    y = torch.randn(B, N, 1, generator=g)
    arrays = {"sd::" + k: (torch.view_as_real(v) if v.is_complex() else v).detach().numpy() for k, v in m.state_dict().items()}
    arrays.update(u=u.numpy(), y=y.numpy())
    out = m(u)
    loss = LpLoss(size_average=True)(out.reshape(B, -1), y.reshape(B, -1))
    loss.backward()
    arrays["out"], arrays["loss"] = out.detach().numpy(), loss.detach().numpy()
    for k, p_ in m.named_parameters():
        if p_.grad is not None:
            gr = p_.grad.detach()
            arrays["grad::" + k] = (torch.view_as_real(gr) if gr.is_complex() else gr).numpy()
    m.zero_grad()
    uc = torch.randn(B, m.s1, m.s2, m.width, generator=g).requires_grad_(True)       # channels-last latent grid
    bias = (0.1 * torch.randn(m.s1, m.s2, m.width, generator=g)).requires_grad_(True)
    r = torch.randn(B, m.s1, m.s2, m.width, generator=g)
    h = uc
    for i in range(1, m.n_layers):
        h = h + m.convs[i](h)[0] + bias
    (h * r).sum().backward()
    arrays.update(uc_in=uc.detach().numpy(), grid_bias=bias.detach().numpy(), r=r.numpy(), uc_out=h.detach().numpy())
    arrays["igrad::uc"], arrays["igrad::bias"] = uc.grad.numpy(), bias.grad.numpy()
    for k, p_ in m.named_parameters():
        if p_.grad is not None and k.startswith("convs."):
            arrays["igrad::" + k] = p_.grad.detach().numpy()
    save(name, kwargs, arrays)


def geo_cases(M):
    geo_case(M, "geo_pointcloud_w32", dict(modes1=4, modes2=4, width=32, in_channels=2, out_channels=1, n_layers=4,
             s1=12, s2=10, share_weight=False), B=2, N=40, seed=40)
    geo_case(M, "geo_pointcloud_shared", dict(modes1=6, modes2=5, width=32, in_channels=2, out_channels=1,
             n_layers=3, s1=14, s2=12, share_weight=True), B=1, N=60, seed=41)


def plus_cases(M):
    """FNOPlus2DBlock (zongyi_fno/grid_plus_2d.py): un-factorized rfft2 spectral layer inside the F-FNO block structure."""
    grid2d_case(M, "plus2d_w32", dict(modes=4, width=32, n_layers=2, input_dim=3, share_weight=False, factor=4,
                ff_weight_norm=True, gain=1), (2, 12, 10, 3), seed=50, cls="FNOPlus2DBlock")
    grid2d_case(M, "plus2d_w64", dict(modes=6, width=64, n_layers=3, input_dim=3, share_weight=True, factor=4,
                ff_weight_norm=True, gain=0.5), (1, 16, 20, 3), seed=52, cls="FNOPlus2DBlock")      # tcgen05 FeedForward
    grid2d_case(M, "plus2d_shared_fork", dict(modes=3, width=16, n_layers=3, input_dim=2, share_weight=True,
                share_fork=True, use_fork=True, factor=2, ff_weight_norm=True, gain=0.5), (1, 6, 9, 2), seed=51,
                cls="FNOPlus2DBlock")


def init_case(M):
    """Seeded-construction checksums: the mirrors must draw the same random numbers in the same order."""
    cases = {
        "FNOFactorized2DBlock": [
            dict(modes=16, width=64, n_layers=3, input_dim=3, share_weight=True, factor=4,
                 ff_weight_norm=True, gain=0.1),
            dict(modes=4, width=32, n_layers=2, input_dim=5, share_weight=False, share_fork=True,
                 use_fork=True, factor=2, ff_weight_norm=False, gain=1, layer_norm=True),
        ],
        "FNOFactorizedMesh2D": [dict(modes_x=6, modes_y=4, width=32, input_dim=4, n_layers=2,
                                     share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2,
                                     layer_norm=False)],
        "FNOFactorizedMesh3D": [dict(modes_x=4, modes_y=3, modes_z=2, width=32, input_dim=4, output_dim=4,
                                     n_layers=2, share_weight=True, factor=4, ff_weight_norm=True,
                                     n_ff_layers=2, layer_norm=False)],
    }
    out = []
    for cls, kws in cases.items():
        for i, kw in enumerate(kws):
            torch.manual_seed(1234 + i)
            m = getattr(M, cls)(**kw)
            sd = m.state_dict()
            out.append({"cls": cls, "kwargs": kw, "seed": 1234 + i,
                        "keys": {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())]
                                 for k, v in sd.items()}})
    with open(os.path.join(OUT, "init_parity.json"), "w") as f:
        json.dump(out, f)
    print("init_parity.json written")


def grad_cases(M, LpLoss):
    # backward pins: shared weights + weight-norm (C2 architecture, reduced), and unshared weights without weight-norm
    grad_case(M, LpLoss, "grad_c2arch_16", dict(modes=8, width=64, n_layers=3, input_dim=3, share_weight=True,
              factor=4, ff_weight_norm=True, gain=0.1, dropout=0.0, in_dropout=0.0), (2, 16, 16, 3), seed=20)
    grad_case(M, LpLoss, "grad_unshared_w32", dict(modes=5, width=32, n_layers=2, input_dim=4, share_weight=False,
              factor=2, ff_weight_norm=False, gain=1), (2, 12, 10, 4), seed=21)
    mesh_grad_cases(M, LpLoss)


def cno_grad_cases(LpLoss):
    """Gradients of the DCT siblings' training loss (experiments/*/fcno): odd coefficient counts, shared weights on the
    periodic grid (width 64: tcgen05 adjoint), per-axis counts with padding and crop on the 2-D mesh (width 32: FP32)."""
    import fourierflow.modules.factorized_cno as CNO
    grad_case(CNO, LpLoss, "grad_cno_grid2d_w64", dict(modes=7, width=64, n_layers=2, input_dim=3, share_weight=True,
              factor=4, ff_weight_norm=True, gain=0.5), (2, 16, 16, 3), seed=24, cls="CNOFactorized2DBlock")
    mesh_grad_case(CNO, LpLoss, "CNOFactorizedMesh2D", "grad_cno_mesh2d_w32", dict(modes_x=5, modes_y=4, width=32,
                   input_dim=4, n_layers=2, share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2,
                   layer_norm=False), (2, 10, 9, 2), 1, seed=25)


def plus_grad_cases(M, LpLoss):
    """Gradients of the un-factorized sibling's training loss (experiments/torus_li/ablation/no_factorization*): 5-D mode
    weights of the two row blocks, shared over the layers (width 32: FP32 kernels throughout; width 64: the forward's
    FeedForward runs on tcgen05)."""
    grad_case(M, LpLoss, "grad_plus2d_w32", dict(modes=4, width=32, n_layers=2, input_dim=3, share_weight=True, factor=4,
              ff_weight_norm=True, gain=0.5), (2, 12, 10, 3), seed=26, cls="FNOPlus2DBlock")
    grad_case(M, LpLoss, "grad_plus2d_w64", dict(modes=3, width=64, n_layers=2, input_dim=3, share_weight=False, factor=4,
              ff_weight_norm=True, gain=0.5), (2, 8, 8, 3), seed=27, cls="FNOPlus2DBlock")


def mesh_grad_cases(M, LpLoss):
    # mesh variants: grid append + padding + crop; width 64 (tcgen05 path) in 3-D, width 32 (FP32 path) in 2-D
    mesh_grad_case(M, LpLoss, "FNOFactorizedMesh3D", "grad_mesh3d_w64", dict(modes_x=6, modes_y=6, modes_z=4, width=64,
                   input_dim=4, output_dim=4, n_layers=2, share_weight=True, factor=4, ff_weight_norm=True,
                   n_ff_layers=2, layer_norm=False), (2, 8, 8, 8, 1), 4, seed=22)
    mesh_grad_case(M, LpLoss, "FNOFactorizedMesh2D", "grad_mesh2d_w32", dict(modes_x=6, modes_y=4, width=32, input_dim=4,
                   n_layers=2, share_weight=False, factor=4, ff_weight_norm=True, n_ff_layers=2, layer_norm=False),
                   (2, 12, 10, 2), 1, seed=23)


def rollout_extras_cases(M, LpLoss, c2):
    rollout_case(M, LpLoss, "rollout_force_mu_16", dict(c2, n_layers=4, modes=8, input_dim=5), B=2, X=16, T=12,
                 n_steps=10, seed=12, force_dims=4, with_mu=True)
    rollout_case(M, LpLoss, "rollout_force_static_16", dict(c2, n_layers=4, modes=8, input_dim=4), B=2, X=16, T=12,
                 n_steps=10, seed=13, force_dims=3)


def main():
    M, LpLoss = import_reference()
    if "--only-grad" in sys.argv:
        grad_cases(M, LpLoss)
        return
    c2 = dict(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
              ff_weight_norm=True, gain=0.1, dropout=0.0, in_dropout=0.0)
    if "--only-plus-grad" in sys.argv:
        plus_grad_cases(M, LpLoss)
        return
    if "--only-cno-grad" in sys.argv:
        cno_grad_cases(LpLoss)
        return
    if "--only-mesh-grad" in sys.argv:
        mesh_grad_cases(M, LpLoss)
        return
    if "--only-plus" in sys.argv:
        plus_cases(M)
        return
    if "--only-geo" in sys.argv:
        geo_cases(M)
        return
    if "--only-geo-grad" in sys.argv:
        geo_grad_case(M, LpLoss, "grad_geo_pointcloud", dict(modes1=4, modes2=4, width=32, in_channels=2, out_channels=1,
                      n_layers=4, s1=12, s2=10, share_weight=False), B=2, N=40, seed=42)
        return
    if "--only-cno" in sys.argv:
        cno_cases()
        return
    if "--only-rollout-ablations" in sys.argv:
        ablation_rollout_cases(M, LpLoss, c2)
        return
    if "--only-rollout-extras" in sys.argv:
        rollout_extras_cases(M, LpLoss, c2)
        return
    init_case(M)
    # (1) the C2 architecture on a reduced grid (32x32 keeps modes=16 legal: 17 rfft bins)
    grid2d_case(M, "grid2d_c2arch_32", dict(c2, n_layers=4), (1, 32, 32, 3), seed=0)
    # (2) full-depth C2 model (24 layers), final forecast + last-layer taps only
    grid2d_case(M, "grid2d_c2_24layers_32", c2, (1, 32, 32, 3), seed=1, taps=False)
    # (3) stress: gain=1, unshared weights, larger activations, C4-like input_dim=5, non-square
    grid2d_case(M, "grid2d_gain1_unshared", dict(modes=8, width=64, n_layers=3, input_dim=5,
                share_weight=False, factor=4, ff_weight_norm=True, gain=1), (2, 16, 24, 5),
                seed=2, x_scale=3.0)
    # (4) generic-path options: no weight-norm, LayerNorm, factor 2, width 32, non-pow2 sizes
    grid2d_case(M, "grid2d_ln_w32", dict(modes=5, width=32, n_layers=2, input_dim=4,
                share_weight=False, factor=2, ff_weight_norm=False, gain=1, layer_norm=True),
                (2, 20, 12, 4), seed=3)
    # (5) fork ablation + shared fork
    grid2d_case(M, "grid2d_fork", dict(modes=4, width=32, n_layers=2, input_dim=3,
                share_weight=True, share_fork=True, use_fork=True, factor=4,
                ff_weight_norm=True, gain=0.5), (1, 16, 16, 3), seed=4)
    # (6) mode switches
    for md in ("low-pass", "no-fourier"):
        grid2d_case(M, "grid2d_" + md.replace("-", ""), dict(modes=4, width=32, n_layers=2,
                    input_dim=3, share_weight=False, factor=4, ff_weight_norm=True, gain=1,
                    mode=md), (1, 16, 16, 3), seed=5)
    # (7) n_modes == L/2+1 on an even length (Nyquist bin kept) and an odd length
    grid2d_case(M, "grid2d_nyquist", dict(modes=5, width=32, n_layers=1, input_dim=2,
                share_weight=False, factor=4, ff_weight_norm=True, gain=1), (1, 8, 9, 2), seed=6)
    # (8) mesh variants: pad +8, linspace grid, per-axis modes, prime lengths after padding
    mesh_case(M, "FNOFactorizedMesh2D", "mesh2d_small", dict(modes_x=6, modes_y=4, width=32,
              input_dim=4, n_layers=2, share_weight=False, factor=4, ff_weight_norm=True,
              n_ff_layers=2, layer_norm=False), (2, 11, 9, 2), seed=7)
    mesh_case(M, "FNOFactorizedMesh3D", "mesh3d_small", dict(modes_x=4, modes_y=3, modes_z=2,
              width=32, input_dim=4, output_dim=4, n_layers=2, share_weight=False, factor=4,
              ff_weight_norm=True, n_ff_layers=2, layer_norm=False), (1, 7, 6, 5, 1), seed=8)
    mesh_case(M, "FNOFactorizedMesh3D", "mesh3d_w64", dict(modes_x=6, modes_y=6, modes_z=4,
              width=64, input_dim=4, output_dim=4, n_layers=2, share_weight=True, factor=4,
              ff_weight_norm=True, n_ff_layers=2, layer_norm=False), (1, 8, 8, 8, 1), seed=9)
    # (9) one bare spectral layer + FF at the exact C2 layer shape
    spectral_case(M, "spectral_c2_layer", 64, 16, (1, 32, 64, 64), seed=10)
    # (10) 10-step Markov rollout with Normalizer / LpLoss
    rollout_case(M, LpLoss, "rollout_c2arch_16", dict(c2, n_layers=4, modes=8), B=2, X=16, T=12,
                 n_steps=10, seed=11)
    # (10b) the torus_vis feature sets: append_force (static and time-varying forcing) and append_mu
    rollout_extras_cases(M, LpLoss, c2)
    ablation_rollout_cases(M, LpLoss, c2)
    # (10c) the factorized cosine (DCT) siblings
    cno_cases()
    # (10d) geo-F-FNO on point clouds (interior layers = the hot path)
    geo_cases(M)
    # (10e) the un-factorized sibling
    plus_cases(M)
    # (11) gradients of the one-step training loss (backward row)
    grad_cases(M, LpLoss)


if __name__ == "__main__":
    main()
