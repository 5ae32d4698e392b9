"""CPU oracle for the F-FNO layer-stack forward — TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU ops, fp32 or fp64) of the reference's
algorithm for the hot path named in BASELINE.json.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``fourierflow_b200/`` imports it, and the product path has no CPU
fallback.

Parity status: PINNED against the reference itself.  ``oracle/make_golden.py`` imports the
unmodified reference modules from /root/reference (stub-package recipe), runs
them on seeded inputs and commits the outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks every function below against those fixtures.  (The reference ships no tests or golden
vectors for this path — SURVEY.md §4 — so the executed reference is the only pin there is.)

Every function cites the reference file:line (relative to /root/reference) it restates.
Parameters are passed as a flat ``dict[str, Tensor]`` using the reference's own ``state_dict``
keys, so a reference checkpoint feeds the oracle unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# WNLinear / FeedForward
# ----------------------------------------------------------------------------------------------

def weight_norm_fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Effective weight of a weight-normalised linear: ``w = g * v / ||v||_2`` (norm per output row).

    Reference: fourierflow/modules/linear.py:41-51 (``weight_norm(self)`` → torch's
    ``_weight_norm(v, g, dim=0)``, recomputed by a forward pre-hook on every call).
    """
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def linear_weight(p: Params, prefix: str) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(weight, bias) of the WNLinear stored under ``prefix`` — plain or weight-normalised keys.

    Reference: fourierflow/modules/linear.py:41-51; key schema SURVEY.md §8(b).
    """
    if prefix + "weight_g" in p:
        w = weight_norm_fold(p[prefix + "weight_g"], p[prefix + "weight_v"])
    else:
        w = p[prefix + "weight"]
    return w, p.get(prefix + "bias")


def feed_forward(p: Params, prefix: str, x: torch.Tensor, n_layers: int = 2,
                 layer_norm: bool = False) -> torch.Tensor:
    """``FeedForward.forward``: n_layers × [WNLinear → (Dropout p=0) → ReLU (all but last) →
    LayerNorm (last only, optional)].

    Reference: fourierflow/modules/feedforward.py:6-24.
    """
    for i in range(n_layers):
        w, b = linear_weight(p, f"{prefix}layers.{i}.0.")
        x = F.linear(x, w, b)
        if i < n_layers - 1:
            x = torch.relu(x)
        elif layer_norm:
            x = F.layer_norm(x, (x.shape[-1],), p[f"{prefix}layers.{i}.3.weight"],
                             p[f"{prefix}layers.{i}.3.bias"])
    return x


# ----------------------------------------------------------------------------------------------
# Spectral convolution along one axis
# ----------------------------------------------------------------------------------------------

def dct_matrix(L: int, dtype=torch.float64) -> torch.Tensor:
    """Ortho DCT-II matrix ``D[j, l] = c_j cos(pi (2 l + 1) j / (2 L))``, c_0 = sqrt(1/L), c_j = sqrt(2/L): what
    ``dct(x, norm='ortho')`` of fourierflow/modules/dct.py:16-45 computes through an FFT (Makhoul); its transpose is
    the ``idct(X, norm='ortho')`` of dct.py:48-88 (DCT-III)."""
    l = torch.arange(L, dtype=torch.float64)
    j = torch.arange(L, dtype=torch.float64)[:, None]
    D = torch.cos(math.pi * (2 * l + 1) * j / (2 * L)) * math.sqrt(2.0 / L)
    D[0] = D[0] / math.sqrt(2.0)
    return D.to(dtype)


def spectral_axis_dct(x: torch.Tensor, w: torch.Tensor, dim: int, n_modes: int) -> torch.Tensor:
    """One axis of the factorized cosine operator on channels-last ``x[..., C]``: ortho DCT-II along ``dim`` → keep the
    first ``n_modes`` coefficients → REAL per-coefficient channel mix ``R[.., j, o] = sum_i X[.., j, i] W[i, o, j]`` →
    zero-pad to L → DCT-III.  Reference: fourierflow/modules/factorized_cno/grid_2d.py:57-69 (last axis) / :72-86;
    mesh_3d.py:63-74, :77-91, :94-108."""
    L = x.shape[dim]
    D = dct_matrix(L, x.dtype)[:n_modes]                # [K, L]
    X = torch.einsum("kl,...lc->...kc", D, x.movedim(dim, -2))
    R = torch.einsum("...ki,iok->...ko", X, w)
    return torch.einsum("kl,...kc->...lc", D, R).movedim(-2, dim)


def spectral_axis(x: torch.Tensor, w: torch.Tensor, dim: int, n_modes: int,
                  mode: str = "full") -> torch.Tensor:
    """One axis of ``forward_fourier`` on a channels-LAST tensor ``x[..., C]``.

    rfft(ortho) along ``dim`` → keep the lowest ``n_modes`` bins → per-mode complex channel mix
    ``R[.., k, o] = sum_i X[.., k, i] * W[i, o, k]`` → zero-pad to L//2+1 → irfft(n=L, ortho).

    Reference: fourierflow/modules/factorized_fno/grid_2d.py:57-72 (Y axis) / :75-90 (X axis);
    mesh_3d.py:61-73, 76-88, 91-103.  The reference permutes to channels-first and transforms
    dim -1/-2/-3; here the same transform is applied along the matching channels-last axis.
    """
    if w is not None and w.dim() == 3:                 # real [in, out, K] weights: the factorized_cno sibling
        return spectral_axis_dct(x, w, dim, n_modes)
    L = x.shape[dim]
    x_ft = torch.fft.rfft(x, dim=dim, norm="ortho")
    x_ft = x_ft.movedim(dim, -2)                       # [..., L/2+1, C]
    out_ft = torch.zeros_like(x_ft)
    if mode == "full":
        wc = torch.view_as_complex(w.contiguous())     # [in, out, K]
        out_ft[..., :n_modes, :] = torch.einsum("...ki,iok->...ko", x_ft[..., :n_modes, :], wc)
    elif mode == "low-pass":
        out_ft[..., :n_modes, :] = x_ft[..., :n_modes, :]
    else:
        raise ValueError(mode)
    out_ft = out_ft.movedim(-2, dim)
    return torch.fft.irfft(out_ft, n=L, dim=dim, norm="ortho")


def forward_fourier_grid2d(x: torch.Tensor, w0: torch.Tensor, w1: torch.Tensor, n_modes: int,
                           mode: str = "full") -> torch.Tensor:
    """``SpectralConv2d.forward_fourier`` of the periodic-grid F-FNO, x:[B,M,N,C].

    NOTE the index↔axis mapping: ``fourier_weight[0]`` acts on the LAST spatial axis (N, "Y"),
    ``fourier_weight[1]`` on M ("X").  Reference: factorized_fno/grid_2d.py:51-99 (:65-68, :83-86).
    """
    xy = spectral_axis(x, w0, 2, n_modes, mode)
    xx = spectral_axis(x, w1, 1, n_modes, mode)
    return xx + xy


def forward_fourier_plus2d(x: torch.Tensor, w0: torch.Tensor, w1: torch.Tensor, n_modes: int) -> torch.Tensor:
    """``SpectralConv2d.forward_fourier`` of the un-factorized sibling (FNOPlus2DBlock), x:[B,M,N,C] channels-last:
    rfft2(ortho) -> rows ``:K`` x columns ``:K`` mixed with ``w0``, rows ``-K:`` x columns ``:K`` with ``w1`` (both
    [in, out, K, K, 2]) -> everything else zero -> irfft2(s=(M, N), ortho).
    Reference: fourierflow/modules/zongyi_fno/grid_plus_2d.py:52-83."""
    B, M, N, C = x.shape
    K = n_modes
    x_ft = torch.fft.rfft2(x, s=(M, N), dim=(1, 2), norm="ortho")              # [B, M, N/2+1, C]
    out_ft = torch.zeros_like(x_ft)
    out_ft[:, :K, :K] = torch.einsum("bxyi,ioxy->bxyo", x_ft[:, :K, :K], torch.view_as_complex(w0.contiguous()))
    out_ft[:, -K:, :K] = torch.einsum("bxyi,ioxy->bxyo", x_ft[:, -K:, :K], torch.view_as_complex(w1.contiguous()))
    return torch.fft.irfft2(out_ft, s=(M, N), dim=(1, 2), norm="ortho")


def forward_fourier_mesh(x: torch.Tensor, ws: Sequence[torch.Tensor], modes: Sequence[int],
                         mode: str = "full") -> torch.Tensor:
    """``forward_fourier`` of the mesh variants: ``fourier_weight[a]`` acts on spatial axis ``a``
    (0→X, 1→Y, 2→Z), per-axis mode counts; result is the sum over axes.

    Reference: factorized_fno/mesh_2d.py:56-104 (:70-73, :88-91); mesh_3d.py:55-112.
    """
    nd = x.dim() - 2
    out = None
    for a in reversed(range(nd)):                       # reference order: last axis first
        t = spectral_axis(x, ws[a], 1 + a, modes[a], mode)
        out = t if out is None else out + t
    return out


# ----------------------------------------------------------------------------------------------
# Block containers
# ----------------------------------------------------------------------------------------------

def _layer_weights(p: Params, l: int, n_axes: int) -> List[torch.Tensor]:
    # share_weight=True registers the same ParameterList on the block and on every layer, so the
    # per-layer keys always exist (SURVEY.md §7 hard part 5).
    return [p[f"spectral_layers.{l}.fourier_weight.{a}"] for a in range(n_axes)]


def _head(p: Params, b: torch.Tensor) -> torch.Tensor:
    # out = Sequential(WNLinear(C,128), WNLinear(128,out_dim)), NO activation in between.
    # Reference: grid_2d.py:150-152; mesh_3d.py:156-158.
    w0, b0 = linear_weight(p, "out.0.")
    w1, b1 = linear_weight(p, "out.1.")
    return F.linear(F.linear(b, w0, b0), w1, b1)


def block_grid2d_forward(p: Params, x: torch.Tensor, *, modes: int, n_layers: int,
                         n_ff_layers: int = 2, layer_norm: bool = False, use_fork: bool = False,
                         mode: str = "full", taps: Optional[dict] = None) -> dict:
    """``FNOFactorized2DBlock.forward``: lift → L × {spectral → FF → residual} → head(last b).

    Reference: factorized_fno/grid_2d.py:154-177 (+ ``SpectralConv2d.forward`` :42-49).
    ``taps`` (optional dict) receives per-layer intermediates for per-layer parity:
    ``lift``, ``s{l}`` (forward_fourier output), ``b{l}``, ``x{l}`` (after the residual).
    """
    w, b = linear_weight(p, "in_proj.")
    x = F.linear(x, w, b)                                # :157 (drop p=0 :158)
    if taps is not None:
        taps["lift"] = x
    forecast = 0
    forecast_list = []
    bb = None
    for l in range(n_layers):
        s = x
        if mode != "no-fourier":                         # :44-45
            w0, w1 = _layer_weights(p, l, 2)
            if w0.dim() == 5:                            # [in, out, K, K, 2]: the un-factorized FNOPlus2DBlock
                s = forward_fourier_plus2d(x, w0, w1, modes)
            else:
                s = forward_fourier_grid2d(x, w0, w1, modes, mode)
        bb = feed_forward(p, f"spectral_layers.{l}.backcast_ff.", s, n_ff_layers, layer_norm)
        if use_fork:                                     # :48, :164-167
            f = feed_forward(p, f"spectral_layers.{l}.forecast_ff.", s, n_ff_layers, layer_norm)
            f_out = _head(p, f)
            forecast = forecast + f_out
            forecast_list.append(f_out)
        x = x + bb                                       # :169
        if taps is not None:
            taps[f"s{l}"], taps[f"b{l}"], taps[f"x{l}"] = s, bb, x
    if not use_fork:
        forecast = _head(p, bb)                          # :171-172 — head on last b, not on x
    return {"forecast": forecast, "forecast_list": forecast_list}


def geo_interior_forward(p: Params, uc: torch.Tensor, grid_bias: torch.Tensor, *, modes: int,
                         n_layers: int) -> torch.Tensor:
    """Interior of the geo-F-FNO (``FNOFactorizedPointCloud2D``): for i = 1 .. n_layers-1, on the channels-last latent
    grid, ``uc = uc + backcast_ff_i(forward_fourier_i(uc)) + bs[0](grid)``; the interior layers are the periodic-grid
    SpectralConv2d with factor 2.  Reference: factorized_fno/point_cloud_2d.py:198-210 (layers built at :170-183)."""
    for i in range(1, n_layers):
        w0, w1 = p[f"convs.{i}.fourier_weight.0"], p[f"convs.{i}.fourier_weight.1"]
        s = forward_fourier_grid2d(uc, w0, w1, modes)
        uc = uc + feed_forward(p, f"convs.{i}.backcast_ff.", s) + grid_bias
    return uc


def mesh_grid_features(shape: Sequence[int], dtype=torch.float32) -> torch.Tensor:
    """``get_grid``: per-axis ``np.linspace(0, 1, size)`` (inclusive) cast to float32, broadcast to
    [B, *spatial, ndim].  Reference: mesh_3d.py:178-189; mesh_2d.py:167-175."""
    B, *sp = shape
    nd = len(sp)
    feats = []
    for a, n in enumerate(sp):
        g = torch.tensor(np.linspace(0, 1, n), dtype=torch.float)
        view = [1] * (nd + 2)
        view[1 + a] = n
        feats.append(g.reshape(view).expand(B, *sp, 1))
    return torch.cat(feats, dim=-1).to(dtype)


def block_mesh_forward(p: Params, x: torch.Tensor, *, modes: Sequence[int], n_layers: int,
                       n_ff_layers: int = 2, layer_norm: bool = False, padding: int = 8,
                       taps: Optional[dict] = None) -> torch.Tensor:
    """``FNOFactorizedMesh2D/3D.forward``: append linspace grid → lift → zero-pad +8 on the high
    side of every spatial axis → L × {spectral → FF → residual} → crop last b → head.

    Reference: mesh_3d.py:160-176; mesh_2d.py:149-165.
    """
    nd = x.dim() - 2
    grid = mesh_grid_features(x.shape[:-1], x.dtype)
    x = torch.cat((x, grid), dim=-1)
    w, b = linear_weight(p, "in_proj.")
    x = F.linear(x, w, b)
    pad = []
    for _ in range(nd):
        pad += [0, padding]
    x = F.pad(x.movedim(-1, 1), pad).movedim(1, -1)      # pad spatial dims, high side only
    if taps is not None:
        taps["lift"] = x
    bb = None
    for l in range(n_layers):
        ws = _layer_weights(p, l, nd)
        s = forward_fourier_mesh(x, ws, modes)
        bb = feed_forward(p, f"spectral_layers.{l}.backcast_ff.", s, n_ff_layers, layer_norm)
        x = x + bb
        if taps is not None:
            taps[f"s{l}"], taps[f"b{l}"], taps[f"x{l}"] = s, bb, x
    sl = (slice(None),) + (slice(0, -padding),) * nd + (slice(None),)
    return _head(p, bb[sl])


# ----------------------------------------------------------------------------------------------
# Rollout glue: Normalizer, LpLoss, Markov rollout
# ----------------------------------------------------------------------------------------------

def normalizer_stats(xs: torch.Tensor) -> dict:
    """Accumulate ``Normalizer`` statistics from one training-mode pass over ``xs[..., H]``.

    Reference: fourierflow/modules/normalizer.py:18-26 (``_accumulate`` over pooled points)."""
    flat = xs.reshape(-1, xs.shape[-1])
    return {"sum": flat.sum(0), "sum_squared": (flat ** 2).sum(0),
            "count": torch.tensor(float(flat.shape[0]), dtype=xs.dtype)}


def normalizer_mean_std(st: dict, std_epsilon: float = 1e-8) -> Tuple[torch.Tensor, torch.Tensor]:
    """``Normalizer.mean`` / ``.std``.  Reference: normalizer.py:68-77."""
    safe = torch.maximum(st["count"], torch.ones_like(st["count"]))
    mean = st["sum"] / safe
    std = torch.sqrt(st["sum_squared"] / safe - mean ** 2)
    return mean, torch.maximum(std, torch.full_like(std, std_epsilon))


def lp_loss_rel(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """``LpLoss.rel`` (p=2, reduction mean): mean_b ||x_b − y_b||₂ / ||y_b||₂.

    Reference: fourierflow/modules/loss.py:33-46."""
    n = x.shape[0]
    diff = torch.norm(x.reshape(n, -1) - y.reshape(n, -1), 2, 1)
    return torch.mean(diff / torch.norm(y.reshape(n, -1), 2, 1))


def position_features(dim_sizes: Sequence[int], low: float = 0.0, high: float = 1.0,
                      dtype=torch.float32) -> torch.Tensor:
    """``encode_positions(fourier=False)``: meshgrid('ij') of ``linspace(low, high, size)``.

    Reference: fourierflow/routines/grid_2d_markov.py:98-112."""
    grids = [torch.linspace(low, high, steps=s, dtype=dtype) for s in dim_sizes]
    return torch.stack(torch.meshgrid(*grids, indexing="ij"), dim=-1)


def rfft_mesh(shape: Sequence[int], domain=((0.0, 2 * math.pi), (0.0, 2 * math.pi))):
    """``jax_cfd.base.grids.Grid(shape, domain).rfft_mesh()`` — THIRD-PARTY arithmetic restated (jax-cfd rev eb4d723,
    pyproject.toml:29 of the reference; the package is not installed here, so this one function is *unpinned* by
    execution and follows its published definition):
        kx, ky = meshgrid(fftfreq(nx, d=step_x), rfftfreq(ny, d=step_y), indexing='ij'),  step = domain length / n.
    Used only to build the ``kx/ky/lap`` buffers of the velocity features (routines/grid_2d_markov.py:82-93)."""
    (x0, x1), (y0, y1) = domain
    nx, ny = shape
    fx = torch.fft.fftfreq(nx, d=(x1 - x0) / nx, dtype=torch.float64)
    fy = torch.fft.rfftfreq(ny, d=(y1 - y0) / ny, dtype=torch.float64)
    return torch.meshgrid(fx, fy, indexing="ij")


def velocity_features(w: torch.Tensor, domain=((0.0, 2 * math.pi), (0.0, 2 * math.pi))) -> Tuple[torch.Tensor, torch.Tensor]:
    """(q, v) = (psi_y, -psi_x) of the stream function of vorticity ``w[B, X, Y, ...]`` (transform over dims 1, 2).

    Reference: fourierflow/routines/grid_2d_markov.py:82-93 (buffers: lap = (2 pi i)^2 (|kx|^2 + |ky|^2), lap[0,0] = 1)
    and :206-220 / :268-285 (rfftn 'backward' -> psi_hat = -w_hat / lap -> 2 pi i ky psi_hat, -2 pi i kx psi_hat
    -> irfftn).  Buffers are cast to the data precision exactly as ``register_buffer(torch.from_numpy(...))`` of the
    float32 / complex64 jax arrays does."""
    X, Y = w.shape[1], w.shape[2]
    kx, ky = rfft_mesh((X, Y), domain)
    lap = (2j * math.pi) ** 2 * (kx.abs() ** 2 + ky.abs() ** 2)
    lap[0, 0] = 1
    rdt = w.dtype
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    kx, ky, lap = kx.to(rdt), ky.to(rdt), lap.to(cdt)
    extra = (1,) * (w.dim() - 3)
    shp = (1, X, Y // 2 + 1) + extra
    w_hat = torch.fft.rfftn(w, dim=[1, 2], norm="backward")
    psi_hat = -w_hat / lap.reshape(shp)
    q = torch.fft.irfftn(2 * math.pi * 1j * ky.reshape(shp) * psi_hat, s=(X, Y), dim=[1, 2], norm="backward")
    v = torch.fft.irfftn(-2 * math.pi * 1j * kx.reshape(shp) * psi_hat, s=(X, Y), dim=[1, 2], norm="backward")
    return q, v


def markov_rollout(p: Params, data: torch.Tensor, stats: dict, *, modes: int, n_layers: int,
                   n_steps: int = 10, low: float = 0.0, high: float = 1.0, use_velocity: bool = False,
                   domain=((0.0, 2 * math.pi), (0.0, 2 * math.pi)), force: torch.Tensor = None,
                   mu: torch.Tensor = None, use_position: bool = True, shuffle=None,
                   learn_difference: bool = False) -> dict:
    """``Grid2DMarkovExperiment._valid_step`` with the torus_li/markov config
    (use_position, should_normalize; no force/mu/shuffle/difference) and, with ``use_velocity``, the torus_kochkov
    feature set [w, q, v, gx, gy] (velocity recomputed from every fed-back forecast, :268-285).

    data:[B,X,Y,T].  Step 0 consumes the ground-truth frame T−n−1; later steps feed back the
    model's own de-normalised forecast, re-concatenated with the position grid.
    ``force`` (append_force, :246-255 / :288-289): [B,X,Y] static forcing or [B,X,Y,T'] whose last n_steps
    frames are used, appended after the position grid; ``mu`` (append_mu, :257-260 / :290-291): [B] viscosity
    broadcast over the grid, appended last.
    Ablation switches: ``use_position=False`` drops the grid features; ``shuffle=(x_idx, y_idx)`` permutes rows / columns
    of the normalised input and undoes it on the forecast (:297-304); ``learn_difference`` treats the forecast as an
    increment — the loss target is yy[t] - yy[t-1] (index -1 at t = 0, as the reference writes it, :309-310) and the
    prediction accumulates from the first input frame (:316-318).
    Reference: fourierflow/routines/grid_2d_markov.py:195-326 (loop :263-321).
    """
    B, X, Y, T = data.shape
    pos = position_features((X, Y), low, high, data.dtype).unsqueeze(0).expand(B, X, Y, 2)
    mean, std = normalizer_mean_std(stats)
    n_steps = n_steps or T - 1
    yy = data[..., -n_steps:]
    loss = 0
    step_losses, preds = [], []
    im = data[..., T - n_steps - 1].unsqueeze(-1)         # :236 / :265
    prev_im = im                                          # :266
    for t in range(n_steps):
        if use_velocity:
            q, v = velocity_features(im, domain)          # :206-220 / :268-285
            im = torch.cat([im, q, v], dim=-1)
        x = torch.cat([im, pos], dim=-1) if use_position else im      # :222-233 / :286-287
        if force is not None:                             # :246-255 / :288-289
            f_t = force if force.dim() == 3 else force[..., -n_steps:][..., t]
            x = torch.cat([x, f_t.unsqueeze(-1)], dim=-1)
        if mu is not None:                                # :257-260 / :290-291
            x = torch.cat([x, mu.reshape(B, 1, 1, 1).expand(B, X, Y, 1)], dim=-1)
        x = (x - mean) / std                              # :296, normalizer.py:51
        if shuffle is not None:                           # :297-298
            x = x[:, shuffle[0]][:, :, shuffle[1]]
        im = block_grid2d_forward(p, x, modes=modes, n_layers=n_layers)["forecast"]   # :300-301
        if shuffle is not None:                           # :303-304
            im = im[:, :, torch.argsort(shuffle[1])][:, torch.argsort(shuffle[0])]
        im = im * std[0] + mean[0]                        # :306, normalizer.py:62
        y = yy[..., t] - yy[..., t - 1] if learn_difference else yy[..., t]           # :309-312
        l = lp_loss_rel(im.reshape(B, -1), y.reshape(B, -1))                          # :313
        step_losses.append(l)
        loss = loss + l
        if learn_difference:                              # :316-318
            im = prev_im + im
            prev_im = im
        preds.append(im)
    return {"loss": loss, "step_losses": torch.stack(step_losses),
            "preds": torch.cat(preds, dim=-1)}            # :319


# ----------------------------------------------------------------------------------------------
# Independent restatement of the transform pair as explicit real matrices (float64).
# These are the matrices the CUDA library builds on the host (csrc/dft_tables.h); the CPU tests
# check them against torch.fft so the exact-semantics traps are pinned:
#   * norm='ortho' scales BOTH directions by 1/sqrt(L)            (grid_2d.py:58,72)
#   * the inverse is a C2R: Im(DC) (and Im(Nyquist) for even L) is discarded, bins k>=1 count twice
# ----------------------------------------------------------------------------------------------

def dft_forward_matrix(L: int, K: int) -> np.ndarray:
    """D[2K, L]: row 2k = Re, row 2k+1 = Im of the ortho rfft bin k (k < K)."""
    l = np.arange(L)[None, :]
    k = np.arange(K)[:, None]
    ang = 2.0 * np.pi * k * l / L
    D = np.empty((2 * K, L))
    D[0::2] = np.cos(ang) / math.sqrt(L)
    D[1::2] = -np.sin(ang) / math.sqrt(L)
    return D


def dft_inverse_matrix(L: int, K: int) -> np.ndarray:
    """E[L, 2K] such that irfft(zero-padded K bins, n=L, ortho) = E @ [Re R0, Im R0, Re R1, ...]."""
    l = np.arange(L)[:, None]
    k = np.arange(K)[None, :]
    ang = 2.0 * np.pi * k * l / L
    c = np.full((1, K), 2.0)
    c[0, 0] = 1.0
    if L % 2 == 0 and K == L // 2 + 1:
        c[0, K - 1] = 1.0
    E = np.empty((L, 2 * K))
    E[:, 0::2] = c * np.cos(ang) / math.sqrt(L)
    E[:, 1::2] = -c * np.sin(ang) / math.sqrt(L)
    return E
