"""Host-side glue between the nn.Module mirrors and libffno_b200's C ABI.

PyTorch supplies device memory and the current stream; every arithmetic step of the forward runs in
the CUDA library.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib


# ----------------------------------------------------------------------------------------------
# checks
# ----------------------------------------------------------------------------------------------

def require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: fourierflow_b200 runs on CUDA (sm_100a) only — got a {x.device} tensor; "
                           "there is no CPU fallback")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {x.dtype}")


def needs_grad(module: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


def require_inference(module: nn.Module, x: torch.Tensor) -> None:
    """Entry points without a CUDA backward (sub-modules called on their own, mesh variants, rollouts: the reference's
    timed path runs under no_grad, fourierflow/routines/base.py:54-56) refuse to silently drop a graph.  The block
    forward of FNOFactorized2DBlock is differentiable (ffno_block_bwd)."""
    if needs_grad(module, x):
        raise RuntimeError(
            f"{type(module).__name__}: this entry point of the B200 backend has the forward pass only; call it under "
            "torch.no_grad() / torch.inference_mode() (the backward exists for FNOFactorized2DBlock.forward)")


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dev_param(t: torch.Tensor, keep: list) -> int:
    """Pointer of a parameter as contiguous fp32 device memory (keeps temporaries alive in `keep`)."""
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    keep.append(t)
    return t.data_ptr()


def linear_params(lin: nn.Linear, keep: list) -> _lib.LinearParams:
    lp = _lib.LinearParams()
    if "weight" in lin._parameters and lin._parameters["weight"] is not None:
        lp.weight = _dev_param(lin._parameters["weight"], keep)
    else:
        lp.weight_g = _dev_param(lin.weight_g, keep)
        lp.weight_v = _dev_param(lin.weight_v, keep)
    if lin.bias is not None:
        lp.bias = _dev_param(lin.bias, keep)
    lp.in_features, lp.out_features = lin.in_features, lin.out_features
    return lp


# ----------------------------------------------------------------------------------------------
# stateless ops (sub-modules called on their own)
# ----------------------------------------------------------------------------------------------

def linear_forward(lin: nn.Linear, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    require_cuda(x, "WNLinear.forward")
    require_inference(lin, x)
    if x.dim() < 1 or x.shape[-1] != lin.in_features:     # nn.Linear raises; a reshape(-1, in) would reinterpret rows
        raise RuntimeError(f"WNLinear.forward: last dimension of the input is {tuple(x.shape)[-1:]}, "
                           f"the layer has in_features={lin.in_features}")
    lib = _lib.load()
    keep: list = []
    with torch.cuda.device(x.device):
        lp = linear_params(lin, keep)
        x2 = x.contiguous().reshape(-1, lin.in_features)
        y = torch.empty(x2.shape[0], lin.out_features, device=x.device, dtype=torch.float32)
        scratch = torch.empty(lin.in_features * lin.out_features, device=x.device, dtype=torch.float32)
        _lib.check(lib.ffno_linear_fwd(C.byref(lp), x2.data_ptr(), x2.shape[0], y.data_ptr(), int(relu),
                                       scratch.data_ptr(), scratch.numel() * 4, _stream(x.device)),
                   "ffno_linear_fwd")
    return y.reshape(*x.shape[:-1], lin.out_features)


def layernorm_forward(ln: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, "LayerNorm")
    if len(ln.normalized_shape) != 1 or x.dim() < 1 or x.shape[-1] != ln.normalized_shape[-1]:
        raise RuntimeError(f"LayerNorm: input {tuple(x.shape)} does not end in normalized_shape={tuple(ln.normalized_shape)} "
                           "(the B200 backend normalises the last dimension only)")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        x2 = x.contiguous().reshape(-1, x.shape[-1])
        y = torch.empty_like(x2)
        w, b = ln.weight.detach().contiguous(), ln.bias.detach().contiguous()
        _lib.check(lib.ffno_layernorm_fwd(x2.data_ptr(), w.data_ptr(), b.data_ptr(), x2.shape[0], x2.shape[1],
                                          y.data_ptr(), _stream(x.device)), "ffno_layernorm_fwd")
    return y.reshape(x.shape)


def rel_l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Per-sample ||x_b − y_b||₂ / ||y_b||₂ (modules/loss.py:33-46) for x, y: [B, ...]."""
    require_cuda(x, "LpLoss.rel")
    require_cuda(y, "LpLoss.rel")
    lib = _lib.load()
    B = x.shape[0]
    # reshape keeps a (uniformly) strided view such as preds[..., t] without copying
    x2, y2 = x.reshape(B, -1), y.reshape(B, -1)
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.ffno_rel_l2(x2.data_ptr(), x2.stride(0), x2.stride(1), y2.data_ptr(), y2.stride(0),
                                   y2.stride(1), B, x2.shape[1], out.data_ptr(), _stream(x.device)),
                   "ffno_rel_l2")
    return out


class _RelL2(torch.autograd.Function):
    """Per-sample relative L2 with the CUDA backward (ffno_rel_l2 / ffno_rel_l2_bwd); y is a constant target."""

    @staticmethod
    def forward(ctx, x, y):
        x2, y2 = x.reshape(x.shape[0], -1).contiguous(), y.reshape(y.shape[0], -1).contiguous()
        ctx.save_for_backward(x2, y2)
        ctx.shape = x.shape
        return rel_l2(x2, y2)

    @staticmethod
    def backward(ctx, g_out):
        x2, y2 = ctx.saved_tensors
        dx = torch.empty_like(x2)
        g = g_out.contiguous().float()
        with torch.cuda.device(x2.device):
            _lib.check(_lib.load().ffno_rel_l2_bwd(x2.data_ptr(), y2.data_ptr(), g.data_ptr(), x2.shape[0], x2.shape[1],
                                                   dx.data_ptr(), _stream(x2.device)), "ffno_rel_l2_bwd")
        return dx.reshape(ctx.shape), None


def rel_l2_differentiable(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    require_cuda(x, "LpLoss.rel")
    require_cuda(y, "LpLoss.rel")
    return _RelL2.apply(x, y)


def velocity_features(w: torch.Tensor, length_x: float, length_y: float):
    """(q, v) = (psi_y, −psi_x) of the vorticity ``w[B, X, Y]`` on a periodic ``length_x × length_y`` domain — the
    ``use_velocity`` features of routines/grid_2d_markov.py:206-220 (ffno_velocity_fwd)."""
    require_cuda(w, "velocity features")
    lib = _lib.load()
    w = w.contiguous()
    B, X, Y = w.shape
    q, v = torch.empty_like(w), torch.empty_like(w)
    with torch.cuda.device(w.device):
        nbytes = lib.ffno_velocity_scratch_bytes(B, X, Y)
        scratch = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=w.device)
        _lib.check(lib.ffno_velocity_fwd(w.data_ptr(), X * Y, 1, B, X, Y, float(length_x), float(length_y), q.data_ptr(),
                                         v.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream(w.device)),
                   "ffno_velocity_fwd")
    return q, v


# ----------------------------------------------------------------------------------------------
# A plan bound to one module tree (block or single spectral layer)
# ----------------------------------------------------------------------------------------------

class LayerSpec:
    """What one spectral layer contributes: fourier weights in TENSOR-AXIS order + its FeedForwards."""

    def __init__(self, fourier_weight: Sequence[Optional[torch.Tensor]], backcast_ff, forecast_ff=None):
        self.fourier_weight = list(fourier_weight)
        self.backcast_ff = backcast_ff
        self.forecast_ff = forecast_ff


class StackPlan:
    """Owns one ``ffno_plan`` (for one device + one spatial shape) and the ctypes parameter structs.

    ``sync_params()`` re-loads (fold weight-norm, re-pack) when any parameter's storage or version
    counter changed since the last load — the CUDA counterpart of the reference recomputing
    ``g·v/‖v‖`` in a forward pre-hook on every call (modules/linear.py:49).
    """

    def __init__(self, device: torch.device, *, size: Sequence[int], pad: Sequence[int], modes: Sequence[int],
                 width: int, in_features: int, append_grid: bool, out_features: int, head_hidden: int,
                 n_layers: int, ff_factor: int, n_ff_layers: int, layer_norm: bool, use_fork: bool,
                 mode: str, path: str = "auto", transform: str = "rfft"):
        self.lib = _lib.load()
        self.device = device
        d = _lib.Desc()
        d.abi_version = _lib.ABI_VERSION
        d.ndim = len(size)
        for a in range(len(size)):
            d.size[a], d.pad[a], d.modes[a] = int(size[a]), int(pad[a]), int(modes[a])
        d.width, d.in_features, d.append_grid = width, in_features, int(append_grid)
        d.out_features, d.head_hidden, d.n_layers = out_features, head_hidden, n_layers
        d.ff_factor, d.n_ff_layers, d.layer_norm = ff_factor, n_ff_layers, int(layer_norm)
        d.use_fork, d.spectral_mode, d.path = int(use_fork), _lib.MODE[mode], _lib.PATH[path]
        d.transform = {"rfft": 0, "dct": 1, "rfft2": 2}[transform]      # ffno_transform
        self.desc = d
        self.ext = [int(size[a]) + int(pad[a]) for a in range(len(size))]
        self.size = [int(s) for s in size]
        self._plan = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.ffno_plan_create(C.byref(d), C.byref(self._plan)), "ffno_plan_create")
        self._keep: list = []
        self._version_key = None
        self._ws: Optional[torch.Tensor] = None
        self._structs = None

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None and self._plan.value:
                self.lib.ffno_plan_destroy(self._plan)
                self._plan = C.c_void_p()
        except Exception:
            pass

    @property
    def uses_umma(self) -> bool:
        return bool(self.lib.ffno_plan_uses_umma(self._plan))

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.ffno_plan_last_launch_count(self._plan))

    def pipeline_unit(self, batch: int) -> int:
        """Samples per unit of the stage-pipelined forward for this batch size; 0 = one launch per stage and layer."""
        return int(self.lib.ffno_plan_pipeline_unit(self._plan, int(batch)))

    @property
    def graph_active(self) -> bool:
        """True once the stack forward / rollout of this plan replays a captured CUDA graph."""
        return bool(self.lib.ffno_plan_graph_active(self._plan))

    # ---- parameters ---------------------------------------------------------------------------
    def sync_params(self, params: List[torch.Tensor], in_proj, out, layers: List[LayerSpec]) -> None:
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key == self._version_key:
            return
        keep: list = []
        bp = _lib.BlockParams()
        if in_proj is not None:
            bp.in_proj = linear_params(in_proj, keep)
            bp.out0 = linear_params(out[0], keep)
            bp.out1 = linear_params(out[1], keep)
        arr = (_lib.LayerParams * len(layers))()
        for l, spec in enumerate(layers):
            for a, w in enumerate(spec.fourier_weight):
                if w is not None:
                    arr[l].fourier_weight[a] = _dev_param(w, keep)
            self._fill_ff(arr[l].backcast_ff, spec.backcast_ff, keep)
            if spec.forecast_ff is not None:
                self._fill_ff(arr[l].forecast_ff, spec.forecast_ff, keep)
        bp.layers = arr
        bp.n_layers = len(layers)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ffno_plan_load_params(self._plan, C.byref(bp), _stream(self.device)),
                       "ffno_plan_load_params")
        self._keep, self._structs, self._version_key = keep, (bp, arr), key

    @staticmethod
    def _fill_ff(dst: _lib.FFParams, ff, keep: list) -> None:
        for i, layer in enumerate(ff.layers):
            dst.linear[i] = linear_params(layer[0], keep)
            if isinstance(layer[3], nn.LayerNorm):
                dst.ln_weight = _dev_param(layer[3].weight, keep)
                dst.ln_bias = _dev_param(layer[3].bias, keep)

    # ---- workspace ------------------------------------------------------------------------------
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- forward entry points ---------------------------------------------------------------------
    def block_forward(self, x: torch.Tensor, want_taps: bool = False, want_forecast_list: bool = False):
        B = x.shape[0]
        d = self.desc
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.ffno_workspace_bytes(self._plan, B))
            out = torch.empty(B, *self.size, d.out_features, device=self.device, dtype=torch.float32)
            taps_struct, taps = None, {}
            if want_taps or want_forecast_list:
                t = _lib.Taps()
                n = d.n_layers
                if want_taps:
                    full = (B, *self.ext, d.width)
                    taps["lift"] = torch.empty(full, device=self.device)
                    taps["x"] = [torch.empty(full, device=self.device) for _ in range(n)]
                    taps["b_last"] = torch.empty(full, device=self.device)
                    t.lift, t.b_last = taps["lift"].data_ptr(), taps["b_last"].data_ptr()
                    xa = (C.c_void_p * n)(*[v.data_ptr() for v in taps["x"]])
                    t.x_after = xa
                    taps["_xa"] = xa
                    if d.spectral_mode != _lib.MODE["no-fourier"]:
                        taps["s"] = [torch.empty(full, device=self.device) for _ in range(n)]
                        sa = (C.c_void_p * n)(*[v.data_ptr() for v in taps["s"]])
                        t.spectral = sa
                        taps["_sa"] = sa
                if want_forecast_list:
                    taps["forecast_list"] = [torch.empty_like(out) for _ in range(n)]
                    fa = (C.c_void_p * n)(*[v.data_ptr() for v in taps["forecast_list"]])
                    t.forecast_list = fa
                    taps["_fa"] = fa
                taps_struct = C.byref(t)
            _lib.check(self.lib.ffno_block_fwd(self._plan, x.data_ptr(), B, out.data_ptr(), taps_struct,
                                               ws.data_ptr(), ws.numel(), _stream(self.device)), "ffno_block_fwd")
        return out, taps

    def set_backward_mode(self, mode: str) -> None:
        """'default' (FP32 forward recompute + tcgen05 spectral adjoint), 'fast' (tcgen05 recompute too) or 'fp32'."""
        _lib.check(self.lib.ffno_plan_set_backward_mode(self._plan, {"default": 0, "fast": 1, "fp32": 2}[mode]),
                   "ffno_plan_set_backward_mode")

    def block_backward(self, x: torch.Tensor, d_forecast: torch.Tensor, in_proj, out, layers: List[LayerSpec],
                       want_dx: bool, layer_bias: Optional[torch.Tensor] = None, want_dbias: bool = False):
        """Gradients of the stack (ffno_block_bwd) -> (dx or None, {id(parameter): gradient tensor}).  Call right after
        ``sync_params`` with the same modules: the raw parameter pointers of that load are what the weight-norm backward
        reads; a parameter shared by several layers gets one buffer that receives the sum.
        ``in_proj is None``: the layer loop alone (ffno_layers_bwd, plans without lift / head): ``x`` is x_0,
        ``d_forecast`` is dL/dx_L, ``layer_bias`` [points, C] is added after every layer -> (dx, gmap, d_bias or None)."""
        B = x.shape[0]
        gmap: dict = {}

        def gbuf(prm: Optional[torch.Tensor]):
            if prm is None or not prm.requires_grad:
                return None
            g = gmap.get(id(prm))
            if g is None:
                g = torch.zeros(prm.shape, device=self.device, dtype=torch.float32)
                gmap[id(prm)] = g
            return g.data_ptr()

        def lin_grads(dst: _lib.LinearGrads, lin: nn.Linear) -> None:
            if "weight" in lin._parameters and lin._parameters["weight"] is not None:
                dst.weight = gbuf(lin._parameters["weight"])
            else:
                dst.weight_g, dst.weight_v = gbuf(lin.weight_g), gbuf(lin.weight_v)
                if (dst.weight_g is None) != (dst.weight_v is None):
                    raise RuntimeError("weight-normed linear: weight_g and weight_v must both (not) require grad")
            dst.bias = gbuf(lin.bias)

        bg = _lib.BlockGrads()
        if in_proj is not None:
            lin_grads(bg.in_proj, in_proj)
            lin_grads(bg.out0, out[0])
            lin_grads(bg.out1, out[1])
        arr = (_lib.LayerGrads * len(layers))()
        for l, spec in enumerate(layers):
            for a, w in enumerate(spec.fourier_weight):
                if w is not None:
                    arr[l].fourier_weight[a] = gbuf(w)
            for i, layer in enumerate(spec.backcast_ff.layers):
                lin_grads(arr[l].backcast_ff[i], layer[0])
        bg.layers = arr
        bg.n_layers = len(layers)
        with torch.cuda.device(self.device):
            need = self.lib.ffno_block_bwd_workspace_bytes(self._plan, B)
            ws = getattr(self, "_ws_bwd", None)
            if ws is None or ws.numel() < need:
                ws = self._ws_bwd = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            dx = torch.empty_like(x) if want_dx else None
            if in_proj is None:
                d_bias = torch.zeros_like(layer_bias) if (want_dbias and layer_bias is not None) else None
                _lib.check(self.lib.ffno_layers_bwd(self._plan, C.byref(self._structs[0]), x.data_ptr(), d_forecast.data_ptr(),
                                                    _ptr(layer_bias), B, C.byref(bg), _ptr(dx), _ptr(d_bias), ws.data_ptr(),
                                                    ws.numel(), _stream(self.device)), "ffno_layers_bwd")
                return dx, gmap, d_bias
            _lib.check(self.lib.ffno_block_bwd(self._plan, C.byref(self._structs[0]), x.data_ptr(), d_forecast.data_ptr(),
                                               B, C.byref(bg), _ptr(dx), ws.data_ptr(), ws.numel(), _stream(self.device)),
                       "ffno_block_bwd")
        return dx, gmap

    def block_forward_host(self, x_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Host-buffer entry point (ffno_block_fwd_host): H2D + forward + D2H + stream sync inside."""
        B = x_host.shape[0]
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.ffno_workspace_bytes_host(self._plan, B))
            _lib.check(self.lib.ffno_block_fwd_host(self._plan, x_host.data_ptr(), B, out_host.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), _stream(self.device)),
                       "ffno_block_fwd_host")

    def spectral_forward(self, layer: int, x: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.ffno_workspace_bytes(self._plan, B))
            s = torch.empty_like(x)
            _lib.check(self.lib.ffno_spectral_fwd(self._plan, layer, x.data_ptr(), B, s.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), _stream(self.device)), "ffno_spectral_fwd")
        return s

    def spectral_split_forward(self, layer: int, x: torch.Tensor, out: Optional[List[torch.Tensor]] = None):
        """The spectral operator as the layer loop runs it: one tensor per spatial axis, whose sum is forward_fourier(x)."""
        B = x.shape[0]
        nd = len(self.size)
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.ffno_workspace_bytes(self._plan, B))
            if out is None:
                out = [torch.empty_like(x) for _ in range(nd)]
            ptrs = (C.c_void_p * 3)(*[t.data_ptr() for t in out], *([None] * (3 - nd)))
            _lib.check(self.lib.ffno_spectral_split_fwd(self._plan, layer, x.data_ptr(), B, ptrs, ws.data_ptr(), ws.numel(),
                                                        _stream(self.device)), "ffno_spectral_split_fwd")
        return out

    def ff_forward(self, layer: int, which: int, s: torch.Tensor, residual: Optional[torch.Tensor]) -> torch.Tensor:
        B = s.shape[0]
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.ffno_workspace_bytes(self._plan, B))
            y = torch.empty_like(s)
            _lib.check(self.lib.ffno_ff_fwd(self._plan, layer, which, s.data_ptr(), _ptr(residual), B, y.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _stream(self.device)), "ffno_ff_fwd")
        return y

    def rollout_forward(self, frame0: torch.Tensor, n_steps: int, mean: Sequence[float], std: Sequence[float],
                        low: float, high: float, domain: Optional[Sequence[float]] = None,
                        force: Optional[torch.Tensor] = None, mu: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``domain`` = (length_x, length_y) of the periodic box: set for velocity-feature rollouts only.
        ``force``: contiguous [B, X, Y, 1 | n_steps] forcing channel, ``mu``: [B] viscosity channel (ffno_rollout_fwd_ex)."""
        B = frame0.shape[0]
        X, Y = self.size
        with torch.cuda.device(self.device):
            if domain is not None:
                _lib.check(self.lib.ffno_plan_set_domain(self._plan, float(domain[0]), float(domain[1])),
                           "ffno_plan_set_domain")
            preds = torch.empty(B, X, Y, n_steps, device=self.device, dtype=torch.float32)
            n_feat = len(mean)
            m = (C.c_float * n_feat)(*[float(v) for v in mean])
            s = (C.c_float * n_feat)(*[float(v) for v in std])
            if force is None and mu is None:
                ws = self._workspace(self.lib.ffno_rollout_workspace_bytes(self._plan, B, n_steps))
                _lib.check(self.lib.ffno_rollout_fwd(self._plan, frame0.data_ptr(), B, n_steps, m, s, float(low),
                                                     float(high), preds.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     _stream(self.device)), "ffno_rollout_fwd")
            else:
                f_steps = 0 if force is None else int(force.shape[-1])
                ex = _lib.RolloutExtras(1 if domain is not None else 0, f_steps, _ptr(force), _ptr(mu))
                ws = self._workspace(self.lib.ffno_rollout_workspace_bytes_ex(
                    self._plan, B, n_steps, ex.use_velocity, f_steps, 0 if mu is None else 1))
                _lib.check(self.lib.ffno_rollout_fwd_ex(self._plan, frame0.data_ptr(), B, n_steps, m, s, float(low),
                                                        float(high), C.byref(ex), preds.data_ptr(), ws.data_ptr(),
                                                        ws.numel(), _stream(self.device)), "ffno_rollout_fwd_ex")
        return preds
