"""Geo-F-FNO on point clouds — host-side mirror of fourierflow/modules/factorized_fno/point_cloud_2d.py.

Only the INTERIOR of this operator is on the hot path: layers 1 .. n_layers-1 are the periodic-grid
factorized spectral layer (grid_2d.py ``SpectralConv2d``) on the s1 x s2 latent grid with a grid-bias term
(point_cloud_2d.py:198-210).  That loop runs device-resident on libffno_b200: one plan for all interior layers,
``ffno_spectral_fwd`` + ``ffno_ff_fwd`` per layer, the residual and the bias folded into the FF kernel's residual input.

The two end layers (:194-196, :212-216) are non-uniform discrete Fourier sums over the mesh points, deformed by the
user's ``iphi`` network; they are torch glue outside the hot path, written here as dense complex matrix products
against an explicitly built basis.  Constructor arguments, parameter names (``fc0 convs ws bs fc1 fc2``) and
initialisation order follow the reference so that its checkpoints load with ``strict=True``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _ops
from ._base import PlanCacheMixin, default_path
from .grid_2d import SpectralConv2d as FactorizedSpectralConv2d


def _wavenumbers(modes1: int, modes2: int, device) -> tuple:
    """Signed wavenumbers of the retained block: 2*modes1 along axis 0 (0..m1-1, -m1..-1), 2*modes2-1 along axis 1
    (0..m2-1, -(m2-1)..-1) (point_cloud_2d.py:92-96)."""
    k1 = torch.cat((torch.arange(0, modes1), torch.arange(-modes1, 0))).to(device=device, dtype=torch.float32)
    k2 = torch.cat((torch.arange(0, modes2), torch.arange(-(modes2 - 1), 0))).to(device=device, dtype=torch.float32)
    return k1, k2


def _basis(x: torch.Tensor, modes1: int, modes2: int, sign: float) -> torch.Tensor:
    """exp(sign * 2 pi i <x_n, k>) for every point n and retained wavenumber k: [B, N, 2*modes1, 2*modes2-1]."""
    k1, k2 = _wavenumbers(modes1, modes2, x.device)
    phase = x[..., 0, None, None] * k1[:, None] + x[..., 1, None, None] * k2[None, :]
    return torch.exp(sign * 2j * np.pi * phase)


class SpectralConv2d(nn.Module):
    """End layers of the geo operator (point_cloud_2d.py:16-153): Fourier sums between the point cloud and the
    latent grid.  ``transform=False`` (first layer) has no weights."""

    def __init__(self, in_channels, out_channels, modes1, modes2, s1=32, s2=32, transform=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2, self.s1, self.s2 = modes1, modes2, s1, s2
        if transform:
            self.scale = (1 / (in_channels * out_channels))
            self.weights1 = nn.Parameter(
                self.scale * torch.rand(in_channels, out_channels, self.modes1, self.modes2, dtype=torch.cfloat))
            self.weights2 = nn.Parameter(
                self.scale * torch.rand(in_channels, out_channels, self.modes1, self.modes2, dtype=torch.cfloat))

    def fft2d(self, u, x_in, iphi=None, code=None):
        """u:[B, C, N] at points x_in:[B, N, 2] -> coefficients [B, C, 2*modes1, 2*modes2-1] (point_cloud_2d.py:82-116)."""
        x = x_in if iphi is None else iphi(x_in, code)
        return torch.einsum("bcn,bnxy->bcxy", u + 0j, _basis(x, self.modes1, self.modes2, -1.0))

    def ifft2d(self, u_ft, x_out, iphi=None, code=None):
        """Half-spectrum u_ft:[B, C, 2*modes1, modes2] -> real values at x_out (point_cloud_2d.py:118-153): the missing
        half is the conjugate of the point-reflected block."""
        x = x_out if iphi is None else iphi(x_out, code)
        full = torch.cat([u_ft, u_ft[..., 1:].flip(-1, -2).conj()], dim=-1)
        return torch.einsum("bcxy,bnxy->bcn", full, _basis(x, self.modes1, self.modes2, 1.0)).real

    def forward(self, u, x_in=None, x_out=None, iphi=None, code=None, transform=True):
        m1, m2 = self.modes1, self.modes2
        if x_in is None:
            u_ft, s1, s2 = torch.fft.rfft2(u), u.size(-2), u.size(-1)
        else:
            u_ft, s1, s2 = self.fft2d(u, x_in, iphi, code), self.s1, self.s2
        lo, hi = u_ft[:, :, :m1, :m2], u_ft[:, :, -m1:, :m2]
        if transform:
            lo = torch.einsum("bixy,ioxy->boxy", lo, self.weights1)
            hi = torch.einsum("bixy,ioxy->boxy", hi, self.weights2)
        if x_out is None:
            out_ft = torch.zeros(u.shape[0], self.out_channels, s1, s2 // 2 + 1, dtype=torch.cfloat, device=u.device)
            out_ft[:, :, :m1, :m2] = lo
            out_ft[:, :, -m1:, :m2] = hi
            return torch.fft.irfft2(out_ft, s=(s1, s2))
        return self.ifft2d(torch.cat([lo, hi], dim=-2), x_out, iphi, code)


class _InteriorFunction(torch.autograd.Function):
    """The interior layer loop as one autograd node: forward = the device-resident loop, backward = ffno_layers_bwd
    (explicit adjoints on the FP32 kernels; the loop is recomputed there, only its input is saved).  The parameters are
    inputs only so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, module, uc, grid_bias, *params):
        ctx.module = module
        ctx.save_for_backward(uc, grid_bias)
        return module._interior_run(uc, grid_bias)

    @staticmethod
    def backward(ctx, d_out):
        uc, grid_bias = ctx.saved_tensors
        module = ctx.module
        plan = module._interior_plan(uc.device)
        d_uc, gmap, d_bias = plan.block_backward(
            uc, d_out.contiguous().float(), None, None, module.__dict__["_spec_cache"], ctx.needs_input_grad[1],
            layer_bias=grid_bias, want_dbias=ctx.needs_input_grad[2])
        return (None, d_uc, d_bias) + tuple(gmap.get(id(p)) for p in module._interior_params())


class FNOFactorizedPointCloud2D(PlanCacheMixin, nn.Module):
    def __init__(self, modes1, modes2, width, in_channels, out_channels, n_layers=4, is_mesh=True, s1=40, s2=40,
                 share_weight=False):
        super().__init__()
        self.modes1, self.modes2, self.width = modes1, modes2, width
        self.is_mesh, self.s1, self.s2, self.n_layers = is_mesh, s1, s2, n_layers

        self.fc0 = nn.Linear(in_channels, self.width)
        self.convs = nn.ModuleList([])
        self.ws = nn.ModuleList([])
        self.bs = nn.ModuleList([])

        self.fourier_weight = None
        if share_weight:
            self.fourier_weight = nn.ParameterList([])
            for _ in range(2):
                param = nn.Parameter(torch.empty(width, width, modes1, 2))
                nn.init.xavier_normal_(param)
                self.fourier_weight.append(param)

        for i in range(self.n_layers + 1):
            if i == 0:
                conv = SpectralConv2d(self.width, self.width, self.modes1, self.modes2, s1, s2, transform=False)
            elif i == self.n_layers:
                conv = SpectralConv2d(self.width, self.width, self.modes1, self.modes2, s1, s2)
            else:
                conv = FactorizedSpectralConv2d(
                    in_dim=width, out_dim=width, n_modes=modes1, forecast_ff=None, backcast_ff=None,
                    fourier_weight=self.fourier_weight, factor=2, ff_weight_norm=True, n_ff_layers=2,
                    layer_norm=False, use_fork=False, dropout=0.0, mode='full')
            self.convs.append(conv)

        self.bs.append(nn.Conv2d(2, self.width, 1))
        self.bs.append(nn.Conv1d(2, self.width, 1))
        for _ in range(self.n_layers - 1):
            self.ws.append(nn.Conv2d(self.width, self.width, 1))      # registered, unused by forward (reference :204)

        self.fc1 = nn.Linear(self.width, 128)
        self.fc2 = nn.Linear(128, out_channels)

    # -- the hot part ---------------------------------------------------------------------------------------------
    def _interior_plan(self, device) -> _ops.StackPlan:
        plan = self._get_plan(
            device, (self.s1, self.s2), pad=(0, 0), modes=(self.modes1, self.modes1), width=self.width, in_features=1,
            append_grid=False, out_features=1, head_hidden=1, n_layers=self.n_layers - 1, ff_factor=2, n_ff_layers=2,
            layer_norm=False, use_fork=False, mode='full', path=default_path())
        params = self._flat_params()
        specs = self.__dict__.get("_spec_cache")
        if specs is None:
            specs = [self.convs[i].layer_spec() for i in range(1, self.n_layers)]
            self.__dict__["_spec_cache"] = specs
        plan.sync_params(params, None, None, specs)
        return plan

    def interior_forward(self, uc: torch.Tensor, grid_bias: torch.Tensor) -> torch.Tensor:
        """uc:[B, s1, s2, width] channels-last, grid_bias:[s1, s2, width] (``bs[0](grid)``, the same for every sample and
        layer) -> the latent field after layers 1 .. n_layers-1:  uc <- uc + backcast_ff(forward_fourier(uc)) + grid_bias
        (point_cloud_2d.py:198-210), device-resident."""
        _ops.require_cuda(uc, "FNOFactorizedPointCloud2D.interior_forward")
        if self.n_layers < 2:
            return uc
        uc, grid_bias = uc.contiguous(), grid_bias.contiguous()
        if torch.is_grad_enabled() and (uc.requires_grad or grid_bias.requires_grad or
                                        any(p.requires_grad for p in self._interior_params())):
            return _InteriorFunction.apply(self, uc, grid_bias, *self._interior_params())      # training: ffno_layers_bwd
        return self._interior_run(uc, grid_bias)

    def _interior_params(self):
        """Distinct parameters of the interior layers, in a fixed order (the autograd node's parameter inputs)."""
        seen, out = set(), []
        for i in range(1, self.n_layers):
            for p in self.convs[i].parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def _interior_run(self, uc: torch.Tensor, grid_bias: torch.Tensor) -> torch.Tensor:
        plan = self._interior_plan(uc.device)
        for l in range(self.n_layers - 1):
            s = plan.spectral_forward(l, uc)
            uc = plan.ff_forward(l, 0, s, uc + grid_bias)      # FF(s) + residual, residual = uc + grid bias
        return uc

    # -- reference API -------------------------------------------------------------------------------------------------
    def forward(self, u, code=None, x_in=None, x_out=None, iphi=None):
        """u:[B, N, 2] mesh coordinates (and features) -> [B, N, out_channels] (point_cloud_2d.py:173-227)."""
        _ops.require_cuda(u, "FNOFactorizedPointCloud2D.forward")
        if self.is_mesh and x_in is None:
            x_in = u
        if self.is_mesh and x_out is None:
            x_out = u
        grid = self.get_grid([1, self.s1, self.s2], u.device)[0]                     # [s1, s2, 2]
        w0 = self.bs[0].weight.reshape(self.width, 2)
        grid_bias = F.linear(grid, w0, self.bs[0].bias)                              # 1x1 conv of the grid: [s1, s2, width]

        v = self.fc0(u).permute(0, 2, 1)                                             # [B, width, N]
        uc = self.convs[0](v, x_in=x_in, iphi=iphi, code=code, transform=False)      # [B, width, s1, s2]
        uc = uc.permute(0, 2, 3, 1) + grid_bias                                      # channels-last latent grid
        uc = self.interior_forward(uc.float(), grid_bias)
        out = self.convs[self.n_layers](uc.permute(0, 3, 1, 2), x_out=x_out, iphi=iphi, code=code)
        out = out + self.bs[-1](x_out.permute(0, 2, 1))
        out = out.permute(0, 2, 1)
        return self.fc2(F.gelu(self.fc1(out)))

    def get_grid(self, shape, device):
        """[B, s1, s2, 2] inclusive linspace(0, 1) coordinates of the latent grid (point_cloud_2d.py:229-238)."""
        b, sx, sy = shape[0], shape[1], shape[2]
        gx = torch.tensor(np.linspace(0, 1, sx), dtype=torch.float).reshape(1, sx, 1, 1).expand(b, sx, sy, 1)
        gy = torch.tensor(np.linspace(0, 1, sy), dtype=torch.float).reshape(1, 1, sy, 1).expand(b, sx, sy, 1)
        return torch.cat((gx, gy), dim=-1).to(device)
