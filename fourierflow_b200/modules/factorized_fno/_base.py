"""Shared host logic of the three F-FNO block mirrors: plan cache, parameter hand-off, forward dispatch."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ... import _ops


# Bumped whenever ANY nn.Module in the process registers a sub-module or a parameter (torch's global registration
# hooks): the cheap way to know that a cached parameter list may no longer describe the module tree.
_STRUCTURE_EPOCH = [0]


def _bump_structure_epoch(*_args):
    _STRUCTURE_EPOCH[0] += 1
    return None


torch.nn.modules.module.register_module_module_registration_hook(_bump_structure_epoch)
torch.nn.modules.module.register_module_parameter_registration_hook(_bump_structure_epoch)


def default_path() -> str:
    """`FFNO_B200_PATH=generic|umma|auto` overrides the kernel family (testing / profiling)."""
    return os.environ.get("FFNO_B200_PATH", "auto")


class PlanCacheMixin:
    """Lazily builds one ``StackPlan`` per (device, spatial size) and keeps parameters in sync."""

    def _plans(self) -> Dict[Tuple, _ops.StackPlan]:
        d = self.__dict__.get("_plan_cache")
        if d is None:
            d = {}
            self.__dict__["_plan_cache"] = d       # not a module attribute: skipped by state_dict/deepcopy hooks
        return d

    def _flat_params(self):
        """Distinct parameters of the module tree, cached — walking a 24-layer tree with ``parameters()`` on every
        forward costs more than launching the CUDA graph.  The cache is re-validated on every call: (i) each cached
        Parameter must still be the object registered under its owner's name (``load_state_dict(assign=True)``,
        re-registration, ``_apply`` replacements), and (ii) no module or parameter may have been registered anywhere in
        the process since the cache was built (swapped sub-modules, added layers) — torch's global registration hooks
        bump ``_STRUCTURE_EPOCH``.  The reference re-folds weight-norm from the live parameters on every forward
        (linear.py:49); this is the same guarantee at ~15 us per call.  In-place edits through ``p.data`` bump no
        version counter and cannot be seen from here: call ``invalidate_plans()`` after such an edit."""
        cached = self.__dict__.get("_param_cache")
        if cached is not None:
            epoch, owners, params = cached
            if epoch == _STRUCTURE_EPOCH[0]:
                for mod, name, p in owners:
                    if mod._parameters.get(name) is not p:
                        break
                else:
                    return params
        owners, params, seen = [], [], set()
        for mod in self.modules():
            for name, p in mod._parameters.items():
                if p is None:
                    continue
                owners.append((mod, name, p))
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.__dict__["_param_cache"] = (_STRUCTURE_EPOCH[0], owners, params)
        self.__dict__.pop("_spec_cache", None)          # layer specs hold Parameter references too
        return params

    def invalidate_plans(self) -> None:
        """Force the next forward to re-read, re-fold and re-pack every parameter (needed only after in-place edits
        through ``p.data``, which no version counter records)."""
        self.__dict__.pop("_param_cache", None)
        self.__dict__.pop("_spec_cache", None)
        for plan in self._plans().values():
            plan._version_key = None

    def __deepcopy__(self, memo):
        # plans hold raw device pointers of THIS module's parameters: never share them with a copy
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_plan_cache", "_param_cache", "_spec_cache"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_plan_cache", "_param_cache", "_spec_cache"):
            st.pop(k, None)
        return st

    def _get_plan(self, device: torch.device, size: Sequence[int], **desc_kwargs) -> _ops.StackPlan:
        key = (device.index if device.index is not None else torch.cuda.current_device(), tuple(size),
               desc_kwargs.get("path", "auto"))
        plans = self._plans()
        plan = plans.get(key)
        if plan is None:
            plan = _ops.StackPlan(device, size=size, **desc_kwargs)
            plans[key] = plan
        return plan


class StackFunction(torch.autograd.Function):
    """The forward of a block mirror (FNOFactorized2DBlock / FNOFactorizedMesh2D / FNOFactorizedMesh3D) as one autograd
    node: forward = ffno_block_fwd (the same kernels as inference), backward = ffno_block_bwd (explicit adjoints; the
    forward is recomputed there, so nothing but the input is saved).  The parameters are passed as inputs only so that
    autograd routes their gradients."""

    @staticmethod
    def forward(ctx, module, x, *params):
        plan = module.plan_for(x.device, x.shape[1:-1])
        forecast, _ = plan.block_forward(x)
        ctx.module = module
        ctx.save_for_backward(x)
        return forecast

    @staticmethod
    def backward(ctx, d_forecast):
        (x,) = ctx.saved_tensors
        module = ctx.module
        plan = module.plan_for(x.device, x.shape[1:-1])        # re-syncs if a parameter changed since the forward
        dx, gmap = plan.block_backward(x, d_forecast.contiguous().float(), module.in_proj, module.out,
                                       module.__dict__["_spec_cache"], ctx.needs_input_grad[1])
        return (None, dx) + tuple(gmap.get(id(p)) for p in module._flat_params())


def check_trainable(module) -> None:
    """Options the CUDA backward covers (every shipped F-FNO config): n_ff_layers = 2, no LayerNorm, no fork, mode 'full'."""
    if getattr(module, "use_fork", False) or module.layer_norm or module.n_ff_layers != 2 or \
            getattr(module, "mode", "full") != "full":
        raise RuntimeError(f"{type(module).__name__}: the CUDA backward covers n_ff_layers=2, no LayerNorm, no fork, "
                           "mode='full' (every shipped F-FNO config); use torch.no_grad() for the others")


def check_input(x: torch.Tensor, ndim: int, features: int, what: str) -> torch.Tensor:
    _ops.require_cuda(x, what)
    if x.dim() != ndim + 2 or x.shape[-1] != features:
        raise RuntimeError(f"{what}: expected [B, {'S, ' * ndim}{features}], got {tuple(x.shape)}")
    return x.contiguous()
