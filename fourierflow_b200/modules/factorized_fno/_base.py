"""Shared host logic of the three F-FNO block mirrors: plan cache, parameter hand-off, forward dispatch."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ... import _ops


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
        """Parameter list in registration order, cached: walking a 24-layer module tree on every forward costs more
        than the launch of the CUDA graph.  The Parameter objects survive load_state_dict / .cuda() / in-place
        updates (their storage pointer and version counter are what StackPlan.sync_params watches)."""
        cached = self.__dict__.get("_param_cache")
        n = sum(1 for _ in self.parameters()) if cached is None else None
        if cached is None or len(cached) != (n if n is not None else len(cached)):
            cached = list(self.parameters())
            self.__dict__["_param_cache"] = cached
        return cached

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


def check_input(x: torch.Tensor, ndim: int, features: int, what: str) -> torch.Tensor:
    _ops.require_cuda(x, what)
    if x.dim() != ndim + 2 or x.shape[-1] != features:
        raise RuntimeError(f"{what}: expected [B, {'S, ' * ndim}{features}], got {tuple(x.shape)}")
    return x.contiguous()
