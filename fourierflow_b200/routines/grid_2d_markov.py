"""``Grid2DMarkovExperiment`` — host-side mirror of fourierflow/routines/grid_2d_markov.py (inference path).

The reference class is a pytorch_lightning ``Routine``; this mirror is a plain ``nn.Module`` with the same
constructor keywords and the same ``forward(batch) -> (loss, step_losses, preds, pred_layer_list)``
contract for the configurations the BASELINE names (torus_li/markov: position features + normaliser;
torus_kochkov: additionally ``use_velocity`` — stream-function velocities recomputed from the vorticity at
every step; no force / mu / grid shuffling / difference learning).  The step loop
(``_valid_step``, grid_2d_markov.py:195-326) runs entirely in libffno_b200 (ffno_rollout_fwd): feature
build → normalise → layer stack → de-normalise, feeding each forecast back, with no host round trip
between steps.  The relative-L2 reduction (modules/loss.py:33-46) is ffno_rel_l2 per sample; the mean over
the batch is the only value a sharded run communicates (see fourierflow_b200/distributed.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import _ops
from ..modules.loss import LpLoss
from .base import RoutineMixin
from ..modules.normalizer import Normalizer


class Grid2DMarkovExperiment(RoutineMixin, nn.Module):
    def __init__(self, conv: nn.Module, n_steps: Optional[int] = None, num_freq_bands: int = 8, freq_base: int = 2,
                 low: float = 0, high: float = 1, use_position: bool = True, append_force: bool = False,
                 append_mu: bool = False, max_accumulations: float = 1e6, should_normalize: bool = True,
                 use_fourier_position: bool = False, noise_std: float = 0.0, shuffle_grid: bool = False,
                 use_velocity: bool = False, learn_difference: bool = False, step_size: float = 1.0,
                 n_test_steps_logged: Optional[int] = None,
                 domain=((0.0, 2 * math.pi), (0.0, 2 * math.pi)), **kwargs):
        super().__init__()
        unsupported = dict(append_force=append_force, append_mu=append_mu, use_fourier_position=use_fourier_position,
                           shuffle_grid=shuffle_grid, learn_difference=learn_difference)
        bad = [k for k, v in unsupported.items() if v]
        if bad or not use_position or not should_normalize:
            raise RuntimeError("Grid2DMarkovExperiment (B200 backend): only the torus_li/markov and torus_kochkov "
                               "feature sets are implemented (use_position + should_normalize [+ use_velocity]); "
                               f"unsupported: {bad}")
        want = 5 if use_velocity else 3
        if conv.input_dim != want:
            raise RuntimeError(f"Grid2DMarkovExperiment: use_velocity={use_velocity} builds {want} input features, "
                               f"conv.input_dim is {conv.input_dim}")
        self.use_velocity = use_velocity
        (x0, x1), (y0, y1) = domain                   # periodic box of the velocity features (grid_2d_markov.py:43,85)
        self.domain_lengths = (float(x1) - float(x0), float(y1) - float(y0))
        self.conv = conv
        self.n_steps = n_steps
        self.l2_loss = LpLoss(size_average=True)
        self.low, self.high = low, high
        self.step_size = step_size
        self.noise_std = noise_std          # training-only in the reference (:150-151); unused at inference
        self.normalizer = Normalizer([conv.input_dim], max_accumulations)
        self.register_buffer('_float', torch.FloatTensor([0.1]))
        self._ms_cache = None

    # -- statistics (reference: epoch 0 of training only accumulates, grid_2d_markov.py:376-378) -----------
    @torch.no_grad()
    def accumulate_statistics(self, data: torch.Tensor) -> None:
        """Fold every one-step input of ``data[B,X,Y,T]`` (frames 0..T-2 + position grid) into the normaliser."""
        B, X, Y, T = data.shape
        pos = self._positions(X, Y, data.device, data.dtype)
        parts = [data[..., :-1].unsqueeze(-1)]
        if self.use_velocity:                         # q, v of every input frame (:130-144), computed by the CUDA library
            frames = data[..., :-1].permute(0, 3, 1, 2).reshape(B * (T - 1), X, Y)
            q, v = _ops.velocity_features(frames, *self.domain_lengths)
            parts += [t.reshape(B, T - 1, X, Y).permute(0, 2, 3, 1).unsqueeze(-1) for t in (q, v)]
        parts.append(pos[None, :, :, None, :].expand(B, X, Y, T - 1, 2))
        feats = torch.cat(parts, dim=-1)
        self.normalizer.accumulate(feats)
        self._ms_cache = None

    def _positions(self, X, Y, device, dtype):
        gx = torch.linspace(self.low, self.high, steps=X, device=device, dtype=dtype)
        gy = torch.linspace(self.low, self.high, steps=Y, device=device, dtype=dtype)
        return torch.stack(torch.meshgrid(gx, gy, indexing='ij'), dim=-1)

    def _mean_std(self):
        key = (self.normalizer.sum._version, self.normalizer.count._version, self.normalizer.sum.data_ptr())
        if self._ms_cache is None or self._ms_cache[0] != key:
            self._ms_cache = (key, self.normalizer.mean.tolist(), self.normalizer.std.tolist())
        return self._ms_cache[1], self._ms_cache[2]

    # -- inference -------------------------------------------------------------------------------------
    def forward(self, data):
        return self._valid_step(data)

    def predict(self, data: torch.Tensor, n_steps: Optional[int] = None) -> torch.Tensor:
        """preds[B,X,Y,n_steps] of the Markov rollout started from frame T−n_steps−1 of ``data``."""
        _ops.require_cuda(data, "Grid2DMarkovExperiment")
        B, X, Y, T = data.shape
        n_steps = n_steps or self.n_steps or T - 1
        if not 1 <= n_steps <= T - 1:      # a negative frame index below would silently wrap to the end of the series
            raise RuntimeError(f"Grid2DMarkovExperiment: a {n_steps}-step rollout needs {n_steps + 1} frames, "
                               f"data has T={T}")
        plan = self.conv.plan_for(data.device, (X, Y))
        mean, std = self._mean_std()
        frame0 = data[..., T - n_steps - 1].contiguous()
        return plan.rollout_forward(frame0, n_steps, mean, std, self.low, self.high,
                                    self.domain_lengths if self.use_velocity else None)

    def per_sample_losses(self, preds: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
        """[n_steps, B] relative L2 of every rollout step against the last n_steps frames of ``data``."""
        n_steps = preds.shape[-1]
        yy = data[..., -n_steps:]
        return torch.stack([self.l2_loss.rel_per_sample(preds[..., t], yy[..., t]) for t in range(n_steps)])

    def _valid_step(self, batch):
        data = batch['data'] if isinstance(batch, dict) else batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.conv.parameters()):
            _ops.require_inference(self.conv, data)
        preds = self.predict(data)
        losses = self.per_sample_losses(preds, data)            # [n_steps, B]
        step_losses = list(losses.mean(dim=1))                  # LpLoss(size_average) per step (:313)
        loss = losses.mean(dim=1).sum()                         # loss += l (:315)
        # one (empty) per-layer forecast list per step, as `pred_layer_list.append(out['forecast_list'])` builds it for
        # a conv without use_fork (grid_2d_markov.py:320-321)
        return loss, step_losses, preds, [[] for _ in range(preds.shape[-1])]
