"""``Grid2DMarkovExperiment`` — host-side mirror of fourierflow/routines/grid_2d_markov.py (inference path).

The reference class is a pytorch_lightning ``Routine``; this mirror is a plain ``nn.Module`` with the same
constructor keywords and the same ``forward(batch) -> (loss, step_losses, preds, pred_layer_list)``
contract for the configurations the BASELINE names (torus_li/markov: position features + normaliser;
torus_kochkov: additionally ``use_velocity`` — stream-function velocities recomputed from the vorticity at
every step) and for the torus_vis / torus_vis_force feature sets (``append_force``: the batch's ``'f'`` forcing,
static [B,X,Y] or time-varying [B,X,Y,T]; ``append_mu``: the batch's ``'mu'`` viscosity per sample).  The ablation
switches (grid shuffling, difference learning, Fourier position features) raise.  The step loop
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
        unsupported = dict(use_fourier_position=use_fourier_position, shuffle_grid=shuffle_grid,
                           learn_difference=learn_difference)
        bad = [k for k, v in unsupported.items() if v]
        if bad or not use_position or not should_normalize:
            raise RuntimeError("Grid2DMarkovExperiment (B200 backend): implemented feature sets are use_position + "
                               "should_normalize [+ use_velocity] [+ append_force] [+ append_mu]; "
                               f"unsupported: {bad}")
        want = 3 + (2 if use_velocity else 0) + (1 if append_force else 0) + (1 if append_mu else 0)
        if conv.input_dim != want:
            raise RuntimeError(f"Grid2DMarkovExperiment: use_velocity={use_velocity}, append_force={append_force}, "
                               f"append_mu={append_mu} build {want} input features, conv.input_dim is {conv.input_dim}")
        self.use_velocity, self.append_force, self.append_mu = use_velocity, append_force, append_mu
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
    def accumulate_statistics(self, data: torch.Tensor, force: Optional[torch.Tensor] = None,
                              mu: Optional[torch.Tensor] = None) -> None:
        """Fold every one-step input of ``data[B,X,Y,T]`` (frames 0..T-2 + position grid [+ force + mu]) into the
        normaliser."""
        B, X, Y, T = data.shape
        self._check_extras(force, mu, B, X, Y)
        pos = self._positions(X, Y, data.device, data.dtype)
        parts = [data[..., :-1].unsqueeze(-1)]
        if self.use_velocity:                         # q, v of every input frame (:130-144), computed by the CUDA library
            frames = data[..., :-1].permute(0, 3, 1, 2).reshape(B * (T - 1), X, Y)
            q, v = _ops.velocity_features(frames, *self.domain_lengths)
            parts += [t.reshape(B, T - 1, X, Y).permute(0, 2, 3, 1).unsqueeze(-1) for t in (q, v)]
        parts.append(pos[None, :, :, None, :].expand(B, X, Y, T - 1, 2))
        if self.append_force:      # frame t of a time-varying forcing accompanies input frame t
            f = force.unsqueeze(-1).expand(B, X, Y, T - 1) if force.dim() == 3 else force[..., :T - 1]
            parts.append(f.unsqueeze(-1))
        if self.append_mu:
            parts.append(mu.reshape(B, 1, 1, 1, 1).expand(B, X, Y, T - 1, 1))
        feats = torch.cat(parts, dim=-1)
        self.normalizer.accumulate(feats)
        self._ms_cache = None

    def _check_extras(self, force, mu, B, X, Y):
        if self.append_force != (force is not None) or self.append_mu != (mu is not None):
            raise RuntimeError(f"Grid2DMarkovExperiment: append_force={self.append_force} / append_mu={self.append_mu} "
                               f"need batch['f'] / batch['mu'] (and only then): got f={'yes' if force is not None else 'no'}, "
                               f"mu={'yes' if mu is not None else 'no'}")
        if force is not None and (force.dim() not in (3, 4) or tuple(force.shape[:3]) != (B, X, Y)):
            raise RuntimeError(f"Grid2DMarkovExperiment: batch['f'] must be [B,X,Y] or [B,X,Y,T], got {tuple(force.shape)}")
        if mu is not None and tuple(mu.shape) != (B,):
            raise RuntimeError(f"Grid2DMarkovExperiment: batch['mu'] must be [B], got {tuple(mu.shape)}")

    def _positions(self, X, Y, device, dtype):
        gx = torch.linspace(self.low, self.high, steps=X, device=device, dtype=dtype)
        gy = torch.linspace(self.low, self.high, steps=Y, device=device, dtype=dtype)
        return torch.stack(torch.meshgrid(gx, gy, indexing='ij'), dim=-1)

    def _mean_std(self):
        key = (self.normalizer.sum._version, self.normalizer.count._version, self.normalizer.sum.data_ptr())
        if self._ms_cache is None or self._ms_cache[0] != key:
            self._ms_cache = (key, self.normalizer.mean.tolist(), self.normalizer.std.tolist())
        return self._ms_cache[1], self._ms_cache[2]

    # -- training (routines/grid_2d_markov.py:122-193, :374-390; routines/base.py:27-52) -------------------------
    def _build_features(self, batch):
        """One-step input features [x, (q, v,) gx, gy, (f,) (mu)] of ``batch['x']`` [B,X,Y,1], normalised (the
        normaliser accumulates when in training mode), + noise_std * N(0, 1) (grid_2d_markov.py:122-170).  Feature
        assembly is host-side tensor plumbing; the velocities come from ffno_velocity_fwd."""
        x = batch['x']
        _ops.require_cuda(x, "Grid2DMarkovExperiment")
        B, X, Y, _ = x.shape
        self._check_extras(batch.get('f') if self.append_force else None,
                           batch.get('mu') if self.append_mu else None, B, X, Y)
        parts = [x]
        if self.use_velocity:
            q, v = _ops.velocity_features(x[..., 0], *self.domain_lengths)
            parts += [q.unsqueeze(-1), v.unsqueeze(-1)]
        parts.append(self._positions(X, Y, x.device, x.dtype).unsqueeze(0).expand(B, X, Y, 2))
        if self.append_force:
            if batch['f'].dim() != 3:
                raise RuntimeError("Grid2DMarkovExperiment: the training batch carries a static forcing batch['f'] [B,X,Y] "
                                   "(grid_2d_markov.py:146-148)")
            parts.append(batch['f'].unsqueeze(-1))
        if self.append_mu:
            parts.append(batch['mu'].reshape(B, 1, 1, 1).expand(B, X, Y, 1))
        feats = self.normalizer(torch.cat(parts, dim=-1))
        self._ms_cache = None
        if self.noise_std:
            feats = feats + torch.randn_like(feats) * self.noise_std
        return feats

    def _training_step(self, batch):
        """loss = LpLoss(de-normalised forecast, batch['y']) of one step (grid_2d_markov.py:172-193); differentiable
        through ffno_block_bwd / ffno_rel_l2_bwd."""
        x = self._build_features(batch)
        im = self.conv(x)['forecast']
        im = self.normalizer.inverse(im, channel=0)
        BN = im.shape[0]
        return self.l2_loss(im.reshape(BN, -1), batch['y'].reshape(BN, -1))

    def training_step(self, batch, batch_idx: int = 0, optimizer=None, scheduler=None, current_epoch: int = 1,
                      clip_val: Optional[float] = None, world_size: int = 1):
        """grid_2d_markov.py:374-390 + routines/base.py:27-52 without pytorch_lightning: epoch 0 only accumulates the
        normaliser statistics; afterwards loss -> zero_grad -> backward -> (data-parallel: one all-reduce of the
        gradients, fourierflow_b200.distributed.allreduce_gradients) -> clip -> optimizer.step -> scheduler.step."""
        if current_epoch == 0:
            with torch.no_grad():
                self._build_features(batch)
            return None
        loss = self._training_step(batch)
        self.optimize_manually(loss, batch_idx, optimizer, scheduler, clip_val, world_size)
        return loss

    # -- inference -------------------------------------------------------------------------------------
    def forward(self, data):
        return self._valid_step(data)

    def predict(self, data: torch.Tensor, n_steps: Optional[int] = None, force: Optional[torch.Tensor] = None,
                mu: Optional[torch.Tensor] = None) -> torch.Tensor:
        """preds[B,X,Y,n_steps] of the Markov rollout started from frame T−n_steps−1 of ``data``.  ``force``
        (batch['f']): [B,X,Y] or [B,X,Y,T'] whose last n_steps frames are used (grid_2d_markov.py:246-255);
        ``mu`` (batch['mu']): [B] (:257-260)."""
        _ops.require_cuda(data, "Grid2DMarkovExperiment")
        B, X, Y, T = data.shape
        self._check_extras(force, mu, B, X, Y)
        n_steps = n_steps or self.n_steps or T - 1
        if not 1 <= n_steps <= T - 1:      # a negative frame index below would silently wrap to the end of the series
            raise RuntimeError(f"Grid2DMarkovExperiment: a {n_steps}-step rollout needs {n_steps + 1} frames, "
                               f"data has T={T}")
        plan = self.conv.plan_for(data.device, (X, Y))
        mean, std = self._mean_std()
        frame0 = data[..., T - n_steps - 1].contiguous()
        if force is not None:
            if force.dim() == 4 and force.shape[-1] < n_steps:
                raise RuntimeError(f"Grid2DMarkovExperiment: batch['f'] has {force.shape[-1]} frames, the rollout needs "
                                   f"{n_steps}")
            force = (force.unsqueeze(-1) if force.dim() == 3 else force[..., -n_steps:]).to(data.dtype).contiguous()
        if mu is not None:
            mu = mu.to(data.dtype).contiguous()
        return plan.rollout_forward(frame0, n_steps, mean, std, self.low, self.high,
                                    self.domain_lengths if self.use_velocity else None, force=force, mu=mu)

    def per_sample_losses(self, preds: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
        """[n_steps, B] relative L2 of every rollout step against the last n_steps frames of ``data``."""
        n_steps = preds.shape[-1]
        yy = data[..., -n_steps:]
        return torch.stack([self.l2_loss.rel_per_sample(preds[..., t], yy[..., t]) for t in range(n_steps)])

    def _valid_step(self, batch):
        data = batch['data'] if isinstance(batch, dict) else batch
        force = batch.get('f') if isinstance(batch, dict) and self.append_force else None
        mu = batch.get('mu') if isinstance(batch, dict) and self.append_mu else None
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.conv.parameters()):
            _ops.require_inference(self.conv, data)
        preds = self.predict(data, force=force, mu=mu)
        losses = self.per_sample_losses(preds, data)            # [n_steps, B]
        step_losses = list(losses.mean(dim=1))                  # LpLoss(size_average) per step (:313)
        loss = losses.mean(dim=1).sum()                         # loss += l (:315)
        # one (empty) per-layer forecast list per step, as `pred_layer_list.append(out['forecast_list'])` builds it for
        # a conv without use_fork (grid_2d_markov.py:320-321)
        return loss, step_losses, preds, [[] for _ in range(preds.shape[-1])]
