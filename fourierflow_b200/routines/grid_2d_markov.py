"""``Grid2DMarkovExperiment`` — host-side mirror of fourierflow/routines/grid_2d_markov.py (inference path).

The reference class is a pytorch_lightning ``Routine``; this mirror is a plain ``nn.Module`` with the same
constructor keywords and the same ``forward(batch) -> (loss, step_losses, preds, pred_layer_list)``
contract for the configurations the BASELINE names (torus_li/markov: position features + normaliser;
torus_kochkov: additionally ``use_velocity`` — stream-function velocities recomputed from the vorticity at
every step) and for the torus_vis / torus_vis_force feature sets (``append_force``: the batch's ``'f'`` forcing,
static [B,X,Y] or time-varying [B,X,Y,T]; ``append_mu``: the batch's ``'mu'`` viscosity per sample).  The step loop
(``_valid_step``, grid_2d_markov.py:195-326) of these runs entirely in libffno_b200 (ffno_rollout_fwd): feature
build → normalise → layer stack → de-normalise, feeding each forecast back, with no host round trip
between steps.  The ablation switches of five shipped configs (``use_position=False``, ``should_normalize=False``,
``shuffle_grid``, ``learn_difference``) take a step loop on the host instead: features by tensor plumbing, one
``ffno_block_fwd`` per step (``_predict_stepwise``).  ``use_fourier_position`` raises: the reference cannot run it either
(``self.k_max`` is never defined, grid_2d_markov.py:116).  The relative-L2 reduction (modules/loss.py:33-46) is ffno_rel_l2 per sample; the mean over
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
        if use_fourier_position:
            raise RuntimeError("Grid2DMarkovExperiment: use_fourier_position=True cannot run in the reference either "
                               "(encode_positions reads self.k_max, which is never set: grid_2d_markov.py:116)")
        want = 1 + (2 if use_velocity else 0) + (2 if use_position else 0) + (1 if append_force else 0) + \
            (1 if append_mu else 0)
        if conv.input_dim != want:
            raise RuntimeError(f"Grid2DMarkovExperiment: use_velocity={use_velocity}, use_position={use_position}, "
                               f"append_force={append_force}, append_mu={append_mu} build {want} input features, "
                               f"conv.input_dim is {conv.input_dim}")
        self.use_position, self.should_normalize = use_position, should_normalize
        self.shuffle_grid, self.learn_difference = shuffle_grid, learn_difference
        # the fused device rollout (ffno_rollout_fwd) covers position features + normaliser; the ablation switches run
        # the step loop on the host around ffno_block_fwd (_predict_stepwise)
        self._fused = use_position and should_normalize and not shuffle_grid and not learn_difference
        if shuffle_grid:                               # grid_2d_markov.py:75-80 (plain attributes there too: not in checkpoints)
            grid_size = kwargs.get("grid_size", [64])
            if len(grid_size) != 1:
                raise RuntimeError("shuffle_grid only supports one size")
            self.x_idx = torch.randperm(grid_size[0])
            self.x_inv = torch.argsort(self.x_idx)
            self.y_idx = torch.randperm(grid_size[0])
            self.y_inv = torch.argsort(self.y_idx)
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
        if self.use_position:
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
        if self.use_position:
            parts.append(self._positions(X, Y, x.device, x.dtype).unsqueeze(0).expand(B, X, Y, 2))
        if self.append_force:
            if batch['f'].dim() != 3:
                raise RuntimeError("Grid2DMarkovExperiment: the training batch carries a static forcing batch['f'] [B,X,Y] "
                                   "(grid_2d_markov.py:146-148)")
            parts.append(batch['f'].unsqueeze(-1))
        if self.append_mu:
            parts.append(batch['mu'].reshape(B, 1, 1, 1).expand(B, X, Y, 1))
        feats = torch.cat(parts, dim=-1)
        if self.should_normalize:
            feats = self.normalizer(feats)
        self._ms_cache = None
        if self.noise_std:
            feats = feats + torch.randn_like(feats) * self.noise_std
        return feats

    def _training_step(self, batch):
        """loss = LpLoss(de-normalised forecast, batch['y']) of one step (grid_2d_markov.py:172-193); differentiable
        through ffno_block_bwd / ffno_rel_l2_bwd."""
        x = self._build_features(batch)
        if self.shuffle_grid:                                     # :177-183
            x = x[:, self.x_idx][:, :, self.y_idx]
        im = self.conv(x)['forecast']
        if self.shuffle_grid:
            im = im[:, :, self.y_inv][:, self.x_inv]
        if self.should_normalize:
            im = self.normalizer.inverse(im, channel=0)
        BN = im.shape[0]
        targets = batch['dy'] if self.learn_difference else batch['y']      # :190
        return self.l2_loss(im.reshape(BN, -1), targets.reshape(BN, -1))

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
        if not self._fused:
            return self._predict_stepwise(data, n_steps, force, mu)[0]
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

    @torch.no_grad()
    def _predict_stepwise(self, data, n_steps, force, mu):
        """The step loop of grid_2d_markov.py:263-321 on the host for the ablation switches: per step the features are
        assembled by tensor plumbing, the operator runs through ffno_block_fwd, and the forecast is fed back.
        -> (preds [B,X,Y,n_steps], outs [B,X,Y,n_steps]: what the loss of step t compares with its target — the forecast,
        or the predicted increment under learn_difference)."""
        B, X, Y, T = data.shape
        pos = self._positions(X, Y, data.device, data.dtype).unsqueeze(0).expand(B, X, Y, 2) if self.use_position else None
        im = data[..., T - n_steps - 1].unsqueeze(-1)
        prev_im = im                                              # :266 (the un-normalised first input frame)
        preds, outs = [], []
        for t in range(n_steps):
            parts = [im]
            if self.use_velocity:
                q, v = _ops.velocity_features(im[..., 0].contiguous(), *self.domain_lengths)
                parts += [q.unsqueeze(-1), v.unsqueeze(-1)]
            if pos is not None:
                parts.append(pos)
            if force is not None:
                parts.append((force if force.dim() == 3 else force[..., -n_steps:][..., t]).unsqueeze(-1))
            if mu is not None:
                parts.append(mu.reshape(B, 1, 1, 1).expand(B, X, Y, 1))
            x = torch.cat(parts, dim=-1)
            if self.should_normalize:
                x = (x - self.normalizer.mean) / self.normalizer.std
            if self.shuffle_grid:
                x = x[:, self.x_idx][:, :, self.y_idx]
            im = self.conv(x.contiguous())['forecast']
            if self.shuffle_grid:
                im = im[:, :, self.y_inv][:, self.x_inv]
            if self.should_normalize:
                im = self.normalizer.inverse(im, channel=0)
            outs.append(im)
            if self.learn_difference:                             # :316-318
                im = prev_im + im
                prev_im = im
            preds.append(im)
        return torch.cat(preds, dim=-1), torch.cat(outs, dim=-1)

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
        if self._fused:
            preds = self.predict(data, force=force, mu=mu)
            losses = self.per_sample_losses(preds, data)            # [n_steps, B]
        else:
            _ops.require_cuda(data, "Grid2DMarkovExperiment")
            B, X, Y, T = data.shape
            self._check_extras(force, mu, B, X, Y)
            n_steps = self.n_steps or T - 1
            if not 1 <= n_steps <= T - 1:
                raise RuntimeError(f"Grid2DMarkovExperiment: a {n_steps}-step rollout needs {n_steps + 1} frames, data has T={T}")
            preds, outs = self._predict_stepwise(data, n_steps, force, mu)
            yy = data[..., -n_steps:]
            # learn_difference compares the increment with yy[t] - yy[t-1]; at t = 0 the reference's index t - 1 = -1
            # wraps to the LAST frame (grid_2d_markov.py:309-310) — kept as it is
            tgt = [yy[..., t] - yy[..., t - 1] if self.learn_difference else yy[..., t] for t in range(n_steps)]
            losses = torch.stack([self.l2_loss.rel_per_sample(outs[..., t], tgt[t]) for t in range(n_steps)])
        step_losses = list(losses.mean(dim=1))                  # LpLoss(size_average) per step (:313)
        loss = losses.mean(dim=1).sum()                         # loss += l (:315)
        # one (empty) per-layer forecast list per step, as `pred_layer_list.append(out['forecast_list'])` builds it for
        # a conv without use_fork (grid_2d_markov.py:320-321)
        return loss, step_losses, preds, [[] for _ in range(preds.shape[-1])]
