"""The inference-side surface of the reference's ``Routine`` base class (fourierflow/routines/base.py:9-102) without
pytorch_lightning: what ``commands/predict.py:91-105`` and ``commands/train.py:125-148`` call on a routine after
training — ``load_lightning_model_state``, ``convert_data``, ``warmup``, ``infer`` — plus ``optimize_manually`` for
the training steps of the routines (the optimizer / scheduler objects are the caller's: ``configure_optimizers`` and the
Lightning trainer stay outside the hot path)."""
from __future__ import annotations

from typing import Any, Dict

import torch

#: buffers of the reference's ``use_velocity`` routines that this backend computes on the fly (ffno_velocity_fwd) and
#: that the reference itself drops on load "to run super-resolution evaluations" (routines/base.py:88-99)
REMOVE_KEYS = ['kx', 'ky', 'lap'] + [f'{n}_{s}' for s in (32, 64, 128, 256) for n in ('kx', 'ky', 'lap')]


class RoutineMixin:
    def warmup(self) -> None:                       # routines/base.py:24
        pass

    def optimize_manually(self, loss, batch_idx: int = 0, optimizer=None, scheduler=None, clip_val=None,
                          world_size: int = 1):
        """routines/base.py:27-52 (accumulate_grad_batches = 1) without pytorch_lightning: zero_grad -> backward ->
        (data-parallel: one all-reduce of the gradients) -> clip -> step -> scheduler step."""
        if optimizer is None:
            return
        optimizer.zero_grad()
        loss.backward()
        if world_size > 1:
            from ..distributed import allreduce_gradients
            allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]], world_size)
        if clip_val:
            for group in optimizer.param_groups:
                torch.nn.utils.clip_grad_value_(group["params"], clip_val)
        optimizer.step()
        if scheduler is not None:
            scheduler.step()

    def infer(self, data):                          # routines/base.py:54-56
        with torch.no_grad():
            return self.forward(data)

    def convert_data(self, data: Dict[str, Any]) -> Dict[str, torch.Tensor]:      # routines/base.py:58-60
        return {k: torch.from_numpy(v).cuda() for k, v in data.items()}

    def load_lightning_model_state(self, checkpoint_path, map_location=None) -> None:
        """Load ``checkpoint['state_dict']`` of a Lightning checkpoint written by the reference (routines/base.py:79-102):
        same key names (``conv.spectral_layers.3.backcast_ff.layers.0.0.weight_g``, ``normalizer.sum``, …), the
        ``kx*/ky*/lap*`` buffers removed, ``strict`` relaxed only when something was removed."""
        if isinstance(checkpoint_path, dict):
            checkpoint = checkpoint_path
        else:
            checkpoint = torch.load(checkpoint_path, map_location=map_location or (lambda storage, loc: storage),
                                    weights_only=False)
        state_dict = dict(checkpoint['state_dict'])
        strict = True
        for key in REMOVE_KEYS:
            if key in state_dict:
                del state_dict[key]
                strict = False
        self.load_state_dict(state_dict, strict=strict)
        if hasattr(self, '_ms_cache'):
            self._ms_cache = None
