"""Lazily built, weight-tracking ``Engine`` for an ``nn.Module`` whose parameters use the
reference's names.  The module stays the owner of the weights (``state_dict`` /
``load_state_dict`` work as in MMPose); the engine keeps packed copies and is refreshed when
a parameter's version counter or the module's device changes."""
from __future__ import annotations

import os
from typing import Dict

import torch

from ..engine import Engine

DEFAULT_PRECISION = os.environ.get("PROBPOSE_B200_PRECISION", "fp16x3")


class EngineCache:
    def __init__(self, engine_kwargs: dict, precision: str):
        self.kwargs = dict(engine_kwargs)
        self.precision = precision or DEFAULT_PRECISION
        self.engine = None
        self.stamp = None
        self.capacity = 0

    def get(self, named_tensors: Dict[str, torch.Tensor], images: int, device: torch.device) -> Engine:
        """``images``: how many images the call pushes through the network - batch x passes, i.e. TWICE the
        batch for a flip_test call.  The engine is rebuilt (next power of two) when a call needs more; the C side
        admits passes x batch <= 2 x max_batch, so ``capacity`` (in images) is exactly that bound."""
        stamp = (str(device),) + tuple((n, t._version, t.data_ptr()) for n, t in named_tensors.items())
        if self.engine is None or images > self.capacity or device != self.engine.device:
            cap = max(32, 1 << (max(images, 1) - 1).bit_length())
            self.engine = None  # release the old workspace first
            self.engine = Engine(precision=self.precision, max_batch=cap // 2, device=device, **self.kwargs)
            self.capacity = cap
            self.stamp = None
        if stamp != self.stamp:
            self.engine.load_state_dict(named_tensors)
            self.stamp = stamp
        return self.engine
