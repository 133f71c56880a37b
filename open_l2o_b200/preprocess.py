"""Gradient preprocessing modules behind the reference's ``preprocess`` surface (DM/preprocess.py)."""
from __future__ import annotations

import torch

from . import engine as _engine


class Clamp(object):
    """DM/preprocess.py:26-39 (plain clamp; elementwise helper, not on the fused path)."""

    def __init__(self, min_value=None, max_value=None, name="clamp"):
        self._min, self._max = min_value, max_value

    def __call__(self, inputs):
        out = inputs
        if self._min is not None:
            out = torch.clamp(out, min=self._min)
        if self._max is not None:
            out = torch.clamp(out, max=self._max)
        return out


class LogAndSign(object):
    """Log and sign preprocessing (DM/preprocess.py:42-70).  ``initializer`` is accepted and ignored exactly
    as in the reference's constructor."""

    def __init__(self, initializer=None, k=5, name="preprocess_log"):
        self._k = k

    def __call__(self, gradients):
        """[d_1..d_n] -> [d_1..d_{n-1}, 2 d_n]: log part first, then sign part."""
        g = gradients.contiguous()
        out = _engine.log_and_sign(g.reshape(-1), float(self._k))      # [2, numel]
        lo = out[0].reshape(g.shape)
        sg = out[1].reshape(g.shape)
        return torch.cat([lo, sg], dim=g.dim() - 1)
