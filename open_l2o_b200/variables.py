"""Variable creation hook: the eager counterpart of the reference's ``mock.patch("tensorflow.get_variable")``
trick (DM/meta.py:88-155).  An optimizee's ``make_loss()`` obtains its tensors through
``get_variable``; the MetaOptimizer installs a getter that first *captures* them
(``_get_variables``, DM/meta.py:102-128) and later *substitutes* the unrolled tensors in creation order
(``_make_with_custom_variables``, DM/meta.py:131-155)."""
from __future__ import annotations

import contextlib
import threading

import torch

_tls = threading.local()


@contextlib.contextmanager
def variable_getter(getter):
    prev = getattr(_tls, "getter", None)
    _tls.getter = getter
    try:
        yield
    finally:
        _tls.getter = prev


def get_variable(name, shape=(), dtype=torch.float32, initializer=None, trainable=True, **kwargs):
    if "custom_getter" in kwargs:  # DM/meta.py:92-94 (the reference's guard never fires; ours does)
        raise AttributeError("Custom getters are not supported for optimizee variables.")
    getter = getattr(_tls, "getter", None)
    if getter is None:
        raise RuntimeError("get_variable() must be called from a make_loss() run by a MetaOptimizer")
    return getter(name, tuple(int(s) for s in shape), dtype, initializer, trainable)


# ---- initializers: callables (shape, generator) -> CPU fp32 tensor --------------------------------
def ones_initializer():
    return lambda shape, gen: torch.ones(shape)


def zeros_initializer():
    return lambda shape, gen: torch.zeros(shape)


def constant_initializer(value):
    def init(shape, gen):
        t = torch.as_tensor(value, dtype=torch.float32)
        return t.reshape(shape).clone() if t.numel() == int(torch.Size(shape).numel()) else t.expand(shape).clone()
    return init


def random_normal_initializer(mean=0.0, stddev=1.0):
    return lambda shape, gen: torch.randn(shape, generator=gen) * stddev + mean


def random_uniform_initializer(minval=0.0, maxval=1.0):
    return lambda shape, gen: torch.rand(shape, generator=gen) * (maxval - minval) + minval
