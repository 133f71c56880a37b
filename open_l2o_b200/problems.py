"""Synthetic optimizees (the callers that feed the hot path).  Each ``problem()`` returns a zero-arg
``build()`` that creates its tensors through ``get_variable`` and returns a scalar loss, exactly like
DM/problems.py; gradients come from torch autograd on the device (the "external-gradient" regime) unless
the builder carries a ``fused`` spec, in which case the unroll kernel evaluates the separable gradient
in-kernel (the "fused" regime)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .variables import (constant_initializer, get_variable, ones_initializer, random_normal_initializer,
                        random_uniform_initializer)


@dataclass
class FusedSpec:
    kind: str        # "rastrigin_sep" | "quadratic_diag" | "quadratic_batch" (in-kernel, include/l2o_b200.h L2O_OPT_*)
                     # | "lasso_batch" (producer kernel l2o_lasso_grad: a = A / w, b = y, alpha = l1 weight)
    var: str         # name of the trainable variable
    a: str           # constant names
    b: str
    alpha: float = 10.0
    fscale: float = 1.0
    group: int = 0   # "quadratic_batch": coordinates per dense group
    extra: dict = None   # producer-specific structure ("mlp_xent": activation, number of layers)


def simple():
    """f(x) = x^2 (DM/problems.py:41-53)."""
    def build():
        x = get_variable("x", shape=[], initializer=ones_initializer())
        return torch.square(x)
    return build


def simple_multi_optimizer(num_dims=2):
    """DM/problems.py:56-70."""
    def build():
        coords = [get_variable("x_{}".format(i), shape=[], initializer=ones_initializer()) for i in range(num_dims)]
        x = torch.stack([c.reshape(()) for c in coords])
        return torch.sum(torch.square(x))
    return build


def quadratic(batch_size=128, num_dims=10, stddev=0.01):
    """f(x) = mean_b ||W_b x_b - y_b||^2 (DM/problems.py:73-101)."""
    def build():
        x = get_variable("x", shape=[batch_size, num_dims], initializer=random_normal_initializer(stddev=stddev))
        w = get_variable("w", shape=[batch_size, num_dims, num_dims], initializer=random_uniform_initializer(),
                         trainable=False)
        y = get_variable("y", shape=[batch_size, num_dims], initializer=random_uniform_initializer(), trainable=False)
        product = torch.bmm(w, x.unsqueeze(-1)).squeeze(-1)
        return torch.mean(torch.sum((product - y) ** 2, dim=1))
    # dense W_b x_b evaluated in-kernel: the whole unroll is one launch.  Only where the
    # kernels are throughput-bound: at BASELINE config #1's 1,280 coordinates a thread walks the whole LSTM serially
    # either way, and the graph-captured step-at-a-time path is measured faster (1.29 vs 1.63 ms per unroll).
    if num_dims <= 128 and batch_size * num_dims >= 16384:
        build.fused = FusedSpec("quadratic_batch", "x", "w", "y", fscale=1.0 / batch_size, group=num_dims)
    return build


def _lasso_loss(x, w, y, l):
    product = torch.bmm(w, x.unsqueeze(-1))
    left = 0.5 * torch.sum((product - y) ** 2, dim=1)
    other = l * torch.sum(torch.abs(x), dim=1, keepdim=True)
    return torch.mean(left + other)


def lasso(batch_size=128, num_dims=10, stddev=0.01, l=0.005):
    """DM/problems.py:103-135."""
    def build():
        x = get_variable("x", shape=[batch_size, num_dims], initializer=random_normal_initializer(stddev=stddev))
        w = get_variable("w", shape=[batch_size, num_dims, num_dims], initializer=random_uniform_initializer(),
                         trainable=False)
        y = get_variable("y", shape=[batch_size, num_dims, 1], initializer=random_uniform_initializer(),
                         trainable=False)
        return _lasso_loss(x, w, y, l)
    build.fused = FusedSpec("lasso_batch", "x", "w", "y", alpha=float(l))   # producer kernel: f and df/dx in one launch
    return build


def lasso_fixed(data_A, data_b, stddev=0.01, l=0.005):
    """DM/problems.py:137-175: A [B, m, n], b [B, m, 1]."""
    a = torch.as_tensor(data_A, dtype=torch.float32)
    b = torch.as_tensor(data_b, dtype=torch.float32)

    def build():
        x = get_variable("x", shape=[a.shape[0], a.shape[2]], initializer=random_normal_initializer(stddev=stddev))
        w = get_variable("w", shape=list(a.shape), initializer=constant_initializer(a), trainable=False)
        y = get_variable("y", shape=list(b.shape), initializer=constant_initializer(b), trainable=False)
        return _lasso_loss(x, w, y, l)
    build.fused = FusedSpec("lasso_batch", "x", "w", "y", alpha=float(l))   # producer kernel: f and df/dx in one launch
    return build


def rastrigin(batch_size=128, num_dims=10, alpha=10, stddev=1):
    """Dense Rastrigin family (DM/problems.py:177-213)."""
    def build():
        x = get_variable("x", shape=[batch_size, num_dims, 1], initializer=random_normal_initializer(stddev=stddev))
        A = get_variable("A", shape=[batch_size, num_dims, num_dims],
                         initializer=random_normal_initializer(stddev=stddev), trainable=False)
        B = get_variable("B", shape=[batch_size, num_dims, 1], initializer=random_normal_initializer(stddev=stddev),
                         trainable=False)
        Cc = get_variable("C", shape=[batch_size, num_dims, 1], initializer=random_normal_initializer(stddev=stddev),
                          trainable=False)
        product = torch.bmm(A, x)
        ras_norm2 = torch.sum((product - B) ** 2, dim=(-2, -1))
        cq = torch.bmm(Cc.transpose(1, 2), torch.cos(2 * math.pi * x)).reshape(-1)
        return torch.mean(0.5 * ras_norm2 - alpha * cq + alpha * num_dims)
    return build


def rastrigin_separable(num_dims=1000000, alpha=10.0, stddev=1.0, normalize=True, shard=None):
    """The A = I member of DM/problems.py:177-213 with batch 1 - the only member that exists at d = 1e6
    (a dense A would be 4 TB).  f = fscale * sum_i (0.5 (x_i-b_i)^2 - alpha c_i cos(2 pi x_i) + alpha).

    ``shard=(lo, hi)``: this rank's contiguous slice of the ``num_dims`` coordinates (SURVEY.md 8(e), strong scaling of
    BASELINE config #5).  The initializers draw the full tensors and keep [lo, hi), so the union over the ranks is
    exactly the single-GPU problem for the same seed; ``fscale`` stays 1/num_dims (the GLOBAL count) and the ranks'
    partial f values are summed by the meta-step's all-reduce."""
    fscale = 1.0 / num_dims if normalize else 1.0
    two_pi = 6.2831855  # fp32(2 pi), the constant the kernel uses
    lo, hi = shard if shard is not None else (0, num_dims)
    n_loc = hi - lo

    def sliced(init):
        if shard is None:
            return init
        return lambda shape, gen: init((num_dims,), gen)[lo:hi].clone()

    def build():
        normal = sliced(random_normal_initializer(stddev=stddev))
        x = get_variable("x", shape=[n_loc], initializer=normal)
        b = get_variable("b", shape=[n_loc], initializer=normal, trainable=False)
        c = get_variable("c", shape=[n_loc], initializer=normal, trainable=False)
        fi = 0.5 * (x - b) ** 2 - alpha * c * torch.cos(two_pi * x) + alpha
        return fscale * torch.sum(fi)
    build.fused = FusedSpec("rastrigin_sep", "x", "b", "c", alpha=float(alpha), fscale=fscale)
    return build


def quadratic_diag(num_dims=1280, stddev=0.01, normalize=True):
    """DM/problems.py:73-101 with diagonal W: f = fscale * sum_i (w_i x_i - y_i)^2."""
    fscale = 1.0 / num_dims if normalize else 1.0

    def build():
        x = get_variable("x", shape=[num_dims], initializer=random_normal_initializer(stddev=stddev))
        w = get_variable("w", shape=[num_dims], initializer=random_uniform_initializer(0.5, 1.5), trainable=False)
        y = get_variable("y", shape=[num_dims], initializer=random_uniform_initializer(), trainable=False)
        return fscale * torch.sum((w * x - y) ** 2)
    build.fused = FusedSpec("quadratic_diag", "x", "w", "y", fscale=fscale)
    return build


def mlp(layers=(100,), in_dim=784, n_classes=10, batch_size=128, activation="sigmoid", init_stddev=0.01):
    """Sigmoid/ReLU MLP with softmax cross-entropy on a fixed synthetic batch (shape of DM/problems.py:254-288;
    the data is synthetic because MNIST cannot be downloaded here)."""
    act = {"sigmoid": torch.sigmoid, "relu": torch.relu}[activation]

    def build():
        data = get_variable("data", shape=[batch_size, in_dim], initializer=random_uniform_initializer(),
                            trainable=False)
        labels = get_variable("labels", shape=[batch_size],
                              initializer=lambda shape, gen: torch.randint(0, n_classes, shape, generator=gen).float(),
                              trainable=False)
        h, k = data, in_dim
        for i, width in enumerate(tuple(layers) + (n_classes,)):
            w = get_variable("mlp/linear_{}/w".format(i), shape=[k, width],
                             initializer=random_normal_initializer(stddev=init_stddev))
            b = get_variable("mlp/linear_{}/b".format(i), shape=[width],
                             initializer=random_normal_initializer(stddev=init_stddev))
            h = h @ w + b
            if i < len(layers):
                h = act(h)
            k = width
        return torch.nn.functional.cross_entropy(h, labels.long())
    # analytic producer: f and df/dx without the autograd engine (mlp_value_and_grad below)
    build.fused = FusedSpec("mlp_xent", "mlp", "data", "labels", extra=dict(activation=activation,
                                                                           n_layers=len(tuple(layers)) + 1))
    return build


def mlp_value_and_grad(params, data, labels, activation, grads_out):
    """Loss and gradient of the `mlp` optimizee written out by hand (the optimizee step feeding the hot path, SURVEY.md
    8(f) row 4): the forward pass keeps the activations, the backward pass writes every dW = h^T dz / db = sum(dz)
    STRAIGHT INTO the caller's gradient views - about a dozen launches (3 GEMMs per layer, fused elementwise) instead of
    the ~30 of the autograd engine with its per-variable copies.  params / grads_out: [w0, b0, w1, b1, ...]."""
    n_layers = len(params) // 2
    hs, h = [data], data
    for i in range(n_layers):
        z = torch.addmm(params[2 * i + 1], h, params[2 * i])
        if i < n_layers - 1:
            h = torch.sigmoid(z) if activation == "sigmoid" else torch.relu(z)
            hs.append(h)
    lab = labels.long()
    logp = torch.log_softmax(z, dim=1)
    loss = -logp.gather(1, lab.unsqueeze(1)).mean()
    dz = torch.exp(logp)
    dz.scatter_add_(1, lab.unsqueeze(1), torch.full((lab.numel(), 1), -1.0, device=dz.device, dtype=dz.dtype))
    dz.mul_(1.0 / lab.numel())
    for i in reversed(range(n_layers)):
        torch.mm(hs[i].t(), dz, out=grads_out[2 * i])
        torch.sum(dz, dim=0, out=grads_out[2 * i + 1])
        if i > 0:
            dh = dz @ params[2 * i].t()
            dz = dh * hs[i] * (1.0 - hs[i]) if activation == "sigmoid" else dh * (hs[i] > 0).to(dh.dtype)
    return loss
