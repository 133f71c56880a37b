"""CPU oracle for the L2O-DM / L2O-RNNProp coordinate-wise LSTM hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``open_l2o_b200``) never does.

It is an op-for-op fp32 (optionally fp64) restatement, on torch-CPU tensors, of the
reference's algorithm.  ``DM/`` abbreviates
``/root/reference/Model_Free_L2O/L2O-DM and L2O-RNNProp/``:

  * ``log_and_sign``            DM/preprocess.py:52-70 (+ Clamp :33-39)
  * ``lstm_cell``               snt.LSTM (dm-sonnet==1.11, NOT vendored in the reference; call site
                                DM/networks.py:197).  Semantics restated from the published Sonnet 1.x
                                ``gated_rnn.LSTM``: gates = [x|h]·w_gates + b_gates, split order i, j, f, o,
                                c' = sigmoid(f + 1.0)*c + sigmoid(i)*tanh(j), h' = tanh(c')*sigmoid(o),
                                state order (hidden, cell) (confirmed by DM/meta_dm_train.py:433-434).
  * ``net_apply``               DM/networks.py:207-232 (StandardDeepLSTM._build), :254-271 (coordinate-wise
                                reshape), :287-300 (RNNprop input stacking order (m, g))
  * ``adam_features``           DM/meta_rnnprop_train.py:383-388
  * ``unroll`` / ``meta_loss``  DM/meta.py:319-376 (update, time_step, while_loop, loss = sum of T+1 fx)
  * ``imitation_loss``          DM/meta_dm_train.py:463-480
  * ``tf_adam_step``            tf.train.AdamOptimizer as called at DM/meta.py:412
  * problem formulas            DM/problems.py:41-53 (simple), :73-101 (quadratic), :103-135 (lasso),
                                :177-213 (rastrigin; the separable member A=I is what scales to 1e6 dims)

PARITY PINNING.  TensorFlow 1.14 / Sonnet 1.11 cannot run in this container, so the oracle is
pinned against the reference's own known-answer tests instead (tests/test_oracle_golden.py):
SW/meta_test.py:50-69 (cost 0.7325327, x 0.8559), SW/networks_test.py:57-69 (zero-initialised
net => zero update), SW/preprocess_test.py:78-98, SW/problems_test.py:99-111.  None of those
exercises the LSTM gate arithmetic itself, so at the ``snt.LSTM`` boundary this oracle is
**parity unpinned** (see DESIGN.md); its gate order / forget bias follow the published
Sonnet 1.11 source and are cross-checked only by hand-computed single-cell vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

EPS_F32 = float(np.finfo(np.float32).eps)  # DM/preprocess.py:63

PRE_IDENTITY = "identity"
PRE_LOGSIGN = "LogAndSign"
PRE_FC = "fc"


@dataclass
class NetSpec:
    """Mirror of StandardDeepLSTM's constructor arguments (DM/networks.py:157-159)."""
    layers: Tuple[int, ...] = (20, 20)
    preprocess_name: str = PRE_IDENTITY
    preprocess_options: dict = field(default_factory=dict)
    scale: float = 1.0
    tanh_output: bool = False
    rnnprop: bool = False  # True: net(m, g, state) with two inputs (DM/networks.py:279-300)

    @property
    def n_in(self) -> int:
        return 2 if self.rnnprop else 1

    @property
    def feat(self) -> int:
        if self.preprocess_name == PRE_FC:
            return int(self.preprocess_options["dim"])
        if self.preprocess_name == PRE_LOGSIGN:
            return 2 * self.n_in
        return self.n_in

    def shapes(self) -> List[Tuple[str, str, Tuple[int, ...]]]:
        """(module, variable, shape) in Sonnet creation order == flat theta order."""
        out = []
        if self.preprocess_name == PRE_FC:
            out.append(("input_projection", "w", (self.n_in, self.feat)))
            out.append(("input_projection", "b", (self.feat,)))
        k = self.feat
        for i, h in enumerate(self.layers, start=1):
            out.append((f"lstm_{i}", "w_gates", (k + h, 4 * h)))
            out.append((f"lstm_{i}", "b_gates", (4 * h,)))
            k = h
        out.append(("linear", "w", (k, 1)))
        out.append(("linear", "b", (1,)))
        return out

    def n_theta(self) -> int:
        return sum(int(np.prod(s)) for _, _, s in self.shapes())


def trunc_normal(shape, std, gen, dtype=torch.float32):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t.to(dtype)


def init_theta(spec: NetSpec, seed: int = 0, initializer=None, out_gain: float = 1.0,
               dtype=torch.float32) -> torch.Tensor:
    """Flat theta.  Default init follows Sonnet 1.x: TruncatedNormal(1/sqrt(fan_in)) for
    w_gates/b_gates/linear.w, zeros for linear biases.  ``initializer`` may be "zeros" or a
    ``{module: {var: ndarray}}`` dict (the .l2l format, DM/networks.py:47-62)."""
    gen = torch.Generator().manual_seed(seed)
    parts = []
    for mod, var, shp in spec.shapes():
        if isinstance(initializer, str):
            if initializer != "zeros":
                raise ValueError(initializer)
            t = torch.zeros(shp, dtype=dtype)
        elif isinstance(initializer, dict) and mod in initializer and var in initializer[mod]:
            t = torch.as_tensor(np.asarray(initializer[mod][var]), dtype=dtype).reshape(shp)
        else:
            if mod.startswith("lstm"):
                t = trunc_normal(shp, 1.0 / math.sqrt(spec_fan_in(spec, mod)), gen, dtype)
            elif var == "w":
                t = trunc_normal(shp, 1.0 / math.sqrt(shp[0]), gen, dtype)
                if mod == "linear":
                    t = t * out_gain
            else:
                t = torch.zeros(shp, dtype=dtype)
        parts.append(t.reshape(-1))
    return torch.cat(parts)


def spec_fan_in(spec: NetSpec, mod: str) -> int:
    for m, v, s in spec.shapes():
        if m == mod and v == "w_gates":
            return s[0]
    raise KeyError(mod)


def unpack_theta(spec: NetSpec, theta: torch.Tensor) -> dict:
    out, off = {}, 0
    for mod, var, shp in spec.shapes():
        n = int(np.prod(shp))
        out.setdefault(mod, {})[var] = theta[off:off + n].reshape(shp)
        off += n
    assert off == theta.numel(), (off, theta.numel())
    return out


# --------------------------------------------------------------------------- preprocess
def log_and_sign(g: torch.Tensor, k: float) -> torch.Tensor:
    """DM/preprocess.py:52-70.  g: [..., d] -> [..., 2d] (log part first, then sign part)."""
    eps = torch.tensor(EPS_F32, dtype=g.dtype)
    log = torch.log(torch.abs(g) + eps)
    clamped_log = torch.clamp(log / k, min=-1.0)
    sign = torch.clamp(g * torch.tensor(float(np.exp(k)), dtype=g.dtype), min=-1.0, max=1.0)
    return torch.cat([clamped_log, sign], dim=-1)


def adam_features(g, m, v, p: float, beta1: float, beta2: float):
    """DM/meta_rnnprop_train.py:383-388.  p = float(step + t).  Returns m', v', m~, g~."""
    m_next = beta1 * m + (1.0 - beta1) * g
    m_hat = m_next / (1 - beta1 ** p)
    v_next = beta2 * v + (1.0 - beta2) * g * g
    v_hat = v_next / (1 - beta2 ** p)
    m_tilde = m_hat / (torch.sqrt(v_hat) + 1e-8)
    g_tilde = g / (torch.sqrt(v_hat) + 1e-8)
    return m_next, v_next, m_tilde, g_tilde


# --------------------------------------------------------------------------- the cell
def lstm_cell(x, h, c, w_gates, b_gates):
    """snt.LSTM (Sonnet 1.11) restated; see module docstring."""
    z = torch.cat([x, h], dim=1) @ w_gates + b_gates
    hsz = h.shape[1]
    i, j, f, o = z[:, 0:hsz], z[:, hsz:2 * hsz], z[:, 2 * hsz:3 * hsz], z[:, 3 * hsz:4 * hsz]
    c_next = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    h_next = torch.tanh(c_next) * torch.sigmoid(o)
    return h_next, c_next


def initial_state(spec: NetSpec, n: int, dtype=torch.float32):
    """snt.DeepRNN.initial_state: zeros, tuple over layers of (hidden, cell)."""
    return tuple((torch.zeros(n, h, dtype=dtype), torch.zeros(n, h, dtype=dtype)) for h in spec.layers)


def net_apply(spec: NetSpec, theta: torch.Tensor, inputs: torch.Tensor, state):
    """inputs: [N, n_in] raw features ([g] or [m~, g~]).  Returns (delta [N], next_state).
    DM/networks.py:207-232."""
    w = unpack_theta(spec, theta)
    if spec.preprocess_name == PRE_FC:
        u = torch.nn.functional.elu(inputs @ w["input_projection"]["w"] + w["input_projection"]["b"])
    elif spec.preprocess_name == PRE_LOGSIGN:
        u = log_and_sign(inputs.unsqueeze(-1), spec.preprocess_options["k"]).reshape(inputs.shape[0], -1)
    elif spec.preprocess_name == PRE_IDENTITY:
        u = inputs
    else:
        raise ValueError(spec.preprocess_name)
    out, nxt = u, []
    for li, _ in enumerate(spec.layers, start=1):
        h, c = state[li - 1]
        hn, cn = lstm_cell(out, h, c, w[f"lstm_{li}"]["w_gates"], w[f"lstm_{li}"]["b_gates"])
        nxt.append((hn, cn))
        out = hn
    y = (out @ w["linear"]["w"] + w["linear"]["b"]).reshape(-1)
    delta = torch.tanh(y) * spec.scale if spec.tanh_output else y * spec.scale
    return delta, tuple(nxt)


# --------------------------------------------------------------------------- optimizees
@dataclass
class FusedProblem:
    """Separable optimizees whose gradient the CUDA unroll evaluates in-kernel."""
    kind: str                 # "rastrigin_sep" | "quadratic_diag"
    a: torch.Tensor           # rastrigin: b ; quadratic: w
    b: torch.Tensor           # rastrigin: c ; quadratic: y
    alpha: float = 10.0
    fscale: float = 1.0

    def f_and_g(self, x):
        dt = x.dtype
        if self.kind == "rastrigin_sep":
            # DM/problems.py:177-213 with A = I, batch 1: 0.5||x-b||^2 - alpha c.cos(2 pi x) + alpha d
            two_pi = torch.tensor(2.0 * math.pi, dtype=torch.float32).to(dt)
            ang = two_pi * x
            fi = 0.5 * (x - self.a) ** 2 - self.alpha * self.b * torch.cos(ang) + self.alpha
            gi = (x - self.a) + (two_pi * self.alpha) * self.b * torch.sin(ang)
        elif self.kind == "quadratic_diag":
            r = self.a * x - self.b
            fi = r * r
            gi = 2.0 * self.a * r
        else:
            raise ValueError(self.kind)
        return self.fscale * fi.to(torch.float64).sum().to(dt), self.fscale * gi


def quadratic_f(x, w, y):
    """DM/problems.py:73-101. x [B,d], w [B,d,d], y [B,d]."""
    product = torch.bmm(w, x.unsqueeze(-1)).squeeze(-1)
    return torch.mean(torch.sum((product - y) ** 2, dim=1))


def lasso_f(x, w, y, l=0.005):
    """DM/problems.py:103-135 / :137-175. x [B,n], w [B,m,n], y [B,m,1]."""
    product = torch.bmm(w, x.unsqueeze(-1))
    left = 0.5 * torch.sum((product - y) ** 2, dim=1)
    other = l * torch.sum(torch.abs(x), dim=1, keepdim=True)
    return torch.mean(left + other)


# --------------------------------------------------------------------------- the unroll
@dataclass
class UnrollResult:
    fx: torch.Tensor          # [T+1]
    loss: torch.Tensor        # scalar = sum fx
    x_final: torch.Tensor
    state_final: tuple
    mv_final: Optional[tuple]
    deltas: List[torch.Tensor]
    grads: List[torch.Tensor]  # g_0..g_{T-1} (inputs fed to the net)


def unroll(spec: NetSpec, theta, x0, state0, f: Callable, T: int, mv0=None, step0: int = 1,
           beta1=0.95, beta2=0.95, grad_of: Optional[Callable] = None) -> UnrollResult:
    """DM/meta.py:338-376 for ONE flat variable.  ``f(x) -> scalar``; the gradient fed to the
    net is stop-gradient'd (meta.py:328-329) but fx_t = f(x_t) stays differentiable w.r.t.
    theta through x_t.  ``grad_of(x) -> (fx, g)`` may supply a closed-form gradient instead of
    autograd (then fx_t must itself be differentiable in x)."""
    x, state, mv = x0, state0, mv0
    fxs, deltas, grads = [], [], []
    for t in range(T):
        if grad_of is not None:
            fx, g = grad_of(x)
            g = g.detach()
        else:
            xg = x.detach().requires_grad_(True)
            with torch.enable_grad():
                fx_d = f(xg)
                (g,) = torch.autograd.grad(fx_d, xg)
            g = g.detach()
            fx = f(x)
        fxs.append(fx)
        gflat = g.reshape(-1)
        if spec.rnnprop:
            m, v = mv
            m, v, mt, gt = adam_features(gflat, m, v, float(step0 + t), beta1, beta2)
            mv = (m, v)
            inp = torch.stack([mt, gt], dim=-1)
        else:
            inp = gflat.unsqueeze(-1)
        delta, state = net_apply(spec, theta, inp, state)
        x = x + delta.reshape(x.shape)
        deltas.append(delta)
        grads.append(gflat)
    fx_final = grad_of(x)[0] if grad_of is not None else f(x)
    fxs.append(fx_final)
    fx = torch.stack([v.reshape(()) for v in fxs])
    return UnrollResult(fx, fx.sum(), x, state, mv, deltas, grads)


def meta_grad(spec: NetSpec, theta, x0, state0, f, T, **kw):
    """dL/dtheta by autograd through ``unroll`` (what tf.gradients emits for DM/meta.py:412)."""
    th = theta.detach().clone().requires_grad_(True)
    res = unroll(spec, th, x0, state0, f, T, **kw)
    (g,) = torch.autograd.grad(res.loss, th)
    return g, res


def imitation_loss(spec: NetSpec, theta, inputs, labels, state0):
    """DM/meta_dm_train.py:463-480: inputs/labels [T, N] (RNNprop: inputs [T, N, 2])."""
    state, total, deltas = state0, 0.0, []
    n = labels.shape[1]
    for t in range(inputs.shape[0]):
        inp = inputs[t] if inputs[t].dim() == 2 else inputs[t].unsqueeze(-1)
        delta, state = net_apply(spec, theta, inp, state)
        total = total + ((labels[t] - delta) ** 2).sum() * 0.5 / n
        deltas.append(delta)
    return total, state, deltas


def tf_adam_step(theta, grad, m, v, k: int, lr=0.01, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (TF 1.14 ApplyAdam): lr_t = lr*sqrt(1-b2^k)/(1-b1^k);
    m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= lr_t*m/(sqrt(v)+eps).  The hyper-parameters are fp32
    scalars in TF, so (1-b) is formed in the tensors' dtype (1 - fp32(0.999) != 0.001)."""
    dt = theta.dtype
    one = torch.ones((), dtype=dt)
    omb1 = one - torch.tensor(b1, dtype=dt)
    omb2 = one - torch.tensor(b2, dtype=dt)
    m = torch.tensor(b1, dtype=dt) * m + omb1 * grad
    v = torch.tensor(b2, dtype=dt) * v + omb2 * grad * grad
    lr_t = lr * math.sqrt(1 - b2 ** k) / (1 - b1 ** k)
    theta = theta - lr_t * m / (torch.sqrt(v) + eps)
    return theta, m, v


class MetaTrainerOracle:
    """``meta_minimize`` + ``run_epoch`` semantics (DM/meta.py:398-414, DM/util.py:31-75) for one
    flat variable: each ``run_unroll`` = one ``sess.run([cost, update, step])``."""

    def __init__(self, spec: NetSpec, theta, f, lr=0.01, grad_of=None, beta1=0.95, beta2=0.95):
        self.spec, self.theta, self.f, self.lr, self.grad_of = spec, theta.clone(), f, lr, grad_of
        self.m = torch.zeros_like(theta)
        self.v = torch.zeros_like(theta)
        self.k = 0
        self.beta1, self.beta2 = beta1, beta2

    def reset(self, x0):
        self.x = x0.clone()
        self.state = initial_state(self.spec, x0.numel(), x0.dtype)
        self.mv = (torch.zeros(x0.numel(), dtype=x0.dtype), torch.zeros(x0.numel(), dtype=x0.dtype)) if self.spec.rnnprop else None
        self.unroll_idx = 0

    def run_unroll(self, T, train=True):
        step0 = self.unroll_idx * T + 1
        if train:
            g, res = meta_grad(self.spec, self.theta, self.x, self.state, self.f, T, mv0=self.mv,
                               step0=step0, beta1=self.beta1, beta2=self.beta2, grad_of=self.grad_of)
            self.k += 1
            self.theta, self.m, self.v = tf_adam_step(self.theta, g, self.m, self.v, self.k, self.lr)
            self.last_grad = g
        else:
            with torch.no_grad():
                res = unroll(self.spec, self.theta, self.x, self.state, self.f, T, mv0=self.mv, step0=step0,
                             beta1=self.beta1, beta2=self.beta2, grad_of=self.grad_of)
        self.x = res.x_final.detach()
        self.state = tuple((h.detach(), c.detach()) for h, c in res.state_final)
        self.mv = tuple(t.detach() for t in res.mv_final) if res.mv_final is not None else None
        self.unroll_idx += 1
        return res


# --------------------------------------------------------------------------- KernelDeepLSTM (DM/networks.py:303-351)
def kernel_net_shapes(kernel_shape, layers, logsign: bool):
    k = int(np.prod(kernel_shape))
    out, kin = [], (2 * k if logsign else k)
    for i, h in enumerate(layers, start=1):
        out += [(f"lstm_{i}", "w_gates", (kin + h, 4 * h)), (f"lstm_{i}", "b_gates", (4 * h,))]
        kin = h
    out += [("linear", "w", (kin, k)), ("linear", "b", (k,))]
    return out


def kernel_net_apply(kernel_shape, layers, theta, inputs, state, preprocess_k=None, scale=1.0, tanh_output=False):
    """KernelDeepLSTM._build: inputs [kw, kh, cin, cout] -> transpose [cin, cout, kw, kh] -> rows [cin*cout, kw*kh]
    (DM/networks.py:325-327) -> StandardDeepLSTM._build (preprocess on the expanded last axis, reshape, DeepRNN, Linear
    with output_size kw*kh, scale; DM/networks.py:207-232) -> transpose back (DM/networks.py:343-346)."""
    shapes = kernel_net_shapes(kernel_shape, layers, preprocess_k is not None)
    w, off = {}, 0
    for mod, var, shp in shapes:
        n = int(np.prod(shp))
        w.setdefault(mod, {})[var] = theta[off:off + n].reshape(shp)
        off += n
    k = int(np.prod(kernel_shape))
    rows = inputs.permute(2, 3, 0, 1).reshape(-1, k)
    if preprocess_k is not None:
        u = log_and_sign(rows.unsqueeze(-1), preprocess_k).reshape(rows.shape[0], -1)
    else:
        u = rows
    out, nxt = u, []
    for li, _ in enumerate(layers, start=1):
        h, c = state[li - 1]
        hn, cn = lstm_cell(out, h, c, w[f"lstm_{li}"]["w_gates"], w[f"lstm_{li}"]["b_gates"])
        nxt.append((hn, cn))
        out = hn
    y = out @ w["linear"]["w"] + w["linear"]["b"]
    y = torch.tanh(y) * scale if tanh_output else y * scale
    return y.t().reshape(inputs.shape), tuple(nxt)
