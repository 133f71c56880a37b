"""Pins the CPU oracle against every known answer the reference's own tests hold for this path
(SURVEY.md 8(c)).  SW/ = /root/reference/Model_Free_L2O/L2O-Swarm/src/ (the DeepMind L2L tests)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import l2o_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _train(tr, x0, num_epochs, num_unrolls, T):
    res = None
    for _ in range(num_epochs):
        tr.reset(x0)
        for _ in range(num_unrolls):
            res = tr.run_unroll(T)
    return float(res.fx[-1]), res.x_final


def test_known_answer_simple_problem():
    """SW/meta_test.py:50-69 testResults: cost 0.7325327, x 0.8559 ("reproducibility of Torch results")."""
    spec = orc.NetSpec(layers=())
    tr = orc.MetaTrainerOracle(spec, orc.init_theta(spec, initializer="zeros"), lambda x: (x * x).sum(), lr=0.01)
    cost, x = _train(tr, torch.ones(1), 1, 2, 5)
    assert abs(cost - 0.7325327) < 1e-4
    assert abs(float(x[0]) - 0.8559) < 1e-4


@pytest.mark.parametrize("init", ["zeros", {"w": np.zeros((20, 1), np.float32), "b": np.zeros((1,), np.float32)}])
def test_zero_linear_gives_zero_update(init):
    """SW/networks_test.py:57-69: zero-initialised output Linear => update == 0."""
    spec = orc.NetSpec(layers=(20, 20))
    initializer = init if isinstance(init, str) else {"linear": init}
    theta = orc.init_theta(spec, seed=3, initializer=initializer)
    g = torch.randn(50)
    d, st = orc.net_apply(spec, theta, g.unsqueeze(-1), orc.initial_state(spec, 50))
    assert float(d.abs().max()) == 0.0
    assert d.shape == g.shape


def test_trainable_variable_count():
    """SW/networks_test.py:47-55: layers=(1,) => 4 variables."""
    assert len(orc.NetSpec(layers=(1,)).shapes()) == 4
    assert orc.NetSpec(layers=(20, 20)).n_theta() == 5061
    assert orc.NetSpec(layers=(20, 20), preprocess_name="LogAndSign", preprocess_options={"k": 5}).n_theta() == 5141
    assert orc.NetSpec(layers=(20, 20), preprocess_name="fc", preprocess_options={"dim": 20}, rnnprop=True).n_theta() == 6641


def test_log_and_sign_known_answers():
    """SW/preprocess_test.py:78-98."""
    out = orc.log_and_sign(torch.ones(10, 1), 1.0)
    assert out.shape == (10, 2)
    assert float(out[:, 0].abs().max()) < 1e-6
    g = torch.randn(64, 1)
    out = orc.log_and_sign(g, 5.0)
    assert torch.equal(torch.sign(out[:, 1:]), torch.sign(g))
    assert float(out[:, 0].min()) >= -1.0 and float(out[:, 1].abs().max()) <= 1.0
    # clamp: tiny gradient -> log part clamps at -1, sign part stays linear (DM/preprocess.py:66-67)
    tiny = orc.log_and_sign(torch.tensor([[1e-20]]), 5.0)
    assert float(tiny[0, 0]) == -1.0 and abs(float(tiny[0, 1]) - 1e-20 * math.exp(5)) < 1e-24


def test_quadratic_value():
    """SW/problems_test.py:99-111: batch 1, dim 1 => ((w x) - y)^2."""
    x, w, y = torch.tensor([[0.7]]), torch.tensor([[[1.3]]]), torch.tensor([[0.2]])
    assert abs(float(orc.quadratic_f(x, w, y)) - (1.3 * 0.7 - 0.2) ** 2) < 1e-7


def test_save_load_roundtrip_initializer():
    """SW/meta_test.py:190-236: a saved {module:{var:ndarray}} dict re-initialises to the same theta."""
    spec = orc.NetSpec(layers=(2, 3))
    theta = orc.init_theta(spec, seed=5)
    d = {m: {v: t.numpy().copy() for v, t in vs.items()} for m, vs in orc.unpack_theta(spec, theta).items()}
    assert torch.equal(orc.init_theta(spec, seed=99, initializer=d), theta)


def test_single_cell_hand_vectors():
    """Hand-computed snt.LSTM cases (the gate arithmetic is pinned by none of the reference tests;
    these encode the Sonnet-1.11 semantics: order i|j|f|o, forget bias +1, state (h, c))."""
    with open(os.path.join(GOLD, "lstm_cell_hand.json")) as f:
        cases = json.load(f)
    for c in cases:
        tt = lambda v: torch.tensor(v, dtype=torch.float32)
        h, cc = orc.lstm_cell(tt(c["x"]), tt(c["h"]), tt(c["c"]), tt(c["w"]), tt(c["b"]))
        assert torch.allclose(h, torch.tensor(c["h_next"]), atol=1e-6), c["name"]
        assert torch.allclose(cc, torch.tensor(c["c_next"]), atol=1e-6), c["name"]


def test_tf_adam_first_step_is_lr_sign():
    th, m, v = orc.tf_adam_step(torch.zeros(3), torch.tensor([2.0, -0.5, 1e-3]), torch.zeros(3), torch.zeros(3), 1, lr=0.01)
    assert torch.allclose(th, torch.tensor([-0.01, 0.01, -0.01]), atol=1e-5)


def test_lambda_suffix_sum_identity():
    """with stop-gradient'd g, dL/dtheta = sum_t (sum_{tau>t} g_tau) . dDelta_t/dtheta."""
    spec = orc.NetSpec(layers=(2, 3))
    theta = orc.init_theta(spec, seed=1)
    gen = torch.Generator().manual_seed(0)
    B, d, T = 3, 4, 5
    w, y, x0 = torch.rand(B, d, d, generator=gen), torch.rand(B, d, generator=gen), torch.randn(B, d, generator=gen) * 0.1
    f = lambda x: orc.quadratic_f(x, w, y)
    g, res = orc.meta_grad(spec, theta, x0, orc.initial_state(spec, B * d), f, T)
    xT = res.x_final.detach().requires_grad_(True)
    (gT,) = torch.autograd.grad(f(xT), xT)
    grads = res.grads + [gT.reshape(-1)]
    th = theta.clone().requires_grad_(True)
    s, total = orc.initial_state(spec, B * d), 0.0
    for t in range(T):
        delta, s = orc.net_apply(spec, th, grads[t].unsqueeze(-1), s)
        lam = torch.stack(grads[t + 1:]).sum(0)
        total = total + (lam * delta).sum()
    (g2,) = torch.autograd.grad(total, th)
    assert torch.allclose(g, g2, rtol=1e-4, atol=1e-7)


def test_lstm_cell_agrees_with_an_independent_lstm_after_gate_relabelling():
    """The oracle's cell against torch.nn.LSTMCell — an independent LSTM implementation whose documented gate order is
    (i, f, g, o) with no built-in forget bias.  Relabelling the Sonnet-order columns (i, j, f, o) and moving the +1.0
    forget bias into the bias vector must reproduce (h', c') exactly; this pins the arithmetic of the restatement
    (what remains unpinned is only Sonnet's own column order / forget-bias convention, see the module header)."""
    from oracle import l2o_oracle as orc
    torch.manual_seed(0)
    n, f, h = 7, 3, 5
    x = torch.randn(n, f, dtype=torch.float64)
    h0, c0 = torch.randn(n, h, dtype=torch.float64), torch.randn(n, h, dtype=torch.float64)
    w = torch.randn(f + h, 4 * h, dtype=torch.float64) * 0.4       # Sonnet layout: rows [x | h], columns i|j|f|o
    b = torch.randn(4 * h, dtype=torch.float64) * 0.1
    h1, c1 = orc.lstm_cell(x, h0, c0, w, b)
    cell = torch.nn.LSTMCell(f, h, dtype=torch.float64)
    i_, j_, f_, o_ = (slice(k * h, (k + 1) * h) for k in range(4))
    order = [i_, f_, j_, o_]                                         # torch rows: i, f, g(=j), o
    with torch.no_grad():
        cell.weight_ih.copy_(torch.cat([w[:f, s].t() for s in order], 0))
        cell.weight_hh.copy_(torch.cat([w[f:, s].t() for s in order], 0))
        bias = torch.cat([b[s] for s in order], 0)
        bias[h:2 * h] += 1.0                                         # forget_bias=1.0, added at run time by snt.LSTM
        cell.bias_ih.copy_(bias)
        cell.bias_hh.zero_()
        h_t, c_t = cell(x, (h0, c0))
    assert torch.allclose(h1, h_t, atol=1e-12) and torch.allclose(c1, c_t, atol=1e-12)


def test_golden_lstm_vectors_match_their_generator():
    """tests/golden/lstm_cell_hand.json is reproducible from its scalar-math generator (no torch, no oracle)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_lstm_cell_hand", os.path.join(GOLD, "make_lstm_cell_hand.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(GOLD, "lstm_cell_hand.json")) as f:
        committed = json.load(f)
    for got, want in zip(mod.regenerate(), committed):
        for key in ("h_next", "c_next"):
            for ra, rb in zip(got[key], want[key]):
                assert all(abs(a - b) <= 1e-12 for a, b in zip(ra, rb)), (want["name"], key)
