"""CPU checks of the HierarchicalRNN oracle (oracle/hrnn_oracle.py) against closed forms read off the reference
code (no reference test exists for L2O-Scale: parity is unpinned, see the oracle's header)."""
import math

import torch

from oracle import hrnn_oracle as H


def test_theta_layout_and_counts():
    assert H.theta_count() == 8349
    P = H.unpack_theta(H.init_theta(0))
    assert P["PerTensor/Layer0_RNN/BiasGRUCell/gates/Affine/Matrix"].shape == (22, 20)
    assert P["PerTensor/Layer1_RNN/BiasGRUCell/gates/Affine/Matrix"].shape == (42, 40)
    assert float(P["PerTensor/Layer0_RNN/BiasGRUCell/gates/Affine/Bias"][0]) == 2.2 or \
        abs(float(P["PerTensor/Layer0_RNN/BiasGRUCell/gates/Affine/Bias"][0]) - 2.2) < 1e-6
    from open_l2o_b200.hierarchical_rnn import THETA_SPEC
    assert [(n, tuple(s)) for n, s in THETA_SPEC] == [(n, tuple(s)) for n, s in H.theta_spec()]


def test_first_step_closed_form():
    """At t=0 the state is zero: decays are 0, so acc = g, ms = g^2 + 1e-12 (utils.py:128-134) and every scaled
    gradient is asinh(g / sqrt(g^2 + 1e-12 + 1e-16)); the 4 log-ms features are centred to exactly 0; the new
    log-lr is lrm*l + (1-lrm)*clip(l + b_l) with zero lr weights (HR:663-689)."""
    th = H.init_theta(1, dtype=torch.float64)
    P = H.unpack_theta(th)
    gen = torch.Generator().manual_seed(2)
    params = [torch.randn(6, 4, generator=gen, dtype=torch.float64), torch.randn(3, generator=gen, dtype=torch.float64)]
    states = [H.initial_state(P, p, gen) for p in params]
    grads = [torch.randn(6, 4, generator=gen, dtype=torch.float64), torch.randn(3, generator=gen, dtype=torch.float64)]
    newp, news, newg, upd = H.step(th, params, grads, states, H.initial_global_state(P, torch.float64))
    for g, st, ns in zip(grads, states, news):
        gv = g.reshape(-1, 1)
        assert torch.allclose(ns["grad_accum1"], gv) and torch.allclose(ns["grad_accum4"], gv)
        assert torch.allclose(ns["ms1"], gv * gv + 1e-12) and torch.allclose(ns["ms3"], gv * gv + 1e-12)
        lrm = 1.0 / (1.0 + math.exp(-3.2))
        assert torch.allclose(ns["log_learning_rate"], st["log_learning_rate"])   # lr weights / bias start at zero
        assert 0.0 < lrm < 1.0
        assert ns["layer"].shape == (1, 20) and ns["parameter"].shape == (g.numel(), 10)
    # the applied step is lr * delta / RMS(delta): per tensor, mean((upd / lr)^2) == 1
    for g, st, u in zip(grads, states, upd):
        lr = torch.exp(st["log_learning_rate"].reshape(-1) - 1.0)
        r = (u.reshape(-1) / lr)
        assert abs(float((r * r).mean()) - 1.0) < 1e-9


def test_global_state_sees_last_tensor_only():
    """HR:426-427 passes `[layer_state]` — the loop's final value — to the global RNN; permuting the earlier tensors
    must not change the new global state, changing the last one must."""
    th = H.init_theta(3, dtype=torch.float64)
    P = H.unpack_theta(th)
    gen = torch.Generator().manual_seed(4)
    ps = [torch.randn(5, generator=gen, dtype=torch.float64) for _ in range(3)]
    gs = [torch.randn(5, generator=gen, dtype=torch.float64) for _ in range(3)]
    sts = [H.initial_state(P, p, gen) for p in ps]
    g0 = H.initial_global_state(P, torch.float64)
    a = H.step(th, ps, gs, sts, g0)[2]
    b = H.step(th, [ps[1], ps[0], ps[2]], [gs[1], gs[0], gs[2]], [sts[1], sts[0], sts[2]], g0)[2]
    c = H.step(th, [ps[0], ps[2], ps[1]], [gs[0], gs[2], gs[1]], [sts[0], sts[2], sts[1]], g0)[2]
    assert torch.allclose(a, b) and not torch.allclose(a, c)


def test_constructor_argument_checks_need_no_gpu():
    """The reference's argument errors (HR:132-144) and this build's flag-set check fire before any device work."""
    import pytest
    from open_l2o_b200 import hierarchical_rnn as hr
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20, 5])
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20.0, 20])
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20], init_lr_range=(1e-6,))
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20], init_lr_range=(1e-2, 1e-6))
    with pytest.raises(NotImplementedError):
        hr.HierarchicalRNN(**dict(hr.metarun_flags(), use_attention=True))
    flags = hr.metarun_flags()
    assert flags["level_sizes"] == [10, 20, 20] and flags["use_problem_lr_mean"] and flags["num_gradient_scales"] == 4
