"""GPU tests of the reference-facing surface (MetaOptimizer / networks / preprocess).  They are written to
read like the reference's own tests (SW/meta_test.py, SW/networks_test.py, SW/preprocess_test.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import l2o_oracle as orc
from tests.helpers import REL_TOL, assert_theta_close, rel_err

pytestmark = pytest.mark.gpu


def train(sess, minimize_ops, num_epochs, num_unrolls):
    """L2L training (SW/meta_test.py:33-43)."""
    step, update, reset, loss_last, x_last = minimize_ops
    for _ in range(num_epochs):
        sess.run(reset)
        for _ in range(num_unrolls):
            cost, final_x, unused_1, unused_2 = sess.run([loss_last, x_last, update, step])
    return cost, final_x


def test_results_known_answer():
    """SW/meta_test.py:50-69 testResults ("reproducibility of Torch results")."""
    from open_l2o_b200 import meta, problems
    problem = problems.simple()
    optimizer = meta.MetaOptimizer(net=dict(net="CoordinateWiseDeepLSTM",
                                            net_options={"layers": (), "initializer": "zeros"}))
    minimize_ops = optimizer.meta_minimize(problem, 5)
    with meta.Session() as sess:
        cost, final_x = train(sess, minimize_ops, 1, 2)
    assert abs(cost - 0.7325327) < 1e-4
    assert abs(float(final_x[0]) - 0.8559) < 1e-4


@pytest.mark.parametrize("net_assignments,net_config", [
    (None, {"net": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1, 1,)}}}),
    ([("net", ["x_0", "x_1"])], {"net": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1,)}}}),
    ([("net1", ["x_0"]), ("net2", ["x_0"])],
     {"net1": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1,)}},
      "net2": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1,)}}}),
])
def test_multi_optimizer(net_assignments, net_config):
    """SW/meta_test.py:71-126 (the Adam-net variant is outside the accelerated path)."""
    from open_l2o_b200 import meta, problems
    problem = problems.simple_multi_optimizer(num_dims=2)
    optimizer = meta.MetaOptimizer(**net_config)
    minimize_ops = optimizer.meta_minimize(problem, 3, net_assignments=net_assignments)
    with meta.Session() as sess:
        cost, x = train(sess, minimize_ops, 1, 2)
    assert np.isfinite(cost)


def test_net_assignment_errors():
    """DM/meta.py:188-190, 206-207."""
    from open_l2o_b200 import meta, problems
    cfg = {"a": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1,)}},
           "b": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (1,)}}}
    with pytest.raises(ValueError):
        meta.MetaOptimizer(**cfg).meta_loss(problems.simple_multi_optimizer(), 3)
    with pytest.raises(ValueError):
        meta.MetaOptimizer(**cfg).meta_loss(problems.simple_multi_optimizer(), 3,
                                            net_assignments=[("a", ["x_0"]), ("a", ["x_1"])])
    with pytest.raises(NotImplementedError):
        meta.MetaOptimizer(**cfg).meta_loss(problems.simple(), 3, second_derivatives=True)


def test_save_and_load():
    """SW/meta_test.py:190-236."""
    from open_l2o_b200 import meta, problems
    net_options = {"layers": (2, 3), "initializer": "zeros"}
    problem = problems.simple()
    optimizer = meta.MetaOptimizer(net=dict(net="CoordinateWiseDeepLSTM", net_options=net_options))
    minimize_ops = optimizer.meta_minimize(problem, 3)
    sess = meta.Session()
    train(sess, minimize_ops, 1, 2)
    tmp_dir = tempfile.mkdtemp()
    save_result = optimizer.save(sess, path=tmp_dir)
    net_path = next(iter(save_result))
    cost, x = train(sess, minimize_ops, 2, 1)

    optimizer2 = meta.MetaOptimizer(net=dict(net="CoordinateWiseDeepLSTM", net_options=net_options,
                                             net_path=net_path))
    minimize_ops2 = optimizer2.meta_minimize(problem, 3)
    cost_loaded, x_loaded = train(meta.Session(), minimize_ops2, 2, 1)
    assert abs(cost - cost_loaded) < 1e-3
    assert abs(float(x[0]) - float(x_loaded[0])) < 1e-3
    os.remove(net_path)
    os.rmdir(tmp_dir)


def test_coordinatewise_net_operator_surface():
    """SW/networks_test.py:29-69: shape, variable count, zero-initialised Linear => zero update."""
    from open_l2o_b200 import networks
    shape = [10, 5]
    gradients = torch.randn(shape, device="cuda")
    net = networks.CoordinateWiseDeepLSTM(layers=(1, 1))
    state = net.initial_state_for_inputs(gradients)
    update, next_state = net(gradients, state)
    assert list(update.shape) == shape
    assert len(networks.CoordinateWiseDeepLSTM(layers=(1,)).variable_shapes()) == 4
    for init in ["zeros", {"w": "zeros", "b": "zeros"}, {"linear": {"w": "zeros", "b": "zeros"}}]:
        net = networks.CoordinateWiseDeepLSTM(layers=(1, 1), initializer=init)
        update, _ = net(gradients, net.initial_state_for_inputs(gradients))
        assert float(update.abs().max()) == 0.0


def test_net_call_matches_oracle_and_chains_state():
    from open_l2o_b200 import networks
    spec = orc.NetSpec(layers=(20, 20), preprocess_name="LogAndSign", preprocess_options={"k": 5}, scale=0.01)
    net = networks.factory("CoordinateWiseDeepLSTM", net_options=dict(layers=(20, 20), preprocess_name="LogAndSign",
                                                                      preprocess_options={"k": 5}, scale=0.01))
    theta = net.theta.cpu()
    g = torch.randn(7, 33)
    s_ref = orc.initial_state(spec, g.numel())
    s = net.initial_state_for_inputs(g.cuda())
    for _ in range(3):
        d_ref, s_ref = orc.net_apply(spec, theta, g.reshape(-1, 1), s_ref)
        d, s = net(g.cuda(), s)
        assert rel_err(d.reshape(-1), d_ref) <= REL_TOL
    for (h, c), (hr, cr) in zip(s, s_ref):
        assert tuple(h.shape) == (g.numel(), 20)
        assert rel_err(h, hr) <= REL_TOL and rel_err(c, cr) <= REL_TOL


def test_rnnprop_operator_surface():
    from open_l2o_b200 import networks
    spec = orc.NetSpec(layers=(20, 20), preprocess_name="fc", preprocess_options={"dim": 20}, scale=0.01,
                       tanh_output=True, rnnprop=True)
    net = networks.factory("RNNprop", net_options=dict(layers=(20, 20), preprocess_name="fc",
                                                       preprocess_options={"dim": 20}, scale=0.01, tanh_output=True))
    m, g = torch.randn(50, 3), torch.randn(50, 3)
    d, s = net(m.cuda(), g.cuda(), net.initial_state_for_inputs(g.cuda()))
    d_ref, _ = orc.net_apply(spec, net.theta.cpu(), torch.stack([m.reshape(-1), g.reshape(-1)], -1),
                             orc.initial_state(spec, g.numel()))
    assert rel_err(d.reshape(-1), d_ref) <= REL_TOL


def test_log_and_sign_module():
    """SW/preprocess_test.py:71-98."""
    from open_l2o_b200 import preprocess
    g = torch.randn(10, 4, device="cuda")
    out = preprocess.LogAndSign(None, k=1)(g)
    assert list(out.shape) == [10, 8]
    out1 = preprocess.LogAndSign(None, k=1.0)(torch.ones(3, 1, device="cuda"))
    assert float(out1[:, 0].abs().max()) < 1e-6
    assert torch.equal(torch.sign(out[:, 4:]), torch.sign(g))
    assert float(preprocess.Clamp(min_value=-1, max_value=1)(g * 10).abs().max()) <= 1.0


def _oracle_trainer_for(prog, spec, lr, f=None, grad_of=None, **kw):
    net = next(iter(prog.nets.values()))
    tr = orc.MetaTrainerOracle(spec, net.theta.cpu().clone(), f, lr=lr, grad_of=grad_of, **kw)
    return tr


@pytest.mark.parametrize("fused", [True, False])
def test_training_trajectory_rastrigin_matches_oracle(fused, monkeypatch):
    """BASELINE config #5 family at small d: 3 x (T-step unroll + BPTT + TF-Adam + carry-over), fused and
    step-at-a-time execution, against the oracle's autograd trainer on identical tensors."""
    from open_l2o_b200 import meta, problems
    if not fused:
        monkeypatch.setenv("L2O_DISABLE_FUSED", "1")
    n, T = 3000, 10
    problem = problems.rastrigin_separable(num_dims=n)
    optimizer = meta.MetaOptimizer(cw={"net": "CoordinateWiseDeepLSTM",
                                       "net_options": {"layers": (20, 20), "scale": 0.1}})
    step, update, reset, fx, x = optimizer.meta_minimize(problem, T, learning_rate=0.001)
    prog = optimizer.program
    assert (prog.fused is not None) == fused
    sess = meta.Session()
    sess.run(reset)
    spec = orc.NetSpec(layers=(20, 20), scale=0.1)
    b, c = prog.const_vals["b"].cpu(), prog.const_vals["c"].cpu()
    prob = orc.FusedProblem("rastrigin_sep", b, c, alpha=10.0, fscale=1.0 / n)
    tr = _oracle_trainer_for(prog, spec, 0.001, grad_of=prob.f_and_g)
    tr.reset(prog.X.cpu().clone())
    for it in range(3):
        cost, xs, _, _ = sess.run([fx, x, update, step])
        res = tr.run_unroll(T)
        assert abs(cost - float(res.fx[-1])) <= 1e-5 * abs(float(res.fx[-1]))
        assert rel_err(xs[0], res.x_final) <= REL_TOL
        assert_theta_close(next(iter(prog.nets.values())).theta, tr, it)


def test_training_trajectory_quadratic_matches_oracle():
    """BASELINE config #1 (L2O-DM quadratic 128x10, LSTM-20x2, unroll 20): external-gradient regime."""
    from open_l2o_b200 import meta, util
    problem, net_config, net_assignments = util.get_config("quadratic")
    optimizer = meta.MetaOptimizer(**net_config)
    ms = optimizer.meta_minimize(problem, 20, learning_rate=0.001, net_assignments=net_assignments)
    prog = optimizer.program
    sess = meta.Session()
    sess.run(ms.reset)
    w, y = prog.const_vals["w"].cpu(), prog.const_vals["y"].cpu()
    spec = orc.NetSpec(layers=(20, 20))
    tr = _oracle_trainer_for(prog, spec, 0.001, f=lambda x: orc.quadratic_f(x, w, y))
    tr.reset(prog.X.cpu().clone().reshape(128, 10))
    for it in range(3):
        cost, xs, _, _ = sess.run([ms.fx, ms.x, ms.update, ms.step])
        res = tr.run_unroll(20)
        assert abs(cost - float(res.fx[-1])) <= 1e-5 * abs(float(res.fx[-1])) + 1e-9
        assert rel_err(xs[0], res.x_final) <= REL_TOL
        assert_theta_close(next(iter(prog.nets.values())).theta, tr, it)


def test_rnnprop_training_matches_oracle():
    """BASELINE config #3 family (RNNProp, MLP optimizee, synthetic batch) at reduced width."""
    from open_l2o_b200 import meta, problems
    T = 8
    problem = problems.mlp(layers=(12,), in_dim=20, n_classes=5, batch_size=16)
    optimizer = meta.RNNpropMetaOptimizer(rp={"net": "RNNprop", "net_options": {
        "layers": (20, 20), "preprocess_name": "fc", "preprocess_options": {"dim": 20}, "scale": 0.01,
        "tanh_output": True}})
    ms = optimizer.meta_minimize(problem, T, learning_rate=0.001)
    prog = optimizer.program
    sess = meta.Session()
    sess.run(ms.reset)
    data, labels = prog.const_vals["data"].cpu(), prog.const_vals["labels"].cpu()
    shapes = [v["shape"] for v in prog.variables]

    def f(xflat):
        off, ts = 0, []
        for s in shapes:
            k = int(np.prod(s))
            ts.append(xflat[off:off + k].view(s))
            off += k
        h = torch.sigmoid(data.to(xflat.dtype) @ ts[0] + ts[1])
        return torch.nn.functional.cross_entropy(h @ ts[2] + ts[3], labels.long())

    spec = orc.NetSpec(layers=(20, 20), preprocess_name="fc", preprocess_options={"dim": 20}, scale=0.01,
                       tanh_output=True, rnnprop=True)
    tr = _oracle_trainer_for(prog, spec, 0.001, f=f)
    tr.reset(prog.X.cpu().clone())
    for it in range(3):
        cost, xs, _, _ = sess.run([ms.fx, ms.x, ms.update, ms.step],
                                  feed_dict={optimizer.step_placeholder: it * T + 1})
        res = tr.run_unroll(T)
        assert abs(cost - float(res.fx[-1])) <= 1e-5 * abs(float(res.fx[-1]))
        assert rel_err(np.concatenate([a.reshape(-1) for a in xs]), res.x_final) <= REL_TOL
        assert_theta_close(next(iter(prog.nets.values())).theta, tr, it)


# ---------------------------------------------------------------------------------------------------------------
# Training-time surface (DM/meta_dm_train.py): scale placeholders + imitation ("mt") tasks + data generator
# ---------------------------------------------------------------------------------------------------------------
def _dm_train_setup(num_mt=1, T=6, lr=0.001):
    from open_l2o_b200 import meta_dm_train, problems
    problem = problems.quadratic(batch_size=16, num_dims=10)
    optimizer = meta_dm_train.MetaOptimizer(num_mt, cw={"net": "CoordinateWiseDeepLSTM",
                                                        "net_options": {"layers": (20, 20), "scale": 0.1}})
    out = optimizer.meta_minimize(problem, T, learning_rate=lr)
    return optimizer, out


def test_imitation_task_matches_oracle():
    """DM/meta_dm_train.py:463-480 + :549-553: two imitation unrolls (loss, state carry-over, per-task Adam)."""
    from open_l2o_b200 import meta
    T, lr = 6, 0.001
    optimizer, (ms, scale, var_x, constants, subsets, loss_mt, steps_mt, update_mt, reset_mt, mt_labels, mt_inputs) = \
        _dm_train_setup(1, T, lr)
    prog = optimizer.program
    n = prog.N
    sess = meta.Session()
    sess.run(ms.reset)
    sess.run(reset_mt[0])
    spec = orc.NetSpec(layers=(20, 20), scale=0.1)
    theta = next(iter(prog.nets.values())).theta.cpu().clone()
    m, v = torch.zeros_like(theta), torch.zeros_like(theta)
    state = orc.initial_state(spec, n)
    gen = torch.Generator().manual_seed(4)
    for it in range(2):
        inputs = torch.randn(T, n, generator=gen) * 0.3
        labels = torch.randn(T, n, generator=gen) * 0.01
        th = theta.clone().requires_grad_(True)
        loss_ref, state_next, _ = orc.imitation_loss(spec, th, inputs, labels, state)
        (g,) = torch.autograd.grad(loss_ref, th)
        theta, m, v = orc.tf_adam_step(theta, g, m, v, it + 1, lr=lr)
        state = tuple((h.detach(), c.detach()) for h, c in state_next)
        cost = sess.run([loss_mt[0], update_mt[0], steps_mt[0]],
                        feed_dict={mt_inputs[0][0]: inputs.numpy(), mt_labels[0][0]: labels.numpy()})[0]
        assert abs(cost - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
        big = g.abs() > 1e-5 * float(g.abs().max())
        assert rel_err(next(iter(prog.nets.values())).theta.cpu()[big], theta[big]) <= REL_TOL, it


def test_scale_placeholders_random_scaling_trick():
    """DM/meta_dm_train.py:336-338,384-385 + DM/util.py:40-54: fx = f(x*scale), g = scale * grad f(x*scale)."""
    from open_l2o_b200 import meta
    T = 5
    optimizer, (ms, scale, var_x, *_rest) = _dm_train_setup(0, T)
    prog = optimizer.program
    sess = meta.Session()
    sess.run(ms.reset)
    w, y = prog.const_vals["w"].cpu(), prog.const_vals["y"].cpu()
    x0 = prog.X.cpu().clone().reshape(16, 10)
    r = torch.exp(torch.rand(16, 10) * 6 - 3)
    spec = orc.NetSpec(layers=(20, 20), scale=0.1)
    theta = next(iter(prog.nets.values())).theta.cpu().clone()
    with torch.no_grad():
        pass
    res = orc.unroll(spec, theta, x0, orc.initial_state(spec, 160), lambda x: orc.quadratic_f(x * r, w, y), T)
    cost, xs = sess.run([ms.fx, ms.x], feed_dict={scale[0]: r.numpy()})
    assert abs(cost - float(res.fx[-1])) <= 1e-5 * abs(float(res.fx[-1]))
    assert rel_err(xs[0], res.x_final.detach()) <= REL_TOL
    assert var_x[0].shape == (16, 10) and var_x[0].value().shape == (16, 10)


def test_data_loader_and_run_epoch_branches():
    """DM/data_generator.py:71-124 shapes + DM/util.py:31-75 branches (random scaling, imitation task)."""
    from open_l2o_b200 import meta, util
    from open_l2o_b200.data_generator import data_loader
    T = 4
    optimizer, (ms, scale, var_x, constants, subsets, loss_mt, steps_mt, update_mt, reset_mt, mt_labels, mt_inputs) = \
        _dm_train_setup(1, T)
    prog = optimizer.program
    sess = meta.Session()
    dl = data_loader(prog, "adam,rmsprop,nag", T)
    for task in range(3):
        data = dl.get_data(task, sess, num_unrolls=2, assign_func=optimizer.assign_func, rd_scale_bound=3.0)
        assert len(data["inputs"]) == 2 and data["inputs"][0][0].shape == (T, prog.N)
        assert np.isfinite(data["labels"][1][0]).all() and np.abs(data["labels"][0][0]).max() > 0
    # Adam's first move is lr * sign(g)
    d0 = dl.get_data(0, sess, num_unrolls=1, if_scale=False)
    assert np.allclose(np.abs(d0["labels"][0][0][0]), 0.01, atol=1e-4)
    t, cost = util.run_epoch(sess, loss_mt[0], [update_mt[0], steps_mt[0]], reset_mt[0], 2, task_i=0, data=data,
                             label_pl=mt_labels[0], input_pl=mt_inputs[0])
    assert np.isfinite(cost)
    t, cost = util.run_epoch(sess, ms.fx, [ms.update, ms.step], ms.reset, 2, scale=scale, rd_scale=True,
                             assign_func=optimizer.assign_func, var_x=var_x)
    assert np.isfinite(cost)


def test_training_quadratic_fused_dense_groups():
    """problems.quadratic at a batch large enough for the fused regime: the dense per-group gradient is evaluated
    in-kernel (L2O_OPT_QUADRATIC_BATCH); two training unrolls against the oracle's autograd trainer."""
    from open_l2o_b200 import meta, problems
    B, d, T = 2048, 10, 6
    problem = problems.quadratic(batch_size=B, num_dims=d)
    optimizer = meta.MetaOptimizer(cw={"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (20, 20)}})
    ms = optimizer.meta_minimize(problem, T, learning_rate=0.001)
    prog = optimizer.program
    assert prog.fused is not None and prog.fused.kind == "quadratic_batch"
    sess = meta.Session()
    sess.run(ms.reset)
    w, y = prog.const_vals["w"].cpu(), prog.const_vals["y"].cpu()
    spec = orc.NetSpec(layers=(20, 20))
    tr = _oracle_trainer_for(prog, spec, 0.001, f=lambda x: orc.quadratic_f(x, w, y))
    tr.reset(prog.X.cpu().clone().reshape(B, d))
    for it in range(2):
        cost, xs, _, _ = sess.run([ms.fx, ms.x, ms.update, ms.step])
        res = tr.run_unroll(T)
        assert abs(cost - float(res.fx[-1])) <= 1e-5 * abs(float(res.fx[-1])) + 1e-9
        assert rel_err(xs[0], res.x_final) <= REL_TOL
        assert_theta_close(next(iter(prog.nets.values())).theta, tr, it)
