"""KernelDeepLSTM (DM/networks.py:303-351) on the run-time-shaped dense engine (l2o_dense_*): operator surface like
SW/networks_test.py:72-112, step parity vs the oracle, and a training trajectory through MetaOptimizer with
net_assignments like SW/meta_test.py:140-167 (testConvolutional)."""
import numpy as np
import pytest
import torch

from oracle import l2o_oracle as orc
from tests.helpers import REL_TOL, assert_theta_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_kernel_net_shape_variables_and_zero_init():
    """SW/networks_test.py:75-112: output shape == input shape; (1,)-layer net has 4 variables; zero-initialised
    output Linear => zero update."""
    from open_l2o_b200 import networks
    kernel_shape = [5, 5]
    shape = kernel_shape + [2, 2]
    gradients = torch.randn(shape, device=DEV)
    net = networks.KernelDeepLSTM(layers=(1, 1), kernel_shape=kernel_shape)
    update, _ = net(gradients, net.initial_state_for_inputs(gradients))
    assert list(update.shape) == shape
    assert len(networks.KernelDeepLSTM(layers=(1,), kernel_shape=kernel_shape).variable_shapes()) == 4
    for init in ["zeros", {"w": "zeros", "b": "zeros", "bad": "bad"}, {"linear": {"w": "zeros", "b": "zeros"}}]:
        net = networks.KernelDeepLSTM(layers=(1, 1), kernel_shape=kernel_shape, initializer=init)
        update, _ = net(gradients, net.initial_state_for_inputs(gradients))
        assert float(update.abs().max()) == 0.0
    with pytest.raises(ValueError):
        net(torch.randn(3, 3, 2, 2, device=DEV), net.initial_state_for_inputs(gradients))


@pytest.mark.parametrize("kernel_shape,layers,cin,cout,logsign", [([3, 3], (20, 20), 5, 7, False),
                                                                    ([5, 5], (5,), 3, 40, True),
                                                                    ([4, 4], (1, 1), 1, 1, False),
                                                                    ([2, 3], (), 6, 6, False)])
def test_kernel_net_steps_match_oracle(kernel_shape, layers, cin, cout, logsign):
    from open_l2o_b200 import networks
    opts = dict(layers=layers, kernel_shape=kernel_shape, scale=0.1, seed=3)
    if logsign:
        opts.update(preprocess_name="LogAndSign", preprocess_options={"k": 5})
    net = networks.factory("KernelDeepLSTM", net_options=opts)
    theta = net.theta.cpu()
    gen = torch.Generator().manual_seed(1)
    shape = kernel_shape + [cin, cout]
    state = net.initial_state_for_inputs(torch.zeros(shape, device=DEV))
    s_ref = tuple((torch.zeros(cin * cout, h), torch.zeros(cin * cout, h)) for h in layers)
    for it in range(3):
        g = torch.randn(shape, generator=gen) * (0.3 if it != 1 else 1e-4)
        d_ref, s_ref = orc.kernel_net_apply(kernel_shape, layers, theta, g, s_ref, preprocess_k=5 if logsign else None,
                                            scale=0.1)
        d, state = net(g.to(DEV), state)
        assert list(d.shape) == shape
        assert rel_err(d, d_ref) <= REL_TOL, it
        for (h, c), (hr, cr) in zip(state, s_ref):
            assert rel_err(h, hr) <= REL_TOL and rel_err(c, cr) <= REL_TOL


def test_convolutional_problem_training_matches_oracle():
    """SW/meta_test.py:140-167 testConvolutional with the reduction made differentiable by hand: a conv filter bank
    trained by a KernelDeepLSTM and its bias by a coordinate-wise net (two nets, net_assignments); three unrolls of
    forward + BPTT + TF-Adam against the oracle's autograd."""
    from open_l2o_b200 import meta
    from open_l2o_b200.variables import get_variable, random_normal_initializer
    kh = kw = 3
    cin, cout, T = 2, 4, 5
    gen = torch.Generator().manual_seed(11)
    data = torch.randn(6, cin, 8, 8, generator=gen)
    target = torch.randn(6, cout, 8, 8, generator=gen)
    data_d, target_d = data.to(DEV), target.to(DEV)

    def conv_loss(w, b, dat, tgt):    # w [kw, kh, cin, cout] (the reference's HWIO filter layout), b [cout]
        out = torch.nn.functional.conv2d(dat, w.permute(3, 2, 0, 1), b, padding=1)
        return ((out - tgt) ** 2).mean()

    def problem():
        w = get_variable("conv/w", shape=[kw, kh, cin, cout], initializer=random_normal_initializer(stddev=0.3))
        b = get_variable("conv/b", shape=[cout], initializer=random_normal_initializer(stddev=0.3))
        return conv_loss(w, b, data_d, target_d)

    net_config = {"conv": {"net": "KernelDeepLSTM", "net_options": {"kernel_shape": [kw, kh], "layers": (5,), "scale": 0.1}},
                  "cw": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (20, 20), "scale": 0.1}}}
    optimizer = meta.MetaOptimizer(**net_config)
    ms = optimizer.meta_minimize(problem, T, learning_rate=0.001,
                                 net_assignments=[("conv", ["conv/w"]), ("cw", ["conv/b"])])
    prog = optimizer.program
    sess = meta.Session()
    sess.run(ms.reset)
    # oracle: both nets, two Adam states, autograd through the unroll
    th = {k: prog.nets[k].theta.cpu().clone() for k in ("conv", "cw")}
    adam = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in th.items()}
    nw = kw * kh * cin * cout
    x = prog.X.cpu().clone()
    spec_cw = orc.NetSpec(layers=(20, 20), scale=0.1)
    s_conv = ((torch.zeros(cin * cout, 5), torch.zeros(cin * cout, 5)),)
    s_cw = orc.initial_state(spec_cw, cout)
    for it in range(3):
        cost, xs, _, _ = sess.run([ms.fx, ms.x, ms.update, ms.step])
        p = {k: v.clone().requires_grad_(True) for k, v in th.items()}
        xc, sc, sw_, total = x, s_conv, s_cw, 0.0
        for t in range(T):
            xg = xc.detach().requires_grad_(True)
            fx_d = conv_loss(xg[:nw].view(kw, kh, cin, cout), xg[nw:], data, target)
            (g,) = torch.autograd.grad(fx_d, xg)
            total = total + conv_loss(xc[:nw].view(kw, kh, cin, cout), xc[nw:], data, target)
            dw, sc = orc.kernel_net_apply([kw, kh], (5,), p["conv"], g[:nw].view(kw, kh, cin, cout).detach(), sc, scale=0.1)
            db, sw_ = orc.net_apply(spec_cw, p["cw"], g[nw:].detach().unsqueeze(-1), sw_)
            xc = xc + torch.cat([dw.reshape(-1), db])
        fx_T = conv_loss(xc[:nw].view(kw, kh, cin, cout), xc[nw:], data, target)
        total = total + fx_T
        grads = torch.autograd.grad(total, [p["conv"], p["cw"]])
        assert abs(cost - float(fx_T)) <= REL_TOL * abs(float(fx_T)), it
        assert rel_err(np.concatenate([a.reshape(-1) for a in xs]), xc.detach()) <= REL_TOL, it
        for k, g in zip(("conv", "cw"), grads):
            th[k], m_, v_ = orc.tf_adam_step(th[k], g, adam[k][0], adam[k][1], it + 1, lr=0.001)
            adam[k] = (m_, v_)
            big = g.abs() > 1e-5 * float(g.abs().max())
            assert rel_err(prog.nets[k].theta.cpu()[big], th[k][big]) <= REL_TOL, (it, k)
        x = xc.detach()
        s_conv = tuple((h.detach(), c.detach()) for h, c in sc)
        s_cw = tuple((h.detach(), c.detach()) for h, c in sw_)
