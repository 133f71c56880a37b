"""GPU parity tests: CUDA kernels (through the C-ABI) against the CPU oracle on identical tensors.

Tolerance: max-norm relative error <= 1e-5 (fp32), the bar BASELINE.json's north_star states.
Gradient (d theta) checks are made against the fp64 oracle and allow the fp32 oracle's own
distance to fp64 as slack (a sum over N*T terms in fp32 is itself only ~1e-6..1e-5 accurate).
"""
import math

import pytest
import torch

from oracle import l2o_oracle as orc
from tests.helpers import (REL_TOL, SPECS, arena_to_state, make_handle, random_state, rel_err, state_to_arena,
                           wild_gradients)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _theta(spec, seed=0, gain=0.1):
    return orc.init_theta(spec, seed=seed, out_gain=gain)


def _oracle_inputs(spec, g, gen):
    if spec.rnnprop:
        g2 = torch.randn(g.numel(), generator=gen)
        return torch.stack([g, g2], dim=-1)
    return g.unsqueeze(-1)


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("n", [1, 127, 1000])
def test_step_parity(name, n):
    spec = SPECS[name]
    gen = torch.Generator().manual_seed(1234 + n)
    theta = _theta(spec)
    g = wild_gradients(n, gen) if spec.preprocess_name == "LogAndSign" else torch.randn(n, generator=gen)
    inp = _oracle_inputs(spec, g, gen)
    st = random_state(spec, n, gen)
    x0 = torch.randn(n, generator=gen)
    d_ref, st_ref = orc.net_apply(spec, theta, inp, st)

    h = make_handle(spec)
    arena_in = state_to_arena(st, n).to(DEV)
    arena_out = torch.zeros_like(arena_in)
    x = x0.to(DEV).clone()
    delta = torch.empty(n, device=DEV)
    in0 = inp[:, 0].contiguous().to(DEV)
    in1 = inp[:, 1].contiguous().to(DEV) if spec.rnnprop else None
    h.step(theta.to(DEV), in0, arena_in, arena_out, in1=in1, x=x, delta=delta)
    torch.cuda.synchronize()
    assert rel_err(delta, d_ref) <= REL_TOL
    assert rel_err(x, x0 + d_ref) <= REL_TOL
    for (hg, cg), (hr, cr) in zip(arena_to_state(arena_out.cpu(), spec.layers, n), st_ref):
        assert rel_err(hg, hr) <= REL_TOL
        assert rel_err(cg, cr) <= REL_TOL
    # in-place state update must give the same answer
    h.step(theta.to(DEV), in0, arena_in, arena_in, in1=in1)
    torch.cuda.synchronize()
    if h.state_floats:
        assert torch.equal(arena_in, arena_out)


def test_step_zero_output_layer_gives_zero_update():
    """SW/networks_test.py:57-69: zero-initialised final Linear => update == 0 through two LSTM layers."""
    spec = SPECS["dm_identity"]
    theta = _theta(spec)
    shapes = spec.shapes()
    n_out = sum(math.prod(s) for m, v, s in shapes if m == "linear")
    theta[-n_out:] = 0
    n = 300
    h = make_handle(spec)
    arena = h.new_state(n, DEV)
    delta = torch.full((n,), 7.0, device=DEV)
    h.step(theta.to(DEV), torch.randn(n, device=DEV), arena, arena, delta=delta)
    assert float(delta.abs().max()) == 0.0
    assert float(arena.abs().max()) > 0.0


def test_log_and_sign_kernel():
    """SW/preprocess_test.py:78-98 + parity with DM/preprocess.py:52-70."""
    from open_l2o_b200.engine import log_and_sign
    gen = torch.Generator().manual_seed(5)
    g = wild_gradients(4097, gen)
    out = log_and_sign(g.to(DEV), 5.0).cpu()
    ref = orc.log_and_sign(g.unsqueeze(-1), 5.0)
    assert rel_err(out[0], ref[:, 0]) <= REL_TOL
    assert rel_err(out[1], ref[:, 1]) <= REL_TOL
    ones = log_and_sign(torch.ones(8, device=DEV), 1.0).cpu()
    assert float(ones[0].abs().max()) < 1e-6          # log(1)/k ~ 0
    nz = g != 0
    assert torch.equal(torch.sign(out[1][nz]), torch.sign(g[nz]))


def _run_prerecorded(spec, n, T, seed, engine=None):
    gen = torch.Generator().manual_seed(seed)
    theta = _theta(spec)
    if spec.rnnprop:
        seq = torch.randn(T, 2, n, generator=gen)
    else:
        seq = torch.randn(T, n, generator=gen) * 0.3
        if spec.preprocess_name == "LogAndSign":
            seq = torch.stack([wild_gradients(n, gen) for _ in range(T)])
    st = random_state(spec, n, gen, amp=0.2)
    x0 = torch.randn(n, generator=gen)
    # oracle
    s, x, deltas, states = st, x0.clone(), [], [st]
    for t in range(T):
        inp = seq[t].t() if spec.rnnprop else seq[t].unsqueeze(-1)
        d, s = orc.net_apply(spec, theta, inp, s)
        x = x + d
        deltas.append(d)
        states.append(s)
    # kernel
    h = make_handle(spec)
    if engine is not None:
        h.set_engine(engine)
    sf = h.state_floats
    arena = state_to_arena(st, n).to(DEV)
    ckpt = torch.zeros((T + 1) * max(sf * n, 1), device=DEV)
    xg = x0.to(DEV).clone()
    dseq = torch.empty(T, n, device=DEV)
    h.unroll_fwd(theta.to(DEV), n, T, arena, in_seq=seq.contiguous().to(DEV), x=xg, ckpt=ckpt, delta_seq=dseq)
    torch.cuda.synchronize()
    return dict(spec=spec, n=n, T=T, theta=theta, seq=seq, st=st, x0=x0, x_ref=x, deltas=deltas, states=states,
                h=h, arena=arena, ckpt=ckpt, xg=xg, dseq=dseq, sf=sf)


@pytest.mark.parametrize("name", ["dm_identity", "dm_logsign", "rnnprop", "empty", "one_one", "two_three"])
def test_unroll_fwd_prerecorded(name):
    r = _run_prerecorded(SPECS[name], n=777, T=20, seed=7)
    spec, n, T, sf = r["spec"], r["n"], r["T"], r["sf"]
    assert rel_err(r["dseq"], torch.stack(r["deltas"])) <= REL_TOL
    assert rel_err(r["xg"], r["x_ref"]) <= REL_TOL
    if sf:
        for t in (0, 1, T // 2, T):
            got = arena_to_state(r["ckpt"][t * sf * n:(t + 1) * sf * n].cpu(), spec.layers, n)
            for (hg, cg), (hr, cr) in zip(got, r["states"][t]):
                assert rel_err(hg, hr) <= REL_TOL and rel_err(cg, cr) <= REL_TOL
        fin = arena_to_state(r["arena"].cpu(), spec.layers, n)
        for (hg, cg), (hr, cr) in zip(fin, r["states"][T]):
            assert rel_err(hg, hr) <= REL_TOL and rel_err(cg, cr) <= REL_TOL


def _fused_problem(kind, n, gen, dtype=torch.float32):
    if kind == "rastrigin_sep":
        a, b = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
        x0 = torch.randn(n, generator=gen)
        return orc.FusedProblem(kind, a.to(dtype), b.to(dtype), alpha=10.0, fscale=1.0 / n), x0.to(dtype)
    a, b = torch.rand(n, generator=gen) + 0.5, torch.rand(n, generator=gen)
    x0 = torch.randn(n, generator=gen) * 0.01
    return orc.FusedProblem(kind, a.to(dtype), b.to(dtype), fscale=1.0 / n), x0.to(dtype)


@pytest.mark.parametrize("kind", ["rastrigin_sep", "quadratic_diag"])
@pytest.mark.parametrize("name", ["dm_identity", "rnnprop"])
def test_unroll_fused_forward_and_backward(kind, name):
    """Fused regime: gradient evaluated in-kernel, T steps in one launch, then BPTT; compared with the
    oracle's autograd through the same unroll (DM/meta.py:338-376, 412)."""
    from open_l2o_b200.engine import OPT_KINDS
    spec = SPECS[name]
    n, T = 1500, 20
    gen = torch.Generator().manual_seed(99)
    theta = _theta(spec, gain=0.05 if name == "dm_identity" else 1.0)
    prob, x0 = _fused_problem(kind, n, gen)
    st0 = orc.initial_state(spec, n)
    mv0 = (torch.zeros(n), torch.zeros(n)) if spec.rnnprop else None
    g32, res = orc.meta_grad(spec, theta, x0, st0, None, T, mv0=mv0, step0=1, grad_of=prob.f_and_g)
    prob64 = orc.FusedProblem(kind, prob.a.double(), prob.b.double(), prob.alpha, prob.fscale)
    mv64 = (torch.zeros(n, dtype=torch.float64),) * 2 if spec.rnnprop else None
    g64, res64 = orc.meta_grad(spec, theta.double(), x0.double(), orc.initial_state(spec, n, torch.float64), None, T,
                               mv0=mv64, step0=1, grad_of=prob64.f_and_g)

    h = make_handle(spec)
    sf = h.state_floats
    th = theta.to(DEV)
    arena = h.new_state(n, DEV)
    ckpt = torch.zeros((T + 1) * sf * n, device=DEV)
    x = x0.to(DEV).clone()
    g_rec = torch.empty(T + 1, n, device=DEV)
    fx = torch.zeros(T + 1, dtype=torch.float64, device=DEV)
    kw = {}
    if spec.rnnprop:
        kw = dict(m=torch.zeros(n, device=DEV), v=torch.zeros(n, device=DEV), beta1=0.95, beta2=0.95, step0=1,
                  feat_rec=torch.empty(T, 2, n, device=DEV))
    h.unroll_fwd(th, n, T, arena, opt_kind=OPT_KINDS[kind], opt_a=prob.a.to(DEV), opt_b=prob.b.to(DEV),
                 opt_alpha=prob.alpha, opt_fscale=prob.fscale, x=x, ckpt=ckpt, g_rec=g_rec, fx=fx, **kw)
    torch.cuda.synchronize()
    assert rel_err(fx, res64.fx) <= REL_TOL
    assert rel_err(x, res64.x_final) <= REL_TOL
    assert rel_err(g_rec[:T], torch.stack(res64.grads)) <= REL_TOL
    if spec.rnnprop:
        assert rel_err(kw["m"], res64.mv_final[0]) <= REL_TOL
        assert rel_err(kw["v"], res64.mv_final[1]) <= REL_TOL

    dtheta = torch.zeros(h.n_theta, dtype=torch.float64, device=DEV)
    in_seq = kw["feat_rec"] if spec.rnnprop else g_rec
    h.unroll_bwd(th, n, T, in_seq, ckpt, dtheta, g_rec=g_rec)
    torch.cuda.synchronize()
    slack = max(REL_TOL, 3.0 * rel_err(g32, g64))
    assert rel_err(dtheta, g64) <= slack, (rel_err(dtheta, g64), rel_err(g32, g64))


@pytest.mark.parametrize("name", ["dm_identity", "dm_logsign", "empty", "one", "one_one", "two_three"])
def test_unroll_bwd_external_gradients(name):
    """External-gradient regime: the optimizee is a dense batched quadratic (DM/problems.py:73-101) whose
    gradients come from outside the kernel; BPTT consumes the recorded g_0..g_T."""
    spec = SPECS[name]
    B, d, T = 16, 10, 12
    n = B * d
    gen = torch.Generator().manual_seed(3)
    theta = _theta(spec, gain=0.05)
    w = torch.rand(B, d, d, generator=gen)
    y = torch.rand(B, d, generator=gen)
    x0 = (torch.randn(B, d, generator=gen) * 0.01)
    f32 = lambda x: orc.quadratic_f(x, w, y)
    f64 = lambda x: orc.quadratic_f(x, w.double(), y.double())
    g32, res = orc.meta_grad(spec, theta, x0, orc.initial_state(spec, n), f32, T)
    g64, res64 = orc.meta_grad(spec, theta.double(), x0.double(), orc.initial_state(spec, n, torch.float64), f64, T)
    xT = res.x_final.detach().requires_grad_(True)
    (gT,) = torch.autograd.grad(f32(xT), xT)
    g_rec = torch.stack(res.grads + [gT.reshape(-1)]).contiguous().to(DEV)

    h = make_handle(spec)
    sf = h.state_floats
    th = theta.to(DEV)
    arena = h.new_state(n, DEV)
    ckpt = torch.zeros((T + 1) * max(sf * n, 1), device=DEV)
    x = x0.reshape(-1).to(DEV).clone()
    h.unroll_fwd(th, n, T, arena, in_seq=g_rec[:T].contiguous(), x=x, ckpt=ckpt)
    assert rel_err(x, res.x_final) <= REL_TOL
    dtheta = torch.zeros(h.n_theta, dtype=torch.float64, device=DEV)
    h.unroll_bwd(th, n, T, g_rec[:T].contiguous(), ckpt, dtheta, g_rec=g_rec)
    torch.cuda.synchronize()
    slack = max(REL_TOL, 3.0 * rel_err(g32, g64))
    assert rel_err(dtheta, g64) <= slack, (rel_err(dtheta, g64), rel_err(g32, g64))


@pytest.mark.parametrize("name", ["dm_identity", "rnnprop"])
def test_imitation_unroll(name):
    """DM/meta_dm_train.py:463-480: LSTM over pre-recorded inputs, loss = sum_t 0.5||label - delta||^2 / N."""
    spec = SPECS[name]
    n, T = 900, 10
    gen = torch.Generator().manual_seed(11)
    theta = _theta(spec, gain=1.0)
    inputs = torch.randn(T, n, 2, generator=gen) if spec.rnnprop else torch.randn(T, n, generator=gen)
    labels = torch.randn(T, n, generator=gen) * 0.01
    th64 = theta.double().requires_grad_(True)
    loss64, _, _ = orc.imitation_loss(spec, th64, inputs.double(), labels.double(), orc.initial_state(spec, n, torch.float64))
    (g64,) = torch.autograd.grad(loss64, th64)
    th32 = theta.clone().requires_grad_(True)
    loss32, _, _ = orc.imitation_loss(spec, th32, inputs, labels, orc.initial_state(spec, n))
    (g32,) = torch.autograd.grad(loss32, th32)

    h = make_handle(spec)
    sf = h.state_floats
    th = theta.to(DEV)
    seq = (inputs.permute(0, 2, 1) if spec.rnnprop else inputs).contiguous().to(DEV)
    arena = h.new_state(n, DEV)
    ckpt = torch.zeros((T + 1) * sf * n, device=DEV)
    il = torch.zeros(1, dtype=torch.float64, device=DEV)
    lab = labels.to(DEV)
    h.unroll_fwd(th, n, T, arena, in_seq=seq, ckpt=ckpt, labels=lab, imit_loss=il, n_total=n)
    assert rel_err(il, loss64) <= REL_TOL
    dtheta = torch.zeros(h.n_theta, dtype=torch.float64, device=DEV)
    h.unroll_bwd(th, n, T, seq, ckpt, dtheta, labels=lab, n_total=n)
    torch.cuda.synchronize()
    slack = max(REL_TOL, 3.0 * rel_err(g32, g64))
    assert rel_err(dtheta, g64) <= slack, (rel_err(dtheta, g64), rel_err(g32, g64))


def test_adam_step_matches_tf_formula():
    from open_l2o_b200.engine import adam_step
    gen = torch.Generator().manual_seed(2)
    n = 5061
    theta = torch.randn(n, generator=gen)
    m = torch.zeros(n)
    v = torch.zeros(n)
    tg, mg, vg = theta.to(DEV).clone(), m.to(DEV), v.to(DEV)
    for k in range(1, 4):
        grad = torch.randn(n, generator=gen, dtype=torch.float64)
        theta, m, v = orc.tf_adam_step(theta, grad.float(), m, v, k, lr=0.01)
        adam_step(tg, grad.to(DEV), mg, vg, k, lr=0.01)
    torch.cuda.synchronize()
    assert rel_err(tg, theta) <= REL_TOL
    assert rel_err(mg, m) <= REL_TOL and rel_err(vg, v) <= REL_TOL


def test_unsupported_shape_is_rejected_loudly():
    from open_l2o_b200.engine import NetHandle
    from open_l2o_b200._lib import L2OError
    with pytest.raises(L2OError):
        NetHandle(layers=(7, 9))
    h = make_handle(SPECS["dm_identity"])
    with pytest.raises(L2OError):
        h.step(torch.zeros(h.n_theta), torch.zeros(4), torch.zeros(320), torch.zeros(320))  # CPU tensors


@pytest.mark.parametrize("B,d", [(128, 10), (37, 7), (3, 128), (50, 1)])
def test_unroll_fused_quadratic_batch(B, d):
    """L2O_OPT_QUADRATIC_BATCH (BASELINE config #1's optimizee, DM/problems.py:73-101): dense W_b x_b evaluated
    in-kernel with the group's x exchanged through shared memory — forward unroll + recorded gradients + BPTT
    against the oracle's autograd through the same unroll.  Group sizes that do / do not divide the 128-thread tile."""
    from open_l2o_b200.engine import OPT_KINDS
    spec = SPECS["dm_identity"]
    n, T = B * d, 12
    gen = torch.Generator().manual_seed(7)
    theta = _theta(spec, gain=0.05)
    w = torch.rand(B, d, d, generator=gen)
    y = torch.rand(B, d, generator=gen)
    x0 = torch.randn(n, generator=gen) * 0.01
    f32 = lambda x: orc.quadratic_f(x.reshape(B, d), w, y)
    f64 = lambda x: orc.quadratic_f(x.reshape(B, d), w.double(), y.double())
    g32, res32 = orc.meta_grad(spec, theta, x0, orc.initial_state(spec, n), f32, T)
    g64, res64 = orc.meta_grad(spec, theta.double(), x0.double(), orc.initial_state(spec, n, torch.float64), f64, T)
    h = make_handle(spec)
    sf = h.state_floats
    th = theta.to(DEV)
    arena = h.new_state(n, DEV)
    ckpt = torch.zeros((T + 1) * sf * n, device=DEV)
    x = x0.to(DEV).clone()
    g_rec = torch.empty(T + 1, n, device=DEV)
    fx = torch.zeros(T + 1, dtype=torch.float64, device=DEV)
    h.unroll_fwd(th, n, T, arena, opt_kind=OPT_KINDS["quadratic_batch"], opt_a=w.reshape(-1).to(DEV),
                 opt_b=y.reshape(-1).to(DEV), opt_fscale=1.0 / B, opt_group=d, x=x, ckpt=ckpt, g_rec=g_rec, fx=fx)
    torch.cuda.synchronize()
    slack = max(REL_TOL, 3.0 * rel_err(res32.x_final, res64.x_final))
    assert rel_err(fx, res64.fx) <= slack
    assert rel_err(x, res64.x_final) <= slack
    dtheta = torch.zeros(h.n_theta, dtype=torch.float64, device=DEV)
    h.unroll_bwd(th, n, T, g_rec, ckpt, dtheta, g_rec=g_rec)
    torch.cuda.synchronize()
    gslack = max(REL_TOL, 3.0 * rel_err(g32, g64))
    assert rel_err(dtheta, g64) <= gslack, (rel_err(dtheta, g64), rel_err(g32, g64))
    with pytest.raises(Exception):   # a group must not straddle the problem
        h.unroll_fwd(th, n, T, arena, opt_kind=OPT_KINDS["quadratic_batch"], opt_a=w.reshape(-1).to(DEV),
                     opt_b=y.reshape(-1).to(DEV), opt_fscale=1.0 / B, opt_group=d + 1 if n % (d + 1) else 0, x=x)


@pytest.mark.parametrize("B,m,n", [(3, 7, 13), (2, 250, 500), (5, 64, 33), (1, 1, 1)])
@pytest.mark.parametrize("scaled", [False, True])
def test_lasso_grad_producer(B, m, n, scaled):
    """l2o_lasso_grad (f and df/dx of DM/problems.py:103-175 in one launch) vs torch autograd of the oracle's loss in
    fp64, incl. exact zeros in x (sign(0) = 0) and the random-scaling chain rule (DM/meta_dm_train.py:384-385)."""
    from open_l2o_b200 import engine
    gen = torch.Generator().manual_seed(17)
    A = torch.randn(B, m, n, generator=gen) / (m ** 0.5)
    y = torch.randn(B, m, 1, generator=gen)
    x = torch.randn(B, n, generator=gen) * 0.3
    x.view(-1)[::5] = 0.0
    sc = torch.exp(torch.rand(B, n, generator=gen) * 6 - 3) if scaled else None
    xd = x.double().requires_grad_(True)
    f_ref = orc.lasso_f(xd * sc.double() if scaled else xd, A.double(), y.double(), 0.005)
    (g_ref,) = torch.autograd.grad(f_ref, xd)
    g = torch.empty(B * n, device=DEV)
    f = torch.zeros((), dtype=torch.float64, device=DEV)
    engine.lasso_grad(A.to(DEV), y.to(DEV), x.reshape(-1).to(DEV), 0.005, g, f=f,
                      scale=sc.reshape(-1).to(DEV) if scaled else None)
    torch.cuda.synchronize()
    assert abs(float(f) - float(f_ref)) <= REL_TOL * abs(float(f_ref))
    assert rel_err(g, g_ref) <= REL_TOL
