"""L2O-Scale HierarchicalRNN step: CUDA path through the C-ABI vs the CPU oracle."""
import pytest
import torch

from oracle import hrnn_oracle as H
from tests.helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(shapes, seed=3):
    from open_l2o_b200 import hierarchical_rnn as hr
    opt = hr.HierarchicalRNN(random_seed=seed, **hr.metarun_flags())
    theta = opt.theta.detach().cpu().clone()
    assert theta.numel() == H.theta_count()
    gen = torch.Generator().manual_seed(seed + 1)
    params = [torch.randn(s, generator=gen) for s in shapes]
    return opt, theta, params, gen


@pytest.mark.parametrize("shapes,steps", [([(3, 3, 3, 8), (8,), (40, 5), (5,)], 6),   # ragged, tiny tensors
                                          ([(700, 300), (1,), (257,)], 4),               # multi-block tensor + size 1
                                          ([(64,)], 3)])
def test_hrnn_steps_match_oracle(shapes, steps):
    opt, theta, params, gen = _setup(shapes)
    gvars = [p.clone().to(DEV) for p in params]
    grads0 = [torch.randn(s, generator=gen) * 0.3 for s in shapes]
    opt.apply_gradients(zip([g.to(DEV) for g in grads0], gvars))          # creates the slots, then steps
    # rebuild the oracle's initial state from the engine's own initial draw (log-lr is random): re-run from scratch
    P = H.unpack_theta(theta)
    opt.reset_state(seed=11)
    for v, p in zip(gvars, params):
        v.data.copy_(p.to(DEV))
    llr = opt.state[12].detach().cpu().clone()
    states32, states64 = [], []
    off = 0
    for p in params:
        n = p.numel()
        st = H.initial_state(P, p, torch.Generator().manual_seed(0))
        st["log_learning_rate"] = llr[off:off + n].reshape(n, 1).clone()
        states32.append(st)
        states64.append({k: v.double() for k, v in st.items()})
        off += n
    g32, g64 = H.initial_global_state(P), H.initial_global_state(P).double()
    p32, p64 = [p.clone() for p in params], [p.double() for p in params]
    th64 = theta.double()
    for t in range(steps):
        grads = [torch.randn(s, generator=gen) * (0.3 if t % 2 == 0 else 3e-3) for s in shapes]
        opt.apply_gradients(zip([g.to(DEV) for g in grads], gvars))
        p32, states32, g32, _ = H.step(theta, p32, grads, states32, g32)
        p64, states64, g64, u64 = H.step(th64, p64, [g.double() for g in grads], states64, g64)
        torch.cuda.synchronize()
        slack = REL_TOL
        for j in range(len(shapes)):
            slack = max(slack, 3.0 * rel_err(states32[j]["parameter"], states64[j]["parameter"]),
                        3.0 * rel_err(p32[j], p64[j]))
        off = 0
        for j, p in enumerate(params):
            n = p.numel()
            assert rel_err(gvars[j], p64[j]) <= slack, (t, j, "x")
            e_u = rel_err(opt.update[off:off + n], u64[j])
            assert e_u <= 3 * slack, (t, j, "update", e_u, slack)
            for key in ("parameter", "scl_decay", "inp_decay", "log_learning_rate", "grad_accum1", "grad_accum4",
                        "ms1", "ms4", "layer"):
                e_k = rel_err(opt.get_slot(j, key), states64[j][key])
                assert e_k <= 3 * slack, (t, j, key, e_k, slack)
            off += n
        assert rel_err(opt.global_state, g64) <= 3 * slack, (t, "global", rel_err(opt.global_state, g64), slack)


def test_hrnn_argument_errors():
    from open_l2o_b200 import hierarchical_rnn as hr
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20, 5])
    with pytest.raises(ValueError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20], init_lr_range=(1e-2, 1e-6))
    with pytest.raises(NotImplementedError):
        hr.HierarchicalRNN(level_sizes=[10, 20, 20])      # the reference's signature defaults are not the built flag set
    opt = hr.HierarchicalRNN(**hr.metarun_flags())
    with pytest.raises(ValueError):
        opt.apply_gradients([(None, torch.zeros(3, device=DEV))])


def test_hrnn_minimizes_a_quadratic():
    """Smoke-level behaviour check: with random-init weights the optimizer still moves downhill on a bowl
    (the update direction is the RMS-normalised gradient shortcut, HR:612-626)."""
    from open_l2o_b200 import hierarchical_rnn as hr
    opt = hr.HierarchicalRNN(random_seed=0, **hr.metarun_flags())
    w = torch.randn(300, 30, device=DEV).requires_grad_(True)
    b = torch.randn(30, device=DEV).requires_grad_(True)
    objs = opt.minimize(lambda w, b: (w ** 2).sum() + (b ** 2).sum(), [w, b], 60)
    assert objs[-1] < objs[0]


def test_hrnn_minimize_graph_replay_matches_eager():
    """HierarchicalRNN.minimize captures one (objective, gradients, step) iteration as a CUDA graph; the objective
    trajectory must equal the eager one."""
    from open_l2o_b200 import hierarchical_rnn as hr
    out = []
    for use_graph in (False, True):
        opt = hr.HierarchicalRNN(random_seed=0, **hr.metarun_flags())
        gen = torch.Generator().manual_seed(5)
        w = torch.randn(200, 20, generator=gen).to(DEV).requires_grad_(True)
        b = torch.randn(20, generator=gen).to(DEV).requires_grad_(True)
        t = torch.randn(64, 200, generator=gen).to(DEV)
        out.append(opt.minimize(lambda w, b: ((t @ w + b) ** 2).mean(), [w, b], 12, cuda_graph=use_graph))
    assert len(out[0]) == len(out[1]) == 12
    assert max(abs(a - c) / (abs(a) + 1e-30) for a, c in zip(out[0], out[1])) <= 1e-6


def test_hrnn_sharded_step_matches_single_gpu():
    """Coordinates of every tensor split over 2 ranks (one all-reduce of the per-tensor sums per step + an all-gather of
    the updated parameters, scripts/hrnn_dist_check.py under torchrun) against the single-GPU optimizer."""
    import os
    import subprocess
    import sys
    # < 2 GPUs: both ranks share cuda:0 and the collectives go through gloo (see the script) - never skipped
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "scripts", "hrnn_dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
