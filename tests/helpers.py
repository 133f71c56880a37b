"""Shared helpers for the parity tests: spec <-> handle construction, error metric."""
import math

import torch

from oracle import l2o_oracle as orc

REL_TOL = 1e-5  # north_star: "within 1e-5 relative fp32"

SPECS = {
    "dm_identity": orc.NetSpec(layers=(20, 20)),
    "dm_logsign": orc.NetSpec(layers=(20, 20), preprocess_name="LogAndSign", preprocess_options={"k": 5}, scale=0.01),
    "rnnprop": orc.NetSpec(layers=(20, 20), preprocess_name="fc", preprocess_options={"dim": 20}, scale=0.01,
                           tanh_output=True, rnnprop=True),
    "empty": orc.NetSpec(layers=()),
    "one": orc.NetSpec(layers=(1,)),
    "one_one": orc.NetSpec(layers=(1, 1)),
    "two_three": orc.NetSpec(layers=(2, 3)),
}


def make_handle(spec):
    from open_l2o_b200.engine import NetHandle
    return NetHandle(layers=spec.layers, preprocess_name=spec.preprocess_name,
                     preprocess_options=spec.preprocess_options, scale=spec.scale, tanh_output=spec.tanh_output,
                     n_in=spec.n_in)


def rel_err(a, b):
    """max-norm relative error of a against reference b."""
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    den = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / den


def state_to_arena(state, n):
    """tuple over layers of (h, c) [n, H] -> flat arena (include/l2o_b200.h layout)."""
    parts = []
    for h, c in state:
        parts += [h.reshape(-1), c.reshape(-1)]
    if not parts:
        return torch.zeros(1)
    return torch.cat(parts).contiguous()


def arena_to_state(arena, layers, n):
    out, off = [], 0
    for h in layers:
        hh = arena[off:off + n * h].view(n, h)
        cc = arena[off + n * h:off + 2 * n * h].view(n, h)
        out.append((hh, cc))
        off += 2 * n * h
    return tuple(out)


def random_state(spec, n, gen, amp=0.5, dtype=torch.float32):
    return tuple(((torch.rand(n, h, generator=gen, dtype=torch.float64) * 2 - 1).mul(amp).to(dtype),
                  (torch.rand(n, h, generator=gen, dtype=torch.float64) * 2 - 1).mul(amp).to(dtype))
                 for h in spec.layers)


def wild_gradients(n, gen):
    """Gradients spanning many magnitudes, with exact zeros and both signs (LogAndSign edge cases)."""
    e = torch.randint(-12, 4, (n,), generator=gen).double()
    g = torch.randn(n, generator=gen, dtype=torch.float64) * (10.0 ** e)
    g[::17] = 0.0
    g[1::29] = 1.0
    g[2::31] = -1e-3
    return g.float()


def assert_theta_close(theta_gpu, trainer, tag=None, tol=REL_TOL, tol_all=5e-5):
    """theta after TF-Adam against the oracle trainer.  Adam's update is ~ lr * sign(g) in its first steps, so entries
    whose meta-gradient sits at round-off level (|g| <= 1e-5 max|g|) may legitimately differ by a fraction of lr; every
    other entry must agree to `tol` (max-norm relative), and all entries to `tol_all`."""
    th = torch.as_tensor(theta_gpu).detach().cpu()
    g = trainer.last_grad.detach().abs()
    big = g > 1e-5 * float(g.max())
    assert rel_err(th[big], trainer.theta[big]) <= tol, (tag, rel_err(th[big], trainer.theta[big]))
    assert rel_err(th, trainer.theta) <= tol_all, (tag, rel_err(th, trainer.theta))
