"""Prints the meta-gradient error of the engine against autograd through the fp64 oracle, block by block of theta, for a few
unroll lengths and a multi-tile problem; then times one meta-training step on the ConvNet of BASELINE config #4."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hrnn_oracle as orc  # noqa: E402
from open_l2o_b200 import hrnn_train as ht  # noqa: E402
from tests.test_hrnn_train_gpu import _groups, _oracle_meta_gradient  # noqa: E402

DEV = "cuda:0"


def problem(sizes, seed, dtype, device):
    gen = torch.Generator().manual_seed(seed)
    tgt = [torch.randn(s, generator=gen, dtype=torch.float64).to(device=device, dtype=dtype) for s in sizes]

    def objective(params):
        return sum(((p - t) ** 2).mean() + 0.05 * torch.cos(2.0 * p).mean() for p, t in zip(params, tgt))
    init = [torch.randn(s, generator=gen, dtype=torch.float64) * 0.5 for s in sizes]
    return objective, init


for sizes, T in (([(300, 40), (40,), (1000,)], 3), ([(30, 7), (7,), (150,)], 20)):
    o64, init = problem(sizes, 1, torch.float64, "cpu")
    o32, _ = problem(sizes, 1, torch.float32, DEV)
    theta = orc.init_theta(seed=4)
    n = sum(int(math.prod(s)) for s in sizes)
    llr = (torch.rand(n, generator=torch.Generator().manual_seed(5), dtype=torch.float64) * 3.0 - 6.0).float()
    mref, gref, oref, xref = _oracle_meta_gradient(theta, o64, init, llr, T)
    tr = ht.MetaTrainer(sizes, theta=theta, device=DEV)
    meta, g, objs, final = tr.meta_gradient(o32, [p.float().to(DEV) for p in init], T, log_learning_rate=llr)
    g = g.detach().cpu().double()
    sc = float(gref.abs().max())
    rows = sorted(((float((g[lo:hi] - gref[lo:hi]).abs().max()) / sc, nm) for nm, lo, hi in _groups()), reverse=True)
    print("sizes %s T=%d: meta %.7f (oracle %.7f)  x err %.2e  worst blocks:" % (
        sizes, T, float(meta), mref, float((final.x.detach().cpu().double() - xref).abs().max() / xref.abs().max())),
        ["%.1e %s" % r for r in rows[:4]])

# timing: ConvNet optimizee of BASELINE config #4 (354 K coordinates), unroll 20
from open_l2o_b200.scale_problems import ConvNet  # noqa: E402
prob = ConvNet((3, 32, 32), 10, [(3, 3, 32), (5, 5, 32)])
params0 = [p.detach() for p in prob.init_tensors(seed=0, device=DEV)]
gen = torch.Generator().manual_seed(0)
data = torch.randn(128, 32, 32, 3, generator=gen).to(DEV)
labels = torch.nn.functional.one_hot(torch.randint(0, 10, (128,), generator=gen), 10).float().to(DEV)
tr = ht.MetaTrainer(prob.param_shapes, device=DEV, random_seed=0)
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.time()
    meta, objs, _ = tr.train_step(lambda ps: prob.objective(ps, data, labels), params0, 20)
    torch.cuda.synchronize()
    print("ConvNet %d coords, unroll 20: meta-train step %.1f ms, meta objective %.5f, f: %.4f -> %.4f" % (
        tr.engine.N, 1e3 * (time.time() - t0), meta, objs[0], objs[-1]))
