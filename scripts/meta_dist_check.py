"""Sharded meta-step check (run under torchrun): ``MetaOptimizer(_distributed=True)`` with the coordinates of
BASELINE config #5's optimizee split over the ranks (no data-path collective, ONE all-reduce of
[dtheta | fx] per meta-step, identical TF-Adam on every rank) must reproduce the single-GPU optimizer on the same
problem: f(x_T), theta after every meta-step, and the final x shard.

Backend: NCCL with one rank per GPU when the box has enough GPUs; otherwise every rank runs its kernels on cuda:0 and
the collective goes through gloo (CUDA tensors staged by the backend) - the engine path (kernels, packing, sharding)
is the same, only the transport differs.  The line printed says which."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_l2o_b200 import meta, problems  # noqa: E402
from open_l2o_b200.dist import shard_range  # noqa: E402

N, T, UNROLLS = 40000, 10, 3
CFG = {"cw": {"net": "CoordinateWiseDeepLSTM", "net_options": {"layers": (20, 20), "scale": 0.1}}}


def run(shard, distributed):
    optimizer = meta.MetaOptimizer(_seed=7, _distributed=distributed, **CFG)
    _out = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        ms = optimizer.meta_minimize(problems.rastrigin_separable(num_dims=N, shard=shard), T, learning_rate=0.001)
    finally:
        sys.stdout = _out
    sess = meta.Session()
    sess.run(ms.reset)
    net = next(iter(optimizer.program.nets.values()))
    trace = []
    for _ in range(UNROLLS):
        cost = sess.run([ms.fx, ms.update, ms.step])[0]
        trace.append((cost, net.theta.detach().clone(), next(iter(optimizer.program.dtheta.values())).detach().clone()))
    return trace, optimizer.program.X.detach().clone()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    one_gpu_each = torch.cuda.device_count() >= world
    local = int(os.environ["LOCAL_RANK"]) if one_gpu_each else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    backend = "nccl" if one_gpu_each else "gloo"
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    lo, hi = shard_range(N, rank, world)
    trace, x_shard = run((lo, hi), True)
    # every rank must hold the identical theta (rank-0 checksum == every rank's)
    th = trace[-1][1]
    ref0 = th.clone()
    dist.broadcast(ref0, src=0)
    same = torch.tensor([1.0 if torch.equal(ref0, th) else 0.0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ok = bool(same.item() == 1.0)
    if rank == 0:
        ref, x_full = run(None, False)

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        # Adam's update is ~ lr * sign(g) in its first steps: entries whose meta-gradient sits at round-off level may
        # differ by a fraction of lr between two summation orders; every other entry must agree to 1e-5
        def rel_big(t, tr, g):
            big = g.abs() > 1e-5 * g.abs().max()
            return rel(t[big], tr[big])
        errs = {"fx": max(abs(c - cr) / abs(cr) for (c, _, _), (cr, _, _) in zip(trace, ref)),
                "dtheta": max(rel(g, gr) for (_, _, g), (_, _, gr) in zip(trace, ref)),
                "theta": max(rel_big(t, tr, gr) for (_, t, _), (_, tr, gr) in zip(trace, ref)),
                "theta_all": max(rel(t, tr) for (_, t, _), (_, tr, _) in zip(trace, ref)),
                "x_shard": rel(x_shard, x_full[lo:hi])}
        # d-theta of this problem is a sum with heavy cancellation: two runs of the EXACT-fp32 engine that differ only in
        # how the coordinates are grouped into CTAs already differ by 1.6e-5 of max|dtheta| (measured, L2O_TC_AUTO=0),
        # so the sharded-vs-single bar for d-theta / theta is 5e-5; f(x_T) and x have no such cancellation: 1e-5
        ok = ok and errs["fx"] <= 1e-5 and errs["x_shard"] <= 1e-5 and \
            all(errs[k] <= 5e-5 for k in ("dtheta", "theta", "theta_all"))
        print("meta sharded vs single-GPU (world %d, backend %s, %d GPU(s)):" % (world, backend, torch.cuda.device_count()),
              errs, "theta identical on all ranks:", bool(same.item() == 1.0), "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
