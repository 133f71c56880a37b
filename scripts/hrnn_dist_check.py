"""2-GPU check of the sharded HierarchicalRNN step (run under torchrun, one rank per GPU): K steps with every tensor's
coordinates split over the ranks must reproduce the single-GPU optimizer on the same problem."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_l2o_b200 import hierarchical_rnn as hr  # noqa: E402
from open_l2o_b200.dist import shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    # NCCL with one rank per GPU when the box has enough GPUs; otherwise all ranks run on cuda:0 and the collectives go
    # through gloo (same kernels and sharding logic, different transport)
    one_gpu_each = torch.cuda.device_count() >= world
    local = int(os.environ["LOCAL_RANK"]) if one_gpu_each else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_gpu_each:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    shapes = [(300, 7), (5,), (1,), (64, 33)]          # includes tensors smaller than / not divisible by the world size
    K = 5

    def problem():
        gen = torch.Generator().manual_seed(3)
        ps = [torch.randn(s, generator=gen).to(dev) for s in shapes]
        gs = [[(torch.randn(s, generator=gen) * (0.3 if t % 2 == 0 else 3e-3)).to(dev) for s in shapes] for t in range(K)]
        return ps, gs

    ps, gs = problem()
    opt = hr.HierarchicalRNN(random_seed=0, device=dev, distributed=True, **hr.metarun_flags())
    for t in range(K):
        opt.apply_gradients(zip(gs[t], ps))
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ps1, gs1 = problem()
        ref = hr.HierarchicalRNN(random_seed=0, device=dev, **hr.metarun_flags())
        for t in range(K):
            ref.apply_gradients(zip(gs1[t], ps1))
        torch.cuda.synchronize()

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        errs = {"x": max(rel(a, b) for a, b in zip(ps, ps1)), "global": rel(opt.global_state, ref.global_state),
                "layer": rel(opt.layer, ref.layer)}
        hid = []
        for j, s in enumerate(shapes):
            n = int(torch.tensor(s).prod())
            lo, hi = shard_range(n, 0, world)
            if hi > lo:
                hid.append(rel(opt.get_slot(j, "parameter"), ref.get_slot(j, "parameter")[lo:hi]))
                hid.append(rel(opt.get_slot(j, "log_learning_rate"), ref.get_slot(j, "log_learning_rate")[lo:hi]))
        errs["state_shard"] = max(hid)
        ok = all(v <= 1e-5 for v in errs.values())
        print("hrnn sharded vs single-GPU (world %d, backend %s):" % (world, dist.get_backend()), errs,
              "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
