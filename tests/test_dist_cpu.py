"""N>1 path on CPU: world_size-2 gloo processes, coordinates sharded, one packed all-reduce of [dtheta | fx]
(open_l2o_b200/dist.py); the sharded result must equal the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import l2o_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, T, out):
    from open_l2o_b200.dist import allreduce_meta_grad, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = orc.NetSpec(layers=(20, 20), scale=0.1)
    theta = orc.init_theta(spec, seed=0, out_gain=1.0)
    gen = torch.Generator().manual_seed(1)
    a, b, x0 = (torch.randn(n, generator=gen) for _ in range(3))
    lo, hi = shard_range(n, rank, world)
    prob = orc.FusedProblem("rastrigin_sep", a[lo:hi], b[lo:hi], 10.0, 1.0 / n)   # fscale uses the GLOBAL n
    g, res = orc.meta_grad(spec, theta, x0[lo:hi], orc.initial_state(spec, hi - lo), None, T, grad_of=prob.f_and_g)
    dtheta = {"cw": g.double().clone()}
    fx = allreduce_meta_grad(dtheta, res.fx.detach().double())
    th, _, _ = orc.tf_adam_step(theta, dtheta["cw"].float(), torch.zeros_like(theta), torch.zeros_like(theta), 1, lr=0.01)
    if rank == 0:
        torch.save({"dtheta": dtheta["cw"], "fx": fx, "theta": th}, out)
    # every rank must hold the identical updated theta
    gathered = [torch.zeros_like(th) for _ in range(world)]
    dist.all_gather(gathered, th)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from open_l2o_b200.dist import shard_range
    for n in (0, 1, 7, 128, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(300)
def test_sharded_meta_gradient_matches_unsharded(tmp_path):
    n, T, world = 600, 6, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, T, out), nprocs=world, join=True)
    got = torch.load(out)
    spec = orc.NetSpec(layers=(20, 20), scale=0.1)
    theta = orc.init_theta(spec, seed=0, out_gain=1.0)
    gen = torch.Generator().manual_seed(1)
    a, b, x0 = (torch.randn(n, generator=gen) for _ in range(3))
    prob = orc.FusedProblem("rastrigin_sep", a, b, 10.0, 1.0 / n)
    g, res = orc.meta_grad(spec, theta, x0, orc.initial_state(spec, n), None, T, grad_of=prob.f_and_g)
    den = float(g.abs().max())
    assert float((got["dtheta"].float() - g).abs().max()) / den <= 2e-5
    assert torch.allclose(got["fx"].float(), res.fx.detach(), rtol=1e-5, atol=1e-7)
    th, _, _ = orc.tf_adam_step(theta, g, torch.zeros_like(theta), torch.zeros_like(theta), 1, lr=0.01)
    # Adam's first step is lr*sign(g): only entries with |g| above the sharding round-off can be compared exactly
    big = g.abs() > 1e-6 * den
    assert torch.allclose(got["theta"][big], th[big], atol=1e-6)


def _gather_worker(rank, world, port, n, out):
    from open_l2o_b200.dist import allgather_shards, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n, dtype=torch.float32) * 0.5 - 3.0
    lo, hi = shard_range(n, rank, world)
    got = allgather_shards(full[lo:hi].clone(), n)
    assert torch.equal(got, full), (rank, got, full)
    if rank == 0:
        torch.save(got, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n", [1, 2, 7, 1000])
def test_allgather_shards_reassembles_uneven_slices(tmp_path, n):
    """Host logic of the sharded HierarchicalRNN step: uneven (and empty) shard_range slices padded to one equal-size
    all-gather must give back the full tensor on every rank (gloo, world size 2)."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_gather_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    assert torch.equal(torch.load(out), torch.arange(n, dtype=torch.float32) * 0.5 - 3.0)
