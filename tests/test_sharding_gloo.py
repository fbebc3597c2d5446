"""world_size-2 gloo tests of the multi-GPU host logic (env blocks, gradient / statistic reductions)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hcr_genesis_lr_cl_b200 import parallel


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Linear(8, 3)
        x = torch.arange(16.0).reshape(2, 8) + rank
        model(x).sum().backward()
        parallel.allreduce_gradients(model.parameters())
        g = model.weight.grad.clone()
        # reference: average of the two per-rank gradients
        expect = sum((torch.arange(16.0).reshape(2, 8) + r).sum(0) for r in range(world)) / world
        ok_grad = torch.allclose(g, expect.expand(3, 8))
        # the edit-free PPO wiring: flat bucket, all-reduce fired by the last gradient hook inside backward(), zero_grad kept in place
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        bucket = parallel.GradientBucket(net.parameters()).attach(opt)
        ok_bucket = True
        for it in range(3):
            opt.zero_grad()                                   # rsl_rl calls it with the default set_to_none=True
            xb = torch.arange(16.0).reshape(2, 8) * 0.01 + rank + it
            net(xb).pow(2).sum().backward()                   # averaged over the ranks when this returns
            ok_bucket &= all(p.grad._base is bucket.flat for p in net.parameters()) and bucket.calls == it + 1
            ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 3))
            ref.load_state_dict(net.state_dict())
            acc = [torch.zeros_like(p) for p in ref.parameters()]
            for r in range(world):
                ref.zero_grad()
                ref(torch.arange(16.0).reshape(2, 8) * 0.01 + r + it).pow(2).sum().backward()
                acc = [a + p.grad / world for a, p in zip(acc, ref.parameters())]
            ok_bucket &= all(torch.allclose(p.grad, a, atol=1e-6) for p, a in zip(net.parameters(), acc))
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
            opt.step()
        kl = parallel.allreduce_mean(torch.tensor([float(rank + 1)]))
        data = torch.arange(10.0) + 10 * rank
        mean, std = parallel.global_mean_std(data)
        full = torch.cat([torch.arange(10.0) + 10 * r for r in range(world)])
        ok_stats = torch.allclose(mean, full.mean()) and torch.allclose(std, full.std() + 1e-8, atol=1e-5)
        stats = torch.tensor([2.0 * (rank + 1), 4.0 * (rank + 1), float(rank + 1)])     # two sums + reset count
        ep = parallel.reduce_episode_stats(stats, 2, 20.0, ["a", "b"])
        ok_ep = abs(float(ep["rew_a"]) - (2 + 4) / (3 * 20.0)) < 1e-6 and abs(float(ep["rew_b"]) - (4 + 8) / (3 * 20.0)) < 1e-6
        q.put((rank, ok_grad, abs(float(kl) - 1.5) < 1e-6, ok_stats, ok_ep, bool(ok_bucket)))
    finally:
        dist.destroy_process_group()


def test_two_rank_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(all(r[1:]) for r in res), res


def test_env_blocks_and_terrain_types():
    assert parallel.env_block(0, 8, 2048) == (0, 16384) and parallel.env_block(3, 8, 2048) == (6144, 16384)
    with pytest.raises(ValueError):
        parallel.env_block(8, 8, 2048)
    full = parallel.terrain_types_for_block(0, 16384, 16384, 10)
    parts = torch.cat([parallel.terrain_types_for_block(*((r * 2048,) + (2048, 16384, 10))) for r in range(8)])
    assert torch.equal(full, parts)                       # sharding keeps the reference's terrain-type distribution
    assert full.min() == 0 and full.max() == 9
