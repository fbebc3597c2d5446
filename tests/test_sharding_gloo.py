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
        kl = parallel.allreduce_mean(torch.tensor([float(rank + 1)]))
        data = torch.arange(10.0) + 10 * rank
        mean, std = parallel.global_mean_std(data)
        full = torch.cat([torch.arange(10.0) + 10 * r for r in range(world)])
        ok_stats = torch.allclose(mean, full.mean()) and torch.allclose(std, full.std() + 1e-8, atol=1e-5)
        stats = torch.tensor([2.0 * (rank + 1), 4.0 * (rank + 1), float(rank + 1)])     # two sums + reset count
        ep = parallel.reduce_episode_stats(stats, 2, 20.0, ["a", "b"])
        ok_ep = abs(float(ep["rew_a"]) - (2 + 4) / (3 * 20.0)) < 1e-6 and abs(float(ep["rew_b"]) - (4 + 8) / (3 * 20.0)) < 1e-6
        q.put((rank, ok_grad, abs(float(kl) - 1.5) < 1e-6, ok_stats, ok_ep))
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
