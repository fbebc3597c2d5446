#!/usr/bin/env python
"""BASELINE config C5: `go2_cat` throughput sweep, 8192 .. 65536 envs per job, with a PPO-shaped learner step whose
gradients are all-reduced over NCCL (`parallel.allreduce_gradients`), the way rsl_rl's PPO would be wired.

    python tools/bench_cat_sweep.py --envs 8192 16384 32768 65536                      # one GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cat_sweep.py --envs 65536   # envs = whole job

One iteration = 24 policy steps of collection (actor MLP in PyTorch -> FusedLeggedEnv.step, no host sync) + 5 epochs x 4
minibatches of a clipped-surrogate update of a 45->512->256->128->12 actor and a 99->512->256->128->1 critic (random
init; the point is the shape of the work, not the learning).  Prints env-substeps/s of the collection alone and of the
whole iteration, and the gradient all-reduce timed with CUDA events around the collective itself (one flat bucket whose
views ARE the .grad tensors, fired by the last gradient hook inside backward(): `parallel.GradientBucket`).  The policy MLP
stays PyTorch (north_star)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import build, parallel, task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402


def mlp(i, o):
    return torch.nn.Sequential(torch.nn.Linear(i, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(),
                               torch.nn.Linear(256, 128), torch.nn.ELU(), torch.nn.Linear(128, o))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[8192, 16384, 32768, 65536], help="envs of the whole job")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--horizon", type=int, default=24)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    build.build()
    spec = T.go2_cat_spec()
    terrain = terrain_for(spec)
    for total in args.envs:
        N = total // world
        env = FusedLeggedEnv(spec, N, dev, terrain=terrain, env_offset=rank * N, num_envs_global=total)
        obs, priv = env.reset()[:2]
        torch.manual_seed(0)
        actor, critic = mlp(obs.shape[1], spec.num_actions).to(dev), mlp(priv.shape[1], 1).to(dev)
        params = list(actor.parameters()) + list(critic.parameters())
        opt = torch.optim.Adam(params, lr=1e-3)
        bucket = parallel.GradientBucket(params).attach(opt)       # edit-free wiring of rsl_rl's PPO.update
        bucket.events = []
        H = args.horizon
        buf_o = torch.zeros(H, N, obs.shape[1], device=dev); buf_p = torch.zeros(H, N, priv.shape[1], device=dev)
        buf_a = torch.zeros(H, N, spec.num_actions, device=dev); buf_r = torch.zeros(H, N, device=dev)
        t_col = t_all = t_ar = 0.0
        for it in range(args.iters + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            with torch.inference_mode():
                for t in range(H):
                    a = actor(obs) + torch.randn(N, spec.num_actions, device=dev)
                    buf_o[t], buf_p[t], buf_a[t] = obs, priv, a
                    out = env.step(a)
                    obs, priv, buf_r[t] = out[0], out[1], out[4]
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            adv = (buf_r - buf_r.mean()).reshape(-1)
            mean, std = parallel.global_mean_std(adv)
            adv = (adv - mean) / std
            fo, fp, fa = buf_o.reshape(H * N, -1), buf_p.reshape(H * N, -1), buf_a.reshape(H * N, -1)
            ar = 0.0
            for epoch in range(5):
                perm = torch.randperm(H * N, device=dev).chunk(4)
                for idx in perm:
                    logp = -0.5 * ((fa[idx] - actor(fo[idx])) ** 2).sum(-1)
                    ratio = torch.exp(logp - logp.detach())
                    loss = -torch.min(ratio * adv[idx], ratio.clamp(0.8, 1.2) * adv[idx]).mean() + (critic(fp[idx]).squeeze(-1) - adv[idx]).pow(2).mean()
                    opt.zero_grad()
                    loss.backward()                                # the bucket's all-reduce fires inside, before the clip (ppo.py:136-137)
                    torch.nn.utils.clip_grad_norm_(params, 1.0)
                    opt.step()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            ar = sum(a.elapsed_time(b) for a, b in bucket.events) * 1e-3
            n_calls = max(len(bucket.events), 1)
            bucket.events = []
            if it > 0:                       # iteration 0 is warm-up
                t_col += t1 - t0; t_all += t2 - t0; t_ar += ar
        times = torch.tensor([t_col, t_all, t_ar], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        if rank == 0:
            sub = total * spec.decimation * H * args.iters
            print(json.dumps({"task": "go2_cat", "envs_total": total, "n_gpus": world, "envs_per_gpu": N,
                              "collection_env_substeps_per_s": sub / float(times[0]), "iteration_env_substeps_per_s": sub / float(times[1]),
                              "ms_per_policy_step_collection": 1e3 * float(times[0]) / (H * args.iters),
                              "allreduce_ms_per_iteration": 1e3 * float(times[2]) / args.iters,
                              "allreduce_us_per_call": 1e6 * float(times[2]) / args.iters / 20, "gradient_bucket_bytes": bucket.nbytes,
                              "allreduce_timing": "CUDA events around all_reduce + 1/world scaling on the flat bucket (device time, max over ranks)",
                              "env_kernel_variant": env.simulator.env_kernel_variant}), flush=True)
        if world > 1 and total == args.envs[-1]:
            # the collective by itself: 50 all-reduces of the gradient bucket back to back (the ranks pace each other, so no
            # arrival skew is inside the timed region, unlike the per-iteration figure above)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5):
                dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG)
            e0.record()
            for _ in range(50):
                dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"allreduce_back_to_back_us": 1e3 * float(t[0]), "bucket_bytes": bucket.nbytes, "n_gpus": world}), flush=True)
        bucket.detach()
        del env
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
