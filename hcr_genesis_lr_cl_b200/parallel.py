"""Multi-GPU plumbing: one process per GPU, envs sharded as contiguous blocks, no collective inside env.step.

The reference has no distributed code at all; what the 8-GPU configs need is
  * the env block of each rank (global ids key the RNG and the terrain-type assignment, so results do not depend on
    the GPU count),
  * a gradient all-reduce between ``loss.backward()`` and ``clip_grad_norm_`` (rsl_rl/algorithms/ppo.py:136-137),
  * all-reduced scalars for the KL-adaptive learning rate (ppo.py:198-206) and advantage normalisation
    (rsl_rl/storage/rollout_storage.py:137-138),
  * summed episode statistics at log time (legged_robot.py:128-141).
All of it is torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests); messages are <= a few MB, i.e.
latency bound, so one flattened bucket per call.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def env_block(rank: int, world_size: int, envs_per_rank: int) -> Tuple[int, int]:
    """(env_offset, num_envs_global) of `rank`: contiguous block, global_id = rank * envs_per_rank + i."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return rank * envs_per_rank, world_size * envs_per_rank


def terrain_types_for_block(env_offset: int, num_envs: int, num_envs_global: int, num_cols: int) -> torch.Tensor:
    """genesis_simulator.py:523-524 evaluated on global ids."""
    gid = torch.arange(num_envs) + env_offset
    return torch.div(gid, num_envs_global / num_cols, rounding_mode="floor").to(torch.long)


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(params: Iterable[torch.nn.Parameter]) -> None:
    """Average gradients over ranks with ONE flattened bucket (SUM then / world)."""
    w = _world()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(w)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def allreduce_mean(x: torch.Tensor) -> torch.Tensor:
    """Mean over ranks of a (small) tensor, e.g. the minibatch KL."""
    w = _world()
    if w == 1:
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / w


def global_mean_std(x: torch.Tensor, eps: float = 1e-8) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean / (unbiased) std of a tensor sharded over ranks (advantage normalisation)."""
    s = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), device=x.device)])
    if _world() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1).clamp(min=1.0)
    return mean, var.clamp(min=0.0).sqrt() + eps


def reduce_episode_stats(stats: torch.Tensor, n_sums: int, episode_length_s: float, names) -> Dict[str, torch.Tensor]:
    """Combine the per-rank `stats` vectors of the fused env kernel (sums over resetting envs + reset count) into the
    job-wide extras["episode"] means."""
    s = stats[:n_sums + 1].clone()
    if _world() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    cnt = s[n_sums].clamp(min=1.0)
    return {"rew_" + n: s[i] / (cnt * episode_length_s) for i, n in enumerate(names)}
