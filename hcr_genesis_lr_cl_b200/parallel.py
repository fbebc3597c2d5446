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


class GradientBucket:
    """All gradients of a model in ONE pre-allocated flat fp32 buffer: every ``p.grad`` is a view of it, so averaging
    over the ranks is one collective on one tensor -- no ``torch.cat``, no copy back (round 1 spent 0.34-0.40 ms per call
    on exactly those at 2 GPUs for a 1.3 MB bucket; the NCCL all-reduce itself is a few tens of microseconds over NVLink).

    ``attach(optimizer)`` makes the wiring edit-free for rsl_rl's PPO (rsl_rl/algorithms/ppo.py:124-148, which calls
    ``optimizer.zero_grad(); loss.backward(); clip_grad_norm_(...); optimizer.step()``):
      * ``optimizer.zero_grad`` is rebound to zero the flat buffer in place (the default ``set_to_none=True`` would drop the
        views),
      * a post-accumulate-grad hook on every parameter launches the all-reduce as soon as the LAST gradient of a backward
        pass has landed, i.e. inside ``loss.backward()`` and therefore before ``clip_grad_norm_`` -- the place
        INTEGRATION.md names (ppo.py:136-137)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, n = self.params[0].device, sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradientBucket expects fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._pending = 0
        self._handles = []
        self.calls = 0
        self.events = None          # optional list of (start, end) CUDA events per all-reduce (profiling)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self, set_to_none: bool = False) -> None:
        self.flat.zero_()
        self._pending = 0

    def allreduce(self) -> None:
        """SUM over ranks, then / world: the average gradient, in place in every ``p.grad``."""
        w = _world()
        self.calls += 1
        if w == 1:
            return
        ev = None
        if self.events is not None and self.flat.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)        # ncclAvg: the 1 / world scaling inside the collective kernel
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / w)
        if ev is not None:
            ev[1].record()
            self.events.append(ev)

    def _hook(self, p) -> None:
        if p.grad is not None and p.grad.data_ptr() != self._view_ptr[id(p)]:
            # something replaced the view (e.g. a zero_grad(set_to_none=True) we did not intercept): fold it back in
            off, n = self._slot[id(p)]
            self.flat[off:off + n].copy_(p.grad.reshape(-1))
            p.grad = self.flat[off:off + n].view_as(p)
        self._pending += 1
        if self._pending == len(self.params):
            self._pending = 0
            self.allreduce()

    def attach(self, optimizer: "torch.optim.Optimizer | None" = None) -> "GradientBucket":
        self._slot, self._view_ptr, off = {}, {}, 0
        for p in self.params:
            self._slot[id(p)] = (off, p.numel())
            self._view_ptr[id(p)] = p.grad.data_ptr()
            off += p.numel()
            self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        if optimizer is not None:
            optimizer.zero_grad = self.zero
        return self

    def detach(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def allreduce_gradients(params: Iterable[torch.nn.Parameter]) -> None:
    """Average gradients over ranks, one collective.  With a ``GradientBucket`` (every ``.grad`` a view of one flat buffer)
    this is a single in-place all-reduce; for loose gradients it falls back to flatten / all-reduce / copy back."""
    w = _world()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    base = grads[0]._base if grads[0]._base is not None else None
    if base is not None and all(g._base is base for g in grads) and sum(g.numel() for g in grads) == base.numel():
        dist.all_reduce(base, op=dist.ReduceOp.SUM)
        base.mul_(1.0 / w)
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
