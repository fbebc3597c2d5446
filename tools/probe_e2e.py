#!/usr/bin/env python
"""Where the host boundary's time goes: K back-to-back steps (no L2 flush, nothing synchronised in between) of
  dev        FusedLeggedEnv.step with device-resident actions, nothing copied
  host_in    step_host with pinned actions, no result buffers
  host_io    step_host with pinned actions and the pinned rew | reset | time_out slab (the bench's e2e)
each timed with one CUDA-event pair around the K steps, after the bench's pre-roll."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402


def main():
    task, N, K = "go2_ts", 4096, 200
    spec = T.PRESETS[task]()
    torch.cuda.set_device(0)
    env = FusedLeggedEnv(spec, N, torch.device('cuda:0'), terrain=terrain_for(spec))
    env.reset()
    dev = env.device
    g = torch.Generator(device=dev); g.manual_seed(1)
    pool = [torch.randn(N, env.num_actions, device=dev, generator=g) for _ in range(16)]
    host_pool = [p.cpu().pin_memory() for p in pool]
    for i in range(300):
        env.step(pool[i % 16])
    rew, rst, tmo = env.simulator.make_host_step_buffers()
    out = {}
    for name, fn in (("dev", lambda i: env.step(pool[i % 16])),
                     ("host_in", lambda i: env.step_host(host_pool[i % 16])),
                     ("host_io", lambda i: env.step_host(host_pool[i % 16], rew, rst, tmo))):
        runs = []
        for _ in range(5):
            for i in range(10):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            runs.append(e0.elapsed_time(e1) / K)
        out[name] = statistics.median(runs)
    # the dynamics kernel alone: device-resident actions vs pinned host actions read in place (UVA: the pinned pointer is a
    # valid device address), K launches back to back
    sim = env.simulator
    def dyn(ptr):
        sim._ck(sim._lib.b200_dynamics_step(sim._handle, ptr, sim._stream()))
    for name, bufs in (("dyn_dev", pool), ("dyn_host", host_pool), ("dyn_dev2", pool), ("dyn_host2", host_pool)):
        runs = []
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                dyn(bufs[i % 16].data_ptr())
            e1.record()
            torch.cuda.synchronize()
            runs.append(e0.elapsed_time(e1) / K)
        out[name] = statistics.median(runs)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
