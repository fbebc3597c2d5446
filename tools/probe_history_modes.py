#!/usr/bin/env python
"""Diagnosis: dynamics / env kernel times with the frame-stack shift on the side stream, inline before the dynamics
kernel, or inside the env kernel (L2 flushed between steps like bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402

N = 4096
spec = T.go2_ts_spec()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
for mode in ("side", "side_first", "inline", "in_env_kernel", "side"):
    env = FusedLeggedEnv(spec, N, "cuda:0", terrain=terrain_for(spec))
    sim = env.simulator
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(1)
    pool = [torch.randn(N, 12, generator=g).cuda() for _ in range(8)]
    sim.fused_histories = mode == "side"
    K = 100
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(20 + K):
        flush.zero_()
        k = max(i - 20, 0)
        ev[k][0].record()
        if mode == "inline":
            sim.history_shift(side_stream=False, reorder=False)
        if mode == "side_first":                 # side stream, enqueued before the dynamics kernel: gets the SMs first
            sim.history_shift(side_stream=True)
        sim.step(pool[i % 8])
        ev[k][1].record()
        env.common_step_counter += 1
        sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
        ev[k][2].record()
    torch.cuda.synchronize()
    td = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    te = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    print(f"{mode:14s} dynamics(+inline shift) {td * 1e3:7.1f} us   env {te * 1e3:6.1f} us   total {(td + te) * 1e3:7.1f} us")
