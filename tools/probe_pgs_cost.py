#!/usr/bin/env python
"""Diagnosis: dynamics kernel time vs the PGS sweep cap (fixed cost of a substep vs cost per sweep of the slowest env)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402

N = 4096
for iters in (1, 5, 10, 20, 30, 60):
    spec = T.go2_ts_spec(pgs_iterations=iters)
    env = FusedLeggedEnv(spec, N, "cuda:0", terrain=terrain_for(spec))
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(1)
    pool = [torch.randn(N, 12, generator=g).cuda() for _ in range(8)]
    for i in range(20):
        env.step(pool[i % 8])
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(50)]
    for i in range(50):
        ev[i][0].record()
        env.simulator.step(pool[i % 8])
        ev[i][1].record()
        env.common_step_counter += 1
        env.simulator.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
    torch.cuda.synchronize()
    t = sum(a.elapsed_time(b) for a, b in ev) / 50
    print(f"pgs_iterations={iters:3d}  dynamics kernel {t * 1e3:7.1f} us")
