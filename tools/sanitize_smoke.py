#!/usr/bin/env python
"""Small end-to-end run to put under compute-sanitizer where it is available (closed on the build pool): every preset, a few policy steps at an
env count that leaves ragged CTAs in both kernels, through b200_env_step and through the call-by-call path.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402

tasks = sys.argv[1:] or sorted(T.PRESETS)
for task in tasks:
    spec = T.PRESETS[task]()
    N = 45                                         # 7 dynamics CTAs (last ragged), 12 env CTAs (last ragged, not staged)
    env = FusedLeggedEnv(spec, N, "cuda:0", terrain=terrain_for(spec))
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(3)
    rew, rst, tmo = env.simulator.make_host_step_buffers()
    for t in range(4):
        a = torch.randn(N, spec.num_actions, generator=g)
        env.step(a.cuda())
        env.step_host(a.pin_memory(), rew, rst, tmo)
        env.step_two_kernels(a.cuda())
    env.episode_length_buf = torch.full((N,), int(env.max_episode_length), dtype=torch.int32)   # force time-outs -> resets
    env.step(torch.zeros(N, spec.num_actions).cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(env.rew_buf).all()
    print(task, "ok", env.simulator.env_kernel_variant, "resets", int(env.reset_buf.sum()))
