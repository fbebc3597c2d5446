#!/usr/bin/env python
"""Diagnosis: per-env completion times inside one dynamics_step_kernel launch (build with -DDYN_TIMING).
Shows how much of the launch is the tail of a few slow envs: B200_NVCC_EXTRA=-DDYN_TIMING python tools/probe_dyn_timing.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200_NVCC_EXTRA", "-DDYN_TIMING")
from hcr_genesis_lr_cl_b200 import build, task_spec as T  # noqa: E402
build.build(force=True)
from hcr_genesis_lr_cl_b200 import _cabi  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402

N = 4096
spec = T.go2_ts_spec()
env = FusedLeggedEnv(spec, N, "cuda:0", terrain=terrain_for(spec))
env.reset()
g = torch.Generator(device="cpu").manual_seed(100)
env.episode_length_buf = torch.randint(0, int(env.max_episode_length), (N,), generator=g).cuda()
pool = [torch.randn(N, 12, generator=g).cuda() for _ in range(16)]
lib = _cabi.load_library()
lib.b200_debug_dyn_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((N, 4), np.uint64)
for i in range(120):
    env.step(pool[i % 16])
    if i in (30, 60, 119):
        lib.b200_debug_dyn_timing(buf.ctypes.data_as(ctypes.c_void_p), N)
        t0 = buf[:, 0].min()
        start, end = (buf[:, 0] - t0).astype(np.float64) / 1e3, (buf[:, 1] - t0).astype(np.float64) / 1e3
        dur = end - start
        sw, rows = buf[:, 2].astype(np.float64), buf[:, 3].astype(np.float64)
        print(f"step {i}: kernel span {end.max():.1f} us; env start p50 {np.median(start):.1f} max {start.max():.1f}; "
              f"env duration mean {dur.mean():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}")
        print(f"   end-time percentiles us: " + " ".join(f"p{q}={np.percentile(end, q):.0f}" for q in (10, 50, 90, 99, 99.9, 100)))
        print(f"   sweeps/env-step mean {sw.mean():.1f} max {sw.max():.0f}; sweep-rows mean {rows.mean():.0f} p99 {np.percentile(rows, 99):.0f} max {rows.max():.0f}")
        c = np.corrcoef(rows, dur)[0, 1]
        A = np.vstack([rows, np.ones(N)]).T
        k, b = np.linalg.lstsq(A, dur, rcond=None)[0]
        print(f"   duration ~ {b:.1f} us + {k * 1e3:.2f} ns x sweep-rows (corr {c:.2f})")
os.environ.pop("B200_NVCC_EXTRA")
build.build(force=True)
