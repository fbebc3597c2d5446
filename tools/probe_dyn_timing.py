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
for i in range(420):
    env.step(pool[i % 16])
    if i in (419,):
        lib.b200_debug_dyn_timing(buf.ctypes.data_as(ctypes.c_void_p), N)
        t0 = buf[:, 0].min()
        start, end = (buf[:, 0] - t0).astype(np.float64) / 1e3, (buf[:, 1] - t0).astype(np.float64) / 1e3
        dur = end - start
        smid, cta = (buf[:, 2] >> np.uint64(32)).astype(np.int64), (buf[:, 3] >> np.uint64(32)).astype(np.int64)
        sw, rows = (buf[:, 2] & np.uint64(0xFFFFFFFF)).astype(np.float64), (buf[:, 3] & np.uint64(0xFFFFFFFF)).astype(np.float64)
        # where does the spread of the end times come from: the SM, the CTA, or the env itself?
        sm_mean = np.array([end[smid == m].mean() for m in np.unique(smid)])
        sm_cnt = np.array([(smid == m).sum() for m in np.unique(smid)])
        cta_max = np.array([end[cta == c].max() for c in np.unique(cta)])
        cta_spread = np.array([end[cta == c].max() - end[cta == c].min() for c in np.unique(cta)])
        sm_rows = np.array([rows[smid == m].sum() for m in np.unique(smid)])
        sm_end = np.array([end[smid == m].max() for m in np.unique(smid)])
        print(f"   SMs used {len(sm_mean)}, envs per SM min {sm_cnt.min()} max {sm_cnt.max()}; per-SM mean end: min {sm_mean.min():.0f} p50 {np.median(sm_mean):.0f} max {sm_mean.max():.0f} us; "
              f"per-SM last end: min {sm_end.min():.0f} p50 {np.median(sm_end):.0f} max {sm_end.max():.0f}")
        print(f"   within-CTA spread of end times: mean {cta_spread.mean():.1f} max {cta_spread.max():.1f} us; CTA end p50 {np.median(cta_max):.0f} max {cta_max.max():.0f}")
        print(f"   corr(per-SM sum of sweep-rows, per-SM last end) = {np.corrcoef(sm_rows, sm_end)[0, 1]:.2f}; corr(per-SM envs, last end) = {np.corrcoef(sm_cnt, sm_end)[0, 1]:.2f}")
        cta_rows_max = np.array([rows[cta == c].max() for c in np.unique(cta)])
        cta_rows_sum = np.array([rows[cta == c].sum() for c in np.unique(cta)])
        print(f"   corr(CTA max sweep-rows, CTA end) = {np.corrcoef(cta_rows_max, cta_max)[0, 1]:.2f}; corr(CTA sum sweep-rows, CTA end) = {np.corrcoef(cta_rows_sum, cta_max)[0, 1]:.2f}")
        print(f"step {i}: kernel span {end.max():.1f} us; env start p50 {np.median(start):.1f} max {start.max():.1f}; "
              f"env duration mean {dur.mean():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}")
        print(f"   end-time percentiles us: " + " ".join(f"p{q}={np.percentile(end, q):.0f}" for q in (10, 50, 90, 99, 99.9, 100)))
        print(f"   sweeps/env-step mean {sw.mean():.1f} max {sw.max():.0f}; sweep-rows mean {rows.mean():.0f} p99 {np.percentile(rows, 99):.0f} max {rows.max():.0f}")
        c = np.corrcoef(rows, dur)[0, 1]
        A = np.vstack([rows, np.ones(N)]).T
        k, b = np.linalg.lstsq(A, dur, rcond=None)[0]
        print(f"   duration ~ {b:.1f} us + {k * 1e3:.2f} ns x sweep-rows (corr {c:.2f})")
        # how the block scheduler placed the CTAs (the cost-ordered slots assume: CTA i of a round on SM i, rounds of one CTA per SM)
        ctas = np.unique(cta)
        cta_sm = np.array([smid[cta == c][0] for c in ctas])
        nsm = len(np.unique(smid))
        first = cta_sm[:nsm]
        print(f"   placement: CTAs 0..{nsm - 1} on {len(np.unique(first))} distinct SMs; second-round CTA {nsm}+k on the SM of CTA k for "
              f"{int((cta_sm[nsm:2 * nsm] == first[:len(cta_sm[nsm:2 * nsm])]).sum())} of {len(cta_sm[nsm:2 * nsm])}; first 16 SM ids: {first[:16].tolist()}; "
              f"SM ids of CTAs {nsm}..{nsm + 15}: {cta_sm[nsm:nsm + 16].tolist()}")
        sm_ids = np.unique(smid)
        order = np.argsort(sm_end)
        print("   slowest SMs (sm, envs, sum sweep-rows, last end us, its CTAs): " +
              "; ".join(f"{sm_ids[j]}, {sm_cnt[j]}, {sm_rows[j]:.0f}, {sm_end[j]:.0f}, {np.unique(cta[smid == sm_ids[j]]).tolist()}" for j in order[-4:]))
        print("   fastest full SMs: " +
              "; ".join(f"{sm_ids[j]}, {sm_cnt[j]}, {sm_rows[j]:.0f}, {sm_end[j]:.0f}, {np.unique(cta[smid == sm_ids[j]]).tolist()}" for j in [q for q in order if sm_cnt[q] == sm_cnt.max()][:4]))
        pick = [0, 1, 2, 37, 74, 110, 146, 147, 148, 149, 185, 220, 256, 290, 291, 292]
        print("   sweep-rows per CTA (index: sum over its envs, max env): " + "; ".join(f"{c}: {rows[cta == c].sum():.0f}, {rows[cta == c].max():.0f}" for c in pick if (cta == c).any()))
        print(f"   per-SM sum of sweep-rows: min {sm_rows.min():.0f} p50 {np.median(sm_rows):.0f} max {sm_rows.max():.0f}; last end = "
              f"{np.polyfit(sm_rows, sm_end, 1)[1]:.1f} us + {np.polyfit(sm_rows, sm_end, 1)[0] * 1e3:.2f} ns x (sum sweep-rows)")
os.environ.pop("B200_NVCC_EXTRA")
build.build(force=True)
