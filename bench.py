#!/usr/bin/env python
"""bench.py -- Go2 rough-terrain env-steps/s (physics substeps x envs), 4096 envs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one policy step of BASELINE config C2 (`go2_ts`, heightfield curriculum, height-scan obs) = one
`FusedLeggedEnv.step` = one C-ABI call `b200_env_step`: the decimated PD + rigid-body kernel and the fused
post_physics_step kernel.  Envs shard as contiguous blocks, one process per
GPU, no collective on the data path (weak scaling).  The line printed by rank 0 carries

  value        whole-job env-substeps/s, inputs resident in HBM: K steps timed one by one with CUDA events (L2 flushed between
               them), MEDIAN step time per rank, max over ranks (the mean is printed beside it: one host hiccup in a 4 ms
               window must not decide a scaling number); the workload is rolled to its steady state un-timed first
               (`config.pre_roll_steps`, `workload_state`)
  e2e          the same metric through FusedLeggedEnv.step_host() = b200_env_step with HOST buffers: the actions come out of
               pinned host memory (a copy kernel reading the mapped buffer over PCIe, overlapped with the dynamics kernel's
               first phase) and rewards / resets / time-outs go back into pinned host memory (written by the env kernel's
               last CTA) inside the timed region, every step; K steps enqueued back to back, three blocks, median block
  roofline     the whole step (all kernels' algorithmic bytes over the step time) against the measured HBM copy bandwidth
               (MEASURED_PEAKS.json), then each kernel on its own bytes and time (timed call by call in a second loop)
  kernels      the step, and its kernels timed call by call: average duration, share of the step, static resources (the
               dynamics kernel is latency / issue bound)
  cpu_baseline the oracle port timed on host cores on a bounded sample (rank 0, N=1 only)

`--impl reference` times the CPU restatement of the reference path (oracle port; the reference's engine cannot be
installed) with all host threads and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (physics substeps x envs), Go2 rough terrain, 4096 envs/GPU"
UNIT = "env-substeps/s"
ENVS_PER_GPU = 4096
TASK = "go2_ts"
WORKLOADS = {   # BASELINE.json configs; the metric is quoted on go2_ts (config 2), the others are selectable with --task
    "go2": "Go2 flat-terrain velocity tracking (base go2 task)",
    "go2_ts": "Go2 rough-terrain heightfield curriculum with height-scan obs",
    "go2_cat": "Go2 CaT constraints-as-terminations, rough terrain",
    "go2_wtw": "Go2 Walk-These-Ways with periodic gait rewards",
    "tron1_pf": "TRON1_PF point-foot biped rough terrain with domain randomisation",
    "tron1_pf_ee": "TRON1_PF biped with explicit estimator, periodic gait, rough terrain",
    "go2_cts": "Go2 concurrent teacher-student, rough terrain",
    "go2_ee": "Go2 explicit estimator, rough terrain",
    "go2_dreamwaq": "Go2 DreamWaQ (labels + next-state decoder target), rough terrain",
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _config(args, spec, N, world):
    """`config` of the JSON line: the workload both arms run (the reference arm steps the same env population on the host)."""
    return {"workload": f"{args.task}: {WORKLOADS[args.task]}, {N} envs/GPU",
            "envs_per_gpu": N, "decimation": spec.decimation, "pgs_iterations": spec.pgs_iterations,
            "actions": "N(0,1) (policy at init)", "pre_roll_steps": args.pre_roll,
            "timing": "median of K per-step CUDA-event times per rank, max over ranks",
            "l2": "flushed between timed steps (256 MB write, then 256 MB read: no dirty lines left)",
            "parallelism": f"env-sharded x{world}, no data-path collective"}


def run_reference(args):
    """CPU arm: the oracle port on all host threads (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    from oracle.cpu_baseline import time_cpu_baseline
    from oracle import physics
    physics.build()
    spec = T.PRESETS[args.task]()
    cores = os.cpu_count() or 1
    sample_envs = (args.envs // cores) * cores or args.envs       # the arm's own config: every step is the whole 4096-env batch, split over the cores
    res = time_cpu_baseline(spec, terrain_for(spec), sample_envs, steps=max(args.steps, 1), threads=cores, warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": _config(args, spec, args.envs, args.gpus),       # the b200 arm's config, key for key (timing / l2 describe that arm)
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": res["sample"] + f"; each step = {sample_envs} envs stepped once, {sample_envs // cores} per core on {cores} cores, "
                                                   "host wall clock, no pre-roll"},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hcr_genesis_lr_cl_b200 import accounting, build, task_spec as T
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    build.build()
    spec = T.PRESETS[args.task](**({"pgs_iterations": args.pgs_iters} if args.pgs_iters else {}))
    terrain = terrain_for(spec)
    N = args.envs
    env = FusedLeggedEnv(spec, N, dev, terrain=terrain, env_offset=rank * N, num_envs_global=world * N)
    sim = env.simulator
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    env.episode_length_buf = torch.randint(0, int(env.max_episode_length), (N,), generator=g).to(dev)  # on_policy_runner.py:169
    pool = [torch.randn(N, spec.num_actions, generator=g).to(dev) for _ in range(16)]          # policy at init: N(0,1) actions
    host_pool = [p.cpu().pin_memory() for p in pool]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)                    # > 126 MB L2
    flush_rd = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)                # 256 MB, only ever read
    sink = torch.zeros((), dtype=torch.float32, device=dev)

    def flush_l2():
        """Evict the step's working set: write 256 MB, then read another 256 MB so that the L2 is left full of CLEAN
        lines (after the write alone it is full of dirty zeros whose write-back competes with the timed kernels' reads:
        a cold streaming copy then measures ~half of the HBM bandwidth, profiles/r1v_sweep_history_shift.log)."""
        flush.zero_()
        sink.copy_(flush_rd.sum())

    K, W = args.steps, args.warmup
    # ---- steady state first (un-timed): with N(0,1) actions the population of fallen / resetting robots and the contact
    # solver's work keep growing for the first few hundred policy steps after reset (0.21 ms at steps 5-25 vs 0.24 ms at
    # steps 20-220 in round 1); every number below is taken on the stationary workload
    for i in range(args.pre_roll):
        env.step(pool[i % 16])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing of the step as shipped (env.step = one b200_env_step call): CUDA events on the launching (current) stream, L2 flushed between steps
    evf = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    for i in range(W):
        env.step(pool[i % 16])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count
    for i in range(K):
        flush_l2()
        evf[i][0].record()
        env.step(pool[(W + i) % 16])
        evf[i][1].record()
    barrier()
    per_step = [e[0].elapsed_time(e[1]) for e in evf]
    t_step, t_step_mean = statistics.median(per_step), sum(per_step) / K
    total_ms = t_step * K
    # what the timed steps worked on (device reads after the timed region)
    b = sim._buf
    in_contact = (b["link_contact_forces"].norm(dim=-1) > 0).sum(1).float()
    workload_state = {
        "policy_steps_since_reset": args.pre_roll + W + K,
        "links_in_contact_per_env": float(in_contact.mean()),
        "fraction_fallen": float((b["projected_gravity"][:, 2] > spec.max_projected_gravity).float().mean()),
        "resets_last_step": float(b["stats"][len(env.sum_names)]),
        "solver_work_per_env": float(b["dyn_cost"].float().mean()),       # sum over the 4 substeps of (sweeps + 4) x constraint rows
        "mean_terrain_level": float(b["terrain_levels"].float().mean()) if spec.heightfield else 0.0,
    }
    # ---- the same step call by call (dynamics kernel | env kernel): per-kernel times that
    # explain the fused number and give the env kernel's own HBM figure
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(K):
        flush_l2()
        a = pool[(W + i) % 16]
        ev[i][0].record()
        sim.step(a)
        ev[i][1].record()
        env.common_step_counter += 1
        env._apply_pending_curriculum()
        env._set_step_flags()
        sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
        ev[i][2].record()
        env._fill_extras()
    barrier()
    launches = sim.launch_count - launches0
    t_dyn = statistics.median(e[0].elapsed_time(e[1]) for e in ev)
    t_env = statistics.median(e[1].elapsed_time(e[2]) for e in ev)
    # ---- end to end through the public API with host buffers
    rew_host, rst_host, tmo_host = sim.make_host_step_buffers()       # one pinned slab: rew f32[N] | reset u8[N] | time_out u8[N]
    # FusedLeggedEnv.step_host = ONE C-ABI call (b200_env_step): pinned actions in (staged by a copy kernel), the kernels, the
    # pinned rew | reset | time_out slab out (written by the env kernel)
    for i in range(W):
        env.step_host(host_pool[i % 16], rew_host, rst_host, tmo_host)
    barrier()
    e2e_runs = []
    for _ in range(3):                       # three back-to-back blocks of K steps, the median block counts
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            env.step_host(host_pool[(W + i) % 16], rew_host, rst_host, tmo_host)
        e1.record()
        barrier()
        e2e_runs.append(e0.elapsed_time(e1))
    e2e_ms = statistics.median(e2e_runs)
    clocks = sampler.stop() if rank == 0 else None
    times = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(times[0]), float(times[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_total = N * world
    value = n_total * spec.decimation * K / (total_ms * 1e-3)
    e2e = n_total * spec.decimation * K / (e2e_ms * 1e-3)
    model = sim._model
    peak, peak_src = _peaks()
    env_bytes = accounting.env_kernel_bytes(spec, model) * N
    dyn_bytes = accounting.dynamics_kernel_bytes(spec, model) * N
    # the reference's post_physics_step = the env kernel (the frame stacks are double-written rings: nothing is moved)
    achieved = env_bytes / (t_env * 1e-3) / 1e9
    step_bytes = env_bytes + dyn_bytes
    whole = step_bytes / (total_ms / K * 1e-3) / 1e9
    on_device = 4 * N * sum(env.widths[k] for k in ("obs", "priv", "hist", "critic"))     # observation tensors the policy reads in HBM
    traffic, tj_all = None, {}              # DRAM bytes per launch of the same kernels from the committed ncu capture
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("workload") == f"{args.task}/{N}":
            traffic, tj_all = tj["env_post_step_kernel"], tj
    ki_env, ki_dyn = sim.kernel_info("env"), sim.kernel_info("dynamics")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": _config(args, spec, N, world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": N * spec.num_actions * 4 * world, "d2h_bytes_per_step": N * 6 * world,
                "ms_per_step": e2e_ms / K, "result_on_device_bytes": on_device * world,
                "note": "actions are read from pinned host memory and rewards / resets / time-outs written back into pinned host memory every step (by the kernels, no copy-engine operation); the observation tensors stay in HBM, where the policy network reads them"},
        "ms_per_step_mean": t_step_mean,
        "workload_state": workload_state,
        "gpu_launches": int(launches),
        "clocks": clocks,
        # the path's figure first (all kernels' algorithmic bytes over the step time: the recipe's number for the dominant part of
        # the step), then each kernel on its own bytes and time
        "roofline": {"kernel": "whole env.step: dynamics_step_kernel (86 % of the time) + env_post_step_kernel", "bound": "hbm",
                     "achieved": whole, "peak": peak, "unit": "GB/s", "frac": whole / peak,
                     "traffic": (tj_all.get("dynamics_step_kernel", 0) + tj_all.get("env_post_step_kernel", 0)) if traffic is not None else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": step_bytes, "avg_launch_ms": total_ms / K,
                     "bytes_variant": "ring: frame stacks are double-written rings handed out as strided views, a step writes each new "
                                      "frame twice and moves nothing (SURVEY 8d names this variant: 5.9 KB instead of 17.5 KB per env "
                                      "and policy step for go2_ts)",
                     "frac_with_round1_bytes": (step_bytes + accounting.shifted_stack_bytes(spec, model) * N) / (total_ms / K * 1e-3) / 1e9 / peak,
                     "dynamics_step_kernel": {"algorithmic_bytes_per_launch": dyn_bytes, "avg_launch_ms": t_dyn,
                                              "achieved": dyn_bytes / (t_dyn * 1e-3) / 1e9, "frac": dyn_bytes / (t_dyn * 1e-3) / 1e9 / peak,
                                              "bound": "latency / issue (SURVEY 8d: 1.3 KB of HBM traffic per env against ~25 k "
                                                       "dependent instructions): see profiles/*_kernels_ncu_summary.txt for issue-slot use"},
                     "env_post_step_kernel": {"algorithmic_bytes_per_launch": env_bytes, "avg_launch_ms": t_env, "achieved": achieved,
                                              "frac": achieved / peak, "traffic": traffic,
                                              "bound": "instruction latency at 28 warps per SM: its working set (19 MB at 4096 envs) "
                                                       "is L2 resident"},
                     "note": "no kernel of this path is HBM-bound at 4096 envs per GPU (one wave of 28 one-env warps per SM): "
                             "time per step, not this fraction, is what to read"},
        "kernels": {
            "step": {"avg_ms": t_step, "note": "one b200_env_step call: dynamics kernel (the env-ordering kernel on a side stream under it), env "
                     "kernel, stats finalize; the entries below are a second loop that times the kernels call by call"},
            "dynamics_step_kernel": {"avg_ms": t_dyn, "share": t_dyn / (t_dyn + t_env), "bound": "latency/issue", **ki_dyn,
                                     "algorithmic_bytes_per_launch": dyn_bytes, "hbm_gbs": dyn_bytes / (t_dyn * 1e-3) / 1e9},
            "env_post_step_kernel": {"variant": sim.env_kernel_variant, "avg_ms": t_env, "share": t_env / (t_dyn + t_env), "bound": "latency/issue", **ki_env,
                                     "algorithmic_bytes_per_launch": env_bytes, "hbm_gbs": env_bytes / (t_env * 1e-3) / 1e9},
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import physics
        from oracle.cpu_baseline import time_cpu_baseline
        physics.build()
        cores = os.cpu_count() or 1
        res = time_cpu_baseline(spec, terrain, (N // cores) * cores or N, steps=20, threads=cores, warmup=2)
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--task", default=TASK, choices=sorted(WORKLOADS), help="task preset (default: the metric's config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pgs-iters", type=int, default=0, help="override the contact solver's sweep cap (tuning experiments)")
    ap.add_argument("--pre-roll", type=int, default=300, help="un-timed policy steps before the warm-up (steady-state workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)          # K and W as given: a step is a bounded sample (256 envs per core), ~15 ms
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
