#!/usr/bin/env python
"""Plugin mode in numbers: the reference's UNMODIFIED task class (legged_gym `Go2TS` etc., torch eager) stepping over
`B200Simulator` through the `Simulator` plugin API, next to `FusedLeggedEnv` (the whole post_physics_step in one kernel),
same task, same env count, N(0,1) actions, steady state after a pre-roll.  Times K `env.step` calls back to back with one
CUDA-event pair (plugin mode is launch bound: hundreds of small torch kernels per step, so per-step events would only
measure the host).

    python tools/bench_plugin.py [--task go2_ts] [--envs 4096] [--steps 100]

The reference tree is the staged copy baseline/_ref (tools/stage_reference.py) or /root/reference.  legged_gym imports
trimesh / matplotlib / pygame / xlsxwriter at module level without using them on this path; empty placeholder modules
stand in when they are not installed."""
import argparse
import json
import os
import statistics
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_root():
    for p in (os.environ.get("HCR_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "legged_gym")):
            return p
    raise SystemExit("no reference tree (baseline/_ref or /root/reference)")


def timed(step, pool, K, reps=3):
    runs = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(pool[i % len(pool)])
        e1.record()
        torch.cuda.synchronize()
        runs.append(e0.elapsed_time(e1) / K)
    return statistics.median(runs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="go2_ts")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--pre-roll", type=int, default=100)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    import hcr_genesis_lr_cl_b200.plugin as b200
    b200.stub_missing_optional_imports()
    backend = b200.install(reference_root=reference_root())
    from legged_gym.utils.task_registry import task_registry
    ns = types.SimpleNamespace(task=args.task, headless=True, cpu=False, num_envs=args.envs, debug=False, max_iterations=None, resume=False,
                               sync_wandb=False, ckpt=-1, load_run=None, export_onnx=False, use_joystick=False, joystick_type=None,
                               follow_robot=False)
    env_cfg, _ = task_registry.get_cfgs(name=args.task)
    env, env_cfg = task_registry.make_env(args.task, args=ns, env_cfg=env_cfg)
    assert isinstance(env.simulator, backend)
    g = torch.Generator(device=dev); g.manual_seed(1)
    pool = [torch.randn(args.envs, env.num_actions, device=dev, generator=g) for _ in range(16)]
    for i in range(args.pre_roll):
        env.step(pool[i % 16])
    l0 = env.simulator.launch_count
    t_plugin = timed(env.step, pool, args.steps)
    own_launches = (env.simulator.launch_count - l0) / (3 * args.steps)

    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    from hcr_genesis_lr_cl_b200.task_spec import TaskSpec
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    spec = TaskSpec.from_reference_cfg(env_cfg, args.task)
    fused = FusedLeggedEnv(spec, args.envs, dev, terrain=terrain_for(spec), cfg=env_cfg)
    fused.reset()
    for i in range(args.pre_roll):
        fused.step(pool[i % 16])
    t_fused = timed(fused.step, pool, args.steps)
    dec = int(env_cfg.control.decimation)
    print(json.dumps({
        "task": args.task, "envs": args.envs, "steps": args.steps, "pre_roll": args.pre_roll,
        "plugin": {"what": f"reference {type(env).__name__}.step (torch eager, unmodified) over B200Simulator", "ms_per_step": t_plugin,
                   "env_substeps_per_s": args.envs * dec / (t_plugin * 1e-3), "backend_launches_per_step": own_launches},
        "fused": {"what": "FusedLeggedEnv.step (b200_env_step)", "ms_per_step": t_fused, "env_substeps_per_s": args.envs * dec / (t_fused * 1e-3)},
        "fused_over_plugin": t_plugin / t_fused}))


if __name__ == "__main__":
    main()
