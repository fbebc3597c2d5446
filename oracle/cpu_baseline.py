"""CPU baseline: the oracle pair (numpy env half + C rigid-body substep) timed on host cores.

TEST/BENCH INFRASTRUCTURE -- used by bench.py's `cpu_baseline` leg and `--impl reference` arm and by smoke().  It is
the "port" kind of baseline: Genesis (the reference's engine) is not installable here or on the GPU box, and the
reference tree itself does not travel, so the reference's path is represented by this restatement
(BASELINE.md section 4.1).  Envs are independent, so the sample is split into one chunk per thread; the C substep
releases the GIL (ctypes), the numpy half mostly does.
"""
from __future__ import annotations

import os
import time

import numpy as np

from oracle.env_oracle import EnvOracle
from oracle.physics import PhysicsOracle, default_params, env_params

PH = dict(base_pos="base_pos", base_quat_wxyz="base_quat_wxyz", base_lin_w="base_lin_w", base_ang_w="base_ang_w", q="q", qd="qd")


class OracleEnv:
    """One chunk of envs stepped by the oracle (same step semantics as FusedLeggedEnv.step)."""

    def __init__(self, spec, num_envs, terrain=None, env_offset=0, precision="f32"):
        self.spec, self.N = spec, num_envs
        hs, origins = terrain if terrain is not None else (None, None)
        self.eo = EnvOracle(spec, num_envs, hs, origins)
        self.model = self.eo.model
        prm = default_params(dt=spec.sim_dt, iters=spec.pgs_iterations, hscale=spec.horizontal_scale, vscale=spec.vertical_scale,
                             border=spec.border_size if spec.heightfield else 0.0, terrain_mu=spec.static_friction, tol=spec.pgs_tolerance)
        self.phys = PhysicsOracle(self.model, prm, hs, precision=precision)
        self.env_offset = env_offset
        st = self.eo.st
        rng = np.random.default_rng(spec.seed + env_offset)
        if spec.heightfield:
            st["terrain_levels"][:] = rng.integers(0, spec.max_init_terrain_level + 1, num_envs)
            st["terrain_types"][:] = rng.integers(0, spec.num_cols, num_envs)
            st["env_origins"][:] = origins[st["terrain_levels"], st["terrain_types"]]
        st["base_pos"][:] = np.asarray(spec.init_pos, np.float32) + st["env_origins"]
        st["q"][:] = np.asarray(spec.default_dof_pos, np.float32)
        st["friction"][:] = 1.0
        st["episode_length"][:] = rng.integers(0, spec.max_episode_length, num_envs)
        self._feet = spec.link_groups(self.model)[0]
        self.warm = np.zeros((num_envs, 48), np.float64)

    def step(self, actions):
        spec, st, N, A = self.spec, self.eo.st, self.N, self.spec.num_actions
        f32 = np.float32
        a = self.eo.pre_step(actions)
        state = np.concatenate([st["base_pos"], st["base_quat_wxyz"], st["base_lin_w"], st["base_ang_w"]], axis=1).astype(np.float64)
        q, qd = st["q"].astype(np.float64), st["qd"].astype(np.float64)
        envp = env_params(spec, st)
        arm = np.tile(st["joint_armature"], (1, A)).astype(np.float64) if spec.randomize_joint_armature else \
            np.tile(self.model.body[1:, 19][None, :], (N, 1)).astype(np.float64)
        dmp = np.tile(st["joint_damping"], (1, A)).astype(np.float64) if spec.randomize_joint_damping else np.zeros((N, A))
        fls = np.tile(st["joint_friction"], (1, A)).astype(np.float64) if spec.randomize_joint_friction else np.zeros((N, A))
        jp = np.concatenate([arm, dmp, fls], axis=1)
        q0 = np.asarray(spec.default_dof_pos, f32)[None, :]
        kp, kd = (st["kp_scale"] * f32(spec.kp)).astype(f32), (st["kd_scale"] * f32(spec.kd)).astype(f32)
        tgt = (a * f32(spec.action_scale) + q0).astype(f32)
        tau, lf = None, None
        for _ in range(spec.decimation):
            tau = (kp * (tgt - q.astype(f32)) - kd * qd.astype(f32)).astype(f32)
            lf, _ = self.phys.substep(state, q, qd, tau, envp, jp, self.warm)
        lp, lv = self.phys.link_kinematics(state, q, qd)
        phys = dict(base_pos=state[:, 0:3], base_quat_wxyz=state[:, 3:7], base_lin_w=state[:, 7:10], base_ang_w=state[:, 10:13],
                    q=q, qd=qd, torques=tau, link_force=lf, feet_pos=lp[:, self._feet], feet_vel=lv[:, self._feet])
        o = self.eo.post_step(phys)
        self.warm[o["env_ids"]] = 0
        return o


def _worker(args):
    task_overrides, per, offset, steps, warmup, seed, barrier = args
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    task, over = task_overrides
    spec = T.PRESETS[task](**over)
    terrain = terrain_for(spec) if spec.heightfield else None      # the task's own heightfield (go2 rough / tron1 rough)
    env = OracleEnv(spec, per, terrain, env_offset=offset)
    rng = np.random.default_rng(seed + offset)
    acts = [rng.normal(size=(per, spec.num_actions)).astype(np.float32) for _ in range(steps + warmup)]
    for t in range(warmup):
        env.step(acts[t])
    barrier.wait()
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        env.step(acts[t])
    return t0, time.perf_counter()


def time_cpu_baseline(spec, terrain, total_envs: int, steps: int, threads: int | None = None, warmup: int = 1, seed: int = 0):
    """Split `total_envs` over one process per core (envs are independent; processes avoid the GIL on the numpy half).
    Returns dict(value=env-substeps/s, cores, sample, seconds)."""
    import multiprocessing as mp
    threads = threads or max(1, os.cpu_count() or 1)
    threads = max(1, min(threads, total_envs))
    per = total_envs // threads
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        barrier = mgr.Barrier(threads)
        over = {"pgs_iterations": spec.pgs_iterations, "seed": spec.seed}
        jobs = [((spec.task, over), per, i * per, steps, warmup, seed, barrier) for i in range(threads)]
        with ctx.Pool(threads) as pool:
            spans = pool.map(_worker, jobs)
    dt = max(e for _, e in spans) - min(s for s, _ in spans)
    n = per * threads
    return dict(value=n * spec.decimation * steps / dt, cores=threads, seconds=dt, ms_per_step=1e3 * dt / steps,
                sample=f"{n} envs x {steps} policy steps ({spec.task}, oracle port: numpy env half + C rigid-body substep, "
                       f"fp32, one process per core)")
