#!/usr/bin/env python
"""Truncation error of the contact solver's sweep cap, measured with the fp64 oracle (TEST / ANALYSIS TOOLING).

Takes a steady-state population of `go2_ts` envs (N(0,1) actions, the bench workload), then advances the SAME states by
1 / 4 / 40 substeps with the sweep cap at 4 ... 60 and compares joint velocities, base velocities, base positions and
link contact forces with the run whose solver is left to converge (cap 400, tolerance 1e-4).  The numbers behind the
default `TaskSpec.pgs_iterations` -- output committed as profiles/r2_solver_cap_accuracy.txt.

    python tools/solver_accuracy.py [action_scale]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
from oracle.cpu_baseline import OracleEnv
from oracle import physics
from oracle.physics import PhysicsOracle, default_params, env_params, P_ITERS
physics.build()
spec=T.PRESETS["go2_ts"]()
terrain=terrain_for(spec)
N=512
env=OracleEnv(spec,N,terrain,precision="f64")
rng=np.random.default_rng(0)
scale=float(sys.argv[1]) if len(sys.argv)>1 else 1.0
for t in range(120): env.step(scale*rng.normal(size=(N,12)).astype(np.float32))
st=env.eo.st
f32=np.float32
def run(cap, nsub):
    ph=env.phys; ph.prm[P_ITERS]=cap
    state=np.concatenate([st["base_pos"],st["base_quat_wxyz"],st["base_lin_w"],st["base_ang_w"]],axis=1).astype(np.float64)
    q,qd=st["q"].astype(np.float64).copy(), st["qd"].astype(np.float64).copy()
    envp=env_params(spec,st); A=12
    arm=np.tile(env.model.body[1:,19][None,:],(N,1)).astype(np.float64); jp=np.concatenate([arm,np.zeros((N,2*A))],axis=1)
    warm=env.warm.copy()
    a=np.clip(act,-100,100); q0=np.asarray(spec.default_dof_pos,f32)[None,:]
    kp,kd=(st["kp_scale"]*f32(spec.kp)).astype(f32),(st["kd_scale"]*f32(spec.kd)).astype(f32)
    tgt=(a*f32(spec.action_scale)+q0).astype(f32)
    for _ in range(nsub):
        tau=(kp*(tgt-q.astype(f32))-kd*qd.astype(f32)).astype(f32)
        lf,nc=ph.substep(state,q,qd,tau,envp,jp,warm)
    return state,q,qd,lf
act=scale*rng.normal(size=(N,12)).astype(np.float32)
for nsub in (1,4,40):
    ref=run(400,nsub)
    print(f"--- {nsub} substeps; scales: |qd| rms {np.sqrt((ref[2]**2).mean()):.3f}, |base vel| rms {np.sqrt((ref[0][:,7:13]**2).mean()):.3f}, |F| per-env max mean {np.abs(ref[3]).reshape(N,-1).max(1).mean():.1f}")
    for cap in (4,6,8,12,16,30,60):
        r=run(cap,nsub)
        eq=np.abs(r[2]-ref[2]); ev=np.abs(r[0][:,7:13]-ref[0][:,7:13]); ef=np.abs(r[3]-ref[3]).reshape(N,-1).max(1)
        ep=np.abs(r[0][:,0:3]-ref[0][:,0:3])
        print(f"cap {cap:3d}: qd err mean {eq.mean():.2e} p99 {np.quantile(eq,0.99):.2e} max {eq.max():.2e} | base vel err mean {ev.mean():.2e} max {ev.max():.2e} | base pos err max {ep.max():.2e} | link force err mean {ef.mean():.3f} p99 {np.quantile(ef,.99):.2f} N")
