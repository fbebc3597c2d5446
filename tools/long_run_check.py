#!/usr/bin/env python
"""Stability soak (B200): 3000 policy steps of 4096 envs under N(0,1) actions with bursts of 3-sigma actions, for a quadruped, the
biped and the walk-these-ways task; reports the non-finite containment counter (must stay 0: the guard is for failures, not a
crutch), whether every state / observation is finite, and extreme joint velocities / base heights.
    python tools/long_run_check.py > profiles/<tag>_long_run_check.json"""
import os
import sys, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import task_spec as T
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
out={}
for task in ("go2_ts","tron1_pf_ee","go2_wtw"):
    spec=T.PRESETS[task](); torch.cuda.set_device(0)
    env=FusedLeggedEnv(spec,4096,torch.device("cuda:0"),terrain=terrain_for(spec)); env.reset()
    g=torch.Generator(device="cuda"); g.manual_seed(3)
    rsum=0.0; resets=0
    for i in range(3000):
        a=torch.randn(4096,env.num_actions,device="cuda",generator=g)*(3.0 if i%500>450 else 1.0)   # bursts of large actions
        o=env.step(a)
        if i%100==99:
            rsum+=float(env.rew_buf.mean()); resets+=int(env.reset_buf.sum())
    b=env.simulator._buf
    out[task]=dict(nonfinite_resets=int(env.simulator.nonfinite_resets), finite_state=bool(torch.isfinite(b["dof_pos"]).all() and torch.isfinite(b["base_pos"]).all() and torch.isfinite(env.obs_buf).all()),
                   max_abs_dof_vel=float(b["dof_vel"].abs().max()), max_base_height=float((b["base_pos"][:,2]-b["env_origins"][:,2]).max()), mean_rew_sampled=rsum/30, resets_sampled=resets)
print(json.dumps(out))
