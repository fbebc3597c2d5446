"""The numpy restatement of the env half (oracle/env_oracle.py) against golden vectors recorded from the reference's
own Python classes (tools/make_golden.py): integer outputs exact, floats to fp32 round-off."""
import numpy as np
import pytest

from golden_util import load_golden, load_terrain, out_at, phys_at, spec_for
from oracle.env_oracle import EnvOracle

INTS = ("reset_buf", "time_out_buf", "episode_length", "fail_buf", "last_contacts", "terrain_levels", "action_delay")


@pytest.mark.parametrize("name", ["go2_ts_n32", "go2_n32", "go2_cat_n32", "tron1_pf_n32", "tron1_pf_ee_n32", "go2_wtw_n32", "go2_cts_n32", "go2_ee_n32",
                                  "go2_dreamwaq_n32", "go2_ts_delay_n32",
                                  # every reward term the Go2TS class defines switched on (incl. dof_*_stand_still, termination,
                                  # thigh_pos, foot_landing_vel, ...) + robots thrown off the terrain inside the window (OOB teleport)
                                  "go2_ts_allrew_oob_n32",
                                  # the "smooth" (von Mises CDF) periodic-gait indicator of go2_wtw / tron1_pf_ee
                                  "go2_wtw_smooth_n32", "tron1_pf_ee_smooth_n32"])
def test_env_oracle_reproduces_reference(name):
    g, s0 = load_golden(name)
    spec = spec_for(g)
    hs, origins = (load_terrain(spec) if spec.heightfield else (None, None))
    N, T_ = g["actions"].shape[1], g["actions"].shape[0]
    eo = EnvOracle(spec, N, hs, origins)
    eo.reproduce_r18 = True            # pin the oracle on every env of the reference run, env 0 included (R18)
    for k, v in s0.items():
        if k in eo.st:
            eo.st[k][...] = v
    eo.common_step_counter = int(s0["common_step_counter"])
    eo.cmd_range_x = [float(x) for x in s0["cmd_range_x"]]
    if "beh_ranges" in s0:             # go2_wtw curriculum state of the recorded window
        eo.beh_ranges = dict(zip(("gait_period", "base_height", "foot_clearance", "pitch"), np.asarray(s0["beh_ranges"], np.float64).tolist()))
        eo.num_gaits = int(s0["num_gaits"])
    resets = 0
    for t in range(T_):
        delay_before = eo.st["action_delay"].copy() if spec.randomize_ctrl_delay else None
        applied = eo.pre_step(g["actions"][t])
        o = eo.post_step(phys_at(g, t))
        ref = out_at(g, t)
        st = eo.st
        mine = dict(o, commands=st["commands"], episode_length=st["episode_length"], fail_buf=st["fail_buf"],
                    feet_air_time=st["feet_air_time"], last_contacts=st["last_contacts"], episode_sums=st["episode_sums"],
                    env_origins=st["env_origins"], base_lin_vel=st["base_lin_vel"], base_ang_vel=st["base_ang_vel"],
                    projected_gravity=o["projected_gravity_obs"], friction=st["friction"], added_mass=st["added_mass"],
                    com_bias=st["com_bias"], kp_scale=st["kp_scale"], kd_scale=st["kd_scale"], rand_push_vels=st["rand_push_vels"],
                    actions_buf=st["actions"], last_actions=st["last_actions"], llast_actions=st["llast_actions"],
                    terrain_levels=st["terrain_levels"], end_q=st["q"], end_qd=st["qd"], gait_state=st["gait_state"])
        if spec.randomize_ctrl_delay:
            mine.update(action_queue=st["action_queue"], action_delay=st["action_delay"])
            # what the simulator was driven with (legged_robot.py:240-245) = the recorded queue at the env's delay slot
            keep = ~ref["reset_buf"].astype(bool)
            rq = ref["action_queue"].reshape(N, -1, spec.num_actions)
            assert np.array_equal(applied[keep], rq[np.arange(N), delay_before][keep]) and (delay_before > 0).any()
        for k, r in ref.items():
            if k not in mine:
                continue
            m = np.asarray(mine[k]).reshape(r.shape)
            if k in INTS:
                assert np.array_equal(m.astype(np.int64), r.astype(np.int64)), f"step {t}: {k}"
            else:
                assert np.allclose(m, r, rtol=2e-5, atol=2e-6), f"step {t}: {k} max err {np.abs(m - r).max():.3e}"
        for hk in ("obs_history", "critic_obs_buf"):
            if f"hist{t}/{hk}" in g:
                assert np.allclose(o[hk], g[f"hist{t}/{hk}"], rtol=2e-5, atol=2e-6), f"step {t}: {hk}"
        resets += int(ref["reset_buf"].sum())
    assert resets > 0
    if "oob" in name:                  # the window did contain out-of-bounds teleports (base x / y jumps by hundreds of metres)
        jump = np.abs(g["out/base_pos"][:, :, :2] - g["phys/base_pos"][:, :, :2]).max(axis=2)
        assert (jump > 100).sum() >= 4
    if "allrew" in name:
        on = {k for k, v in spec.reward_scales.items() if v != 0}
        assert {"dof_pos_stand_still", "dof_vel_stand_still", "termination", "thigh_pos", "foot_landing_vel", "keep_balance"} <= on
