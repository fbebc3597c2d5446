"""CPU checks of the CUDA kernel *sources* on the single-warp emulator (tests/warp_emu): the same .cuh files that
nvcc compiles for sm_100a are compiled by g++ with lock-step coroutine lanes, so kernel logic (and missing
__syncwarp() hazards) is caught here before GPU time is spent.  This is test infrastructure only: the product
library is built by nvcc and has no CPU path."""
import numpy as np
import pytest

from hcr_genesis_lr_cl_b200 import _cabi
from emu_util import EmuSim, oracle_params, oracle_policy_step
from golden_util import load_golden, load_terrain, out_at, phys_at, spec_for
from oracle.physics import PhysicsOracle

PH = dict(base_pos="base_pos", base_quat_wxyz="base_quat_wxyz", base_lin_w="base_lin_w", base_ang_w="base_ang_w",
          q="dof_pos", qd="dof_vel", torques="torques", link_force="link_contact_forces", feet_pos="feet_pos", feet_vel="feet_vel")
INTS = ("reset_buf", "time_out_buf", "episode_length", "fail_buf", "last_contacts", "terrain_levels", "action_delay")


@pytest.mark.parametrize("name,steps", [("go2_ts_n32", 5), ("go2_n32", 4), ("go2_cat_n32", 5), ("tron1_pf_n32", 5), ("tron1_pf_ee_n32", 5), ("go2_wtw_n32", 5),
                                        ("go2_cts_n32", 5), ("go2_ee_n32", 5), ("go2_dreamwaq_n32", 5)])
def test_emulated_env_kernel_matches_reference_golden(name, steps):
    """The preset-specialised instantiation b200_create selects for this descriptor (asserted: not the generic one)."""
    _run_golden(name, steps)


@pytest.mark.parametrize("name", ["go2_ts_allrew_oob_n32", "go2_wtw_smooth_n32", "tron1_pf_ee_smooth_n32"])
def test_emulated_env_kernel_matches_reference_golden_of_edited_configs(name):
    """Edited configurations (generic instantiation): every reward term of the Go2TS class on + an out-of-bounds teleport
    inside the window; the von Mises ("smooth") periodic-gait indicator, computed on the device instead of by scipy."""
    _run_golden(name, 6, specialized=False)


def test_emulated_env_kernel_with_control_delay_matches_reference_golden():
    """domain_rand.randomize_ctrl_delay on (no shipped config enables it -> generic instantiation): the env kernel clears
    the action queue and redraws the delay of envs that reset (legged_robot.py:144-148)."""
    _run_golden("go2_ts_delay_n32", 8, specialized=False)


@pytest.mark.parametrize("name", ["go2_ts_n32", "tron1_pf_ee_n32", "go2_wtw_n32", "go2_cat_n32"])
def test_emulated_generic_env_kernel_matches_reference_golden(name):
    """The generic instantiation (any edited configuration runs on it): descriptor ints read at run time."""
    _run_golden(name, 3, specialized=False)


def test_emulated_frame_stack_rings_wrap_around():
    """More steps than the critic stack has slots (5): the ring position wraps and the window stays the last K frames."""
    _run_golden("go2_ts_n32", 9)


@pytest.mark.parametrize("name", ["tron1_pf_ee_n32", "go2_wtw_n32"])
def test_emulated_env_kernel_reproduces_the_env0_coupling_on_request(name):
    """TaskSpec.reproduce_r18: the reference's flattened-nonzero() indexing makes env 0 follow "any env" for the gait-clock
    wrap and the swing / stance indicator.  With the switch on the dynamics kernel leaves the "any env"
    bits and the env kernel applies them to row 0: every env of the golden is compared, env 0 included."""
    _run_golden(name, 5, specialized=False, r18=True)


def _run_golden(name, steps, specialized=True, r18=False):
    g, s0 = load_golden(name)
    spec = spec_for(g)
    spec.reproduce_r18 = r18
    hs, origins = (load_terrain(spec) if spec.heightfield else (None, None))
    N = g["actions"].shape[1]
    sim = EmuSim(spec, N, hs, origins, specialized=specialized)
    sim.load_state(s0)
    B = sim.buf
    for t in range(steps):
        # pre-step bookkeeping of the dynamics kernel, then the recorded post-physics state
        a = np.clip(g["actions"][t], -spec.clip_actions, spec.clip_actions)
        if r18:                                # the real (emulated) dynamics kernel: pre-step bookkeeping + the R18 "any env" bits
            sim.dynamics_step(g["actions"][t])
        else:
            B["llast_actions"][:] = B["last_actions"]; B["last_actions"][:] = B["actions"]; B["actions"][:] = a
            if spec.randomize_ctrl_delay:          # the queue push of the dynamics kernel's pre-step (legged_robot.py:240-245)
                qu = B["action_queue"].reshape(N, -1, spec.num_actions)
                qu[:, 1:] = qu[:, :-1].copy(); qu[:, 0] = a
            B["last_dof_vel"][:] = B["dof_vel"]; B["last_feet_vel"][:] = B["feet_vel"]
        for k, b in PH.items():
            B[b][...] = phys_at(g, t)[k].reshape(B[b].shape)
        # what the dynamics kernel would leave for these joint velocities (CaT R4, bit 0); its R18 "any env" bits stay
        B["global_flags"][0] = (int(B["global_flags"][0]) & ~1) | int((np.abs(phys_at(g, t)["qd"]) > 4).any())
        B["stats"][:] = 0
        sim.env_post_step()
        assert (sim.last_preset >= 0) == specialized, "preset selection"
        ref = out_at(g, t)
        mine = dict(B, actions_buf=B["actions"], end_q=B["dof_pos"], end_qd=B["dof_vel"])
        if spec.obs_kind in ("tron1_pf", "tron1_pf_ee", "go2_wtw", "go2_ee"):     # the returned obs / privileged obs are the frame stacks
            mine["estimator_labels_buf"] = B["privileged_obs_buf"]
            mine["obs_buf"], mine["privileged_obs_buf"] = sim.obs_history, sim.critic_obs
        if spec.obs_kind == "go2_dreamwaq":                              # labels travel in privileged_obs_buf, the critic stack is returned
            mine["explicit_labels_buf"], mine["privileged_obs_buf"] = B["privileged_obs_buf"], sim.critic_obs
        skip0 = spec.obs_kind in ("tron1_pf_ee", "go2_wtw") and not r18  # R18: the reference couples env 0 to all envs (reproduced on request only)
        for k, r in ref.items():
            if k not in mine or k == "end_state":
                continue
            m = np.asarray(mine[k]).reshape(r.shape)
            if skip0 and m.ndim >= 1 and m.shape[0] == N:
                m, r = m[1:], r[1:]
            if k in INTS:
                assert np.array_equal(m.astype(np.int64), r.astype(np.int64)), f"step {t}: {k}"
            else:
                assert np.allclose(m, r, rtol=1e-4, atol=1e-5), f"step {t}: {k} max err {np.abs(m - r).max():.3e}"
        if f"hist{t}/obs_history" in g:
            assert np.allclose(sim.obs_history, g[f"hist{t}/obs_history"], rtol=1e-4, atol=1e-5)
            assert np.allclose(sim.critic_obs, g[f"hist{t}/critic_obs_buf"], rtol=1e-4, atol=1e-5)
        n = len(spec.episode_sum_names())
        assert B["stats"][n] == ref["reset_buf"].sum()            # reset counter of the per-step reductions
        # the last block finalised extras["episode"] into the ring slot of this step and re-armed the ticket counter
        assert B["global_flags"][1] == 0
        we = n + _cabi.H["B200_STATS_EXTRA"]
        ring = B["stats"][2 * n + 4 + (sim.step_counter % 32) * we:][:we]
        assert ring[n + 3] == ref["reset_buf"].sum()              # ... and again in the ring (device-stepped curricula read it there)
        cnt = max(float(ref["reset_buf"].sum()), 1.0)
        assert np.allclose(ring[:n], B["stats"][:n] / cnt / np.float32(spec.episode_length_s), rtol=1e-6, atol=0)
        if spec.terrain_curriculum:
            assert np.isclose(ring[n], B["terrain_levels"].mean(), rtol=1e-5)


def test_emulated_dynamics_kernel_matches_oracle():
    g, s0 = load_golden("go2_ts_n32")
    spec = spec_for(g)
    hs, origins = load_terrain()
    N = 6
    sub = {k: (v[:N] if getattr(v, "ndim", 0) > 0 and v.shape[0] == 32 else v) for k, v in s0.items()}
    sim = EmuSim(spec, N, hs, origins)
    sim.load_state(sub)
    before = {k: v.copy() for k, v in sim.buf.items()}
    a = g["actions"][0][:N]
    sim.dynamics_step(a)
    orc = PhysicsOracle(sim.model, oracle_params(spec, sim.model), hs, precision="f32")
    ref = oracle_policy_step(spec, sim.model, orc, before, a)
    assert ref["ncontact"].max() >= 3
    for k, name in PH.items():
        r = np.asarray(ref[k], np.float64)
        err = np.abs(sim.buf[name].reshape(r.shape) - r).max()
        tol = (2e-3 if k == "link_force" else 1e-4) * max(1.0, np.abs(r).max())
        assert err < tol, f"{k}: {err:.3e} >= {tol:.1e}"
    # pre-step bookkeeping (legged_robot.py:230-252, genesis_simulator.py:21-24)
    assert np.array_equal(sim.buf["actions"], np.clip(a, -100, 100))
    assert np.array_equal(sim.buf["last_actions"], before["actions"])
    assert np.array_equal(sim.buf["llast_actions"], before["last_actions"])
    assert np.array_equal(sim.buf["last_dof_vel"], before["dof_vel"])



def test_emulated_dynamics_without_domain_randomisation_keeps_nominal_mass_and_friction():
    """randomize_base_mass / randomize_friction off (go2_sysid-style cfg): the reference never calls set_mass_shift /
    set_friction_ratio, so the engine runs the nominal robot although `_added_base_mass` is ones and `_friction_values`
    zeros (genesis_simulator.py:644-650).  Kernel vs oracle with the switches off, and: the result must differ from what
    the buffers' values would give if they leaked into the physics."""
    g, s0 = load_golden("go2_ts_n32")
    hs, origins = load_terrain()
    N = 6
    sub = {k: (v[:N] if getattr(v, "ndim", 0) > 0 and v.shape[0] == 32 else v) for k, v in s0.items()}
    outs = {}
    for off in (True, False):
        spec = spec_for(g)
        if off:
            spec.randomize_base_mass = spec.randomize_friction = False
        sim = EmuSim(spec, N, hs, origins)
        sim.load_state(sub)
        sim.buf["added_mass"][:] = 1.0
        sim.buf["friction"][:] = 0.0
        before = {k: v.copy() for k, v in sim.buf.items()}
        a = g["actions"][0][:N]
        sim.dynamics_step(a)
        orc = PhysicsOracle(sim.model, oracle_params(spec, sim.model), hs, precision="f32")
        ref = oracle_policy_step(spec, sim.model, orc, before, a)
        for k, name in PH.items():
            r = np.asarray(ref[k], np.float64)
            err = np.abs(sim.buf[name].reshape(r.shape) - r).max()
            tol = (2e-3 if k == "link_force" else 1e-4) * max(1.0, np.abs(r).max())
            assert err < tol, f"DR off={off} {k}: {err:.3e} >= {tol:.1e}"
        outs[off] = sim.buf["dof_vel"].copy()
    assert np.abs(outs[True] - outs[False]).max() > 1e-3      # +1 kg and mu = terrain_mu only would have changed the step


def test_emulated_simulator_step_skips_the_pre_step_bookkeeping():
    """b200_simulator_step (plugin mode): same physics as b200_dynamics_step on already clipped actions, but the action
    history / delay queue of B200Buffers stay untouched (LeggedRobot._pre_sim_step owns them, legged_robot.py:230-252)."""
    import ctypes
    import torch
    from emu_backend import EmuB200Simulator
    g, s0 = load_golden("go2_ts_n32")
    spec = spec_for(g)
    hs, origins = load_terrain()
    N = 4
    sub = {k: (v[:N] if getattr(v, "ndim", 0) > 0 and v.shape[0] == 32 else v) for k, v in s0.items()}
    res = {}
    for mode in ("full", "sim_only"):
        sim = EmuB200Simulator(spec, None, "cpu", True, num_envs=N, terrain=(hs, origins))
        sim.load_state(sub)
        before = sim.get_state()
        a = torch.from_numpy(np.clip(g["actions"][0][:N], -5, 5).copy())
        if mode == "full":
            sim.step(a)
        else:
            sim._ck(sim._lib.b200_simulator_step(sim._handle, a.data_ptr(), ctypes.c_void_p(0)))
        res[mode] = (before, sim.get_state())
    (b0, full), (b1, only) = res["full"], res["sim_only"]
    for k in ("base_pos", "base_quat_wxyz", "dof_pos", "dof_vel", "torques", "link_contact_forces", "feet_pos", "last_dof_vel", "last_feet_vel"):
        assert np.array_equal(full[k], only[k]), k
    for k in ("actions", "last_actions", "llast_actions"):
        assert np.array_equal(only[k], b1[k]), k
    assert not np.array_equal(full["actions"], b0["actions"])


def test_emulated_nonfinite_state_is_contained_and_reset():
    """An env whose state goes non-finite inside the dynamics kernel (here: a NaN joint velocity on entry) is not written
    back: it keeps its pose, at rest, is flagged, and the env kernel resets it -- its neighbours are untouched."""
    import torch
    from emu_backend import EmuFusedLeggedEnv
    from hcr_genesis_lr_cl_b200 import task_spec as T
    hs, origins = load_terrain()
    spec = T.go2_ts_spec()
    N = 4
    env = EmuFusedLeggedEnv(spec, N, "cpu", terrain=(hs, origins))
    ref = EmuFusedLeggedEnv(spec, N, "cpu", terrain=(hs, origins))
    for e in (env, ref):
        e.reset()
        e.step(torch.zeros(N, 12))
    b = env.simulator._buf
    pose = b["base_pos"][2].clone()
    b["dof_vel"][2, 5] = float("nan")
    a = 0.3 * torch.ones(N, 12)
    env.step(a)
    ref.step(a)
    assert int(b["global_flags"][2]) == 1 and bool(env.reset_buf[2]) and not bool(env.reset_buf[0])
    for k in ("base_pos", "dof_pos", "dof_vel", "base_lin_w", "obs_buf", "rew_buf"):
        assert torch.isfinite(b[k]).all(), k
        assert torch.equal(b[k][[0, 1, 3]], ref.simulator._buf[k][[0, 1, 3]]), k          # the other envs never noticed
    assert int(b["nonfinite"].sum()) == 0 and int(b["episode_length"][2]) == 0              # flag consumed, episode restarted
    assert not torch.equal(b["base_pos"][2], pose)                                          # re-placed by reset_idx
    env.step(a)
    assert torch.isfinite(b["obs_buf"]).all() and int(b["global_flags"][2]) == 1


@pytest.mark.parametrize("zero_copy", ["2", "1", "0"])
def test_emulated_host_buffer_step_equals_the_device_step(monkeypatch, zero_copy):
    """b200_env_step with HOST buffers -- actions read in place and the rew | reset | time_out slab written by the env kernel's
    finalising CTA (zero copy, the default), or staged through copies (B200_ZERO_COPY_*=0) -- leaves the same state and the
    same host results as the step with device-resident actions.  Ragged slab sizes (6 N not a multiple of 16) take the copy path."""
    import numpy as np
    import torch
    from emu_backend import EmuFusedLeggedEnv
    from hcr_genesis_lr_cl_b200 import task_spec as TS
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    monkeypatch.setenv("B200_ZERO_COPY_ACTIONS", zero_copy)
    monkeypatch.setenv("B200_ZERO_COPY_RESULTS", "0" if zero_copy == "0" else "1")
    spec = TS.PRESETS["go2_ts"]()
    for N in (16, 13):
        a_env = EmuFusedLeggedEnv(spec, N, torch.device("cpu"), terrain=terrain_for(spec))
        b_env = EmuFusedLeggedEnv(spec, N, torch.device("cpu"), terrain=terrain_for(spec))
        a_env.reset(); b_env.reset()
        ep = a_env.episode_length_buf
        ep[0::3] = int(a_env.max_episode_length) - 2          # time-outs (and resets) inside the window
        b_env.episode_length_buf = ep.clone()
        rew, rst, tmo = b_env.simulator.make_host_step_buffers()
        rng = np.random.default_rng(5)
        for t in range(4):
            act = torch.from_numpy(rng.normal(size=(N, spec.num_actions)).astype(np.float32))
            host_act = act.clone(); host_act._emu_host = True
            out_a = a_env.step(act)
            out_b = b_env.step_host(host_act, rew, rst, tmo)
            for x, y in zip(out_a[:6], out_b[:6]):
                assert torch.equal(x, y)
            assert torch.equal(rew, a_env.rew_buf) and torch.equal(rst, a_env.reset_buf) and torch.equal(tmo, a_env.time_out_buf)
        assert int(a_env.reset_buf.sum()) >= 0


@pytest.mark.parametrize("N", [5, 13, 32, 56, 57, 100])
def test_emulated_dynamics_order_is_a_permutation(N):
    """dynamics_order_kernel (cost-ordered warp slots; env-granular dealing over the SMs when the launch is one wave, heaviest
    first otherwise) must hand every env to exactly one warp slot, whatever the env count: checked on the kernel's own source."""
    import numpy as np
    import torch
    from emu_backend import EmuFusedLeggedEnv
    from hcr_genesis_lr_cl_b200 import task_spec as TS
    spec = TS.PRESETS["go2"]()
    env = EmuFusedLeggedEnv(spec, N, torch.device("cpu"))
    env.reset()
    rng = np.random.default_rng(N)
    for t in range(4):
        env.step(torch.from_numpy(rng.normal(size=(N, spec.num_actions)).astype(np.float32)))
        order = env.simulator._buf["dyn_order"].numpy().reshape(2, N)
        for half in order:
            envs = np.arange(N) + half                   # delta-encoded: slot + delta = env (a zeroed buffer is the identity)
            assert sorted(envs.tolist()) == list(range(N)), (N, t, envs)
    cost = env.simulator._buf["dyn_cost"].numpy()
    assert (cost >= 0).all()


def test_emulated_host_buffer_step_with_misaligned_buffers_takes_the_copy_path():
    """Host buffers that are not 16-byte aligned (views at an odd offset of a larger pinned allocation) cannot be read / written
    with the 16-byte accesses of the zero-copy paths: b200_env_step must fall back to copies and still hand back the same values."""
    import numpy as np
    import torch
    from emu_backend import EmuFusedLeggedEnv
    from hcr_genesis_lr_cl_b200 import task_spec as TS
    spec = TS.PRESETS["go2"]()
    N = 16
    a_env = EmuFusedLeggedEnv(spec, N, torch.device("cpu"))
    b_env = EmuFusedLeggedEnv(spec, N, torch.device("cpu"))
    a_env.reset(); b_env.reset()
    slab = torch.zeros(6 * N + 4, dtype=torch.uint8)[4:]                 # 4 bytes into the allocation
    rew, rst, tmo = slab[:4 * N].view(torch.float32), slab[4 * N:5 * N].view(torch.bool), slab[5 * N:].view(torch.bool)
    big = torch.zeros(N * spec.num_actions + 1, dtype=torch.float32)
    rng = np.random.default_rng(9)
    for t in range(3):
        act = torch.from_numpy(rng.normal(size=(N, spec.num_actions)).astype(np.float32))
        host_act = big[1:].view(N, spec.num_actions)                     # 4 bytes into the allocation
        host_act.copy_(act); host_act._emu_host = True
        out_a = a_env.step(act)
        out_b = b_env.step_host(host_act, rew, rst, tmo)
        assert torch.equal(out_a[0], out_b[0])
        assert torch.equal(rew, a_env.rew_buf) and torch.equal(rst, a_env.reset_buf) and torch.equal(tmo, a_env.time_out_buf)
