"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference-recorded golden vectors.

Tolerances (north_star: "bit-exact reset/termination/terrain-cell indices; 1e-4 relative on rewards, observations
and single-step states"):
  * integer/bool outputs: exact;
  * env-half floats on injected identical inputs: rtol 1e-4 with an absolute floor of 1e-5;
  * one decimated physics step vs the fp32 oracle: 1e-4 of the quantity's scale (the fp32 oracle itself sits 1e-3
    from the fp64 one on velocities because penetration depths are differences of ~20 m world coordinates);
  * contact forces vs the fp32 oracle: 5e-3 of the largest force in the env (the two implementations may leave the
    block Gauss-Seidel solver one sweep apart, and on a slowly converging contact set the force still moves by a few
    1e-3 per sweep when the stopping rule -- last change <= 1e-4 (1 + |f|max) -- fires; the states, which only see
    J^T f, stay within 1e-4).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from golden_util import load_golden, load_terrain, out_at, phys_at, spec_for  # noqa: E402

PH = dict(base_pos="base_pos", base_quat_wxyz="base_quat_wxyz", base_lin_w="base_lin_w", base_ang_w="base_ang_w",
          q="dof_pos", qd="dof_vel", torques="torques", link_force="link_contact_forces", feet_pos="feet_pos",
          feet_vel="feet_vel")


def _env(spec, N, terrain=None, **kw):
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    return FusedLeggedEnv(spec, N, "cuda:0", terrain=terrain, debug_cells=True, **kw)


def _close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.asarray(a).shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"{what}: max violation {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.fixture(scope="module")
def golden():
    g, s0 = load_golden("go2_ts_n32")
    return g, s0, spec_for(g), load_terrain()


def test_library_is_the_cuda_path(built_library):
    from hcr_genesis_lr_cl_b200 import _cabi
    lib = _cabi.load_library()
    assert all(hasattr(lib, s) for s in _cabi.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("name", ["go2_ts_n32", "go2_n32", "go2_cat_n32", "tron1_pf_n32", "tron1_pf_ee_n32", "go2_wtw_n32", "go2_cts_n32", "go2_ee_n32",
                                  "go2_dreamwaq_n32"])
def test_env_kernel_matches_reference_golden(name):
    """Injected post-physics states from the reference run -> every output of the fused kernel."""
    _golden_run(name)


@pytest.mark.parametrize("name", ["go2_ts_allrew_oob_n32", "go2_wtw_smooth_n32", "tron1_pf_ee_smooth_n32"])
def test_env_kernel_matches_reference_golden_of_edited_configs(name):
    """Generic instantiation: all reward terms of the Go2TS class on + an out-of-bounds teleport inside the window; the
    von Mises ("smooth") periodic-gait indicator evaluated on the device (the reference calls scipy on the host)."""
    _golden_run(name)


def test_control_delay_queue_matches_reference_golden():
    """domain_rand.randomize_ctrl_delay (legged_robot.py:144-148,240-245): the dynamics kernel pushes the clipped action
    into the env's queue and drives the joints with the delayed slot, the env kernel clears the queue and redraws the
    delay on reset; runs on the generic env-kernel instantiation (no shipped config enables the delay)."""
    _golden_run("go2_ts_delay_n32")


def test_frame_stack_views_are_strided_windows_of_the_rings():
    """The frame stacks handed to the runner are [N, K * width] views of the double-written rings: row stride 2 K width,
    contiguous rows, no copy -- and they hold exactly the reference's stacks (checked against the goldens above)."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    spec = T.go2_ts_spec()
    env = _env(spec, 64, load_terrain(spec))
    env.reset()
    for _ in range(7):
        env.step(torch.randn(64, 12, device="cuda"))
        h, c = env.obs_history, env.critic_obs_buf
        assert h.shape == (64, 900) and c.shape == (64, 885) and h.stride() == (2 * 21 * 45, 1) and c.stride() == (2 * 6 * 177, 1)
        assert h.untyped_storage().data_ptr() == env.simulator._buf["obs_history"].untyped_storage().data_ptr()
        assert torch.equal(h[:, -45:], env.obs_buf)                      # newest frame last (legged_robot_ts.py:41-47)
        if _ > 0:                                                        # the views of the previous step are still intact (ring's spare slot)
            assert torch.equal(prev[0], prev[1]) and torch.equal(prev[2], prev[3])
        prev = (h, h.clone(), c, c.clone())
        w = torch.nn.Linear(900, 8, device="cuda")
        assert torch.allclose(w(h), w(h.contiguous()), atol=1e-5)         # a strided batch feeds nn.Linear as it is


@pytest.mark.parametrize("name", ["tron1_pf_ee_n32", "go2_wtw_n32"])
def test_env_kernel_reproduces_the_env0_coupling_on_request(name):
    """TaskSpec.reproduce_r18 (bug-compatible mode): every env of the reference golden, env 0 included."""
    _golden_run(name, r18=True)


def _golden_run(name, r18=False):
    g, s0 = load_golden(name)
    spec = spec_for(g)
    spec.reproduce_r18 = r18
    terrain = load_terrain(spec) if spec.heightfield else None
    N, T_ = g["actions"].shape[1], g["actions"].shape[0]
    env = _env(spec, N, terrain)
    sim = env.simulator
    sim.load_state(s0)
    env.common_step_counter = int(s0["common_step_counter"])
    env.command_ranges["lin_vel_x"] = [float(x) for x in s0["cmd_range_x"]]
    _load_behavior(env, s0)
    ints = ("reset_buf", "time_out_buf", "episode_length", "fail_buf", "last_contacts", "terrain_levels", "action_delay")
    for t in range(T_):
        a = torch.from_numpy(g["actions"][t]).cuda()
        sim.step(a)                                   # pre-step bookkeeping (+ our own physics, overwritten below)
        sim.load_state({PH[k]: v for k, v in phys_at(g, t).items()})
        # what the dynamics kernel would leave for these joint velocities (CaT R4, bit 0); its R18 "any env" bits stay
        sim._buf["global_flags"][0] = (int(sim._buf["global_flags"][0]) & ~1) | int((np.abs(phys_at(g, t)["qd"]) > 4).any())
        env.common_step_counter += 1
        env._set_step_flags()
        sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
        st = sim.get_state()
        ref = out_at(g, t)
        mine = dict(st, actions_buf=st["actions"], end_q=st["dof_pos"], end_qd=st["dof_vel"])
        if spec.obs_kind in ("tron1_pf", "tron1_pf_ee", "go2_wtw", "go2_ee"):     # the returned obs / privileged obs are the frame stacks
            mine["estimator_labels_buf"] = st["privileged_obs_buf"]
            mine["obs_buf"], mine["privileged_obs_buf"] = st["obs_history"], st["critic_obs"]
        if spec.obs_kind == "go2_dreamwaq":                              # labels travel in privileged_obs_buf, the critic stack is returned
            mine["explicit_labels_buf"], mine["privileged_obs_buf"] = st["privileged_obs_buf"], st["critic_obs"]
        skip0 = spec.obs_kind in ("tron1_pf_ee", "go2_wtw") and not r18  # R18: env 0 is coupled to all envs in the reference (reproduced on request only)
        for k, r in ref.items():
            if k not in mine or k in ("end_state",):
                continue
            m = np.asarray(mine[k]).reshape(r.shape)
            if skip0 and m.ndim >= 1 and m.shape[0] == N:
                m, r = m[1:], r[1:]
            if k in ints:
                assert np.array_equal(m.astype(np.int64), r.astype(np.int64)), f"step {t}: {k} not bit-exact"
            else:
                _close(m, r, what=f"step {t}: {k}")
        for hk, name in (("obs_history", "obs_history"), ("critic_obs_buf", "critic_obs")):
            if f"hist{t}/{hk}" in g:
                _close(st[name], g[f"hist{t}/{hk}"], what=f"step {t}: {hk}")
    assert int(g["out/reset_buf"].sum()) > 0          # the window did contain resets


def _load_behavior(env, s0):
    """go2_wtw curriculum state (behaviour ranges, unlocked gaits) of a recorded window."""
    if "beh_ranges" in s0:
        r = np.asarray(s0["beh_ranges"], np.float64).tolist()
        env.gait_period_range, env.base_height_target_range, env.foot_clearance_target_range, env.pitch_target_range = r
        env.num_gaits = int(s0["num_gaits"])


def _random_state(spec, N, terrain, seed):
    """A seeded, physically plausible state: robots dropped near their origins with joint/velocity noise."""
    rng = np.random.default_rng(seed)
    model = spec.load_model()
    A = spec.num_actions
    hs, origins = terrain if terrain is not None else (None, None)
    st = {}
    if spec.heightfield:
        lv = rng.integers(0, spec.num_rows, N)
        ty = rng.integers(0, spec.num_cols, N)
        org = origins[lv, ty]
        st["terrain_levels"], st["terrain_types"] = lv.astype(np.int64), ty.astype(np.int64)
    else:
        org = np.zeros((N, 3), np.float32)
        org[:, :2] = rng.uniform(-20, 20, (N, 2))
    st["env_origins"] = org.astype(np.float32)
    pos = org + np.array(spec.init_pos, np.float32)
    pos[:, :2] += rng.uniform(-1.5, 1.5, (N, 2))
    pos[:, 2] += rng.uniform(-0.12, 0.05, N)
    st["base_pos"] = pos.astype(np.float32)
    qt = rng.normal(size=(N, 4)) * np.array([1, 0.15, 0.15, 0.5])
    qt[:, 0] = np.abs(qt[:, 0]) + 1.0
    st["base_quat_wxyz"] = (qt / np.linalg.norm(qt, axis=1, keepdims=True)).astype(np.float32)
    st["base_lin_w"] = rng.normal(size=(N, 3)).astype(np.float32) * 0.5
    st["base_ang_w"] = rng.normal(size=(N, 3)).astype(np.float32)
    st["dof_pos"] = (np.array(spec.default_dof_pos, np.float32) + rng.uniform(-0.4, 0.4, (N, A))).astype(np.float32)
    st["dof_vel"] = (rng.normal(size=(N, A)) * 2).astype(np.float32)
    st["friction"] = rng.uniform(0.2, 1.7, (N, 1)).astype(np.float32)
    st["added_mass"] = rng.uniform(-1, 1, (N, 1)).astype(np.float32)
    st["com_bias"] = rng.uniform(-0.03, 0.03, (N, 3)).astype(np.float32)
    st["kp_scale"] = rng.uniform(0.8, 1.2, (N, A)).astype(np.float32)
    st["kd_scale"] = rng.uniform(0.8, 1.2, (N, A)).astype(np.float32)
    st["joint_armature"] = rng.uniform(0.11, 0.13, (N, 1)).astype(np.float32)     # consumed only when the task randomises them
    st["joint_friction"] = rng.uniform(0.0, 0.01, (N, 1)).astype(np.float32)
    st["joint_damping"] = rng.uniform(1.4, 1.45, (N, 1)).astype(np.float32)
    return st, model


@pytest.mark.parametrize("task", ["go2_ts", "go2", "tron1_pf", "tron1_pf_ee", "go2_wtw", "go2_ts:trimesh", "tron1_pf_ee:trimesh"])
def test_dynamics_kernel_matches_oracle(task):
    """One policy step (4 substeps) of the warp-per-env kernel vs the fp32 C oracle on seeded states.  ":trimesh" =
    terrain.mesh_type "trimesh": the same height samples triangulated like convert_heightfield_to_trimesh does (the oracle's
    surface for that mode is pinned against the reference's triangles in test_oracle_physics.py)."""
    from emu_util import oracle_params, oracle_policy_step
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.physics import PhysicsOracle
    task, _, mesh = task.partition(":")
    spec = T.PRESETS[task](**({"mesh_type": mesh} if mesh else {}))
    terrain = load_terrain(spec) if spec.heightfield else None
    N = 256
    st, model = _random_state(spec, N, terrain, seed=7)
    env = _env(spec, N, terrain)
    sim = env.simulator
    sim.load_state(st)
    full = sim.get_state()
    actions = np.random.default_rng(3).normal(size=(N, spec.num_actions)).astype(np.float32)
    sim.step(torch.from_numpy(actions).cuda())
    out = sim.get_state()
    orc = PhysicsOracle(model, oracle_params(spec, model), terrain[0] if terrain else None, precision="f32")
    ref = oracle_policy_step(spec, model, orc, full, actions)
    assert ref["ncontact"].max() > 0 and (ref["ncontact"] == 0).any()      # both contact and flight are covered
    scale = lambda x: max(1.0, float(np.abs(x).max()))
    for k, name in PH.items():
        r = np.asarray(ref[k], np.float64)
        tol = 5e-3 if k == "link_force" else 1e-4
        err = np.abs(out[name].reshape(r.shape) - r)
        if k == "link_force":
            per_env = np.abs(r).reshape(N, -1).max(1)[:, None, None] + 1.0
            assert (err / per_env).max() < tol, f"{k}: {(err / per_env).max():.3e}"
        else:
            assert err.max() < tol * scale(r), f"{k}: abs err {err.max():.3e} (scale {scale(r):.2f})"


def test_dynamics_kernel_without_domain_randomisation():
    """randomize_base_mass / randomize_friction off: mass shift 0 and friction ratio 1 whatever the observation-side
    buffers hold (ones / zeros as the reference initialises them, genesis_simulator.py:644-650)."""
    from emu_util import oracle_params, oracle_policy_step
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.physics import PhysicsOracle
    spec = T.go2_ts_spec(randomize_base_mass=False, randomize_friction=False, randomize_com_displacement=False)
    terrain = load_terrain(spec)
    N = 256
    st, model = _random_state(spec, N, terrain, seed=17)
    st["added_mass"] = np.ones((N, 1), np.float32)
    st["friction"] = np.zeros((N, 1), np.float32)
    env = _env(spec, N, terrain)
    sim = env.simulator
    assert sim.env_kernel_variant == "generic"          # an edited configuration runs the generic env kernel
    sim.load_state(st)
    full = sim.get_state()
    actions = np.random.default_rng(5).normal(size=(N, spec.num_actions)).astype(np.float32)
    sim.step(torch.from_numpy(actions).cuda())
    out = sim.get_state()
    orc = PhysicsOracle(model, oracle_params(spec, model), terrain[0], precision="f32")
    ref = oracle_policy_step(spec, model, orc, full, actions)
    assert ref["ncontact"].max() > 0
    for k in ("base_pos", "q", "qd", "base_lin_w", "base_ang_w"):
        r = np.asarray(ref[k], np.float64)
        err = np.abs(out[PH[k]].reshape(r.shape) - r).max()
        assert err < 1e-4 * max(1.0, float(np.abs(r).max())), f"{k}: {err:.3e}"


def test_full_size_one_step_matches_the_oracle_pair():
    """BASELINE config C2 at its full size: 4096 envs of go2_ts, ONE policy step, CUDA path vs the oracle pair (C substep
    + numpy env half) on the same seeded state -- every output, ints exact (the CPU port steps 4096 envs in well under a second)."""
    from emu_util import oracle_params, oracle_policy_step
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.env_oracle import EnvOracle
    from oracle.physics import PhysicsOracle
    spec = T.go2_ts_spec()
    terrain = load_terrain(spec)
    N = 4096
    st, model = _random_state(spec, N, terrain, seed=23)
    rng = np.random.default_rng(31)
    st["episode_length"] = rng.integers(0, 1001, N).astype(np.int32)      # time-outs, command resampling (every 500 steps)
    st["fail_buf"] = rng.integers(0, 6, N).astype(np.int32)
    st["commands"] = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    env = _env(spec, N, terrain)
    sim = env.simulator
    sim.load_state(st)
    env.common_step_counter = 1499           # the step under test is a push step (push_interval 500)
    full = sim.get_state()
    actions = np.random.default_rng(29).normal(size=(N, spec.num_actions)).astype(np.float32)
    # dynamics kernel vs the C oracle
    sim.step(torch.from_numpy(actions).cuda())
    mid = sim.get_state()
    orc = PhysicsOracle(model, oracle_params(spec, model), terrain[0], precision="f32")
    ref = oracle_policy_step(spec, model, orc, full, actions)
    for k in ("base_pos", "q", "qd", "base_lin_w", "base_ang_w"):
        r = np.asarray(ref[k], np.float64)
        err = np.abs(mid[PH[k]].reshape(r.shape) - r).max()
        assert err < 1e-4 * max(1.0, float(np.abs(r).max())), f"{k}: {err:.3e}"
    # env kernel vs the numpy oracle on the kernel's own post-physics state
    eo = EnvOracle(spec, N, terrain[0], terrain[1])
    alias = {"dof_pos": "q", "dof_vel": "qd", "obs_history": "obs_hist", "critic_obs": "critic_hist"}
    for k, v in full.items():
        kk = alias.get(k, k)
        if kk in eo.st:
            eo.st[kk][...] = v.reshape(eo.st[kk].shape)
    eo.common_step_counter = env.common_step_counter
    eo.pre_step(actions)
    o = eo.post_step(dict(base_pos=mid["base_pos"], base_quat_wxyz=mid["base_quat_wxyz"], base_lin_w=mid["base_lin_w"],
                          base_ang_w=mid["base_ang_w"], q=mid["dof_pos"], qd=mid["dof_vel"], torques=mid["torques"],
                          link_force=mid["link_contact_forces"], feet_pos=mid["feet_pos"], feet_vel=mid["feet_vel"]))
    env.common_step_counter += 1
    sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
    out = sim.get_state()
    assert int(o["reset_buf"].sum()) > 0
    for k in ("reset_buf", "time_out_buf"):
        assert np.array_equal(out[k].astype(bool), np.asarray(o[k]).astype(bool)), k
    assert np.array_equal(out["height_cells"], o["height_cells"]), "height-scan cells"
    _close(out["rew_buf"], o["rew_buf"], what="rew_buf")
    _close(out["obs_buf"], o["obs_buf"], what="obs_buf")
    _close(out["privileged_obs_buf"], o["privileged_obs_buf"], what="privileged_obs_buf")
    _close(out["critic_obs"], o["critic_obs_buf"], what="critic stack")
    _close(out["obs_history"], o["obs_history"], what="obs history")
    _close(out["rand_push_vels"], eo.st["rand_push_vels"], what="push")


@pytest.mark.parametrize("task", ["go2_ts", "go2_cat", "tron1_pf_ee", "go2_wtw", "go2_cts", "go2_ee", "go2_dreamwaq"])
def test_env_kernel_matches_numpy_oracle_seeded(task):
    """Seeded random states at N=512 through several fused steps vs the numpy restatement (all phases, with resets)."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.env_oracle import EnvOracle
    spec = T.PRESETS[task]()
    terrain = load_terrain(spec) if spec.heightfield else None
    F = len(spec.link_groups(spec.load_model())[0])
    N = 512
    st, model = _random_state(spec, N, terrain, seed=11)
    rng = np.random.default_rng(5)
    st["base_pos"][7::61, 0] += 500.0                        # out of the terrain bounds: teleported home (genesis_simulator.py:612-628)
    st["base_pos"][11::61, 1] -= 500.0
    st["episode_length"] = rng.integers(0, 1001, N).astype(np.int32)
    st["episode_length"][::7] = 499
    st["fail_buf"] = rng.integers(0, 6, N).astype(np.int32)
    st["commands"] = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    st["feet_air_time"] = rng.uniform(0, 0.4, (N, F)).astype(np.float32)
    st["last_contacts"] = rng.integers(0, 2, (N, F)).astype(np.uint8)
    from hcr_genesis_lr_cl_b200._cabi import H
    gs = np.zeros((N, H["B200_GAIT_STATE"]), np.float32)
    gs[:, H["B200_GS_GT"]] = rng.integers(0, 15, N) * np.float32(spec.dt)
    if spec.behavior_enabled:                                # go2_wtw: per-env behaviour parameters, one of the four gaits
        st["episode_length"][3::7] = 249                     # behaviour resampling interval (250 steps) hit in the callback
        gs[:, H["B200_GS_TH"]:H["B200_GS_TH"] + 4] = np.asarray(spec.gait_theta_lists, np.float32)[rng.integers(0, 4, N)]
        for col, r in ((H["B200_GS_PER"], spec.gait_period_range), (H["B200_GS_BH"], spec.base_height_target_range),
                       (H["B200_GS_FC"], spec.foot_clearance_target_range), (H["B200_GS_PT"], spec.pitch_target_range)):
            gs[:, col] = rng.uniform(r[0], r[1], N)
        gs[:, H["B200_GS_PHI"]] = gs[:, H["B200_GS_GT"]] / gs[:, H["B200_GS_PER"]]
    else:
        gs[:, 0], gs[:, 1] = rng.uniform(0, 1, N), rng.uniform(0, 1, N) + 0.5
        gs[:, H["B200_GS_PER"]] = spec.gait_period
        gs[:, H["B200_GS_PHI"]] = gs[:, H["B200_GS_GT"]] / np.float32(spec.gait_period)
    st["gait_state"] = gs
    env = _env(spec, N, terrain)
    sim = env.simulator
    sim.load_state(st)
    env.common_step_counter = spec.push_interval - 2
    eo = EnvOracle(spec, N, *(terrain if terrain is not None else (None, None)))
    if spec.behavior_enabled:                                # a mid-curriculum state: wide ranges, all gaits unlocked
        rr = [[0.35, 0.55], [0.24, 0.32], [0.05, 0.1], [-0.2, 0.2]]
        _load_behavior(env, dict(beh_ranges=rr, num_gaits=4))
        eo.beh_ranges = dict(zip(("gait_period", "base_height", "foot_clearance", "pitch"), rr))
        eo.num_gaits = 4
    alias = {"dof_pos": "q", "dof_vel": "qd"}
    # explicit phases: the oracle is fed exactly the kernel's post-physics state (this test isolates the env half)
    s_now = sim.get_state()
    for k, v in s_now.items():
        kk = alias.get(k, k)
        if kk in eo.st:
            eo.st[kk][...] = v.reshape(eo.st[kk].shape)
    eo.st["obs_hist"][:] = 0
    eo.st["critic_hist"][:] = 0
    eo.common_step_counter = env.common_step_counter
    total_resets = 0
    for t in range(4):
        a = rng.normal(size=(N, spec.num_actions)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda())
        mid = sim.get_state()
        phys = {k: mid[name] for k, name in PH.items()}
        eo.pre_step(a)
        o = eo.post_step(phys)
        env.common_step_counter += 1
        env._set_step_flags()
        sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
        out = sim.get_state()
        total_resets += int(o["reset_buf"].sum())
        assert np.array_equal(out["reset_buf"].astype(bool), o["reset_buf"]), f"step {t}: reset_buf"
        assert np.array_equal(out["time_out_buf"].astype(bool), o["time_out_buf"]), f"step {t}: time_out_buf"
        if spec.heightfield:
            assert np.array_equal(out["height_cells"], o["height_cells"]), f"step {t}: height-scan cell indices"
            assert np.array_equal(out["terrain_levels"], eo.st["terrain_levels"]), f"step {t}: terrain levels"
        assert np.array_equal(out["episode_length"], eo.st["episode_length"])
        assert np.array_equal(out["fail_buf"], eo.st["fail_buf"])
        _close(out["rew_buf"], o["rew_buf"], what=f"step {t}: rew")
        if spec.cat_enabled:
            assert np.array_equal(out["cstr_prob"], o["cstr_prob"]), f"step {t}: cstr_prob"
        if spec.obs_kind == "tron1_pf_ee":
            _close(out["privileged_obs_buf"], o["estimator_labels_buf"], what=f"step {t}: labels")
            _close(out["gait_state"], eo.st["gait_state"], what=f"step {t}: gait_state")
        elif spec.obs_kind == "go2_ee":
            _close(out["privileged_obs_buf"], o["estimator_labels_buf"], what=f"step {t}: labels")
        elif spec.obs_kind == "go2_dreamwaq":
            _close(out["obs_buf"], o["obs_buf"], what=f"step {t}: obs")
            _close(out["privileged_obs_buf"], o["explicit_labels_buf"], what=f"step {t}: labels")
            _close(out["next_state_buf"], o["next_state_buf"], what=f"step {t}: next_state")
        elif spec.obs_kind == "go2_wtw":
            _close(out["gait_state"], eo.st["gait_state"], what=f"step {t}: gait_state")
        else:
            _close(out["obs_buf"], o["obs_buf"], what=f"step {t}: obs")
            _close(out["privileged_obs_buf"], o["privileged_obs_buf"], what=f"step {t}: priv")
        _close(out["obs_history"], o["obs_history"], what=f"step {t}: obs_history")
        _close(out["critic_obs"], o["critic_obs_buf"], what=f"step {t}: critic")
        _close(out["episode_sums"], eo.st["episode_sums"], what=f"step {t}: episode_sums")
        _close(out["commands"], eo.st["commands"], what=f"step {t}: commands")
        _close(out["dof_pos"], eo.st["q"], what=f"step {t}: reset dof_pos")
        _close(out["base_pos"], eo.st["base_pos"], what=f"step {t}: reset base_pos")
        if t == 0:                                           # the out-of-bounds envs are back inside after the first post step
            assert np.abs(out["base_pos"][7::61, 0] - st["base_pos"][7::61, 0]).min() > 100 and np.abs(mid["base_pos"][7::61, 0] - st["base_pos"][7::61, 0]).max() < 10
            _close(out["feet_pos"], o["feet_pos"], what="step 0: feet_pos after the out-of-bounds teleport")
        for k in ("friction", "added_mass", "com_bias", "kp_scale", "kd_scale"):
            _close(out[k], eo.st[k], what=f"step {t}: {k}")
        if o["episode_means"] is not None:                                   # extras["episode"] of this step (device ring)
            n = len(eo.sum_names)
            we = n + H["B200_STATS_EXTRA"]
            base = 2 * n + 4 + (env.common_step_counter % 32) * we
            ring = out["stats"][base:base + we]
            for i, name in enumerate(eo.sum_names):
                assert abs(ring[i] - o["episode_means"]["rew_" + name]) <= 1e-4 * abs(o["episode_means"]["rew_" + name]) + 1e-6, name
            if spec.terrain_curriculum:
                assert abs(ring[n] - eo.st["terrain_levels"].mean()) < 1e-3
    assert total_resets > 0


def test_full_size_properties():
    """BASELINE config C2 at full size (4096 envs): size-independent properties instead of an oracle run."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    spec = T.go2_ts_spec()
    terrain = load_terrain()
    N = 4096
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(rng.normal(size=(N, 12)).astype(np.float32)).cuda() for _ in range(6)]

    def rollout(n, offset=0, total=None, a_slice=slice(None)):
        env = _env(spec, n, terrain, env_offset=offset, num_envs_global=total or n)
        env.reset()
        hist_prev = None
        for a in acts:
            out = env.step(a[a_slice].contiguous())
            hist = out[2].clone()
            if hist_prev is not None:
                keep = ~out[5]
                # history stack = previous stack shifted by one frame + the new observation (unclipped == clipped here)
                assert torch.equal(hist[keep][:, :-45], hist_prev[keep][:, 45:])
                assert torch.allclose(hist[:, -45:], out[0], atol=0)
            hist_prev = hist
        return env, {k: v.clone() for k, v in env.simulator._buf.items()}

    env_a, sa = rollout(N)
    env_b, sb = rollout(N)
    for k in sa:                                                     # determinism: bit-identical reruns
        if k in ("stats", "dyn_order"):            # float atomics (logging only) / scheduling state (ties broken by atomics)
            continue
        assert torch.equal(sa[k], sb[k]), f"non-deterministic buffer {k}"
    for par in range(2):                           # the scheduling state is a permutation, whatever the tie-breaking
        assert torch.equal(torch.sort(sa["dyn_order"][par] + torch.arange(N, device="cuda", dtype=torch.int32)).values,
                           torch.arange(N, device="cuda", dtype=torch.int32))
    # sharding invariance: the second half computed as its own shard equals the same envs inside the full job
    half = N // 2
    env_c, sc = rollout(half, offset=half, total=N, a_slice=slice(half, N))
    for k in ("obs_buf", "rew_buf", "reset_buf", "dof_pos", "base_pos", "commands", "episode_sums", "measured_heights"):
        assert torch.equal(sa[k][half:], sc[k]), f"shard != full job for {k}"
    assert torch.isfinite(sa["obs_buf"]).all() and torch.isfinite(sa["rew_buf"]).all()
    assert sa["height_cells"].min() >= 0 and sa["height_cells"].max() <= 1198
    n_steps = len(acts) + 1                                           # reset() = reset_all + one zero-action step
    assert env_a.simulator.launch_count == 1 + 2 * n_steps + (n_steps - 1)   # (dynamics, env) per step + the env-order kernel from the 2nd step on


def test_short_rollout_against_oracle(golden):
    """Closed-loop (physics + env) rollout of 5 policy steps vs the fp32 oracle pair; contact dynamics are chaotic,
    so only a short horizon and a loose bound are meaningful (north_star)."""
    from emu_util import oracle_params, oracle_policy_step
    from oracle.env_oracle import EnvOracle
    from oracle.physics import PhysicsOracle
    g, s0, spec, terrain = golden
    N = 32
    env = _env(spec, N, terrain)
    sim = env.simulator
    sim.load_state(s0)
    env.common_step_counter = int(s0["common_step_counter"])
    model = sim._model
    orc = PhysicsOracle(model, oracle_params(spec, model), terrain[0], precision="f32")
    eo = EnvOracle(spec, N, terrain[0], terrain[1])
    for k, v in s0.items():
        if k in eo.st:
            eo.st[k][...] = v
    eo.common_step_counter = env.common_step_counter
    worst = 0.0
    warm = np.zeros((N, 48))
    for t in range(5):
        a = g["actions"][t]
        cur = {k: eo.st[k2] for k, k2 in (("base_pos", "base_pos"), ("base_quat_wxyz", "base_quat_wxyz"), ("base_lin_w", "base_lin_w"),
                                         ("base_ang_w", "base_ang_w"), ("dof_pos", "q"), ("dof_vel", "qd"), ("added_mass", "added_mass"),
                                         ("com_bias", "com_bias"), ("friction", "friction"), ("kp_scale", "kp_scale"), ("kd_scale", "kd_scale"),
                                         ("joint_armature", "joint_armature"), ("joint_damping", "joint_damping"), ("joint_friction", "joint_friction"))}
        cur["contact_warm"] = warm
        eo.pre_step(a)
        ref = oracle_policy_step(spec, model, orc, cur, a)
        o = eo.post_step({k: np.asarray(v, np.float32) for k, v in ref.items() if k not in ("ncontact", "contact_warm")})
        warm = ref["contact_warm"]
        warm[o["env_ids"]] = 0
        out = env.step(torch.from_numpy(a).cuda())
        same = ~o["reset_buf"]
        worst = max(worst, float(np.abs(out[0].cpu().numpy() - o["obs_buf"])[same].max()))
    assert worst < 5e-2, f"5-step rollout observation drift {worst:.3e}"


@pytest.mark.parametrize("zero_copy", ["2", "1", "0"])
def test_one_call_host_step_equals_the_three_call_step(golden, monkeypatch, zero_copy):
    """b200_env_step (pinned host actions in, rew / reset / time_out out, ONE C-ABI call) is bit-identical to the step
    made call by call (b200_dynamics_step, b200_env_post_step) -- with the pinned buffers read / written in place by the
    kernels (zero copy, the default) and with copy-engine transfers (B200_ZERO_COPY_*=0)."""
    monkeypatch.setenv("B200_ZERO_COPY_ACTIONS", zero_copy)
    monkeypatch.setenv("B200_ZERO_COPY_RESULTS", "0" if zero_copy == "0" else "1")
    g, s0, spec, terrain = golden
    N = 512
    rng = np.random.default_rng(2)
    a_env, b_env = _env(spec, N, terrain), _env(spec, N, terrain)
    a_env.reset(); b_env.reset()
    rew, rst, tmo = b_env.simulator.make_host_step_buffers()        # one slab -> one device->host copy per step
    rew2 = torch.empty(N, dtype=torch.float32).pin_memory()         # separate buffers -> three copies; same values
    rst2 = torch.empty(N, dtype=torch.bool).pin_memory()
    for t in range(6):
        act = torch.from_numpy(rng.normal(size=(N, spec.num_actions)).astype(np.float32)).pin_memory()
        out_a = a_env.step_two_kernels(act.cuda())
        out_b = b_env.step_host(act, rew, rst, tmo) if t % 2 == 0 else b_env.step_host(act, rew2, rst2, tmo)
        torch.cuda.synchronize()
        if t % 2:
            rew.copy_(rew2); rst.copy_(rst2)
        for x, y in zip(out_a[:6], out_b[:6]):
            assert torch.equal(x, y)
        assert torch.equal(rew, a_env.rew_buf.cpu()) and torch.equal(rst, a_env.reset_buf.cpu()) and torch.equal(tmo, a_env.time_out_buf.cpu())
    with pytest.raises(ValueError):
        b_env.step_host(torch.zeros(N, spec.num_actions))          # pageable host memory is refused, not silently staged


@pytest.mark.parametrize("N", [1, 5, 13, 130])
def test_ragged_env_counts_match_the_same_envs_in_a_full_block(golden, N):
    """Env counts that are no multiple of the CTA shapes (7 envs per dynamics CTA, 4 per env CTA, ragged last CTA not
    staged by TMA): the first N envs of a 132-env job computed alone are bit-identical (Philox is keyed by global id)."""
    g, s0, spec, terrain = golden
    big, small = _env(spec, 132, terrain), _env(spec, N, terrain, num_envs_global=132)
    big.reset(); small.reset()
    rng = np.random.default_rng(4)
    for t in range(4):
        a = torch.from_numpy(rng.normal(size=(132, spec.num_actions)).astype(np.float32)).cuda()
        ob = big.step(a)
        os_ = small.step(a[:N].contiguous())
        for x, y in zip(ob[:6], os_[:6]):
            assert torch.equal(x[:N], y), f"step {t}"
    for k in ("dof_pos", "base_pos", "episode_sums", "commands", "terrain_levels", "contact_warm"):
        assert torch.equal(big.simulator._buf[k][:N], small.simulator._buf[k]), k


def test_edited_config_runs_on_the_generic_instantiation(golden):
    """A descriptor that is not one of the built-in presets (here: one reward term switched off, a coarser height scan)
    must select the generic env kernel and still match the numpy oracle."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.env_oracle import EnvOracle
    g, s0, _, terrain = golden
    spec = T.go2_ts_spec()
    spec.reward_scales = dict(spec.reward_scales, dof_power=0.0)
    spec.measured_points_x = spec.measured_points_x[::2]
    N = 64
    env = _env(spec, N, terrain)
    sim = env.simulator
    assert sim.env_kernel_variant == "generic"
    assert _env(T.go2_ts_spec(), 8, terrain).simulator.env_kernel_variant == "go2_ts"
    st, _ = _random_state(spec, N, terrain, seed=21)
    sim.load_state(st)
    eo = EnvOracle(spec, N, *terrain)
    for k, v in sim.get_state().items():
        kk = {"dof_pos": "q", "dof_vel": "qd"}.get(k, k)
        if kk in eo.st:
            eo.st[kk][...] = v.reshape(eo.st[kk].shape)
    eo.st["obs_hist"][:] = 0
    eo.st["critic_hist"][:] = 0
    eo.common_step_counter = env.common_step_counter
    rng = np.random.default_rng(8)
    for t in range(3):
        a = rng.normal(size=(N, spec.num_actions)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda())
        mid = sim.get_state()
        eo.pre_step(a)
        o = eo.post_step({k: mid[name] for k, name in PH.items()})
        env.common_step_counter += 1
        sim.fused_post_step(env.common_step_counter, env.command_ranges["lin_vel_x"])
        out = sim.get_state()
        assert np.array_equal(out["reset_buf"].astype(bool), o["reset_buf"])
        assert np.array_equal(out["height_cells"], o["height_cells"])
        _close(out["rew_buf"], o["rew_buf"], what=f"step {t}: rew")
        _close(out["obs_buf"], o["obs_buf"], what=f"step {t}: obs")
        _close(out["critic_obs"], o["critic_obs_buf"], what=f"step {t}: critic")


@pytest.mark.parametrize("task", ["go2", "go2_ts", "go2_cat", "go2_wtw", "go2_cts", "go2_ee", "go2_dreamwaq", "tron1_pf", "tron1_pf_ee"])
def test_every_preset_steps_and_resets_with_ragged_ctas(task):
    """Every preset through all three step entry points at an env count that leaves ragged CTAs in both kernels, then a
    forced time-out of every env: finite outputs, all envs reset, the preset instantiation of the env kernel in use, and
    the return tuple of the reference's step for that task family."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    spec = T.PRESETS[task]()
    N = 45
    env = _env(spec, N, load_terrain(spec) if spec.heightfield else None)
    assert env.simulator.env_kernel_variant == task
    first = env.reset()
    g = torch.Generator(device="cpu").manual_seed(3)
    rew, rst, tmo = env.simulator.make_host_step_buffers()
    arity = {"go2": 5, "go2_wtw": 5, "tron1_pf": 5, "go2_ts": 7, "go2_cat": 7, "go2_cts": 7, "go2_ee": 6, "tron1_pf_ee": 6, "go2_dreamwaq": 8}[task]
    for t in range(3):
        a = torch.randn(N, spec.num_actions, generator=g)
        for out in (env.step(a.cuda()), env.step_host(a.pin_memory(), rew, rst, tmo), env.step_two_kernels(a.cuda())):
            assert len(out) == arity and len(first) == arity - 3
    env.episode_length_buf = torch.full((N,), int(env.max_episode_length), dtype=torch.int32)
    out = env.step(torch.zeros(N, spec.num_actions).cuda())
    torch.cuda.synchronize()
    assert bool(env.reset_buf.all()) and bool(env.time_out_buf.all()) and int(env.episode_length_buf.max()) == 0
    for x in out[:arity - 3]:
        assert x is None or torch.isfinite(x).all()
    assert torch.isfinite(env.rew_buf).all()


@pytest.mark.parametrize("task", ["go2_ts", "tron1_pf_ee"])
def test_long_closed_loop_rollout_tracks_the_oracle(task):
    """80 policy steps (320 physics substeps) of the closed loop -- dynamics kernel + env kernel, with resets, pushes and
    command resampling -- against the oracle pair fed the same actions.  With moderate actions (0.5 sigma) the loop does
    not amplify fp32 noise much, so far more than a statistical agreement can be asked: the same envs reset at the
    same steps, base positions stay within 1e-4 m, windowed mean rewards within 1e-4 absolute."""
    from emu_util import oracle_params, oracle_policy_step
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.env_oracle import EnvOracle
    from oracle.physics import PhysicsOracle
    spec = T.PRESETS[task]()
    terrain = load_terrain(spec) if spec.heightfield else None
    N, TT = 256, 80
    env = _env(spec, N, terrain)
    sim = env.simulator
    env.reset()
    for _ in range(5):
        env.step(torch.zeros(N, spec.num_actions).cuda())
    s0 = sim.get_state()
    model = sim._model
    orc = PhysicsOracle(model, oracle_params(spec, model), terrain[0] if terrain else None, precision="f32")
    eo = EnvOracle(spec, N, *(terrain if terrain else (None, None)))
    alias = {"dof_pos": "q", "dof_vel": "qd", "obs_history": "obs_hist", "critic_obs": "critic_hist"}
    for k, v in s0.items():
        kk = alias.get(k, k)
        if kk in eo.st:
            eo.st[kk][...] = v.reshape(eo.st[kk].shape)
    eo.common_step_counter = env.common_step_counter
    warm = s0["contact_warm"].astype(np.float64)
    rng = np.random.default_rng(0)
    same = np.ones(N, bool)
    rew_g, rew_o, resets = [], [], 0
    keys = (("base_pos", "base_pos"), ("base_quat_wxyz", "base_quat_wxyz"), ("base_lin_w", "base_lin_w"), ("base_ang_w", "base_ang_w"),
            ("dof_pos", "q"), ("dof_vel", "qd"), ("added_mass", "added_mass"), ("com_bias", "com_bias"), ("friction", "friction"),
            ("kp_scale", "kp_scale"), ("kd_scale", "kd_scale"), ("joint_armature", "joint_armature"), ("joint_damping", "joint_damping"),
            ("joint_friction", "joint_friction"))
    for t in range(TT):
        a = (rng.normal(size=(N, spec.num_actions)) * 0.5).astype(np.float32)
        cur = {k: eo.st[k2] for k, k2 in keys}
        cur["contact_warm"] = warm
        eo.pre_step(a)
        ref = oracle_policy_step(spec, model, orc, cur, a)
        o = eo.post_step({k: np.asarray(v, np.float32) for k, v in ref.items() if k not in ("ncontact", "contact_warm")})
        warm = ref["contact_warm"]
        warm[o["env_ids"]] = 0
        env.step(torch.from_numpy(a).cuda())
        same &= env.reset_buf.cpu().numpy() == o["reset_buf"]
        resets += int(o["reset_buf"].sum())
        rew_g.append(float(env.rew_buf.mean())); rew_o.append(float(o["rew_buf"].mean()))
    assert resets > 0
    assert same.mean() >= 0.99, f"only {same.sum()} of {N} envs reset at the same steps as the oracle"
    d = np.abs(sim.get_state()["base_pos"] - eo.st["base_pos"]).max(1)[same]
    assert np.percentile(d, 90) < 1e-4, f"base position drift p90 {np.percentile(d, 90):.2e} m after {TT} steps"
    for w in range(0, TT, 10):
        g, r = np.mean(rew_g[w:w + 10]), np.mean(rew_o[w:w + 10])
        assert abs(g - r) <= 1e-3 * abs(r) + 1e-4, f"steps {w}-{w + 9}: mean reward {g:.6f} vs {r:.6f}"   # rewards are O(0.03), means cancel


def test_dynamics_kernel_drives_joints_with_the_delayed_action():
    """randomize_ctrl_delay: joints are driven with action_queue[env, action_delay[env]] after the push of the new
    action (legged_robot.py:240-245); compared with the oracle substep fed that action."""
    from emu_util import oracle_params, oracle_policy_step
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.physics import PhysicsOracle
    spec = T.go2_spec()
    spec.randomize_ctrl_delay, spec.ctrl_delay_step_range = True, [0, 3]
    N, A, depth = 128, spec.num_actions, 4
    st, model = _random_state(spec, N, None, seed=17)
    rng = np.random.default_rng(6)
    queue = rng.normal(size=(N, depth, A)).astype(np.float32)
    delay = rng.integers(0, depth, N).astype(np.int32)
    st["action_queue"], st["action_delay"] = queue.reshape(N, -1), delay
    env = _env(spec, N, None)
    sim = env.simulator
    assert sim.env_kernel_variant == "generic"
    sim.load_state(st)
    full = sim.get_state()
    actions = rng.normal(size=(N, A)).astype(np.float32)
    sim.step(torch.from_numpy(actions).cuda())
    out = sim.get_state()
    pushed = np.concatenate([actions[:, None], queue[:, :-1]], axis=1)
    assert np.array_equal(out["action_queue"].reshape(N, depth, A), pushed)
    assert np.array_equal(out["actions"], actions)                       # the undelayed action feeds obs / rewards
    applied = pushed[np.arange(N), delay]
    assert (delay > 0).any() and not np.array_equal(applied, actions)
    orc = PhysicsOracle(model, oracle_params(spec, model), None, precision="f32")
    ref = oracle_policy_step(spec, model, orc, full, applied)
    for k, name in (("q", "dof_pos"), ("qd", "dof_vel"), ("torques", "torques"), ("base_pos", "base_pos")):
        r = np.asarray(ref[k], np.float64)
        assert np.abs(out[name].reshape(r.shape) - r).max() < 1e-4 * max(1.0, np.abs(r).max()), k


def test_nonfinite_state_is_contained_and_reset():
    """Failure containment (no reference counterpart): an env whose state goes non-finite inside the dynamics
    kernel is not written back -- it keeps its pose, at rest, is flagged, and the env kernel resets it; the other envs of the
    batch (same CTA included) are bit-identical to a run without the fault."""
    from hcr_genesis_lr_cl_b200 import task_spec as T
    spec = T.go2_ts_spec()
    terrain = load_terrain(spec)
    N = 64
    env, ref = _env(spec, N, terrain), _env(spec, N, terrain)
    for e in (env, ref):
        e.reset()
        e.step(torch.zeros(N, 12, device="cuda"))
    b = env.simulator._buf
    bad = [5, 33]
    b["dof_vel"][5, 7] = float("nan")
    b["base_lin_w"][33, 0] = float("inf")
    a = 0.3 * torch.ones(N, 12, device="cuda")
    env.step(a)
    ref.step(a)
    ok = [i for i in range(N) if i not in bad]
    assert env.simulator.nonfinite_resets == 2 and bool(env.reset_buf[bad].all())
    for k in ("base_pos", "base_quat_wxyz", "dof_pos", "dof_vel", "base_lin_w", "obs_buf", "rew_buf", "privileged_obs_buf"):
        assert torch.isfinite(b[k]).all(), k
        assert torch.equal(b[k][ok], ref.simulator._buf[k][ok]), k
    assert int(b["nonfinite"].sum()) == 0 and int(b["episode_length"][bad].sum()) == 0
    for _ in range(3):
        env.step(a)
    assert torch.isfinite(b["obs_buf"]).all() and torch.isfinite(env.obs_history).all() and env.simulator.nonfinite_resets == 2
