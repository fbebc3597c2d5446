"""Plugin mode: the reference's UNMODIFIED task classes (GO2, Go2TS, GO2WTW, Go2CaT, TRON1PF_EE, ...) stepping over
``B200Simulator`` through the ``Simulator`` plugin API (north_star sentence 2), compared step by step with
``FusedLeggedEnv`` started from the same state.

Both runs use the same dynamics kernel (b200_simulator_step vs b200_dynamics_step); everything else is the reference's
torch code over the backend's properties on one side and the fused env kernel on the other.  Every random draw of the
reference is redirected to the Philox stream (oracle/inject.py); the backend's own draws (reset DR, pushes, curriculum
levels) come from that stream natively.  Tolerances as in tests/test_gpu_parity.py: integers / bools exact, floats 1e-4
relative with a 1e-5 floor (the closed loop runs >= 20 policy steps; physics state gets a wider absolute floor because a
1-ulp difference in a reset pose is carried through the contact solver).

CPU variant: B200Simulator over tests/warp_emu/libb200step_emu.so (csrc/b200_step.cu compiled for the warp emulator).
GPU variant (`-m gpu`): the real library on cuda:0; the reference tree is the staged copy baseline/_ref."""
import numpy as np
import pytest
import torch

from plugin_util import make_plugin_env, reference_root, transfer_state, uninstall_plugin

needs_reference = pytest.mark.skipif(reference_root() is None, reason="no reference tree (/root/reference or baseline/_ref)")

INT_KEYS = ("reset_buf", "time_out_buf", "episode_length", "fail_buf", "last_contacts", "terrain_levels")


def _close(a, b, what, rtol=1e-4, atol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.asarray(a).shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f"{what}: max violation {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)} ({a.flat[err.argmax()]} vs {b.flat[err.argmax()]})"


def _n(t):
    return t.detach().float().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _ref_outputs(env, ret, task, sum_names):
    """What the reference step returned / left behind, keyed like the fused env's buffers."""
    sim = env.simulator
    o = dict(rew_buf=_n(env.rew_buf), reset_buf=_n(env.reset_buf), time_out_buf=_n(env.time_out_buf), commands=_n(env.commands),
             episode_length=_n(env.episode_length_buf), fail_buf=_n(env.fail_buf), feet_air_time=_n(env.feet_air_time),
             last_contacts=_n(env.last_contacts), actions=_n(env.actions), last_actions=_n(env.last_actions),
             llast_actions=_n(env.llast_actions),
             episode_sums=np.stack([_n(env.episode_sums[k]) for k in sum_names], axis=1))
    for k in ("base_pos", "dof_pos", "dof_vel", "base_lin_vel", "base_ang_vel", "projected_gravity", "friction", "added_mass", "com_bias",
              "kp_scale", "kd_scale", "rand_push_vels", "env_origins", "terrain_levels", "measured_heights", "link_contact_forces",
              "feet_pos", "feet_vel", "base_quat_wxyz", "base_lin_w", "base_ang_w", "height_around_feet", "link_contact_states"):
        o[k] = _n(sim._buf[k])
    if hasattr(env, "cstr_prob"):
        o["cstr_prob"] = _n(env.cstr_prob)
    return o


def _fused_returns(env):
    """obs-like tensors of the fused step in the order the reference's step returns them."""
    out = env._returns()
    return [x for x in out[:-1] if isinstance(x, torch.Tensor)]


def _run(task, impl, fused_cls, cpu, N=16, steps=24, cfg_edit=None, skip_env0=False):
    from hcr_genesis_lr_cl_b200 import task_spec as T
    from oracle.inject import Injector
    dev = "cpu" if cpu else "cuda:0"
    g = torch.Generator().manual_seed(4321)
    try:
        ref, cfg, _ = make_plugin_env(task, N, impl=impl, cpu=cpu, cfg_edit=cfg_edit)
        sim = ref.simulator
        assert not sim.fused and sim.spec.task == T.PLUGIN_TASK
        spec = T.TaskSpec.from_reference_cfg(cfg, task)
        A = ref.num_actions
        with Injector(ref, spec.seed):
            ref.reset()
            for _ in range(6):
                ref.step((0.8 * torch.randn(N, A, generator=g)).to(dev))
            # spread the interesting events over the compared window (as tools/make_golden.py does)
            L = int(ref.max_episode_length)
            ep = ref.episode_length_buf
            ep[:] = torch.randint(10, 400, (N,), generator=g).to(ep.dtype).to(dev)
            ep[0::8] = L - 3 - (torch.arange(len(ep[0::8])) % 3).to(ep.dtype).to(dev)            # time-outs
            ep[1::8] = 500 - 2 - (torch.arange(len(ep[1::8])) % 4).to(ep.dtype).to(dev)          # command resampling
            ref.fail_buf[2::8] = 3
            b = sim._buf
            b["base_quat_wxyz"][2::8] = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev)           # upside down -> fails -> reset
            b["base_pos"][2::8, 2] += 0.3 if A == 12 else 0.8
            ref.common_step_counter = sim._step_counter = int(spec.push_interval) * 3 - steps // 2   # a push inside the window
            if hasattr(ref, "num_gaits"):
                ref.num_gaits = ref.num_gait_max
                ref.gait_period_range, ref.base_height_target_range = [0.35, 0.55], [0.24, 0.32]
                ref.foot_clearance_target_range, ref.pitch_target_range = [0.05, 0.10], [-0.2, 0.2]
                ep[4::8] = 250 - 2 - (torch.arange(len(ep[4::8])) % 4).to(ep.dtype).to(dev)
            ref.step(torch.zeros(N, A, device=dev))                                                 # API buffers consistent with the edits
            sum_names = spec.episode_sum_names()          # the reference's dict keys, "termination" moved behind the other terms
            assert sorted(sum_names) == sorted(ref.episode_sums.keys()), (sum_names, list(ref.episode_sums.keys()))
            terrain = (sim._height_samples.cpu().numpy(), sim._terrain_origins.cpu().numpy()) if spec.heightfield else None
            fused = fused_cls(spec, N, dev, terrain=terrain)
            transfer_state(ref, fused, sum_names)
            n_reset = 0
            sl = slice(1, None) if skip_env0 else slice(None)
            for t in range(steps):
                a = (0.8 * torch.randn(N, A, generator=g)).to(dev)
                if t == 1:
                    a[0] = 150.0                                                                    # exercises clip_actions
                r_ret = ref.step(a.clone())
                fused.step(a.clone())
                o = _ref_outputs(ref, r_ret, task, sum_names)
                fb = fused.simulator.get_state()
                for k, rv in o.items():
                    if k not in fb:
                        continue
                    fv = np.asarray(fb[k]).reshape(rv.shape)
                    if k in INT_KEYS:
                        assert np.array_equal(fv[sl].astype(np.int64), rv[sl].astype(np.int64)), f"{task} step {t}: {k} not bit-exact"
                    else:
                        # physics state: a 1-ulp difference in a reset pose is amplified by the contact solve
                        phys = k in ("base_pos", "dof_pos", "dof_vel", "base_lin_vel", "base_ang_vel", "projected_gravity", "link_contact_forces",
                                     "feet_pos", "feet_vel", "base_quat_wxyz", "base_lin_w", "base_ang_w")
                        _close(fv[sl], rv[sl], f"{task} step {t}: {k}", atol=2e-4 if phys else 1e-5,
                               rtol=2e-3 if k == "link_contact_forces" else 1e-4)
                # the tensors the two step() calls hand to the runner, in order
                r_obs = [x for x in r_ret[:-1] if isinstance(x, torch.Tensor) and x.dtype == torch.float32 and x.dim() == 2]
                f_obs = [x for x in _fused_returns(fused) if x.dtype == torch.float32 and x.dim() == 2]
                assert len(r_obs) == len(f_obs), (len(r_obs), len(f_obs))
                for i, (rv, fv) in enumerate(zip(r_obs, f_obs)):
                    assert rv.shape == fv.shape, (task, i, rv.shape, fv.shape)
                    _close(_n(fv)[sl], _n(rv)[sl], f"{task} step {t}: returned tensor {i}", atol=5e-4)
                n_reset += int(o["reset_buf"].sum())
            assert n_reset > 0, "the window held no reset"
            return ref, fused
    finally:
        uninstall_plugin()


TASKS = ["go2", "go2_ts", "go2_wtw", "go2_cat", "tron1_pf_ee"]


@needs_reference
@pytest.mark.parametrize("task", TASKS)
def test_reference_task_classes_run_over_the_backend_emulated(task):
    """CPU: the host layer + the kernels' own sources on the warp emulator."""
    from emu_backend import EmuB200Simulator, EmuFusedLeggedEnv
    _run(task, EmuB200Simulator, EmuFusedLeggedEnv, cpu=True, N=16, steps=20, skip_env0=task in ("go2_wtw", "tron1_pf_ee"))


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("task", TASKS + ["go2_cts", "go2_ee", "go2_dreamwaq", "tron1_pf"])
def test_reference_task_classes_run_over_the_backend_gpu(task):
    """B200: stock LeggedRobot / task classes over libb200step.so vs FusedLeggedEnv, 24 policy steps, 64 envs."""
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    _run(task, None, FusedLeggedEnv, cpu=False, N=64, steps=24, skip_env0=task in ("go2_wtw", "tron1_pf_ee"))


# ---------------------------------------------------------------------------------------------------------------------
# The two ways of wiring the backend in, each in a fresh interpreter
def _smoke(variant, impl, task, n_envs=8, steps=6):
    import json
    import os
    import shutil
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = reference_root()
    tmp = None
    try:
        if variant == "patched":
            tmp = tempfile.mkdtemp(prefix="b200_overlay_")
            tree = os.path.join(tmp, "ref")
            shutil.copytree(ref, tree, ignore=shutil.ignore_patterns("*.STL", "*.stl", "*.dae", "*.obj", "*.png", "*.jpg", ".git"))
            subprocess.run(["patch", "-p1", "-s", "-i", os.path.join(root, "overlay", "b200_backend.patch")], cwd=tree, check=True)
            shutil.copy(os.path.join(root, "overlay", "legged_gym", "simulator", "b200_simulator.py"), os.path.join(tree, "legged_gym", "simulator"))
            ref = tree
        env = dict(os.environ)
        env.pop("SIMULATOR", None)
        out = subprocess.run([sys.executable, os.path.join(root, "tests", "run_backend_smoke.py"), ref, variant, impl, task, str(n_envs), str(steps)],
                             capture_output=True, text=True, env=env, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("DIGEST ")][-1]
        return json.loads(line[7:])
    finally:
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)


def _check_wiring(impl, task):
    hook = _smoke("hook", impl, task)
    patched = _smoke("patched", impl, task)
    # import-hook variant: unmodified tree, SIMULATOR=genesis branches, the placeholder stands in for the absent engine
    assert hook["simulator"] == "genesis" and hook["backend"] == "B200Simulator" and hook["is_simulator_abc"] and hook["finite"]
    assert "placeholder" in hook["genesis"]
    # maintainer variant: patch + overlay file, SIMULATOR=b200, no `genesis` module at all
    assert patched["simulator"] == "b200" and patched["backend"] == "B200Simulator" and patched["is_simulator_abc"] and patched["finite"]
    assert patched["backend_module"] == "legged_gym.simulator.b200_simulator" and patched["genesis"] == ""
    assert hook["mesh_type"] == patched["mesh_type"]
    assert hook["sha"] == patched["sha"], "the two wirings step differently"
    assert hook["launches"] == patched["launches"] > 0


@needs_reference
@pytest.mark.parametrize("task", ["go2_ts", "tron1_pf_ee"])
def test_overlay_patch_and_import_hook_wire_the_same_backend_emulated(task):
    _check_wiring("emu", task)


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("task", ["go2_ts", "tron1_pf_ee"])
def test_overlay_patch_and_import_hook_wire_the_same_backend_gpu(task):
    _check_wiring("cuda", task)
