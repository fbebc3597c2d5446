"""Replace the reference's random-number call sites by the counter-based Philox stream of the fused kernels.

TEST INFRASTRUCTURE (used by tools/make_golden.py and tests/test_plugin_mode.py).  The reference draws from torch's
global generator in data-dependent order and sizes, which no fused kernel can reproduce; for parity runs
every draw site of the imported, unmodified reference classes is redirected to
u(seed, step, env, site, idx) (oracle/philox.py == csrc/philox.cuh).  ``install()`` patches module-level functions of
torch / numpy / legged_gym; ``uninstall()`` puts every patched name back (a pytest process runs other tests afterwards).
"""
from __future__ import annotations

import sys

import numpy as np

from hcr_genesis_lr_cl_b200 import task_spec as T
from oracle import philox


class Injector:
    """Replaces the reference's RNG call sites by philox draws keyed (seed, step, env, site, idx)."""

    def __init__(self, env, seed, env_offset=0):
        import torch
        self.torch, self.env, self.sim, self.seed = torch, env, env.simulator, seed
        self.site, self.ids, self.col, self.call, self.in_reset = None, None, 0, 0, False
        self.N = env.num_envs
        self.env_offset = env_offset
        self._undo = []

    def _patch(self, obj, name, value):
        """setattr that uninstall() reverts (instance attributes that shadowed a class method are deleted again)."""
        had = name in getattr(obj, "__dict__", {}) or not hasattr(obj, "__dict__")
        self._undo.append((obj, name, getattr(obj, name, None), had))
        setattr(obj, name, value)

    def uninstall(self):
        for obj, name, old, had in reversed(self._undo):
            if had:
                setattr(obj, name, old)
            else:
                try:
                    delattr(obj, name)
                except AttributeError:
                    pass
        self._undo.clear()

    def __enter__(self):
        self.install()
        return self

    def __exit__(self, *exc):
        self.uninstall()

    def u(self, site, ids, cols):
        step = self.env.common_step_counter
        return philox.uniform(self.seed, step, np.asarray(ids)[:, None] + self.env_offset, site, np.asarray(cols)[None, :])

    def draw(self, shape):
        assert self.site is not None, "random draw outside a known site"
        k = shape[0]
        ncol = shape[1] if len(shape) > 1 else 1
        site = self.site
        if site == T.SITE_DOF and ncol * 3 == self.env.num_actions:
            cols = 3 * np.arange(ncol) + self.call         # go2_ts.py:86-91 / tron1_pf_ee.py:276-281: hips, thighs, calves
        elif site == T.SITE_KP:
            site, cols = (T.SITE_KP if self.call == 0 else T.SITE_KD), np.arange(ncol)
        elif site == T.SITE_ROOT:                      # fixed slots: xy -> 0,1; lin vel -> 2..4; ang vel -> 5..7 (legged_robot.py:283-298)
            if ncol == 2:
                cols = np.arange(2)
            elif ncol == 1:                                # torch_rand_float(c, c, ...) constants of the sit pose (tron1_pf_ee.py:294-300)
                cols = np.arange(1) + 8
            else:
                cols = 2 + 3 * self.col + np.arange(3)
                self.col += 1
        else:
            cols = self.col + np.arange(ncol)
            self.col += ncol
        self.call += 1
        ids = self.ids
        assert len(ids) == k, (len(ids), shape, site)
        return self.torch.from_numpy(self.u(site, ids, cols).reshape(shape).astype(np.float32)).to(self.env.device)

    def ctx(self, fn, site, ids_arg=True, name=None):
        inj = self

        def wrapped(*a, **kw):
            prev = (inj.site, inj.ids, inj.col, inj.call)
            ids = a[0] if ids_arg else np.arange(inj.N)
            ids = ids.cpu().numpy() if hasattr(ids, "cpu") else np.asarray(ids)
            s = site
            if name == "resample":
                s = T.SITE_CMD_RESET if inj.in_reset else T.SITE_CMD_RESAMPLE
            if name == "behavior":
                s = T.SITE_BEHAVIOR_RESET if inj.in_reset else T.SITE_BEHAVIOR
            inj.site, inj.ids, inj.col, inj.call = s, ids, 0, 0
            try:
                return fn(*a, **kw)
            finally:
                inj.site, inj.ids, inj.col, inj.call = prev
        return wrapped

    def install(self):
        torch, env, sim = self.torch, self.env, self.sim
        import genesis as gs

        def rand_float(lower, upper, shape, device="cpu"):
            return (upper - lower) * self.draw(tuple(shape)) + lower

        for mod in list(sys.modules.values()):
            if mod is not None and getattr(mod, "__name__", "").startswith("legged_gym") and hasattr(mod, "torch_rand_float"):
                self._patch(mod, "torch_rand_float", rand_float)
        self._patch(gs, "_rand_hook", lambda shape: self.draw(shape))
        self._patch(torch, "rand_like", lambda t, **k: self.draw(tuple(t.shape)))
        self._patch(torch, "rand", lambda *shape, **k: self.draw(tuple(shape[0]) if isinstance(shape[0], (tuple, list)) else tuple(shape)))

        def randint_like(t, high, **k):
            return torch.clamp((self.draw(tuple(t.shape)) * high).to(t.dtype), max=high - 1)
        self._patch(torch, "randint_like", randint_like)
        # tron1_pf_ee.py:204: one host coin per reset batch (R8)
        self._patch(np.random, "random", lambda *a, **k: float(philox.uniform(self.seed, self.env.common_step_counter, 0xFFFFFFFF, T.SITE_HOST, 0)))

        self._patch(env, "_resample_commands", self.ctx(env._resample_commands, None, name="resample"))
        if hasattr(env, "_resample_behavior_params"):
            self._patch(env, "_resample_behavior_params", self.ctx(env._resample_behavior_params, None, name="behavior"))

            def randint(low, high, size, **k):              # go2_wtw.py:205-206: one gait index per call (R7)
                u = float(philox.uniform(self.seed, self.env.common_step_counter, 0xFFFFFFFF, T.SITE_HOST, 2 if self.in_reset else 1))
                return torch.tensor([min(int(u * high), high - 1)])
            self._patch(torch, "randint", randint)
        elif getattr(env.cfg.domain_rand, "randomize_ctrl_delay", False):
            def randint_delay(low, high, size, **k):        # legged_robot.py:147-148: per-env action delay on reset
                prev = (self.site, self.col, self.call)
                self.site, self.col, self.call = T.SITE_CTRL_DELAY, 0, 0
                try:
                    u = self.draw(tuple(size))
                finally:
                    self.site, self.col, self.call = prev
                return low + torch.clamp((u * (high - low)).to(torch.long), max=high - low - 1)
            self._patch(torch, "randint", randint_delay)
        self._patch(env, "_reset_dofs", self.ctx(env._reset_dofs, T.SITE_DOF))
        self._patch(env, "_reset_root_states", self.ctx(env._reset_root_states, T.SITE_ROOT))
        if hasattr(env, "_reset_root_states_sit_pose"):
            self._patch(env, "_reset_root_states_sit_pose", self.ctx(env._reset_root_states_sit_pose, T.SITE_ROOT))
            self._patch(env, "_reset_dofs_sit_pose", self.ctx(env._reset_dofs_sit_pose, T.SITE_DOF))
        self._patch(env, "compute_observations", self.ctx(env.compute_observations, T.SITE_OBS_NOISE, ids_arg=False))
        self._patch(sim, "update_terrain_curriculum", self.ctx(sim.update_terrain_curriculum, T.SITE_LEVEL))
        self._patch(sim, "push_robots", self.ctx(sim.push_robots, T.SITE_PUSH, ids_arg=False))
        for nm, site in (("_randomize_friction", T.SITE_FRICTION), ("_randomize_base_mass", T.SITE_MASS),
                         ("_randomize_com_displacement", T.SITE_COM), ("_randomize_pd_gain", T.SITE_KP),
                         ("_randomize_joint_armature", T.SITE_ARMATURE), ("_randomize_joint_friction", T.SITE_JFRICTION),
                         ("_randomize_joint_damping", T.SITE_JDAMPING)):
            self._patch(sim, nm, self.ctx(getattr(sim, nm), site))
        orig_reset = env.reset_idx

        def reset_idx(env_ids):
            self.in_reset = True
            prev = (self.site, self.ids, self.col, self.call)
            ids = env_ids.cpu().numpy() if hasattr(env_ids, "cpu") else np.asarray(env_ids)
            self.site, self.ids, self.col, self.call = T.SITE_GAIT, ids, 0, 0     # direct torch.rand calls (gait phase)
            try:
                return orig_reset(env_ids)
            finally:
                self.in_reset = False
                self.site, self.ids, self.col, self.call = prev
        self._patch(env, "reset_idx", reset_idx)
