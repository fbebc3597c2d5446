"""ctypes front-end of the CPU physics oracle (oracle/rigid_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
P_DT, P_GRAV, P_TC, P_DAMPRATIO, P_D0, P_DMAX, P_WIDTH, P_MID, P_POWER, P_TERRAIN_MU, P_GEOM_MU, P_ITERS, \
    P_HSCALE, P_VSCALE, P_BORDER, P_TOL, P_TRIMESH = range(17)


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "rigid_oracle.c")
    for name in ("liboracle_f64.so", "liboracle_f32.so"):
        out = os.path.join(_HERE, name)
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, name])


def default_params(dt=0.005, iters=30, hscale=0.1, vscale=0.005, border=0.0, terrain_mu=1.0, geom_mu=1.0, tol=1e-4, trimesh=False) -> np.ndarray:
    p = np.zeros(20, np.float32)
    p[P_DT], p[P_GRAV] = dt, 9.81
    p[P_TC], p[P_DAMPRATIO] = 2 * dt, 1.0
    p[P_D0], p[P_DMAX], p[P_WIDTH], p[P_MID], p[P_POWER] = 0.9, 0.95, 0.001, 0.5, 2.0
    p[P_TERRAIN_MU], p[P_GEOM_MU], p[P_ITERS] = terrain_mu, geom_mu, iters
    p[P_HSCALE], p[P_VSCALE], p[P_BORDER] = hscale, vscale, border
    p[P_TOL] = tol
    p[P_TRIMESH] = 1.0 if trimesh else 0.0
    return p


class PhysicsOracle:
    """Batched substep over ``n`` envs; all state arrays are float64 [n, k] numpy arrays."""

    def __init__(self, model, params: np.ndarray, heightfield: np.ndarray | None = None, precision: str = "f64"):
        build()
        self.lib = ctypes.CDLL(os.path.join(_HERE, f"liboracle_{precision}.so"))
        self.model = model
        self.mi = model.packed_ints()
        self.mf = model.packed_floats()
        self.prm = np.ascontiguousarray(params, np.float32)
        if heightfield is None:
            self.hf, self.rows, self.cols = np.zeros(1, np.int16), 0, 0
        else:
            self.hf = np.ascontiguousarray(heightfield, np.int16)
            self.rows, self.cols = self.hf.shape

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def substep(self, state, q, qd, tau, envp, jparam, warm=None):
        """In-place substep. Returns (link_force [n, nlinks, 3], ncontact [n]).  `warm` [n, 48] float64 (in/out) carries the
        contact-solver warm start between substeps; None = cold start."""
        n = state.shape[0]
        for a in (state, q, qd):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        tau = np.ascontiguousarray(tau, np.float64)
        envp = np.ascontiguousarray(envp, np.float64)
        jparam = np.ascontiguousarray(jparam, np.float64)
        lf = np.zeros((n, self.model.nlinks, 3), np.float64)
        nc = np.zeros(n, np.int32)
        if warm is None:
            warm = np.zeros((n, 48), np.float64)
        assert warm.dtype == np.float64 and warm.flags.c_contiguous and warm.shape == (n, 48)
        self.lib.oracle_substep(self._p(self.mi), self._p(self.mf), self._p(self.prm), self._p(self.hf),
                                ctypes.c_int(self.rows), ctypes.c_int(self.cols), ctypes.c_int(n),
                                self._p(state), self._p(q), self._p(qd), self._p(tau), self._p(envp), self._p(jparam),
                                self._p(lf), self._p(nc), self._p(warm))
        return lf, nc

    def link_kinematics(self, state, q, qd):
        n = state.shape[0]
        lp = np.zeros((n, self.model.nlinks, 3), np.float64)
        lv = np.zeros((n, self.model.nlinks, 3), np.float64)
        self.lib.oracle_link_kinematics(self._p(self.mi), self._p(self.mf), ctypes.c_int(n),
                                        self._p(np.ascontiguousarray(state, np.float64)),
                                        self._p(np.ascontiguousarray(q, np.float64)),
                                        self._p(np.ascontiguousarray(qd, np.float64)), self._p(lp), self._p(lv))
        return lp, lv


def terrain_query(prm, hf, xy, lib=None):
    """Height and unit normal of the oracle's collision surface under world points `xy` [n, 2] (float64)."""
    build()
    lib = lib or ctypes.CDLL(os.path.join(_HERE, "liboracle_f64.so"))
    hf = np.ascontiguousarray(hf, np.int16)
    xy = np.ascontiguousarray(xy, np.float64)
    h, nrm = np.zeros(len(xy)), np.zeros((len(xy), 3))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.oracle_terrain_query(vp(np.ascontiguousarray(prm, np.float32)), vp(hf), ctypes.c_int(hf.shape[0]), ctypes.c_int(hf.shape[1]),
                             ctypes.c_int(len(xy)), vp(xy), vp(h), vp(nrm))
    return h, nrm


def oracle_params(spec, model=None, hf=None):
    return default_params(dt=spec.sim_dt, iters=spec.pgs_iterations, hscale=spec.horizontal_scale,
                                vscale=spec.vertical_scale, border=spec.border_size if spec.heightfield else 0.0,
                                terrain_mu=spec.static_friction, geom_mu=1.0, tol=spec.pgs_tolerance,
                                trimesh=spec.mesh_type == "trimesh")


def env_params(spec, st) -> np.ndarray:
    """Per-env engine parameters [added base mass, COM shift xyz, geom friction ratio] as the reference hands them to the
    engine: `set_mass_shift` / `set_friction_ratio` are only ever called from `_randomize_base_mass` / `_randomize_friction`
    (genesis_simulator.py:62-82,384-405,665-697), i.e. with the DR switch off the engine keeps mass shift 0 and ratio 1
    whatever the observation-side buffers (`_added_base_mass` starts at ones, `_friction_values` at zeros, :644-650) hold."""
    n = st["added_mass"].shape[0]
    mass = st["added_mass"] if spec.randomize_base_mass else np.zeros((n, 1))
    fric = st["friction"] if spec.randomize_friction else np.ones((n, 1))
    # the COM shift needs no switch: `_base_com_bias` starts at zeros and only _randomize_com_displacement writes it
    return np.concatenate([mass, st["com_bias"], fric], axis=1).astype(np.float64)


def oracle_policy_step(spec, model, oracle, st, actions, clip=True):
    """Reference statement of one decimated physics step (genesis_simulator.py:20-33) on a state dict `st` of fp32
    arrays (keys as B200Buffers).  Returns the post-physics dict the env half consumes (float64 internally).
    `clip=False`: the actions are what LeggedRobot._pre_sim_step hands over (plugin mode: already clipped / delayed)."""
    f32 = np.float32
    N, A = actions.shape
    a = np.clip(actions.astype(f32), f32(-spec.clip_actions), f32(spec.clip_actions)) if clip else actions.astype(f32)
    state = np.concatenate([st["base_pos"], st["base_quat_wxyz"], st["base_lin_w"], st["base_ang_w"]], axis=1).astype(np.float64)
    q, qd = st["dof_pos"].astype(np.float64).copy(), st["dof_vel"].astype(np.float64).copy()
    envp = env_params(spec, st)
    arm = np.tile(model.body[1:, 19][None, :], (N, 1)).astype(np.float64)
    if spec.randomize_joint_armature:
        arm = np.tile(st["joint_armature"], (1, A)).astype(np.float64)
    dmp = np.tile(st["joint_damping"], (1, A)).astype(np.float64) if spec.randomize_joint_damping else np.zeros((N, A))
    fls = np.tile(st["joint_friction"], (1, A)).astype(np.float64) if spec.randomize_joint_friction else np.zeros((N, A))
    jp = np.concatenate([arm, dmp, fls], axis=1)
    q0 = np.asarray(spec.default_dof_pos, f32)[None, :]
    kp, kd = (st["kp_scale"] * f32(spec.kp)).astype(f32), (st["kd_scale"] * f32(spec.kd)).astype(f32)
    tgt = (a * f32(spec.action_scale) + q0).astype(f32)
    tau = np.zeros((N, A), f32)
    lf = nc = None
    warm = np.ascontiguousarray(st.get("contact_warm", np.zeros((N, 48))), np.float64).copy()
    for _ in range(spec.decimation):
        tau = (kp * (tgt - q.astype(f32)) - kd * qd.astype(f32)).astype(f32)
        lf, nc = oracle.substep(state, q, qd, tau, envp, jp, warm)
    lp, lv = oracle.link_kinematics(state, q, qd)
    feet = spec.link_groups(model)[0]
    return dict(base_pos=state[:, 0:3], base_quat_wxyz=state[:, 3:7], base_lin_w=state[:, 7:10], base_ang_w=state[:, 10:13],
                q=q, qd=qd, torques=tau, link_force=lf, feet_pos=lp[:, feet], feet_vel=lv[:, feet], ncontact=nc, contact_warm=warm)
