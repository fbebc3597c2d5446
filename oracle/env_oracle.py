"""CPU restatement (numpy, fp32) of the *environment half* of the hot path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg; never by hcr_genesis_lr_cl_b200.  Pinned against golden vectors recorded from the reference's own
Python code (tools/make_golden.py -> tests/golden/*.npz, checked in tests/test_env_oracle_golden.py).

Each method cites the reference lines it restates:
    pre_step            LeggedRobot._pre_sim_step            legged_robot.py:230-252
                        GenesisSimulator.step (last_* copy)  genesis_simulator.py:21-24
    pd_torque           _compute_torques                     genesis_simulator.py:630-642
    post_step           LeggedRobot.post_physics_step        legged_robot.py:55-76
      _sim_post         GenesisSimulator.post_physics_step   genesis_simulator.py:35-60
      _heights          _update_surrounding_heights          genesis_simulator.py:552-577
      _feet_terrain     _calc_terrain_info_around_feet       genesis_simulator.py:579-610
      _callback         _post_physics_step_callback          legged_robot.py:300-315 (+ push 150-158)
      _termination      check_termination                    legged_robot.py:78-92
      _reward           compute_reward + _reward_*           legged_robot.py:150-168,458-608; go2_ts.py:133-176
      _reset            reset_idx and what it calls          legged_robot.py:94-148,254-298,317-334;
                                                             go2_ts.py:75-96; genesis_simulator.py:62-148,665-739;
                                                             legged_robot_ts.py:120-125
      _observe          compute_observations                 go2.py:40-90, go2_ts.py:5-84
Random draws come from oracle/philox.py keyed (seed, step, env, site, idx) -- see SURVEY section 8d.
"""
from __future__ import annotations

import numpy as np

from hcr_genesis_lr_cl_b200 import task_spec as T  # data format (TaskSpec, site ids) only
from oracle import philox

f32 = np.float32
G_TH, G_GT, G_PHI, G_PER, G_BH, G_FC, G_PT, G_CLK = 0, 4, 5, 6, 7, 8, 9, 10      # columns of gait_state


def _rot_inv(q, v):
    """quat_rotate_inverse, q xyzw (math_utils.py:63-76)."""
    w = q[:, 3:4]
    qv = q[:, :3]
    a = v * (f32(2.0) * w * w - f32(1.0))
    b = np.cross(qv, v) * (f32(2.0) * w)
    c = qv * (f32(2.0) * np.sum(qv * v, axis=1, keepdims=True, dtype=f32))
    return (a - b + c).astype(f32)


def _quat_apply(q, v):
    """quat_apply, q xyzw (math_utils.py:33-40)."""
    xyz = q[:, :3]
    t = np.cross(xyz, v) * f32(2.0)
    return (v + q[:, 3:4] * t + np.cross(xyz, t)).astype(f32)


def _euler_xyz(q):
    """get_euler_xyz (math_utils.py:89-109)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    sinr = f32(2.0) * (w * x + y * z)
    cosr = w * w - x * x - y * y + z * z
    roll = np.arctan2(sinr, cosr)
    sinp = f32(2.0) * (w * y - z * x)
    pitch = np.where(np.abs(sinp) >= 1, np.abs(f32(np.pi / 2.0)) * np.sign(sinp), np.arcsin(np.clip(sinp, -1, 1)))
    siny = f32(2.0) * (w * z + x * y)
    cosy = w * w + x * x - y * y - z * z
    yaw = np.arctan2(siny, cosy)
    return np.stack([roll, pitch, yaw], axis=-1).astype(f32)


def _norm3(v):
    """torch.norm(dim=-1) of 3-vectors in fp32."""
    return np.sqrt(np.sum(v * v, axis=-1, dtype=f32)).astype(f32)


class EnvOracle:
    def __init__(self, spec: T.TaskSpec, num_envs: int, height_samples=None, terrain_origins=None):
        self.spec = s = spec
        self.N = N = num_envs
        self.model = m = spec.load_model()
        self.A = A = spec.num_actions
        self.feet, self.pen, self.term, self.cs = spec.link_groups(m)
        self.F = F = len(self.feet)
        self.L = m.nlinks
        self.widths = spec.obs_widths(m)
        self.sum_names = spec.episode_sum_names()
        self.hs = None if height_samples is None else np.ascontiguousarray(height_samples, np.int16)
        self.terrain_origins = None if terrain_origins is None else np.asarray(terrain_origins, f32)
        self.q0 = np.asarray(spec.default_dof_pos, f32)[None, :]
        lim = m.dof_limits.astype(f32).copy()                      # soft limits, genesis_simulator.py:373-382
        for i in range(A):
            mid = (lim[i, 0] + lim[i, 1]) / f32(2)
            r = lim[i, 1] - lim[i, 0]
            lim[i, 0] = mid - f32(0.5) * r * f32(s.soft_dof_pos_limit)
            lim[i, 1] = mid + f32(0.5) * r * f32(s.soft_dof_pos_limit)
        self.dof_pos_limits = lim
        gx, gy = np.meshgrid(np.asarray(s.measured_points_x, f32), np.asarray(s.measured_points_y, f32), indexing="ij")
        self.height_points = np.stack([gx.ravel(), gy.ravel()], axis=-1).astype(f32)   # [P,2]
        self.P = self.height_points.shape[0]
        if s.heightfield:
            self.x_range = (f32(-s.border_size + 1.0), f32(s.border_size + s.num_rows * s.terrain_length - 1.0))
            self.y_range = (f32(-s.border_size + 1.0), f32(s.border_size + s.num_cols * s.terrain_width - 1.0))
        else:
            self.x_range = self.y_range = (f32(-s.plane_length / 2 + 1), f32(s.plane_length / 2 - 1))
        self.cmd_range_x = [float(s.cmd_lin_vel_x[0]), float(s.cmd_lin_vel_x[1])]
        self.common_step_counter = 0
        self.noise_vec = spec.noise_scale_vec()
        z = lambda *sh: np.zeros(sh, f32)
        self.st = dict(
            base_pos=z(N, 3), base_quat_wxyz=z(N, 4), base_lin_w=z(N, 3), base_ang_w=z(N, 3), q=z(N, A), qd=z(N, A),
            friction=z(N, 1), added_mass=np.ones((N, 1), f32), com_bias=z(N, 3), kp_scale=np.ones((N, A), f32),
            kd_scale=np.ones((N, A), f32), joint_armature=z(N, 1), joint_friction=z(N, 1), joint_damping=z(N, 1),
            rand_push_vels=z(N, 3),
            actions=z(N, A), last_actions=z(N, A), llast_actions=z(N, A), commands=z(N, 4),
            episode_length=np.zeros(N, np.int32), fail_buf=np.zeros(N, np.int32), feet_air_time=z(N, F),
            last_contacts=np.zeros((N, F), np.uint8), episode_sums=z(N, len(self.sum_names)),
            terrain_levels=np.zeros(N, np.int32), terrain_types=np.zeros(N, np.int32), env_origins=z(N, 3),
            base_lin_vel=z(N, 3), base_ang_vel=z(N, 3), feet_vel=z(N, F, 3), last_dof_vel=z(N, A),
            last_feet_vel=z(N, F, 3), last_base_lin_vel=z(N, 3), last_base_ang_vel=z(N, 3),
            obs_hist=z(N, max(self.widths["hist"], 1)), critic_hist=z(N, max(self.widths["critic"], 1)),
            gait_state=z(N, 20),     # theta[4], gait_time, phi, gait_period, base_height_t, foot_clearance_t, pitch_t, clock[8]
        )
        self.st["gait_state"][:, G_PER] = spec.gait_period
        if spec.randomize_ctrl_delay:                          # legged_robot.py:405-409
            self.st["action_queue"] = z(N, int(spec.ctrl_delay_step_range[1]) + 1, A)
            self.st["action_delay"] = np.zeros(N, np.int32)
        # go2_wtw host-side behaviour state (go2_wtw.py:352-372): curriculum-mutable ranges and the number of unlocked gaits
        mid = lambda r: [(r[0] + r[1]) / 2] * 2
        self.beh_ranges = dict(gait_period=mid(spec.gait_period_range), base_height=mid(spec.base_height_target_range),
                               foot_clearance=[spec.foot_clearance_target_range[0]] * 2, pitch=mid(spec.pitch_target_range))
        self.num_gaits = 1
        self.gait_choice = None      # (callback, reset) gait indices of the step; None -> drawn from the SITE_HOST stream
        # R18 (reference defect, found here): tron1_pf_ee.py:32-34,414-421 flatten `nonzero()` of [N,1] masks, so the column
        # index 0 joins the row indices: env 0's gait time is zeroed whenever ANY env wraps and its swing/stance indicator
        # is overwritten whenever any env is in swing/stance.  The product does not reproduce this coupling;
        # the oracle can, so that it is pinned on every env of the reference goldens.
        self.reproduce_r18 = False
        self.sit_coin = None         # host coin of the step (tron1_pf_ee.py:204); None -> drawn from the SITE_HOST stream
        self.st["base_quat_wxyz"][:, 0] = 1
        self.out = {}

    # ------------------------------------------------------------------ randoms
    def u(self, site, idx, env=None):
        env = np.arange(self.N) if env is None else env
        return philox.uniform(self.spec.seed, self.common_step_counter, np.asarray(env)[:, None], site,
                              np.atleast_1d(idx)[None, :])

    @staticmethod
    def _range(lo, hi, u):
        """torch_rand_float: (upper - lower) * rand + lower in fp32 (math_utils.py:78-81)."""
        return (f32(hi - lo) * u + f32(lo)).astype(f32)

    # ------------------------------------------------------------------ step halves
    def pre_step(self, actions):
        s, st = self.spec, self.st
        a = np.clip(np.asarray(actions, f32), f32(-s.clip_actions), f32(s.clip_actions))
        st["llast_actions"][:] = st["last_actions"]
        st["last_actions"][:] = st["actions"]
        st["actions"][:] = a
        st["last_base_lin_vel"][:] = st["base_lin_vel"]
        st["last_base_ang_vel"][:] = st["base_ang_vel"]
        st["last_feet_vel"][:] = st["feet_vel"]
        st["last_dof_vel"][:] = st["qd"]
        if s.randomize_ctrl_delay:                             # legged_robot.py:240-245: the simulator gets a delayed action
            qu = st["action_queue"]
            qu[:, 1:] = qu[:, :-1].copy()
            qu[:, 0] = a
            a = qu[np.arange(self.N), st["action_delay"]].copy()
        return a

    def pd_torque(self, a, q, qd):
        s, st = self.spec, self.st
        scaled = a * f32(s.action_scale)
        return (st["kp_scale"] * f32(s.kp) * (scaled + self.q0 - q) - st["kd_scale"] * f32(s.kd) * qd).astype(f32)

    def post_step(self, phys):
        """phys: dict(base_pos, base_quat_wxyz, base_lin_w, base_ang_w, q, qd, torques, link_force[N,L,3],
        feet_pos[N,F,3], feet_vel[N,F,3]) -- the state right after the decimated physics loop."""
        st = self.st
        for k in ("base_pos", "base_quat_wxyz", "base_lin_w", "base_ang_w", "q", "qd"):
            st[k][:] = np.asarray(phys[k], f32)
        st["episode_length"] += 1
        self.common_step_counter += 1
        o = self.out = {}
        self._sim_post(phys, o)
        self._callback(o)
        self._termination(o)
        if self.spec.cat_enabled:
            self._constraints_cat(o)
        self._reward(o)
        if self.spec.gait_enabled:                        # tron1_pf_ee.py:27-35
            g = st["gait_state"]
            g[:, G_GT] = (g[:, G_GT] + f32(self.spec.dt)).astype(f32)
            over = g[:, G_GT] >= (g[:, G_PER] - f32(self.spec.dt / 2)).astype(f32)
            g[over, G_GT] = 0
            if self.reproduce_r18 and over.any():
                g[0, G_GT] = 0
            g[:, G_PHI] = (g[:, G_GT] / g[:, G_PER]).astype(f32)
        self._reset(o)
        if self.spec.gait_enabled:                        # _calc_periodic_reward_obs, tron1_pf_ee.py:258-263
            g = st["gait_state"]
            for i in range(self.F):
                arg = (f32(2 * np.pi) * (g[:, G_PHI] + g[:, G_TH + i]).astype(f32)).astype(f32)
                g[:, G_CLK + i] = np.sin(arg)
                g[:, G_CLK + self.F + i] = np.cos(arg)
        self._observe(o)
        if self.spec.double_shift_actions:               # go2_cat.py:127-130
            st["llast_actions"][:] = st["last_actions"]
            st["last_actions"][:] = st["actions"]
            st["last_dof_vel"][:] = st["qd"]
            st["last_feet_vel"][:] = st["feet_vel"]
        return o

    # ------------------------------------------------------------------ genesis_simulator.py:35-60
    def _sim_post(self, phys, o):
        s, st = self.spec, self.st
        bp = st["base_pos"]
        feet_pos = np.asarray(phys["feet_pos"], f32).copy()
        oob = (bp[:, 0] >= self.x_range[1]) | (bp[:, 0] <= self.x_range[0]) | (bp[:, 1] >= self.y_range[1]) | (bp[:, 1] <= self.y_range[0])
        if oob.any():                                            # genesis_simulator.py:612-628
            new = (np.asarray(s.init_pos, f32)[None, :] + st["env_origins"][oob]).astype(f32)
            feet_pos[oob] += (new - bp[oob])[:, None, :]
            bp[oob] = new
        o["oob"] = oob
        qw = st["base_quat_wxyz"]
        quat = np.concatenate([qw[:, 1:4], qw[:, 0:1]], axis=1).astype(f32)      # wxyz -> xyzw
        o["base_quat"] = quat
        o["base_euler"] = _euler_xyz(quat)
        st["base_lin_vel"][:] = _rot_inv(quat, st["base_lin_w"])
        st["base_ang_vel"][:] = _rot_inv(quat, st["base_ang_w"])
        g = np.zeros((self.N, 3), f32)
        g[:, 2] = -1
        o["projected_gravity"] = _rot_inv(quat, g)
        o["dof_pos"], o["dof_vel"] = st["q"].copy(), st["qd"].copy()
        o["torques"] = np.asarray(phys["torques"], f32).copy()
        o["link_contact_forces"] = lf = np.asarray(phys["link_force"], f32).copy()
        o["feet_pos"] = feet_pos
        st["feet_vel"][:] = np.asarray(phys["feet_vel"], f32)
        o["feet_vel"] = st["feet_vel"].copy()
        o["base_pos"] = bp.copy()
        if s.obtain_link_contact_states:
            o["link_contact_states"] = (_norm3(lf[:, self.cs, :]) > f32(1.0)).astype(f32)
        else:
            o["link_contact_states"] = np.zeros((self.N, 0), f32)
        if s.measure_heights and s.heightfield:
            self._heights(o)
            if s.obtain_terrain_info_around_feet:
                self._feet_terrain(o)
        else:                                                     # scan never updated on a plane
            o["measured_heights"] = np.zeros((self.N, self.P), f32)
            o["height_cells"] = np.zeros((self.N, self.P, 2), np.int32)

    def _cell(self, p):
        """trunc((p + border) / hscale) with fp32 IEEE ops (SURVEY a-kernel notes)."""
        s = self.spec
        g = ((p + f32(s.border_size)).astype(f32) / f32(s.horizontal_scale)).astype(f32)
        return np.trunc(g).astype(np.int64)

    def _heights(self, o):
        s, st = self.spec, self.st
        quat = o["base_quat"]
        z, w = quat[:, 2], quat[:, 3]
        nrm = np.maximum(np.sqrt((z * z + w * w).astype(f32)).astype(f32), f32(1e-9))
        zn, wn = (z / nrm).astype(f32)[:, None], (w / nrm).astype(f32)[:, None]
        px, py = self.height_points[None, :, 0], self.height_points[None, :, 1]
        tx = (-(zn * py) * f32(2.0)).astype(f32)
        ty = ((zn * px) * f32(2.0)).astype(f32)
        rx = ((px + wn * tx).astype(f32) + (-(zn * ty)).astype(f32)).astype(f32)
        ry = ((py + wn * ty).astype(f32) + (zn * tx).astype(f32)).astype(f32)
        wx = (rx + st["base_pos"][:, 0:1]).astype(f32)
        wy = (ry + st["base_pos"][:, 1:2]).astype(f32)
        cx = np.clip(self._cell(wx), 0, self.hs.shape[0] - 2)
        cy = np.clip(self._cell(wy), 0, self.hs.shape[1] - 2)
        h = np.minimum(np.minimum(self.hs[cx, cy], self.hs[cx + 1, cy]), self.hs[cx, cy + 1])
        o["measured_heights"] = (h.astype(f32) * f32(s.vertical_scale)).astype(f32)
        o["height_cells"] = np.stack([cx, cy], axis=-1).astype(np.int32)

    def _feet_terrain(self, o):
        s = self.spec
        fp = o["feet_pos"]
        cx = np.clip(self._cell(fp[:, :, 0]), 0, self.hs.shape[0] - 2)
        cy = np.clip(self._cell(fp[:, :, 1]), 0, self.hs.shape[1] - 2)
        H = lambda dx, dy: self.hs[cx + dx, cy + dy]
        hh = [H(-1, 0), H(1, 0), H(0, -1), H(0, 1), H(0, 0), H(-1, -1), H(1, 1), H(-1, 1), H(1, -1)]
        den = f32(s.horizontal_scale * 2)
        # int16 difference (torch keeps int16) then true division -> fp32
        dx = ((hh[1] - hh[0]).astype(np.int16).astype(f32) / den).astype(f32)
        dy = ((hh[3] - hh[2]).astype(np.int16).astype(f32) / den).astype(f32)
        nv = np.stack([dx, dy, -np.ones_like(dx)], axis=-1).astype(f32)
        nv = (nv / _norm3(nv)[..., None]).astype(f32)
        o["normal_vector_around_feet"] = nv.reshape(self.N, -1)
        o["height_around_feet"] = (np.stack(hh, axis=-1).astype(f32) * f32(s.vertical_scale)).astype(f32)
        o["feet_cells"] = np.stack([cx, cy], axis=-1).astype(np.int32)

    # ------------------------------------------------------------------ legged_robot.py:300-334
    def _resample(self, ids, site):
        s, st = self.spec, self.st
        if len(ids) == 0:
            return
        c = st["commands"]
        c[ids, 0] = self._range(self.cmd_range_x[0], self.cmd_range_x[1], self.u(site, [0], ids)[:, 0])
        c[ids, 1] = self._range(s.cmd_lin_vel_y[0], s.cmd_lin_vel_y[1], self.u(site, [1], ids)[:, 0])
        if s.heading_command:
            c[ids, 3] = self._range(s.cmd_heading[0], s.cmd_heading[1], self.u(site, [2], ids)[:, 0])
        else:
            c[ids, 2] = self._range(s.cmd_ang_vel_yaw[0], s.cmd_ang_vel_yaw[1], self.u(site, [2], ids)[:, 0])
        keep = (_norm3(c[ids, :3]) > f32(0.2)).astype(f32)
        c[ids, :3] *= keep[:, None]

    def _callback(self, o):
        s, st = self.spec, self.st
        ids = np.nonzero(st["episode_length"] % s.resample_interval == 0)[0]
        self._resample(ids, T.SITE_CMD_RESAMPLE)
        if s.heading_command:
            fwd = _quat_apply(o["base_quat"], np.tile(np.array([[1, 0, 0]], f32), (self.N, 1)))
            heading = np.arctan2(fwd[:, 1], fwd[:, 0]).astype(f32)
            ang = (st["commands"][:, 3] - heading).astype(f32)
            two_pi = f32(2 * np.pi)
            # wrap_to_pi, math_utils.py:49-53.  TorchScript compiles `angles %= 2*pi` to aten::fmod_ (sign of the
            # dividend), so negative angles are NOT wrapped: the result lies in (-2*pi, pi]  (quirk R16)
            ang = np.fmod(ang, two_pi).astype(f32)
            ang = (ang - two_pi * (ang > f32(np.pi)).astype(f32)).astype(f32)
            st["commands"][:, 2] = np.clip(f32(0.5) * ang, f32(s.cmd_ang_vel_yaw[0]), f32(s.cmd_ang_vel_yaw[1]))
        o["pushed"] = False
        if s.push_robots and self.common_step_counter % s.push_interval == 0:
            push = self._range(-s.max_push_vel_xy, s.max_push_vel_xy, self.u(T.SITE_PUSH, [0, 1]))
            st["rand_push_vels"][:, :2] = push
            st["base_lin_w"][:, :2] += push
            o["pushed"] = True
        if s.behavior_enabled:                                     # go2_wtw.py:258-263
            ids = np.nonzero(st["episode_length"] % int(s.behavior_resampling_time / s.dt) == 0)[0]
            self._resample_behavior(ids, T.SITE_BEHAVIOR, self._gait_choice(0))

    def _gait_choice(self, which):
        """ONE randint per resampling call for all envs (go2_wtw.py:205-206); host draw idx 1 (callback) / 2 (reset)."""
        if self.gait_choice is not None:
            return self.gait_choice[which]
        u = float(philox.uniform(self.spec.seed, self.common_step_counter, 0xFFFFFFFF, T.SITE_HOST, 1 + which))
        return min(int(u * self.num_gaits), self.num_gaits - 1)

    def _resample_behavior(self, ids, site, gait):
        """go2_wtw.py:180-217 (the pronk/bound clearance clamp is applied per resampled env; DESIGN.md deviation D4)."""
        s, g = self.spec, self.st["gait_state"]
        if len(ids) == 0:
            return
        u = self.u(site, [0, 1, 2, 3], ids)
        r = self.beh_ranges
        g[ids, G_PER] = self._range(r["gait_period"][0], r["gait_period"][1], u[:, 0])
        g[ids, G_BH] = self._range(r["base_height"][0], r["base_height"][1], u[:, 1])
        g[ids, G_FC] = self._range(r["foot_clearance"][0], r["foot_clearance"][1], u[:, 2])
        g[ids, G_PT] = self._range(r["pitch"][0], r["pitch"][1], u[:, 3])
        th = np.asarray(s.gait_theta_lists[gait], f32)
        g[ids, G_TH:G_TH + 4] = th[None, :]
        if (th[0] == 0 and th[1] == 0) and ((th[2] == 0 and th[3] == 0) or (th[2] == 0.5 and th[3] == 0.5)):
            g[ids, G_FC] = f32(r["foot_clearance"][0])

    # ------------------------------------------------------------------ legged_robot.py:78-92
    def _termination(self, o):
        s, st = self.spec, self.st
        lf = o["link_contact_forces"]
        fail = np.zeros(self.N, bool)
        if len(self.term):
            fail = np.any(_norm3(lf[:, self.term, :]) > f32(10.0), axis=1)
        fail |= o["projected_gravity"][:, 2] > f32(s.max_projected_gravity)
        st["fail_buf"] += fail.astype(np.int32)
        o["time_out_buf"] = st["episode_length"] > s.max_episode_length
        o["reset_buf"] = (st["fail_buf"] > s.fail_limit) | o["time_out_buf"]

    # ------------------------------------------------------------------ go2_cat.py:135-215 + constraint_manager.py:23-71
    def _constraints_cat(self, o):
        s, st = self.spec, self.st
        lf = o["link_contact_forces"]
        tl = self.model.effort.astype(f32)[None, :]
        vl = np.asarray(s.dof_vel_limits, f32)[None, :]
        lim = self.dof_pos_limits
        c = {}
        c["torque"] = np.any(np.abs(o["torques"]) > tl, axis=1)
        c["dof_vel"] = np.any(np.abs(o["dof_vel"]) > vl, axis=1)
        c["action_rate"] = np.any((np.abs(st["actions"] - st["last_actions"]) / f32(s.dt)).astype(f32) > f32(s.cat_action_rate), axis=1)
        c["base_height"] = np.mean(st["base_pos"][:, 2:3] - o["measured_heights"], axis=1, dtype=f32) < f32(s.cat_min_base_height)
        c["collision"] = np.any(_norm3(lf[:, self.pen, :]) > f32(10.0), axis=1)
        ff = lf[:, self.feet, :]
        c["feet_stumble"] = np.any(_norm3(ff) > (f32(4) * np.abs(ff[:, :, 2])).astype(f32), axis=1)
        c["dof_pos"] = np.any(o["dof_pos"] < lim[None, :, 0], axis=1) & np.any(o["dof_pos"] > lim[None, :, 1], axis=1)
        c["base_orientation"] = o["projected_gravity"][:, 2] > f32(s.cat_max_projected_gravity)
        fast = np.any(np.abs(o["dof_vel"]) > f32(4.0), axis=1)
        still = _norm3(st["commands"][:, :3]) < f32(0.1)
        # as shipped the product is an [N,N] broadcast: env i is flagged when it stands still and ANY env moves fast (R4)
        c["stand_still"] = (still & fast.any()) if s.cat_stand_still_global else (still & fast)
        soft = f32(s.cat_soft_p)
        pmax = dict(torque=soft, dof_vel=soft, action_rate=soft, base_height=soft, collision=f32(1), feet_stumble=f32(1),
                    dof_pos=f32(1), base_orientation=f32(1), stand_still=soft)
        # violations are 0/1 and the running maximum never exceeds 1, so the probability is max_p wherever violated
        prob = np.zeros(self.N, f32)
        base = len(self.sum_names) - len(T.CAT_CONSTRAINTS)
        for k, name in enumerate(T.CAT_CONSTRAINTS):
            prob = np.maximum(prob, np.where(c[name], pmax[name], f32(0)))
            st["episode_sums"][:, base + k] += c[name].astype(f32)
        o["cstr"] = c
        o["cstr_prob"] = prob

    # ------------------------------------------------------------------ legged_robot.py:150-168 + _reward_*
    def _reward(self, o):
        s, st = self.spec, self.st
        lf, cmd = o["link_contact_forces"], st["commands"]
        dt = f32(s.dt)
        ssum = lambda x: np.sum(x, axis=1, dtype=f32).astype(f32)
        terms = {}

        def feet_air_time():
            contact = lf[:, self.feet, 2] > f32(1.0)
            filt = contact | (st["last_contacts"] > 0)
            st["last_contacts"][:] = contact
            first = (st["feet_air_time"] > 0) & filt
            st["feet_air_time"] += dt
            r = ssum((st["feet_air_time"] - f32(s.feet_air_time_threshold)) * first)
            r = r * (_norm3(np.concatenate([cmd[:, :2], np.zeros((self.N, 1), f32)], 1)) > f32(0.1))
            st["feet_air_time"] *= ~filt
            return r

        def foot_clearance():
            vxy = np.sqrt(np.sum(o["feet_vel"][:, :, :2] ** 2, axis=-1, dtype=f32)).astype(f32)
            z = o["feet_pos"][:, :, 2]
            if s.foot_clearance_mode == 1:
                z = z - np.mean(o["height_around_feet"], axis=-1, dtype=f32).astype(f32)
            elif s.foot_clearance_mode == 2:                         # tron1_pf_ee.py:446-456
                z = z - np.max(o["height_around_feet"], axis=-1)
            err = ssum(vxy * np.square(z - f32(s.foot_clearance_target) - f32(s.foot_height_offset)))
            return np.exp(-err / f32(s.foot_clearance_tracking_sigma)).astype(f32)

        def foot_landing_vel():
            zv = o["feet_vel"][:, :, 2]
            contacts = lf[:, self.feet, 2] > f32(0.1)
            about = ((o["feet_pos"][:, :, 2] - f32(s.foot_height_offset)) < f32(s.about_landing_threshold)) & ~contacts & (zv < 0)
            return ssum(np.square(np.where(about, zv, f32(0))))

        small_cmd = (_norm3(cmd[:, :3]) < f32(0.1)).astype(f32)
        dq = o["dof_pos"] - self.q0
        lim = self.dof_pos_limits
        fn = {
            "action_rate": lambda: ssum(np.square(st["last_actions"] - st["actions"])),
            "action_smoothness": lambda: ssum(np.square(st["actions"] - f32(2) * st["last_actions"] + st["llast_actions"])),
            "ang_vel_xy": lambda: ssum(np.square(st["base_ang_vel"][:, :2])),
            "base_height": lambda: np.square(np.mean(st["base_pos"][:, 2:3] - o["measured_heights"], axis=1, dtype=f32)
                                             - f32(s.base_height_target)),
            "collision": lambda: ssum((_norm3(lf[:, self.pen, :]) > f32(0.1)).astype(f32)),
            "dof_acc": lambda: ssum(np.square((st["last_dof_vel"] - o["dof_vel"]) / dt)),
            "dof_close_to_default": lambda: ssum(np.square(dq)),
            "dof_pos_limits": lambda: ssum(-np.minimum(o["dof_pos"] - lim[None, :, 0], f32(0)) + np.maximum(o["dof_pos"] - lim[None, :, 1], f32(0))),
            "dof_pos_stand_still": lambda: ssum(np.square(dq)) * small_cmd,
            "dof_power": lambda: ssum(np.abs(o["torques"] * o["dof_vel"])),
            "dof_vel": lambda: ssum(np.square(o["dof_vel"])),
            "dof_vel_stand_still": lambda: ssum(np.abs(o["dof_vel"])) * small_cmd,
            "feet_air_time": feet_air_time,
            "feet_contact_stand_still": lambda: (ssum((lf[:, self.feet, 2] > f32(0.1)).astype(f32)) == self.F).astype(f32) * small_cmd,
            "feet_distance": lambda: np.maximum(f32(0), f32(s.foot_distance_threshold) - np.sqrt(np.sum(np.square(
                o["feet_pos"][:, 0, :2] - o["feet_pos"][:, 1, :2]), axis=-1, dtype=f32)).astype(f32)),      # tron1_pf.py:146-151
            "no_fly": lambda: (ssum((lf[:, self.feet, 2] > f32(0.1)).astype(f32)) == 1).astype(f32),       # tron1_pf.py:153-156
            "biped_periodic_gait": lambda: self._periodic_gait(o, 2),
            "quad_periodic_gait": lambda: self._periodic_gait(o, 4),
            "tracking_base_height": lambda: np.exp(-np.square(np.mean(st["base_pos"][:, 2:3] - o["measured_heights"], axis=1, dtype=f32)
                                                              - (st["gait_state"][:, G_BH] if s.behavior_enabled else f32(s.base_height_target)))
                                                   / f32(s.base_height_tracking_sigma)).astype(f32),
            "tracking_orientation": lambda: np.exp(-(np.square(o["base_euler"][:, 0]) + np.square(o["base_euler"][:, 1] - st["gait_state"][:, G_PT]))
                                                   / f32(s.euler_tracking_sigma)).astype(f32),                      # go2_wtw.py:497-500
            "tracking_foot_clearance": lambda: np.exp(-ssum(np.sqrt(np.sum(o["feet_vel"][:, :, :2] ** 2, axis=-1, dtype=f32)).astype(f32) * np.square(
                o["feet_pos"][:, :, 2] - st["gait_state"][:, G_FC:G_FC + 1] - f32(s.foot_height_offset))) / f32(s.foot_clearance_tracking_sigma)).astype(f32),
            "foot_acc": lambda: np.sum(np.square((o["feet_vel"] - st["last_feet_vel"]) / dt), axis=(1, 2), dtype=f32),
            "foot_clearance": foot_clearance,
            "foot_landing_vel": foot_landing_vel,
            "hip_pos": lambda: ssum(np.square(dq[:, 0::3])),
            "keep_balance": lambda: np.ones(self.N, f32),
            "lin_vel_z": lambda: np.square(st["base_lin_vel"][:, 2]),
            "orientation": lambda: ssum(np.square(o["projected_gravity"][:, :2])),
            "thigh_pos": lambda: ssum(np.square(dq[:, 1::3])),
            "torques": lambda: ssum(np.square(o["torques"])),
            "tracking_ang_vel": lambda: np.exp(-np.square(cmd[:, 2] - st["base_ang_vel"][:, 2]) / f32(s.tracking_sigma)),
            "tracking_lin_vel": lambda: np.exp(-ssum(np.square(cmd[:, :2] - st["base_lin_vel"][:, :2])) / f32(s.tracking_sigma)),
        }
        rew = np.zeros(self.N, f32)
        for i, name in enumerate(s.active_rewards()):
            r = (fn[name]().astype(f32) * s.scaled_reward(name)).astype(f32)
            terms[name] = r
            rew = (rew + r).astype(f32)
            st["episode_sums"][:, i] += r
        if s.only_positive_rewards:
            if s.cat_enabled:                                      # go2_cat.py:229-231
                rew = (rew * (f32(1.0) - o["cstr_prob"])).astype(f32)
            rew = np.maximum(rew, f32(0))
        if s.reward_scales.get("termination", 0) != 0:
            r = ((o["reset_buf"] & ~o["time_out_buf"]).astype(f32) * s.scaled_reward("termination")).astype(f32)
            rew = (rew + r).astype(f32)
            st["episode_sums"][:, self.sum_names.index("termination")] += r
            terms["termination"] = r
        o["rew_buf"], o["reward_terms"] = rew, terms

    # ------------------------------------------------------------------ tron1_pf_ee.py:335-437 ("step" indicator)
    def _periodic_gait(self, o, nfeet):
        s, st = self.spec, self.st
        g = st["gait_state"]
        lf = o["link_contact_forces"]
        total = np.zeros(self.N, np.float64 if s.gait_smooth else f32)
        o["exp_C_frc"] = np.zeros((self.N, nfeet), f32)
        for i in range(nfeet):
            q_frc = _norm3(lf[:, self.feet[i], :])
            q_spd = _norm3(o["feet_vel"][:, i, :])
            phi = (np.mod((g[:, G_PHI] + g[:, G_TH + i]).astype(f32), f32(1.0)) * f32(2 * np.pi)).astype(f32)
            b_swing = f32(s.gait_b_swing * 2 * np.pi)
            swing = (phi >= 0) & (phi < b_swing)
            stance = (phi >= b_swing) & (phi < f32(2 * np.pi))
            if s.gait_smooth:
                # go2_wtw.py:415-453 / tron1_pf_ee.py:369-407: von Mises CDFs from scipy on the host, in float64
                from scipy.stats import vonmises
                FA = np.clip(vonmises.cdf(loc=0.0, kappa=s.gait_kappa, x=phi), 0.0, 1.0)
                FB = np.clip(vonmises.cdf(loc=np.full(self.N, b_swing, f32), kappa=s.gait_kappa, x=phi), 0.0, 1.0)
                FS = np.clip(vonmises.cdf(loc=2 * np.pi, kappa=s.gait_kappa, x=phi), 0.0, 1.0)
                sw_ind, st_ind = FA * (1 - FB), FB * (1 - FS)
                spd_ori, frc_ori = -st_ind, -sw_ind
                c_frc = -0.5 + (-0.5 - spd_ori)
                c_spd = spd_ori.copy()
                in_swing = swing.copy()
                if self.reproduce_r18 and swing.any():
                    in_swing[0] = True                     # flattened nonzero() of the [N,1] mask also indexes row 0 (R18)
                c_frc[in_swing] = frc_ori[in_swing]
                c_spd[in_swing] = -0.5 + (-0.5 - frc_ori[in_swing])
                if b_swing == 0:
                    c_frc[:], c_spd[:] = 0, -1
                o["exp_C_frc"][:, i] = c_frc.astype(f32)
                total = total + (c_spd * q_spd + c_frc * q_frc)          # float64, like the reference's mixed-dtype sum
                continue
            c_frc = np.where(swing, f32(-1), f32(0))
            c_spd = np.where(stance, f32(-1), f32(0))
            if self.reproduce_r18:
                if swing.any():
                    c_frc[0], c_spd[0] = -1, 0
                if stance.any():
                    c_frc[0], c_spd[0] = 0, -1
            o["exp_C_frc"][:, i] = c_frc
            total = (total + (c_spd * q_spd + c_frc * q_frc).astype(f32)).astype(f32)
        return np.exp(total).astype(f32)

    # ------------------------------------------------------------------ legged_robot.py:94-148
    def _reset(self, o):
        s, st = self.spec, self.st
        ids = np.nonzero(o["reset_buf"])[0]
        o["env_ids"] = ids
        o["projected_gravity_obs"] = o["projected_gravity"]
        o["episode_means"] = None
        if len(ids) == 0:
            return
        if s.terrain_curriculum:                                   # legged_robot.py:254-272; genesis_simulator.py:140-148
            dist = np.sqrt(np.sum(np.square(st["base_pos"][ids, :2] - st["env_origins"][ids, :2]), axis=1, dtype=f32)).astype(f32)
            up = dist > f32(s.terrain_length / 2)
            cn = np.sqrt(np.sum(np.square(st["commands"][ids, :2]), axis=1, dtype=f32)).astype(f32)
            down = (dist < (cn * f32(s.episode_length_s) * f32(0.5)).astype(f32)) & ~up
            lv = st["terrain_levels"][ids] + up.astype(np.int32) - down.astype(np.int32)
            rnd = np.minimum((self.u(T.SITE_LEVEL, [0], ids)[:, 0] * f32(s.num_rows)).astype(np.int32), s.num_rows - 1)
            lv = np.where(lv >= s.num_rows, rnd, np.maximum(lv, 0))
            st["terrain_levels"][ids] = lv
            st["env_origins"][ids] = self.terrain_origins[lv, st["terrain_types"][ids]]
        # command curriculum (legged_robot.py:336-348) is applied by the host one step later: DESIGN.md "deviations"
        if s.behavior_enabled:                                     # go2_wtw.py:124-126: behaviour params before commands
            self._resample_behavior(ids, T.SITE_BEHAVIOR_RESET, self._gait_choice(1))
        self._resample(ids, T.SITE_CMD_RESET)
        A = self.A
        sit = False
        if s.sit_init_percent > 0:                                 # tron1_pf_ee.py:204-210: ONE coin per reset batch (R8)
            coin = self.sit_coin if self.sit_coin is not None else float(philox.uniform(s.seed, self.common_step_counter, 0xFFFFFFFF, T.SITE_HOST, 0))
            sit = coin < s.sit_init_percent
        o["sit"] = sit
        noise = np.asarray(s.reset_dof_noise, f32)[None, :]        # go2_ts.py:86-91
        udof = self.u(T.SITE_DOF, np.arange(A), ids)
        if sit:
            st["q"][ids] = np.asarray(s.sit_joint_angles, f32)[None, :]
        else:
            st["q"][ids] = (self.q0 + (f32(2) * noise * udof + (-noise)).astype(f32)).astype(f32)
        st["qd"][ids] = 0
        o["dof_pos"][ids], o["dof_vel"][ids] = st["q"][ids], 0
        uroot = self.u(T.SITE_ROOT, np.arange(8), ids)             # legged_robot.py:283-298
        pos = (np.asarray(s.sit_pos if sit else s.init_pos, f32)[None, :] + st["env_origins"][ids]).astype(f32)
        if s.heightfield:
            pos[:, :2] += self._range(-s.reset_root_xy, s.reset_root_xy, uroot[:, 0:2])
        st["base_pos"][ids] = pos
        qx = np.asarray(s.init_quat_xyzw, f32)
        if sit:                                                    # quat_from_euler_xyz(0, pitch, 0), tron1_pf_ee.py:294-300
            half = f32(s.sit_pitch_angle) * f32(0.5)
            qx = np.array([0, np.sin(half), 0, np.cos(half)], f32)
        st["base_quat_wxyz"][ids] = np.array([qx[3], qx[0], qx[1], qx[2]], f32)
        lin = self._range(-s.reset_root_vel, s.reset_root_vel, uroot[:, 2:5])
        ang = self._range(-s.reset_root_vel, s.reset_root_vel, uroot[:, 5:8])
        if sit:
            lin, ang = np.zeros_like(lin), np.zeros_like(ang)
        st["base_lin_w"][ids], st["base_ang_w"][ids] = lin, ang
        st["base_lin_vel"][ids], st["base_ang_vel"][ids] = lin, ang
        quat = o["base_quat"].copy()
        quat[ids] = qx
        g = np.zeros((self.N, 3), f32)
        g[:, 2] = -1
        o["projected_gravity_obs"] = _rot_inv(quat, g)             # genesis_simulator.py:125
        o["base_pos"][ids] = pos
        # domain randomisation, genesis_simulator.py:62-82,665-739
        if s.randomize_friction:
            st["friction"][ids, 0] = self._range(s.friction_range[0], s.friction_range[1], self.u(T.SITE_FRICTION, [0], ids)[:, 0])
        if s.randomize_base_mass:
            st["added_mass"][ids, 0] = self._range(s.added_mass_range[0], s.added_mass_range[1], self.u(T.SITE_MASS, [0], ids)[:, 0])
        if s.randomize_com_displacement:
            uc = self.u(T.SITE_COM, [0, 1, 2], ids)
            for k, rg in enumerate((s.com_pos_x_range, s.com_pos_y_range, s.com_pos_z_range)):
                st["com_bias"][ids, k] = self._range(rg[0], rg[1], uc[:, k])
        for flag, key, rg, site in ((s.randomize_joint_armature, "joint_armature", s.joint_armature_range, T.SITE_ARMATURE),
                                    (s.randomize_joint_friction, "joint_friction", s.joint_friction_range, T.SITE_JFRICTION),
                                    (s.randomize_joint_damping, "joint_damping", s.joint_damping_range, T.SITE_JDAMPING)):
            if flag:
                st[key][ids, 0] = self._range(rg[0], rg[1], self.u(site, [0], ids)[:, 0])
        if s.randomize_pd_gain:
            st["kp_scale"][ids] = self._range(s.kp_range[0], s.kp_range[1], self.u(T.SITE_KP, np.arange(A), ids))
            st["kd_scale"][ids] = self._range(s.kd_range[0], s.kd_range[1], self.u(T.SITE_KD, np.arange(A), ids))
        for k in ("last_dof_vel", "last_feet_vel", "last_base_lin_vel", "last_base_ang_vel",
                  "llast_actions", "last_actions", "actions", "feet_air_time"):
            st[k][ids] = 0
        st["episode_length"][ids] = 0
        st["fail_buf"][ids] = 0
        if s.gait_enabled:                                         # tron1_pf_ee.py:221-228
            g = st["gait_state"]
            if s.behavior_enabled:                                 # go2_wtw.py:139-142
                g[ids, G_GT] = 0
                g[ids, G_PHI] = 0
            else:
                ug = self.u(T.SITE_GAIT, [0, 1], ids)
                g[ids, G_TH] = (f32(s.gait_theta_left) + ug[:, 0]).astype(f32)
                g[ids, G_TH + 1] = (g[ids, G_TH] + f32(s.gait_theta_right - s.gait_theta_left)).astype(f32)
                g[ids, G_GT] = (ug[:, 1] * g[ids, G_PER]).astype(f32)
                g[ids, G_PHI] = (g[ids, G_GT] / g[ids, G_PER]).astype(f32)
            g[ids, G_CLK:G_CLK + 8] = 0
        o["episode_means"] = {"rew_" + n: np.mean(st["episode_sums"][ids, i], dtype=f32) / f32(s.episode_length_s)
                              for i, n in enumerate(self.sum_names)}
        st["episode_sums"][ids] = 0
        if s.randomize_ctrl_delay:                             # legged_robot.py:144-148: randint(lo, hi + 1) from one uniform draw
            lo, hi = (int(v) for v in s.ctrl_delay_step_range)
            st["action_queue"][ids] = 0
            ud = self.u(T.SITE_CTRL_DELAY, [0], ids)[:, 0]
            st["action_delay"][ids] = lo + np.minimum((ud * f32(hi - lo + 1)).astype(np.int32), hi - lo)
        if self.widths["hist"]:
            st["obs_hist"][ids] = 0
            st["critic_hist"][ids] = 0

    # ------------------------------------------------------------------ go2.py:40-90 / go2_ts.py:5-84
    def _observe(self, o):
        s, st = self.spec, self.st
        cs = np.array([s.obs_scale_lin_vel, s.obs_scale_lin_vel, s.obs_scale_ang_vel], f32)
        obs = np.concatenate([
            st["commands"][:, :3] * cs, o["projected_gravity_obs"], st["base_ang_vel"] * f32(s.obs_scale_ang_vel),
            (o["dof_pos"] - self.q0) * f32(s.obs_scale_dof_pos), o["dof_vel"] * f32(s.obs_scale_dof_vel), st["actions"]],
            axis=1).astype(f32)
        clean = obs.copy()
        if s.add_noise and s.obs_kind not in ("tron1_pf_ee", "go2_wtw"):
            un = self.u(T.SITE_OBS_NOISE, np.arange(obs.shape[1]))
            obs = (obs + ((f32(2) * un - f32(1)).astype(f32) * self.noise_vec[None, :]).astype(f32)).astype(f32)
        c = f32(s.clip_observations)
        if s.obs_kind == "go2_wtw":                                # go2_wtw.py:53-111
            g = st["gait_state"]
            obs61 = np.concatenate([clean, g[:, G_CLK:G_CLK + 8], g[:, G_PER:G_PT + 1], g[:, G_TH:G_TH + 4]], axis=1).astype(f32)
            lin = (st["base_lin_vel"] * f32(s.obs_scale_lin_vel)).astype(f32)
            crit = np.concatenate([obs61, lin, st["rand_push_vels"][:, :2], st["added_mass"], st["friction"], st["com_bias"],
                                   st["kp_scale"], st["kd_scale"], o.get("exp_C_frc", np.zeros((self.N, 4), f32))], axis=1).astype(f32)
            noisy = obs61.copy()
            if s.add_noise:
                un = self.u(T.SITE_OBS_NOISE, np.arange(obs61.shape[1]))
                noisy = (obs61 + ((f32(2) * un - f32(1)).astype(f32) * self.noise_vec[None, :]).astype(f32)).astype(f32)
            sc, so = self.widths["single_critic"], self.widths["obs"]
            st["critic_hist"][:] = np.concatenate([st["critic_hist"][:, sc:], crit], axis=1)
            st["obs_hist"][:] = np.concatenate([st["obs_hist"][:, so:], noisy], axis=1)
            o["obs_buf"] = np.clip(st["obs_hist"], -c, c)
            o["privileged_obs_buf"] = np.clip(st["critic_hist"], -c, c)
            o["obs_history"], o["critic_obs_buf"] = o["obs_buf"], o["privileged_obs_buf"]
            return
        if s.obs_kind == "tron1_pf_ee":                            # tron1_pf_ee.py:53-141
            g = st["gait_state"]
            obs31 = np.concatenate([clean, g[:, G_CLK:G_CLK + 4]], axis=1).astype(f32)
            noisy = obs31.copy()
            if s.add_noise:
                un = self.u(T.SITE_OBS_NOISE, np.arange(obs31.shape[1]))
                noisy = (obs31 + ((f32(2) * un - f32(1)).astype(f32) * self.noise_vec[None, :]).astype(f32)).astype(f32)
            dr = np.concatenate([st["friction"] - f32((s.friction_range[0] + s.friction_range[1]) / 2), st["added_mass"], st["com_bias"],
                                 st["rand_push_vels"][:, :2], st["kp_scale"] - f32((s.kp_range[0] + s.kp_range[1]) / 2),
                                 st["kd_scale"] - f32((s.kd_range[0] + s.kd_range[1]) / 2), st["joint_armature"], st["joint_friction"],
                                 st["joint_damping"]], axis=1).astype(f32)
            hobs = (np.clip(st["base_pos"][:, 2:3] - f32(s.height_obs_offset) - o["measured_heights"], f32(-1), f32(1))
                    * f32(s.obs_scale_height)).astype(f32)
            fz = o["feet_pos"][:, :, 2:3]
            crit = np.concatenate([obs31, dr, o.get("exp_C_frc", np.zeros((self.N, 2), f32)), o["link_contact_states"], hobs,
                                   o["normal_vector_around_feet"],
                                   np.clip((fz - o["height_around_feet"]).reshape(self.N, -1), f32(-1), f32(1))], axis=1).astype(f32)
            sc, so = self.widths["single_critic"], self.widths["obs"]
            st["critic_hist"][:] = np.concatenate([st["critic_hist"][:, sc:], crit], axis=1)
            st["obs_hist"][:] = np.concatenate([st["obs_hist"][:, so:], noisy], axis=1)
            lin = (st["base_lin_vel"] * f32(s.obs_scale_lin_vel)).astype(f32)
            clr = np.clip(o["feet_pos"][:, :, 2] - np.max(o["height_around_feet"], axis=-1) - f32(s.foot_height_offset), f32(-1), f32(1))
            o["estimator_labels_buf"] = np.concatenate([lin, o["link_contact_states"], clr, o["normal_vector_around_feet"]], axis=1).astype(f32)
            o["estimator_features_buf"] = np.clip(st["obs_hist"], -c, c)
            o["privileged_obs_buf"] = np.clip(st["critic_hist"], -c, c)
            o["obs_history"], o["critic_obs_buf"] = o["estimator_features_buf"], o["privileged_obs_buf"]
            o["obs_buf"] = o["estimator_features_buf"]
            return
        if s.obs_kind == "tron1_pf":                               # tron1_pf.py:15-70
            lin = (st["base_lin_vel"] * f32(s.obs_scale_lin_vel)).astype(f32)
            priv = np.concatenate([lin, clean, st["last_actions"], st["friction"] - f32((s.friction_range[0] + s.friction_range[1]) / 2),
                                   st["added_mass"], st["com_bias"], st["rand_push_vels"][:, :2], st["feet_air_time"]], axis=1).astype(f32)
            sc, so = self.widths["single_critic"], self.widths["obs"]
            st["critic_hist"][:] = np.concatenate([st["critic_hist"][:, sc:], priv], axis=1)
            st["obs_hist"][:] = np.concatenate([st["obs_hist"][:, so:], obs], axis=1)
            o["obs_buf"] = np.clip(st["obs_hist"], -c, c)
            o["privileged_obs_buf"] = np.clip(st["critic_hist"], -c, c)
            o["obs_history"], o["critic_obs_buf"] = o["obs_buf"], o["privileged_obs_buf"]
            return
        if s.obs_kind == "go2":
            o["obs_buf"] = np.clip(obs, -c, c)
            return
        dr = [st["friction"] - f32((s.friction_range[0] + s.friction_range[1]) / 2), st["added_mass"], st["com_bias"],
              st["rand_push_vels"][:, :2], st["kp_scale"] - f32((s.kp_range[0] + s.kp_range[1]) / 2),
              st["kd_scale"] - f32((s.kd_range[0] + s.kd_range[1]) / 2)]
        kind = s.obs_kind
        cat = kind == "go2_cat"
        if cat:                                                    # go2_cat.py:40-42
            dr += [st["joint_armature"], st["joint_friction"], st["joint_damping"]]
        dr = np.concatenate(dr, axis=1).astype(f32)
        lin = (st["base_lin_vel"] * f32(s.obs_scale_lin_vel)).astype(f32)
        if kind == "go2_dreamwaq":                                 # go2_dreamwaq.py:31-35: base_lin_vel leads the critic frame
            parts = [lin, clean, dr]
        elif cat or kind == "go2_ee":                              # go2_cat.py:45-48 / go2_ee.py:34-37: no base_lin_vel
            parts = [clean, dr]
        else:                                                      # go2_ts.py:30-34 / go2_cts.py:36-40
            parts = [clean, dr, lin]
        if s.obtain_link_contact_states:
            parts.append(o["link_contact_states"])
        if s.measure_heights:
            hobs = (np.clip(st["base_pos"][:, 2:3] - f32(s.height_obs_offset) - o["measured_heights"], f32(-1), f32(1))
                    * f32(s.obs_scale_height)).astype(f32)
            parts.append(hobs)
        critic = np.concatenate(parts, axis=1).astype(f32)
        sc, so = self.widths["single_critic"], self.widths["obs"]
        st["critic_hist"][:] = np.concatenate([st["critic_hist"][:, sc:], critic], axis=1)
        st["obs_hist"][:] = np.concatenate([st["obs_hist"][:, so:], obs], axis=1)
        fz = o["feet_pos"][:, :, 2:3]
        if kind in ("go2_ee", "go2_dreamwaq"):                     # go2_ee.py:66-73 / go2_dreamwaq.py:83-90: estimator labels
            clr = np.clip(o["feet_pos"][:, :, 2] - np.mean(o["height_around_feet"], axis=-1, dtype=f32) - f32(s.foot_height_offset),
                          f32(-1), f32(1)).astype(f32)
            lab_lin = lin if kind == "go2_ee" else (lin * f32(0.5)).astype(f32)
            labels = np.concatenate([lab_lin, o["link_contact_states"], clr], axis=1).astype(f32)
            if kind == "go2_ee":                                   # LeggedRobotEE.step, legged_robot_ee.py:56-73
                o["estimator_labels_buf"] = labels
                o["estimator_features_buf"] = np.clip(st["obs_hist"], -c, c)
                o["privileged_obs_buf"] = np.clip(st["critic_hist"], -c, c)
                o["obs_history"], o["critic_obs_buf"] = o["estimator_features_buf"], o["privileged_obs_buf"]
                o["obs_buf"] = o["estimator_features_buf"]
            else:                                                  # LeggedRobotDreamwaq.step, legged_robot_dreamwaq.py:63-79
                A = s.num_actions
                o["explicit_labels_buf"] = labels
                o["next_state_buf"] = np.concatenate([clean[:, :9 + 2 * A], (st["actions"] * f32(s.action_scale)).astype(f32)], axis=1)
                o["obs_buf"] = np.clip(obs, -c, c)
                o["privileged_obs_buf"] = np.clip(st["critic_hist"], -c, c)
                o["obs_history"], o["critic_obs_buf"] = st["obs_hist"].copy(), o["privileged_obs_buf"]
            return
        if cat:                                                    # go2_cat.py:82-99: raw heights, no base_lin_vel
            priv = [dr, o["height_around_feet"].reshape(self.N, -1), o["normal_vector_around_feet"]]
        elif kind == "go2_cts":                                    # go2_cts.py:74-81: raw heights
            priv = [dr, o["height_around_feet"].reshape(self.N, -1), o["normal_vector_around_feet"], lin]
        else:
            priv = [dr, np.clip((fz - o["height_around_feet"]).reshape(self.N, -1), f32(-1), f32(1)),
                    o["normal_vector_around_feet"], lin]
        if s.obtain_link_contact_states:
            priv.append(o["link_contact_states"])
        o["obs_buf"] = np.clip(obs, -c, c)
        o["privileged_obs_buf"] = np.clip(np.concatenate(priv, axis=1).astype(f32), -c, c)
        o["obs_history"] = st["obs_hist"].copy()
        o["critic_obs_buf"] = st["critic_hist"].copy()
