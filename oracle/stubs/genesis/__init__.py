"""A stand-in for the `genesis` engine API that the reference backend calls
(legged_gym/simulator/genesis_simulator.py:27-51,96-133,153-158,231-320,328-371,490-491,627,669-733,768-778),
implemented over the CPU physics oracle (oracle/rigid_oracle.c).

TEST INFRASTRUCTURE ONLY.  It lets the reference's *own, unmodified* Python environment code
(LeggedRobot, Go2TS, GenesisSimulator, ...) run in the build container so that
tools/make_golden.py can record golden vectors and bench.py can time the CPU baseline
("reference env code + restated physics").  The product package never imports it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from hcr_genesis_lr_cl_b200.robot_model import load_robot_model  # noqa: E402  (data format only)
from oracle import physics as _phys  # noqa: E402

cpu, gpu = "cpu", "gpu"
tc_int = torch.int32
device = "cpu"

#: hook the golden harness replaces to inject counter-based random numbers
_rand_hook = None


def init(*a, **k):
    pass


def rand(shape, dtype=float, **k):
    if _rand_hook is not None:
        return _rand_hook(tuple(shape))
    return torch.rand(*shape, dtype=torch.float32)


class _Bag:
    def __init__(self, *a, **k):
        self.__dict__.update(k)


options = types.SimpleNamespace(SimOptions=_Bag, ViewerOptions=_Bag, VisOptions=_Bag, RigidOptions=_Bag)
constraint_solver = types.SimpleNamespace(Newton="Newton", CG="CG")


class _URDF(_Bag):
    kind = "urdf"


class _Terrain(_Bag):
    kind = "terrain"


class _Mesh(_Bag):
    kind = "mesh"


morphs = types.SimpleNamespace(URDF=_URDF, Terrain=_Terrain, Mesh=_Mesh)
sensors = types.SimpleNamespace(DepthCameraPattern=_Bag, DepthCamera=_Bag)

_ASSET_BY_FILE = {"go2.urdf": "go2", "PF_TRON1A/urdf/robot.urdf": "tron1_pf"}


class _Link:
    def __init__(self, name, idx):
        self.name, self.idx = name, idx


class _Joint:
    def __init__(self, dof_start):
        self.dof_start = dof_start


class TerrainEntity:
    def __init__(self, morph):
        self.morph = morph
        self.friction = 1.0

    def set_friction(self, f):
        self.friction = float(f)


class RobotEntity:
    """Robot articulation; dof order = URDF document order (what dof_start resolves against)."""

    def __init__(self, scene, morph):
        import json
        self.scene = scene
        key = next(k for k in _ASSET_BY_FILE if morph.file.replace("\\", "/").endswith(k))
        asset = _ASSET_BY_FILE[key]
        from hcr_genesis_lr_cl_b200.robot_model import ASSET_DIR
        with open(os.path.join(ASSET_DIR, asset + ".json")) as f:
            raw = json.load(f)
        doc_names = [raw["bodies"][b]["joint_name"] for ch in raw["chains"] for b in ch]
        self.model = load_robot_model(asset, doc_names)
        self.joint_names = doc_names
        self.links = [_Link(n, i) for i, n in enumerate(self.model.link_names)]
        self.link_start, self.base_link_idx, self.n_links, self.idx = 0, 0, self.model.nlinks, 0
        self.init_pos = np.array(morph.pos, np.float64)
        self.init_quat = np.array(morph.quat, np.float64)

    # ---- build ----
    def _build(self, n):
        nj = self.model.nj
        self.n = n
        self.state = np.zeros((n, 13))
        self.state[:, 0:3] = self.init_pos
        self.state[:, 3:7] = self.init_quat
        self.q = np.zeros((n, nj))
        self.qd = np.zeros((n, nj))
        self.tau = np.zeros((n, nj))
        self.envp = np.zeros((n, 5))
        self.envp[:, 4] = 1.0
        self.jparam = np.zeros((n, 3 * nj))
        self.jparam[:, :nj] = self.model.body[1:, 19]
        self.link_force = np.zeros((n, self.model.nlinks, 3))
        self.ncontact = np.zeros(n, np.int32)
        self.warm = np.zeros((n, 48))
        self._refresh()

    def _refresh(self):
        self.link_pos, self.link_vel = self.scene.oracle.link_kinematics(self.state, self.q, self.qd)

    def _step(self):
        self.link_force, self.ncontact = self.scene.oracle.substep(self.state, self.q, self.qd, self.tau, self.envp, self.jparam, self.warm)
        self._refresh()

    # ---- helpers ----
    @staticmethod
    def _np(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy().astype(np.float64)
        return np.asarray(x, np.float64)

    @staticmethod
    def _ids(envs_idx, n):
        if envs_idx is None:
            return np.arange(n)
        if isinstance(envs_idx, torch.Tensor):
            return envs_idx.detach().cpu().numpy().astype(np.int64)
        return np.asarray(envs_idx, np.int64)

    @staticmethod
    def _t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)

    def _jidx(self, dofs):
        return [int(d) - 6 for d in dofs]

    # ---- getters ----
    def get_joint(self, name):
        return _Joint(6 + self.joint_names.index(name))

    def get_dofs_limit(self, dofs):
        lim = np.array([self.model.dof_limits[j] for j in self._jidx(dofs)])
        return self._t(lim[:, 0]), self._t(lim[:, 1])

    def get_dofs_force_range(self, dofs):
        e = np.array([self.model.effort[j] for j in self._jidx(dofs)])
        return self._t(-e), self._t(e)

    def get_dofs_position(self, dofs=None):
        if dofs is None:
            raise NotImplementedError
        return self._t(self.q[:, self._jidx(dofs)])

    def get_dofs_velocity(self, dofs=None):
        if dofs is None:
            return self._t(np.concatenate([self.state[:, 7:13], self.qd], axis=1))
        return self._t(self.qd[:, self._jidx(dofs)])

    def get_pos(self):
        return self._t(self.state[:, 0:3])

    def get_quat(self):
        return self._t(self.state[:, 3:7])

    def get_vel(self):
        return self._t(self.state[:, 7:10])

    def get_ang(self):
        return self._t(self.state[:, 10:13])

    def get_links_net_contact_force(self):
        return self._t(self.link_force)

    def get_links_pos(self):
        return self._t(self.link_pos)

    def get_links_vel(self):
        return self._t(self.link_vel)

    # ---- setters ----
    def set_dofs_kp(self, *a, **k):
        pass

    def set_dofs_kv(self, *a, **k):
        pass

    def control_dofs_force(self, force, dofs):
        self.tau[:, self._jidx(dofs)] = self._np(force)

    def set_pos(self, pos, zero_velocity=True, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.state[ids, 0:3] = self._np(pos)
        if zero_velocity:
            self.state[ids, 7:13] = 0
            self.qd[ids] = 0
        self._refresh()

    def set_quat(self, quat, zero_velocity=True, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.state[ids, 3:7] = self._np(quat)
        if zero_velocity:
            self.state[ids, 7:13] = 0
            self.qd[ids] = 0
        self._refresh()

    def zero_all_dofs_velocity(self, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.state[ids, 7:13] = 0
        self.qd[ids] = 0
        self._refresh()

    def set_dofs_position(self, position, dofs_idx_local=None, zero_velocity=True, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.q[np.ix_(ids, self._jidx(dofs_idx_local))] = self._np(position)
        if zero_velocity:
            self.state[ids, 7:13] = 0
            self.qd[ids] = 0
        self._refresh()

    def set_dofs_velocity(self, velocity, dofs_idx_local=None, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        vel = self._np(velocity)
        dofs = list(range(6 + self.model.nj)) if dofs_idx_local is None else [int(d) for d in dofs_idx_local]
        full = np.concatenate([self.state[ids, 7:13], self.qd[ids]], axis=1)
        full[:, dofs] = vel
        self.state[ids, 7:13] = full[:, :6]
        self.qd[ids] = full[:, 6:]
        self._refresh()

    def set_friction_ratio(self, ratios, links_idx, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.envp[ids, 4] = self._np(ratios)[:, 0]

    def set_mass_shift(self, shift, link_idx, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.envp[ids, 0] = self._np(shift).reshape(len(ids))

    def set_COM_shift(self, shift, link_idx, envs_idx=None):
        ids = self._ids(envs_idx, self.n)
        self.envp[ids, 1:4] = self._np(shift).reshape(len(ids), 3)

    def _set_jparam(self, k, values, dofs, envs_idx):
        ids = self._ids(envs_idx, self.n)
        nj = self.model.nj
        cols = [k * nj + j for j in self._jidx(dofs)]
        self.jparam[np.ix_(ids, cols)] = self._np(values)

    def set_dofs_armature(self, values, dofs, envs_idx=None):
        self._set_jparam(0, values, dofs, envs_idx)

    def set_dofs_damping(self, values, dofs, envs_idx=None):
        self._set_jparam(1, values, dofs, envs_idx)

    def set_dofs_frictionloss(self, values, dofs, envs_idx=None):
        self._set_jparam(2, values, dofs, envs_idx)


class Scene:
    #: physics parameter overrides applied by the harness before build (e.g. PGS iterations)
    param_overrides = {}

    def __init__(self, sim_options=None, viewer_options=None, vis_options=None, rigid_options=None, show_viewer=False, **k):
        self.sim_options, self.rigid_options = sim_options, rigid_options
        self.robot, self.terrain = None, None
        self.viewer = types.SimpleNamespace(set_camera_pose=lambda **kw: None)
        self.n_steps = 0

    def add_entity(self, morph, **k):
        if morph.kind == "urdf" and "plane" in morph.file:
            self.terrain = TerrainEntity(morph)
            return self.terrain
        if morph.kind == "urdf":
            self.robot = RobotEntity(self, morph)
            return self.robot
        self.terrain = TerrainEntity(morph)
        return self.terrain

    def add_sensor(self, *a, **k):
        raise NotImplementedError("depth sensors are outside the hot path")

    def build(self, n_envs=1, **k):
        tm = self.terrain.morph
        hf = None
        prm = _phys.default_params(dt=float(self.sim_options.dt))
        if tm.kind == "terrain":
            hf = np.asarray(tm.height_field, np.int16)
            prm[_phys.P_HSCALE], prm[_phys.P_VSCALE], prm[_phys.P_BORDER] = tm.horizontal_scale, tm.vertical_scale, -tm.pos[0]
        prm[_phys.P_TERRAIN_MU] = self.terrain.friction
        for k_, v in Scene.param_overrides.items():
            prm[k_] = v
        self.params, self.heightfield = prm, hf
        self.oracle = _phys.PhysicsOracle(self.robot.model, prm, hf, precision=os.environ.get("ORACLE_PRECISION", "f64"))
        self.robot._build(n_envs)

    def step(self):
        # terrain friction may be set after build (genesis_simulator.py:276)
        self.oracle.prm[_phys.P_TERRAIN_MU] = self.terrain.friction
        self.robot._step()
        self.n_steps += 1

    def clear_debug_objects(self):
        pass
