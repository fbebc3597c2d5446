"""Rigid-body model of one robot, packed for the C-ABI.

The numbers come from the URDF the reference loads with
``gs.morphs.URDF(merge_fixed_links=True, links_to_keep=...)``
(legged_gym/simulator/genesis_simulator.py:303-311); ``tools/extract_robot_model.py``
derives ``assets/<robot>.json`` from it once.  At load time the serial chains are
re-ordered to follow ``cfg.asset.dof_names`` (the order of actions/observations,
legged_gym/envs/base/common_cfgs.py:52-65) while the *reporting links* keep URDF
document order, which is what the reference resolves link names against
(genesis_simulator.py:333-363): feet enumerate FL,FR,RL,RR although dofs
enumerate FR,FL,RR,RL.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
BODY_STRIDE = 20  # jp[3] ax[3] com[3] mass I[6] lo hi effort armature
MAX_SPHERES = 64
MAX_JOINTS = 12

#: joint armature the engine applies to URDF joints when none is given
#: [UPSTREAM-UNVERIFIED, SURVEY Appendix C.1]
DEFAULT_ARMATURE = 0.1


@dataclass
class RobotModel:
    name: str
    num_chains: int
    chain_len: int
    dof_names: List[str]
    link_names: List[str]
    body: np.ndarray        # [nb, BODY_STRIDE] float32, body 0 = base, then chains in dof order
    link_off: np.ndarray    # [nlinks, 3] float32
    link_body: np.ndarray   # [nlinks] int32
    sph: np.ndarray         # [nspheres, 4] float32
    sph_body: np.ndarray    # [nspheres] int32
    sph_link: np.ndarray    # [nspheres] int32
    total_mass: float
    dof_limits: np.ndarray = field(default=None)   # [nj, 2] hard URDF limits (dof order)
    effort: np.ndarray = field(default=None)       # [nj]
    velocity: np.ndarray = field(default=None)     # [nj]

    @property
    def nb(self) -> int:
        return self.body.shape[0]

    @property
    def nj(self) -> int:
        return self.nb - 1

    @property
    def nlinks(self) -> int:
        return self.link_off.shape[0]

    @property
    def nspheres(self) -> int:
        return self.sph.shape[0]

    def find_link_indices(self, names: Sequence[str]) -> List[int]:
        """Substring match in link order (genesis_simulator.py:333-342)."""
        return [i for i, ln in enumerate(self.link_names) if any(n in ln for n in names)]

    def packed_ints(self) -> np.ndarray:
        head = np.array([self.num_chains, self.chain_len, self.nb, self.nj, self.nlinks, self.nspheres, 0, 0], np.int32)
        return np.ascontiguousarray(np.concatenate([head, self.link_body, self.sph_body, self.sph_link]).astype(np.int32))

    def packed_floats(self) -> np.ndarray:
        return np.ascontiguousarray(
            np.concatenate([self.body.ravel(), self.link_off.ravel(), self.sph.ravel()]).astype(np.float32))


def load_robot_model(robot: str, dof_names: Sequence[str], armature: float = DEFAULT_ARMATURE) -> RobotModel:
    path = robot if os.path.isabs(robot) else os.path.join(ASSET_DIR, f"{robot}.json")
    with open(path) as f:
        m = json.load(f)
    bodies, chains = m["bodies"], m["chains"]
    D = m["chain_len"]
    by_joint = {b["joint_name"]: i for i, b in enumerate(bodies) if b["joint_name"]}
    if len(dof_names) != len(by_joint):
        raise ValueError(f"{robot}: expected {len(by_joint)} dof names, got {len(dof_names)}")
    if len(dof_names) > MAX_JOINTS:
        raise ValueError(f"{robot}: more than {MAX_JOINTS} actuated joints are not supported")
    # chain order that makes internal joint order == dof_names order
    order = []
    for c in range(len(chains)):
        first = dof_names[c * D]
        if first not in by_joint:
            raise ValueError(f"{robot}: unknown joint {first!r}")
        ch = next((ch for ch in chains if ch[0] == by_joint[first]), None)
        if ch is None:
            raise ValueError(f"{robot}: dof_names must list each leg root-to-tip ({first!r} is not a chain root)")
        for k in range(D):
            if bodies[ch[k]]["joint_name"] != dof_names[c * D + k]:
                raise ValueError(f"{robot}: dof_names must list each leg root-to-tip ({dof_names[c*D+k]!r})")
        order.append(ch)
    new_index = {0: 0}
    for c, ch in enumerate(order):
        for k, b in enumerate(ch):
            new_index[b] = 1 + c * D + k
    nb = len(bodies)
    body = np.zeros((nb, BODY_STRIDE), np.float32)
    for old, b in enumerate(bodies):
        r = body[new_index[old]]
        r[0:3], r[3:6], r[6:9], r[9] = b["joint_pos"], b["joint_axis"], b["com"], b["mass"]
        r[10:16] = b["inertia"]
        r[16:18], r[18] = b["limit"], b["effort"]
        r[19] = armature if old != 0 else 0.0
    links = m["links"]
    sph = m["spheres"]
    if len(sph) > MAX_SPHERES:
        raise ValueError(f"{robot}: {len(sph)} collision spheres > {MAX_SPHERES}")
    jb = [by_joint[n] for n in dof_names]
    return RobotModel(
        name=m["name"], num_chains=len(chains), chain_len=D, dof_names=list(dof_names),
        link_names=[l["name"] for l in links], body=body,
        link_off=np.array([l["offset"] for l in links], np.float32),
        link_body=np.array([new_index[l["body"]] for l in links], np.int32),
        sph=np.array([s["pos"] + [s["radius"]] for s in sph], np.float32),
        sph_body=np.array([new_index[s["body"]] for s in sph], np.int32),
        sph_link=np.array([s["link"] for s in sph], np.int32),
        total_mass=float(m["total_mass"]),
        dof_limits=np.array([bodies[i]["limit"] for i in jb], np.float32),
        effort=np.array([bodies[i]["effort"] for i in jb], np.float32),
        velocity=np.array([bodies[i]["velocity"] for i in jb], np.float32),
    )


GO2_DOF_NAMES = [f"{leg}_{j}_joint" for leg in ("FR", "FL", "RR", "RL") for j in ("hip", "thigh", "calf")]
TRON1_PF_DOF_NAMES = [f"{j}_{s}_Joint" for s in ("L", "R") for j in ("abad", "hip", "knee")]
