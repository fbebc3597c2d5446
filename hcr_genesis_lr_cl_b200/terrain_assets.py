"""Packaged terrain for the Go2 rough-terrain tasks.

``assets/go2_rough_terrain.npz`` is the int16 heightfield and the sub-terrain origins that the reference's own
generator produces for ``Go2RoughCommonCfg.terrain`` (legged_gym/utils/terrain.py:37-83 with
legged_gym/envs/base/common_cfgs.py:75-93, ``np.random.seed(1)`` via set_seed): 1200 x 1200 samples, 10 x 10
sub-terrains of 8 m, 20 m border, 0.1 m / 0.005 m scales.  It is written by tools/make_golden.py in the build
container; inside LeggedGym-Ex the backend calls the reference generator directly instead
(simulator.py ``_create_sim``).
"""
import os

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def load_terrain(name: str):
    """`go2_rough` (Go2RoughCommonCfg.terrain) or `tron1_rough` (TRON1PF_EECfg.terrain: 15 m border, 1100 x 1100)."""
    z = np.load(os.path.join(_ASSETS, f"{name}_terrain.npz"))
    return z["height_samples"].astype(np.int16), z["terrain_origins"].astype(np.float32)


def load_go2_rough_terrain():
    return load_terrain("go2_rough")


def terrain_for(spec):
    """Packaged terrain matching a task preset (None for plane tasks)."""
    if not spec.heightfield:
        return None
    return load_terrain("tron1_rough" if spec.robot == "tron1_pf" else "go2_rough")
