"""Helpers shared by the golden-vector tests (CPU oracle and CUDA path replay the same recordings)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    s0 = {k[3:]: v for k, v in g.items() if k.startswith("s0/")}
    return g, s0


def load_terrain(spec=None):
    from hcr_genesis_lr_cl_b200.terrain_assets import load_go2_rough_terrain, terrain_for
    return load_go2_rough_terrain() if spec is None else terrain_for(spec)


def spec_for(g):
    from hcr_genesis_lr_cl_b200 import task_spec as T
    spec = T.PRESETS[str(g["meta/task"])]()
    spec.seed = int(g["meta/seed"])
    spec.contact_state_link_names = [str(x) for x in g["meta/contact_links"]]
    if "meta/gait_smooth" in g and int(g["meta/gait_smooth"]):          # recorded with the von Mises ("smooth") gait indicator
        spec.gait_smooth = True
    if "meta/reward_names" in g:                                        # the reward scales of the recording (an edited config)
        spec.reward_scales = {str(k): float(v) for k, v in zip(g["meta/reward_names"], g["meta/reward_scales"])}
    if "meta/ctrl_delay" in g and int(g["meta/ctrl_delay"]) > 0:       # recorded with domain_rand.randomize_ctrl_delay on
        spec.randomize_ctrl_delay, spec.ctrl_delay_step_range = True, [0, int(g["meta/ctrl_delay"])]
    return spec


def phys_at(g, t):
    return {k[5:]: v[t] for k, v in g.items() if k.startswith("phys/")}


def out_at(g, t):
    return {k[4:]: v[t] for k, v in g.items() if k.startswith("out/")}
