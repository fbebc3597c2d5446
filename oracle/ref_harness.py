"""Run the reference's own (unmodified) environment code over the CPU physics oracle.

TEST INFRASTRUCTURE ONLY; needs /root/reference (build container).  Used by
tools/make_golden.py (golden vectors) and tests that cross-check oracle/env_oracle.py.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference() -> str:
    """/root/reference in the build container; on the GPU box the staged copy under baseline/_ref (git-ignored, made by
    tools/stage_reference.py: python sources only, nothing of it is committed)."""
    for p in (os.environ.get("HCR_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "legged_gym")):
            return p
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legged_gym"))


def import_reference():
    """Put the stubs + reference on sys.path and import legged_gym.envs (registers the tasks)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    os.environ["SIMULATOR"] = "genesis"
    for p in (REFERENCE_ROOT, os.path.join(_HERE, "stubs"), os.path.join(_HERE, "stubs_thirdparty")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import legged_gym.envs  # noqa: F401
    from legged_gym.envs.base.legged_robot import LeggedRobot
    # go2_wtw / tron1_pf_ee call a method name that does not exist; harness-level alias.
    if not hasattr(LeggedRobot, "update_command_curriculum"):
        LeggedRobot.update_command_curriculum = LeggedRobot._update_command_curriculum
    from legged_gym.simulator.genesis_simulator import GenesisSimulator
    # tron1_pf_ee needs simulator.dof_names.
    if not hasattr(GenesisSimulator, "dof_names"):
        GenesisSimulator.dof_names = property(lambda self: self._dof_names)
    from legged_gym.utils.task_registry import task_registry
    return task_registry


def make_env(task: str, num_envs: int, cfg_edit=None):
    task_registry = import_reference()
    args = SimpleNamespace(task=task, headless=True, cpu=True, num_envs=num_envs, debug=False, max_iterations=None,
                           resume=False, sync_wandb=False, ckpt=-1, load_run=None, export_onnx=False,
                           use_joystick=False, joystick_type=None, follow_robot=False, sim_device="cpu", rl_device="cpu")
    env_cfg, train_cfg = task_registry.get_cfgs(name=task)
    if cfg_edit is not None:
        cfg_edit(env_cfg)
    import torch
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", os.cpu_count() or 1)))
    env, env_cfg = task_registry.make_env(task, args=args, env_cfg=env_cfg)
    return env, env_cfg, train_cfg
