"""B200-native physics + environment step for LeggedGym-Ex (oscar-youngquist/HCR_Genesis_LR_CL).

Public surface:
    B200Simulator   -- backend for legged_gym.simulator.Simulator (plugin API)      simulator.py
    FusedLeggedEnv  -- VecEnv whose step() is two CUDA kernels                      fused_env.py
    TaskSpec, go2_spec, go2_ts_spec -- task descriptors                             task_spec.py
    build           -- compile libb200step.so in-tree with nvcc for sm_100a        build.py
Importing the package does not need a GPU; constructing a simulator does (no CPU fallback).
"""
from .task_spec import TaskSpec, go2_spec, go2_ts_spec, PRESETS  # noqa: F401
from .robot_model import load_robot_model, RobotModel  # noqa: F401


def __getattr__(name):
    if name in ("B200Simulator",):
        from .simulator import B200Simulator
        return B200Simulator
    if name in ("FusedLeggedEnv", "make_env"):
        from . import fused_env
        return getattr(fused_env, name)
    raise AttributeError(name)
