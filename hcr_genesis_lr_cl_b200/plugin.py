"""Wire the B200 backend into an UNMODIFIED LeggedGym-Ex tree at run time (the import-hook twin of
``overlay/b200_backend.patch``).

``install()`` makes ``BaseTask.__init__`` (legged_gym/envs/base/base_task.py:41-48) build a ``B200Simulator`` -- a
subclass of the reference's own ``Simulator`` ABC -- wherever it would have built a ``GenesisSimulator``; nothing in the
reference tree is edited:

* ``SIMULATOR=genesis`` is what the unpatched ``legged_gym/__init__.py:9-15`` accepts and what makes every config class
  pick the heightfield terrain and the Genesis run names (common_cfgs.py:76-79, tron1_pf_ee_config.py:19-22), i.e. the
  configuration this backend implements;
* when the ``genesis`` engine is not installed a placeholder module satisfies ``import genesis as gs``
  (legged_gym/__init__.py:21, scripts/train.py:10-12 call ``gs.init``); the B200 path never calls into it;
* ``base_task.GenesisSimulator`` (the module-global name the dispatch uses) is rebound to the B200 backend class.

Usage::

    import hcr_genesis_lr_cl_b200.plugin as b200
    b200.install()                      # imports legged_gym.envs (registers the tasks) and rebinds the backend
    from legged_gym.utils import task_registry ...   # train.py / play.py as shipped
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Optional

_BACKEND = None
_ORIGINAL = None     # base_task.GenesisSimulator before install()


def _placeholder_genesis() -> types.ModuleType:
    m = types.ModuleType("genesis")
    m.__doc__ = "placeholder installed by hcr_genesis_lr_cl_b200.plugin: the Genesis engine is absent and the B200 backend does not use it"
    m.cpu, m.gpu = "cpu", "gpu"
    m.init = lambda *a, **k: None

    def _unavailable(name):
        raise AttributeError(f"genesis.{name}: the Genesis engine is not installed (B200 backend active)")
    m.__getattr__ = _unavailable
    return m


def backend_class(impl=None):
    """``B200Simulator`` as a subclass of ``legged_gym.simulator.simulator.Simulator`` (what the overlay file defines).
    ``impl`` replaces the implementation class (tests run the host layer over the warp emulator this way)."""
    from legged_gym.simulator.simulator import Simulator
    if impl is None:
        from .simulator import B200Simulator as impl

    class B200Simulator(impl, Simulator):
        def __init__(self, cfg, sim_params, sim_device="cuda:0", headless=True):
            impl.__init__(self, cfg, sim_params, sim_device, headless)

    B200Simulator.__doc__ = impl.__doc__
    return B200Simulator


_MAKE_ENV = None     # task_registry.make_env before install(fused=True)


def stub_missing_optional_imports() -> list:
    """legged_gym imports trimesh (utils/terrain.py:32), matplotlib (utils/logger.py:4-5), pygame (the joystick script) and
    xlsxwriter (utils/logger.py:6) at module level; none of them is called on the training path of a heightfield task.  Where
    one is not installed, an empty placeholder module lets the import pass (returns the names it stood in for)."""
    made = []
    for name in ("trimesh", "matplotlib", "matplotlib.pyplot", "pygame", "xlsxwriter"):
        try:
            importlib.import_module(name)
        except ImportError:
            m = types.ModuleType(name)
            m.__doc__ = "placeholder installed by hcr_genesis_lr_cl_b200.plugin.stub_missing_optional_imports"
            if name == "matplotlib":
                m.use = lambda *a, **k: None
            if name == "matplotlib.pyplot":
                m.rcParams = {}                      # legged_gym/utils/logger.py updates it at import time
            if name == "trimesh":
                m.Trimesh = type("Trimesh", (), {"__init__": lambda self, *a, **k: None})
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
            made.append(name)
    return made


def _route_make_env_to_fused(env_class=None):
    """`task_registry.make_env(name, args)` (task_registry.py:35-72, what scripts/train.py calls) hands out a `FusedLeggedEnv`
    for every task whose cfg has a fused descriptor -- the reference's cfg handling (get_cfgs, update_cfg_from_args, set_seed)
    and its terrain generator are used as they are; anything else (an unsupported task, reward term or asset) falls back to
    the task class over the plugin backend with a warning."""
    global _MAKE_ENV
    import warnings
    tr = importlib.import_module("legged_gym.utils.task_registry")
    reg = tr.task_registry
    if _MAKE_ENV is None:
        _MAKE_ENV = reg.make_env

    def make_env(name, args=None, env_cfg=None):
        from .fused_env import FusedLeggedEnv
        from .task_spec import TaskSpec
        if args is None:
            args = tr.get_args()
        if name not in reg.task_classes:
            raise ValueError(f"Task with name: {name} was not registered")
        if env_cfg is None:
            env_cfg, _ = reg.get_cfgs(name)
        env_cfg, _ = tr.update_cfg_from_args(env_cfg, None, args)
        try:
            spec = TaskSpec.from_reference_cfg(env_cfg, name)
        except (ValueError, KeyError, AttributeError) as e:
            warnings.warn(f"B200 backend: task {name!r} has no fused descriptor ({e}); running the reference task class in plugin mode")
            return _MAKE_ENV(name, args=args, env_cfg=env_cfg)
        tr.set_seed(env_cfg.seed)
        terrain = None
        if spec.heightfield:
            from legged_gym.utils.terrain import Terrain        # the reference generator, unchanged (genesis_simulator.py:267-269)
            t = Terrain(env_cfg.terrain)
            terrain = (t.heightsamples, t.env_origins)
        cls = env_class or FusedLeggedEnv
        env = cls(spec, int(env_cfg.env.num_envs), "cpu" if args.cpu else "cuda:0", terrain=terrain, cfg=env_cfg)
        return env, env_cfg

    reg.make_env = make_env


def install(reference_root: Optional[str] = None, impl=None, fix_reference_defects: bool = True, fused: bool = False, fused_env_class=None):
    """Route the reference's backend dispatch to the B200 backend; returns the backend class.

    ``fused=True``: additionally ``task_registry.make_env`` returns a ``FusedLeggedEnv`` (whole ``env.step`` = one C-ABI
    call, same return tuples) for the tasks with a fused descriptor, so that ``scripts/train.py --task go2_ts`` runs fused
    without any edit of the reference tree.

    ``fix_reference_defects``: also add the two harness-level aliases without which ``go2_wtw`` / ``tron1_pf_ee`` crash on
    any backend as shipped (``update_command_curriculum`` is called but only ``_update_command_curriculum``
    exists, go2_wtw.py:121, tron1_pf_ee.py:201).  ``dof_names`` (R3) is a property of the backend itself."""
    global _BACKEND
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    os.environ.setdefault("SIMULATOR", "genesis")
    if os.environ["SIMULATOR"] != "genesis":
        raise RuntimeError("plugin.install() rides on the SIMULATOR=genesis configuration branches; unset SIMULATOR or use the patch overlay")
    if "genesis" not in sys.modules:
        try:
            importlib.import_module("genesis")
        except ImportError:
            sys.modules["genesis"] = _placeholder_genesis()
    importlib.import_module("legged_gym.envs")     # the entry module of the reference's own scripts (train.py:4); importing
    import legged_gym.simulator as sim_pkg         # legged_gym.simulator first trips over the tree's circular imports
    cls = backend_class(impl)
    sim_pkg.B200Simulator = cls
    bt = importlib.import_module("legged_gym.envs.base.base_task")
    global _ORIGINAL
    if _ORIGINAL is None:
        _ORIGINAL = bt.GenesisSimulator
    bt.GenesisSimulator = cls
    if fix_reference_defects:
        from legged_gym.envs.base.legged_robot import LeggedRobot
        if not hasattr(LeggedRobot, "update_command_curriculum"):
            LeggedRobot.update_command_curriculum = LeggedRobot._update_command_curriculum
    _BACKEND = cls
    if fused:
        _route_make_env_to_fused(fused_env_class)
    return cls


def installed():
    return _BACKEND


def uninstall() -> None:
    """Give BaseTask its original Genesis backend back (tests that run both harnesses in one process)."""
    global _BACKEND, _ORIGINAL, _MAKE_ENV
    if _ORIGINAL is not None:
        bt = importlib.import_module("legged_gym.envs.base.base_task")
        bt.GenesisSimulator = _ORIGINAL
    if _MAKE_ENV is not None:
        importlib.import_module("legged_gym.utils.task_registry").task_registry.__dict__.pop("make_env", None)
    _BACKEND = _ORIGINAL = _MAKE_ENV = None
