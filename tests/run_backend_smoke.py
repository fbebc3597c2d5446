"""TEST-ONLY subprocess body: build a reference task over the B200 backend in a fresh interpreter and print a digest.

  python tests/run_backend_smoke.py <reference_root> <variant: hook|patched> <impl: emu|cuda> <task> <envs> <steps>

`hook`    = hcr_genesis_lr_cl_b200.plugin.install() on the unmodified tree (SIMULATOR=genesis, placeholder `genesis` module);
`patched` = a tree with overlay/b200_backend.patch applied and overlay/legged_gym/simulator/b200_simulator.py dropped in
            (SIMULATOR=b200), exactly what INTEGRATION.md section 2 asks a maintainer to do.
The reference's own draws come from torch's global generator (seeded by task_registry.make_env -> set_seed), the
backend's from Philox: both variants must print the same digest."""
import hashlib
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref_root, variant, impl, task, n_envs, steps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    for p in (os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "oracle", "stubs_thirdparty"), ref_root):
        sys.path.insert(0, p)
    import torch
    emu = None
    if impl == "emu":
        from emu_backend import EmuB200Simulator as emu
    if variant == "patched":
        os.environ["SIMULATOR"] = "b200"
        if emu is not None:                      # the overlay file imports the implementation class by this name
            import hcr_genesis_lr_cl_b200.simulator as impl_mod
            impl_mod.B200Simulator = emu
        import legged_gym.envs  # noqa: F401
        from legged_gym.envs.base.legged_robot import LeggedRobot
        if not hasattr(LeggedRobot, "update_command_curriculum"):      # SURVEY R2 (reference defect, any backend)
            LeggedRobot.update_command_curriculum = LeggedRobot._update_command_curriculum
    else:
        os.environ.pop("SIMULATOR", None)
        from hcr_genesis_lr_cl_b200 import plugin
        plugin.install(impl=emu)
    import legged_gym
    from legged_gym.simulator import Simulator
    from legged_gym.utils.task_registry import task_registry
    args = SimpleNamespace(task=task, headless=True, cpu=impl == "emu", num_envs=n_envs, debug=False, max_iterations=None, resume=False,
                           sync_wandb=False, ckpt=-1, load_run=None, export_onnx=False, use_joystick=False, joystick_type=None, follow_robot=False)
    env, cfg = task_registry.make_env(task, args=args)
    dev = env.device
    env.reset()
    g = torch.Generator().manual_seed(7)
    h = hashlib.sha256()
    ret = None
    for _ in range(steps):
        ret = env.step((0.5 * torch.randn(n_envs, env.num_actions, generator=g)).to(dev))
        h.update(ret[0].detach().cpu().numpy().tobytes())
        h.update(env.rew_buf.detach().cpu().numpy().tobytes())
    sim = env.simulator
    print("DIGEST " + json.dumps(dict(
        variant=variant, simulator=legged_gym.SIMULATOR, backend=type(sim).__name__, backend_module=type(sim).__module__,
        is_simulator_abc=isinstance(sim, Simulator), genesis=getattr(sys.modules.get("genesis"), "__doc__", "") or "",
        task_class=type(env).__name__, mesh_type=cfg.terrain.mesh_type, obs_shape=list(ret[0].shape),
        finite=bool(torch.isfinite(ret[0]).all()), sha=h.hexdigest(), launches=int(sim.launch_count))))


if __name__ == "__main__":
    main()
