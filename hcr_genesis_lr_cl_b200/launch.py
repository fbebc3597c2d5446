"""Run one of the reference's own scripts, unmodified, on the B200 backend:

    python -m hcr_genesis_lr_cl_b200.launch [--b200-mode fused|plugin] [--b200-reference-root DIR] [--b200-stub-missing-imports] \\
        legged_gym/scripts/train.py --task go2_ts --headless --num_envs 4096 --max_iterations 1500

does what `python legged_gym/scripts/train.py ...` does (scripts/train.py:8-36: `gs.init`, `task_registry.make_env`,
`task_registry.make_alg_runner`, `runner.learn`) after `plugin.install()`: nothing in the reference tree is edited.

* ``--b200-mode fused`` (default): `task_registry.make_env` hands out a `FusedLeggedEnv` for every task with a fused
  descriptor (the whole `env.step` is one C-ABI call); other tasks fall back to plugin mode with a warning.
* ``--b200-mode plugin``: the reference's task classes (torch eager) over `B200Simulator`.
* ``--b200-reference-root``: the LeggedGym-Ex checkout (default: the directory two levels above the script, i.e. the tree the
  script lives in).
* ``--b200-stub-missing-imports``: legged_gym imports trimesh / matplotlib / pygame / xlsxwriter at module level although the
  training path never calls them; in a container without those packages, empty placeholder modules let the imports pass.
"""
from __future__ import annotations

import os
import runpy
import sys

from . import plugin


def main(argv=None, impl=None, fused_env_class=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    mode, root, stub = "fused", None, False
    while argv and argv[0].startswith("--b200-"):
        flag = argv.pop(0)
        if flag == "--b200-mode":
            mode = argv.pop(0)
        elif flag == "--b200-reference-root":
            root = argv.pop(0)
        elif flag == "--b200-stub-missing-imports":
            stub = True
        else:
            raise SystemExit(f"unknown option {flag}")
    if mode not in ("fused", "plugin"):
        raise SystemExit("--b200-mode takes fused or plugin")
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    if root is None:
        root = os.path.dirname(os.path.dirname(os.path.dirname(script)))      # <root>/legged_gym/scripts/train.py
    if not os.path.isdir(os.path.join(root, "legged_gym")):
        raise SystemExit(f"{root} is not a LeggedGym-Ex checkout (no legged_gym/): pass --b200-reference-root")
    if stub:
        plugin.stub_missing_optional_imports()
    plugin.install(reference_root=root, impl=impl, fused=(mode == "fused"), fused_env_class=fused_env_class)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
