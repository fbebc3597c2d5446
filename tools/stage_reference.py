#!/usr/bin/env python
"""Stage the reference's Python tree under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that it travels to the GPU
box, where /root/reference does not exist: `legged_gym/`, `rsl_rl/`, `pyproject.toml` and the URDF text files (no meshes).
Used only by the plugin-mode tests (tests/test_plugin_mode.py: the reference's unmodified task classes stepping over the
B200 backend).  Nothing under baseline/_ref is committed or imported by the product package."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(src: str = "/root/reference", dst: str = os.path.join(ROOT, "baseline", "_ref")) -> str:
    if not os.path.isdir(os.path.join(src, "legged_gym")):
        raise RuntimeError(f"{src} holds no reference tree")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    keep = (".py", ".toml", ".urdf", ".md", ".txt")
    for top in ("legged_gym", "rsl_rl", "resources", "pyproject.toml", "LICENSE"):
        s = os.path.join(src, top)
        if os.path.isfile(s):
            shutil.copy2(s, os.path.join(dst, top))
            continue
        for dp, _, files in os.walk(s):
            for f in files:
                if f.endswith(keep):
                    out = os.path.join(dst, os.path.relpath(os.path.join(dp, f), src))
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copy2(os.path.join(dp, f), out)
    return dst


if __name__ == "__main__":
    print(stage(*(sys.argv[1:3])))
