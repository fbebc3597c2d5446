"""Build libb200step.so in-tree with nvcc for sm_100a (B200).  `python -m hcr_genesis_lr_cl_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC))] + [os.path.join(os.path.dirname(_PKG), "include", "b200_step.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(_PKG, "libb200step.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in sources() if os.path.exists(s)):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (and there is no fallback)")
    extra = os.environ.get("B200_NVCC_EXTRA", "").split()      # tuning experiments, e.g. -DDYN_MIN_BLOCKS=5
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out, os.path.join(_CSRC, "b200_step.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(_PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see hcr_genesis_lr_cl_b200/build.log")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
