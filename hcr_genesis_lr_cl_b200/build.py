"""Build libb200step.so in-tree with nvcc for sm_100a (B200).  `python -m hcr_genesis_lr_cl_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# translation units, compiled side by side: the C ABI with the dynamics kernels | the host-stepped env kernels | the
# device-stepped env kernels (b200_env_step_device).  The env kernels restate torch fp32 eager arithmetic: no FMA contraction
# beyond the explicit fmaf() calls (csrc/env_kernel.cuh).
UNITS = [("b200_step.cu", []), ("env_kernels_host.cu", ["-fmad=false"]), ("env_kernels_dev.cu", ["-fmad=false"])]


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
    objdir = os.path.join(_PKG, "_build")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, u.replace(".cu", ".o")) for u, _ in UNITS]
    cmds = [[nvcc] + NVCC_FLAGS + fl + extra + ["-c", "-o", o, os.path.join(_CSRC, u)] for (u, fl), o in zip(UNITS, objs)]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
    log, failed = "", False
    for c, p in zip(cmds, procs):
        o, _ = p.communicate()
        log += " ".join(c) + "\n" + o
        failed = failed or p.returncode != 0
    if not failed:
        link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs
        res = subprocess.run(link, capture_output=True, text=True)
        log += " ".join(link) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(_PKG, "build.log"), "w") as f:
        f.write(log)
    if verbose or failed:
        print(log)
    if failed:
        raise RuntimeError("nvcc failed, see hcr_genesis_lr_cl_b200/build.log")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
