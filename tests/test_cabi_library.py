"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, exports every symbol the header declares,
fails loudly without a GPU, and the product package never touches the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from hcr_genesis_lr_cl_b200 import _cabi, task_spec as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols(built_library):
    header = open(os.path.join(ROOT, "include", "b200_step.h")).read()
    declared = set(re.findall(r"\b(b200_\w+)\s*\(", header))
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", built_library], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200_\w+)", out))
    assert declared <= exported, declared - exported
    lib = _cabi.load_library()
    for s in declared:
        assert hasattr(lib, s)


def test_library_contains_sm100a_code(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_library):
    lib = _cabi.load_library()
    spec = T.go2_spec()
    model = spec.load_model()
    tf, ti = _cabi.pack_task(spec, model, 8)
    mi, mf = model.packed_ints(), model.packed_floats()
    h = ctypes.c_void_p()
    rc = lib.b200_create(mi.ctypes.data, mi.size, mf.ctypes.data, mf.size, ti.ctypes.data, ti.size, tf.ctypes.data, tf.size, ctypes.byref(h))
    assert rc != 0 and lib.b200_last_error()
    from hcr_genesis_lr_cl_b200.fused_env import make_env
    with pytest.raises(RuntimeError, match="no CPU path"):
        make_env("go2", 8, device="cuda:0")
    with pytest.raises(RuntimeError):
        make_env("go2", 8, device="cpu")


def test_descriptor_size_mismatch_is_rejected(built_library):
    lib = _cabi.load_library()
    z = np.zeros(4, np.float32)
    zi = np.zeros(4, np.int32)
    h = ctypes.c_void_p()
    assert lib.b200_create(zi.ctypes.data, 4, z.ctypes.data, 4, zi.ctypes.data, 4, z.ctypes.data, 4, ctypes.byref(h)) != 0
    assert b"descriptor" in lib.b200_last_error() or b"CUDA" in lib.b200_last_error() or b"cuda" in lib.b200_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hcr_genesis_lr_cl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                # the two files that can be compiled for the TEST-ONLY warp emulator name it (under #ifdef B200_WARP_EMU)
                assert "warp_emu" not in src or f in ("cuda_compat.cuh", "b200_step.cu"), f


def test_simulator_api_covers_the_reference_abc():
    """Every abstract method / property of legged_gym.simulator.Simulator exists on B200Simulator."""
    from oracle.ref_harness import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import_reference()
    from legged_gym.simulator.simulator import Simulator
    from hcr_genesis_lr_cl_b200.simulator import B200Simulator
    missing = [n for n in dir(Simulator) if not n.startswith("__") and n != "_abc_impl" and not hasattr(B200Simulator, n)]
    assert not missing, missing
    for n in ("_friction_values", "_added_base_mass", "_base_com_bias", "_rand_push_vels", "_kp_scale", "_kd_scale",
              "_joint_armature", "_joint_friction", "_joint_damping", "dof_names"):
        assert hasattr(B200Simulator, n), n
