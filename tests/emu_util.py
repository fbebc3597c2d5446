"""TEST-ONLY helpers: run the CUDA kernel sources on the host warp emulator (tests/warp_emu) with numpy buffers,
and drive the CPU oracle through the same policy step for comparison."""
import ctypes
import os
import subprocess

import numpy as np

from hcr_genesis_lr_cl_b200 import _cabi
from oracle import physics as ophys

_EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_emu")
_CSRC = os.path.join(os.path.dirname(_EMU_DIR), "..", "hcr_genesis_lr_cl_b200", "csrc")


def build_emu(with_env=True):
    out = os.path.join(_EMU_DIR, "libwarpemu.so")
    srcs = [os.path.join(_EMU_DIR, f) for f in ("emu.cpp", "emu_main.cpp", "emu.h")]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    srcs.append(os.path.join(os.path.dirname(_EMU_DIR), "..", "include", "b200_step.h"))
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cmd = ["g++", "-DEMU_WITH_ENV", "-DENV_WARPS_PER_BLOCK=1", "-O1", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-I", _EMU_DIR, "-o", out,
               os.path.join(_EMU_DIR, "emu.cpp"), os.path.join(_EMU_DIR, "emu_main.cpp")]
        subprocess.check_call(cmd)
    return ctypes.CDLL(out)


def build_emu_cabi():
    """TEST-ONLY: hcr_genesis_lr_cl_b200/csrc/b200_step.cu compiled by g++ over the warp emulator and a synchronous CUDA-runtime
    shim (warp_emu/cuda_shim.h) -> tests/warp_emu/libb200step_emu.so with the same C ABI as the product library, host memory
    instead of device memory.  Lets the host layers above the C ABI (B200Simulator, FusedLeggedEnv, the plugin overlay) run in
    a container without a GPU; the product never loads it."""
    out = os.path.join(_EMU_DIR, "libb200step_emu.so")
    srcs = [os.path.join(_EMU_DIR, f) for f in ("emu.cpp", "emu.h", "cuda_shim.h")]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)]
    srcs.append(os.path.join(os.path.dirname(_EMU_DIR), "..", "include", "b200_step.h"))
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cmd = ["g++", "-DB200_WARP_EMU=1", "-DENV_WARPS_PER_BLOCK=1", "-DDYN_WARPS_PER_BLOCK=1", "-O1", "-fPIC", "-shared", "-std=c++17",
               "-ffp-contract=off", "-I", _EMU_DIR, "-o", out, os.path.join(_EMU_DIR, "emu.cpp"), "-x", "c++", os.path.join(_CSRC, "b200_step.cu")]
        subprocess.check_call(cmd)
    return _cabi.bind(ctypes.CDLL(out))


class EmuSim:
    """numpy-buffer twin of the device-side state; the emulated kernels read/write it in place."""

    def __init__(self, spec, num_envs, height_samples=None, terrain_origins=None, with_env=True, specialized=True):
        self.lib = build_emu(with_env)
        self.specialized, self.last_preset = specialized, -1      # preset instantiation of the env kernel, like b200_create picks it
        self.spec, self.N = spec, num_envs
        self.model = spec.load_model()
        self.hs = None if height_samples is None else np.ascontiguousarray(height_samples, np.int16)
        self.origins = None if terrain_origins is None else np.ascontiguousarray(terrain_origins, np.float32)
        shape = self.hs.shape if self.hs is not None else (0, 0)
        self.tf, self.ti = _cabi.pack_task(spec, self.model, num_envs, shape)
        self.mi, self.mf = self.model.packed_ints(), self.model.packed_floats()
        self.buf = {k: np.zeros(s, d) for k, (s, d) in _cabi.buffer_shapes(spec, self.model, num_envs).items()}
        self.buf["base_quat_wxyz"][:, 0] = 1
        self.buf["added_mass"][:] = 1
        self.buf["kp_scale"][:] = 1
        self.buf["kd_scale"][:] = 1
        self.cbuf = _cabi.fill_buffers(lambda n: self.buf[n].ctypes.data)
        self.hist_count = 0          # observation frames appended so far (ring position of the frame stacks)
        self.step_counter = 0
        self.cmd_range_x = [float(spec.cmd_lin_vel_x[0]), float(spec.cmd_lin_vel_x[1])]
        self.beh_ranges = np.zeros((4, 2), np.float64)      # go2_wtw behaviour ranges (period, height, clearance, pitch)
        self.num_gaits = 1

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def _stack(self, name):
        w = self.spec.obs_widths(self.model)
        return (self.spec.frame_stack, w["obs"]) if name == "obs_history" else (self.spec.c_frame_stack, w["single_critic"])

    def _load_stack(self, name, window):
        K, W = self._stack(name)
        win = np.asarray(window, np.float32).reshape(self.N, K, W)
        M = K + 1
        ring = self.buf[name].reshape(self.N, 2 * M, W)
        for i in range(K):
            slot = (self.hist_count - K + i) % M
            ring[:, slot] = win[:, i]
            ring[:, slot + M] = win[:, i]

    def _window(self, name):
        K, W = self._stack(name)
        last = (self.hist_count - 1) % (K + 1)
        return self.buf[name][:, (last + 2) * W:(last + 2 + K) * W]

    def load_state(self, st):
        alias = {"q": "dof_pos", "qd": "dof_vel"}
        w = self.spec.obs_widths(self.model)
        for k, v in st.items():
            name = alias.get(k, k)
            if name in ("obs_hist", "obs_history") and w["hist"] and np.asarray(v).size == self.N * w["hist"]:
                self._load_stack("obs_history", v)
            elif name in ("critic_hist", "critic_obs") and w["critic"] and np.asarray(v).size == self.N * w["critic"]:
                self._load_stack("critic_obs", v)
            elif name in self.buf and np.asarray(v).size == self.buf[name].size:
                self.buf[name][...] = np.asarray(v).reshape(self.buf[name].shape)
        if "common_step_counter" in st:
            self.step_counter = int(st["common_step_counter"])
        if "cmd_range_x" in st:
            self.cmd_range_x = [float(x) for x in st["cmd_range_x"]]
        if "beh_ranges" in st:
            self.beh_ranges = np.asarray(st["beh_ranges"], np.float64).reshape(4, 2)
            self.num_gaits = int(st["num_gaits"])

    def dynamics_step(self, actions):
        a = np.ascontiguousarray(actions, np.float32)
        rows, cols = (self.hs.shape if self.hs is not None else (0, 0))
        self.buf["global_flags"][0] = 0          # the memset launch_dynamics (csrc/b200_step.cu) enqueues before the kernel
        self.lib.emu_dynamics_step(self._p(self.tf), self._p(self.ti), self._p(self.mi), self._p(self.mf), self._p(self.hs),
                                   ctypes.c_int(rows), ctypes.c_int(cols), ctypes.byref(self.cbuf), self._p(a))

    def env_post_step(self, phase_mask=63, force_reset=0, sit_pose=None):
        rows, cols = (self.hs.shape if self.hs is not None else (0, 0))
        lv, ty = (self.origins.shape[:2] if self.origins is not None else (0, 0))
        self.step_counter += 1
        lo, hi = self.cmd_range_x
        if sit_pose is None:
            from hcr_genesis_lr_cl_b200.host_rng import host_uniform
            from hcr_genesis_lr_cl_b200 import task_spec as T
            sit_pose = self.spec.sit_init_percent > 0 and host_uniform(self.spec.seed, self.step_counter, T.SITE_HOST, 0) < self.spec.sit_init_percent
        from hcr_genesis_lr_cl_b200.host_rng import host_uniform as _hu
        from hcr_genesis_lr_cl_b200 import task_spec as _T
        gaits = [min(int(_hu(self.spec.seed, self.step_counter, _T.SITE_HOST, 1 + w) * self.num_gaits), self.num_gaits - 1) for w in (0, 1)]
        beh = np.zeros(8, np.float32)
        beh[0::2] = self.beh_ranges[:, 0]
        beh[1::2] = self.beh_ranges[:, 1] - self.beh_ranges[:, 0]
        self.lib.emu_set_env_specialized(ctypes.c_int(int(self.specialized)))
        self.last_preset = self.lib.emu_env_post_step(self._p(self.tf), self._p(self.ti), self._p(self.hs), ctypes.c_int(rows), ctypes.c_int(cols),
                                   self._p(self.origins), ctypes.c_int(lv), ctypes.c_int(ty), ctypes.byref(self.cbuf),
                                   ctypes.c_longlong(self.step_counter), ctypes.c_float(lo), ctypes.c_float(np.float32(hi - lo)),
                                   ctypes.c_longlong(self.hist_count), ctypes.c_int(phase_mask), ctypes.c_int(force_reset), ctypes.c_int(int(bool(sit_pose))),
                                   self._p(beh), ctypes.c_int(gaits[0]), ctypes.c_int(gaits[1]))
        if phase_mask & 32:
            self.hist_count += 1

    @property
    def obs_history(self):
        return self._window("obs_history")

    @property
    def critic_obs(self):
        return self._window("critic_obs")


from oracle.physics import oracle_params, oracle_policy_step  # noqa: E402,F401  (historic home of these helpers)
