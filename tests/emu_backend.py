"""TEST-ONLY: the host layers above the C ABI (B200Simulator, FusedLeggedEnv, the plugin overlay) over
tests/warp_emu/libb200step_emu.so -- the product's own csrc/b200_step.cu compiled for the single-warp emulator -- with
torch CPU tensors standing in for device memory.  This is how plugin mode (the reference's task classes on top of the
backend) is exercised in a container without a GPU; the `-m gpu` twins of these tests run the real library."""
import contextlib
import ctypes
import os

import torch

from emu_util import build_emu_cabi
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
from hcr_genesis_lr_cl_b200.simulator import B200Simulator


class EmuB200Simulator(B200Simulator):
    def __init__(self, *a, **k):
        prev = os.environ.get("B200_DYN_ORDER")
        os.environ["B200_DYN_ORDER"] = "0"      # dynamics_order_kernel needs real shared memory between lanes
        try:
            super().__init__(*a, **k)
        finally:
            if prev is None:
                os.environ.pop("B200_DYN_ORDER", None)
            else:
                os.environ["B200_DYN_ORDER"] = prev

    def _load_library(self):
        return build_emu_cabi()

    def _check_device(self, sim_device):
        return torch.device("cpu")

    def _device_ctx(self):
        return contextlib.nullcontext()

    def _sync(self):
        pass

    def _stream(self):
        return ctypes.c_void_p(0)

    def _pinned(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    @staticmethod
    def _on_device(t):
        return not getattr(t, "_emu_host", False)      # every CPU tensor is "device" memory here unless a test marks it as host

    @staticmethod
    def _is_pinned(t):
        return True


class EmuFusedLeggedEnv(FusedLeggedEnv):
    simulator_class = EmuB200Simulator
