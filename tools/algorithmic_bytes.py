#!/usr/bin/env python
"""Print the algorithmic-bytes table for a task (SURVEY section 8d accounting; used by bench.py and DESIGN.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import accounting, task_spec as T  # noqa: E402

for task in (sys.argv[1:] or sorted(T.PRESETS)):
    spec = T.PRESETS[task]()
    model = spec.load_model()
    r, w = accounting.env_kernel_items(spec, model)
    print(f"== {task}: fused env kernel")
    for k, v in r.items():
        print(f"   read  {v:6d} B  {k}")
    for k, v in w.items():
        print(f"   write {v:6d} B  {k}")
    e, d = accounting.env_kernel_bytes(spec, model), accounting.dynamics_kernel_bytes(spec, model)
    print(f"   env kernel {e} B/env/policy-step; dynamics kernel {d} B; whole step {e + d} B ({(e + d) / spec.decimation:.0f} B per env-substep); "
          f"moving the kept frames of the stacks every step would add {accounting.shifted_stack_bytes(spec, model)} B")
