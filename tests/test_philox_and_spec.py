"""CPU tests: Philox known answers, task descriptors, header parsing, accounting."""
import dataclasses

import numpy as np
import pytest

from hcr_genesis_lr_cl_b200 import _cabi, accounting, task_spec as T
from oracle import philox


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(x) for x in philox.philox4x32(*ctr, *key)) == out


def test_uniform_range_and_determinism():
    u = philox.uniform(1, 7, np.arange(1000)[:, None], 3, np.arange(12)[None, :])
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(u, philox.uniform(1, 7, np.arange(1000)[:, None], 3, np.arange(12)[None, :]))
    assert not np.array_equal(u, philox.uniform(1, 8, np.arange(1000)[:, None], 3, np.arange(12)[None, :]))


def test_header_enums_and_struct():
    H = _cabi.H
    assert H["TF_COUNT"] > H["TF_POINTS_Y"] and H["TI_COUNT"] > H["TI_CS_LINKS"]
    assert H["RW_COUNT"] == len(T.REWARD_TERMS) == H["B200_MAX_REWARDS"]
    for i, name in enumerate(T.REWARD_TERMS):
        assert H["RW_" + name.upper()] == i
    assert [H[s] for s in ("SITE_CMD_RESAMPLE", "SITE_PUSH", "SITE_LEVEL", "SITE_CMD_RESET", "SITE_DOF", "SITE_ROOT")] == \
        [T.SITE_CMD_RESAMPLE, T.SITE_PUSH, T.SITE_LEVEL, T.SITE_CMD_RESET, T.SITE_DOF, T.SITE_ROOT]
    assert H["SITE_OBS_NOISE"] == T.SITE_OBS_NOISE
    names = [n for n, _ in _cabi.BUFFER_FIELDS]
    assert names[0] == "base_pos" and names[-1] == "nonfinite" and "dyn_order" in names and "stats" in names and "obs_history" in names and "critic_obs" in names and len(set(names)) == len(names)


@pytest.mark.parametrize("task", ["go2", "go2_ts"])
def test_pack_task_and_buffers(task):
    spec = T.PRESETS[task]()
    model = spec.load_model()
    f, i = _cabi.pack_task(spec, model, 64, (1200, 1200) if spec.heightfield else (0, 0))
    H = _cabi.H
    assert f.shape == (H["TF_COUNT"],) and i.shape == (H["TI_COUNT"],)
    assert i[H["TI_A"]] == 12 and i[H["TI_C"]] == 4 and i[H["TI_L"]] == 17 and i[H["TI_F"]] == 4
    assert list(i[H["TI_FEET_LINKS"]:H["TI_FEET_LINKS"] + 4]) == [4, 8, 12, 16]          # FL, FR, RL, RR
    ids = list(i[H["TI_REWARD_IDS"]:H["TI_REWARD_IDS"] + i[H["TI_N_REWARDS"]]])
    assert ids == sorted(ids) and [T.REWARD_TERMS[k] for k in ids] == spec.active_rewards()
    assert np.isclose(f[H["TF_FAIL_LIMIT"]], 5.0) and i[H["TI_MAX_EPISODE_LENGTH"]] == 1000
    shapes = _cabi.buffer_shapes(spec, model, 64)
    w = spec.obs_widths(model)
    assert shapes["obs_buf"][0] == (64, 45)
    if task == "go2_ts":
        assert w == dict(obs=45, priv=99, single_critic=177, hist=900, critic=885)        # SURVEY Appendix A (as run)
        assert shapes["obs_history"][0] == (64, 2 * 21 * 45) and shapes["critic_obs"][0] == (64, 2 * 6 * 177)      # double-written rings, period K + 1
        assert i[H["TI_RESAMPLE_INTERVAL"]] == 500 and i[H["TI_PUSH_INTERVAL"]] == 500
    else:
        assert i[H["TI_PUSH_INTERVAL"]] == 750 and list(i[H["TI_TERM_LINKS"]:H["TI_TERM_LINKS"] + 1]) == [0]


def test_documented_widths_with_12_contact_links():
    """with contact_state_link_names = thigh/calf/foot the widths are the documented 94 / 860."""
    spec = T.go2_ts_spec(contact_state_link_names=["thigh", "calf", "foot"])
    w = spec.obs_widths(spec.load_model())
    assert w["priv"] == 94 and w["critic"] == 860


def test_presets_match_reference_configs():
    from oracle.ref_harness import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    reg = import_reference()
    for task in ("go2", "go2_ts", "go2_cat", "tron1_pf"):
        cfg, _ = reg.get_cfgs(task)
        a, b = dataclasses.asdict(T.TaskSpec.from_reference_cfg(cfg, task)), dataclasses.asdict(T.PRESETS[task]())
        ra, rb = a.pop("reward_scales"), b.pop("reward_scales")
        assert a == b, {k: (a[k], b[k]) for k in a if a[k] != b[k]}
        assert {k: v for k, v in ra.items() if v != 0} == {k: v for k, v in rb.items() if v != 0}


def test_algorithmic_bytes_close_to_survey():
    """SURVEY 8d's hand count is 16.7 KB per env and policy step with the frame stacks re-written every step and names the
    ring variant; the accounting here is that variant (frames written twice, nothing moved), and adding back what moving
    the kept frames would cost must land on SURVEY's figure."""
    spec = T.go2_ts_spec()
    model = spec.load_model()
    whole = accounting.step_bytes(spec, model)
    assert 5_000 < whole < 7_000
    assert 15_000 < whole + accounting.shifted_stack_bytes(spec, model) - 4 * (45 + 177) < 19_000
    assert accounting.env_path_bytes(spec, model) > 3 * accounting.dynamics_kernel_bytes(spec, model)
    assert accounting.env_path_bytes(spec, model) == accounting.env_kernel_bytes(spec, model)


def test_robot_model_invariants():
    for robot, names, mass in (("go2", T.GO2_DOF_NAMES, 16.087), ("tron1_pf", None, 18.593)):
        from hcr_genesis_lr_cl_b200.robot_model import TRON1_PF_DOF_NAMES, load_robot_model
        m = load_robot_model(robot, names or TRON1_PF_DOF_NAMES)
        assert abs(m.body[:, 9].sum() - mass) < 2e-3                      # SURVEY Appendix B total mass
        assert m.nspheres <= 64 and m.chain_len == 3
        assert np.allclose(np.linalg.norm(m.body[1:, 3:6], axis=1), 1.0)
    with pytest.raises(ValueError):
        load_robot_model("go2", list(reversed(T.GO2_DOF_NAMES)))


def test_generated_env_presets_are_current():
    """csrc/env_presets.inc (compile-time int descriptors of the built-in presets) matches what pack_task produces now."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sp = importlib.util.spec_from_file_location("gen_env_presets", os.path.join(root, "tools", "gen_env_presets.py"))
    gen = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(gen)
    assert open(gen.path()).read() == gen.render(), "run tools/gen_env_presets.py"


def test_depth_sensor_configs_are_refused_loudly():
    """A cfg that asks for depth images (sensor.add_depth, genesis_simulator.py:135-138) must not silently get a backend without
    a camera."""
    from types import SimpleNamespace
    import pytest
    from hcr_genesis_lr_cl_b200.task_spec import TaskSpec
    cfg = SimpleNamespace(sensor=SimpleNamespace(add_depth=True))
    with pytest.raises(ValueError, match="depth"):
        TaskSpec.from_reference_cfg(cfg, "go2_ts")
