"""``FusedLeggedEnv`` -- the VecEnv the rsl_rl runners drive, with the whole ``env.step`` on two CUDA kernels.

Mirrors the call contract of ``LeggedRobot.step`` / ``LeggedRobotTS.step`` (legged_gym/envs/base/legged_robot.py:37-53,
legged_robot_ts.py:59-76): same return tuples, same attribute names the runners read (``num_envs``, ``num_obs``,
``num_privileged_obs``, ``num_actions``, ``max_episode_length``, ``episode_length_buf``, ``extras``, ``device``,
``get_observations``, ``reset``).  Per policy step it makes ONE C-ABI call, b200_env_step, which launches

    dynamics_step_kernel   (clip/shift actions, 4 x [PD torque + rigid-body substep])
    env_post_step_kernel   (fused post_physics_step; its last CTA finalises extras["episode"])

(plus the one-CTA dynamics_order_kernel on a side stream).  The frame stacks (`obs_history`, `critic_obs_buf`) are strided
[N, K * width] views of double-written rings in HBM (include/b200_step.h): same values and frame order as the tensors the
reference concatenates from its deques every step, row stride 2 K * width, nothing is moved between steps.  The view of a
step stays valid until K further steps have overwritten its slots; copy it (`RolloutStorage.add_transitions` does) to keep it.

and nothing else on the GPU; there is no host synchronisation inside ``step`` (SURVEY 8b "Threading").
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import task_spec as T
from .host_rng import host_uniform
from .simulator import B200Simulator


class FusedLeggedEnv:
    simulator_class = B200Simulator

    def __init__(self, spec: T.TaskSpec, num_envs: int, device: str = "cuda:0", terrain=None, env_offset: int = 0,
                 num_envs_global: Optional[int] = None, cfg=None, debug_cells: bool = False):
        self.spec = spec
        self.cfg = cfg
        self.device = device
        self.headless = True
        self.num_envs = int(num_envs)
        self.simulator = self.simulator_class(spec, None, device, True, num_envs=num_envs, terrain=terrain, env_offset=env_offset,
                                       num_envs_global=num_envs_global, debug_cells=debug_cells)
        sim = self.simulator
        self._b = sim._buf
        self.widths = spec.obs_widths(sim._model)
        self.stacked = spec.obs_kind in ("tron1_pf", "tron1_pf_ee", "go2_wtw", "go2_ee")   # obs / privileged obs are the frame stacks themselves
        self.estimator = spec.obs_kind in ("tron1_pf_ee", "go2_ee")             # LeggedRobotEE return tuples
        self.teacher_student = spec.obs_kind in ("go2_ts", "go2_cat", "go2_cts")  # LeggedRobotTS / LeggedRobotCTS return tuples
        self.num_obs = self.widths["hist"] if self.stacked else self.widths["obs"]
        self.num_estimator_features, self.num_estimator_labels = self.widths["hist"], self.widths["priv"]
        self.num_privileged_obs = self.widths["critic"] if self.stacked else (
            self.widths["priv"] if self.teacher_student else (self.widths["critic"] if spec.obs_kind == "go2_dreamwaq" else None))
        self.num_actions = spec.num_actions
        self.num_history_obs = self.widths["hist"]
        self.num_critic_obs = self.widths["critic"]
        self.dt = spec.dt
        self.max_episode_length_s = spec.episode_length_s
        self.max_episode_length = float(spec.max_episode_length)
        self.common_step_counter = 0
        self.command_ranges = dict(lin_vel_x=list(spec.cmd_lin_vel_x), lin_vel_y=list(spec.cmd_lin_vel_y),
                                   ang_vel_yaw=list(spec.cmd_ang_vel_yaw), heading=list(spec.cmd_heading))
        self.sum_names = spec.episode_sum_names()
        self.reward_scales = {n: float(spec.reward_scales[n] * spec.dt) for n in self.sum_names if n in spec.reward_scales}
        self.extras = {}
        self._pending_curriculum = None
        self.init_done = True
        # live views (same storage the kernels write)
        b = self._b
        self.rew_buf = b["rew_buf"]
        # The per-step rows a step hands out (obs_buf, privileged_obs_buf / estimator labels, next_state_buf) alternate between
        # two buffers: rsl_rl's algorithms keep a reference to the tensors they acted on and copy them into the rollout storage
        # only AFTER env.step (ppo.py:91-116), which works in the reference because its step builds fresh tensors.  The kernel
        # writes them through the rollout-target mechanism (b200_set_rollout_targets), so no copy is involved.
        self._pp = [{k: torch.zeros_like(b[k]) for k in ("obs_buf", "privileged_obs_buf", "next_state_buf")} for _ in range(2)]
        self._pp_i = 0
        self._targets_set = False
        if not self.stacked:
            self.obs_buf = b["obs_buf"]
            if spec.obs_kind != "go2_dreamwaq":                  # dreamwaq: the critic stack (property below)
                self.privileged_obs_buf = b["privileged_obs_buf"] if self.num_privileged_obs is not None else None
        self.reset_buf = b["reset_buf"].view(torch.bool)
        self.time_out_buf = b["time_out_buf"].view(torch.bool)
        self.commands, self.actions = b["commands"], b["actions"]
        self.last_actions, self.llast_actions = b["last_actions"], b["llast_actions"]
        self.feet_air_time, self.fail_buf = b["feet_air_time"], b["fail_buf"]
        self.episode_sums = {n: b["episode_sums"][:, i] for i, n in enumerate(self.sum_names)}
        self.cstr_prob = b["cstr_prob"]
        if spec.randomize_ctrl_delay:                         # legged_robot.py:405-409
            self.action_queue = b["action_queue"].view(self.num_envs, -1, spec.num_actions)
            self.action_delay = b["action_delay"]
        if spec.behavior_enabled:                             # go2_wtw.py:354-375 + 320-352: ranges start at mid / min
            def mid(r):
                return [(r[0] + r[1]) / 2] * 2
            self.gait_period_range, self.base_height_target_range = mid(spec.gait_period_range), mid(spec.base_height_target_range)
            self.foot_clearance_target_range, self.pitch_target_range = [spec.foot_clearance_target_range[0]] * 2, mid(spec.pitch_target_range)
            self.num_gaits, self.num_gait_max = 1, len(spec.gait_theta_lists)
            from ._cabi import H
            g = b["gait_state"]
            g.zero_()
            g[:, H["B200_GS_TH"]:H["B200_GS_TH"] + 4] = torch.tensor(spec.gait_theta_lists[0], device=g.device)
            g[:, H["B200_GS_PER"]], g[:, H["B200_GS_BH"]] = self.gait_period_range[0], self.base_height_target_range[0]
            g[:, H["B200_GS_FC"]], g[:, H["B200_GS_PT"]] = self.foot_clearance_target_range[0], self.pitch_target_range[0]
        self._extras_ring = self._build_extras_ring()
        # go2_cat as shipped couples ALL envs through its stand-still constraint (go2_cat.py:177-178): under env
        # sharding the "any env is moving fast" flag is a scalar of the whole job.  cat_global_allreduce=True keeps it one
        # (a 4-byte NCCL MAX between the dynamics and the env kernel: the step then takes two C-ABI calls and ~one collective
        # latency more); the default leaves the flag rank-local, i.e. results then depend on the sharding for this one task --
        # which the per-env switch (cat_stand_still_global=False) avoids altogether.
        import os
        self.cat_global_allreduce = bool(spec.cat_enabled and spec.cat_stand_still_global and os.environ.get("B200_CAT_ALLREDUCE", "0") == "1")

    # runners assign a fresh tensor to env.episode_length_buf (on_policy_runner.py:169); keep the bound storage
    @property
    def episode_length_buf(self):
        return self._b["episode_length"]

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        self._b["episode_length"].copy_(value.to(self._b["episode_length"].dtype))

    # tron1_pf: LeggedRobot.step returns the stacks as obs_buf / privileged_obs_buf (tron1_pf.py:57-70)
    def __getattr__(self, name):
        if name == "obs_buf" and self.__dict__.get("stacked"):
            return self.obs_history
        if name == "privileged_obs_buf" and (self.__dict__.get("stacked") or self.__dict__["spec"].obs_kind == "go2_dreamwaq"):
            return self.critic_obs_buf
        # sizes the reference's runners read off the env (ts_runner.py:56-67, ee_runner.py, dreamwaq_runner.py, cts_runner.py:
        # num_latent_dims, num_explicit_dims, num_decoder_output, num_teacher, ...) live in cfg.env, like in the task classes
        # (legged_robot_ts.py:92, legged_robot_dreamwaq.py:99-101, legged_robot_ee.py:88-89)
        cfg = self.__dict__.get("cfg")
        if name.startswith("num_") and cfg is not None and hasattr(getattr(cfg, "env", None), name):
            return getattr(cfg.env, name)
        raise AttributeError(name)

    @property
    def obs_history(self):
        return self.simulator.obs_history

    @property
    def critic_obs_buf(self):
        return self.simulator.critic_obs

    # ------------------------------------------------------------------ VecEnv API
    def step(self, actions: torch.Tensor):
        """LeggedRobot.step: one C-ABI call (b200_env_step) = _pre_sim_step + simulator.step + post_physics_step:
        dynamics kernel -> env kernel, with the frame-stack shift on a side stream under the dynamics kernel."""
        if self.cat_global_allreduce and self._world_size() > 1:
            return self._step_with_global_flag(actions)
        self.common_step_counter += 1
        self._apply_pending_curriculum()
        self._set_step_flags()
        self._set_output_targets()
        self.simulator.fused_env_step(actions, self.common_step_counter, self.command_ranges["lin_vel_x"])
        self._fill_extras()
        return self._returns()

    def set_rollout_targets(self, targets) -> None:
        """Extra destinations of the NEXT step's results (`_cabi.B200RolloutTargets`, used by rollout.FusedRolloutCollector);
        rows it does not redirect still alternate between the env's two output buffers."""
        import ctypes
        pp = self._pp[self._pp_i ^ 1]
        if not targets.obs:
            targets.obs = pp["obs_buf"].data_ptr()
        if not targets.privileged_obs:
            targets.privileged_obs = pp["privileged_obs_buf"].data_ptr()
        if not targets.next_state and self.spec.obs_kind == "go2_dreamwaq":
            targets.next_state = pp["next_state_buf"].data_ptr()
        sim = self.simulator
        sim._ck(sim._lib.b200_set_rollout_targets(sim._handle, ctypes.byref(targets)))
        self._targets_set = True

    def _set_output_targets(self) -> None:
        """Point this step's per-step rows at the buffer the PREVIOUS step did not return."""
        from ._cabi import B200RolloutTargets
        if not self._targets_set:
            self.set_rollout_targets(B200RolloutTargets())
        self._targets_set = False
        self._pp_i ^= 1
        pp = self._pp[self._pp_i]
        if not self.stacked:
            self.obs_buf = pp["obs_buf"]
            if self.spec.obs_kind != "go2_dreamwaq" and self.num_privileged_obs is not None:
                self.privileged_obs_buf = pp["privileged_obs_buf"]

    @staticmethod
    def _world_size() -> int:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _step_with_global_flag(self, actions: torch.Tensor):
        """go2_cat under env sharding with the job-wide stand-still flag (see __init__): dynamics kernel, 4-byte all-reduce
        (MAX) of global_flags[0], env kernel."""
        import torch.distributed as dist
        sim = self.simulator
        sim.step(actions)
        dist.all_reduce(self._b["global_flags"][:1], op=dist.ReduceOp.MAX)
        self.common_step_counter += 1
        self._apply_pending_curriculum()
        self._set_step_flags()
        self._set_output_targets()
        sim.fused_post_step(self.common_step_counter, self.command_ranges["lin_vel_x"])
        self._fill_extras()
        return self._returns()

    def step_two_kernels(self, actions: torch.Tensor):
        """The same step call by call (b200_dynamics_step, b200_env_post_step): what the plugin path
        does between `Simulator.step` and `post_physics_step`; kept for per-kernel timing and as a cross-check."""
        sim = self.simulator
        sim.step(actions)                                      # _pre_sim_step + simulator.step
        self.common_step_counter += 1
        self._apply_pending_curriculum()
        self._set_step_flags()
        self._set_output_targets()
        sim.fused_post_step(self.common_step_counter, self.command_ranges["lin_vel_x"])
        self._fill_extras()
        return self._returns()

    def step_host(self, actions_pinned: torch.Tensor, rew_out: Optional[torch.Tensor] = None,
                  reset_out: Optional[torch.Tensor] = None, time_out_out: Optional[torch.Tensor] = None):
        """`step` for a rollout loop that keeps actions / rewards / dones in pinned host memory: the host->device copy of
        the actions, the three kernels and the device->host copies of rew / reset / time_out are enqueued by ONE C-ABI
        call (b200_env_step) on the current stream.  Nothing is synchronised here; wait on the stream (or an event)
        before reading the host buffers.  Returns the same tuple as `step` (device views)."""
        self.common_step_counter += 1
        self._apply_pending_curriculum()
        self._set_step_flags()
        self._set_output_targets()
        self.simulator.fused_env_step(actions_pinned, self.common_step_counter, self.command_ranges["lin_vel_x"],
                                      rew_out, reset_out, time_out_out)
        self._fill_extras()
        return self._returns()

    # ------------------------------------------------------------------ device-stepped mode (CUDA-graph capturable)
    def sync_step_state(self) -> None:
        """Host -> device copy of the step's host-owned scalars (step counter, command range, behaviour ranges, number of gaits):
        call before the first `step_device` and whenever the host changed one of them (curricula)."""
        beh = [self.gait_period_range, self.base_height_target_range, self.foot_clearance_target_range, self.pitch_target_range] \
            if self.spec.behavior_enabled else None
        self.simulator.write_step_state(self.common_step_counter, self.command_ranges["lin_vel_x"], beh,
                                        self.num_gaits if self.spec.behavior_enabled else 1)
        self._curriculum_done_to = self.common_step_counter      # curricula up to here were the host-stepped path's business

    def step_device(self, actions: torch.Tensor):
        """`step` with every per-step scalar taken from device memory (b200_env_step_device): the launches take no value from
        the host that changes between steps, so T such steps -- with the policy's kernels in between -- can be captured into
        ONE CUDA graph and replayed (rollout.GraphedRolloutCollector).  The host-drawn scalars of `_set_step_flags` are drawn
        on the device from the same Philox stream; the command / behaviour curricula are NOT evaluated here (the caller
        runs `device_steps_done` at its rollout boundaries)."""
        if self.cat_global_allreduce and self._world_size() > 1:
            raise RuntimeError("go2_cat with the job-wide stand-still flag needs a collective inside the step: host-stepped only")
        self.common_step_counter += 1
        self._set_output_targets()
        self.simulator.fused_env_step_device(actions, self.spec.sit_init_percent)
        self._fill_extras(curriculum=False)
        return self._returns()

    def device_steps_done(self, n_replayed: int = 0) -> None:
        """Book-keeping after device-stepped steps: `n_replayed` steps ran on the device WITHOUT their python side (a graph
        replay) -> move the host counters, the output ping-pong and the frame-stack position along; then apply the command /
        behaviour curricula for every multiple of max_episode_length the window crossed (legged_robot.py:110-111, read from
        the statistics ring: late by at most the window) and push changed ranges back to the device."""
        if n_replayed:
            self.common_step_counter += n_replayed
            self.simulator._hist_count += n_replayed
            if n_replayed % 2:
                self._pp_i ^= 1
            self._fill_extras(curriculum=False)
        last = self.common_step_counter
        L = int(self.max_episode_length)
        done_to = getattr(self, "_curriculum_done_to", last)
        if not (self.spec.cmd_curriculum and "tracking_lin_vel" in self.sum_names):
            self._curriculum_done_to = last
            return
        changed = False
        first = max(done_to + 1, last - self.stats_ring + 1, 1)      # older slots of the ring have been overwritten
        for k in range((first + L - 1) // L * L, last + 1, L):
            changed |= self._curriculum_from_ring(k)
        self._curriculum_done_to = last
        if changed:
            self.sync_step_state()

    def _curriculum_from_ring(self, step: int) -> bool:
        """Command / behaviour curriculum of policy step `step` from its slot of the statistics ring (means over the envs that
        reset in that step, and their count)."""
        from ._cabi import STATS_RING, H
        n = len(self.sum_names)
        base, w = 2 * max(n, 1) + 4, n + H["B200_STATS_EXTRA"]
        row = self._b["stats"][base + (step % STATS_RING) * w: base + (step % STATS_RING + 1) * w].clone()
        import torch.distributed as dist
        cnt = row[n + 3].clone()
        sums = row[:n] * cnt                    # back to sums so that an env-sharded job can add its ranks up
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            pack = torch.cat([sums, cnt.view(1)])
            dist.all_reduce(pack, op=dist.ReduceOp.SUM)
            sums, cnt = pack[:n], pack[n]
        if float(cnt) <= 0:
            return False
        # ring entries are sum / count / episode_length_s; the curricula compare sum / count / max_episode_length
        means = sums / cnt * (self.max_episode_length_s / float(self.max_episode_length))
        before = (list(self.command_ranges["lin_vel_x"]), self.num_gaits,
                  [list(r) for r in (self.gait_period_range, self.base_height_target_range, self.foot_clearance_target_range, self.pitch_target_range)]
                  if self.spec.behavior_enabled else None)
        self._pending_curriculum = (means[self.sum_names.index("tracking_lin_vel")], cnt, means if self.spec.behavior_enabled else None)
        self._apply_pending_curriculum()
        after = (list(self.command_ranges["lin_vel_x"]), self.num_gaits,
                 [list(r) for r in (self.gait_period_range, self.base_height_target_range, self.foot_clearance_target_range, self.pitch_target_range)]
                 if self.spec.behavior_enabled else None)
        return before != after

    def _set_step_flags(self):
        """Host scalars of the step: the sit-pose coin (one per reset batch, tron1_pf_ee.py:204-210)."""
        if self.spec.sit_init_percent > 0:
            coin = host_uniform(self.spec.seed, self.common_step_counter, T.SITE_HOST, 0)
            self.simulator.set_step_flags(coin < self.spec.sit_init_percent)
        if self.spec.behavior_enabled:
            # ONE gait index per resampling call, shared by all its envs (go2_wtw.py:205-206): host draws
            # 1 (callback) and 2 (reset) of the step
            gaits = [min(int(host_uniform(self.spec.seed, self.common_step_counter, T.SITE_HOST, 1 + w) * self.num_gaits),
                         self.num_gaits - 1) for w in (0, 1)]
            self.simulator.set_behavior([self.gait_period_range, self.base_height_target_range, self.foot_clearance_target_range,
                                         self.pitch_target_range], gaits[0], gaits[1])

    @property
    def estimator_features_buf(self):
        return self.obs_history

    @property
    def estimator_labels_buf(self):
        return self._pp[self._pp_i]["privileged_obs_buf"]

    @property
    def explicit_labels_buf(self):
        return self._pp[self._pp_i]["privileged_obs_buf"]

    @property
    def next_state_buf(self):
        return self._pp[self._pp_i]["next_state_buf"]

    def _returns(self):
        if self.estimator:                               # LeggedRobotEE.step, legged_robot_ee.py:56-73
            return (self.obs_history, self.estimator_labels_buf, self.critic_obs_buf, self.rew_buf, self.reset_buf, self.extras)
        if self.spec.obs_kind == "go2_dreamwaq":         # LeggedRobotDreamwaq.step, legged_robot_dreamwaq.py:63-79
            return (self.obs_buf, self.critic_obs_buf, self.obs_history, self.explicit_labels_buf, self.next_state_buf,
                    self.rew_buf, self.reset_buf, self.extras)
        if self.teacher_student:
            return (self.obs_buf, self.privileged_obs_buf, self.obs_history, self.critic_obs_buf, self.rew_buf,
                    self.reset_buf, self.extras)
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def reset(self):
        """BaseTask.reset (base_task.py:60-64): reset_idx(all) then one zero-action step."""
        self._set_step_flags()
        self.simulator.fused_reset_all(self.common_step_counter, self.command_ranges["lin_vel_x"])
        out = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        if self.estimator:
            return out[:3]
        if self.spec.obs_kind == "go2_dreamwaq":
            return out[:5]
        return out[:4] if self.teacher_student else out[:2]

    def get_observations(self):
        if self.estimator:
            return self.obs_history, self.estimator_labels_buf, self.critic_obs_buf
        if self.spec.obs_kind == "go2_dreamwaq":
            return self.obs_buf, self.critic_obs_buf, self.obs_history, self.explicit_labels_buf, self.next_state_buf
        if self.teacher_student:
            return self.obs_buf, self.privileged_obs_buf, self.obs_history, self.critic_obs_buf
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    # ------------------------------------------------------------------ extras / curricula (device side, no sync)
    def _build_extras_ring(self):
        """Per-slot dicts of 0-dim views into the statistics ring the finalize kernel fills (no launch per step)."""
        from ._cabi import STATS_RING, H
        n = len(self.sum_names)
        stats = self._b["stats"]
        base = 2 * max(n, 1) + 4
        ring, w = [], n + H["B200_STATS_EXTRA"]
        for slot in range(STATS_RING):
            row = stats[base + slot * w: base + (slot + 1) * w]
            ep = {"rew_" + name: row[i] for i, name in enumerate(self.sum_names)}
            if self.spec.terrain_curriculum:
                ep["terrain_level"] = row[n]
            if self.spec.cat_enabled:
                ep["cstr_probs"] = row[n + 1]               # go2_cat.py:101-104
            if self.spec.num_teacher > 0 and self.spec.terrain_curriculum:    # go2_cts.py:93-99
                ep["teacher_terrain_level"], ep["student_terrain_level"] = row[n + 1], row[n + 2]
            ring.append(ep)
        return ring

    #: a step's extras["episode"] values are views into slot (step % stats_ring) of a device ring: a runner that keeps them
    #: for more than `stats_ring` steps before reading (num_steps_per_env > stats_ring) must clone them at append time
    stats_ring = property(lambda self: int(self.simulator._lib.b200_stats_ring()))

    def _fill_extras(self, curriculum: bool = True):
        """extras["episode"]["rew_*"] = mean over resetting envs of episode_sums / episode_length_s
        (legged_robot.py:127-141): computed on the device by the env kernel's last CTA into slot step % 32."""
        from ._cabi import STATS_RING
        n = len(self.sum_names)
        stats = self._b["stats"]
        ep = dict(self._extras_ring[self.common_step_counter % STATS_RING])
        if self.spec.cmd_curriculum:
            ep["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.spec.behavior_enabled:                      # go2_wtw.py:161-165
            ep["gait_period_max"], ep["base_height_target_max"] = self.gait_period_range[1], self.base_height_target_range[1]
            ep["foot_clearance_target_max"], ep["pitch_target_max"] = self.foot_clearance_target_range[1], self.pitch_target_range[1]
            ep["num_gaits"] = self.num_gaits
        self.extras["episode"] = ep
        if self.spec.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        # command curriculum (legged_robot.py:110-111,336-348): evaluated on the reductions of this step and applied
        # before the next one (one-step delay, DESIGN.md "deviations"); the only host read, once per max_episode_length.
        if curriculum and self.spec.cmd_curriculum and self.common_step_counter % int(self.max_episode_length) == 0 \
                and "tracking_lin_vel" in self.sum_names:
            i = self.sum_names.index("tracking_lin_vel")
            tot = stats[:n + 1].clone()                       # episode sums of the envs that reset this step + their count
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # env-sharded job: the curricula are scalars of the WHOLE job; 4 (n + 1) bytes once per max_episode_length steps
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            cnt = tot[n].clamp(min=1.0)
            self._pending_curriculum = (tot[i] / cnt / self.max_episode_length, tot[n].clone(),
                                        (tot[:n] / cnt / self.max_episode_length) if self.spec.behavior_enabled else None)

    def _apply_pending_curriculum(self):
        if self._pending_curriculum is None:
            return
        mean_track, cnt, means = self._pending_curriculum
        self._pending_curriculum = None
        if float(cnt) <= 0:
            return
        if float(mean_track) > self.spec.curriculum_threshold * self.reward_scales["tracking_lin_vel"]:
            r = self.command_ranges["lin_vel_x"]
            r[0] = float(np.clip(r[0] - 0.5, -self.spec.max_curriculum, 0.0))
            r[1] = float(np.clip(r[1] + 0.5, 0.0, self.spec.max_curriculum))
        if means is not None:                                 # go2_wtw.py:220-249
            self._update_behavior_param_curriculum(dict(zip(self.sum_names, means.tolist())))

    def _update_behavior_param_curriculum(self, mean_sums):
        """Widen the behaviour ranges from the mean per-step episode sums of the resetting envs (go2_wtw.py:220-249)."""
        s, sc = self.spec, self.reward_scales

        def widen(rng, step, lim):
            rng[0], rng[1] = max(rng[0] - step, lim[0]), min(rng[1] + step, lim[1])

        if mean_sums.get("quad_periodic_gait", -np.inf) > 0.8 * sc.get("quad_periodic_gait", np.inf):
            widen(self.gait_period_range, 0.05, s.gait_period_range)
            self.num_gaits = min(self.num_gaits + 1, self.num_gait_max)
        if mean_sums.get("tracking_base_height", -np.inf) > 0.9 * sc.get("tracking_base_height", np.inf):
            widen(self.base_height_target_range, 0.02, s.base_height_target_range)
        if mean_sums.get("tracking_foot_clearance", -np.inf) > 0.8 * sc.get("tracking_foot_clearance", np.inf):
            widen(self.foot_clearance_target_range, 0.01, s.foot_clearance_target_range)
        if mean_sums.get("tracking_orientation", -np.inf) > 0.9 * sc.get("tracking_orientation", np.inf):
            widen(self.pitch_target_range, 0.05, s.pitch_target_range)


def make_env(task: str, num_envs: int, device: str = "cuda:0", terrain=None, env_offset: int = 0,
             num_envs_global: Optional[int] = None, **spec_overrides) -> FusedLeggedEnv:
    """Build a fused env from a built-in task preset (`task_spec.PRESETS`)."""
    if task not in T.PRESETS:
        raise ValueError(f"no fused descriptor for task {task!r} (available: {sorted(T.PRESETS)})")
    spec = T.PRESETS[task](**spec_overrides)
    return FusedLeggedEnv(spec, num_envs, device, terrain=terrain, env_offset=env_offset, num_envs_global=num_envs_global)
