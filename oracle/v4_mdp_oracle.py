"""TEST INFRASTRUCTURE (oracle) -- numpy float32 restatement of the MDP half of the reference task ``Zbot6SEnvV4``
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py``; each function cites the
lines it follows), including the EventManager's "reset" / "interval" modes it relies on ([IL-upstream]).
PINNED by ``tests/golden/v4_*.npz`` (outputs of the reference's own unmodified code, ``tests/golden/make_golden.py``).
Only tests/, smoke() and bench.py's CPU legs may import this.

Random numbers: ``rnd`` (N,10) uniforms per step, one column per draw the reference makes (V4RandSlot in
``csrc/zbot_core.h``); an env consumes a column only when the corresponding event fires for it.
"""
from __future__ import annotations

import numpy as np

from .il_semantics import quat_apply

F = np.float32
#: …env_v4.py:522-556 (dict order = evaluation order); bare weights, multiplied by step_dt at evaluation (:857)
REWARD_SCALES_V4 = {
    "track_lin_vel_x": 1.0, "track_heading_yaw": 1.0, "lin_vel_y": -1.0, "action_rate": -0.1, "torques": -2e-4,
    "joint_vel": -0.001, "joint_acc": -2.5e-7, "feet_downward": -1.0, "feet_forward": -0.5, "step_length": 5.0,
    "feet_air_time_biped": 1.0, "airtime_variance": -5.0, "feet_slide": -1.0, "feet_harmony": 0.0, "feet_close": -10.0,
}
#: EventCfg (…env_v4.py:331-418)
EVENTS_V4 = {"pose_lo": (-0.5, -0.5, -3.14), "pose_hi": (0.5, 0.5, 3.14), "velocity_range": (0.3, 0.3),
             "yaw_range": (-0.1, 0.1), "dual_sign": True, "offset": 0.0, "prob_pos": 1.0, "interval_range_s": (3.0, 6.0)}


def wrap_to_pi(a):
    """isaaclab.utils.math.wrap_to_pi in float32."""
    a = np.asarray(a, F)
    w = np.mod(a + F(np.pi), F(2 * np.pi)).astype(F)
    return np.where((w == 0) & (a > 0), F(np.pi), w - F(np.pi)).astype(F)


class V4MdpOracle:
    def __init__(self, num_envs, index_sets, default_joint_pos, default_root_pos, fk, reward_scales=None, events=None,
                 step_dt=0.02, max_episode_length=1000, termination_height=0.20):
        n = self.n = int(num_envs)
        self.feet_ids = list(index_sets["feet_ids"])
        self.undesired_ids = list(index_sets["undesired_ids"])
        self.base_body_idx = int(index_sets["base_body_idx"][0])
        self.feet_body_idx = list(index_sets["feet_body_idx"])
        self.default_joint_pos = np.asarray(default_joint_pos, F)
        self.default_root_pos = np.asarray(default_root_pos, F)      # per env, WORLD frame (default + origin)
        self.fk = fk                                                  # (root_pos, root_quat) -> link pos (12,3), quat (12,4)
        self.step_dt, self.max_episode_length, self.termination_height = step_dt, max_episode_length, termination_height
        self.reward_scales = dict(REWARD_SCALES_V4 if reward_scales is None else reward_scales)
        self.events = dict(EVENTS_V4 if events is None else events)
        self.episode_sums = {k: np.zeros(n, F) for k in self.reward_scales}
        self.actions = np.zeros((n, 6), F)
        self.prev_actions = np.zeros((n, 6), F)
        self.p_delta = np.zeros((n, 6), F)
        self.commands = np.zeros((n, 2), F)
        self.current_yaw = np.zeros(n, F)
        self.target_heading_yaw = np.zeros(n, F)
        self.interval_time_left = np.zeros(n, F)
        self.feet_contact_forces_last = np.full((n, 2), 15.0, F)      # :637
        self.feet_down_pos_last = np.zeros((n, 2, 3), F)
        self.feet_step_length = np.zeros((n, 2), F)
        self.episode_length_buf = np.zeros(n, np.int64)
        self.reset_terminated = np.zeros(n, bool)
        self.reset_time_outs = np.zeros(n, bool)
        self.S = None
        self.log = None

    def attach(self, S):
        self.S = {k: np.array(v, copy=True) for k, v in S.items()}

    def pre_physics_step(self, actions):                                                 # :776-785
        self.actions = np.tanh(np.asarray(actions, F)).astype(F)
        self.p_delta = self.p_delta + (F(np.pi) * self.actions * F(1.0) * F(self.step_dt))
        self.p_delta = np.clip(self.p_delta, F(-np.pi), F(np.pi)).astype(F)
        self.processed_actions = self.p_delta + self.default_joint_pos

    def compute_intermediate_values(self):                                               # :792-826
        S, n = self.S, self.n
        self.base_pos_w = S["body_link_pos_w"][:, self.base_body_idx].copy()
        self.base_quat_w = S["body_link_quat_w"][:, self.base_body_idx].copy()
        self.feet_quat_w = S["body_link_quat_w"][:, self.feet_body_idx].copy()
        self.feet_pos_w = S["body_link_pos_w"][:, self.feet_body_idx].copy()
        self.base_shoulder_w = quat_apply(self.base_quat_w, np.tile(np.array([0, 0, 1], F), (n, 1)))
        self.base_dir_forward_w = np.cross(np.tile(np.array([0, 0, -1], F), (n, 1)), self.base_shoulder_w).astype(F)
        self.current_yaw = np.arctan2(self.base_dir_forward_w[:, 1], self.base_dir_forward_w[:, 0]).astype(F)
        diff = (self.target_heading_yaw - self.current_yaw).astype(F)
        self.heading_err = np.arctan2(np.sin(diff), np.cos(diff)).astype(F)
        self.base_lin_vel_w = S["body_link_lin_vel_w"][:, self.base_body_idx].copy()
        self.base_lin_vel_forward_w = np.sum(self.base_lin_vel_w * self.base_dir_forward_w, axis=-1, dtype=F)
        h = S["net_forces_w_history"][:, :, self.feet_ids, 2]
        self.feet_contact_forces = (((h[:, 0] + h[:, 1]) + h[:, 2]) / F(h.shape[1])).astype(F)

    def get_observations(self):                                                          # :828-851
        S = self.S
        self.prev_actions = self.actions.copy()
        self.base_quat_w = S["body_link_quat_w"][:, self.base_body_idx].copy()
        diff = (self.target_heading_yaw - self.current_yaw).astype(F)
        self.heading_err = np.arctan2(np.sin(diff), np.cos(diff)).astype(F)
        return np.concatenate([self.base_quat_w, S["joint_pos"] - self.default_joint_pos, S["joint_vel"], self.actions,
                               self.commands[:, 0:1], self.heading_err[:, None]], axis=-1).astype(F)

    def get_dones(self):                                                                 # :868-886
        self.compute_intermediate_values()
        time_out = self.episode_length_buf >= self.max_episode_length - 1
        hist = self.S["net_forces_w_history"][:, :, self.undesired_ids]
        norms = np.sqrt(hist[..., 0] * hist[..., 0] + hist[..., 1] * hist[..., 1] + hist[..., 2] * hist[..., 2])
        died = np.any(norms.max(axis=1) > F(0.5), axis=1)
        died |= self.base_pos_w[:, 2] < F(self.termination_height)
        return died, time_out

    # ---- reward terms (:1013-1199) ----
    def _reward_track_lin_vel_x(self):
        e = np.square(self.commands[:, 0] - self.base_lin_vel_forward_w)
        return np.exp(-e / F(0.25)).astype(F)

    def _reward_track_heading_yaw(self):
        return np.exp(-np.square(self.heading_err) / F(0.25)).astype(F)

    def _reward_lin_vel_x(self):
        return np.square(self.base_lin_vel_forward_w)

    def _reward_lin_vel_y(self):
        return np.square(np.sum(self.base_lin_vel_w * self.base_shoulder_w, axis=-1, dtype=F))

    def _reward_feet_forward(self):
        fx = quat_apply(self.feet_quat_w, np.tile(np.array([1, 0, 0], F), (self.n, 2, 1)))
        d = fx - self.base_dir_forward_w[:, None, :]
        return np.sum(np.sqrt(np.sum(d * d, axis=-1, dtype=F)), axis=-1, dtype=F)

    def _reward_feet_downward(self):
        fz = quat_apply(self.feet_quat_w, np.tile(np.array([[0, 0, 1], [0, 0, -1]], F), (self.n, 1, 1)))
        d = fz - np.tile(np.array([0, 0, 1], F), (self.n, 2, 1))
        return np.sum(np.sqrt(np.sum(d * d, axis=-1, dtype=F)), axis=-1, dtype=F)

    def _reward_step_length(self):
        F_, Fl = self.feet_contact_forces, self.feet_contact_forces_last
        down = (F_ > F(10.0)) & (Fl < F(10.0))
        vec = self.feet_pos_w - self.feet_down_pos_last
        length = np.sum(vec * self.base_dir_forward_w[:, None, :], axis=-1, dtype=F)
        sgn = np.sign(self.commands[:, 0:1]).astype(F)
        self.feet_step_length = np.where(down, length * sgn, self.feet_step_length).astype(F)
        rew = self.feet_step_length.min(axis=-1)
        self.feet_step_length = (self.feet_step_length * F(0.99)).astype(F)
        self.feet_down_pos_last = np.where(down[..., None], self.feet_pos_w, self.feet_down_pos_last).astype(F)
        self.feet_contact_forces_last = F_.copy()
        return np.tanh(F(15.0) * rew).astype(F)

    def _reward_airtime_variance(self):
        la = np.minimum(self.S["last_air_time"][:, self.feet_ids], F(0.5))
        lc = np.minimum(self.S["last_contact_time"][:, self.feet_ids], F(0.5))
        return (np.var(la, axis=1, ddof=1, dtype=F) + np.var(lc, axis=1, ddof=1, dtype=F)).astype(F)

    def _reward_airtime_sum(self):
        return np.minimum(np.sum(self.S["last_air_time"][:, self.feet_ids], axis=-1, dtype=F), F(2.0))

    def _reward_feet_air_time_biped(self):
        air = self.S["current_air_time"][:, self.feet_ids]
        con = self.S["current_contact_time"][:, self.feet_ids]
        in_contact = con > 0.0
        in_mode = np.where(in_contact, con, air)
        single = np.sum(in_contact.astype(np.int32), axis=1) == 1
        r = np.where(single[:, None], in_mode, F(0.0)).min(axis=1)
        return np.minimum(r, F(2.0)).astype(F)

    def _reward_feet_slide(self):
        contacts = self.feet_contact_forces > F(1.0)
        v = self.S["body_com_lin_vel_w"][:, self.feet_body_idx, :2]
        return np.sum(np.sqrt(np.sum(v * v, axis=-1, dtype=F)) * contacts, axis=1, dtype=F)

    def _reward_feet_harmony(self):
        la = self.S["last_air_time"][:, self.feet_ids]
        return (np.sum(la, axis=-1, dtype=F) - F(3.0) * np.abs(la[:, 0] - la[:, 1])).astype(F)

    def _reward_feet_close(self):
        d = self.feet_pos_w[:, 0, :2] - self.feet_pos_w[:, 1, :2]
        return np.maximum(F(0.115) - np.sqrt(np.sum(d * d, axis=-1, dtype=F)), F(0.0)).astype(F)

    def _reward_action_rate(self):
        d = self.actions - self.prev_actions
        return np.sum(d * d, axis=1, dtype=F)

    def _reward_torques(self):
        t = self.S["applied_torque"]
        return np.sum(t * t, axis=1, dtype=F)

    def _reward_joint_vel(self):
        v = self.S["joint_vel"]
        return np.sum(v * v, axis=1, dtype=F)

    def _reward_joint_acc(self):
        a = self.S["joint_acc"]
        return np.sum(a * a, axis=1, dtype=F)

    def get_rewards(self):                                                               # :853-866
        reward = np.zeros(self.n, F)
        for name, scale in self.reward_scales.items():
            rew = ((getattr(self, "_reward_" + name)() * F(scale)).astype(F) * F(self.step_dt)).astype(F)
            reward = reward + rew
            self.episode_sums[name] = self.episode_sums[name] + rew
        return np.where(self.reset_terminated, reward - F(20.0), reward).astype(F)

    # ---- events ----
    def resample_commands(self, ids, u_sign, u_vel, u_yaw):                              # :109-136
        ev = self.events
        low, high = F(ev["velocity_range"][0]), F(ev["velocity_range"][1])
        if ev["dual_sign"]:
            sign = ((u_sign < F(ev["prob_pos"])).astype(F) * F(2.0) - F(1.0)).astype(F)
            high = (high + F(ev["offset"]) * (sign - F(1.0))).astype(F)
            self.commands[ids, 0] = ((u_vel * (high - low) + low) * sign).astype(F)
        else:
            self.commands[ids, 0] = (u_vel * (high - low) + low).astype(F)
        ylo, yhi = F(ev["yaw_range"][0]), F(ev["yaw_range"][1])
        self.commands[ids, 1] = (u_yaw * (yhi - ylo) + ylo).astype(F)
        self.target_heading_yaw[ids] = wrap_to_pi(self.current_yaw[ids] + self.commands[ids, 1])

    def reset_idx(self, ids, rnd):                                                       # :888-976 + reset-mode events
        S = self.S
        dur = np.maximum(self.episode_length_buf[ids].astype(F) * F(self.step_dt), F(self.step_dt))
        log = {}
        for k in self.episode_sums:
            log["Episode_Reward/" + k] = F(np.mean(self.episode_sums[k][ids] / dur, dtype=F))
            self.episode_sums[k][ids] = 0
        log["Episode_Termination/died"] = int(np.count_nonzero(self.reset_terminated[ids]))
        log["Episode_Termination/time_out"] = int(np.count_nonzero(self.reset_time_outs[ids]))
        self.log = log
        # reset_base (reset_root_state_uniform, :59-106)
        ev = self.events
        lo, hi = np.asarray(ev["pose_lo"], F), np.asarray(ev["pose_hi"], F)
        smp = (rnd[ids, 0:3] * (hi - lo) + lo).astype(F)
        self.current_yaw[ids] = smp[:, 2]
        pos = self.default_root_pos[ids].copy()
        pos[:, 0] += smp[:, 0]
        pos[:, 1] += smp[:, 1]
        quat = np.stack([np.cos(smp[:, 2] * F(0.5)), np.zeros(len(ids), F), np.zeros(len(ids), F),
                         np.sin(smp[:, 2] * F(0.5))], -1).astype(F)
        for row, e in enumerate(ids):
            lp, lq = self.fk(pos[row], quat[row])
            S["body_link_pos_w"][e] = lp
            S["body_link_quat_w"][e] = lq
        for k in ("body_link_lin_vel_w", "body_com_lin_vel_w", "joint_vel", "joint_acc", "applied_torque",
                  "net_forces_w_history", "last_air_time", "last_contact_time", "current_air_time", "current_contact_time"):
            S[k][ids] = 0
        S["joint_pos"][ids] = self.default_joint_pos[ids] if self.default_joint_pos.ndim == 2 else self.default_joint_pos
        # reset_command_resample
        self.resample_commands(ids, rnd[ids, 3], rnd[ids, 4], rnd[ids, 5])
        self.episode_length_buf[ids] = 0
        self.actions[ids] = 0
        self.prev_actions[ids] = 0
        self.p_delta[ids] = 0
        self.feet_contact_forces_last[ids] = F(15.0)
        self.feet_down_pos_last[ids] = S["body_link_pos_w"][:, self.feet_body_idx][ids]
        self.feet_step_length[ids] = 0

    def interval_events(self, rnd):                                                      # [IL-upstream] EventManager.apply("interval")
        self.interval_time_left = (self.interval_time_left - F(self.step_dt)).astype(F)
        ids = np.nonzero(self.interval_time_left < F(1e-6))[0]
        if len(ids) > 0:
            lo, hi = self.events["interval_range_s"]
            self.interval_time_left[ids] = (rnd[ids, 6] * F(hi - lo) + F(lo)).astype(F)
            self.resample_commands(ids, rnd[ids, 7], rnd[ids, 8], rnd[ids, 9])
        return ids

    def observe(self, S):
        self.attach(S)
        self.compute_intermediate_values()
        return self.get_observations()

    def step(self, actions, S1, rnd):
        rnd = np.asarray(rnd, F)
        self.pre_physics_step(actions)
        self.attach(S1)
        self.episode_length_buf += 1
        self.reset_terminated, self.reset_time_outs = self.get_dones()
        rew = self.get_rewards()
        ids = np.nonzero(self.reset_terminated | self.reset_time_outs)[0]
        log = None
        if len(ids) > 0:
            self.reset_idx(ids, rnd)
            log = self.log
        interval_ids = self.interval_events(rnd)
        obs = self.get_observations()
        return obs, rew, self.reset_terminated.copy(), self.reset_time_outs.copy(), ids, interval_ids, log

    def mdp_state(self):
        out = {"p_delta": self.p_delta, "actions": self.actions, "commands": self.commands,
               "target_heading_yaw": self.target_heading_yaw, "current_yaw": self.current_yaw,
               "feet_contact_forces_last": self.feet_contact_forces_last, "feet_down_pos_last": self.feet_down_pos_last,
               "feet_step_length": self.feet_step_length, "episode_length_buf": self.episode_length_buf,
               "interval_time_left": self.interval_time_left}
        for k, v in self.episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: np.array(v, copy=True) for k, v in out.items()}
