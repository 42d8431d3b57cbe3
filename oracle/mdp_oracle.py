"""TEST INFRASTRUCTURE (oracle) -- numpy float32 restatement of the MDP half of
``ZbotDirectEnvV2`` (reference: ``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/
zbot_direct_6dof_bipedal_env_v2.py``; each function cites the lines it follows).

PINNED: ``tests/test_oracle_golden.py`` checks this restatement against
``tests/golden/mdp_v2_*.npz``, which were produced by running the reference's own
unmodified code (``oracle/ref_loader.py`` + ``tests/golden/make_golden.py``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
legs may import this module; the product path never does.
"""
from __future__ import annotations

import numpy as np

from .il_semantics import quat_apply

F = np.float32

#: …env_v2.py:190-206 (dict insertion order == evaluation order, SURVEY C-4)
REWARD_SCALES_V2 = {
    "base_vel_forward": 1.0, "feet_downward": -2.0, "feet_forward": -1.0, "base_heading_x": -1.0,
    "base_heading_x_sum": -5.0, "step_length": 5.0, "airtime_balance": -15.0, "action_rate": -0.1,
    "torques": -0.002, "feet_slide": -10.0, "base_pos_y_err": -2.0, "base_pos_y_err_sum": -2.0,
    "airtime_sum": 3.0,
}


def _norm3(v):
    return np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2])


class MdpOracle:
    """State + methods named after the reference's (``_pre_physics_step`` ...)."""

    def __init__(self, num_envs, env_origins, reset_tables, index_sets, default_joint_pos,
                 reward_scales=None, step_dt=0.02, termination_height=0.22,
                 max_episode_length=1000, max_episode_length_s=20.0):
        n = self.n = int(num_envs)
        self.origins = np.asarray(env_origins, F)
        self.reset_tables = {k: np.asarray(v, F) for k, v in reset_tables.items()}
        self.feet_ids = list(index_sets["feet_ids"])
        self.undesired_ids = list(index_sets["undesired_ids"])
        self.base_body_idx = int(index_sets["base_body_idx"][0])
        self.feet_body_idx = list(index_sets["feet_body_idx"])
        self.default_joint_pos = np.asarray(default_joint_pos, F)
        self.step_dt = step_dt
        self.termination_height = termination_height
        self.max_episode_length = max_episode_length
        self.max_episode_length_s = max_episode_length_s
        scales = dict(REWARD_SCALES_V2 if reward_scales is None else reward_scales)
        # …env_v2.py:250-251: python-float product, used as a python scalar afterwards
        self.reward_scales = {k: v * step_dt for k, v in scales.items()}
        self.episode_sums = {k: np.zeros(n, F) for k in scales}
        self.actions = np.zeros((n, 6), F)
        self.prev_actions = np.zeros((n, 6), F)
        self.p_delta = np.zeros((n, 6), F)
        self.feet_contact_forces_last = np.zeros((n, 2), F)
        self.feet_down_pos_last = np.zeros((n, 2, 3), F)
        self.feet_step_length = np.zeros((n, 2), F)
        self.feet_force_sum = np.zeros(n, F)
        self.base_heading_x_sum = np.zeros(n, F)
        self.base_pos_y_err_sum = np.zeros(n, F)
        self.joint_speed_limit = np.ones((n, 1), F)
        self.episode_length_buf = np.zeros(n, np.int64)
        self.reset_terminated = np.zeros(n, bool)
        self.reset_time_outs = np.zeros(n, bool)
        self.S = None
        self.log = None

    # ------------------------------------------------------------------ data
    def attach(self, S):
        self.S = {k: np.array(v, copy=True) for k, v in S.items()}

    # ------------------------------------------------------------------ …env_v2.py:276-287
    def pre_physics_step(self, actions):
        self.actions = np.tanh(np.asarray(actions, F)).astype(F)
        self.p_delta = self.p_delta + (F(np.pi) * self.actions * self.joint_speed_limit * F(self.step_dt))
        self.p_delta = np.clip(self.p_delta, F(-np.pi), F(np.pi)).astype(F)
        self.processed_actions = self.p_delta + self.default_joint_pos

    # ------------------------------------------------------------------ …env_v2.py:312-369
    def get_observations(self):
        S = self.S
        self.prev_actions = self.actions.copy()
        self.base_pos_w = S["body_link_pos_w"][:, self.base_body_idx].copy()
        self.base_quat_w = S["body_link_quat_w"][:, self.base_body_idx].copy()
        self.feet_quat_w = S["body_link_quat_w"][:, self.feet_body_idx].copy()
        self.feet_pos_w = S["body_link_pos_w"][:, self.feet_body_idx].copy()
        n = self.n
        axis_z = np.tile(np.array([0, 0, 1], F), (n, 1))
        self.base_shoulder_w = quat_apply(self.base_quat_w, axis_z)
        gravity = np.tile(np.array([0, 0, -1], F), (n, 1))
        self.base_dir_forward_w = np.cross(gravity, self.base_shoulder_w).astype(F)  # NOT normalised (C-2)
        self.base_heading_x_err = -self.base_dir_forward_w[:, 1]
        self.base_lin_vel_w = S["body_com_lin_vel_w"][:, self.base_body_idx].copy()
        self.base_lin_vel_forward_w = np.sum(self.base_lin_vel_w * self.base_dir_forward_w, axis=-1, dtype=F)
        self.z_w = np.tile(np.array([0, 0, 1], F), (n, 2, 1))
        axis_x_feet = np.tile(np.array([1, 0, 0], F), (n, 2, 1))
        axis_z_feet = np.tile(np.array([[0, 0, 1], [0, 0, -1]], F), (n, 1, 1))
        self.feet_z_w = quat_apply(self.feet_quat_w, axis_z_feet)
        self.feet_x_w = quat_apply(self.feet_quat_w, axis_x_feet)
        obs = np.concatenate(
            [self.base_quat_w, S["joint_pos"] - self.default_joint_pos, S["joint_vel"], self.actions,
             self.joint_speed_limit], axis=-1).astype(F)
        return obs

    # ------------------------------------------------------------------ …env_v2.py:384-411
    def get_dones(self):
        S = self.S
        time_out = self.episode_length_buf >= self.max_episode_length - 1
        hist = S["net_forces_w_history"]
        fz = hist[:, :, self.feet_ids, 2]  # (N,5,2)
        acc = fz[:, 0].copy()
        for t in range(1, fz.shape[1]):
            acc = acc + fz[:, t]
        self.feet_contact_forces = (acc / F(fz.shape[1])).astype(F)
        self.feet_air_times = S["last_air_time"][:, self.feet_ids]
        self.feet_contact_times = S["current_contact_time"][:, self.feet_ids]
        norms = _norm3(hist[:, :, self.undesired_ids])           # (N,5,U)
        died = np.any(norms.max(axis=1) > F(1.0), axis=1)
        died_1 = self.base_pos_w[:, 2] < F(self.termination_height)
        self.base_pos_y_err = self.base_pos_w[:, 1] - self.origins[:, 1]
        died_6 = np.abs(self.base_pos_y_err) > F(0.5)
        return died | died_1 | died_6, time_out

    # ------------------------------------------------------------------ reward terms
    def _reward_feet_forward(self):  # :461-469
        return np.sum(_norm3(self.feet_x_w - self.base_dir_forward_w[:, None, :]), axis=-1, dtype=F)

    def _reward_feet_downward(self):  # :471-479
        return np.sum(_norm3(self.feet_z_w - self.z_w), axis=-1, dtype=F)

    def _reward_base_heading_x(self):  # :481-482
        return np.abs(self.base_heading_x_err)

    def _reward_base_heading_x_sum(self):  # :484-487
        self.base_heading_x_sum = self.base_heading_x_sum + F(0.01) * self.base_heading_x_err
        self.base_heading_x_sum = np.clip(self.base_heading_x_sum, F(-1), F(1))
        return np.abs(self.base_heading_x_sum)

    def _reward_base_vel_forward(self):  # :489-491
        return np.tanh(F(10.0) * self.base_lin_vel_forward_w / self.joint_speed_limit[:, 0]).astype(F)

    def _reward_base_pos_y_err(self):  # :493-495
        oy = self.origins[:, 1]
        return (np.abs(self.feet_pos_w[:, 0, 1] + self.feet_pos_w[:, 1, 1] - F(2.0) * oy)
                + np.abs(self.base_pos_w[:, 1] - oy))

    def _reward_base_pos_y_err_sum(self):  # :497-500
        self.base_pos_y_err_sum = self.base_pos_y_err_sum + F(0.01) * self.base_pos_y_err
        self.base_pos_y_err_sum = np.clip(self.base_pos_y_err_sum, F(-1), F(1))
        return np.abs(self.base_pos_y_err_sum)

    def _reward_action_rate(self):  # :502-507
        d = self.actions - self.prev_actions
        return np.sum(d * d, axis=1, dtype=F)

    def _reward_step_length(self):  # :509-533
        force_c = F(10.0)
        down = (self.feet_contact_forces > force_c) & (self.feet_contact_forces_last < force_c)
        vec = self.feet_pos_w - self.feet_down_pos_last
        length = np.sum(vec * self.base_dir_forward_w[:, None, :], axis=-1, dtype=F)
        self.feet_step_length = np.where(down, length, self.feet_step_length)
        rew = np.min(self.feet_step_length, axis=-1)
        self.feet_down_pos_last = np.where(down[..., None], self.feet_pos_w, self.feet_down_pos_last)
        self.feet_contact_forces_last = self.feet_contact_forces.copy()
        return np.tanh(F(15.0) * rew).astype(F)

    def _reward_airtime_balance(self):  # :535-539
        return np.abs(self.feet_air_times[:, 0] - self.feet_air_times[:, 1])

    def _reward_airtime_sum(self):  # :541-543
        return np.tanh(np.sum(self.feet_air_times, axis=-1, dtype=F)).astype(F)

    def _reward_feet_slide(self):  # :545-556
        contacts = self.feet_contact_forces > F(1.0)
        v = self.S["body_com_lin_vel_w"][:, self.feet_body_idx, :2]
        speed = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1])
        return np.sum(speed * contacts.astype(F), axis=1, dtype=F)

    def _reward_torques(self):  # :558-561
        t = self.S["applied_torque"]
        return np.sum(t * t, axis=1, dtype=F)

    def _reward_feet_force_diff(self):  # :563-565 (inactive in v2's scale dict)
        return (self.feet_contact_forces[:, 1] - self.feet_contact_forces[:, 0]) * np.sign(self.feet_force_sum)

    def _reward_feet_force_sum(self):  # :567-571 (inactive in v2's scale dict)
        self.feet_force_sum = self.feet_force_sum + F(0.001) * (
            self.feet_contact_forces[:, 0] - self.feet_contact_forces[:, 1])
        return np.abs(self.feet_force_sum)

    # ------------------------------------------------------------------ …env_v2.py:371-382
    def get_rewards(self):
        reward = np.zeros(self.n, F)
        self.last_terms = {}
        for name, scale in self.reward_scales.items():
            rew = (getattr(self, "_reward_" + name)() * F(scale)).astype(F)
            reward = reward + rew
            self.episode_sums[name] = self.episode_sums[name] + rew
            self.last_terms[name] = rew
        reward = np.where(self.reset_terminated, reward - F(20.0), reward).astype(F)
        return reward

    # ------------------------------------------------------------------ …env_v2.py:413-459
    def reset_idx(self, ids, rng_episode_lengths=None):
        S = self.S
        t = self.reset_tables
        # DirectRLEnv._reset_idx + robot.reset/scene.reset effects on the data
        self.episode_length_buf[ids] = 0
        if len(ids) == self.n and rng_episode_lengths is not None:
            self.episode_length_buf[:] = rng_episode_lengths  # …env_v2.py:418-422 (torch RNG, host-supplied)
        self.actions[ids] = 0
        self.prev_actions[ids] = 0
        S["body_link_pos_w"][ids] = t["body_link_pos_local"][None] + self.origins[ids][:, None, :]
        S["body_link_quat_w"][ids] = t["body_link_quat"][None]
        S["body_com_lin_vel_w"][ids] = 0
        S["joint_pos"][ids] = self.default_joint_pos[ids]
        S["joint_vel"][ids] = 0
        S["applied_torque"][ids] = 0
        S["net_forces_w_history"][ids] = 0
        S["last_air_time"][ids] = 0
        S["current_contact_time"][ids] = 0
        self.p_delta[ids] = 0
        self.feet_down_pos_last[ids] = S["body_link_pos_w"][:, self.feet_body_idx][ids]
        self.feet_force_sum[ids] = 0
        self.base_heading_x_sum[ids] = 0
        self.base_pos_y_err_sum[ids] = 0
        log = {}
        for k in self.episode_sums:
            log["Episode_Reward/" + k] = F(np.mean(self.episode_sums[k][ids], dtype=F)) / F(self.max_episode_length_s)
            self.episode_sums[k][ids] = 0
        log["Episode_Termination/body_contact"] = int(np.count_nonzero(self.reset_terminated[ids]))
        log["Episode_Termination/time_out"] = int(np.count_nonzero(self.reset_time_outs[ids]))
        self.log = log

    # ------------------------------------------------------------------ protocol (SURVEY D)
    def observe(self, S):
        self.attach(S)
        return self.get_observations()

    def step(self, actions, S1):
        self.pre_physics_step(actions)
        self.attach(S1)
        self.episode_length_buf += 1
        self.reset_terminated, self.reset_time_outs = self.get_dones()
        rew = self.get_rewards()
        ids = np.nonzero(self.reset_terminated | self.reset_time_outs)[0]
        log = None
        if len(ids) > 0:
            self.reset_idx(ids)
            log = self.log
        obs = self.get_observations()
        return obs, rew, self.reset_terminated.copy(), self.reset_time_outs.copy(), ids, log

    def mdp_state(self):
        out = {
            "p_delta": self.p_delta, "actions": self.actions, "prev_actions": self.prev_actions,
            "feet_contact_forces_last": self.feet_contact_forces_last,
            "feet_down_pos_last": self.feet_down_pos_last, "feet_step_length": self.feet_step_length,
            "base_heading_x_sum": self.base_heading_x_sum, "base_pos_y_err_sum": self.base_pos_y_err_sum,
            "episode_length_buf": self.episode_length_buf,
        }
        for k, v in self.episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: np.array(v, copy=True) for k, v in out.items()}
