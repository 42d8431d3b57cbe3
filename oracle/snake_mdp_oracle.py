"""TEST INFRASTRUCTURE (oracle) -- numpy float32 restatement of the MDP half of the reference snake task
``ZbotDirectEnvV0`` (``/root/reference/source/zbot/zbot/tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py``;
each function cites the lines it follows).  PINNED by ``tests/golden/snake_v0_*.npz`` (outputs of the
reference's own unmodified code).  Only tests/, smoke() and bench.py's CPU legs may import this."""
from __future__ import annotations

import numpy as np

from .il_semantics import quat_apply

F = np.float32
#: zbot_direct_6dof_snake_v0.py:88-98 (dict order = evaluation order)
REWARD_SCALES_SNAKE = {
    "base_vel_forward": 5.0, "base_up_z": -0.5, "base_heading_y": -1.0, "base_heading_y_sum": -1.0,
    "base_pos_x_err": -1.0, "action_rate": -0.1, "torques": -0.002,
}
BASE_LINK = 6   # zbot_direct_6dof_snake_v0.py:203-204: index 6 = middle link (a4), 0 / 11 = the two ends


class SnakeMdpOracle:
    def __init__(self, num_envs, env_origins, reset_tables, joint_speed_limit, reward_scales=None, step_dt=0.02,
                 max_episode_length=800, max_episode_length_s=16.0):
        n = self.n = int(num_envs)
        self.origins = np.asarray(env_origins, F)
        self.reset_tables = {k: np.asarray(v, F) for k, v in reset_tables.items()}
        self.step_dt, self.max_episode_length, self.max_episode_length_s = step_dt, max_episode_length, max_episode_length_s
        scales = dict(REWARD_SCALES_SNAKE if reward_scales is None else reward_scales)
        self.reward_scales = {k: v * step_dt for k, v in scales.items()}                 # :127-128
        self.episode_sums = {k: np.zeros(n, F) for k in scales}
        self.actions = np.zeros((n, 6), F)
        self.prev_actions = np.zeros((n, 6), F)
        self.p_delta = np.zeros((n, 6), F)
        self.base_heading_y_sum = np.zeros(n, F)
        self.base_pos_x_err_sum = np.zeros(n, F)
        self.joint_speed_limit = np.asarray(joint_speed_limit, F).reshape(n, 1)
        self.episode_length_buf = np.zeros(n, np.int64)
        self.reset_terminated = np.zeros(n, bool)
        self.reset_time_outs = np.zeros(n, bool)
        self.S = None
        self.log = None

    def attach(self, S):
        self.S = {k: np.array(v, copy=True) for k, v in S.items()}

    def pre_physics_step(self, actions):                                                 # :160-170
        self.actions = np.tanh(np.asarray(actions, F)).astype(F)
        self.p_delta = self.p_delta + self.actions * self.joint_speed_limit * F(self.step_dt)
        self.p_delta = np.clip(self.p_delta, F(-np.pi), F(np.pi)).astype(F)
        self.processed_actions = self.p_delta.copy()                                     # default_joint_pos = 0

    def get_observations(self):                                                          # :175-208
        S, n = self.S, self.n
        self.prev_actions = self.actions.copy()
        self.base_pos_w = S["body_link_pos_w"][:, BASE_LINK].copy()
        self.base_quat_w = S["body_link_quat_w"][:, BASE_LINK].copy()
        heading_vec = np.tile(np.array([0, -1, 0], F), (n, 1))
        up_vec = np.tile(np.array([-1, 0, 0], F), (n, 1))
        self.base_heading_w = quat_apply(self.base_quat_w, heading_vec)
        self.base_up_w = quat_apply(self.base_quat_w, up_vec)
        self.base_heading_y_err = -self.base_heading_w[:, 0]
        self.base_lin_vel_w = S["body_link_vel_w"][:, BASE_LINK, :3].copy()
        self.base_lin_vel_forward_w = np.sum(self.base_lin_vel_w * self.base_heading_w, axis=-1, dtype=F)
        return np.concatenate([self.base_quat_w, S["joint_pos"], S["joint_vel"], self.actions,
                               self.joint_speed_limit], axis=-1).astype(F)

    def get_dones(self):                                                                 # :222-240
        S = self.S
        time_out = self.episode_length_buf >= self.max_episode_length - 1
        fm = np.concatenate([S[f"force_matrix_w_{i}"] for i in (1, 2, 3, 4)], axis=2)     # (N,1,14,3)
        norms = np.sqrt(fm[..., 0] * fm[..., 0] + fm[..., 1] * fm[..., 1] + fm[..., 2] * fm[..., 2])
        died = np.any(norms.max(axis=1) > F(1.0), axis=1)
        self.base_pos_x_err = self.base_pos_w[:, 0] - self.origins[:, 0] + F(0.318)
        died |= np.abs(self.base_pos_x_err) > F(0.2)
        return died, time_out

    def _reward_base_vel_forward(self):                                                  # :300-302
        return np.tanh(F(10.0) * self.base_lin_vel_forward_w / self.joint_speed_limit[:, 0]).astype(F)

    def _reward_base_up_z(self):                                                         # :304-305
        return np.abs(self.base_up_w[:, 1])

    def _reward_base_heading_y(self):                                                    # :307-308
        return np.abs(self.base_heading_y_err)

    def _reward_base_heading_y_sum(self):                                                # :310-313
        self.base_heading_y_sum = np.clip(self.base_heading_y_sum + F(0.01) * self.base_heading_y_err, F(-1), F(1))
        return np.abs(self.base_heading_y_sum)

    def _reward_base_pos_x_err(self):                                                    # :329-335
        # the reference returns ONLY this first term: the "+ abs(base_pos_x_err)" on the next source line is a
        # dangling unary-plus statement (SURVEY C-9)
        c = self.S["body_com_pos_w"]
        return np.abs(c[:, 0, 0] + c[:, 11, 0] - F(2.0) * self.origins[:, 0] + F(0.636))

    def _reward_base_pos_x_err_sum(self):                                                # :337-340 (inactive)
        self.base_pos_x_err_sum = np.clip(self.base_pos_x_err_sum + F(0.01) * self.base_pos_x_err, F(-1), F(1))
        return np.abs(self.base_pos_x_err_sum)

    def _reward_action_rate(self):                                                       # :342-346
        d = self.actions - self.prev_actions
        return np.sum(d * d, axis=1, dtype=F)

    def _reward_torques(self):                                                           # :348-350
        t = self.S["applied_torque"]
        return np.sum(t * t, axis=1, dtype=F)

    def get_rewards(self):                                                               # :210-220
        reward = np.zeros(self.n, F)
        for name, scale in self.reward_scales.items():
            rew = (getattr(self, "_reward_" + name)() * F(scale)).astype(F)
            reward = reward + rew
            self.episode_sums[name] = self.episode_sums[name] + rew
        return np.where(self.reset_terminated, reward - F(20.0), reward).astype(F)

    def reset_idx(self, ids):                                                            # :242-298
        S, t = self.S, self.reset_tables
        self.episode_length_buf[ids] = 0
        self.actions[ids] = 0
        self.prev_actions[ids] = 0
        S["body_link_pos_w"][ids] = t["body_link_pos_local"][None] + self.origins[ids][:, None, :]
        S["body_link_quat_w"][ids] = t["body_link_quat"][None]
        S["body_com_pos_w"][ids] = t["body_com_pos_local"][None] + self.origins[ids][:, None, :]
        S["body_link_vel_w"][ids] = 0
        S["joint_pos"][ids] = 0
        S["joint_vel"][ids] = 0
        S["applied_torque"][ids] = 0
        for i in (1, 2, 3, 4):
            S[f"force_matrix_w_{i}"][ids] = 0
        self.p_delta[ids] = 0
        self.base_heading_y_sum[ids] = 0
        self.base_pos_x_err_sum[ids] = 0
        log = {}
        for k in self.episode_sums:
            log["Episode_Reward/" + k] = F(np.mean(self.episode_sums[k][ids], dtype=F)) / F(self.max_episode_length_s)
            self.episode_sums[k][ids] = 0
        log["Episode_Termination/died"] = int(np.count_nonzero(self.reset_terminated[ids]))
        log["Episode_Termination/time_out"] = int(np.count_nonzero(self.reset_time_outs[ids]))
        self.log = log

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
        out = {"p_delta": self.p_delta, "actions": self.actions, "base_heading_y_sum": self.base_heading_y_sum,
               "base_pos_x_err_sum": self.base_pos_x_err_sum, "episode_length_buf": self.episode_length_buf}
        for k, v in self.episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: np.array(v, copy=True) for k, v in out.items()}
