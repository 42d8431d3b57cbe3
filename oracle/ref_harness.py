"""TEST INFRASTRUCTURE -- drives the reference's own MDP code through one control step.

Restates the *framework* half the reference relies on but does not contain
(``DirectRLEnv.step`` ordering, the effect of ``write_*_to_sim`` on ``robot.data``,
``ContactSensor.reset``), following SURVEY.md §3.2 / Appendix B.1 / Appendix D, so the
reference's unmodified ``_pre_physics_step / _get_dones / _get_rewards / _reset_idx /
_get_observations`` (``…env_v2.py:276-459``) can be run on synthetic articulation
state.  Used by ``tests/golden/make_golden.py`` (build container only).
"""
from __future__ import annotations

import torch

from . import ref_loader


class RefMdpHarness:
    """One reference env + the call order of ``DirectRLEnv.step`` steps 1,3-8 (physics
    replaced by attaching a synthetic end-of-physics state)."""

    def __init__(self, num_envs, env_origins, reset_tables, index_sets, default_joint_pos,
                 default_root_state):
        self.n = num_envs
        self.reset_tables = {k: torch.as_tensor(v).clone() for k, v in reset_tables.items()}
        self.env = ref_loader.make_reference_env(
            num_envs, default_joint_pos=default_joint_pos, default_root_state=default_root_state,
            env_origins=env_origins, **index_sets)
        rob = self.env._robot
        harness = self

        # CPU-PhysX semantics: link poses follow the root/joint writes immediately, so
        # ``feet_down_pos_last[ids]`` (…env_v2.py:436) sees the POST-reset pose (SURVEY C-5).
        def write_root_pose_to_sim(pose, env_ids):
            harness._apply_reset_rows(env_ids)

        rob.write_root_pose_to_sim = write_root_pose_to_sim
        rob.write_root_velocity_to_sim = lambda vel, env_ids: None
        rob.write_joint_state_to_sim = lambda p, v, _i, env_ids: None

    # -- state attachment ------------------------------------------------------------
    def attach(self, S: dict):
        d = self.env._robot.data
        for k in ("body_link_pos_w", "body_link_quat_w", "body_com_lin_vel_w", "joint_pos",
                  "joint_vel", "applied_torque"):
            setattr(d, k, torch.as_tensor(S[k]).clone())
        c = self.env._contact_sensor.data
        for k in ("net_forces_w_history", "last_air_time", "current_contact_time"):
            setattr(c, k, torch.as_tensor(S[k]).clone())

    def _apply_reset_rows(self, env_ids):
        d = self.env._robot.data
        c = self.env._contact_sensor.data
        t = self.reset_tables
        org = self.env._terrain.env_origins[env_ids]
        d.body_link_pos_w[env_ids] = t["body_link_pos_local"].unsqueeze(0) + org.unsqueeze(1)
        d.body_link_quat_w[env_ids] = t["body_link_quat"].unsqueeze(0).expand(len(env_ids), -1, -1)
        d.body_com_lin_vel_w[env_ids] = 0.0
        d.joint_pos[env_ids] = d.default_joint_pos[env_ids]
        d.joint_vel[env_ids] = 0.0
        d.applied_torque[env_ids] = 0.0
        # ContactSensor.reset(ids): forces, history and all timers -> 0
        c.net_forces_w_history[env_ids] = 0.0
        c.last_air_time[env_ids] = 0.0
        c.current_contact_time[env_ids] = 0.0

    # -- protocol ----------------------------------------------------------------------
    def observe(self):
        return self.env._get_observations()["policy"]

    def step(self, actions, S1):
        e = self.env
        e._pre_physics_step(torch.as_tensor(actions))
        self.attach(S1)
        e.episode_length_buf += 1
        e.reset_terminated, e.reset_time_outs = e._get_dones()
        rew = e._get_rewards()
        reset_buf = e.reset_terminated | e.reset_time_outs
        ids = reset_buf.nonzero(as_tuple=False).squeeze(-1)
        log = None
        if len(ids) > 0:
            e._reset_idx(ids)
            log = dict(e.extras["log"])
        obs = e._get_observations()["policy"]
        return obs, rew, e.reset_terminated.clone(), e.reset_time_outs.clone(), ids, log

    def mdp_state(self) -> dict:
        e = self.env
        out = {
            "p_delta": e.p_delta, "actions": e._actions, "prev_actions": e._previous_actions,
            "feet_contact_forces_last": e.feet_contact_forces_last,
            "feet_down_pos_last": e.feet_down_pos_last, "feet_step_length": e.feet_step_length,
            "base_heading_x_sum": e.base_heading_x_sum, "base_pos_y_err_sum": e.base_pos_y_err_sum,
            "episode_length_buf": e.episode_length_buf,
        }
        for k, v in e._episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: v.clone() for k, v in out.items()}


SNAKE_SENSOR_WIDTHS = (5, 4, 3, 2)   # filter bodies per sensor (zbot_direct_6dof_snake_v0.py:23-48)


class RefSnakeHarness:
    """Same protocol as :class:`RefMdpHarness` for the reference snake task (``ZbotDirectEnvV0``)."""

    ROBOT_KEYS = ("body_link_pos_w", "body_link_quat_w", "body_link_vel_w", "body_com_pos_w", "joint_pos",
                  "joint_vel", "applied_torque")

    def __init__(self, num_envs, env_origins, reset_tables, default_root_state, joint_speed_limit):
        self.n = num_envs
        self.reset_tables = {k: torch.as_tensor(v).clone() for k, v in reset_tables.items()}
        self.env = ref_loader.make_reference_snake_env(num_envs, default_root_state=default_root_state,
                                                       env_origins=env_origins, joint_speed_limit=joint_speed_limit)
        rob = self.env._robot
        harness = self
        rob.write_root_pose_to_sim = lambda pose, env_ids: harness._apply_reset_rows(env_ids)
        rob.write_root_velocity_to_sim = lambda vel, env_ids: None
        rob.write_joint_state_to_sim = lambda p, v, _i, env_ids: None

    def attach(self, S: dict):
        d = self.env._robot.data
        for k in self.ROBOT_KEYS:
            setattr(d, k, torch.as_tensor(S[k]).clone())
        for i in (1, 2, 3, 4):
            getattr(self.env, f"_contact_sensor_{i}").data.force_matrix_w = torch.as_tensor(S[f"force_matrix_w_{i}"]).clone()

    def _apply_reset_rows(self, env_ids):
        d, t = self.env._robot.data, self.reset_tables
        org = self.env._terrain.env_origins[env_ids]
        d.body_link_pos_w[env_ids] = t["body_link_pos_local"].unsqueeze(0) + org.unsqueeze(1)
        d.body_link_quat_w[env_ids] = t["body_link_quat"].unsqueeze(0).expand(len(env_ids), -1, -1)
        d.body_com_pos_w[env_ids] = t["body_com_pos_local"].unsqueeze(0) + org.unsqueeze(1)
        d.body_link_vel_w[env_ids] = 0.0
        d.joint_pos[env_ids] = 0.0
        d.joint_vel[env_ids] = 0.0
        d.applied_torque[env_ids] = 0.0
        for i in (1, 2, 3, 4):
            getattr(self.env, f"_contact_sensor_{i}").data.force_matrix_w[env_ids] = 0.0

    def observe(self):
        return self.env._get_observations()["policy"]

    def step(self, actions, S1):
        e = self.env
        e._pre_physics_step(torch.as_tensor(actions))
        self.attach(S1)
        e.episode_length_buf += 1
        e.reset_terminated, e.reset_time_outs = e._get_dones()
        rew = e._get_rewards()
        ids = (e.reset_terminated | e.reset_time_outs).nonzero(as_tuple=False).squeeze(-1)
        log = None
        if len(ids) > 0:
            e._reset_idx(ids)
            log = dict(e.extras["log"])
        obs = e._get_observations()["policy"]
        return obs, rew, e.reset_terminated.clone(), e.reset_time_outs.clone(), ids, log

    def mdp_state(self) -> dict:
        e = self.env
        out = {"p_delta": e.p_delta, "actions": e._actions, "base_heading_y_sum": e.base_heading_y_sum,
               "base_pos_x_err_sum": e.base_pos_x_err_sum, "episode_length_buf": e.episode_length_buf}
        for k, v in e._episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: v.clone() for k, v in out.items()}
