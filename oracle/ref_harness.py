"""TEST INFRASTRUCTURE -- drives the reference's own MDP code through one control step.

Restates the *framework* half the reference relies on but does not contain
(``DirectRLEnv.step`` ordering, the effect of ``write_*_to_sim`` on ``robot.data``,
``ContactSensor.reset``), following SURVEY.md §3.2 / Appendix B.1 / Appendix D, so the
reference's unmodified ``_pre_physics_step / _get_dones / _get_rewards / _reset_idx /
_get_observations`` (``…env_v2.py:276-459``) can be run on synthetic articulation
state.  Used by ``tests/golden/make_golden.py`` (build container only).
"""
from __future__ import annotations

import numpy as np
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


class RefV4Harness:
    """The reference ``Zbot6SEnvV4`` driven in ``DirectRLEnv.step`` order, including the EventManager's "reset" and
    "interval" modes ([IL-upstream] ``DirectRLEnv._reset_idx`` / ``step`` and ``EventManager.apply``): the
    reference's OWN ``reset_root_state_uniform`` / ``my_curriculum`` / ``range_curriculum`` / ``resample_commands``
    (…env_v4.py:59-265) are called in cfg order.  Random numbers: ``torch.rand`` / ``torch.bernoulli`` are patched
    for the duration of each event call so that env e consumes ``rnd[e, slot]`` (slot map: V4RandSlot in
    csrc/zbot_core.h) -- bernoulli(p) := (u < p)."""

    ROBOT_KEYS = ("body_link_pos_w", "body_link_quat_w", "body_link_lin_vel_w", "body_com_lin_vel_w", "joint_pos",
                  "joint_vel", "joint_acc", "applied_torque")
    SENSOR_KEYS = ("net_forces_w_history", "last_air_time", "last_contact_time", "current_air_time", "current_contact_time")

    def __init__(self, num_envs, env_origins, index_sets, default_joint_pos, default_root_state, interval_time_left):
        from zbot_lab_b200.assets import zbot_6s as Z
        self.Z = Z
        self.n = num_envs
        self.ref = ref_loader.load_reference_module(ref_loader.REF_ENV_V4)
        self.env = ref_loader.make_reference_v4_env(num_envs, default_joint_pos=default_joint_pos,
                                                    default_root_state=default_root_state, env_origins=env_origins,
                                                    **index_sets)
        self.time_left = torch.as_tensor(interval_time_left, dtype=torch.float32).clone()
        self.rnd = None
        rob = self.env._robot
        harness = self
        rob.write_root_pose_to_sim = lambda pose, env_ids: harness._write_root_pose(pose, env_ids)
        rob.write_root_velocity_to_sim = lambda vel, env_ids: None
        rob.write_joint_state_to_sim = lambda p, v, _i, env_ids: None
        self.env._ref_reset_events = self._reset_events

    # -- RNG plumbing ------------------------------------------------------------------
    class _Patched:
        def __init__(self, queue):
            self.queue = list(queue)

        def __enter__(self):
            self._rand, self._bern = torch.rand, torch.bernoulli
            q = self.queue

            def rand(*size, **kw):
                u = q.pop(0)
                want = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list)) else tuple(size)
                assert tuple(u.shape) == want, (u.shape, want)
                return u.clone()

            def bernoulli(p, **kw):
                u = q.pop(0)
                return (u < p).to(p.dtype)

            torch.rand, torch.bernoulli = rand, bernoulli
            return self

        def __exit__(self, *a):
            torch.rand, torch.bernoulli = self._rand, self._bern
            assert not self.queue, "event drew fewer random tensors than queued"

    # -- framework half ----------------------------------------------------------------
    def attach(self, S: dict):
        d = self.env._robot.data
        for k in self.ROBOT_KEYS:
            setattr(d, k, torch.as_tensor(S[k]).clone())
        c = self.env._contact_sensor.data
        for k in self.SENSOR_KEYS:
            setattr(c, k, torch.as_tensor(S[k]).clone())

    def _write_root_pose(self, pose, env_ids):
        """CPU-PhysX semantics (SURVEY C-5): link poses follow the root write; joints = default; sensors reset."""
        d, c = self.env._robot.data, self.env._contact_sensor.data
        import numpy as np
        for row, e in enumerate(env_ids.tolist()):
            p, q = self.Z.fk_links(pose[row, :3].double().numpy(), pose[row, 3:7].double().numpy(),
                                   np.asarray(self.Z.DEFAULT_JOINT_POS, np.float64))
            d.body_link_pos_w[e] = torch.from_numpy(p).float()
            d.body_link_quat_w[e] = torch.from_numpy(q).float()
        d.body_link_lin_vel_w[env_ids] = 0.0
        d.body_com_lin_vel_w[env_ids] = 0.0
        d.joint_pos[env_ids] = d.default_joint_pos[env_ids]
        d.joint_vel[env_ids] = 0.0
        d.joint_acc[env_ids] = 0.0
        d.applied_torque[env_ids] = 0.0
        for k in self.SENSOR_KEYS:
            getattr(c, k)[env_ids] = 0.0

    def _reset_events(self, env_ids):
        e, ev, r = self.env, self.env.cfg.events, self.rnd
        k = len(env_ids)
        pose = torch.full((k, 6), 0.5)
        pose[:, 0], pose[:, 1], pose[:, 5] = r[env_ids, 0], r[env_ids, 1], r[env_ids, 2]
        with self._Patched([pose, torch.full((k, 6), 0.5)]):
            ev.reset_base.func(e, env_ids, **ev.reset_base.params)
        ev.my_curric.func(e, env_ids)
        ev.vel_range.func(e, env_ids, **ev.vel_range.params)
        p = ev.reset_command_resample.params
        with self._Patched(([r[env_ids, 3]] if p["dual_sign"] else []) + [r[env_ids, 4], r[env_ids, 5]]):
            ev.reset_command_resample.func(e, env_ids, **p)

    def _interval_events(self):
        e, ev, r = self.env, self.env.cfg.events, self.rnd
        term = ev.interval_command_resample
        self.time_left -= e.step_dt
        ids = (self.time_left < 1e-6).nonzero().flatten()
        if len(ids) > 0:
            lower, upper = term.interval_range_s
            with self._Patched([r[ids, 6]]):
                self.time_left[ids] = torch.rand(len(ids)) * (upper - lower) + lower
            p = term.params
            with self._Patched(([r[ids, 7]] if p["dual_sign"] else []) + [r[ids, 8], r[ids, 9]]):
                term.func(e, ids, **p)
        return ids

    # -- protocol ----------------------------------------------------------------------
    def observe(self):
        self.env._compute_intermediate_values()
        return self.env._get_observations()["policy"]

    def step(self, actions, S1, rnd):
        e = self.env
        self.rnd = torch.as_tensor(rnd)
        e._pre_physics_step(torch.as_tensor(actions))
        self.attach(S1)
        e.episode_length_buf += 1
        e.common_step_counter += 1
        e.reset_terminated, e.reset_time_outs = e._get_dones()
        rew = e._get_rewards()
        ids = (e.reset_terminated | e.reset_time_outs).nonzero(as_tuple=False).squeeze(-1)
        log = None
        if len(ids) > 0:
            e._reset_idx(ids)
            log = dict(e.extras["log"])
        interval_ids = self._interval_events()
        obs = e._get_observations()["policy"]
        return obs, rew, e.reset_terminated.clone(), e.reset_time_outs.clone(), ids, interval_ids, log

    def mdp_state(self) -> dict:
        e = self.env
        out = {"p_delta": e.p_delta, "actions": e._actions, "commands": e.commands,
               "target_heading_yaw": e.target_heading_yaw, "current_yaw": e.current_yaw,
               "feet_contact_forces_last": e.feet_contact_forces_last, "feet_down_pos_last": e.feet_down_pos_last,
               "feet_step_length": e.feet_step_length, "episode_length_buf": e.episode_length_buf,
               "interval_time_left": self.time_left}
        for k, v in e._episode_sums.items():
            out["episode_sum/" + k] = v
        return {k: v.clone() for k, v in out.items()}


class RefMHarness:
    """The reference's OWN manager-task term functions (``zbotlab_manager/mdp/rewards.py``, ``mdp/terminations.py``),
    loaded unmodified behind the stub modules and called on a fake ``env`` whose ``scene`` / ``command_manager`` hand out
    plain CPU tensors.  ``call(func_name, view, params)`` evaluates ONE term exactly as RewardManager would
    (``func(env, **params)``); the per-env attributes of ``init_my_data`` (rewards.py:29-35) live on ``self.env`` and are
    mutated by the reference code itself.  Isaac Lab functions that are not in the reference tree are not here (they are
    restated, flagged [IL-upstream], in ``oracle/m_mdp_oracle.py``)."""

    FEET_IDS = [10, 11]          # articulation indices of foot0 / foot1 (assets/zbot_6s_v2.py LINK_NAMES)

    def __init__(self, n, step_dt=0.02):
        import os
        self.rew = ref_loader.load_reference_module(os.path.join(ref_loader.REF_M_MDP_DIR, "rewards.py"))
        self.term = ref_loader.load_reference_module(os.path.join(ref_loader.REF_M_MDP_DIR, "terminations.py"))
        self.n = n
        NS = ref_loader._NS
        env = NS()
        env.num_envs, env.device, env.step_dt = n, torch.device("cpu"), step_dt
        env.sim = NS()
        env.sim.device = "cpu"
        env.episode_length_buf = torch.zeros(n, dtype=torch.long)
        self.asset, self.sensor = NS(), NS()
        self.asset.data, self.sensor.data, self.sensor.cfg = NS(), NS(), NS()
        self.sensor.cfg.track_air_time = True

        class Scene(dict):
            pass
        scene = Scene(robot=self.asset)
        scene.sensors = {"contact_forces": self.sensor}
        env.scene = scene
        env.command_manager = NS()
        self.command = torch.zeros(n, 3)
        env.command_manager.get_command = lambda name: self.command
        self.env = env
        self.rew.init_my_data(env, None)                                  # rewards.py:29-35
        cfg = ref_loader._Cfg
        self.asset_cfg = cfg(name="robot", body_ids=self.FEET_IDS)
        self.sensor_cfg = cfg(name="contact_forces", body_ids=self.FEET_IDS)

    def attach(self, view: dict):
        """Scatter a view (oracle.m_mdp_oracle.synth_m_views layout) into robot.data / contact sensor tensors."""
        n, t = self.n, lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
        d, sd = self.asset.data, self.sensor.data
        pos, quat, vel = torch.zeros(n, 12, 3), torch.zeros(n, 12, 4), torch.zeros(n, 12, 3)
        quat[..., 0] = 1.0
        pos[:, self.FEET_IDS], quat[:, self.FEET_IDS], vel[:, self.FEET_IDS] = t(view["feet_pos"]), t(view["feet_quat"]), t(view["feet_com_vel"])
        pos[:, 0], quat[:, 0] = t(view["root_pos"]), t(view["root_quat"])
        d.body_link_pos_w = d.body_pos_w = pos
        d.body_link_quat_w = quat
        d.body_lin_vel_w = vel                                             # CoM velocity [IL-upstream naming]
        d.root_quat_w = t(view["root_quat"])
        d.root_link_lin_vel_w = t(view["root_lin_vel"])
        d.root_link_ang_vel_w = t(view["root_ang_vel"])
        d.GRAVITY_VEC_W = torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1)
        hist = torch.zeros(n, 3, 12, 3)
        hist[:, :, self.FEET_IDS] = t(view["feet_force_hist"])
        sd.net_forces_w_history = hist
        for name, key in (("last_air_time", "last_air"), ("last_contact_time", "last_contact"),
                          ("current_air_time", "cur_air"), ("current_contact_time", "cur_contact")):
            full = torch.zeros(n, 12)
            full[:, self.FEET_IDS] = t(view[key])
            setattr(sd, name, full)

    def call(self, func: str, params: dict) -> torch.Tensor:
        kw = dict(params)
        fn = getattr(self.rew, func)
        import inspect
        sig = inspect.signature(fn).parameters
        if "asset_cfg" in sig and func != "base_vel_forward" and not func.startswith("track_"):
            kw["asset_cfg"] = self.asset_cfg
        if "sensor_cfg" in sig:
            kw["sensor_cfg"] = self.sensor_cfg
        if func in ("base_vel_forward", "track_lin_vel_xy_yaw_frame_exp", "track_ang_vel_z_world_exp"):
            kw["asset_cfg"] = ref_loader._Cfg(name="robot")
        return fn(self.env, **kw)

    def feet_close(self, minimum_distance: float) -> torch.Tensor:
        return self.term.feet_close(self.env, minimum_distance, self.asset_cfg)

    def reset_my_data(self, env_ids: torch.Tensor):
        self.rew.reset_my_data(self.env, env_ids, self.asset_cfg)
