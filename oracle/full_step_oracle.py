"""TEST INFRASTRUCTURE (oracle) -- the whole ``zbot-6b-walking-v2`` control step on the CPU,
composed from the independent pieces: float64 dynamics (``dyn_oracle.py``), the Isaac Lab
ContactSensor / ImplicitActuator / DirectRLEnv.step semantics (``il_semantics.py``,
SURVEY.md Appendix B) and the reference-pinned MDP restatement (``mdp_oracle.py``).

Order follows ``DirectRLEnv.step``: _pre_physics_step, 4 x (apply action,
actuator bookkeeping, physics, sensor update), episode_length += 1, dones, rewards, partial
reset, observations.  Positions are env-LOCAL (origin subtracted), matching the fused kernel.
PARITY UNPINNED for the dynamics half (see dyn_oracle.py).  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

from zbot_lab_b200.assets import zbot_6s as Z
from zbot_lab_b200.utils import synthetic as syn

from .dyn_oracle import DynOracle, DynParams
from .il_semantics import ContactSensorState
from .mdp_oracle import MdpOracle

F = np.float32
#: sensor index (prim order) of the link each reduced body's contact force is attributed to
BODY_TO_SENSOR = [Z.SENSOR_BODY_NAMES.index(nm) for nm in ("foot_0", "a2", "a3", "base", "a5", "a6", "foot_1")]


class FullStepOracle:
    def __init__(self, n, params: DynParams | None = None, reward_scales=None):
        self.n = n
        self.dyn = DynOracle(n, params)
        self.sensor = ContactSensorState(n, 12, history=5, dtype=np.float64)
        self.mdp = MdpOracle(n, np.zeros((n, 3), F), syn.reset_tables(), syn.index_sets(),
                             np.tile(np.asarray(Z.DEFAULT_JOINT_POS, F), (n, 1)), reward_scales=reward_scales)
        self.decimation = Z.DECIMATION

    def _robot_data(self):
        ls = self.dyn.link_state()
        return {
            "body_link_pos_w": ls["body_link_pos"].astype(F), "body_link_quat_w": ls["body_link_quat"].astype(F),
            "body_com_lin_vel_w": ls["body_com_lin_vel"].astype(F), "joint_pos": self.dyn.q.astype(F),
            "joint_vel": self.dyn.qd.astype(F), "applied_torque": self.dyn.applied_torque.astype(F),
            "net_forces_w_history": self.sensor.net_forces_w_history.astype(F),
            "last_air_time": self.sensor.last_air_time.astype(F),
            "current_contact_time": self.sensor.current_contact_time.astype(F),
        }

    def reset_all(self):
        ids = np.arange(self.n)
        self.dyn.reset(ids)
        self.sensor.reset(ids)
        self.mdp.attach(self._robot_data())
        self.mdp.reset_idx(ids)
        return self.mdp.get_observations()

    def observe(self):
        self.mdp.attach(self._robot_data())
        return self.mdp.get_observations()

    def step(self, actions):
        m, d = self.mdp, self.dyn
        if m.S is None:
            self.observe()
        m.pre_physics_step(actions)
        target = m.processed_actions.astype(np.float64)
        for _ in range(self.decimation):
            d.substep(target)
            net = np.zeros((self.n, 12, 3))
            net[:, BODY_TO_SENSOR[0]] = d.body_force[:, 0]
            net[:, BODY_TO_SENSOR[6]] = d.body_force[:, 6]
            for b in range(1, 6):
                net[:, BODY_TO_SENSOR[b]] = d.body_force_pred[:, b]
            self.sensor.update(net, d.P.dt)
        m.attach(self._robot_data())
        m.episode_length_buf += 1
        m.reset_terminated, m.reset_time_outs = m.get_dones()
        rew = m.get_rewards()
        ids = np.nonzero(m.reset_terminated | m.reset_time_outs)[0]
        log = None
        if len(ids) > 0:
            d.reset(ids)
            self.sensor.reset(ids)
            m.reset_idx(ids)
            log = m.log
        obs = m.get_observations()
        return obs, rew, m.reset_terminated.copy(), m.reset_time_outs.copy(), ids, log


class SnakeFullStepOracle:
    """The whole ``zbot-6s-snake-v0`` control step on the CPU: float64 dynamics of the snake robot model +
    the reference-pinned snake MDP restatement.  The filtered self-contact sensors (snake_v0.py:23-48) are
    modelled as sphere overlaps between link-cylinder centres times the contact spring (a termination proxy
    that is not fed back into the dynamics -- DESIGN.md §3)."""

    def __init__(self, n, joint_speed_limit, params: DynParams | None = None):
        from zbot_lab_b200.assets import zbot_d_6s as S
        from .snake_mdp_oracle import SnakeMdpOracle
        self.n = n
        self.model = S.model_f32()
        self.dyn = DynOracle(n, params, model=self.model)
        self.mdp = SnakeMdpOracle(n, np.zeros((n, 3), F), syn.snake_reset_tables(), joint_speed_limit)
        self.decimation = Z.DECIMATION

    def _robot_data(self):
        ls, P, m = self.dyn.link_state(), self.dyn.P, self.model
        vel = np.zeros((self.n, 12, 6))
        vel[..., :3] = ls["body_link_lin_vel"]
        out = {
            "body_link_pos_w": ls["body_link_pos"].astype(F), "body_link_quat_w": ls["body_link_quat"].astype(F),
            "body_link_vel_w": vel.astype(F), "body_com_pos_w": ls["body_com_pos"].astype(F),
            "joint_pos": self.dyn.q.astype(F), "joint_vel": self.dyn.qd.astype(F),
            "applied_torque": self.dyn.applied_torque.astype(F),
        }
        k_n = P.alpha * P.erp / P.dt
        ctr = ls["link_centre"]
        pair_force = {}
        for (a, b) in m.self_pairs:
            d = np.linalg.norm(ctr[:, a] - ctr[:, b], axis=-1)
            pair_force[(a, b)] = k_n * np.maximum(2.0 * m.sphere_radius - d, 0.0)
        from zbot_lab_b200.assets.zbot_d_6s import LINK_NAMES, SELF_CONTACT_SENSORS
        for i, (s_name, filt) in enumerate(SELF_CONTACT_SENSORS, start=1):
            fm = np.zeros((self.n, 1, len(filt), 3), F)
            for j, f_name in enumerate(filt):
                fm[:, 0, j, 0] = pair_force[(LINK_NAMES.index(s_name), LINK_NAMES.index(f_name))]
            out[f"force_matrix_w_{i}"] = fm
        return out

    def reset_all(self):
        ids = np.arange(self.n)
        self.dyn.reset(ids)
        self.mdp.attach(self._robot_data())
        self.mdp.reset_idx(ids)
        return self.mdp.get_observations()

    def step(self, actions):
        m, d = self.mdp, self.dyn
        if m.S is None:
            m.attach(self._robot_data())
            m.get_observations()
        m.pre_physics_step(actions)
        target = m.processed_actions.astype(np.float64)
        for _ in range(self.decimation):
            d.substep(target)
        m.attach(self._robot_data())
        m.episode_length_buf += 1
        m.reset_terminated, m.reset_time_outs = m.get_dones()
        rew = m.get_rewards()
        ids = np.nonzero(m.reset_terminated | m.reset_time_outs)[0]
        if len(ids) > 0:
            d.reset(ids)
            m.reset_idx(ids)
        obs = m.get_observations()
        return obs, rew, m.reset_terminated.copy(), m.reset_time_outs.copy(), ids, m.log if len(ids) else None
