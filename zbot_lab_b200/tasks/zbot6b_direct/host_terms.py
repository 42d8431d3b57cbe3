"""Host-side view of one fused step and the reference's subclass-hook surface for ``zbot-6b-walking-v2``.

The reference discovers its reward terms by name -- ``getattr(self, "_reward_" + name)`` over
``cfg.reward_cfg["reward_scales"]`` (``…/zbot_direct_6dof_bipedal_env_v2.py:246-252``) -- so a user subclass can add a
term or override one.  The fused kernel evaluates the 15 terms it knows (``native.TERM_IDS``); everything else is a HOST
term: a ``_reward_<name>`` method of the (sub)class, called after the kernel on a :class:`StepView`, ``value * scale``
added to the reward (and to its own episode sum / ``extras["log"]`` entry).  A host term is off the fused path: the step
then runs through the export hook (``zbot_step_export``: the same kernel instantiation plus plain stores, two small
view kernels around it) and a handful of torch launches per term.

:class:`StepView` is what such a method reads -- the tensors the reference's methods read, with the reference's
fresh / stale semantics:

* ``env._robot.data.{body_link_pos_w, body_link_quat_w, body_com_lin_vel_w, joint_pos, joint_vel, applied_torque}`` and
  ``env._contact_sensor.data.{net_forces_w_history, last_air_time, current_contact_time}``: END of physics of this step,
  BEFORE any reset (what ``_get_dones`` / ``_get_rewards`` see, …env_v2.py:386-394, 554, 560);
* ``env.base_pos_w, base_quat_w, feet_pos_w, feet_quat_w, base_dir_forward_w, base_heading_x_err, base_lin_vel_w,
  base_lin_vel_forward_w, feet_z_w, feet_x_w``: the values the PREVIOUS ``_get_observations`` cached, i.e. the START of this
  step (…env_v2.py:315-345).

The built-in ``_reward_<name>`` methods of ``ZbotDirectEnvV2`` are torch restatements of the fused terms on this view
(formulas: SURVEY.md §8 a8.1-a8.13; reference lines cited per function).  They are *views*: the three stateful terms (``base_heading_x_sum``, ``step_length``, ``base_pos_y_err_sum``;
also the inactive ``feet_force_sum``) return the value of the integrator / touchdown memory the KERNEL's evaluation of
that term advanced, they never advance it a second time -- so a subclass that overrides one of those must carry its own
state (as the reference's method does), while the ten stateless ones can be moved to the host as they are
(``tests/test_gpu_env.py::test_subclass_reward_hooks_host_terms_and_step_view`` checks all thirteen views against the
kernel term by term).
"""
from __future__ import annotations

import torch

from ...assets import zbot_6s as Z


def quat_apply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """isaaclab.utils.math.quat_apply, wxyz: v + w t + q_xyz x t with t = 2 (q_xyz x v)."""
    xyz = q[..., 1:]
    t = 2.0 * torch.cross(xyz, v, dim=-1)
    return v + q[..., 0:1] * t + torch.cross(xyz, t, dim=-1)


class _Data:
    pass


class StepView:
    """Tensors of ONE step, filled from the export hook of the fused kernel (device tensors, world frame)."""

    def __init__(self, env):
        self.env = env
        self.valid = False
        self.export = env._stepper.alloc_export()
        self.robot = _Data()
        self.sensor = _Data()
        self.stale = {}

    def refresh(self):
        env, ex = self.env, self.export
        org = env._terrain.env_origins.unsqueeze(1)
        r, s = self.robot, self.sensor
        r.body_link_pos_w = ex["body_link_pos_w1"] + org
        r.body_link_quat_w = ex["body_link_quat_w1"]
        r.body_com_lin_vel_w = ex["body_com_lin_vel_w1"]
        r.joint_pos, r.joint_vel, r.applied_torque = ex["joint_pos1"], ex["joint_vel1"], ex["applied_torque1"]
        s.net_forces_w_history = ex["net_forces_w_history1"]
        s.last_air_time, s.current_contact_time = ex["last_air_time1"], ex["current_contact_time1"]
        # what the previous _get_observations cached (…env_v2.py:315-345), from the start-of-step view
        b, f = env.base_body_idx[0], env.feet_body_idx
        pos0 = ex["body_link_pos_w0"] + org
        quat0, vel0 = ex["body_link_quat_w0"], ex["body_com_lin_vel_w0"]
        st = self.stale
        st["base_pos_w"], st["base_quat_w"] = pos0[:, b], quat0[:, b]
        st["feet_pos_w"], st["feet_quat_w"] = pos0[:, f], quat0[:, f]
        n, dev = env.num_envs, env.device
        ez = torch.tensor([0.0, 0.0, 1.0], device=dev).expand(n, 3)
        st["base_shoulder_w"] = quat_apply(st["base_quat_w"], ez)                                   # :322
        st["base_dir_forward_w"] = torch.cross(env._robot.data.GRAVITY_VEC_W, st["base_shoulder_w"], dim=-1)   # :323, not normalised
        st["base_heading_x_err"] = -st["base_dir_forward_w"][:, 1]                                   # :324
        st["base_lin_vel_w"] = vel0[:, b]                                                           # :326
        st["base_lin_vel_forward_w"] = (st["base_lin_vel_w"] * st["base_dir_forward_w"]).sum(-1)     # :327
        axis_z = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]], device=dev).expand(n, 2, 3)       # :341-343
        axis_x = torch.tensor([1.0, 0.0, 0.0], device=dev).expand(n, 2, 3)
        st["feet_z_w"] = quat_apply(st["feet_quat_w"], axis_z)                                       # :344
        st["feet_x_w"] = quat_apply(st["feet_quat_w"], axis_x)                                       # :345
        self.valid = True


class ContactSensor:
    """``env._contact_sensor``: the ContactSensor surface of the reference (…env_v2.py:30-36: all 12 links, 5-deep force
    history, air / contact timers), served from the step view."""

    def __init__(self, env):
        self._env = env
        self.body_names = list(Z.SENSOR_BODY_NAMES)

    def find_bodies(self, pattern):
        return Z.find_bodies(pattern, Z.SENSOR_BODY_NAMES)

    @property
    def data(self):
        return self._env._need_view().sensor


# ---------------------------------------------------------------------------------------------------------------------
# torch restatements of the fused reward terms (views; see the module docstring)
# ---------------------------------------------------------------------------------------------------------------------
class RewardTermViews:
    """Mixin of ``ZbotDirectEnvV2``: ``_reward_<name>`` for the 15 names of ``native.TERM_IDS``."""

    # -- inputs the reference computes in _get_dones (…env_v2.py:386-394) and keeps on self
    def _fresh_feet_forces(self):
        h = self._contact_sensor.data.net_forces_w_history
        return torch.mean(h[:, :, self._feet_ids, 2], dim=1)                                  # :387-390

    def _fresh_air_times(self):
        return self._contact_sensor.data.last_air_time[:, self._feet_ids]                     # :391

    def _reward_base_vel_forward(self):                                                       # :489-491
        return torch.tanh(10.0 * self.base_lin_vel_forward_w / self.joint_speed_limit.squeeze(-1))

    def _reward_feet_downward(self):                                                          # :471-479
        tgt = torch.tensor([0.0, 0.0, 1.0], device=self.device)
        return torch.linalg.norm(self.feet_z_w - tgt, dim=-1).sum(-1)

    def _reward_feet_forward(self):                                                           # :461-469
        return torch.linalg.norm(self.feet_x_w - self.base_dir_forward_w.unsqueeze(1), dim=-1).sum(-1)

    def _reward_base_heading_x(self):                                                         # :481-482
        return self.base_heading_x_err.abs()

    def _reward_base_heading_x_sum(self):                                                     # :484-487 (integrator: kernel state)
        return self.base_heading_x_sum.abs()

    def _reward_step_length(self):                                                            # :509-533 (touchdown memory: kernel state)
        return torch.tanh(15.0 * self.feet_step_length.min(dim=-1).values)

    def _reward_airtime_balance(self):                                                        # :535-539
        a = self._fresh_air_times()
        return (a[:, 0] - a[:, 1]).abs()

    def _reward_action_rate(self):                                                            # :502-507
        v = self._need_view()
        return torch.square(v.actions_now - v.actions_prev).sum(-1)

    def _reward_torques(self):                                                                # :558-561
        return torch.square(self._robot.data.applied_torque).sum(-1)

    def _reward_feet_slide(self):                                                             # :545-556
        vel = self._robot.data.body_com_lin_vel_w[:, self.feet_body_idx, :2]
        return (torch.linalg.norm(vel, dim=-1) * (self._fresh_feet_forces() > 1.0)).sum(-1)

    def _reward_base_pos_y_err(self):                                                         # :493-495
        oy = self._terrain.env_origins[:, 1]
        return (self.feet_pos_w[:, 0, 1] + self.feet_pos_w[:, 1, 1] - 2.0 * oy).abs() + (self.base_pos_w[:, 1] - oy).abs()

    def _reward_base_pos_y_err_sum(self):                                                     # :497-500 (integrator: kernel state)
        return self.base_pos_y_err_sum.abs()

    def _reward_airtime_sum(self):                                                            # :541-543
        return torch.tanh(self._fresh_air_times().sum(-1))

    def _reward_feet_force_diff(self):                                                        # :563-565
        f = self._fresh_feet_forces()
        return (f[:, 1] - f[:, 0]) * torch.sign(self.feet_force_sum)

    def _reward_feet_force_sum(self):                                                         # :567-571 (integrator: kernel state)
        return self.feet_force_sum.abs()
