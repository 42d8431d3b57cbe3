"""``ZbotDirectEnvV0`` -- drop-in for the reference task class of ``zbot-6s-snake-v0``
(``/root/reference/source/zbot/zbot/tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py:101-352``) over the
fused sm_100a step (``zbot_snake_step_kernel``).  Same surface as the walking task's class; what differs:

* robot ``ZBOT_D_6S_CFG`` (kp 20 / kd 0.5, lying on the ground, root = a1; assets/zbot_cfg.py:109-168);
* per-env ``joint_speed_limit = (U[0,1) * 1.8 + 0.2) * pi`` drawn ONCE in the constructor on the torch
  generator (snake_v0.py:121) -- it survives resets and is the last observation column;
* terminations: filtered self-contact force > 1 N, |base_pos_x_err| > 0.2 (snake_v0.py:222-240); log key
  ``Episode_Termination/died`` (snake_v0.py:289);
* reward table ``cfg.reward_cfg["reward_scales"]`` over the snake term names (snake_v0.py:88-98, 300-350).
"""
from __future__ import annotations

import math

import torch

from ... import native
from ...assets import zbot_d_6s as S
from ...assets import zbot_6s as Z
from ..zbot6b_direct.walking_v2 import ZbotDirectEnvV2
from .snake_v0_cfg import ZbotDirectEnvCfgV0


class _SnakeArticulationData:
    """``robot.data`` of the snake robot: defaults + joint state (snake_v0.py:167, 175-176, 193-194, 265-267)."""

    def __init__(self, env):
        self._env = env
        n, dev = env.num_envs, env.device
        m = S.model_f32()
        self.default_joint_pos = torch.zeros(n, 6, device=dev)
        self.default_joint_vel = torch.zeros(n, 6, device=dev)
        drs = torch.zeros(n, 13, device=dev)
        drs[:, :3] = torch.tensor(m.default_root_pos, dtype=torch.float32, device=dev)
        drs[:, 3:7] = torch.tensor(m.default_root_quat, dtype=torch.float32, device=dev)
        self.default_root_state = drs

    @property
    def joint_pos(self):
        return self._env._stepper.state.get("joint_pos")

    @property
    def joint_vel(self):
        return self._env._stepper.state.get("joint_vel")


class _SnakeRobot:
    def __init__(self, env):
        self.data = _SnakeArticulationData(env)
        self._ALL_INDICES = torch.arange(env.num_envs, dtype=torch.long, device=env.device)
        self.body_names = list(S.LINK_NAMES)
        self.joint_names = list(S.JOINT_NAMES)

    def find_bodies(self, pattern):
        return Z.find_bodies(pattern, S.LINK_NAMES)

    def find_joints(self, pattern):
        return Z.find_bodies(pattern, S.JOINT_NAMES)


class ZbotDirectEnvV0(ZbotDirectEnvV2):
    cfg: ZbotDirectEnvCfgV0

    _TASK = native.TASK_SNAKE_V0
    _TERM_IDS = native.SNAKE_TERM_IDS
    _HOST_TERMS_SUPPORTED = False
    _DIED_LOG_KEY = "Episode_Termination/died"      # snake_v0.py:289

    def __init__(self, cfg: ZbotDirectEnvCfgV0 | None = None, render_mode: str | None = None, **kwargs):
        super().__init__(cfg if cfg is not None else ZbotDirectEnvCfgV0(), render_mode, **kwargs)
        # snake_v0.py:121 -- one torch.rand(num_envs, 1) call on the env device, after the base-class setup
        self.joint_speed_limit = (torch.rand(self.num_envs, 1, device=self.device) * 1.8 + 0.2) * math.pi

    def _setup_scene(self):
        self._robot = _SnakeRobot(self)
        # snake_v0.py:118-119
        self.heading_vec = torch.tensor([0, -1, 0], dtype=torch.float32, device=self.device).repeat((self.num_envs, 1))
        self.up_vec = torch.tensor([-1, 0, 0], dtype=torch.float32, device=self.device).repeat((self.num_envs, 1))

    def __getattr__(self, name):
        fields = {"base_heading_y_sum": "base_heading_x_sum", "base_pos_x_err_sum": "base_pos_y_err_sum"}
        if name in fields and "_stepper" in self.__dict__:
            return self._stepper.state.get(fields[name])[:, 0]
        raise AttributeError(name)
