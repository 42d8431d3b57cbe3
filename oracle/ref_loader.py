"""TEST INFRASTRUCTURE -- loads the reference's own MDP code behind stub modules.

Only ``tests/`` (and ``tests/golden/make_golden.py``) may import this.  It needs
``/root/reference`` and therefore works only in the build container; nothing that
runs on the GPU box imports it.  The GPU box sees the committed fixtures under
``tests/golden/`` that this loader produced.

What it does (recipe from SURVEY.md Appendix D): registers stand-ins for
``gymnasium``, ``isaaclab.*`` and ``zbot.assets`` in ``sys.modules`` and then executes
``source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py`` *unmodified*
via ``importlib``.  An instance is built with ``object.__new__`` (the real ``__init__``
needs a live PhysX scene) and given plain CPU tensors for ``_robot.data``,
``_contact_sensor.data`` and ``_terrain.env_origins``; the reference's unbound methods
``_pre_physics_step / _get_dones / _get_rewards / _reset_idx / _get_observations``
(``...env_v2.py:276-459``) are then called as-is.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("ZBOT_REFERENCE_ROOT", "/root/reference")
REF_ENV_V2 = os.path.join(
    REF_ROOT, "source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py"
)
REF_ENV_SNAKE = os.path.join(REF_ROOT, "source/zbot/zbot/tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py")
REF_ENV_V4 = os.path.join(REF_ROOT, "source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py")
REF_M_MDP_DIR = os.path.join(REF_ROOT, "source/zbot/zbot/tasks/zbotlab_manager/mdp")      # rewards.py, terminations.py, curriculums.py


def reference_available() -> bool:
    return os.path.isfile(REF_ENV_V2)


def _quat_apply(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    # isaaclab.utils.math.quat_apply: wxyz, t = 2 q_xyz x v; v + w t + q_xyz x t
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec + quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


# ---- isaaclab.utils.math functions the v4 task calls (restated from Isaac Lab; [IL-upstream]) ----
def _quat_from_euler_xyz(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                        sy * cr * cp - cy * sr * sp], dim=-1)


def _quat_mul(q1, q2):
    shape = q1.shape
    q1, q2 = q1.reshape(-1, 4), q2.reshape(-1, 4)
    w1, x1, y1, z1 = q1[:, 0], q1[:, 1], q1[:, 2], q1[:, 3]
    w2, x2, y2, z2 = q2[:, 0], q2[:, 1], q2[:, 2], q2[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([w, x, y, z], dim=-1).view(shape)


def _wrap_to_pi(angles):
    wrapped = (angles + torch.pi) % (2 * torch.pi)
    return torch.where((wrapped == 0) & (angles > 0), torch.pi, wrapped - torch.pi)


def _quat_apply_inverse(quat, vec):
    # isaaclab.utils.math.quat_apply_inverse [IL-upstream]: v - w t + q_xyz x t, t = 2 q_xyz x v
    shape = vec.shape
    quat = quat.reshape(-1, 4)
    vec = vec.reshape(-1, 3)
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return (vec - quat[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def _yaw_quat(quat):
    # isaaclab.utils.math.yaw_quat [IL-upstream]
    shape = quat.shape
    q = quat.view(-1, 4)
    qw, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    yaw = torch.atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
    out = torch.zeros_like(q)
    out[:, 3] = torch.sin(yaw / 2)
    out[:, 0] = torch.cos(yaw / 2)
    out = torch.nn.functional.normalize(out, p=2.0, dim=-1)
    return out.view(shape)


def _sample_uniform(lower, upper, size, device=None):
    if isinstance(size, int):
        size = (size,)
    return torch.rand(*size) * (upper - lower) + lower


class _Cfg:
    """kwargs-accepting dummy with ``replace`` (stands in for every ``*Cfg`` class)."""

    def __init__(self, *a, **kw):
        self.__dict__.update(kw)

    def replace(self, **kw):
        new = _Cfg(**self.__dict__)
        new.__dict__.update(kw)
        return new


class _DirectRLEnv:
    """Stand-in base: only the part of ``DirectRLEnv._reset_idx`` the subclass relies on
    (``episode_length_buf[ids] = 0`` after scene/event/noise resets)."""

    def _reset_idx(self, env_ids):
        hook = getattr(self, "_ref_reset_events", None)
        if hook is not None:           # EventManager.apply(mode="reset") of the v4 task (ref_harness.RefV4Harness)
            hook(env_ids)
        self.episode_length_buf[env_ids] = 0

    def set_debug_vis(self, flag):
        return False


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in (
        "gymnasium", "gymnasium.spaces", "isaaclab", "isaaclab.sim", "isaaclab.utils",
        "isaaclab.utils.math", "isaaclab.assets", "isaaclab.envs", "isaaclab.scene",
        "isaaclab.sensors", "isaaclab.terrains", "zbot", "zbot.assets", "isaaclab.envs.mdp", "isaaclab.managers",
        "isaaclab.markers", "isaaclab.markers.config")}
    spaces = mod("gymnasium.spaces", flatdim=lambda s: int(s))
    mod("gymnasium", spaces=spaces)
    sim = mod("isaaclab.sim", RigidBodyMaterialCfg=_Cfg, SimulationCfg=_Cfg, DomeLightCfg=_Cfg)
    umath = mod("isaaclab.utils.math", quat_apply=_quat_apply, quat_from_euler_xyz=_quat_from_euler_xyz,
                quat_mul=_quat_mul, wrap_to_pi=_wrap_to_pi, sample_uniform=_sample_uniform,
                quat_apply_inverse=_quat_apply_inverse, yaw_quat=_yaw_quat)
    utils = mod("isaaclab.utils", configclass=lambda c: c, math=umath)
    assets = mod("isaaclab.assets", Articulation=object, ArticulationCfg=_Cfg, RigidObject=object)
    envs_mdp = mod("isaaclab.envs.mdp")
    envs = mod("isaaclab.envs", DirectRLEnv=_DirectRLEnv, DirectRLEnvCfg=object, mdp=envs_mdp)
    managers = mod("isaaclab.managers", EventTermCfg=_Cfg, SceneEntityCfg=_Cfg)
    arrow = lambda: _Cfg(markers={"arrow": _Cfg(scale=(1.0, 1.0, 1.0))})
    markers_cfg = mod("isaaclab.markers.config", RED_ARROW_X_MARKER_CFG=arrow(), GREEN_ARROW_X_MARKER_CFG=arrow())
    markers = mod("isaaclab.markers", VisualizationMarkers=object, VisualizationMarkersCfg=_Cfg, config=markers_cfg)
    scene = mod("isaaclab.scene", InteractiveSceneCfg=_Cfg)
    sensors = mod("isaaclab.sensors", ContactSensor=object, ContactSensorCfg=_Cfg)
    terrains = mod("isaaclab.terrains", TerrainImporterCfg=_Cfg, TerrainImporter=object)
    mod("isaaclab", sim=sim, utils=utils, assets=assets, envs=envs, scene=scene,
        sensors=sensors, terrains=terrains, managers=managers, markers=markers)
    zassets = mod("zbot.assets", ZBOT_6S_CFG=_Cfg(), ZBOT_D_6S_CFG=_Cfg())
    mod("zbot", assets=zassets)
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


_REF_MODULES = {}


def load_reference_module(path: str = REF_ENV_V2):
    """Execute a reference task file unmodified behind the stubs; returns the module."""
    if path in _REF_MODULES:
        return _REF_MODULES[path]
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    saved = _install_stubs()
    try:
        name = "_zbot_ref_" + os.path.splitext(os.path.basename(path))[0]
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        _restore(saved)
    _REF_MODULES[path] = m
    return m


class _NS:
    pass


class _Recorder:
    """no-op articulation recording the writes ``_reset_idx`` makes (…env_v2.py:416,431-433)."""

    def __init__(self, n):
        self.data = _NS()
        self._ALL_INDICES = torch.arange(n, dtype=torch.long)
        self.calls = []

    def reset(self, env_ids):
        self.calls.append(("reset", env_ids.clone()))

    def write_root_pose_to_sim(self, pose, env_ids):
        self.calls.append(("root_pose", pose.clone(), env_ids.clone()))

    def write_root_velocity_to_sim(self, vel, env_ids):
        self.calls.append(("root_vel", vel.clone(), env_ids.clone()))

    def write_joint_state_to_sim(self, pos, vel, _ids, env_ids):
        self.calls.append(("joint_state", pos.clone(), vel.clone(), env_ids.clone()))


def make_reference_env(num_envs: int, *, feet_ids, undesired_ids, base_body_idx, feet_body_idx,
                       default_joint_pos, default_root_state, env_origins):
    """Build a ``ZbotDirectEnvV2`` (reference class) without its ``__init__``.

    Mirrors the attribute set of ``…env_v2.py:211-257`` with plain CPU tensors.
    ``reward_scales`` is a *copy* of the cfg dict scaled by step_dt once (SURVEY C-3).
    """
    ref = load_reference_module()
    n = num_envs
    env = object.__new__(ref.ZbotDirectEnvV2)
    env.cfg = _NS()
    env.cfg.termination_height = ref.ZbotDirectEnvCfgV2.termination_height
    env.cfg.reward_cfg = ref.ZbotDirectEnvCfgV2.reward_cfg
    env.num_envs = n
    env.device = torch.device("cpu")
    env.sim = _NS()
    env.sim.device = "cpu"
    env.step_dt = ref.ZbotDirectEnvCfgV2.decimation * (1 / 200.0)
    env.max_episode_length = 1000
    env.max_episode_length_s = ref.ZbotDirectEnvCfgV2.episode_length_s
    env.extras = {}

    env._robot = _Recorder(n)
    d = env._robot.data
    d.default_joint_pos = default_joint_pos.clone()
    d.default_joint_vel = torch.zeros(n, 6)
    d.default_root_state = default_root_state.clone()
    d.GRAVITY_VEC_W = torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1)
    env._contact_sensor = _NS()
    env._contact_sensor.data = _NS()
    env._terrain = _NS()
    env._terrain.env_origins = env_origins.clone()

    env._feet_ids = list(feet_ids)
    env._undesired_contact_body_ids = list(undesired_ids)
    env.base_body_idx = list(base_body_idx)
    env.feet_body_idx = list(feet_body_idx)

    env._actions = torch.zeros(n, 6)
    env._previous_actions = torch.zeros(n, 6)
    env.feet_contact_forces_last = torch.zeros(n, 2)
    env.feet_down_pos_last = torch.zeros(n, 2, 3)
    env.feet_step_length = torch.zeros(n, 2)
    env.feet_air_times = torch.zeros(n, 2)
    env.feet_force_sum = torch.zeros(n)
    env.base_heading_x_sum = torch.zeros(n)
    env.base_pos_y_err_sum = torch.zeros(n)
    env.joint_speed_limit = torch.ones(n, 1)
    env.p_delta = torch.zeros(n, 6)
    env.episode_length_buf = torch.zeros(n, dtype=torch.long)
    env.reset_terminated = torch.zeros(n, dtype=torch.bool)
    env.reset_time_outs = torch.zeros(n, dtype=torch.bool)

    scales = dict(ref.ZbotDirectEnvCfgV2.reward_cfg["reward_scales"])
    env.reward_scales = {k: v * env.step_dt for k, v in scales.items()}
    env.reward_functions = {k: getattr(env, "_reward_" + k) for k in env.reward_scales}
    env._episode_sums = {k: torch.zeros(n) for k in env.reward_scales}
    return env


def reference_reward_scales() -> dict:
    """Unscaled term->weight dict, in the reference's insertion order (…env_v2.py:190-206)."""
    return dict(load_reference_module().ZbotDirectEnvCfgV2.reward_cfg["reward_scales"])


def make_reference_snake_env(num_envs: int, *, default_root_state, env_origins, joint_speed_limit):
    """Build the reference ``ZbotDirectEnvV0`` (snake task, zbot_direct_6dof_snake_v0.py:102-146) without its
    ``__init__``; attribute set mirrors that ``__init__`` with plain CPU tensors."""
    ref = load_reference_module(REF_ENV_SNAKE)
    n = num_envs
    cfgc = ref.ZbotDirectEnvCfgV0
    env = object.__new__(ref.ZbotDirectEnvV0)
    env.cfg = _NS()
    env.cfg.reward_cfg = cfgc.reward_cfg
    env.num_envs = n
    env.device = torch.device("cpu")
    env.sim = _NS()
    env.sim.device = "cpu"
    env.step_dt = cfgc.decimation * (1 / 200.0)
    env.max_episode_length_s = cfgc.episode_length_s
    env.max_episode_length = int(round(cfgc.episode_length_s / env.step_dt))
    env.extras = {}
    env._robot = _Recorder(n)
    d = env._robot.data
    d.default_joint_pos = torch.zeros(n, 6)
    d.default_joint_vel = torch.zeros(n, 6)
    d.default_root_state = default_root_state.clone()
    for i in (1, 2, 3, 4):
        sensor = _NS()
        sensor.data = _NS()
        setattr(env, f"_contact_sensor_{i}", sensor)
    env._terrain = _NS()
    env._terrain.env_origins = env_origins.clone()
    env._actions = torch.zeros(n, 6)
    env._previous_actions = torch.zeros(n, 6)
    env.heading_vec = torch.tensor([0, -1, 0], dtype=torch.float32).repeat((n, 1))
    env.up_vec = torch.tensor([-1, 0, 0], dtype=torch.float32).repeat((n, 1))
    env.base_heading_y_sum = torch.zeros(n)
    env.base_pos_x_err_sum = torch.zeros(n)
    env.joint_speed_limit = joint_speed_limit.clone().reshape(n, 1)
    env.p_delta = torch.zeros(n, 6)
    env.episode_length_buf = torch.zeros(n, dtype=torch.long)
    env.reset_terminated = torch.zeros(n, dtype=torch.bool)
    env.reset_time_outs = torch.zeros(n, dtype=torch.bool)
    scales = dict(cfgc.reward_cfg["reward_scales"])
    env.reward_scales = {k: v * env.step_dt for k, v in scales.items()}
    env.reward_functions = {k: getattr(env, "_reward_" + k) for k in env.reward_scales}
    env._episode_sums = {k: torch.zeros(n) for k in env.reward_scales}
    return env


def reference_snake_reward_scales() -> dict:
    return dict(load_reference_module(REF_ENV_SNAKE).ZbotDirectEnvCfgV0.reward_cfg["reward_scales"])


class _EventManager:
    """``env.event_manager.get_term_cfg(name)`` over the reference's own ``EventCfg`` instance."""

    def __init__(self, events_cfg):
        self._cfg = events_cfg

    def get_term_cfg(self, name):
        return getattr(self._cfg, name)


def make_reference_v4_env(num_envs: int, *, feet_ids, undesired_ids, base_body_idx, feet_body_idx, default_joint_pos,
                          default_root_state, env_origins):
    """Build the reference ``Zbot6SEnvV4`` (…env_v4.py:560-651) without its ``__init__``; attribute set mirrors that
    ``__init__`` with plain CPU tensors.  ``events`` is a fresh ``EventCfg()`` of the reference's own class."""
    from collections import deque
    ref = load_reference_module(REF_ENV_V4)
    n = num_envs
    cfgc = ref.Zbot6SEnvV4Cfg
    env = object.__new__(ref.Zbot6SEnvV4)
    env.cfg = _NS()
    env.cfg.termination_height = cfgc.termination_height
    env.cfg.reward_cfg = cfgc.reward_cfg
    env.cfg.events = ref.EventCfg()
    env.cfg.events.vel_range.params["limit_yaw_ranges"] = (-0.5, 0.5)        # …env_v4.py:558
    env.cfg.debug_vis = False
    env.event_manager = _EventManager(env.cfg.events)
    env.num_envs = n
    env.device = torch.device("cpu")
    env.sim = _NS()
    env.sim.device = "cpu"
    env.step_dt = cfgc.decimation * (1 / 200.0)
    env.max_episode_length_s = cfgc.episode_length_s
    env.max_episode_length = 1000
    env.common_step_counter = 0
    env.extras = {}
    env.scene = _NS()
    env.scene.env_origins = env_origins.clone()
    env._robot = _Recorder(n)
    d = env._robot.data
    d.default_joint_pos = default_joint_pos.clone()
    d.default_joint_vel = torch.zeros(n, 6)
    d.default_root_state = default_root_state.clone()
    d.GRAVITY_VEC_W = torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1)
    env._contact_sensor = _NS()
    env._contact_sensor.data = _NS()
    env._terrain = _NS()
    env._terrain.env_origins = env_origins.clone()
    env._feet_ids = list(feet_ids)
    env._undesired_contact_body_ids = list(undesired_ids)
    env.base_body_idx = list(base_body_idx)
    env.feet_body_idx = list(feet_body_idx)
    # …env_v4.py:563-651
    env.base_lin_vel_forward_w = torch.zeros(n)
    env.commands = torch.zeros(n, 2)
    env.current_yaw = torch.zeros(n)
    env.target_heading_yaw = torch.zeros(n)
    env.curriculum_stage = 0
    env.curriculum_vel_reward_buffer = deque(maxlen=24)
    env.curriculum_yaw_reward_buffer = deque(maxlen=24)
    env._actions = torch.zeros(n, 6)
    env._previous_actions = torch.zeros(n, 6)
    env.p_delta = torch.zeros(n, 6)
    env.joint_speed_limit = torch.ones(n, 1)
    env.z_w = torch.tensor([0, 0, 1], dtype=torch.float32).repeat((n, 2, 1))
    env.axis_x_feet = torch.tensor([1, 0, 0], dtype=torch.float32).repeat((n, 2, 1))
    env.axis_z_feet = torch.tensor([[0, 0, 1], [0, 0, -1]], dtype=torch.float32).repeat((n, 1, 1))
    env.feet_contact_forces_last = 15.0 * torch.ones(n, 2)
    env.feet_down_pos_last = torch.zeros(n, 2, 3)
    env.feet_step_length = torch.zeros(n, 2)
    env.feet_force_sum = torch.zeros(n)
    env.heading_err_sum = torch.zeros(n)
    env.episode_length_buf = torch.zeros(n, dtype=torch.long)
    env.reset_terminated = torch.zeros(n, dtype=torch.bool)
    env.reset_time_outs = torch.zeros(n, dtype=torch.bool)
    env.reward_scales = dict(cfgc.reward_cfg["reward_scales"])
    env.reward_functions = {k: getattr(env, "_reward_" + k) for k in env.reward_scales}
    env._episode_sums = {k: torch.zeros(n) for k in env.reward_scales}
    return env


def reference_v4_reward_scales() -> dict:
    return dict(load_reference_module(REF_ENV_V4).Zbot6SEnvV4Cfg.reward_cfg["reward_scales"])
