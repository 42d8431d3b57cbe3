"""Cfg tree of ``zbot-6b-walking-m-v0`` -- mirrors what ``ZbotLabRoughEnvCfg``
(``/root/reference/source/zbot/zbot/tasks/zbotlab_manager/zbotlab_env_cfg.py:414-452``) carries after the robot
overrides of ``Zbot6BRoughEnvCfg`` (``config/zbot6b_manager/rough_env_cfg.py:22-56``) and the flat overrides of
``Zbot6BFlatEnvCfg`` (``config/zbot6b_manager/flat_env_cfg.py:11-112``): same section / term names, weights and params, so
code that edits ``cfg.rewards.<term>.weight``, sets a term to ``None`` or swaps ``params`` works as on the reference."""
from __future__ import annotations

import math

from ...utils.configclass import Cfg
from ..zbot6b_direct.walking_v2_cfg import (ContactModelCfg, InteractiveSceneCfg, RslRlPpoActorCriticCfg,
                                            RslRlPpoAlgorithmCfg, SimulationCfg)
from . import mdp
from .mdp import CurriculumTermCfg as CurrTerm
from .mdp import EventTermCfg as EventTerm
from .mdp import ObservationTermCfg as ObsTerm
from .mdp import RewardTermCfg as RewTerm
from .mdp import SceneEntityCfg
from .mdp import TerminationTermCfg as DoneTerm

FEET = {"asset_cfg": SceneEntityCfg("robot", body_names="foot.*")}
FEET_SENSOR = {"sensor_cfg": SceneEntityCfg("contact_forces", body_names="foot.*")}


class TerrainCfg(Cfg):
    terrain_type = "plane"             # flat_env_cfg.py:107-108; the rough cfg: "generator" (zbotlab_env_cfg.py:44-48)
    terrain_generator = None           # zbot_lab_b200.terrain.TerrainGeneratorCfg (restated ROUGH_TERRAINS_CFG [IL-upstream])
    max_init_terrain_level = 5
    static_friction = 1.0              # zbotlab_env_cfg.py:50-55, combine mode "multiply"
    dynamic_friction = 1.0


class ContactSensorCfg(Cfg):
    prim_path = "{ENV_REGEX_NS}/Robot/.*"
    history_length = 3                 # zbotlab_env_cfg.py:68
    track_air_time = True
    update_period = 0.005


class RobotCfg(Cfg):
    """ZBOT_6S_V2_CFG (assets/zbot_cfg.py:959-1005) on zbot_6s_v09.usd -> assets/zbot_6s_v2.py."""
    usd = "zbot_6s_v09.usd"
    stiffness = 20.0
    damping = 0.5
    effort_limit = 20.0
    velocity_limit = 10.0              # not applied by implicit actuators [IL-upstream]


class MySceneCfg(InteractiveSceneCfg):
    num_envs = 4096
    env_spacing = 2.5                  # zbotlab_env_cfg.py:421
    terrain = TerrainCfg()
    robot = RobotCfg()
    contact_forces = ContactSensorCfg()


class CommandsCfg(Cfg):
    base_velocity = mdp.UniformLevelVelocityCommandCfg(           # zbotlab_env_cfg.py:99-117
        resampling_time_range=(10.0, 10.0), rel_standing_envs=0.02, rel_heading_envs=1.0, heading_command=False,
        debug_vis=True,
        ranges=mdp.UniformVelocityCommandCfg.Ranges(lin_vel_x=(-0.1, 0.1), lin_vel_y=(0.0, 0.0), ang_vel_z=(0.0, 0.0)),
        limit_ranges=mdp.UniformVelocityCommandCfg.Ranges(lin_vel_x=(-0.3, 0.3), lin_vel_y=(0.0, 0.0), ang_vel_z=(0.0, 0.0)))


class ActionsCfg(Cfg):
    joint_pos = mdp.RelativeJointPositionActionCfg(               # :124-130
        clip={"joint.*": [-0.04 * math.pi, 0.04 * math.pi]}, joint_names=["joint.*"], scale=0.04 * math.pi,
        use_zero_offset=True)


class PolicyCfg(mdp.ObservationGroupCfg):                          # :137-160, order preserved
    base_quat = ObsTerm(func=mdp.root_quat_w, noise=mdp.AdditiveUniformNoiseCfg(n_min=-0.01, n_max=0.01))
    velocity_commands = ObsTerm(func=mdp.generated_commands, params={"command_name": "base_velocity"})
    joint_pos = ObsTerm(func=mdp.joint_pos_rel, noise=mdp.AdditiveUniformNoiseCfg(n_min=-0.01, n_max=0.01))
    joint_vel = ObsTerm(func=mdp.joint_vel_rel, noise=mdp.AdditiveUniformNoiseCfg(n_min=-1.5, n_max=1.5))
    actions = ObsTerm(func=mdp.last_action)
    enable_corruption = True
    concatenate_terms = True


class ObservationsCfg(Cfg):
    policy = PolicyCfg()


class EventCfg(Cfg):                                               # :166-236 with rough_env_cfg.py:38-42 applied
    init_my_data = EventTerm(func=mdp.init_my_data, mode="startup")
    physics_material = EventTerm(func=mdp.randomize_rigid_body_material, mode="startup", params={
        "asset_cfg": SceneEntityCfg("robot", body_names=".*"), "static_friction_range": (0.3, 1.0),
        "dynamic_friction_range": (0.3, 1.0), "restitution_range": (0.0, 0.0), "num_buckets": 64})
    add_base_mass = None
    base_com = None
    reset_base = EventTerm(func=mdp.reset_root_state_uniform, mode="reset", params={
        "pose_range": {"x": (-0.5, 0.5), "y": (-0.5, 0.5), "yaw": (-3.14, 3.14)},
        "velocity_range": {k: (0.0, 0.0) for k in ("x", "y", "z", "roll", "pitch", "yaw")}})
    reset_robot_joints = EventTerm(func=mdp.reset_joints_by_scale, mode="reset",
                                   params={"position_range": (1.0, 1.0), "velocity_range": (1.0, 1.0)})
    reset_my_data = EventTerm(func=mdp.reset_my_data, mode="reset", params=dict(FEET))
    push_robot = None


class RewardsCfg(Cfg):                                             # :240-352 with flat_env_cfg.py:96-112 applied
    track_lin_vel_xy_exp = RewTerm(func=mdp.track_lin_vel_xy_yaw_frame_exp, weight=1.0,
                                   params={"command_name": "base_velocity", "std": math.sqrt(0.25)})
    track_ang_vel_z_exp = RewTerm(func=mdp.track_ang_vel_z_world_exp, weight=0.5,
                                  params={"command_name": "base_velocity", "std": math.sqrt(0.25)})
    termination_penalty = RewTerm(func=mdp.is_terminated, weight=-200.0)
    dof_torques_l2 = RewTerm(func=mdp.joint_torques_l2, weight=-1.0e-5)
    dof_acc_l2 = RewTerm(func=mdp.joint_acc_l2, weight=-2.5e-7)
    action_rate_l2 = RewTerm(func=mdp.action_rate_l2, weight=-0.01)
    foot_step_length = RewTerm(func=mdp.foot_step_length, weight=5.0, params={**FEET, **FEET_SENSOR, "command_name": None})
    foot_downward = RewTerm(func=mdp.foot_downward, weight=-1.0, params=dict(FEET))
    foot_forward = RewTerm(func=mdp.foot_forward, weight=-0.5, params=dict(FEET))
    gait = None
    feet_slide = RewTerm(func=mdp.feet_slide, weight=-6.5, params={**FEET, **FEET_SENSOR})
    feet_clearance = None
    feet_air_time = None
    air_time_variance = RewTerm(func=mdp.air_time_balance_penalty, weight=-15.0, params=dict(FEET_SENSOR))
    base_vel_forward = None
    feet_force_pattern = None
    undesired_contacts = None


class TerminationsCfg(Cfg):                                        # :371-393 with rough_env_cfg.py:46
    time_out = DoneTerm(func=mdp.time_out, time_out=True)
    base_contact = None
    base_height = DoneTerm(func=mdp.root_height_below_minimum, params={"minimum_height": 0.2})
    feet_close = DoneTerm(func=mdp.feet_close, params={"minimum_distance": 0.12, **FEET})


class CurriculumCfg(Cfg):                                          # :396-401 with flat_env_cfg.py:110
    terrain_levels = None
    lin_vel_cmd_levels = CurrTerm(func=mdp.lin_vel_cmd_levels)


class ManagerBasedRLEnvCfg(Cfg):
    decimation = 4
    episode_length_s = 20.0
    sim = SimulationCfg()
    scene = MySceneCfg()
    contact = ContactModelCfg()
    seed = None
    log_dir = None
    is_finite_horizon = False
    output_ring = 4


class Zbot6BFlatEnvCfg(ManagerBasedRLEnvCfg):
    observations = ObservationsCfg()
    actions = ActionsCfg()
    commands = CommandsCfg()
    rewards = RewardsCfg()
    terminations = TerminationsCfg()
    events = EventCfg()
    curriculum = CurriculumCfg()


class Zbot6BFlatEnvCfg_PLAY(Zbot6BFlatEnvCfg):
    """flat_env_cfg.py:115-126: 64 envs, no observation corruption, commands drawn from the limit ranges."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.scene.num_envs = 64
        self.scene.env_spacing = 2.5
        self.observations.policy.enable_corruption = False
        self.commands.base_velocity.ranges = self.commands.base_velocity.limit_ranges


# ---------------------------------------------------------------------------------------------------------------------
# zbot-6b-walking-m-rough-v0: ZbotLabRoughEnvCfg (zbotlab_env_cfg.py:414-452) + Zbot6BRoughEnvCfg (rough_env_cfg.py:22-56):
# generated terrain + terrain_levels curriculum, ALL RewTerms of RewardsCfg at their base weights (:240-371), the robot
# overrides of the rough cfg (add_base_mass / base_com / push_robot / base_contact = None)
# ---------------------------------------------------------------------------------------------------------------------
class RoughTerrainCfg(TerrainCfg):
    terrain_type = "generator"

    def __init__(self, **kw):
        from ...terrain import rough_terrains_cfg
        super().__init__(**kw)
        if self.terrain_generator is None:
            self.terrain_generator = rough_terrains_cfg()


class RoughSceneCfg(MySceneCfg):
    terrain = RoughTerrainCfg()


class RoughRewardsCfg(RewardsCfg):                                 # zbotlab_env_cfg.py:240-371, nothing switched off
    foot_step_length = RewTerm(func=mdp.foot_step_length, weight=2.0, params={**FEET, **FEET_SENSOR, "command_name": None})
    gait = RewTerm(func=mdp.feet_gait, weight=0.5, params={"period": 2.0, "offset": [0.0, 0.5], "threshold": 0.55,
                                                            "command_name": "base_velocity", **FEET_SENSOR})
    feet_slide = RewTerm(func=mdp.feet_slide, weight=-0.2, params={**FEET, **FEET_SENSOR})
    feet_clearance = RewTerm(func=mdp.foot_clearance_reward, weight=1.0,
                             params={"std": 0.05, "tanh_mult": 2.0, "target_height": 0.01, **FEET})
    feet_air_time = RewTerm(func=mdp.feet_air_time_positive_biped, weight=2.5,
                            params={**FEET_SENSOR, "command_name": "base_velocity", "threshold": 0.3})
    air_time_variance = RewTerm(func=mdp.air_time_balance_penalty, weight=-1.0, params=dict(FEET_SENSOR))
    base_vel_forward = RewTerm(func=mdp.base_vel_forward, weight=1.0, params={"which_forward": 1})
    feet_force_pattern = RewTerm(func=mdp.feet_force_pattern, weight=1.0, params=dict(FEET_SENSOR))
    undesired_contacts = RewTerm(func=mdp.undesired_contacts, weight=-1.0, params={
        "sensor_cfg": SceneEntityCfg("contact_forces", body_names="base|a.*|b.*"), "threshold": 1.0})


class RoughCurriculumCfg(Cfg):                                     # :396-401
    terrain_levels = CurrTerm(func=mdp.terrain_levels_vel)
    lin_vel_cmd_levels = CurrTerm(func=mdp.lin_vel_cmd_levels)


class Zbot6BRoughEnvCfg(Zbot6BFlatEnvCfg):
    scene = RoughSceneCfg()
    rewards = RoughRewardsCfg()
    curriculum = RoughCurriculumCfg()


class Zbot6BRoughEnvCfg_PLAY(Zbot6BRoughEnvCfg):
    """rough_env_cfg.py:59-83: 64 envs, 5 x 5 tiles without the level curriculum, robots spawned on random levels,
    forward commands, no observation corruption."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.scene.num_envs = 64
        self.scene.env_spacing = 2.5
        self.scene.terrain.max_init_terrain_level = None
        g = self.scene.terrain.terrain_generator
        g.num_rows, g.num_cols, g.curriculum = 5, 5, False
        self.commands.base_velocity.ranges.lin_vel_x = (0.7, 1.0)
        self.commands.base_velocity.ranges.lin_vel_y = (0.0, 0.0)
        self.commands.base_velocity.ranges.heading = (0.0, 0.0)
        self.observations.policy.enable_corruption = False


class Zbot6BFlatPPORunnerCfg(Cfg):
    """config/zbot6b_manager/agents/rsl_rl_ppo_cfg.py:11-50"""
    class_name = "OnPolicyRunner"
    seed = 42
    device = "cuda:0"
    num_steps_per_env = 24
    max_iterations = 1000
    save_interval = 100
    experiment_name = "zbot_6b_flat_mana_v1"
    run_name = ""
    empirical_normalization = False
    clip_actions = None
    resume = False
    load_run = ".*"
    load_checkpoint = "model_.*.pt"
    logger = "tensorboard"
    policy = RslRlPpoActorCriticCfg(actor_hidden_dims=[128, 128, 128], critic_hidden_dims=[128, 128, 128])
    algorithm = RslRlPpoAlgorithmCfg(entropy_coef=0.01)


class Zbot6BRoughPPORunnerCfg(Zbot6BFlatPPORunnerCfg):
    """config/zbot6b_manager/agents/rsl_rl_ppo_cfg.py:11-38"""
    max_iterations = 1500
    experiment_name = "zbot_6b_rough_mana_v1"
    policy = RslRlPpoActorCriticCfg(actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
