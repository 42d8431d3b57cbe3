"""Static task parameters of ``zbot-6b-walking-v4`` -- mirror of ``Zbot6SEnvV4Cfg`` / ``EventCfg``
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py:267-558``) and of its
agent cfg ``Zbot6SEnvV4PPOCfg`` (``.../zbot6b_direct/agents/rsl_rl_ppo_cfg.py:205-233``)."""
from __future__ import annotations

from ...utils.configclass import Cfg
from .walking_v2_cfg import (ActuatorCfg, ContactModelCfg, InteractiveSceneCfg, RslRlPpoActorCriticCfg,
                             RslRlPpoAlgorithmCfg, SimulationCfg)

#: …env_v4.py:522-556; dict ORDER = evaluation order.  Bare weights: v4 multiplies by step_dt when it evaluates (:857)
REWARD_SCALES_V4 = {
    "track_lin_vel_x": 1.0,
    "track_heading_yaw": 1.0,
    "lin_vel_y": -1.0,
    "action_rate": -0.1,
    "torques": -2e-4,
    "joint_vel": -0.001,
    "joint_acc": -2.5e-7,
    "feet_downward": -1.0,
    "feet_forward": -0.5,
    "step_length": 5.0,
    "feet_air_time_biped": 1.0,
    "airtime_variance": -5.0,
    "feet_slide": -1.0,
    "feet_harmony": 0.0,
    "feet_close": -10.0,
}


class ResetBaseCfg(Cfg):
    """EventTerm ``reset_base`` (func reset_root_state_uniform, mode "reset"; …env_v4.py:331-347)."""
    pose_range = {"x": (-0.5, 0.5), "y": (-0.5, 0.5), "yaw": (-3.14, 3.14)}


class CommandResampleCfg(Cfg):
    """EventTerms ``reset_command_resample`` (mode "reset") and ``interval_command_resample`` (mode "interval",
    interval_range_s (3, 6)); …env_v4.py:392-418.  The reference's curricula always write both terms."""
    velocity_range = (0.3, 0.3)
    yaw_range = (-0.1, 0.1)
    dual_sign = True
    offset = 0.0
    prob_pos = 1.0
    interval_range_s = (3.0, 6.0)


class RangeCurriculumCfg(Cfg):
    """EventTerm ``vel_range`` (func range_curriculum; …env_v4.py:381-388, limit_yaw_ranges overridden at :558)."""
    limit_ranges = (0.0, 0.3)
    limit_yaw_ranges = (-0.5, 0.5)


class EventCfg(Cfg):
    reset_base = ResetBaseCfg()
    my_curric = True                 # my_curriculum (…env_v4.py:138-198)
    vel_range = RangeCurriculumCfg()
    command_resample = CommandResampleCfg()


class Zbot6SEnvV4Cfg(Cfg):
    episode_length_s = 20.0          # …env_v4.py:448
    decimation = 4
    action_space = 6
    observation_space = 24
    state_space = 0
    termination_height = 0.20        # :520
    sim = SimulationCfg()
    scene = InteractiveSceneCfg()
    actuator = ActuatorCfg()
    contact = ContactModelCfg()
    events = EventCfg()
    reward_cfg = {"reward_scales": dict(REWARD_SCALES_V4)}
    debug_vis = False
    seed = None
    log_dir = None
    is_finite_horizon = False
    # additive uniform observation noise {term: (n_min, n_max)} over base_quat / joint_pos / joint_vel / actions / extra
    # (ObservationManager `Unoise` semantics of the manager-based task, zbotlab_manager/zbotlab_env_cfg.py PolicyCfg:
    # base_quat +-0.01, joint_pos +-0.01, joint_vel +-1.5); None = the direct tasks' behaviour (no corruption)
    observation_noise = None
    check_all_envs_reset = None
    output_ring = 4


class Zbot6SEnvV4PPOCfg(Cfg):
    """agents/rsl_rl_ppo_cfg.py:205-233"""
    class_name = "OnPolicyRunner"
    seed = 42
    device = "cuda:0"
    num_steps_per_env = 24
    max_iterations = 2000
    save_interval = 1000
    experiment_name = "zbot_6b_flat_direct_v4"
    run_name = ""
    empirical_normalization = False
    clip_actions = None
    resume = False
    load_run = ".*"
    load_checkpoint = "model_.*.pt"
    logger = "tensorboard"
    policy = RslRlPpoActorCriticCfg(actor_hidden_dims=[256, 256, 128], critic_hidden_dims=[256, 256, 128])
    algorithm = RslRlPpoAlgorithmCfg()
