"""Static task parameters of ``zbot-6s-snake-v0`` -- mirror of ``ZbotDirectEnvCfgV0``
(``/root/reference/source/zbot/zbot/tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py:19-99``) and of its agent
cfg ``PPORunnerCfgV1`` (``.../zbot6_direct/agents/rsl_rl_ppo_cfg.py:38-63``)."""
from __future__ import annotations

from ...utils.configclass import Cfg
from ..zbot6b_direct.walking_v2_cfg import (ContactModelCfg, InteractiveSceneCfg, RslRlPpoActorCriticCfg,
                                            RslRlPpoAlgorithmCfg, SimulationCfg)

#: snake_v0.py:88-98; dict ORDER = evaluation order
REWARD_SCALES_SNAKE_V0 = {
    "base_vel_forward": 5.0,
    "base_up_z": -0.5,
    "base_heading_y": -1.0,
    "base_heading_y_sum": -1.0,
    "base_pos_x_err": -1.0,
    "action_rate": -0.1,
    "torques": -0.002,
}


class SnakeActuatorCfg(Cfg):
    """ImplicitActuatorCfg "zbot_six" of ZBOT_D_6S_CFG (assets/zbot_cfg.py:157-166)."""
    stiffness = 20.0
    damping = 0.5
    effort_limit = 20.0
    velocity_limit = 10.0   # ignored for implicit actuators


class ZbotDirectEnvCfgV0(Cfg):
    episode_length_s = 16.0          # snake_v0.py:50
    decimation = 4
    action_space = 6
    observation_space = 23
    state_space = 0
    sim = SimulationCfg()
    scene = InteractiveSceneCfg()
    actuator = SnakeActuatorCfg()
    contact = ContactModelCfg()
    reward_cfg = {"reward_scales": dict(REWARD_SCALES_SNAKE_V0)}
    seed = None
    log_dir = None
    is_finite_horizon = False
    # additive uniform observation noise {term: (n_min, n_max)} over base_quat / joint_pos / joint_vel / actions / extra
    # (ObservationManager `Unoise` semantics of the manager-based task, zbotlab_manager/zbotlab_env_cfg.py PolicyCfg:
    # base_quat +-0.01, joint_pos +-0.01, joint_vel +-1.5); None = the direct tasks' behaviour (no corruption)
    observation_noise = None
    check_all_envs_reset = None
    output_ring = 4


class PPORunnerCfgV1(Cfg):
    """zbot6_direct/agents/rsl_rl_ppo_cfg.py:38-63"""
    class_name = "OnPolicyRunner"
    seed = 42
    device = "cuda:0"
    num_steps_per_env = 16
    max_iterations = 1000
    save_interval = 100
    experiment_name = "zbot_6s_flat_snake_v1"
    run_name = ""
    empirical_normalization = False
    clip_actions = None
    resume = False
    load_run = ".*"
    load_checkpoint = "model_.*.pt"
    logger = "tensorboard"
    policy = RslRlPpoActorCriticCfg(actor_hidden_dims=[256, 256, 128], critic_hidden_dims=[256, 256, 128])
    algorithm = RslRlPpoAlgorithmCfg()
