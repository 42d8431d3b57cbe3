"""Static task parameters of ``zbot-6b-walking-v2`` -- mirror of ``ZbotDirectEnvCfgV2``
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py:26-206``)
and of the agent cfg ``PPORunnerCfgV2`` (``.../agents/rsl_rl_ppo_cfg.py:65-91``)."""
from __future__ import annotations

from ...utils.configclass import Cfg

#: "train reward 2000 step4" (…env_v2.py:190-206); dict ORDER = evaluation order (SURVEY C-4)
REWARD_SCALES_V2 = {
    "base_vel_forward": 1.0,
    "feet_downward": -2.0,
    "feet_forward": -1.0,
    "base_heading_x": -1.0,
    "base_heading_x_sum": -5.0,
    "step_length": 5.0,
    "airtime_balance": -15.0,
    "action_rate": -0.1,
    "torques": -0.002,
    "feet_slide": -10.0,
    "base_pos_y_err": -2.0,
    "base_pos_y_err_sum": -2.0,
    "airtime_sum": 3.0,
}


class SimulationCfg(Cfg):
    dt = 1 / 200.0          # …env_v2.py:48
    render_interval = 4
    device = "cuda:0"
    gravity = (0.0, 0.0, -9.81)


class InteractiveSceneCfg(Cfg):
    num_envs = 4096         # …env_v2.py:73-75
    env_spacing = 4.0
    replicate_physics = True


class ContactModelCfg(Cfg):
    """Ground-contact law of the B200 step (ours; PhysX's is closed -- DESIGN.md §3)."""
    alpha = 1000.0
    erp = 0.2
    max_depenetration_velocity = 1.0   # assets/zbot_cfg.py:633
    beta_max = 3000.0
    friction = 1.0                     # 1.0 x 1.0, "multiply" combine (…env_v2.py:50-68)
    ramp = 5.0e-4
    margin = 0.02


class ActuatorCfg(Cfg):
    """ImplicitActuatorCfg "zbot_six" (assets/zbot_cfg.py:658-668)."""
    stiffness = 50.0
    damping = 5.0
    effort_limit = 20.0
    velocity_limit = 20.0   # ignored for implicit actuators


class ZbotDirectEnvCfgV2(Cfg):
    # env (…env_v2.py:38-44)
    episode_length_s = 20.0
    decimation = 4
    action_space = 6
    observation_space = 23
    state_space = 0
    termination_height = 0.22
    # simulation / scene
    sim = SimulationCfg()
    scene = InteractiveSceneCfg()
    actuator = ActuatorCfg()
    contact = ContactModelCfg()
    # rewards (…env_v2.py:190-206)
    reward_cfg = {"reward_scales": dict(REWARD_SCALES_V2)}
    # DirectRLEnvCfg fields the scripts touch (scripts/rsl_rl/train.py:114-155)
    seed = None
    log_dir = None
    is_finite_horizon = False
    # B200 step specifics
    # additive uniform observation noise {term: (n_min, n_max)} over base_quat / joint_pos / joint_vel / actions / extra
    # (ObservationManager `Unoise` semantics of the manager-based task, zbotlab_manager/zbotlab_env_cfg.py PolicyCfg:
    # base_quat +-0.01, joint_pos +-0.01, joint_vel +-1.5); None = the direct tasks' behaviour (no corruption)
    observation_noise = None
    check_all_envs_reset = None   # None: host sync-check when num_envs <= 256, device-side spread above; True / False: force host path / off (…env_v2.py:418-422)
    output_ring = 4               # step() outputs rotate through this many buffers


class RslRlPpoActorCriticCfg(Cfg):
    class_name = "ActorCritic"
    init_noise_std = 1.0
    actor_hidden_dims = [128, 128, 128]
    critic_hidden_dims = [128, 128, 128]
    activation = "elu"


class RslRlPpoAlgorithmCfg(Cfg):
    class_name = "PPO"
    value_loss_coef = 1.0
    use_clipped_value_loss = True
    clip_param = 0.2
    entropy_coef = 0.005
    num_learning_epochs = 5
    num_mini_batches = 4
    learning_rate = 1.0e-3
    schedule = "adaptive"
    gamma = 0.99
    lam = 0.95
    desired_kl = 0.01
    max_grad_norm = 1.0


class PPORunnerCfgV2(Cfg):
    """agents/rsl_rl_ppo_cfg.py:65-91"""
    class_name = "OnPolicyRunner"
    seed = 42
    device = "cuda:0"
    num_steps_per_env = 24
    max_iterations = 1000
    save_interval = 100
    experiment_name = "zbot_6b_flat_direct_v2"
    run_name = ""
    empirical_normalization = False
    clip_actions = None
    resume = False
    load_run = ".*"
    load_checkpoint = "model_.*.pt"
    logger = "tensorboard"
    policy = RslRlPpoActorCriticCfg()
    algorithm = RslRlPpoAlgorithmCfg()
