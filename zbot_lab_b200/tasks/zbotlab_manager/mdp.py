"""Term vocabulary of the manager-based task: the cfg classes (``RewTerm`` / ``ObsTerm`` / ``DoneTerm`` / ``EventTerm`` /
``CurrTerm`` / ``SceneEntityCfg`` / ``Unoise`` / action and command cfgs) and the term FUNCTIONS by name.

In the reference a term function is Python executed by a manager every step
(``/root/reference/source/zbot/zbot/tasks/zbotlab_manager/mdp/rewards.py`` etc., plus ``isaaclab.envs.mdp``).  Here the
functions are *names*: ``ManagerBasedRLEnv`` (``manager_env.py``) reads ``term.func.__name__`` / ``term.params`` /
``term.weight`` from the cfg and compiles them into the term table of the fused kernel (``native.make_m_cfg``); calling
one raises.  An unknown function name fails loudly at env construction -- nothing is silently dropped."""
from __future__ import annotations

from ...utils.configclass import Cfg


class SceneEntityCfg(Cfg):
    name = "robot"
    body_names = None
    joint_names = None

    def __init__(self, name: str = "robot", **kw):
        super().__init__(name=name, **kw)


class _TermCfg(Cfg):
    func = None
    params = {}

    def __init__(self, func=None, **kw):
        kw.setdefault("params", {})
        super().__init__(func=func, **kw)


class RewardTermCfg(_TermCfg):
    weight = 0.0


class TerminationTermCfg(_TermCfg):
    time_out = False


class ObservationTermCfg(_TermCfg):
    noise = None
    clip = None
    scale = None


class EventTermCfg(_TermCfg):
    mode = "reset"
    interval_range_s = None


class CurriculumTermCfg(_TermCfg):
    pass


class ObservationGroupCfg(Cfg):
    enable_corruption = False
    concatenate_terms = True


class AdditiveUniformNoiseCfg(Cfg):
    n_min = -1.0
    n_max = 1.0


class RelativeJointPositionActionCfg(Cfg):
    """[IL-upstream] processed = clip(raw * scale + 0); applied each substep as processed + current joint position."""
    asset_name = "robot"
    joint_names = ["joint.*"]
    scale = 1.0
    clip = None
    use_zero_offset = True


class UniformVelocityCommandCfg(Cfg):
    class Ranges(Cfg):
        lin_vel_x = (0.0, 0.0)
        lin_vel_y = (0.0, 0.0)
        ang_vel_z = (0.0, 0.0)
        heading = None

    asset_name = "robot"
    resampling_time_range = (10.0, 10.0)
    rel_standing_envs = 0.0
    rel_heading_envs = 1.0
    heading_command = False
    heading_control_stiffness = 1.0
    debug_vis = False
    ranges = Ranges()


class UniformLevelVelocityCommandCfg(UniformVelocityCommandCfg):
    """mdp/commands/velocity_command.py:9-11: the command cfg plus the curriculum's ``limit_ranges``."""
    limit_ranges = UniformVelocityCommandCfg.Ranges()


def _term(name: str, doc: str):
    def f(env, *a, **k):
        raise RuntimeError(f"mdp.{name} is evaluated inside the fused CUDA step; it is a cfg name here, not host code")
    f.__name__ = f.__qualname__ = name
    f.__doc__ = doc
    return f


# rewards: reference mdp/rewards.py (file:line of each in oracle/m_mdp_oracle.py) + isaaclab.envs.mdp [IL-upstream]
for _n in ("track_lin_vel_xy_yaw_frame_exp", "track_ang_vel_z_world_exp", "is_terminated", "joint_torques_l2", "joint_acc_l2",
           "action_rate_l2", "foot_step_length", "foot_downward", "foot_forward", "feet_gait", "feet_slide",
           "foot_clearance_reward", "feet_air_time_positive_biped", "air_time_variance_penalty", "air_time_balance_penalty",
           "base_vel_forward", "feet_force_pattern", "undesired_contacts", "feet_air_time", "stand_still_joint_deviation_l1",
           # terminations
           "time_out", "illegal_contact", "root_height_below_minimum", "feet_close", "terrain_out_of_bounds",
           # observations
           "root_quat_w", "generated_commands", "joint_pos_rel", "joint_vel_rel", "last_action", "base_lin_vel",
           "base_ang_vel", "projected_gravity",
           # events
           "init_my_data", "reset_my_data", "randomize_rigid_body_material", "randomize_rigid_body_mass",
           "randomize_rigid_body_com", "reset_root_state_uniform", "reset_joints_by_scale", "push_by_setting_velocity",
           "apply_external_force_torque",
           # curricula
           "terrain_levels_vel", "lin_vel_cmd_levels", "ang_vel_cmd_levels"):
    globals()[_n] = _term(_n, "cfg name of a manager term; see the module docstring")
del _n
