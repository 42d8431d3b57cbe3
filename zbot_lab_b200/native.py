"""ctypes binding of the C-ABI shared library (``include/zbot_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` /
``python -m zbot_lab_b200.build`` into ``zbot_lab_b200/csrc/libzbot_b200.so``.
There is NO fallback: if the library is missing, :func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ZBOT_B200_LIB: load a tuning build of the SAME sources instead (tools/sweep_step.py); never a different implementation
LIB_PATH = os.environ.get("ZBOT_B200_LIB") or os.path.join(HERE, "csrc", "libzbot_b200.so")

ZBOT_ABI_VERSION = 7
TASK_WALKING_V2, TASK_SNAKE_V0, TASK_WALKING_V4, TASK_WALKING_M = 0, 1, 2, 3
M_NUM_OBS, M_NUM_RAND, M_EXPORT_WORDS = 25, 22, 72
V4_NUM_OBS, V4_NUM_RAND, V4_EXPORT_WORDS = 24, 10, 69
MAX_TERMS = 16
HOST_ROW_WORDS = 25   # zbot_step_host result row: obs 23 | reward | flags word
STATE_WORDS = 80
MDP_STATE_WORDS = 72
STATS_WORDS = 32
NUM_OBS = 23
NUM_ACTIONS = 6

# stats slot words (zbot_b200.h: zbot_step)
STAT_NUM_RESET = 16
STAT_NUM_TERMINATED_RESET = 17
STAT_NUM_TIMEOUT_RESET = 18
STAT_REWARD_SUM = 19
STAT_NUM_TERMINATED = 20
STAT_NUM_TRUNCATED = 21

TERM_IDS = {
    "base_vel_forward": 0, "feet_downward": 1, "feet_forward": 2, "base_heading_x": 3,
    "base_heading_x_sum": 4, "step_length": 5, "airtime_balance": 6, "action_rate": 7,
    "torques": 8, "feet_slide": 9, "base_pos_y_err": 10, "base_pos_y_err_sum": 11,
    "airtime_sum": 12, "feet_force_diff": 13, "feet_force_sum": 14,
}
#: snake task (zbot-6s-snake-v0): its own name -> id table (three names are shared with the walking task)
SNAKE_TERM_IDS = {
    "base_vel_forward": 0, "action_rate": 7, "torques": 8, "base_up_z": 15, "base_heading_y": 16,
    "base_heading_y_sum": 17, "base_pos_x_err": 18, "base_pos_x_err_sum": 19,
}


#: zbot-6b-walking-v4: its own name -> id table (…env_v4.py:1013-1199)
V4_TERM_IDS = {
    "track_lin_vel_x": 20, "track_heading_yaw": 21, "lin_vel_y": 22, "action_rate": 7, "torques": 8, "joint_vel": 23,
    "joint_acc": 24, "feet_downward": 1, "feet_forward": 2, "step_length": 25, "feet_air_time_biped": 26,
    "airtime_variance": 27, "feet_slide": 9, "feet_harmony": 28, "feet_close": 29, "lin_vel_x": 30, "airtime_sum": 31,
    "feet_height": 32, "base_height": 33,
}


class ZbotCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("task", C.c_int32), ("num_envs", C.c_int32), ("decimation", C.c_int32),
        ("max_episode_length", C.c_int32), ("sim_dt", C.c_float), ("termination_height", C.c_float),
        ("y_err_limit", C.c_float), ("terminated_penalty", C.c_float), ("contact_died_force", C.c_float),
        ("kp", C.c_float), ("kd", C.c_float), ("effort_limit", C.c_float), ("gravity", C.c_float),
        ("contact_alpha", C.c_float), ("contact_erp", C.c_float), ("contact_vdep", C.c_float),
        ("contact_beta_max", C.c_float), ("contact_mu", C.c_float), ("contact_ramp", C.c_float),
        ("contact_vt_eps", C.c_float), ("contact_margin", C.c_float),
        ("num_terms", C.c_int32), ("term_id", C.c_int32 * MAX_TERMS), ("term_weight", C.c_float * MAX_TERMS),
        ("ev_vel_lo", C.c_float), ("ev_vel_hi", C.c_float), ("ev_yaw_lo", C.c_float), ("ev_yaw_hi", C.c_float),
        ("ev_offset", C.c_float), ("ev_prob_pos", C.c_float), ("ev_dual_sign", C.c_int32),
        ("ev_pose_lo", C.c_float * 3), ("ev_pose_hi", C.c_float * 3),
        ("ev_interval_lo", C.c_float), ("ev_interval_hi", C.c_float), ("rng_seed", C.c_uint64),
        ("obs_noise_enable", C.c_int32), ("obs_noise_lo", C.c_float * 24), ("obs_noise_hi", C.c_float * 24),
        ("term_param", (C.c_float * 4) * MAX_TERMS), ("cmd_lo", C.c_float * 3), ("cmd_hi", C.c_float * 3),
        ("cmd_rel_standing", C.c_float), ("cmd_resample_lo", C.c_float), ("cmd_resample_hi", C.c_float),
        ("act_scale", C.c_float), ("act_clip", C.c_float), ("feet_close_min", C.c_float),
        ("is_terminated_weight", C.c_float),
        ("illegal_contact_threshold", C.c_float), ("illegal_contact_mask", C.c_int32), ("cmd_heading", C.c_int32),
        ("cmd_heading_lo", C.c_float), ("cmd_heading_hi", C.c_float), ("cmd_heading_stiffness", C.c_float),
        ("cmd_rel_heading", C.c_float), ("push_interval_lo", C.c_float), ("push_interval_hi", C.c_float),
        ("push_lo", C.c_float * 2), ("push_hi", C.c_float * 2),
    ]


class ZbotExport(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "body_link_pos_w0", "body_link_quat_w0", "body_com_lin_vel_w0",
        "body_link_pos_w1", "body_link_quat_w1", "body_com_lin_vel_w1",
        "joint_pos1", "joint_vel1", "applied_torque1",
        "net_forces_w_history1", "last_air_time1", "current_contact_time1")]


class ZbotMdpInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "body_link_pos_w", "body_link_quat_w", "body_com_lin_vel_w", "joint_pos", "joint_vel",
        "applied_torque", "net_forces_w_history", "last_air_time", "env_origins")]


class ZbotPolicy(C.Structure):
    """include/zbot_b200.h ``ZbotPolicy``: nn.Linear-layout weight pointers of the 3 x 128 ELU actor / critic."""
    _fields_ = [("actor_w", C.c_void_p * 4), ("actor_b", C.c_void_p * 4), ("critic_w", C.c_void_p * 4),
                ("critic_b", C.c_void_p * 4), ("std", C.c_void_p), ("num_obs", C.c_int32), ("num_actions", C.c_int32),
                ("hidden", C.c_int32), ("activation", C.c_int32)]


#: observation columns per term, in concatenation order (…env_v2.py:351-366 / …env_v4.py:835-847 / snake_v0.py:189-206)
OBS_TERMS = {"base_quat": (0, 4), "joint_pos": (4, 10), "joint_vel": (10, 16), "actions": (16, 22), "extra": (22, 24)}


def set_obs_noise(cfg: "ZbotCfg", noise: dict | None):
    """``noise`` = {term: (n_min, n_max)} over OBS_TERMS -- additive uniform corruption of the emitted observation
    (ObservationManager ``Unoise`` semantics, zbotlab_manager/zbotlab_env_cfg.py PolicyCfg); None / {} disables it."""
    for i in range(24):
        cfg.obs_noise_lo[i] = cfg.obs_noise_hi[i] = 0.0
    cfg.obs_noise_enable = 0
    for term, (lo, hi) in (noise or {}).items():
        if term not in OBS_TERMS:
            raise KeyError(f"unknown observation term {term!r}; known: {sorted(OBS_TERMS)}")
        a, b = OBS_TERMS[term]
        for i in range(a, b):
            cfg.obs_noise_lo[i], cfg.obs_noise_hi[i] = float(lo), float(hi)
        cfg.obs_noise_enable = 1
    return cfg


def make_cfg(num_envs: int, reward_scales: dict | None = None, step_dt: float | None = None,
             task: int = TASK_WALKING_V2, **overrides) -> ZbotCfg:
    """Build a ``ZbotCfg`` from the task constants (``zbot_lab_b200/assets/zbot_6s.py``) and a
    reward-scale dict in cfg order (``…env_v2.py:190-206``); weights are multiplied by
    ``step_dt`` here exactly once (the reference mutates its class-level dict, SURVEY C-3)."""
    from .assets import zbot_6s as Z

    cfg = ZbotCfg()
    cfg.abi_version = ZBOT_ABI_VERSION
    cfg.task = int(task)
    cfg.num_envs = int(num_envs)
    cfg.decimation = Z.DECIMATION
    cfg.max_episode_length = 800 if task == TASK_SNAKE_V0 else 1000
    cfg.sim_dt = Z.SIM_DT
    cfg.termination_height = 0.20 if task == TASK_WALKING_V4 else 0.22       # …env_v4.py:520 / …env_v2.py:44
    cfg.y_err_limit = 0.5
    cfg.terminated_penalty = 20.0
    cfg.contact_died_force = 0.5 if task == TASK_WALKING_V4 else 1.0          # …env_v4.py:880 / …env_v2.py:400
    cfg.kp, cfg.kd, cfg.effort_limit = Z.KP, Z.KD, Z.EFFORT_LIMIT
    if task == TASK_SNAKE_V0:
        from .assets import zbot_d_6s as S
        cfg.kp, cfg.kd, cfg.effort_limit = S.KP, S.KD, S.EFFORT_LIMIT
    cfg.gravity = Z.GRAVITY
    cfg.contact_alpha, cfg.contact_erp, cfg.contact_vdep = Z.CONTACT_ALPHA, Z.CONTACT_ERP, Z.CONTACT_VDEP
    cfg.contact_beta_max, cfg.contact_mu = Z.CONTACT_BETA_MAX, Z.CONTACT_MU
    cfg.contact_ramp, cfg.contact_vt_eps = Z.CONTACT_RAMP, Z.CONTACT_VT_EPS
    cfg.contact_margin = Z.CONTACT_MARGIN
    if task == TASK_WALKING_V4:
        # EventCfg (…env_v4.py:331-418): reset_base pose ranges, reset / interval command resampling
        cfg.ev_vel_lo, cfg.ev_vel_hi, cfg.ev_yaw_lo, cfg.ev_yaw_hi = 0.3, 0.3, -0.1, 0.1
        cfg.ev_offset, cfg.ev_prob_pos, cfg.ev_dual_sign = 0.0, 1.0, 1
        cfg.ev_pose_lo[:] = (-0.5, -0.5, -3.14)
        cfg.ev_pose_hi[:] = (0.5, 0.5, 3.14)
        cfg.ev_interval_lo, cfg.ev_interval_hi = 3.0, 6.0
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"ZbotCfg has no field {k!r}")
        if isinstance(v, (tuple, list)):
            getattr(cfg, k)[:] = v
        else:
            setattr(cfg, k, v)
    term_ids = {TASK_WALKING_V2: TERM_IDS, TASK_SNAKE_V0: SNAKE_TERM_IDS, TASK_WALKING_V4: V4_TERM_IDS}[task]
    if reward_scales is None:
        if task == TASK_WALKING_V2:
            from .tasks.zbot6b_direct.walking_v2_cfg import REWARD_SCALES_V2 as reward_scales
        elif task == TASK_WALKING_V4:
            from .tasks.zbot6b_direct.walking_v4_cfg import REWARD_SCALES_V4 as reward_scales
        else:
            from .tasks.zbot6_direct.snake_v0_cfg import REWARD_SCALES_SNAKE_V0 as reward_scales
    if step_dt is None:
        step_dt = cfg.decimation * Z.SIM_DT
    if len(reward_scales) > MAX_TERMS:
        raise ValueError(f"at most {MAX_TERMS} reward terms are supported")
    cfg.num_terms = len(reward_scales)
    for i, (name, w) in enumerate(reward_scales.items()):
        if name not in term_ids:
            raise KeyError(f"unknown reward term {name!r}; known: {sorted(term_ids)}")
        cfg.term_id[i] = term_ids[name]
        # v4 multiplies by step_dt at evaluation time (…env_v4.py:857): the kernel gets the bare weight
        cfg.term_weight[i] = float(w) if task == TASK_WALKING_V4 else float(w) * float(step_dt)
    return cfg


_LIB = None


def _declare(lib):
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    lib.zbot_abi_version.restype = C.c_int
    lib.zbot_cfg_sizeof.restype = C.c_int
    lib.zbot_build_info.restype = C.c_char_p
    lib.zbot_last_error.restype = C.c_char_p
    lib.zbot_default_cfg.argtypes = [P(ZbotCfg), i32]
    lib.zbot_state_word.argtypes = [C.c_char_p]
    lib.zbot_mdp_state_word.argtypes = [C.c_char_p]
    lib.zbot_create.argtypes = [P(ZbotCfg), C.c_int, P(vp)]
    lib.zbot_destroy.argtypes = [vp]
    lib.zbot_bind.argtypes = [vp, vp, vp, vp, i32]
    lib.zbot_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.zbot_step_export.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, P(ZbotExport), vp]
    lib.zbot_snake_step_export.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.zbot_step_host.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.zbot_v4_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.zbot_v4_step_export.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.zbot_m_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.zbot_m_step_export.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.zbot_update_cfg.argtypes = [vp, P(ZbotCfg)]
    lib.zbot_set_all_reset_spread.argtypes = [vp, i32]
    lib.zbot_bind_terrain.argtypes = [vp, vp, i32, i32, C.c_float, C.c_float, C.c_float, vp, i32, i32, C.c_float, vp, i32]
    lib.zbot_reset_idx.argtypes = [vp, vp, i64, vp, vp, i32, vp]
    lib.zbot_observe.argtypes = [vp, vp, vp]
    lib.zbot_articulation_view.argtypes = [vp, vp, vp, vp, vp]
    lib.zbot_mdp_bind.argtypes = [vp, vp, vp, vp, i32]
    lib.zbot_mdp_observe.argtypes = [vp, P(ZbotMdpInputs), vp, vp]
    lib.zbot_mdp_step.argtypes = [vp, P(ZbotMdpInputs), vp, vp, vp, vp, vp, i32, i32, vp]
    lib.zbot_policy_act.argtypes = [vp, P(ZbotPolicy), vp, vp, vp, vp, vp, vp, vp, C.c_uint64, vp]
    lib.zbot_rollout_store.argtypes = [vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]
    lib.zbot_launch_count.argtypes = [vp]
    lib.zbot_launch_count.restype = i64
    lib.zbot_step_kernel_name.argtypes = [vp]
    lib.zbot_step_kernel_name.restype = C.c_char_p
    for name in ("zbot_default_cfg", "zbot_state_word", "zbot_mdp_state_word", "zbot_create", "zbot_destroy",
                 "zbot_bind", "zbot_step", "zbot_step_export", "zbot_snake_step_export", "zbot_step_host", "zbot_v4_step",
                 "zbot_v4_step_export", "zbot_m_step", "zbot_m_step_export", "zbot_update_cfg", "zbot_reset_idx",
                 "zbot_observe", "zbot_set_all_reset_spread", "zbot_bind_terrain", "zbot_policy_act",
                 "zbot_rollout_store", "zbot_articulation_view", "zbot_mdp_bind", "zbot_mdp_observe", "zbot_mdp_step"):
        getattr(lib, name).restype = C.c_int


EXPORTED_SYMBOLS = (
    "zbot_abi_version", "zbot_cfg_sizeof", "zbot_build_info", "zbot_last_error", "zbot_default_cfg", "zbot_state_word",
    "zbot_mdp_state_word", "zbot_create", "zbot_destroy", "zbot_bind", "zbot_step", "zbot_step_export",
    "zbot_snake_step_export", "zbot_step_host", "zbot_v4_step", "zbot_v4_step_export", "zbot_m_step", "zbot_m_step_export",
    "zbot_update_cfg", "zbot_reset_idx", "zbot_observe", "zbot_articulation_view", "zbot_mdp_bind", "zbot_mdp_observe",
    "zbot_mdp_step", "zbot_launch_count", "zbot_set_all_reset_spread", "zbot_step_kernel_name", "zbot_bind_terrain",
    "zbot_policy_act", "zbot_rollout_store",
)


def lib():
    """Load ``libzbot_b200.so``; raise (never fall back) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m zbot_lab_b200.build`). "
                "There is no CPU fallback.")
        _LIB = C.CDLL(LIB_PATH)
        _declare(_LIB)
        if _LIB.zbot_abi_version() != ZBOT_ABI_VERSION:
            raise RuntimeError("libzbot_b200.so ABI version mismatch; rebuild")
        if _LIB.zbot_cfg_sizeof() != C.sizeof(ZbotCfg):
            raise RuntimeError(f"libzbot_b200.so was built with sizeof(ZbotCfg) = {_LIB.zbot_cfg_sizeof()}, the binding has "
                               f"{C.sizeof(ZbotCfg)}: stale library, rebuild")
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().zbot_last_error()
        raise RuntimeError(f"{what or 'zbot call'} failed (rc={rc}): {msg.decode() if msg else ''}")


# --------------------------------------------------------------------------------------------------------------
# zbot-6b-walking-m-v0 (manager-based task): RewTerm function name -> (term id, params -> term_param row)
# (reference: tasks/zbotlab_manager/mdp/rewards.py; joint_torques_l2 / joint_acc_l2 / action_rate_l2 / is_terminated are
# isaaclab.envs.mdp functions [IL-upstream])
# --------------------------------------------------------------------------------------------------------------
def _need(cond, msg):
    if not cond:
        raise NotImplementedError(msg)


def _gait_params(p):
    _need(p.get("command_name") is not None, "feet_gait without command gating is not built")
    off = list(p["offset"])
    _need(len(off) == 2, "feet_gait: two feet")
    return [p["period"], off[0], off[1], p.get("threshold", 0.5)]


def _step_length_params(p):
    _need(p.get("command_name") is None, "foot_step_length(command_name=...) (in-place command overwrite, rewards.py:79-80) is not built")
    return []


M_REWARD_FUNCS = {
    "track_lin_vel_xy_yaw_frame_exp": (34, lambda p: [p["std"] ** 2]),
    "track_ang_vel_z_world_exp": (35, lambda p: [p["std"] ** 2]),
    "joint_torques_l2": (8, lambda p: []),
    "joint_acc_l2": (24, lambda p: []),
    "action_rate_l2": (7, lambda p: []),
    "foot_step_length": (36, _step_length_params),
    "foot_downward": (1, lambda p: []),
    "foot_forward": (2, lambda p: []),
    "feet_gait": (37, _gait_params),
    "feet_slide": (38, lambda p: []),
    "foot_clearance_reward": (39, lambda p: [p["std"], p["tanh_mult"], p["target_height"]]),
    "feet_air_time_positive_biped": (40, lambda p: [p["threshold"]]),
    "air_time_balance_penalty": (6, lambda p: []),
    "air_time_variance_penalty": (27, lambda p: []),
    "base_vel_forward": (41, lambda p: [float(p.get("which_forward", 1))]),
    "feet_force_pattern": (42, lambda p: []),
    "undesired_contacts": (43, lambda p: [p.get("threshold", 1.0)]),     # isaaclab.envs.mdp [IL-upstream]
}


def make_m_cfg(num_envs: int, terms, *, is_terminated_weight: float = 0.0, minimum_height: float = 0.2,
               feet_close_min: float = 0.12, cmd_ranges=((-0.1, 0.1), (0.0, 0.0), (0.0, 0.0)), rel_standing_envs: float = 0.02,
               resampling_time_range=(10.0, 10.0), act_scale: float = 0.04 * 3.141592653589793, act_clip: float | None = None,
               pose_range=((-0.5, 0.5), (-0.5, 0.5), (-3.14, 3.14)), episode_length_s: float = 20.0, friction: float = 1.0,
               rng_seed: int = 0, illegal_contact: tuple | None = None, heading: dict | None = None, push: dict | None = None,
               **overrides) -> ZbotCfg:
    """``ZbotCfg`` of the manager-based task.  ``terms`` = [(func_name, weight, params_dict), ...] in cfg order
    (zero-weight terms are skipped, as RewardManager does); the defaults are ``Zbot6BFlatEnvCfg``
    (config/zbot6b_manager/flat_env_cfg.py) over ``ZbotLabRoughEnvCfg`` (zbotlab_env_cfg.py:99-452)."""
    from .assets import zbot_6s as Z
    from .assets import zbot_6s_v2 as V

    cfg = make_cfg(num_envs, reward_scales={}, task=TASK_WALKING_V2)
    cfg.task = TASK_WALKING_M
    cfg.kp, cfg.kd, cfg.effort_limit = V.KP, V.KD, V.EFFORT_LIMIT
    cfg.max_episode_length = int(-(-episode_length_s // (Z.SIM_DT * Z.DECIMATION)))
    cfg.termination_height = float(minimum_height)
    cfg.feet_close_min = float(feet_close_min) if feet_close_min else 0.0
    cfg.is_terminated_weight = float(is_terminated_weight)
    cfg.contact_mu = float(friction)
    for i in range(3):
        cfg.cmd_lo[i], cfg.cmd_hi[i] = float(cmd_ranges[i][0]), float(cmd_ranges[i][1])
        cfg.ev_pose_lo[i], cfg.ev_pose_hi[i] = float(pose_range[i][0]), float(pose_range[i][1])
    cfg.cmd_rel_standing = float(rel_standing_envs)
    cfg.cmd_resample_lo, cfg.cmd_resample_hi = float(resampling_time_range[0]), float(resampling_time_range[1])
    cfg.act_scale = float(act_scale)
    cfg.act_clip = float(act_clip if act_clip is not None else 3.0e38)
    cfg.rng_seed = int(rng_seed)
    if illegal_contact is not None:      # DoneTerm mdp.illegal_contact: (threshold, merged-body mask)
        cfg.illegal_contact_threshold, cfg.illegal_contact_mask = float(illegal_contact[0]), int(illegal_contact[1])
    if heading is not None:              # UniformVelocityCommandCfg(heading_command=True, ...)
        cfg.cmd_heading = 1
        cfg.cmd_heading_lo, cfg.cmd_heading_hi = float(heading["range"][0]), float(heading["range"][1])
        cfg.cmd_heading_stiffness, cfg.cmd_rel_heading = float(heading["stiffness"]), float(heading["rel_heading_envs"])
    if push is not None:                 # EventTerm push_by_setting_velocity, mode="interval"
        cfg.push_interval_lo, cfg.push_interval_hi = float(push["interval_range_s"][0]), float(push["interval_range_s"][1])
        for i, k in enumerate(("x", "y")):
            cfg.push_lo[i], cfg.push_hi[i] = (float(v) for v in push["velocity_range"].get(k, (0.0, 0.0)))
    active = [(f, w, p) for f, w, p in terms if float(w) != 0.0]
    if len(active) > MAX_TERMS - 0:
        raise ValueError(f"at most {MAX_TERMS} weighted reward terms are supported (termination_penalty excluded)")
    cfg.num_terms = len(active)
    for i, (func, w, p) in enumerate(active):
        if func not in M_REWARD_FUNCS:
            raise NotImplementedError(f"reward function {func!r} is not built into the fused step; known: {sorted(M_REWARD_FUNCS)}")
        tid, par = M_REWARD_FUNCS[func]
        cfg.term_id[i] = tid
        cfg.term_weight[i] = float(w)            # bare weight: RewardManager multiplies by dt per evaluation
        row = [float(x) for x in par(dict(p or {}))]
        for j in range(4):
            cfg.term_param[i][j] = row[j] if j < len(row) else 0.0
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"ZbotCfg has no field {k!r}")
        if isinstance(v, (tuple, list)):
            getattr(cfg, k)[:] = v
        else:
            setattr(cfg, k, v)
    return cfg


#: Zbot6BFlatEnvCfg's active RewTerms in cfg order (zbotlab_env_cfg.py:240-352 with flat_env_cfg.py:96-112 applied)
M_FLAT_TERMS = [
    ("track_lin_vel_xy_exp", "track_lin_vel_xy_yaw_frame_exp", 1.0, {"std": 0.5}),
    ("track_ang_vel_z_exp", "track_ang_vel_z_world_exp", 0.5, {"std": 0.5}),
    ("termination_penalty", "is_terminated", -200.0, {}),
    ("dof_torques_l2", "joint_torques_l2", -1.0e-5, {}),
    ("dof_acc_l2", "joint_acc_l2", -2.5e-7, {}),
    ("action_rate_l2", "action_rate_l2", -0.01, {}),
    ("foot_step_length", "foot_step_length", 5.0, {"command_name": None}),
    ("foot_downward", "foot_downward", -1.0, {}),
    ("foot_forward", "foot_forward", -0.5, {}),
    ("feet_slide", "feet_slide", -6.5, {}),
    ("air_time_variance", "air_time_balance_penalty", -15.0, {}),
]
