"""Shared helpers for the parity tests."""
import os

import numpy as np

from zbot_lab_b200.assets import zbot_6s as Z
from zbot_lab_b200.utils import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["mdp_v2_n64", "mdp_v2_n7", "mdp_v2_n300"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = syn.synth_mdp_case(int(g["seed"]), int(g["n"]), int(g["steps"]))
    return g, case


def make_mdp_oracle(n, origins):
    from oracle.mdp_oracle import MdpOracle
    return MdpOracle(n, origins, syn.reset_tables(), syn.index_sets(),
                     np.tile(np.asarray(Z.DEFAULT_JOINT_POS, np.float32), (n, 1)))


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


# ---- zbot-6b-walking-v4: V4Export row (csrc/zbot_core.h) -> the tensors the reference task reads ----
V4_EXPORT = {"base_pos": (0, 3), "base_quat": (3, 7), "base_link_vel": (7, 10), "feet_pos": (10, 16), "feet_quat": (16, 24),
             "feet_com_vel": (24, 30), "feet_fz_hist": (30, 36), "undesired_max": (36, 37), "last_air": (37, 39),
             "last_contact": (39, 41), "cur_air": (41, 43), "cur_contact": (43, 45), "joint_pos": (45, 51),
             "joint_vel": (51, 57), "applied_torque": (57, 63), "joint_acc": (63, 69)}


def v4_export_to_S(ex, n):
    ids = syn.index_sets()
    col = lambda k: np.asarray(ex[:, V4_EXPORT[k][0]:V4_EXPORT[k][1]], np.float32)
    base, (f0, f1), (s0, s1) = ids["base_body_idx"][0], ids["feet_body_idx"], ids["feet_ids"]
    S = {"body_link_pos_w": np.zeros((n, 12, 3), np.float32), "body_link_quat_w": np.zeros((n, 12, 4), np.float32),
         "body_link_lin_vel_w": np.zeros((n, 12, 3), np.float32), "body_com_lin_vel_w": np.zeros((n, 12, 3), np.float32),
         "net_forces_w_history": np.zeros((n, 3, 12, 3), np.float32)}
    S["body_link_quat_w"][..., 0] = 1.0
    S["body_link_pos_w"][:, base] = col("base_pos")
    S["body_link_quat_w"][:, base] = col("base_quat")
    S["body_link_lin_vel_w"][:, base] = col("base_link_vel")
    fp, fq, fv = col("feet_pos").reshape(n, 2, 3), col("feet_quat").reshape(n, 2, 4), col("feet_com_vel").reshape(n, 2, 3)
    for j, b in enumerate((f0, f1)):
        S["body_link_pos_w"][:, b], S["body_link_quat_w"][:, b], S["body_com_lin_vel_w"][:, b] = fp[:, j], fq[:, j], fv[:, j]
    fz = col("feet_fz_hist").reshape(n, 3, 2)
    S["net_forces_w_history"][:, :, s0, 2] = fz[:, :, 0]
    S["net_forces_w_history"][:, :, s1, 2] = fz[:, :, 1]
    S["net_forces_w_history"][:, 0, ids["undesired_ids"][0], 0] = col("undesired_max")[:, 0]   # the max, in one slot
    for k_src, k_dst in (("last_air", "last_air_time"), ("last_contact", "last_contact_time"),
                         ("cur_air", "current_air_time"), ("cur_contact", "current_contact_time")):
        t = np.zeros((n, 12), np.float32)
        t[:, s0], t[:, s1] = col(k_src)[:, 0], col(k_src)[:, 1]
        S[k_dst] = t
    for k in ("joint_pos", "joint_vel", "applied_torque", "joint_acc"):
        S[k] = col(k).copy()
    return S


def make_v4_oracle(n, origins):
    from oracle.v4_mdp_oracle import V4MdpOracle
    drp = (np.asarray(Z.DEFAULT_ROOT_POS, np.float32)[None] + origins).astype(np.float32)
    dq = np.asarray(Z.DEFAULT_JOINT_POS, np.float64)
    fk = lambda p, q: Z.fk_links(np.asarray(p, np.float64), np.asarray(q, np.float64), dq)
    return V4MdpOracle(n, syn.index_sets(), np.tile(np.asarray(Z.DEFAULT_JOINT_POS, np.float32), (n, 1)), drp, fk)


def v4_check_step(o, a, rnd, ex, obs, rew, term, trunc, ep_len, state_get, rtol=1e-5):
    """One step of the pinned v4 oracle on the exported view `ex`; asserts parity with the implementation's
    outputs / state (`state_get(name) -> (N, w)` array).  Returns (reset ids, interval ids)."""
    n = o.n
    obs_o, rew_o, term_o, trunc_o, ids_o, iv_o, log_o = o.step(a, v4_export_to_S(ex, n), rnd)
    assert np.array_equal(np.asarray(term, bool), term_o) and np.array_equal(np.asarray(trunc, bool), trunc_o)
    assert np.array_equal(np.asarray(ep_len), o.episode_length_buf)
    assert rel_err(rew, rew_o) <= rtol
    assert rel_err(obs, obs_o) <= 2 * rtol
    cmd = state_get("carry_feet_fz")
    assert rel_err(cmd, o.commands) <= rtol
    assert rel_err(state_get("carry_mid_max")[:, 0], o.target_heading_yaw) <= rtol
    assert rel_err(state_get("base_pos_y_err_sum")[:, 0], o.interval_time_left) <= rtol
    assert rel_err(state_get("p_delta"), o.p_delta) <= rtol
    assert rel_err(state_get("feet_step_length"), o.feet_step_length) <= rtol
    assert rel_err(state_get("feet_contact_forces_last"), o.feet_contact_forces_last) <= rtol
    assert rel_err(state_get("feet_down_pos_last").reshape(n, 2, 3), o.feet_down_pos_last) <= 2 * rtol
    eps = state_get("episode_sums")
    for i, nm in enumerate(o.episode_sums):
        assert rel_err(eps[:, i], o.episode_sums[nm]) <= rtol, nm
    return ids_o, iv_o, log_o


# ---------------------------------------------------------------------------------------------
# zbot-6b-walking-m-v0 (manager-based task)
# ---------------------------------------------------------------------------------------------
#: every RewTerm the fused step knows, with the reference cfg's params (zbotlab_env_cfg.py:240-352): 16 slots + is_terminated
M_ALL_TERMS = [
    ("track_lin_vel_xy_exp", "track_lin_vel_xy_yaw_frame_exp", 1.0, {"std": 0.5}),
    ("track_ang_vel_z_exp", "track_ang_vel_z_world_exp", 0.5, {"std": 0.5}),
    ("termination_penalty", "is_terminated", -200.0, {}),
    ("dof_torques_l2", "joint_torques_l2", -1.0e-5, {}),
    ("dof_acc_l2", "joint_acc_l2", -2.5e-7, {}),
    ("action_rate_l2", "action_rate_l2", -0.01, {}),
    ("foot_step_length", "foot_step_length", 2.0, {"command_name": None}),
    ("foot_downward", "foot_downward", -1.0, {}),
    ("foot_forward", "foot_forward", -0.5, {}),
    ("gait", "feet_gait", 0.5, {"period": 2.0, "offset": [0.0, 0.5], "threshold": 0.55, "command_name": "base_velocity"}),
    ("feet_slide", "feet_slide", -0.2, {}),
    ("feet_clearance", "foot_clearance_reward", 1.0, {"std": 0.05, "tanh_mult": 2.0, "target_height": 0.01}),
    ("feet_air_time", "feet_air_time_positive_biped", 2.5, {"command_name": "base_velocity", "threshold": 0.3}),
    ("air_time_variance", "air_time_balance_penalty", -1.0, {}),
    ("air_time_variance2", "air_time_variance_penalty", -1.0, {}),
    ("base_vel_forward", "base_vel_forward", 1.0, {"which_forward": 1}),
    ("feet_force_pattern", "feet_force_pattern", 1.0, {}),
]
M_PARAMS = {"minimum_height": 0.2, "feet_close_min": 0.12, "cmd_ranges": ((-0.3, 0.3), (-0.1, 0.1), (-0.2, 0.2)),
            "rel_standing_envs": 0.1, "resampling_time_range": (0.1, 0.3), "pose_range": ((-0.5, 0.5), (-0.5, 0.5), (-3.14, 3.14)),
            "max_episode_length": 1000, "step_dt": 0.02}


#: the cfg features the reference's registered cfgs switch off (zbotlab_env_cfg.py:86-97, 253-258, 367-371, 385-388), all on
M_EXTRA_TERMS = [t for t in M_ALL_TERMS if t[0] != "air_time_variance2"] + [
    ("undesired_contacts", "undesired_contacts", -1.0, {"threshold": 1.0})]
M_PARAMS_EXTRA = dict(M_PARAMS, cmd_ranges=((-0.3, 0.3), (-0.1, 0.1), (-1.0, 1.0)), illegal_contact=(1.0, 0b00100),
                      heading={"range": (-np.pi, np.pi), "stiffness": 0.5, "rel_heading_envs": 0.7},
                      push={"interval_range_s": (0.1, 0.4), "velocity_range": {"x": (-0.5, 0.5), "y": (-0.5, 0.5)}})


def m_native_cfg(n, terms, P=M_PARAMS, **kw):
    from zbot_lab_b200 import native
    slot = [(f, w, p) for _, f, w, p in terms if f != "is_terminated"]
    wt = [w for _, f, w, _ in terms if f == "is_terminated"]
    return native.make_m_cfg(n, slot, is_terminated_weight=wt[0] if wt else 0.0, minimum_height=P["minimum_height"],
                             feet_close_min=kw.pop("feet_close_min", P["feet_close_min"]), cmd_ranges=P["cmd_ranges"],
                             rel_standing_envs=P["rel_standing_envs"], resampling_time_range=P["resampling_time_range"],
                             pose_range=P["pose_range"], act_clip=0.04 * np.pi, illegal_contact=P.get("illegal_contact"),
                             heading=P.get("heading"), push=P.get("push"), **kw)


def m_make_oracle(n, terms, state_get, ep_len, P=M_PARAMS):
    """MMdpOracle initialised from the step implementation's own state words (``state_get(name, width)``)."""
    from oracle.m_mdp_oracle import MMdpOracle
    o = MMdpOracle(n, terms, P)
    o.s["episode_length_buf"] = np.asarray(ep_len, np.int64).copy()
    c2, c1 = state_get("carry_feet_fz", 2), state_get("carry_mid_max", 1)
    o.cmd[:] = np.concatenate([c2, c1], 1)
    o.standing[:] = state_get("base_heading_x_sum", 1)[:, 0] != 0
    o.time_left[:] = state_get("base_pos_y_err_sum", 1)[:, 0]
    o.s["feet_down_pos_last"][:] = state_get("feet_down_pos_last", 6).reshape(n, 2, 3)
    o.s["feet_contact_forces_last"][:] = state_get("feet_contact_forces_last", 2)
    o.s["feet_step_length"][:] = state_get("feet_step_length", 2)
    o.s["feet_force_sum"][:] = state_get("feet_force_sum", 1)[:, 0]
    o.s["action"][:] = state_get("actions", 6)
    pd = state_get("p_delta", 6)
    o.heading_target[:], o.is_heading[:], o.push_left[:] = pd[:, 0], pd[:, 1] != 0, pd[:, 2]
    return o


def m_check_step(o, a, rnd, ex, obs, rew, term, trunc, ep_len, state_get, rs=None, rtol=1e-5):
    """One step of the manager task: the implementation's outputs / state words against the reference-pinned oracle
    evaluated on the view the implementation exported.  Integer work bit-exact, floats <= rtol."""
    from oracle.m_mdp_oracle import split_view
    n = o.n
    view = split_view(np.asarray(ex, np.float32))
    view["qd_chain"] = np.asarray(state_get("joint_vel", 6), np.float32)       # reset envs are zeroed by the oracle anyway
    r = o.step(view, a, rnd)
    assert np.array_equal(r["terminated"], np.asarray(term, bool)), "terminated flags"
    assert np.array_equal(r["time_outs"], np.asarray(trunc, bool)), "time-out flags"
    assert np.array_equal(o.s["episode_length_buf"], np.asarray(ep_len)), "episode counters"
    assert rel_err(rew, r["reward"], 1.0) <= rtol, ("reward", rel_err(rew, r["reward"], 1.0))
    assert np.abs(np.asarray(obs) - r["obs"]).max() <= 5e-6, ("obs", np.abs(np.asarray(obs) - r["obs"]).max(0))
    cmd = np.concatenate([state_get("carry_feet_fz", 2), state_get("carry_mid_max", 1)], 1)
    assert np.abs(cmd - o.cmd).max() <= 1e-6, "commands"
    assert np.array_equal(state_get("base_heading_x_sum", 1)[:, 0] != 0, o.standing), "standing envs"
    assert np.abs(state_get("base_pos_y_err_sum", 1)[:, 0] - o.time_left).max() <= 1e-5, "command time_left"
    pd = state_get("p_delta", 6)
    if o.P.get("heading"):
        assert np.abs(pd[:, 0] - o.heading_target).max() <= 1e-6 and np.array_equal(pd[:, 1] != 0, o.is_heading), "heading command state"
    if o.P.get("push"):
        assert np.abs(pd[:, 2] - o.push_left).max() <= 1e-5, "push_robot interval timer"
    for k, w in (("feet_step_length", 2), ("feet_contact_forces_last", 2), ("feet_down_pos_last", 6)):
        assert np.abs(state_get(k, w).reshape(n, -1) - o.s[k].reshape(n, -1)).max() <= 2e-5, k
    assert np.abs(state_get("feet_force_sum", 1)[:, 0] - o.s["feet_force_sum"]).max() <= 1e-6
    assert np.abs(state_get("actions", 6) - o.s["action"]).max() == 0
    sums = state_get("episode_sums", 16)
    slot_names = [nm for nm, f, w, p in o.terms if f != "is_terminated" and float(w) != 0.0]
    for i, nm in enumerate(slot_names):
        assert rel_err(sums[:, i], o.ep_sums[nm], 1e-2) <= 20 * rtol, ("episode sum", nm)
    ids = r["reset_ids"]
    if rs is not None and len(ids):
        for i, nm in enumerate(slot_names):
            want = r["log"][nm] * (o.P["max_episode_length"] * o.P["step_dt"])     # mean episodic sum over the reset envs
            got = float(np.mean(np.asarray(rs)[ids, i], dtype=np.float32))
            assert abs(got - want) <= 20 * rtol * max(1e-2, abs(want)), ("log", nm, got, want)
        rs_ = np.asarray(rs)
        cols = [(13, 14, 15)] if len(slot_names) <= 13 else []           # mirrored in the spare term slots when there are any
        if rs_.shape[1] >= 20:
            cols.append((16, 17, 18))                                     # statistics words 22..24 (+ 25 = illegal_contact)
            assert int(rs_[ids, 19].sum()) == r["log"]["#illegal_contact"]
        for c_pen, c_low, c_close in cols:
            pen = [nm for nm, f, w, p in o.terms if f == "is_terminated"]
            if pen:
                want = r["log"][pen[0]] * (o.P["max_episode_length"] * o.P["step_dt"])
                assert abs(float(np.mean(rs_[ids, c_pen], dtype=np.float32)) - want) <= 1e-5 * max(1e-2, abs(want))
            assert int(rs_[ids, c_low].sum()) == r["log"]["#base_height"]
            assert int(rs_[ids, c_close].sum()) == r["log"]["#feet_close"]
    return r
