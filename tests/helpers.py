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
