// zbot_layout.h -- word layout of the per-env state buffers and ZbotCfg -> Params conversion.
// Shared by the CUDA kernels (state held as [NQ][N] float4, loaded 16 B per thread, coalesced)
// and by the CPU port under oracle/ (state held as [N][WORDS]).
#pragma once
#include <string.h>

#include "../../include/zbot_b200.h"
#include "zbot_core.h"

namespace zbot {

// fused-step state (ZBOT_STATE_WORDS = 80)
enum StateWord : int {
  // "early" words (quads 0..10): everything the physics phase touches
  W_ROOT_POS = 0, W_ROOT_QUAT = 3, W_ROOT_LIN = 7, W_ROOT_ANG = 10, W_Q = 13, W_QD = 19,
  W_PDELTA = 25, W_SPEED = 31, W_CARRY_FZ = 32, W_CARRY_MID = 34,
  W_CUR_AIR = 36, W_CUR_CONTACT = 38, W_LAST_AIR = 40, W_LAST_CONTACT = 42,
  // "late" words (quads 11..19): MDP state only needed after the physics
  W_ACT = 44, W_FLAST = 50, W_FDPL = 52, W_FSL = 58, W_HSUM = 60, W_YSUM = 61, W_FFSUM = 62,
  W_EPSUM = 64
};
constexpr int EARLY_QUADS = 11, SIM_QUADS = 7;   // quads [0,11) early, [11,20) late; quads [0,7) hold the SimState
// MDP-only state (ZBOT_MDP_STATE_WORDS = 72)
enum MdpWord : int {
  M_PDELTA = 0, M_ACT = 6, M_FLAST = 12, M_FDPL = 14, M_FSL = 20, M_HSUM = 22, M_YSUM = 23, M_FFSUM = 24,
  M_SPEED = 25, M_ST_BASE_POS = 26, M_ST_FORWARD = 29, M_ST_FEET_X = 32, M_ST_FEET_Z = 38,
  M_ST_FEET_POS = 44, M_ST_VFWD = 50, M_EPSUM = 56
};

struct FieldName { const char* name; int word; };
static const FieldName kStateFields[] = {
    {"root_pos", W_ROOT_POS}, {"root_quat", W_ROOT_QUAT}, {"root_lin_vel", W_ROOT_LIN},
    {"root_ang_vel", W_ROOT_ANG}, {"joint_pos", W_Q}, {"joint_vel", W_QD}, {"p_delta", W_PDELTA},
    {"actions", W_ACT}, {"carry_feet_fz", W_CARRY_FZ}, {"carry_mid_max", W_CARRY_MID},
    {"current_air_time", W_CUR_AIR}, {"current_contact_time", W_CUR_CONTACT},
    {"last_air_time", W_LAST_AIR}, {"last_contact_time", W_LAST_CONTACT},
    {"feet_contact_forces_last", W_FLAST}, {"feet_down_pos_last", W_FDPL},
    {"feet_step_length", W_FSL}, {"base_heading_x_sum", W_HSUM}, {"base_pos_y_err_sum", W_YSUM},
    {"feet_force_sum", W_FFSUM}, {"joint_speed_limit", W_SPEED}, {"episode_sums", W_EPSUM}};
static const FieldName kMdpFields[] = {
    {"p_delta", M_PDELTA}, {"actions", M_ACT}, {"feet_contact_forces_last", M_FLAST},
    {"feet_down_pos_last", M_FDPL}, {"feet_step_length", M_FSL}, {"base_heading_x_sum", M_HSUM},
    {"base_pos_y_err_sum", M_YSUM}, {"feet_force_sum", M_FFSUM}, {"joint_speed_limit", M_SPEED},
    {"stale_base_pos", M_ST_BASE_POS}, {"stale_forward", M_ST_FORWARD}, {"stale_feet_x", M_ST_FEET_X},
    {"stale_feet_z", M_ST_FEET_Z}, {"stale_feet_pos", M_ST_FEET_POS}, {"stale_v_fwd", M_ST_VFWD},
    {"episode_sums", M_EPSUM}};

inline int find_word(const FieldName* tab, int n, const char* f) {
  if (!f) return -1;
  for (int i = 0; i < n; ++i)
    if (strcmp(tab[i].name, f) == 0) return tab[i].word;
  return -1;
}

template <typename T>
ZB_HD void mdp_state_unpack(const T* w, int pd, int act, int fl, int fdpl, int fsl, int hs, int ys, int ffs,
                            int sp, int eps, MdpState<T>& m) {
  ZB_UNROLL for (int i = 0; i < 6; ++i) { m.p_delta[i] = w[pd + i]; m.actions[i] = w[act + i]; }
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    m.feet_force_last[j] = w[fl + j];
    m.feet_step_length[j] = w[fsl + j];
    ZB_UNROLL for (int i = 0; i < 3; ++i) m.feet_down_pos_last[j][i] = w[fdpl + 3 * j + i];
  }
  m.heading_sum = w[hs]; m.y_err_sum = w[ys]; m.feet_force_sum = w[ffs]; m.speed_limit = w[sp];
  ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) m.ep_sums[i] = w[eps + i];
}
template <typename T>
ZB_HD void mdp_state_pack(const MdpState<T>& m, T* w, int pd, int act, int fl, int fdpl, int fsl, int hs,
                          int ys, int ffs, int sp, int eps) {
  ZB_UNROLL for (int i = 0; i < 6; ++i) { w[pd + i] = m.p_delta[i]; w[act + i] = m.actions[i]; }
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    w[fl + j] = m.feet_force_last[j];
    w[fsl + j] = m.feet_step_length[j];
    ZB_UNROLL for (int i = 0; i < 3; ++i) w[fdpl + 3 * j + i] = m.feet_down_pos_last[j][i];
  }
  w[hs] = m.heading_sum; w[ys] = m.y_err_sum; w[ffs] = m.feet_force_sum; w[sp] = m.speed_limit;
  ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) w[eps + i] = m.ep_sums[i];
}

template <typename T>
ZB_HD void env_state_unpack(const T* w, EnvState<T>& e) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { e.sim.p[i] = w[W_ROOT_POS + i]; e.sim.v[i] = w[W_ROOT_LIN + i]; e.sim.w[i] = w[W_ROOT_ANG + i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) e.sim.Q[i] = w[W_ROOT_QUAT + i];
  ZB_UNROLL for (int i = 0; i < 6; ++i) { e.sim.q[i] = w[W_Q + i]; e.sim.qd[i] = w[W_QD + i]; }
  mdp_state_unpack(w, W_PDELTA, W_ACT, W_FLAST, W_FDPL, W_FSL, W_HSUM, W_YSUM, W_FFSUM, W_SPEED, W_EPSUM, e.mdp);
  e.carry_feet_fz[0] = w[W_CARRY_FZ]; e.carry_feet_fz[1] = w[W_CARRY_FZ + 1];
  e.carry_mid_max = w[W_CARRY_MID];
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    e.timers[j].cur_air = w[W_CUR_AIR + j]; e.timers[j].cur_contact = w[W_CUR_CONTACT + j];
    e.timers[j].last_air = w[W_LAST_AIR + j]; e.timers[j].last_contact = w[W_LAST_CONTACT + j];
  }
}
template <typename T>
ZB_HD void env_state_pack(const EnvState<T>& e, T* w) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { w[W_ROOT_POS + i] = e.sim.p[i]; w[W_ROOT_LIN + i] = e.sim.v[i]; w[W_ROOT_ANG + i] = e.sim.w[i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) w[W_ROOT_QUAT + i] = e.sim.Q[i];
  ZB_UNROLL for (int i = 0; i < 6; ++i) { w[W_Q + i] = e.sim.q[i]; w[W_QD + i] = e.sim.qd[i]; }
  mdp_state_pack(e.mdp, w, W_PDELTA, W_ACT, W_FLAST, W_FDPL, W_FSL, W_HSUM, W_YSUM, W_FFSUM, W_SPEED, W_EPSUM);
  w[W_CARRY_FZ] = e.carry_feet_fz[0]; w[W_CARRY_FZ + 1] = e.carry_feet_fz[1];
  w[W_CARRY_MID] = e.carry_mid_max;
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    w[W_CUR_AIR + j] = e.timers[j].cur_air; w[W_CUR_CONTACT + j] = e.timers[j].cur_contact;
    w[W_LAST_AIR + j] = e.timers[j].last_air; w[W_LAST_CONTACT + j] = e.timers[j].last_contact;
  }
  w[35] = T(0); w[63] = T(0);
}

// --- split views of the same 80 words (GPU: load the early quads, run the physics, then load the rest) ---
template <typename T>
ZB_HD void sim_state_unpack(const T* w, SimState<T>& s) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { s.p[i] = w[W_ROOT_POS + i]; s.v[i] = w[W_ROOT_LIN + i]; s.w[i] = w[W_ROOT_ANG + i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) s.Q[i] = w[W_ROOT_QUAT + i];
  ZB_UNROLL for (int i = 0; i < 6; ++i) { s.q[i] = w[W_Q + i]; s.qd[i] = w[W_QD + i]; }
}
template <typename T>
ZB_HD void env_early_unpack(const T* w /*first 44 words*/, EnvState<T>& e) {
  sim_state_unpack(w, e.sim);
  ZB_UNROLL for (int i = 0; i < 6; ++i) e.mdp.p_delta[i] = w[W_PDELTA + i];
  e.mdp.speed_limit = w[W_SPEED];
  e.carry_feet_fz[0] = w[W_CARRY_FZ]; e.carry_feet_fz[1] = w[W_CARRY_FZ + 1];
  e.carry_mid_max = w[W_CARRY_MID];
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    e.timers[j].cur_air = w[W_CUR_AIR + j]; e.timers[j].cur_contact = w[W_CUR_CONTACT + j];
    e.timers[j].last_air = w[W_LAST_AIR + j]; e.timers[j].last_contact = w[W_LAST_CONTACT + j];
  }
}
template <typename T>
ZB_HD void env_late_unpack(const T* w /*words 44..79, indexed from 0*/, EnvState<T>& e) {
  constexpr int O = 4 * EARLY_QUADS;
  ZB_UNROLL for (int i = 0; i < 6; ++i) e.mdp.actions[i] = w[W_ACT - O + i];
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    e.mdp.feet_force_last[j] = w[W_FLAST - O + j];
    e.mdp.feet_step_length[j] = w[W_FSL - O + j];
    ZB_UNROLL for (int i = 0; i < 3; ++i) e.mdp.feet_down_pos_last[j][i] = w[W_FDPL - O + 3 * j + i];
  }
  e.mdp.heading_sum = w[W_HSUM - O]; e.mdp.y_err_sum = w[W_YSUM - O]; e.mdp.feet_force_sum = w[W_FFSUM - O];
  ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) e.mdp.ep_sums[i] = w[W_EPSUM - O + i];
}

template <typename T>
ZB_HD void stale_unpack(const T* w, StaleCache<T>& c) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { c.base_pos[i] = w[M_ST_BASE_POS + i]; c.forward[i] = w[M_ST_FORWARD + i]; }
  ZB_UNROLL for (int j = 0; j < 2; ++j)
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      c.feet_x[j][i] = w[M_ST_FEET_X + 3 * j + i];
      c.feet_z[j][i] = w[M_ST_FEET_Z + 3 * j + i];
      c.feet_pos[j][i] = w[M_ST_FEET_POS + 3 * j + i];
    }
  c.v_fwd = w[M_ST_VFWD];
}
template <typename T>
ZB_HD void stale_pack(const StaleCache<T>& c, T* w) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { w[M_ST_BASE_POS + i] = c.base_pos[i]; w[M_ST_FORWARD + i] = c.forward[i]; }
  ZB_UNROLL for (int j = 0; j < 2; ++j)
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      w[M_ST_FEET_X + 3 * j + i] = c.feet_x[j][i];
      w[M_ST_FEET_Z + 3 * j + i] = c.feet_z[j][i];
      w[M_ST_FEET_POS + 3 * j + i] = c.feet_pos[j][i];
    }
  w[M_ST_VFWD] = c.v_fwd;
}

// ZbotCfg (C ABI) -> Params<T>
template <typename T>
inline void params_from_cfg(const ZbotCfg& c, Params<T>& P) {
  P.dt = T(c.sim_dt);
  P.kp = T(c.kp); P.kd = T(c.kd); P.effort = T(c.effort_limit);
  P.arm = P.dt * P.kd + P.dt * P.dt * P.kp;
  P.gravity = T(c.gravity);
  P.c_k = T(c.contact_alpha) * T(c.contact_erp) / P.dt;
  P.c_d = T(c.contact_alpha) * (T(1) - T(c.contact_erp));
  P.c_fcap = T(c.contact_alpha) * T(c.contact_vdep);
  P.c_beta_max = T(c.contact_beta_max);
  P.c_mu = T(c.contact_mu);
  P.c_inv_ramp = T(1) / T(c.contact_ramp);
  P.c_vt_eps = T(c.contact_vt_eps);
  P.c_margin = T(c.contact_margin);
  P.c_inv_fband = T(1) / (T(0.25) * P.c_k * T(c.contact_ramp));   // activation band: 5 N at the defaults
  P.decimation = c.decimation;
  P.step_dt = T(c.sim_dt * (float)c.decimation);
  P.termination_height = T(c.termination_height);
  P.y_limit = T(c.y_err_limit);
  P.term_penalty = T(c.terminated_penalty);
  P.contact_died_threshold = T(c.contact_died_force);
  P.max_episode_length = c.max_episode_length;
  P.num_terms = c.num_terms;
  for (int i = 0; i < MAX_TERMS; ++i) { P.term_id[i] = c.term_id[i]; P.term_w[i] = T(c.term_weight[i]); }
  P.ev_vel_lo = T(c.ev_vel_lo); P.ev_vel_hi = T(c.ev_vel_hi); P.ev_yaw_lo = T(c.ev_yaw_lo); P.ev_yaw_hi = T(c.ev_yaw_hi);
  P.ev_offset = T(c.ev_offset); P.ev_prob_pos = T(c.ev_prob_pos); P.ev_dual_sign = c.ev_dual_sign;
  for (int i = 0; i < 3; ++i) { P.ev_pose_lo[i] = T(c.ev_pose_lo[i]); P.ev_pose_hi[i] = T(c.ev_pose_hi[i]); }
  P.ev_interval_lo = T(c.ev_interval_lo); P.ev_interval_hi = T(c.ev_interval_hi);
  P.obs_noise_enable = c.obs_noise_enable;
  for (int i = 0; i < 24; ++i) { P.obs_noise_lo[i] = T(c.obs_noise_lo[i]); P.obs_noise_w[i] = T(c.obs_noise_hi[i]) - T(c.obs_noise_lo[i]); }
  P.rng_seed = c.rng_seed;
  for (int i = 0; i < MAX_TERMS; ++i)
    for (int j = 0; j < 4; ++j) P.term_par[i][j] = T(c.term_param[i][j]);
  for (int i = 0; i < 3; ++i) { P.cmd_lo[i] = T(c.cmd_lo[i]); P.cmd_hi[i] = T(c.cmd_hi[i]); }
  P.cmd_rel_standing = T(c.cmd_rel_standing);
  P.cmd_resample_lo = T(c.cmd_resample_lo); P.cmd_resample_hi = T(c.cmd_resample_hi);
  P.act_scale = T(c.act_scale); P.act_clip = T(c.act_clip);
  P.feet_close_min = T(c.feet_close_min);
  P.term_penalty_w = T(c.is_terminated_weight) * P.step_dt;       // value * weight * dt with value = 1 (RewardManager)
  P.illegal_thr = T(c.illegal_contact_threshold); P.illegal_mask = c.illegal_contact_mask;
  P.cmd_heading = c.cmd_heading;
  P.cmd_heading_lo = T(c.cmd_heading_lo); P.cmd_heading_hi = T(c.cmd_heading_hi);
  P.cmd_heading_stiffness = T(c.cmd_heading_stiffness); P.cmd_rel_heading = T(c.cmd_rel_heading);
  P.push_interval_lo = T(c.push_interval_lo); P.push_interval_hi = T(c.push_interval_hi);
  for (int i = 0; i < 2; ++i) { P.push_lo[i] = T(c.push_lo[i]); P.push_hi[i] = T(c.push_hi[i]); }
  P.default_terms = (c.num_terms == 13) && (c.task == ZBOT_TASK_WALKING_V2);
  for (int i = 0; i < 13 && P.default_terms; ++i) P.default_terms = (c.term_id[i] == i);
}

inline int cfg_validate(const ZbotCfg& c, const char** why) {
  if (c.abi_version != ZBOT_ABI_VERSION) { *why = "ZbotCfg.abi_version mismatch"; return ZBOT_E_INVALID; }
  if (c.num_envs < 1) { *why = "num_envs must be >= 1"; return ZBOT_E_INVALID; }
  if (c.task != ZBOT_TASK_WALKING_V2 && c.task != ZBOT_TASK_SNAKE_V0 && c.task != ZBOT_TASK_WALKING_V4 &&
      c.task != ZBOT_TASK_WALKING_M) {
    *why = "unknown task"; return ZBOT_E_INVALID;
  }
  if (c.decimation != 4) { *why = "only decimation == 4 is supported (5-deep contact history)"; return ZBOT_E_INVALID; }
  if (c.num_terms < 0 || c.num_terms > ZBOT_MAX_TERMS) { *why = "num_terms out of range"; return ZBOT_E_INVALID; }
  for (int i = 0; i < c.num_terms; ++i)
    if (c.term_id[i] < 0 || c.term_id[i] >= NUM_TERM_IDS) { *why = "unknown reward term id"; return ZBOT_E_INVALID; }
  if (!(c.sim_dt > 0.f) || !(c.contact_ramp > 0.f)) { *why = "sim_dt and contact_ramp must be > 0"; return ZBOT_E_INVALID; }
  return ZBOT_OK;
}

inline void cfg_defaults(ZbotCfg& c, int num_envs) {
  memset(&c, 0, sizeof(c));
  c.abi_version = ZBOT_ABI_VERSION;
  c.task = ZBOT_TASK_WALKING_V2;
  c.num_envs = num_envs;
  c.decimation = 4;
  c.max_episode_length = 1000;
  c.sim_dt = 1.0f / 200.0f;
  c.termination_height = 0.22f;
  c.y_err_limit = 0.5f;
  c.terminated_penalty = 20.0f;
  c.contact_died_force = 1.0f;
  c.kp = 50.0f; c.kd = 5.0f; c.effort_limit = 20.0f;
  c.gravity = (float)model::GRAVITY;
  c.contact_alpha = 1000.0f; c.contact_erp = 0.2f; c.contact_vdep = 1.0f; c.contact_beta_max = 3000.0f;
  c.contact_mu = 1.0f; c.contact_ramp = 5.0e-4f; c.contact_vt_eps = 1.0e-6f; c.contact_margin = 0.02f;
  // …env_v2.py:190-206, dict order; weight * step_dt as a python (double) product (…:250-251)
  static const int ids[13] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};
  static const double wts[13] = {1.0, -2.0, -1.0, -1.0, -5.0, 5.0, -15.0, -0.1, -0.002, -10.0, -2.0, -2.0, 3.0};
  const double step_dt = 4 * (1 / 200.0);
  c.num_terms = 13;
  for (int i = 0; i < 13; ++i) { c.term_id[i] = ids[i]; c.term_weight[i] = (float)(wts[i] * step_dt); }
}

}  // namespace zbot
