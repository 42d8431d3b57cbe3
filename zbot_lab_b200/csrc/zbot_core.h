// zbot_core.h -- per-environment math of the zbot-6b-walking-v2 control step.
//
// One environment = one call chain on scalars held in registers.  The header is
// compiled by nvcc for the sm_100a kernels (T = float) and by g++ for the CPU port
// under oracle/ (T = float or double), so the same arithmetic can be checked on the
// build box (no GPU) against the independent float64 oracle (oracle/dyn_oracle.py).
//
// What it restates (reference = /root/reference/source/zbot/zbot/...):
//   * tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py:276-287  action -> joint target
//   * ...:384-411 terminations, :371-382 + :461-571 reward terms, :413-459 partial reset,
//     :312-369 observation (mdp_* functions below)
//   * assets/zbot_cfg.py:621-669 + zbot_6s_new.usd: the articulation that PhysX steps in the
//     reference.  PhysX is closed, so the dynamics here are OUR model:
//     7-body floating-base chain, articulated-body algorithm in world-aligned Pluecker
//     coordinates about the root origin, implicit joint PD folded into the joint-space
//     diagonal, linearly-implicit soft ground contact folded into the articulated inertia.
#pragma once
#include <math.h>
#include <stdint.h>

#include "zbot_model_constants.h"

#ifndef ZB_LINKKIN_NOINLINE
#define ZB_LINKKIN_NOINLINE 0   // 1 = link_kinematics as a real device function (smaller SASS).  Measured slower: the
#endif                          // by-reference structs go through a 680 B stack frame; 34.9 -> 37.0 us @4096, 86.4 -> 93.8 @65536
#ifndef ZB_PIPELINED_SWEEP
#define ZB_PIPELINED_SWEEP 1   // software-pipelined backward sweep (physics_substep); 0 = the plain loop
#endif
#if defined(__CUDACC__)
#define ZB_HD_NOINLINE __host__ __device__ __noinline__
#define ZB_HD __host__ __device__ __forceinline__
#define ZB_UNROLL _Pragma("unroll")
#else
#define ZB_HD_NOINLINE inline
#define ZB_HD inline __attribute__((always_inline))
#define ZB_UNROLL
#endif

#if ZB_LINKKIN_NOINLINE
#define ZB_LINKKIN_ATTR ZB_HD_NOINLINE
#else
#define ZB_LINKKIN_ATTR ZB_HD
#endif

namespace zbot {

// ------------------------------------------------------------------------------------
// reward-term ids (names = reference method suffixes, ...env_v2.py:461-571)
// ------------------------------------------------------------------------------------
enum TermId : int {
  TERM_BASE_VEL_FORWARD = 0,
  TERM_FEET_DOWNWARD = 1,
  TERM_FEET_FORWARD = 2,
  TERM_BASE_HEADING_X = 3,
  TERM_BASE_HEADING_X_SUM = 4,
  TERM_STEP_LENGTH = 5,
  TERM_AIRTIME_BALANCE = 6,
  TERM_ACTION_RATE = 7,
  TERM_TORQUES = 8,
  TERM_FEET_SLIDE = 9,
  TERM_BASE_POS_Y_ERR = 10,
  TERM_BASE_POS_Y_ERR_SUM = 11,
  TERM_AIRTIME_SUM = 12,
  TERM_FEET_FORCE_DIFF = 13,  // defined but inactive in v2's scale dict (:563-565)
  TERM_FEET_FORCE_SUM = 14,   // defined but inactive in v2's scale dict (:567-571)
  // snake task (tasks/zbot6_direct/zbot_direct_6dof_snake_v0.py:300-350); "base_vel_forward", "action_rate"
  // and "torques" share ids 0, 7, 8 with the walking task
  TERM_SNAKE_BASE_UP_Z = 15,
  TERM_SNAKE_BASE_HEADING_Y = 16,
  TERM_SNAKE_BASE_HEADING_Y_SUM = 17,
  TERM_SNAKE_BASE_POS_X_ERR = 18,
  TERM_SNAKE_BASE_POS_X_ERR_SUM = 19,  // defined but inactive in the snake scale dict
  // zbot-6b-walking-v4 (tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py:1013-1199); "feet_downward",
  // "feet_forward", "action_rate", "torques", "feet_slide" share ids 1, 2, 7, 8, 9 (same formulas, FRESH inputs)
  TERM_V4_TRACK_LIN_VEL_X = 20,
  TERM_V4_TRACK_HEADING_YAW = 21,
  TERM_V4_LIN_VEL_Y = 22,
  TERM_V4_JOINT_VEL = 23,
  TERM_V4_JOINT_ACC = 24,
  TERM_V4_STEP_LENGTH = 25,
  TERM_V4_FEET_AIR_TIME_BIPED = 26,
  TERM_V4_AIRTIME_VARIANCE = 27,
  TERM_V4_FEET_HARMONY = 28,
  TERM_V4_FEET_CLOSE = 29,
  TERM_V4_LIN_VEL_X = 30,      // defined, not in the default v4 table
  TERM_V4_AIRTIME_SUM = 31,    // defined, not in the default v4 table
  TERM_V4_FEET_HEIGHT = 32,    // defined, not in the default v4 table
  TERM_V4_BASE_HEIGHT = 33,    // defined, not in the default v4 table
  // zbot-6b-walking-m-v0 (tasks/zbotlab_manager/mdp/rewards.py + isaaclab.envs.mdp [IL-upstream]).  RewTerm functions with
  // a twin above reuse its id on FRESH inputs: joint_torques_l2 = 8, joint_acc_l2 = 24, action_rate_l2 = 7 (raw actions),
  // foot_downward = 1, foot_forward = 2, air_time_balance_penalty = 6, air_time_variance_penalty = 27
  TERM_M_TRACK_LIN_VEL_XY_EXP = 34,   // track_lin_vel_xy_yaw_frame_exp (rewards.py:286-297); par = {std^2}
  TERM_M_TRACK_ANG_VEL_Z_EXP = 35,    // track_ang_vel_z_world_exp (:300-309); par = {std^2}
  TERM_M_FOOT_STEP_LENGTH = 36,       // foot_step_length (:45-107), command_name = None
  TERM_M_GAIT = 37,                   // feet_gait (:155-186); par = {period, offset0, offset1, threshold}
  TERM_M_FEET_SLIDE = 38,             // feet_slide (:247-262)
  TERM_M_FEET_CLEARANCE = 39,         // foot_clearance_reward (:143-153); par = {std, tanh_mult, target_height}
  TERM_M_FEET_AIR_TIME_BIPED = 40,    // feet_air_time_positive_biped (:211-230); par = {threshold}
  TERM_M_BASE_VEL_FORWARD = 41,       // base_vel_forward (:264-274); par = {which_forward}
  TERM_M_FEET_FORCE_PATTERN = 42,     // feet_force_pattern (:276-284)
  TERM_M_UNDESIRED_CONTACTS = 43,     // isaaclab.envs.mdp.undesired_contacts [IL-upstream] (zbotlab_env_cfg.py:367-371); par = {threshold}
  NUM_TERM_IDS = 44
};
constexpr int MAX_TERMS = 16;

// ------------------------------------------------------------------------------------
// uniform parameters
// ------------------------------------------------------------------------------------
template <typename T>
struct Params {
  // dynamics
  T dt, kp, kd, effort, arm;  // arm = dt*kd + dt*dt*kp
  T gravity;
  T c_k, c_d, c_fcap, c_beta_max, c_mu, c_inv_ramp, c_vt_eps, c_margin, c_inv_fband;  // contact law (zbot_6s.py)
  int decimation;
  // MDP
  T step_dt;
  T termination_height, y_limit, term_penalty, contact_died_threshold;
  int max_episode_length;
  int default_terms;   // 1: the table is exactly the 13 v2 terms in dict order (fast path)
  int num_terms;
  int term_id[MAX_TERMS];
  T term_w[MAX_TERMS];  // weight * step_dt, rounded the way the reference rounds it (v4: the bare weight)
  // zbot-6b-walking-v4 event parameters (EventCfg, …env_v4.py:331-418; curriculum-adjusted by the host)
  T ev_vel_lo, ev_vel_hi, ev_yaw_lo, ev_yaw_hi, ev_offset, ev_prob_pos;
  int ev_dual_sign;
  T ev_pose_lo[3], ev_pose_hi[3];     // reset_base pose_range x, y, yaw
  T ev_interval_lo, ev_interval_hi;   // interval_range_s of interval_command_resample
  // additive uniform observation noise (ObservationManager `Unoise`); applied to the EMITTED row only
  int obs_noise_enable;
  T obs_noise_lo[24], obs_noise_w[24];   // lower bound, width (hi - lo)
  unsigned long long rng_seed;
  // zbot-6b-walking-m-v0 (manager-based task): RewTerm params, command term, action term, termination terms
  T term_par[MAX_TERMS][4];             // RewTerm `params` of the term in the same slot (meaning per TermId)
  T cmd_lo[3], cmd_hi[3];               // UniformVelocityCommandCfg.ranges lin_vel_x / lin_vel_y / ang_vel_z
  T cmd_rel_standing;                   // rel_standing_envs
  T cmd_resample_lo, cmd_resample_hi;   // resampling_time_range
  T act_scale, act_clip;                // RelativeJointPositionActionCfg scale, symmetric clip of the processed action
  T feet_close_min;                     // DoneTerm feet_close minimum_distance (<= 0: term absent)
  T term_penalty_w;                     // RewTerm is_terminated: weight * step_dt (0: term absent)
  T illegal_thr;                        // DoneTerm illegal_contact threshold (zbotlab_env_cfg.py:385-388; <= 0: term absent)
  int illegal_mask;                     // ... over the merged bodies 1..5 (bit b-1)
  int cmd_heading;                      // UniformVelocityCommandCfg.heading_command
  T cmd_heading_lo, cmd_heading_hi, cmd_heading_stiffness, cmd_rel_heading;   // ranges.heading, heading_control_stiffness, rel_heading_envs
  T push_interval_lo, push_interval_hi; // EventTerm push_robot interval_range_s (<= 0: term absent; zbotlab_env_cfg.py:253-258)
  T push_lo[2], push_hi[2];             // push_by_setting_velocity velocity_range x / y
};

// ------------------------------------------------------------------------------------
// math shims
// ------------------------------------------------------------------------------------
ZB_HD float zb_sqrt(float x) { return sqrtf(x); }
ZB_HD double zb_sqrt(double x) { return sqrt(x); }
ZB_HD float zb_tanh(float x) { return tanhf(x); }
ZB_HD double zb_tanh(double x) { return tanh(x); }
ZB_HD float zb_exp(float x) { return expf(x); }
ZB_HD double zb_exp(double x) { return exp(x); }
ZB_HD float zb_atan2(float y, float x) { return atan2f(y, x); }
ZB_HD double zb_atan2(double y, double x) { return atan2(y, x); }
ZB_HD float zb_sin(float x) { return sinf(x); }
ZB_HD double zb_sin(double x) { return sin(x); }
ZB_HD float zb_cos(float x) { return cosf(x); }
ZB_HD double zb_cos(double x) { return cos(x); }
ZB_HD float zb_fmod(float a, float b) { return fmodf(a, b); }
ZB_HD double zb_fmod(double a, double b) { return fmod(a, b); }
ZB_HD float zb_abs(float x) { return fabsf(x); }
ZB_HD double zb_abs(double x) { return fabs(x); }
ZB_HD float zb_min(float a, float b) { return fminf(a, b); }
ZB_HD double zb_min(double a, double b) { return fmin(a, b); }
ZB_HD float zb_max(float a, float b) { return fmaxf(a, b); }
ZB_HD double zb_max(double a, double b) { return fmax(a, b); }
// sin/cos of a joint half-angle (|x| < ~8 rad): Cody-Waite reduction to [-pi/4, pi/4] + the classic
// single-precision minimax polynomials (error ~1 ulp).  ~25 instructions, no slow path, no local
// memory -- CUDA's sincosf() carries a Payne-Hanek fallback that cost ~1000 SASS instructions and a
// stack frame per inlined call site.
ZB_HD void zb_sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  const float k = rintf(x * 0.63661977236758134f);
  const int q = (int)k;
  float r = fmaf(k, -1.57079601287841796875f, x);
  r = fmaf(k, -3.1391647326017846353352069854736328125e-7f, r);
  r = fmaf(k, -5.3903025299577647655e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp, r2, -0.5f);
  cp = fmaf(cp, r2, 1.0f);
  const float ss = (q & 1) ? cp : sp;
  const float cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
#else
  *s = sinf(x);
  *c = cosf(x);
#endif
}
ZB_HD void zb_sincos(double x, double* s, double* c) {
  *s = sin(x);
  *c = cos(x);
}
// reciprocal / reciprocal square root for the DYNAMICS only (float on the GPU: one MUFU + one multiply,
// ~2 ulp; the MDP terms keep IEEE division and sqrt for the 1e-5 parity with the reference)
ZB_HD float zb_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __fdividef(1.0f, x);
#else
  return 1.0f / x;
#endif
}
ZB_HD double zb_rcp(double x) { return 1.0 / x; }
ZB_HD float zb_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
ZB_HD double zb_rsqrt(double x) { return 1.0 / sqrt(x); }

// predicate helpers: generic code that must also run with T = F2 (two lanes, zbot_pair.h) uses these instead of `if`
ZB_HD bool zb_gt(float a, float b) { return a > b; }
ZB_HD bool zb_gt(double a, double b) { return a > b; }
ZB_HD bool zb_any(bool m) { return m; }
ZB_HD int zb_min_i(int a, int b) { return a < b ? a : b; }
ZB_HD float zb_sel(bool m, float a, float b) { return m ? a : b; }
ZB_HD double zb_sel(bool m, double a, double b) { return m ? a : b; }

template <typename T>
ZB_HD T zb_clamp(T x, T lo, T hi) {
  return zb_min(zb_max(x, lo), hi);
}

// PhysX keeps the position of a revolute joint WITHOUT limits inside [-2 pi, 2 pi]: past either end it re-enters from
// the other one, i.e. shifts by 4 pi (two full turns: the same physical angle, and the same half-angle quaternion).
// The reference's author checked it on these robots: "joint_pos is normalised into [-2pi, 2pi] by default even without
// joint limits; only with limits set is it kept inside the limits" (assets/test_articulation.py:18-20).  Targets are
// p_delta (clipped to +-pi) + the default pose, |target| <= 5.16 rad, so a tracked joint never gets there; a free-spinning
// one (saturated drive after a fall) does, and then the PD error jumps by 4 pi exactly as it does in the reference.
template <typename T>
ZB_HD T zb_wrap_joint(T q) {
  const T two_pi = T(6.283185307179586476925286766559);
  const T four_pi = T(12.566370614359172953850573533118);
  return zb_sel(zb_gt(q, two_pi), q - four_pi, zb_sel(zb_gt(-two_pi, q), q + four_pi, q));
}

template <typename T>
ZB_HD void cross3(const T* a, const T* b, T* o) {
  T x = a[1] * b[2] - a[2] * b[1];
  T y = a[2] * b[0] - a[0] * b[2];
  T z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <typename T>
ZB_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// isaaclab.utils.math.quat_apply: wxyz; t = 2 (q_xyz x v); v + q_w t + q_xyz x t
template <typename T>
ZB_HD void quat_apply(const T* q, const T* v, T* o) {
  T t[3], u[3];
  cross3(q + 1, v, t);
  t[0] *= T(2); t[1] *= T(2); t[2] *= T(2);
  cross3(q + 1, t, u);
  o[0] = v[0] + q[0] * t[0] + u[0];
  o[1] = v[1] + q[0] * t[1] + u[1];
  o[2] = v[2] + q[0] * t[2] + u[2];
}

// rotation matrix (row-major R[3*r+c]) of a unit quaternion wxyz
template <typename T>
ZB_HD void quat_to_mat(const T* q, T* R) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  T xx = x * x, yy = y * y, zz = z * z;
  T xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  R[0] = T(1) - T(2) * (yy + zz); R[1] = T(2) * (xy - wz);        R[2] = T(2) * (xz + wy);
  R[3] = T(2) * (xy + wz);        R[4] = T(1) - T(2) * (xx + zz); R[5] = T(2) * (yz - wx);
  R[6] = T(2) * (xz - wy);        R[7] = T(2) * (yz + wx);        R[8] = T(1) - T(2) * (xx + yy);
}

// Q <- Q (x) (c, sx*s, 0, sz*s): child orientation after a revolute joint about (sx, 0, sz)
template <typename T>
ZB_HD void quat_mul_joint(T* Q, T c, T jx, T jz) {
  T w1 = Q[0], x1 = Q[1], y1 = Q[2], z1 = Q[3];
  Q[0] = w1 * c - x1 * jx - z1 * jz;
  Q[1] = w1 * jx + x1 * c + y1 * jz;
  Q[2] = -x1 * jz + y1 * c + z1 * jx;
  Q[3] = w1 * jz - y1 * jx + z1 * c;
}

// ------------------------------------------------------------------------------------
// simulation state of one environment (root pose is env-LOCAL: relative to the env origin)
// ------------------------------------------------------------------------------------
template <typename T>
struct SimState {
  T p[3];   // root (foot_0 link origin) position
  T Q[4];   // root orientation wxyz
  T v[3];   // root origin linear velocity (world frame)
  T w[3];   // root angular velocity (world frame)
  T q[6];   // joint positions
  T qd[6];  // joint velocities
};


// ------------------------------------------------------------------------------------
// symmetric 6x6 articulated inertia  [[I, H], [H^T, M]]  (motion = (w; vO), force = (n; f))
// ------------------------------------------------------------------------------------
template <typename T>
struct SpInertia {
  T I[6];  // xx xy xz yy yz zz
  T H[9];  // row-major
  T M[6];  // xx xy xz yy yz zz
};

template <typename T>
ZB_HD void sym3_mul(const T* S, const T* v, T* o) {
  o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
  o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
  o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}

// (top; bot) = IA * (a; m)
template <typename T>
ZB_HD void spi_mul(const SpInertia<T>& A, const T* a, const T* m, T* top, T* bot) {
  T t0[3], t1[3];
  sym3_mul(A.I, a, t0);
  top[0] = t0[0] + A.H[0] * m[0] + A.H[1] * m[1] + A.H[2] * m[2];
  top[1] = t0[1] + A.H[3] * m[0] + A.H[4] * m[1] + A.H[5] * m[2];
  top[2] = t0[2] + A.H[6] * m[0] + A.H[7] * m[1] + A.H[8] * m[2];
  sym3_mul(A.M, m, t1);
  bot[0] = t1[0] + A.H[0] * a[0] + A.H[3] * a[1] + A.H[6] * a[2];
  bot[1] = t1[1] + A.H[1] * a[0] + A.H[4] * a[1] + A.H[7] * a[2];
  bot[2] = t1[2] + A.H[2] * a[0] + A.H[5] * a[1] + A.H[8] * a[2];
}

// A -= s * (ut; ub)(ut; ub)^T
template <typename T>
ZB_HD void spi_rank1_sub(SpInertia<T>& A, const T* ut, const T* ub, T s) {
  T st[3] = {s * ut[0], s * ut[1], s * ut[2]};
  T sb[3] = {s * ub[0], s * ub[1], s * ub[2]};
  A.I[0] -= st[0] * ut[0]; A.I[1] -= st[0] * ut[1]; A.I[2] -= st[0] * ut[2];
  A.I[3] -= st[1] * ut[1]; A.I[4] -= st[1] * ut[2]; A.I[5] -= st[2] * ut[2];
  ZB_UNROLL for (int r = 0; r < 3; ++r)
    ZB_UNROLL for (int c = 0; c < 3; ++c) A.H[3 * r + c] -= st[r] * ub[c];
  A.M[0] -= sb[0] * ub[0]; A.M[1] -= sb[0] * ub[1]; A.M[2] -= sb[0] * ub[2];
  A.M[3] -= sb[1] * ub[1]; A.M[4] -= sb[1] * ub[2]; A.M[5] -= sb[2] * ub[2];
}

// A += rigid body (mass m, world inertia about CoM Iw[6], CoM at c relative to O)
template <typename T>
ZB_HD void spi_add_rigid(SpInertia<T>& A, T m, const T* Iw, const T* c) {
  T hx = m * c[0], hy = m * c[1], hz = m * c[2];
  A.I[0] += Iw[0] + (hy * c[1] + hz * c[2]);
  A.I[1] += Iw[1] - hx * c[1];
  A.I[2] += Iw[2] - hx * c[2];
  A.I[3] += Iw[3] + (hx * c[0] + hz * c[2]);
  A.I[4] += Iw[4] - hy * c[2];
  A.I[5] += Iw[5] + (hx * c[0] + hy * c[1]);
  // H = skew(h)
  A.H[1] -= hz; A.H[2] += hy;
  A.H[3] += hz; A.H[5] -= hx;
  A.H[6] -= hy; A.H[7] += hx;
  A.M[0] += m; A.M[3] += m; A.M[5] += m;
}

// A += k * g g^T with g = (rho x e_j ; e_j), e_j the j-th world axis (sparse rank-1)
template <typename T>
ZB_HD void spi_add_contact(SpInertia<T>& A, const T* r, T kx, T ky, T kz) {
  // g_x = ((0, rz, -ry); (1,0,0))
  A.I[3] += kx * r[2] * r[2]; A.I[4] -= kx * r[2] * r[1]; A.I[5] += kx * r[1] * r[1];
  A.H[3] += kx * r[2]; A.H[6] -= kx * r[1]; A.M[0] += kx;
  // g_y = ((-rz, 0, rx); (0,1,0))
  A.I[0] += ky * r[2] * r[2]; A.I[2] -= ky * r[2] * r[0]; A.I[5] += ky * r[0] * r[0];
  A.H[1] -= ky * r[2]; A.H[7] += ky * r[0]; A.M[3] += ky;
  // g_z = ((ry, -rx, 0); (0,0,1))
  A.I[0] += kz * r[1] * r[1]; A.I[1] -= kz * r[1] * r[0]; A.I[3] += kz * r[0] * r[0];
  A.H[2] += kz * r[1]; A.H[5] -= kz * r[0]; A.M[5] += kz;
}

// Solve the SPD system  A x = b  for the floating base (6x6 LDL^T, fully unrolled).
template <typename T>
ZB_HD void spi_solve(const SpInertia<T>& A, const T* bt, const T* bb, T* xt, T* xb) {
  T a[6][6];
  a[0][0] = A.I[0]; a[1][0] = A.I[1]; a[2][0] = A.I[2]; a[1][1] = A.I[3]; a[2][1] = A.I[4]; a[2][2] = A.I[5];
  ZB_UNROLL for (int r = 0; r < 3; ++r)
    ZB_UNROLL for (int c = 0; c < 3; ++c) a[3 + c][r] = A.H[3 * r + c];  // lower block = H^T
  a[3][3] = A.M[0]; a[4][3] = A.M[1]; a[5][3] = A.M[2]; a[4][4] = A.M[3]; a[5][4] = A.M[4]; a[5][5] = A.M[5];
  T x[6] = {bt[0], bt[1], bt[2], bb[0], bb[1], bb[2]};
  T dinv[6];
  ZB_UNROLL for (int j = 0; j < 6; ++j) {
    T d = a[j][j];
    ZB_UNROLL for (int k = 0; k < j; ++k) d -= a[j][k] * a[j][k] * a[k][k];
    a[j][j] = d;  // D_j
    dinv[j] = zb_rcp(d);
    ZB_UNROLL for (int i = j + 1; i < 6; ++i) {
      T s = a[i][j];
      ZB_UNROLL for (int k = 0; k < j; ++k) s -= a[i][k] * a[j][k] * a[k][k];
      a[i][j] = s * dinv[j];  // L_ij
    }
  }
  ZB_UNROLL for (int i = 0; i < 6; ++i)
    ZB_UNROLL for (int k = 0; k < i; ++k) x[i] -= a[i][k] * x[k];
  ZB_UNROLL for (int i = 0; i < 6; ++i) x[i] *= dinv[i];
  ZB_UNROLL for (int i = 5; i >= 0; --i)
    ZB_UNROLL for (int k = i + 1; k < 6; ++k) x[i] -= a[k][i] * x[k];
  xt[0] = x[0]; xt[1] = x[1]; xt[2] = x[2];
  xb[0] = x[3]; xb[1] = x[4]; xb[2] = x[5];
}

// ------------------------------------------------------------------------------------
// contact: one candidate point at rho (relative to O) on a body with spatial velocity (w; vO)
// ------------------------------------------------------------------------------------
template <typename T>
struct ContactAgg {  // enough to evaluate  sum_c (F0_c - dt K_c (A_lin + A_ang x rho_c))  later
  T F0[3];
  T sb, sg;         // sum beta, sum gamma
  T sbr[3];         // sum beta * rho
  T sgrx, sgry;     // sum gamma * rho_x, rho_y
};
template <typename T>
ZB_HD void contact_agg_zero(ContactAgg<T>& g) {
  g.F0[0] = g.F0[1] = g.F0[2] = T(0);
  g.sb = g.sg = T(0);
  g.sbr[0] = g.sbr[1] = g.sbr[2] = T(0);
  g.sgrx = g.sgry = T(0);
}
template <typename T>
ZB_HD void contact_agg_force(const ContactAgg<T>& g, T dt, const T* At, const T* Ab, T* f) {
  // K (A_lin + A_ang x rho), K = diag(beta, beta, gamma)
  f[0] = g.F0[0] - dt * (g.sb * Ab[0] + At[1] * g.sbr[2] - At[2] * g.sbr[1]);
  f[1] = g.F0[1] - dt * (g.sb * Ab[1] + At[2] * g.sbr[0] - At[0] * g.sbr[2]);
  f[2] = g.F0[2] - dt * (g.sg * Ab[2] + At[0] * g.sgry - At[1] * g.sgrx);
}

// Adds dt*J^T K J to IA, subtracts J^T F0 from pA.  Branch-free per lane (T may be F2: two environments): a lane
// outside the speculative margin, or below the activation band, contributes exact zeros; the early-outs only fire when
// EVERY lane is out.  PS = scalar type of the uniform parameters.
template <typename PS, typename T>
ZB_HD void contact_point(const Params<PS>& P, T mu, const T* rho, T height, const T* w, const T* vO,
                         SpInertia<T>& IA, T* pAt, T* pAb, ContactAgg<T>* agg, T* f0out) {
  const T pen = -height;
  const auto in_margin = zb_gt(pen, T(-P.c_margin));   // speculative margin: points this far above ground are skipped
  if (!zb_any(in_margin)) return;
  T vp[3], wxv[3];
  cross3(w, rho, vp);
  vp[0] += vO[0]; vp[1] += vO[1]; vp[2] += vO[2];
  cross3(w, vp, wxv);
  const T dt = T(P.dt);
  T sx = vp[0] + dt * wxv[0], sy = vp[1] + dt * wxv[1], sz = vp[2] + dt * wxv[2];
  T s = zb_clamp(pen * T(P.c_inv_ramp), T(0), T(1));
  T fs = zb_min(T(P.c_k) * pen, T(P.c_fcap));
  T gamma = T(P.c_k * P.dt) + T(P.c_d) * s;
  T fn0 = fs - gamma * sz;
  // Activation is CONTINUOUS in the state: the implicit normal stiffness ramps in over a band of predictor
  // force [-fband, 0] (there F0 = 0 and only -dt*act*gamma*a_n acts: an added-mass term, PSD), so a point
  // sitting exactly on the threshold (a snake at rest on the plane: all 12 spheres at pen = 0) does not flip a
  // coin on float round-off.  Full stiffness for fn0 >= 0 as before.
  T act = zb_clamp(fn0 * T(P.c_inv_fband) + T(1), T(0), T(1));
  act = zb_sel(in_margin, act, T(0));
  if (!zb_any(zb_gt(act, T(0)))) return;
  fn0 = zb_sel(in_margin, zb_max(fn0, T(0)), T(0));
  gamma *= act;
  // beta = min(beta_max, mu fn0 / max(|v_t|, eps))
  T beta = zb_min(T(P.c_beta_max), mu * fn0 * zb_rsqrt(zb_max(sx * sx + sy * sy, T(P.c_vt_eps * P.c_vt_eps))));
  T F0[3] = {-(beta * sx), -(beta * sy), fn0};
  T n[3];
  cross3(rho, F0, n);
  pAt[0] -= n[0]; pAt[1] -= n[1]; pAt[2] -= n[2];
  pAb[0] -= F0[0]; pAb[1] -= F0[1]; pAb[2] -= F0[2];
  spi_add_contact(IA, rho, dt * beta, dt * beta, dt * gamma);
  if (agg) {
    agg->F0[0] += F0[0]; agg->F0[1] += F0[1]; agg->F0[2] += F0[2];
    agg->sb += beta; agg->sg += gamma;
    agg->sbr[0] += beta * rho[0]; agg->sbr[1] += beta * rho[1]; agg->sbr[2] += beta * rho[2];
    agg->sgrx += gamma * rho[0]; agg->sgry += gamma * rho[1];
  }
  if (f0out) { f0out[0] = F0[0]; f0out[1] = F0[1]; f0out[2] = F0[2]; }
}

// ------------------------------------------------------------------------------------
// per-body rigid terms: adds the body's spatial inertia about O and its bias force
//   p = V x* (I V) - f_gravity
// ------------------------------------------------------------------------------------
template <typename PS, typename T>
ZB_HD void body_rigid_terms(const Params<PS>& P, T mass, T cx, T cz, T ixx, T iyy, T izz, T ixz,
                            const T* R, const T* r, const T* w, const T* vO,
                            SpInertia<T>& IA, T* pAt, T* pAb) {
  // CoM relative to O (body CoM_y == 0)
  T c[3] = {r[0] + R[0] * cx + R[2] * cz, r[1] + R[3] * cx + R[5] * cz, r[2] + R[6] * cx + R[8] * cz};
  // Iw = R diag/sparse R^T
  T t0[3] = {ixx * R[0] + ixz * R[2], ixx * R[3] + ixz * R[5], ixx * R[6] + ixz * R[8]};
  T t1[3] = {iyy * R[1], iyy * R[4], iyy * R[7]};
  T t2[3] = {ixz * R[0] + izz * R[2], ixz * R[3] + izz * R[5], ixz * R[6] + izz * R[8]};
  T Iw[6];
  Iw[0] = t0[0] * R[0] + t1[0] * R[1] + t2[0] * R[2];
  Iw[1] = t0[0] * R[3] + t1[0] * R[4] + t2[0] * R[5];
  Iw[2] = t0[0] * R[6] + t1[0] * R[7] + t2[0] * R[8];
  Iw[3] = t0[1] * R[3] + t1[1] * R[4] + t2[1] * R[5];
  Iw[4] = t0[1] * R[6] + t1[1] * R[7] + t2[1] * R[8];
  Iw[5] = t0[2] * R[6] + t1[2] * R[7] + t2[2] * R[8];
  // momentum about O:  L = Iw w + c x (m vc),  Pl = m vc,  vc = vO + w x c
  T wxc[3];
  cross3(w, c, wxc);
  T Pl[3] = {mass * (vO[0] + wxc[0]), mass * (vO[1] + wxc[1]), mass * (vO[2] + wxc[2])};
  T L[3], cxP[3];
  sym3_mul(Iw, w, L);
  cross3(c, Pl, cxP);
  L[0] += cxP[0]; L[1] += cxP[1]; L[2] += cxP[2];
  // p = (w x L + vO x Pl ; w x Pl) - (c x m g ; m g),  g = (0,0,-G)
  T a0[3], a1[3], a2[3];
  cross3(w, L, a0);
  cross3(vO, Pl, a1);
  cross3(w, Pl, a2);
  T mg = mass * T(P.gravity);
  pAt[0] += a0[0] + a1[0] + mg * c[1];
  pAt[1] += a0[1] + a1[1] - mg * c[0];
  pAt[2] += a0[2] + a1[2];
  pAb[0] += a2[0];
  pAb[1] += a2[1];
  pAb[2] += a2[2] + mg;
  spi_add_rigid(IA, mass, Iw, c);
}

// Same terms for a body with a full CoM vector c_b and a full symmetric inertia Ib = (xx, xy, xz, yy, yz, zz) in its body
// frame (the manager-based biped, whose authored mass properties sit off the chain's plane).
template <typename PS, typename T>
ZB_HD void body_rigid_terms_full(const Params<PS>& P, T mass, const T* cb, const T* Ib,
                                 const T* R, const T* r, const T* w, const T* vO,
                                 SpInertia<T>& IA, T* pAt, T* pAb) {
  T c[3] = {r[0] + R[0] * cb[0] + R[1] * cb[1] + R[2] * cb[2], r[1] + R[3] * cb[0] + R[4] * cb[1] + R[5] * cb[2],
            r[2] + R[6] * cb[0] + R[7] * cb[1] + R[8] * cb[2]};
  // t_i = Ib * (row i of R)  ->  Iw = R Ib R^T
  T t0[3], t1[3], t2[3];
  sym3_mul(Ib, R + 0, t0);
  sym3_mul(Ib, R + 3, t1);
  sym3_mul(Ib, R + 6, t2);
  T Iw[6];
  Iw[0] = R[0] * t0[0] + R[1] * t0[1] + R[2] * t0[2];
  Iw[1] = R[0] * t1[0] + R[1] * t1[1] + R[2] * t1[2];
  Iw[2] = R[0] * t2[0] + R[1] * t2[1] + R[2] * t2[2];
  Iw[3] = R[3] * t1[0] + R[4] * t1[1] + R[5] * t1[2];
  Iw[4] = R[3] * t2[0] + R[4] * t2[1] + R[5] * t2[2];
  Iw[5] = R[6] * t2[0] + R[7] * t2[1] + R[8] * t2[2];
  T wxc[3];
  cross3(w, c, wxc);
  T Pl[3] = {mass * (vO[0] + wxc[0]), mass * (vO[1] + wxc[1]), mass * (vO[2] + wxc[2])};
  T L[3], cxP[3];
  sym3_mul(Iw, w, L);
  cross3(c, Pl, cxP);
  L[0] += cxP[0]; L[1] += cxP[1]; L[2] += cxP[2];
  T a0[3], a1[3], a2[3];
  cross3(w, L, a0);
  cross3(vO, Pl, a1);
  cross3(w, Pl, a2);
  T mg = mass * T(P.gravity);
  pAt[0] += a0[0] + a1[0] + mg * c[1];
  pAt[1] += a0[1] + a1[1] - mg * c[0];
  pAt[2] += a0[2] + a1[2];
  pAb[0] += a2[0];
  pAb[1] += a2[1];
  pAb[2] += a2[2] + mg;
  spi_add_rigid(IA, mass, Iw, c);
}

// rigid terms of body k of `Model` (sparse xz-symmetric table, or the full one)
template <typename Model, typename PS, typename T>
ZB_HD void model_body_terms(const Params<PS>& P, int k, const T* R, const T* r, const T* w, const T* vO,
                            SpInertia<T>& IA, T* pAt, T* pAb) {
  if constexpr (Model::kFullInertia) {
    T mass, cb[3], Ib[6];
    Model::body_full(k, mass, cb, Ib);
    body_rigid_terms_full(P, mass, cb, Ib, R, r, w, vO, IA, pAt, pAb);
  } else {
    T mass, cx, cz, ixx, iyy, izz, ixz;
    Model::body(k, mass, cx, cz, ixx, iyy, izz, ixz);
    body_rigid_terms(P, mass, cx, cz, ixx, iyy, izz, ixz, R, r, w, vO, IA, pAt, pAb);
  }
}

// ------------------------------------------------------------------------------------
// robot models: per-body mass tables, ground-contact candidate points, default pose.  The chain geometry
// (joint positions / axes in the body frames) is common to both; a Model only supplies constants, selected
// with compile-time-literal ternaries so they fold into FFMA immediates.
// ------------------------------------------------------------------------------------
struct ModelWalk {   // ZBOT_6S_CFG + zbot_6s_new.usd (zbot-6b-walking-*): stands on two foot discs
  static constexpr int kTask = 0;
  static constexpr bool kPerBodyForces = false;    // the direct tasks only need the max over the merged bodies
  static constexpr bool kFullInertia = false;      // bodies symmetric about the chain's xz plane
  static constexpr bool kPerEnvFriction = false;
  static constexpr bool kFresh = false;            // MDP reads one-step-stale quantities, 5-deep force history
  static constexpr bool kGroundForceSensor = true;   // feet: applied force; merged bodies: predictor force
  static ZB_HD bool wraps(int) { return true; }      // zbot_6s_new.usd: no joint has limits -> PhysX wraps all six at +-2 pi
  template <typename T>
  static ZB_HD void body(int k, T& mass, T& cx, T& cz, T& ixx, T& iyy, T& izz, T& ixz) {
    using namespace model;
    const bool foot0 = (k == 0), foot1 = (k == 6);
    mass = (foot0 || foot1) ? T(FOOT0_MASS) : T(MID_MASS);
    cx = foot0 ? T(FOOT0_COM_X) : foot1 ? T(FOOT1_COM_X) : T(MID_COM_X);
    cz = foot0 ? T(FOOT0_COM_Z) : foot1 ? T(FOOT1_COM_Z) : T(MID_COM_Z);
    ixx = foot0 ? T(FOOT0_IXX) : foot1 ? T(FOOT1_IXX) : T(MID_IXX);
    iyy = foot0 ? T(FOOT0_IYY) : foot1 ? T(FOOT1_IYY) : T(MID_IYY);
    izz = foot0 ? T(FOOT0_IZZ) : foot1 ? T(FOOT1_IZZ) : T(MID_IZZ);
    ixz = foot0 ? T(FOOT0_IXZ) : foot1 ? T(FOOT1_IXZ) : T(MID_IXZ);
  }
  static ZB_HD int npts(int k) { return (k == 0 || k == 6) ? 4 : 1; }
  // 4 rim points of a foot sole, or the bottom of a merged body's sphere (drop = world-z offset)
  template <typename T>
  static ZB_HD void point(int k, int j, T& lx, T& ly, T& lz, T& drop) {
    using namespace model;
    const bool foot0 = (k == 0), foot1 = (k == 6), foot = foot0 || foot1;
    lx = !foot ? T(0) : (j == 0) ? T(FOOT_R) : (j == 2) ? T(-FOOT_R) : T(0);
    ly = !foot ? T(0) : (j == 1) ? T(FOOT_R) : (j == 3) ? T(-FOOT_R) : T(0);
    lz = foot0 ? T(FOOT0_SOLE_Z) : foot1 ? T(FOOT1_SOLE_Z) : T(SPHERE_Z);
    drop = foot ? T(0) : T(SPHERE_R);
  }
  template <typename T>
  static ZB_HD T default_q(int k) {
    using namespace model;
    return k == 0 ? T(DQ0) : k == 1 ? T(DQ1) : k == 2 ? T(DQ2) : k == 3 ? T(DQ3) : k == 4 ? T(DQ4) : T(DQ5);
  }
  template <typename T>
  static ZB_HD void default_root(T* p, T* Q) {
    p[0] = T(model::DEFAULT_ROOT_X); p[1] = T(model::DEFAULT_ROOT_Y); p[2] = T(model::DEFAULT_ROOT_Z);
    Q[0] = T(1); Q[1] = T(0); Q[2] = T(0); Q[3] = T(0);
  }
};

struct ModelWalkV4 : ModelWalk {   // zbot-6b-walking-v4: same robot, 3-deep force history, FRESH MDP inputs
  static constexpr int kTask = 2;
  static constexpr bool kFresh = true;             // MDP reads end-of-physics quantities, 3-deep force history
};

struct ModelSnake {  // ZBOT_D_6S_CFG + zbot_6s_v03.usd (zbot-6s-snake-v0): lies on the ground
  static constexpr int kTask = 1;
  static constexpr bool kPerBodyForces = false;
  static constexpr bool kFullInertia = false;
  static constexpr bool kPerEnvFriction = false;
  static constexpr bool kFresh = false;
  static constexpr bool kGroundForceSensor = false;  // the task only senses filtered SELF contacts
  static ZB_HD bool wraps(int k) { return k != 5; }  // zbot_6s_v03.usd: only joint6 has limits (+-720 deg): it is not wrapped
  template <typename T>
  static ZB_HD void body(int k, T& mass, T& cx, T& cz, T& ixx, T& iyy, T& izz, T& ixz) {
    using namespace model_snake;
    const bool e0 = (k == 0), e6 = (k == 6), odd = (k & 1);
    mass = (e0 || e6) ? T(A1_MASS) : T(MIDO_MASS);
    cx = e0 ? T(A1_COM_X) : e6 ? T(B6_COM_X) : odd ? T(MIDO_COM_X) : T(MIDE_COM_X);
    cz = e0 ? T(A1_COM_Z) : e6 ? T(B6_COM_Z) : T(MIDO_COM_Z);
    ixx = e0 ? T(A1_IXX) : e6 ? T(B6_IXX) : T(MIDO_IXX);
    iyy = e0 ? T(A1_IYY) : e6 ? T(B6_IYY) : T(MIDO_IYY);
    izz = e0 ? T(A1_IZZ) : e6 ? T(B6_IZZ) : T(MIDO_IZZ);
    ixz = e0 ? T(A1_IXZ) : e6 ? T(B6_IXZ) : odd ? T(MIDO_IXZ) : T(MIDE_IXZ);
  }
  // a line of r = 0.05 spheres on the chain axis: body origin (joint centre) and, except for a1, z = 0.053
  static ZB_HD int npts(int k) { return (k == 0) ? 1 : 2; }
  template <typename T>
  static ZB_HD void point(int k, int j, T& lx, T& ly, T& lz, T& drop) {
    (void)k;
    lx = T(0); ly = T(0);
    lz = (j == 0) ? T(0) : T(model_snake::LINK_Z);
    drop = T(model_snake::SPHERE_R);
  }
  template <typename T>
  static ZB_HD T default_q(int) { return T(0); }
  template <typename T>
  static ZB_HD void default_root(T* p, T* Q) {
    using namespace model_snake;
    p[0] = T(DEFAULT_ROOT_X); p[1] = T(DEFAULT_ROOT_Y); p[2] = T(DEFAULT_ROOT_Z);
    Q[0] = T(DEFAULT_ROOT_QW); Q[1] = T(DEFAULT_ROOT_QX); Q[2] = T(DEFAULT_ROOT_QY); Q[3] = T(DEFAULT_ROOT_QZ);
  }
};

struct ModelWalkM : ModelWalk {   // ZBOT_6S_V2_CFG + zbot_6s_v09.usd (zbot-6b-walking-m-v0): same chain, authored mass tables
  static constexpr int kTask = 3;
  static constexpr bool kFullInertia = true;       // CoM / inertia sit off the chain's plane (assets/zbot_6s_v2.py)
  static constexpr bool kPerEnvFriction = true;    // EventCfg.physics_material: per-env friction drawn at startup
  static constexpr bool kPerBodyForces = true;     // undesired_contacts / illegal_contact need |F| of every merged body
  static constexpr bool kFresh = true;             // ManagerBasedRLEnv: every term reads end-of-physics data; history_length = 3
  static ZB_HD bool wraps(int) { return false; }   // zbot_6s_v09.usd: all six joints carry +-360 deg limits (hard stops in PhysX; not modelled)
  template <typename T>
  static ZB_HD void body_full(int k, T& mass, T* c, T* I) {
    using namespace model_m;
    const bool f0 = (k == 0), f1 = (k == 6), b3 = (k == 3), lo = (k < 3);   // bodies 1,2 = MIDA, 4,5 = MIDB
    mass = (f0 || f1) ? T(FOOT0_MASS) : T(MIDA_MASS);
    c[0] = f0 ? T(FOOT0_CX) : f1 ? T(FOOT1_CX) : b3 ? T(BASE_CX) : T(MIDA_CX);
    c[1] = f0 ? T(FOOT0_CY) : f1 ? T(FOOT1_CY) : b3 ? T(BASE_CY) : T(MIDA_CY);
    c[2] = f0 ? T(FOOT0_CZ) : f1 ? T(FOOT1_CZ) : b3 ? T(BASE_CZ) : lo ? T(MIDA_CZ) : T(MIDB_CZ);
    I[0] = (f0 || f1) ? T(FOOT0_IXX) : b3 ? T(BASE_IXX) : T(MIDA_IXX);
    I[1] = (f0 || f1) ? T(FOOT0_IXY) : b3 ? T(BASE_IXY) : T(MIDA_IXY);
    I[2] = (f0 || f1) ? T(FOOT0_IXZ) : b3 ? T(BASE_IXZ) : lo ? T(MIDA_IXZ) : T(MIDB_IXZ);
    I[3] = (f0 || f1) ? T(FOOT0_IYY) : b3 ? T(BASE_IYY) : T(MIDA_IYY);
    I[4] = (f0 || f1) ? T(FOOT0_IYZ) : b3 ? T(BASE_IYZ) : lo ? T(MIDA_IYZ) : T(MIDB_IYZ);
    I[5] = (f0 || f1) ? T(FOOT0_IZZ) : b3 ? T(BASE_IZZ) : T(MIDA_IZZ);
  }
  template <typename T>
  static ZB_HD T default_q(int k) {
    using namespace model_m;
    return k == 0 ? T(DQ0) : k == 1 ? T(DQ1) : k == 2 ? T(DQ2) : k == 3 ? T(DQ3) : k == 4 ? T(DQ4) : T(DQ5);
  }
  template <typename T>
  static ZB_HD void default_root(T* p, T* Q) {
    using namespace model_m;
    p[0] = T(DEFAULT_ROOT_X); p[1] = T(DEFAULT_ROOT_Y); p[2] = T(DEFAULT_ROOT_Z);
    Q[0] = T(DEFAULT_ROOT_QW); Q[1] = T(DEFAULT_ROOT_QX); Q[2] = T(DEFAULT_ROOT_QY); Q[3] = T(DEFAULT_ROOT_QZ);
  }
  // Isaac Lab joint order (joint1, joint7, joint2, joint8, joint3, joint9) <-> chain joints (joint3, 2, 1, 7, 8, 9)
  static ZB_HD int il_col(int chain_k) { return chain_k == 0 ? 4 : chain_k == 1 ? 2 : chain_k == 2 ? 0 : chain_k == 3 ? 1 : chain_k == 4 ? 3 : 5; }
};
static_assert(model_m::FOOT1_MASS == model_m::FOOT0_MASS && model_m::MIDB_MASS == model_m::MIDA_MASS &&
              model_m::BASE_MASS == model_m::MIDA_MASS && model_m::FOOT1_IXX == model_m::FOOT0_IXX &&
              model_m::FOOT1_IXY == model_m::FOOT0_IXY && model_m::FOOT1_IXZ == model_m::FOOT0_IXZ &&
              model_m::FOOT1_IYY == model_m::FOOT0_IYY && model_m::FOOT1_IYZ == model_m::FOOT0_IYZ &&
              model_m::FOOT1_IZZ == model_m::FOOT0_IZZ && model_m::MIDB_IXX == model_m::MIDA_IXX &&
              model_m::MIDB_IXY == model_m::MIDA_IXY && model_m::MIDB_IYY == model_m::MIDA_IYY &&
              model_m::MIDB_IZZ == model_m::MIDA_IZZ && model_m::MIDB_CX == model_m::MIDA_CX &&
              model_m::MIDB_CY == model_m::MIDA_CY, "ModelWalkM::body_full shares table entries between bodies");

template <typename Model, typename T>
ZB_HD void sim_state_default(SimState<T>& s) {
  Model::default_root(s.p, s.Q);
  ZB_UNROLL for (int i = 0; i < 3; ++i) { s.v[i] = T(0); s.w[i] = T(0); }
  ZB_UNROLL for (int i = 0; i < 6; ++i) { s.q[i] = Model::template default_q<T>(i); s.qd[i] = T(0); }
}
template <typename T>
ZB_HD void sim_state_default(SimState<T>& s) { sim_state_default<ModelWalk>(s); }
template <typename T>
ZB_HD T default_joint_pos(int k) { return ModelWalk::default_q<T>(k); }

// ------------------------------------------------------------------------------------
// per-thread scratch for the joint-indexed quantities of one substep.  The sweeps over the
// chain are REAL loops (not unrolled: the fully unrolled step was ~160 KB of SASS and stalled
// on instruction fetch), so everything indexed by the joint lives in indexable storage:
// shared memory on the GPU ([slot][thread], conflict-free), a plain array on the CPU.
// ------------------------------------------------------------------------------------
// ZB_STORE_Q: the kinematics sweep parks each body's quaternion (4 words per joint) and the backward sweep reloads it
// instead of unwinding the joint rotation from (sin, cos) (2 words): 16 instructions and one link of the dependent chain less
// per joint, 12 words more scratch per thread.
#ifndef ZB_STORE_Q
#define ZB_STORE_Q 0
#endif
constexpr int SCR_PER_JOINT = ZB_STORE_Q ? 19 : 17;
constexpr int SCR_WORDS = 6 * SCR_PER_JOINT;   // 102 (114) words per environment
enum ScrSlot : int { SC_SA = 0, SC_SM = 3, SC_UT = 6, SC_UB = 9, SC_DINV = 12, SC_U = 13 /* tau, then u */,
                     SC_SN = 14, SC_CS = 15, SC_QJ = 14 /* ZB_STORE_Q: body quaternion, 4 words */, SC_QD = ZB_STORE_Q ? 18 : 16 };

template <typename T>
struct ArrayScratch {
  T a[SCR_WORDS];
  ZB_HD T& operator()(int j, int slot) { return a[j * SCR_PER_JOINT + slot]; }
};
#if defined(__CUDACC__)
// Each thread owns SCR_STRIDE consecutive words of shared memory.  SCR_STRIDE is odd, so the 32 lanes of a
// warp hit 32 different banks for any (j, slot); `slot` is a compile-time constant at every use, so it folds
// into the LDS/STS immediate and only `j * SCR_PER_JOINT` costs an instruction per loop iteration.
constexpr int SCR_RAW_ACT = SCR_WORDS;       // 6 words: this step's raw actions, parked across the physics phase
constexpr int SCR_STRIDE = SCR_WORDS + 7;   // 109 (odd)
struct SmemScratch {
  float* base;   // this thread's row
  __device__ __forceinline__ float& operator()(int j, int slot) { return base[j * SCR_PER_JOINT + slot]; }
};
#endif

// ------------------------------------------------------------------------------------
// the ground under a contact candidate: height of the terrain at env-LOCAL (x, y), relative to the env origin's z.
// FlatGround = the plane z = 0 of every flat task (folds away).  TerrainGround = the rough manager task's height field
// (zbot_lab_b200/terrain.py: one float32 grid over all tiles, bilinear sample; normals stay vertical).
// ------------------------------------------------------------------------------------
struct FlatGround {
  template <typename T>
  ZB_HD T operator()(T, T) const { return T(0); }
};
template <typename T>
struct TerrainGround {
  const float* H;      // [nx][ny] heights, world frame
  int nx, ny;
  T x0, y0, inv_cell;  // world coordinate of H[0][0], 1 / cell size
  T ox, oy, oz;        // this env's origin (world): local + origin = world
  ZB_HD T operator()(T xl, T yl) const {
    T fx = (xl + ox - x0) * inv_cell, fy = (yl + oy - y0) * inv_cell;
    fx = zb_clamp(fx, T(0), T(nx) - T(1.001));
    fy = zb_clamp(fy, T(0), T(ny) - T(1.001));
    const int ix = (int)fx, iy = (int)fy;
    const T tx = fx - T(ix), ty = fy - T(iy);
    const float* p = H + (size_t)ix * ny + iy;
    const T h00 = T(p[0]), h01 = T(p[1]), h10 = T(p[ny]), h11 = T(p[ny + 1]);
    const T a = h00 + tx * (h10 - h00), b = h01 + tx * (h11 - h01);
    return (a + ty * (b - a)) - oz;
  }
};

// ------------------------------------------------------------------------------------
// one physics substep (dt = P.dt): implicit PD + contact + ABA + semi-implicit Euler
// ------------------------------------------------------------------------------------
template <typename T>
struct SubstepOut {
  T foot_force[2][3];  // net contact force on foot_0 / foot_1 (applied, world frame)
  T mid_force2_max;    // max over bodies 1..5 of |predictor contact force|^2
  T applied_torque[6]; // ImplicitActuator bookkeeping evaluated BEFORE this substep
  T mid_force2[5];     // Model::kPerBodyForces only: |predictor contact force|^2 of each merged body 1..5
};

// mid_force_out: optional [5][3] predictor forces of bodies 1..5 (export / debug only)
// kUnroll: unroll factor of the three sweeps over the chain (FK, backward, forward).  1 = smallest code (best for a lone
// warp per SM: instruction fetch is 11 % of its time); 2 halves the loop-carried register shuffles (MOV was 8.5 % of the
// executed instructions) and wins ~4 % once two or more warps share a sub-partition (profiles/r1_notes.md).
// per-sweep unroll factors (tuning builds override them with -D; profiles/r1_notes.md "third session"): the backward and
// forward sweeps follow kUnroll, the short kinematics sweep is always fully unrolled
#define ZB_DO_PRAGMA(x) _Pragma(#x)
#define ZB_PRAGMA_UNROLL(n) ZB_DO_PRAGMA(unroll n)
#ifndef ZB_UNROLL_FK
#define ZB_UNROLL_FK 6   // kinematics sweep fully unrolled in every instantiation: 35.8 -> 34.9 us at 4096 envs, 84.0 -> 82.0 at 65536
#endif
#ifndef ZB_UNROLL_BWD
#define ZB_UNROLL_BWD kUnroll
#endif
#ifndef ZB_UNROLL_PTS
// ground-contact candidate loop of a body (4 rim points on a foot, 1 sphere otherwise): the four rim points are independent
// until they are accumulated, so the throughput instantiation unrolls them (82.0 -> 80.0 us at 65536 envs); the rolled
// kernel keeps the loop (a lone warp per SM is bound by instruction fetch: 34.9 us rolled vs 39.0 us unrolled at 4096 envs)
#define ZB_UNROLL_PTS (kUnroll == 1 ? 1 : 4)
#endif
#ifndef ZB_UNROLL_SUB
#define ZB_UNROLL_SUB 1   // the decimation loop of env_step_physics
#endif
// Tuning experiment (profiles/r2_notes.md "instruction fetch"): CTA barriers at phase boundaries keep the warps of a CTA in the
// same stretch of the ~70 KB substep loop, so they share instruction-cache lines.  Level 1: once per substep; 2: also before the
// backward and forward sweeps; 3: every backward-sweep iteration.  Only legal when every thread of the CTA is live (the
// experiment's launches are); off (0) in the product build.
#ifndef ZB_PHASE_BAR_LEVEL
#define ZB_PHASE_BAR_LEVEL 0
#endif
#if defined(__CUDA_ARCH__) && ZB_PHASE_BAR_LEVEL > 0
#define ZB_PHASE_BAR(lvl) do { if (ZB_PHASE_BAR_LEVEL >= (lvl)) __syncthreads(); } while (0)
#else
#define ZB_PHASE_BAR(lvl) do { } while (0)
#endif
#ifndef ZB_UNROLL_FWD
#define ZB_UNROLL_FWD kUnroll
#endif
template <typename Model, bool kPipe = (ZB_PIPELINED_SWEEP != 0), int kUnroll = 1, typename PS, typename T, typename Scr,
          typename Ground = FlatGround>
ZB_HD void physics_substep(const Params<PS>& P, SimState<T>& s, const T* target, SubstepOut<T>& out, Scr& scr,
                           T* mid_force_out, const Ground& ground = Ground()) {
  using namespace model;
  const T dt = T(P.dt);
  // friction coefficient: the uniform cfg value, or per env (`target[6]`) for models with randomised materials
  const T mu = Model::kPerEnvFriction ? target[6] : T(P.c_mu);
  // ---- PD (implicit part lives in P.arm) + park q, qd (static indices: registers -> scratch) ----
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    const T e = target[k] - s.q[k];
    out.applied_torque[k] = zb_clamp(T(P.kp) * e - T(P.kd) * s.qd[k], T(-P.effort), T(P.effort));
    scr(k, SC_U) = zb_clamp(T(P.kp) * (e - dt * s.qd[k]) - T(P.kd) * s.qd[k], T(-P.effort), T(P.effort));
    scr(k, SC_QD) = s.qd[k];
  }
  // ---- forward kinematics sweep: motion subspaces S_k = (a_k ; r_k x a_k), arrive at body 6 ----
  T Q[4] = {s.Q[0], s.Q[1], s.Q[2], s.Q[3]};
  T r[3] = {T(0), T(0), T(0)};               // body origin relative to O (= root origin)
  T w[3] = {s.w[0], s.w[1], s.w[2]};         // spatial velocity of the current body about O
  T vO[3] = {s.v[0], s.v[1], s.v[2]};
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_FK)
#endif
  for (int k = 0; k < 6; ++k) {
    T R[9];
    quat_to_mat(Q, R);
    const T jz = (k == 0) ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
    const T sg = (k & 1) ? T(-AXIS_S) : T(AXIS_S);
    r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
    T a[3] = {sg * R[0] + T(AXIS_S) * R[2], sg * R[3] + T(AXIS_S) * R[5], sg * R[6] + T(AXIS_S) * R[8]};
    T m[3];
    cross3(r, a, m);
    const T qd = scr(k, SC_QD);
    T sn, cs;
    // select (not index) the joint angle: q stays in registers
    const T qk = (k == 0) ? s.q[0] : (k == 1) ? s.q[1] : (k == 2) ? s.q[2] : (k == 3) ? s.q[3] : (k == 4) ? s.q[4] : s.q[5];
    zb_sincos(T(0.5) * qk, &sn, &cs);
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      scr(k, SC_SA + i) = a[i];
      scr(k, SC_SM + i) = m[i];
      w[i] += a[i] * qd;
      vO[i] += m[i] * qd;
    }
    if (ZB_STORE_Q) { ZB_UNROLL for (int i = 0; i < 4; ++i) scr(k, SC_QJ + i) = Q[i]; }      // body k, before joint k turns it
    else { scr(k, SC_SN) = sn; scr(k, SC_CS) = cs; }
    quat_mul_joint(Q, cs, sg * sn, T(AXIS_S) * sn);
  }
  ZB_PHASE_BAR(2);
  // ---- backward sweep over bodies 6..0: rigid + contact terms, then eliminate the joint above ----
  SpInertia<T> IA;
  ZB_UNROLL for (int i = 0; i < 6; ++i) { IA.I[i] = T(0); IA.M[i] = T(0); }
  ZB_UNROLL for (int i = 0; i < 9; ++i) IA.H[i] = T(0);
  T pAt[3] = {T(0), T(0), T(0)}, pAb[3] = {T(0), T(0), T(0)};
  ContactAgg<T> agg, agg1;   // running / foot_1 (body 6); after the sweep `agg` holds foot_0 (body 0)
  contact_agg_zero(agg1);
  T mid2 = T(0);
  if constexpr (Model::kPerBodyForces) { ZB_UNROLL for (int b = 0; b < 5; ++b) out.mid_force2[b] = T(0); }
  if constexpr (kPipe) {
  // Software-pipelined form: the articulated-body elimination of joint k-1 is one long dependent chain
  // (IA S -> D -> 1/D -> rank-1 update); the kinematics, world inertia and bias force of the NEXT body (k-1) do not
  // depend on it.  Both are issued in the same basic block, into separate accumulators (IB, pB), so the scheduler
  // interleaves two instruction streams per thread (DESIGN.md §4 "ILP"); IA += IB afterwards.
  {
    T R[9];
    quat_to_mat(Q, R);
    model_body_terms<Model>(P, 6, R, r, w, vO, IA, pAt, pAb);
    contact_agg_zero(agg);
    const int npts = Model::npts(6);
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_PTS)
#endif
    for (int j = 0; j < npts; ++j) {
      T lx, ly, lz, drop;
      Model::point(6, j, lx, ly, lz, drop);
      T rho[3] = {r[0] + R[0] * lx + R[1] * ly + R[2] * lz, r[1] + R[3] * lx + R[4] * ly + R[5] * lz,
                  r[2] + R[6] * lx + R[7] * ly + R[8] * lz - drop};
      contact_point(P, mu, rho, s.p[2] + rho[2] - ground(s.p[0] + rho[0], s.p[1] + rho[1]), w, vO, IA, pAt, pAb,
                    Model::kGroundForceSensor ? &agg : (ContactAgg<T>*)nullptr, (T*)nullptr);
    }
    if (Model::kGroundForceSensor) agg1 = agg;
  }
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_BWD)
#endif
  for (int k = 6; k >= 1; --k) {
    const int j = k - 1;          // joint between body k and body k-1
    ZB_PHASE_BAR(3);
    const T Sa[3] = {scr(j, SC_SA), scr(j, SC_SA + 1), scr(j, SC_SA + 2)};
    const T Sm[3] = {scr(j, SC_SM), scr(j, SC_SM + 1), scr(j, SC_SM + 2)};
    const T qd = scr(j, SC_QD);
    const T sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
    const T sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
    // ---- stream B: kinematics + rigid terms of body k-1 (independent of IA) ----
    T wn[3], vn[3], Qn[4], rn[3], Rn[9];
    {
      const bool root = (j == 0);                          // root: exact kinematics from the state
      if (ZB_STORE_Q) {
        ZB_UNROLL for (int i = 0; i < 4; ++i) Qn[i] = scr(j, SC_QJ + i);   // parked by the kinematics sweep (j == 0: s.Q itself)
      } else {
        const T sg = (j & 1) ? T(-AXIS_S) : T(AXIS_S);
        const T sn = scr(j, SC_SN), cs = scr(j, SC_CS);
        ZB_UNROLL for (int i = 0; i < 4; ++i) Qn[i] = Q[i];
        quat_mul_joint(Qn, cs, -sg * sn, -T(AXIS_S) * sn);   // Q_{k-1} = Q_k (x) conj(qj)
        ZB_UNROLL for (int i = 0; i < 4; ++i) Qn[i] = root ? s.Q[i] : Qn[i];
      }
      quat_to_mat(Qn, Rn);
      const T jz = root ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
      ZB_UNROLL for (int i = 0; i < 3; ++i) {
        rn[i] = root ? T(0) : (r[i] - jz * Rn[3 * i + 2]);
        wn[i] = root ? s.w[i] : (w[i] - sa[i]);
        vn[i] = root ? s.v[i] : (vO[i] - sm[i]);
      }
    }
    SpInertia<T> IB;
    ZB_UNROLL for (int i = 0; i < 6; ++i) { IB.I[i] = T(0); IB.M[i] = T(0); }
    ZB_UNROLL for (int i = 0; i < 9; ++i) IB.H[i] = T(0);
    T pBt[3] = {T(0), T(0), T(0)}, pBb[3] = {T(0), T(0), T(0)};
    model_body_terms<Model>(P, k - 1, Rn, rn, wn, vn, IB, pBt, pBb);
    // ---- stream A: velocity-product term c = V x (S qd), articulated-body elimination of joint j ----
    T ct[3], cb[3], tmp[3];
    cross3(w, sa, ct);
    cross3(w, sm, cb);
    cross3(vO, sa, tmp);
    cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
    T Ut[3], Ub[3];
    spi_mul(IA, Sa, Sm, Ut, Ub);
    const T D = dot3(Sa, Ut) + dot3(Sm, Ub) + T(P.arm);
    const T Dinv = zb_rcp(D);
    const T u = scr(j, SC_U) - (dot3(Sa, pAt) + dot3(Sm, pAb));
    T Ict[3], Icb[3];
    spi_mul(IA, ct, cb, Ict, Icb);
    const T g = (u - (dot3(Ut, ct) + dot3(Ub, cb))) * Dinv;
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      pAt[i] += Ict[i] + Ut[i] * g;
      pAb[i] += Icb[i] + Ub[i] * g;
      scr(j, SC_UT + i) = Ut[i];
      scr(j, SC_UB + i) = Ub[i];
    }
    scr(j, SC_DINV) = Dinv;
    scr(j, SC_U) = u;
    spi_rank1_sub(IA, Ut, Ub, Dinv);
    // ---- merge, then the ground contacts of body k-1 ----
    ZB_UNROLL for (int i = 0; i < 6; ++i) { IA.I[i] += IB.I[i]; IA.M[i] += IB.M[i]; }
    ZB_UNROLL for (int i = 0; i < 9; ++i) IA.H[i] += IB.H[i];
    ZB_UNROLL for (int i = 0; i < 3; ++i) { pAt[i] += pBt[i]; pAb[i] += pBb[i]; }
    ZB_UNROLL for (int i = 0; i < 3; ++i) { r[i] = rn[i]; w[i] = wn[i]; vO[i] = vn[i]; }
    ZB_UNROLL for (int i = 0; i < 4; ++i) Q[i] = Qn[i];
    contact_agg_zero(agg);
    const int npts = Model::npts(k - 1);
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_PTS)
#endif
    for (int c = 0; c < npts; ++c) {
      T lx, ly, lz, drop;
      Model::point(k - 1, c, lx, ly, lz, drop);
      T rho[3] = {r[0] + Rn[0] * lx + Rn[1] * ly + Rn[2] * lz, r[1] + Rn[3] * lx + Rn[4] * ly + Rn[5] * lz,
                  r[2] + Rn[6] * lx + Rn[7] * ly + Rn[8] * lz - drop};
      contact_point(P, mu, rho, s.p[2] + rho[2] - ground(s.p[0] + rho[0], s.p[1] + rho[1]), w, vO, IA, pAt, pAb,
                    Model::kGroundForceSensor ? &agg : (ContactAgg<T>*)nullptr, (T*)nullptr);
    }
    if (Model::kGroundForceSensor) {
      if (k - 1 != 0) {
        const T f2 = agg.F0[0] * agg.F0[0] + agg.F0[1] * agg.F0[1] + agg.F0[2] * agg.F0[2];
        mid2 = zb_max(mid2, f2);
        if constexpr (Model::kPerBodyForces) { ZB_UNROLL for (int b = 0; b < 5; ++b) out.mid_force2[b] = (b == k - 2) ? f2 : out.mid_force2[b]; }
        if (mid_force_out) { mid_force_out[3 * (k - 2)] = agg.F0[0]; mid_force_out[3 * (k - 2) + 1] = agg.F0[1];
                             mid_force_out[3 * (k - 2) + 2] = agg.F0[2]; }
      }
    }
  }
  } else {
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_BWD)
#endif
  for (int k = 6; k >= 0; --k) {
    if (k == 0) {   // root: exact kinematics from the state (no unwinding round-off)
      ZB_UNROLL for (int i = 0; i < 3; ++i) { r[i] = T(0); w[i] = s.w[i]; vO[i] = s.v[i]; }
      ZB_UNROLL for (int i = 0; i < 4; ++i) Q[i] = s.Q[i];
    }
    T R[9];
    quat_to_mat(Q, R);
    const bool foot1 = (k == 6), foot0 = (k == 0);
    model_body_terms<Model>(P, k, R, r, w, vO, IA, pAt, pAb);
    // ground-contact candidates of this body (Model::point)
    contact_agg_zero(agg);
    const int npts = Model::npts(k);
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_PTS)
#endif
    for (int j = 0; j < npts; ++j) {
      T lx, ly, lz, drop;
      Model::point(k, j, lx, ly, lz, drop);
      T rho[3] = {r[0] + R[0] * lx + R[1] * ly + R[2] * lz, r[1] + R[3] * lx + R[4] * ly + R[5] * lz,
                  r[2] + R[6] * lx + R[7] * ly + R[8] * lz - drop};
      contact_point(P, mu, rho, s.p[2] + rho[2] - ground(s.p[0] + rho[0], s.p[1] + rho[1]), w, vO, IA, pAt, pAb,
                    Model::kGroundForceSensor ? &agg : (ContactAgg<T>*)nullptr, (T*)nullptr);
    }
    if (Model::kGroundForceSensor) {
      if (foot1) {
        agg1 = agg;
      } else if (!foot0) {
        const T f2 = agg.F0[0] * agg.F0[0] + agg.F0[1] * agg.F0[1] + agg.F0[2] * agg.F0[2];
        mid2 = zb_max(mid2, f2);
        if constexpr (Model::kPerBodyForces) { ZB_UNROLL for (int b = 0; b < 5; ++b) out.mid_force2[b] = (b == k - 1) ? f2 : out.mid_force2[b]; }
        if (mid_force_out) { mid_force_out[3 * (k - 1)] = agg.F0[0]; mid_force_out[3 * (k - 1) + 1] = agg.F0[1];
                             mid_force_out[3 * (k - 1) + 2] = agg.F0[2]; }
      }
    }
    if (k == 0) break;
    // joint k (index j): velocity-product term c = V x (S qd), articulated-body elimination
    const int j = k - 1;
    const T Sa[3] = {scr(j, SC_SA), scr(j, SC_SA + 1), scr(j, SC_SA + 2)};
    const T Sm[3] = {scr(j, SC_SM), scr(j, SC_SM + 1), scr(j, SC_SM + 2)};
    const T qd = scr(j, SC_QD);
    T sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
    T sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
    T ct[3], cb[3], tmp[3];
    cross3(w, sa, ct);
    cross3(w, sm, cb);
    cross3(vO, sa, tmp);
    cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
    T Ut[3], Ub[3];
    spi_mul(IA, Sa, Sm, Ut, Ub);
    const T D = dot3(Sa, Ut) + dot3(Sm, Ub) + T(P.arm);
    const T Dinv = zb_rcp(D);
    const T u = scr(j, SC_U) - (dot3(Sa, pAt) + dot3(Sm, pAb));
    // pa = pA + IA c + U (u - U.c)/D   (== pA + Ia c + U u/D)
    T Ict[3], Icb[3];
    spi_mul(IA, ct, cb, Ict, Icb);
    const T g = (u - (dot3(Ut, ct) + dot3(Ub, cb))) * Dinv;
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      pAt[i] += Ict[i] + Ut[i] * g;
      pAb[i] += Icb[i] + Ub[i] * g;
      scr(j, SC_UT + i) = Ut[i];
      scr(j, SC_UB + i) = Ub[i];
    }
    scr(j, SC_DINV) = Dinv;
    scr(j, SC_U) = u;
    spi_rank1_sub(IA, Ut, Ub, Dinv);
    // unwind kinematics to body k-1
    ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] -= sa[i]; vO[i] -= sm[i]; }
    if (ZB_STORE_Q) {
      ZB_UNROLL for (int i = 0; i < 4; ++i) Q[i] = scr(j, SC_QJ + i);
    } else {
      const T sg = (j & 1) ? T(-AXIS_S) : T(AXIS_S);
      const T sn = scr(j, SC_SN), cs = scr(j, SC_CS);
      quat_mul_joint(Q, cs, -sg * sn, -T(AXIS_S) * sn);   // Q_{k-1} = Q_k (x) conj(qj)
    }
    const T jz = (j == 0) ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
    T Rp2[3] = {T(2) * (Q[1] * Q[3] + Q[0] * Q[2]), T(2) * (Q[2] * Q[3] - Q[0] * Q[1]),
                T(1) - T(2) * (Q[1] * Q[1] + Q[2] * Q[2])};     // third column of R_{k-1}
    r[0] -= jz * Rp2[0]; r[1] -= jz * Rp2[1]; r[2] -= jz * Rp2[2];
  }
  }   // kPipe
  // ---- floating base:  IA a0 = -pA ----
  T At[3], Ab[3];
  {
    T nt[3] = {-pAt[0], -pAt[1], -pAt[2]}, nb[3] = {-pAb[0], -pAb[1], -pAb[2]};
    spi_solve(IA, nt, nb, At, Ab);
  }
  if (Model::kGroundForceSensor) contact_agg_force(agg, dt, At, Ab, out.foot_force[0]);
  // classical acceleration of the root origin = spatial + w x v
  T wxv[3];
  cross3(s.w, s.v, wxv);
  T wk[3] = {s.w[0], s.w[1], s.w[2]}, vk[3] = {s.v[0], s.v[1], s.v[2]};  // V_{k-1} (pre-update velocities)
  ZB_UNROLL for (int i = 0; i < 3; ++i) {
    s.w[i] += dt * At[i];
    s.v[i] += dt * (Ab[i] + wxv[i]);
  }
  ZB_PHASE_BAR(2);
  // ---- forward sweep: joint accelerations ----
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_FWD)
#endif
  for (int j = 0; j < 6; ++j) {
    const T Sa[3] = {scr(j, SC_SA), scr(j, SC_SA + 1), scr(j, SC_SA + 2)};
    const T Sm[3] = {scr(j, SC_SM), scr(j, SC_SM + 1), scr(j, SC_SM + 2)};
    const T Ut[3] = {scr(j, SC_UT), scr(j, SC_UT + 1), scr(j, SC_UT + 2)};
    const T Ub[3] = {scr(j, SC_UB), scr(j, SC_UB + 1), scr(j, SC_UB + 2)};
    T qd = scr(j, SC_QD);
    T sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
    T sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
    T ct[3], cb[3], tmp[3];
    cross3(wk, sa, ct);
    cross3(wk, sm, cb);
    cross3(vk, sa, tmp);
    cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
    ZB_UNROLL for (int i = 0; i < 3; ++i) { At[i] += ct[i]; Ab[i] += cb[i]; wk[i] += sa[i]; vk[i] += sm[i]; }
    const T qdd = (scr(j, SC_U) - (dot3(Ut, At) + dot3(Ub, Ab))) * scr(j, SC_DINV);
    ZB_UNROLL for (int i = 0; i < 3; ++i) { At[i] += Sa[i] * qdd; Ab[i] += Sm[i] * qdd; }
    qd += dt * qdd;
    scr(j, SC_QD) = qd;
  }
  if (Model::kGroundForceSensor) contact_agg_force(agg1, dt, At, Ab, out.foot_force[1]);
  out.mid_force2_max = mid2;
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    s.qd[k] = scr(k, SC_QD);
    s.q[k] += dt * s.qd[k];
    if (Model::wraps(k)) s.q[k] = zb_wrap_joint(s.q[k]);
  }
  // ---- root pose ----
  ZB_UNROLL for (int i = 0; i < 3; ++i) s.p[i] += dt * s.v[i];
  {
    T h = T(0.5) * dt;
    T qw = s.Q[0], qx = s.Q[1], qy = s.Q[2], qz = s.Q[3];
    T nw = qw + h * (-s.w[0] * qx - s.w[1] * qy - s.w[2] * qz);
    T nx = qx + h * (s.w[0] * qw + s.w[1] * qz - s.w[2] * qy);
    T ny = qy + h * (-s.w[0] * qz + s.w[1] * qw + s.w[2] * qx);
    T nz = qz + h * (s.w[0] * qy - s.w[1] * qx + s.w[2] * qw);
    T inv = zb_rsqrt(nw * nw + nx * nx + ny * ny + nz * nz);
    s.Q[0] = nw * inv; s.Q[1] = nx * inv; s.Q[2] = ny * inv; s.Q[3] = nz * inv;
  }
}

// ------------------------------------------------------------------------------------
// kinematic quantities the MDP reads from robot.data (…env_v2.py:315-326, 554)
// ------------------------------------------------------------------------------------
template <typename T>
struct LinkKin {
  T base_pos[3], base_quat[4], base_com_vel[3];
  T feet_pos[2][3], feet_quat[2][4], feet_com_vel[2][3];
  T base_link_vel[3];   // body_link_lin_vel_w[:, base] (velocity of the LINK origin; v4 reads this, …env_v4.py:812)
};

template <typename T>
ZB_LINKKIN_ATTR void link_kinematics(const SimState<T>& s, LinkKin<T>& o) {
  using namespace model;
  T Q[4] = {s.Q[0], s.Q[1], s.Q[2], s.Q[3]};
  T r[3] = {T(0), T(0), T(0)};
  T w[3] = {s.w[0], s.w[1], s.w[2]};
  T vO[3] = {s.v[0], s.v[1], s.v[2]};
  // park the joint state where a real (not unrolled) loop can index it
  T qk[6], qdk[6];
  ZB_UNROLL for (int k = 0; k < 6; ++k) { qk[k] = s.q[k]; qdk[k] = s.qd[k]; }
  T R[9];
  quat_to_mat(Q, R);
  {  // body 0 = foot_0 (a-type link)
    T c[3] = {R[0] * T(A_COM_X) + R[2] * T(A_COM_Z), R[3] * T(A_COM_X) + R[5] * T(A_COM_Z),
              R[6] * T(A_COM_X) + R[8] * T(A_COM_Z)};
    T wxc[3];
    cross3(w, c, wxc);
    ZB_UNROLL for (int i = 0; i < 3; ++i) { o.feet_pos[0][i] = s.p[i]; o.feet_com_vel[0][i] = vO[i] + wxc[i]; }
    ZB_UNROLL for (int i = 0; i < 4; ++i) o.feet_quat[0][i] = Q[i];
  }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = 0; k < 6; ++k) {
    const T jz = (k == 0) ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
    const T sg = (k & 1) ? T(-AXIS_S) : T(AXIS_S);
    r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
    T a[3] = {sg * R[0] + T(AXIS_S) * R[2], sg * R[3] + T(AXIS_S) * R[5], sg * R[6] + T(AXIS_S) * R[8]};
    T m[3];
    cross3(r, a, m);
    // select (not index) the joint state: keeps q/qd in registers
    const T qv = (k == 0) ? qk[0] : (k == 1) ? qk[1] : (k == 2) ? qk[2] : (k == 3) ? qk[3] : (k == 4) ? qk[4] : qk[5];
    const T qdv = (k == 0) ? qdk[0] : (k == 1) ? qdk[1] : (k == 2) ? qdk[2] : (k == 3) ? qdk[3] : (k == 4) ? qdk[4] : qdk[5];
    ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] += a[i] * qdv; vO[i] += m[i] * qdv; }
    T sn, cs;
    zb_sincos(T(0.5) * qv, &sn, &cs);
    quat_mul_joint(Q, cs, sg * sn, T(AXIS_S) * sn);
    quat_to_mat(Q, R);
    if (k == 2 || k == 5) {
      // k == 2: body 3 = b3 + base; base LINK origin = body origin + R (0,0,LINK_OFFSET_Z), a-type CoM
      // k == 5: body 6 = foot_1 (b-type link), link origin = body origin
      const T oz = (k == 2) ? T(LINK_OFFSET_Z) : T(0);
      const T cx = (k == 2) ? T(A_COM_X) : T(B_COM_X), cz = (k == 2) ? T(A_COM_Z) : T(B_COM_Z);
      T lo[3] = {r[0] + R[2] * oz, r[1] + R[5] * oz, r[2] + R[8] * oz};
      T c[3] = {lo[0] + R[0] * cx + R[2] * cz, lo[1] + R[3] * cx + R[5] * cz, lo[2] + R[6] * cx + R[8] * cz};
      T wxc[3];
      cross3(w, c, wxc);
      if (k == 2) {
        T wxl[3];
        cross3(w, lo, wxl);
        ZB_UNROLL for (int i = 0; i < 3; ++i) o.base_link_vel[i] = vO[i] + wxl[i];
        ZB_UNROLL for (int i = 0; i < 3; ++i) { o.base_pos[i] = s.p[i] + lo[i]; o.base_com_vel[i] = vO[i] + wxc[i]; }
        ZB_UNROLL for (int i = 0; i < 4; ++i) o.base_quat[i] = Q[i];
      } else {
        ZB_UNROLL for (int i = 0; i < 3; ++i) { o.feet_pos[1][i] = s.p[i] + lo[i]; o.feet_com_vel[1][i] = vO[i] + wxc[i]; }
        ZB_UNROLL for (int i = 0; i < 4; ++i) o.feet_quat[1][i] = Q[i];
      }
    }
  }
}

// all 12 link poses + CoM velocities (articulation order) -- export / debug only
template <typename T>
ZB_HD void all_link_kinematics(const SimState<T>& s, T* pos /*12x3*/, T* quat /*12x4*/, T* comvel /*12x3*/) {
  using namespace model;
  T Q[4] = {s.Q[0], s.Q[1], s.Q[2], s.Q[3]};
  T r[3] = {T(0), T(0), T(0)};
  T w[3] = {s.w[0], s.w[1], s.w[2]};
  T vO[3] = {s.v[0], s.v[1], s.v[2]};
  for (int b = 0; b < 7; ++b) {
    T R[9];
    quat_to_mat(Q, R);
    const int nl = (b == 0 || b == 6) ? 1 : 2;
    for (int h = 0; h < nl; ++h) {
      const int link = (b == 0) ? 0 : (b == 6) ? 11 : (2 * b - 1 + h);
      const bool a_type = (b == 0) || (h == 1);
      const T oz = (h == 1) ? T(LINK_OFFSET_Z) : T(0);
      const T cx = a_type ? T(A_COM_X) : T(B_COM_X), cz = a_type ? T(A_COM_Z) : T(B_COM_Z);
      T lo[3] = {r[0] + R[2] * oz, r[1] + R[5] * oz, r[2] + R[8] * oz};
      T c[3] = {lo[0] + R[0] * cx + R[2] * cz, lo[1] + R[3] * cx + R[5] * cz, lo[2] + R[6] * cx + R[8] * cz};
      T wxc[3];
      cross3(w, c, wxc);
      for (int i = 0; i < 3; ++i) { pos[3 * link + i] = s.p[i] + lo[i]; comvel[3 * link + i] = vO[i] + wxc[i]; }
      for (int i = 0; i < 4; ++i) quat[4 * link + i] = Q[i];
    }
    if (b < 6) {
      const int k = b;
      const T jz = (k == 0) ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
      const T sg = (k & 1) ? T(-AXIS_S) : T(AXIS_S);
      r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
      T a[3] = {sg * R[0] + T(AXIS_S) * R[2], sg * R[3] + T(AXIS_S) * R[5], sg * R[6] + T(AXIS_S) * R[8]};
      T m[3];
      cross3(r, a, m);
      for (int i = 0; i < 3; ++i) { w[i] += a[i] * s.qd[k]; vO[i] += m[i] * s.qd[k]; }
      T sn, cs;
      zb_sincos(T(0.5) * s.q[k], &sn, &cs);
      quat_mul_joint(Q, cs, sg * sn, T(AXIS_S) * sn);
    }
  }
}

// ------------------------------------------------------------------------------------
// MDP state carried between control steps (reference attributes, ...env_v2.py:215-245)
// ------------------------------------------------------------------------------------
template <typename T>
struct MdpState {
  T p_delta[6];
  T actions[6];              // _actions == _previous_actions between steps (:313)
  T feet_force_last[2];      // feet_contact_forces_last
  T feet_down_pos_last[2][3];
  T feet_step_length[2];
  T heading_sum, y_err_sum, feet_force_sum;
  T speed_limit;             // joint_speed_limit
  T ep_sums[MAX_TERMS];      // _episode_sums in cfg dict order
};

// quantities cached by the PREVIOUS _get_observations (:315-345) -- "stale" (SURVEY C-1)
template <typename T>
struct StaleCache {
  T base_pos[3];      // base_pos_w (world; fused path: env-local with origin 0)
  T forward[3];       // base_dir_forward_w (NOT normalised, SURVEY C-2)
  T feet_x[2][3], feet_z[2][3], feet_pos[2][3];
  T v_fwd;            // base_lin_vel_forward_w
  T aux;              // snake task: base_up_w.y (snake_v0.py:181); `forward` then holds base_heading_w
};

template <typename T>
ZB_HD void stale_from_links(const T* base_pos, const T* base_quat, const T* base_com_vel,
                            const T feet_pos[2][3], const T feet_quat[2][4], StaleCache<T>& c) {
  const T ez[3] = {T(0), T(0), T(1)};
  const T enz[3] = {T(0), T(0), T(-1)};
  const T ex[3] = {T(1), T(0), T(0)};
  T shoulder[3];
  quat_apply(base_quat, ez, shoulder);                 // :322
  const T grav[3] = {T(0), T(0), T(-1)};
  cross3(grav, shoulder, c.forward);                   // :323
  c.v_fwd = base_com_vel[0] * c.forward[0] + base_com_vel[1] * c.forward[1] + base_com_vel[2] * c.forward[2];  // :327
  quat_apply(feet_quat[0], ez, c.feet_z[0]);           // :344  axis_z_feet = [[0,0,1],[0,0,-1]]
  quat_apply(feet_quat[1], enz, c.feet_z[1]);
  quat_apply(feet_quat[0], ex, c.feet_x[0]);           // :345
  quat_apply(feet_quat[1], ex, c.feet_x[1]);
  ZB_UNROLL for (int i = 0; i < 3; ++i) {
    c.base_pos[i] = base_pos[i];
    c.feet_pos[0][i] = feet_pos[0][i];
    c.feet_pos[1][i] = feet_pos[1][i];
  }
}

// fresh (end-of-physics) inputs of _get_dones/_get_rewards
template <typename T>
struct FreshInputs {
  T feet_force[2];        // mean_t history[:, t, foot, z]  (:387-390)
  T last_air_time[2];     // (:391)
  T undesired_force_max;  // max_{t,b} |history[:, t, b, :]|  over the undesired bodies (:396-402)
  T feet_vel_xy[2][2];    // body_com_lin_vel_w[:, feet, :2]  (:554)
  T applied_torque[6];    // (:560)
  T origin_y;             // env_origins[:, 1]  (0 in the fused path: env-local coordinates)
  T com_x_sum;            // snake task: body_com_pos_w[:, 0, 0] + body_com_pos_w[:, 11, 0] - 2 origin_x (snake_v0.py:330-333)
};

// …env_v2.py:276-287.  raw -> post-tanh actions, p_delta integration/clip, joint targets
template <typename Model, typename T>
ZB_HD void mdp_pre_physics(const Params<T>& P, const T* raw, MdpState<T>& m, T* new_actions, T* target) {
  const T pi = T(3.14159265358979323846);
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    T a = zb_tanh(raw[k]);
    new_actions[k] = a;
    // walking: p_delta += pi * a * speed * dt (…env_v2.py:280-285); snake: p_delta += a * speed * dt, the pi lives
    // in its per-env joint_speed_limit (zbot_direct_6dof_snake_v0.py:121, 162-166)
    T pd = (Model::kTask != 1) ? (m.p_delta[k] + pi * a * m.speed_limit * P.step_dt)
                               : (m.p_delta[k] + a * m.speed_limit * P.step_dt);
    pd = zb_clamp(pd, -pi, pi);
    m.p_delta[k] = pd;
    target[k] = pd + Model::template default_q<T>(k);
  }
}

// One reward term (…env_v2.py:461-571).  `id` is a compile-time constant on the default-table fast
// path (the switch folds away) and a runtime value on the generic path.
template <typename T>
ZB_HD T mdp_term_value(int id, const StaleCache<T>& c, const FreshInputs<T>& f, const T* new_actions,
                       MdpState<T>& m, T heading_err, T y_err) {
  T val = T(0);
  switch (id) {
    case TERM_BASE_VEL_FORWARD:                                                // :489-491
      val = zb_tanh(T(10.0) * c.v_fwd / m.speed_limit);
      break;
    case TERM_FEET_DOWNWARD: {                                                 // :471-479
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        T dx = c.feet_z[j][0], dy = c.feet_z[j][1], dz = c.feet_z[j][2] - T(1);
        val += zb_sqrt(dx * dx + dy * dy + dz * dz);
      }
    } break;
    case TERM_FEET_FORWARD: {                                                  // :461-469
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        T dx = c.feet_x[j][0] - c.forward[0], dy = c.feet_x[j][1] - c.forward[1], dz = c.feet_x[j][2] - c.forward[2];
        val += zb_sqrt(dx * dx + dy * dy + dz * dz);
      }
    } break;
    case TERM_BASE_HEADING_X:                                                  // :481-482
      val = zb_abs(heading_err);
      break;
    case TERM_BASE_HEADING_X_SUM:                                              // :484-487
      m.heading_sum = zb_clamp(m.heading_sum + T(0.01) * heading_err, T(-1), T(1));
      val = zb_abs(m.heading_sum);
      break;
    case TERM_STEP_LENGTH: {                                                   // :509-533
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        bool down = (f.feet_force[j] > T(10.0)) && (m.feet_force_last[j] < T(10.0));
        if (down) {
          T d[3] = {c.feet_pos[j][0] - m.feet_down_pos_last[j][0], c.feet_pos[j][1] - m.feet_down_pos_last[j][1],
                    c.feet_pos[j][2] - m.feet_down_pos_last[j][2]};
          m.feet_step_length[j] = d[0] * c.forward[0] + d[1] * c.forward[1] + d[2] * c.forward[2];
          m.feet_down_pos_last[j][0] = c.feet_pos[j][0];
          m.feet_down_pos_last[j][1] = c.feet_pos[j][1];
          m.feet_down_pos_last[j][2] = c.feet_pos[j][2];
        }
        m.feet_force_last[j] = f.feet_force[j];
      }
      val = zb_tanh(T(15.0) * zb_min(m.feet_step_length[0], m.feet_step_length[1]));
    } break;
    case TERM_AIRTIME_BALANCE:                                                 // :535-539
      val = zb_abs(f.last_air_time[0] - f.last_air_time[1]);
      break;
    case TERM_ACTION_RATE: {                                                   // :502-507
      ZB_UNROLL for (int k = 0; k < 6; ++k) {
        T d = new_actions[k] - m.actions[k];
        val += d * d;
      }
    } break;
    case TERM_TORQUES: {                                                       // :558-561
      ZB_UNROLL for (int k = 0; k < 6; ++k) val += f.applied_torque[k] * f.applied_torque[k];
    } break;
    case TERM_FEET_SLIDE: {                                                    // :545-556
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        T sp = zb_sqrt(f.feet_vel_xy[j][0] * f.feet_vel_xy[j][0] + f.feet_vel_xy[j][1] * f.feet_vel_xy[j][1]);
        val += (f.feet_force[j] > T(1.0)) ? sp : T(0);
      }
    } break;
    case TERM_BASE_POS_Y_ERR:                                                  // :493-495
      val = zb_abs(c.feet_pos[0][1] + c.feet_pos[1][1] - T(2.0) * f.origin_y) + zb_abs(c.base_pos[1] - f.origin_y);
      break;
    case TERM_BASE_POS_Y_ERR_SUM:                                              // :497-500
      m.y_err_sum = zb_clamp(m.y_err_sum + T(0.01) * y_err, T(-1), T(1));
      val = zb_abs(m.y_err_sum);
      break;
    case TERM_AIRTIME_SUM:                                                     // :541-543
      val = zb_tanh(f.last_air_time[0] + f.last_air_time[1]);
      break;
    case TERM_FEET_FORCE_DIFF: {                                               // :563-565
      T sgn = (m.feet_force_sum > T(0)) ? T(1) : (m.feet_force_sum < T(0)) ? T(-1) : T(0);
      val = (f.feet_force[1] - f.feet_force[0]) * sgn;
    } break;
    case TERM_FEET_FORCE_SUM:                                                  // :567-571
      m.feet_force_sum += T(0.001) * (f.feet_force[0] - f.feet_force[1]);
      val = zb_abs(m.feet_force_sum);
      break;
    // ---- snake task (zbot_direct_6dof_snake_v0.py); heading_err = -base_heading_w.x, y_err = base_pos_x_err ----
    case TERM_SNAKE_BASE_UP_Z:                                                 // snake :304-305
      val = zb_abs(c.aux);
      break;
    case TERM_SNAKE_BASE_HEADING_Y:                                            // snake :307-308
      val = zb_abs(heading_err);
      break;
    case TERM_SNAKE_BASE_HEADING_Y_SUM:                                        // snake :310-313
      m.heading_sum = zb_clamp(m.heading_sum + T(0.01) * heading_err, T(-1), T(1));
      val = zb_abs(m.heading_sum);
      break;
    case TERM_SNAKE_BASE_POS_X_ERR:                                            // snake :329-335 -- only the first
      val = zb_abs(f.com_x_sum + T(0.636));                                    // abs(): the second is a dangling statement (SURVEY C-9)
      break;
    case TERM_SNAKE_BASE_POS_X_ERR_SUM:                                        // snake :337-340
      m.y_err_sum = zb_clamp(m.y_err_sum + T(0.01) * y_err, T(-1), T(1));
      val = zb_abs(m.y_err_sum);
      break;
    default:
      break;
  }
  return val;
}

template <typename T>
ZB_HD T mdp_reward_sum(const Params<T>& P, const StaleCache<T>& c, const FreshInputs<T>& f, const T* new_actions,
                       MdpState<T>& m, T heading_err, T y_err, bool terminated);

// …env_v2.py:384-411 + 371-382 + 461-571.  `new_actions` = this step's post-tanh actions,
// m.actions = previous step's.  Returns the reward; updates the stateful terms and ep_sums.
template <typename T>
ZB_HD T mdp_dones_rewards(const Params<T>& P, const StaleCache<T>& c, const FreshInputs<T>& f,
                          const T* new_actions, MdpState<T>& m, int64_t ep_len, bool& terminated,
                          bool& time_out) {
  time_out = ep_len >= (int64_t)(P.max_episode_length - 1);                     // :385
  bool died = f.undesired_force_max > P.contact_died_threshold;                  // :396-402
  died |= c.base_pos[2] < P.termination_height;                                  // :405
  const T y_err = c.base_pos[1] - f.origin_y;                                    // :406
  died |= zb_abs(y_err) > P.y_limit;                                             // :407
  terminated = died;
  const T heading_err = -c.forward[1];                                           // :324
  return mdp_reward_sum(P, c, f, new_actions, m, heading_err, y_err, terminated);
}

// sum_k term_k * scale_k in cfg-dict order, episode sums, -penalty when terminated (…env_v2.py:371-382)
template <typename T>
ZB_HD T mdp_reward_sum(const Params<T>& P, const StaleCache<T>& c, const FreshInputs<T>& f, const T* new_actions,
                       MdpState<T>& m, T heading_err, T y_err, bool terminated) {
  T reward = T(0);
  if (P.default_terms) {
    // the v2 table (…env_v2.py:190-206) in dict order: static indices, everything stays in registers
    ZB_UNROLL for (int i = 0; i < 13; ++i) {
      const T rew = mdp_term_value(i, c, f, new_actions, m, heading_err, y_err) * P.term_w[i];   // :375
      reward += rew;                                                             // :376
      m.ep_sums[i] += rew;                                                       // :377
    }
  } else {
    // any subset / order / weights of the 15 known terms (cfg-driven)
    for (int i = 0; i < P.num_terms; ++i) {
      const T rew = mdp_term_value(P.term_id[i], c, f, new_actions, m, heading_err, y_err) * P.term_w[i];
      reward += rew;
      // select, do not index: a dynamically indexed ep_sums[] would push the array to local memory
      ZB_UNROLL for (int k = 0; k < MAX_TERMS; ++k) m.ep_sums[k] += (k == i) ? rew : T(0);
    }
  }
  if (terminated) reward -= P.term_penalty;                                      // :379-380
  return reward;
}

// local part of _reset_idx (:423-424, 435-439, 448); the caller resets the articulation /
// sensor state and supplies the post-reset feet link positions (SURVEY C-5).
template <typename T>
ZB_HD void mdp_reset(MdpState<T>& m, const T feet_pos[2][3], int num_terms) {
  ZB_UNROLL for (int k = 0; k < 6; ++k) { m.p_delta[k] = T(0); m.actions[k] = T(0); }
  ZB_UNROLL for (int j = 0; j < 2; ++j)
    ZB_UNROLL for (int i = 0; i < 3; ++i) m.feet_down_pos_last[j][i] = feet_pos[j][i];
  m.feet_force_sum = T(0);
  m.heading_sum = T(0);
  m.y_err_sum = T(0);
  (void)num_terms;
  ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) m.ep_sums[i] = T(0);   // slots >= num_terms are always 0
  // NOT reset in v2 (SURVEY C-5): feet_force_last, feet_step_length
}

// :351-365   obs = [base_quat_w(4), q - q_default(6), qd(6), actions(6), joint_speed_limit(1)]
template <typename T>
ZB_HD void mdp_observation(const T* base_quat, const T* q, const T* qd, const T* actions, T speed_limit, T* obs) {
  ZB_UNROLL for (int i = 0; i < 4; ++i) obs[i] = base_quat[i];
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    obs[4 + k] = q[k] - default_joint_pos<T>(k);
    obs[10 + k] = qd[k];
    obs[16 + k] = actions[k];
  }
  obs[22] = speed_limit;
}

// ------------------------------------------------------------------------------------
// ContactSensor air/contact timers for one body, dt = physics dt
// ------------------------------------------------------------------------------------
template <typename T>
struct ContactTimers {
  T cur_air, cur_contact, last_air, last_contact;
};
template <typename T>
ZB_HD void contact_timers_update(ContactTimers<T>& t, bool is_contact, T dt) {
  const bool first_contact = (t.cur_air > T(0)) && is_contact;
  const bool first_detached = (t.cur_contact > T(0)) && !is_contact;
  t.last_air = first_contact ? (t.cur_air + dt) : t.last_air;
  t.cur_air = (!is_contact) ? (t.cur_air + dt) : T(0);
  t.last_contact = first_detached ? (t.cur_contact + dt) : t.last_contact;
  t.cur_contact = is_contact ? (t.cur_contact + dt) : T(0);
}

// ------------------------------------------------------------------------------------
// the whole control step of one environment (DirectRLEnv.step order)
// ------------------------------------------------------------------------------------
template <typename T>
struct EnvState {
  SimState<T> sim;
  MdpState<T> mdp;
  T carry_feet_fz[2];     // feet Fz of the LAST substep of the previous control step (history slot 4)
  T carry_mid_max;        // max |F| over the undesired bodies in that substep
  ContactTimers<T> timers[2];
};

template <typename T>
struct StepOut {
  T obs[23];
  T reward;
  bool terminated, time_out;
};

// optional export of the articulation/sensor view the MDP saw (test hook: feeds the pinned
// MDP oracle with the kernel's own physics)
template <typename T>
struct StepExport {
  LinkKin<T> k0, k1;
  T pos0[36], quat0[48], vel0[36];  // all 12 links at the start of the step (articulation order)
  T pos1[36], quat1[48], vel1[36];  // ... at the end of physics (before any reset)
  T feet_force_hist[5][2][3];  // newest first
  T mid_force_hist[5][5][3];   // newest first; slot 4 = carry (norm only: stored in [..][0])
  T applied_torque[6];
  T q1[6], qd1[6];
  T last_air[2], cur_contact[2];
};

template <typename T>
ZB_HD void env_reset(const Params<T>& P, EnvState<T>& e, const T default_feet_pos[2][3]) {
  sim_state_default(e.sim);
  mdp_reset(e.mdp, default_feet_pos, P.num_terms);
  e.carry_feet_fz[0] = e.carry_feet_fz[1] = T(0);
  e.carry_mid_max = T(0);
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    e.timers[j].cur_air = e.timers[j].cur_contact = e.timers[j].last_air = e.timers[j].last_contact = T(0);
  }
}

template <typename T>
ZB_HD void env_observe(const EnvState<T>& e, T* obs) {
  LinkKin<T> k;
  link_kinematics(e.sim, k);
  mdp_observation(k.base_quat, e.sim.q, e.sim.qd, e.mdp.actions, e.mdp.speed_limit, obs);
}

// What the physics phase hands to the MDP phase (everything else the MDP needs is re-derivable
// from the start-of-step state S0, which is still in global memory, and from the end state).
template <typename T>
struct PhysOut {
  T fz[5][2];           // feet Fz history, newest first; slot 4 = last substep of the PREVIOUS step
  T mid2;               // max over the 5 slots of |F|^2 on the undesired bodies
  T applied_torque[6];  // ImplicitActuator bookkeeping before the last substep
  T mid2_h3;            // v4 (history_length = 3): max over the LAST THREE substeps only
  T qd_prev[6];         // v4: joint velocities before the last substep (joint_acc finite difference)
  T fn2_h3[2];          // manager task: max over the last three substeps of |F_foot|^2 (feet_slide, rewards.py:256)
  T midb2_h3[5];        // manager task: the same per merged body 1..5 (undesired_contacts / illegal_contact)
};

// Phase B of the control step: _pre_physics_step (…env_v2.py:276-287) + decimation x (physics substep +
// ContactSensor.update).  Touches only e.sim, e.mdp.p_delta / speed_limit and the contact carry /
// timers, so a GPU thread can run it before the rest of the MDP state has even been loaded.
template <typename Model, int kUnroll = 1, typename T, typename Scr, typename Ground = FlatGround>
ZB_HD void env_step_physics(const Params<T>& P, EnvState<T>& e, const T* raw_actions, PhysOut<T>& po, Scr& scr,
                            StepExport<T>* ex, const Ground& ground = Ground()) {
  T new_actions[6], target[7], proc[6];
  if constexpr (Model::kTask == 3) {
    // RelativeJointPositionAction [IL-upstream]: processed = clip(raw * scale); applied at EVERY substep as
    // processed + the joint position of that moment.  Action columns are in Isaac Lab joint order.
    ZB_UNROLL for (int k = 0; k < 6; ++k) {
      proc[k] = zb_clamp(raw_actions[Model::il_col(k)] * P.act_scale, -P.act_clip, P.act_clip);
      new_actions[k] = T(0); target[k] = T(0);
    }
    target[6] = e.mdp.speed_limit;      // per-env friction coefficient (startup material randomisation)
    po.fn2_h3[0] = po.fn2_h3[1] = T(0);
    ZB_UNROLL for (int b = 0; b < 5; ++b) po.midb2_h3[b] = T(0);
  } else {
    mdp_pre_physics<Model>(P, raw_actions, e.mdp, new_actions, target);
    target[6] = T(0);
    ZB_UNROLL for (int k = 0; k < 6; ++k) proc[k] = T(0);
  }
  // (v4 keeps its commands in the carry slots: 3-deep history needs no carry-over from the previous step)
  po.fz[4][0] = (Model::kFresh) ? T(0) : e.carry_feet_fz[0];
  po.fz[4][1] = (Model::kFresh) ? T(0) : e.carry_feet_fz[1];
  po.mid2 = (Model::kFresh) ? T(0) : e.carry_mid_max * e.carry_mid_max;
  if (ex) {
    ZB_UNROLL for (int b = 0; b < 5; ++b) { ex->mid_force_hist[4][b][0] = (b == 0) ? e.carry_mid_max : T(0);
      ex->mid_force_hist[4][b][1] = T(0); ex->mid_force_hist[4][b][2] = T(0); }
    ZB_UNROLL for (int j = 0; j < 2; ++j) { ex->feet_force_hist[4][j][0] = T(0); ex->feet_force_hist[4][j][1] = T(0);
      ex->feet_force_hist[4][j][2] = e.carry_feet_fz[j]; }
  }
  SubstepOut<T> so;
  if (Model::kFresh) po.mid2_h3 = T(0);
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_UNROLL_SUB)
#endif
  for (int sub = 0; sub < P.decimation; ++sub) {
    T midf[15];
    ZB_PHASE_BAR(1);
    if (Model::kFresh) {
      if (sub == P.decimation - 1) { ZB_UNROLL for (int k = 0; k < 6; ++k) po.qd_prev[k] = e.sim.qd[k]; }
    }
    if constexpr (Model::kTask == 3) { ZB_UNROLL for (int k = 0; k < 6; ++k) target[k] = proc[k] + e.sim.q[k]; }
    physics_substep<Model, (ZB_PIPELINED_SWEEP != 0), kUnroll>(P, e.sim, target, so, scr, ex ? midf : (T*)nullptr, ground);
    if (!Model::kGroundForceSensor) continue;
    // ContactSensor.update
    const int slot = P.decimation - 1 - sub;  // newest first
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      const T* ff = so.foot_force[j];
      const T nrm = zb_sqrt(ff[0] * ff[0] + ff[1] * ff[1] + ff[2] * ff[2]);
      contact_timers_update(e.timers[j], nrm > T(1.0), P.dt);
      if constexpr (Model::kTask == 3) { if (slot < 3) po.fn2_h3[j] = zb_max(po.fn2_h3[j], nrm * nrm); }
      ZB_UNROLL for (int k = 0; k < 4; ++k) po.fz[k][j] = (slot == k) ? ff[2] : po.fz[k][j];   // select, not index
      if (ex && slot < 4) { ex->feet_force_hist[slot][j][0] = ff[0]; ex->feet_force_hist[slot][j][1] = ff[1];
        ex->feet_force_hist[slot][j][2] = ff[2]; }
    }
    if (slot < 4) po.mid2 = zb_max(po.mid2, so.mid_force2_max);
    if (Model::kFresh) { if (slot < 3) po.mid2_h3 = zb_max(po.mid2_h3, so.mid_force2_max); }
    if constexpr (Model::kPerBodyForces) { if (slot < 3) { ZB_UNROLL for (int b = 0; b < 5; ++b) po.midb2_h3[b] = zb_max(po.midb2_h3[b], so.mid_force2[b]); } }
    if (ex && slot < 4) {
      ZB_UNROLL for (int b = 0; b < 5; ++b)
        ZB_UNROLL for (int i = 0; i < 3; ++i) ex->mid_force_hist[slot][b][i] = midf[3 * b + i];
    }
  }
  if (Model::kGroundForceSensor && !Model::kFresh) {
    e.carry_feet_fz[0] = so.foot_force[0][2];
    e.carry_feet_fz[1] = so.foot_force[1][2];
    e.carry_mid_max = zb_sqrt(so.mid_force2_max);
  }
  ZB_UNROLL for (int k = 0; k < 6; ++k) po.applied_torque[k] = so.applied_torque[k];
}

// Phase C: episode counter, dones, rewards, partial reset, observation (DirectRLEnv.step steps 3-8,
// SURVEY 3.2).  `s0` = articulation state at the START of the step: the quantities the reference caches in
// _get_observations one step earlier ("stale", SURVEY C-1) are recomputed from it.  `raw_actions` again:
// tanh is recomputed instead of carrying 6 registers across the physics phase.
template <typename T>
ZB_HD void env_step_finish(const Params<T>& P, EnvState<T>& e, const SimState<T>& s0, const T* raw_actions,
                           const PhysOut<T>& po, int64_t& ep_len, const T default_feet_pos[2][3],
                           const T* default_base_quat, StepOut<T>& out, T* reset_ep_sums, StepExport<T>* ex) {
  StaleCache<T> stale;
  {
    LinkKin<T> k0;
    link_kinematics(s0, k0);
    stale_from_links(k0.base_pos, k0.base_quat, k0.base_com_vel, k0.feet_pos, k0.feet_quat, stale);
    if (ex) { ex->k0 = k0; all_link_kinematics(s0, ex->pos0, ex->quat0, ex->vel0); }
  }
  T new_actions[6];
  ZB_UNROLL for (int k = 0; k < 6; ++k) new_actions[k] = zb_tanh(raw_actions[k]);   // :278
  ep_len += 1;                                             // DirectRLEnv.step
  // fresh view
  LinkKin<T> k1;
  link_kinematics(e.sim, k1);
  FreshInputs<T> f;
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    f.feet_force[j] = ((((po.fz[0][j] + po.fz[1][j]) + po.fz[2][j]) + po.fz[3][j]) + po.fz[4][j]) / T(5);
    f.last_air_time[j] = e.timers[j].last_air;
    f.feet_vel_xy[j][0] = k1.feet_com_vel[j][0];
    f.feet_vel_xy[j][1] = k1.feet_com_vel[j][1];
  }
  f.undesired_force_max = zb_sqrt(po.mid2);
  ZB_UNROLL for (int k = 0; k < 6; ++k) f.applied_torque[k] = po.applied_torque[k];
  f.origin_y = T(0);
  if (ex) {
    ex->k1 = k1;
    all_link_kinematics(e.sim, ex->pos1, ex->quat1, ex->vel1);
    ZB_UNROLL for (int k = 0; k < 6; ++k) { ex->applied_torque[k] = po.applied_torque[k]; ex->q1[k] = e.sim.q[k]; ex->qd1[k] = e.sim.qd[k]; }
    ZB_UNROLL for (int j = 0; j < 2; ++j) { ex->last_air[j] = e.timers[j].last_air; ex->cur_contact[j] = e.timers[j].cur_contact; }
  }
  bool terminated, time_out;
  out.reward = mdp_dones_rewards(P, stale, f, new_actions, e.mdp, ep_len, terminated, time_out);
  out.terminated = terminated;
  out.time_out = time_out;
  ZB_UNROLL for (int k = 0; k < 6; ++k) e.mdp.actions[k] = new_actions[k];   // :313 (next obs pass)
  if (terminated || time_out) {
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) reset_ep_sums[i] = e.mdp.ep_sums[i];
    const T speed = e.mdp.speed_limit;
    env_reset(P, e, default_feet_pos);
    ep_len = 0;
    mdp_observation(default_base_quat, e.sim.q, e.sim.qd, e.mdp.actions, speed, out.obs);
  } else {
    mdp_observation(k1.base_quat, e.sim.q, e.sim.qd, e.mdp.actions, e.mdp.speed_limit, out.obs);
  }
}

// ------------------------------------------------------------------------------------
// snake task (zbot-6s-snake-v0): per-link quantities its MDP reads (snake_v0.py:175-208, 222-240, 329-335)
// ------------------------------------------------------------------------------------
template <typename T>
struct SnakeKin {
  T base_pos[3], base_quat[4], base_vel[3];   // link 6 = a4: body_link_pos_w / quat_w / body_link_vel_w[:, 6, :3]
  T com_x[2];                                 // body_com_pos_w[:, 0, 0] and [:, 11, 0] (a1 and b6)
  T self_pen_max;                             // max sphere overlap over the 14 filtered self-contact pairs
};

template <typename T>
ZB_HD void snake_kinematics(const SimState<T>& s, SnakeKin<T>& o, bool want_self) {
  using namespace model_snake;
  T Q[4] = {s.Q[0], s.Q[1], s.Q[2], s.Q[3]};
  T r[3] = {T(0), T(0), T(0)};
  T w[3] = {s.w[0], s.w[1], s.w[2]};
  T vO[3] = {s.v[0], s.v[1], s.v[2]};
  T ctr[12][3];   // self-contact sphere centres, articulation order a1 b1 a2 b2 a3 b3 a4 b4 a5 b5 a6 b6
  ZB_UNROLL for (int b = 0; b < 7; ++b) {
    T R[9];
    quat_to_mat(Q, R);
    // sphere centres lie on the body's z axis: origin + z * R[:,2]
    if (b == 0) {
      ZB_UNROLL for (int i = 0; i < 3; ++i) ctr[0][i] = r[i] + R[3 * i + 2] * T(CENTRE_A_Z);
      o.com_x[0] = s.p[0] + r[0] + R[0] * T(A1_COM_X) + R[2] * T(A1_COM_Z);
    } else if (b == 6) {
      ZB_UNROLL for (int i = 0; i < 3; ++i) ctr[11][i] = r[i] + R[3 * i + 2] * T(CENTRE_B_Z);
      o.com_x[1] = s.p[0] + r[0] + R[0] * T(B6_COM_X) + R[2] * T(B6_COM_Z);
    } else {
      ZB_UNROLL for (int i = 0; i < 3; ++i) {
        ctr[2 * b - 1][i] = r[i] + R[3 * i + 2] * T(CENTRE_B_Z);
        ctr[2 * b][i] = r[i] + R[3 * i + 2] * (T(LINK_Z) + T(CENTRE_A_Z));
      }
    }
    if (b == 3) {   // a4 = second link of body 3: origin + (0,0,LINK_Z), frame rotated 180 deg about z
      T lo[3] = {r[0] + R[2] * T(LINK_Z), r[1] + R[5] * T(LINK_Z), r[2] + R[8] * T(LINK_Z)};
      T wxl[3];
      cross3(w, lo, wxl);
      ZB_UNROLL for (int i = 0; i < 3; ++i) { o.base_pos[i] = s.p[i] + lo[i]; o.base_vel[i] = vO[i] + wxl[i]; }
      o.base_quat[0] = -Q[3]; o.base_quat[1] = Q[2]; o.base_quat[2] = -Q[1]; o.base_quat[3] = Q[0];   // Q (x) (0,0,0,1)
    }
    if (b < 6) {
      const T jz = (b == 0) ? T(model::JOINT_Z_FIRST) : T(model::JOINT_Z_REST);
      const T sg = (b & 1) ? T(-model::AXIS_S) : T(model::AXIS_S);
      r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
      T a[3] = {sg * R[0] + T(model::AXIS_S) * R[2], sg * R[3] + T(model::AXIS_S) * R[5], sg * R[6] + T(model::AXIS_S) * R[8]};
      T m[3];
      cross3(r, a, m);
      ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] += a[i] * s.qd[b]; vO[i] += m[i] * s.qd[b]; }
      T sn, cs;
      zb_sincos(T(0.5) * s.q[b], &sn, &cs);
      quat_mul_joint(Q, cs, sg * sn, T(model::AXIS_S) * sn);
    }
  }
  o.self_pen_max = T(0);
  if (want_self) {
    // filtered contact sensors (snake_v0.py:23-48): a1|{b4 a5 b5 a6 b6}, b6|{a3 b2 a2 b1}, b1|{a5 b5 a6}, a6|{b2 a2}
    const int pa[14] = {0, 0, 0, 0, 0, 11, 11, 11, 11, 1, 1, 1, 10, 10};
    const int pb[14] = {7, 8, 9, 10, 11, 4, 3, 2, 1, 8, 9, 10, 3, 2};
    T pen = T(0);
    ZB_UNROLL for (int k = 0; k < 14; ++k) {
      const T dx = ctr[pa[k]][0] - ctr[pb[k]][0], dy = ctr[pa[k]][1] - ctr[pb[k]][1], dz = ctr[pa[k]][2] - ctr[pb[k]][2];
      pen = zb_max(pen, T(2) * T(SPHERE_R) - zb_sqrt(dx * dx + dy * dy + dz * dz));
    }
    o.self_pen_max = pen;
  }
}

template <typename Model, typename T>
ZB_HD void env_reset_model(const Params<T>& P, EnvState<T>& e, const T default_feet_pos[2][3]) {
  sim_state_default<Model>(e.sim);
  mdp_reset(e.mdp, default_feet_pos, P.num_terms);
  e.carry_feet_fz[0] = e.carry_feet_fz[1] = T(0);
  e.carry_mid_max = T(0);
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    e.timers[j].cur_air = e.timers[j].cur_contact = e.timers[j].last_air = e.timers[j].last_contact = T(0);
  }
}

template <typename T>
ZB_HD void snake_observe(const EnvState<T>& e, T* obs) {
  SnakeKin<T> k;
  snake_kinematics(e.sim, k, false);
  ZB_UNROLL for (int i = 0; i < 4; ++i) obs[i] = k.base_quat[i];
  ZB_UNROLL for (int j = 0; j < 6; ++j) { obs[4 + j] = e.sim.q[j]; obs[10 + j] = e.sim.qd[j]; obs[16 + j] = e.mdp.actions[j]; }
  obs[22] = e.mdp.speed_limit;
}

// self-contact proxy exported for the parity test: what the MDP saw
template <typename T>
struct SnakeExport {
  T base_pos0[3], base_quat0[4], base_vel0[3];
  T base_pos1[3], base_quat1[4], base_vel1[3];
  T com_x1[2], self_force1;
  T q1[6], qd1[6], tau1[6];
};

// Phase C of the snake task's control step (same ordering as the walking task)
template <typename T>
ZB_HD void snake_step_finish(const Params<T>& P, EnvState<T>& e, const SimState<T>& s0, const T* raw_actions,
                             const PhysOut<T>& po, int64_t& ep_len, const T* default_base_quat, StepOut<T>& out,
                             T* reset_ep_sums, SnakeExport<T>* ex) {
  StaleCache<T> stale;
  SnakeKin<T> k0;
  snake_kinematics(s0, k0, false);
  {
    const T heading_vec[3] = {T(0), T(-1), T(0)}, up_vec[3] = {T(-1), T(0), T(0)};   // snake_v0.py:118-119
    T up[3];
    quat_apply(k0.base_quat, heading_vec, stale.forward);                            // base_heading_w (:180)
    quat_apply(k0.base_quat, up_vec, up);                                            // base_up_w (:181)
    stale.aux = up[1];
    ZB_UNROLL for (int i = 0; i < 3; ++i) stale.base_pos[i] = k0.base_pos[i];
    stale.v_fwd = k0.base_vel[0] * stale.forward[0] + k0.base_vel[1] * stale.forward[1] + k0.base_vel[2] * stale.forward[2];  // :184
  }
  T new_actions[6];
  ZB_UNROLL for (int k = 0; k < 6; ++k) new_actions[k] = zb_tanh(raw_actions[k]);
  ep_len += 1;
  SnakeKin<T> k1;
  snake_kinematics(e.sim, k1, true);
  FreshInputs<T> f;
  f.feet_force[0] = f.feet_force[1] = T(0);
  f.last_air_time[0] = f.last_air_time[1] = T(0);
  f.feet_vel_xy[0][0] = f.feet_vel_xy[0][1] = f.feet_vel_xy[1][0] = f.feet_vel_xy[1][1] = T(0);
  // filtered self-contact force proxy: contact spring x sphere overlap (not fed back into the dynamics)
  f.undesired_force_max = P.c_k * k1.self_pen_max;
  ZB_UNROLL for (int k = 0; k < 6; ++k) f.applied_torque[k] = po.applied_torque[k];
  f.origin_y = T(0);
  f.com_x_sum = k1.com_x[0] + k1.com_x[1];                                           // env-local: origin_x = 0
  if (ex) {
    ZB_UNROLL for (int i = 0; i < 3; ++i) { ex->base_pos0[i] = k0.base_pos[i]; ex->base_vel0[i] = k0.base_vel[i];
                                            ex->base_pos1[i] = k1.base_pos[i]; ex->base_vel1[i] = k1.base_vel[i]; }
    ZB_UNROLL for (int i = 0; i < 4; ++i) { ex->base_quat0[i] = k0.base_quat[i]; ex->base_quat1[i] = k1.base_quat[i]; }
    ex->com_x1[0] = k1.com_x[0]; ex->com_x1[1] = k1.com_x[1]; ex->self_force1 = f.undesired_force_max;
    ZB_UNROLL for (int k = 0; k < 6; ++k) { ex->q1[k] = e.sim.q[k]; ex->qd1[k] = e.sim.qd[k]; ex->tau1[k] = po.applied_torque[k]; }
  }
  const bool time_out = ep_len >= (int64_t)(P.max_episode_length - 1);               // snake :223
  const T x_err = stale.base_pos[0] + T(0.318);                                      // snake :236 (origin_x = 0)
  const bool died = (f.undesired_force_max > P.contact_died_threshold) || (zb_abs(x_err) > T(0.2));   // :228-239
  const T heading_err = -stale.forward[0];                                           // snake :182
  out.reward = mdp_reward_sum(P, stale, f, new_actions, e.mdp, heading_err, x_err, died);
  out.terminated = died;
  out.time_out = time_out;
  ZB_UNROLL for (int k = 0; k < 6; ++k) e.mdp.actions[k] = new_actions[k];
  if (died || time_out) {
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) reset_ep_sums[i] = e.mdp.ep_sums[i];
    const T speed = e.mdp.speed_limit;
    const T zero_feet[2][3] = {{T(0), T(0), T(0)}, {T(0), T(0), T(0)}};
    env_reset_model<ModelSnake>(P, e, zero_feet);
    e.mdp.speed_limit = speed;
    ep_len = 0;
    ZB_UNROLL for (int i = 0; i < 4; ++i) out.obs[i] = default_base_quat[i];
    ZB_UNROLL for (int j = 0; j < 6; ++j) { out.obs[4 + j] = e.sim.q[j]; out.obs[10 + j] = e.sim.qd[j]; out.obs[16 + j] = e.mdp.actions[j]; }
    out.obs[22] = speed;
  } else {
    ZB_UNROLL for (int i = 0; i < 4; ++i) out.obs[i] = k1.base_quat[i];
    ZB_UNROLL for (int j = 0; j < 6; ++j) { out.obs[4 + j] = e.sim.q[j]; out.obs[10 + j] = e.sim.qd[j]; out.obs[16 + j] = e.mdp.actions[j]; }
    out.obs[22] = e.mdp.speed_limit;
  }
}

// whole snake control step in one call (CPU port, export kernel)
template <typename T, typename Scr>
ZB_HD void snake_env_step(const Params<T>& P, EnvState<T>& e, const T* raw_actions, int64_t& ep_len,
                          const T* default_base_quat, StepOut<T>& out, T* reset_ep_sums, SnakeExport<T>* ex, Scr& scr) {
  const SimState<T> s0 = e.sim;
  PhysOut<T> po;
  env_step_physics<ModelSnake>(P, e, raw_actions, po, scr, (StepExport<T>*)nullptr);
  snake_step_finish(P, e, s0, raw_actions, po, ep_len, default_base_quat, out, reset_ep_sums, ex);
}

// ------------------------------------------------------------------------------------
// zbot-6b-walking-v4 (tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py): velocity / heading commands with
// reset- and interval-mode resampling, randomised reset pose, 24-wide observation, 15 reward terms on FRESH
// (end-of-physics) quantities, 3-deep contact history.  Same robot, actuator and physics substep as v2.
// State reuse inside the 80-word layout: carry_feet_fz[0..1] = commands[0..1], carry_mid_max = target_heading_yaw,
// heading_sum = current_yaw (what the last observation used), y_err_sum = interval_command_resample time_left.
// ------------------------------------------------------------------------------------
constexpr int V4_NUM_RAND = 10;
// uniform [0,1) numbers of one env-step: the draws the reference makes through torch.rand / torch.bernoulli
enum V4RandSlot : int { VR_POSE_X = 0, VR_POSE_Y = 1, VR_POSE_YAW = 2,          // reset_root_state_uniform (:84-86)
                        VR_RESET_SIGN = 3, VR_RESET_VEL = 4, VR_RESET_YAW = 5,  // resample_commands, mode "reset"
                        VR_INT_TIME = 6,                                        // EventManager interval re-arm
                        VR_INT_SIGN = 7, VR_INT_VEL = 8, VR_INT_YAW = 9 };      // resample_commands, mode "interval"

// isaaclab.utils.math.wrap_to_pi: ((a + pi) mod 2 pi) - pi with python-style modulo, +pi kept for a > 0
template <typename T>
ZB_HD T v4_wrap_to_pi(T a) {
  const T pi = T(3.14159265358979323846), two_pi = T(2) * T(3.14159265358979323846);
  T r = zb_fmod(a + pi, two_pi);
  if (r != T(0) && r < T(0)) r += two_pi;
  return (r == T(0) && a > T(0)) ? pi : (r - pi);
}

// resample_commands (…env_v4.py:109-135)
template <typename T>
ZB_HD void v4_resample_commands(const Params<T>& P, T u_sign, T u_vel, T u_yaw, T current_yaw, T& cmd0, T& cmd1,
                                T& target_yaw) {
  const T low = P.ev_vel_lo;
  if (P.ev_dual_sign) {
    const T vel_sign = ((u_sign < P.ev_prob_pos) ? T(1) : T(0)) * T(2) - T(1);   // torch.bernoulli(prob_pos) * 2 - 1
    const T high = P.ev_vel_hi + P.ev_offset * (vel_sign - T(1));               // :126
    cmd0 = (u_vel * (high - low) + low) * vel_sign;                              // :127
  } else {
    cmd0 = u_vel * (P.ev_vel_hi - low) + low;                                    // :129
  }
  cmd1 = u_yaw * (P.ev_yaw_hi - P.ev_yaw_lo) + P.ev_yaw_lo;                      // :133-134
  target_yaw = v4_wrap_to_pi(current_yaw + cmd1);                                // :136
}

template <typename T>
struct V4Fresh {   // _compute_intermediate_values (…env_v4.py:792-826) + the sensor / articulation reads of the terms
  T forward[3], shoulder[3], current_yaw, heading_err, v_fwd, vel_y;
  T feet_pos[2][3], feet_force[2];
  T last_air[2], last_contact[2], cur_air[2], cur_contact[2];
  T joint_vel2, joint_acc2, base_height;
};

template <typename T>
ZB_HD T v4_term_value(int id, const StaleCache<T>& c, const FreshInputs<T>& f, const V4Fresh<T>& v, const T* new_actions,
                      MdpState<T>& m, T cmd0) {
  T val = T(0);
  switch (id) {
    case TERM_V4_TRACK_LIN_VEL_X: {                                              // :1013-1016
      const T d = cmd0 - v.v_fwd;
      val = zb_exp(-(d * d) / T(0.25));
    } break;
    case TERM_V4_TRACK_HEADING_YAW:                                              // :1018-1020
      val = zb_exp(-(v.heading_err * v.heading_err) / T(0.25));
      break;
    case TERM_V4_LIN_VEL_X:                                                      // :1022-1024
      val = v.v_fwd * v.v_fwd;
      break;
    case TERM_V4_LIN_VEL_Y:                                                      // :1026-1030
      val = v.vel_y * v.vel_y;
      break;
    case TERM_V4_JOINT_VEL:                                                      // :1168-1171
      val = v.joint_vel2;
      break;
    case TERM_V4_JOINT_ACC:                                                      // :1173-1176
      val = v.joint_acc2;
      break;
    case TERM_V4_STEP_LENGTH: {                                                  // :1056-1079
      const T sgn = (cmd0 > T(0)) ? T(1) : (cmd0 < T(0)) ? T(-1) : T(0);
      bool down[2];
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        down[j] = (f.feet_force[j] > T(10.0)) && (m.feet_force_last[j] < T(10.0));
        if (down[j]) {
          const T d[3] = {v.feet_pos[j][0] - m.feet_down_pos_last[j][0], v.feet_pos[j][1] - m.feet_down_pos_last[j][1],
                          v.feet_pos[j][2] - m.feet_down_pos_last[j][2]};
          m.feet_step_length[j] = (d[0] * v.forward[0] + d[1] * v.forward[1] + d[2] * v.forward[2]) * sgn;
        }
      }
      const T rew_len = zb_min(m.feet_step_length[0], m.feet_step_length[1]);
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        m.feet_step_length[j] *= T(0.99);                                        // decay (:1073)
        if (down[j]) { ZB_UNROLL for (int i = 0; i < 3; ++i) m.feet_down_pos_last[j][i] = v.feet_pos[j][i]; }
        m.feet_force_last[j] = f.feet_force[j];
      }
      val = zb_tanh(T(15.0) * rew_len);
    } break;
    case TERM_V4_FEET_AIR_TIME_BIPED: {                                          // :1110-1124
      const bool c0 = v.cur_contact[0] > T(0), c1 = v.cur_contact[1] > T(0);
      const T t0 = c0 ? v.cur_contact[0] : v.cur_air[0], t1 = c1 ? v.cur_contact[1] : v.cur_air[1];
      const bool single = (c0 != c1);
      val = zb_min(zb_min(single ? t0 : T(0), single ? t1 : T(0)), T(2.0));
    } break;
    case TERM_V4_AIRTIME_VARIANCE: {                                             // :1081-1087 (torch.var, unbiased, n = 2)
      const T a0 = zb_min(v.last_air[0], T(0.5)), a1 = zb_min(v.last_air[1], T(0.5));
      const T b0 = zb_min(v.last_contact[0], T(0.5)), b1 = zb_min(v.last_contact[1], T(0.5));
      const T ma = (a0 + a1) * T(0.5), mb = (b0 + b1) * T(0.5);
      val = ((a0 - ma) * (a0 - ma) + (a1 - ma) * (a1 - ma)) + ((b0 - mb) * (b0 - mb) + (b1 - mb) * (b1 - mb));
    } break;
    case TERM_V4_AIRTIME_SUM:                                                    // :1089-1093
      val = zb_min(v.last_air[0] + v.last_air[1], T(2.0));
      break;
    case TERM_V4_FEET_HARMONY:                                                   // :1139-1144
      val = (v.last_air[0] + v.last_air[1]) - T(3.0) * zb_abs(v.last_air[0] - v.last_air[1]);
      break;
    case TERM_V4_FEET_CLOSE: {                                                   // :1146-1151
      const T dx = v.feet_pos[0][0] - v.feet_pos[1][0], dy = v.feet_pos[0][1] - v.feet_pos[1][1];
      val = zb_max(T(0.115) - zb_sqrt(dx * dx + dy * dy), T(0));
    } break;
    case TERM_V4_FEET_HEIGHT:                                                    // :1178-1183
      val = v.feet_pos[0][2] + (v.feet_pos[1][2] - T(0.053));
      break;
    case TERM_V4_BASE_HEIGHT:                                                    // :1185-1187 (env-local: origin z = 0)
      val = v.base_height - T(0.25);
      break;
    default:   // feet_downward, feet_forward, action_rate, torques, feet_slide: the v2 formulas on fresh inputs
      val = mdp_term_value(id, c, f, new_actions, m, T(0), T(0));
      break;
  }
  return val;
}

template <typename T>
struct V4Export {   // what the v4 MDP saw at the end of physics (test hook), 60 words
  T base_pos[3], base_quat[4], base_link_vel[3];
  T feet_pos[2][3], feet_quat[2][4], feet_com_vel[2][3];
  T feet_fz_hist[3][2];       // newest first
  T undesired_max;            // max_t,b |F| over the 3-deep history
  T last_air[2], last_contact[2], cur_air[2], cur_contact[2];
  T q1[6], qd1[6], tau1[6], joint_acc1[6];
};

// Phase C of the v4 control step.  `rnd` = this env-step's V4_NUM_RAND uniforms.  obs has 24 entries.
template <typename T>
ZB_HD void v4_step_finish(const Params<T>& P, EnvState<T>& e, const T* raw_actions, const PhysOut<T>& po, int64_t& ep_len,
                          const T* rnd, T* obs24, StepOut<T>& out, T* reset_ep_sums, V4Export<T>* ex) {
  T& cmd0 = e.carry_feet_fz[0];
  T& cmd1 = e.carry_feet_fz[1];
  T& target_yaw = e.carry_mid_max;
  T& time_left = e.mdp.y_err_sum;
  T new_actions[6];
  ZB_UNROLL for (int k = 0; k < 6; ++k) new_actions[k] = zb_tanh(raw_actions[k]);
  ep_len += 1;
  // ---- _compute_intermediate_values (:792-826): everything FRESH ----
  LinkKin<T> k1;
  link_kinematics(e.sim, k1);
  StaleCache<T> c;     // the v2 container, filled with fresh values (feet_x / feet_z / forward / feet_pos)
  stale_from_links(k1.base_pos, k1.base_quat, k1.base_link_vel, k1.feet_pos, k1.feet_quat, c);
  V4Fresh<T> v;
  {
    const T ez[3] = {T(0), T(0), T(1)};
    quat_apply(k1.base_quat, ez, v.shoulder);
    ZB_UNROLL for (int i = 0; i < 3; ++i) v.forward[i] = c.forward[i];
    v.current_yaw = zb_atan2(v.forward[1], v.forward[0]);                         // :806
    const T diff = target_yaw - v.current_yaw;
    v.heading_err = zb_atan2(zb_sin(diff), zb_cos(diff));                         // :809
    v.v_fwd = c.v_fwd;                                                            // :812-813 (link velocity . forward)
    v.vel_y = k1.base_link_vel[0] * v.shoulder[0] + k1.base_link_vel[1] * v.shoulder[1] + k1.base_link_vel[2] * v.shoulder[2];
    v.base_height = k1.base_pos[2];
    v.joint_vel2 = T(0); v.joint_acc2 = T(0);
    const T inv_dt = T(1) / P.dt;
    ZB_UNROLL for (int k = 0; k < 6; ++k) {
      v.joint_vel2 += e.sim.qd[k] * e.sim.qd[k];
      const T acc = (e.sim.qd[k] - po.qd_prev[k]) * inv_dt;
      v.joint_acc2 += acc * acc;
    }
  }
  FreshInputs<T> f;
  ZB_UNROLL for (int j = 0; j < 2; ++j) {
    f.feet_force[j] = ((po.fz[0][j] + po.fz[1][j]) + po.fz[2][j]) / T(3);         // mean over history_length = 3 (:822-825)
    f.last_air_time[j] = e.timers[j].last_air;
    f.feet_vel_xy[j][0] = k1.feet_com_vel[j][0];
    f.feet_vel_xy[j][1] = k1.feet_com_vel[j][1];
    v.feet_force[j] = f.feet_force[j];
    v.last_air[j] = e.timers[j].last_air; v.last_contact[j] = e.timers[j].last_contact;
    v.cur_air[j] = e.timers[j].cur_air; v.cur_contact[j] = e.timers[j].cur_contact;
    ZB_UNROLL for (int i = 0; i < 3; ++i) v.feet_pos[j][i] = k1.feet_pos[j][i];
  }
  f.undesired_force_max = zb_sqrt(po.mid2_h3);
  ZB_UNROLL for (int k = 0; k < 6; ++k) f.applied_torque[k] = po.applied_torque[k];
  f.origin_y = T(0);
  f.com_x_sum = T(0);
  if (ex) {
    ZB_UNROLL for (int i = 0; i < 3; ++i) { ex->base_pos[i] = k1.base_pos[i]; ex->base_link_vel[i] = k1.base_link_vel[i]; }
    ZB_UNROLL for (int i = 0; i < 4; ++i) ex->base_quat[i] = k1.base_quat[i];
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      ZB_UNROLL for (int i = 0; i < 3; ++i) { ex->feet_pos[j][i] = k1.feet_pos[j][i]; ex->feet_com_vel[j][i] = k1.feet_com_vel[j][i]; }
      ZB_UNROLL for (int i = 0; i < 4; ++i) ex->feet_quat[j][i] = k1.feet_quat[j][i];
      ZB_UNROLL for (int t = 0; t < 3; ++t) ex->feet_fz_hist[t][j] = po.fz[t][j];
      ex->last_air[j] = v.last_air[j]; ex->last_contact[j] = v.last_contact[j];
      ex->cur_air[j] = v.cur_air[j]; ex->cur_contact[j] = v.cur_contact[j];
    }
    ex->undesired_max = f.undesired_force_max;
    ZB_UNROLL for (int k = 0; k < 6; ++k) {
      ex->q1[k] = e.sim.q[k]; ex->qd1[k] = e.sim.qd[k]; ex->tau1[k] = po.applied_torque[k];
      ex->joint_acc1[k] = (e.sim.qd[k] - po.qd_prev[k]) * (T(1) / P.dt);
    }
  }
  // ---- _get_dones (:868-886) ----
  const bool time_out = ep_len >= (int64_t)(P.max_episode_length - 1);
  const bool died = (f.undesired_force_max > P.contact_died_threshold) || (k1.base_pos[2] < P.termination_height);
  // ---- _get_rewards (:853-866): rew = f() * scale * step_dt, in cfg-dict order ----
  T reward = T(0);
  for (int i = 0; i < P.num_terms; ++i) {
    const T rew = v4_term_value(P.term_id[i], c, f, v, new_actions, e.mdp, cmd0) * P.term_w[i] * P.step_dt;
    reward += rew;
    ZB_UNROLL for (int k = 0; k < MAX_TERMS; ++k) e.mdp.ep_sums[k] += (k == i) ? rew : T(0);
  }
  if (died) reward -= P.term_penalty;
  out.reward = reward;
  out.terminated = died;
  out.time_out = time_out;
  ZB_UNROLL for (int k = 0; k < 6; ++k) e.mdp.actions[k] = new_actions[k];
  T current_yaw = v.current_yaw;
  // ---- _reset_idx (:888-976) ----
  if (died || time_out) {
    // log: episodic sum per second of the ACTUAL episode duration (:893-901)
    const T dur = zb_max(T(ep_len) * P.step_dt, P.step_dt);
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) reset_ep_sums[i] = e.mdp.ep_sums[i] / dur;
    // EventManager mode "reset", cfg order: reset_base (:84-106), [curricula: host], reset_command_resample
    sim_state_default<ModelWalk>(e.sim);
    const T yaw = rnd[VR_POSE_YAW] * (P.ev_pose_hi[2] - P.ev_pose_lo[2]) + P.ev_pose_lo[2];
    e.sim.p[0] += rnd[VR_POSE_X] * (P.ev_pose_hi[0] - P.ev_pose_lo[0]) + P.ev_pose_lo[0];
    e.sim.p[1] += rnd[VR_POSE_Y] * (P.ev_pose_hi[1] - P.ev_pose_lo[1]) + P.ev_pose_lo[1];
    // quat_mul(default (1,0,0,0), quat_from_euler_xyz(0, 0, yaw)) = (cos(yaw/2), 0, 0, sin(yaw/2))
    e.sim.Q[0] = zb_cos(yaw * T(0.5)); e.sim.Q[1] = T(0); e.sim.Q[2] = T(0); e.sim.Q[3] = zb_sin(yaw * T(0.5));
    current_yaw = yaw;                                                            // env.current_yaw[env_ids] = rand yaw (:88)
    v4_resample_commands(P, rnd[VR_RESET_SIGN], rnd[VR_RESET_VEL], rnd[VR_RESET_YAW], current_yaw, cmd0, cmd1, target_yaw);
    ep_len = 0;
    ZB_UNROLL for (int k = 0; k < 6; ++k) { e.mdp.p_delta[k] = T(0); e.mdp.actions[k] = T(0); }
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      e.timers[j].cur_air = e.timers[j].cur_contact = e.timers[j].last_air = e.timers[j].last_contact = T(0);
      e.mdp.feet_force_last[j] = T(15.0);                                         // :966
      e.mdp.feet_step_length[j] = T(0);                                           // :970
    }
    e.mdp.feet_force_sum = T(0);
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) e.mdp.ep_sums[i] = T(0);
    link_kinematics(e.sim, k1);                                                   // post-reset view
    ZB_UNROLL for (int j = 0; j < 2; ++j)
      ZB_UNROLL for (int i = 0; i < 3; ++i) e.mdp.feet_down_pos_last[j][i] = k1.feet_pos[j][i];   // :967-969
  }
  // ---- EventManager mode "interval" (per-env timer, [IL-upstream] EventManager.apply) ----
  time_left -= P.step_dt;
  if (time_left < T(1e-6)) {
    time_left = rnd[VR_INT_TIME] * (P.ev_interval_hi - P.ev_interval_lo) + P.ev_interval_lo;
    v4_resample_commands(P, rnd[VR_INT_SIGN], rnd[VR_INT_VEL], rnd[VR_INT_YAW], current_yaw, cmd0, cmd1, target_yaw);
  }
  e.mdp.heading_sum = current_yaw;
  // ---- _get_observations (:828-851) ----
  const T diff = target_yaw - current_yaw;
  ZB_UNROLL for (int i = 0; i < 4; ++i) obs24[i] = k1.base_quat[i];
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    obs24[4 + k] = e.sim.q[k] - default_joint_pos<T>(k);
    obs24[10 + k] = e.sim.qd[k];
    obs24[16 + k] = e.mdp.actions[k];
  }
  obs24[22] = cmd0;
  obs24[23] = zb_atan2(zb_sin(diff), zb_cos(diff));
}

// whole v4 control step in one call (CPU port)
template <typename T, typename Scr>
ZB_HD void v4_env_step(const Params<T>& P, EnvState<T>& e, const T* raw_actions, int64_t& ep_len, const T* rnd, T* obs24,
                       StepOut<T>& out, T* reset_ep_sums, V4Export<T>* ex, Scr& scr) {
  PhysOut<T> po;
  env_step_physics<ModelWalkV4>(P, e, raw_actions, po, scr, (StepExport<T>*)nullptr);
  v4_step_finish(P, e, raw_actions, po, ep_len, rnd, obs24, out, reset_ep_sums, ex);
}

// ------------------------------------------------------------------------------------
// zbot-6b-walking-m-v0: the manager-based task (tasks/zbotlab_manager/zbotlab_env_cfg.py:82-452, config/zbot6b_manager/
// flat_env_cfg.py, mdp/rewards.py, mdp/terminations.py) in ManagerBasedRLEnv.step order [IL-upstream]: action term ->
// decimation x (apply action, physics, sensor update) -> episode_length += 1 -> TerminationManager -> RewardManager ->
// reset of the done envs (EventManager mode "reset", then the managers' own reset: the command term resamples) ->
// CommandManager.compute -> ObservationManager (25 columns, additive uniform noise applied by the caller).
// Robot = ModelWalkM.  State reuse inside the 80-word layout: carry_feet_fz[0..1] = command lin_vel x / y, carry_mid_max =
// command ang_vel z, heading_sum = is_standing_env, y_err_sum = command time_left, speed_limit = per-env friction,
// actions = last RAW action (ActionManager.action, Isaac Lab joint order); p_delta is unused.
// ------------------------------------------------------------------------------------
constexpr int M_NUM_OBS = 25;
constexpr int M_NUM_RAND = 22;
// uniforms of one env-step, in the order the reference would draw them for that env
enum MRandSlot : int { MR_POSE_X = 0, MR_POSE_Y = 1, MR_POSE_YAW = 2,                       // reset_root_state_uniform
                       MR_RESET_TIME = 3, MR_RESET_VX = 4, MR_RESET_VY = 5, MR_RESET_WZ = 6, MR_RESET_STAND = 7,   // CommandTerm.reset
                       MR_INT_TIME = 8, MR_INT_VX = 9, MR_INT_VY = 10, MR_INT_WZ = 11, MR_INT_STAND = 12,          // CommandTerm.compute
                       MR_RESET_HEADING = 13, MR_RESET_ISHEAD = 14, MR_INT_HEADING = 15, MR_INT_ISHEAD = 16,       // heading_command=True
                       MR_PUSH_RESET_TIME = 17, MR_PUSH_TIME = 18, MR_PUSH_X = 19, MR_PUSH_Y = 20,                 // EventTerm push_robot
                       MR_TERRAIN_LEVEL = 21 };                                                                    // terrain curriculum: random restart level

template <typename T>
ZB_HD void quat_mul(const T* a, const T* b, T* o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

template <typename T>
struct MKin {   // what the manager terms read from `robot.data` (root = LINK `base`, feet = LINKS foot0 / foot1)
  T root_pos[3], root_quat[4], root_lin_vel[3], root_ang_vel[3], shoulder[3];   // shoulder = root_quat applied to (0,1,0)
  T feet_pos[2][3], feet_up[2][3], feet_x[2][3], feet_com_vel[2][3];           // feet_up = link (0,+-1,0), feet_x = link (1,0,0)
  T feet_bq[2][4];                                                             // BODY (chain-frame) quaternions of the feet
};

// Link axes in the body (chain) frame: x_link = -x_c, y_link = z_c, z_link = y_c (assets/zbot_6s_v2.py); `base` differs
// from that by a rotation about its own y axis, which leaves the shoulder axis y_link = z_c in place.
template <typename T>
ZB_HD void m_kinematics(const SimState<T>& s, MKin<T>& o) {
  using namespace model;
  T Q[4] = {s.Q[0], s.Q[1], s.Q[2], s.Q[3]};
  T r[3] = {T(0), T(0), T(0)};
  T w[3] = {s.w[0], s.w[1], s.w[2]};
  T vO[3] = {s.v[0], s.v[1], s.v[2]};
  T qk[6], qdk[6];
  ZB_UNROLL for (int k = 0; k < 6; ++k) { qk[k] = s.q[k]; qdk[k] = s.qd[k]; }
  T R[9];
  quat_to_mat(Q, R);
  {  // body 0 = foot0: link origin at (0,0,FOOT0_LINK_Z), CoM = body CoM
    const T lz = T(model_m::FOOT0_LINK_Z);
    const T cx = T(model_m::FOOT0_CX), cy = T(model_m::FOOT0_CY), cz = T(model_m::FOOT0_CZ);
    T c[3] = {R[0] * cx + R[1] * cy + R[2] * cz, R[3] * cx + R[4] * cy + R[5] * cz, R[6] * cx + R[7] * cy + R[8] * cz};
    T wxc[3];
    cross3(w, c, wxc);
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      o.feet_pos[0][i] = s.p[i] + R[3 * i + 2] * lz;
      o.feet_com_vel[0][i] = vO[i] + wxc[i];
      o.feet_up[0][i] = R[3 * i + 2];
      o.feet_x[0][i] = -R[3 * i];
    }
    ZB_UNROLL for (int i = 0; i < 4; ++i) o.feet_bq[0][i] = Q[i];
  }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = 0; k < 6; ++k) {
    const T jz = (k == 0) ? T(JOINT_Z_FIRST) : T(JOINT_Z_REST);
    const T sg = (k & 1) ? T(-AXIS_S) : T(AXIS_S);
    r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
    T a[3] = {sg * R[0] + T(AXIS_S) * R[2], sg * R[3] + T(AXIS_S) * R[5], sg * R[6] + T(AXIS_S) * R[8]};
    T m[3];
    cross3(r, a, m);
    const T qv = (k == 0) ? qk[0] : (k == 1) ? qk[1] : (k == 2) ? qk[2] : (k == 3) ? qk[3] : (k == 4) ? qk[4] : qk[5];
    const T qdv = (k == 0) ? qdk[0] : (k == 1) ? qdk[1] : (k == 2) ? qdk[2] : (k == 3) ? qdk[3] : (k == 4) ? qdk[4] : qdk[5];
    ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] += a[i] * qdv; vO[i] += m[i] * qdv; }
    T sn, cs;
    zb_sincos(T(0.5) * qv, &sn, &cs);
    quat_mul_joint(Q, cs, sg * sn, T(AXIS_S) * sn);
    quat_to_mat(Q, R);
    if (k == 2) {   // body 3 = base + a7: `base` link origin at (0,0,BASE_LINK_Z)
      const T lz = T(model_m::BASE_LINK_Z);
      T lo[3] = {r[0] + R[2] * lz, r[1] + R[5] * lz, r[2] + R[8] * lz};
      T wxl[3];
      cross3(w, lo, wxl);
      const T lq[4] = {T(model_m::BASE_LQW), T(model_m::BASE_LQX), T(model_m::BASE_LQY), T(model_m::BASE_LQZ)};
      quat_mul(Q, lq, o.root_quat);
      ZB_UNROLL for (int i = 0; i < 3; ++i) {
        o.root_pos[i] = s.p[i] + lo[i];
        o.root_lin_vel[i] = vO[i] + wxl[i];
        o.root_ang_vel[i] = w[i];
        o.shoulder[i] = R[3 * i + 2];
      }
    }
    if (k == 5) {   // body 6 = foot1: link origin = body origin
      const T cx = T(model_m::FOOT1_CX), cy = T(model_m::FOOT1_CY), cz = T(model_m::FOOT1_CZ);
      T c[3] = {r[0] + R[0] * cx + R[1] * cy + R[2] * cz, r[1] + R[3] * cx + R[4] * cy + R[5] * cz,
                r[2] + R[6] * cx + R[7] * cy + R[8] * cz};
      T wxc[3];
      cross3(w, c, wxc);
      ZB_UNROLL for (int i = 0; i < 3; ++i) {
        o.feet_pos[1][i] = s.p[i] + r[i];
        o.feet_com_vel[1][i] = vO[i] + wxc[i];
        o.feet_up[1][i] = -R[3 * i + 2];
        o.feet_x[1][i] = -R[3 * i];
      }
      ZB_UNROLL for (int i = 0; i < 4; ++i) o.feet_bq[1][i] = Q[i];
    }
  }
}

// UniformVelocityCommand._resample [IL-upstream]: time_left, lin_vel_x / y, ang_vel_z ~ U(range); standing ~ U(0,1) <= rel
template <typename T>
ZB_HD void m_resample_command(const Params<T>& P, T u_time, T ux, T uy, T uw, T u_stand, T& cmd0, T& cmd1, T& cmd2,
                              T& standing, T& time_left) {
  time_left = u_time * (P.cmd_resample_hi - P.cmd_resample_lo) + P.cmd_resample_lo;
  cmd0 = ux * (P.cmd_hi[0] - P.cmd_lo[0]) + P.cmd_lo[0];
  cmd1 = uy * (P.cmd_hi[1] - P.cmd_lo[1]) + P.cmd_lo[1];
  cmd2 = uw * (P.cmd_hi[2] - P.cmd_lo[2]) + P.cmd_lo[2];
  standing = (u_stand <= P.cmd_rel_standing) ? T(1) : T(0);
}
// ... with heading_command=True it also draws heading_target ~ U(ranges.heading) and is_heading_env ~ U(0,1) <= rel_heading_envs
template <typename T>
ZB_HD void m_resample_heading(const Params<T>& P, T u_heading, T u_is, T& heading_target, T& is_heading) {
  heading_target = u_heading * (P.cmd_heading_hi - P.cmd_heading_lo) + P.cmd_heading_lo;
  is_heading = (u_is <= P.cmd_rel_heading) ? T(1) : T(0);
}

template <typename T>
struct MFresh {
  T cmd[3], forward[3];
  T root_lin_vel[3], root_ang_z, yaw_c, yaw_s;   // (cos, sin) of the root yaw
  T feet_pos[2][3], feet_force[2], feet_fnorm_max[2];
  T cur_air[2], cur_contact[2];
  T phase_time;                                  // episode_length_buf * step_dt (float32 product)
  T midb_max[5];                                 // max over the 3-deep history of |F| on each merged body 1..5
};

template <typename T>
ZB_HD T m_term_value(int id, const T* par, const StaleCache<T>& c, const FreshInputs<T>& f, const V4Fresh<T>& v4,
                     const MFresh<T>& v, const T* new_actions, MdpState<T>& m) {
  T val = T(0);
  switch (id) {
    case TERM_M_TRACK_LIN_VEL_XY_EXP: {                                          // rewards.py:286-297
      const T vx = v.yaw_c * v.root_lin_vel[0] + v.yaw_s * v.root_lin_vel[1];   // quat_apply_inverse(yaw_quat(root_quat_w), v)
      const T vy = v.yaw_c * v.root_lin_vel[1] - v.yaw_s * v.root_lin_vel[0];
      const T dx = v.cmd[0] - vx, dy = v.cmd[1] - vy;
      val = zb_exp(-(dx * dx + dy * dy) / par[0]);
    } break;
    case TERM_M_TRACK_ANG_VEL_Z_EXP: {                                           // :300-309
      const T d = v.cmd[2] - v.root_ang_z;
      val = zb_exp(-(d * d) / par[0]);
    } break;
    case TERM_M_FOOT_STEP_LENGTH: {                                              // :45-107 (command_name = None)
      const T inv = T(1) / (zb_sqrt(v.forward[0] * v.forward[0] + v.forward[1] * v.forward[1] + v.forward[2] * v.forward[2]) + T(1e-6));
      const T fw[3] = {v.forward[0] * inv, v.forward[1] * inv, v.forward[2] * inv};
      bool down[2];
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        down[j] = (v.feet_force[j] > T(10.0)) && (m.feet_force_last[j] < T(10.0));
        if (down[j]) {
          const T d[3] = {v.feet_pos[j][0] - m.feet_down_pos_last[j][0], v.feet_pos[j][1] - m.feet_down_pos_last[j][1],
                          v.feet_pos[j][2] - m.feet_down_pos_last[j][2]};
          m.feet_step_length[j] = zb_abs(d[0] * fw[0] + d[1] * fw[1] + d[2] * fw[2]);
        }
      }
      const T rew_len = zb_min(m.feet_step_length[0], m.feet_step_length[1]);
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        if (down[j]) { ZB_UNROLL for (int i = 0; i < 3; ++i) m.feet_down_pos_last[j][i] = v.feet_pos[j][i]; }
        m.feet_force_last[j] = v.feet_force[j];
      }
      val = zb_tanh(T(15.0) * rew_len);
    } break;
    case TERM_M_GAIT: {                                                          // :155-186, par = period, offsets, threshold
      const T g = zb_fmod(v.phase_time, par[0]) / par[0];
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        const T ph = zb_fmod(g + par[1 + j], T(1.0));
        const bool stance = ph < par[3], contact = v.cur_contact[j] > T(0);
        val += (stance == contact) ? T(1) : T(0);
      }
      const T n = zb_sqrt(v.cmd[0] * v.cmd[0] + v.cmd[1] * v.cmd[1] + v.cmd[2] * v.cmd[2]);
      val = (n > T(0.05)) ? val : T(0);
    } break;
    case TERM_M_FEET_SLIDE:                                                      // :247-262
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        const T sp = zb_sqrt(f.feet_vel_xy[j][0] * f.feet_vel_xy[j][0] + f.feet_vel_xy[j][1] * f.feet_vel_xy[j][1]);
        val += (v.feet_fnorm_max[j] > T(1.0)) ? sp : T(0);
      }
      break;
    case TERM_M_FEET_CLEARANCE: {                                                // :143-153, par = std, tanh_mult, target
      T acc = T(0);
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        const T dz = v.feet_pos[j][2] - par[2];
        const T sp = zb_sqrt(f.feet_vel_xy[j][0] * f.feet_vel_xy[j][0] + f.feet_vel_xy[j][1] * f.feet_vel_xy[j][1]);
        acc += (dz * dz) * zb_tanh(par[1] * sp);
      }
      val = zb_exp(-acc / par[0]);
    } break;
    case TERM_M_FEET_AIR_TIME_BIPED: {                                           // :211-230, par = threshold
      const bool c0 = v.cur_contact[0] > T(0), c1 = v.cur_contact[1] > T(0);
      const T t0 = c0 ? v.cur_contact[0] : v.cur_air[0], t1 = c1 ? v.cur_contact[1] : v.cur_air[1];
      const bool single = (c0 != c1);
      val = zb_min(zb_min(single ? t0 : T(0), single ? t1 : T(0)), par[0]);
      val = (zb_sqrt(v.cmd[0] * v.cmd[0] + v.cmd[1] * v.cmd[1]) > T(0.1)) ? val : T(0);
    } break;
    case TERM_M_BASE_VEL_FORWARD:                                                // :264-274, par = which_forward
      val = par[0] * (v.root_lin_vel[0] * v.forward[0] + v.root_lin_vel[1] * v.forward[1] + v.root_lin_vel[2] * v.forward[2]);
      break;
    case TERM_M_FEET_FORCE_PATTERN: {                                            // :276-284
      const T sg = (m.feet_force_sum > T(0)) ? T(1) : (m.feet_force_sum < T(0)) ? T(-1) : T(0);
      const T diff = (v.feet_force[1] - v.feet_force[0]) * sg;
      m.feet_force_sum += T(0.001) * (v.feet_force[0] - v.feet_force[1]);
      val = T(0.5) * diff - T(0.1) * zb_abs(m.feet_force_sum);
    } break;
    case TERM_M_UNDESIRED_CONTACTS:                                              // [IL-upstream] sum_b( max_t |F_b| > threshold )
      ZB_UNROLL for (int b = 0; b < 5; ++b) val += (v.midb_max[b] > par[0]) ? T(1) : T(0);
      break;
    default:   // ids shared with the direct tasks (torques, joint_acc, action_rate, foot_downward / forward, air-time terms)
      val = v4_term_value(id, c, f, v4, new_actions, m, T(0));
      break;
  }
  return val;
}

template <typename T>
struct MExport {   // what the manager terms saw at the end of physics (test hook), 72 words; feet_quat = LINK quaternions
  T root_pos[3], root_quat[4], root_lin_vel[3], root_ang_vel[3];
  T feet_pos[2][3], feet_quat[2][4], feet_com_vel[2][3];
  T feet_fz_hist[3][2], feet_fnorm_max[2];
  T last_air[2], last_contact[2], cur_air[2], cur_contact[2];
  T q1[6], tau1[6], joint_acc1[6];     // chain joint order
  T midb_max[5];                       // max over the 3-deep history of |F| on the merged bodies 1..5
};
constexpr int M_EXPORT_WORDS = 13 + 20 + 8 + 8 + 18 + 5;

// Rough-terrain context of one env (zbot-6b-walking-m-rough-v0): its origin on the tile grid (world frame; the state's root
// position is relative to it) and what the `terrain_levels_vel` curriculum needs (mdp/curriculums.py:26-55 +
// TerrainImporter.update_env_origins [IL-upstream]).  State words: p_delta[3] = terrain level, p_delta[4] = terrain type.
template <typename T>
struct MTerrainCtx {
  T origin[3];                 // in: this step's origin; out: the origin after a reset moved the env to another tile
  const float* tile_origins;   // [rows][cols][3]
  int rows, cols;
  T tile_size, episode_s;
  int curriculum;              // CurriculumCfg.terrain_levels present
};

// Phase C of the manager task's control step.  `rnd` = this env-step's M_NUM_RAND uniforms; obs25 = the clean row.
template <typename T>
ZB_HD void m_step_finish(const Params<T>& P, EnvState<T>& e, const T* raw_actions, const PhysOut<T>& po, int64_t& ep_len,
                         const T* rnd, T* obs25, StepOut<T>& out, T* reset_ep_sums, MExport<T>* ex,
                         MTerrainCtx<T>* terrain = nullptr) {
  T& cmd0 = e.carry_feet_fz[0];
  T& cmd1 = e.carry_feet_fz[1];
  T& cmd2 = e.carry_mid_max;
  T& standing = e.mdp.heading_sum;
  T& time_left = e.mdp.y_err_sum;
  T& heading_target = e.mdp.p_delta[0];     // p_delta is unused by this task's action term: four of its words carry
  T& is_heading = e.mdp.p_delta[1];         // the heading command state and the push_robot interval timer
  T& push_left = e.mdp.p_delta[2];
  ep_len += 1;
  MKin<T> k1;
  m_kinematics(e.sim, k1);
  // ---- the shared containers of the direct tasks, filled with this task's fresh values ----
  StaleCache<T> c;
  FreshInputs<T> f;
  V4Fresh<T> v4;
  MFresh<T> v;
  {
    const T grav[3] = {T(0), T(0), T(-1)};
    cross3(grav, k1.shoulder, v.forward);                                       // cross(GRAVITY_VEC_W, R (0,1,0)), not normalised
    const T qw = k1.root_quat[0], qx = k1.root_quat[1], qy = k1.root_quat[2], qz = k1.root_quat[3];
    const T sy = T(2) * (qw * qz + qx * qy), cy = T(1) - T(2) * (qy * qy + qz * qz);   // yaw_quat [IL-upstream]
    const T inv = zb_rsqrt(zb_max(sy * sy + cy * cy, T(1e-30)));
    v.yaw_c = cy * inv; v.yaw_s = sy * inv;
    v.cmd[0] = cmd0; v.cmd[1] = cmd1; v.cmd[2] = cmd2;
    v.root_ang_z = k1.root_ang_vel[2];
    v.phase_time = T(ep_len) * P.step_dt;
    v4.joint_vel2 = T(0); v4.joint_acc2 = T(0);
    const T inv_dt = T(1) / P.dt;
    ZB_UNROLL for (int k = 0; k < 6; ++k) {
      const T acc = (e.sim.qd[k] - po.qd_prev[k]) * inv_dt;                      // ArticulationData.joint_acc [IL-upstream]
      v4.joint_vel2 += e.sim.qd[k] * e.sim.qd[k];
      v4.joint_acc2 += acc * acc;
      f.applied_torque[k] = po.applied_torque[k];
    }
    ZB_UNROLL for (int i = 0; i < 3; ++i) { v.root_lin_vel[i] = k1.root_lin_vel[i]; c.forward[i] = v.forward[i]; c.base_pos[i] = k1.root_pos[i]; }
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      v.feet_force[j] = ((po.fz[0][j] + po.fz[1][j]) + po.fz[2][j]) / T(3);      // mean over history_length = 3
      v.feet_fnorm_max[j] = zb_sqrt(po.fn2_h3[j]);
      v.cur_air[j] = e.timers[j].cur_air; v.cur_contact[j] = e.timers[j].cur_contact;
      v4.last_air[j] = e.timers[j].last_air; v4.last_contact[j] = e.timers[j].last_contact;
      v4.cur_air[j] = v.cur_air[j]; v4.cur_contact[j] = v.cur_contact[j];
      f.feet_force[j] = v.feet_force[j];
      f.last_air_time[j] = e.timers[j].last_air;
      f.feet_vel_xy[j][0] = k1.feet_com_vel[j][0]; f.feet_vel_xy[j][1] = k1.feet_com_vel[j][1];
      v4.feet_force[j] = v.feet_force[j];
      ZB_UNROLL for (int i = 0; i < 3; ++i) {
        v.feet_pos[j][i] = k1.feet_pos[j][i]; v4.feet_pos[j][i] = k1.feet_pos[j][i]; c.feet_pos[j][i] = k1.feet_pos[j][i];
        c.feet_z[j][i] = k1.feet_up[j][i]; c.feet_x[j][i] = k1.feet_x[j][i];
      }
    }
    ZB_UNROLL for (int i = 0; i < 3; ++i) { v4.forward[i] = v.forward[i]; v4.shoulder[i] = k1.shoulder[i]; }
    v4.current_yaw = T(0); v4.heading_err = T(0); v4.v_fwd = T(0); v4.vel_y = T(0); v4.base_height = k1.root_pos[2];
    c.v_fwd = T(0); c.aux = T(0);
    f.undesired_force_max = zb_sqrt(po.mid2_h3);
    f.origin_y = T(0); f.com_x_sum = T(0);
    ZB_UNROLL for (int b = 0; b < 5; ++b) v.midb_max[b] = zb_sqrt(po.midb2_h3[b]);
  }
  if (ex) {
    ZB_UNROLL for (int i = 0; i < 3; ++i) { ex->root_pos[i] = k1.root_pos[i]; ex->root_lin_vel[i] = k1.root_lin_vel[i]; ex->root_ang_vel[i] = k1.root_ang_vel[i]; }
    ZB_UNROLL for (int i = 0; i < 4; ++i) ex->root_quat[i] = k1.root_quat[i];
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      ZB_UNROLL for (int i = 0; i < 3; ++i) { ex->feet_pos[j][i] = k1.feet_pos[j][i]; ex->feet_com_vel[j][i] = k1.feet_com_vel[j][i]; }
      const T lq[4] = {T(model_m::LINK_LQW), T(model_m::LINK_LQX), T(model_m::LINK_LQY), T(model_m::LINK_LQZ)};
      quat_mul(k1.feet_bq[j], lq, ex->feet_quat[j]);
      ZB_UNROLL for (int t = 0; t < 3; ++t) ex->feet_fz_hist[t][j] = po.fz[t][j];
      ex->feet_fnorm_max[j] = v.feet_fnorm_max[j];
      ex->last_air[j] = v4.last_air[j]; ex->last_contact[j] = v4.last_contact[j];
      ex->cur_air[j] = v.cur_air[j]; ex->cur_contact[j] = v.cur_contact[j];
    }
    ZB_UNROLL for (int k = 0; k < 6; ++k) { ex->q1[k] = e.sim.q[k]; ex->tau1[k] = po.applied_torque[k];
                                            ex->joint_acc1[k] = (e.sim.qd[k] - po.qd_prev[k]) * (T(1) / P.dt); }
    ZB_UNROLL for (int b = 0; b < 5; ++b) ex->midb_max[b] = v.midb_max[b];
  }
  // ---- TerminationManager.compute (TerminationsCfg, zbotlab_env_cfg.py:371-393; base_contact is None for this robot) ----
  const bool time_out = ep_len >= (int64_t)P.max_episode_length;                 // mdp.time_out [IL-upstream]: >= max_episode_length
  // root_height_below_minimum reads root_pos_w: the WORLD height (on a rough tile the origin's z is not 0)
  const bool low = (k1.root_pos[2] + (terrain ? terrain->origin[2] : T(0))) < P.termination_height;
  bool close = false;
  if (P.feet_close_min > T(0)) {                                                 // terminations.py:55-60
    const T dx = k1.feet_pos[0][0] - k1.feet_pos[1][0], dy = k1.feet_pos[0][1] - k1.feet_pos[1][1], dz = k1.feet_pos[0][2] - k1.feet_pos[1][2];
    close = zb_sqrt(dx * dx + dy * dy + dz * dz) < P.feet_close_min;
  }
  bool illegal = false;                                                          // mdp.illegal_contact [IL-upstream]: any selected body with max_t |F| > threshold
  if (P.illegal_thr > T(0)) {
    ZB_UNROLL for (int b = 0; b < 5; ++b) illegal |= ((P.illegal_mask >> b) & 1) && (v.midb_max[b] > P.illegal_thr);
  }
  const bool died = low || close || illegal;
  // ---- RewardManager.compute: value = func(env) * weight * dt, in cfg order; raw actions are the action term's `action` ----
  T reward = T(0);
  for (int i = 0; i < P.num_terms; ++i) {
    const T par[4] = {P.term_par[i][0], P.term_par[i][1], P.term_par[i][2], P.term_par[i][3]};
    const T rew = m_term_value(P.term_id[i], par, c, f, v4, v, raw_actions, e.mdp) * P.term_w[i] * P.step_dt;
    reward += rew;
    ZB_UNROLL for (int k = 0; k < MAX_TERMS; ++k) e.mdp.ep_sums[k] += (k == i) ? rew : T(0);
  }
  if (died) reward += P.term_penalty_w;                                          // RewTerm is_terminated
  out.reward = reward;
  out.terminated = died;
  out.time_out = time_out;
  ZB_UNROLL for (int k = 0; k < 6; ++k) e.mdp.actions[k] = raw_actions[k];      // ActionManager.action (prev_action next step)
  // ---- _reset_idx of the done envs ----
  if (died || time_out) {
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) reset_ep_sums[i] = e.mdp.ep_sums[i];
    if (P.num_terms <= MAX_TERMS - 3) {   // spare slots: the termination_penalty term's episodic sum and the per-DoneTerm counts
      reset_ep_sums[MAX_TERMS - 3] = died ? P.term_penalty_w : T(0);
      reset_ep_sums[MAX_TERMS - 2] = low ? T(1) : T(0);
      reset_ep_sums[MAX_TERMS - 1] = close ? T(1) : T(0);
    }
    // the same four, always, for the statistics words 22..25 (a cfg with more than 13 terms has no spare term slots)
    reset_ep_sums[MAX_TERMS + 0] = died ? P.term_penalty_w : T(0);
    reset_ep_sums[MAX_TERMS + 1] = low ? T(1) : T(0);
    reset_ep_sums[MAX_TERMS + 2] = close ? T(1) : T(0);
    reset_ep_sums[MAX_TERMS + 3] = illegal ? T(1) : T(0);
    if (terrain && terrain->curriculum) {
      // CurriculumManager.compute runs first in _reset_idx: terrain_levels_vel (mdp/curriculums.py:26-55) on the pre-reset
      // root position and command, then TerrainImporter.update_env_origins [IL-upstream]
      const T dist = zb_sqrt(k1.root_pos[0] * k1.root_pos[0] + k1.root_pos[1] * k1.root_pos[1]);   // |root_pos_w.xy - env_origin.xy|
      const bool up = dist > terrain->tile_size * T(0.5);
      const bool down = !up && (dist < zb_sqrt(cmd0 * cmd0 + cmd1 * cmd1) * terrain->episode_s * T(0.5));
      int level = (int)e.mdp.p_delta[3] + (up ? 1 : 0) - (down ? 1 : 0);
      if (level >= terrain->rows) level = zb_min_i((int)(rnd[MR_TERRAIN_LEVEL] * T(terrain->rows)), terrain->rows - 1);
      level = level < 0 ? 0 : level;
      e.mdp.p_delta[3] = T(level);
      const float* o = terrain->tile_origins + ((size_t)level * terrain->cols + (int)e.mdp.p_delta[4]) * 3;
      terrain->origin[0] = T(o[0]); terrain->origin[1] = T(o[1]); terrain->origin[2] = T(o[2]);
    }
    // EventManager mode "reset": reset_base (reset_root_state_uniform on the root LINK `base`), reset_robot_joints
    // (default * U(1,1)), reset_my_data (rewards.py:37-43)
    sim_state_default<ModelWalkM>(e.sim);
    {
      using namespace model_m;
      const T yaw = rnd[MR_POSE_YAW] * (P.ev_pose_hi[2] - P.ev_pose_lo[2]) + P.ev_pose_lo[2];
      const T bx = T(BASE_DEFAULT_X) + (rnd[MR_POSE_X] * (P.ev_pose_hi[0] - P.ev_pose_lo[0]) + P.ev_pose_lo[0]);
      const T by = T(BASE_DEFAULT_Y) + (rnd[MR_POSE_Y] * (P.ev_pose_hi[1] - P.ev_pose_lo[1]) + P.ev_pose_lo[1]);
      const T cyw = zb_cos(yaw), syw = zb_sin(yaw);
      const T ox = T(DEFAULT_ROOT_X - BASE_DEFAULT_X), oy = T(DEFAULT_ROOT_Y - BASE_DEFAULT_Y);   // chain root relative to `base`
      e.sim.p[0] = bx + (cyw * ox - syw * oy);
      e.sim.p[1] = by + (syw * ox + cyw * oy);
      const T qz[4] = {zb_cos(yaw * T(0.5)), T(0), T(0), zb_sin(yaw * T(0.5))};
      const T q0[4] = {e.sim.Q[0], e.sim.Q[1], e.sim.Q[2], e.sim.Q[3]};
      quat_mul(qz, q0, e.sim.Q);
    }
    ep_len = 0;
    ZB_UNROLL for (int k = 0; k < 6; ++k) e.mdp.actions[k] = T(0);                // ActionManager.reset (p_delta carries command / event / terrain state)
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      e.timers[j].cur_air = e.timers[j].cur_contact = e.timers[j].last_air = e.timers[j].last_contact = T(0);
      e.mdp.feet_force_last[j] = T(0);
      e.mdp.feet_step_length[j] = T(0);
    }
    e.mdp.feet_force_sum = T(0);
    ZB_UNROLL for (int i = 0; i < MAX_TERMS; ++i) e.mdp.ep_sums[i] = T(0);
    m_kinematics(e.sim, k1);                                                     // post-reset view
    ZB_UNROLL for (int j = 0; j < 2; ++j)
      ZB_UNROLL for (int i = 0; i < 3; ++i) e.mdp.feet_down_pos_last[j][i] = k1.feet_pos[j][i];
    // CommandManager.reset -> UniformVelocityCommand._resample
    m_resample_command(P, rnd[MR_RESET_TIME], rnd[MR_RESET_VX], rnd[MR_RESET_VY], rnd[MR_RESET_WZ], rnd[MR_RESET_STAND],
                       cmd0, cmd1, cmd2, standing, time_left);
    if (P.cmd_heading) m_resample_heading(P, rnd[MR_RESET_HEADING], rnd[MR_RESET_ISHEAD], heading_target, is_heading);
    // EventManager.reset: the interval term's per-env timer is re-drawn
    if (P.push_interval_hi > T(0)) push_left = rnd[MR_PUSH_RESET_TIME] * (P.push_interval_hi - P.push_interval_lo) + P.push_interval_lo;
  }
  // ---- CommandManager.compute(dt): timer, resample, standing envs get a zero command ----
  time_left -= P.step_dt;
  if (time_left <= T(0)) {
    m_resample_command(P, rnd[MR_INT_TIME], rnd[MR_INT_VX], rnd[MR_INT_VY], rnd[MR_INT_WZ], rnd[MR_INT_STAND],
                       cmd0, cmd1, cmd2, standing, time_left);
    if (P.cmd_heading) m_resample_heading(P, rnd[MR_INT_HEADING], rnd[MR_INT_ISHEAD], heading_target, is_heading);
  }
  if (P.cmd_heading && is_heading != T(0)) {
    // UniformVelocityCommand._update_command [IL-upstream]: ang_vel_z = clip(stiffness * wrap_to_pi(target - heading_w), range),
    // heading_w = atan2 of the root's x axis in the world (ArticulationData.heading_w)
    const T qw = k1.root_quat[0], qx = k1.root_quat[1], qy = k1.root_quat[2], qz = k1.root_quat[3];
    const T fx = T(1) - T(2) * (qy * qy + qz * qz), fy = T(2) * (qx * qy + qw * qz);
    const T d = heading_target - zb_atan2(fy, fx);
    const T err = zb_atan2(zb_sin(d), zb_cos(d));
    cmd2 = zb_clamp(P.cmd_heading_stiffness * err, P.cmd_lo[2], P.cmd_hi[2]);
  }
  if (standing != T(0)) { cmd0 = T(0); cmd1 = T(0); cmd2 = T(0); }
  // ---- EventManager.apply(mode="interval"): push_robot = push_by_setting_velocity [IL-upstream]: root velocity += U(range),
  //      i.e. every link of the robot gains the same linear velocity ----
  if (P.push_interval_hi > T(0)) {
    push_left -= P.step_dt;
    if (push_left < T(1e-6)) {
      push_left = rnd[MR_PUSH_TIME] * (P.push_interval_hi - P.push_interval_lo) + P.push_interval_lo;
      e.sim.v[0] += rnd[MR_PUSH_X] * (P.push_hi[0] - P.push_lo[0]) + P.push_lo[0];
      e.sim.v[1] += rnd[MR_PUSH_Y] * (P.push_hi[1] - P.push_lo[1]) + P.push_lo[1];
    }
  }
  // ---- ObservationManager: base_quat, velocity_commands, joint_pos_rel, joint_vel_rel, last_action (Isaac Lab joint order) ----
  ZB_UNROLL for (int i = 0; i < 4; ++i) obs25[i] = k1.root_quat[i];
  obs25[4] = cmd0; obs25[5] = cmd1; obs25[6] = cmd2;
  ZB_UNROLL for (int k = 0; k < 6; ++k) {
    const int col = ModelWalkM::il_col(k);
    obs25[7 + col] = e.sim.q[k] - ModelWalkM::default_q<T>(k);
    obs25[13 + col] = e.sim.qd[k];
    obs25[19 + k] = e.mdp.actions[k];
  }
}

// whole manager-task control step in one call (CPU port)
template <typename T, typename Scr>
ZB_HD void m_env_step(const Params<T>& P, EnvState<T>& e, const T* raw_actions, int64_t& ep_len, const T* rnd, T* obs25,
                      StepOut<T>& out, T* reset_ep_sums, MExport<T>* ex, Scr& scr) {
  PhysOut<T> po;
  env_step_physics<ModelWalkM>(P, e, raw_actions, po, scr, (StepExport<T>*)nullptr);
  m_step_finish(P, e, raw_actions, po, ep_len, rnd, obs25, out, reset_ep_sums, ex);
}

// the whole control step in one call (CPU port, export kernel)
template <typename T, typename Scr>
ZB_HD void env_step(const Params<T>& P, EnvState<T>& e, const T* raw_actions, int64_t& ep_len,
                    const T default_feet_pos[2][3], const T* default_base_quat, StepOut<T>& out,
                    T* reset_ep_sums /*MAX_TERMS, valid when a reset happened*/, StepExport<T>* ex, Scr& scr) {
  const SimState<T> s0 = e.sim;
  PhysOut<T> po;
  env_step_physics<ModelWalk>(P, e, raw_actions, po, scr, ex);
  env_step_finish(P, e, s0, raw_actions, po, ep_len, default_feet_pos, default_base_quat, out, reset_ep_sums, ex);
}

}  // namespace zbot
