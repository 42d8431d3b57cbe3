// TEST INFRASTRUCTURE / CPU BASELINE -- host build of the per-env step math.
//
// Compiles zbot_lab_b200/csrc/zbot_core.h (the same templates the sm_100a kernels
// instantiate with T = float) for the CPU with T = float and T = double and loops over
// envs with OpenMP.  Two uses, both outside the product path:
//   1. tests/: lets the build box (no GPU) check the kernel's arithmetic against the
//      independent float64 oracle (oracle/dyn_oracle.py) and the pinned MDP oracle;
//   2. bench.py cpu_baseline / --impl reference: "the path on the host cores", standing in
//      for the reference's torch + PhysX-CPU step (PhysX is closed and absent).
// Nothing under zbot_lab_b200/ links or loads this file; the product fails loudly without
// its CUDA library.  State layout: [N][ZBOT_STATE_WORDS] (AoS on the CPU).
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../zbot_lab_b200/csrc/zbot_layout.h"
#include "../zbot_lab_b200/csrc/zbot_pair.h"
#include "../zbot_lab_b200/csrc/zbot_halves.h"
#include "../zbot_lab_b200/csrc/zbot_h2.h"
#include <type_traits>

using namespace zbot;

static int g_threads = 0;   // 0: OpenMP default
static inline int port_threads() { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

template <typename T>
static void default_pose_constants(T feet_pos[2][3], T base_quat[4]) {
  SimState<T> s;
  sim_state_default(s);
  LinkKin<T> k;
  link_kinematics(s, k);
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 3; ++i) feet_pos[j][i] = k.feet_pos[j][i];
  for (int i = 0; i < 4; ++i) base_quat[i] = k.base_quat[i];
}

template <typename T>
static int port_step(const ZbotCfg* cfg, T* state, int64_t* ep_len, const T* actions, T* obs, T* rew,
                     uint8_t* term, uint8_t* trunc, T* reset_sums /*[N][16] or null*/,
                     T* export_buf /*[N][EXPORT_WORDS] or null*/, int n) {
  const char* why = nullptr;
  if (cfg_validate(*cfg, &why) != ZBOT_OK) return ZBOT_E_INVALID;
  Params<T> P;
  params_from_cfg(*cfg, P);
  T dfp[2][3], dbq[4];
  default_pose_constants(dfp, dbq);
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; ++e) {
    EnvState<T> es;
    env_state_unpack(state + (size_t)e * ZBOT_STATE_WORDS, es);
    StepOut<T> out;
    T rs[MAX_TERMS];
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = T(0);
    StepExport<T> ex;
    ArrayScratch<T> scr;
    env_step(P, es, actions + (size_t)e * 6, ep_len[e], dfp, dbq, out, rs, export_buf ? &ex : (StepExport<T>*)nullptr, scr);
    env_state_pack(es, state + (size_t)e * ZBOT_STATE_WORDS);
    for (int i = 0; i < 23; ++i) obs[(size_t)e * 23 + i] = out.obs[i];
    rew[e] = out.reward;
    term[e] = out.terminated ? 1 : 0;
    trunc[e] = out.time_out ? 1 : 0;
    if (reset_sums)
      for (int i = 0; i < MAX_TERMS; ++i) reset_sums[(size_t)e * MAX_TERMS + i] = rs[i];
    if (export_buf) memcpy(export_buf + (size_t)e * (sizeof(StepExport<T>) / sizeof(T)), &ex, sizeof(ex));
  }
  return ZBOT_OK;
}

// snake task: whole control step
template <typename T>
static int port_snake_step(const ZbotCfg* cfg, T* state, int64_t* ep_len, const T* actions, T* obs, T* rew,
                           uint8_t* term, uint8_t* trunc, T* reset_sums, T* export_buf, int n) {
  const char* why = nullptr;
  if (cfg_validate(*cfg, &why) != ZBOT_OK) return ZBOT_E_INVALID;
  Params<T> P;
  params_from_cfg(*cfg, P);
  SimState<T> sd;
  sim_state_default<ModelSnake>(sd);
  SnakeKin<T> kd;
  snake_kinematics(sd, kd, false);
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; ++e) {
    EnvState<T> es;
    env_state_unpack(state + (size_t)e * ZBOT_STATE_WORDS, es);
    StepOut<T> out;
    T rs[MAX_TERMS];
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = T(0);
    SnakeExport<T> ex;
    ArrayScratch<T> scr;
    snake_env_step(P, es, actions + (size_t)e * 6, ep_len[e], kd.base_quat, out, rs,
                   export_buf ? &ex : (SnakeExport<T>*)nullptr, scr);
    env_state_pack(es, state + (size_t)e * ZBOT_STATE_WORDS);
    for (int i = 0; i < 23; ++i) obs[(size_t)e * 23 + i] = out.obs[i];
    rew[e] = out.reward;
    term[e] = out.terminated ? 1 : 0;
    trunc[e] = out.time_out ? 1 : 0;
    if (reset_sums)
      for (int i = 0; i < MAX_TERMS; ++i) reset_sums[(size_t)e * MAX_TERMS + i] = rs[i];
    if (export_buf) memcpy(export_buf + (size_t)e * (sizeof(SnakeExport<T>) / sizeof(T)), &ex, sizeof(ex));
  }
  return ZBOT_OK;
}

// zbot-6b-walking-v4: whole control step; rnd [N][V4_NUM_RAND] uniforms, obs [N][24]
template <typename T>
static int port_v4_step(const ZbotCfg* cfg, T* state, int64_t* ep_len, const T* actions, const T* rnd, T* obs, T* rew,
                        uint8_t* term, uint8_t* trunc, T* reset_sums, T* export_buf, int n) {
  const char* why = nullptr;
  if (cfg_validate(*cfg, &why) != ZBOT_OK || cfg->task != ZBOT_TASK_WALKING_V4) return ZBOT_E_INVALID;
  Params<T> P;
  params_from_cfg(*cfg, P);
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; ++e) {
    EnvState<T> es;
    env_state_unpack(state + (size_t)e * ZBOT_STATE_WORDS, es);
    StepOut<T> out;
    T rs[MAX_TERMS];
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = T(0);
    V4Export<T> ex;
    ArrayScratch<T> scr;
    v4_env_step(P, es, actions + (size_t)e * 6, ep_len[e], rnd + (size_t)e * V4_NUM_RAND, obs + (size_t)e * ZBOT_V4_NUM_OBS, out,
                rs, export_buf ? &ex : (V4Export<T>*)nullptr, scr);
    env_state_pack(es, state + (size_t)e * ZBOT_STATE_WORDS);
    rew[e] = out.reward;
    term[e] = out.terminated ? 1 : 0;
    trunc[e] = out.time_out ? 1 : 0;
    if (reset_sums)
      for (int i = 0; i < MAX_TERMS; ++i) reset_sums[(size_t)e * MAX_TERMS + i] = rs[i];
    if (export_buf) memcpy(export_buf + (size_t)e * ZBOT_V4_EXPORT_WORDS, &ex, sizeof(ex));
  }
  return ZBOT_OK;
}

// rough ground of the manager task (host twin of zbot_bind_terrain); heights == nullptr / t == nullptr: flat
struct PortTerrain {
  const float* heights;
  int nx, ny;
  float x0, y0, cell;
  const float* tile_origins;
  int rows, cols;
  float tile_size;
  float* env_origins;     // [N][4]
  int curriculum;
};

// zbot-6b-walking-m-v0: whole control step; rnd [N][M_NUM_RAND] uniforms, obs [N][25] (no observation noise here)
template <typename T>
static int port_m_step(const ZbotCfg* cfg, T* state, int64_t* ep_len, const T* actions, const T* rnd, T* obs, T* rew,
                       uint8_t* term, uint8_t* trunc, T* reset_sums, T* export_buf, int n, const PortTerrain* pt = nullptr) {
  const char* why = nullptr;
  if (cfg_validate(*cfg, &why) != ZBOT_OK || cfg->task != ZBOT_TASK_WALKING_M) return ZBOT_E_INVALID;
  Params<T> P;
  params_from_cfg(*cfg, P);
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; ++e) {
    EnvState<T> es;
    env_state_unpack(state + (size_t)e * ZBOT_STATE_WORDS, es);
    StepOut<T> out;
    T rs[MAX_TERMS + 4];      // + is_terminated sum, base_height / feet_close / illegal_contact counts (statistics words 22..25)
    for (int i = 0; i < MAX_TERMS + 4; ++i) rs[i] = T(0);
    MExport<T> ex;
    ArrayScratch<T> scr;
    if (pt && pt->heights) {
      float* o = pt->env_origins + (size_t)e * 4;
      const TerrainGround<T> ground{pt->heights, pt->nx, pt->ny, T(pt->x0), T(pt->y0), T(1.0f / pt->cell), T(o[0]), T(o[1]), T(o[2])};
      MTerrainCtx<T> tc{{T(o[0]), T(o[1]), T(o[2])}, pt->tile_origins, pt->rows, pt->cols, T(pt->tile_size),
                        T((float)cfg->max_episode_length * cfg->sim_dt * (float)cfg->decimation), pt->curriculum};
      PhysOut<T> po;
      env_step_physics<ModelWalkM>(P, es, actions + (size_t)e * 6, po, scr, (StepExport<T>*)nullptr, ground);
      m_step_finish(P, es, actions + (size_t)e * 6, po, ep_len[e], rnd + (size_t)e * M_NUM_RAND, obs + (size_t)e * M_NUM_OBS, out,
                    rs, export_buf ? &ex : (MExport<T>*)nullptr, &tc);
      if (out.terminated || out.time_out) { o[0] = (float)tc.origin[0]; o[1] = (float)tc.origin[1]; o[2] = (float)tc.origin[2]; }
    } else {
      m_env_step(P, es, actions + (size_t)e * 6, ep_len[e], rnd + (size_t)e * M_NUM_RAND, obs + (size_t)e * M_NUM_OBS, out,
                 rs, export_buf ? &ex : (MExport<T>*)nullptr, scr);
    }
    env_state_pack(es, state + (size_t)e * ZBOT_STATE_WORDS);
    rew[e] = out.reward;
    term[e] = out.terminated ? 1 : 0;
    trunc[e] = out.time_out ? 1 : 0;
    if (reset_sums)      // manager task: [N][MAX_TERMS + 4]
      for (int i = 0; i < MAX_TERMS + 4; ++i) reset_sums[(size_t)e * (MAX_TERMS + 4) + i] = rs[i];
    if (export_buf) memcpy(export_buf + (size_t)e * ZBOT_M_EXPORT_WORDS, &ex, sizeof(ex));
  }
  return ZBOT_OK;
}

// dynamics only: sim [N][25] (root_pos3 quat4 lin3 ang3 q6 qd6), target [N][6]
// forces [N][7][3] (body 0 and 6 = applied foot forces, 1..5 = predictor), tau [N][6]
template <typename T>
static int port_substeps(const ZbotCfg* cfg, T* sim, const T* target, T* forces, T* tau, int n, int nsub, int model = 0,
                         const PortTerrain* pt = nullptr) {
  const bool snake = (model == 1);
  Params<T> P;
  params_from_cfg(*cfg, P);
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; ++e) {
    SimState<T> s;
    T* w = sim + (size_t)e * 25;
    for (int i = 0; i < 3; ++i) { s.p[i] = w[i]; s.v[i] = w[7 + i]; s.w[i] = w[10 + i]; }
    for (int i = 0; i < 4; ++i) s.Q[i] = w[3 + i];
    for (int i = 0; i < 6; ++i) { s.q[i] = w[13 + i]; s.qd[i] = w[19 + i]; }
    SubstepOut<T> so;
    ArrayScratch<T> scr;
    T midf[15];
    T tgt7[7];
    for (int i = 0; i < 6; ++i) tgt7[i] = target[(size_t)e * 6 + i];
    tgt7[6] = T(cfg->contact_mu);     // ModelWalkM reads its friction coefficient per env
    for (int k = 0; k < nsub; ++k) {
      if (snake) physics_substep<ModelSnake>(P, s, tgt7, so, scr, midf);
      else if (model == 2 && pt && pt->heights) {
        const float* o = pt->env_origins + (size_t)e * 4;
        const TerrainGround<T> ground{pt->heights, pt->nx, pt->ny, T(pt->x0), T(pt->y0), T(1.0f / pt->cell), T(o[0]), T(o[1]), T(o[2])};
        physics_substep<ModelWalkM>(P, s, tgt7, so, scr, midf, ground);
      }
      else if (model == 2) physics_substep<ModelWalkM>(P, s, tgt7, so, scr, midf);
      else if (model == 3) physics_substep_halves<ModelWalk>(P, s, tgt7, so, midf);   // two-halves elimination (csrc/zbot_halves.h)
      else if (model == 4) {    // both halves packed in the two FP32 lanes of one thread (csrc/zbot_h2.h); float only
        if constexpr (std::is_same<T, float>::value) physics_substep_h2_sim<ModelWalk>(P, s, tgt7, so, midf);
        else physics_substep_halves<ModelWalk>(P, s, tgt7, so, midf);
      }
      else physics_substep<ModelWalk>(P, s, tgt7, so, scr, midf);
    }
    for (int i = 0; i < 3; ++i) { w[i] = s.p[i]; w[7 + i] = s.v[i]; w[10 + i] = s.w[i]; }
    for (int i = 0; i < 4; ++i) w[3 + i] = s.Q[i];
    for (int i = 0; i < 6; ++i) { w[13 + i] = s.q[i]; w[19 + i] = s.qd[i]; }
    T* f = forces + (size_t)e * 21;
    for (int i = 0; i < 3; ++i) { f[i] = so.foot_force[0][i]; f[18 + i] = so.foot_force[1][i]; }
    for (int b = 0; b < 5; ++b)
      for (int i = 0; i < 3; ++i) f[3 * (b + 1) + i] = midf[3 * b + i];
    for (int i = 0; i < 6; ++i) tau[(size_t)e * 6 + i] = so.applied_torque[i];
  }
  return ZBOT_OK;
}

// the same substeps with TWO envs per call chain (T = F2, the packed GPU kernel's instantiation): envs (2i, 2i+1)
static int port_substeps_pair(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub,
                              bool snake) {
  Params<float> P;
  params_from_cfg(*cfg, P);
  if (n & 1) return ZBOT_E_INVALID;
#pragma omp parallel for schedule(static) num_threads(port_threads())
  for (int e = 0; e < n; e += 2) {
    SimState<F2> s;
    float *w0 = sim + (size_t)e * 25, *w1 = w0 + 25;
    for (int i = 0; i < 3; ++i) { s.p[i] = F2(w0[i], w1[i]); s.v[i] = F2(w0[7 + i], w1[7 + i]); s.w[i] = F2(w0[10 + i], w1[10 + i]); }
    for (int i = 0; i < 4; ++i) s.Q[i] = F2(w0[3 + i], w1[3 + i]);
    for (int i = 0; i < 6; ++i) { s.q[i] = F2(w0[13 + i], w1[13 + i]); s.qd[i] = F2(w0[19 + i], w1[19 + i]); }
    F2 tgt[6];
    for (int i = 0; i < 6; ++i) tgt[i] = F2(target[(size_t)e * 6 + i], target[(size_t)(e + 1) * 6 + i]);
    SubstepOut<F2> so;
    ArrayScratch<F2> scr;
    F2 midf[15];
    for (int k = 0; k < nsub; ++k) {
      if (snake) physics_substep<ModelSnake>(P, s, tgt, so, scr, midf);
      else physics_substep<ModelWalk>(P, s, tgt, so, scr, midf);
    }
    for (int l = 0; l < 2; ++l) {
      float* w = l ? w1 : w0;
      auto L = [l](F2 v) { return l ? v.y : v.x; };
      for (int i = 0; i < 3; ++i) { w[i] = L(s.p[i]); w[7 + i] = L(s.v[i]); w[10 + i] = L(s.w[i]); }
      for (int i = 0; i < 4; ++i) w[3 + i] = L(s.Q[i]);
      for (int i = 0; i < 6; ++i) { w[13 + i] = L(s.q[i]); w[19 + i] = L(s.qd[i]); }
      float* f = forces + (size_t)(e + l) * 21;
      for (int i = 0; i < 3; ++i) { f[i] = L(so.foot_force[0][i]); f[18 + i] = L(so.foot_force[1][i]); }
      for (int b = 0; b < 5; ++b)
        for (int i = 0; i < 3; ++i) f[3 * (b + 1) + i] = L(midf[3 * b + i]);
      for (int i = 0; i < 6; ++i) tau[(size_t)(e + l) * 6 + i] = L(so.applied_torque[i]);
    }
  }
  return ZBOT_OK;
}

template <typename T>
static int port_link_view(const T* sim, T* pos, T* quat, T* vel, int n) {
  for (int e = 0; e < n; ++e) {
    SimState<T> s;
    const T* w = sim + (size_t)e * 25;
    for (int i = 0; i < 3; ++i) { s.p[i] = w[i]; s.v[i] = w[7 + i]; s.w[i] = w[10 + i]; }
    for (int i = 0; i < 4; ++i) s.Q[i] = w[3 + i];
    for (int i = 0; i < 6; ++i) { s.q[i] = w[13 + i]; s.qd[i] = w[19 + i]; }
    all_link_kinematics(s, pos + (size_t)e * 36, quat + (size_t)e * 48, vel + (size_t)e * 36);
  }
  return 0;
}

extern "C" {
// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must use all host threads
int zbot_port_set_threads(int n) { if (n > 0) { g_threads = n; omp_set_num_threads(n); } return port_threads(); }
int zbot_port_default_cfg(ZbotCfg* c, int n) { cfg_defaults(*c, n); return 0; }
int zbot_port_export_words_f32(void) { return (int)(sizeof(StepExport<float>) / sizeof(float)); }
int zbot_port_export_words_f64(void) { return (int)(sizeof(StepExport<double>) / sizeof(double)); }
int zbot_port_state_word(const char* f) { return find_word(kStateFields, (int)(sizeof(kStateFields) / sizeof(kStateFields[0])), f); }

int zbot_port_step_f32(const ZbotCfg* cfg, float* state, int64_t* ep_len, const float* actions, float* obs,
                       float* rew, uint8_t* term, uint8_t* trunc, float* reset_sums, float* export_buf, int n) {
  return port_step<float>(cfg, state, ep_len, actions, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_step_f64(const ZbotCfg* cfg, double* state, int64_t* ep_len, const double* actions, double* obs,
                       double* rew, uint8_t* term, uint8_t* trunc, double* reset_sums, double* export_buf, int n) {
  return port_step<double>(cfg, state, ep_len, actions, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_v4_step_f32(const ZbotCfg* cfg, float* state, int64_t* ep_len, const float* actions, const float* rnd,
                          float* obs, float* rew, uint8_t* term, uint8_t* trunc, float* reset_sums, float* export_buf, int n) {
  return port_v4_step<float>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_v4_step_f64(const ZbotCfg* cfg, double* state, int64_t* ep_len, const double* actions, const double* rnd,
                          double* obs, double* rew, uint8_t* term, uint8_t* trunc, double* reset_sums, double* export_buf, int n) {
  return port_v4_step<double>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_snake_export_words(void) { return (int)(sizeof(SnakeExport<float>) / sizeof(float)); }
int zbot_port_snake_step_f32(const ZbotCfg* cfg, float* state, int64_t* ep_len, const float* actions, float* obs,
                             float* rew, uint8_t* term, uint8_t* trunc, float* reset_sums, float* export_buf, int n) {
  return port_snake_step<float>(cfg, state, ep_len, actions, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_snake_step_f64(const ZbotCfg* cfg, double* state, int64_t* ep_len, const double* actions, double* obs,
                             double* rew, uint8_t* term, uint8_t* trunc, double* reset_sums, double* export_buf, int n) {
  return port_snake_step<double>(cfg, state, ep_len, actions, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_substeps_pair_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub, int snake) {
  return port_substeps_pair(cfg, sim, target, forces, tau, n, nsub, snake != 0);
}
int zbot_port_substeps_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub);
}
int zbot_port_substeps_f64(const ZbotCfg* cfg, double* sim, const double* target, double* forces, double* tau, int n, int nsub) {
  return port_substeps<double>(cfg, sim, target, forces, tau, n, nsub);
}
int zbot_port_substeps_halves_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub, 3);
}
int zbot_port_substeps_halves_f64(const ZbotCfg* cfg, double* sim, const double* target, double* forces, double* tau, int n, int nsub) {
  return port_substeps<double>(cfg, sim, target, forces, tau, n, nsub, 3);
}
int zbot_port_substeps_h2_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub, 4);
}
int zbot_port_substeps_snake_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub, 1);
}
int zbot_port_substeps_snake_f64(const ZbotCfg* cfg, double* sim, const double* target, double* forces, double* tau, int n, int nsub) {
  return port_substeps<double>(cfg, sim, target, forces, tau, n, nsub, 1);
}
int zbot_port_substeps_m_terrain_f64(const ZbotCfg* cfg, double* sim, const double* target, double* forces, double* tau, int n, int nsub,
                                     const PortTerrain* pt) {
  return port_substeps<double>(cfg, sim, target, forces, tau, n, nsub, 2, pt);
}
int zbot_port_substeps_m_terrain_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub,
                                     const PortTerrain* pt) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub, 2, pt);
}
int zbot_port_substeps_m_f32(const ZbotCfg* cfg, float* sim, const float* target, float* forces, float* tau, int n, int nsub) {
  return port_substeps<float>(cfg, sim, target, forces, tau, n, nsub, 2);
}
int zbot_port_substeps_m_f64(const ZbotCfg* cfg, double* sim, const double* target, double* forces, double* tau, int n, int nsub) {
  return port_substeps<double>(cfg, sim, target, forces, tau, n, nsub, 2);
}
int zbot_port_m_step_terrain_f32(const ZbotCfg* cfg, float* state, int64_t* ep_len, const float* actions, const float* rnd,
                                 float* obs, float* rew, uint8_t* term, uint8_t* trunc, float* reset_sums, float* export_buf, int n,
                                 const PortTerrain* pt) {
  return port_m_step<float>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n, pt);
}
int zbot_port_m_step_terrain_f64(const ZbotCfg* cfg, double* state, int64_t* ep_len, const double* actions, const double* rnd,
                                 double* obs, double* rew, uint8_t* term, uint8_t* trunc, double* reset_sums, double* export_buf, int n,
                                 const PortTerrain* pt) {
  return port_m_step<double>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n, pt);
}
int zbot_port_m_step_f32(const ZbotCfg* cfg, float* state, int64_t* ep_len, const float* actions, const float* rnd,
                         float* obs, float* rew, uint8_t* term, uint8_t* trunc, float* reset_sums, float* export_buf, int n) {
  return port_m_step<float>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_m_step_f64(const ZbotCfg* cfg, double* state, int64_t* ep_len, const double* actions, const double* rnd,
                         double* obs, double* rew, uint8_t* term, uint8_t* trunc, double* reset_sums, double* export_buf, int n) {
  return port_m_step<double>(cfg, state, ep_len, actions, rnd, obs, rew, term, trunc, reset_sums, export_buf, n);
}
int zbot_port_link_view_f64(const double* sim, double* pos, double* quat, double* vel, int n) {
  return port_link_view<double>(sim, pos, quat, vel, n);
}
int zbot_port_link_view_f32(const float* sim, float* pos, float* quat, float* vel, int n) {
  return port_link_view<float>(sim, pos, quat, vel, n);
}
}
