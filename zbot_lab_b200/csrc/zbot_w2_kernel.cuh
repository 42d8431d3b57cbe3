// zbot_w2_kernel.cuh -- the fused zbot-6b-walking-v2 control step with TWO WARPS per 32 environments.
// Included by zbot_kernels.cu (inside its anonymous namespace, after the helpers it uses).
//
// A CTA is one warp PAIR = 64 threads = 32 environments.  Warp 0 ("side A") owns foot_0, bodies 0..2, joints 0..2, the
// floating root and the whole MDP phase; warp 1 ("side B") owns foot_1, bodies 4..6, joints 3..5 (csrc/zbot_halves.h).
// Lane l of both warps works on env e0 + l.  Per physics substep the two warps meet three times through shared memory and
// named barriers (bar.sync / bar.arrive with 64 threads):
//     FRAME  A -> B   pose / twist of body 3 after A's kinematics (14 words)   -- B's kinematics start from it
//     ROOT   A <-> B  each side's share of body 3's articulated inertia + bias force (27 words each way); both solve
//     (next FRAME)    also orders the re-use of the exchange buffers
// and once per control step at the start (A hands B its joints / targets / foot-1 sensor state) and at the end (B hands
// back its joints, foot-1 force history, timers).  All global memory traffic (state quads, actions, outputs) is A's, with
// exactly the access pattern of the one-thread-per-env kernel; B never touches global memory.
//
// Why two warps and not two lanes: the kinematics chain is serial (B's starts where A's ends).  In a lane pair both lanes
// would execute both halves of it; as two warps, A simply runs ahead into its elimination sweep while B does its
// kinematics -- no instruction is issued twice, each warp carries half the state and half the dependent chain, and the
// scheduler sees twice as many independent instruction streams at the same register budget.
#pragma once

constexpr int W2_RAW = HALF_SCR_WORDS;            // side A row: 6 raw actions, parked across the physics phase
constexpr int W2_PD = W2_RAW + 6;                 // side A row: 6 p_delta words (only needed again for the state store)
constexpr int W2_STRIDE = HALF_SCR_WORDS + 12;    // 63 words per thread (odd: conflict-free rows)
// exchange buffer of the pair: [slot][32 lanes]
constexpr int X_FRAME = 0, X_ROOTA = X_FRAME + FRAME3_WORDS, X_ROOTB = X_ROOTA + ROOT_SHARE_WORDS,
              X_WORDS = X_ROOTB + ROOT_SHARE_WORDS;   // 68
// Test hook (zbot_step_export): the SAME kernel instantiation writes, when `xp.hist1` is set, the contact forces of every
// substep, the end-of-physics articulation state and the sensor timers to global memory (uniform run-time branches around
// plain stores: no register is held for it); the host side of zbot_step_export turns them into the reference-layout view.
// offsets (floats) inside one [12][3] history slot of the merged bodies a side senses at k = 1, 2, 3 (-1: none); sensor body
// order b1 a2 b2 a3 b3 b4 a5 b5 a6 foot_0 foot_1 base, a merged body reports on its "a" half (kMidSensorIdx)
__device__ __constant__ int kW2MidOff[2][3] = {{1 * 3, 3 * 3, -1}, {8 * 3, 6 * 3, 11 * 3}};
// words side A leaves in side B's row before START, and side B leaves there before END
constexpr int W2_START_WORDS = 14, W2_END_WORDS = 19;
constexpr int BAR_START = 1, BAR_FRAME = 2, BAR_ROOT = 3, BAR_END = 4;

__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

struct W2Scratch {
  float* base;   // this thread's row
  __device__ __forceinline__ float& operator()(int j, int slot) { return base[j * SCR_PER_JOINT + slot]; }
};

__device__ __forceinline__ void w2_put_share(float* x, const RootShare<float>& r) {
#pragma unroll
  for (int i = 0; i < 6; ++i) { x[i * 32] = r.IA.I[i]; x[(15 + i) * 32] = r.IA.M[i]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) x[(6 + i) * 32] = r.IA.H[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { x[(21 + i) * 32] = r.pt[i]; x[(24 + i) * 32] = r.pb[i]; }
}
__device__ __forceinline__ void w2_get_share(const float* x, RootShare<float>& r) {
#pragma unroll
  for (int i = 0; i < 6; ++i) { r.IA.I[i] = x[i * 32]; r.IA.M[i] = x[(15 + i) * 32]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) r.IA.H[i] = x[(6 + i) * 32];
#pragma unroll
  for (int i = 0; i < 3; ++i) { r.pt[i] = x[(21 + i) * 32]; r.pb[i] = x[(24 + i) * 32]; }
}

// rows of ROW floats held by the side-A threads (tid < 32) -> the CTA's rows, written contiguously by all 64 threads
template <int ROW>
__device__ __forceinline__ void w2_store_rows(float* __restrict__ dst, const float* row, int n_end, int e0, float* smem) {
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < ROW; ++i) smem[threadIdx.x * ROW + i] = row[i];   // ROW odd -> conflict-free
  }
  __syncthreads();
  const int valid = min(32, n_end - e0);
  const int total = valid * ROW;
  float* base = dst + (size_t)e0 * ROW;
  if ((((size_t)e0 * ROW) & 3) == 0 && ((uintptr_t)dst & 15) == 0) {
    const int nv = total >> 2;
    for (int i = threadIdx.x; i < nv; i += 64) reinterpret_cast<float4*>(base)[i] = reinterpret_cast<const float4*>(smem)[i];
    for (int i = (nv << 2) + threadIdx.x; i < total; i += 64) base[i] = smem[i];
  } else {
    for (int i = threadIdx.x; i < total; i += 64) base[i] = smem[i];
  }
}

__device__ __forceinline__ void
zbot_step_w2_body(const Params<float>& P, const DefaultPose& dp, float4* __restrict__ state, int64_t* __restrict__ ep_len_buf,
                  const float* __restrict__ actions, float* __restrict__ obs, float* __restrict__ rew,
                  uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, int n, int e_begin, int e_end,
                  StatsCtx sc, ExportPtrs xp) {
  extern __shared__ float smem[];   // 64 rows of W2_STRIDE words, then the pair's exchange buffer
  pdl_wait();
  if (sc.pdl_early) pdl_trigger();
  const int side = threadIdx.x >> 5, lane = threadIdx.x & 31;     // warp-uniform side
  const int e0 = e_begin + blockIdx.x * 32;
  const int e = e0 + lane;
  const bool live = e < e_end;
  const int el = live ? e : (e_end - 1);        // a dead lane shadows the last env (it must keep the barriers company); never stored
  float* row = smem + threadIdx.x * W2_STRIDE;
  float* xch = smem + 64 * W2_STRIDE + lane;    // slot s of this lane: xch[s * 32]
  W2Scratch scr{row};
  const float dt = P.dt;
  const float mu = P.c_mu;
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  const bool kPacked = sc.packed_rows != 0;
  float* ex_hist = (live && xp.hist1) ? xp.hist1 + (size_t)e * (5 * 12 * 3) : nullptr;   // export hook (null in production)
  float obs_row[ZBOT_HOST_ROW_WORDS];
#pragma unroll
  for (int i = 0; i < ZBOT_HOST_ROW_WORDS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;

  // ------------------------------------------------------------------------------------------------------------------
  // prologue.  Side A loads the early quads and the actions, turns the actions into joint targets and hands side B its
  // joints / targets / foot-1 sensor state (START); from then on both warps run ONE copy of the substep loop below
  // (`side` is a run-time, warp-uniform value: the instruction footprint of the loop is that of half a robot).
  // ------------------------------------------------------------------------------------------------------------------
  float* row_b = row + 32 * W2_STRIDE;           // side A: the partner thread's row
  HalfState<float> h;
  float target[3];
  ContactTimers<float> tm;
  float speed_limit = 0.f, carry_fz0 = 0.f, carry_fz1 = 0.f, carry_mid = 0.f;
  if (side == 0) {
    EnvState<float> es;                          // prologue only: nothing of it stays live across the substep loop
    {
      float w[4 * EARLY_QUADS];
      load_words<EARLY_QUADS>(state, n, el, w);
      env_early_unpack(w, es);
    }
    float tgt6[7];
    {
      const float2* a2p = reinterpret_cast<const float2*>(actions + (size_t)el * 6);
      const float2 a0 = __ldg(a2p), a1 = __ldg(a2p + 1), a2v = __ldg(a2p + 2);
      const float raw[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
#pragma unroll
      for (int k = 0; k < 6; ++k) row[W2_RAW + k] = raw[k];     // read ONCE (may be pinned host memory), parked for the MDP phase
      float new_actions[6];
      mdp_pre_physics<ModelWalk>(P, raw, es.mdp, new_actions, tgt6);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) { row_b[t] = es.sim.q[3 + t]; row_b[3 + t] = es.sim.qd[3 + t]; row_b[6 + t] = tgt6[3 + t]; }
    row_b[9] = es.timers[1].cur_air; row_b[10] = es.timers[1].cur_contact; row_b[11] = es.timers[1].last_air;
    row_b[12] = es.timers[1].last_contact;
    named_bar_arrive(BAR_START);
#pragma unroll
    for (int k = 0; k < 6; ++k) row[W2_PD + k] = es.mdp.p_delta[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) { h.p[i] = es.sim.p[i]; h.v[i] = es.sim.v[i]; h.w[i] = es.sim.w[i]; h.q[i] = es.sim.q[i]; h.qd[i] = es.sim.qd[i]; target[i] = tgt6[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) h.Q[i] = es.sim.Q[i];
    tm = es.timers[0];
    speed_limit = es.mdp.speed_limit;
    carry_fz0 = es.carry_feet_fz[0]; carry_fz1 = es.carry_feet_fz[1]; carry_mid = es.carry_mid_max;
  } else {
    named_bar_sync(BAR_START);
#pragma unroll
    for (int t = 0; t < 3; ++t) { h.q[t] = row[t]; h.qd[t] = row[3 + t]; target[t] = row[6 + t]; }
    tm.cur_air = row[9]; tm.cur_contact = row[10]; tm.last_air = row[11]; tm.last_contact = row[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) { h.p[i] = 0.f; h.v[i] = 0.f; h.w[i] = 0.f; }
    h.Q[0] = 1.f; h.Q[1] = 0.f; h.Q[2] = 0.f; h.Q[3] = 0.f;
  }
  float fz[4] = {0.f, 0.f, 0.f, 0.f};
  float mid2_all = 0.f, mid2_last = 0.f, tau[3] = {0.f, 0.f, 0.f};
  float* x_mine = xch + (X_ROOTA + side * ROOT_SHARE_WORDS) * 32;
  float* x_other = xch + (X_ROOTA + (1 - side) * ROOT_SHARE_WORDS) * 32;
#pragma unroll 1
  for (int sub = 0; sub < P.decimation; ++sub) {
    half_pd(P, h, target, scr, tau);
    // ---- kinematics: A from the root to body 3, THEN B from body 3 to foot_1 (FRAME) ----
    BodyKin<float> k;
    float pz;
    if (side) {
      named_bar_sync(BAR_FRAME);
#pragma unroll
      for (int i = 0; i < 4; ++i) k.Q[i] = xch[(X_FRAME + i) * 32];
#pragma unroll
      for (int i = 0; i < 3; ++i) { k.r[i] = xch[(X_FRAME + 4 + i) * 32]; k.w[i] = xch[(X_FRAME + 7 + i) * 32]; k.vO[i] = xch[(X_FRAME + 10 + i) * 32]; }
      pz = xch[(X_FRAME + 13) * 32];
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) { k.r[i] = 0.f; k.w[i] = h.w[i]; k.vO[i] = h.v[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) k.Q[i] = h.Q[i];
      pz = h.p[2];
    }
    half_fk(side, h, k, scr);
    if (!side) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xch[(X_FRAME + i) * 32] = k.Q[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) { xch[(X_FRAME + 4 + i) * 32] = k.r[i]; xch[(X_FRAME + 7 + i) * 32] = k.w[i]; xch[(X_FRAME + 10 + i) * 32] = k.vO[i]; }
      xch[(X_FRAME + 13) * 32] = pz;
      named_bar_sync(BAR_FRAME);       // a full sync (not an arrive): it also orders the re-use of the exchange buffers
      // the elimination sweep starts at this side's FOOT: the root for side A, where the kinematics arrived for side B
#pragma unroll
      for (int i = 0; i < 3; ++i) { k.r[i] = 0.f; k.w[i] = h.w[i]; k.vO[i] = h.v[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) k.Q[i] = h.Q[i];
    }
    // ---- elimination from the foot to body 3, exchange of the two shares of body 3, solve, sweep back out ----
    RootShare<float> mine, other;
    float w3[3], vO3[3], mid2;
    ContactAgg<float> agg;
    const int slot = P.decimation - 1 - sub;   // ContactSensor history slot of this substep, newest first
    half_backward<ModelWalk>(P, side, mu, pz, k, scr, mine, w3, vO3, agg, mid2, ex_hist ? ex_hist + slot * 36 : (float*)nullptr,
                             kW2MidOff[side]);
    w2_put_share(x_mine, mine);
    named_bar_sync(BAR_ROOT);
    w2_get_share(x_other, other);
    float At[3], Ab[3];
    half_root_solve(mine, other, At, Ab);
    half_forward(P, side, scr, w3, vO3, At, Ab);
    float ff[3];
    contact_agg_force(agg, dt, At, Ab, ff);
    half_integrate(P, side, h, scr, At, Ab);
    // ---- ContactSensor.update of this side's foot ----
    const float nrm = sqrtf(ff[0] * ff[0] + ff[1] * ff[1] + ff[2] * ff[2]);
    contact_timers_update(tm, nrm > 1.0f, dt);
#pragma unroll
    for (int q = 0; q < 4; ++q) fz[q] = (slot == q) ? ff[2] : fz[q];
    mid2_all = fmaxf(mid2_all, mid2);
    mid2_last = mid2;
    if (ex_hist) {
      float* f = ex_hist + slot * 36 + (side ? kFoot1Sensor : kFoot0Sensor) * 3;
      f[0] = ff[0]; f[1] = ff[1]; f[2] = ff[2];
    }
  }
  if (side == 1) {
    // END: hand the results back through this thread's own row (its scratch is dead now)
#pragma unroll
    for (int t = 0; t < 3; ++t) { row[t] = h.q[t]; row[3 + t] = h.qd[t]; row[6 + t] = tau[t]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) row[9 + q] = fz[q];
    row[13] = mid2_all; row[14] = mid2_last;
    row[15] = tm.cur_air; row[16] = tm.cur_contact; row[17] = tm.last_air; row[18] = tm.last_contact;
    named_bar_arrive(BAR_END);
  } else {
    StepOut<float> out;
    float rs[MAX_TERMS];
#pragma unroll
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = 0.f;
    named_bar_sync(BAR_END);
    // ---- assemble the end-of-physics state and what the MDP phase reads (env_step_physics' outputs) ----
    EnvState<float> es;
    PhysOut<float> po;
#pragma unroll
    for (int i = 0; i < 3; ++i) { es.sim.p[i] = h.p[i]; es.sim.v[i] = h.v[i]; es.sim.w[i] = h.w[i]; es.sim.q[i] = h.q[i]; es.sim.qd[i] = h.qd[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) es.sim.Q[i] = h.Q[i];
#pragma unroll
    for (int t = 0; t < 3; ++t) { es.sim.q[3 + t] = row_b[t]; es.sim.qd[3 + t] = row_b[3 + t]; po.applied_torque[t] = tau[t]; po.applied_torque[3 + t] = row_b[6 + t]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { po.fz[q][0] = fz[q]; po.fz[q][1] = row_b[9 + q]; }
    po.fz[4][0] = carry_fz0; po.fz[4][1] = carry_fz1;
    const float mid2_b_all = row_b[13], mid2_b_last = row_b[14];
    po.mid2 = fmaxf(carry_mid * carry_mid, fmaxf(mid2_all, mid2_b_all));
    es.timers[0] = tm;
    es.timers[1].cur_air = row_b[15]; es.timers[1].cur_contact = row_b[16]; es.timers[1].last_air = row_b[17]; es.timers[1].last_contact = row_b[18];
    es.carry_feet_fz[0] = fz[0]; es.carry_feet_fz[1] = po.fz[0][1];          // slot 0 = the last substep
    es.carry_mid_max = sqrtf(fmaxf(mid2_last, mid2_b_last));
    es.mdp.speed_limit = speed_limit;
#pragma unroll
    for (int k = 0; k < 6; ++k) es.mdp.p_delta[k] = row[W2_PD + k];
    if (ex_hist) {
      // end-of-physics articulation state (before any reset) and sensor timers; the 13 root words ride in the first
      // words of this env's body_link_pos_w1 row until zbot_export_view_kernel expands them to the 12-link view
      float* r13 = xp.pos1 + (size_t)e * 36;
#pragma unroll
      for (int i = 0; i < 3; ++i) { r13[i] = es.sim.p[i]; r13[7 + i] = es.sim.v[i]; r13[10 + i] = es.sim.w[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) r13[3 + i] = es.sim.Q[i];
#pragma unroll
      for (int k = 0; k < 6; ++k) { xp.q1[(size_t)e * 6 + k] = es.sim.q[k]; xp.qd1[(size_t)e * 6 + k] = es.sim.qd[k]; xp.tau1[(size_t)e * 6 + k] = po.applied_torque[k]; }
      xp.last_air1[(size_t)e * 12 + kFoot0Sensor] = es.timers[0].last_air;
      xp.last_air1[(size_t)e * 12 + kFoot1Sensor] = es.timers[1].last_air;
      xp.cur_contact1[(size_t)e * 12 + kFoot0Sensor] = es.timers[0].cur_contact;
      xp.cur_contact1[(size_t)e * 12 + kFoot1Sensor] = es.timers[1].cur_contact;
    }
    // ---- phase C, exactly as in the one-thread kernel: late quads, S0 again (L2 hit), raw actions, the MDP ----
    {
      float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
      load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, el, w);
      env_late_unpack(w, es);
    }
    SimState<float> s0;
    {
      float w[4 * SIM_QUADS];
      load_words<SIM_QUADS>(state, n, el, w);
      sim_state_unpack(w, s0);
    }
    float raw[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) raw[k] = row[W2_RAW + k];
    int64_t ep = ep_len_buf[el];
    env_step_finish(P, es, s0, raw, po, ep, dp.feet_pos, dp.base_quat, out, rs, (StepExport<float>*)nullptr);
    if (live) {
      float w[ZBOT_STATE_WORDS];
      env_state_pack(es, w);
      store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
      ep_len_buf[e] = ep;
      if (kPacked) {
        obs_row[ZBOT_NUM_OBS] = out.reward;
        obs_row[ZBOT_NUM_OBS + 1] = __uint_as_float((out.terminated ? 1u : 0u) | (out.time_out ? 0x100u : 0u));
      } else {
        rew[e] = out.reward;
        terminated[e] = out.terminated ? 1 : 0;
        truncated[e] = out.time_out ? 1 : 0;
      }
#pragma unroll
      for (int i = 0; i < ZBOT_NUM_OBS; ++i) obs_row[i] = out.obs[i];
      obs_add_noise<ZBOT_NUM_OBS>(P, sc, e, obs_row);
      did_reset = out.terminated || out.time_out;
      if (did_reset) {
#pragma unroll
        for (int i = 0; i < MAX_TERMS; ++i) stat[i] = rs[i];
        stat[S_NUM_RESET] = 1.f;
        stat[S_NUM_TERM_RESET] = out.terminated ? 1.f : 0.f;
        stat[S_NUM_TO_RESET] = out.time_out ? 1.f : 0.f;
      }
      stat[S_REW_SUM] = out.reward;
      stat[S_NUM_TERM] = out.terminated ? 1.f : 0.f;
      stat[S_NUM_TRUNC] = out.time_out ? 1.f : 0.f;
    }
  }
  __syncthreads();   // both warps are done with the rows and the exchange buffer before the output rows are staged
  if (kPacked) w2_store_rows<ZBOT_HOST_ROW_WORDS>(obs, obs_row, e_end, e0, smem);   // block-uniform branch
  else w2_store_rows<ZBOT_NUM_OBS>(obs, obs_row, e_end, e0, smem);
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

// register budget = resident 64-thread CTAs per SM the kernel is compiled for: 3 (uncapped) is the product default -- the
// library launches this kernel while an SM holds at most two pairs; 6 (168 registers), 8 (128: 4 warps per scheduler, the
// MDP phase spills) and 10 are the tuning variants of the occupancy experiments (ZBOT_W2_CTAS, profiles/r2_notes.md)
template <int kMinBlocks>
__global__ void __launch_bounds__(64, kMinBlocks) zbot_step_w2_kernel(ZB_STEP_ARGS) {
  zbot_step_w2_body(ZB_STEP_CALL);
}
constexpr size_t kW2Smem = (size_t)(64 * W2_STRIDE + X_WORDS * 32) * sizeof(float);                      // 24.8 KB

// ---- host-side helpers of the export hook (zbot_step_export on a w2 handle) ----
// before the step: zero the history / timer tensors and fill history slot 4 (the carry-over of the previous step's last substep)
__global__ void zbot_w2_export_pre_kernel(const float4* __restrict__ state, ExportPtrs xp, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* h = xp.hist1 + (size_t)e * 180;
  for (int i = 0; i < 180; ++i) h[i] = 0.f;
  for (int b = 0; b < 12; ++b) { xp.last_air1[(size_t)e * 12 + b] = 0.f; xp.cur_contact1[(size_t)e * 12 + b] = 0.f; }
  const float4 c = state[(size_t)(W_CARRY_FZ / 4) * n + e];       // words 32..35: carry_feet_fz[2], carry_mid_max, pad
  h[(4 * 12 + kFoot0Sensor) * 3 + 2] = c.x;
  h[(4 * 12 + kFoot1Sensor) * 3 + 2] = c.y;
  h[(4 * 12 + kMidSensorIdx[0]) * 3 + 0] = c.z;
}
// after the step: expand the exported end-of-physics state (13 root words parked in the pos1 row + q1 / qd1) to the 12-link view
__global__ void zbot_w2_export_view_kernel(ExportPtrs xp, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  SimState<float> s;
  const float* r13 = xp.pos1 + (size_t)e * 36;
  for (int i = 0; i < 3; ++i) { s.p[i] = r13[i]; s.v[i] = r13[7 + i]; s.w[i] = r13[10 + i]; }
  for (int i = 0; i < 4; ++i) s.Q[i] = r13[3 + i];
  for (int k = 0; k < 6; ++k) { s.q[k] = xp.q1[(size_t)e * 6 + k]; s.qd[k] = xp.qd1[(size_t)e * 6 + k]; }
  float p[36], q[48], v[36];
  all_link_kinematics(s, p, q, v);
  for (int i = 0; i < 36; ++i) { xp.pos1[(size_t)e * 36 + i] = p[i]; xp.vel1[(size_t)e * 36 + i] = v[i]; }
  for (int i = 0; i < 48; ++i) xp.quat1[(size_t)e * 48 + i] = q[i];
}
