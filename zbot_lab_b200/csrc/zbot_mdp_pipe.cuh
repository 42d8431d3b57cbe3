// zbot_mdp_pipe.cuh -- the MDP-only step (BASELINE configs[0] shape; SURVEY section 8 rows a3, a7-a10 on caller-supplied
// articulation / contact tensors) as a PERSISTENT, TMA-fed kernel.  Same arithmetic as `zbot_mdp_kernel<true>` (the
// per-env device functions of zbot_core.h are shared); what changes is how the bytes travel.
//
// This is the one HBM-bound kernel of the path (1570 algorithmic bytes per env-step, 46 % of them the 5 x 12 x 3 contact-force
// history).  The one-shot kernel (one CTA per 128 envs, history tile by TMA, everything else by per-thread
// loads) ran 512 CTAs on 296 resident slots (1.73 waves), exposed the latency of ~70 scattered loads per thread at the head of
// every CTA and needed a second launch for the statistics.  Here:
//   * grid = one CTA per SM, 4 shared-memory stages of 52 KB (one 32-env tile each) and kShare warps per stage that take the
//     stage's tiles in turn (the arithmetic of one tile is a ~8000-clock dependent chain for a lone warp, ncu: 4 warps per SM
//     left the kernel latency-bound at 0.11 instructions per clock and scheduler; with 2-3 warps per stage one computes while
//     the next tile of its neighbour lands);
//   * EVERY input of a tile (history, the three [N][12][*] link tensors, joint state, timers, origins, actions, episode
//     counters and the 18 state rows) arrives by 1-D bulk async copies (`cp.async.bulk`, 29 per tile, issued by 29 lanes at
//     once) that complete on the warp's mbarrier: 4 x 52 KB in flight per SM, no load instruction waits on DRAM;
//   * a warp drains the stage into registers, hands the stage straight back to the copy engine for the NEXT tile of the stage
//     (which completes on the mbarrier of the warp whose turn is next) and only then computes -- the next tile's DRAM time
//     hides behind this tile's arithmetic;
//   * tiles are dealt warp-major (worker id = warp * gridDim + cta), so the surplus tiles of the last round spread over all
//     SMs (65536 envs: 13 or 14 tiles per SM instead of 12 or 16);
//   * the statistics are accumulated per lane over all tiles of the warp and leave as ONE partial row per CTA (148 rows
//     instead of 512 for the finalize pass).
// Measured on B200 (profiles/r2_notes.md section 8; CUDA-graph replay over six rotating input sets, i.e. inputs larger than the
// L2 and no flush): 262144 envs 91.9 us (0.68 of the measured copy peak) against 96.5-99 us for the one-shot kernel; 65536 envs
// 29.2 us against 28.4 us -- there the one-shot kernel's 410 KB of requests per SM issued at once beat this kernel's 4 x 52 KB,
// and a lone warp needs ~2.4 us to drain and reduce a stage.  Hence: default from 131072 envs (zbot_mdp_step), ZBOT_MDP_PIPE =
// 42 | 43 | 33 | 41 forces a <stages><warps per stage> shape, 0 the one-shot kernel; ZBOT_MDP_FUSE_STATS=1 runs the grid-level
// statistics pass in the last CTA to finish instead of a second launch (no gain: the separate launch overlaps through PDL).
// Ragged tail (n % 32 != 0, and then only when the live count is not a multiple of 4): the arrays whose tile span is not a
// multiple of 16 bytes are copied by the warp itself; everything else still goes by TMA.
#pragma once
// (included inside zbot_kernels.cu's anonymous namespace, after zbot_mdp_kernel)

constexpr int kPT = 32;                          // envs per tile = one warp
constexpr int kPS_HIST = 0;                      // 32 x 720
constexpr int kPS_POS = kPS_HIST + kPT * 720;    // 32 x 144
constexpr int kPS_QUAT = kPS_POS + kPT * 144;    // 32 x 192
constexpr int kPS_VEL = kPS_QUAT + kPT * 192;    // 32 x 144
constexpr int kPS_Q = kPS_VEL + kPT * 144;       // 32 x 24
constexpr int kPS_QD = kPS_Q + kPT * 24;
constexpr int kPS_TAU = kPS_QD + kPT * 24;
constexpr int kPS_ACT = kPS_TAU + kPT * 24;
constexpr int kPS_AIR = kPS_ACT + kPT * 24;      // 32 x 48
constexpr int kPS_ORG = kPS_AIR + kPT * 48;      // 32 x 12
constexpr int kPS_EP = kPS_ORG + kPT * 12;       // 32 x 8
constexpr int kPS_MST = kPS_EP + kPT * 8;        // 18 x 32 x 16
constexpr int kPS_BYTES = kPS_MST + (ZBOT_MDP_STATE_WORDS / 4) * kPT * 16;
constexpr int kPS_OBS = (kPT / 2) * ZBOT_NUM_OBS * 4;  // per-warp observation staging: half a tile at a time
// kStages stages, kShare warps per stage
constexpr size_t pipe_smem(int stages, int share) {
  return (size_t)stages * kPS_BYTES + (size_t)stages * share * (kPS_OBS + kStats * sizeof(float) + 8);
}
static_assert(kPS_BYTES % 128 == 0 && kPS_OBS % 16 == 0, "stage alignment");
static_assert(pipe_smem(4, 3) <= 232448 - 1024 && pipe_smem(3, 3) <= 232448 - 1024, "shared memory budget of one SM");

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kStages, int kShare>
__global__ void __launch_bounds__(kStages * kShare * 32, 1)
zbot_mdp_pipe_kernel(const __grid_constant__ Params<float> P, const __grid_constant__ DefaultPose dp, MdpIn in,
                     float4* __restrict__ mstate, int64_t* __restrict__ ep_len_buf, const float* __restrict__ actions,
                     float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated,
                     uint8_t* __restrict__ truncated, int n, StatsCtx sc) {
  extern __shared__ __align__(128) unsigned char mpsm[];
  constexpr int kWarps = kStages * kShare;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sidx = w % kStages, slot = w / kStages;       // stage of this warp, its turn among the stage's warps
  unsigned char* stage = mpsm + (size_t)sidx * kPS_BYTES;
  float* obs_stage = reinterpret_cast<float*>(mpsm + (size_t)kStages * kPS_BYTES + (size_t)w * kPS_OBS);
  float* stat_sm = reinterpret_cast<float*>(mpsm + (size_t)kStages * kPS_BYTES + (size_t)kWarps * kPS_OBS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat_sm + kWarps * kStats);   // [stage][slot]
  uint64_t* my_bar = bars + sidx * kShare + slot;
  const int ntiles = (n + kPT - 1) / kPT;
  const int nstages = gridDim.x * kStages;
  const int sg = sidx * gridDim.x + blockIdx.x;                 // stage-major id: the surplus tiles spread over all SMs
  constexpr int NQ = ZBOT_MDP_STATE_WORDS / 4;
  if (threadIdx.x < kWarps) mbar_init(bars + threadIdx.x, 1);
  __syncthreads();
  pdl_wait();

  // hand the warp's stage to the copy engine for `tile`
  auto issue = [&](int tile, uint64_t* bar) {
    const int e0 = tile * kPT;
    const int valid = min(kPT, n - e0);
    const bool all16 = (valid & 3) == 0;        // every span is a multiple of 16 bytes
    const uint32_t v = (uint32_t)valid;
    if (!all16) {      // ragged tail: the small rows by the warp itself (4-byte words, lane = env) -- BEFORE the arrive below, whose
                       // release makes them visible to the consuming warp together with the bulk copies
      if (lane < valid) {
        const int e = e0 + lane;
        float* sq = reinterpret_cast<float*>(stage + kPS_Q) + lane * 6;
        float* sqd = reinterpret_cast<float*>(stage + kPS_QD) + lane * 6;
        float* st = reinterpret_cast<float*>(stage + kPS_TAU) + lane * 6;
        float* sa = reinterpret_cast<float*>(stage + kPS_ACT) + lane * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          sq[k] = __ldg(in.q + (size_t)e * 6 + k); sqd[k] = __ldg(in.qd + (size_t)e * 6 + k);
          st[k] = __ldg(in.tau + (size_t)e * 6 + k); sa[k] = __ldg(actions + (size_t)e * 6 + k);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) reinterpret_cast<float*>(stage + kPS_ORG)[lane * 3 + k] = __ldg(in.origins + (size_t)e * 3 + k);
        reinterpret_cast<int64_t*>(stage + kPS_EP)[lane] = ep_len_buf[e];
      }
      __syncwarp();
    }
    if (lane == 0) {
      uint32_t tx = v * (720u + 144u + 192u + 144u + 48u + 16u * NQ);
      if (all16) tx += v * (24u * 4u + 12u + 8u);
      mbar_expect_tx(bar, tx);
    }
    __syncwarp();
    if (lane < NQ) {
      bulk_g2s(stage + kPS_MST + lane * (kPT * 16), mstate + (size_t)lane * n + e0, v * 16u, bar);
    } else {
      const void* src = nullptr;
      int off = 0;
      uint32_t row = 0;
      bool small = false;      // span = valid * row is a multiple of 16 only when valid % 4 == 0 (or % 2 for the 24 / 8 byte rows)
      switch (lane - NQ) {
        case 0: src = in.hist + (size_t)e0 * kHistRow; off = kPS_HIST; row = 720; break;
        case 1: src = in.pos + (size_t)e0 * 36; off = kPS_POS; row = 144; break;
        case 2: src = in.quat + (size_t)e0 * 48; off = kPS_QUAT; row = 192; break;
        case 3: src = in.vel + (size_t)e0 * 36; off = kPS_VEL; row = 144; break;
        case 4: src = in.last_air + (size_t)e0 * 12; off = kPS_AIR; row = 48; break;
        case 5: src = in.q + (size_t)e0 * 6; off = kPS_Q; row = 24; small = true; break;
        case 6: src = in.qd + (size_t)e0 * 6; off = kPS_QD; row = 24; small = true; break;
        case 7: src = in.tau + (size_t)e0 * 6; off = kPS_TAU; row = 24; small = true; break;
        case 8: src = actions + (size_t)e0 * 6; off = kPS_ACT; row = 24; small = true; break;
        case 9: src = in.origins + (size_t)e0 * 3; off = kPS_ORG; row = 12; small = true; break;
        case 10: src = ep_len_buf + e0; off = kPS_EP; row = 8; small = true; break;
        default: break;
      }
      if (src && (!small || all16)) bulk_g2s(stage + off, src, v * row, bar);
    }
  };

  float stat_acc[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat_acc[j] = 0.f;
  bool any_reset_acc = false;
  uint32_t phase = 0;
  // the k-th tile of the stage (tile index sg + k * nstages) is taken by the warp in slot k % kShare
  if (slot == 0 && sg < ntiles) issue(sg, my_bar);

  for (int k = slot;; k += kShare) {
    const int tile = sg + k * nstages;
    if (tile >= ntiles) break;
    const int e0 = tile * kPT;
    const int e = e0 + lane;
    const int valid = min(kPT, n - e0);
    const bool live = lane < valid;
    mbar_wait(my_bar, phase);
    phase ^= 1u;

    // ---- drain the stage into registers -----------------------------------------------------------------------------
    float wv[ZBOT_MDP_STATE_WORDS];
    float base_pos[3], base_quat[4], base_vel[3], feet_pos[2][3], feet_quat[2][4], feet_vel[2][3];
    float q[6], qd[6], raw[6];
    FreshInputs<float> f;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    int64_t ep = 0;
    if (live) {
      const float4* ms4 = reinterpret_cast<const float4*>(stage + kPS_MST);
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) {
        const float4 v4 = ms4[qq * kPT + lane];
        wv[4 * qq] = v4.x; wv[4 * qq + 1] = v4.y; wv[4 * qq + 2] = v4.z; wv[4 * qq + 3] = v4.w;
      }
      const float* p = reinterpret_cast<const float*>(stage + kPS_POS) + lane * 36;
      const float* qv = reinterpret_cast<const float*>(stage + kPS_QUAT) + lane * 48;
      const float* v = reinterpret_cast<const float*>(stage + kPS_VEL) + lane * 36;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        base_pos[i] = p[kBaseLink * 3 + i]; base_vel[i] = v[kBaseLink * 3 + i];
        feet_pos[0][i] = p[kFoot0Link * 3 + i]; feet_pos[1][i] = p[kFoot1Link * 3 + i];
        feet_vel[0][i] = v[kFoot0Link * 3 + i]; feet_vel[1][i] = v[kFoot1Link * 3 + i];
      }
      {
        const float4 b = *reinterpret_cast<const float4*>(qv + kBaseLink * 4);
        const float4 f0 = *reinterpret_cast<const float4*>(qv + kFoot0Link * 4);
        const float4 f1 = *reinterpret_cast<const float4*>(qv + kFoot1Link * 4);
        base_quat[0] = b.x; base_quat[1] = b.y; base_quat[2] = b.z; base_quat[3] = b.w;
        feet_quat[0][0] = f0.x; feet_quat[0][1] = f0.y; feet_quat[0][2] = f0.z; feet_quat[0][3] = f0.w;
        feet_quat[1][0] = f1.x; feet_quat[1][1] = f1.y; feet_quat[1][2] = f1.z; feet_quat[1][3] = f1.w;
      }
      const float* sq = reinterpret_cast<const float*>(stage + kPS_Q) + lane * 6;
      const float* sqd = reinterpret_cast<const float*>(stage + kPS_QD) + lane * 6;
      const float* stau = reinterpret_cast<const float*>(stage + kPS_TAU) + lane * 6;
      const float* sact = reinterpret_cast<const float*>(stage + kPS_ACT) + lane * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) { q[k] = sq[k]; qd[k] = sqd[k]; f.applied_torque[k] = stau[k]; raw[k] = sact[k]; }
      const float* so = reinterpret_cast<const float*>(stage + kPS_ORG) + lane * 3;
      ox = so[0]; oy = so[1]; oz = so[2];
      const float* sair = reinterpret_cast<const float*>(stage + kPS_AIR) + lane * 12;
      f.last_air_time[0] = sair[kFoot0Sensor];
      f.last_air_time[1] = sair[kFoot1Sensor];
      ep = reinterpret_cast<const int64_t*>(stage + kPS_EP)[lane] + 1;
      {
        // the lane's own 45 float4 of the history tile (conflict-free: row stride 45 float4 is odd)
        const float4* row4 = reinterpret_cast<const float4*>(stage + kPS_HIST) + lane * (kHistRow / 4);
        float fz0 = 0.f, fz1 = 0.f, mx2 = 0.f;
#pragma unroll
        for (int t = 0; t < ZBOT_HISTORY; ++t) {
          float h[36];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const float4 v4 = row4[t * 9 + i];
            h[4 * i] = v4.x; h[4 * i + 1] = v4.y; h[4 * i + 2] = v4.z; h[4 * i + 3] = v4.w;
          }
          const float a = h[kFoot0Sensor * 3 + 2], b = h[kFoot1Sensor * 3 + 2];
          fz0 = (t == 0) ? a : (fz0 + a);                      // (((h0+h1)+h2)+h3)+h4, newest first
          fz1 = (t == 0) ? b : (fz1 + b);
#pragma unroll
          for (int b12 = 0; b12 < 12; ++b12) {
            if (b12 == kFoot0Sensor || b12 == kFoot1Sensor) continue;
            const float x = h[b12 * 3], y = h[b12 * 3 + 1], z = h[b12 * 3 + 2];
            mx2 = fmaxf(mx2, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
          }
        }
        f.feet_force[0] = fz0 / 5.0f;                            // torch.mean = sum / count (:387-390)
        f.feet_force[1] = fz1 / 5.0f;
        f.undesired_force_max = sqrtf(mx2);                      // max_t |F| > 1.0  (:396-402)
      }
    }
    // ---- the stage is drained: give it back to the copy engine for this warp's next tile ------------------------------
    __syncwarp();
    if (tile + nstages < ntiles) {
      fence_proxy_async();
      issue(tile + nstages, bars + sidx * kShare + (slot + 1) % kShare);
    }

    // ---- the MDP of one env (identical to zbot_mdp_kernel<true>) -------------------------------------------------------
    float stat[kStatUsed];
#pragma unroll
    for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
    float obs_row[ZBOT_NUM_OBS];
#pragma unroll
    for (int i = 0; i < ZBOT_NUM_OBS; ++i) obs_row[i] = 0.f;
    bool did_reset = false;
    if (live) {
      MdpState<float> m;
      mdp_state_unpack(wv, M_PDELTA, M_ACT, M_FLAST, M_FDPL, M_FSL, M_HSUM, M_YSUM, M_FFSUM, M_SPEED, M_EPSUM, m);
      StaleCache<float> stale;
      stale_unpack(wv, stale);
      float new_actions[6], target[6];
      mdp_pre_physics<ModelWalk>(P, raw, m, new_actions, target);
#pragma unroll
      for (int j = 0; j < 2; ++j) { f.feet_vel_xy[j][0] = feet_vel[j][0]; f.feet_vel_xy[j][1] = feet_vel[j][1]; }
      f.origin_y = oy;
      bool term, tout;
      const float r = mdp_dones_rewards(P, stale, f, new_actions, m, ep, term, tout);
#pragma unroll
      for (int k = 0; k < 6; ++k) m.actions[k] = new_actions[k];
      did_reset = term || tout;
      if (did_reset) {
#pragma unroll
        for (int i = 0; i < MAX_TERMS; ++i) stat[i] = (i < P.num_terms) ? m.ep_sums[i] : 0.f;
        stat[S_NUM_RESET] = 1.f;
        stat[S_NUM_TERM_RESET] = term ? 1.f : 0.f;
        stat[S_NUM_TO_RESET] = tout ? 1.f : 0.f;
        // post-reset articulation view = default pose + env origin (SURVEY C-5)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float o = (i == 0) ? ox : (i == 1) ? oy : oz;
          base_pos[i] = dp.base_pos[i] + o; base_vel[i] = 0.f;
          feet_pos[0][i] = dp.feet_pos[0][i] + o; feet_pos[1][i] = dp.feet_pos[1][i] + o;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { base_quat[i] = dp.base_quat[i]; feet_quat[0][i] = dp.feet_quat[0][i]; feet_quat[1][i] = dp.feet_quat[1][i]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { q[k] = default_joint_pos<float>(k); qd[k] = 0.f; }
        mdp_reset(m, feet_pos, P.num_terms);
        ep = 0;
      }
      stat[S_REW_SUM] = r;
      stat[S_NUM_TERM] = term ? 1.f : 0.f;
      stat[S_NUM_TRUNC] = tout ? 1.f : 0.f;
      ep_len_buf[e] = ep;
      rew[e] = r;
      terminated[e] = term ? 1 : 0;
      truncated[e] = tout ? 1 : 0;
      // _get_observations: refresh the stale cache from the (possibly reset) articulation view
      stale_from_links(base_pos, base_quat, base_vel, feet_pos, feet_quat, stale);
      mdp_observation(base_quat, q, qd, m.actions, m.speed_limit, obs_row);
      mdp_state_pack(m, wv, M_PDELTA, M_ACT, M_FLAST, M_FDPL, M_FSL, M_HSUM, M_YSUM, M_FFSUM, M_SPEED, M_EPSUM);
      stale_pack(stale, wv);
      store_words<NQ>(mstate, n, e, wv);
    }
    // observation rows of the tile: staged per warp (half a tile at a time), written as contiguous spans
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if ((lane >> 4) == half) {
#pragma unroll
        for (int i = 0; i < ZBOT_NUM_OBS; ++i) obs_stage[(lane & 15) * ZBOT_NUM_OBS + i] = obs_row[i];   // row of 23 words: odd, conflict-free
      }
      __syncwarp();
      const int total = max(0, min(16, valid - 16 * half)) * ZBOT_NUM_OBS;
      float* dst = obs + (size_t)(e0 + 16 * half) * ZBOT_NUM_OBS;
      if (((uintptr_t)obs & 15) == 0) {          // (e0 + 16 half) * 92 bytes is a multiple of 16
        const int nv = total >> 2;
        for (int i = lane; i < nv; i += 32) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(obs_stage)[i];
        for (int i = (nv << 2) + lane; i < total; i += 32) dst[i] = obs_stage[i];
      } else {
        for (int i = lane; i < total; i += 32) dst[i] = obs_stage[i];
      }
      __syncwarp();
    }
    any_reset_acc |= did_reset;
#pragma unroll
    for (int j = 0; j < kStatUsed; ++j) stat_acc[j] += stat[j];
  }

  // ---- statistics: lanes -> warp -> CTA -> one partial row ----------------------------------------------------------------
  {
    const bool any_reset = __any_sync(0xffffffffu, any_reset_acc);
#pragma unroll
    for (int j = 0; j < kStatUsed; ++j) {
      float v = 0.f;
      if ((j >= S_REW_SUM) || any_reset) v = warp_sum(stat_acc[j]);
      if (lane == 0) stat_sm[w * kStats + j] = v;
    }
    __syncthreads();
    float acc = 0.f;
    if (threadIdx.x < kStats) {
      if (threadIdx.x < kStatUsed)
        for (int ww = 0; ww < kWarps; ++ww) acc += stat_sm[ww * kStats + threadIdx.x];
      if (!sc.acc) sc.partials[(size_t)(blockIdx.x + sc.block_offset) * kStats + threadIdx.x] = acc;
    }
    if (sc.acc) {      // fused statistics (fixed-point accumulators + ticket): see stats_fused_commit
      __syncthreads();
      stats_fused_commit(sc, acc, stat_sm);
      return;
    }
  }
  // ---- grid-level pass in the last CTA to finish (no second launch): ticket, then the one-block finalize body -----------------
  if (sc.ticket) {
    __shared__ int s_last;
    float (*red)[33] = reinterpret_cast<float (*)[33]>(mpsm);   // every warp is past its tile loop: the stages are dead
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(sc.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
      __threadfence();
      stats_finalize_body(sc, gridDim.x, red);
      if (threadIdx.x == 0) *sc.ticket = 0u;
    }
  }
}
