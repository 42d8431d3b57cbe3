// zbot_kernels.cu -- sm_100a kernels + C ABI (include/zbot_b200.h) of the batched
// zbot-6b-walking-v2 environment step.
//
// Layout / execution model:
//   * one thread per environment; every per-env quantity lives in registers for the whole
//     control step (4 physics substeps + MDP + partial reset), so state is read once and
//     written once per step;
//   * state in HBM is AoSoA `float4 state[20][N]`: thread e issues 20 independent 16-byte
//     loads, a warp touches 512 contiguous bytes per load (fully coalesced);
//   * the [N][23] observation rows (AoS, as the policy consumes them) are staged through
//     shared memory and written back as contiguous float4 per block;
//   * per-step reset statistics (`extras["log"]`) are reduced with warp shuffles -> shared
//     memory -> one partial row per block; a one-block finalize kernel sums the rows in a fixed
//     order (deterministic, no float atomics, no host sync).
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "zbot_layout.h"
#include "zbot_pair.h"
#include "zbot_h2.h"
#include "zbot_halves.h"

using namespace zbot;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
#define ZB_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) return fail(ZBOT_E_CUDA, #call ": %s", cudaGetErrorString(e__)); \
  } while (0)

// step-kernel launch of a task entry point: `grid`, `block`, `smem`, `s` and the handle `h` are in scope at the call site
#define ZB_CUDA_LAUNCH(kernel, ...) ZB_CUDA(launch_pdl(kernel, dim3(grid), dim3(block), smem, s, h->pdl, __VA_ARGS__))

struct DefaultPose {      // FK of the ZBOT_6S_CFG init state, computed ON THE DEVICE at create time
  float feet_pos[2][3];   // env-local
  float feet_quat[2][4];
  float base_pos[3];
  float base_quat[4];
};

constexpr int kStats = ZBOT_STATS_WORDS;   // 32
constexpr int kStatUsed = 22;
constexpr int S_NUM_RESET = 16, S_NUM_TERM_RESET = 17, S_NUM_TO_RESET = 18, S_REW_SUM = 19, S_NUM_TERM = 20,
              S_NUM_TRUNC = 21;
// manager task: words 22..25 = episodic sum of the is_terminated RewTerm (normalised like a term) and the number of reset
// envs that tripped base_height / feet_close / illegal_contact (raw counts)
constexpr int kStatUsedM = 26, S_M_TERM_PENALTY = 22;

struct StatsCtx {
  float* partials;        // [max_blocks][32]
  float* ring;            // [slots][32]
  int slot, prev_slot;
  float inv_episode_s;    // 1 / max_episode_length_s  (…env_v2.py:444-447)
  int block_offset;       // env-range launches (zbot_step_host): index of this launch's first partial row
  const unsigned long long* rng_ctr;  // DEVICE counter = stream position of the in-kernel generator (v4 / manager events,
                                      // observation noise).  Read by the step kernel, bumped by the statistics kernel that
                                      // follows it, so it also advances when a captured CUDA graph is replayed (a host
                                      // counter passed by value is frozen at capture: every replay would re-draw the same
                                      // uniforms)
  int packed_rows;         // 1: host-facing output layout (zbot_step_host), see zbot_step_body
  int raw_tail;            // this many trailing term slots (…, 14, 15) hold raw counts: summed, not normalised (manager task)
  // all-envs-reset spread (…env_v2.py:418-422: `episode_length_buf[:] = randint_like(high=max_episode_length)` when EVERY env
  // reset in this step): done on the device by the statistics kernel, which holds the reset count -- no host sync, no work
  // unless the (rare) event fires.  Values come from the in-kernel counter generator (slot 64), not from torch's.
  int64_t* spread_ep_len;  // nullptr: off
  int spread_n, spread_high;
  unsigned long long spread_seed;
  int norm_word22;         // manager task: word 22 is an Episode_Reward value (normalised), not a raw count
  unsigned int* ticket;    // non-null: the producing kernel runs the grid-level pass itself in its LAST CTA to finish (a device
                           // counter, left at zero); no separate statistics launch follows (zbot_mdp_pipe_kernel)
  unsigned long long* acc; // non-null: FUSED statistics.  Every CTA adds its partial row to 32 fixed-point (2^-30) 64-bit accumulators
                           // with integer atomics -- integer addition commutes, so the totals are bit-reproducible whatever the
                           // order the CTAs finish in -- and the LAST CTA to finish (`ticket`) converts them, writes the ring slot,
                           // does the all-env-reset spread, bumps the generator position and clears the accumulators: the
                           // control step is ONE launch, and the next step's kernel follows it directly
  int pdl_early;           // 1: the step kernel releases its dependents (the statistics CTA, and through it the next step's CTAs)
                           // as soon as every one of its CTAs is running, so they are RESIDENT -- parked at their own
                           // griddepcontrol.wait -- when this grid completes, instead of being launched then
};

// ---------------------------------------------------------------------------------------------
// deterministic block + grid reduction of the per-thread statistics vector
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be
// scheduled while its predecessor in the stream drains; pdl_wait() blocks until that predecessor has completed and its
// writes are visible (a no-op for an ordinary launch); pdl_trigger() lets the successor's CTAs be scheduled early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint64_t rng_position(const StatsCtx& sc) { return sc.rng_ctr ? __ldcg(sc.rng_ctr) : 0ull; }

__device__ __forceinline__ float v4_uniform(uint64_t seed, uint64_t call, uint32_t env, uint32_t slot) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (call * 0x100000000ull + env) + 0xD1B54A32D192ED03ull * (slot + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

// ObservationManager-style corruption of the emitted observation row: obs[i] += lo[i] + u * (hi[i] - lo[i]), one
// counter-based uniform per (launch, env, column).  Slots 32.. keep clear of the v4 event slots 0..9.
template <int kCols>
__device__ __forceinline__ void obs_add_noise(const Params<float>& P, const StatsCtx& sc, int e, float* row) {
  if (!P.obs_noise_enable) return;
  const uint64_t call = rng_position(sc);
#pragma unroll
  for (int i = 0; i < kCols; ++i)
    row[i] = fmaf(v4_uniform(P.rng_seed, call, (uint32_t)e, 32u + (uint32_t)i), P.obs_noise_w[i], row[i] + P.obs_noise_lo[i]);
}

// Grid-level pass of the statistics: ONE block of 1024 threads.  Warp w sums the partial rows
// b = w, w+32, ... (a row is 128 contiguous bytes: one coalesced load per warp) with four independent
// accumulators so four L2 loads are in flight per thread; everything is combined in a fixed order, so the
// result is bit-reproducible (no float atomics).
// (any block size that is a multiple of 32; with 1024 threads the summation order is the one-block kernel's)
// The part after the totals (`tot[0..31]`, shared memory, visible to the block): normalisation, ring slot, all-env-reset spread,
// generator position.  Any block size that is a multiple of 32.
__device__ __forceinline__ void stats_finalize_tail(const StatsCtx& sc, const float* tot) {
  if (threadIdx.x < kStats) {
    const float nreset = tot[S_NUM_RESET];
    float v = tot[threadIdx.x];
    // words 0..15 leave the kernel as the reference's `Episode_Reward/<term>` values:
    // mean over the reset envs of the episodic sum, divided by max_episode_length_s
    if ((threadIdx.x < MAX_TERMS - sc.raw_tail || (sc.norm_word22 && threadIdx.x == S_M_TERM_PENALTY)) && nreset > 0.f)
      v = (v / nreset) * sc.inv_episode_s;
    // the reference only rewrites extras["log"] when something reset (…env_v2.py:450): keep the previous log
    // (word 16, the number of envs reset THIS step, is always the live count)
    if ((threadIdx.x < S_REW_SUM || threadIdx.x >= kStatUsed) && threadIdx.x != S_NUM_RESET && !(nreset > 0.f))
      v = (sc.prev_slot >= 0) ? sc.ring[(size_t)sc.prev_slot * kStats + threadIdx.x] : 0.f;
    sc.ring[(size_t)sc.slot * kStats + threadIdx.x] = v;
  }
  if (sc.spread_ep_len && tot[S_NUM_RESET] == (float)sc.spread_n) {   // block-uniform: every env reset in this step
    const uint64_t call = rng_position(sc);
    for (int i = threadIdx.x; i < sc.spread_n; i += blockDim.x) {
      const float u = v4_uniform(sc.spread_seed, call, (uint32_t)i, 64u);
      sc.spread_ep_len[i] = min((int)(u * (float)sc.spread_high), sc.spread_high - 1);
    }
    __syncthreads();                                                     // all reads of the position precede the bump
  }
  // advance the in-kernel generator's stream position: every CTA of this control step's kernel has finished (the separate
  // statistics kernel: pdl_wait; the fused pass: the ticket), the next one reads it only after this grid has completed
  if (threadIdx.x == 0 && sc.rng_ctr) *const_cast<unsigned long long*>(sc.rng_ctr) = *sc.rng_ctr + 1ull;
}

__device__ __forceinline__ void stats_finalize_body(const StatsCtx& sc, unsigned int nblocks, float (*red)[33]) {
  const int j = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int nw = blockDim.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* p = sc.partials + j;
  unsigned int b = w;
  for (; b + 3 * nw < nblocks; b += 4 * nw) {
    const float v0 = __ldcg(p + (size_t)b * kStats), v1 = __ldcg(p + (size_t)(b + nw) * kStats),
                v2 = __ldcg(p + (size_t)(b + 2 * nw) * kStats), v3 = __ldcg(p + (size_t)(b + 3 * nw) * kStats);
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; b < nblocks; b += nw) a0 += __ldcg(p + (size_t)b * kStats);
  red[w][j] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < kStats) {
    float acc = 0.f;
    for (unsigned int g = 0; g < nw; ++g) acc += red[g][threadIdx.x];
    red[0][threadIdx.x] = acc;           // words a kernel does not produce are zero in every partial row
  }
  __syncthreads();
  stats_finalize_tail(sc, red[0]);
}

// FUSED statistics (StatsCtx::acc): this CTA's total of word `threadIdx.x` (threads 0..31 carry one each, the others pass 0)
// goes to the fixed-point accumulators; the last CTA of the grid to get here runs the tail.  `tot`: >= 32 floats of shared
// memory nobody else uses any more.  Must be reached by every thread of every CTA of the grid.
constexpr float kStatFix = 1073741824.f;          // 2^30: a float32 partial of magnitude >= 2^-7 is represented exactly
__device__ __forceinline__ void stats_fused_commit(const StatsCtx& sc, float cta_total, float* tot) {
  __shared__ int s_last;
  if (threadIdx.x < kStats && cta_total != 0.f)
    atomicAdd(sc.acc + threadIdx.x, (unsigned long long)__float2ll_rn(cta_total * kStatFix));   // two's complement: signed add
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(sc.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < kStats) {
    const long long raw = (long long)atomicExch(sc.acc + threadIdx.x, 0ull);      // read and clear for the next control step
    tot[threadIdx.x] = (float)((double)raw * (1.0 / (double)kStatFix));
  }
  if (threadIdx.x == 0) *sc.ticket = 0u;
  __syncthreads();
  stats_finalize_tail(sc, tot);
}

__global__ void __launch_bounds__(1024) zbot_stats_finalize_kernel(StatsCtx sc, unsigned int nblocks) {
  __shared__ float red[32][33];
  pdl_trigger();   // the next step kernel may be scheduled behind this one-block kernel
  pdl_wait();      // the step kernel's partial rows are complete and visible
  stats_finalize_body(sc, nblocks, red);
}

// warp shuffles -> shared memory -> this block's partial row.  `vals[0..18]` are non-zero only for
// threads that reset this step; 19..21 for every thread.
// kN = statistics words this kernel produces (22 for the direct tasks, 26 for the manager task: words 22..25 are reset-only)
template <int kN>
__device__ __forceinline__ void stats_block_partial(float (&vals)[kN], bool did_reset, float* smem, const StatsCtx& sc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  const bool any_reset = __any_sync(0xffffffffu, did_reset);
#pragma unroll
  for (int j = 0; j < kN; ++j) {
    float v = 0.f;
    if ((j >= S_REW_SUM && j < kStatUsed) || any_reset) v = warp_sum(vals[j]);   // any_reset is warp-uniform
    if (lane == 0) smem[warp * kN + j] = v;
  }
  __syncthreads();
  float acc = 0.f;
  if (threadIdx.x < kStats) {            // a whole 32-word row: the words this kernel does not produce are written as zeros
    if (threadIdx.x < kN)
      for (int w = 0; w < nwarps; ++w) acc += smem[w * kN + threadIdx.x];
    if (!sc.acc) sc.partials[(size_t)(blockIdx.x + sc.block_offset) * kStats + threadIdx.x] = acc;
  }
  if (sc.acc) {                          // grid-uniform
    __syncthreads();                     // the warp rows have been read: smem is free for the totals
    stats_fused_commit(sc, acc, smem);
  }
}

// ---------------------------------------------------------------------------------------------
// state load / store: 20 x float4 per thread, [quad][env] layout
// ---------------------------------------------------------------------------------------------
template <int NQ>
__device__ __forceinline__ void load_words(const float4* __restrict__ st, int n, int e, float* w) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 v = __ldg(st + (size_t)q * n + e);
    w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
}
#ifndef ZB_PREFETCH_LATE
#define ZB_PREFETCH_LATE 1
#endif
// L2 prefetch of the quads a thread will load AFTER the physics phase (the MDP state) and of its episode counter: when the
// state is not L2-resident (a cold step: another kernel ran through the L2 in between) their DRAM latency would be exposed at
// the start of the MDP phase, where every warp of a wave arrives at once; issued next to the early loads it is hidden
// behind the four substeps.  A hint: no register, no dependency; an L2 hit makes it a no-op.  Used by the rolled (small-N)
// instantiations only: bench with the L2 flushed 37.4 -> 36.1 us at 4096 envs, but 85.4 -> 85.9 us at 65536 envs.
template <int NQ>
__device__ __forceinline__ void prefetch_words_l2(const float4* __restrict__ st, int n, int e) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(st + (size_t)q * n + e));
}
template <int NQ>
__device__ __forceinline__ void store_words(float4* __restrict__ st, int n, int e, const float* w) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) st[(size_t)q * n + e] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// stage per-thread rows of ROW floats through shared memory, write the block's rows contiguously
template <int ROW>
__device__ __forceinline__ void store_rows_coalesced(float* __restrict__ dst, const float* row, int n, int e0,
                                                     float* smem /*blockDim*ROW*/, int tile = 0 /*envs of this CTA; 0: blockDim*/) {
  const int e = e0 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < ROW; ++i) smem[threadIdx.x * ROW + i] = row[i];   // ROW odd -> conflict-free
  __syncthreads();
  const int valid = min(tile ? tile : (int)blockDim.x, n - e0);
  const int total = valid * ROW;
  float* base = dst + (size_t)e0 * ROW;
  if ((((size_t)e0 * ROW) & 3) == 0 && ((uintptr_t)dst & 15) == 0) {
    const int nv = total >> 2;
    for (int i = threadIdx.x; i < nv; i += blockDim.x)
      reinterpret_cast<float4*>(base)[i] = reinterpret_cast<const float4*>(smem)[i];
    for (int i = (nv << 2) + threadIdx.x; i < total; i += blockDim.x) base[i] = smem[i];
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) base[i] = smem[i];
  }
  (void)e;
}

struct ExportPtrs {
  float *pos0, *quat0, *vel0, *pos1, *quat1, *vel1, *q1, *qd1, *tau1, *hist1, *last_air1, *cur_contact1;
};

// sensor-body index (USD prim order: b1 a2 b2 a3 b3 b4 a5 b5 a6 foot_0 foot_1 base) of the link a
// merged body's sphere force is attributed to (the "a" half: a2 a3 base a5 a6)
__device__ __constant__ int kMidSensorIdx[5] = {1, 3, 11, 6, 8};
constexpr int kFoot0Sensor = 9, kFoot1Sensor = 10;

// kExport flavours: write the articulation / contact-sensor view the MDP phase of THIS launch saw (test hook: the
// reference-pinned MDP oracle is evaluated on it).  Layouts = the reference's robot.data / contact_sensor.data tensors.
__device__ __noinline__ void step_export_store(const StepExport<float>& ex, const ExportPtrs& xp, int e) {
  for (int i = 0; i < 36; ++i) { xp.pos0[(size_t)e * 36 + i] = ex.pos0[i]; xp.vel0[(size_t)e * 36 + i] = ex.vel0[i];
                                 xp.pos1[(size_t)e * 36 + i] = ex.pos1[i]; xp.vel1[(size_t)e * 36 + i] = ex.vel1[i]; }
  for (int i = 0; i < 48; ++i) { xp.quat0[(size_t)e * 48 + i] = ex.quat0[i]; xp.quat1[(size_t)e * 48 + i] = ex.quat1[i]; }
  for (int i = 0; i < 6; ++i) { xp.q1[(size_t)e * 6 + i] = ex.q1[i]; xp.qd1[(size_t)e * 6 + i] = ex.qd1[i];
                                xp.tau1[(size_t)e * 6 + i] = ex.applied_torque[i]; }
  float* h = xp.hist1 + (size_t)e * (5 * 12 * 3);
  for (int i = 0; i < 5 * 12 * 3; ++i) h[i] = 0.f;
  for (int t = 0; t < 5; ++t) {
    for (int i = 0; i < 3; ++i) {
      h[(t * 12 + kFoot0Sensor) * 3 + i] = ex.feet_force_hist[t][0][i];
      h[(t * 12 + kFoot1Sensor) * 3 + i] = ex.feet_force_hist[t][1][i];
      for (int b = 0; b < 5; ++b) h[(t * 12 + kMidSensorIdx[b]) * 3 + i] = ex.mid_force_hist[t][b][i];
    }
  }
  for (int b = 0; b < 12; ++b) { xp.last_air1[(size_t)e * 12 + b] = 0.f; xp.cur_contact1[(size_t)e * 12 + b] = 0.f; }
  xp.last_air1[(size_t)e * 12 + kFoot0Sensor] = ex.last_air[0];
  xp.last_air1[(size_t)e * 12 + kFoot1Sensor] = ex.last_air[1];
  xp.cur_contact1[(size_t)e * 12 + kFoot0Sensor] = ex.cur_contact[0];
  xp.cur_contact1[(size_t)e * 12 + kFoot1Sensor] = ex.cur_contact[1];
}

// ---------------------------------------------------------------------------------------------
// the fused control step
// ---------------------------------------------------------------------------------------------
// sc.packed_rows: host-facing output layout -- ONE row of 25 words per env [obs 23 | reward | flags word
// (terminated | truncated << 8)] written to `obs` ([N][25]); `rew` / `terminated` / `truncated` are not touched.
// A run-time flag of the SAME kernel (not a second instantiation), so step_host and step are bit-identical.
// kPhys: 0 = one chain per thread (physics_substep), 1 = both halves of the chain in the two FP32 lanes (zbot_h2.h)
constexpr int kH2RowF2 = HALF_SCR_WORDS + 4;   // 55 float2 per thread (odd: conflict-free 64-bit rows): 51 scratch words + the raw actions
static_assert((kH2RowF2 & 1) == 1 && 2 * kH2RowF2 <= SCR_STRIDE + 1, "h2 scratch row");
constexpr int kStepRowWords = SCR_STRIDE + 1;  // floats of dynamic shared memory per thread of a walking step kernel
template <bool kExport, int kUnroll = 1, int kPhys = 0>
__device__ __forceinline__ void
zbot_step_body(const Params<float>& P, const DefaultPose& dp,
                 float4* __restrict__ state, int64_t* __restrict__ ep_len_buf,
                 const float* __restrict__ actions, float* __restrict__ obs, float* __restrict__ rew,
                 uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, int n, int e_begin, int e_end,
                 StatsCtx sc, ExportPtrs xp) {
  extern __shared__ float smem[];   // blockDim*SCR_STRIDE floats: substep scratch, then obs rows, then stats
  pdl_wait();                       // ordered after the previous kernel of the stream (no-op unless launched with PDL)
  if (sc.pdl_early) pdl_trigger();
  // this launch covers envs [e_begin, e_end) of the n-env state (whole range: 0, n)
  const int e0 = e_begin + blockIdx.x * blockDim.x;
  const int e = e0 + threadIdx.x;
  const bool live = e < e_end;
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  const bool kPacked = !kExport && sc.packed_rows != 0;
  float obs_row[ZBOT_HOST_ROW_WORDS];
#pragma unroll
  for (int i = 0; i < ZBOT_HOST_ROW_WORDS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;
  if (live) {
    EnvState<float> es;
    StepOut<float> out;
    float rs[MAX_TERMS];
#pragma unroll
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = 0.f;
    float* const row = smem + threadIdx.x * (kPhys == 1 ? 2 * kH2RowF2 : SCR_STRIDE);
    float* const raw_park = row + (kPhys == 1 ? 2 * HALF_SCR_WORDS : SCR_RAW_ACT);
    const float2* a2p = reinterpret_cast<const float2*>(actions + (size_t)e * 6);
    int64_t ep;
    StepExport<float> ex;   // kExport only (dead otherwise): the articulation / sensor view the MDP phase saw
    {
      // ---- phase A: only the 11 "early" quads (articulation state, p_delta, contact carry, timers) ----
      {
        float w[4 * EARLY_QUADS];
        load_words<EARLY_QUADS>(state, n, e, w);
        if (ZB_PREFETCH_LATE && kUnroll == 1) { prefetch_words_l2<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(ep_len_buf + e)); }
        env_early_unpack(w, es);
      }
      PhysOut<float> po;
      {
        const float2 a0 = __ldg(a2p), a1 = __ldg(a2p + 1), a2v = __ldg(a2p + 2);
        const float raw[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
        // read ONCE (the buffer may be pinned host memory: zero-copy over PCIe) and parked in this thread's
        // shared-memory row for phase C
#pragma unroll
        for (int k = 0; k < 6; ++k) raw_park[k] = raw[k];
        // ---- phase B: 4 physics substeps; the MDP state is not even loaded yet (register budget) ----
        if constexpr (kPhys == 1) {
          SmemScratch2 scr{reinterpret_cast<float2*>(row)};
          env_step_physics_h2<ModelWalk>(P, es, raw, po, scr, kExport ? &ex : (StepExport<float>*)nullptr);
        } else {
          SmemScratch scr{row};
          env_step_physics<ModelWalk, kUnroll>(P, es, raw, po, scr, kExport ? &ex : (StepExport<float>*)nullptr);
        }
      }
      // ---- phase C: the 9 "late" quads, the start-of-step state S0 again (still unmodified in global
      //      memory -> L2 hit) for the one-step-stale quantities, the raw actions again, then the MDP ----
      {
        float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
        load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e, w);
        env_late_unpack(w, es);
      }
      SimState<float> s0;
      {
        float w[4 * SIM_QUADS];
        load_words<SIM_QUADS>(state, n, e, w);
        sim_state_unpack(w, s0);
      }
      float raw[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) raw[k] = raw_park[k];
      ep = ep_len_buf[e];
      env_step_finish(P, es, s0, raw, po, ep, dp.feet_pos, dp.base_quat, out, rs, kExport ? &ex : (StepExport<float>*)nullptr);
      if (kExport) step_export_store(ex, xp, e);
    }
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
  __syncthreads();   // every thread is done with its scratch column before the rows are staged
  if (kPacked) store_rows_coalesced<ZBOT_HOST_ROW_WORDS>(obs, obs_row, e_end, e0, smem);   // block-uniform branch
  else store_rows_coalesced<ZBOT_NUM_OBS>(obs, obs_row, e_end, e0, smem);
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

#define ZB_STEP_ARGS                                                                                   \
  const __grid_constant__ Params<float> P, const __grid_constant__ DefaultPose dp, float4 *__restrict__ state, \
      int64_t *__restrict__ ep_len_buf, const float *__restrict__ actions, float *__restrict__ obs,      \
      float *__restrict__ rew, uint8_t *__restrict__ terminated, uint8_t *__restrict__ truncated, int n, \
      int e_begin, int e_end, StatsCtx sc, ExportPtrs xp
#define ZB_STEP_CALL P, dp, state, ep_len_buf, actions, obs, rew, terminated, truncated, n, e_begin, e_end, sc, xp

// register-budget variants of the same body (DESIGN.md §4 "occupancy"): kMinBlocks resident CTAs of
// kMaxThreads threads per SM
template <bool kExport, int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) zbot_step_kernel(ZB_STEP_ARGS) {
  zbot_step_body<kExport>(ZB_STEP_CALL);
}
// ---------------------------------------------------------------------------------------------
// the fused control step, TWO ENVIRONMENTS PER THREAD: the four physics substeps run once per thread with T = F2
// (packed FFMA2 / FMUL2 / FADD2: every FP32 instruction of the articulated-body recursion serves both envs); the MDP
// phase runs scalar, once per lane.  A CTA of B threads owns 2B consecutive envs: lane 0 = env e0 + t, lane 1 =
// env e0 + B + t, so every state load / store is as coalesced as in the one-env kernel.
// ---------------------------------------------------------------------------------------------
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) zbot_step2_kernel(ZB_STEP_ARGS) {
  extern __shared__ float smem[];   // blockDim * SCR_STRIDE float2: substep scratch, then obs rows, then stats
  const int B = blockDim.x;
  const int e0 = e_begin + blockIdx.x * 2 * B;
  const int ea_i = e0 + threadIdx.x, eb_i = e0 + B + threadIdx.x;
  const bool live_a = ea_i < e_end, live_b = eb_i < e_end;
  const int eb_ld = live_b ? eb_i : ea_i;          // a dead lane shadows lane 0 (never stored)
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  float obs_a[ZBOT_NUM_OBS], obs_b[ZBOT_NUM_OBS];
#pragma unroll
  for (int i = 0; i < ZBOT_NUM_OBS; ++i) { obs_a[i] = 0.f; obs_b[i] = 0.f; }
  bool did_reset = false;
  if (live_a) {
    EnvState<float> es_a, es_b;
    PhysOut<float> po_a, po_b;
    SmemScratch2 scr{reinterpret_cast<float2*>(smem) + threadIdx.x * SCR_STRIDE};
    {
      float w[4 * EARLY_QUADS];
      load_words<EARLY_QUADS>(state, n, ea_i, w);
      env_early_unpack(w, es_a);
      load_words<EARLY_QUADS>(state, n, eb_ld, w);
      env_early_unpack(w, es_b);
    }
    {
      const float2* pa = reinterpret_cast<const float2*>(actions + (size_t)ea_i * 6);
      const float2* pb = reinterpret_cast<const float2*>(actions + (size_t)eb_ld * 6);
      const float2 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2v = __ldg(pa + 2);
      const float2 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2v = __ldg(pb + 2);
      const float raw_a[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
      const float raw_b[6] = {b0.x, b0.y, b1.x, b1.y, b2v.x, b2v.y};
#pragma unroll
      for (int k = 0; k < 6; ++k) scr.base[SCR_RAW_ACT + k] = make_float2(raw_a[k], raw_b[k]);
      env_step_physics2<ModelWalk>(P, es_a, es_b, raw_a, raw_b, po_a, po_b, scr);
    }
    // ---- phase C, once per lane (scalar) ----
#pragma unroll 1
    for (int l = 0; l < 2; ++l) {
      const bool live = l ? live_b : true;
      if (!live) break;
      const int e = l ? eb_i : ea_i;
      // one code instance of the MDP phase: lane 1's physics results are moved into the lane-0 structs after lane 0 is
      // done (register moves; a run-time choice between two structs would push both into local memory)
      if (l) { es_a = es_b; po_a = po_b; }
      EnvState<float>& es = es_a;
      const PhysOut<float>& po = po_a;
      {
        float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
        load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e, w);
        env_late_unpack(w, es);
      }
      SimState<float> s0;
      {
        float w[4 * SIM_QUADS];
        load_words<SIM_QUADS>(state, n, e, w);
        sim_state_unpack(w, s0);
      }
      float raw[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const float2 v = scr.base[SCR_RAW_ACT + k]; raw[k] = l ? v.y : v.x; }
      int64_t ep = ep_len_buf[e];
      StepOut<float> out;
      float rs[MAX_TERMS];
#pragma unroll
      for (int i = 0; i < MAX_TERMS; ++i) rs[i] = 0.f;
      env_step_finish(P, es, s0, raw, po, ep, dp.feet_pos, dp.base_quat, out, rs, (StepExport<float>*)nullptr);
      float w[ZBOT_STATE_WORDS];
      env_state_pack(es, w);
      store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
      ep_len_buf[e] = ep;
      rew[e] = out.reward;
      terminated[e] = out.terminated ? 1 : 0;
      truncated[e] = out.time_out ? 1 : 0;
      obs_add_noise<ZBOT_NUM_OBS>(P, sc, e, out.obs);
#pragma unroll
      for (int i = 0; i < ZBOT_NUM_OBS; ++i) { if (l) obs_b[i] = out.obs[i]; else obs_a[i] = out.obs[i]; }
      if (out.terminated || out.time_out) {
        did_reset = true;
#pragma unroll
        for (int i = 0; i < MAX_TERMS; ++i) stat[i] += rs[i];
        stat[S_NUM_RESET] += 1.f;
        stat[S_NUM_TERM_RESET] += out.terminated ? 1.f : 0.f;
        stat[S_NUM_TO_RESET] += out.time_out ? 1.f : 0.f;
      }
      stat[S_REW_SUM] += out.reward;
      stat[S_NUM_TERM] += out.terminated ? 1.f : 0.f;
      stat[S_NUM_TRUNC] += out.time_out ? 1.f : 0.f;
    }
  }
  __syncthreads();   // every thread is done with its scratch row before the rows are staged
  store_rows_coalesced<ZBOT_NUM_OBS>(obs, obs_a, e_end, e0, smem);
  __syncthreads();
  if (e0 + B < e_end) store_rows_coalesced<ZBOT_NUM_OBS>(obs, obs_b, e_end, e0 + B, smem);   // block-uniform condition
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

// chain sweeps unrolled by two (large N: two or more warps per sub-partition)
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) zbot_step_u2_kernel(ZB_STEP_ARGS) {
  zbot_step_body<false, 2>(ZB_STEP_CALL);
}
// export flavour of the unrolled instantiation (test hook): the SAME phased body and sweep unroll as the kernel that is
// launched -- and benchmarked -- above 18944 envs, plus the stores of the view its MDP phase saw
__global__ void __launch_bounds__(128, 1) zbot_step_u2_export_kernel(ZB_STEP_ARGS) {
  zbot_step_body<true, 2>(ZB_STEP_CALL);
}
// both halves of the chain in the two FP32 lanes of one thread (zbot_h2.h): "h128x2"
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) zbot_step_h2_kernel(ZB_STEP_ARGS) {
  zbot_step_body<false, 1, 1>(ZB_STEP_CALL);
}
__global__ void __launch_bounds__(128, 1) zbot_step_h2_export_kernel(ZB_STEP_ARGS) {
  zbot_step_body<true, 1, 1>(ZB_STEP_CALL);
}
#include "zbot_w2_kernel.cuh"   // two warps per 32 envs (the default walking-v2 step kernel)
#include "zbot_policy.cuh"      // the act / store halves of the PPO rollout (f4)
#include "zbot_policy_tc.cuh"   // the act half on the tensor cores, mma.sync build (3 x TF32 split products: FP32 accuracy)
// unrolled sweeps under a direct register cap (single-wave experiments: 14 warps/SM hold 65536 envs at <= 146 registers)
template <int kMaxRegs>
__global__ void __maxnreg__(kMaxRegs) zbot_step_u2_kernel_r(ZB_STEP_ARGS) {
  zbot_step_body<false, 2>(ZB_STEP_CALL);
}
// same body, register cap given directly (ptxas snaps __launch_bounds__ caps to a few occupancy steps:
// 197 -> 168 -> 128; __maxnreg__ gives the steps in between)
template <int kMaxRegs>
__global__ void __maxnreg__(kMaxRegs) zbot_step_kernel_r(ZB_STEP_ARGS) {
  zbot_step_body<false>(ZB_STEP_CALL);
}
// ---------------------------------------------------------------------------------------------
// snake task (zbot-6s-snake-v0): same state layout, same physics substep (ModelSnake), its own MDP.
// Phased like the walking kernel: early quads -> 4 substeps -> late quads + S0 re-read (L2) -> MDP.
// ---------------------------------------------------------------------------------------------
constexpr int kSnakeExportWords = (int)(sizeof(SnakeExport<float>) / sizeof(float));   // 41

template <bool kExport, int kUnroll = 1, int kMinBlocks = 2>
__global__ void __launch_bounds__(128, kMinBlocks)
zbot_snake_step_kernel(const __grid_constant__ Params<float> P, const __grid_constant__ DefaultPose dp,
                       float4* __restrict__ state, int64_t* __restrict__ ep_len_buf, const float* __restrict__ actions,
                       float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated,
                       uint8_t* __restrict__ truncated, int n, StatsCtx sc, float* __restrict__ export_buf) {
  extern __shared__ float smem[];
  pdl_wait();
  const int e0 = blockIdx.x * blockDim.x;
  const int e = e0 + threadIdx.x;
  const bool live = e < n;
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  float obs_row[ZBOT_NUM_OBS];
#pragma unroll
  for (int i = 0; i < ZBOT_NUM_OBS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;
  if (live) {
    EnvState<float> es;
    StepOut<float> out;
    float rs[MAX_TERMS];
#pragma unroll
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = 0.f;
    SmemScratch scr{smem + threadIdx.x * SCR_STRIDE};
    const float2* a2p = reinterpret_cast<const float2*>(actions + (size_t)e * 6);
    {
      float w[4 * EARLY_QUADS];
      load_words<EARLY_QUADS>(state, n, e, w);
        if (ZB_PREFETCH_LATE && kUnroll == 1) { prefetch_words_l2<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(ep_len_buf + e)); }
      env_early_unpack(w, es);
    }
    PhysOut<float> po;
    {
      const float2 a0 = __ldg(a2p), a1 = __ldg(a2p + 1), a2v = __ldg(a2p + 2);
      const float raw[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
#pragma unroll
      for (int k = 0; k < 6; ++k) scr.base[SCR_RAW_ACT + k] = raw[k];
      env_step_physics<ModelSnake, kUnroll>(P, es, raw, po, scr, (StepExport<float>*)nullptr);
    }
    {
      float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
      load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e, w);
      env_late_unpack(w, es);
    }
    SimState<float> s0;
    {
      float w[4 * SIM_QUADS];
      load_words<SIM_QUADS>(state, n, e, w);
      sim_state_unpack(w, s0);
    }
    float raw[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) raw[k] = scr.base[SCR_RAW_ACT + k];
    int64_t ep = ep_len_buf[e];
    SnakeExport<float> ex;
    snake_step_finish(P, es, s0, raw, po, ep, dp.base_quat, out, rs, kExport ? &ex : (SnakeExport<float>*)nullptr);
    if (kExport) {
      const float* src = reinterpret_cast<const float*>(&ex);
      for (int i = 0; i < kSnakeExportWords; ++i) export_buf[(size_t)e * kSnakeExportWords + i] = src[i];
    }
    float w[ZBOT_STATE_WORDS];
    env_state_pack(es, w);
    store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
    ep_len_buf[e] = ep;
    rew[e] = out.reward;
    terminated[e] = out.terminated ? 1 : 0;
    truncated[e] = out.time_out ? 1 : 0;
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
  __syncthreads();
  store_rows_coalesced<ZBOT_NUM_OBS>(obs, obs_row, n, e0, smem);
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

// ---------------------------------------------------------------------------------------------
// zbot-6b-walking-v4: same phased structure; all MDP inputs are FRESH, so the start-of-step state is not
// re-read.  Random numbers: caller-supplied [N][10] uniforms, or a counter-based generator
// (splitmix64 finaliser of seed / call counter / env / slot -> 24-bit mantissa, like torch.rand's float32).
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(V4Export<float>) / sizeof(float) == ZBOT_V4_EXPORT_WORDS, "V4Export layout");

// kH2: the physics phase as packed halves (zbot_h2.h), like zbot_step_h2_kernel
template <bool kExport, int kUnroll = 1, int kMinBlocks = 2, bool kH2 = false>
__global__ void __launch_bounds__(128, kMinBlocks)
zbot_v4_step_kernel(const __grid_constant__ Params<float> P, float4* __restrict__ state, int64_t* __restrict__ ep_len_buf,
                    const float* __restrict__ actions, const float* __restrict__ rand, uint64_t seed,
                    float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated,
                    uint8_t* __restrict__ truncated, int n, StatsCtx sc, float* __restrict__ export_buf) {
  extern __shared__ float smem[];
  pdl_wait();
  const int e0 = blockIdx.x * blockDim.x;
  const int e = e0 + threadIdx.x;
  const bool live = e < n;
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  float obs_row[ZBOT_V4_NUM_OBS];
#pragma unroll
  for (int i = 0; i < ZBOT_V4_NUM_OBS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;
  if (live) {
    EnvState<float> es;
    StepOut<float> out;
    float rs[MAX_TERMS];
#pragma unroll
    for (int i = 0; i < MAX_TERMS; ++i) rs[i] = 0.f;
    float* const row = smem + threadIdx.x * (kH2 ? 2 * kH2RowF2 : SCR_STRIDE);
    float* const raw_park = row + (kH2 ? 2 * HALF_SCR_WORDS : SCR_RAW_ACT);
    const float2* a2p = reinterpret_cast<const float2*>(actions + (size_t)e * 6);
    {
      float w[4 * EARLY_QUADS];
      load_words<EARLY_QUADS>(state, n, e, w);
        if (ZB_PREFETCH_LATE && kUnroll == 1) { prefetch_words_l2<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(ep_len_buf + e)); }
      env_early_unpack(w, es);
    }
    PhysOut<float> po;
    {
      const float2 a0 = __ldg(a2p), a1 = __ldg(a2p + 1), a2v = __ldg(a2p + 2);
      const float raw[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
#pragma unroll
      for (int k = 0; k < 6; ++k) raw_park[k] = raw[k];
      if constexpr (kH2) {
        SmemScratch2 scr{reinterpret_cast<float2*>(row)};
        env_step_physics_h2<ModelWalkV4>(P, es, raw, po, scr, (StepExport<float>*)nullptr);
      } else {
        SmemScratch scr{row};
        env_step_physics<ModelWalkV4, kUnroll>(P, es, raw, po, scr, (StepExport<float>*)nullptr);
      }
    }
    {
      float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
      load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e, w);
      env_late_unpack(w, es);
    }
    float raw[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) raw[k] = raw_park[k];
    float rnd[V4_NUM_RAND];
    if (rand) {
      const float2* rp = reinterpret_cast<const float2*>(rand + (size_t)e * V4_NUM_RAND);
#pragma unroll
      for (int i = 0; i < V4_NUM_RAND / 2; ++i) { const float2 v = __ldg(rp + i); rnd[2 * i] = v.x; rnd[2 * i + 1] = v.y; }
    } else {
      const uint64_t call = rng_position(sc);
#pragma unroll
      for (int i = 0; i < V4_NUM_RAND; ++i) rnd[i] = v4_uniform(seed, call, (uint32_t)e, (uint32_t)i);
    }
    int64_t ep = ep_len_buf[e];
    V4Export<float> ex;
    v4_step_finish(P, es, raw, po, ep, rnd, obs_row, out, rs, kExport ? &ex : (V4Export<float>*)nullptr);
    obs_add_noise<ZBOT_V4_NUM_OBS>(P, sc, e, obs_row);
    if (kExport) {
      const float* src = reinterpret_cast<const float*>(&ex);
      for (int i = 0; i < ZBOT_V4_EXPORT_WORDS; ++i) export_buf[(size_t)e * ZBOT_V4_EXPORT_WORDS + i] = src[i];
    }
    float w[ZBOT_STATE_WORDS];
    env_state_pack(es, w);
    store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
    ep_len_buf[e] = ep;
    rew[e] = out.reward;
    terminated[e] = out.terminated ? 1 : 0;
    truncated[e] = out.time_out ? 1 : 0;
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
  __syncthreads();
  store_rows_coalesced<ZBOT_V4_NUM_OBS>(obs, obs_row, n, e0, smem);
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

// ---------------------------------------------------------------------------------------------
// zbot-6b-walking-m-v0 (manager-based task): same phased structure as the v4 kernel -- early quads, 4 substeps with the
// relative joint-position action re-targeted at every substep, late quads, manager-ordered MDP -- on ModelWalkM (full
// inertia tables, per-env friction).  Random numbers: caller-supplied [N][13] uniforms or the counter-based generator.
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(MExport<float>) / sizeof(float) == ZBOT_M_EXPORT_WORDS && M_EXPORT_WORDS == ZBOT_M_EXPORT_WORDS, "MExport layout");
static_assert(M_NUM_OBS == ZBOT_M_NUM_OBS && M_NUM_RAND == ZBOT_M_NUM_RAND, "manager task widths");

// rough-terrain arguments of the manager kernel (zbot_bind_terrain); heights == nullptr: flat ground
struct TerrainArgs {
  const float* heights;        // [nx][ny] world-frame height field
  int nx, ny;
  float x0, y0, inv_cell;
  float4* env_origins;         // [N] (x, y, z, unused): read every step, rewritten when the curriculum moves an env
  const float* tile_origins;   // [rows][cols][3]
  int rows, cols;
  float tile_size, episode_s;
  int curriculum;
};

template <bool kExport, int kUnroll = 1, int kMinBlocks = 2, bool kTerrain = false>
__global__ void __launch_bounds__(128, kMinBlocks)
zbot_m_step_kernel(const __grid_constant__ Params<float> P, float4* __restrict__ state, int64_t* __restrict__ ep_len_buf,
                   const float* __restrict__ actions, const float* __restrict__ rand, uint64_t seed,
                   float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated,
                   uint8_t* __restrict__ truncated, int n, StatsCtx sc, float* __restrict__ export_buf, TerrainArgs ta) {
  extern __shared__ float smem[];
  pdl_wait();
  const int e0 = blockIdx.x * blockDim.x;
  const int e = e0 + threadIdx.x;
  const bool live = e < n;
  float stat[kStatUsedM];
#pragma unroll
  for (int j = 0; j < kStatUsedM; ++j) stat[j] = 0.f;
  float obs_row[M_NUM_OBS];
#pragma unroll
  for (int i = 0; i < M_NUM_OBS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;
  if (live) {
    EnvState<float> es;
    StepOut<float> out;
    float rs[MAX_TERMS + 4];
#pragma unroll
    for (int i = 0; i < MAX_TERMS + 4; ++i) rs[i] = 0.f;
    SmemScratch scr{smem + threadIdx.x * SCR_STRIDE};
    const float2* a2p = reinterpret_cast<const float2*>(actions + (size_t)e * 6);
    {
      float w[4 * EARLY_QUADS];
      load_words<EARLY_QUADS>(state, n, e, w);
        if (ZB_PREFETCH_LATE && kUnroll == 1) { prefetch_words_l2<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(ep_len_buf + e)); }
      env_early_unpack(w, es);
    }
    PhysOut<float> po;
    float4 org = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kTerrain) org = ta.env_origins[e];
    {
      const float2 a0 = __ldg(a2p), a1 = __ldg(a2p + 1), a2v = __ldg(a2p + 2);
      const float raw[6] = {a0.x, a0.y, a1.x, a1.y, a2v.x, a2v.y};
#pragma unroll
      for (int k = 0; k < 6; ++k) scr.base[SCR_RAW_ACT + k] = raw[k];
      if (kTerrain) {
        const TerrainGround<float> ground{ta.heights, ta.nx, ta.ny, ta.x0, ta.y0, ta.inv_cell, org.x, org.y, org.z};
        env_step_physics<ModelWalkM, kUnroll>(P, es, raw, po, scr, (StepExport<float>*)nullptr, ground);
      } else {
        env_step_physics<ModelWalkM, kUnroll>(P, es, raw, po, scr, (StepExport<float>*)nullptr);
      }
    }
    {
      float w[ZBOT_STATE_WORDS - 4 * EARLY_QUADS];
      load_words<ZBOT_STATE_WORDS / 4 - EARLY_QUADS>(state + (size_t)EARLY_QUADS * n, n, e, w);
      env_late_unpack(w, es);
    }
    float raw[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) raw[k] = scr.base[SCR_RAW_ACT + k];
    float rnd[M_NUM_RAND];
    if (rand) {
#pragma unroll
      for (int i = 0; i < M_NUM_RAND; ++i) rnd[i] = __ldg(rand + (size_t)e * M_NUM_RAND + i);
    } else {
      const uint64_t call = rng_position(sc);
#pragma unroll
      for (int i = 0; i < M_NUM_RAND; ++i) rnd[i] = v4_uniform(seed, call, (uint32_t)e, (uint32_t)i);
    }
    int64_t ep = ep_len_buf[e];
    MExport<float> ex;
    if (kTerrain) {
      MTerrainCtx<float> tc{{org.x, org.y, org.z}, ta.tile_origins, ta.rows, ta.cols, ta.tile_size, ta.episode_s, ta.curriculum};
      m_step_finish(P, es, raw, po, ep, rnd, obs_row, out, rs, kExport ? &ex : (MExport<float>*)nullptr, &tc);
      if (out.terminated || out.time_out) ta.env_origins[e] = make_float4(tc.origin[0], tc.origin[1], tc.origin[2], 0.f);
    } else {
      m_step_finish(P, es, raw, po, ep, rnd, obs_row, out, rs, kExport ? &ex : (MExport<float>*)nullptr);
    }
    // ObservationManager corruption (PolicyCfg: base_quat +-0.01, joint_pos +-0.01, joint_vel +-1.5); columns 0..23
    if (P.obs_noise_enable) {
      const uint64_t call = rng_position(sc);
#pragma unroll
      for (int i = 0; i < 24; ++i)
        obs_row[i] = fmaf(v4_uniform(P.rng_seed, call, (uint32_t)e, 32u + (uint32_t)i), P.obs_noise_w[i], obs_row[i] + P.obs_noise_lo[i]);
    }
    if (kExport) {
      const float* src = reinterpret_cast<const float*>(&ex);
      for (int i = 0; i < ZBOT_M_EXPORT_WORDS; ++i) export_buf[(size_t)e * ZBOT_M_EXPORT_WORDS + i] = src[i];
    }
    float w[ZBOT_STATE_WORDS];
    env_state_pack(es, w);
    store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
    ep_len_buf[e] = ep;
    rew[e] = out.reward;
    terminated[e] = out.terminated ? 1 : 0;
    truncated[e] = out.time_out ? 1 : 0;
    did_reset = out.terminated || out.time_out;
    if (did_reset) {
#pragma unroll
      for (int i = 0; i < MAX_TERMS; ++i) stat[i] = rs[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) stat[kStatUsed + i] = rs[MAX_TERMS + i];
      stat[S_NUM_RESET] = 1.f;
      stat[S_NUM_TERM_RESET] = out.terminated ? 1.f : 0.f;
      stat[S_NUM_TO_RESET] = out.time_out ? 1.f : 0.f;
    }
    stat[S_REW_SUM] = out.reward;
    stat[S_NUM_TERM] = out.terminated ? 1.f : 0.f;
    stat[S_NUM_TRUNC] = out.time_out ? 1.f : 0.f;
  }
  __syncthreads();
  store_rows_coalesced<M_NUM_OBS>(obs, obs_row, n, e0, smem);
  __syncthreads();
  stats_block_partial(stat, did_reset, smem, sc);
}

// ---------------------------------------------------------------------------------------------
// reset / observe / articulation view / init
// ---------------------------------------------------------------------------------------------
template <bool kSnake>
__global__ void zbot_reset_kernel(const __grid_constant__ Params<float> P, const __grid_constant__ DefaultPose dp,
                                  float4* __restrict__ state, int64_t* __restrict__ ep_len_buf,
                                  const int64_t* __restrict__ ids, int64_t nids, const uint8_t* __restrict__ terminated,
                                  const uint8_t* __restrict__ truncated, int n, StatsCtx sc) {
  extern __shared__ float smem[];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  bool did = false;
  if (i < nids) {
    const int64_t e64 = ids ? ids[i] : i;
    if (e64 >= 0 && e64 < n) {
      const int e = (int)e64;
      float w[ZBOT_STATE_WORDS];
      load_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
      EnvState<float> es;
      env_state_unpack(w, es);
#pragma unroll
      for (int k = 0; k < MAX_TERMS; ++k) stat[k] = (k < P.num_terms) ? es.mdp.ep_sums[k] : 0.f;
      stat[S_NUM_RESET] = 1.f;
      stat[S_NUM_TERM_RESET] = (terminated && terminated[e]) ? 1.f : 0.f;
      stat[S_NUM_TO_RESET] = (truncated && truncated[e]) ? 1.f : 0.f;
      if (kSnake) {
        const float speed = es.mdp.speed_limit;       // per-env random constant of the snake task: survives resets
        env_reset_model<ModelSnake>(P, es, dp.feet_pos);
        es.mdp.speed_limit = speed;
      } else {
        env_reset(P, es, dp.feet_pos);
      }
      env_state_pack(es, w);
      store_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
      ep_len_buf[e] = 0;
      did = true;
    }
  }
  stats_block_partial(stat, did, smem, sc);
}

template <bool kSnake>
__global__ void zbot_observe_kernel(const float4* __restrict__ state, float* __restrict__ obs, int n) {
  extern __shared__ float smem[];
  const int e0 = blockIdx.x * blockDim.x;
  const int e = e0 + threadIdx.x;
  float row[ZBOT_NUM_OBS];
#pragma unroll
  for (int i = 0; i < ZBOT_NUM_OBS; ++i) row[i] = 0.f;
  if (e < n) {
    float w[ZBOT_STATE_WORDS];
    load_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
    EnvState<float> es;
    env_state_unpack(w, es);
    if (kSnake) snake_observe(es, row); else env_observe(es, row);
  }
  store_rows_coalesced<ZBOT_NUM_OBS>(obs, row, n, e0, smem);
}

__global__ void zbot_view_kernel(const float4* __restrict__ state, float* pos, float* quat, float* vel, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float w[ZBOT_STATE_WORDS];
  load_words<ZBOT_STATE_WORDS / 4>(state, n, e, w);
  EnvState<float> es;
  env_state_unpack(w, es);
  float p[36], q[48], v[36];
  all_link_kinematics(es.sim, p, q, v);
  if (pos) for (int i = 0; i < 36; ++i) pos[(size_t)e * 36 + i] = p[i];
  if (quat) for (int i = 0; i < 48; ++i) quat[(size_t)e * 48 + i] = q[i];
  if (vel) for (int i = 0; i < 36; ++i) vel[(size_t)e * 36 + i] = v[i];
}

__global__ void zbot_default_pose_kernel(DefaultPose* out, int task) {
  SimState<float> s;
  if (task == ZBOT_TASK_SNAKE_V0) {
    sim_state_default<ModelSnake>(s);
    SnakeKin<float> k;
    snake_kinematics(s, k, false);
    for (int j = 0; j < 2; ++j) {
      for (int i = 0; i < 3; ++i) out->feet_pos[j][i] = 0.f;   // the snake MDP has no feet (mdp_reset zeroes the slots)
      for (int i = 0; i < 4; ++i) out->feet_quat[j][i] = (i == 0) ? 1.f : 0.f;
    }
    for (int i = 0; i < 3; ++i) out->base_pos[i] = k.base_pos[i];
    for (int i = 0; i < 4; ++i) out->base_quat[i] = k.base_quat[i];
    return;
  }
  sim_state_default(s);
  LinkKin<float> k;
  link_kinematics(s, k);
  for (int j = 0; j < 2; ++j) {
    for (int i = 0; i < 3; ++i) out->feet_pos[j][i] = k.feet_pos[j][i];
    for (int i = 0; i < 4; ++i) out->feet_quat[j][i] = k.feet_quat[j][i];
  }
  for (int i = 0; i < 3; ++i) out->base_pos[i] = k.base_pos[i];
  for (int i = 0; i < 4; ++i) out->base_quat[i] = k.base_quat[i];
}

// ---------------------------------------------------------------------------------------------
// MDP-only step: the reference's MDP code on caller-supplied articulation / contact tensors
// ---------------------------------------------------------------------------------------------
struct MdpIn {
  const float *pos, *quat, *vel, *q, *qd, *tau, *hist, *last_air, *origins;
};
constexpr int kHistRow = ZBOT_HISTORY * ZBOT_NUM_LINKS * 3;   // 180 floats = 45 float4 per env
// articulation order: base = 6, feet = {0, 11}; sensor order: feet = {9, 10}, undesired = the other 10
constexpr int kBaseLink = 6, kFoot0Link = 0, kFoot1Link = 11;

__device__ __forceinline__ void mdp_load_links(const MdpIn& in, int e, float* base_pos, float* base_quat, float* base_vel,
                                               float feet_pos[2][3], float feet_quat[2][4], float feet_vel[2][3]) {
  const float* p = in.pos + (size_t)e * 36;
  const float* q = in.quat + (size_t)e * 48;
  const float* v = in.vel + (size_t)e * 36;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    base_pos[i] = __ldg(p + kBaseLink * 3 + i); base_vel[i] = __ldg(v + kBaseLink * 3 + i);
    feet_pos[0][i] = __ldg(p + kFoot0Link * 3 + i); feet_pos[1][i] = __ldg(p + kFoot1Link * 3 + i);
    feet_vel[0][i] = __ldg(v + kFoot0Link * 3 + i); feet_vel[1][i] = __ldg(v + kFoot1Link * 3 + i);
  }
  {
    const float4 b = __ldg(reinterpret_cast<const float4*>(q + kBaseLink * 4));
    const float4 f0 = __ldg(reinterpret_cast<const float4*>(q + kFoot0Link * 4));
    const float4 f1 = __ldg(reinterpret_cast<const float4*>(q + kFoot1Link * 4));
    base_quat[0] = b.x; base_quat[1] = b.y; base_quat[2] = b.z; base_quat[3] = b.w;
    feet_quat[0][0] = f0.x; feet_quat[0][1] = f0.y; feet_quat[0][2] = f0.z; feet_quat[0][3] = f0.w;
    feet_quat[1][0] = f1.x; feet_quat[1][1] = f1.y; feet_quat[1][2] = f1.z; feet_quat[1][3] = f1.w;
  }
}

// ---- 1-D bulk async copy (TMA) global -> shared, completion on an mbarrier ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// envs per CTA of the MDP step kernel (128 threads, the last 16 idle): 79 KB history tile, 2 CTAs / SM, and 65536 / 131072 / 262144
// envs are 1.98 / 3.95 / 7.9 waves of 296 CTAs instead of 1.73 / 3.46 / 6.9 (tile sweep: profiles/r2_notes.md section 9; the
// per-thread row loads lean on the L1, so the tile sizes for which the driver picks a large shared-memory carve-out -- 96, 72, 56,
// 48, 32 -- and an explicit maximum carve-out are 25 % slower)
constexpr int kMdpTile = 112;

// kStep = false: `_get_observations` only (fills the stale cache)
template <bool kStep>
__global__ void __launch_bounds__(128)
zbot_mdp_kernel(const __grid_constant__ Params<float> P, const __grid_constant__ DefaultPose dp, MdpIn in,
                float4* __restrict__ mstate, int64_t* __restrict__ ep_len_buf, const float* __restrict__ actions,
                float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated,
                uint8_t* __restrict__ truncated, int n, StatsCtx sc, int tile) {
  extern __shared__ __align__(128) float smem[];   // [blockDim][180] history tile, reused for obs rows / stats
  pdl_wait();
  // `tile` envs per CTA (<= blockDim, a multiple of 4): chosen by the host so that the grid is a whole number of full waves
  const int e0 = blockIdx.x * tile;
  const int e = e0 + threadIdx.x;
  const bool live = (int)threadIdx.x < tile && e < n;
  const int valid = min(tile, n - e0);
  float stat[kStatUsed];
#pragma unroll
  for (int j = 0; j < kStatUsed; ++j) stat[j] = 0.f;
  float obs_row[ZBOT_NUM_OBS];
#pragma unroll
  for (int i = 0; i < ZBOT_NUM_OBS; ++i) obs_row[i] = 0.f;
  bool did_reset = false;

  FreshInputs<float> f;
  __shared__ uint64_t tile_bar;
  if (kStep) {
    // the block's history tile is one contiguous span of valid*720 bytes: a single 1-D bulk async copy
    // (TMA) brings it to shared memory while the threads issue their own global loads below
    if (threadIdx.x == 0) mbar_init(&tile_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)valid * kHistRow * sizeof(float);
      mbar_expect_tx(&tile_bar, bytes);
      bulk_g2s(smem, in.hist + (size_t)e0 * kHistRow, bytes, &tile_bar);
    }
  }
  if (live) {
    float w[ZBOT_MDP_STATE_WORDS];
    load_words<ZBOT_MDP_STATE_WORDS / 4>(mstate, n, e, w);
    MdpState<float> m;
    mdp_state_unpack(w, M_PDELTA, M_ACT, M_FLAST, M_FDPL, M_FSL, M_HSUM, M_YSUM, M_FFSUM, M_SPEED, M_EPSUM, m);
    StaleCache<float> stale;
    stale_unpack(w, stale);
    float base_pos[3], base_quat[4], base_vel[3], feet_pos[2][3], feet_quat[2][4], feet_vel[2][3];
    mdp_load_links(in, e, base_pos, base_quat, base_vel, feet_pos, feet_quat, feet_vel);
    float q[6], qd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { q[k] = __ldg(in.q + (size_t)e * 6 + k); qd[k] = __ldg(in.qd + (size_t)e * 6 + k); }
    const float ox = __ldg(in.origins + (size_t)e * 3), oy = __ldg(in.origins + (size_t)e * 3 + 1),
                oz = __ldg(in.origins + (size_t)e * 3 + 2);
    if (kStep) {
      float raw[6], new_actions[6], target[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) raw[k] = __ldg(actions + (size_t)e * 6 + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) f.applied_torque[k] = __ldg(in.tau + (size_t)e * 6 + k);
      f.last_air_time[0] = __ldg(in.last_air + (size_t)e * 12 + kFoot0Sensor);
      f.last_air_time[1] = __ldg(in.last_air + (size_t)e * 12 + kFoot1Sensor);
      int64_t ep = ep_len_buf[e] + 1;
      // ---- all global loads are in flight; now consume the history tile the TMA delivered ----
      mbar_wait(&tile_bar, 0);
      {
        // per-thread pass over its own 45 float4 (conflict-free: row stride 45 float4 is odd)
        const float4* row4 = reinterpret_cast<const float4*>(smem) + threadIdx.x * (kHistRow / 4);
        float fz0 = 0.f, fz1 = 0.f, mx2 = 0.f;
#pragma unroll
        for (int t = 0; t < ZBOT_HISTORY; ++t) {
          float h[36];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const float4 v = row4[t * 9 + i];
            h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
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
    }
    // _get_observations: refresh the stale cache from the (possibly reset) articulation view
    stale_from_links(base_pos, base_quat, base_vel, feet_pos, feet_quat, stale);
    mdp_observation(base_quat, q, qd, m.actions, m.speed_limit, obs_row);
    mdp_state_pack(m, w, M_PDELTA, M_ACT, M_FLAST, M_FDPL, M_FSL, M_HSUM, M_YSUM, M_FFSUM, M_SPEED, M_EPSUM);
    stale_pack(stale, w);
    store_words<ZBOT_MDP_STATE_WORDS / 4>(mstate, n, e, w);
  }
  if (kStep) __syncthreads();   // every thread has consumed its history row: the tile is dead, smem is reused
  store_rows_coalesced<ZBOT_NUM_OBS>(obs, obs_row, n, e0, smem, tile);
  if (kStep) {
    __syncthreads();
    stats_block_partial(stat, did_reset, smem, sc);
  }
}

#include "zbot_policy_tc5.cuh"  // the act half on tcgen05.mma / TMEM, the default (uses the mbarrier helpers above)
#include "zbot_mdp_pipe.cuh"    // the same step as a persistent, TMA-fed kernel (the default; ZBOT_MDP_PIPE=0 restores the one above)

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
struct ZbotHandle {
  ZbotCfg cfg;
  Params<float> P;
  DefaultPose dp;
  int device;
  int num_sms;
  bool unroll2;    // chain sweeps unrolled by two (more than one warp per scheduler)
  bool pdl;        // launch the step / statistics kernels with programmatic stream serialization
  bool fused_stats;  // the grid-level statistics pass runs in the last CTA of the producing kernel (StatsCtx::acc): no second launch
  bool pdl_early;  // ... and let the step kernel release its dependents at its start (StatsCtx::pdl_early; ZBOT_PDL_EARLY=0/1)
  bool ctas3;      // 3 CTAs/SM need fewer waves than 2 at this N (the register-budget rule of zbot_create)
  float4* state;
  int64_t* ep_len;
  float* ring;
  int ring_slots;
  float4* mstate;
  int64_t* m_ep_len;
  float* m_ring;
  int m_ring_slots;
  float* partials;
  int max_blocks;
  int64_t launches;
  float inv_episode_s;
  int variant;      // index into kStepVariants
  int force_block;
  unsigned long long* rng_ctr;   // device: stream position of the in-kernel generator (see StatsCtx::rng_ctr)
  DefaultPose* d_dp;             // device scratch of the create-time default-pose FK
  int spread_all_reset;          // zbot_set_all_reset_spread
  int w2_ctas;                   // resident 64-thread CTAs per SM the w2 kernel is compiled for (register budget); ZBOT_W2_CTAS
  bool w2;                       // walking-v2: the two-warps-per-32-envs kernel (zbot_w2_kernel.cuh); ZBOT_W2=0 / a ZBOT_STEP_VARIANT restore the one-thread-per-env kernels
  bool v4_h2;                    // walking-v4: physics phase as packed halves (same rule as the walking-v2 kernel)
  char kernel_name[96];          // zbot_step_kernel_name
  TerrainArgs terrain;           // zbot_bind_terrain (heights == nullptr: flat)
  int mdp_tile;
};

namespace {

// Register-budget / CTA-shape variants of the walking step kernel: (threads per CTA, resident CTAs per SM the
// kernel is compiled for).  The register cap follows from 65536 / (threads * ctas); smaller CTAs give a finer
// occupancy grid (e.g. 5 x 64 threads = 10 warps/SM at the full 197 registers).
typedef void (*StepFn)(Params<float>, DefaultPose, float4*, int64_t*, const float*, float*, float*, uint8_t*, uint8_t*, int,
                       int, int, StatsCtx, ExportPtrs);
struct StepVariant { int threads, ctas; StepFn fn; int envs_per_thread; };
const StepVariant kStepVariants[] = {
    {128, 2, zbot_step_kernel<false, 128, 2>, 1}, {128, 3, zbot_step_kernel<false, 128, 3>, 1},
    {128, 4, zbot_step_kernel<false, 128, 4>, 1},
    {64, 5, zbot_step_kernel_r<200>, 1},  {32, 10, zbot_step_kernel_r<200>, 1}, {32, 11, zbot_step_kernel_r<184>, 1},
    {64, 6, zbot_step_kernel<false, 64, 6>, 1}, {32, 12, zbot_step_kernel<false, 32, 12>, 1},
    {32, 13, zbot_step_kernel_r<152>, 1}, {32, 14, zbot_step_kernel_r<144>, 1}, {64, 7, zbot_step_kernel_r<144>, 1},
    {32, 16, zbot_step_kernel<false, 32, 16>, 1},
    // chain sweeps unrolled by two: "u128x2" / "u128x3"
    {2128, 2, zbot_step_u2_kernel<128, 2>, 1}, {2128, 3, zbot_step_u2_kernel<128, 3>, 1},
    {2064, 7, zbot_step_u2_kernel_r<144>, 1}, {2032, 14, zbot_step_u2_kernel_r<144>, 1},   // "u64x7" / "u32x14": one wave at 65536 envs
    // (nine warps per SM at the full budget -- 3 x 96 or 9 x 32 threads -- do not exist: the register file is per scheduler,
    //  16384 words each, and the ninth warp puts three on one of them: ptxas caps such a shape at 168 registers and spills)
    // EXPERIMENTAL, opt-in (ZBOT_STEP_VARIANT=p128x2): two envs per thread, packed FP32 (zbot_step2_kernel).  31 % fewer
    // warp instructions per env, but 255 registers + spills at 1.7 warps per sub-partition: 84.0 vs 86.4 us at 65536 envs,
    // 59 vs 35 us at 4096 (profiles/r1_notes.md).
    {1128, 2, zbot_step2_kernel<128, 2>, 2},
    // both halves of one env packed in the two FP32 lanes of one thread (zbot_h2.h): "h128x2"
    {3128, 2, zbot_step_h2_kernel<128, 2>, 1},
};
constexpr int kNumStepVariants = (int)(sizeof(kStepVariants) / sizeof(kStepVariants[0]));

// launch with the programmatic-stream-serialization attribute (see pdl_wait)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

StepFn w2_fn(int ctas) {
  switch (ctas) {
    case 3: return zbot_step_w2_kernel<3>;
    case 6: return zbot_step_w2_kernel<6>;
    case 10: return zbot_step_w2_kernel<10>;
    default: return zbot_step_w2_kernel<8>;
  }
}

int pick_block(const ZbotHandle* h, int n) {
  // fill the SMs first: the step is latency/issue bound, not bandwidth bound
  const StepVariant& v = kStepVariants[h->variant];
  int block = v.threads % 1000;
  if (h->force_block == 32 || h->force_block == 64 || h->force_block == 128) block = min(block, h->force_block);
  while (block > 32 && (n + block * v.envs_per_thread - 1) / (block * v.envs_per_thread) < 2 * h->num_sms) block >>= 1;
  return block;
}

int find_variant(int threads, int ctas) {
  for (int v = 0; v < kNumStepVariants; ++v)
    if (kStepVariants[v].threads == threads && kStepVariants[v].ctas == ctas) return v;
  return -1;
}

// Makes the handle's device current for the duration of an entry point and restores the caller's device afterwards
// (a handle used while another device is current would otherwise launch on a foreign device / stream).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <typename H>
void ctx_fuse(StatsCtx& sc, const H* h) {
  if (!h->fused_stats) return;
  sc.acc = h->rng_ctr + 2;
  sc.ticket = reinterpret_cast<unsigned int*>(h->rng_ctr + 1);
}

template <typename H>
void ctx_spread(StatsCtx& sc, const H* h) {
  sc.pdl_early = (h->pdl && h->pdl_early) ? 1 : 0;
  if (!h->spread_all_reset) return;
  sc.spread_ep_len = h->ep_len;
  sc.spread_n = h->cfg.num_envs;
  sc.spread_high = h->cfg.max_episode_length;
  sc.spread_seed = h->cfg.rng_seed ^ 0xA11E5E7ull;
}

int check_slot(int slot, int prev, int slots) {
  if (slot < 0 || slot >= slots) return fail(ZBOT_E_INVALID, "stats_slot out of range%s");
  if (prev >= slots) return fail(ZBOT_E_INVALID, "prev_slot out of range%s");
  return ZBOT_OK;
}

}  // namespace

extern "C" {

int zbot_abi_version(void) { return ZBOT_ABI_VERSION; }
int zbot_cfg_sizeof(void) { return (int)sizeof(ZbotCfg); }
const char* zbot_build_info(void) { return "zbot_b200 sm_100a, nvcc " __DATE__ " " __TIME__; }
const char* zbot_last_error(void) { return g_err; }

int zbot_default_cfg(ZbotCfg* cfg, int32_t num_envs) {
  if (!cfg) return fail(ZBOT_E_INVALID, "cfg is NULL%s");
  cfg_defaults(*cfg, num_envs);
  return ZBOT_OK;
}
int zbot_state_word(const char* f) { return find_word(kStateFields, (int)(sizeof(kStateFields) / sizeof(kStateFields[0])), f); }
int zbot_mdp_state_word(const char* f) { return find_word(kMdpFields, (int)(sizeof(kMdpFields) / sizeof(kMdpFields[0])), f); }

static int create_impl(const ZbotCfg* cfg, int device, ZbotHandle* h);

int zbot_create(const ZbotCfg* cfg, int device, ZbotHandle** out) {
  if (!cfg || !out) return fail(ZBOT_E_INVALID, "cfg/out is NULL%s");
  const char* why = "";
  if (cfg_validate(*cfg, &why) != ZBOT_OK) return fail(ZBOT_E_INVALID, "%s", why);
  int ndev = 0;
  ZB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ZBOT_E_INVALID, "no such CUDA device%s");
  DeviceGuard guard(device);             // the caller's current device is restored on every path
  ZbotHandle* h = new (std::nothrow) ZbotHandle();
  if (!h) return fail(ZBOT_E_INVALID, "out of host memory%s");
  memset(h, 0, sizeof(*h));
  h->device = device;
  const int rc = create_impl(cfg, device, h);
  if (rc != ZBOT_OK) {                   // nothing leaks on an early CUDA error: the handle owns every allocation
    zbot_destroy(h);
    return rc;
  }
  *out = h;
  return ZBOT_OK;
}

static int create_impl(const ZbotCfg* cfg, int device, ZbotHandle* h) {
  h->cfg = *cfg;
  params_from_cfg(*cfg, h->P);
  h->device = device;
  h->inv_episode_s = 1.0f / ((float)cfg->max_episode_length * cfg->sim_dt * (float)cfg->decimation);
  if (cfg->task == ZBOT_TASK_WALKING_V4) h->inv_episode_s = 1.0f;   // v4 logs sum / ACTUAL duration, per env, in the kernel
  cudaDeviceProp prop;
  ZB_CUDA(cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  h->max_blocks = (cfg->num_envs + 31) / 32 + 1;
  ZB_CUDA(cudaMalloc(&h->partials, (size_t)h->max_blocks * kStats * sizeof(float)));
  // [0] stream position, [1] last-CTA ticket (StatsCtx::ticket), [2..33] fixed-point statistics accumulators (StatsCtx::acc)
  ZB_CUDA(cudaMalloc(&h->rng_ctr, (2 + kStats) * sizeof(unsigned long long)));
  ZB_CUDA(cudaMemset(h->rng_ctr, 0, (2 + kStats) * sizeof(unsigned long long)));
  ZB_CUDA(cudaMalloc(&h->d_dp, sizeof(DefaultPose)));          // scratch of the default-pose FK; freed by zbot_destroy
  zbot_default_pose_kernel<<<1, 1>>>(h->d_dp, cfg->task);
  ZB_CUDA(cudaGetLastError());
  ZB_CUDA(cudaMemcpy(&h->dp, h->d_dp, sizeof(DefaultPose), cudaMemcpyDeviceToHost));
  ZB_CUDA(cudaFuncSetAttribute(zbot_mdp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kHistRow * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_mdp_pipe_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem(4, 2)));
  ZB_CUDA(cudaFuncSetAttribute(zbot_mdp_pipe_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem(4, 3)));
  ZB_CUDA(cudaFuncSetAttribute(zbot_mdp_pipe_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem(3, 3)));
  ZB_CUDA(cudaFuncSetAttribute(zbot_mdp_pipe_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem(4, 1)));
  for (int v = 0; v < kNumStepVariants; ++v)
    ZB_CUDA(cudaFuncSetAttribute((const void*)kStepVariants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 128 * kStepRowWords * 4 * kStepVariants[v].envs_per_thread));
  ZB_CUDA(cudaFuncSetAttribute(zbot_step_kernel<true, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_step_u2_export_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_step_h2_export_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_snake_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_snake_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_v4_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_v4_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_snake_step_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_v4_step_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_m_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_m_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_m_step_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_snake_step_kernel<false, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_snake_step_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_v4_step_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_m_step_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_m_step_kernel<true, 2, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_m_step_kernel<true, 1, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_m_step_kernel<false, 2, 3, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_m_step_kernel<false, 2, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_m_step_kernel<false, 1, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_v4_step_kernel<false, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_v4_step_kernel<false, 2, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute((zbot_v4_step_kernel<true, 2, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kStepRowWords * 4));
  ZB_CUDA(cudaFuncSetAttribute(zbot_m_step_kernel<false, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * SCR_STRIDE * 4));
  {
    // register-budget variant of the step kernel = resident 128-thread CTAs per SM it is compiled for.
    // Measured (profiles/r1_notes.md): 2 CTAs/SM (197 regs, no spill) is fastest while the grid fits two
    // rounds of 2 x 148 CTAs; beyond that 3 CTAs/SM (168 regs) wins on throughput (+18 % at 131072 envs).
    const char* sv = getenv("ZBOT_STEP_VARIANT");      // tuning override, "<threads>x<ctas>", e.g. 64x5
    int vt = 0, vc = 0;
    h->variant = -1;
    if (sv && sv[0] == 'p' && sscanf(sv + 1, "%dx%d", &vt, &vc) == 2) h->variant = find_variant(1000 + vt, vc);
    else if (sv && sv[0] == 'h' && sscanf(sv + 1, "%dx%d", &vt, &vc) == 2) h->variant = find_variant(3000 + vt, vc);
    else if (sv && sv[0] == 'u' && sscanf(sv + 1, "%dx%d", &vt, &vc) == 2) h->variant = find_variant(2000 + vt, vc);
    else if (sv && sscanf(sv, "%dx%d", &vt, &vc) == 2) h->variant = find_variant(vt, vc);
    // default register budget: by wave count, below; the sweeps unrolled by two as soon as a
    // scheduler holds more than one warp (N > 148 * 4 * 32), measured -2.3 us at 32768, -3.3 us at 65536 envs
    // programmatic dependent launch: the statistics kernel and the next step kernel are scheduled while their predecessor
    // drains (34.9 -> 30.9 us per step at 4096 envs, 80.0 -> 76.7 at 65536); ZBOT_PDL=0 restores plain launches
    { const char* sp = getenv("ZBOT_PDL"); h->pdl = sp ? (atoi(sp) != 0) : true; }
    // statistics fused into the producing kernel (fixed-point accumulators + last-CTA pass); ZBOT_FUSED_STATS=0 restores the
    // separate one-block zbot_stats_finalize_kernel behind every step
    { const char* sp = getenv("ZBOT_FUSED_STATS"); h->fused_stats = sp ? (atoi(sp) != 0) : true; }
    { const char* sp = getenv("ZBOT_PDL_EARLY"); h->pdl_early = sp ? (atoi(sp) != 0) : false; }
    h->unroll2 = cfg->num_envs > 4 * 32 * h->num_sms;
    if (const char* su = getenv("ZBOT_SWEEP_UNROLL")) h->unroll2 = (atoi(su) == 2);   // test / tuning override, all tasks
    if (h->variant < 0) {
      // 3 CTAs/SM (168 registers, 8 B spill) exactly when it needs FEWER waves than 2 CTAs/SM (208-218 registers, no spill):
      // 49152 envs = one wave of three (59.4 us) against two waves of two (83.9 us); 65536 envs = two waves either way
      // (80.0 us with two, 90.4 us with three); 131072 envs = three waves against four (143 us vs 154 us)
      const int cap2 = 2 * h->num_sms * 128, cap3 = 3 * h->num_sms * 128;
      const int w2 = (cfg->num_envs + cap2 - 1) / cap2, w3 = (cfg->num_envs + cap3 - 1) / cap3;
      h->variant = find_variant((h->unroll2 ? 2000 : 0) + 128, (w3 < w2) ? 3 : 2);
      // Round 2: both halves of the chain in the two FP32 lanes of one thread (zbot_h2.h, "h128x2": 250 registers, no spill,
      // 27 % fewer warp instructions).  Back to back on B200 (tools/time_h2.py): 31.6 vs 32.2 us at 16384 envs, 39.9 vs 41.4
      // at 32768, 73.2 vs 77.8 at 65536, 125.2 vs 139.6 at 131072, 241.8 vs 248.0 at 262144.  The one window it loses is
      // where three CTAs/SM of the 168-register kernel hold ALL envs in one wave and two CTAs/SM do not
      // (37888 < N <= 56832: 55.2 vs 59.2 us at 49152).  ZBOT_H2=0 restores the one-chain kernels.
      const char* sh = getenv("ZBOT_H2");
      const bool h2_on = sh ? (atoi(sh) != 0) : true;
      const bool one_wave_of_three = cfg->num_envs > cap2 && cfg->num_envs <= cap3;
      if (h2_on && cfg->task == ZBOT_TASK_WALKING_V2 && !one_wave_of_three) h->variant = find_variant(3128, 2);
    }
    {
      const int cap2 = 2 * h->num_sms * 128, cap3 = 3 * h->num_sms * 128;
      const char* sh = getenv("ZBOT_H2");
      h->v4_h2 = (sh ? (atoi(sh) != 0) : true) && cfg->task == ZBOT_TASK_WALKING_V4 && h->unroll2 &&
                 !(cfg->num_envs > cap2 && cfg->num_envs <= cap3);
      h->ctas3 = (cfg->num_envs + cap3 - 1) / cap3 < (cfg->num_envs + cap2 - 1) / cap2;
      if (const char* sc3 = getenv("ZBOT_CTAS3")) h->ctas3 = (atoi(sc3) != 0);   // tuning override (snake / v4 / manager kernels)
    }
    const char* mt = getenv("ZBOT_MDP_TILE");          // envs (= threads) per CTA of the MDP-only step kernel
    h->mdp_tile = mt ? atoi(mt) : kMdpTile;
    if (h->mdp_tile < 32 || h->mdp_tile > 128 || (h->mdp_tile & 3)) h->mdp_tile = kMdpTile;
    const char* bs = getenv("ZBOT_STEP_BLOCK");
    h->force_block = bs ? atoi(bs) : 0;
    // Two warps per 32 envs (zbot_w2_kernel.cuh) while an SM holds at most two warp pairs, i.e. while the step is bound by
    // the dependent-issue latency of one env's chain: 24.8 vs 31.2 us per step at 4096 envs, 25.4 vs 31.7 at 8192.  Beyond
    // that the one-thread-per-env kernels win on instruction count (the split costs ~27 % more warp instructions: loop /
    // select overhead of the side-generic sweep, the redundant 6x6 solve, the exchanges): 102 vs 77 us at 65536 envs
    // (profiles/r2_notes.md).  ZBOT_W2=0 / 1 forces the choice.
    h->w2 = (cfg->task == ZBOT_TASK_WALKING_V2) && !sv && cfg->num_envs <= 2 * 32 * h->num_sms;
    if (const char* sw = getenv("ZBOT_W2")) h->w2 = (cfg->task == ZBOT_TASK_WALKING_V2) && !sv && (atoi(sw) != 0);
    h->w2_ctas = 3;      // register budget: uncapped (<= 2 pairs per SM where this kernel is the default); 6 / 8 / 10 = tuning variants
    if (const char* sc2 = getenv("ZBOT_W2_CTAS")) { const int c = atoi(sc2); if (c == 3 || c == 6 || c == 8 || c == 10) h->w2_ctas = c; }
  }
  {
    const StepVariant& v = kStepVariants[h->variant];
    const int u = h->unroll2 ? 2 : 1, c = h->ctas3 ? 3 : 2;
    switch (cfg->task) {
      case ZBOT_TASK_SNAKE_V0: snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_snake_step_kernel<false,%d,%d>", u, u == 2 ? c : 2); break;
      case ZBOT_TASK_WALKING_V4:
        if (h->v4_h2) snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_v4_step_kernel<false,2,2,h2>");
        else snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_v4_step_kernel<false,%d,%d>", u, u == 2 ? c : 2);
        break;
      case ZBOT_TASK_WALKING_M: snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_m_step_kernel<false,%d,%d>", u, u == 2 ? c : 2); break;
      default:
        if (h->w2) snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_step_w2_kernel<%d>", h->w2_ctas);
        else if (v.threads >= 3000) snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_step_h2_kernel<%d,%d>", v.threads % 1000, v.ctas);
        else if (v.threads >= 2000) snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_step_u2_kernel<%d,%d>", v.threads % 1000, v.ctas);
        else if (v.threads >= 1000) snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_step2_kernel<%d,%d>", v.threads % 1000, v.ctas);
        else snprintf(h->kernel_name, sizeof(h->kernel_name), "zbot_step_kernel<false,%d,%d>", v.threads, v.ctas);
    }
  }
  return ZBOT_OK;
}

int zbot_destroy(ZbotHandle* h) {
  if (!h) return ZBOT_OK;
  {
    DeviceGuard guard(h->device);
    cudaFree(h->partials);
    cudaFree(h->rng_ctr);
    cudaFree(h->d_dp);
  }
  delete h;
  return ZBOT_OK;
}

int zbot_bind(ZbotHandle* h, float* state, int64_t* episode_length, float* stats_ring, int32_t stats_slots) {
  if (!h || !state || !episode_length || !stats_ring || stats_slots < 1) return fail(ZBOT_E_INVALID, "zbot_bind: NULL buffer%s");
  if (((uintptr_t)state & 15) != 0) return fail(ZBOT_E_INVALID, "state must be 16-byte aligned%s");
  h->state = reinterpret_cast<float4*>(state);
  h->ep_len = episode_length;
  h->ring = stats_ring;
  h->ring_slots = stats_slots;
  return ZBOT_OK;
}

static int step_impl(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated,
                     int32_t slot, int32_t prev, const ZbotExport* ex, void* stream, float* snake_export = nullptr) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (!actions || !obs || !rew || !terminated || !truncated) return fail(ZBOT_E_INVALID, "zbot_step: NULL buffer%s");
  if (((uintptr_t)actions & 7) != 0) return fail(ZBOT_E_INVALID, "actions must be 8-byte aligned%s");
  if (int rc = check_slot(slot, prev, h->ring_slots)) return rc;
  const int n = h->cfg.num_envs;
  const bool walk = (h->cfg.task == ZBOT_TASK_WALKING_V2) && !ex;
  const bool walk_any = (h->cfg.task == ZBOT_TASK_WALKING_V2);
  const int ept = walk ? kStepVariants[h->variant].envs_per_thread : 1;
  int block = pick_block(h, n);
  if (!walk) { block = 128; while (block > 32 && (n + block - 1) / block < 2 * h->num_sms) block >>= 1; }
  const int grid = (walk_any && h->w2) ? (n + 31) / 32 : (n + block * ept - 1) / (block * ept);
  const size_t smem = (size_t)block * kStepRowWords * sizeof(float) * ept;   // >= obs rows (23/thread) and stats (704 floats)
  StatsCtx sc{h->partials, h->ring, slot, prev, h->inv_episode_s, 0, h->rng_ctr, 0};
  ctx_fuse(sc, h);
  ctx_spread(sc, h);
  cudaStream_t s = (cudaStream_t)stream;
  ExportPtrs xp{};
  if (h->cfg.task == ZBOT_TASK_WALKING_V4) return fail(ZBOT_E_INVALID, "task zbot-6b-walking-v4 steps through zbot_v4_step%s");
  if (h->cfg.task == ZBOT_TASK_WALKING_M) return fail(ZBOT_E_INVALID, "task zbot-6b-walking-m-v0 steps through zbot_m_step%s");
  if (h->cfg.task == ZBOT_TASK_SNAKE_V0) {
    if (ex) return fail(ZBOT_E_INVALID, "zbot_step_export is a walking-task hook; use zbot_snake_step_export%s");
    if (snake_export && h->unroll2)   // export flavour of the SAME sweep unroll the product launch uses at this N
      ZB_CUDA_LAUNCH((zbot_snake_step_kernel<true, 2>), h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                         truncated, n, sc, snake_export);
    else if (snake_export)
      ZB_CUDA_LAUNCH(zbot_snake_step_kernel<true>, h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                         truncated, n, sc, snake_export);
    else if (h->unroll2 && h->ctas3)   // more than one warp per scheduler: sweeps unrolled by two (see zbot_create)
      ZB_CUDA_LAUNCH((zbot_snake_step_kernel<false, 2, 3>), h->P, h->dp, h->state, h->ep_len, actions, obs, rew,
                                                             terminated, truncated, n, sc, nullptr);
    else if (h->unroll2)   // more than one warp per scheduler: sweeps unrolled by two (see zbot_create)
      ZB_CUDA_LAUNCH((zbot_snake_step_kernel<false, 2>), h->P, h->dp, h->state, h->ep_len, actions, obs, rew,
                                                             terminated, truncated, n, sc, nullptr);
    else
      ZB_CUDA_LAUNCH(zbot_snake_step_kernel<false>, h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                          truncated, n, sc, nullptr);
  } else if (snake_export) {
    return fail(ZBOT_E_INVALID, "zbot_snake_step_export needs a handle created with task = ZBOT_TASK_SNAKE_V0%s");
  } else if (ex) {
    xp = ExportPtrs{ex->body_link_pos_w0, ex->body_link_quat_w0, ex->body_com_lin_vel_w0, ex->body_link_pos_w1,
                    ex->body_link_quat_w1, ex->body_com_lin_vel_w1, ex->joint_pos1, ex->joint_vel1, ex->applied_torque1,
                    ex->net_forces_w_history1, ex->last_air_time1, ex->current_contact_time1};
    const float* const* pp = reinterpret_cast<const float* const*>(&xp);
    for (int i = 0; i < 12; ++i)
      if (!pp[i]) return fail(ZBOT_E_INVALID, "zbot_step_export: NULL export buffer%s");
    // the export flavour of the instantiation `zbot_step` launches for this handle (same phased body, same sweep unroll)
    if (h->w2) {
      // The two-warp kernel exports from the PRODUCT instantiation itself (run-time hook: plain stores behind uniform
      // branches), so the launch the oracle tests read is bit for bit the launch that is benchmarked.  Around it: the
      // start-of-step 12-link view and history slot 4 from the state before the step, the end-of-physics view from the
      // exported state after it.
      const int vb = 64, vg = (n + vb - 1) / vb;
      zbot_view_kernel<<<vg, vb, 0, s>>>(h->state, xp.pos0, xp.quat0, xp.vel0, n);
      zbot_w2_export_pre_kernel<<<vg, vb, 0, s>>>(h->state, xp, n);
      w2_fn(h->w2_ctas)<<<(n + 31) / 32, 64, kW2Smem, s>>>(h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                          truncated, n, 0, n, sc, xp);
      zbot_w2_export_view_kernel<<<vg, vb, 0, s>>>(xp, n);
    } else if (kStepVariants[h->variant].threads >= 3000)
      zbot_step_h2_export_kernel<<<grid, block, smem, s>>>(h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                         truncated, n, 0, n, sc, xp);
    else if (h->unroll2)
      zbot_step_u2_export_kernel<<<grid, block, smem, s>>>(h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                         truncated, n, 0, n, sc, xp);
    else
      zbot_step_kernel<true, 128, 1><<<grid, block, smem, s>>>(h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                         truncated, n, 0, n, sc, xp);
  } else if (h->w2) {
    const int g2 = (n + 31) / 32;
    ZB_CUDA(launch_pdl(w2_fn(h->w2_ctas), dim3(g2), dim3(64), kW2Smem, s, h->pdl, h->P, h->dp, h->state, h->ep_len,
                       actions, obs, rew, terminated, truncated, n, 0, n, sc, xp));
    h->launches += 1;
    if (!sc.acc) {
      ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, h->pdl, sc, (unsigned int)g2));
      h->launches += 1;
    }
    return ZBOT_OK;
  } else if (h->pdl) {
    ZB_CUDA(launch_pdl(kStepVariants[h->variant].fn, dim3(grid), dim3(block), smem, s, true, h->P, h->dp, h->state, h->ep_len,
                       actions, obs, rew, terminated, truncated, n, 0, n, sc, xp));
    h->launches += 1;
    if (!sc.acc) {
      ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, true, sc, (unsigned int)grid));
      h->launches += 1;
    }
    return ZBOT_OK;
  } else {
    kStepVariants[h->variant].fn<<<grid, block, smem, s>>>(h->P, h->dp, h->state, h->ep_len, actions, obs, rew, terminated,
                                                        truncated, n, 0, n, sc, xp);
  }
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  if (!sc.acc) {
    ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, h->pdl, sc, (unsigned int)grid));
    h->launches += 1;
  }
  return ZBOT_OK;
}

int zbot_step(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated,
              int32_t stats_slot, int32_t prev_slot, void* stream) {
  return step_impl(h, actions, obs, rew, terminated, truncated, stats_slot, prev_slot, nullptr, stream);
}

int zbot_step_export(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated, uint8_t* truncated,
                     int32_t stats_slot, int32_t prev_slot, const ZbotExport* ex, void* stream) {
  if (!ex) return fail(ZBOT_E_INVALID, "zbot_step_export: ex is NULL%s");
  return step_impl(h, actions, obs, rew, terminated, truncated, stats_slot, prev_slot, ex, stream);
}

// Whole control step for a HOST-resident caller (DESIGN.md §4 "end to end").  ONE launch of the fused kernel with
// pinned-host pointers: each thread reads its 24 B of actions straight from host memory (once -- they are parked in
// shared memory for the MDP phase) and each CTA writes its envs' result rows [obs 23 | reward | flags] as one
// contiguous run of float4 stores straight to host memory, so both directions cross PCIe inside the kernel,
// overlapped with the other CTAs' compute, with no DMA operation at all.  Measured against the alternatives on
// B200 / PCIe 5 x16 (tools/probe_host_step.py, 65536 envs): H2D + kernel + D2H back to back 297 us, four env
// ranges pipelined on streams with one DMA each way per range 219 us (a DMA op costs ~5 us however small), this
// 196 us.  Returns when the result is complete in host memory.
int zbot_step_host(ZbotHandle* h, const float* host_actions, float* host_rows, int32_t stats_slot, int32_t prev_slot,
                   void* stream) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (h->cfg.task != ZBOT_TASK_WALKING_V2) return fail(ZBOT_E_INVALID, "zbot_step_host: walking task only%s");
  if (!host_actions || !host_rows) return fail(ZBOT_E_INVALID, "zbot_step_host: NULL buffer%s");
  if (((uintptr_t)host_actions & 7) != 0 || ((uintptr_t)host_rows & 15) != 0)
    return fail(ZBOT_E_INVALID, "zbot_step_host: host_actions must be 8-byte and host_rows 16-byte aligned%s");
  if (int rc = check_slot(stats_slot, prev_slot, h->ring_slots)) return rc;
  {
    // the buffers must be device-accessible (pinned + mapped: cudaHostAlloc / cudaHostRegister / torch pin_memory)
    const void *da = nullptr, *dr = nullptr;
    if (cudaHostGetDevicePointer((void**)&da, (void*)host_actions, 0) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&dr, (void*)host_rows, 0) != cudaSuccess) {
      cudaGetLastError();
      return fail(ZBOT_E_INVALID, "zbot_step_host: buffers must be pinned (page-locked, mapped) host memory%s");
    }
    host_actions = static_cast<const float*>(da);
    host_rows = static_cast<float*>(const_cast<void*>(dr));
  }
  const int n = h->cfg.num_envs;
  // the SAME kernel instantiation and CTA shape `zbot_step` uses for this handle (bit-identical results)
  const int vi = (kStepVariants[h->variant].envs_per_thread == 1) ? h->variant : find_variant(128, 2);
  int block = kStepVariants[vi].threads % 1000;
  while (block > 32 && (n + block - 1) / block < 2 * h->num_sms) block >>= 1;
  const int grid = (n + block - 1) / block;
  const size_t smem = (size_t)block * kStepRowWords * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  StatsCtx sc{h->partials, h->ring, stats_slot, prev_slot, h->inv_episode_s, 0, h->rng_ctr, 1};
  ctx_fuse(sc, h);
  ctx_spread(sc, h);
  ExportPtrs xp{};
  const int grid_used = h->w2 ? (n + 31) / 32 : grid;
  if (h->w2)
    w2_fn(h->w2_ctas)<<<grid_used, 64, kW2Smem, s>>>(h->P, h->dp, h->state, h->ep_len, host_actions, host_rows, nullptr,
                                                    nullptr, nullptr, n, 0, n, sc, xp);
  else
    kStepVariants[vi].fn<<<grid, block, smem, s>>>(h->P, h->dp, h->state, h->ep_len, host_actions, host_rows, nullptr, nullptr,
                                                   nullptr, n, 0, n, sc, xp);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  if (!sc.acc) {
    ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, h->pdl, sc, (unsigned int)grid_used));
    h->launches += 1;
  }
  ZB_CUDA(cudaStreamSynchronize(s));
  return ZBOT_OK;
}

static int v4_step_impl(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                        uint8_t* truncated, int32_t slot, int32_t prev, float* export_buf, void* stream) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (h->cfg.task != ZBOT_TASK_WALKING_V4) return fail(ZBOT_E_INVALID, "zbot_v4_step needs task = ZBOT_TASK_WALKING_V4%s");
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (!actions || !obs || !rew || !terminated || !truncated) return fail(ZBOT_E_INVALID, "zbot_v4_step: NULL buffer%s");
  if (((uintptr_t)actions & 7) != 0 || ((uintptr_t)rand & 7) != 0)
    return fail(ZBOT_E_INVALID, "actions / rand must be 8-byte aligned%s");
  if (int rc = check_slot(slot, prev, h->ring_slots)) return rc;
  const int n = h->cfg.num_envs;
  int block = 128;
  while (block > 32 && (n + block - 1) / block < 2 * h->num_sms) block >>= 1;
  const int grid = (n + block - 1) / block;
  const size_t smem = (size_t)block * kStepRowWords * sizeof(float);
  StatsCtx sc{h->partials, h->ring, slot, prev, h->inv_episode_s, 0, h->rng_ctr, 0};
  ctx_fuse(sc, h);
  ctx_spread(sc, h);
  cudaStream_t s = (cudaStream_t)stream;
  if (export_buf && h->v4_h2)
    ZB_CUDA_LAUNCH((zbot_v4_step_kernel<true, 2, 2, true>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew,
                                                    terminated, truncated, n, sc, export_buf);
  else if (h->v4_h2)
    ZB_CUDA_LAUNCH((zbot_v4_step_kernel<false, 2, 2, true>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew,
                                                    terminated, truncated, n, sc, nullptr);
  else if (export_buf && h->unroll2)
    ZB_CUDA_LAUNCH((zbot_v4_step_kernel<true, 2>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew,
                                                    terminated, truncated, n, sc, export_buf);
  else if (export_buf)
    ZB_CUDA_LAUNCH(zbot_v4_step_kernel<true>, h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew,
                                                    terminated, truncated, n, sc, export_buf);
  else if (h->unroll2 && h->ctas3)
    ZB_CUDA_LAUNCH((zbot_v4_step_kernel<false, 2, 3>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs,
                                                        rew, terminated, truncated, n, sc, nullptr);
  else if (h->unroll2)
    ZB_CUDA_LAUNCH((zbot_v4_step_kernel<false, 2>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs,
                                                        rew, terminated, truncated, n, sc, nullptr);
  else
    ZB_CUDA_LAUNCH(zbot_v4_step_kernel<false>, h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew,
                                                     terminated, truncated, n, sc, nullptr);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  if (!sc.acc) {
    ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, h->pdl, sc, (unsigned int)grid));
    h->launches += 1;
  }
  return ZBOT_OK;
}

int zbot_v4_step(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                 uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream) {
  return v4_step_impl(h, actions, rand, obs, rew, terminated, truncated, stats_slot, prev_slot, nullptr, stream);
}

int zbot_v4_step_export(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                        uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, float* export_buf, void* stream) {
  if (!export_buf) return fail(ZBOT_E_INVALID, "zbot_v4_step_export: export_buf is NULL%s");
  return v4_step_impl(h, actions, rand, obs, rew, terminated, truncated, stats_slot, prev_slot, export_buf, stream);
}

static int m_step_impl(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                       uint8_t* truncated, int32_t slot, int32_t prev, float* export_buf, void* stream) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (h->cfg.task != ZBOT_TASK_WALKING_M) return fail(ZBOT_E_INVALID, "zbot_m_step needs task = ZBOT_TASK_WALKING_M%s");
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (!actions || !obs || !rew || !terminated || !truncated) return fail(ZBOT_E_INVALID, "zbot_m_step: NULL buffer%s");
  if (((uintptr_t)actions & 7) != 0) return fail(ZBOT_E_INVALID, "actions must be 8-byte aligned%s");
  if (int rc = check_slot(slot, prev, h->ring_slots)) return rc;
  const int n = h->cfg.num_envs;
  int block = 128;
  while (block > 32 && (n + block - 1) / block < 2 * h->num_sms) block >>= 1;
  const int grid = (n + block - 1) / block;
  const size_t smem = (size_t)block * SCR_STRIDE * sizeof(float);
  StatsCtx sc{h->partials, h->ring, slot, prev, h->inv_episode_s, 0, h->rng_ctr, 0,
              (h->cfg.num_terms <= MAX_TERMS - 3) ? 2 : 0};
  ctx_fuse(sc, h);
  ctx_spread(sc, h);
  sc.norm_word22 = 1;
  cudaStream_t s = (cudaStream_t)stream;
#define ZB_M_LAUNCH(...) ZB_CUDA_LAUNCH((zbot_m_step_kernel<__VA_ARGS__>), h->P, h->state, h->ep_len, actions, rand, h->cfg.rng_seed, obs, rew, \
                                      terminated, truncated, n, sc, export_buf, h->terrain)
  const bool rough = h->terrain.heights != nullptr;
  if (rough) {     // height-field ground + terrain curriculum (zbot_bind_terrain)
    if (export_buf && h->unroll2) ZB_M_LAUNCH(true, 2, 2, true);
    else if (export_buf) ZB_M_LAUNCH(true, 1, 2, true);
    else if (h->unroll2 && h->ctas3) ZB_M_LAUNCH(false, 2, 3, true);
    else if (h->unroll2) ZB_M_LAUNCH(false, 2, 2, true);
    else ZB_M_LAUNCH(false, 1, 2, true);
  } else {
    if (export_buf && h->unroll2) ZB_M_LAUNCH(true, 2);
    else if (export_buf) ZB_M_LAUNCH(true);
    else if (h->unroll2 && h->ctas3) ZB_M_LAUNCH(false, 2, 3);
    else if (h->unroll2) ZB_M_LAUNCH(false, 2);
    else ZB_M_LAUNCH(false);
  }
#undef ZB_M_LAUNCH
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  if (!sc.acc) {
    ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, s, h->pdl, sc, (unsigned int)grid));
    h->launches += 1;
  }
  return ZBOT_OK;
}

int zbot_m_step(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream) {
  return m_step_impl(h, actions, rand, obs, rew, terminated, truncated, stats_slot, prev_slot, nullptr, stream);
}

int zbot_m_step_export(ZbotHandle* h, const float* actions, const float* rand, float* obs, float* rew, uint8_t* terminated,
                       uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, float* export_buf, void* stream) {
  if (!export_buf) return fail(ZBOT_E_INVALID, "zbot_m_step_export: export_buf is NULL%s");
  return m_step_impl(h, actions, rand, obs, rew, terminated, truncated, stats_slot, prev_slot, export_buf, stream);
}

int zbot_bind_terrain(ZbotHandle* h, const float* heights, int32_t nx, int32_t ny, float x0, float y0, float cell,
                      const float* tile_origins, int32_t rows, int32_t cols, float tile_size, float* env_origins,
                      int32_t curriculum) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  if (h->cfg.task != ZBOT_TASK_WALKING_M) return fail(ZBOT_E_INVALID, "zbot_bind_terrain: manager task only%s");
  if (!heights) { memset(&h->terrain, 0, sizeof(h->terrain)); return ZBOT_OK; }     // back to the flat ground
  if (!tile_origins || !env_origins || nx < 2 || ny < 2 || rows < 1 || cols < 1 || !(cell > 0.f))
    return fail(ZBOT_E_INVALID, "zbot_bind_terrain: bad argument%s");
  if (((uintptr_t)env_origins & 15) != 0) return fail(ZBOT_E_INVALID, "env_origins must be 16-byte aligned ([N][4] floats)%s");
  h->terrain = TerrainArgs{heights, nx, ny, x0, y0, 1.0f / cell, reinterpret_cast<float4*>(env_origins), tile_origins, rows, cols,
                           tile_size, (float)h->cfg.max_episode_length * h->cfg.sim_dt * (float)h->cfg.decimation, curriculum ? 1 : 0};
  return ZBOT_OK;
}

int zbot_set_all_reset_spread(ZbotHandle* h, int32_t enable) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  h->spread_all_reset = enable ? 1 : 0;
  return ZBOT_OK;
}

int zbot_update_cfg(ZbotHandle* h, const ZbotCfg* cfg) {
  if (!h || !cfg) return fail(ZBOT_E_INVALID, "zbot_update_cfg: NULL argument%s");
  const char* why = "";
  if (cfg_validate(*cfg, &why) != ZBOT_OK) return fail(ZBOT_E_INVALID, "%s", why);
  if (cfg->num_envs != h->cfg.num_envs || cfg->task != h->cfg.task || cfg->sim_dt != h->cfg.sim_dt ||
      cfg->max_episode_length != h->cfg.max_episode_length)
    return fail(ZBOT_E_INVALID, "zbot_update_cfg: num_envs / task / sim_dt / max_episode_length must not change%s");
  h->cfg = *cfg;
  params_from_cfg(*cfg, h->P);
  return ZBOT_OK;
}

int zbot_snake_step_export(ZbotHandle* h, const float* actions, float* obs, float* rew, uint8_t* terminated,
                           uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, float* export41, void* stream) {
  if (!export41) return fail(ZBOT_E_INVALID, "zbot_snake_step_export: export41 is NULL%s");
  return step_impl(h, actions, obs, rew, terminated, truncated, stats_slot, prev_slot, nullptr, stream, export41);
}

int zbot_reset_idx(ZbotHandle* h, const int64_t* env_ids, int64_t nids, const uint8_t* terminated, const uint8_t* truncated,
                   int32_t stats_slot, void* stream) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (int rc = check_slot(stats_slot, -1, h->ring_slots)) return rc;
  if (h->cfg.task == ZBOT_TASK_WALKING_V4 || h->cfg.task == ZBOT_TASK_WALKING_M)
    return fail(ZBOT_E_INVALID, "zbot_reset_idx: the v4 / manager tasks' randomised reset is written by the caller (state words)%s");
  const int n = h->cfg.num_envs;
  if (!env_ids || nids < 0) { env_ids = nullptr; nids = n; }
  if (nids == 0) return ZBOT_OK;
  if (nids > n) return fail(ZBOT_E_INVALID, "more env ids than envs%s");
  const int block = 64;
  const int grid = (int)((nids + block - 1) / block);
  StatsCtx sc{h->partials, h->ring, stats_slot, -1, h->inv_episode_s, 0, h->rng_ctr, 0};
  ctx_fuse(sc, h);
  if (h->cfg.task == ZBOT_TASK_SNAKE_V0)
    zbot_reset_kernel<true><<<grid, block, 32 * kStatUsed * sizeof(float), (cudaStream_t)stream>>>(
        h->P, h->dp, h->state, h->ep_len, env_ids, nids, terminated, truncated, n, sc);
  else
    zbot_reset_kernel<false><<<grid, block, 32 * kStatUsed * sizeof(float), (cudaStream_t)stream>>>(
        h->P, h->dp, h->state, h->ep_len, env_ids, nids, terminated, truncated, n, sc);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  if (!sc.acc) {
    zbot_stats_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sc, (unsigned int)grid);
    ZB_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  return ZBOT_OK;
}

int zbot_observe(ZbotHandle* h, float* obs, void* stream) {
  if (!h || !obs) return fail(ZBOT_E_INVALID, "zbot_observe: NULL argument%s");
  DeviceGuard guard(h->device);
  if (h->cfg.task == ZBOT_TASK_WALKING_V4 || h->cfg.task == ZBOT_TASK_WALKING_M)
    return fail(ZBOT_E_INVALID, "zbot_observe: v2 / snake observation layout only%s");
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  const int n = h->cfg.num_envs, block = 64, grid = (n + block - 1) / block;
  if (h->cfg.task == ZBOT_TASK_SNAKE_V0)
    zbot_observe_kernel<true><<<grid, block, block * ZBOT_NUM_OBS * sizeof(float), (cudaStream_t)stream>>>(h->state, obs, n);
  else
    zbot_observe_kernel<false><<<grid, block, block * ZBOT_NUM_OBS * sizeof(float), (cudaStream_t)stream>>>(h->state, obs, n);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  return ZBOT_OK;
}

int zbot_articulation_view(ZbotHandle* h, float* pos, float* quat, float* vel, void* stream) {
  if (!h) return fail(ZBOT_E_INVALID, "handle is NULL%s");
  DeviceGuard guard(h->device);
  if (!h->state) return fail(ZBOT_E_UNBOUND, "zbot_bind has not been called%s");
  if (h->cfg.task == ZBOT_TASK_SNAKE_V0) return fail(ZBOT_E_INVALID, "zbot_articulation_view: walking tasks only%s");
  const int n = h->cfg.num_envs, block = 64, grid = (n + block - 1) / block;
  zbot_view_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(h->state, pos, quat, vel, n);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  return ZBOT_OK;
}

int zbot_mdp_bind(ZbotHandle* h, float* mdp_state, int64_t* episode_length, float* stats_ring, int32_t stats_slots) {
  if (!h || !mdp_state || !episode_length || !stats_ring || stats_slots < 1) return fail(ZBOT_E_INVALID, "zbot_mdp_bind: NULL buffer%s");
  if (((uintptr_t)mdp_state & 15) != 0) return fail(ZBOT_E_INVALID, "mdp_state must be 16-byte aligned%s");
  h->mstate = reinterpret_cast<float4*>(mdp_state);
  h->m_ep_len = episode_length;
  h->m_ring = stats_ring;
  h->m_ring_slots = stats_slots;
  return ZBOT_OK;
}

static int mdp_check(const ZbotHandle* h, const ZbotMdpInputs* in, bool step) {
  if (!h || !in) return fail(ZBOT_E_INVALID, "zbot_mdp: NULL argument%s");
  if (h->cfg.task != ZBOT_TASK_WALKING_V2) return fail(ZBOT_E_INVALID, "zbot_mdp_*: walking task only%s");
  if (!h->mstate) return fail(ZBOT_E_UNBOUND, "zbot_mdp_bind has not been called%s");
  if (!in->body_link_pos_w || !in->body_link_quat_w || !in->body_com_lin_vel_w || !in->joint_pos || !in->joint_vel ||
      !in->env_origins)
    return fail(ZBOT_E_INVALID, "zbot_mdp: NULL input tensor%s");
  if (step && (!in->applied_torque || !in->net_forces_w_history || !in->last_air_time))
    return fail(ZBOT_E_INVALID, "zbot_mdp_step: NULL input tensor%s");
  if (((uintptr_t)in->body_link_quat_w & 15) != 0 || (step && ((uintptr_t)in->net_forces_w_history & 15) != 0))
    return fail(ZBOT_E_INVALID, "quat / history tensors must be 16-byte aligned%s");
  return ZBOT_OK;
}

int zbot_mdp_observe(ZbotHandle* h, const ZbotMdpInputs* in, float* obs, void* stream) {
  if (int rc = mdp_check(h, in, false)) return rc;
  DeviceGuard guard(h->device);
  if (!obs) return fail(ZBOT_E_INVALID, "obs is NULL%s");
  const int n = h->cfg.num_envs, block = 128, grid = (n + block - 1) / block;
  MdpIn mi{in->body_link_pos_w, in->body_link_quat_w, in->body_com_lin_vel_w, in->joint_pos, in->joint_vel,
           in->applied_torque, in->net_forces_w_history, in->last_air_time, in->env_origins};
  StatsCtx sc{h->partials, h->m_ring, 0, -1, h->inv_episode_s, 0, h->rng_ctr, 0};
  zbot_mdp_kernel<false><<<grid, block, block * ZBOT_NUM_OBS * sizeof(float), (cudaStream_t)stream>>>(
      h->P, h->dp, mi, h->mstate, h->m_ep_len, nullptr, obs, nullptr, nullptr, nullptr, n, sc, block);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  return ZBOT_OK;
}

int zbot_mdp_step(ZbotHandle* h, const ZbotMdpInputs* in, const float* actions, float* obs, float* rew, uint8_t* terminated,
                  uint8_t* truncated, int32_t stats_slot, int32_t prev_slot, void* stream) {
  if (int rc = mdp_check(h, in, true)) return rc;
  DeviceGuard guard(h->device);
  if (!actions || !obs || !rew || !terminated || !truncated) return fail(ZBOT_E_INVALID, "zbot_mdp_step: NULL buffer%s");
  if (int rc = check_slot(stats_slot, prev_slot, h->m_ring_slots)) return rc;
  const int n = h->cfg.num_envs, tile = h->mdp_tile, block = (tile + 31) & ~31;
  int grid = (n + tile - 1) / tile;
  MdpIn mi{in->body_link_pos_w, in->body_link_quat_w, in->body_com_lin_vel_w, in->joint_pos, in->joint_vel,
           in->applied_torque, in->net_forces_w_history, in->last_air_time, in->env_origins};
  StatsCtx sc{h->partials, h->m_ring, stats_slot, prev_slot, h->inv_episode_s, 0, h->rng_ctr, 0};
  ctx_fuse(sc, h);
  // persistent TMA-fed kernel (zbot_mdp_pipe.cuh) whenever every tensor can be a bulk-copy source (16-byte aligned base)
  const char* pipe_env = getenv("ZBOT_MDP_PIPE");      // read per call: tests A/B the two kernels in one process
  const bool pipe_off = pipe_env && atoi(pipe_env) == 0;
  const uintptr_t align_or = (uintptr_t)mi.pos | (uintptr_t)mi.quat | (uintptr_t)mi.vel | (uintptr_t)mi.q | (uintptr_t)mi.qd |
                             (uintptr_t)mi.tau | (uintptr_t)mi.hist | (uintptr_t)mi.last_air | (uintptr_t)mi.origins |
                             (uintptr_t)actions | (uintptr_t)h->m_ep_len | (uintptr_t)h->mstate;
  // measured (profiles/r2_notes.md sections 8, 9; CUDA-graph replay over rotating input sets): with 112-env tiles the one-shot
  // kernel is ahead at every size (65536 envs 27.0 us vs 28.6 persistent, 131072: 49.0 vs 49.1, 262144: 87.0 vs 89.7) and is the
  // default; ZBOT_MDP_PIPE=<shape> selects the persistent kernel
  if (!pipe_off && (align_or & 15) == 0 && pipe_env) {
    const int ntiles = (n + kPT - 1) / kPT;
    // ZBOT_MDP_PIPE = <stages><warps per stage> (tuning switch): 42 (default), 43, 33, 41
    const int shape = pipe_env ? atoi(pipe_env) : 0;
    const int stages = (shape == 33) ? 3 : 4;
    const char* fuse_env = getenv("ZBOT_MDP_FUSE_STATS");               // 1: grid-level statistics in the last CTA instead of a
    if (fuse_env && atoi(fuse_env) == 1) sc.ticket = reinterpret_cast<unsigned int*>(h->rng_ctr + 1);   // second launch (A/B switch)
    grid = std::min(h->num_sms, (ntiles + stages - 1) / stages);
#define ZB_PIPE_LAUNCH(ST, SH)                                                                                             \
  ZB_CUDA(launch_pdl(zbot_mdp_pipe_kernel<ST, SH>, dim3(grid), dim3(ST * SH * 32), pipe_smem(ST, SH), (cudaStream_t)stream, \
                     h->pdl, h->P, h->dp, mi, h->mstate, h->m_ep_len, actions, obs, rew, terminated, truncated, n, sc))
    if (shape == 43) { ZB_PIPE_LAUNCH(4, 3); }
    else if (shape == 33) { ZB_PIPE_LAUNCH(3, 3); }
    else if (shape == 41) { ZB_PIPE_LAUNCH(4, 1); }
    else { ZB_PIPE_LAUNCH(4, 2); }
#undef ZB_PIPE_LAUNCH
  } else {
    ZB_CUDA(launch_pdl(zbot_mdp_kernel<true>, dim3(grid), dim3(block), (size_t)tile * kHistRow * sizeof(float), (cudaStream_t)stream,
                       h->pdl, h->P, h->dp, mi, h->mstate, h->m_ep_len, actions, obs, rew, terminated, truncated, n, sc, tile));
  }
  h->launches += 1;
  if (!sc.ticket) {
    {
      ZB_CUDA(launch_pdl(zbot_stats_finalize_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, h->pdl, sc, (unsigned int)grid));
      h->launches += 1;
    }
  }
  return ZBOT_OK;
}

int zbot_policy_act(ZbotHandle* h, const ZbotPolicy* p, const float* obs, float* obs_out, float* act, float* logp,
                    float* value, float* mu, float* sigma, uint64_t seed, void* stream) {
  if (!h || !p || !obs || !act || !logp || !value || !mu || !sigma) return fail(ZBOT_E_INVALID, "zbot_policy_act: NULL argument%s");
  if (p->hidden != kPolHid || p->activation != 0)
    return fail(ZBOT_E_INVALID, "zbot_policy_act: only 3 x 128 ELU networks are built%s");
  if (p->num_obs < 1 || p->num_obs > kPolMaxObs || p->num_actions < 1 || p->num_actions > kPolMaxAct)
    return fail(ZBOT_E_INVALID, "zbot_policy_act: num_obs must be 1..64 and num_actions 1..8%s");
  if (!p->std) return fail(ZBOT_E_INVALID, "zbot_policy_act: std is NULL%s");
  PolicyArgs a{};
  for (int l = 0; l < 4; ++l) {
    a.w[0][l] = p->actor_w[l]; a.b[0][l] = p->actor_b[l];
    a.w[1][l] = p->critic_w[l]; a.b[1][l] = p->critic_b[l];
    if (!a.w[0][l] || !a.b[0][l] || !a.w[1][l] || !a.b[1][l]) return fail(ZBOT_E_INVALID, "zbot_policy_act: NULL weight%s");
    if (((uintptr_t)a.w[0][l] | (uintptr_t)a.w[1][l]) & 15) return fail(ZBOT_E_INVALID, "zbot_policy_act: weights must be 16-byte aligned%s");
  }
  DeviceGuard guard(h->device);
  a.std = p->std; a.obs = obs; a.obs_out = obs_out; a.act = act; a.logp = logp; a.value = value; a.mu = mu; a.sigma = sigma;
  a.ctr = h->rng_ctr; a.seed = seed; a.call = 0;
  a.n = h->cfg.num_envs; a.num_obs = p->num_obs; a.num_actions = p->num_actions;
  // rows per thread: 4 (256 threads per 64-env tile; default) or 8 (128 threads) -- ZBOT_POLICY_ROWS, tuning switch
  static const int rows = [] { const char* e = getenv("ZBOT_POLICY_ROWS"); return (e && atoi(e) == 8) ? 8 : 4; }();
  static bool attr_set[64] = {};
  if (h->device < 64 && !attr_set[h->device]) {
    ZB_CUDA(cudaFuncSetAttribute(zbot_policy_act_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolSmem));
    ZB_CUDA(cudaFuncSetAttribute(zbot_policy_act_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolSmem));
    attr_set[h->device] = true;
  }
  const dim3 grid((a.n + kPolTile - 1) / kPolTile, 2);
  // tcgen05 kernel (zbot_policy_tc5.cuh) by default; ZBOT_POLICY_TC=1 selects the mma.sync kernel (zbot_policy_tc.cuh), =0 the
  // CUDA-core FFMA2 kernels (read per call: tests compare the three in one process)
  const char* tc_env = getenv("ZBOT_POLICY_TC");
  if (!tc_env || atoi(tc_env) == 2) {              // tcgen05.mma / TMEM build (zbot_policy_tc5.cuh): the default
    static bool t5_attr_set[64] = {};
    if (h->device < 64 && !t5_attr_set[h->device]) {
      ZB_CUDA(cudaFuncSetAttribute(zbot_policy_act_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolT5Smem));
      t5_attr_set[h->device] = true;
    }
    zbot_policy_act_tc5_kernel<<<grid, kT5Threads, kPolT5Smem, (cudaStream_t)stream>>>(a);
  } else if (atoi(tc_env) != 0) {                  // ZBOT_POLICY_TC=1: mma.sync build (zbot_policy_tc.cuh)
    static bool tc_attr_set[64] = {};
    if (h->device < 64 && !tc_attr_set[h->device]) {
      ZB_CUDA(cudaFuncSetAttribute(zbot_policy_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolTcSmem));
      tc_attr_set[h->device] = true;
    }
    zbot_policy_act_tc_kernel<<<grid, kTcThreads, kPolTcSmem, (cudaStream_t)stream>>>(a);
  } else if (rows == 8)
    zbot_policy_act_kernel<8><<<grid, 128, kPolSmem, (cudaStream_t)stream>>>(a);
  else
    zbot_policy_act_kernel<4><<<grid, 256, kPolSmem, (cudaStream_t)stream>>>(a);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  return ZBOT_OK;
}

int zbot_rollout_store(ZbotHandle* h, const float* rew, const uint8_t* terminated, const uint8_t* truncated,
                       const float* value, float gamma, float* rew_out, float* done_out, void* stream) {
  if (!h || !rew || !terminated || !truncated || !value || !rew_out || !done_out)
    return fail(ZBOT_E_INVALID, "zbot_rollout_store: NULL argument%s");
  DeviceGuard guard(h->device);
  const int n = h->cfg.num_envs, block = 256;
  zbot_rollout_store_kernel<<<(n + block - 1) / block, block, 0, (cudaStream_t)stream>>>(rew, terminated, truncated, value, gamma,
                                                                                         rew_out, done_out, n);
  ZB_CUDA(cudaGetLastError());
  h->launches += 1;
  return ZBOT_OK;
}

int64_t zbot_launch_count(const ZbotHandle* h) { return h ? h->launches : 0; }

const char* zbot_step_kernel_name(const ZbotHandle* h) { return h ? h->kernel_name : ""; }

}  // extern "C"
