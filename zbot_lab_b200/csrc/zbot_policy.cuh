// zbot_policy.cuh -- the act half of the PPO rollout as ONE launch (BASELINE configs[4]).
//
// Between two env steps the reference's rollout (rsl_rl OnPolicyRunner over `agents/rsl_rl_ppo_cfg.py:65-91`: actor and critic
// MLP num_obs -> 128 -> 128 -> 128 -> {num_actions | 1}, ELU, state-independent std) evaluates the actor, samples a Gaussian
// action, evaluates its log-probability and the critic, and stores everything in the rollout buffer: ~45 small torch launches
// per step even inside a CUDA graph (115 us of a 140 us step at 4096 envs; the fused env step is 25 us of it).  Here it is one
// kernel: blockIdx.y = 0 runs the actor on a tile of 64 envs (+ sampling, log-prob, stores), blockIdx.y = 1 the critic.
//
// (This file: the CUDA-core build, ZBOT_POLICY_TC=0, 27.1 us at 4096 envs.  The default since the end of round 2 is
// zbot_policy_tc5.cuh -- the same kernel with the three hidden layers as tcgen05.mma kind::tf32 (accumulators in TMEM), every
// product split three ways so the result is FP32 to round-off: 16.9 us; zbot_policy_tc.cuh is the mma.sync build of that
// arithmetic, ZBOT_POLICY_TC=1, 19.2 us.)
//
// GEMM layout (FP32 on the CUDA cores -- the update phase differentiates the same weights in FP32 through torch, so the
// rollout must see the same numbers to round-off; plain TF32 / BF16 tensor-core products would not).  Packed FP32 (FFMA2: the FMA pipe of
// sm_100a retires a scalar FFMA every other cycle per sub-partition, tools/micro/ffma2_probe.cu -- 39 TFLOP/s scalar against
// 57.5 packed at two warps per sub-partition).  A CTA of 256 threads owns a 64-env x 128-neuron output tile per layer.  Thread
// (cg = t % 16, rg = t / 16) accumulates rows 4 rg .. 4 rg + 3 x columns {cg + 16 c, c = 0..7} as 4 x 4 register pairs:
//     (acc[r][2p], acc[r][2p+1]) += x[r] (scalar, broadcast operand form) * (w[2p], w[2p+1])        one FFMA2
// Activations live K-major in shared memory (`xs[k][m]`, row stride 68 floats: the 128-bit read of a k-step is two addresses
// per warp, the epilogue's 128-bit stores spread evenly over the banks); the layer's weights stream through shared memory
// transposed in chunks of 32 input neurons (`ws[kk][pos]`), column j = cg + 16 c parked at pos = 4 cg + (c & 3) + 64 (c >> 2) so a
// thread's eight weights of a k-step are two 128-bit reads and the register pairs are the natural ones.  The next chunk is
// prefetched into registers while the current one is consumed.  Weights stay in torch's nn.Linear layout (W[out][in], b[out])
// -- the optimizer updates them in place, so a captured graph keeps seeing the live parameters.
#pragma once
// (included inside zbot_kernels.cu's anonymous namespace, after v4_uniform)

constexpr int kPolTile = 64;      // envs per CTA
constexpr int kPolHid = 128;      // hidden width
constexpr int kPolChunk = 32;     // input neurons per weight chunk
constexpr int kPolXS = 68;        // row stride of xs (floats)
constexpr int kPolMaxObs = 64;
constexpr int kPolMaxAct = 8;
constexpr size_t kPolSmem = (size_t)(2 * kPolHid * kPolXS + 2 * kPolChunk * kPolHid + 4 * kPolTile * kPolMaxAct + kPolMaxAct * kPolHid + kPolMaxAct) * sizeof(float);

struct PolicyArgs {
  const float* w[2][4];        // [net: 0 actor, 1 critic][layer]  W[out][in]
  const float* b[2][4];
  const float* std;            // [num_actions]
  const float* obs;            // [n][num_obs]
  float* obs_out;              // [n][num_obs] or null: the rollout buffer's copy of the observation
  float* act;                  // [n][num_actions]
  float* logp;                 // [n]
  float* value;                // [n]
  float* mu;                   // [n][num_actions]
  float* sigma;                // [n][num_actions]
  const unsigned long long* ctr;   // device counter (stream position) or null
  unsigned long long seed, call;
  int n, num_obs, num_actions;
};

// torch.nn.ELU(alpha = 1) as ATen evaluates it: x > 0 ? x : exp(x) - 1.  Branch-free; ex2.approx on the clamped argument
// (absolute error <= 2^-22 on (-1, 0], the branch that is taken there)
__device__ __forceinline__ float pol_elu(float x) {
  const float e = __expf(fminf(x, 0.f)) - 1.f;
  return x > 0.f ? x : e;
}

// one hidden layer: xs_in[K][68] -> xs_out[128][68], ELU.  kFull: K is a multiple of the chunk (the 128-wide layers) -- the 32
// k-steps of a chunk are unrolled completely so the shared-memory reads of the steps ahead are in flight under the FFMA2s.
// kRows = rows per thread (4: 256 threads, 8: 128 threads).  A 128-bit shared-memory read is served per HALF-warp, one
// wavefront per 128 distinct bytes (ncu: the 16-column-groups-per-half-warp mapping cost 4 + 4 + 2 wavefronts per k-step), so a
// half-warp is 8 column groups x 2 row groups: 1 + 1 wavefronts for the weights, 1 per four rows for the activations.
// Measured (profiles/r2_notes.md section 7): the k-loop runs at ~105 clocks per k-step and SM = 78 FMA / clock / SM against the
// 110 a pure FFMA2 stream reaches (tools/micro/ffma2_bcast_probe.cu); what bounds it is the shared-memory -> register return
// path (128 B / clock / SM: 256 threads x 48 B per k-step = 96 clocks), not wavefronts and not the FMA pipe.
// weight staging: thread t parks weight row sj (shared-memory position sp = t % 128) of a chunk, k-part sh = t / 128
template <int kStage>
__device__ __forceinline__ void pol_prefetch(const float* __restrict__ W, int K, int c, float (&wreg)[kStage]) {
  const int sp = threadIdx.x & 127, sh = threadIdx.x >> 7;
  const int sj = ((sp & 63) >> 2) + 16 * ((sp & 3) + 4 * (sp >> 6));
  const int k0 = c * kPolChunk + kStage * sh;
  const float* row = W + (size_t)sj * K + k0;
  if ((K & 3) == 0 && k0 + kStage <= K) {
#pragma unroll
    for (int i = 0; i < kStage / 4; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
      wreg[4 * i] = v.x; wreg[4 * i + 1] = v.y; wreg[4 * i + 2] = v.z; wreg[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kStage; ++i) wreg[i] = (k0 + i < K) ? __ldg(row + i) : 0.f;
  }
}
template <int kStage>
__device__ __forceinline__ void pol_stash(float* ws, int buf, const float (&wreg)[kStage]) {
  const int sp = threadIdx.x & 127, sh = threadIdx.x >> 7;
  float* dst = ws + buf * kPolChunk * kPolHid + kStage * sh * kPolHid + sp;
#pragma unroll
  for (int i = 0; i < kStage; ++i) dst[i * kPolHid] = wreg[i];
}

// On entry chunk 0 of this layer is already in ws[buf] (staged by the caller / the previous layer, a barrier has passed); during
// the last chunk the first chunk of the NEXT layer (Wn, Kn; null: none) is fetched and parked in the other buffer, so no layer
// starts with an exposed L2 round trip.  On exit `buf` names the buffer holding the next layer's chunk 0.
template <bool kFull, int kRows>
__device__ __forceinline__ void pol_hidden_layer(const float* __restrict__ W, const float* __restrict__ bias, int K,
                                                 const float* xs_in, float* xs_out, float* ws, int& buf,
                                                 const float* __restrict__ Wn, int Kn) {
  constexpr int kThreads = (kPolTile / kRows) * 16;
  constexpr int kStage = kPolChunk * kPolHid / kThreads;     // weights of a chunk staged per thread (16 or 32)
  const int t = threadIdx.x, hw = t >> 4;
  const int cg = (t & 7) + 8 * (hw & 1), rg = ((t >> 3) & 1) + 2 * (hw >> 1);
  float2 acc[kRows][4];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[r][p] = make_float2(0.f, 0.f);
  const int nchunk = (K + kPolChunk - 1) / kPolChunk;
  float wreg[kStage];
  for (int c = 0; c < nchunk; ++c) {
    const bool last = c + 1 == nchunk;
    if (!last) pol_prefetch<kStage>(W, K, c + 1, wreg);
    else if (Wn) pol_prefetch<kStage>(Wn, Kn, 0, wreg);
    const float* wb = ws + buf * kPolChunk * kPolHid + 4 * cg;
    const float* xb = xs_in + (size_t)c * kPolChunk * kPolXS + kRows * rg;
    auto kstep = [&](int kk) {
      float x[kRows];
#pragma unroll
      for (int q = 0; q < kRows / 4; ++q) {
        const float4 xa = *reinterpret_cast<const float4*>(xb + kk * kPolXS + 4 * q);
        x[4 * q] = xa.x; x[4 * q + 1] = xa.y; x[4 * q + 2] = xa.z; x[4 * q + 3] = xa.w;
      }
      const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * kPolHid);
      const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * kPolHid + 64);
      const float2 wp[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[r][p] = __ffma2_rn(make_float2(x[r], x[r]), wp[p], acc[r][p]);
    };
    if (kFull) {
#pragma unroll
      for (int kk = 0; kk < kPolChunk; ++kk) kstep(kk);
    } else {
      const int kc = min(kPolChunk, K - c * kPolChunk);
#pragma unroll 4
      for (int kk = 0; kk < kc; ++kk) kstep(kk);
    }
    if (!last || Wn) pol_stash<kStage>(ws, buf ^ 1, wreg);     // the other buffer was last read before the previous barrier
    if (!last) {
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int j = cg + 16 * c;
    const float bj = __ldg(bias + j);
#pragma unroll
    for (int q = 0; q < kRows / 4; ++q) {
      float o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) o[r] = pol_elu(((c & 1) ? acc[4 * q + r][c >> 1].y : acc[4 * q + r][c >> 1].x) + bj);
      *reinterpret_cast<float4*>(xs_out + (size_t)j * kPolXS + kRows * rg + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  buf ^= 1;
}

template <int kRows>
__global__ void __launch_bounds__((kPolTile / kRows) * 16, 1) zbot_policy_act_kernel(const PolicyArgs a) {
  constexpr int kPolThreads = (kPolTile / kRows) * 16;
  extern __shared__ __align__(16) float psm[];
  float* xs0 = psm;
  float* xs1 = psm + kPolHid * kPolXS;
  float* ws = psm + 2 * kPolHid * kPolXS;
  float* outs = ws + 2 * kPolChunk * kPolHid;          // [4][64][8] partial head outputs
  float* hws = outs + 4 * kPolTile * kPolMaxAct;       // [8][128] head weights, [8] head bias
  const int net = blockIdx.y;
  const int e0 = blockIdx.x * kPolTile;
  const int valid = min(kPolTile, a.n - e0);
  const int t = threadIdx.x;
  // first weight chunk of the first layer: in flight while the observation tile is loaded
  constexpr int kStage = kPolChunk * kPolHid / kPolThreads;
  {
    float wreg[kStage];
    pol_prefetch<kStage>(a.w[net][0], a.num_obs, 0, wreg);
    pol_stash<kStage>(ws, 0, wreg);
  }
  // head weights + bias -> shared memory now (consumed after the third layer; nothing waits for them there)
  const int nout = net == 0 ? a.num_actions : 1;
  for (int i = t; i < nout * (kPolHid / 4); i += kPolThreads)
    reinterpret_cast<float4*>(hws)[i] = __ldg(reinterpret_cast<const float4*>(a.w[net][3]) + i);
  if (t < nout) hws[kPolMaxAct * kPolHid + t] = __ldg(a.b[net][3] + t);
  // observation tile -> xs0[k][m] (rows of dead envs are zero): warp w takes envs w, w + 8, ..., lane = column (+ 32); the
  // actor CTA also writes the rollout buffer's copy
  {
    const int lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int i = 0; i < kPolTile / (kPolThreads / 32); ++i) {
      const int m = wid + (kPolThreads / 32) * i;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        if (k < a.num_obs) {
          float v = 0.f;
          if (m < valid) {
            v = __ldg(a.obs + (size_t)(e0 + m) * a.num_obs + k);
            if (net == 0 && a.obs_out) a.obs_out[(size_t)(e0 + m) * a.num_obs + k] = v;
          }
          xs0[k * kPolXS + m] = v;
        }
      }
    }
  }
  __syncthreads();
  int buf = 0;
  pol_hidden_layer<false, kRows>(a.w[net][0], a.b[net][0], a.num_obs, xs0, xs1, ws, buf, a.w[net][1], kPolHid);
  pol_hidden_layer<true, kRows>(a.w[net][1], a.b[net][1], kPolHid, xs1, xs0, ws, buf, a.w[net][2], kPolHid);
  pol_hidden_layer<true, kRows>(a.w[net][2], a.b[net][2], kPolHid, xs0, xs1, ws, buf, nullptr, 0);
  // head: num_actions (actor) or 1 (critic) outputs per env.  Thread (m = t % 64, kq = t / 64) sums its quarter (half) of the
  // 128 inputs for every output; the partial sums meet in shared memory and are added in a fixed order.
  constexpr int kG = kPolThreads / 64;
  const int m = t & 63, kq = t >> 6;
  {
    float part[kPolMaxAct];
#pragma unroll
    for (int o = 0; o < kPolMaxAct; ++o) part[o] = 0.f;
    const int kb = kq * (kPolHid / kG);
#pragma unroll 4
    for (int k = kb; k < kb + kPolHid / kG; ++k) {
      const float x = xs1[k * kPolXS + m];
#pragma unroll
      for (int o = 0; o < kPolMaxAct; ++o)
        if (o < nout) part[o] = fmaf(x, hws[o * kPolHid + k], part[o]);
    }
#pragma unroll
    for (int o = 0; o < kPolMaxAct; ++o)
      if (o < nout) outs[(kq * kPolTile + m) * kPolMaxAct + o] = part[o];
  }
  __syncthreads();
  auto head_out = [&](int mm, int o) {
    float v = outs[mm * kPolMaxAct + o];
#pragma unroll
    for (int g = 1; g < kG; ++g) v += outs[(g * kPolTile + mm) * kPolMaxAct + o];
    return v + hws[kPolMaxAct * kPolHid + o];
  };
  if (net == 1) {
    if (t < valid) a.value[e0 + t] = head_out(t, 0);
    return;
  }
  // Gaussian sample (Box-Muller on the counter-based uniforms of the step kernels: seed / stream position / env / slot), one
  // (env, action) pair per thread and round; the log-probability terms are evaluated from the stored action exactly as
  // torch.distributions.Normal.log_prob does and summed over the actions in index order
  const unsigned long long call = a.ctr ? __ldcg(a.ctr) : a.call;
  float* lpc = xs0;                                    // [64][8] log-prob terms (xs0 is dead after the third layer)
  for (int o = kq; o < a.num_actions; o += kG) {
    if (m < valid) {
      const int e = e0 + m;
      const float mean = head_out(m, o);
      const float sd = fmaxf(__ldg(a.std + o), 1e-6f);
      const float u1 = 1.0f - v4_uniform(a.seed, call, (uint32_t)e, 128u + 2u * (uint32_t)o);       // (0, 1]
      const float u2 = v4_uniform(a.seed, call, (uint32_t)e, 129u + 2u * (uint32_t)o);
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      const float z = sqrtf(-2.0f * logf(u1)) * cs;
      const float act = fmaf(sd, z, mean);
      const float d = act - mean;
      lpc[m * kPolMaxAct + o] = -(d * d) / (2.0f * sd * sd) - logf(sd) - 0.91893853320467274178f;
      a.act[(size_t)e * a.num_actions + o] = act;
      a.mu[(size_t)e * a.num_actions + o] = mean;
      a.sigma[(size_t)e * a.num_actions + o] = sd;
    }
  }
  __syncthreads();
  if (t < valid) {
    float lp = 0.f;
    for (int o = 0; o < a.num_actions; ++o) lp += lpc[t * kPolMaxAct + o];
    a.logp[e0 + t] = lp;
  }
}

// the store half of a rollout step: reward with the time-out bootstrap (rew + gamma * V(s_t) * time_out) and the
// done flag (terminated | truncated) as float, into the rollout buffer
__global__ void zbot_rollout_store_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ terminated,
                                          const uint8_t* __restrict__ truncated, const float* __restrict__ value, float gamma,
                                          float* __restrict__ rew_out, float* __restrict__ done_out, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool to = truncated[e] != 0;
  float r = rew[e];
  if (to) r = fmaf(gamma, value[e], r);
  rew_out[e] = r;
  done_out[e] = (to || terminated[e]) ? 1.f : 0.f;
}

