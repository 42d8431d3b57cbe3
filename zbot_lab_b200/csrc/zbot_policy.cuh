// zbot_policy.cuh -- the act half of the PPO rollout as ONE launch (BASELINE configs[4]).
//
// Between two env steps the reference's rollout (rsl_rl OnPolicyRunner over `agents/rsl_rl_ppo_cfg.py:65-91`: actor and critic
// MLP num_obs -> 128 -> 128 -> 128 -> {num_actions | 1}, ELU, state-independent std) evaluates the actor, samples a Gaussian
// action, evaluates its log-probability and the critic, and stores everything in the rollout buffer: ~45 small torch launches
// per step even inside a CUDA graph (115 us of a 140 us step at 4096 envs; the fused env step is 25 us of it).  Here it is one
// kernel: blockIdx.y = 0 runs the actor on a tile of 32 envs (+ sampling, log-prob, stores), blockIdx.y = 1 the critic.
//
// GEMM layout (FP32 on the CUDA cores -- the update phase differentiates the same weights in FP32 through torch, so the
// rollout must see the same numbers to round-off; TF32 / BF16 tensor cores would not): a CTA of 128 threads owns a 32-env x
// 128-neuron output tile per layer.  Thread (cg = t % 32, rg = t / 32) accumulates rows 8 rg .. 8 rg + 7 x columns
// {cg, cg + 32, cg + 64, cg + 96}.  Activations live K-major in shared memory (`xs[k][m]`, row stride 36 floats: the two
// 128-bit reads of a k-step are warp-wide broadcasts, the epilogue's 128-bit stores are conflict-free); the layer's weights
// stream through shared memory transposed in chunks of 32 input neurons (`ws[kk][j]`), the next chunk prefetched into
// registers while the current one is consumed.  Weights stay in torch's nn.Linear layout (W[out][in], b[out]) -- the
// optimizer updates them in place, so a captured graph keeps seeing the live parameters.
#pragma once
// (included inside zbot_kernels.cu's anonymous namespace, after v4_uniform)

constexpr int kPolTile = 32;      // envs per CTA
constexpr int kPolHid = 128;      // hidden width = threads per CTA
constexpr int kPolChunk = 32;     // input neurons per weight chunk
constexpr int kPolXS = 36;        // row stride of xs (floats)
constexpr int kPolMaxObs = 64;
constexpr int kPolMaxAct = 8;
constexpr size_t kPolSmem = (size_t)(2 * kPolHid * kPolXS + 2 * kPolChunk * kPolHid + kPolTile * kPolMaxAct) * sizeof(float);

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

__device__ __forceinline__ float pol_elu(float x) { return x > 0.f ? x : expm1f(x); }

// one hidden layer: xs_in[K][36] -> xs_out[128][36], ELU
__device__ __forceinline__ void pol_hidden_layer(const float* __restrict__ W, const float* __restrict__ bias, int K,
                                                 const float* xs_in, float* xs_out, float* ws) {
  const int t = threadIdx.x, cg = t & 31, rg = t >> 5;
  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const int nchunk = (K + kPolChunk - 1) / kPolChunk;
  const bool vec = (K & 3) == 0;
  float wreg[kPolChunk];
  auto prefetch = [&](int c) {      // this thread's output neuron t: W[t][c*32 .. c*32+31]
    const int k0 = c * kPolChunk;
    const float* row = W + (size_t)t * K + k0;
    if (vec && k0 + kPolChunk <= K) {
#pragma unroll
      for (int i = 0; i < kPolChunk / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        wreg[4 * i] = v.x; wreg[4 * i + 1] = v.y; wreg[4 * i + 2] = v.z; wreg[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kPolChunk; ++i) wreg[i] = (k0 + i < K) ? __ldg(row + i) : 0.f;
    }
  };
  auto stash = [&](int buf) {
    float* dst = ws + buf * kPolChunk * kPolHid + t;
#pragma unroll
    for (int i = 0; i < kPolChunk; ++i) dst[i * kPolHid] = wreg[i];
  };
  prefetch(0);
  stash(0);
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) prefetch(c + 1);
    const float* wb = ws + (c & 1) * kPolChunk * kPolHid;
    const float* xb = xs_in + (size_t)c * kPolChunk * kPolXS + 8 * rg;
    const int kc = min(kPolChunk, K - c * kPolChunk);
#pragma unroll 4
    for (int kk = 0; kk < kc; ++kk) {
      const float4 xa = *reinterpret_cast<const float4*>(xb + kk * kPolXS);
      const float4 xc = *reinterpret_cast<const float4*>(xb + kk * kPolXS + 4);
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
      float w4[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) w4[cc] = wb[kk * kPolHid + cg + 32 * cc];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[r][cc] = fmaf(x[r], w4[cc], acc[r][cc]);
    }
    if (c + 1 < nchunk) stash((c + 1) & 1);
    __syncthreads();
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int j = cg + 32 * cc;
    const float bj = __ldg(bias + j);
    float o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) o[r] = pol_elu(acc[r][cc] + bj);
    float* dst = xs_out + (size_t)j * kPolXS + 8 * rg;
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPolHid) zbot_policy_act_kernel(const PolicyArgs a) {
  extern __shared__ __align__(16) float psm[];
  float* xs0 = psm;
  float* xs1 = psm + kPolHid * kPolXS;
  float* ws = psm + 2 * kPolHid * kPolXS;
  float* outs = ws + 2 * kPolChunk * kPolHid;          // [32][8] head outputs
  const int net = blockIdx.y;
  const int e0 = blockIdx.x * kPolTile;
  const int valid = min(kPolTile, a.n - e0);
  const int t = threadIdx.x;
  // observation tile -> xs0[k][m] (rows of dead envs are zero); the actor CTA also writes the rollout buffer's copy
  for (int idx = t; idx < kPolTile * a.num_obs; idx += kPolHid) {
    const int m = idx / a.num_obs, k = idx - m * a.num_obs;
    float v = 0.f;
    if (m < valid) {
      v = __ldg(a.obs + (size_t)e0 * a.num_obs + idx);
      if (net == 0 && a.obs_out) a.obs_out[(size_t)e0 * a.num_obs + idx] = v;
    }
    xs0[k * kPolXS + m] = v;
  }
  __syncthreads();
  pol_hidden_layer(a.w[net][0], a.b[net][0], a.num_obs, xs0, xs1, ws);
  pol_hidden_layer(a.w[net][1], a.b[net][1], kPolHid, xs1, xs0, ws);
  pol_hidden_layer(a.w[net][2], a.b[net][2], kPolHid, xs0, xs1, ws);
  // head: num_actions (actor) or 1 (critic) outputs per env; thread (m = t % 32, o = t / 32 [+ 4])
  const int nout = net == 0 ? a.num_actions : 1;
  const int m = t & 31;
  for (int o = t >> 5; o < nout; o += 4) {
    const float* wrow = a.w[net][3] + (size_t)o * kPolHid;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < kPolHid; k += 2) {
      s0 = fmaf(xs1[k * kPolXS + m], __ldg(wrow + k), s0);
      s1 = fmaf(xs1[(k + 1) * kPolXS + m], __ldg(wrow + k + 1), s1);
    }
    outs[m * kPolMaxAct + o] = s0 + s1 + __ldg(a.b[net][3] + o);
  }
  __syncthreads();
  if (t >= valid) return;
  const int e = e0 + t;
  if (net == 1) {
    a.value[e] = outs[t * kPolMaxAct];
    return;
  }
  // Gaussian sample (Box-Muller on the counter-based uniforms of the step kernels: seed / stream position / env / slot),
  // log-probability evaluated from the stored action exactly as torch.distributions.Normal.log_prob does
  const unsigned long long call = a.ctr ? __ldcg(a.ctr) : a.call;
  float lp = 0.f;
  for (int o = 0; o < a.num_actions; ++o) {
    const float mean = outs[t * kPolMaxAct + o];
    const float sd = fmaxf(__ldg(a.std + o), 1e-6f);
    const float u1 = 1.0f - v4_uniform(a.seed, call, (uint32_t)e, 128u + 2u * (uint32_t)o);       // (0, 1]
    const float u2 = v4_uniform(a.seed, call, (uint32_t)e, 129u + 2u * (uint32_t)o);
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float z = sqrtf(-2.0f * logf(u1)) * cs;
    const float act = fmaf(sd, z, mean);
    const float d = act - mean;
    lp += -(d * d) / (2.0f * sd * sd) - logf(sd) - 0.91893853320467274178f;
    a.act[(size_t)e * a.num_actions + o] = act;
    a.mu[(size_t)e * a.num_actions + o] = mean;
    a.sigma[(size_t)e * a.num_actions + o] = sd;
  }
  a.logp[e] = lp;
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

