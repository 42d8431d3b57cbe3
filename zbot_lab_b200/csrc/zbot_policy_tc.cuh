// zbot_policy_tc.cuh -- the act half of the PPO rollout on the tensor cores through the legacy warp-level path (mma.sync), at
// FP32 accuracy (3 x TF32 split products).  ZBOT_POLICY_TC=1; the default is the tcgen05 / TMEM build, zbot_policy_tc5.cuh.
//
// Same contract as zbot_policy_act_kernel (zbot_policy.cuh): blockIdx.y = 0 actor (+ Gaussian sample, log-prob, stores),
// blockIdx.y = 1 critic, a tile of 64 envs per CTA, weights read live in torch's nn.Linear layout (W[out][in]).  The three
// hidden layers (64 x K) x (K x 128) are warp-level MMAs (`mma.sync.aligned.m16n8k8 ... tf32`, FP32 accumulate); the head
// (128 -> num_actions | 1) and everything after it are the FP32 code of the CUDA-core kernel.
//
// Why this is still FP32 to round-off (the PPO update differentiates the same weights in FP32 through torch and compares
// log-probabilities with the rollout's): every operand x is split as x = hi + lo, hi = x rounded to TF32 (round half away, as cvt.rna), lo = x - hi
// (exact in FP32; the tensor core reads its upper 19 bits), and a product is evaluated as lo_a hi_b + hi_a lo_b + hi_a hi_b with
// the small terms first.  The dropped lo_a lo_b term is 2^-22 relative; a numpy emulation with a round-toward-zero accumulator
// puts the network outputs within 1e-6 of float64, torch's own FP32 path within 3e-7 (tests/test_gpu_policy.py holds the
// kernel to 2e-5 absolute against torch FP32 and to 4 x torch's own distance from float64 + 2e-6).
//
// Shared memory (211 KB, one CTA per SM): activations `xs[m][k]` and weights `ws[n][k]`, both with a row stride of 132 floats,
// so the fragment loads -- a: (row g | g + 8, k = tig | tig + 4), b: (n = g, k = tig | tig + 4), g = lane / 4, tig = lane % 4 --
// hit bank 4 g + tig: conflict-free without a swizzle.  A layer's whole weight matrix (128 x 128 floats) is one buffer; the
// next layer's arrives by cp.async (16 B per request, straight from the nn.Linear rows) while the current one is multiplied.
// A warp owns a 32 x 32 output tile: 2 x 4 MMA tiles, 8 + 8 fragment loads, 24 MMAs per k-step of 8.
#pragma once
// (included inside zbot_kernels.cu's anonymous namespace, after zbot_policy.cuh)

constexpr int kTcXS = 132;                                   // row stride of xs / ws in floats (= 4 mod 32)
constexpr int kTcThreads = 256;
constexpr size_t kPolTcSmem =
    (size_t)(2 * kPolTile * kTcXS + 2 * kPolHid * kTcXS + 4 * kPolTile * kPolMaxAct + kPolMaxAct * kPolHid + kPolMaxAct) * sizeof(float);

// x rounded to TF32 (10 mantissa bits, round half away from zero = cvt.rna.tf32.f32, which ptxas expands to ~6 instructions on
// sm_100a): add half an ulp of the kept field to the bit pattern, clear the 13 dropped bits.  Finite inputs only (activations and
// weights); a carry out of the mantissa moves to the next exponent, which is the correctly rounded result.
__device__ __forceinline__ uint32_t tc_tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void tc_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tc_tf32_hi(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void tc_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void tc_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void tc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// a 128 x 128 nn.Linear weight (16-byte aligned rows) -> ws[n][k], asynchronously
__device__ __forceinline__ void tc_stage_weights_async(const float* __restrict__ W, float* ws) {
  for (int i = threadIdx.x; i < kPolHid * (kPolHid / 4); i += kTcThreads) {
    const int n = i >> 5, kq = i & 31;
    tc_cp_async16(ws + n * kTcXS + 4 * kq, W + (size_t)n * kPolHid + 4 * kq);
  }
  tc_cp_commit();
}
// the first layer's 128 x K weight (K = num_obs: any value, rows not aligned) -> ws[n][k], zero-padded to K8
__device__ __forceinline__ void tc_stage_weights_first(const float* __restrict__ W, int K, int K8, float* ws) {
  // thread -> (row n = t / 2 + 128 j ... ) without a division: two threads per weight row, alternating columns
  const int n = threadIdx.x >> 1, h = threadIdx.x & 1;
  for (int k = h; k < K8; k += 2) ws[n * kTcXS + k] = (k < K) ? __ldg(W + (size_t)n * K + k) : 0.f;
}

// xs_out[m][n] = ELU(sum_k xs_in[m][k] ws[n][k] + bias[n]),  m < 64, n < 128, k < K8 (a multiple of 8; padding is zero)
__device__ __forceinline__ void tc_hidden_layer(const float* xs_in, const float* ws, const float* __restrict__ bias, int K8,
                                                float* xs_out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const int m0 = 32 * (w & 1), n0 = 32 * (w >> 1);
  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  const float* ap = xs_in + (m0 + g) * kTcXS + tig;
  const float* bp = ws + (n0 + g) * kTcXS + tig;
#pragma unroll 2
  for (int k0 = 0; k0 < K8; k0 += 8) {
    uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* p = ap + 16 * mt * kTcXS + k0;
      tc_split(p[0], ah[mt][0], al[mt][0]);
      tc_split(p[8 * kTcXS], ah[mt][1], al[mt][1]);
      tc_split(p[4], ah[mt][2], al[mt][2]);
      tc_split(p[8 * kTcXS + 4], ah[mt][3], al[mt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* p = bp + 8 * nt * kTcXS + k0;
      tc_split(p[0], bh[nt][0], bl[nt][0]);
      tc_split(p[4], bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        tc_mma(acc[mt][nt], al[mt], bh[nt]);       // small terms first
        tc_mma(acc[mt][nt], ah[mt], bl[nt]);
        tc_mma(acc[mt][nt], ah[mt], bh[nt]);
      }
  }
  // epilogue: c0 (row g, col 2 tig), c1 (g, 2 tig + 1), c2 (g + 8, 2 tig), c3 (g + 8, 2 tig + 1)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = n0 + 8 * nt + 2 * tig;
    const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int row = m0 + 16 * mt + g;
      *reinterpret_cast<float2*>(xs_out + row * kTcXS + col) =
          make_float2(pol_elu(acc[mt][nt][0] + b0), pol_elu(acc[mt][nt][1] + b1));
      *reinterpret_cast<float2*>(xs_out + (row + 8) * kTcXS + col) =
          make_float2(pol_elu(acc[mt][nt][2] + b0), pol_elu(acc[mt][nt][3] + b1));
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) zbot_policy_act_tc_kernel(const PolicyArgs a) {
  extern __shared__ __align__(16) float psm[];
  float* xs0 = psm;                                    // [64][132]
  float* xs1 = xs0 + kPolTile * kTcXS;
  float* wsA = xs1 + kPolTile * kTcXS;                 // [128][132]
  float* wsB = wsA + kPolHid * kTcXS;
  float* outs = wsB + kPolHid * kTcXS;                 // [4][64][8] partial head outputs
  float* hws = outs + 4 * kPolTile * kPolMaxAct;       // [8][128] head weights, [8] head bias
  const int net = blockIdx.y;
  const int e0 = blockIdx.x * kPolTile;
  const int valid = min(kPolTile, a.n - e0);
  const int t = threadIdx.x;
  const int K8 = (a.num_obs + 7) & ~7;
  // second layer's weights: in flight (cp.async) behind everything below
  tc_stage_weights_async(a.w[net][1], wsB);
  tc_stage_weights_first(a.w[net][0], a.num_obs, K8, wsA);
  const int nout = net == 0 ? a.num_actions : 1;
  for (int i = t; i < nout * (kPolHid / 4); i += kTcThreads)
    reinterpret_cast<float4*>(hws)[i] = __ldg(reinterpret_cast<const float4*>(a.w[net][3]) + i);
  if (t < nout) hws[kPolMaxAct * kPolHid + t] = __ldg(a.b[net][3] + t);
  // observation tile -> xs0[m][k], zero beyond num_obs and for dead envs; the actor CTA also writes the rollout buffer's copy
  // (warp w takes envs w, w + 8, ...; lane = column, + 32 for the second half of a wide observation)
  {
    const int lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int i = 0; i < kPolTile / (kTcThreads / 32); ++i) {
      const int m = wid + (kTcThreads / 32) * i;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = lane + 32 * hh;
        if (k < K8) {
          float v = 0.f;
          if (m < valid && k < a.num_obs) {
            v = __ldg(a.obs + (size_t)(e0 + m) * a.num_obs + k);
            if (net == 0 && a.obs_out) a.obs_out[(size_t)(e0 + m) * a.num_obs + k] = v;
          }
          xs0[m * kTcXS + k] = v;
        }
      }
    }
  }
  __syncthreads();
  tc_hidden_layer(xs0, wsA, a.b[net][0], K8, xs1);
  __syncthreads();                                     // wsA and xs0 are free, xs1 is complete
  tc_stage_weights_async(a.w[net][2], wsA);            // third layer's weights behind the second layer's MMAs
  tc_cp_wait<1>();                                     // this thread's share of the second layer's weights has landed
  __syncthreads();
  tc_hidden_layer(xs1, wsB, a.b[net][1], kPolHid, xs0);
  tc_cp_wait<0>();
  __syncthreads();
  tc_hidden_layer(xs0, wsA, a.b[net][2], kPolHid, xs1);
  __syncthreads();
  // head: num_actions (actor) or 1 (critic) outputs per env.  Thread (m = t % 64, kq = t / 64) sums its quarter of the 128
  // inputs for every output (128-bit reads along k: a quarter-warp covers all 32 banks); the partial sums meet in shared
  // memory and are added in a fixed order.
  constexpr int kG = kTcThreads / 64;
  const int m = t & 63, kq = t >> 6;
  {
    float part[kPolMaxAct];
#pragma unroll
    for (int o = 0; o < kPolMaxAct; ++o) part[o] = 0.f;
    const int kb = kq * (kPolHid / kG);
#pragma unroll 2
    for (int k = kb; k < kb + kPolHid / kG; k += 4) {
      const float4 x = *reinterpret_cast<const float4*>(xs1 + m * kTcXS + k);
#pragma unroll
      for (int o = 0; o < kPolMaxAct; ++o)
        if (o < nout) {
          const float4 wv = *reinterpret_cast<const float4*>(hws + o * kPolHid + k);
          part[o] = fmaf(x.w, wv.w, fmaf(x.z, wv.z, fmaf(x.y, wv.y, fmaf(x.x, wv.x, part[o]))));
        }
    }
#pragma unroll
    for (int o = 0; o < kPolMaxAct; ++o)
      if (o < nout) outs[(kq * kPolTile + m) * kPolMaxAct + o] = part[o];
  }
  __syncthreads();
  auto head_out = [&](int mm, int o) {
    float v = outs[mm * kPolMaxAct + o];
#pragma unroll
    for (int gq = 1; gq < kG; ++gq) v += outs[(gq * kPolTile + mm) * kPolMaxAct + o];
    return v + hws[kPolMaxAct * kPolHid + o];
  };
  if (net == 1) {
    if (t < valid) a.value[e0 + t] = head_out(t, 0);
    return;
  }
  // Gaussian sample, log-probability, stores: identical to zbot_policy_act_kernel (same generator, same slots, same formulas)
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
