// zbot_policy_tc5.cuh -- the act half of the PPO rollout on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Same contract and the same FP32-accurate arithmetic as zbot_policy_tc.cuh (every product split as lo*hi + hi*lo + hi*hi over
// TF32 operands, FP32 accumulate), but the three hidden layers are `tcgen05.mma.cta_group::1.kind::tf32` instructions issued by
// ONE thread of the CTA: D[64 x 128] (TMEM, 128 columns) += A[64 x 8] (shared memory) * B[128 x 8]^T (shared memory), 16 k-steps
// x 3 split terms per 128-wide layer.  The legacy `mma.sync` path of the other file runs at ~145 TFLOP/s of TF32 on sm_100a and
// is what bounds that kernel (12 of its 19 us); this path is not MMA-bound any more.
//
// Operand layout = the canonical K-major, no-swizzle UMMA layout (cute::UMMA::LayoutType::SWIZZLE_NONE: mma_traits_sm100.hpp
// "make_umma_desc<Major::K>"): 8-row x 16-byte core matrices stored as 128 contiguous bytes; the core matrices of one 8-row group
// follow each other along K (leading byte offset 128), the 8-row groups are 4096 bytes apart (stride byte offset = 32 chunks x
// 128 B).  Element (r, k) of a tile therefore sits at (r / 8) * 4096 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4 bytes; one MMA
// consumes two 16-byte chunks (K = 8 TF32), so k-step j starts 256 j bytes into the tile.  Four tiles: A_hi, A_lo (64 rows, 32 KB
// each), B_hi, B_lo (128 rows = output neurons, 64 KB each) = 192 KB.
//
// Accumulator layout for M = 64 (cute tmem_frg_1sm: ((16,4),N):((1,32),128)): row r lives in TMEM lane (r % 16) + 32 (r / 16), so
// the 16 rows of lane quadrant q belong to warps q and q + 4 (a warp may only touch lanes 32 (warp % 4) ..+31); warp w
// reads rows 16 (w % 4) ..+15, columns 64 (w / 4) ..+63 with one `tcgen05.ld.16x256b.x8` (accumulator-fragment distribution:
// every thread gets two rows x 16 column pairs), applies bias + ELU, splits the result and writes it straight into the A_hi / A_lo
// tiles of the next layer.
//
// Order of events per layer: [all] next layer's weights -> registers (global loads in flight) | wait for the MMAs of this layer
// (mbarrier armed by tcgen05.commit) | epilogue TMEM -> A tiles | weights registers -> B tiles | fence.proxy.async, barrier |
// [thread 0] 48 x tcgen05.mma, tcgen05.commit.
//
// Measured (profiles/r2_notes.md section 12): 17.0 us at 4096 envs and 16.4 us for a dozen CTAs -- the kernel is a chain of
// dependent latencies (L2 round trips of the operand loads, MMA phase, TMEM read-back, barriers), not a throughput problem.
// Requesting the weights a whole layer ahead, or all prologue loads before the first dependent store, or the biases through shared
// memory each made it SLOWER (17.7 / 19.0 / 18.5 us): loads that are still in flight hold the scoreboards the next dependent
// instruction waits on.  Issuing a layer as two N = 64 halves with a commit each (read-back of one half under the MMAs of the
// other) costs twice the MMA time -- an N = 64 instruction takes as long as an N = 128 one -- 19.5 us.
#pragma once
// (included inside zbot_kernels.cu's anonymous namespace, after zbot_policy_tc.cuh)

constexpr int kT5Threads = 256;
constexpr int kT5TileA = kPolTile * kPolHid;                 // floats per A tile (64 x 128)
constexpr int kT5TileB = kPolHid * kPolHid;                  // floats per B tile (128 x 128)
constexpr size_t kPolT5Smem = (size_t)(2 * kT5TileA + 2 * kT5TileB) * sizeof(float);

__device__ __forceinline__ int t5_off(int r, int k) { return (r >> 3) * 1024 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }   // in floats

__device__ __forceinline__ uint64_t t5_desc(const void* smem_tile) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)(128u >> 4) << 16;              // leading byte offset (between the two 16-byte K chunks of an MMA), bits [16,30)
  d |= (uint64_t)(4096u >> 4) << 32;             // stride byte offset (between 8-row groups), bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  return d;                                      // base offset 0, LBO mode 0, layout type 0 = no swizzle
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0),
// N = 128 (bits 17-22 = N / 8), M = 64 (bits 24-28 = M / 16)
constexpr uint32_t kT5Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPolHid >> 3) << 17) | ((uint32_t)(kPolTile >> 4) << 24);

__device__ __forceinline__ void t5_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kT5Idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void t5_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 TMEM lanes x 64 columns -> 32 registers per thread, distributed like an MMA accumulator fragment: thread t holds, for each
// group j of 8 columns, (lane t / 4, columns 8 j + 2 (t % 4) + {0, 1}) in registers 4 j + {0, 1} and (lane t / 4 + 8, same
// columns) in registers 4 j + {2, 3} -- all 32 threads get data although an M = 64 accumulator only fills 16 lanes per quadrant
__device__ __forceinline__ void t5_ld16x64(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// weights of a 128-wide layer -> registers: thread t takes the float4 chunks (row n, chunk c) with n = 8 * (j / 4 ...) laid out so
// that a quarter-warp later writes 128 contiguous bytes of a core matrix: item i = t + 256 j, n = (i & 7) + 8 * (i >> 8), c = (i >> 3) & 31
__device__ __forceinline__ void t5_weights_to_regs(const float* __restrict__ W, float4 (&wr)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = threadIdx.x + kT5Threads * j;
    const int n = (i & 7) + 8 * (i >> 8), c = (i >> 3) & 31;
    wr[j] = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * kPolHid) + c);
  }
}
__device__ __forceinline__ void t5_split4(const float4 v, float4& hi, float4& lo) {
  uint32_t h, l;
  tc_split(v.x, h, l); hi.x = __uint_as_float(h); lo.x = __uint_as_float(l);
  tc_split(v.y, h, l); hi.y = __uint_as_float(h); lo.y = __uint_as_float(l);
  tc_split(v.z, h, l); hi.z = __uint_as_float(h); lo.z = __uint_as_float(l);
  tc_split(v.w, h, l); hi.w = __uint_as_float(h); lo.w = __uint_as_float(l);
}
__device__ __forceinline__ void t5_regs_to_tiles(const float4 (&wr)[16], float* b_hi, float* b_lo) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = threadIdx.x + kT5Threads * j;
    const int n = (i & 7) + 8 * (i >> 8), c = (i >> 3) & 31;
    float4 hi, lo;
    t5_split4(wr[j], hi, lo);
    const int off = t5_off(n, 4 * c);
    *reinterpret_cast<float4*>(b_hi + off) = hi;
    *reinterpret_cast<float4*>(b_lo + off) = lo;
  }
}
// [thread 0] the MMAs of one layer: ksteps k-steps of 8, three split terms each (small terms first), then the commit
__device__ __forceinline__ void t5_issue_layer(uint32_t tmem_d, const float* a_hi, const float* a_lo, const float* b_hi,
                                               const float* b_lo, int ksteps, uint64_t* bar) {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint64_t dah = t5_desc(a_hi), dal = t5_desc(a_lo), dbh = t5_desc(b_hi), dbl = t5_desc(b_lo);
  for (int j = 0; j < ksteps; ++j) {
    const uint64_t o = (uint64_t)((256u * (uint32_t)j) >> 4);       // 256 bytes per k-step, in the descriptor's 16-byte units
    t5_mma(tmem_d, dal + o, dbh + o, j > 0 ? 1u : 0u);
    t5_mma(tmem_d, dah + o, dbl + o, 1u);
    t5_mma(tmem_d, dah + o, dbh + o, 1u);
  }
  t5_commit(bar);
}

__global__ void __launch_bounds__(kT5Threads, 1) zbot_policy_act_tc5_kernel(const PolicyArgs a) {
  extern __shared__ __align__(1024) uint8_t t5_raw[];   // (the no-swizzle descriptors only need 16-byte alignment)
  float* a_hi = reinterpret_cast<float*>(t5_raw);        // no integer round trip: the compiler keeps the shared address space (STS / LDS)
  float* a_lo = a_hi + kT5TileA;
  float* b_hi = a_lo + kT5TileA;
  float* b_lo = b_hi + kT5TileB;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_sm;
  const int net = blockIdx.y;
  const int e0 = blockIdx.x * kPolTile;
  const int valid = min(kPolTile, a.n - e0);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int K8 = (a.num_obs + 7) & ~7;
  // ---- TMEM: 128 columns (one FP32 accumulator column per output neuron), allocated by warp 0 ----
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_sm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) mbar_init(&mma_bar, 1);
  // ---- first layer's operands: observation tile -> A tiles (zero beyond num_obs / dead envs), W0 (128 x num_obs) -> B tiles ----
  // head weights / bias: one float4 per thread at most, parked in registers until the layers are done
  const int nout = net == 0 ? a.num_actions : 1;
  float4 hw_reg = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < nout * (kPolHid / 4)) hw_reg = __ldg(reinterpret_cast<const float4*>(a.w[net][3]) + t);
  const float hb_reg = (t < nout) ? __ldg(a.b[net][3] + t) : 0.f;
  {
    // observation tile: warp w takes envs w, w + 8, ...; lane = column (+ 32 for a wide observation): 16 loads in flight
    float ov[kPolTile / (kT5Threads / 32)][2];
#pragma unroll
    for (int i = 0; i < kPolTile / (kT5Threads / 32); ++i) {
      const int m = wid + (kT5Threads / 32) * i;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = lane + 32 * hh;
        ov[i][hh] = (m < valid && k < a.num_obs) ? __ldg(a.obs + (size_t)(e0 + m) * a.num_obs + k) : 0.f;
      }
    }
    // W0 (128 x num_obs, rows not aligned): two threads per row, alternating columns, four loads per batch
    const int wn = t >> 1, wh = t & 1;
    const float* wrow = a.w[net][0] + (size_t)wn * a.num_obs;
    for (int k0 = wh; k0 < K8; k0 += 8) {
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = (k0 + 2 * u < a.num_obs) ? __ldg(wrow + k0 + 2 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 2 * u;
        if (k < K8) {
          uint32_t h, l;
          tc_split(wv[u], h, l);
          b_hi[t5_off(wn, k)] = __uint_as_float(h);
          b_lo[t5_off(wn, k)] = __uint_as_float(l);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kPolTile / (kT5Threads / 32); ++i) {
      const int m = wid + (kT5Threads / 32) * i;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = lane + 32 * hh;
        if (k < K8) {
          if (net == 0 && a.obs_out && m < valid && k < a.num_obs) a.obs_out[(size_t)(e0 + m) * a.num_obs + k] = ov[i][hh];
          uint32_t h, l;
          tc_split(ov[i][hh], h, l);
          a_hi[t5_off(m, k)] = __uint_as_float(h);
          a_lo[t5_off(m, k)] = __uint_as_float(l);
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_sm;
  if (t == 0) t5_issue_layer(tmem_d, a_hi, a_lo, b_hi, b_lo, K8 >> 3, &mma_bar);

  const int q = wid & 3, ch = wid >> 2;                  // lane quadrant, column half
  const int rowA = 16 * q + (lane >> 2);                 // this thread's two rows (TMEM lanes 32 q + lane / 4 and + 8) ...
  const int colq = 64 * ch + 2 * (lane & 3);             // ... and its column pair inside every group of 8 columns
  const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * ch);
  float* xs = b_hi;                                      // after the last layer: final activations [64][132] (the B tiles are dead)
#pragma unroll 1
  for (int layer = 0; layer < 3; ++layer) {
    float4 wr[16];
    if (layer < 2) t5_weights_to_regs(a.w[net][layer + 1], wr);        // in flight across the wait below
    float2 bias2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias2[j] = make_float2(__ldg(a.b[net][layer] + colq + 8 * j), __ldg(a.b[net][layer] + colq + 8 * j + 1));
    mbar_wait(&mma_bar, (uint32_t)(layer & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: TMEM -> bias + ELU -> next layer's A tiles (or the plain activation tile for the head) ----
    {
      float v[32];
      t5_ld16x64(taddr, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = colq + 8 * j;
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int row = rowA + 8 * hrow;
          const float o0 = pol_elu(v[4 * j + 2 * hrow] + bias2[j].x), o1 = pol_elu(v[4 * j + 2 * hrow + 1] + bias2[j].y);
          if (layer < 2) {
            uint32_t h0, l0, h1, l1;
            tc_split(o0, h0, l0);
            tc_split(o1, h1, l1);
            const int off = t5_off(row, col);
            *reinterpret_cast<float2*>(a_hi + off) = make_float2(__uint_as_float(h0), __uint_as_float(h1));
            *reinterpret_cast<float2*>(a_lo + off) = make_float2(__uint_as_float(l0), __uint_as_float(l1));
          } else {
            *reinterpret_cast<float2*>(xs + row * kTcXS + col) = make_float2(o0, o1);
          }
        }
      }
    }
    if (layer < 2) {
      t5_regs_to_tiles(wr, b_hi, b_lo);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (t == 0) t5_issue_layer(tmem_d, a_hi, a_lo, b_hi, b_lo, kPolHid >> 3, &mma_bar);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
  // ---- head, sampling, stores: the FP32 code of the other two kernels; scratch lives in the dead B_lo tile ----
  float* outs = b_lo;                                    // [4][64][8] partial head outputs
  float* hws = outs + 4 * kPolTile * kPolMaxAct;         // [8][128] head weights, [8] head bias
  float* lpc = hws + kPolMaxAct * kPolHid + kPolMaxAct;  // [64][8] log-prob terms
  if (t < nout * (kPolHid / 4)) reinterpret_cast<float4*>(hws)[t] = hw_reg;
  if (t < nout) hws[kPolMaxAct * kPolHid + t] = hb_reg;
  __syncthreads();
  constexpr int kG = kT5Threads / 64;
  const int m = t & 63, kq = t >> 6;
  {
    float part[kPolMaxAct];
#pragma unroll
    for (int o = 0; o < kPolMaxAct; ++o) part[o] = 0.f;
    const int kb = kq * (kPolHid / kG);
#pragma unroll 2
    for (int k = kb; k < kb + kPolHid / kG; k += 4) {
      const float4 x = *reinterpret_cast<const float4*>(xs + m * kTcXS + k);
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
  const unsigned long long call = a.ctr ? __ldcg(a.ctr) : a.call;
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
