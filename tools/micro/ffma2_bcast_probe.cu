// Packed-FP32 operand-form probe for sm_100a (B200): does the scalar-broadcast operand form of FFMA2
// (`FFMA2 Rd, Ra.F32, Rb.F32x2.HI_LO, Rc.F32x2.HI_LO`, what a register-tiled FP32 GEMM uses: acc2 += x * (w0, w1)) issue at the
// same rate as the plain pair form?  16 independent accumulator pairs per thread, outer-product pattern 4 x 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bcast_probe ffma2_bcast_probe.cu && ./ffma2_bcast_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 16384;

template <int MODE>
__global__ void probe(float* out, float a, float b, long long* cyc) {
  float2 acc[4][4];
  float x[4];
  float2 xp[4], w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = a + threadIdx.x * 1e-6f + i;
    xp[i] = make_float2(x[i], x[i] + b);
    w[i] = make_float2(b + i, b * 0.5f + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(i, j);
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 0) acc[i][j] = __ffma2_rn(make_float2(x[i], x[i]), w[j], acc[i][j]);   // scalar-broadcast form
        if (MODE == 1) acc[i][j] = __ffma2_rn(xp[i], w[j], acc[i][j]);                     // pair form
        if (MODE == 2) { acc[i][j].x = fmaf(x[i], w[j].x, acc[i][j].x); acc[i][j].y = fmaf(x[i], w[j].y, acc[i][j].y); }   // 2 x FFMA
      }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
}

template <int MODE>
void run(const char* name, float* out, int fma_inst_per_iter, double flop_per_inst) {
  for (int w = 1; w <= 4; w *= 2) {
    long long* cyc;
    cudaMalloc(&cyc, 8);
    long long hc = 0, best = 1LL << 60;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(cyc, 0, 8);
      probe<MODE><<<148, 128 * w>>>(out, 0.999f, 0.001f, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
      if (rep > 0 && hc < best) best = hc;
    }
    cudaFree(cyc);
    const double inst = (double)ITER * fma_inst_per_iter * w;
    printf("%-34s warps/SMSP %d: %.3f FMA-inst/clk/SMSP  %.1f FMA/clk/SM\n", name, w, inst / best, flop_per_inst / 2 * 32 * 4 * inst / best);
  }
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  run<0>("FFMA2 scalar-broadcast operand", out, 16, 4);
  run<1>("FFMA2 pair operands", out, 16, 4);
  run<2>("2 x FFMA", out, 32, 2);
  return 0;
}
