// FP32 issue-rate probe for sm_100a (B200): what can one SM sub-partition (SMSP) actually issue per clock?
//   scalar FFMA with 3 register operands, FFMA with immediate operands, packed FFMA2, and an FFMA + integer mix.
// Evidence for DESIGN.md §4 ("the fused step is at the measured FP32 issue ceiling").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 32768, ACC = 8;

template <int MODE>
__global__ void probe(float* out, float a, float b, int ia, long long* cyc) {
  const long long t0 = clock64();
  float x[ACC];
  float2 y[ACC];
  int z[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = make_float2(x[i], x[i] * 0.5f); z[i] = threadIdx.x + i; }
  const float m = a + threadIdx.x * 1e-9f, c = b;
  const float2 m2 = make_float2(m, a), c2 = make_float2(b, b * 0.5f);
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], m, c);                       // FFMA R, R, R, R
      if (MODE == 1) x[i] = fmaf(x[i], 0.99990f, 0.5f);             // FFMA R, R, imm, imm
      if (MODE == 2) y[i] = __ffma2_rn(y[i], m2, c2);               // FFMA2 (two FP32 FMAs per lane)
      if (MODE == 3) { if (i & 1) z[i] = (z[i] ^ ia) + it; else x[i] = fmaf(x[i], m, c); }   // FFMA / integer 1:1
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += x[i] + y[i].x + y[i].y + (float)z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)cyc, (unsigned long long)(t1 - t0));   // SM clock cycles, slowest warp
}

template <int MODE>
void run(const char* name, float* out, double flops_per_inst) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 1; w <= 8; w *= 2) {
    const int threads = 32 * 4 * w;   // one CTA per SM, w warps per SMSP
    float ms = 0, best = 1e9f;
    long long* cyc;
    cudaMalloc(&cyc, 8);
    long long hc = 0, best_c = 1LL << 60;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(cyc, 0, 8);
      cudaEventRecord(e0);
      probe<MODE><<<148, threads>>>(out, 0.999f, 0.001f, 12345, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
      if (rep > 0 && ms < best) { best = ms; best_c = hc; }
    }
    cudaFree(cyc);
    const double inst = (double)ITER * ACC * w;               // warp instructions per SMSP (loop overhead ~3 / 8 more)
    const double cycles = (double)best_c;                     // measured with clock64 inside the kernel
    printf("%-26s warps/SMSP %d: %7.3f ms  %.3f warp-inst/clk/SMSP", name, w, best, inst / cycles);
    printf("  (%.0f MHz)", cycles / (best * 1e-3) / 1e6);
    if (flops_per_inst > 0) printf("  %.1f TFLOP/s", flops_per_inst * 32 * inst * 4 * 148 / (best * 1e-3) / 1e12);
    printf("\n");
  }
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  run<0>("FFMA reg,reg,reg", out, 2);
  run<1>("FFMA reg,imm,imm", out, 2);
  run<2>("FFMA2 (packed 2 x FP32)", out, 4);
  run<3>("FFMA : integer 1:1", out, 0);
  return 0;
}
