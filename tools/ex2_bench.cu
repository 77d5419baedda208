// Micro-benchmark: MUFU ex2 throughput, f32 vs packed f16x2 vs bf16x2, and an FMA-pipe polynomial exp2.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  unsigned h0 = 0x3c003c00u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    } else {
      // degree-3 polynomial 2^x on the FMA pipe (x <= 0): split integer / fraction with the magic-add trick
      float* v[4] = {&a0, &a1, &a2, &a3};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = fmaxf(*v[j], -126.f);
        float r = x + 12582912.f;           // round to nearest integer
        float fl = r - 12582912.f;
        float f = x - fl;                   // in [-0.5, 0.5]
        float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
        int e = __float_as_int(r) << 23;    // integer part into the exponent field
        *v[j] = __int_as_float(__float_as_int(p) + e) - 1.5f;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 4; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 1024>>>(d, iters);
      if (mode == 1) k<1><<<148 * 8, 1024>>>(d, iters);
      if (mode == 2) k<2><<<148 * 8, 1024>>>(d, iters);
      if (mode == 3) k<3><<<148 * 8, 1024>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * 8 * 1024 * iters * 4 * (mode == 1 || mode == 2 ? 2 : 1);
    printf("mode %d: %.3f ms, %.2f Texp/s, %.1f exp/clk/SM @1.9GHz\n", mode, ms, n / ms / 1e9, n / ms / 1e6 / 148 / 1.9e3 * 1e-3 * 1e3);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
