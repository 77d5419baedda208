// Shared host/device helpers for the tiny-SD B200 library.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace tsd {

typedef __nv_bfloat16 bf16;

// Error plumbing for the C ABI: every entry point returns 0 on success, non-zero otherwise, and
// tsd_last_error() returns the message.
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define TSD_CHECK(cond, ...)        \
  do {                              \
    if (!(cond)) {                  \
      tsd::set_error(__VA_ARGS__);  \
      return 1;                     \
    }                               \
  } while (0)

#define TSD_CUDA(expr)                                   \
  do {                                                   \
    if (tsd::check_cuda((expr), #expr)) return 2;        \
  } while (0)

void count_launch();
#define TSD_LAUNCH_CHECK()          \
  do {                              \
    tsd::count_launch();            \
    TSD_CUDA(cudaGetLastError());   \
  } while (0)

int num_sms();
// "done once" flag per CUDA device (cudaFuncSetAttribute and friends are per device, not per process)
struct PerDeviceFlag {
  bool done[64] = {};
  bool& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return done[d & 63];
  }
};

// Tensor-map (TMA descriptor) builders; bf16 / f32 elements, SWIZZLE_128B, zero OOB fill.
// 2-D: tensor [rows][cols] with cols contiguous; box = box_cols x box_rows.
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows);
// Same with a 32 / 64 / 128-byte swizzle; the inner box spans exactly the swizzle width.
int make_tmap_2d_sw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                    uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes);
// 4-D NHWC activation [N][H][W][C]; box = (box_c, bw, bh, bn) elements *loaded*; traversal stride s
// along W and H (s = 2 gives the stride-2 convolution gather).
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W,
                   uint64_t C, uint32_t box_c, uint32_t bw, uint32_t bh, uint32_t bn, uint32_t s);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// silu(x) = x * sigmoid(x) = 0.5 x (1 + tanh(x/2)): ONE MUFU op per element (tanh.approx, rel. err ~2^-11, well
// below the bf16 rounding of the result) instead of ex2 + rcp -- the norm kernels are otherwise MUFU-limited.
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_fast(h), h);
}
// erf(x), |abs err| < 2e-7 (Abramowitz-Stegun 7.1.26): one MUFU.EX2 + one MUFU.RCP + 7 FMA instead of libm erff
__device__ __forceinline__ float erf_fast(float x) {
  const float a = fabsf(x);
  const float t = rcp_fast(fmaf(0.3275911f, a, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.f - p * t * __expf(-a * a);
  return copysignf(r, x);
}
// d/dx silu(x) = s + x*s*(1-s), s = sigmoid(x) = 0.5 (1 + tanh(x/2))
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
  return s * fmaf(x, 1.f - s, 1.f);
}
__device__ __forceinline__ float gelu_f(float x) {  // exact erf form (F.gelu default)
  return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752440f));
}
// value * gelu(gate) with ONE MUFU op: erf(u) ~= tanh(u (a + b u^2)), u = g / sqrt(2), (a, b) fitted for the minimax
// error of gelu itself (2.7e-4 absolute; tanh.approx adds <= 2.5e-4 |g|) -- an order of magnitude below the bf16
// rounding of the product.  Used where the activation sits in a GEMM epilogue and instruction issue is the bound
// (inference GEGLU, diffusion.py:151-152); the stand-alone training kernels keep the erf form above.
__device__ __forceinline__ float geglu_fast_f(float x, float g) {
  const float t = g * g;
  const float z = g * fmaf(0.0347008941f, t, 0.8001570768f);  // a / sqrt(2), b / (2 sqrt(2))
  const float hg = 0.5f * x * g;
  return fmaf(hg, tanh_fast(z), hg);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf = 0.5f * (1.f + erf_fast(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Philox4x32-10 counter RNG (same generator family torch's CUDA RNG uses; the stream layout is ours).
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const { return rounds<10>(ctr_lo, ctr_hi); }
  // Philox4x32-R: R = 10 is the standard generator (Gaussian noise); R = 7 is the smallest variant that passes
  // BigCrush (Salmon et al., SC'11) and is used for dropout masks, where the generator is the ALU cost of the kernel.
  template <int R>
  __device__ __forceinline__ uint4 rounds(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi,
             c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// Device-resident part of a Philox stream position: {number of calls so far, global index of this rank's first sample}.
// Kernels that draw random numbers take an optional pointer to it, so that (a) a captured CUDA graph draws fresh
// numbers on every replay (the counter is advanced by a kernel inside the graph) and (b) the numbers of a sample depend
// on its GLOBAL index only -- a batch sharded over N ranks draws exactly what one rank would.
struct RngPos {
  uint64_t calls, sample0;
};
__device__ __forceinline__ RngPos load_rng_pos(const uint64_t* p) {
  RngPos r{0ull, 0ull};
  if (p) { r.calls = p[0]; r.sample0 = p[1]; }
  return r;
}
// Two uniform u32 -> two N(0,1) via Box-Muller.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = (a + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
  float u2 = (b + 0.5f) * 2.3283064365386963e-10f;
  float r = sqrtf(-2.f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

}  // namespace tsd
