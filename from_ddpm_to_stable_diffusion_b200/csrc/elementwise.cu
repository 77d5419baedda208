// Vectorised elementwise / small-reduction kernels on bf16 channels-last tensors.
#include "../../include/tinysd_b200.h"
#include "common.cuh"

using namespace tsd;

namespace {

__device__ __forceinline__ void load8(const bf16* p, float* e) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = d.x; e[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float* e) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
}

// out = a + b
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float x[8], y[8];
    load8(a + i * 8, x);
    load8(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    store8(out + i * 8, x);
  }
}

// GEGLU: h8 [M][2H] (value half then gate half) -> out [M][H] = value * gelu(gate)   (diffusion.py:151-152)
__global__ void geglu_fwd_kernel(const bf16* __restrict__ h8, bf16* __restrict__ out, size_t M, int H) {
  const int vec_per_row = H / 8;
  const size_t total = M * vec_per_row;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / vec_per_row;
    const int c = (int)(i - row * vec_per_row) * 8;
    float v[8], g[8];
    load8(h8 + row * 2 * H + c, v);
    load8(h8 + row * 2 * H + H + c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= gelu_f(g[j]);
    store8(out + row * H + c, v);
  }
}
// dh8 = [dout * gelu(gate), dout * value * gelu'(gate)]; optionally dbias[2H] += column sums of dh8 (the bias gradient
// of the C -> 8C linear, diffusion.py:133: a separate pass over the 8C-wide tensor otherwise).  A thread owns one
// 8-column vector of both halves and walks rows, so the sums stay in registers.  blockDim = 256, 256 % (H / 8) == 0.
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const bf16* __restrict__ h8, const bf16* __restrict__ dout,
                                                        bf16* __restrict__ dh8, size_t M, int H, size_t rows_per_cta,
                                                        float* __restrict__ dbias) {
  extern __shared__ float s_db[];  // [2H], only when dbias != null
  const int vec_per_row = H / 8;
  const int slots = blockDim.x / vec_per_row;
  const int c = (threadIdx.x % vec_per_row) * 8;
  const int slot = threadIdx.x / vec_per_row;
  if (dbias) {
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) s_db[i] = 0.f;
    __syncthreads();
  }
  float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t r0 = blockIdx.x * rows_per_cta;
  const size_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  for (size_t row = r0 + slot; row < r1; row += slots) {
    float v[8], g[8], d[8], dv[8], dg[8];
    load8(h8 + row * 2 * H + c, v);
    load8(h8 + row * 2 * H + H + c, g);
    load8(dout + row * H + c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dv[j] = d[j] * gelu_f(g[j]);
      dg[j] = d[j] * v[j] * gelu_grad_f(g[j]);
      sv[j] += dv[j];
      sg[j] += dg[j];
    }
    store8(dh8 + row * 2 * H + c, dv);
    store8(dh8 + row * 2 * H + H + c, dg);
  }
  if (dbias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_db[c + j], sv[j]);
      atomicAdd(&s_db[H + c + j], sg[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) atomicAdd(&dbias[i], s_db[i]);
  }
}

// nearest x2 upsample (diffusion.py:167): out[n][2h][2w][c] = in[n][h][w][c]
__global__ void upsample2_fwd_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int n_img, int H, int W, int C) {
  const int vec = C / 8;
  const size_t total = (size_t)n_img * 4 * H * W * vec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    size_t p = i / vec;
    const int x = (int)(p % (2 * W)); p /= (2 * W);
    const int y = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    const uint4 u = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + (y >> 1)) * W + (x >> 1)) * C + cv * 8);
    *reinterpret_cast<uint4*>(out + i * 8) = u;
  }
}
// adjoint: din[n][h][w][c] = sum of the 2x2 block of dout
__global__ void upsample2_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ din, int n_img, int H, int W, int C) {
  const int vec = C / 8;
  const size_t total = (size_t)n_img * H * W * vec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    size_t p = i / vec;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int n = (int)(p / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float e[8];
        load8(dout + (((size_t)n * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C + cv * 8, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += e[j];
      }
    store8(din + i * 8, acc);
  }
}
// zero-stuffing for the stride-2 data gradient: out[n][2h][2w][c] = (y,x both even) ? in[n][y/2][x/2][c] : 0
__global__ void zero_stuff2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int n_img, int H, int W, int C) {
  const int vec = C / 8;
  const size_t total = (size_t)n_img * 4 * H * W * vec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    size_t p = i / vec;
    const int x = (int)(p % (2 * W)); p /= (2 * W);
    const int y = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    uint4 u = make_uint4(0, 0, 0, 0);
    if (!(x & 1) && !(y & 1)) u = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + (y >> 1)) * W + (x >> 1)) * C + cv * 8);
    *reinterpret_cast<uint4*>(out + i * 8) = u;
  }
}

// Per-sample column sums: out[n][c] (+)= sum over the rows of sample n of x[row][c].
// grid = (chunks, n_img); used for bias gradients (sum over n afterwards) and the time-bias gradient.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, int rows_per_sample, int C, int CW,
                                                     int rows_per_cta, float* __restrict__ out, float* __restrict__ total) {
  // blockIdx.z selects a chunk of CW <= 2048 channels (row pitch stays C)
  extern __shared__ float s_acc[];  // [CW]
  for (int i = threadIdx.x; i < CW; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int cbase = blockIdx.z * CW;
  const int vec = CW / 8;
  const int slots = blockDim.x / vec;
  const int cv = (threadIdx.x % vec) * 8;
  const int slot = threadIdx.x / vec;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows_per_sample, r0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (slot < slots) {
    for (int r = r0 + slot; r < r1; r += slots) {
      float e[8];
      load8(x + ((size_t)n * rows_per_sample + r) * C + cbase + cv, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += e[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[cv + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CW; i += blockDim.x) {
    atomicAdd(&out[(size_t)n * C + cbase + i], s_acc[i]);
    if (total) atomicAdd(&total[cbase + i], s_acc[i]);  // sum over all samples as well (bias gradient): no second kernel
  }
}

// out[c] += sum_n in[n][c]
__global__ void reduce_rows_f32_kernel(const float* __restrict__ in, int n_rows, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int n = 0; n < n_rows; ++n) a += in[(size_t)n * C + c];
  out[c] += a;
}

// fp32 -> bf16 cast with an optional row permutation (GEGLU packing) -- weights
__global__ void cast_rows_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int cols, int geglu) {
  const size_t total = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    int sr = r;
    if (geglu) {  // packed tile t (128 rows) = value rows [64t, 64t+64) then gate rows H + [64t, 64t+64)
      const int t = r >> 7, j = r & 127, H = rows >> 1;
      sr = j < 64 ? t * 64 + j : H + t * 64 + (j - 64);
    }
    dst[i] = __float2bfloat16(src[(size_t)sr * cols + c]);
  }
}
__global__ void permute_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int t = r >> 7, j = r & 127, H = rows >> 1;
  dst[r] = src[j < 64 ? t * 64 + j : H + t * 64 + (j - 64)];
}
// OIHW fp32 [co][ci][3][3] -> packed bf16 [co][tap][ci]
__global__ void pack_conv3x3_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int co, int ci) {
  const size_t total = (size_t)co * ci * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ci);
    const int tap = (int)((i / ci) % 9);
    const int o = (int)(i / ((size_t)ci * 9));
    dst[i] = __float2bfloat16(src[((size_t)o * ci + c) * 9 + tap]);
  }
}
// OIHW fp32 [co][ci][3][3] -> data-gradient weight, packed bf16 [ci][tap'][co] with tap' = 8 - tap (mirrored taps,
// in/out channels swapped): dX = conv3x3(dY, this) runs the K-major forward path (40 % faster than reading the
// forward weight as an MN-major operand).
__global__ void pack_conv3x3_dgrad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int co, int ci) {
  const size_t total = (size_t)co * ci * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % co);
    const int tap = (int)((i / co) % 9);
    const int c = (int)(i / ((size_t)co * 9));
    dst[i] = __float2bfloat16(src[((size_t)o * ci + c) * 9 + (8 - tap)]);
  }
}
// packed fp32 gradient [co][tap][ci] -> accumulate into OIHW fp32 gradient
__global__ void unpack_conv3x3_grad_kernel(const float* __restrict__ src, float* __restrict__ dst, int co, int ci) {
  const size_t total = (size_t)co * ci * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int c = (int)((i / 9) % ci);
    const int o = (int)(i / ((size_t)ci * 9));
    dst[i] += src[((size_t)o * 9 + tap) * ci + c];
  }
}

// ------------------------------------------------------------------------------------------
// All weight packings of a step in ONE launch.  The UNet has ~140 GEMM weights; refreshing their bf16 copies one
// tensor at a time costs ~140 launch latencies per optimiser step for 0.3 GB of traffic.  A descriptor per tensor
// (built once on the host; the parameter and packed buffers do not move) turns it into a single grid-stride pass.
// ------------------------------------------------------------------------------------------
struct PackDesc {
  const float* src;
  void* dst;
  long long begin;  // first flat output element of this tensor
  int rows, cols;   // linear: [rows][cols]; conv: rows = co, cols = ci
  int kind;         // TSD_PACK_*
  int pad_;
};
__global__ void __launch_bounds__(256) pack_many_kernel(const PackDesc* __restrict__ table, int n_desc, long long total) {
  extern __shared__ long long s_begin[];  // [n_desc + 1]
  for (int i = threadIdx.x; i < n_desc; i += blockDim.x) s_begin[i] = table[i].begin;
  if (threadIdx.x == 0) s_begin[n_desc] = total;
  __syncthreads();
  int d = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!(i >= s_begin[d] && i < s_begin[d + 1])) {  // consecutive elements mostly stay in the same tensor
      int lo = 0, hi = n_desc - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_begin[mid] <= i) lo = mid; else hi = mid - 1;
      }
      d = lo;
    }
    const PackDesc t = table[d];
    const long long e = i - t.begin;
    if (t.kind == TSD_PACK_LINEAR || t.kind == TSD_PACK_LINEAR_GEGLU) {
      const int r = (int)(e / t.cols), c = (int)(e - (long long)r * t.cols);
      int sr = r;
      if (t.kind == TSD_PACK_LINEAR_GEGLU) {  // packed tile q (128 rows) = value rows [64q, 64q+64) then gate rows H + [64q, ..)
        const int q = r >> 7, j = r & 127, H = t.rows >> 1;
        sr = j < 64 ? q * 64 + j : H + q * 64 + (j - 64);
      }
      reinterpret_cast<bf16*>(t.dst)[e] = __float2bfloat16(t.src[(size_t)sr * t.cols + c]);
    } else if (t.kind == TSD_PACK_CONV3X3) {  // OIHW -> [co][tap][ci]
      const int ci = t.cols;
      const int c = (int)(e % ci), tap = (int)((e / ci) % 9), o = (int)(e / ((long long)ci * 9));
      reinterpret_cast<bf16*>(t.dst)[e] = __float2bfloat16(t.src[((size_t)o * ci + c) * 9 + tap]);
    } else if (t.kind == TSD_PACK_CONV3X3_DGRAD) {  // OIHW -> [ci][8 - tap][co]
      const int co = t.rows, ci = t.cols;
      const int o = (int)(e % co), tap = (int)((e / co) % 9), c = (int)(e / ((long long)co * 9));
      reinterpret_cast<bf16*>(t.dst)[e] = __float2bfloat16(t.src[((size_t)o * ci + c) * 9 + (8 - tap)]);
    } else {  // TSD_PACK_GEGLU_BIAS: fp32 vector, same row permutation as the GEGLU weight
      const int r = (int)e, q = r >> 7, j = r & 127, H = t.rows >> 1;
      reinterpret_cast<float*>(t.dst)[e] = t.src[j < 64 ? q * 64 + j : H + q * 64 + (j - 64)];
    }
  }
}

inline int ew_grid(size_t work_items) {
  size_t g = (work_items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int tsd_add_bf16(void* stream, const void* a, const void* b, void* out, int64_t numel) {
  TSD_CHECK(numel % 8 == 0, "add_bf16: numel must be a multiple of 8");
  add_kernel<<<ew_grid(numel / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, numel / 8);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_geglu_fwd(void* stream, const void* h8, void* out, int64_t M, int H) {
  TSD_CHECK(H % 8 == 0, "geglu_fwd: H must be a multiple of 8");
  geglu_fwd_kernel<<<ew_grid(M * (H / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)h8, (bf16*)out, M, H);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_geglu_bwd(void* stream, const void* h8, const void* dout, void* dh8, int64_t M, int H, float* dbias) {
  TSD_CHECK(H % 8 == 0 && H / 8 <= 256 && 256 % (H / 8) == 0, "geglu_bwd: H=%d (H / 8 must divide 256)", H);
  const int slots = 256 / (H / 8);
  size_t ctas = (size_t)num_sms() * 8;
  size_t rpc = ((size_t)M + ctas - 1) / ctas;
  rpc = (rpc + slots - 1) / slots * slots;
  if (rpc < (size_t)slots) rpc = slots;
  const int grid = (int)(((size_t)M + rpc - 1) / rpc);
  geglu_bwd_kernel<<<grid, 256, dbias ? 2 * H * sizeof(float) : 0, (cudaStream_t)stream>>>(
      (const bf16*)h8, (const bf16*)dout, (bf16*)dh8, (size_t)M, H, rpc, dbias);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_upsample2_fwd(void* stream, const void* in, void* out, int n_img, int H, int W, int C) {
  TSD_CHECK(C % 8 == 0, "upsample2: C must be a multiple of 8");
  upsample2_fwd_kernel<<<ew_grid((size_t)n_img * 4 * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, n_img, H, W, C);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_upsample2_bwd(void* stream, const void* dout, void* din, int n_img, int H, int W, int C) {
  TSD_CHECK(C % 8 == 0, "upsample2: C must be a multiple of 8");
  upsample2_bwd_kernel<<<ew_grid((size_t)n_img * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, (bf16*)din, n_img, H, W, C);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_zero_stuff2(void* stream, const void* in, void* out, int n_img, int H, int W, int C) {
  TSD_CHECK(C % 8 == 0, "zero_stuff2: C must be a multiple of 8");
  zero_stuff2_kernel<<<ew_grid((size_t)n_img * 4 * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, n_img, H, W, C);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_colsum(void* stream, const void* x, int n_samples, int rows_per_sample, int C, float* out,
                          float* total) {
  int CW = C;
  while (CW > 2048) CW /= 2;
  TSD_CHECK(C % 8 == 0 && C % CW == 0 && 256 % (CW / 8) == 0, "colsum: unsupported C=%d", C);
  int want = ceil_div(4 * num_sms(), n_samples * (C / CW));
  int rpc = ceil_div(rows_per_sample, want < 1 ? 1 : want);
  if (rpc < 16) rpc = 16;
  if (rpc > rows_per_sample) rpc = rows_per_sample;
  colsum_kernel<<<dim3(ceil_div(rows_per_sample, rpc), n_samples, C / CW), 256, CW * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)x, rows_per_sample, C, CW, rpc, out, total);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_reduce_rows_f32(void* stream, const float* in, int n_rows, int C, float* out) {
  reduce_rows_f32_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(in, n_rows, C, out);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_pack_linear(void* stream, const float* src, void* dst, int rows, int cols, int geglu) {
  TSD_CHECK(!geglu || rows % 256 == 0, "pack_linear: GEGLU packing needs rows %% 256 == 0");
  cast_rows_kernel<<<ew_grid((size_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, rows, cols, geglu);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_pack_geglu_bias(void* stream, const float* src, float* dst, int rows) {
  permute_vec_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_pack_conv3x3(void* stream, const float* src, void* dst, int co, int ci) {
  pack_conv3x3_kernel<<<ew_grid((size_t)co * ci * 9), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, co, ci);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_unpack_conv3x3_grad(void* stream, const float* src, float* dst, int co, int ci) {
  unpack_conv3x3_grad_kernel<<<ew_grid((size_t)co * ci * 9), 256, 0, (cudaStream_t)stream>>>(src, dst, co, ci);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_pack_conv3x3_dgrad(void* stream, const float* src, void* dst, int co, int ci) {
  pack_conv3x3_dgrad_kernel<<<ew_grid((size_t)co * ci * 9), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, co, ci);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_pack_many(void* stream, const void* table_dev, int n_desc, int64_t total) {
  TSD_CHECK(n_desc > 0 && n_desc <= 4096 && total > 0, "pack_many: n_desc=%d total=%lld", n_desc, (long long)total);
  static_assert(sizeof(PackDesc) == 40, "PackDesc layout is part of the C ABI (host side builds the table)");
  pack_many_kernel<<<ew_grid((size_t)total / 4), 256, (n_desc + 1) * sizeof(long long), (cudaStream_t)stream>>>(
      reinterpret_cast<const PackDesc*>(table_dev), n_desc, (long long)total);
  TSD_LAUNCH_CHECK();
  return 0;
}
