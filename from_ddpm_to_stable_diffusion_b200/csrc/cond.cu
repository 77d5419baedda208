// Conditioning path (tiny, fp32 CUDA-core work): sinusoidal timestep embedding, label embedding
// lookup and the small-M linears (time / label MLPs, per-ResBlock linear_time, the degenerate
// cross-attention v_proj -> out_proj vectors).  Weights are read directly in the reference's
// fp32 [out][in] layout, no packing.
//
// Reference: TimestepEmbedder diffusion.py:13-37; label_embedding :196-201; linear_time :101-104,
// :112; CrossAttention :61-82 (one key/value token => out_proj(v_proj(ctx)) per sample).
#include "../../include/tinysd_b200.h"
#include "common.cuh"
#include <cstring>

using namespace tsd;

namespace {

constexpr int ROWS = 8;  // rows of x handled per CTA (weight rows are re-used across them)

// out[m][n] = sum_k f(x[m][k]) * w[n][k] + b[n],  f = SiLU if silu_in.  grid = (ceil(N/64), ceil(M/ROWS))
__device__ __forceinline__ void small_linear_fwd_body(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out, int M, int N,
                                                      int K, int silu_in, int bx, int by) {
  extern __shared__ float sx[];  // [ROWS][K]
  const int m0 = by * ROWS;
  for (int i = threadIdx.x; i < ROWS * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    float v = (m0 + r) < M ? x[(size_t)(m0 + r) * K + k] : 0.f;
    if (silu_in) v = v / (1.f + expf(-v));
    sx[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = bx * 64 + warp; n < min(N, bx * 64 + 64); n += 8) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = w[(size_t)n * K + k];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] += wv * sx[r * K + k];
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bv = bias ? bias[n] : 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
        if (m0 + r < M) out[(size_t)(m0 + r) * N + n] = acc[r] + bv;
    }
  }
}

__global__ void __launch_bounds__(256) small_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ out,
                                                               int M, int N, int K, int silu_in) {
  small_linear_fwd_body(x, w, bias, out, M, N, K, silu_in, blockIdx.x, blockIdx.y);
}

// A group of small linears in ONE launch (blockIdx.z = entry): the 14 linear_time layers, the 10 cross-attention
// v_proj and the 10 out_proj of a step share M but are separate parameters; launched one by one each is a 20-40 us
// latency-bound kernel on a handful of CTAs.  Entries may differ in N and K; pointers are passed by value.
constexpr int SLB_MAX = 16;
struct SmallLinearBatch {
  const float* x[SLB_MAX];     // forward input [M][K_i] (also needed by the backward)
  const float* w[SLB_MAX];     // [N_i][K_i]
  const float* bias[SLB_MAX];  // forward: bias or null
  float* out[SLB_MAX];         // forward: [M][N_i]
  const float* dy[SLB_MAX];    // backward: [M][N_i]
  float* dx[SLB_MAX];          // backward: [M][K_i] or null; entries may alias (accumulated with atomics)
  float* dw[SLB_MAX];          // backward: [N_i][K_i] (+=) or null
  float* db[SLB_MAX];          // backward: [N_i] (+=) or null
  int N[SLB_MAX], K[SLB_MAX];
  int n, M, silu_in, accumulate_dx;
};
__global__ void __launch_bounds__(256) small_linear_many_fwd_kernel(const __grid_constant__ SmallLinearBatch b) {
  const int e = blockIdx.z;
  if ((int)blockIdx.x * 64 >= b.N[e]) return;
  small_linear_fwd_body(b.x[e], b.w[e], b.bias[e], b.out[e], b.M, b.N[e], b.K[e], b.silu_in, blockIdx.x, blockIdx.y);
}

// dx[m][k] (+)= f'(x[m][k]) * sum_n dy[m][n] * w[n][k].  grid = (ceil(K/256), ceil(M/ROWS))
__global__ void __launch_bounds__(64) small_linear_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                 const float* __restrict__ x, float* __restrict__ dx,
                                                                 int M, int N, int K, int silu_in, int accumulate) {
  extern __shared__ float sdy[];  // [ROWS][N]
  const int m0 = blockIdx.y * ROWS;
  for (int i = threadIdx.x; i < ROWS * N; i += blockDim.x) {
    const int r = i / N, n = i - r * N;
    sdy[i] = (m0 + r) < M ? dy[(size_t)(m0 + r) * N + n] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
  int n = 0;
  for (; n + 8 <= N; n += 8) {  // eight weight rows in flight
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (size_t)(n + u) * K + k);  // coalesced over k
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(wv[u], sdy[r * N + n + u], acc[r]);
  }
  for (; n < N; ++n) {
    const float wv = w[(size_t)n * K + k];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] += wv * sdy[r * N + n];
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (m0 + r >= M) break;
    float v = acc[r];
    if (silu_in) {
      const float xv = x[(size_t)(m0 + r) * K + k];
      const float s = 1.f / (1.f + expf(-xv));
      v *= s * (1.f + xv * (1.f - s));
    }
    float* p = dx + (size_t)(m0 + r) * K + k;
    *p = accumulate ? *p + v : v;
  }
}

// dw[n][k] += sum_m dy[m][n] * f(x[m][k]);  db[n] += sum_m dy[m][n].  grid = (ceil(K/256), ceil(N/NT)):
// a thread owns one k and NT consecutive n, so f(x[m][k]) is evaluated once per (m, k) and dy rows are broadcast.
constexpr int NT = 16;
__global__ void __launch_bounds__(64) small_linear_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 float* __restrict__ dw, float* __restrict__ db, int M,
                                                                 int N, int K, int silu_in) {
  const int n0 = blockIdx.y * NT;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float sdy[32][NT];
  float acc[NT], accb[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) { acc[j] = 0.f; accb[j] = 0.f; }
  for (int m0 = 0; m0 < M; m0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * NT; i += blockDim.x) {
      const int mm = i / NT, j = i - mm * NT;
      sdy[mm][j] = (m0 + mm < M && n0 + j < N) ? dy[(size_t)(m0 + mm) * N + n0 + j] : 0.f;
    }
    __syncthreads();
    if (k < K) {
      // all 32 activations of the chunk are requested before the first is used (the loop was bound by the latency of
      // one dependent global load per row)
      float xv[32];
#pragma unroll
      for (int mm = 0; mm < 32; ++mm) xv[mm] = (m0 + mm < M) ? __ldg(x + (size_t)(m0 + mm) * K + k) : 0.f;
#pragma unroll
      for (int mm = 0; mm < 32; ++mm) {
        float v = xv[mm];
        if (silu_in) v = v / (1.f + expf(-v));
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] = fmaf(sdy[mm][j], v, acc[j]);
      }
    }
    if (db && blockIdx.x == 0 && threadIdx.x < NT) {
      const int mend = min(32, M - m0);
      for (int mm = 0; mm < mend; ++mm) accb[0] += sdy[mm][threadIdx.x];
    }
  }
  if (k < K) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (n0 + j < N) dw[(size_t)(n0 + j) * K + k] += acc[j];
  }
  if (db && blockIdx.x == 0 && threadIdx.x < NT && n0 + threadIdx.x < N) db[n0 + threadIdx.x] += accb[0];
}

// ------------------------------------------------------------------------------------------
// Tiled fp32 versions of the two kernels above for batch-sized M (>= 32): a CTA of 256 threads owns a 32 x 64 output
// tile (2 x 4 per thread) and walks the reduction dimension in steps of 32 through shared memory, so every operand
// element is read from global memory once per tile and from shared memory once per 4-8 FMAs (the row-per-thread
// kernels read one shared-memory word per FMA and run on 64-128 tiny CTAs).
//   dgrad: dx[m][k] (+)= f'(x[m][k]) * sum_n dy[m][n] * w[n][k]      tile = 32 m x 64 k, reduce over n
//   wgrad: dw[n][k] += sum_m dy[m][n] * f(x[m][k]); db[n] += sum_m dy[m][n]   tile = 32 n x 64 k, reduce over m
// ------------------------------------------------------------------------------------------
constexpr int TI = 32, TJ = 64, TR = 32;

template <int WGRAD, int ATOMIC>
__device__ __forceinline__ void small_linear_bwd_tiled_body(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ w, float* __restrict__ dx,
                                                            float* __restrict__ dw, float* __restrict__ db, int M, int N,
                                                            int K, int silu_in, int accumulate, int bx, int by) {
  // A(r, i): the dy operand; B(r, j): f(x) (wgrad) or w (dgrad).  sA[r][i], sB[r][j]
  __shared__ float sA[TR][TI + 1];
  __shared__ float sB[TR][TJ];
  const int i0 = by * TI;  // wgrad: n; dgrad: m
  const int j0 = bx * TJ;  // k
  const int R = WGRAD ? M : N;     // reduction length
  const int ti = (threadIdx.x >> 4) * 2, tj = (threadIdx.x & 15) * 4;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float accb[2] = {0.f, 0.f};
  for (int r0 = 0; r0 < R; r0 += TR) {
    __syncthreads();
    for (int e = threadIdx.x; e < TR * TI; e += blockDim.x) {
      // wgrad: A(r, i) = dy[m = r][n = i] (i fastest: coalesced); dgrad: A(r, i) = dy[m = i][n = r] (r fastest)
      const int a = WGRAD ? e / TI : e % TR, bq = WGRAD ? e % TI : e / TR;  // (r, i) local
      const int r = r0 + a, i = i0 + bq;
      const int m = WGRAD ? r : i, n = WGRAD ? i : r;
      sA[a][bq] = (m < M && n < N) ? dy[(size_t)m * N + n] : 0.f;
    }
    for (int e = threadIdx.x; e < TR * TJ; e += blockDim.x) {
      const int a = e / TJ, bq = e - a * TJ;
      const int r = r0 + a, k = j0 + bq;
      float v = 0.f;
      if (k < K && r < R) {
        if (WGRAD) {
          v = x[(size_t)r * K + k];
          if (silu_in) v = v / (1.f + expf(-v));
        } else {
          v = w[(size_t)r * K + k];
        }
      }
      sB[a][bq] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int a = 0; a < TR; ++a) {
      const float a0 = sA[a][ti], a1 = sA[a][ti + 1];
      const float4 b = *reinterpret_cast<const float4*>(&sB[a][tj]);
      acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
      if (WGRAD) { accb[0] += a0; accb[1] += a1; }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + ti + u;
    if (i >= (WGRAD ? N : M)) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = j0 + tj + v;
      if (k >= K) continue;
      if (WGRAD) {
        dw[(size_t)i * K + k] += acc[u][v];
      } else {
        float val = acc[u][v];
        if (silu_in) {
          const float xv = x[(size_t)i * K + k];
          const float sg = 1.f / (1.f + expf(-xv));
          val *= sg * (1.f + xv * (1.f - sg));
        }
        float* pd = dx + (size_t)i * K + k;
        if (ATOMIC && accumulate) atomicAdd(pd, val);  // several entries of a batch add into the same dx
        else *pd = accumulate ? *pd + val : val;
      }
    }
    if (WGRAD && db && bx == 0 && (threadIdx.x & 15) == 0) db[i] += accb[u];
  }
}
template <int WGRAD>
__global__ void __launch_bounds__(256) small_linear_bwd_tiled_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                     const float* __restrict__ w, float* __restrict__ dx,
                                                                     float* __restrict__ dw, float* __restrict__ db, int M,
                                                                     int N, int K, int silu_in, int accumulate) {
  small_linear_bwd_tiled_body<WGRAD, 0>(dy, x, w, dx, dw, db, M, N, K, silu_in, accumulate, blockIdx.x, blockIdx.y);
}
// grid = (ceil(max K / TJ), ceil(max(N or M) / TI), entries)
template <int WGRAD>
__global__ void __launch_bounds__(256) small_linear_many_bwd_kernel(const __grid_constant__ SmallLinearBatch b) {
  const int e = blockIdx.z;
  if ((int)blockIdx.x * TJ >= b.K[e] || (int)blockIdx.y * TI >= (WGRAD ? b.N[e] : b.M)) return;
  if (WGRAD ? b.dw[e] == nullptr : b.dx[e] == nullptr) return;
  small_linear_bwd_tiled_body<WGRAD, 1>(b.dy[e], b.x[e], b.w[e], b.dx[e], b.dw[e], b.db[e], b.M, b.N[e], b.K[e], b.silu_in,
                                        b.accumulate_dx, blockIdx.x, blockIdx.y);
}

// emb[m][0:half] = cos(t*f_i), emb[m][half:] = sin(t*f_i), f_i = exp(-ln(10000) * i / half)  (diffusion.py:24-28)
// freqs are passed in (built by the host shell with the reference's own torch expression).
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ emb, int M, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * half) return;
  const int m = i / half, j = i - m * half;
  const float a = (float)t[m] * freqs[j];
  emb[(size_t)m * 2 * half + j] = cosf(a);
  emb[(size_t)m * 2 * half + half + j] = sinf(a);
}

__global__ void embedding_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table,
                                     float* __restrict__ out, int M, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int m = i / D, d = i - m * D;
  out[i] = table[(size_t)idx[m] * D + d];
}
// dtable[idx[m]] += dy[m]; padding row (padding_idx) receives no gradient (diffusion.py:197)
__global__ void embedding_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dy,
                                     float* __restrict__ dtable, int M, int D, int padding_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int m = i / D, d = i - m * D;
  const int64_t r = idx[m];
  if (r == padding_idx) return;
  atomicAdd(&dtable[(size_t)r * D + d], dy[i]);
}

}  // namespace

extern "C" int tsd_small_linear_fwd(void* stream, const float* x, const float* w, const float* bias, float* out, int M,
                                    int N, int K, int silu_in) {
  TSD_CHECK(K <= 4096, "small_linear_fwd: K=%d too large", K);
  dim3 grid(ceil_div(N, 64), ceil_div(M, ROWS));
  small_linear_fwd_kernel<<<grid, 256, ROWS * K * sizeof(float), (cudaStream_t)stream>>>(x, w, bias, out, M, N, K, silu_in);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_small_linear_bwd(void* stream, const float* dy, const float* x, const float* w, float* dx, float* dw,
                                    float* db, int M, int N, int K, int silu_in, int accumulate_dx) {
  TSD_CHECK(N <= 4096, "small_linear_bwd: N=%d too large", N);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tiled = M >= 32;  // batch-sized M: the shared-memory tiled kernels
  if (dx) {
    if (tiled) {
      small_linear_bwd_tiled_kernel<0><<<dim3(ceil_div(K, TJ), ceil_div(M, TI)), 256, 0, st>>>(dy, x, w, dx, nullptr, nullptr, M, N,
                                                                                              K, silu_in, accumulate_dx);
    } else {
      dim3 grid(ceil_div(K, 64), ceil_div(M, ROWS));
      small_linear_dgrad_kernel<<<grid, 64, ROWS * N * sizeof(float), st>>>(dy, w, x, dx, M, N, K, silu_in, accumulate_dx);
    }
    TSD_LAUNCH_CHECK();
  }
  if (dw) {
    if (tiled) {
      small_linear_bwd_tiled_kernel<1><<<dim3(ceil_div(K, TJ), ceil_div(N, TI)), 256, 0, st>>>(dy, x, w, nullptr, dw, db, M, N, K,
                                                                                              silu_in, 0);
    } else {
      dim3 grid(ceil_div(K, 64), ceil_div(N, NT));
      small_linear_wgrad_kernel<<<grid, 64, 0, st>>>(dy, x, dw, db, M, N, K, silu_in);
    }
    TSD_LAUNCH_CHECK();
  }
  return 0;
}
static int slb_fill(SmallLinearBatch& b, int n, const float* const* x, const float* const* w, const int* N, const int* K,
                    int M, int silu_in) {
  TSD_CHECK(n >= 1 && n <= SLB_MAX, "small_linear_many: %d entries (1..%d)", n, SLB_MAX);
  memset(&b, 0, sizeof(b));
  b.n = n; b.M = M; b.silu_in = silu_in;
  for (int i = 0; i < n; ++i) {
    TSD_CHECK(N[i] > 0 && K[i] > 0 && K[i] <= 4096 && N[i] <= 4096, "small_linear_many: entry %d has N=%d K=%d", i, N[i], K[i]);
    b.x[i] = x[i]; b.w[i] = w[i]; b.N[i] = N[i]; b.K[i] = K[i];
  }
  return 0;
}
extern "C" int tsd_small_linear_many_fwd(void* stream, int n, const float* const* x, const float* const* w,
                                         const float* const* bias, float* const* out, const int* N, const int* K, int M,
                                         int silu_in) {
  SmallLinearBatch b;
  if (slb_fill(b, n, x, w, N, K, M, silu_in)) return 1;
  int maxN = 0, maxK = 0;
  for (int i = 0; i < n; ++i) { b.bias[i] = bias ? bias[i] : nullptr; b.out[i] = out[i]; maxN = max(maxN, N[i]); maxK = max(maxK, K[i]); }
  dim3 grid(ceil_div(maxN, 64), ceil_div(M, ROWS), n);
  small_linear_many_fwd_kernel<<<grid, 256, ROWS * maxK * sizeof(float), (cudaStream_t)stream>>>(b);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_small_linear_many_bwd(void* stream, int n, const float* const* dy, const float* const* x,
                                         const float* const* w, float* const* dx, float* const* dw, float* const* db,
                                         const int* N, const int* K, int M, int silu_in, int accumulate_dx) {
  TSD_CHECK(M >= 32, "small_linear_many_bwd: batch-sized M only (M=%d); use tsd_small_linear_bwd per layer", M);
  SmallLinearBatch b;
  if (slb_fill(b, n, x, w, N, K, M, silu_in)) return 1;
  b.accumulate_dx = accumulate_dx;
  int maxN = 0, maxK = 0;
  bool any_dx = false, any_dw = false;
  for (int i = 0; i < n; ++i) {
    b.dy[i] = dy[i];
    b.dx[i] = dx ? dx[i] : nullptr; b.dw[i] = dw ? dw[i] : nullptr; b.db[i] = db ? db[i] : nullptr;
    any_dx |= b.dx[i] != nullptr; any_dw |= b.dw[i] != nullptr;
    maxN = max(maxN, N[i]); maxK = max(maxK, K[i]);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (any_dx) {
    small_linear_many_bwd_kernel<0><<<dim3(ceil_div(maxK, TJ), ceil_div(M, TI), n), 256, 0, st>>>(b);
    TSD_LAUNCH_CHECK();
  }
  if (any_dw) {
    small_linear_many_bwd_kernel<1><<<dim3(ceil_div(maxK, TJ), ceil_div(maxN, TI), n), 256, 0, st>>>(b);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}
extern "C" int tsd_timestep_embedding(void* stream, const int64_t* t, const float* freqs, float* emb, int M, int half) {
  timestep_embedding_kernel<<<ceil_div(M * half, 256), 256, 0, (cudaStream_t)stream>>>(t, freqs, emb, M, half);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_embedding_fwd(void* stream, const int64_t* idx, const float* table, float* out, int M, int D) {
  embedding_fwd_kernel<<<ceil_div(M * D, 256), 256, 0, (cudaStream_t)stream>>>(idx, table, out, M, D);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_embedding_bwd(void* stream, const int64_t* idx, const float* dy, float* dtable, int M, int D,
                                 int padding_idx) {
  embedding_bwd_kernel<<<ceil_div(M * D, 256), 256, 0, (cudaStream_t)stream>>>(idx, dy, dtable, M, D, padding_idx);
  TSD_LAUNCH_CHECK();
  return 0;
}
