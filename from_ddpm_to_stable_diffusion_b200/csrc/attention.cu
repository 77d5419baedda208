// Self-attention on warp-level MMA: forward and two-pass backward for head_dim 16 / 32 / 64 (any L), and the one-pass
// backward for head_dim 16 with L % 256 == 0.  (The head_dim-16 forward of the large stages is in attention_tc.cu.)
//
// Reference: SelfAttention.forward, diffusion.py:46-58 -- softmax(q k^T / sqrt(dh)) v over 8 heads,
// q, k, v = consecutive C-wide column blocks of the in_proj output, head h = channels [h*dh, (h+1)*dh).
//
// With dh = 16 the score GEMM has K = 16 and the kernel is bound by exp throughput (MUFU), not by
// tensor throughput, so the contractions use warp-level mma.sync (m16n8k16, bf16 in,
// fp32 accumulate) with operands staged by cp.async + ldmatrix; the online softmax runs in the
// log2 domain with the 1/sqrt(dh) scale folded into one FFMA per score.
//
// Layout: qkv bf16 [B*L][3C]; out bf16 [B*L][C]; lse2 fp32 [B][heads][L] (log2-domain logsumexp).
#include "../../include/tinysd_b200.h"
#include "common.cuh"
#include "attention_tc.cuh"
#include <cstdlib>

using namespace tsd;

namespace {

constexpr int KV_TILE = 64;

__device__ __forceinline__ uint32_t smem_u32_(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// B fragments for two n-tiles (rows row0..row0+15 of a row-major [n][k] tile), k = cols col0..col0+15.
// r[0],r[1] = (b0,b1) of n-tile row0; r[2],r[3] = n-tile row0+8.
__device__ __forceinline__ void ldsm_nt(uint32_t* r, uint32_t base, int RS, int row0, int col0, int lane) {
  const int mat = lane >> 3, rr = lane & 7;
  const uint32_t addr = base + (row0 + 8 * (mat >> 1) + rr) * RS + (col0 + 8 * (mat & 1)) * 2;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// B fragments from a row-major [k][n] tile (transposed load): k = rows row0..row0+15, two n-tiles
// cols col0..col0+7 and col0+8..col0+15.  r[0],r[1] = n-tile col0; r[2],r[3] = n-tile col0+8.
__device__ __forceinline__ void ldsm_t(uint32_t* r, uint32_t base, int RS, int row0, int col0, int lane) {
  const int mat = lane >> 3, rr = lane & 7;
  const uint32_t addr = base + (row0 + 8 * (mat & 1) + rr) * RS + (col0 + 8 * (mat >> 1)) * 2;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// A fragment (16 rows x 16 cols at column col0) straight from global memory; rows >= row_limit read row 0.
__device__ __forceinline__ void load_a_frag(uint32_t* a, const bf16* base, size_t ld, int row0, int row_limit,
                                            int col0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  int r0 = row0 + g, r1 = row0 + g + 8;
  if (r0 >= row_limit) r0 = 0;
  if (r1 >= row_limit) r1 = 0;
  const bf16* p0 = base + (size_t)r0 * ld + col0 + 2 * t;
  const bf16* p1 = base + (size_t)r1 * ld + col0 + 2 * t;
  a[0] = *reinterpret_cast<const uint32_t*>(p0);
  a[1] = *reinterpret_cast<const uint32_t*>(p1);
  a[2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
}

// Cooperative async load of a [64 rows][DH] bf16 tile (rows row0.. of a strided matrix) into padded smem.
template <int DH>
__device__ __forceinline__ void load_tile_async(uint32_t smem, const bf16* base, size_t ld, int row0, int row_limit) {
  constexpr int RS = DH * 2 + 16;
  constexpr int CH = DH / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < KV_TILE * CH; i += blockDim.x) {
    const int r = i / CH, c = i - r * CH;
    const int row = row0 + r;
    const bool ok = row < row_limit;
    cp_async16(smem + r * RS + c * 16, base + (size_t)(ok ? row : 0) * ld + c * 8, ok ? 16 : 0);
  }
}

// Same, for an arbitrary number of rows (a stage of several 64-key tiles).
template <int DH>
__device__ __forceinline__ void load_rows_async(uint32_t smem, const bf16* base, size_t ld, int row0, int row_limit,
                                                int nrows) {
  constexpr int RS = DH * 2 + 16;
  constexpr int CH = DH / 8;
  for (int i = threadIdx.x; i < nrows * CH; i += blockDim.x) {
    const int r = i / CH, c = i - r * CH;
    const int row = row0 + r;
    const bool ok = row < row_limit;
    cp_async16(smem + r * RS + c * 16, base + (size_t)(ok ? row : 0) * ld + c * 8, ok ? 16 : 0);
  }
}

// ------------------------------------------------------------------------------------------
// Forward.  grid = (ceil(L / (warps*16*MT)), heads, B); each warp owns MT m-tiles of 16 query rows.
// K/V arrive in stages of NSUB*64 keys (cp.async, double buffered): one barrier pair per stage, so the
// warps of a CTA drift apart inside a stage and MUFU / FMA / tensor work of different warps overlap.
// Softmax: running max on the raw scores, p = ex2(fma(s, c, -m*c)) with c = log2(e)/sqrt(dh).
// ------------------------------------------------------------------------------------------
template <int DH, int MT, int NSUB>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                       float* __restrict__ lse2, int L, int C, float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int STAGE_KEYS = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_KEYS * RS;  // per operand
  extern __shared__ __align__(16) uint8_t smem_dyn[];  // [2][K | V]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const int q0 = (blockIdx.x * nwarps + warp) * 16 * MT;
  const uint32_t smem0 = smem_u32_(smem_dyn);

  uint32_t qf[MT][KT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) load_a_frag(qf[mt][kk], qbase, ld, q0 + mt * 16, L, kk * 16, lane);

  float oacc[MT][ND][4];
  float mrow[MT][2], lrow[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    mrow[mt][0] = mrow[mt][1] = -INFINITY;
    lrow[mt][0] = lrow[mt][1] = 0.f;
#pragma unroll
    for (int j = 0; j < ND; ++j) oacc[mt][j][0] = oacc[mt][j][1] = oacc[mt][j][2] = oacc[mt][j][3] = 0.f;
  }

  const int nstages = (L + STAGE_KEYS - 1) / STAGE_KEYS;
  load_rows_async<DH>(smem0, kbase, ld, 0, L, STAGE_KEYS);
  load_rows_async<DH>(smem0 + STAGE_BYTES, vbase, ld, 0, L, STAGE_KEYS);
  cp_async_commit();

  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, kbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      load_rows_async<DH>(nb + STAGE_BYTES, vbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
      const int key0 = st * STAGE_KEYS + sub * KV_TILE;
      if (key0 >= L) break;
      const uint32_t kS = smem0 + buf * 2 * STAGE_BYTES + sub * KV_TILE * RS;
      const uint32_t vS = kS + STAGE_BYTES;

      float sacc[MT][8][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc[mt][j][0] = sacc[mt][j][1] = sacc[mt][j][2] = sacc[mt][j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          uint32_t r[4];
          ldsm_nt(r, kS, RS, 16 * jp, 16 * kk, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(sacc[mt][2 * jp], qf[mt][kk], r[0], r[1]);
            mma_bf16(sacc[mt][2 * jp + 1], qf[mt][kk], r[2], r[3]);
          }
        }
      const bool partial = key0 + KV_TILE > L;
      float msc[MT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (partial) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (key0 + 8 * j + 2 * t + (e & 1) >= L) sacc[mt][j][e] = -INFINITY;
          }
          mx0 = fmaxf(mx0, fmaxf(sacc[mt][j][0], sacc[mt][j][1]));
          mx1 = fmaxf(mx1, fmaxf(sacc[mt][j][2], sacc[mt][j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(mrow[mt][0], mx0), mn1 = fmaxf(mrow[mt][1], mx1);
        const float c0 = ex2((mrow[mt][0] - mn0) * scale_log2), c1 = ex2((mrow[mt][1] - mn1) * scale_log2);
        mrow[mt][0] = mn0; mrow[mt][1] = mn1;
        msc[mt][0] = mn0 * scale_log2; msc[mt][1] = mn1 * scale_log2;
        lrow[mt][0] *= c0; lrow[mt][1] *= c1;
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          oacc[mt][j][0] *= c0; oacc[mt][j][1] *= c0; oacc[mt][j][2] *= c1; oacc[mt][j][3] *= c1;
        }
      }
      // exponentials of k-step kk (two n-tiles) are produced right before the PV MMAs that consume them
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pf[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * kk + jj;
            const float p0 = ex2(fmaf(sacc[mt][j][0], scale_log2, -msc[mt][0]));
            const float p1 = ex2(fmaf(sacc[mt][j][1], scale_log2, -msc[mt][0]));
            const float p2 = ex2(fmaf(sacc[mt][j][2], scale_log2, -msc[mt][1]));
            const float p3 = ex2(fmaf(sacc[mt][j][3], scale_log2, -msc[mt][1]));
            lrow[mt][0] += p0 + p1;
            lrow[mt][1] += p2 + p3;
            pf[mt][jj * 2] = pack_bf16(p0, p1);
            pf[mt][jj * 2 + 1] = pack_bf16(p2, p3);
          }
        }
#pragma unroll
        for (int jp = 0; jp < ND / 2; ++jp) {
          uint32_t r[4];
          ldsm_t(r, vS, RS, 16 * kk, 16 * jp, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(oacc[mt][2 * jp], pf[mt], r[0], r[1]);
            mma_bf16(oacc[mt][2 * jp + 1], pf[mt], r[2], r[3]);
          }
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float l0 = lrow[mt][0], l1 = lrow[mt][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      if (r0 < L)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + r0) * C + h * DH + 8 * j + 2 * t) =
            pack_bf16(oacc[mt][j][0] * i0, oacc[mt][j][1] * i0);
      if (r1 < L)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + r1) * C + h * DH + 8 * j + 2 * t) =
            pack_bf16(oacc[mt][j][2] * i1, oacc[mt][j][3] * i1);
    }
    if (lse2 && t == 0) {
      if (r0 < L) lse2[((size_t)b * H + h) * L + r0] = mrow[mt][0] * scale_log2 + log2f(l0);
      if (r1 < L) lse2[((size_t)b * H + h) * L + r1] = mrow[mt][1] * scale_log2 + log2f(l1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Forward, tuned variant (one m-tile per warp, more CTAs per SM).  Differences from attn_fwd_kernel:
//   * row sums come out of the PV MMA: every V row carries a bf16 1.0 in its 16-byte pad, read as one
//     extra n-tile, so sum_k P[q][k] accumulates in the tensor pipe on exactly the bf16 P the numerator uses;
//   * 3-input max (FMNMX3) and packed fp32 FMA (FFMA2) halve the softmax instruction count;
//   * POLY of the 8 n-tiles per 64-key tile evaluate 2^x with a cubic on the FMA pipe instead of MUFU
//     (Cody-Waite split, magic-number rounding), because MUFU.EX2 (16/clk/SM) is the binding unit.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for a packed pair, x <= 0 (clamped at -126): cubic on [-0.5, 0.5], |rel err| < 2e-4 (bf16 P needs 4e-3)
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& o0, float& o1) {
  float x0, x1;
  unpack2(x, x0, x1);
  const uint64_t xc = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t r = fadd2(xc, pack2(12582912.f, 12582912.f));
  const uint64_t fl = fadd2(r, pack2(-12582912.f, -12582912.f));
  const uint64_t f = ffma2(fl, pack2(-1.f, -1.f), xc);
  uint64_t pv = ffma2(f, pack2(0.05550411f, 0.05550411f), pack2(0.24022651f, 0.24022651f));
  pv = ffma2(pv, f, pack2(0.69314718f, 0.69314718f));
  pv = ffma2(pv, f, pack2(1.f, 1.f));
  float r0, r1, p0, p1;
  unpack2(r, r0, r1);
  unpack2(pv, p0, p1);
  o0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  o1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}
template <int DH, int MT, int NSUB, int POLY, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) attn_fwd2_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                               float* __restrict__ lse2, int L, int C, float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int STAGE_KEYS = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_KEYS * RS;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const int q0 = (blockIdx.x * nwarps + warp) * 16 * MT;
  const uint32_t smem0 = smem_u32_(smem_dyn);
  // B fragment of the constant "ones" n-tile (column 0 = 1): b0 = b1 = {1, 1} on lanes with g == 0
  const uint32_t ones = g == 0 ? 0x3f803f80u : 0u;

  uint32_t qf[MT][KT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) load_a_frag(qf[mt][kk], qbase, ld, q0 + mt * 16, L, kk * 16, lane);

  float oacc[MT][ND + 1][4];
  float m0[MT], m1[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    m0[mt] = m1[mt] = -INFINITY;
#pragma unroll
    for (int j = 0; j <= ND; ++j) oacc[mt][j][0] = oacc[mt][j][1] = oacc[mt][j][2] = oacc[mt][j][3] = 0.f;
  }
  const uint64_t c2 = pack2(scale_log2, scale_log2);

  const int nstages = (L + STAGE_KEYS - 1) / STAGE_KEYS;
  load_rows_async<DH>(smem0, kbase, ld, 0, L, STAGE_KEYS);
  load_rows_async<DH>(smem0 + STAGE_BYTES, vbase, ld, 0, L, STAGE_KEYS);
  cp_async_commit();

  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, kbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      load_rows_async<DH>(nb + STAGE_BYTES, vbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
      const int key0 = st * STAGE_KEYS + sub * KV_TILE;
      if (key0 >= L) break;
      const uint32_t kS = smem0 + buf * 2 * STAGE_BYTES + sub * KV_TILE * RS;
      const uint32_t vS = kS + STAGE_BYTES;
      float sacc[MT][8][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc[mt][j][0] = sacc[mt][j][1] = sacc[mt][j][2] = sacc[mt][j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          uint32_t r[4];
          ldsm_nt(r, kS, RS, 16 * jp, 16 * kk, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(sacc[mt][2 * jp], qf[mt][kk], r[0], r[1]);
            mma_bf16(sacc[mt][2 * jp + 1], qf[mt][kk], r[2], r[3]);
          }
        }
      uint64_t nm0[MT], nm1[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (key0 + KV_TILE > L) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (key0 + 8 * j + 2 * t + (e & 1) >= L) sacc[mt][j][e] = -INFINITY;
        }
        float mx0 = m0[mt], mx1 = m1[mt];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mx0 = max3(mx0, sacc[mt][j][0], sacc[mt][j][1]);
          mx1 = max3(mx1, sacc[mt][j][2], sacc[mt][j][3]);
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float cr0 = ex2((m0[mt] - mx0) * scale_log2), cr1 = ex2((m1[mt] - mx1) * scale_log2);
        m0[mt] = mx0; m1[mt] = mx1;
        nm0[mt] = pack2(-mx0 * scale_log2, -mx0 * scale_log2);
        nm1[mt] = pack2(-mx1 * scale_log2, -mx1 * scale_log2);
#pragma unroll
        for (int j = 0; j <= ND; ++j) {
          oacc[mt][j][0] *= cr0; oacc[mt][j][1] *= cr0; oacc[mt][j][2] *= cr1; oacc[mt][j][3] *= cr1;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pf[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * kk + jj;
            const uint64_t x01 = ffma2(pack2(sacc[mt][j][0], sacc[mt][j][1]), c2, nm0[mt]);
            const uint64_t x23 = ffma2(pack2(sacc[mt][j][2], sacc[mt][j][3]), c2, nm1[mt]);
            float p0, p1, p2, p3;
            const bool use_poly = POLY > 0 && ((j + 1) % (8 / (POLY > 0 ? POLY : 1)) == 0);
            if (use_poly) {
              exp2_poly2(x01, p0, p1);
              exp2_poly2(x23, p2, p3);
            } else {
              float a0, a1, a2, a3;
              unpack2(x01, a0, a1);
              unpack2(x23, a2, a3);
              p0 = ex2(a0); p1 = ex2(a1); p2 = ex2(a2); p3 = ex2(a3);
            }
            pf[mt][jj * 2] = pack_bf16(p0, p1);
            pf[mt][jj * 2 + 1] = pack_bf16(p2, p3);
          }
#pragma unroll
        for (int jp = 0; jp < ND / 2; ++jp) {
          uint32_t r[4];
          ldsm_t(r, vS, RS, 16 * kk, 16 * jp, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(oacc[mt][2 * jp], pf[mt], r[0], r[1]);
            mma_bf16(oacc[mt][2 * jp + 1], pf[mt], r[2], r[3]);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_bf16(oacc[mt][ND], pf[mt], ones, ones);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    // row sums sit in column 0 of the extra n-tile (threads with t == 0): broadcast over the quad
    const float l0 = __shfl_sync(0xffffffffu, oacc[mt][ND][0], lane & ~3);
    const float l1 = __shfl_sync(0xffffffffu, oacc[mt][ND][2], lane & ~3);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      if (r0 < L)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + r0) * C + h * DH + 8 * j + 2 * t) =
            pack_bf16(oacc[mt][j][0] * i0, oacc[mt][j][1] * i0);
      if (r1 < L)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + r1) * C + h * DH + 8 * j + 2 * t) =
            pack_bf16(oacc[mt][j][2] * i1, oacc[mt][j][3] * i1);
    }
    if (lse2 && t == 0) {
      if (r0 < L) lse2[((size_t)b * H + h) * L + r0] = m0[mt] * scale_log2 + log2f(l0);
      if (r1 < L) lse2[((size_t)b * H + h) * L + r1] = m1[mt] * scale_log2 + log2f(l1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward prep: delta[b][h][row] = sum_d dO * O
// ------------------------------------------------------------------------------------------
template <int DH>
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                     int B, int L, int C) {
  const int H = C / DH;
  const size_t total = (size_t)B * L * H;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const size_t row = i / H;  // b*L + l
    const bf16* po = o + row * C + h * DH;
    const bf16* pd = dout + row * C + h * DH;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(po + c), d = *reinterpret_cast<const uint4*>(pd + c);
      const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
      const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
      acc += a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y;
    }
    const size_t b = row / L, l = row % L;
    delta[(b * H + h) * L + l] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// Backward dQ: query-outer, recompute P, dS = P * (dP - delta), dQ += dS K.  One m-tile per warp.
// ------------------------------------------------------------------------------------------
template <int DH, int NSUB, int MINB>
__global__ void __launch_bounds__(256, MINB) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                          const float* __restrict__ lse2, const float* __restrict__ delta,
                                                          bf16* __restrict__ dqkv, int L, int C, float scale,
                                                          float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int STAGE_KEYS = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_KEYS * RS;
  extern __shared__ __align__(16) uint8_t smem_dyn[];  // [2][K | V]
  const uint32_t smem0 = smem_u32_(smem_dyn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const bf16* dobase = dout + (size_t)b * L * C + h * DH;
  const int q0 = (blockIdx.x * nwarps + warp) * 16;

  uint32_t qf[KT][4], dof[KT][4];
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    load_a_frag(qf[kk], qbase, ld, q0, L, kk * 16, lane);
    load_a_frag(dof[kk], dobase, C, q0, L, kk * 16, lane);
  }
  const int r0 = q0 + g, r1 = q0 + g + 8;
  const size_t sbase = ((size_t)b * H + h) * L;
  const float lse0 = r0 < L ? lse2[sbase + r0] : 0.f, lse1 = r1 < L ? lse2[sbase + r1] : 0.f;
  const float dl0 = r0 < L ? delta[sbase + r0] : 0.f, dl1 = r1 < L ? delta[sbase + r1] : 0.f;

  float dq[ND][4];
#pragma unroll
  for (int j = 0; j < ND; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;

  const int nstages = (L + STAGE_KEYS - 1) / STAGE_KEYS;
  load_rows_async<DH>(smem0, kbase, ld, 0, L, STAGE_KEYS);
  load_rows_async<DH>(smem0 + STAGE_BYTES, vbase, ld, 0, L, STAGE_KEYS);
  cp_async_commit();
  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, kbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      load_rows_async<DH>(nb + STAGE_BYTES, vbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
    const int key0 = st * STAGE_KEYS + sub * KV_TILE;
    if (key0 >= L) break;
    const uint32_t kS = smem0 + buf * 2 * STAGE_BYTES + sub * KV_TILE * RS, vS = kS + STAGE_BYTES;
    float sacc[8][4], pacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
      pacc[j][0] = pacc[j][1] = pacc[j][2] = pacc[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t r[4];
        ldsm_nt(r, kS, RS, 16 * jp, 16 * kk, lane);
        mma_bf16(sacc[2 * jp], qf[kk], r[0], r[1]);
        mma_bf16(sacc[2 * jp + 1], qf[kk], r[2], r[3]);
        ldsm_nt(r, vS, RS, 16 * jp, 16 * kk, lane);
        mma_bf16(pacc[2 * jp], dof[kk], r[0], r[1]);
        mma_bf16(pacc[2 * jp + 1], dof[kk], r[2], r[3]);
      }
    const bool partial = key0 + KV_TILE > L;
    uint32_t dsf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lse = (e < 2) ? lse0 : lse1;
        const float dl = (e < 2) ? dl0 : dl1;
        float p = ex2(sacc[j][e] * scale_log2 - lse);
        if (partial && (key0 + 8 * j + 2 * t + (e & 1)) >= L) p = 0.f;
        ds[e] = p * (pacc[j][e] - dl);
      }
      dsf[j >> 1][(j & 1) * 2] = pack_bf16(ds[0], ds[1]);
      dsf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int jp = 0; jp < ND / 2; ++jp) {
        uint32_t r[4];
        ldsm_t(r, kS, RS, 16 * kk, 16 * jp, lane);
        mma_bf16(dq[2 * jp], dsf[kk], r[0], r[1]);
        mma_bf16(dq[2 * jp + 1], dsf[kk], r[2], r[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    if (r0 < L)
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + h * DH + 8 * j + 2 * t) =
          pack_bf16(dq[j][0] * scale, dq[j][1] * scale);
    if (r1 < L)
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + h * DH + 8 * j + 2 * t) =
          pack_bf16(dq[j][2] * scale, dq[j][3] * scale);
  }
}

// ------------------------------------------------------------------------------------------
// Backward dK/dV: key-outer (each warp owns 16 keys), loops over query tiles of 64.
//   S^T = K Q^T,  P^T = exp2(S^T*c - lse[q]),  dV += P^T dO,  dP^T = V dO^T,
//   dS^T = P^T * (dP^T - delta[q]),  dK += dS^T Q   (scaled by 1/sqrt(dh) at the end)
// ------------------------------------------------------------------------------------------
template <int DH, int NSUB, int MINB>
__global__ void __launch_bounds__(256, MINB) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                           const float* __restrict__ lse2, const float* __restrict__ delta,
                                                           bf16* __restrict__ dqkv, int L, int C, float scale,
                                                           float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int STAGE_Q = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_Q * RS;
  extern __shared__ __align__(16) uint8_t smem_dyn[];  // [2][Q | dO] then lse[2][STAGE_Q], delta[2][STAGE_Q]
  const uint32_t smem0 = smem_u32_(smem_dyn);
  float* sLse = reinterpret_cast<float*>(smem_dyn + 4 * STAGE_BYTES);
  float* sDl = sLse + 2 * STAGE_Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const bf16* dobase = dout + (size_t)b * L * C + h * DH;
  const size_t sbase = ((size_t)b * H + h) * L;
  const int k0 = (blockIdx.x * nwarps + warp) * 16;

  uint32_t kf[KT][4], vf[KT][4];
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    load_a_frag(kf[kk], kbase, ld, k0, L, kk * 16, lane);
    load_a_frag(vf[kk], vbase, ld, k0, L, kk * 16, lane);
  }
  const bool key_ok0 = (k0 + g) < L, key_ok1 = (k0 + g + 8) < L;
  float dk[ND][4], dv[ND][4];
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const int nstages = (L + STAGE_Q - 1) / STAGE_Q;
  auto load_stats = [&](int buf, int stage) {
    for (int i = threadIdx.x; i < STAGE_Q; i += blockDim.x) {
      const int q = stage * STAGE_Q + i;
      // padded query rows: lse = +inf makes P = 0
      sLse[buf * STAGE_Q + i] = q < L ? lse2[sbase + q] : INFINITY;
      sDl[buf * STAGE_Q + i] = q < L ? delta[sbase + q] : 0.f;
    }
  };
  load_rows_async<DH>(smem0, qbase, ld, 0, L, STAGE_Q);
  load_rows_async<DH>(smem0 + STAGE_BYTES, dobase, C, 0, L, STAGE_Q);
  cp_async_commit();
  load_stats(0, 0);
  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, qbase, ld, (st + 1) * STAGE_Q, L, STAGE_Q);
      load_rows_async<DH>(nb + STAGE_BYTES, dobase, C, (st + 1) * STAGE_Q, L, STAGE_Q);
      cp_async_commit();
      load_stats(buf ^ 1, st + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
    if (st * STAGE_Q + sub * KV_TILE >= L) break;
    const uint32_t qS = smem0 + buf * 2 * STAGE_BYTES + sub * KV_TILE * RS, dS_ = qS + STAGE_BYTES;
    const float* lseS = sLse + buf * STAGE_Q + sub * KV_TILE;
    const float* dlS = sDl + buf * STAGE_Q + sub * KV_TILE;
    float sacc[8][4], pacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
      pacc[j][0] = pacc[j][1] = pacc[j][2] = pacc[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t r[4];
        ldsm_nt(r, qS, RS, 16 * jp, 16 * kk, lane);
        mma_bf16(sacc[2 * jp], kf[kk], r[0], r[1]);
        mma_bf16(sacc[2 * jp + 1], kf[kk], r[2], r[3]);
        ldsm_nt(r, dS_, RS, 16 * jp, 16 * kk, lane);
        mma_bf16(pacc[2 * jp], vf[kk], r[0], r[1]);
        mma_bf16(pacc[2 * jp + 1], vf[kk], r[2], r[3]);
      }
    uint32_t pf[4][4], dsf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = 8 * j + 2 * t + (e & 1);
        float pe = ex2(sacc[j][e] * scale_log2 - lseS[qi]);
        if (!((e < 2) ? key_ok0 : key_ok1)) pe = 0.f;
        p[e] = pe;
        ds[e] = pe * (pacc[j][e] - dlS[qi]);
      }
      pf[j >> 1][(j & 1) * 2] = pack_bf16(p[0], p[1]);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
      dsf[j >> 1][(j & 1) * 2] = pack_bf16(ds[0], ds[1]);
      dsf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int jp = 0; jp < ND / 2; ++jp) {
        uint32_t r[4];
        ldsm_t(r, dS_, RS, 16 * kk, 16 * jp, lane);
        mma_bf16(dv[2 * jp], pf[kk], r[0], r[1]);
        mma_bf16(dv[2 * jp + 1], pf[kk], r[2], r[3]);
        ldsm_t(r, qS, RS, 16 * kk, 16 * jp, lane);
        mma_bf16(dk[2 * jp], dsf[kk], r[0], r[1]);
        mma_bf16(dk[2 * jp + 1], dsf[kk], r[2], r[3]);
      }
    }
    __syncthreads();
  }
  const int r0 = k0 + g, r1 = k0 + g + 8;
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    const int col = h * DH + 8 * j + 2 * t;
    if (r0 < L) {
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + C + col) = pack_bf16(dk[j][0] * scale, dk[j][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + 2 * C + col) = pack_bf16(dv[j][0], dv[j][1]);
    }
    if (r1 < L) {
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + C + col) = pack_bf16(dk[j][2] * scale, dk[j][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + 2 * C + col) = pack_bf16(dv[j][2], dv[j][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward dK/dV, register-blocked variant: each warp owns MT*16 keys and walks the query stage in sub-tiles of
// QT = 64/MT queries, so every ldmatrix of Q / dO feeds MT m-tiles (half the shared-memory traffic per MMA,
// which is what bounds the one-m-tile kernel above: ncu shows its smem pipe at 63 %).
// ------------------------------------------------------------------------------------------
template <int DH, int MT, int NSUB, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) attn_bwd_dkv2_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                   const float* __restrict__ lse2, const float* __restrict__ delta,
                                                                   bf16* __restrict__ dqkv, int L, int C, float scale,
                                                                   float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int QT = KV_TILE / MT;   // queries per sub-tile
  constexpr int NJ = QT / 8;         // n-tiles per sub-tile
  constexpr int STAGE_Q = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_Q * RS;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const uint32_t smem0 = smem_u32_(smem_dyn);
  float* sLse = reinterpret_cast<float*>(smem_dyn + 4 * STAGE_BYTES);
  float* sDl = sLse + 2 * STAGE_Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const bf16* dobase = dout + (size_t)b * L * C + h * DH;
  const size_t sbase = ((size_t)b * H + h) * L;
  const int k0 = (blockIdx.x * nwarps + warp) * 16 * MT;

  uint32_t kf[MT][KT][4], vf[MT][KT][4];
  bool key_ok[MT][2];
  float dk[MT][ND][4], dv[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      load_a_frag(kf[mt][kk], kbase, ld, k0 + 16 * mt, L, kk * 16, lane);
      load_a_frag(vf[mt][kk], vbase, ld, k0 + 16 * mt, L, kk * 16, lane);
    }
    key_ok[mt][0] = (k0 + 16 * mt + g) < L;
    key_ok[mt][1] = (k0 + 16 * mt + g + 8) < L;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      dk[mt][j][0] = dk[mt][j][1] = dk[mt][j][2] = dk[mt][j][3] = 0.f;
      dv[mt][j][0] = dv[mt][j][1] = dv[mt][j][2] = dv[mt][j][3] = 0.f;
    }
  }
  const bool partial_keys = k0 + 16 * MT > L;  // warp-uniform: only the last key block needs masking
  const int nstages = (L + STAGE_Q - 1) / STAGE_Q;
  auto load_stats = [&](int buf, int stage) {
    for (int i = threadIdx.x; i < STAGE_Q; i += blockDim.x) {
      const int q = stage * STAGE_Q + i;
      sLse[buf * STAGE_Q + i] = q < L ? lse2[sbase + q] : INFINITY;  // padded queries: P = 0
      sDl[buf * STAGE_Q + i] = q < L ? delta[sbase + q] : 0.f;
    }
  };
  load_rows_async<DH>(smem0, qbase, ld, 0, L, STAGE_Q);
  load_rows_async<DH>(smem0 + STAGE_BYTES, dobase, C, 0, L, STAGE_Q);
  cp_async_commit();
  load_stats(0, 0);
  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, qbase, ld, (st + 1) * STAGE_Q, L, STAGE_Q);
      load_rows_async<DH>(nb + STAGE_BYTES, dobase, C, (st + 1) * STAGE_Q, L, STAGE_Q);
      cp_async_commit();
      load_stats(buf ^ 1, st + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < STAGE_Q / QT; ++sub) {
      if (st * STAGE_Q + sub * QT >= L) break;
      const uint32_t qS = smem0 + buf * 2 * STAGE_BYTES + sub * QT * RS, dS_ = qS + STAGE_BYTES;
      const float* lseS = sLse + buf * STAGE_Q + sub * QT;
      const float* dlS = sDl + buf * STAGE_Q + sub * QT;
      float sacc[MT][NJ][4], pacc[MT][NJ][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          sacc[mt][j][0] = sacc[mt][j][1] = sacc[mt][j][2] = sacc[mt][j][3] = 0.f;
          pacc[mt][j][0] = pacc[mt][j][1] = pacc[mt][j][2] = pacc[mt][j][3] = 0.f;
        }
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int jp = 0; jp < NJ / 2; ++jp) {
          uint32_t r[4], r2[4];
          ldsm_nt(r, qS, RS, 16 * jp, 16 * kk, lane);
          ldsm_nt(r2, dS_, RS, 16 * jp, 16 * kk, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(sacc[mt][2 * jp], kf[mt][kk], r[0], r[1]);
            mma_bf16(sacc[mt][2 * jp + 1], kf[mt][kk], r[2], r[3]);
            mma_bf16(pacc[mt][2 * jp], vf[mt][kk], r2[0], r2[1]);
            mma_bf16(pacc[mt][2 * jp + 1], vf[mt][kk], r2[2], r2[3]);
          }
        }
      uint32_t pf[MT][NJ / 2][4], dsf[MT][NJ / 2][4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float2 ls = *reinterpret_cast<const float2*>(lseS + 8 * j + 2 * t);
        const float2 dl = *reinterpret_cast<const float2*>(dlS + 8 * j + 2 * t);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float pe = ex2(fmaf(sacc[mt][j][e], scale_log2, -((e & 1) ? ls.y : ls.x)));
            if (partial_keys && !key_ok[mt][e >> 1]) pe = 0.f;
            p[e] = pe;
            ds[e] = pe * (pacc[mt][j][e] - ((e & 1) ? dl.y : dl.x));
          }
          pf[mt][j >> 1][(j & 1) * 2] = pack_bf16(p[0], p[1]);
          pf[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
          dsf[mt][j >> 1][(j & 1) * 2] = pack_bf16(ds[0], ds[1]);
          dsf[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < NJ / 2; ++kk)
#pragma unroll
        for (int jp = 0; jp < ND / 2; ++jp) {
          uint32_t r[4], r2[4];
          ldsm_t(r, dS_, RS, 16 * kk, 16 * jp, lane);
          ldsm_t(r2, qS, RS, 16 * kk, 16 * jp, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(dv[mt][2 * jp], pf[mt][kk], r[0], r[1]);
            mma_bf16(dv[mt][2 * jp + 1], pf[mt][kk], r[2], r[3]);
            mma_bf16(dk[mt][2 * jp], dsf[mt][kk], r2[0], r2[1]);
            mma_bf16(dk[mt][2 * jp + 1], dsf[mt][kk], r2[2], r2[3]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r0 = k0 + 16 * mt + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int col = h * DH + 8 * j + 2 * t;
      if (r0 < L) {
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + C + col) = pack_bf16(dk[mt][j][0] * scale, dk[mt][j][1] * scale);
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + 2 * C + col) = pack_bf16(dv[mt][j][0], dv[mt][j][1]);
      }
      if (r1 < L) {
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + C + col) = pack_bf16(dk[mt][j][2] * scale, dk[mt][j][3] * scale);
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + 2 * C + col) = pack_bf16(dv[mt][j][2], dv[mt][j][3]);
      }
    }
  }
}

// Same idea for dQ: each warp owns MT*16 query rows and walks the key stage in sub-tiles of 64/MT keys.
template <int DH, int MT, int NSUB, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) attn_bwd_dq2_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                  const float* __restrict__ lse2, const float* __restrict__ delta,
                                                                  bf16* __restrict__ dqkv, int L, int C, float scale,
                                                                  float scale_log2) {
  constexpr int RS = DH * 2 + 16;
  constexpr int KT = DH / 16;
  constexpr int ND = DH / 8;
  constexpr int KTILE = KV_TILE / MT;  // keys per sub-tile
  constexpr int NJ = KTILE / 8;
  constexpr int STAGE_KEYS = KV_TILE * NSUB;
  constexpr int STAGE_BYTES = STAGE_KEYS * RS;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const uint32_t smem0 = smem_u32_(smem_dyn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const bf16* dobase = dout + (size_t)b * L * C + h * DH;
  const int q0 = (blockIdx.x * nwarps + warp) * 16 * MT;
  const size_t sbase = ((size_t)b * H + h) * L;

  uint32_t qf[MT][KT][4], dof[MT][KT][4];
  float nlse[MT][2], dl[MT][2], dq[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      load_a_frag(qf[mt][kk], qbase, ld, q0 + 16 * mt, L, kk * 16, lane);
      load_a_frag(dof[mt][kk], dobase, C, q0 + 16 * mt, L, kk * 16, lane);
    }
    const int r0 = q0 + 16 * mt + g, r1 = r0 + 8;
    nlse[mt][0] = r0 < L ? -lse2[sbase + r0] : 0.f;
    nlse[mt][1] = r1 < L ? -lse2[sbase + r1] : 0.f;
    dl[mt][0] = r0 < L ? delta[sbase + r0] : 0.f;
    dl[mt][1] = r1 < L ? delta[sbase + r1] : 0.f;
#pragma unroll
    for (int j = 0; j < ND; ++j) dq[mt][j][0] = dq[mt][j][1] = dq[mt][j][2] = dq[mt][j][3] = 0.f;
  }
  const int nstages = (L + STAGE_KEYS - 1) / STAGE_KEYS;
  load_rows_async<DH>(smem0, kbase, ld, 0, L, STAGE_KEYS);
  load_rows_async<DH>(smem0 + STAGE_BYTES, vbase, ld, 0, L, STAGE_KEYS);
  cp_async_commit();
  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, kbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      load_rows_async<DH>(nb + STAGE_BYTES, vbase, ld, (st + 1) * STAGE_KEYS, L, STAGE_KEYS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < STAGE_KEYS / KTILE; ++sub) {
      const int key0 = st * STAGE_KEYS + sub * KTILE;
      if (key0 >= L) break;
      const uint32_t kS = smem0 + buf * 2 * STAGE_BYTES + sub * KTILE * RS, vS = kS + STAGE_BYTES;
      float sacc[MT][NJ][4], pacc[MT][NJ][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          sacc[mt][j][0] = sacc[mt][j][1] = sacc[mt][j][2] = sacc[mt][j][3] = 0.f;
          pacc[mt][j][0] = pacc[mt][j][1] = pacc[mt][j][2] = pacc[mt][j][3] = 0.f;
        }
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int jp = 0; jp < NJ / 2; ++jp) {
          uint32_t r[4], r2[4];
          ldsm_nt(r, kS, RS, 16 * jp, 16 * kk, lane);
          ldsm_nt(r2, vS, RS, 16 * jp, 16 * kk, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(sacc[mt][2 * jp], qf[mt][kk], r[0], r[1]);
            mma_bf16(sacc[mt][2 * jp + 1], qf[mt][kk], r[2], r[3]);
            mma_bf16(pacc[mt][2 * jp], dof[mt][kk], r2[0], r2[1]);
            mma_bf16(pacc[mt][2 * jp + 1], dof[mt][kk], r2[2], r2[3]);
          }
        }
      const bool partial = key0 + KTILE > L;
      uint32_t dsf[MT][NJ / 2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float ds[4], pe4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pe4[e] = ex2(fmaf(sacc[mt][j][e], scale_log2, nlse[mt][e >> 1]));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float pe = pe4[e];
            if (partial && (key0 + 8 * j + 2 * t + (e & 1)) >= L) pe = 0.f;
            ds[e] = pe * (pacc[mt][j][e] - dl[mt][e >> 1]);
          }
          dsf[mt][j >> 1][(j & 1) * 2] = pack_bf16(ds[0], ds[1]);
          dsf[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
      for (int kk = 0; kk < NJ / 2; ++kk)
#pragma unroll
        for (int jp = 0; jp < ND / 2; ++jp) {
          uint32_t r[4];
          ldsm_t(r, kS, RS, 16 * kk, 16 * jp, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(dq[mt][2 * jp], dsf[mt][kk], r[0], r[1]);
            mma_bf16(dq[mt][2 * jp + 1], dsf[mt][kk], r[2], r[3]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r0 = q0 + 16 * mt + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      if (r0 < L)
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + h * DH + 8 * j + 2 * t) =
            pack_bf16(dq[mt][j][0] * scale, dq[mt][j][1] * scale);
      if (r1 < L)
        *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + h * DH + 8 * j + 2 * t) =
            pack_bf16(dq[mt][j][2] * scale, dq[mt][j][3] * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------
// One-pass backward (head_dim 16, L % 256 == 0): the key-outer dK/dV kernel above also produces dQ, so S, dP and the
// exponentials are evaluated once instead of twice.
//   * per 32-query sub-tile a warp holds dS^T (its 32 keys x 32 queries) as MMA fragments; one stmatrix.trans /
//     ldmatrix round trip through a private shared-memory scratch turns them into the B operand of
//     dQ^T[16 x 32 q] += K^T[16 x 32 keys] dS^T   (8 more HMMA; K^T fragments are loaded once per warp);
//   * the eight warps' partial dQ^T tiles (each over 32 of the CTA's 256 keys) go to per-warp slots; after the
//     sub-tile's CTA barrier every thread sums two of the 512 outputs over the eight slots and writes them, scaled, to
//     a [32 q][16] fp32 staging tile; one thread adds that tile to the fp32 dQ workspace [B][H][L][16] with a bulk
//     reduce (cp.reduce.async.bulk ... add.f32: 2 KB per 8192 scores of L2 reduction traffic).  Slots and staging are
//     double buffered, so there is exactly one barrier per sub-tile;
//   * a small kernel converts the workspace to the bf16 dq columns of dqkv.
// dQ partials of different key blocks are summed by the L2 in arrival order: like the split-K weight gradients, dQ is
// reproducible only up to fp32 summation order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stsm_t4(uint32_t addr, const uint32_t* r) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};"
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void fb_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fb_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fb_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // a protocol bug becomes a launch error, not a hung box
  }
}

constexpr int FB_MT = 2, FB_NSUB = 4, FB_NTHR = 256;
constexpr int FB_QT = KV_TILE / FB_MT;               // 32 queries per sub-tile
constexpr int FB_RS = 16 * 2 + 16;                   // padded Q / dO row
constexpr int FB_STAGE_Q = KV_TILE * FB_NSUB;        // 256 queries per stage
constexpr int FB_STAGE_BYTES = FB_STAGE_Q * FB_RS;   // 12 KB
constexpr int FB_SCR_PITCH = 80;                     // [32 q][32 keys] bf16 + 16 B pad: conflict-free ldmatrix
constexpr int FB_SCR_BYTES = FB_QT * FB_SCR_PITCH;   // 2560 B per warp
constexpr int FB_SLOT_FLOATS = 16 * FB_QT;           // 512 floats = dQ^T tile of one warp
constexpr int FB_OFF_STATS = 4 * FB_STAGE_BYTES;
constexpr int FB_OFF_SCR = FB_OFF_STATS + 4 * FB_STAGE_Q * 4;
constexpr int FB_OFF_SLOTS = FB_OFF_SCR + 8 * FB_SCR_BYTES;
constexpr int FB_OFF_STAGING = FB_OFF_SLOTS + 2 * 8 * FB_SLOT_FLOATS * 4;
constexpr int FB_OFF_BARS = FB_OFF_STAGING + 3 * FB_SLOT_FLOATS * 4;  // three staging tiles, then 4 mbarriers
constexpr int FB_SMEM = FB_OFF_BARS + 64;

__global__ void __launch_bounds__(FB_NTHR, 2)
attn_bwd_fused_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse2,
                      const float* __restrict__ delta, bf16* __restrict__ dqkv, float* __restrict__ dq_ws, int L, int C,
                      float scale, float scale_log2) {
  constexpr int DH = 16, MT = FB_MT, RS = FB_RS, ND = 2, QT = FB_QT, NJ = QT / 8, STAGE_Q = FB_STAGE_Q,
                STAGE_BYTES = FB_STAGE_BYTES;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const uint32_t smem0 = smem_u32_(smem_dyn);
  float* sLse = reinterpret_cast<float*>(smem_dyn + FB_OFF_STATS);
  float* sDl = sLse + 2 * STAGE_Q;
  float* sSlots = reinterpret_cast<float*>(smem_dyn + FB_OFF_SLOTS);
  float* sStaging = reinterpret_cast<float*>(smem_dyn + FB_OFF_STAGING);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const size_t ld = 3 * (size_t)C;
  const bf16* qbase = qkv + (size_t)b * L * ld + h * DH;
  const bf16* kbase = qbase + C;
  const bf16* vbase = qbase + 2 * C;
  const bf16* dobase = dout + (size_t)b * L * C + h * DH;
  const size_t sbase = ((size_t)b * H + h) * L;
  float* dq_head = dq_ws + sbase * DH;  // [L][16] fp32 of this (sample, head)
  const int k0 = (blockIdx.x * 8 + warp) * 16 * MT;
  const uint32_t scr = smem0 + FB_OFF_SCR + warp * FB_SCR_BYTES;

  uint32_t kf[MT][4], vf[MT][4], kT[MT][4];
  float dk[MT][ND][4], dv[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    load_a_frag(kf[mt], kbase, ld, k0 + 16 * mt, L, 0, lane);
    load_a_frag(vf[mt], vbase, ld, k0 + 16 * mt, L, 0, lane);
    // A fragment of K^T (rows = head dim, cols = the 16 keys of this m-tile): a0 = K^T[g][2t,2t+1], a1 = rows g+8,
    // a2 / a3 = keys +8
    const bf16* kr = kbase + (size_t)(k0 + 16 * mt) * ld;
    auto pair = [&](int key, int d) {
      const uint32_t lo = *reinterpret_cast<const uint16_t*>(kr + (size_t)key * ld + d);
      const uint32_t hi = *reinterpret_cast<const uint16_t*>(kr + (size_t)(key + 1) * ld + d);
      return lo | (hi << 16);
    };
    kT[mt][0] = pair(2 * t, g);
    kT[mt][1] = pair(2 * t, g + 8);
    kT[mt][2] = pair(2 * t + 8, g);
    kT[mt][3] = pair(2 * t + 8, g + 8);
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      dk[mt][j][0] = dk[mt][j][1] = dk[mt][j][2] = dk[mt][j][3] = 0.f;
      dv[mt][j][0] = dv[mt][j][1] = dv[mt][j][2] = dv[mt][j][3] = 0.f;
    }
  }
  // where this thread's two reduced outputs (slot elements tid and tid + 256, layout [n-tile][lane][4]) land in the
  // [32 q][16] staging tile
  int stg_off[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int e = threadIdx.x + 256 * i;
    const int nt = e >> 7, ln = (e >> 2) & 31, c = e & 3;
    const int d = (ln >> 2) + 8 * (c >> 1), q = 8 * nt + 2 * (ln & 3) + (c & 1);
    stg_off[i] = q * 16 + d;
  }
  const int nstages = L / STAGE_Q;
  const int nsub_total = nstages * (STAGE_Q / QT);
  auto load_stats = [&](int buf, int stage) {
    for (int i = threadIdx.x; i < STAGE_Q; i += blockDim.x) {
      const int q = stage * STAGE_Q + i;
      sLse[buf * STAGE_Q + i] = lse2[sbase + q];
      sDl[buf * STAGE_Q + i] = delta[sbase + q];
    }
  };
  // Split-phase hand-off (mbarriers, 256 arrivals each), so the warps need not run in lockstep:
  //   slot_bar[k & 1]: every thread has written its warp's dQ^T slot of sub-tile k;
  //   stg_bar[k & 1] : every thread has reduced sub-tile k into staging tile k % 3 (its slots are free again).
  // Sub-tile k is reduced while sub-tile k + 1 is in flight and its staging tile is handed to the bulk reduce one
  // sub-tile later still; three staging tiles make the read-completion wait of that copy fall two sub-tiles back.
  const uint32_t bars = smem0 + FB_OFF_BARS;
  auto slot_bar = [&](int k) { return bars + 8u * (k & 1); };
  auto stg_bar = [&](int k) { return bars + 16u + 8u * (k & 1); };
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) fb_mbar_init(bars + 8u * i, FB_NTHR);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto reduce_subtile = [&](int k) {  // after slot_bar(k) completed
    fb_mbar_wait(slot_bar(k), (k >> 1) & 1);
    const float* sl = sSlots + (k & 1) * 8 * FB_SLOT_FLOATS;
    float* stg = sStaging + (k % 3) * FB_SLOT_FLOATS;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = threadIdx.x + 256 * i;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += sl[w * FB_SLOT_FLOATS + e];
      stg[stg_off[i]] = a * scale;
    }
    fence_async_smem();
    fb_mbar_arrive(stg_bar(k));
  };
  auto flush_subtile = [&](int k) {  // thread 0, after stg_bar(k) completed
    bulk_reduce_add_f32(dq_head + (size_t)k * QT * DH, smem0 + FB_OFF_STAGING + (k % 3) * FB_SLOT_FLOATS * 4,
                        FB_SLOT_FLOATS * 4);
    bulk_commit();
  };

  load_rows_async<DH>(smem0, qbase, ld, 0, L, STAGE_Q);
  load_rows_async<DH>(smem0 + STAGE_BYTES, dobase, C, 0, L, STAGE_Q);
  cp_async_commit();
  load_stats(0, 0);
  int it = 0;  // global sub-tile counter
  for (int st = 0; st < nstages; ++st) {
    const int buf = st & 1;
    if (st + 1 < nstages) {
      const uint32_t nb = smem0 + (buf ^ 1) * 2 * STAGE_BYTES;
      load_rows_async<DH>(nb, qbase, ld, (st + 1) * STAGE_Q, L, STAGE_Q);
      load_rows_async<DH>(nb + STAGE_BYTES, dobase, C, (st + 1) * STAGE_Q, L, STAGE_Q);
      cp_async_commit();
      load_stats(buf ^ 1, st + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < STAGE_Q / QT; ++sub, ++it) {
      const uint32_t qS = smem0 + buf * 2 * STAGE_BYTES + sub * QT * RS, dS_ = qS + STAGE_BYTES;
      const float* lseS = sLse + buf * STAGE_Q + sub * QT;
      const float* dlS = sDl + buf * STAGE_Q + sub * QT;
      float sacc[MT][NJ][4], pacc[MT][NJ][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          sacc[mt][j][0] = sacc[mt][j][1] = sacc[mt][j][2] = sacc[mt][j][3] = 0.f;
          pacc[mt][j][0] = pacc[mt][j][1] = pacc[mt][j][2] = pacc[mt][j][3] = 0.f;
        }
#pragma unroll
      for (int jp = 0; jp < NJ / 2; ++jp) {
        uint32_t r[4], r2[4];
        ldsm_nt(r, qS, RS, 16 * jp, 0, lane);
        ldsm_nt(r2, dS_, RS, 16 * jp, 0, lane);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(sacc[mt][2 * jp], kf[mt], r[0], r[1]);
          mma_bf16(sacc[mt][2 * jp + 1], kf[mt], r[2], r[3]);
          mma_bf16(pacc[mt][2 * jp], vf[mt], r2[0], r2[1]);
          mma_bf16(pacc[mt][2 * jp + 1], vf[mt], r2[2], r2[3]);
        }
      }
      uint32_t pf[MT][NJ / 2][4], dsf[MT][NJ / 2][4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float2 ls = *reinterpret_cast<const float2*>(lseS + 8 * j + 2 * t);
        const float2 dl = *reinterpret_cast<const float2*>(dlS + 8 * j + 2 * t);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = ex2(fmaf(sacc[mt][j][e], scale_log2, -((e & 1) ? ls.y : ls.x)));
            ds[e] = p[e] * (pacc[mt][j][e] - ((e & 1) ? dl.y : dl.x));
          }
          pf[mt][j >> 1][(j & 1) * 2] = pack_bf16(p[0], p[1]);
          pf[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
          dsf[mt][j >> 1][(j & 1) * 2] = pack_bf16(ds[0], ds[1]);
          dsf[mt][j >> 1][(j & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
      }
      // dS^T fragments -> scratch as [query][key] (transposing store), for the dQ product below
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kk = 0; kk < NJ / 2; ++kk) {
          const int mtx = lane >> 3, i = lane & 7;  // matrix 0/1: queries 16kk + i, keys +0 / +8; 2/3: queries + 8
          stsm_t4(scr + (16 * kk + 8 * (mtx >> 1) + i) * FB_SCR_PITCH + (16 * mt + 8 * (mtx & 1)) * 2, dsf[mt][kk]);
        }
#pragma unroll
      for (int kk = 0; kk < NJ / 2; ++kk)
#pragma unroll
        for (int jp = 0; jp < ND / 2; ++jp) {
          uint32_t r[4], r2[4];
          ldsm_t(r, dS_, RS, 16 * kk, 16 * jp, lane);
          ldsm_t(r2, qS, RS, 16 * kk, 16 * jp, lane);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(dv[mt][2 * jp], pf[mt][kk], r[0], r[1]);
            mma_bf16(dv[mt][2 * jp + 1], pf[mt][kk], r[2], r[3]);
            mma_bf16(dk[mt][2 * jp], dsf[mt][kk], r2[0], r2[1]);
            mma_bf16(dk[mt][2 * jp + 1], dsf[mt][kk], r2[2], r2[3]);
          }
        }
      // dQ^T[16 x 32 q] = K^T[16 x 32 keys] dS^T: B fragments come back from the scratch ([n = query][k = key])
      __syncwarp();
      float dqT[NJ][4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) dqT[j][0] = dqT[j][1] = dqT[j][2] = dqT[j][3] = 0.f;
#pragma unroll
      for (int qh = 0; qh < QT / 16; ++qh)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t r[4];
          ldsm_nt(r, scr, FB_SCR_PITCH, 16 * qh, 16 * mt, lane);
          mma_bf16(dqT[2 * qh], kT[mt], r[0], r[1]);
          mma_bf16(dqT[2 * qh + 1], kT[mt], r[2], r[3]);
        }
      __syncwarp();  // the scratch is rewritten in the next sub-tile
      if (it >= 2) {
        fb_mbar_wait(stg_bar(it), ((it - 2) >> 1) & 1);  // sub-tile it-2 fully reduced: its slots are free, its staging final
        if (threadIdx.x == 0) {
          flush_subtile(it - 2);
          bulk_wait_read<1>();  // everything but the copy just issued has been read out of its staging tile
        }
      }
      float4* slot = reinterpret_cast<float4*>(sSlots + ((it & 1) * 8 + warp) * FB_SLOT_FLOATS);
#pragma unroll
      for (int j = 0; j < NJ; ++j) slot[j * 32 + lane] = make_float4(dqT[j][0], dqT[j][1], dqT[j][2], dqT[j][3]);
      fb_mbar_arrive(slot_bar(it));
      if (it >= 1) reduce_subtile(it - 1);
    }
    __syncthreads();  // stage boundary: the Q / dO buffers are about to be refilled
  }
  reduce_subtile(nsub_total - 1);
  if (threadIdx.x == 0) {
    for (int k = nsub_total - 2; k < nsub_total; ++k) {
      if (k < 0) continue;
      fb_mbar_wait(stg_bar(k), (k >> 1) & 1);
      flush_subtile(k);
    }
    bulk_wait_all<0>();
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r0 = k0 + 16 * mt + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int col = h * DH + 8 * j + 2 * t;
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + C + col) = pack_bf16(dk[mt][j][0] * scale, dk[mt][j][1] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r0) * ld + 2 * C + col) = pack_bf16(dv[mt][j][0], dv[mt][j][1]);
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + C + col) = pack_bf16(dk[mt][j][2] * scale, dk[mt][j][3] * scale);
      *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * L + r1) * ld + 2 * C + col) = pack_bf16(dv[mt][j][2], dv[mt][j][3]);
    }
  }
}

// dq workspace fp32 [B][H][L][DH] -> bf16 dq columns of dqkv [B*L][3C]; one thread per (row, head)
template <int DH>
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ ws, bf16* __restrict__ dqkv, int B,
                                                              int L, int C) {
  const int H = C / DH;
  const size_t total = (size_t)B * L * H;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const size_t row = i / H;  // b * L + l: consecutive threads write consecutive head slices of one row
    const size_t bb = row / L, l = row - bb * L;
    const float4* src = reinterpret_cast<const float4*>(ws + (((bb * H + h) * L) + l) * DH);
    uint4* dst = reinterpret_cast<uint4*>(dqkv + row * 3 * C + h * DH);
#pragma unroll
    for (int q = 0; q < DH / 8; ++q) {
      const float4 a = src[2 * q], b4 = src[2 * q + 1];
      dst[q] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b4.x, b4.y), pack_bf16(b4.z, b4.w));
    }
  }
}

struct LaunchShape { int warps, mt, grid_x; };
LaunchShape pick_shape(int L, bool allow_mt2) {
  LaunchShape s;
  if (allow_mt2 && L % 256 == 0) { s.warps = 8; s.mt = 2; }
  else if (L >= 128) { s.warps = 8; s.mt = 1; }
  else if (L >= 64) { s.warps = 4; s.mt = 1; }
  else { s.warps = (L + 15) / 16; if (s.warps < 1) s.warps = 1; s.mt = 1; }
  s.grid_x = ceil_div(L, s.warps * 16 * s.mt);
  return s;
}

}  // namespace

template <int DH, int MT, int NSUB>
static int launch_fwd(cudaStream_t st, dim3 grid, dim3 block, const void* qkv, void* out, float* lse2, int L, int C,
                      float scale_log2) {
  auto kern = attn_fwd_kernel<DH, MT, NSUB>;
  constexpr int smem = 2 * 2 * KV_TILE * NSUB * (DH * 2 + 16);
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.cur() = true;
  }
  kern<<<grid, block, smem, st>>>((const bf16*)qkv, (bf16*)out, lse2, L, C, scale_log2);
  TSD_LAUNCH_CHECK();
  return 0;
}

template <int DH, int MT, int NSUB, int POLY, int NTHR, int MINB>
static int launch_fwd2(cudaStream_t st, int B, int heads, const void* qkv, void* out, float* lse2, int L, int C,
                       float scale_log2) {
  auto kern = attn_fwd2_kernel<DH, MT, NSUB, POLY, NTHR, MINB>;
  constexpr int smem = 2 * 2 * KV_TILE * NSUB * (DH * 2 + 16);
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.cur() = true;
  }
  dim3 grid(ceil_div(L, (NTHR / 32) * 16 * MT), heads, B);
  kern<<<grid, NTHR, smem, st>>>((const bf16*)qkv, (bf16*)out, lse2, L, C, scale_log2);
  TSD_LAUNCH_CHECK();
  return 0;
}

static int g_attn_variant = -1;  // experiment knob: TSD_ATTN_FWD = mt*10 + nsub

static int attn_fwd_impl(void* stream, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C, int heads);
extern "C" int tsd_attn_fwd(void* stream, const void* qkv, void* out, float* lse2, int B, int L, int C, int heads) {
  return attn_fwd_impl(stream, qkv, out, lse2, nullptr, B, L, C, heads);
}
extern "C" int tsd_attn_fwd_ws(void* stream, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C,
                               int heads) {
  return attn_fwd_impl(stream, qkv, out, lse2, ws, B, L, C, heads);
}
static int attn_fwd_impl(void* stream, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C, int heads) {
  TSD_CHECK(C % heads == 0, "attn_fwd: C %% heads != 0");
  const int dh = C / heads;
  TSD_CHECK(dh == 16 || dh == 32 || dh == 64, "attn_fwd: head_dim %d not in {16, 32, 64}", dh);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)dh);
  {  // tcgen05 / TMEM path (attention_tc.cu): TSD_ATTN_TC=0 falls back to the mma.sync kernels below
    static int use_tc = -1, tc_poly = 4, tc_bound = 1;
    if (use_tc < 0) {
      const char* e = getenv("TSD_ATTN_TC");
      use_tc = e ? atoi(e) : 1;
      const char* pe = getenv("TSD_ATTN_TC_POLY");
      if (pe) tc_poly = atoi(pe);
      const char* be = getenv("TSD_ATTN_TC_BOUND");
      if (be) tc_bound = atoi(be);
    }
    if (use_tc && attn_tc_supported(L, C, heads))
      return launch_attn_fwd_tc((cudaStream_t)stream, qkv, out, lse2, tc_bound ? ws : nullptr, B, L, C, heads, tc_poly);
  }
  if (g_attn_variant < 0) {
    const char* e = getenv("TSD_ATTN_FWD");
    g_attn_variant = e ? atoi(e) : 112;
  }
  LaunchShape s = pick_shape(L, dh == 16);
  int nsub = L >= 256 ? 4 : 1;
  if (g_attn_variant > 0 && dh == 16 && L % 256 == 0) {
    s.mt = g_attn_variant / 10; nsub = g_attn_variant % 10; s.warps = 8;
    s.grid_x = ceil_div(L, s.warps * 16 * s.mt);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (g_attn_variant >= 100 && dh == 16 && L % 256 == 0) {
#define TSD_F2(MT, NS, P, NT, MB) return launch_fwd2<16, MT, NS, P, NT, MB>(st, B, heads, qkv, out, lse2, L, C, scale_log2)
    switch (g_attn_variant) {
      case 100: TSD_F2(1, 4, 0, 256, 4);
      case 101: TSD_F2(1, 4, 1, 256, 4);
      case 110: TSD_F2(2, 4, 0, 256, 2);
      case 111: TSD_F2(2, 4, 1, 256, 2);
      case 112: TSD_F2(2, 4, 2, 256, 2);
      case 120: TSD_F2(2, 4, 0, 128, 4);
      case 121: TSD_F2(2, 4, 1, 128, 4);
      case 122: TSD_F2(2, 4, 2, 128, 4);
      case 130: TSD_F2(2, 2, 0, 128, 4);
      case 140: TSD_F2(1, 4, 0, 128, 8);
      default: break;
    }
#undef TSD_F2
  }
  dim3 grid(s.grid_x, heads, B), block(s.warps * 32);
  if (dh == 16) {
    if (s.mt == 2 && nsub == 4) return launch_fwd<16, 2, 4>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    if (s.mt == 2 && nsub == 2) return launch_fwd<16, 2, 2>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    if (s.mt == 2 && nsub == 1) return launch_fwd<16, 2, 1>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    if (s.mt == 1 && nsub == 4) return launch_fwd<16, 1, 4>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    if (s.mt == 1 && nsub == 2) return launch_fwd<16, 1, 2>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    return launch_fwd<16, 1, 1>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
  }
  if (dh == 64) {  // the wider [1,2,4,4] configuration (diffusion.py:203): 512 channels / 8 heads
    if (nsub == 4) return launch_fwd<64, 1, 2>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
    return launch_fwd<64, 1, 1>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
  }
  if (nsub == 4) return launch_fwd<32, 1, 4>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
  return launch_fwd<32, 1, 1>(st, grid, block, qkv, out, lse2, L, C, scale_log2);
}

template <int NSUB>
static int launch_bwd(cudaStream_t st, dim3 grid, dim3 block, int dh, const void* qkv, const void* dout, const float* lse2,
                      const float* delta, void* dqkv, int L, int C, float scale, float scale_log2) {
  const int smem_dq16 = 4 * KV_TILE * NSUB * 48, smem_dq32 = 4 * KV_TILE * NSUB * 80, smem_dq64 = 4 * KV_TILE * NSUB * 144;
  const int smem_kv16 = smem_dq16 + 4 * KV_TILE * NSUB * 4, smem_kv32 = smem_dq32 + 4 * KV_TILE * NSUB * 4;
  const int smem_kv64 = smem_dq64 + 4 * KV_TILE * NSUB * 4;
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<16, NSUB, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq16));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<32, NSUB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq32));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<16, NSUB, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv16));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<32, NSUB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv32));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<64, NSUB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq64));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<64, NSUB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv64));
    configured.cur() = true;
  }
  if (dh == 16) {
    attn_bwd_dq_kernel<16, NSUB, 3><<<grid, block, smem_dq16, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
    TSD_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<16, NSUB, 3><<<grid, block, smem_kv16, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
  } else if (dh == 64) {
    attn_bwd_dq_kernel<64, NSUB, 1><<<grid, block, smem_dq64, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
    TSD_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<64, NSUB, 1><<<grid, block, smem_kv64, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
  } else {
    attn_bwd_dq_kernel<32, NSUB, 2><<<grid, block, smem_dq32, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
    TSD_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<32, NSUB, 2><<<grid, block, smem_kv32, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2);
  }
  TSD_LAUNCH_CHECK();
  return 0;
}

// dqkv [B*L][3C] receives (dq, dk, dv); delta is scratch fp32 [B][heads][L].
static int attn_bwd_impl(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2, float* delta,
                         void* dqkv, float* ws, int B, int L, int C, int heads);
extern "C" int tsd_attn_bwd(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2,
                            float* delta, void* dqkv, int B, int L, int C, int heads) {
  return attn_bwd_impl(stream, qkv, out, dout, lse2, delta, dqkv, nullptr, B, L, C, heads);
}
extern "C" int tsd_attn_bwd_ws(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2,
                               float* delta, void* dqkv, float* ws, int B, int L, int C, int heads) {
  return attn_bwd_impl(stream, qkv, out, dout, lse2, delta, dqkv, ws, B, L, C, heads);
}
static int attn_bwd_impl(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2, float* delta,
                         void* dqkv, float* ws, int B, int L, int C, int heads) {
  TSD_CHECK(C % heads == 0, "attn_bwd: C %% heads != 0");
  const int dh = C / heads;
  TSD_CHECK(dh == 16 || dh == 32 || dh == 64, "attn_bwd: head_dim %d not in {16, 32, 64}", dh);
  const float scale = 1.f / sqrtf((float)dh);
  const float scale_log2 = 1.4426950408889634f * scale;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * L * heads;
  int pg = (int)((total + 255) / 256);
  if (pg > num_sms() * 16) pg = num_sms() * 16;
  if (dh == 16) attn_bwd_prep_kernel<16><<<pg, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, B, L, C);
  else if (dh == 32) attn_bwd_prep_kernel<32><<<pg, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, B, L, C);
  else attn_bwd_prep_kernel<64><<<pg, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, B, L, C);
  TSD_LAUNCH_CHECK();
  static int bwd_variant = -1, bwd_fused = 1;
  if (bwd_variant < 0) {
    const char* e = getenv("TSD_ATTN_BWD");
    bwd_variant = e ? atoi(e) : 44;
    const char* f = getenv("TSD_ATTN_BWD_FUSED");
    if (f) bwd_fused = atoi(f);
  }
  static int bwd_tc = -1;
  if (bwd_tc < 0) {
    const char* e = getenv("TSD_ATTN_BWD_TC");
    bwd_tc = e ? atoi(e) : 1;
  }
  if (bwd_tc && ws != nullptr && attn_bwd_tc_supported(L, C, heads)) {
    // one-pass backward on tcgen05: dK/dV in TMEM, dQ through the fp32 workspace [B][heads][L][16]
    TSD_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)B * L * C, st));
    if (launch_attn_bwd_tc(st, qkv, dout, lse2, delta, dqkv, ws, B, L, C, heads)) return 1;
    int cg = (int)((total + 255) / 256);
    if (cg > num_sms() * 16) cg = num_sms() * 16;
    attn_dq_convert_kernel<16><<<cg, 256, 0, st>>>(ws, (bf16*)dqkv, B, L, C);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  if (bwd_tc && ws != nullptr && attn_bwd_tc32_supported(L, C, heads)) {
    // the same for head_dim 32 (attention_bwd_tc32.cu): one 128-key block per CTA, workspace [B][heads][L][32]
    TSD_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)B * L * C, st));
    if (launch_attn_bwd_tc32(st, qkv, dout, lse2, delta, dqkv, ws, B, L, C, heads)) return 1;
    int cg = (int)((total + 255) / 256);
    if (cg > num_sms() * 16) cg = num_sms() * 16;
    attn_dq_convert_kernel<32><<<cg, 256, 0, st>>>(ws, (bf16*)dqkv, B, L, C);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  if (bwd_fused && ws != nullptr && dh == 16 && L % 256 == 0) {
    // one-pass backward: dK/dV in registers, dQ through the fp32 workspace [B][heads][L][16]
    static tsd::PerDeviceFlag cfgd;
    if (!cfgd.cur()) {
      TSD_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      cfgd.cur() = true;
    }
    TSD_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)B * L * C, st));
    attn_bwd_fused_kernel<<<dim3(L / 256, heads, B), FB_NTHR, FB_SMEM, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta,
                                                                           (bf16*)dqkv, ws, L, C, scale, scale_log2);
    TSD_LAUNCH_CHECK();
    int cg = (int)((total + 255) / 256);
    if (cg > num_sms() * 16) cg = num_sms() * 16;
    attn_dq_convert_kernel<16><<<cg, 256, 0, st>>>(ws, (bf16*)dqkv, B, L, C);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  if (bwd_variant > 0 && dh == 16 && L % 256 == 0) {
    const int vq = bwd_variant / 10, vk = bwd_variant % 10;
    const int smem_q = 4 * KV_TILE * 4 * 48, smem_k = smem_q + 4 * KV_TILE * 4 * 4;
#define TSD_BWD_LAUNCH(KERN, MT, NT, SM)                                                                          \
    do {                                                                                                          \
      static tsd::PerDeviceFlag cfgd;                                                                                   \
      if (!cfgd.cur()) { TSD_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); cfgd.cur() = true; } \
      dim3 gr(ceil_div(L, (NT / 32) * 16 * MT), heads, B);                                                        \
      KERN<<<gr, NT, SM, st>>>((const bf16*)qkv, (const bf16*)dout, lse2, delta, (bf16*)dqkv, L, C, scale, scale_log2); \
      TSD_LAUNCH_CHECK();                                                                                         \
    } while (0)
    if (vq == 1) TSD_BWD_LAUNCH((attn_bwd_dq2_kernel<16, 2, 4, 256, 1>), 2, 256, smem_q);
    else if (vq == 2) TSD_BWD_LAUNCH((attn_bwd_dq2_kernel<16, 2, 4, 128, 3>), 2, 128, smem_q);
    else if (vq == 3) TSD_BWD_LAUNCH((attn_bwd_dq2_kernel<16, 1, 4, 256, 3>), 1, 256, smem_q);
    else TSD_BWD_LAUNCH((attn_bwd_dq2_kernel<16, 2, 4, 256, 2>), 2, 256, smem_q);
    if (vk == 1) TSD_BWD_LAUNCH((attn_bwd_dkv2_kernel<16, 2, 4, 256, 1>), 2, 256, smem_k);
    else if (vk == 2) TSD_BWD_LAUNCH((attn_bwd_dkv2_kernel<16, 2, 4, 128, 3>), 2, 128, smem_k);
    else if (vk == 3) TSD_BWD_LAUNCH((attn_bwd_dkv2_kernel<16, 1, 4, 256, 3>), 1, 256, smem_k);
    else TSD_BWD_LAUNCH((attn_bwd_dkv2_kernel<16, 2, 4, 256, 2>), 2, 256, smem_k);
#undef TSD_BWD_LAUNCH
    return 0;
  }
  const LaunchShape s = pick_shape(L, false);
  dim3 grid(s.grid_x, heads, B), block(s.warps * 32);
  const int rc = (L >= 256 && dh != 64) ? launch_bwd<4>(st, grid, block, dh, qkv, dout, lse2, delta, dqkv, L, C, scale, scale_log2)
                                        : launch_bwd<1>(st, grid, block, dh, qkv, dout, lse2, delta, dqkv, L, C, scale, scale_log2);
  if (rc) return rc;
  TSD_LAUNCH_CHECK();
  return 0;
}
