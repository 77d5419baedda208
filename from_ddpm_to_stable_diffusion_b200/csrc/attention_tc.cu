// tcgen05 / TMEM self-attention forward for head_dim 16 and 32 with L % 256 == 0 (nine of the ten attention blocks of the
// 64x64 configuration; the text below describes head_dim 16, head_dim 32 only changes the row width and swizzle).
//
// Reference: SelfAttention.forward, diffusion.py:46-58.  Same contract as the mma.sync kernels in attention.cu
// (qkv bf16 [B*L][3C], out bf16 [B*L][C], lse2 fp32 [B][heads][L] in the log2 domain).
//
// Why a second implementation: with dh = 16 the work per score is one exponential plus a handful of FMA-pipe
// instructions; in the mma.sync kernels the warp that does the softmax also issues every HMMA / LDSM and carries
// the accumulator fragments, and ends up issue-bound (5.6 issue slots per 32 scores, 57 % issue utilisation)
// well below the MUFU rate.  Here the contractions leave the softmax warps entirely:
//   * one thread issues  S = Q K^T  (UMMA 128 x 64 x 16, operands straight from TMA-written 32-byte-swizzled tiles)
//     and  O += P V  (UMMA 128 x 16 x 16, A = P read from TMEM, B = V as an MN-major tile);
//   * a softmax thread owns one query row (its TMEM lane): it reads 64 scores with tcgen05.ld, takes the maximum,
//     exponentiates, writes bf16 P back over S with tcgen05.st -- no shuffles, no fragments, no ldmatrix;
//   * the running maximum is only refreshed when it grows by more than 2^8 (P stays <= 256, exact in the final
//     normalisation because the same reference is used for the row sum), so the O rescale in TMEM is rare;
//   * every POLY-th pair of exponentials runs as a cubic on the FMA pipe (Cody-Waite), because MUFU.EX2
//     (16 / clk / SM) is the binding unit.
// Two softmax warpgroups (2 x 128 query rows) share one K/V stream; each has two 64-column S buffers so the next
// score tile is ready before the current one is finished.
#include "attention_tc.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <type_traits>

namespace tsd {
namespace {

constexpr int QT = 128;  // query rows per softmax warpgroup (= TMEM lanes)
constexpr int NWG = 2;
constexpr int KB = 64;   // keys per block
constexpr int NST = 6;   // K/V ring stages
constexpr int TC_THREADS = NWG * 128 + 64;  // warps 0-7 softmax (warp % 4 = TMEM sub-partition), 8 / 9 = MMA issuers
constexpr int WG_COLS = 128;  // TMEM columns per warpgroup: S [0,64), P [64,96), O [96, 96 + DH)
constexpr int S_COL = 0, P_COL = 64, O_COL = 96;
constexpr float BOUND_LIMIT = 50.f;  // log2 units: |s*c| <= 50 for every key, so 2^(s*c - bound) >= 2^-100
constexpr int TMEM_COLS = NWG * WG_COLS;  // 256: two CTAs per SM
constexpr float RESCALE_THRESHOLD = 8.f;
// head_dim 16: 32-byte rows, SWIZZLE_32B (descriptor code 6); head_dim 32: 64-byte rows, SWIZZLE_64B (code 4)
template <int DH>
struct Geo {
  static constexpr int ROWB = DH * 2;
  static constexpr uint32_t SWZ = DH == 16 ? 6u : 4u;
  static constexpr int Q_BYTES = NWG * QT * ROWB;
  static constexpr int KV_BYTES = KB * ROWB;
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int ONES_BYTES = KB * ROWB;  // LSUM: constant [64 keys][16] tile, 1 in column 0 (see the kernel)
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + NST * STAGE_BYTES + ONES_BYTES + 256;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2_(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2_(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for a packed pair, x <= 8: cubic on [-0.5, 0.5] after a Cody-Waite split, |rel err| < 2e-4 (bf16 P needs 4e-3)
template <bool CLAMP>
__device__ __forceinline__ void exp2_poly2_(uint64_t x, float& o0, float& o1) {
  uint64_t xc = x;
  if (CLAMP) {  // the exponent insertion below needs x >= -126 (guaranteed without a clamp in bound mode)
    float x0, x1;
    upk2(x, x0, x1);
    xc = pk2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  }
  const uint64_t r = fadd2_(xc, pk2(12582912.f, 12582912.f));
  const uint64_t fl = fadd2_(r, pk2(-12582912.f, -12582912.f));
  const uint64_t f = ffma2_(fl, pk2(-1.f, -1.f), xc);
  uint64_t pv = ffma2_(f, pk2(0.05550411f, 0.05550411f), pk2(0.24022651f, 0.24022651f));
  pv = ffma2_(pv, f, pk2(0.69314718f, 0.69314718f));
  pv = ffma2_(pv, f, pk2(1.f, 1.f));
  float r0, r1, p0, p1;
  upk2(r, r0, r1);
  upk2(pv, p0, p1);
  o0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  o1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// max_k |k|^2 per (sample, head): one thread per (key row, head), 16 bf16 each.  knmax2 must be zeroed.
template <int DH>
__global__ void __launch_bounds__(512) attn_knorm_kernel(const bf16* __restrict__ qkv, float* __restrict__ knmax2,
                                                         int L, int C, int H) {
  __shared__ int s_max[32];
  if (threadIdx.x < 32) s_max[threadIdx.x] = 0;
  __syncthreads();
  const int rows_per_block = blockDim.x / H;
  const int r = threadIdx.x / H, hd = threadIdx.x - r * H;
  const int row = blockIdx.x * rows_per_block + r;
  const int b = blockIdx.y;
  if (r < rows_per_block && row < L) {
    const uint4* p = reinterpret_cast<const uint4*>(qkv + ((size_t)b * L + row) * 3 * C + C + hd * DH);
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < DH / 8; ++v) {
      const uint4 a = p[v];
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16(w[i]);
        s = fmaf(f.x, f.x, s);
        s = fmaf(f.y, f.y, s);
      }
    }
    atomicMax(&s_max[hd], __float_as_int(s));  // non-negative floats order like ints
  }
  __syncthreads();
  if (threadIdx.x < H) atomicMax(reinterpret_cast<int*>(knmax2) + b * H + threadIdx.x, s_max[threadIdx.x]);
}

// LSUM (head_dim 16 only): the softmax row sum comes out of the tensor core instead of one FADD per score.  The P V
// product is issued with N = 32: the second 16-column chunk of its MN-major B operand is a constant tile (leading-
// dimension byte offset of the descriptor pointing at it) with ones in column 0, so TMEM column O_COL + 16 accumulates
// sum_k bf16(P) next to O -- the same rounded P values that form the numerator.
template <int DH, int POLY, bool LSUM>
__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse2,
                   const float* __restrict__ knmax2, int L, int C, float scale_log2) {
  constexpr int ROWB = Geo<DH>::ROWB, Q_BYTES = Geo<DH>::Q_BYTES, KV_BYTES = Geo<DH>::KV_BYTES,
                STAGE_BYTES = Geo<DH>::STAGE_BYTES;
  constexpr uint32_t SWZ = Geo<DH>::SWZ;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + Q_BYTES;
  const uint32_t sOnes = sKV + NST * STAGE_BYTES;
  const uint32_t bar_base = sOnes + Geo<DH>::ONES_BYTES;
  // barriers (8 B each): q_full, kv_full[NST], kv_empty[NST], s_full[NWG], s_free[NWG], p_full[NWG], pv_done[NWG]
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + NST + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 2 * NST + g); };
  auto s_free = [&](int g) { return bar_base + 8u * (1 + 2 * NST + NWG + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (1 + 2 * NST + 2 * NWG + g); };
  auto pv_done = [&](int g) { return bar_base + 8u * (1 + 2 * NST + 3 * NWG + g); };
  const uint32_t tmem_slot = bar_base + 8u * (1 + 2 * NST + 4 * NWG);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // warp index made warp-uniform for the compiler: the issuer warps run converged and elect a lane only around the
  // async instructions, so the UMMAs issue back to back from uniform registers (no per-instruction elect loop)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const int q0 = blockIdx.x * (NWG * QT);
  const int nkb = L / KB;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), NWG);  // one commit per warpgroup's P V
    }
    for (int g = 0; g < NWG; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_free(g), QT);
      mbar_init(p_full(g), QT);
      mbar_init(pv_done(g), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (LSUM) {
    // 64 rows of 32 bytes: 1.0 in the first element of BOTH 16-byte halves (the 32-byte swizzle may swap the halves of a
    // row; either way output column 16 -- and its twin, column 24 -- receives the row sum), zero elsewhere
    for (int i = threadIdx.x; i < Geo<DH>::ONES_BYTES / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(smem_raw + (sOnes - smem_u32(smem_raw)))[i] = (i % 4 == 0) ? 0x00003f80u : 0u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= 8) {
    {
      // =========================================================== MMA issuer of warpgroup g (g == 0 also feeds TMA)
      // The softmax warps raise their events in program order (s_free(j), p_full(j), s_free(j+1), ...), so the issuer
      // simply blocks on them in that order: no polling, and one issuer per warpgroup keeps the two independent.
      const int g = warp - 8;
      constexpr uint32_t idescS = umma_idesc_bf16(QT, KB, 0, 0);
      constexpr uint32_t idescPV = umma_idesc_bf16(QT, LSUM ? 2 * DH : DH, 0, 1);
      constexpr int AHEAD = NST - 2;  // K/V blocks in flight beyond the one being consumed
      const int row_base = b * L;
      auto load_kv = [&](int jn) {
        const int s = jn % NST;
        mbar_wait(kv_empty(s), ((jn / NST) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full(s), STAGE_BYTES);
          tma_load_2d(sKV + s * STAGE_BYTES, &tmQKV, kv_full(s), C + h * DH, row_base + jn * KB);
          tma_load_2d(sKV + s * STAGE_BYTES + KV_BYTES, &tmQKV, kv_full(s), 2 * C + h * DH, row_base + jn * KB);
        }
        __syncwarp();
      };
      if (g == 0) {
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
          for (int i = 0; i < NWG * QT / 64; ++i)
            tma_load_2d(sQ + i * 64 * ROWB, &tmQKV, q_full, h * DH, row_base + q0 + i * 64);
        }
        __syncwarp();
        for (int jn = 0; jn < AHEAD && jn < nkb; ++jn) load_kv(jn);
      }
      const uint64_t descQ = umma_smem_desc_sw(sQ + g * QT * ROWB, 0, 8 * ROWB, SWZ);
      const uint32_t tW = tmem_base + g * WG_COLS;
      auto issue_S = [&](int j) {
        const int s = j % NST;
        mbar_wait(kv_full(s), (j / NST) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t descK = umma_smem_desc_sw(sKV + s * STAGE_BYTES, 0, 8 * ROWB, SWZ);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)  // K-major: a 16-element k-step is 32 bytes further along the swizzled row
            umma_bf16(tW + S_COL, descQ + (uint64_t)(k * 2), descK + (uint64_t)(k * 2), idescS, k > 0 ? 1u : 0u);
          umma_commit(s_full(g));
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      for (int j = 0; j < nkb; ++j) {
        if (j + 1 < nkb) {
          mbar_wait(s_free(g), j & 1);  // S_j sits in the softmax warps' registers
          issue_S(j + 1);
        }
        mbar_wait(p_full(g), j & 1);
        tc_fence_after();
        const int s = j % NST;
        const uint32_t sV = sKV + s * STAGE_BYTES + KV_BYTES;
        if (elect_one()) {
          // MN-major B: 16-column chunks LBO bytes apart -- the second chunk is the constant ones tile (LSUM)
          const uint64_t descV = umma_smem_desc_sw(sV, LSUM ? sOnes - sV : 0u, 8 * ROWB, SWZ);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_bf16_ts(tW + O_COL, tW + P_COL + k * 8, descV + (uint64_t)(k * (16 * ROWB / 16)), idescPV,
                         (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(pv_done(g));
          umma_commit(kv_empty(s));
        }
        __syncwarp();
        if (g == 0 && j + AHEAD < nkb) load_kv(j + AHEAD);  // its stage was released by block j - 2
      }
    }
  } else {
    // =========================================================== softmax: one query row per thread
    const int g = warp >> 2;
    const int sub = warp & 3;  // TMEM sub-partition of this warp
    const int row = sub * 32 + lane;
    const uint32_t tW = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + g * WG_COLS;
    const uint32_t tS = tW + S_COL, tP = tW + P_COL, tO = tW + O_COL;
    const uint64_t c2 = pk2(scale_log2, scale_log2);
    // Cauchy-Schwarz bound on this row's scores: |q||k|max.  When it is small enough that 2^(s*c - bound) cannot
    // underflow for any key, it serves as the softmax reference from the start: no maximum pass, no rescale.
    float m_ref = -INFINITY;
    bool bound_mode = false;
    if (knmax2 != nullptr) {
      mbar_wait(q_full, 0);
      const uint4* qrow = reinterpret_cast<const uint4*>(smem_raw + (sQ - smem_u32(smem_raw)) + (g * QT + row) * ROWB);
      float qn2 = 0.f;  // the swizzle permutes 16-byte chunks inside the row: irrelevant for a norm
#pragma unroll
      for (int v = 0; v < DH / 8; ++v) {
        const uint4 a = qrow[v];
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16(w[i]);
          qn2 = fmaf(f.x, f.x, qn2);
          qn2 = fmaf(f.y, f.y, qn2);
        }
      }
      const float bound = sqrtf(qn2 * __ldg(knmax2 + b * H + h)) * scale_log2 * 1.001f + 1e-3f;
      bound_mode = __all_sync(0xffffffffu, bound <= BOUND_LIMIT);
      if (bound_mode) m_ref = bound;
    }
    uint64_t l2 = pk2(0.f, 0.f);
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(s_full(g), j & 1);
      tc_fence_after();
      float alpha = 1.f;
      bool rescale = false;
      if (!bound_mode) {
        // ---- exact mode: block maximum first, lazy refresh of the reference
        float mx0, mx1;
        uint32_t sv[32];
        tmem_ld32(tS, sv);
        tmem_ld_wait();
        mx0 = __uint_as_float(sv[0]);
        mx1 = __uint_as_float(sv[1]);
#pragma unroll
        for (int i = 2; i < 32; i += 4) {
          mx0 = max3f(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          if (i + 2 < 32) mx1 = max3f(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        }
        tmem_ld32(tS + 32, sv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = max3f(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          mx1 = max3f(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        }
        const float bm = fmaxf(mx0, mx1) * scale_log2;
        const bool need = bm > m_ref + RESCALE_THRESHOLD;  // first block: m_ref = -inf
        rescale = __any_sync(0xffffffffu, need);
        if (need) {
          alpha = ex2f(m_ref - bm);  // 0 on the first block
          m_ref = bm;
          if (!LSUM) {  // (LSUM: the sum lives in TMEM next to O and is rescaled with it)
            float la, lb;
            upk2(l2, la, lb);
            l2 = pk2(la * alpha, lb * alpha);
          }
        }
      }
      // ---- P = 2^(s*c - m_ref) as bf16 pairs; the S buffer is handed back as soon as it sits in registers
      const uint64_t nm2 = pk2(-m_ref, -m_ref);
      uint32_t pk[KB / 2];
      auto exp_pass = [&](auto clamp_tag) {
        constexpr bool CLAMP = decltype(clamp_tag)::value;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sv[32];
          tmem_ld32(tS + half * 32, sv);
          tmem_ld_wait();
          if (half == 1) {
            tc_fence_before();
            mbar_arrive(s_free(g));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x = ffma2_(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
            float p0, p1;
            if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
              exp2_poly2_<CLAMP>(x, p0, p1);
            } else {
              float a0, a1;
              upk2(x, a0, a1);
              p0 = ex2f(a0);
              p1 = ex2f(a1);
            }
            if (!LSUM) l2 = fadd2_(l2, pk2(p0, p1));
            pk[half * 16 + i] = pack_bf16(p0, p1);
          }
        }
      };
      if (bound_mode) exp_pass(std::false_type{});  // x >= -2 * BOUND_LIMIT: no clamp needed
      else exp_pass(std::true_type{});
      // ---- the P region (and O) are free once P_{j-1} V_{j-1} has completed
      if (j > 0) {
        mbar_wait(pv_done(g), (j - 1) & 1);
        tc_fence_after();
        if (rescale) {
#pragma unroll
          for (int part = 0; part < (LSUM ? 2 * DH : DH) / 16; ++part) {
            uint32_t o[16];
            tmem_ld16(tO + part * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + part * 16, o);
          }
        }
      }
      tmem_st32(tP, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(g));
    }
    // ---- epilogue: O / l -> bf16, lse
    mbar_wait(pv_done(g), (nkb - 1) & 1);
    tc_fence_after();
    float l;
    if (LSUM) {
      uint32_t o2[16];
      tmem_ld16(tO + DH, o2);
      tmem_ld_wait();
      l = __uint_as_float(o2[0]);
    } else {
      float la, lb;
      upk2(l2, la, lb);
      l = la + lb;
    }
    const float inv = 1.f / l;
    const size_t grow = (size_t)b * L + q0 + g * QT + row;
    uint4* dst = reinterpret_cast<uint4*>(out + grow * C + h * DH);
#pragma unroll
    for (int part = 0; part < DH / 16; ++part) {
      uint32_t o[16];
      tmem_ld16(tO + part * 16, o);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w[i] = pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
      dst[2 * part] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[2 * part + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    if (lse2) lse2[((size_t)b * H + h) * L + q0 + g * QT + row] = m_ref + log2f(l);
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int DH, int POLY, bool LSUM>
int launch_fwd_t(cudaStream_t st, const CUtensorMap& tm, void* out, float* lse2, const float* knmax2, int B, int L,
                 int C, int heads, float scale_log2) {
  auto kern = attn_fwd_tc_kernel<DH, POLY, LSUM>;
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<DH>::SMEM_BYTES));
    configured.cur() = true;
  }
  dim3 grid(L / (NWG * QT), heads, B);
  kern<<<grid, TC_THREADS, Geo<DH>::SMEM_BYTES, st>>>(tm, (bf16*)out, lse2, knmax2, L, C, scale_log2);
  TSD_LAUNCH_CHECK();
  return 0;
}

template <int DH>
int launch_fwd_dh(cudaStream_t st, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C, int heads,
                  int poly) {
  CUtensorMap tm;
  if (make_tmap_2d_sw(&tm, qkv, 2, (uint64_t)B * L, 3 * (uint64_t)C, 3 * (uint64_t)C, DH, 64, DH * 2)) return 1;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  if (ws) {  // max |k|^2 per (sample, head) for the score bound
    TSD_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * B * heads, st));
    const int rows_per_block = 512 / heads;
    attn_knorm_kernel<DH><<<dim3(ceil_div(L, rows_per_block), B), 512, 0, st>>>((const bf16*)qkv, ws, L, C, heads);
    TSD_LAUNCH_CHECK();
  }
  static int lsum = -1;  // TSD_ATTN_TC_LSUM=1: softmax row sum from the tensor core (head_dim 16)
  if (lsum < 0) { const char* e = getenv("TSD_ATTN_TC_LSUM"); lsum = e ? atoi(e) : 0; }
  if constexpr (DH == 16) {
    if (lsum) {
      switch (poly) {
        case 0: return launch_fwd_t<DH, 0, true>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
        case 2: return launch_fwd_t<DH, 2, true>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
        case 3: return launch_fwd_t<DH, 3, true>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
        default: return launch_fwd_t<DH, 4, true>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
      }
    }
  }
  switch (poly) {
    case 0: return launch_fwd_t<DH, 0, false>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
    case 2: return launch_fwd_t<DH, 2, false>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
    case 3: return launch_fwd_t<DH, 3, false>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
    default: return launch_fwd_t<DH, 4, false>(st, tm, out, lse2, ws, B, L, C, heads, scale_log2);
  }
}

}  // namespace

bool attn_tc_supported(int L, int C, int heads) {
  if (C % heads != 0 || heads > 32) return false;
  const int dh = C / heads;
  return (dh == 16 || dh == 32) && L % (NWG * QT) == 0 && L >= NWG * QT;
}

int launch_attn_fwd_tc(cudaStream_t st, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C,
                       int heads, int poly) {
  TSD_CHECK(attn_tc_supported(L, C, heads), "attn_fwd_tc: unsupported shape L=%d C=%d heads=%d", L, C, heads);
  if (C / heads == 16) return launch_fwd_dh<16>(st, qkv, out, lse2, ws, B, L, C, heads, poly);
  return launch_fwd_dh<32>(st, qkv, out, lse2, ws, B, L, C, heads, poly);
}

}  // namespace tsd
