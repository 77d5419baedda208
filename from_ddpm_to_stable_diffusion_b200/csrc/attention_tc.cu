// tcgen05 / TMEM self-attention for head_dim 16 (the three 64x64 and two 32x32 attention stages, L % 256 == 0).
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

namespace tsd {
namespace {

constexpr int DH = 16;
constexpr int QT = 128;  // query rows per softmax warpgroup (= TMEM lanes)
constexpr int NWG = 2;
constexpr int KB = 64;   // keys per block
constexpr int NST = 6;   // K/V ring stages
constexpr int ROWB = DH * 2;
constexpr int Q_BYTES = NWG * QT * ROWB;
constexpr int KV_BYTES = KB * ROWB;
constexpr int STAGE_BYTES = 2 * KV_BYTES;
constexpr int TC_THREADS = NWG * 128 + 64;  // warps 0-7 softmax (warp % 4 = TMEM sub-partition), 8 = TMA, 9 = MMA
constexpr int WG_COLS = 128;  // TMEM columns per warpgroup: S [0,64) (P overwrites [0,32)), O [64,80)
constexpr int O_COL = 64;
constexpr int TMEM_COLS = NWG * WG_COLS;  // 256: two CTAs per SM
constexpr uint32_t SW32 = 6;
constexpr int SMEM_BYTES = 1024 + Q_BYTES + NST * STAGE_BYTES + 256;
constexpr float RESCALE_THRESHOLD = 8.f;

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
__device__ __forceinline__ void exp2_poly2_(uint64_t x, float& o0, float& o1) {
  float x0, x1;
  upk2(x, x0, x1);
  const uint64_t xc = pk2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
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

template <int POLY>
__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse2,
                   int L, int C, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + Q_BYTES;
  const uint32_t bar_base = sKV + NST * STAGE_BYTES;
  // barriers (8 B each): q_full, kv_full[NST], kv_empty[NST], s_full[NWG], p_full[NWG]
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + NST + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 2 * NST + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (1 + 2 * NST + NWG + g); };
  const uint32_t tmem_slot = bar_base + 8u * (1 + 2 * NST + 2 * NWG);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
      mbar_init(p_full(g), QT);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 8) {
    if (lane == 0) {
      // =========================================================== TMA producer
      const int row_base = b * L;
      mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int i = 0; i < NWG * QT / 64; ++i)
        tma_load_2d(sQ + i * 64 * ROWB, &tmQKV, q_full, h * DH, row_base + q0 + i * 64);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % NST;
        mbar_wait(kv_empty(s), ((j / NST) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full(s), STAGE_BYTES);
        tma_load_2d(sKV + s * STAGE_BYTES, &tmQKV, kv_full(s), C + h * DH, row_base + j * KB);
        tma_load_2d(sKV + s * STAGE_BYTES + KV_BYTES, &tmQKV, kv_full(s), 2 * C + h * DH, row_base + j * KB);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // =========================================================== MMA issuer (serves whichever warpgroup is ready)
      constexpr uint32_t idescS = umma_idesc_bf16(QT, KB, 0, 0);
      constexpr uint32_t idescPV = umma_idesc_bf16(QT, DH, 0, 1);
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_S = [&](int j, int g) {
        const int s = j % NST;
        mbar_wait(kv_full(s), (j / NST) & 1);
        tc_fence_after();
        const uint64_t da = umma_smem_desc_sw(sQ + g * QT * ROWB, 0, 8 * ROWB, SW32);
        const uint64_t db = umma_smem_desc_sw(sKV + s * STAGE_BYTES, 0, 8 * ROWB, SW32);
        umma_bf16(tmem_base + g * WG_COLS, da, db, idescS, 0u);
        umma_commit(s_full(g));
      };
      for (int g = 0; g < NWG; ++g) issue_S(0, g);
      int jg[NWG];
      for (int g = 0; g < NWG; ++g) jg[g] = 0;
      int remaining = NWG * nkb;
      uint32_t spins = 0;
      while (remaining > 0) {
#pragma unroll
        for (int g = 0; g < NWG; ++g) {
          const int j = jg[g];
          if (j >= nkb || !mbar_try_wait(p_full(g), j & 1)) continue;
          spins = 0;
          tc_fence_after();
          const int s = j % NST;
          const uint32_t sV = sKV + s * STAGE_BYTES + KV_BYTES;
          const uint32_t tP = tmem_base + g * WG_COLS;
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t db = umma_smem_desc_sw(sV + k * 16 * ROWB, 0, 8 * ROWB, SW32);
            umma_bf16_ts(tP + O_COL, tP + k * 8, db, idescPV, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(kv_empty(s));
          // S_{j+1} overwrites P_j: the tensor pipe executes in issue order.  Its commit also covers P_j V_j,
          // so "S_{j+1} ready" tells the softmax warps that O is up to date; the last commit stands in for it.
          if (j + 1 < nkb) issue_S(j + 1, g);
          else umma_commit(s_full(g));
          jg[g] = j + 1;
          --remaining;
        }
        if (++spins > TSD_SPIN_LIMIT) {
          printf("tsd: attention MMA issuer timeout (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
          __trap();
        }
      }
    }
  } else {
    // =========================================================== softmax: one query row per thread
    const int g = warp >> 2;
    const int sub = warp & 3;  // TMEM sub-partition of this warp
    const int row = sub * 32 + lane;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + g * WG_COLS;
    const uint32_t tO = tS + O_COL;
    const uint64_t c2 = pk2(scale_log2, scale_log2);
    float m_ref = -INFINITY;
    uint64_t l2 = pk2(0.f, 0.f);
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(s_full(g), j & 1);
      tc_fence_after();
      // ---- pass 1: block maximum
      float mx0, mx1;
      {
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
      }
      const float bm = fmaxf(mx0, mx1) * scale_log2;
      const bool need = bm > m_ref + RESCALE_THRESHOLD;  // first block: m_ref = -inf
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2f(m_ref - bm) : 1.f;  // 0 on the first block
        if (need) m_ref = bm;
        float la, lb;
        upk2(l2, la, lb);
        l2 = pk2(la * alpha, lb * alpha);
        if (j > 0) {  // O already holds P_{j-1} V_{j-1}: the commit behind s_full covered it
          uint32_t o[16];
          tmem_ld16(tO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tO, o);
        }
      }
      // ---- pass 2: P = 2^(s*c - m_ref) as bf16 pairs, written over S columns [0, 32)
      const uint64_t nm2 = pk2(-m_ref, -m_ref);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32];
        tmem_ld32(tS + half * 32, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t x = ffma2_(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
          float p0, p1;
          if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
            exp2_poly2_(x, p0, p1);
          } else {
            float a0, a1;
            upk2(x, a0, a1);
            p0 = ex2f(a0);
            p1 = ex2f(a1);
          }
          l2 = fadd2_(l2, pk2(p0, p1));
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st16(tS + half * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(g));
    }
    // ---- epilogue: O / l -> bf16, lse
    mbar_wait(s_full(g), nkb & 1);
    tc_fence_after();
    uint32_t o[16];
    tmem_ld16(tO, o);
    tmem_ld_wait();
    float la, lb;
    upk2(l2, la, lb);
    const float l = la + lb;
    const float inv = 1.f / l;
    const size_t grow = (size_t)b * L + q0 + g * QT + row;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      w[i] = pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
    uint4* dst = reinterpret_cast<uint4*>(out + grow * C + h * DH);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
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

template <int POLY>
int launch_fwd_t(cudaStream_t st, const CUtensorMap& tm, void* out, float* lse2, int B, int L, int C, int heads,
                 float scale_log2) {
  auto kern = attn_fwd_tc_kernel<POLY>;
  static bool configured = false;
  if (!configured) {
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid(L / (NWG * QT), heads, B);
  kern<<<grid, TC_THREADS, SMEM_BYTES, st>>>(tm, (bf16*)out, lse2, L, C, scale_log2);
  TSD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool attn_tc_supported(int L, int C, int heads) {
  return C % heads == 0 && C / heads == DH && L % (NWG * QT) == 0 && L >= NWG * QT;
}

int launch_attn_fwd_tc(cudaStream_t st, const void* qkv, void* out, float* lse2, int B, int L, int C, int heads,
                       int poly) {
  TSD_CHECK(attn_tc_supported(L, C, heads), "attn_fwd_tc: unsupported shape L=%d C=%d heads=%d", L, C, heads);
  CUtensorMap tm;
  if (make_tmap_2d_sw(&tm, qkv, 2, (uint64_t)B * L, 3 * (uint64_t)C, 3 * (uint64_t)C, DH, 64, 32)) return 1;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  switch (poly) {
    case 0: return launch_fwd_t<0>(st, tm, out, lse2, B, L, C, heads, scale_log2);
    case 2: return launch_fwd_t<2>(st, tm, out, lse2, B, L, C, heads, scale_log2);
    case 4: return launch_fwd_t<4>(st, tm, out, lse2, B, L, C, heads, scale_log2);
    default: return launch_fwd_t<3>(st, tm, out, lse2, B, L, C, heads, scale_log2);
  }
}

}  // namespace tsd
