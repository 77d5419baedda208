// tcgen05 / TMEM self-attention backward for head_dim 32 with L % 128 == 0: the one-pass kernel of attention_bwd_tc.cu
// (head_dim 16) re-laid-out for 64-byte rows.  Reference: autograd of SelfAttention.forward, diffusion.py:46-58, for the
// C = 256 blocks (diffusion.py:212,216,225,237,240).  Same contract: qkv bf16 [B*L][3C], dout bf16 [B*L][C], lse2 / delta
// fp32 [B][heads][L]; dK / dV go to dqkv as bf16, dQ is reduced into the fp32 workspace [B][heads][L][32] and converted
// by attn_dq_convert_kernel<32>.
//
// What differs from head_dim 16 (see that file for the schedule: softmax warpgroups own 32 query columns of every
// 128 x 128 sub-tile, scores transposed with keys on the TMEM lanes, lse2 / delta folded into the score products through a
// ones tile, P^T to TMEM, dS^T to one shared-memory tile read K-major for dK and MN-major for dQ):
//   * TMEM: with 32-column accumulators the head_dim-16 layout needs 616 columns.  Here a CTA owns ONE 128-key block
//     (dV / dK / dQ 32 columns each) and the K tile stays in shared memory as the A operand of S'^T = K Q^T (an SS-mode
//     UMMA; V still sits in TMEM for dP'^T = V dO^T): S 256 | P 128 | dV 32 | dK 32 | dQ 32 | V 16 | ones 8 = 504 columns.
//   * Q / K / V / dO tiles have 64-byte rows (SWIZZLE_64B boxes {32 ch, 64 rows}); each score product takes two K-steps.
//   * a query tile is complete after one sub-tile, so dQ leaves TMEM every sub-tile and the Q ring is refilled two tiles
//     ahead of the score issuers.
#include "attention_tc.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace tsd {
namespace {

constexpr int DH = 32, ROWB = 64;
constexpr int KT = 128;        // keys per CTA = TMEM lanes
constexpr int QT = 128;        // queries per tile
constexpr int NWG = 4;         // softmax warpgroups
constexpr int CW = QT / NWG;   // query columns per warpgroup and tile
constexpr int NSI = 2;         // issuers of S'^T / dP'^T, NWG / NSI warpgroups each
constexpr int NSTQ = 4;        // Q / dO / lse / delta ring
constexpr int W_DRAIN = NWG * 4, W_S = W_DRAIN + 4, W_P = W_S + NSI;  // first drain warp, S issuer, product issuer
constexpr int BT_THREADS = (W_P + 3) * 32;
// TMEM columns (fp32 unless noted)
constexpr int S_COL = 0;       // + g * 2 CW: S'^T [0, CW), dP'^T [CW, 2 CW)
constexpr int P_COL = 256;     // + buf * 64: P^T bf16 pairs, 128 queries
constexpr int DV_COL = 384, DK_COL = 416, DQ_COL = 448;
constexpr int VA_COL = 480;    // bf16 A operand: V (32 elements = 16 columns)
constexpr int ONES_COL = 496;  // bf16 A operand: 1 in K slots 0-2, 0 elsewhere
constexpr int TMEM_COLS = 512;
// shared memory (offsets from a 1024-byte aligned base)
constexpr int OFF_K = 0, OFF_V = KT * ROWB;                      // 8 KB each
constexpr int OFF_Q = 2 * KT * ROWB;                             // ring of stages:
constexpr int ST_DO = QT * ROWB;                                 //   Q 8 KB, dO 8 KB,
constexpr int ST_LSE = 2 * QT * ROWB, ST_DELTA = ST_LSE + QT * 4;  // lse2, delta fp32 (512 B each),
constexpr int ST_LSET = ST_DELTA + QT * 4, ST_DELT = ST_LSET + QT * 32;  // their split bf16 tiles [q][16] (4 KB each)
constexpr int QSTAGE = ST_DELT + QT * 32;                        // 25600
constexpr int Q_TX = ST_LSET;                                    // bytes written by TMA per stage
constexpr int OFF_DS = OFF_Q + NSTQ * QSTAGE;
constexpr int DS_BYTES = KT * QT * 2;                            // 32 KB, two 64-query chunks of 16 KB
constexpr int OFF_BAR = OFF_DS + 2 * DS_BYTES;
constexpr int NBAR = 2 + 3 * NSTQ + 2 * NWG + 2 * 2 + 2;
constexpr int BT_SMEM = 1024 + OFF_BAR + NBAR * 8 + 16;
static_assert(OFF_DS % 1024 == 0 && QSTAGE % 1024 == 0 && ST_LSET % 1024 == 0, "swizzled tiles must keep their alignment");

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fmul2_(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_tc32_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const float* __restrict__ lse2, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                     float* __restrict__ ws, int L, int C, float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base + OFF_K, sV = smem_base + OFF_V;
  auto sQ = [&](int s) { return smem_base + OFF_Q + s * QSTAGE; };
  auto sDS = [&](int buf) { return smem_base + OFF_DS + buf * DS_BYTES; };
  const uint32_t bar_base = smem_base + OFF_BAR;
  const uint32_t kv_full = bar_base;                 // K / V tiles landed (TMA)
  const uint32_t ka_full = bar_base + 8u;            // V / ones copied to TMEM
  auto q_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto q_empty = [&](int s) { return bar_base + 8u * (2 + NSTQ + s); };
  auto ld_full = [&](int s) { return bar_base + 8u * (2 + 2 * NSTQ + s); };  // split lse2 / delta tiles built
  constexpr int B0 = 2 + 3 * NSTQ;
  auto s_full = [&](int g) { return bar_base + 8u * (B0 + g); };
  auto s_free = [&](int g) { return bar_base + 8u * (B0 + NWG + g); };
  auto pds_full = [&](int buf) { return bar_base + 8u * (B0 + 2 * NWG + buf); };
  auto mma_done = [&](int buf) { return bar_base + 8u * (B0 + 2 * NWG + 2 + buf); };
  const uint32_t dq_full = bar_base + 8u * (B0 + 2 * NWG + 4);
  const uint32_t dq_free = bar_base + 8u * (B0 + 2 * NWG + 5);
  const uint32_t tmem_slot = bar_base + 8u * NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * NBAR);

  // warp-uniform for the compiler: the issuer warps run converged and only elect a lane around the async instructions
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const int kb0 = blockIdx.x * KT;
  const int nq = L / QT;  // sub-tile t = query tile t (one key block per CTA)
  const int row_base = b * L;

  if (warp == W_P && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    mbar_init(ka_full, NWG * 128);
    for (int s = 0; s < NSTQ; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 2);  // the dV and dK issuers read the Q / dO tiles
      mbar_init(ld_full(s), 128);
    }
    for (int g = 0; g < NWG; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_free(g), 128);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(pds_full(i), NWG * 128);
      mbar_init(mma_done(i), 3);  // one commit per product issuer
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    fence_mbar_init();
  }
  if (warp == W_S) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // descriptors: the start-address field is (addr >> 4) in the low word, so a byte offset adds (offset >> 4)
  auto d64 = [&](uint32_t addr) { return umma_smem_desc_sw(addr, 0, 8 * ROWB, 4); };  // 64-byte rows, SWIZZLE_64B
  auto d32 = [&](uint32_t addr) { return umma_smem_desc_sw(addr, 0, 8 * 32, 6); };    // 32-byte rows, SWIZZLE_32B

  if (warp >= W_S && warp < W_P) {
    // =========================================================== issuers of S'^T / dP'^T (the first one also feeds TMA)
    constexpr uint32_t idescS = umma_idesc_bf16(KT, CW, 0, 0);
    const int si = warp - W_S;
    const int g_lo = si * (NWG / NSI);
    const float* lse_h = lse2 + ((size_t)b * H + h) * L;
    const float* delta_h = delta + ((size_t)b * H + h) * L;
    auto load_q = [&](int j) {
      const int s = j % NSTQ;
      mbar_wait(q_empty(s), ((j / NSTQ) & 1) ^ 1);
      const uint32_t st = sQ(s);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full(s), Q_TX);
#pragma unroll
        for (int i = 0; i < QT / 64; ++i) {
          tma_load_2d(st + i * 64 * ROWB, &tmQKV, q_full(s), h * DH, row_base + j * QT + i * 64);
          tma_load_2d(st + ST_DO + i * 64 * ROWB, &tmDO, q_full(s), h * DH, row_base + j * QT + i * 64);
        }
        bulk_load_1d(st + ST_LSE, lse_h + j * QT, QT * 4, q_full(s));
        bulk_load_1d(st + ST_DELTA, delta_h + j * QT, QT * 4, q_full(s));
      }
      __syncwarp();
    };
    if (si == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * KT * ROWB);
#pragma unroll
        for (int i = 0; i < KT / 64; ++i) {
          tma_load_2d(sK + i * 64 * ROWB, &tmQKV, kv_full, C + h * DH, row_base + kb0 + i * 64);
          tma_load_2d(sV + i * 64 * ROWB, &tmQKV, kv_full, 2 * C + h * DH, row_base + kb0 + i * 64);
        }
      }
      __syncwarp();
      for (int j = 0; j < 2 && j < nq; ++j) load_q(j);
    }
    const uint32_t tOnes = tmem_base + ONES_COL, tVA = tmem_base + VA_COL;
    const uint64_t dk = d64(sK);
    auto issue_S = [&](int t, int g) {
      const uint32_t qs = sQ(t % NSTQ);
      const uint64_t dq = d64(qs + g * CW * ROWB);
      const uint64_t ddo = d64(qs + ST_DO + g * CW * ROWB);
      const uint64_t dl = d32(qs + ST_LSET + g * CW * 32);
      const uint64_t dd = d32(qs + ST_DELT + g * CW * 32);
      const uint32_t tS = tmem_base + S_COL + g * 2 * CW;
      if (elect_one()) {
        umma_bf16(tS, dk, dq, idescS, 0u);                      // K Q^T: A from shared memory, two K-steps of 32 bytes
        umma_bf16(tS, dk + 2, dq + 2, idescS, 1u);
        umma_bf16_ts(tS, tOnes, dl, idescS, 1u);                // - lse2 / c
        umma_bf16_ts(tS + CW, tVA, ddo, idescS, 0u);            // V dO^T: A from TMEM
        umma_bf16_ts(tS + CW, tVA + 8, ddo + 2, idescS, 1u);
        umma_bf16_ts(tS + CW, tOnes, dd, idescS, 1u);           // - delta
        umma_commit(s_full(g));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    mbar_wait(ka_full, 0);
    mbar_wait(q_full(0), 0);
    mbar_wait(ld_full(0), 0);
    tc_fence_after();
#pragma unroll
    for (int gi = 0; gi < NWG / NSI; ++gi) issue_S(0, g_lo + gi);
    for (int t = 0; t + 1 < nq; ++t) {
      const int jn = t + 1;
      mbar_wait(q_full(jn % NSTQ), (jn / NSTQ) & 1);
      mbar_wait(ld_full(jn % NSTQ), (jn / NSTQ) & 1);
#pragma unroll
      for (int gi = 0; gi < NWG / NSI; ++gi) {
        mbar_wait(s_free(g_lo + gi), t & 1);  // sub-tile t of this warpgroup sits in its registers
        tc_fence_after();
        issue_S(t + 1, g_lo + gi);
      }
      // refill the ring two tiles ahead: tile t + 2 goes where tile t - 2 was (released long ago, no stall here)
      if (si == 0 && t + 2 < nq) load_q(t + 2);
    }
  } else if (warp >= W_P) {
    // =========================================================== issuers of dV, dK, dQ
    constexpr uint32_t idescKN = umma_idesc_bf16(KT, DH, 0, 1);   // A K-major (or TMEM), B MN-major
    constexpr uint32_t idescDQ = umma_idesc_bf16(QT, DH, 1, 1);   // A MN-major, B MN-major
    const int which = warp - W_P;
    mbar_wait(kv_full, 0);
    const uint64_t dKt = d64(sK);
    for (int t = 0; t < nq; ++t) {
      const int buf = t & 1;
      const uint32_t st = sQ(t % NSTQ);
      const uint32_t accKV = t > 0 ? 1u : 0u;
      mbar_wait(pds_full(buf), (t >> 1) & 1);
      if (which == 2 && t > 0) mbar_wait(dq_free, (t - 1) & 1);  // dQ of tile t - 1 has left TMEM
      tc_fence_after();
      if (!elect_one()) {
      } else if (which == 0) {
        const uint32_t tP = tmem_base + P_COL + buf * 64;
        const uint64_t db = d64(st + ST_DO);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          umma_bf16_ts(tmem_base + DV_COL, tP + k * 8, db + (uint64_t)(k * (16 * ROWB / 16)), idescKN, k > 0 ? 1u : accKV);
      } else if (which == 1) {
        const uint64_t da = umma_smem_desc(sDS(buf), 0, 1024);
        const uint64_t db = d64(st);
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)
          umma_bf16(tmem_base + DK_COL, da + (uint64_t)(((k >> 2) * (DS_BYTES / 2) + (k & 3) * 32) / 16),
                    db + (uint64_t)(k * (16 * ROWB / 16)), idescKN, k > 0 ? 1u : accKV);
      } else {
        const uint64_t da = umma_smem_desc(sDS(buf), DS_BYTES / 2, 1024);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_bf16(tmem_base + DQ_COL, da + (uint64_t)(k * (2048 / 16)), dKt + (uint64_t)(k * (16 * ROWB / 16)), idescDQ,
                    k > 0 ? 1u : 0u);
      }
      if (elect_one()) {
        umma_commit(mma_done(buf));
        if (which == 2) umma_commit(dq_full);
        else umma_commit(q_empty(t % NSTQ));
      }
      __syncwarp();
    }
  } else if (warp < W_DRAIN) {
    // =========================================================== softmax: one key row per thread
    const int g = warp >> 2, sub = warp & 3;
    const int r = sub * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
    // ---- once: the V rows (and the ones tile) become TMEM A operands
    mbar_wait(kv_full, 0);
    if (g == 0) {
      const uint32_t sw = static_cast<uint32_t>((r >> 1) & 3);  // 64-byte swizzle: 16-byte chunk index ^ address bits 7-8
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(smem_gen + OFF_V + r * ROWB + ((static_cast<uint32_t>(c) ^ sw) << 4));
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      tmem_st16(lane_base + VA_COL, w);
    }
    if (g == 1) {
      const uint32_t w[8] = {0x3f803f80u, 0x00003f80u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(lane_base + ONES_COL, w);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(ka_full);

    const uint32_t tS = lane_base + S_COL + g * 2 * CW, tDP = tS + CW;
    const uint64_t c2 = pk2(scale_log2, scale_log2);
    // dS^T tile: 64-query chunk (g * CW) / 64, 16-byte unit ((g * CW) % 64) / 8 + ..., XOR-swizzled with the row
    const uint32_t ds_row = ((g * CW) >> 6) * (DS_BYTES / 2) + r * 128;
    const uint32_t u0 = ((g * CW) & 63) >> 3;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    for (int t = 0; t < nq; ++t) {
      const int buf = t & 1;
      mbar_wait(s_full(g), t & 1);
      tc_fence_after();
      constexpr int NCH = CW / 16;
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t sv[16], dv[16];
        tmem_ld16(tS + cc * 16, sv);
        tmem_ld16(tDP + cc * 16, dv);
        tmem_ld_wait();
        if (cc == NCH - 1) {
          tc_fence_before();
          mbar_arrive(s_free(g));
        }
        uint32_t pP[8], pD[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // x = c (s - lse2 / c) <= 0 up to rounding: P <= 1
          float a0, a1;
          upk2(fmul2_(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2), a0, a1);
          const float p0 = ex2f(a0), p1 = ex2f(a1);
          float e0, e1;
          upk2(fmul2_(pk2(p0, p1), pk2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1]))), e0, e1);
          pP[i] = pack_bf16(p0, p1);
          pD[i] = pack_bf16(e0, e1);
        }
        if (cc == 0 && t >= 2) {  // P^T / dS^T buffers of sub-tile t - 2 have been consumed
          mbar_wait(mma_done(buf), ((t >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        tmem_st8(lane_base + P_COL + buf * 64 + (g * CW + cc * 16) / 2, pP);
        const uint32_t drow = sDS(buf) + ds_row;
#pragma unroll
        for (int i4 = 0; i4 < 2; ++i4)
          sts128(drow + (((u0 + cc * 2 + i4) ^ sw) << 4), pD[4 * i4], pD[4 * i4 + 1], pD[4 * i4 + 2], pD[4 * i4 + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full(buf));
    }
    // ---- epilogue: dV (warpgroup 0) and dK (warpgroup 1)
    mbar_wait(mma_done((nq - 1) & 1), ((nq - 1) >> 1) & 1);
    tc_fence_after();
    if (g < 2) {
      const bool is_dk = g == 1;
      uint32_t a[32];
      tmem_ld32(lane_base + (is_dk ? DK_COL : DV_COL), a);
      tmem_ld_wait();
      const float mul = is_dk ? scale : 1.f;
      const size_t grow = (size_t)row_base + kb0 + r;
      uint4* dst = reinterpret_cast<uint4*>(dqkv + grow * 3 * C + (is_dk ? C : 2 * C) + h * DH);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(a[2 * i]) * mul, __uint_as_float(a[2 * i + 1]) * mul);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  } else {
    // =========================================================== drain warpgroup: one query row per thread
    const int sub = warp & 3;
    const int r = sub * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
    float* dq_head = ws + (((size_t)b * H + h) * L) * DH;
    const float inv_c = 1.f / scale_log2;
    // -v as three bf16 pieces (24 significant bits), the same 16 bytes in both halves of the 32-byte row so that the
    // tile reads the same with or without the 32-byte swizzle
    auto split3 = [](float v, uint32_t& w0, uint32_t& w1) {
      const bf16 b0 = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(b0);
      const bf16 b1 = __float2bfloat16_rn(r1);
      const bf16 b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));
      w0 = static_cast<uint32_t>(__bfloat16_as_ushort(b0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b1)) << 16);
      w1 = static_cast<uint32_t>(__bfloat16_as_ushort(b2));
    };
    auto build = [&](int j) {
      const int s = j % NSTQ;
      mbar_wait(q_full(s), (j / NSTQ) & 1);
      uint8_t* st = smem_gen + OFF_Q + s * QSTAGE;
      const float lv = reinterpret_cast<const float*>(st + ST_LSE)[r];
      const float dl = reinterpret_cast<const float*>(st + ST_DELTA)[r];
      uint32_t w0, w1;
      split3(-lv * inv_c, w0, w1);
      const int x = (r >> 2) & 1;  // which half goes first: 8 consecutive rows then cover 8 distinct 16-byte bank groups
      uint4* row = reinterpret_cast<uint4*>(st + ST_LSET + r * 32);
      row[x] = make_uint4(w0, w1, 0u, 0u);
      row[x ^ 1] = make_uint4(w0, w1, 0u, 0u);
      split3(-dl, w0, w1);
      row = reinterpret_cast<uint4*>(st + ST_DELT + r * 32);
      row[x] = make_uint4(w0, w1, 0u, 0u);
      row[x ^ 1] = make_uint4(w0, w1, 0u, 0u);
      fence_proxy_async_smem();
      mbar_arrive(ld_full(s));
    };
    build(0);
    if (nq > 1) build(1);
    for (int j = 0; j < nq; ++j) {
      if (j + 2 < nq) build(j + 2);
      mbar_wait(dq_full, j & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(lane_base + DQ_COL, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      float* dst = dq_head + ((size_t)j * QT + r) * DH;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        red_add_v4(dst + 4 * i, __uint_as_float(o[4 * i]) * scale, __uint_as_float(o[4 * i + 1]) * scale,
                   __uint_as_float(o[4 * i + 2]) * scale, __uint_as_float(o[4 * i + 3]) * scale);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == W_S) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

bool attn_bwd_tc32_supported(int L, int C, int heads) {
  static int on = -1;  // TSD_ATTN_BWD_TC32=0: head_dim 32 keeps the two-pass mma.sync kernels (A/B switch)
  if (on < 0) { const char* e = getenv("TSD_ATTN_BWD_TC32"); on = e ? atoi(e) : 1; }
  return on && C % heads == 0 && C / heads == DH && L % KT == 0 && L >= 2 * KT;
}

int launch_attn_bwd_tc32(cudaStream_t st, const void* qkv, const void* dout, const float* lse2, const float* delta,
                         void* dqkv, float* ws, int B, int L, int C, int heads) {
  TSD_CHECK(attn_bwd_tc32_supported(L, C, heads), "attn_bwd_tc32: unsupported shape L=%d C=%d heads=%d", L, C, heads);
  CUtensorMap tmQKV, tmDO;
  if (make_tmap_2d_sw(&tmQKV, qkv, 2, (uint64_t)B * L, 3 * (uint64_t)C, 3 * (uint64_t)C, DH, 64, ROWB)) return 1;
  if (make_tmap_2d_sw(&tmDO, dout, 2, (uint64_t)B * L, (uint64_t)C, (uint64_t)C, DH, 64, ROWB)) return 1;
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    configured.cur() = true;
  }
  const float scale = 1.f / sqrtf((float)DH);
  const dim3 grid(L / KT, heads, B);
  attn_bwd_tc32_kernel<<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale,
                                                        1.4426950408889634f * scale);
  TSD_LAUNCH_CHECK();
  return 0;
}

}  // namespace tsd
