// tcgen05 / TMEM self-attention backward for head_dim 16 with L % 256 == 0: one pass, every contraction on the 5th-gen
// tensor cores, the softmax warps do nothing but the element-wise part.
//
// Reference: autograd of SelfAttention.forward, diffusion.py:46-58.  Same contract as attn_bwd_fused_kernel (attention.cu):
// qkv bf16 [B*L][3C], dout bf16 [B*L][C], lse2 / delta fp32 [B][heads][L]; dK / dV go to dqkv as bf16, dQ is reduced
// into the fp32 workspace [B][heads][L][16] (bulk reduce-add) and converted by attn_dq_convert_kernel.
//
// A CTA owns 256 keys of one (sample, head) as two 128-key halves (TMEM lanes = keys) and walks over all 128-query
// tiles; one "sub-tile" t = (query tile j, key half kh).  Per sub-tile and softmax warpgroup g (CW query columns):
//   S'^T  [128 k x CW q] = K_kh Q_g^T - lse2[q] / c      two UMMAs 128 x CW x 16 with A from TMEM: K_kh (copied there
//   dP'^T [128 k x CW q] = V_kh dO_g^T - delta[q]         once), then a constant tile of ones against a [q][16] tile that
//        holds -lse2/c (resp. -delta) split into three bf16 pieces: the per-column terms come out of the tensor core
//        instead of 2 x 128 broadcast shared-memory loads per thread, which saturated the shared-memory pipe;
//   softmax thread = one key row: P^T = 2^(c S'^T), dS^T = P^T dP'^T;
//        P^T  -> TMEM as bf16 (A operand of dV),
//        dS^T -> shared memory [128 k][128 q] bf16, 128-byte swizzled: read K-major for dK and MN-major for dQ,
//                so the transposition the mma.sync kernel does with stmatrix/ldmatrix disappears
//   dV_kh += P^T dO    8 x UMMA 128 x 16 x 16, A from TMEM, B = dO tile MN-major
//   dK_kh += dS^T Q    8 x UMMA 128 x 16 x 16, A = dS^T tile K-major, B = Q tile MN-major
//   dQ_j  += dS K_kh   8 x UMMA 128 x 16 x 16, A = dS^T tile MN-major, B = K tile MN-major (accumulated over both halves)
// dK / dV stay in TMEM for the CTA's life; dQ_j is drained by a separate warpgroup (tcgen05.ld -> red.global.add.v4.f32
// into the workspace: the shared-memory pipe is the busiest unit of this kernel, so the drain stays off it), which also
// builds the split lse2 / delta tiles of the coming tiles.
//
// Warps: NWG softmax warpgroups (g owns query columns [CW g, CW g + CW) of every tile), one drain warpgroup, NSI issuers of
// S'^T / dP'^T (the first one is also the TMA producer and allocates TMEM), three issuers for dV / dK / dQ: a single
// issuing thread needs ~75 clk per tcgen05.mma and was the critical path.  One CTA per SM (all 512 TMEM columns).
#include "attention_tc.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace tsd {
namespace {

constexpr int DH = 16, ROWB = 32;
constexpr int KT = 128;        // keys per half = TMEM lanes
constexpr int KH = 2;          // key halves per CTA
constexpr int QT = 128;        // queries per tile
constexpr int NWG = 4;         // softmax warpgroups
constexpr int CW = QT / NWG;   // query columns per warpgroup and tile
constexpr int NSI = 2;         // issuers of S'^T / dP'^T, NWG / NSI warpgroups each
constexpr int NSTQ = 4;        // Q / dO / lse / delta ring
constexpr int W_DRAIN = NWG * 4, W_S = W_DRAIN + 4, W_P = W_S + NSI;  // first drain warp, S issuer, product issuer
constexpr int BT_THREADS = (W_P + 3) * 32;
// PIPE variant: the CTA is padded to whole warpgroups (7 x 128 threads) so that setmaxnreg can move registers from the
// service warps (40 each) to the softmax warps (96 each), which then pull a whole sub-tile of scores out of TMEM at
// once and hand the score buffer back before the first exponential.
constexpr int BT_THREADS_PIPE = 7 * 128;
constexpr int REGS_SOFTMAX = 96, REGS_SERVICE = 40;
// TMEM columns (fp32 unless noted)
constexpr int S_COL = 0;       // + g * 2 CW: S'^T [0, CW), dP'^T [CW, 2 CW)
constexpr int P_COL = 256;     // + buf * 64: P^T bf16 pairs, 128 queries
constexpr int DV_COL = 384;    // + kh * 16
constexpr int DK_COL = 416;    // + kh * 16
constexpr int DQ_COL = 448;
constexpr int KA_COL = 464;    // bf16 A operands: K_kh at + kh * 8, V_kh at + 16 + kh * 8
constexpr int ONES_COL = 496;  // bf16 A operand: 1 in K slots 0-2, 0 elsewhere
constexpr int TMEM_COLS = 512;
// shared memory (offsets from a 1024-byte aligned base)
constexpr int OFF_K = 0, OFF_V = KH * KT * ROWB;                 // 8 KB each
constexpr int OFF_Q = 2 * KH * KT * ROWB;                        // ring of stages:
constexpr int ST_DO = QT * ROWB;                                 //   Q 4 KB, dO 4 KB,
constexpr int ST_LSE = 2 * QT * ROWB, ST_DELTA = ST_LSE + QT * 4;  // lse2, delta fp32 (512 B each),
constexpr int ST_LSET = ST_DELTA + QT * 4, ST_DELT = ST_LSET + QT * ROWB;  // their split bf16 tiles [q][16] (4 KB each)
constexpr int QSTAGE = ST_DELT + QT * ROWB;                      // 17408
constexpr int Q_TX = ST_LSET;                                    // bytes written by TMA per stage
constexpr int OFF_DS = OFF_Q + NSTQ * QSTAGE;
constexpr int DS_BYTES = KT * QT * 2;                            // 32 KB, two 64-query chunks of 16 KB
constexpr int OFF_BAR = OFF_DS + 2 * DS_BYTES;
constexpr int NBAR = 2 + 3 * NSTQ + 2 * NWG + 2 * 2 + 2;
constexpr int BT_SMEM = 1024 + OFF_BAR + NBAR * 8 + 16;
static_assert(OFF_DS % 1024 == 0 && QSTAGE % 1024 == 0, "swizzled tiles must keep their alignment");
static_assert(NWG % NSI == 0 && (NWG == 2 || NWG == 4), "warpgroup split");

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
// 2^x for a packed pair on the FMA pipe (the forward kernel's Cody-Waite cubic, attention_tc.cu): |rel err| < 2e-4, far
// under the bf16 rounding of P.  x <= 0 up to rounding here; clamped at -126 for the exponent insertion (2^-126 ~ 0).
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

// SHARED: the score products of a sub-tile are issued ONCE for all warpgroups (four UMMAs with N = 128 instead of
// sixteen with N = 32: the in-order tensor pipe is paid per instruction, not per column), the warpgroups read their
// 32-column slices of the shared S'^T / dP'^T regions and hand them back together.
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// POLY > 0: every POLY-th pair of exponentials runs as a cubic on the FMA pipe instead of MUFU.EX2 (as in the forward).
template <bool SHARED, bool PIPE, int POLY = 0>
__global__ void __launch_bounds__(PIPE ? BT_THREADS_PIPE : BT_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse2, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                   float* __restrict__ ws, int L, int C, float scale, float scale_log2, int stagger) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base + OFF_K, sV = smem_base + OFF_V;
  auto sQ = [&](int s) { return smem_base + OFF_Q + s * QSTAGE; };
  auto sDS = [&](int buf) { return smem_base + OFF_DS + buf * DS_BYTES; };
  const uint32_t bar_base = smem_base + OFF_BAR;
  const uint32_t kv_full = bar_base;                 // K / V tiles landed (TMA)
  const uint32_t ka_full = bar_base + 8u;            // K / V / ones copied to TMEM
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
  const int kb0 = blockIdx.x * (KH * KT);
  const int nq = L / QT;
  const int NT = nq * KH;
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
      mbar_init(s_free(g), SHARED ? NWG * 128 : 128);
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
  // register hand-over: one instruction for all service warpgroups here, one for the softmax warpgroups at the top of
  // their branch (every thread of a warpgroup must execute the same setmaxnreg)
  if (PIPE && warp >= W_DRAIN) setmaxnreg_dec<REGS_SERVICE>();
  // descriptors: the start-address field is (addr >> 4) in the low word, so a byte offset adds (offset >> 4)
  auto d32 = [&](uint32_t addr) { return umma_smem_desc_sw(addr, 0, 8 * ROWB, 6); };

  if (SHARED && warp >= W_S && warp < W_P) {
    if (warp == W_S) {
      // =========================================================== one issuer of S'^T / dP'^T for all warpgroups (+ TMA)
      constexpr uint32_t idescS = umma_idesc_bf16(KT, QT, 0, 0);
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
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * KH * KT * ROWB);
#pragma unroll
        for (int i = 0; i < KH * KT / 64; ++i) {
          tma_load_2d(sK + i * 64 * ROWB, &tmQKV, kv_full, C + h * DH, row_base + kb0 + i * 64);
          tma_load_2d(sV + i * 64 * ROWB, &tmQKV, kv_full, 2 * C + h * DH, row_base + kb0 + i * 64);
        }
      }
      __syncwarp();
      for (int j = 0; j < NSTQ - 1 && j < nq; ++j) load_q(j);
      const uint32_t tOnes = tmem_base + ONES_COL;
      auto issue_S = [&](int t) {
        const int j = t >> 1, kh = t & 1;
        const uint64_t dq = d32(sQ(j % NSTQ));
        const uint32_t tS = tmem_base + S_COL, tDP = tS + QT;
        if (elect_one()) {
          umma_bf16_ts(tS, tmem_base + KA_COL + kh * 8, dq, idescS, 0u);
          umma_bf16_ts(tS, tOnes, dq + (uint64_t)(ST_LSET / 16), idescS, 1u);
          umma_bf16_ts(tDP, tmem_base + KA_COL + 16 + kh * 8, dq + (uint64_t)(ST_DO / 16), idescS, 0u);
          umma_bf16_ts(tDP, tOnes, dq + (uint64_t)(ST_DELT / 16), idescS, 1u);
          umma_commit(s_full(0));
        }
        __syncwarp();
      };
      mbar_wait(ka_full, 0);
      mbar_wait(q_full(0), 0);
      mbar_wait(ld_full(0), 0);
      tc_fence_after();
      issue_S(0);
      for (int t = 0; t + 1 < NT; ++t) {
        const int j = t >> 1, kh = t & 1;
        if (kh == 1) {
          const int jn = j + 1;
          mbar_wait(q_full(jn % NSTQ), (jn / NSTQ) & 1);
          mbar_wait(ld_full(jn % NSTQ), (jn / NSTQ) & 1);
        }
        mbar_wait(s_free(0), t & 1);  // every warpgroup holds its slice of sub-tile t in registers
        tc_fence_after();
        issue_S(t + 1);
        if (kh == 0 && j + NSTQ - 1 < nq) load_q(j + NSTQ - 1);
      }
    }
  } else if (warp >= W_S && warp < W_P) {
    {
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
          mbar_arrive_expect_tx(kv_full, 2 * KH * KT * ROWB);
#pragma unroll
          for (int i = 0; i < KH * KT / 64; ++i) {
            tma_load_2d(sK + i * 64 * ROWB, &tmQKV, kv_full, C + h * DH, row_base + kb0 + i * 64);
            tma_load_2d(sV + i * 64 * ROWB, &tmQKV, kv_full, 2 * C + h * DH, row_base + kb0 + i * 64);
          }
        }
        __syncwarp();
        for (int j = 0; j < NSTQ - 1 && j < nq; ++j) load_q(j);
      }
      const uint32_t tOnes = tmem_base + ONES_COL;
      auto issue_S = [&](int t, int g) {
        const int j = t >> 1, kh = t & 1;
        const uint64_t dq = d32(sQ(j % NSTQ) + g * CW * ROWB);
        const uint32_t tS = tmem_base + S_COL + g * 2 * CW;
        if (elect_one()) {
          umma_bf16_ts(tS, tmem_base + KA_COL + kh * 8, dq, idescS, 0u);
          umma_bf16_ts(tS, tOnes, dq + (uint64_t)(ST_LSET / 16), idescS, 1u);
          umma_bf16_ts(tS + CW, tmem_base + KA_COL + 16 + kh * 8, dq + (uint64_t)(ST_DO / 16), idescS, 0u);
          umma_bf16_ts(tS + CW, tOnes, dq + (uint64_t)(ST_DELT / 16), idescS, 1u);
          umma_commit(s_full(g));
        }
        __syncwarp();
      };
      mbar_wait(ka_full, 0);
      mbar_wait(q_full(0), 0);
      mbar_wait(ld_full(0), 0);
      // stagger: the second issuer's warpgroups start half a sub-tile after the first one's, so that the four softmax
      // warps of a scheduler are not all in the same phase (TMEM load / exponentials / stores) at the same time
      if ((stagger & 0xff) && si > 0) mbar_wait(s_free(0), 0);
      tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < NWG / NSI; ++gi) issue_S(0, g_lo + gi);
      for (int t = 0; t + 1 < NT; ++t) {
        const int j = t >> 1, kh = t & 1;
        if (kh == 1) {
          const int jn = j + 1;
          mbar_wait(q_full(jn % NSTQ), (jn / NSTQ) & 1);
          mbar_wait(ld_full(jn % NSTQ), (jn / NSTQ) & 1);
        }
#pragma unroll
        for (int gi = 0; gi < NWG / NSI; ++gi) {
          mbar_wait(s_free(g_lo + gi), t & 1);  // sub-tile t of this warpgroup sits in its registers
          tc_fence_after();
          issue_S(t + 1, g_lo + gi);
        }
        // refill the ring: tile j + NSTQ - 1 goes where tile j - 1 was (released by its key half 1 products)
        if (si == 0 && kh == 0 && j + NSTQ - 1 < nq) load_q(j + NSTQ - 1);
      }
    }
  } else if (warp >= W_P + 3) {
    // padding warps of the PIPE variant: nothing to do until the final barrier
  } else if (warp >= W_P) {
    {
      // =========================================================== issuers of dV, dK, dQ
      constexpr uint32_t idescKN = umma_idesc_bf16(KT, DH, 0, 1);   // A K-major (or TMEM), B MN-major
      constexpr uint32_t idescDQ = umma_idesc_bf16(QT, DH, 1, 1);   // A MN-major, B MN-major
      const int which = warp - W_P;
      const int dbg = stagger >> 8;  // TSD_ATTN_BWD_TC_DBG (timing experiments, WRONG results): 1 / 2 / 4 = no dV / dK / dQ products
      mbar_wait(kv_full, 0);
      const uint64_t dKt = d32(sK);
      for (int t = 0; t < NT; ++t) {
        const int j = t >> 1, kh = t & 1, buf = t & 1;
        const uint32_t st = sQ(j % NSTQ);
        const uint32_t accKV = j > 0 ? 1u : 0u;
        mbar_wait(pds_full(buf), (t >> 1) & 1);
        if (which == 2 && kh == 0 && j > 0) mbar_wait(dq_free, (j - 1) & 1);  // dQ of tile j - 1 has left TMEM
        tc_fence_after();
        if (!elect_one() || ((dbg >> which) & 1)) {
        } else if (which == 0) {
          const uint32_t tP = tmem_base + P_COL + buf * 64;
          const uint64_t db = d32(st + ST_DO);
#pragma unroll
          for (int k = 0; k < QT / 16; ++k)
            umma_bf16_ts(tmem_base + DV_COL + kh * 16, tP + k * 8, db + (uint64_t)(k * (16 * ROWB / 16)), idescKN,
                         k > 0 ? 1u : accKV);
        } else if (which == 1) {
          const uint64_t da = umma_smem_desc(sDS(buf), 0, 1024);
          const uint64_t db = d32(st);
#pragma unroll
          for (int k = 0; k < QT / 16; ++k)
            umma_bf16(tmem_base + DK_COL + kh * 16, da + (uint64_t)(((k >> 2) * (DS_BYTES / 2) + (k & 3) * 32) / 16),
                      db + (uint64_t)(k * (16 * ROWB / 16)), idescKN, k > 0 ? 1u : accKV);
        } else {
          const uint64_t da = umma_smem_desc(sDS(buf), DS_BYTES / 2, 1024);
          const uint64_t db = dKt + (uint64_t)(kh * (KT * ROWB / 16));
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            umma_bf16(tmem_base + DQ_COL, da + (uint64_t)(k * (2048 / 16)), db + (uint64_t)(k * (16 * ROWB / 16)),
                      idescDQ, (k > 0 || kh > 0) ? 1u : 0u);
        }
        if (elect_one()) {
          umma_commit(mma_done(buf));
          if (kh == 1) {
            if (which == 2) umma_commit(dq_full);
            else umma_commit(q_empty(j % NSTQ));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < W_DRAIN) {
    // =========================================================== softmax: one key row per thread
    if (PIPE) setmaxnreg_inc<REGS_SOFTMAX>();
    const int g = warp >> 2, sub = warp & 3;
    const int r = sub * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(sub * 32) << 16);
    // ---- once: K / V rows of both halves (and the ones tile) become TMEM A operands
    mbar_wait(kv_full, 0);
    for (int it = g; it < 4; it += NWG) {
      const int kh = it & 1;
      const uint32_t src = ((it >> 1) ? OFF_V : OFF_K) + (kh * KT + r) * ROWB;
      const uint32_t x = static_cast<uint32_t>((r >> 2) & 1) << 4;  // 32-byte swizzle: 16-byte halves swap every 4 rows
      const uint4 lo = *reinterpret_cast<const uint4*>(smem_gen + src + x);
      const uint4 hi = *reinterpret_cast<const uint4*>(smem_gen + src + (x ^ 16u));
      const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      tmem_st8(lane_base + KA_COL + (it >> 1) * 16 + kh * 8, w);
    }
    if (g == 0) {
      const uint32_t w[8] = {0x3f803f80u, 0x00003f80u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(lane_base + ONES_COL, w);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(ka_full);

    const uint32_t tS = lane_base + S_COL + (SHARED ? g * CW : g * 2 * CW), tDP = SHARED ? tS + QT : tS + CW;
    const int gs = SHARED ? 0 : g;  // whose score barriers this warpgroup uses
    const uint64_t c2 = pk2(scale_log2, scale_log2);
    // dS^T tile: 64-query chunk (g * CW) / 64, 16-byte unit ((g * CW) % 64) / 8 + ..., XOR-swizzled with the row
    const uint32_t ds_row = ((g * CW) >> 6) * (DS_BYTES / 2) + r * 128;
    const uint32_t u0 = ((g * CW) & 63) >> 3;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const int dbg = stagger >> 8;  // 8 = no dS^T shared-memory stores, 16 = no exponentials (timing experiments), 32 = late s_free in PIPE mode
    for (int t = 0; t < NT; ++t) {
      const int buf = t & 1;
      mbar_wait(s_full(gs), t & 1);
      tc_fence_after();
      constexpr int NCH = CW / 16;
      uint32_t svp[PIPE ? NCH : 1][16], dvp[PIPE ? NCH : 1][16];
      if (PIPE) {
        // the whole sub-tile of this warpgroup leaves TMEM now: the next score products can be issued before the
        // first exponential of this one (they queue behind up to 24 product UMMAs in the in-order tensor pipe)
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          tmem_ld16(tS + cc * 16, svp[cc]);
          tmem_ld16(tDP + cc * 16, dvp[cc]);
        }
        tmem_ld_wait();
        if (!(dbg & 32)) {
          tc_fence_before();
          mbar_arrive(s_free(gs));
        }
      }
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t svl[16], dvl[16];
        if (!PIPE) {
          tmem_ld16(tS + cc * 16, svl);
          tmem_ld16(tDP + cc * 16, dvl);
          tmem_ld_wait();
          if (cc == NCH - 1) {
            tc_fence_before();
            mbar_arrive(s_free(gs));
          }
        }
        const uint32_t* sv = PIPE ? svp[cc] : svl;
        const uint32_t* dv = PIPE ? dvp[cc] : dvl;
        uint32_t pP[8], pD[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // x = c (s - lse2 / c) <= 0 up to rounding: P <= 1
          float a0, a1, p0, p1;
          const uint64_t a01 = fmul2_(pk2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2);
          if (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
            exp2_poly2_(a01, p0, p1);
          } else {
            upk2(a01, a0, a1);
            p0 = (dbg & 16) ? a0 : ex2f(a0);
            p1 = (dbg & 16) ? a1 : ex2f(a1);
          }
          float e0, e1;
          upk2(fmul2_(pk2(p0, p1), pk2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1]))), e0, e1);
          pP[i] = pack_bf16(p0, p1);
          pD[i] = pack_bf16(e0, e1);
        }
        if (PIPE && (dbg & 32) && cc == 0) {  // A/B: whole-sub-tile load, but the score buffer goes back where the default does
          tc_fence_before();
          mbar_arrive(s_free(gs));
        }
        if (cc == 0 && t >= 2) {  // P^T / dS^T buffers of sub-tile t - 2 have been consumed
          mbar_wait(mma_done(buf), ((t >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        tmem_st8(lane_base + P_COL + buf * 64 + (g * CW + cc * 16) / 2, pP);
        const uint32_t drow = sDS(buf) + ds_row;
        if (!(dbg & 8))
#pragma unroll
        for (int i4 = 0; i4 < 2; ++i4)
          sts128(drow + (((u0 + cc * 2 + i4) ^ sw) << 4), pD[4 * i4], pD[4 * i4 + 1], pD[4 * i4 + 2], pD[4 * i4 + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full(buf));
    }
    // ---- epilogue: dV / dK of both key halves, one (matrix, half) per warpgroup round
    mbar_wait(mma_done(1), ((NT - 1) >> 1) & 1);
    tc_fence_after();
    for (int it = g; it < 4; it += NWG) {
      const int kh_e = it & 1;
      const bool is_dk = it >= 2;
      uint32_t a[16];
      tmem_ld16(lane_base + (is_dk ? DK_COL : DV_COL) + kh_e * 16, a);
      tmem_ld_wait();
      const float mul = is_dk ? scale : 1.f;
      const size_t grow = (size_t)row_base + kb0 + kh_e * KT + r;
      uint4* dst = reinterpret_cast<uint4*>(dqkv + grow * 3 * C + (is_dk ? C : 2 * C) + h * DH);
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = pack_bf16(__uint_as_float(a[2 * i]) * mul, __uint_as_float(a[2 * i + 1]) * mul);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
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
      uint4* row = reinterpret_cast<uint4*>(st + ST_LSET + r * ROWB);
      row[x] = make_uint4(w0, w1, 0u, 0u);
      row[x ^ 1] = make_uint4(w0, w1, 0u, 0u);
      split3(-dl, w0, w1);
      row = reinterpret_cast<uint4*>(st + ST_DELT + r * ROWB);
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
      uint32_t o[16];
      tmem_ld16(lane_base + DQ_COL, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      float* dst = dq_head + ((size_t)j * QT + r) * DH;
#pragma unroll
      for (int i = 0; i < 4; ++i)
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

bool attn_bwd_tc_supported(int L, int C, int heads) {
  return C % heads == 0 && C / heads == DH && L % (KH * KT) == 0 && L >= KH * KT;
}

int launch_attn_bwd_tc(cudaStream_t st, const void* qkv, const void* dout, const float* lse2, const float* delta,
                       void* dqkv, float* ws, int B, int L, int C, int heads) {
  TSD_CHECK(attn_bwd_tc_supported(L, C, heads), "attn_bwd_tc: unsupported shape L=%d C=%d heads=%d", L, C, heads);
  CUtensorMap tmQKV, tmDO;
  if (make_tmap_2d_sw(&tmQKV, qkv, 2, (uint64_t)B * L, 3 * (uint64_t)C, 3 * (uint64_t)C, DH, 64, ROWB)) return 1;
  if (make_tmap_2d_sw(&tmDO, dout, 2, (uint64_t)B * L, (uint64_t)C, (uint64_t)C, DH, 64, ROWB)) return 1;
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    configured.cur() = true;
  }
  static int stagger = -1, shared = -1, pipe = -1, poly = 0;
  if (stagger < 0) {
    // every n-th pair of exponentials on the FMA pipe (n = 2, 4 or 8; 0 = all on MUFU.EX2, the default).  The kernel alone,
    // 64 x L = 4096, one box, alternating: 0: 3.297 ms, 8: 3.235, 4: 3.205 / 3.207, 2: 3.399 (gradients identical) -- but
    // the whole training step at batch 256 (power-capped, tools/ncu_step.py, alternating 4 / 0 / 4 / 0): 128.06 / 126.16 /
    // 127.20 / 125.59 ms: the extra FMA-pipe work costs the step more than the shorter MUFU queue gains.  Kept for A/B.
    if (const char* ep = getenv("TSD_ATTN_BWD_TC_POLY")) poly = atoi(ep);
    const char* e = getenv("TSD_ATTN_BWD_TC_STAGGER");
    stagger = e ? atoi(e) : 0;
    e = getenv("TSD_ATTN_BWD_TC_DBG");
    if (e) stagger |= atoi(e) << 8;
    // 1: score products issued once per sub-tile for all warpgroups (N = 128).  Measured SLOWER (3.54 vs 3.05 ms per 64
    // samples at L = 4096): sharing the hand-off locks the four warpgroups into the same phase.  Kept for A/B runs.
    e = getenv("TSD_ATTN_BWD_TC_SHARED");
    shared = e ? atoi(e) : 0;
    // 1: setmaxnreg register hand-over + whole-sub-tile score loads (see BT_THREADS_PIPE).  Measured SLOWER as well
    // (3.27 vs 3.09 ms): handing the score buffer back earlier puts the next score products ahead of the pending
    // dV / dK / dQ products in the in-order tensor pipe and lengthens their critical path.  Kept for A/B runs.
    e = getenv("TSD_ATTN_BWD_TC_PIPE");
    pipe = e ? atoi(e) : 0;
  }
  const float scale = 1.f / sqrtf((float)DH);
  const dim3 grid(L / (KH * KT), heads, B);
  const float sl2 = 1.4426950408889634f * scale;
  if (shared)
    attn_bwd_tc_kernel<true, false><<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  else if (pipe)
    attn_bwd_tc_kernel<false, true><<<grid, BT_THREADS_PIPE, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  else if (poly == 2)
    attn_bwd_tc_kernel<false, false, 2><<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  else if (poly == 4)
    attn_bwd_tc_kernel<false, false, 4><<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  else if (poly == 8)
    attn_bwd_tc_kernel<false, false, 8><<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  else
    attn_bwd_tc_kernel<false, false><<<grid, BT_THREADS, BT_SMEM, st>>>(tmQKV, tmDO, lse2, delta, (bf16*)dqkv, ws, L, C, scale, sl2, stagger);
  TSD_LAUNCH_CHECK();
  return 0;
}

}  // namespace tsd
