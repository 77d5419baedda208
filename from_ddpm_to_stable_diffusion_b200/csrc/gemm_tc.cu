// tcgen05 / TMEM / TMA GEMM core (see gemm_tc.cuh for the operand modes).
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace tsd {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
// Epilogue warps: EPI_Q per TMEM sub-partition, each draining 128 / EPI_Q accumulator columns.  Measured (same box,
// alternating builds): 16 epilogue warps of 96 registers (-DTSD_EPI_Q=4) against 8 of 168 -- training step 136.4 vs
// 135.8 ms, reverse step 77.7 vs 77.2 ms: the epilogue-heavy K = 128 layers are not short of warps, and the HBM-bound
// ones lose a little.  8 it stays.
#ifndef TSD_EPI_Q
#define TSD_EPI_Q 2
#endif
constexpr int EPI_Q = TSD_EPI_Q;
constexpr int EPI_THREADS = 4 * EPI_Q * 32;
constexpr int NUM_THREADS = 128 + EPI_THREADS;  // 4 control warps + the epilogue warps
constexpr int EPI_CH = 4 / EPI_Q;               // 32-column chunks per epilogue warp (plain / fp32 paths)
constexpr int EPI_PAIRS = 64 / EPI_Q;           // (value, gate) pairs per thread (GEGLU paths)
constexpr int GN_ROWS = 128 * 128 / EPI_THREADS;  // rows per thread of the column-wise by-products (32)
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 32) tmem_ld32(taddr, r);
  else tmem_ld16(taddr, r);
}
// packed fp32 pairs (one FMUL2 / FFMA2 / FADD2 issue slot for two values): the GEGLU epilogues are bound by instruction
// issue -- 8 epilogue warps, ~20 instructions per output element pair in scalar form
__device__ __forceinline__ uint64_t pk2f(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2f(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2f(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2f(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2f(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
constexpr int TMEM_COLS = 256;     // two 128-column fp32 accumulators
constexpr int TMEM_COLS_WG = 512;  // weight-gradient instantiation: room for the three tap accumulators of the patch mode
constexpr int MAX_RING = 8, MAX_ARING = 4;  // mbarrier pairs reserved for the operand rings

// ---- halo mode of the 3x3 convolution (A_KHALO): the M tile is a 16-row x 8-pixel patch of ONE image, and the
// activations of a 64-channel block are loaded once per patch instead of once per tap.  An 8-pixel image row of 64
// channels is exactly one SWIZZLE_128B atom (8 x 128 B), so a K-major UMMA descriptor with SBO = one halo row walks down
// the 16 patch rows, and a tap's vertical shift is a whole number of atoms:
//   halo == 1: three x-shifted copies {64 ch, 8 px, 18 rows} (18 KB each, one ring slot per copy); tap (dy, dx) reads
//              copy dx from row dy on -- every descriptor start stays 1024-byte aligned.  A traffic: 54 KB per patch and
//              channel block instead of 9 x 16 KB.
//   halo == 2: one copy {64 ch, 10 px, 18 rows} (22.5 KB); tap (dy, dx) starts dy rows + dx pixels in, i.e. NOT on an
//              atom boundary -- correct as it stands, because TMA and UMMA both derive the swizzle phase from the
//              absolute shared-memory address.  A traffic: 22.5 KB instead of 9 x 16 KB.
//   halo == 3: the tile is TWO such patches side by side (16 rows x 16 pixels, one copy {64 ch, 18 px, 18 rows}, 40.5 KB)
//              with one accumulator each: every weight tile now feeds eight UMMAs instead of four, which halves the
//              weight stream and the per-k-block issue overhead of the MMA warp (measured: that overhead, not the
//              operand streams, bounded halo == 2).  The default where the width is a multiple of 16.
// The weight tiles (16 KB per tap and channel block) keep streaming through their own ring.
constexpr int HALO_TW = 8, HALO_TH = 16, HALO_ROWS = HALO_TH + 2;
// ---- patch mode of the 3x3 weight gradient (p.wg_halo, instantiation <1,1,1>): a k-block is an 8 x 8 pixel patch of one
// image.  A work item owns one kernel ROW dy of a (128 co x 128 ci) block and keeps the three taps dx = 0, 1, 2 in three
// TMEM accumulators: dY of the patch (16 KB) and the input rows y + dy - 1 with a one-pixel margin left and right
// ({64 ci, 10 px, 8 rows} per 64-channel chunk, 20 KB) are loaded once and serve all three taps -- the tap is a
// 128-byte shift of the MN-major B descriptor.  Operand traffic per tap and 64 pixels: 12 KB instead of 32 KB.
constexpr int WG_A_BYTES = 64 * 128 * 2;             // dY patch: two 64-channel chunks of [64 px][128 B]
constexpr int WG_B_CHUNK = 8 * (HALO_TW + 2) * 128;  // one 64-channel chunk of the input rows: 10 KB
constexpr int WG_STAGE = WG_A_BYTES + 2 * WG_B_CHUNK;  // 36 KB
constexpr int WG_STAGES = 4;
struct HaloCfg {
  int taps_per_item;     // taps served by one ring slot of activations
  int n_a, n_b;          // ring depths
  uint32_t a_bytes;      // bytes one activation slot receives
  uint32_t a_stage;      // slot pitch (1024-byte aligned)
  uint32_t row_pitch;    // bytes between halo rows inside a slot (= SBO)
};
__device__ __forceinline__ HaloCfg halo_cfg(int mode) {
  HaloCfg c;
  if (mode == 1) {
    c.taps_per_item = 3; c.n_a = 4; c.n_b = 5;
    c.a_bytes = HALO_ROWS * HALO_TW * 128; c.a_stage = c.a_bytes; c.row_pitch = HALO_TW * 128;
  } else if (mode == 2) {
    c.taps_per_item = 9; c.n_a = 2; c.n_b = 7;
    c.a_bytes = HALO_ROWS * (HALO_TW + 2) * 128; c.a_stage = 23 * 1024; c.row_pitch = (HALO_TW + 2) * 128;
  } else {  // 3: two patches side by side (16 rows x 16 pixels, M = 256 in two accumulators)
    c.taps_per_item = 9; c.n_a = 2; c.n_b = 4;
    c.a_bytes = HALO_ROWS * (2 * HALO_TW + 2) * 128; c.a_stage = 41 * 1024; c.row_pitch = (2 * HALO_TW + 2) * 128;
  }
  return c;
}

template <int OUT_F32>
struct Cfg {
  static constexpr int STAGES = 5;
  // fp32 out: one 64 KB staging tile; bf16 out: two 32 KB staging tiles (double buffered; a tile's residual is
  // TMA-prefetched into the buffer its result will be stored from)
  static constexpr int STAGING_BYTES = 64 * 1024;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
};

template <int A_MN, int B_MN, int OUT_F32, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
               const GemmParams p) {
  constexpr int STAGES = Cfg<OUT_F32>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage0 = smem_base;
  const uint32_t smem_staging = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = smem_staging + Cfg<OUT_F32>::STAGING_BYTES;
  // barrier map (8 bytes each): full[8], empty[8], tmem_full[2], tmem_empty[2], res[2], halo a_full[4], a_empty[4],
  // then the tmem ptr.  The plain modes use STAGES of the full / empty pairs; the halo mode (below) uses up to 7 of
  // them for its weight ring and the a_* pairs for its activation ring.
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_RING + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_RING + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_RING + 2 + s); };
  auto res_bar = [&](int s) { return bar_base + 8u * (2 * MAX_RING + 4 + s); };
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_RING + 6 + s); };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_RING + 6 + MAX_ARING + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_RING + 6 + 2 * MAX_ARING);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to aligned base
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // warp index made warp-uniform for the compiler: the MMA warp runs converged and elects a lane only around the
  // tcgen05 instructions, so the four UMMAs of a k-block issue back to back from uniform registers (a lane-0 branch
  // costs an elect loop and vector-to-uniform moves per instruction: ~75 clk per UMMA, more than the UMMA itself)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB0);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_RING; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < MAX_ARING; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_THREADS);
      mbar_init(res_bar(s), 1);
    }
    fence_mbar_init();
  }
  constexpr int kTmemCols = TMEM_COLS_WG;  // four accumulators (halo == 3: two pairs; patch-mode weight gradient: three)
  constexpr bool kWG = A_MN && B_MN && OUT_F32;  // the instantiation that carries the patch-mode 3x3 weight gradient
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_mn * p.splits;

  // Patch geometry of an M tile in halo mode: m_blk = (image, patch row, patch column).
  auto halo_geom = [&](int m_blk, int& img, int& y0, int& x0) {
    img = fdiv(m_blk, p.fd_tpi);
    const int r = m_blk - img * p.halo_tpi;
    const int ty = fdiv(r, p.fd_tx);
    y0 = ty * HALO_TH;
    x0 = (r - ty * p.halo_tx) * HALO_TW;
  };
  // One weight (B) k-block of the tile column starting at n0 into ring slot sb.
  auto load_b = [&](int kb, uint32_t sb, uint32_t fb, int n0, int b_c, int b_tap, const CUtensorMap* b_map) {
    if (B_MN == 0) {
      tma_load_2d(sb, &tmB0, fb, kb * BK, n0);
    } else if (p.b_mode == B_MN2D) {
      const int tap = kb / p.b_cpt;
      const int row = (kb - tap * p.b_cpt) * BK;
      const int t = p.b_flip ? (p.b_ntaps - 1 - tap) : tap;
      const int col = b_c + t * p.b_tapstride;
      tma_load_2d(sb, b_map, fb, col, row);
      tma_load_2d(sb + B_STAGE_BYTES / 2, b_map, fb, col + 64, row);
    } else {  // B_MNCONV: 64 output pixels starting at kb*64, shifted by the tap
      const int p0 = kb * BK;
      const int hw = p.Ho * p.Wo;
      const int n_i = p0 / hw;
      const int rem = p0 - n_i * hw;
      const int y = rem / p.Wo, x = rem - (rem / p.Wo) * p.Wo;
      const int dy = b_tap / 3, dx = b_tap - dy * 3;
      const int xs = x * p.stride + dx - 1, ys = y * p.stride + dy - 1;
      tma_load_4d(sb, b_map, fb, b_c, xs, ys, n_i);
      tma_load_4d(sb + B_STAGE_BYTES / 2, b_map, fb, b_c + 64, xs, ys, n_i);
    }
  };

  if (p.halo && A_MN == 0 && warp == 3 && lane == 0) {
    // =========================================================== halo mode: activation producer
    const HaloCfg hc = halo_cfg(p.halo);
    const int items = p.a_cpt * (9 / hc.taps_per_item);
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = fdiv(tile - fdiv(tile, p.fd_tiles_mn) * tiles_mn, p.fd_tiles_n) * (p.halo == 3 ? 2 : 1);  // (first) 16 x 8 patch
      int img, y0, x0;
      halo_geom(m_blk, img, y0, x0);
      for (int item = 0; item < items; ++item) {
        const int cb = hc.taps_per_item == 3 ? item / 3 : item;
        const int dx = hc.taps_per_item == 3 ? item - cb * 3 : 0;
        int c = cb * BK;
        const CUtensorMap* m = &tmA0;
        if (c >= p.a_c0) { c -= p.a_c0; m = &tmA1; }
        mbar_wait(aempty_bar(as), aph ^ 1);
        if ((p.dbg == 2 || p.dbg == 5) && (item >= hc.n_a || tile != (int)blockIdx.x)) {  // timing experiment: no A loads
          mbar_arrive(afull_bar(as));
        } else {
          mbar_arrive_expect_tx(afull_bar(as), hc.a_bytes);
          tma_load_4d(smem_stage0 + as * hc.a_stage, m, afull_bar(as), c, x0 + dx - 1, y0 - 1, img);
        }
        if (++as == hc.n_a) { as = 0; aph ^= 1; }
      }
    }
  } else if (p.halo && A_MN == 0 && warp == 0 && lane == 0) {
    // =========================================================== halo mode: weight producer
    const HaloCfg hc = halo_cfg(p.halo);
    const int items = p.a_cpt * (9 / hc.taps_per_item);
    const uint32_t smem_b0 = smem_stage0 + hc.n_a * hc.a_stage;
    int bs = 0;
    uint32_t bph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int t2b = tile - fdiv(tile, p.fd_tiles_mn) * tiles_mn;
      const int n0 = (t2b - fdiv(t2b, p.fd_tiles_n) * p.tiles_n) * BN;
      for (int item = 0; item < items; ++item) {
        const int cb = hc.taps_per_item == 3 ? item / 3 : item;
        const int dx = hc.taps_per_item == 3 ? item - cb * 3 : 0;
        for (int j = 0; j < hc.taps_per_item; ++j) {
          const int tap = hc.taps_per_item == 3 ? j * 3 + dx : j;
          mbar_wait(empty_bar(bs), bph ^ 1);
          if ((p.dbg == 1 || p.dbg == 5) && (item * hc.taps_per_item + j >= hc.n_b || tile != (int)blockIdx.x)) {  // no B loads
            mbar_arrive(full_bar(bs));
          } else {
            mbar_arrive_expect_tx(full_bar(bs), B_STAGE_BYTES);
            load_b(tap * p.a_cpt + cb, smem_b0 + bs * B_STAGE_BYTES, full_bar(bs), n0, n0, 0, &tmB0);
          }
          if (++bs == hc.n_b) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (kWG && p.wg_halo && warp == 0 && lane == 0) {
    // =========================================================== patch-mode weight gradient: producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int t2 = tile - split * tiles_mn;
      const int n_blk = t2 % p.tiles_n;           // (kernel row dy, 128-channel block of the input)
      const int m0 = (t2 / p.tiles_n) * BM;
      const int ci_blocks = p.tiles_n / 3;
      const int dy = n_blk / ci_blocks;
      int b_c = (n_blk - dy * ci_blocks) * BN;
      const CUtensorMap* b_map = &tmB0;
      if (b_c >= p.b_c0) { b_c -= p.b_c0; b_map = &tmB1; }
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        int img, y0, x0;
        halo_geom(kb, img, y0, x0);
        y0 >>= 1;  // halo_geom counts 16-row patches; these are 8 rows tall
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_stage0 + stage * WG_STAGE;
        const uint32_t sb = sa + WG_A_BYTES;
        const uint32_t fb = full_bar(stage);
        mbar_arrive_expect_tx(fb, WG_STAGE);
        tma_load_4d(sa, &tmA0, fb, m0, x0, y0, img);
        tma_load_4d(sa + WG_A_BYTES / 2, &tmA0, fb, m0 + 64, x0, y0, img);
        tma_load_4d(sb, b_map, fb, b_c, x0 - 1, y0 + dy - 1, img);
        tma_load_4d(sb + WG_B_CHUNK, b_map, fb, b_c + 64, x0 - 1, y0 + dy - 1, img);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0 && lane == 0 && !(p.halo && A_MN == 0)) {
    // =========================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      // (m, n) tiles fastest, K split slowest: CTAs running together share operand tiles in L2.
      const int split = fdiv(tile, p.fd_tiles_mn);
      const int t2 = tile - split * tiles_mn;
      const int m_blk = fdiv(t2, p.fd_tiles_n);
      const int n_blk = t2 - m_blk * p.tiles_n;
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);

      // Conv geometry of this M tile (A_KCONV): the 128 rows are a (bn, bh, bw) box of output pixels.
      int a_n0 = 0, a_y0 = 0, a_x0 = 0;
      if (p.a_mode == A_KCONV) {
        const int hw = p.Ho * p.Wo;
        a_n0 = m0 / hw;
        const int rem = m0 - a_n0 * hw;
        a_y0 = rem / p.Wo;
        a_x0 = rem - a_y0 * p.Wo;
      }
      // B_MNCONV: the N tile is (tap, 128 channels of one source).
      int b_tap = 0, b_c = n0;
      const CUtensorMap* b_map = &tmB0;
      if (p.b_mode == B_MNCONV) {
        b_tap = n0 / p.b_ctot;
        b_c = n0 - b_tap * p.b_ctot;
      }
      if (p.b_mode != B_K2D && b_c >= p.b_c0) {
        b_c -= p.b_c0;
        b_map = &tmB1;
      }

      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_stage0 + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint32_t fb = full_bar(stage);
        // p.dbg (TSD_GEMM_DBG, timing experiments only, WRONG results): 1 = no B loads after the first ring fill,
        // 2 = no A loads after the first ring fill -- which operand stream bounds the main loop?
        const bool skip_b = p.dbg == 1 && (kb - kb_begin >= STAGES || tile != (int)blockIdx.x);
        const bool skip_a = p.dbg == 2 && (kb - kb_begin >= STAGES || tile != (int)blockIdx.x);
        mbar_arrive_expect_tx(fb, (skip_a ? 0 : A_STAGE_BYTES) + (skip_b ? 0 : B_STAGE_BYTES));
        // ---- A
        if (skip_a) {
        } else if (A_MN == 0) {
          if (p.a_mode == A_K2D) {
            int c = kb * BK;
            const CUtensorMap* m = &tmA0;
            if (c >= p.a_c0) { c -= p.a_c0; m = &tmA1; }
            tma_load_2d(sa, m, fb, c, m0);
          } else {  // A_KCONV
            const int tap = kb / p.a_cpt;
            int c = (kb - tap * p.a_cpt) * BK;
            const CUtensorMap* m = &tmA0;
            if (c >= p.a_c0) { c -= p.a_c0; m = &tmA1; }
            const int dy = tap / 3, dx = tap - dy * 3;
            tma_load_4d(sa, m, fb, c, a_x0 * p.stride + dx - 1, a_y0 * p.stride + dy - 1, a_n0);
          }
        } else {  // A_MN2D: [K][M], two 64-wide M chunks
          tma_load_2d(sa, &tmA0, fb, m0, kb * BK);
          tma_load_2d(sa + A_STAGE_BYTES / 2, &tmA0, fb, m0 + 64, kb * BK);
        }
        // ---- B
        if (!skip_b) load_b(kb, sb, fb, n0, b_c, b_tap, b_map);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    // K-major: 32 B per UMMA_K step inside the 128 B swizzle row; MN-major: 16 k-rows of 128 B.
    constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32;
    constexpr uint32_t B_KSTEP = B_MN ? 2048 : 32;
    constexpr uint32_t A_LBO = A_MN ? (A_STAGE_BYTES / 2) : 0;
    constexpr uint32_t B_LBO = B_MN ? (B_STAGE_BYTES / 2) : 0;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (kWG && p.wg_halo) {
      // patch-mode weight gradient: three taps (dx) of one kernel row accumulate side by side in TMEM
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int split = tile / tiles_mn;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
        mbar_wait(tempty_bar(0), (it & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_stage0 + stage * WG_STAGE;
          const uint32_t sb = sa + WG_A_BYTES;
          if (elect_one()) {
            const uint64_t da = umma_smem_desc(sa, WG_A_BYTES / 2, 1024);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              // B rows are pixels of the margin-extended input rows: 8-pixel groups one image row (10 pixels) apart,
              // the tap's horizontal shift is dx pixels into the row (not on an atom boundary: see the halo mode)
              const uint64_t db = umma_smem_desc(sb + dx * 128, WG_B_CHUNK, (HALO_TW + 2) * 128);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)  // 16 pixels = two image rows of the patch per step
                umma_bf16(tmem_base + dx * BN, da + (uint64_t)(k * (2048 / 16)),
                          db + (uint64_t)(k * (2 * (HALO_TW + 2) * 128 / 16)), idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            if (kb + 1 == kb_end) umma_commit(tfull_bar(0));
          }
          __syncwarp();
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb_end <= kb_begin) {
          if (elect_one()) umma_commit(tfull_bar(0));
          __syncwarp();
        }
      }
    } else if (p.halo && A_MN == 0) {
      // halo mode: the taps of an activation slot are shifted descriptors over the same shared-memory patch
      const HaloCfg hc = halo_cfg(p.halo);
      const int items = p.a_cpt * (9 / hc.taps_per_item);
      const uint32_t smem_b0 = smem_stage0 + hc.n_a * hc.a_stage;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const int subs = p.halo == 3 ? 2 : 1;
        const uint32_t d_tmem = tmem_base + acc * subs * BN;
        for (int item = 0; item < items; ++item) {
          mbar_wait(afull_bar(as), aph);
          const uint32_t sa = smem_stage0 + as * hc.a_stage;
          for (int j = 0; j < hc.taps_per_item; ++j) {
            if (p.dbg != 4) mbar_wait(full_bar(stage), phase);  // (4: timing experiment, issue without waiting)
            tc_fence_after();
            const uint32_t sb = smem_b0 + stage * B_STAGE_BYTES;
            const uint32_t a0 = hc.taps_per_item == 3 ? sa + j * hc.row_pitch
                                                      : sa + (j / 3) * hc.row_pitch + (j % 3) * 128;
            if (elect_one()) {
              const uint64_t da = umma_smem_desc(a0, A_LBO, hc.row_pitch);
              // (halo >= 2: a0 is not on a 1024-byte atom boundary.  The swizzle is a function of the absolute
              // shared-memory address, on the TMA side and on the UMMA side alike, so the descriptor's base-offset
              // field stays 0 -- measured: with the field set to (a0 >> 7) & 7 the results are wrong.)
              const uint64_t db = umma_smem_desc(sb, B_LBO, 1024);
              for (int h = 0; h < subs; ++h) {  // halo == 3: the second patch starts 8 pixels to the right
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                  umma_bf16(d_tmem + h * BN, da + (uint64_t)(h * (HALO_TW * 128 / 16) + k * (A_KSTEP / 16)),
                            db + (uint64_t)(k * (B_KSTEP / 16)), idesc, (item > 0 || j > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(empty_bar(stage));
              if (j + 1 == hc.taps_per_item) umma_commit(aempty_bar(as));
              if (j + 1 == hc.taps_per_item && item + 1 == items) umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++stage == hc.n_b) { stage = 0; phase ^= 1; }
          }
          if (++as == hc.n_a) { as = 0; aph ^= 1; }
        }
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int split = tile / tiles_mn;
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_stage0 + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        if (elect_one()) {
          const uint64_t da = umma_smem_desc(sa, A_LBO, 1024);
          const uint64_t db = umma_smem_desc(sb, B_LBO, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)  // the start-address field holds addr >> 4
            umma_bf16(d_tmem, da + (uint64_t)(k * (A_KSTEP / 16)), db + (uint64_t)(k * (B_KSTEP / 16)), idesc,
                      (kb > kb_begin || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
          if (kb + 1 == kb_end) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb_end <= kb_begin) {  // empty K range (never produced by the host-side split): keep the hand-off alive
        if (elect_one()) umma_commit(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // =========================================================== epilogue
    // Epilogue warps w, w+4, ... share TMEM sub-partition (w % 4) and split the 128 columns into EPI_Q groups.
    const int ep_warp = (warp - 4) & 3;      // TMEM sub-partition
    const int half = (warp - 4) >> 2;        // column group handled by this warp (0 .. EPI_Q-1)
    const int ep_tid = ep_warp * 32 + lane;  // == TMEM lane == row inside the tile
    const bool ep_leader = (warp == 4 && lane == 0);
    const uint32_t lane_off = static_cast<uint32_t>(ep_warp * 32) << 16;
    uint8_t* staging0 = smem_gen + (smem_staging - smem_base);
    const bool has_res = (!OUT_F32) && p.residual != nullptr;
    // residual tile of (m0, n0): two 64-column boxes, prefetched by TMA into a staging buffer
    const int subs = (!OUT_F32 && p.halo == 3) ? 2 : 1;  // 128-row sub-tiles (accumulators) per tile
    auto prefetch_residual = [&](int tile_idx, int hsub, int bufsel) {
      const int t2r = tile_idx - fdiv(tile_idx, p.fd_tiles_mn) * tiles_mn;
      const int mq = fdiv(t2r, p.fd_tiles_n);
      const int nb = t2r - mq * p.tiles_n, mb = mq * subs + hsub;
      const uint32_t dst = smem_staging + bufsel * (BM * BN * 2);
      if (p.halo) {  // the tile is a 16 x 8 pixel patch: same bytes in shared memory, 4-D box in global memory
        int img, y0, x0;
        halo_geom(mb, img, y0, x0);
        mbar_arrive_expect_tx(res_bar(bufsel), BM * BN * 2);
        tma_load_4d(dst, &tmR, res_bar(bufsel), nb * BN, x0, y0, img);
        tma_load_4d(dst + BM * 128, &tmR, res_bar(bufsel), nb * BN + 64, x0, y0, img);
        return;
      }
      if (EPI == EPI_GEGLU_BWD) {  // the 64 d(gg) columns that belong to this tile's 64 value + 64 gate columns
        mbar_arrive_expect_tx(res_bar(bufsel), BM * 64 * 2);
        tma_load_2d(dst, &tmR, res_bar(bufsel), nb * 64, mb * BM);
        return;
      }
      mbar_arrive_expect_tx(res_bar(bufsel), BM * BN * 2);
      tma_load_2d(dst, &tmR, res_bar(bufsel), nb * BN, mb * BM);
      tma_load_2d(dst + BM * 128, &tmR, res_bar(bufsel), nb * BN + 64, mb * BM);
    };
    float cs_acc[16];  // EPI_GEGLU_BWD: this thread's column-sum partials per n-block (bias gradient), flushed at the end
#pragma unroll
    for (int i = 0; i < 16; ++i) cs_acc[i] = 0.f;
    if (has_res && ep_leader && (int)blockIdx.x < total_tiles) prefetch_residual(blockIdx.x, 0, 0);
    int it = 0;
    if (kWG && p.wg_halo) {
      // patch-mode weight gradient: the item's three tap accumulators leave one after the other through the single
      // fp32 staging tile (TMA reduce-add into dW[co][(dy*3 + dx) * cin + ci])
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int t2 = tile % tiles_mn;
        const int n_blk = t2 % p.tiles_n;
        const int m0 = (t2 / p.tiles_n) * BM;
        const int ci_blocks = p.tiles_n / 3;
        const int dy = n_blk / ci_blocks;
        const int ci0 = (n_blk - dy * ci_blocks) * BN;
        mbar_wait(tfull_bar(0), it & 1);
        tc_fence_after();
        const int r7 = ep_tid & 7;
#pragma unroll 1
        for (int dx = 0; dx < 3; ++dx) {
          if (ep_leader) tma_store_wait_read<0>();  // the previous reduce-add has read the staging tile
          named_bar_sync(1, EPI_THREADS);
          const uint32_t t_addr = tmem_base + lane_off + dx * BN;
#pragma unroll 1
          for (int ch = half * EPI_CH; ch < half * EPI_CH + EPI_CH; ++ch) {
            uint32_t v[32];
            tmem_ld32(t_addr + ch * 32, v);
            tmem_ld_wait();
            uint8_t* dst = staging0 + ch * (BM * 128) + ep_tid * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint4 o = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              *reinterpret_cast<uint4*>(dst + ((q ^ r7) << 4)) = o;
            }
          }
          if (dx == 2) {
            tc_fence_before();
            mbar_arrive(tempty_bar(0));
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          if (ep_leader) {
            const int col0 = (dy * 3 + dx) * p.b_ctot + ci0;
#pragma unroll
            for (int ch = 0; ch < BN / 32; ++ch)
              tma_reduce_add_2d(&tmD, smem_staging + ch * (BM * 128), col0 + ch * 32, m0);
            tma_store_commit();
          }
        }
      }
    } else
    for (int tile = blockIdx.x, sit = 0; tile < total_tiles; tile += gridDim.x, ++it)
    for (int hsub = 0; hsub < subs; ++hsub, ++sit) {  // sit counts sub-tiles: staging buffers / residual barriers alternate
      const int sbuf = OUT_F32 ? 0 : (sit & 1);
      uint8_t* staging = staging0 + sbuf * (BM * BN * 2);
      const uint32_t staging_s = smem_staging + sbuf * (BM * BN * 2);
      // (no integer divisions per tile: with K = 128 the epilogue is the critical path, and the division sequences of
      // all 256 epilogue threads were a quarter of its instructions)
      const int t2 = tile - fdiv(tile, p.fd_tiles_mn) * tiles_mn;
      const int mq = fdiv(t2, p.fd_tiles_n);
      const int n_blk = t2 - mq * p.tiles_n;
      const int m_blk = mq * subs + hsub;  // index of the 128-row sub-tile
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int h_img = 0, h_y0 = 0, h_x0 = 0;
      if (p.halo) halo_geom(m_blk, h_img, h_y0, h_x0);
      // halo mode: all 128 rows of the patch belong to image h_img
      const int row = p.halo ? h_img * (p.Ho * p.Wo) : m0 + ep_tid;
      const bool row_ok = row < p.M;
      const int sample = (p.row_bias && row_ok) ? fdiv(row, p.fd_rps) : 0;

      if (hsub == 0) mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (p.dbg == 3) {  // timing experiment: no epilogue work at all
        tc_fence_before();
        if (hsub == subs - 1) mbar_arrive(tempty_bar(acc));
        continue;
      }
      if (OUT_F32) {
        // single staging tile: it must have been fully read by the previous tile's TMA reduce-add
        if (ep_leader) tma_store_wait_read<0>();
        named_bar_sync(1, EPI_THREADS);
      } else if (has_res) {
        mbar_wait(res_bar(sbuf), (sit >> 1) & 1);  // residual tile has landed in this staging buffer
      }

      const uint32_t t_addr = tmem_base + lane_off + (acc * subs + hsub) * BN;
      const int r7 = ep_tid & 7;
      if (OUT_F32) {
#pragma unroll 1
        for (int ch = half * EPI_CH; ch < half * EPI_CH + EPI_CH; ++ch) {
          uint32_t v[32];
          tmem_ld32(t_addr + ch * 32, v);
          tmem_ld_wait();
          uint8_t* dst = staging + ch * (BM * 128) + ep_tid * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint4 o = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            *reinterpret_cast<uint4*>(dst + ((q ^ r7) << 4)) = o;
          }
        }
      } else if (EPI == EPI_GEGLU_BWD) {
        // Backward of value * gelu(gate) with the pre-activations RECOMPUTED by this GEMM (accumulator columns [0,64) =
        // value x, [64,128) = gate g, same packing as the forward) and d(gg) of the tile TMA-loaded into the staging
        // buffer: dx = d * gelu(g) goes back over d, dg = d * x * gelu'(g) into the second half; gelu in the tanh form of
        // the forward epilogue (geglu_fast_f), one MUFU per element for value and derivative together.
        const int ch = half;
        uint32_t xv[EPI_PAIRS], gv[EPI_PAIRS];
        tmem_ld_n<EPI_PAIRS>(t_addr + ch * EPI_PAIRS, xv);
        tmem_ld_n<EPI_PAIRS>(t_addr + 64 + ch * EPI_PAIRS, gv);
        tmem_ld_wait();
        uint8_t* row_d = staging + ep_tid * 128;
        uint8_t* row_g = staging + BM * 128 + ep_tid * 128;
#pragma unroll
        for (int q = 0; q < EPI_PAIRS / 8; ++q) {
          const int cj = ch * (EPI_PAIRS / 8) + q;
          const uint4 dv4 = *reinterpret_cast<const uint4*>(row_d + ((cj ^ r7) << 4));
          const uint32_t dw[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
          uint32_t ox[4], og[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q * 8 + e * 2;
            uint64_t X = pk2f(__uint_as_float(xv[j]), __uint_as_float(xv[j + 1]));
            uint64_t G = pk2f(__uint_as_float(gv[j]), __uint_as_float(gv[j + 1]));
            if (p.bias) {
              const float2 bx = __ldg(reinterpret_cast<const float2*>(p.bias + n0 + ch * EPI_PAIRS + j));
              const float2 bg = __ldg(reinterpret_cast<const float2*>(p.bias + n0 + 64 + ch * EPI_PAIRS + j));
              X = add2f(X, pk2f(bx.x, bx.y));
              G = add2f(G, pk2f(bg.x, bg.y));
            }
            const float2 d = unpack_bf16(dw[e]);
            const uint64_t D = pk2f(d.x, d.y);
            // two columns per instruction: th = tanh(g (a + b g^2)), hp = 0.5 (1 + th), gelu'(g) = hp + 0.5 g (a + 3 b g^2) (1 - th^2)
            const uint64_t T = mul2f(G, G);
            const uint64_t Z = mul2f(G, fma2f(pk2f(0.0347008941f, 0.0347008941f), T, pk2f(0.8001570768f, 0.8001570768f)));
            float z0, z1;
            upk2f(Z, z0, z1);
            const uint64_t TH = pk2f(tanh_fast(z0), tanh_fast(z1));
            const uint64_t HALF = pk2f(0.5f, 0.5f);
            const uint64_t HP = fma2f(HALF, TH, HALF);
            const uint64_t DZ = fma2f(pk2f(0.1041026823f, 0.1041026823f), T, pk2f(0.8001570768f, 0.8001570768f));
            const uint64_t SECH = fma2f(mul2f(TH, pk2f(-1.f, -1.f)), TH, pk2f(1.f, 1.f));
            const uint64_t GP = fma2f(mul2f(mul2f(HALF, G), DZ), SECH, HP);
            float rx0, rx1, rg0, rg1;
            upk2f(mul2f(D, mul2f(G, HP)), rx0, rx1);
            upk2f(mul2f(mul2f(D, X), GP), rg0, rg1);
            ox[e] = pack_bf16(rx0, rx1);
            og[e] = pack_bf16(rg0, rg1);
          }
          *reinterpret_cast<uint4*>(row_d + ((cj ^ r7) << 4)) = make_uint4(ox[0], ox[1], ox[2], ox[3]);
          *reinterpret_cast<uint4*>(row_g + ((cj ^ r7) << 4)) = make_uint4(og[0], og[1], og[2], og[3]);
        }
      } else if (EPI == EPI_GEGLU) {
        // columns [0,64) = value half, [64,128) = gate half (weights are packed that way)
        {
          const int ch = half;
          uint32_t xv[EPI_PAIRS], gv[EPI_PAIRS];
          tmem_ld_n<EPI_PAIRS>(t_addr + ch * EPI_PAIRS, xv);
          tmem_ld_n<EPI_PAIRS>(t_addr + 64 + ch * EPI_PAIRS, gv);
          tmem_ld_wait();
          uint8_t* dst = staging + ep_tid * 128;
#pragma unroll
          for (int q = 0; q < EPI_PAIRS / 8; ++q) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // value * gelu(gate) in the tanh form of geglu_fast_f, two columns per instruction
              const int j = q * 8 + e * 2;
              uint64_t X = pk2f(__uint_as_float(xv[j]), __uint_as_float(xv[j + 1]));
              uint64_t G = pk2f(__uint_as_float(gv[j]), __uint_as_float(gv[j + 1]));
              if (p.bias) {
                const float2 bx = __ldg(reinterpret_cast<const float2*>(p.bias + n0 + ch * EPI_PAIRS + j));
                const float2 bg = __ldg(reinterpret_cast<const float2*>(p.bias + n0 + 64 + ch * EPI_PAIRS + j));
                X = add2f(X, pk2f(bx.x, bx.y));
                G = add2f(G, pk2f(bg.x, bg.y));
              }
              const uint64_t T = mul2f(G, G);
              const uint64_t Z = mul2f(G, fma2f(pk2f(0.0347008941f, 0.0347008941f), T, pk2f(0.8001570768f, 0.8001570768f)));
              const uint64_t HG = mul2f(mul2f(X, pk2f(0.5f, 0.5f)), G);
              float z0, z1;
              upk2f(Z, z0, z1);
              float r0, r1;
              upk2f(fma2f(HG, pk2f(tanh_fast(z0), tanh_fast(z1)), HG), r0, r1);
              o[e] = pack_bf16(r0, r1);
            }
            const int cj = ch * (EPI_PAIRS / 8) + q;
            *reinterpret_cast<uint4*>(dst + ((cj ^ r7) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      } else {
#pragma unroll 1
        for (int ch = half * EPI_CH; ch < half * EPI_CH + EPI_CH; ++ch) {
          uint32_t v[32];
          tmem_ld32(t_addr + ch * 32, v);
          tmem_ld_wait();
          const int col0 = n0 + ch * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + col0 + j);
          }
          if (p.row_bias) {
            const float* rb = p.row_bias + (size_t)sample * p.N + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(rb + j);
          }
          if (has_res) {
            const uint8_t* rsrc = staging + (ch >> 1) * (BM * 128) + ep_tid * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int cjr = (ch & 1) * 4 + q;
              const uint4 r = *reinterpret_cast<const uint4*>(rsrc + ((cjr ^ r7) << 4));
              const float2 a = unpack_bf16(r.x), b = unpack_bf16(r.y), c = unpack_bf16(r.z),
                           d = unpack_bf16(r.w);
              f[q * 8 + 0] += a.x; f[q * 8 + 1] += a.y; f[q * 8 + 2] += b.x; f[q * 8 + 3] += b.y;
              f[q * 8 + 4] += c.x; f[q * 8 + 5] += c.y; f[q * 8 + 6] += d.x; f[q * 8 + 7] += d.y;
            }
          }
          if (p.act == ACT_LRELU) {  // nn.LeakyReLU() default slope 0.01
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : 0.01f * f[j];
          } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (p.act == ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);  // rel. err 2^-11, below the bf16 rounding of the result
          }
          uint8_t* dst = staging + (ch >> 1) * (BM * 128) + ep_tid * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o = make_uint4(pack_bf16(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16(f[q * 8 + 2], f[q * 8 + 3]),
                                 pack_bf16(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16(f[q * 8 + 6], f[q * 8 + 7]));
            const int cj = (ch & 1) * 4 + q;
            *reinterpret_cast<uint4*>(dst + ((cj ^ r7) << 4)) = o;
          }
        }
      }
      // TMEM accumulator (pair) drained -> hand it back to the MMA warp.
      tc_fence_before();
      if (hsub == subs - 1) mbar_arrive(tempty_bar(acc));
      // smem writes (generic proxy) -> visible to the TMA engine (async proxy).
      fence_proxy_async_smem();
      // bf16: before anyone moves on, the store of tile it-1 must have finished reading the OTHER staging buffer
      // (the next tile's residual prefetch and, one tile later, its result are written there)
      if (!OUT_F32 && ep_leader) tma_store_wait_read<0>();
      named_bar_sync(1, EPI_THREADS);
      if (ep_leader) {
        if (OUT_F32) {
#pragma unroll
          for (int ch = 0; ch < BN / 32; ++ch)
            tma_reduce_add_2d(&tmD, smem_staging + ch * (BM * 128), n0 + ch * 32, m0);
        } else {
          // the other staging buffer was the source of tile it-1's store: once that has been read it can take
          // the next tile's residual (prefetch) and, one tile later, the next result
          const int ntile = hsub + 1 < subs ? tile : tile + (int)gridDim.x;
          if (has_res && ntile < total_tiles) prefetch_residual(ntile, hsub + 1 < subs ? hsub + 1 : 0, sbuf ^ 1);
          if (EPI == EPI_GEGLU) {
            tma_store_2d(&tmD, staging_s, n0 / 2, m0);
          } else if (EPI == EPI_GEGLU_BWD) {  // dh8 in the plain layout: [d value (n_half) | d gate (n_half)]
            tma_store_2d(&tmD, staging_s, n0 / 2, m0);
            tma_store_2d(&tmD, staging_s + BM * 128, p.n_half + n0 / 2, m0);
          } else if (p.halo) {
            tma_store_4d(&tmD, staging_s, n0, h_x0, h_y0, h_img);
            tma_store_4d(&tmD, staging_s + BM * 128, n0 + 64, h_x0, h_y0, h_img);
          } else {
            tma_store_2d(&tmD, staging_s, n0, m0);
            tma_store_2d(&tmD, staging_s + BM * 128, n0 + 64, m0);
          }
        }
        tma_store_commit();
      }
      // by-products computed from the finished staging tile AFTER the leader has issued the store and the next residual
      // prefetch (doing them first delayed both by a microsecond per tile and made the epilogue the critical path)
      if (!OUT_F32 && p.gn_part) {
        // GroupNorm statistics of the tensor being written, as a by-product: per-channel sum and sum of squares of each
        // 32-row quarter of the finished tile, read back from the staging buffer (thread = one column of one half; the
        // bf16-rounded values, i.e. exactly what a later pass over the tensor would read).  Partials are stored, not
        // accumulated: the consumer reduces them in a fixed order, so the statistics stay bit-reproducible.
        const int et = (warp - 4) * 32 + lane;
        const int col = et & 127, rh = et >> 7;
        if (m0 + rh * GN_ROWS < p.M) {
          const uint8_t* base = staging + (col >> 6) * (BM * 128) + (col & 7) * 2;
          const int cc = (col & 63) >> 3;
#pragma unroll
          for (int sub = 0; sub < GN_ROWS / 32; ++sub) {  // partials have a fixed 32-row granularity
            const int r0 = rh * GN_ROWS + sub * 32;
            float sm = 0.f, sq = 0.f;
#pragma unroll 8
            for (int r = r0; r < r0 + 32; ++r) {
              const uint16_t v = *reinterpret_cast<const uint16_t*>(base + r * 128 + ((cc ^ (r & 7)) << 4));
              const float f = __uint_as_float(static_cast<uint32_t>(v) << 16);
              sm += f;
              sq = fmaf(f, f, sq);
            }
            if (m0 + r0 < p.M)
              *reinterpret_cast<float2*>(p.gn_part + ((size_t)(m_blk * 4 + (r0 >> 5)) * p.N + n0 + col) * 2) = make_float2(sm, sq);
          }
        }
      }
      if (!OUT_F32 && EPI == EPI_GEGLU_BWD && p.colsum) {
        // bias gradient of the C -> 8C linear as a by-product: column sums of the finished tile, read back from the
        // staging buffer (thread = one column of one 64-row half), kept per n-block in registers across the CTA's tiles
        const int et = (warp - 4) * 32 + lane;
        const int col = et & 63, which = (et >> 6) & 1, r_lo = (et >> 7) * GN_ROWS;
        const uint8_t* base = staging + which * (BM * 128) + (col & 7) * 2;
        float part = 0.f;
#pragma unroll 8
        for (int r = r_lo; r < r_lo + GN_ROWS; ++r) {
          const uint16_t v = *reinterpret_cast<const uint16_t*>(base + r * 128 + (((col >> 3) ^ (r & 7)) << 4));
          part += __uint_as_float(static_cast<uint32_t>(v) << 16);
        }
        cs_acc[n_blk & 15] += part;
      }
    }
    if (ep_leader) tma_store_wait_all<0>();
    if (!OUT_F32 && EPI == EPI_GEGLU_BWD && p.colsum) {
      const int et = (warp - 4) * 32 + lane;
      const int col = et & 63, which = (et >> 6) & 1;
      for (int nb = 0; nb < p.tiles_n && nb < 16; ++nb)
        if (cs_acc[nb] != 0.f) atomicAdd(p.colsum + which * p.n_half + nb * 64 + col, cs_acc[nb]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int A_MN, int B_MN, int OUT_F32, int EPI>
static int launch_t(cudaStream_t stream, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
                    const CUtensorMap& tmB0, const CUtensorMap& tmB1, const CUtensorMap& tmD,
                    const CUtensorMap& tmR, const GemmParams& p) {
  auto kern = gemm_tc_kernel<A_MN, B_MN, OUT_F32, EPI>;
  constexpr int smem = Cfg<OUT_F32>::SMEM_BYTES;
  static tsd::PerDeviceFlag configured;
  if (!configured.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.cur() = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, NUM_THREADS, smem, stream>>>(tmA0, tmA1, tmB0, tmB1, tmD, tmR, p);
  TSD_LAUNCH_CHECK();
  return 0;
}

int launch_gemm(cudaStream_t stream, int a_mn, int b_mn, int out_f32, const CUtensorMap& tmA0,
                const CUtensorMap& tmA1, const CUtensorMap& tmB0, const CUtensorMap& tmB1,
                const CUtensorMap& tmD, const GemmParams& p) {
  // residual [M][ldr] bf16 is read through its own tensor map (same tiling as D)
  CUtensorMap tmR = tmD;
  if (p.residual && !out_f32 && p.halo) {
    if (make_tmap_nhwc(&tmR, p.residual, p.M / (p.Ho * p.Wo), p.Ho, p.Wo, p.ldr, 64, HALO_TW, HALO_TH, 1, 1)) return 1;
  } else if (p.residual && !out_f32) {
    // EPI_GEGLU_BWD: the "residual" operand is d(gg) [M][n_half], one 64-column box per tile
    const int rcols = p.epi == EPI_GEGLU_BWD ? p.n_half : p.N;
    if (make_tmap_2d(&tmR, p.residual, 2, p.M, rcols, p.ldr, 64, 128)) return 1;
  }
  TSD_CHECK(p.gn_part == nullptr || (!out_f32 && p.epi == EPI_NONE && p.M % 64 == 0),
            "gemm: GroupNorm partials need a plain bf16 epilogue and M %% 64 == 0");
  TSD_CHECK(p.epi != EPI_GEGLU_BWD || (p.residual && !out_f32 && p.tiles_n <= 16 && p.n_half * 2 == p.N),
            "gemm: GEGLU backward epilogue needs d(gg), bf16 output and N = 2 * n_half <= 2048");
  TSD_CHECK(p.N % BN == 0, "gemm: N=%d must be a multiple of %d", p.N, BN);
  TSD_CHECK(!p.wg_halo || (a_mn && b_mn && out_f32 && p.stride == 1 && p.tiles_n % 3 == 0 && p.Ho % 8 == 0 && p.Wo % 8 == 0 &&
                           p.halo_tx == p.Wo / 8 && p.halo_tpi == (p.Ho / 8) * (p.Wo / 8) && p.b_ctot == (p.tiles_n / 3) * BN),
            "gemm: patch-mode weight gradient needs a stride-1 3x3 convolution on an 8 x 8 pixel patch grid");
  TSD_CHECK(p.halo != 3 || (p.Wo % (2 * HALO_TW) == 0 && p.tiles_m * 2 * BM == p.M), "gemm: halo mode 3 needs a width multiple of 16");
  TSD_CHECK(!p.halo || (!a_mn && !out_f32 && p.splits == 1 && p.stride == 1 && p.epi == EPI_NONE && p.Ho % HALO_TH == 0 &&
                        p.Wo % HALO_TW == 0 && p.halo_tx == p.Wo / HALO_TW &&
                        p.halo_tpi == (p.Ho / HALO_TH) * (p.Wo / HALO_TW)),
            "gemm: halo mode needs a stride-1 3x3 convolution on a %d x %d pixel patch grid", HALO_TH, HALO_TW);
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("TSD_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
  GemmParams pd = p;
  pd.dbg = dbg;
  const GemmParams& p2 = pd;
  TSD_CHECK(p.num_kb > 0 && p.splits > 0 && p.kb_per_split > 0, "gemm: empty K loop");
  pd.fd_tiles_mn = make_fastdiv(p.tiles_m * p.tiles_n);
  pd.fd_tiles_n = make_fastdiv(p.tiles_n);
  pd.fd_rps = make_fastdiv(p.rows_per_sample > 0 ? p.rows_per_sample : 1);
  pd.fd_tpi = make_fastdiv(p.halo_tpi > 0 ? p.halo_tpi : 1);
  pd.fd_tx = make_fastdiv(p.halo_tx > 0 ? p.halo_tx : 1);
  // the epilogue kind is a template parameter: each instantiation carries only its own epilogue code
  if (!a_mn && !b_mn && !out_f32 && p.epi == EPI_NONE) return launch_t<0, 0, 0, EPI_NONE>(stream, tmA0, tmA1, tmB0, tmB1, tmD, tmR, p2);
  if (!a_mn && !b_mn && !out_f32 && p.epi == EPI_GEGLU) return launch_t<0, 0, 0, EPI_GEGLU>(stream, tmA0, tmA1, tmB0, tmB1, tmD, tmR, p2);
  if (!a_mn && !b_mn && !out_f32 && p.epi == EPI_GEGLU_BWD) return launch_t<0, 0, 0, EPI_GEGLU_BWD>(stream, tmA0, tmA1, tmB0, tmB1, tmD, tmR, p2);
  if (!a_mn && b_mn && !out_f32 && p.epi == EPI_NONE) return launch_t<0, 1, 0, EPI_NONE>(stream, tmA0, tmA1, tmB0, tmB1, tmD, tmR, p2);
  if (a_mn && b_mn && out_f32 && p.epi == EPI_NONE) return launch_t<1, 1, 1, EPI_NONE>(stream, tmA0, tmA1, tmB0, tmB1, tmD, tmR, p2);
  set_error("gemm: unsupported operand-major / output combination (%d,%d,%d)", a_mn, b_mn, out_f32);
  return 1;
}

}  // namespace tsd
