// One tcgen05/TMEM/TMA GEMM core for every dense contraction on the UNet path.
//
//   D[M, N] = A[M, K] * B[N, K]^T  (+ epilogue), bf16 operands, fp32 accumulation in TMEM.
//
// Tile 128 x 128 x 64 (UMMA M=128, N=128, K=16 x 4 per stage), cta_group::1, persistent CTAs,
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM -> registers -> swizzled smem -> TMA store / TMA reduce-add).
// Two TMEM accumulators (2 x 128 columns) so the epilogue of tile i overlaps the main loop of i+1.
//
// Operand "modes" (how a k-block index turns into TMA coordinates):
//   A: K2D    row-major [M][K] (optionally the channel-concat of two tensors)        -> K-major smem
//      KCONV  NHWC activations, 3x3 taps by shifted 4-D boxes, zero fill = padding  -> K-major smem
//      MN2D   [K][M] (M contiguous), used for weight gradients (A = dY^T)           -> MN-major smem
//   B: K2D    packed weights [N][K]                                                 -> K-major smem
//      MN2D   packed weights read "transposed" (data gradients), or activations
//             [K][N] for the weight gradient of linears / 1x1 convs                 -> MN-major smem
//      MNCONV shifted NHWC boxes as the [K=pixels][N=channels] operand of the 3x3
//             weight gradient                                                       -> MN-major smem
#pragma once
#include "common.cuh"

namespace tsd {

enum : int { A_K2D = 0, A_KCONV = 1, A_MN2D = 2 };
enum : int { B_K2D = 0, B_MN2D = 1, B_MNCONV = 2 };
enum : int { EPI_NONE = 0, EPI_GEGLU = 1, EPI_GEGLU_BWD = 2 };
enum : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_TANH = 3 };  // applied last: act(acc + bias + residual)

// n / d for 0 <= n < 2^31 as one wide multiply and a shift (d >= 1, set up on the host)
struct FastDiv { uint32_t mul, sh; };
inline FastDiv make_fastdiv(int d) {
  int l = 0;
  while ((1ll << l) < d) ++l;
  const int p = 31 + l;
  FastDiv f;
  f.mul = static_cast<uint32_t>(((1ull << p) + static_cast<uint64_t>(d) - 1) / static_cast<uint64_t>(d));
  f.sh = static_cast<uint32_t>(p);
  return f;
}
__device__ __forceinline__ int fdiv(int n, FastDiv f) {
  return static_cast<int>((static_cast<uint64_t>(static_cast<uint32_t>(n)) * f.mul) >> f.sh);
}

struct GemmParams {
  int M, N;              // rows / cols of D (GEGLU: N counts the 2x-wide pre-activation columns)
  int tiles_m, tiles_n;
  int num_kb;            // k-blocks of 64
  int splits, kb_per_split;
  // ---- A operand
  int a_mode;
  int a_c0;              // channels (K2D: k elements) served by source 0; the rest come from source 1
  int a_cpt;             // KCONV: k-blocks per tap
  // conv geometry in OUTPUT space (shared by A_KCONV and B_MNCONV)
  int Ho, Wo, stride;
  // ---- B operand
  int b_mode;
  int b_c0;              // MN2D/MNCONV: n-columns (channels) served by source 0
  int b_cpt;             // MN2D: k-blocks per tap (data gradient); huge for plain [K][N]
  int b_tapstride;       // MN2D: column offset per tap in the packed weight
  int b_ntaps, b_flip;   // MN2D: taps and whether they are visited mirrored (conv data gradient)
  int b_ctot;            // MNCONV: channels per tap (C0 + C1)
  // ---- epilogue
  int epi;
  const float* bias;      // [N] or null
  const float* row_bias;  // [M / rows_per_sample][N] or null (time-embedding / cross-attention bias)
  int rows_per_sample;
  const bf16* residual;   // [M][ldr] or null
  int ldr;
  float* gn_part;         // null, or [ceil(M/64)][N][2] fp32: per-channel (sum, sum of squares) of every 64-row half-tile
  float* colsum;          // EPI_GEGLU_BWD: bias gradient [N] (+=) or null
  int n_half;             // EPI_GEGLU_BWD: N / 2 (column offset of the gate half in the plain layout)
  int dbg;                // timing experiments only (TSD_GEMM_DBG): 1 = skip B loads, 2 = skip A loads after the ring fill
  int halo;               // 3x3 stride-1 convolution in halo mode (gemm_tc.cu): 0 off, 1 = three aligned copies, 2 = one copy
  int halo_tx, halo_tpi;  // patches per image row / per image
  int wg_halo;            // 3x3 stride-1 weight gradient in patch mode (gemm_tc.cu); halo_tx / halo_tpi count 8 x 8 patches
  FastDiv fd_tiles_mn, fd_tiles_n, fd_rps, fd_tpi, fd_tx;  // filled in by launch_gemm
  int act;                // ACT_*: pointwise activation on the finished value (codec convolutions, vqvae models.py:286-341)
};

// Host launcher (defined in gemm_tc.cu).  tm* are fully-built tensor maps.
int launch_gemm(cudaStream_t stream, int a_mn, int b_mn, int out_f32, const CUtensorMap& tmA0,
                const CUtensorMap& tmA1, const CUtensorMap& tmB0, const CUtensorMap& tmB1,
                const CUtensorMap& tmD, const GemmParams& p);

}  // namespace tsd
