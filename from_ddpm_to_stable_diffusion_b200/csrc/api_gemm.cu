// C-ABI entry points for the dense contractions (linear / 1x1 conv / 3x3 conv; forward, data
// gradient, weight gradient).  All of them route to the single tcgen05 GEMM core in gemm_tc.cu.
#include "../../include/tinysd_b200.h"
#include "gemm_tc.cuh"
#include <cstdlib>
#include <cstring>

using namespace tsd;

namespace {

// Decompose P consecutive output pixels (P = 128 or 64) of an [n][Ho][Wo] grid into a TMA box.
int pixel_box(int Ho, int Wo, int P, uint32_t* bw, uint32_t* bh, uint32_t* bn) {
  if (Wo >= P) {
    TSD_CHECK(Wo % P == 0, "conv: output width %d not a multiple of the %d-pixel tile", Wo, P);
    *bw = P; *bh = 1; *bn = 1;
    return 0;
  }
  TSD_CHECK(P % Wo == 0, "conv: output width %d must divide the %d-pixel tile", Wo, P);
  const int rows = P / Wo;
  if (Ho >= rows) {
    TSD_CHECK(Ho % rows == 0, "conv: output height %d not a multiple of %d tile rows", Ho, rows);
    *bw = Wo; *bh = rows; *bn = 1;
    return 0;
  }
  TSD_CHECK(rows % Ho == 0, "conv: output height %d must divide %d tile rows", Ho, rows);
  *bw = Wo; *bh = Ho; *bn = rows / Ho;
  return 0;
}

// Halo mode of the stride-1 3x3 convolutions (gemm_tc.cu): the M tile becomes a 16 x 8 pixel patch (or two of them side
// by side) whose activations are loaded once per channel block instead of once per tap.  TSD_CONV_HALO: 0 = off
// (tap-by-tap boxes), 1 = three aligned x-shifted copies, 2 = one 10-pixel-wide copy with unaligned descriptor starts,
// 3 = two patches per tile (default; widths that are not a multiple of 16 fall back to 2) -- A/B switches for the
// measurements in DESIGN.md.  TSD_CONV_HALO_MINH: smallest image height that takes the halo path.
int halo_mode(int H, int W, int stride) {
  static int mode = -1, min_h = 16;
  if (mode < 0) {
    const char* e = getenv("TSD_CONV_HALO");
    mode = e ? atoi(e) : 3;
    if (mode < 0 || mode > 3) mode = 3;
    const char* h = getenv("TSD_CONV_HALO_MINH");
    if (h) min_h = atoi(h);
  }
  if (!mode || stride != 1 || H % 16 != 0 || W % 8 != 0 || H < min_h) return 0;
  return (mode == 3 && W % 16 != 0) ? 2 : mode;
}
// Tensor maps and tile geometry of a halo-mode convolution: a0 / a1 = the (up to two) NHWC sources, d = NHWC output.
int setup_halo(GemmParams& p, int mode, CUtensorMap* tA0, CUtensorMap* tA1, CUtensorMap* tD, const void* a0, const void* a1,
               int c0, int c1, void* d, int n_img, int H, int W, int cout) {
  const uint32_t bw = mode == 1 ? 8 : mode == 2 ? 10 : 18;
  if (make_tmap_nhwc(tA0, a0, n_img, H, W, c0, 64, bw, 18, 1, 1)) return 1;
  if (c1 > 0) { if (make_tmap_nhwc(tA1, a1, n_img, H, W, c1, 64, bw, 18, 1, 1)) return 1; } else *tA1 = *tA0;
  if (make_tmap_nhwc(tD, d, n_img, H, W, cout, 64, 8, 16, 1, 1)) return 1;
  p.halo = mode;
  p.halo_tx = W / 8;
  p.halo_tpi = (H / 16) * (W / 8);
  return 0;
}

void zero_params(GemmParams& p) { memset(&p, 0, sizeof(p)); p.splits = 1; p.b_cpt = 1 << 30; p.a_cpt = 1; }

}  // namespace

extern "C" int tsd_gemm_fwd_gn(void* stream, const void* a0, const void* a1, int c0, int c1, int M,
                               const void* w, int N, const float* bias, const float* row_bias,
                               int rows_per_sample, const void* residual, int epi, void* d, float* gn_part);
extern "C" int tsd_gemm_fwd(void* stream, const void* a0, const void* a1, int c0, int c1, int M,
                            const void* w, int N, const float* bias, const float* row_bias,
                            int rows_per_sample, const void* residual, int epi, void* d) {
  return tsd_gemm_fwd_gn(stream, a0, a1, c0, c1, M, w, N, bias, row_bias, rows_per_sample, residual, epi, d, nullptr);
}
extern "C" int tsd_gemm_fwd_gn(void* stream, const void* a0, const void* a1, int c0, int c1, int M,
                               const void* w, int N, const float* bias, const float* row_bias,
                               int rows_per_sample, const void* residual, int epi, void* d, float* gn_part) {
  const int K = c0 + c1;
  TSD_CHECK(M > 0 && N % 128 == 0 && K % 64 == 0 && c0 % 64 == 0, "gemm_fwd: bad shape M=%d N=%d K=%d c0=%d", M, N, K, c0);
  TSD_CHECK(c1 == 0 || a1 != nullptr, "gemm_fwd: second source missing");
  CUtensorMap tA0, tA1, tB, tD;
  if (make_tmap_2d(&tA0, a0, 2, M, c0, c0, 64, 128)) return 1;
  if (c1 > 0) { if (make_tmap_2d(&tA1, a1, 2, M, c1, c1, 64, 128)) return 1; } else tA1 = tA0;
  if (make_tmap_2d(&tB, w, 2, N, K, K, 64, 128)) return 1;
  TSD_CHECK(epi >= 0 && epi <= TSD_EPI_TANH, "gemm_fwd: unknown epilogue %d", epi);
  const int act = epi >= TSD_EPI_LRELU ? epi - TSD_EPI_LRELU + ACT_LRELU : ACT_NONE;  // pointwise activations
  if (act != ACT_NONE) epi = EPI_NONE;
  const int Nd = epi == EPI_GEGLU ? N / 2 : N;
  if (make_tmap_2d(&tD, d, 2, M, Nd, Nd, 64, 128)) return 1;
  GemmParams p; zero_params(p);
  p.M = M; p.N = N; p.tiles_m = ceil_div(M, 128); p.tiles_n = N / 128;
  p.num_kb = K / 64; p.kb_per_split = p.num_kb;
  p.a_mode = A_K2D; p.a_c0 = c0; p.b_mode = B_K2D;
  p.epi = epi; p.act = act; p.bias = bias; p.row_bias = row_bias; p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  p.gn_part = gn_part;
  p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = N;
  TSD_CHECK(!(epi == EPI_GEGLU && (residual || row_bias)), "gemm_fwd: GEGLU epilogue takes no residual/row bias");
  return launch_gemm((cudaStream_t)stream, 0, 0, 0, tA0, tA1, tB, tB, tD, p);
}

// dh8[M][2H] = backward of GEGLU(a w^T + bias) given d(gg)[M][H], with the pre-activations recomputed by the GEMM and
// the activation backward in its epilogue (w / bias in the GEGLU packing of tsd_pack_linear(geglu = 1)).
extern "C" int tsd_gemm_geglu_bwd(void* stream, const void* a, int M, int K, const void* w_geglu, int N,
                                  const float* bias_geglu, const void* dgg, void* dh8, float* dbias) {
  TSD_CHECK(M > 0 && N % 256 == 0 && N <= 2048 && K % 64 == 0, "gemm_geglu_bwd: bad shape M=%d N=%d K=%d", M, N, K);
  CUtensorMap tA, tB, tD;
  if (make_tmap_2d(&tA, a, 2, M, K, K, 64, 128)) return 1;
  if (make_tmap_2d(&tB, w_geglu, 2, N, K, K, 64, 128)) return 1;
  if (make_tmap_2d(&tD, dh8, 2, M, N, N, 64, 128)) return 1;
  GemmParams p; zero_params(p);
  p.M = M; p.N = N; p.tiles_m = ceil_div(M, 128); p.tiles_n = N / 128;
  p.num_kb = K / 64; p.kb_per_split = p.num_kb;
  p.a_mode = A_K2D; p.a_c0 = K; p.b_mode = B_K2D;
  p.epi = EPI_GEGLU_BWD; p.bias = bias_geglu; p.rows_per_sample = 1;
  p.residual = reinterpret_cast<const bf16*>(dgg); p.ldr = N / 2; p.n_half = N / 2; p.colsum = dbias;
  return launch_gemm((cudaStream_t)stream, 0, 0, 0, tA, tA, tB, tB, tD, p);
}

extern "C" int tsd_conv3x3_fwd_gn(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img,
                                  int H, int W, int stride, const void* w, int cout, const float* bias,
                                  const float* row_bias, int rows_per_sample, const void* residual, int act, void* d,
                                  float* gn_part);
extern "C" int tsd_conv3x3_fwd_act(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img,
                                   int H, int W, int stride, const void* w, int cout, const float* bias,
                                   const float* row_bias, int rows_per_sample, const void* residual, int act, void* d) {
  return tsd_conv3x3_fwd_gn(stream, x0, x1, c0, c1, n_img, H, W, stride, w, cout, bias, row_bias, rows_per_sample,
                            residual, act, d, nullptr);
}
extern "C" int tsd_conv3x3_fwd(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img,
                               int H, int W, int stride, const void* w, int cout, const float* bias,
                               const float* row_bias, int rows_per_sample, const void* residual, void* d) {
  return tsd_conv3x3_fwd_act(stream, x0, x1, c0, c1, n_img, H, W, stride, w, cout, bias, row_bias, rows_per_sample,
                             residual, 0, d);
}
extern "C" int tsd_conv3x3_fwd_gn(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img,
                                  int H, int W, int stride, const void* w, int cout, const float* bias,
                                  const float* row_bias, int rows_per_sample, const void* residual, int act, void* d,
                                  float* gn_part) {
  const int cin = c0 + c1;
  TSD_CHECK(act == 0 || (act >= TSD_EPI_LRELU && act <= TSD_EPI_TANH), "conv3x3_fwd_act: unknown activation %d", act);
  TSD_CHECK(stride == 1 || stride == 2, "conv3x3_fwd: stride must be 1 or 2");
  TSD_CHECK(cin % 64 == 0 && c0 % 64 == 0 && cout % 128 == 0, "conv3x3_fwd: bad channels c0=%d c1=%d cout=%d", c0, c1, cout);
  TSD_CHECK(H % stride == 0 && W % stride == 0, "conv3x3_fwd: H, W must be multiples of the stride");
  const int Ho = H / stride, Wo = W / stride;
  const int M = n_img * Ho * Wo;
  CUtensorMap tA0, tA1, tB, tD;
  GemmParams p; zero_params(p);
  if (make_tmap_2d(&tB, w, 2, cout, 9 * cin, 9 * cin, 64, 128)) return 1;
  if (const int hm = halo_mode(H, W, stride)) {
    if (setup_halo(p, hm, &tA0, &tA1, &tD, x0, x1, c0, c1, d, n_img, H, W, cout)) return 1;
  } else {
    uint32_t bw, bh, bn;
    if (pixel_box(Ho, Wo, 128, &bw, &bh, &bn)) return 1;
    if (make_tmap_nhwc(&tA0, x0, n_img, H, W, c0, 64, bw, bh, bn, stride)) return 1;
    if (c1 > 0) { if (make_tmap_nhwc(&tA1, x1, n_img, H, W, c1, 64, bw, bh, bn, stride)) return 1; } else tA1 = tA0;
    if (make_tmap_2d(&tD, d, 2, M, cout, cout, 64, 128)) return 1;
  }
  p.M = M; p.N = cout; p.tiles_m = p.halo == 3 ? M / 256 : ceil_div(M, 128); p.tiles_n = cout / 128;
  p.a_mode = A_KCONV; p.a_c0 = c0; p.a_cpt = cin / 64;
  p.num_kb = 9 * p.a_cpt; p.kb_per_split = p.num_kb;
  p.Ho = Ho; p.Wo = Wo; p.stride = stride; p.b_mode = B_K2D;
  p.bias = bias; p.row_bias = row_bias; p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : Ho * Wo;
  p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = cout;
  p.act = act ? act - TSD_EPI_LRELU + ACT_LRELU : ACT_NONE;
  p.gn_part = gn_part;
  return launch_gemm((cudaStream_t)stream, 0, 0, 0, tA0, tA1, tB, tB, tD, p);
}

// dX[M][K] = dY[M][N] * W[N][K] (+ residual): the packed forward weight is read as an MN-major B.
extern "C" int tsd_gemm_dgrad(void* stream, const void* dy, int M, int N, const void* w, int K,
                              const void* residual, void* dx) {
  TSD_CHECK(M > 0 && N % 64 == 0 && K % 128 == 0, "gemm_dgrad: bad shape M=%d N=%d K=%d", M, N, K);
  CUtensorMap tA, tB, tD;
  if (make_tmap_2d(&tA, dy, 2, M, N, N, 64, 128)) return 1;
  if (make_tmap_2d(&tB, w, 2, N, K, K, 64, 64)) return 1;
  if (make_tmap_2d(&tD, dx, 2, M, K, K, 64, 128)) return 1;
  GemmParams p; zero_params(p);
  p.M = M; p.N = K; p.tiles_m = ceil_div(M, 128); p.tiles_n = K / 128;
  p.num_kb = N / 64; p.kb_per_split = p.num_kb;
  p.a_mode = A_K2D; p.a_c0 = N;
  p.b_mode = B_MN2D; p.b_c0 = K; p.b_cpt = p.num_kb; p.b_ntaps = 1; p.b_tapstride = 0;
  p.rows_per_sample = 1;
  p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = K;
  return launch_gemm((cudaStream_t)stream, 0, 1, 0, tA, tA, tB, tB, tD, p);
}

// Stride-1 3x3 data gradient: dX = conv3x3(dY, W mirrored and transposed), same implicit GEMM with
// the forward weight [cout][9*cin] consumed as an MN-major operand, taps visited mirrored.
extern "C" int tsd_conv3x3_dgrad(void* stream, const void* dy, int n_img, int H, int W, int cout,
                                 const void* w, int cin, const void* residual, void* dx) {
  TSD_CHECK(cout % 64 == 0 && cin % 128 == 0, "conv3x3_dgrad: bad channels cin=%d cout=%d", cin, cout);
  const int M = n_img * H * W;
  CUtensorMap tA, tB, tD;
  GemmParams p; zero_params(p);
  if (make_tmap_2d(&tB, w, 2, cout, 9 * cin, 9 * cin, 64, 64)) return 1;
  if (const int hm = halo_mode(H, W, 1)) {
    CUtensorMap tA1;
    if (setup_halo(p, hm, &tA, &tA1, &tD, dy, nullptr, cout, 0, dx, n_img, H, W, cin)) return 1;
  } else {
    uint32_t bw, bh, bn;
    if (pixel_box(H, W, 128, &bw, &bh, &bn)) return 1;
    if (make_tmap_nhwc(&tA, dy, n_img, H, W, cout, 64, bw, bh, bn, 1)) return 1;
    if (make_tmap_2d(&tD, dx, 2, M, cin, cin, 64, 128)) return 1;
  }
  p.M = M; p.N = cin; p.tiles_m = p.halo == 3 ? M / 256 : ceil_div(M, 128); p.tiles_n = cin / 128;
  p.a_mode = A_KCONV; p.a_c0 = cout; p.a_cpt = cout / 64;
  p.num_kb = 9 * p.a_cpt; p.kb_per_split = p.num_kb;
  p.Ho = H; p.Wo = W; p.stride = 1;
  p.b_mode = B_MN2D; p.b_c0 = cin; p.b_cpt = cout / 64; p.b_ntaps = 9; p.b_flip = 1; p.b_tapstride = cin;
  p.rows_per_sample = H * W;
  p.residual = reinterpret_cast<const bf16*>(residual); p.ldr = cin;
  return launch_gemm((cudaStream_t)stream, 0, 1, 0, tA, tA, tB, tB, tD, p);
}

static void pick_splits(GemmParams& p) {
  // Work items = tiles * splits are dealt round-robin to one persistent CTA per SM: aim for (just under) two full
  // rounds so that no CTA runs a third item alone (wave quantisation), never more splits than k-blocks.
  const int tiles = p.tiles_m * p.tiles_n;
  int want = (2 * num_sms()) / tiles;
  if (want < 1) want = 1;
  if (want > p.num_kb) want = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, want);
  p.splits = ceil_div(p.num_kb, p.kb_per_split);
}

// dW[N][K] (fp32, accumulated) += dY[M][N]^T * [X0 | X1][M][K]
extern "C" int tsd_gemm_wgrad(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1,
                              int M, int N, float* dw) {
  const int K = c0 + c1;
  TSD_CHECK(M > 0 && N % 64 == 0 && K % 128 == 0 && c0 % 128 == 0, "gemm_wgrad: bad shape M=%d N=%d K=%d c0=%d", M, N, K, c0);
  CUtensorMap tA, tB0, tB1, tD;
  if (make_tmap_2d(&tA, dy, 2, M, N, N, 64, 64)) return 1;
  if (make_tmap_2d(&tB0, x0, 2, M, c0, c0, 64, 64)) return 1;
  if (c1 > 0) { if (make_tmap_2d(&tB1, x1, 2, M, c1, c1, 64, 64)) return 1; } else tB1 = tB0;
  if (make_tmap_2d(&tD, dw, 4, N, K, K, 32, 128)) return 1;
  GemmParams p; zero_params(p);
  p.M = N; p.N = K; p.tiles_m = ceil_div(N, 128); p.tiles_n = K / 128;
  p.num_kb = ceil_div(M, 64);
  p.a_mode = A_MN2D; p.b_mode = B_MN2D; p.b_c0 = c0; p.b_cpt = 1 << 30; p.b_ntaps = 1;
  p.rows_per_sample = 1;
  pick_splits(p);
  return launch_gemm((cudaStream_t)stream, 1, 1, 1, tA, tA, tB0, tB1, tD, p);
}

// dW[cout][9*cin] (fp32, accumulated) += sum over output pixels of dY[p][co] * X[p*s + tap][ci]
extern "C" int tsd_conv3x3_wgrad(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1,
                                 int n_img, int H, int W, int stride, int cout, float* dw) {
  const int cin = c0 + c1;
  TSD_CHECK(stride == 1 || stride == 2, "conv3x3_wgrad: stride must be 1 or 2");
  TSD_CHECK(cin % 128 == 0 && c0 % 128 == 0 && cout % 64 == 0, "conv3x3_wgrad: bad channels c0=%d c1=%d cout=%d", c0, c1, cout);
  const int Ho = H / stride, Wo = W / stride;
  const int M = n_img * Ho * Wo;
  static int wg_mode = -1;  // TSD_WGRAD_HALO=0: tap-by-tap boxes for every shape (A/B switch)
  if (wg_mode < 0) { const char* e = getenv("TSD_WGRAD_HALO"); wg_mode = e ? atoi(e) : 1; }
  if (wg_mode && stride == 1 && H % 8 == 0 && W % 8 == 0 && cout % 128 == 0) {
    // patch mode (gemm_tc.cu): k-blocks are 8 x 8 pixel patches, a work item = (kernel row, co block, ci block, pixel range)
    CUtensorMap tA, tB0, tB1, tD;
    if (make_tmap_nhwc(&tA, dy, n_img, H, W, cout, 64, 8, 8, 1, 1)) return 1;
    if (make_tmap_nhwc(&tB0, x0, n_img, H, W, c0, 64, 10, 8, 1, 1)) return 1;
    if (c1 > 0) { if (make_tmap_nhwc(&tB1, x1, n_img, H, W, c1, 64, 10, 8, 1, 1)) return 1; } else tB1 = tB0;
    if (make_tmap_2d(&tD, dw, 4, cout, 9 * cin, 9 * cin, 32, 128)) return 1;
    GemmParams p; zero_params(p);
    p.M = cout; p.N = 9 * cin; p.tiles_m = cout / 128; p.tiles_n = 3 * (cin / 128);
    p.num_kb = n_img * (H / 8) * (W / 8);
    p.a_mode = A_MN2D; p.b_mode = B_MNCONV; p.b_c0 = c0; p.b_ctot = cin;
    p.Ho = H; p.Wo = W; p.stride = 1; p.rows_per_sample = 1;
    p.wg_halo = 1; p.halo_tx = W / 8; p.halo_tpi = (H / 8) * (W / 8);
    // one round of work items over the SMs: every item ends with three 64 KB reduce-adds, so fewer, longer items
    const int tiles = p.tiles_m * p.tiles_n;
    int want = num_sms() / tiles;
    if (want < 1) want = 1;
    if (want > p.num_kb) want = p.num_kb;
    p.kb_per_split = ceil_div(p.num_kb, want);
    p.splits = ceil_div(p.num_kb, p.kb_per_split);
    return launch_gemm((cudaStream_t)stream, 1, 1, 1, tA, tA, tB0, tB1, tD, p);
  }
  uint32_t bw, bh, bn;
  if (pixel_box(Ho, Wo, 64, &bw, &bh, &bn)) return 1;
  CUtensorMap tA, tB0, tB1, tD;
  if (make_tmap_2d(&tA, dy, 2, M, cout, cout, 64, 64)) return 1;
  if (make_tmap_nhwc(&tB0, x0, n_img, H, W, c0, 64, bw, bh, bn, stride)) return 1;
  if (c1 > 0) { if (make_tmap_nhwc(&tB1, x1, n_img, H, W, c1, 64, bw, bh, bn, stride)) return 1; } else tB1 = tB0;
  if (make_tmap_2d(&tD, dw, 4, cout, 9 * cin, 9 * cin, 32, 128)) return 1;
  GemmParams p; zero_params(p);
  p.M = cout; p.N = 9 * cin; p.tiles_m = ceil_div(cout, 128); p.tiles_n = 9 * cin / 128;
  p.num_kb = ceil_div(M, 64);
  p.a_mode = A_MN2D; p.b_mode = B_MNCONV; p.b_c0 = c0; p.b_ctot = cin;
  p.Ho = Ho; p.Wo = Wo; p.stride = stride; p.rows_per_sample = 1;
  pick_splits(p);
  return launch_gemm((cudaStream_t)stream, 1, 1, 1, tA, tA, tB0, tB1, tD, p);
}
