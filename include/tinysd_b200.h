/*
 * tinysd_b200.h -- C ABI of the B200-native (sm_100a) kernels behind the tiny-Stable-Diffusion
 * DDPM hot path of JAYANDJEAN/From_DDPM_to_Stable_Diffusion.
 *
 * The reference has no FFI of its own: its boundary is three Python classes and one function
 * (06_tiny_stable_diffusion/diffusion.py:183 Diffusion, utils.py:96 TrainerDDPM, utils.py:122
 * SamplerDDPM, utils.py:32 extract).  The Python shells in from_ddpm_to_stable_diffusion_b200/
 * keep those signatures and bind this library with ctypes; every entry point below states which
 * reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*
 *   - activations are bf16, channels-last: [n][h][w][c] == row-major [n*h*w][c]
 *   - every function returns 0 on success; on failure tsd_last_error() describes it
 *   - nothing allocates, nothing synchronises the stream
 *   - there is no CPU fallback: without an sm_100a device every call fails with a CUDA error
 */
#ifndef TINYSD_B200_H
#define TINYSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* tsd_last_error(void);
/* ABI version of this header; bumped whenever a signature changes. */
int tsd_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Weights are "packed": bf16 [n_out][k] with k contiguous; for a 3x3 conv k = tap*cin + ci,
 * tap = ky*3 + kx (tsd_pack_conv3x3 converts from the reference's OIHW fp32 layout).
 * ------------------------------------------------------------------------------------------ */

/* epilogue selector for tsd_gemm_fwd */
#define TSD_EPI_NONE 0
#define TSD_EPI_GEGLU 1 /* d[m][j] = (acc[x_j]+b) * gelu(acc[g_j]+b); weights packed by tsd_pack_geglu */

/* d[M][N] = [a0 | a1][M][c0+c1] * w[N][c0+c1]^T + bias[N] + row_bias[m / rows_per_sample][N] + residual[M][N]
 * Replaces nn.Linear / 1x1 nn.Conv2d calls: diffusion.py:43-44 (in/out_proj), :106 (shortcut),
 * :123,:136 (1x1 convs), :133-134 (GEGLU linears).  a1/c1 = second tensor of a channel concat
 * (diffusion.py:273) or NULL/0.  bias, row_bias, residual may be NULL. */
int tsd_gemm_fwd(void* stream, const void* a0, const void* a1, int c0, int c1, int M, const void* w, int N,
                 const float* bias, const float* row_bias, int rows_per_sample, const void* residual,
                 int epi, void* d);

/* 3x3 convolution, padding 1, stride 1 or 2, NHWC bf16, implicit GEMM (9 shifted TMA boxes, OOB
 * zero fill = padding).  x = channel concat of x0 (c0) and x1 (c1).  row_bias[row / rows_per_sample][cout]
 * is the time-embedding bias of diffusion.py:113 (rows_per_sample <= 0 means one row per image).
 * Replaces nn.Conv2d(k=3) at diffusion.py:92,98,164,210,214,218. */
int tsd_conv3x3_fwd(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int H, int W,
                    int stride, const void* w, int cout, const float* bias, const float* row_bias,
                    int rows_per_sample, const void* residual, void* d);

/* dx[M][K] = dy[M][N] * w[N][K] + residual[M][K]   (autograd of the ops above; reference relies on
 * torch autograd, 02_train_direct.py:71) */
int tsd_gemm_dgrad(void* stream, const void* dy, int M, int N, const void* w, int K, const void* residual,
                   void* dx);
/* stride-1 3x3 data gradient; w is the forward packed weight [cout][9*cin] */
int tsd_conv3x3_dgrad(void* stream, const void* dy, int n_img, int H, int W, int cout, const void* w, int cin,
                      const void* residual, void* dx);
/* dw[N][c0+c1] += dy[M][N]^T * [x0 | x1]  (fp32, split-K partials reduced by TMA reduce-add) */
int tsd_gemm_wgrad(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1, int M, int N,
                   float* dw);
/* dw[cout][9*(c0+c1)] += 3x3 weight gradient (packed layout), stride 1 or 2; H, W are INPUT dims */
int tsd_conv3x3_wgrad(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1, int n_img,
                      int H, int W, int stride, int cout, float* dw);

#ifdef __cplusplus
}
#endif
#endif /* TINYSD_B200_H */
