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
/* pointwise activation applied last, act(acc + bias + row_bias + residual): the codec convolutions of
 * 03_variational_autoencoder/models.py:286-341 (nn.LeakyReLU() slope 0.01, nn.ReLU, nn.Tanh) */
#define TSD_EPI_LRELU 2
#define TSD_EPI_RELU 3
#define TSD_EPI_TANH 4

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

/* Backward of the GEGLU feed-forward input (diffusion.py:151-152) without a stored pre-activation tensor:
 * dh8[M][N] = d/dh8 of (value * gelu(gate)) where h8 = a[M][K] w^T + bias is RECOMPUTED by this GEMM (w_geglu /
 * bias_geglu in the packing of tsd_pack_linear(geglu = 1) / tsd_pack_geglu_bias) and dgg[M][N/2] is the gradient of
 * the GEGLU output; the activation backward runs in the epilogue (tanh-form gelu, as the forward epilogue).  dh8 is
 * written in the plain layout [d value (N/2) | d gate (N/2)]; dbias (optional, fp32 [N], +=) receives its column sums. */
int tsd_gemm_geglu_bwd(void* stream, const void* a, int M, int K, const void* w_geglu, int N, const float* bias_geglu,
                       const void* dgg, void* dh8, float* dbias);
/* GroupNorm statistics as a by-product of the kernel that writes the tensor (north star: norm fused into the producing
 * epilogue).  tsd_gemm_fwd_gn / tsd_conv3x3_fwd_gn are tsd_gemm_fwd / tsd_conv3x3_fwd_act with one more output:
 * gn_part fp32 [M / 32][N][2] = (sum, sum of squares) per channel of every 32-row quarter-tile of d (the bf16-rounded
 * values; M % 64 == 0, plain epilogue).  tsd_gn_stats_from_parts reduces the partials of x0 (and x1, for a channel
 * concat) to the (mean, rstd) of tsd_gn_stats in a fixed order, without reading the tensors (hw % 64 == 0). */
int tsd_gemm_fwd_gn(void* stream, const void* a0, const void* a1, int c0, int c1, int M, const void* w, int N,
                    const float* bias, const float* row_bias, int rows_per_sample, const void* residual, int epi, void* d,
                    float* gn_part);
int tsd_conv3x3_fwd_gn(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int H, int W,
                       int stride, const void* w, int cout, const float* bias, const float* row_bias,
                       int rows_per_sample, const void* residual, int act, void* d, float* gn_part);
int tsd_gn_stats_from_parts(void* stream, const float* part0, const float* part1, int c0, int c1, int n_img, int hw,
                            float eps, float* stats);
/* tsd_conv3x3_fwd with a pointwise activation (0 or TSD_EPI_LRELU / RELU / TANH) on the finished value */
int tsd_conv3x3_fwd_act(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int H, int W,
                        int stride, const void* w, int cout, const float* bias, const float* row_bias,
                        int rows_per_sample, const void* residual, int act, void* d);

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

/* ------------------------------------------------------------------------------------------
 * Normalisation (HBM-bound, vectorised, fp32 statistics).
 * ------------------------------------------------------------------------------------------ */

/* GroupNorm(32 groups) statistics of x = cat(x0 [.., c0], x1 [.., c1]) per image: stats[n][32][2] =
 * (mean, rstd).  scratch: fp32 workspace of at least tsd_gn_scratch_floats(n_img) elements (per-CTA partial sums,
 * reduced in a fixed order: the statistics are bit-reproducible).
 * nn.GroupNorm at diffusion.py:90,95 (eps 1e-5), :122 (eps 1e-6), :258. */
int64_t tsd_gn_scratch_floats(int n_img);
int tsd_gn_stats(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int hw, float eps,
                 float* scratch, float* stats);
/* Random-stream position kept in DEVICE memory: rng_dev[0] = number of calls so far, rng_dev[1] = GLOBAL index of this
 * rank's first sample.  Every kernel that draws random numbers (dropout, q_sample noise, timesteps, sampler z) takes an
 * optional pointer to it (NULL = {0, 0}): a captured CUDA graph then draws fresh numbers on each replay (advance the
 * counter with tsd_counter_add_u64 inside the graph), and what a sample draws depends on its global index only, so a
 * batch sharded over N ranks reproduces the single-rank stream. */
int tsd_counter_add_u64(void* stream, uint64_t* counter, uint64_t delta);

/* out = dropout_p(silu?(gamma * (x - mean) * rstd + beta)), bf16 [n*hw][c0+c1].  The dropout mask is a
 * counter-based Philox function of (seed, rng_dev position, element index); nn.SiLU / nn.Dropout at diffusion.py:91,96-97. */
int tsd_gn_apply(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int hw, const float* stats,
                 const float* gamma, const float* beta, int act_silu, float drop_p, uint64_t seed, void* out,
                 const uint64_t* rng_dev);
/* Backward of tsd_gn_apply (+ optional residual add `radd` [n*hw][c0+c1]); dx is written split as dx0 [.., c0],
 * dx1 [.., c1]; dgamma/dbeta (fp32 [c0+c1]) are accumulated.  ab: fp32 scratch [n_img][c0+c1][2].
 * colsum_out (optional, fp32 [n_img][c0+c1], +=) / colsum_total (optional, [c0+c1], +=): column sums of dx per image and
 * over all images as a by-product (the time-embedding and conv-bias gradients of diffusion.py:112-113). */
int tsd_gn_bwd(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1, int n_img, int hw,
               const float* stats, const float* gamma, const float* beta, int act_silu, float drop_p, uint64_t seed,
               float* ab, const void* radd, void* dx0, void* dx1, float* dgamma, float* dbeta, const uint64_t* rng_dev,
               float* colsum_out, float* colsum_total);
/* LayerNorm over C in {128, 256, 512} per token row; nn.LayerNorm at diffusion.py:127,132 */
int tsd_ln_fwd(void* stream, const void* x, int M, int C, const float* gamma, const float* beta, float eps, void* out);
/* colsum_out (optional, fp32 [M / rows_per_sample][C], +=) / colsum_total (optional, [C], +=): column sums of dx per
 * sample and over all rows as a by-product (gradients of the bias / per-sample vector added before the LayerNorm) */
int tsd_ln_bwd(void* stream, const void* dy, const void* x, int M, int C, const float* gamma, float eps,
               const void* radd, void* dx, float* dgamma, float* dbeta, int rows_per_sample, float* colsum_out,
               float* colsum_total);

/* ------------------------------------------------------------------------------------------
 * Self-attention, head_dim 16 / 32 (SelfAttention.forward, diffusion.py:46-58).
 * qkv bf16 [B*L][3C] (q | k | v, head h = columns [h*dh, (h+1)*dh)); out bf16 [B*L][C];
 * lse2 fp32 [B][heads][L] = log2-domain logsumexp (NULL when no backward is needed).
 * ------------------------------------------------------------------------------------------ */
int tsd_attn_fwd(void* stream, const void* qkv, void* out, float* lse2, int B, int L, int C, int heads);
/* Same, with a caller-provided fp32 scratch of B*heads floats: lets the tcgen05 path (head_dim 16 / 32, L % 256 == 0) bound
 * every score row by |q| max|k| and skip the running-maximum pass.  ws == NULL behaves like tsd_attn_fwd. */
int tsd_attn_fwd_ws(void* stream, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C, int heads);
/* dqkv [B*L][3C] receives (dq | dk | dv); delta: fp32 scratch [B][heads][L] */
int tsd_attn_bwd(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2, float* delta,
                 void* dqkv, int B, int L, int C, int heads);
/* Same, with a caller-provided fp32 scratch of B*L*C floats: enables the one-pass backward for head_dim 16 and
 * L % 256 == 0 (dK/dV and dQ from one evaluation of the scores; dQ partials are summed in the scratch by bulk
 * reduce-adds, so dQ is reproducible up to fp32 summation order).  ws == NULL behaves like tsd_attn_bwd. */
int tsd_attn_bwd_ws(void* stream, const void* qkv, const void* out, const void* dout, const float* lse2, float* delta,
                    void* dqkv, float* ws, int B, int L, int C, int heads);

/* ------------------------------------------------------------------------------------------
 * Elementwise / small reductions on bf16 channels-last tensors.
 * ------------------------------------------------------------------------------------------ */
int tsd_add_bf16(void* stream, const void* a, const void* b, void* out, int64_t numel);
/* GEGLU: out[M][H] = h8[:, :H] * gelu(h8[:, H:]) (exact erf GELU; diffusion.py:151-152) and its backward */
int tsd_geglu_fwd(void* stream, const void* h8, void* out, int64_t M, int H);
/* dbias (optional, fp32 [2H]) += column sums of dh8: the bias gradient of the C -> 8C linear as a by-product */
int tsd_geglu_bwd(void* stream, const void* h8, const void* dout, void* dh8, int64_t M, int H, float* dbias);
/* nearest x2 upsample (F.interpolate, diffusion.py:167) and its adjoint (2x2 block sums) */
int tsd_upsample2_fwd(void* stream, const void* in, void* out, int n_img, int H, int W, int C);
int tsd_upsample2_bwd(void* stream, const void* dout, void* din, int n_img, int H, int W, int C);
/* zero-stuffing [n][H][W][C] -> [n][2H][2W][C]: turns the stride-2 conv data gradient into a stride-1 one */
int tsd_zero_stuff2(void* stream, const void* in, void* out, int n_img, int H, int W, int C);
/* out[n][C] += column sums of sample n's rows (time-bias / cross-attention gradients) and, when total != NULL,
 * total[C] += the sums over all samples (bias gradients); tsd_reduce_rows_f32: out[C] += sum_n in[n][C] */
int tsd_colsum(void* stream, const void* x, int n_samples, int rows_per_sample, int C, float* out, float* total);
int tsd_reduce_rows_f32(void* stream, const float* in, int n_rows, int C, float* out);

/* ------------------------------------------------------------------------------------------
 * Conditioning path, fp32 (TimestepEmbedder diffusion.py:13-37, label_embedding :196-201, linear_time
 * :101-104, degenerate CrossAttention :61-82).  Weights in the reference's [out][in] fp32 layout.
 * ------------------------------------------------------------------------------------------ */
/* out[M][N] = f(x)[M][K] * w[N][K]^T + bias, f = SiLU when silu_in */
int tsd_small_linear_fwd(void* stream, const float* x, const float* w, const float* bias, float* out, int M, int N,
                         int K, int silu_in);
/* dx (=|+=) f'(x) * (dy * w); dw += dy^T f(x); db += colsum(dy).  dx, dw, db may each be NULL. */
int tsd_small_linear_bwd(void* stream, const float* dy, const float* x, const float* w, float* dx, float* dw,
                         float* db, int M, int N, int K, int silu_in, int accumulate_dx);
/* n (<= 16) independent small linears with a common M in one launch each way (the 14 linear_time layers of the
 * ResBlocks, the 10 v_proj and 10 out_proj of the degenerate cross-attention: separate parameters, same batch).  x, w,
 * bias, out, dy, dx, dw, db are HOST arrays of n device pointers, N / K host arrays of n ints; entries of dx may alias
 * (accumulate_dx: summed with atomics), NULL entries of dx / dw / db are skipped.  Backward needs M >= 32. */
int tsd_small_linear_many_fwd(void* stream, int n, const float* const* x, const float* const* w, const float* const* bias,
                              float* const* out, const int* N, const int* K, int M, int silu_in);
int tsd_small_linear_many_bwd(void* stream, int n, const float* const* dy, const float* const* x, const float* const* w,
                              float* const* dx, float* const* dw, float* const* db, const int* N, const int* K, int M,
                              int silu_in, int accumulate_dx);
/* emb[m] = [cos(t_m * freqs), sin(t_m * freqs)] (cos first, diffusion.py:28); freqs built by the host shell */
int tsd_timestep_embedding(void* stream, const int64_t* t, const float* freqs, float* emb, int M, int half);
int tsd_embedding_fwd(void* stream, const int64_t* idx, const float* table, float* out, int M, int D);
int tsd_embedding_bwd(void* stream, const int64_t* idx, const float* dy, float* dtable, int M, int D, int padding_idx);

/* ------------------------------------------------------------------------------------------
 * Image-side convolutions (fp32 NCHW image <-> bf16 NHWC features) and the DDPM process.
 * ------------------------------------------------------------------------------------------ */
/* head conv channel_img(<=4) -> co (diffusion.py:206); weights OIHW fp32 */
int tsd_head_conv_fwd(void* stream, const float* x, const float* w, const float* bias, void* out, int n_img, int ci,
                      int H, int W, int co);
/* operands that put the two skinny weight gradients on the tensor cores: patch[p][c*9+tap] (bf16, zero-padded to KP
 * columns) of the fp32 NCHW image for tsd_gemm_wgrad, and d(eps) as bf16 NHWC padded to CP channels for tsd_conv3x3_wgrad */
int tsd_im2col_head(void* stream, const float* x, void* patch, int n_img, int ci, int H, int W, int KP);
int tsd_nchw_to_nhwc_pad(void* stream, const float* src, void* dst, int n_img, int co, int HW, int CP);
/* tail conv 128 -> co in {3,4} on the GroupNorm+SiLU'ed features (diffusion.py:260); out fp32 NCHW; image width in
 * {16, 32, 64} and H*W a multiple of 128 (tensor-core implicit GEMM, 128-pixel row blocks) */
int tsd_tail_conv_fwd(void* stream, const void* a, const float* w, const float* bias, float* out, int n_img, int H,
                      int W, int c_in, int co);
int tsd_tail_conv_dgrad(void* stream, const float* dy, const float* w, void* da, int n_img, int H, int W, int c_in,
                        int co);
/* x_t = sqrt_ab[t_n] x0 + sqrt_1mab[t_n] noise (utils.py:115-116); noise_in NULL => Philox N(0,1), written to
 * noise_out.  Tables are the fp32 casts of the reference's fp64 buffers (what extract() returns). */
int tsd_q_sample(void* stream, const float* x0, const int64_t* t, const float* sqrt_ab, const float* sqrt_1mab,
                 const float* noise_in, uint64_t seed, uint64_t offset, float* x_t, float* noise_out, int n_img,
                 int64_t per_sample, const uint64_t* rng_dev);
/* t[n] ~ U{0..T-1} (torch.randint at utils.py:112), Philox keyed by (seed, rng_dev position, n) */
int tsd_draw_timesteps(void* stream, int64_t* t, int n, int T, uint64_t seed, const uint64_t* rng_dev);
/* loss = (pred - noise)^2 un-reduced (utils.py:118); dpred = 2 (pred - noise) gout */
int tsd_mse_fwd(void* stream, const float* pred, const float* noise, float* loss, int64_t total);
int tsd_mse_bwd(void* stream, const float* pred, const float* noise, const float* gout, float* dpred, int64_t total);
/* One reverse step (utils.py:149-166): eps = (1+w) eps[0:total] - w eps[total:2 total]; x' = c1[t] x - c2[t] eps +
 * sigma[t] z; z = 0 at t = 0, Philox keyed by (seed, t) unless noise_in; t = *step_ptr (device memory, so a CUDA
 * graph can replay the launch); NaN sets *nan_flag (utils.py:167); clip_last clamps to [-1,1] at t = 0 (:171);
 * dup also writes x' to x_out[total:2 total] (the unconditional copy of the 2B batch). */
int tsd_sampler_update(void* stream, const float* x, const float* eps, const int* step_ptr, const float* c1,
                       const float* c2, const float* sigma, float w, const float* noise_in, uint64_t seed,
                       float* x_out, int* nan_flag, int64_t total, int clip_last, int dup, const uint64_t* rng_dev,
                       int n_img);
/* Fused sampling tail: final conv 128 -> co over BOTH halves of the 2B batch `a` (rows [0,B) conditional, [B,2B)
 * unconditional; diffusion.py:260) with the reverse-step update above in its epilogue: x ([2B,co,H,W] fp32, both
 * halves identical) is read once and overwritten once per step and eps never reaches HBM (eps_out != NULL dumps
 * eps_c | eps_u for tests).  Replaces utils.py:151-166 after the UNet body. */
int tsd_tail_conv_sample(void* stream, const void* a, const float* w, const float* bias, float* x, const int* step_ptr,
                         const float* c1, const float* c2, const float* sigma, float wcfg, const float* noise_in,
                         uint64_t seed, int* nan_flag, float* eps_out, int B, int H, int W, int c_in, int co,
                         int clip_last, const uint64_t* rng_dev);
int tsd_step_add(void* stream, int* step_ptr, int delta);
/* out[0:len] = table[*step_ptr][0:len] (per-step time-embedding rows, indexed on the device) */
int tsd_gather_row_f32(void* stream, const float* table, const int* step_ptr, int len, float* out);

/* ------------------------------------------------------------------------------------------
 * Weight packing and the caller-side optimiser step (02_train_direct.py:72-73).
 * ------------------------------------------------------------------------------------------ */
/* Every weight packing of a step in one launch.  table_dev: DEVICE array of n_desc records
 *   struct { const float* src; void* dst; int64_t begin; int32_t rows, cols, kind, pad; }   (40 bytes)
 * sorted by `begin` (first flat output element of the tensor; total = sum of all element counts).  kind selects the
 * layouts of the single-tensor entry points below; conv tensors pass rows = co, cols = ci. */
#define TSD_PACK_LINEAR 0
#define TSD_PACK_LINEAR_GEGLU 1
#define TSD_PACK_CONV3X3 2
#define TSD_PACK_CONV3X3_DGRAD 3
#define TSD_PACK_GEGLU_BIAS 4
int tsd_pack_many(void* stream, const void* table_dev, int n_desc, int64_t total);
int tsd_pack_linear(void* stream, const float* src, void* dst, int rows, int cols, int geglu);
int tsd_pack_geglu_bias(void* stream, const float* src, float* dst, int rows);
int tsd_pack_conv3x3(void* stream, const float* src, void* dst, int co, int ci);          /* OIHW -> [co][tap][ci] */
/* OIHW -> [ci][8-tap][co]: the data gradient of a stride-1 3x3 conv is tsd_conv3x3_fwd(dy, this weight) */
int tsd_pack_conv3x3_dgrad(void* stream, const float* src, void* dst, int co, int ci);
int tsd_unpack_conv3x3_grad(void* stream, const float* src, float* dst, int co, int ci);  /* dst(OIHW) += src */
/* out[0] += sum g^2, bit-reproducible (fixed-order reduction of per-CTA partials through `scratch`,
 * tsd_sumsq_scratch_floats() floats, zero-filled once before the first call): data-parallel ranks must derive the
 * identical clip coefficient from their identical all-reduced gradients */
int64_t tsd_sumsq_scratch_floats(void);
int tsd_sumsq_f32(void* stream, const float* g, int64_t n, float* out, float* scratch);
/* clip_grad_norm_(max_norm) + torch.optim.AdamW step over flat fp32 buffers; sumsq = squared global grad norm */
int tsd_adamw_clip(void* stream, float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float wd, int step, float max_norm, const float* sumsq, int write_clipped_grad);
/* Same with the learning rate and the step count (already incremented, >= 1) read from device memory: the launch can
 * be replayed from a CUDA graph while the LR schedule and the bias correction advance */
int tsd_adamw_clip_dev(void* stream, float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                       const int* step_dev, float beta1, float beta2, float eps, float wd, float max_norm,
                       const float* sumsq, int write_clipped_grad);
int tsd_scale_f32(void* stream, float* x, int64_t n, float s);
/* shadow = (1-decay)*p + decay*shadow over a flat fp32 buffer (EMA.update, utils.py:54-58); one_minus_decay is the
 * host's float(1.0 - decay) and every step is separately rounded, so the result equals the torch expression bit for bit */
int tsd_ema_update(void* stream, float* ema, const float* p, int64_t n, float decay, float one_minus_decay);
/* dst[r][c] += src[r][c], c < cols (row pitches ldd / lds): folds padded tensor-core gradients into parameter grads */
int tsd_add_cols_f32(void* stream, float* dst, const float* src, int rows, int cols, int ldd, int lds);
/* ------------------------------------------------------------------------------------------
 * Image input / output either side of the denoiser (utils.py:10-29; 02_train_direct.py:24-27).
 * in_hwc uint8 [N][H][W][C] -> out_chw fp32 [N][C][H][W] = ((in / 255) - mean[c]) / std[c]   (ToTensor + Normalize);
 * mean / std are HOST arrays of C floats.  Bit-exact with the torch expressions.
 * ------------------------------------------------------------------------------------------ */
int tsd_u8_to_f32_norm(void* stream, const void* in_hwc, float* out_chw, int N, int C, int H, int W, const float* mean,
                       const float* stdv);
/* x fp32 [N][C][H][W] -> out_hwc uint8 [(H+pad)*ceil(N/min(nrow,N))+pad][(W+pad)*min(nrow,N)+pad][C or 3]:
 * denormalize (x*std+mean), torchvision make_grid(nrow, padding, pad_value 0) and save_image's uint8 conversion
 * (N == 1: the image itself, no padding frame, as make_grid returns it) */
int tsd_denorm_grid_u8(void* stream, const float* x, void* out_hwc, int N, int C, int H, int W, int nrow, int padding,
                       const float* mean, const float* stdv);
/* ------------------------------------------------------------------------------------------
 * Latent codec either side of the denoiser: the repo's VQ-VAE,
 * 03_variational_autoencoder/models.py:135-185 (VectorQuantizer) and :268-378 (VQVAE encode / decode).  Its
 * convolutions run on the GEMM core above (tsd_gemm_fwd / tsd_conv3x3_fwd_act with TSD_EPI_LRELU / RELU / TANH);
 * these are the data-movement and quantisation kernels around them.  Activations bf16 NHWC as everywhere else.
 * ------------------------------------------------------------------------------------------ */
/* patch[p][(ky*KW + kx)*C + c] = x[n][oy*stride - pad + ky][ox*stride - pad + kx][c], zero outside the image and for
 * columns >= KH*KW*C; patch is bf16 [n*Ho*Wo][Kp]: the A operand of a KHxKW strided nn.Conv2d (models.py:286) */
int tsd_im2col_nhwc(void* stream, const void* x, void* patch, int n_img, int H, int W, int C, int KH, int KW, int stride,
                    int pad, int Kp);
/* src [n][H][W][ld] holding the four output parities q = py*2 + px as channel blocks [q*Cq, (q+1)*Cq) -> dst
 * [n][2H][2W][Cq]: the shuffle behind nn.ConvTranspose2d(k=4, s=2, p=1) (models.py:330-341) */
int tsd_depth_to_space2(void* stream, const void* src, void* dst, int n_img, int H, int W, int Cq, int ld);
/* same shuffle for the last layer, straight to the fp32 NCHW image [n][Co][2H][2W] */
int tsd_d2s_to_nchw_f32(void* stream, const void* src, float* dst, int n_img, int H, int W, int Co, int ld);
/* first D channels of bf16 [n*hw][ld] -> fp32 NCHW [n][D][hw] (the encoder's latent output, models.py:351-352) */
int tsd_nhwc_to_nchw_f32(void* stream, const void* src, float* dst, int n_img, int hw, int D, int ld);
/* VectorQuantizer.forward (models.py:149-185): idx[n*hw] = argmin_k |z|^2 + |e_k|^2 - 2 z.e_k (fp32, first minimum),
 * zq = z + (codebook[idx] - z) (the reference's straight-through expression, :180; NCHW fp32 like z),
 * *loss = (1 + beta) * mean((codebook[idx] - z)^2) when loss != NULL.
 * scratch: tsd_vq_scratch_floats(n_img, hw) floats.  embedding_dim D <= 16. */
int64_t tsd_vq_scratch_floats(int n_img, int hw);
int tsd_vq_nearest(void* stream, const float* z, const float* codebook, int64_t* idx, float* zq, float* scratch,
                   float* loss, float beta, int n_img, int hw, int D, int K);
/* number of kernel launches issued by this library so far (host-side counter, for bench.py's gpu_launches) */
unsigned long long tsd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TINYSD_B200_H */
