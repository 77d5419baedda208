// Latent codec either side of the denoiser: the data-movement and quantisation kernels of the repo's
// VQ-VAE (03_variational_autoencoder/models.py:135-185 VectorQuantizer, :268-378 VQVAE).  Every convolution of the codec
// runs on the tcgen05 GEMM core (gemm_tc.cu) with the activation in its epilogue:
//   4x4 stride-2 conv        = im2col (this file) + GEMM
//   3x3 / 1x1 conv           = the implicit-GEMM conv / plain GEMM of the UNet path
//   4x4 stride-2 ConvTranspose = ONE 3x3 implicit GEMM that produces the four output parities as 4 x Cout channels
//                              (each parity uses a 2x2 subset of the 3x3 window, the rest of the packed weight is zero)
//                              followed by a depth-to-space shuffle (this file)
//   nearest codebook entry   = fp32 distances in the reference's formula |z|^2 + |e|^2 - 2 z.e, first minimum
#include "../../include/tinysd_b200.h"
#include "common.cuh"

using namespace tsd;

namespace {

inline int ew_grid(size_t items) {
  size_t g = (items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// patch[p][(ky*KW + kx)*C + c] = x[n][oy*s - pad + ky][ox*s - pad + kx][c]  (zero outside the image and for k >= KH*KW*C)
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const bf16* __restrict__ x, bf16* __restrict__ patch, int n_img,
                                                          int H, int W, int C, int KH, int KW, int stride, int pad, int Ho,
                                                          int Wo, int Kp) {
  const int vec = Kp / 8, cvec = C / 8;
  const size_t total = (size_t)n_img * Ho * Wo * vec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kv = (int)(i % vec);
    size_t p = i / vec;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int n = (int)(p / Ho);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int tap = kv / cvec;
    if (tap < KH * KW) {
      const int c = (kv - tap * cvec) * 8;
      const int iy = oy * stride - pad + tap / KW, ix = ox * stride - pad + tap % KW;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = *reinterpret_cast<const uint4*>(x + (((size_t)n * H + iy) * W + ix) * C + c);
    }
    *reinterpret_cast<uint4*>(patch + i * 8) = v;
  }
}

// src [n][H][W][4][Cq] (row pitch ld >= 4*Cq; parity q = py*2 + px) -> dst [n][2H][2W][Cq]
__global__ void __launch_bounds__(256) depth_to_space2_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst,
                                                              int n_img, int H, int W, int Cq, int ld) {
  const int cvec = Cq / 8;
  const size_t total = (size_t)n_img * 4 * H * W * cvec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cvec) * 8;
    size_t p = i / cvec;
    const int ox = (int)(p % (2 * W)); p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    const int q = (oy & 1) * 2 + (ox & 1);
    const size_t row = ((size_t)n * H + (oy >> 1)) * W + (ox >> 1);
    *reinterpret_cast<uint4*>(dst + i * 8) = *reinterpret_cast<const uint4*>(src + row * ld + q * Cq + c);
  }
}

// src [n][H][W][ld] with channel q*Co + c (q = parity, c < Co) -> dst fp32 NCHW [n][Co][2H][2W]
__global__ void __launch_bounds__(256) d2s_to_nchw_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst,
                                                              int n_img, int H, int W, int Co, int ld) {
  const size_t total = (size_t)n_img * Co * 4 * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t p = i;
    const int ox = (int)(p % (2 * W)); p /= (2 * W);
    const int oy = (int)(p % (2 * H)); p /= (2 * H);
    const int c = (int)(p % Co);
    const int n = (int)(p / Co);
    const int q = (oy & 1) * 2 + (ox & 1);
    const size_t row = ((size_t)n * H + (oy >> 1)) * W + (ox >> 1);
    dst[i] = __bfloat162float(src[row * ld + q * Co + c]);
  }
}

// src bf16 [n*hw][ld] (first D channels) -> dst fp32 NCHW [n][D][hw]
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst,
                                                               int n_img, int hw, int D, int ld) {
  const size_t total = (size_t)n_img * D * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int px = (int)(i % hw);
    const int c = (int)((i / hw) % D);
    const int n = (int)(i / ((size_t)hw * D));
    dst[i] = __bfloat162float(src[((size_t)n * hw + px) * ld + c]);
  }
}

// VectorQuantizer.forward (models.py:149-176): for every latent vector z (D values at one spatial position) the index of
// the nearest codebook row under dist = sum z^2 + sum e^2 - 2 z.e (the reference's expression, fp32), first minimum on
// ties (torch.argmin); the quantised latent is z + (row - z), the reference's straight-through expression.  z / zq are
// NCHW fp32; thread = one latent vector.
constexpr int VQ_MAX_D = 16;
constexpr int VQ_TILE = 256;  // codebook rows staged in shared memory at a time
__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ codebook,
                                                         int64_t* __restrict__ idx, float* __restrict__ zq,
                                                         float* __restrict__ partial_sqerr, int n_img, int hw, int D, int K) {
  __shared__ float s_e[VQ_TILE * VQ_MAX_D];
  __shared__ float s_e2[VQ_TILE];
  __shared__ float s_red[8];
  const size_t total = (size_t)n_img * hw;
  const size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const bool live = v < total;
  const int n = live ? (int)(v / hw) : 0, px = live ? (int)(v % hw) : 0;
  float zr[VQ_MAX_D];
  float z2 = 0.f;
#pragma unroll
  for (int d = 0; d < VQ_MAX_D; ++d) {
    zr[d] = (live && d < D) ? z[((size_t)n * D + d) * hw + px] : 0.f;
    z2 += zr[d] * zr[d];
  }
  float best = INFINITY;
  int best_k = 0;
  for (int k0 = 0; k0 < K; k0 += VQ_TILE) {
    const int kt = min(VQ_TILE, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * D; i += blockDim.x) s_e[(i / D) * VQ_MAX_D + (i % D)] = codebook[(size_t)k0 * D + i];
    __syncthreads();
    for (int i = threadIdx.x; i < kt; i += blockDim.x) {
      float e2 = 0.f;
      for (int d = 0; d < D; ++d) e2 += s_e[i * VQ_MAX_D + d] * s_e[i * VQ_MAX_D + d];
      s_e2[i] = e2;
    }
    __syncthreads();
    for (int k = 0; k < kt; ++k) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < VQ_MAX_D; ++d)
        if (d < D) dot = fmaf(zr[d], s_e[k * VQ_MAX_D + d], dot);
      const float dist = (z2 + s_e2[k]) - 2.f * dot;
      if (dist < best) { best = dist; best_k = k0 + k; }
    }
  }
  float err = 0.f;
  if (live) {
    idx[v] = best_k;
    for (int d = 0; d < D; ++d) {
      const float e = codebook[(size_t)best_k * D + d];
      const float df = __fsub_rn(e, zr[d]);
      // the reference returns latents + (quantised - latents).detach() (models.py:180): the codebook row up to rounding
      zq[((size_t)n * D + d) * hw + px] = __fadd_rn(zr[d], df);
      err += df * df;
    }
  }
  // per-CTA partial of sum (zq - z)^2, reduced in a fixed order (the loss is reproducible run to run)
  err = warp_sum(err);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    partial_sqerr[blockIdx.x] = t;
  }
}
// vq_loss = beta * mse(zq.detach(), z) + mse(zq, z.detach()) = (1 + beta) * mean (zq - z)^2   (models.py:168-171)
__global__ void vq_loss_kernel(const float* __restrict__ partial, int n_part, float scale, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n_part; ++i) t += (double)partial[i];
    *loss = (float)(t * (double)scale);
  }
}

}  // namespace

extern "C" int tsd_im2col_nhwc(void* stream, const void* x, void* patch, int n_img, int H, int W, int C, int KH, int KW,
                               int stride, int pad, int Kp) {
  TSD_CHECK(C % 8 == 0 && Kp % 8 == 0 && Kp >= KH * KW * C, "im2col_nhwc: C=%d Kp=%d k=%dx%d", C, Kp, KH, KW);
  TSD_CHECK(stride >= 1 && (H + 2 * pad - KH) >= 0 && (W + 2 * pad - KW) >= 0, "im2col_nhwc: bad geometry");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  im2col_nhwc_kernel<<<ew_grid((size_t)n_img * Ho * Wo * (Kp / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (bf16*)patch, n_img, H, W, C, KH, KW, stride, pad, Ho, Wo, Kp);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_depth_to_space2(void* stream, const void* src, void* dst, int n_img, int H, int W, int Cq, int ld) {
  TSD_CHECK(Cq % 8 == 0 && ld >= 4 * Cq && ld % 8 == 0, "depth_to_space2: Cq=%d ld=%d", Cq, ld);
  depth_to_space2_kernel<<<ew_grid((size_t)n_img * 4 * H * W * (Cq / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)src, (bf16*)dst, n_img, H, W, Cq, ld);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_d2s_to_nchw_f32(void* stream, const void* src, float* dst, int n_img, int H, int W, int Co, int ld) {
  TSD_CHECK(Co >= 1 && ld >= 4 * Co, "d2s_to_nchw_f32: Co=%d ld=%d", Co, ld);
  d2s_to_nchw_f32_kernel<<<ew_grid((size_t)n_img * Co * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)src, dst, n_img, H, W, Co, ld);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_nhwc_to_nchw_f32(void* stream, const void* src, float* dst, int n_img, int hw, int D, int ld) {
  TSD_CHECK(D >= 1 && ld >= D, "nhwc_to_nchw_f32: D=%d ld=%d", D, ld);
  nhwc_to_nchw_f32_kernel<<<ew_grid((size_t)n_img * D * hw), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n_img,
                                                                                             hw, D, ld);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int64_t tsd_vq_scratch_floats(int n_img, int hw) { return ((int64_t)n_img * hw + 255) / 256; }
extern "C" int tsd_vq_nearest(void* stream, const float* z, const float* codebook, int64_t* idx, float* zq, float* scratch,
                              float* loss, float beta, int n_img, int hw, int D, int K) {
  TSD_CHECK(D >= 1 && D <= VQ_MAX_D && K >= 1, "vq_nearest: embedding_dim %d not in [1, %d] or empty codebook", D, VQ_MAX_D);
  const size_t total = (size_t)n_img * hw;
  const int grid = (int)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  vq_nearest_kernel<<<grid, 256, 0, st>>>(z, codebook, idx, zq, scratch, n_img, hw, D, K);
  TSD_LAUNCH_CHECK();
  if (loss) {
    vq_loss_kernel<<<1, 32, 0, st>>>(scratch, grid, (1.f + beta) / (float)(total * (size_t)D), loss);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}
