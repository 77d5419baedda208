// Image input / output either side of the denoiser.
//
// Reference: 06_tiny_stable_diffusion/utils.py:10-29 -- the loader's ToTensor (uint8 HWC -> float CHW / 255) and
// Normalize((x - mean) / std), and `denormalize` (x * std + mean) followed by torchvision.utils.save_image
// (02_train_direct.py:24-27: make_grid(nrow, padding, pad_value 0), then mul(255).add(0.5).clamp(0, 255) -> uint8 HWC).
// Both directions are one pass over the pixels, HBM-bound, and bit-exact with the torch expressions: every
// floating-point step is an explicitly rounded IEEE operation in the reference's order (no FMA contraction).
#include "../../include/tinysd_b200.h"
#include "common.cuh"

using namespace tsd;

namespace {

constexpr int MAX_C = 4;
struct ChanStats { float mean[MAX_C], stdv[MAX_C]; };

// in  uint8 [N][H][W][C]  ->  out fp32 [N][C][H][W] = ((in / 255) - mean[c]) / std[c]
// A pixel value has 256 possibilities per channel, so every block first tabulates the exact result (two IEEE divisions
// and a subtraction per entry, 256*C entries in shared memory) and the streaming loop is a table lookup: one thread per
// 4 consecutive pixels (H*W % 4 == 0), 4*C bytes read, C float4 stores -- HBM-bound instead of division-bound.
template <int C>
__global__ void __launch_bounds__(256) u8_to_f32_norm_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                             size_t groups, int HW, ChanStats st) {
  __shared__ float lut[C][256];
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), st.mean[c]), st.stdv[c]);
  }
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < groups; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i * 4;
    const size_t n = pix / HW;
    const int p = (int)(pix - n * HW);
    uint8_t v[4 * C];
    if (C == 3) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(in + pix * 3);  // 12 bytes, 4-byte aligned (pix % 4 == 0)
      const uint32_t w[3] = {src[0], src[1], src[2]};
#pragma unroll
      for (int k = 0; k < 12; ++k) v[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    } else if (C == 4) {
      const uint4 w4 = *reinterpret_cast<const uint4*>(in + pix * 4);
      const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
      for (int k = 0; k < 4 * C; ++k) v[k] = in[pix * C + k];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float4 o;
      o.x = lut[c][v[0 * C + c]];
      o.y = lut[c][v[1 * C + c]];
      o.z = lut[c][v[2 * C + c]];
      o.w = lut[c][v[3 * C + c]];
      *reinterpret_cast<float4*>(out + (n * C + c) * (size_t)HW + p) = o;
    }
  }
}

__device__ __forceinline__ uint32_t denorm_q(float v, float sd, float mn) {
  const float d = __fadd_rn(__fmul_rn(v, sd), mn);
  float q = __fadd_rn(__fmul_rn(d, 255.f), 0.5f);
  q = fminf(fmaxf(q, 0.f), 255.f);
  return (uint32_t)q;  // float -> uint8 truncates
}

// x fp32 [N][C][H][W] -> grid uint8 [GH][GW][CO]: tile k at (k / xmaps, k % xmaps), `padding` pixels of 0 between
// tiles and around the border, value = trunc(clamp((x * std + mean) * 255 + 0.5, 0, 255)); C == 1 is replicated to 3.
// General path: one thread per grid pixel.
__global__ void __launch_bounds__(256) denorm_grid_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int N,
                                                             int C, int H, int W, int xmaps, int padding, int GH, int GW,
                                                             int CO, ChanStats st) {
  const size_t total = (size_t)GH * GW;
  const int th = H + padding, tw = W + padding;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int gy = (int)(i / GW), gx = (int)(i - (size_t)gy * GW);
    const int ty = (gy - padding) / th, tx = (gx - padding) / tw;
    const int y = gy - padding - ty * th, xx = gx - padding - tx * tw;
    const int k = ty * xmaps + tx;
    const bool inside = gy >= padding && gx >= padding && y < H && xx < W && tx < xmaps && k < N;
    for (int c = 0; c < CO; ++c) {
      uint8_t o = 0;
      if (inside) {
        const int cs = C == 1 ? 0 : c;
        o = (uint8_t)denorm_q(x[(((size_t)k * C + cs) * H + y) * W + xx], st.stdv[cs], st.mean[cs]);
      }
      out[i * CO + c] = o;
    }
  }
}

// Fast path for the reference's use (3 channels, padding 0, W % 4 == 0): one thread per 4 consecutive grid pixels of a
// row, which lie in one tile: three float4 loads (one per channel plane) and three 32-bit stores of packed RGB bytes.
__global__ void __launch_bounds__(256) denorm_grid_u8_rgb4_kernel(const float* __restrict__ x, uint8_t* __restrict__ out,
                                                                  int N, int H, int W, int xmaps, int GH, int GW,
                                                                  ChanStats st) {
  const int gw4 = GW / 4;
  const size_t total = (size_t)GH * gw4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int gy = (int)(i / gw4), gx = (int)(i - (size_t)gy * gw4) * 4;
    const int ty = gy / H, tx = gx / W;
    const int y = gy - ty * H, xx = gx - tx * W;
    const int k = ty * xmaps + tx;
    uint32_t q[12];
    if (k < N) {
      const float* base = x + (((size_t)k * 3) * H + y) * W + xx;
      const size_t plane = (size_t)H * W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(base + c * plane);
        q[0 * 3 + c] = denorm_q(v.x, st.stdv[c], st.mean[c]);
        q[1 * 3 + c] = denorm_q(v.y, st.stdv[c], st.mean[c]);
        q[2 * 3 + c] = denorm_q(v.z, st.stdv[c], st.mean[c]);
        q[3 * 3 + c] = denorm_q(v.w, st.stdv[c], st.mean[c]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) q[j] = 0;  // tiles past the last image keep pad_value 0
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((size_t)gy * GW + gx) * 3);
    dst[0] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
    dst[1] = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
    dst[2] = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
  }
}

inline int io_grid(size_t items) {
  size_t g = (items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int tsd_u8_to_f32_norm(void* stream, const void* in_hwc, float* out_chw, int N, int C, int H, int W,
                                  const float* mean, const float* stdv) {
  TSD_CHECK(C >= 1 && C <= MAX_C, "u8_to_f32_norm: C=%d not in [1, %d]", C, MAX_C);
  TSD_CHECK(((size_t)H * W) % 4 == 0, "u8_to_f32_norm: H*W must be a multiple of 4");
  ChanStats st;
  for (int c = 0; c < C; ++c) { st.mean[c] = mean[c]; st.stdv[c] = stdv[c]; }
  const size_t groups = (size_t)N * H * W / 4;
  if (groups == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const uint8_t* in = (const uint8_t*)in_hwc;
  switch (C) {
    case 1: u8_to_f32_norm_kernel<1><<<io_grid(groups), 256, 0, s>>>(in, out_chw, groups, H * W, st); break;
    case 2: u8_to_f32_norm_kernel<2><<<io_grid(groups), 256, 0, s>>>(in, out_chw, groups, H * W, st); break;
    case 3: u8_to_f32_norm_kernel<3><<<io_grid(groups), 256, 0, s>>>(in, out_chw, groups, H * W, st); break;
    default: u8_to_f32_norm_kernel<4><<<io_grid(groups), 256, 0, s>>>(in, out_chw, groups, H * W, st); break;
  }
  TSD_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsd_denorm_grid_u8(void* stream, const float* x, void* out_hwc, int N, int C, int H, int W, int nrow,
                                  int padding, const float* mean, const float* stdv) {
  TSD_CHECK(C >= 1 && C <= MAX_C && N >= 1 && nrow >= 1 && padding >= 0, "denorm_grid_u8: bad arguments");
  ChanStats st;
  for (int c = 0; c < C; ++c) { st.mean[c] = mean[c]; st.stdv[c] = stdv[c]; }
  if (N == 1) padding = 0;  // make_grid hands a single image back without the padding frame
  const int xmaps = nrow < N ? nrow : N;
  const int ymaps = (N + xmaps - 1) / xmaps;
  const int GH = (H + padding) * ymaps + padding, GW = (W + padding) * xmaps + padding;
  const int CO = C == 1 ? 3 : C;
  if (C == 3 && padding == 0 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out_hwc) & 3) == 0)
    denorm_grid_u8_rgb4_kernel<<<io_grid((size_t)GH * GW / 4), 256, 0, (cudaStream_t)stream>>>(x, (uint8_t*)out_hwc, N, H, W,
                                                                                               xmaps, GH, GW, st);
  else
    denorm_grid_u8_kernel<<<io_grid((size_t)GH * GW), 256, 0, (cudaStream_t)stream>>>(x, (uint8_t*)out_hwc, N, C, H, W, xmaps,
                                                                                      padding, GH, GW, CO, st);
  TSD_LAUNCH_CHECK();
  return 0;
}
