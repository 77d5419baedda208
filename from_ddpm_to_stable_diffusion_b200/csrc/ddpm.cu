// DDPM process kernels and the two skinny convolutions that touch the fp32 NCHW image tensor.
//
//   head conv  channel_img -> 128, K = 27/36: not tensor-core shaped, CUDA cores     (diffusion.py:206)
//   tail conv  128 -> channel_img, N = 3/4                                            (diffusion.py:260)
//   q_sample + on-device Philox noise                                                 (utils.py:112-116)
//   noise-MSE forward / backward                                                      (utils.py:118)
//   CFG combine + posterior update + Philox z + NaN flag                              (utils.py:149-167)
#include "../../include/tinysd_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

using namespace tsd;

namespace {

constexpr int MAX_CI = 4;    // image / latent channels supported (3 or 4)
constexpr int HEAD_CO = 128;

// ------------------------------------------------------------------------------------------ head conv
// x fp32 NCHW [n][ci][H][W] -> out bf16 NHWC [n][H][W][co]; thread = (pixel, 8 output channels)
__global__ void __launch_bounds__(256) head_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ out,
                                                            int n_img, int ci, int H, int W, int co) {
  extern __shared__ float sw[];  // [ci*9][co] (transposed for conflict-free reads) + bias[co]
  float* sb = sw + ci * 9 * co;
  for (int i = threadIdx.x; i < co * ci * 9; i += blockDim.x) {
    const int o = i / (ci * 9), r = i - o * (ci * 9);  // OIHW: r = c*9 + tap
    sw[r * co + o] = w[i];
  }
  for (int i = threadIdx.x; i < co; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  // thread = (4 consecutive pixels of a row, 8 output channels): every weight vector read from shared memory is
  // used for 4 pixels and every input value for up to 3 taps (the smem weight reads bound the 1-pixel version)
  constexpr int PX = 4;
  const int groups = co / 8;
  const int wq = W / PX;
  const int total = n_img * H * wq * groups;  // < 2^31 for any batch that fits the GPU
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int gsel = i % groups;
    int p = i / groups;
    const int x0 = (p % wq) * PX; p /= wq;
    const int yy = p % H;
    const int n = p / H;
    float acc[PX][8];
#pragma unroll
    for (int q = 0; q < PX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = sb[gsel * 8 + j];
    for (int c = 0; c < ci; ++c) {
      const float* xp = x + ((size_t)n * ci + c) * H * W;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int y2 = yy + dy - 1;
        if (y2 < 0 || y2 >= H) continue;
        float v[PX + 2];
#pragma unroll
        for (int k = 0; k < PX + 2; ++k) {
          const int x2 = x0 + k - 1;
          v[k] = (x2 >= 0 && x2 < W) ? __ldg(xp + (size_t)y2 * W + x2) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wr = sw + (c * 9 + dy * 3 + dx) * co + gsel * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(v[q + dx], wv[j], acc[q][j]);
        }
      }
    }
    const size_t pix = ((size_t)n * H + yy) * W + x0;
#pragma unroll
    for (int q = 0; q < PX; ++q)
      *reinterpret_cast<uint4*>(out + ((pix + q) * groups + gsel) * 8) =
          make_uint4(pack_bf16(acc[q][0], acc[q][1]), pack_bf16(acc[q][2], acc[q][3]), pack_bf16(acc[q][4], acc[q][5]),
                     pack_bf16(acc[q][6], acc[q][7]));
  }
}


// The same convolution as an implicit GEMM on warp-level TF32 MMA (m16n8k8): M = pixels, N = 128, K = 9 ci (27 / 36,
// padded to 32 / 40).  x and w are rounded to TF32 (2^-11; the bf16 output rounds at 2^-9), accumulation in fp32.  The CUDA-core
// kernel above needs 3 456 FMAs per pixel and was FMA-bound at a quarter of the HBM rate of its output.
// DGRAD = 1 turns it into the DATA GRADIENT of the tail conv (128 -> co): da[p][c] = sum_{o,tap} dy[p - off(tap)][o] w[o][c][tap]
// is the same contraction over (o, tap') with tap' = 8 - tap, no bias.
// A persistent CTA walks over 256-pixel tiles; the fp32 halo tile (ci x (TH + 2) x (W + 2)) is staged in shared memory as TF32 and
// every A fragment element is one shared-memory load at (k's tap offset) + (pixel offset); a warp owns 64 of the 128 output
// channels (its B fragments stay in registers) and every fourth m-tile; the bf16 rows leave through a per-warp staging tile
// as 16-byte stores.
template <int CI, int W, int DGRAD>
__global__ void __launch_bounds__(256, 2)
in_conv_tf32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    bf16* __restrict__ out, int n_img, int H, int n_tiles) {
  constexpr int TH = 256 / W, HWP = W + 2, HR = TH + 2, K = CI * 9, KC = (K + 7) / 8, CO = 128;
  constexpr int XS = CI * HR * HWP;
  __shared__ uint32_t xs[XS];
  __shared__ __align__(16) uint32_t so[8][16][36];  // [warp][pixel][32 words of 2 bf16 + 4 pad]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int half = warp & 1, mq = warp >> 1;
  auto tf32 = [](float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
  };
  // B[k = c * 9 + tap][n]: forward w[n][c][tap] (OIHW); data gradient w[o = c][n][8 - tap]
  uint32_t bfr[KC][8][2];
  int aoff[KC][2];
#pragma unroll
  for (int kc = 0; kc < KC; ++kc)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int k = kc * 8 + t + 4 * h2;
      const bool valid = k < K;
      const int c = valid ? k / 9 : 0, tap = valid ? k - c * 9 : 0;
      aoff[kc][h2] = c * HR * HWP + (tap / 3) * HWP + (tap % 3);  // k >= K: any finite value, its B row is zero
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = half * 64 + nt * 8 + g;
        float v = 0.f;
        if (valid) v = DGRAD ? __ldg(w + ((size_t)c * CO + n) * 9 + 8 - tap) : __ldg(w + ((size_t)n * CI + c) * 9 + tap);
        bfr[kc][nt][h2] = tf32(v);
      }
    }
  __shared__ __align__(8) float sb[CO];
  if (tid < CO) sb[tid] = DGRAD ? 0.f : bias[tid];
  const int tiles_per_img = H * W / 256;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_img;
    const int y0 = (tile - n * tiles_per_img) * TH;
    __syncthreads();  // the previous tile's fragments have been read
    for (int i = tid; i < XS; i += 256) {
      const int c = i / (HR * HWP), r = i - c * (HR * HWP);
      const int hy = r / HWP, hx = r - hy * HWP;
      const int yy = y0 + hy - 1, xx = hx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      xs[i] = ok ? tf32(__ldg(x + (((size_t)n * CI + c) * H + yy) * W + xx)) : 0u;
    }
    __syncthreads();
#pragma unroll 1
    for (int mi = 0; mi < 4; ++mi) {
      const int mt = mq + 4 * mi;
      const int p0 = mt * 16 + g, p1 = p0 + 8;
      const int pb0 = (p0 / W) * HWP + (p0 % W), pb1 = (p1 / W) * HWP + (p1 % W);
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 b2 = *reinterpret_cast<const float2*>(&sb[half * 64 + nt * 8 + 2 * t]);
        acc[nt][0] = acc[nt][2] = b2.x;
        acc[nt][1] = acc[nt][3] = b2.y;
      }
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const uint32_t a0 = xs[aoff[kc][0] + pb0], a1 = xs[aoff[kc][0] + pb1];
        const uint32_t a2 = xs[aoff[kc][1] + pb0], a3 = xs[aoff[kc][1] + pb1];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfr[kc][nt][0]), "r"(bfr[kc][nt][1]));
      }
      __syncwarp();  // the staging tile's previous rows have been read
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        so[warp][g][nt * 4 + t] = pack_bf16(acc[nt][0], acc[nt][1]);
        so[warp][g + 8][nt * 4 + t] = pack_bf16(acc[nt][2], acc[nt][3]);
      }
      __syncwarp();
      bf16* dst = out + ((size_t)tile * 256 + mt * 16) * CO + half * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int id = lane + 32 * i, row = id >> 3, c16 = id & 7;
        *reinterpret_cast<uint4*>(dst + (size_t)row * CO + c16 * 8) = *reinterpret_cast<const uint4*>(&so[warp][row][c16 * 4]);
      }
    }
  }
}
static bool in_conv_tf32_ok(int ci, int H, int W, int co) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TSD_IN_CONV_TF32");  // 0: the CUDA-core kernels (A/B)
    on = e ? atoi(e) : 1;
  }
  return on && co == 128 && (ci == 3 || ci == 4) && (W == 16 || W == 32 || W == 64) && (H * W) % 256 == 0 && H % (256 / W) == 0;
}
template <int CI, int DGRAD>
static int launch_in_conv_tf32(cudaStream_t st, const float* x, const float* w, const float* bias, bf16* out, int n_img,
                               int H, int W) {
  const int n_tiles = n_img * (H * W / 256);
  const int grid = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
  if (W == 64) in_conv_tf32_kernel<CI, 64, DGRAD><<<grid, 256, 0, st>>>(x, w, bias, out, n_img, H, n_tiles);
  else if (W == 32) in_conv_tf32_kernel<CI, 32, DGRAD><<<grid, 256, 0, st>>>(x, w, bias, out, n_img, H, n_tiles);
  else in_conv_tf32_kernel<CI, 16, DGRAD><<<grid, 256, 0, st>>>(x, w, bias, out, n_img, H, n_tiles);
  TSD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ tail conv
// Data gradient of the tail conv: da[p][c] = sum_tap sum_co dy[p - off(tap)][co] * w[co][c][tap], register-blocked:
// thread = (4 consecutive pixels of a row, 8 channels).  Every weight vector read from shared memory serves 4 pixels
// and every dy value up to 3 taps.  W % 4 == 0.
template <int CO>
__global__ void __launch_bounds__(256) tail_conv_dgrad4_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                               bf16* __restrict__ da, int n_img, int H, int W) {
  constexpr int C = 128, PX = 4;
  __shared__ float sw[CO * 9 * C];  // [co][tap][c]
  for (int i = threadIdx.x; i < CO * C * 9; i += blockDim.x) {
    const int o = i / (C * 9), r = i - o * (C * 9), c = r / 9, tap = r - c * 9;
    sw[(o * 9 + tap) * C + c] = w[i];
  }
  __syncthreads();
  const int wq = W / PX;
  const int total = n_img * H * wq * (C / 8);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (i % (C / 8)) * 8;
    int p = i / (C / 8);
    const int x0 = (p % wq) * PX; p /= wq;
    const int yy = p % H;
    const size_t n = p / H;
    float acc[PX][8];
#pragma unroll
    for (int q = 0; q < PX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      const float* dp = dy + (n * CO + o) * (size_t)H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        // forward: out[q] += a[q + off(tap)] * w[tap]  =>  da[p] += dy[p - off(tap)] * w[tap]
        const int y2 = yy - (ky - 1);
        if (y2 < 0 || y2 >= H) continue;
        float v[PX + 2];
#pragma unroll
        for (int k = 0; k < PX + 2; ++k) {
          const int x2 = x0 + k - 1;
          v[k] = (x2 >= 0 && x2 < W) ? __ldg(dp + (size_t)y2 * W + x2) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wr = sw + (o * 9 + ky * 3 + kx) * C + cv;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int q = 0; q < PX; ++q) {
            const float d = v[q + 2 - kx];  // dy at x = x0 + q - (kx - 1)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(d, wv[j], acc[q][j]);
          }
        }
      }
    }
    bf16* dst = da + (((size_t)n * H + yy) * W + x0) * C + cv;
#pragma unroll
    for (int q = 0; q < PX; ++q)
      *reinterpret_cast<uint4*>(dst + (size_t)q * C) =
          make_uint4(pack_bf16(acc[q][0], acc[q][1]), pack_bf16(acc[q][2], acc[q][3]), pack_bf16(acc[q][4], acc[q][5]),
                     pack_bf16(acc[q][6], acc[q][7]));
  }
}


// ------------------------------------------------------------------------------------------ DDPM elementwise
// x_t = sqrt_ab[t_n] * x0 + sqrt_1mab[t_n] * noise (utils.py:115-116); noise ~ N(0,1) from Philox unless given.
__global__ void q_sample_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t,
                                const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab,
                                const float* __restrict__ noise_in, uint64_t seed, uint64_t offset,
                                float* __restrict__ x_t, float* __restrict__ noise_out, size_t per_sample, size_t total4,
                                const uint64_t* __restrict__ rng_dev) {
  const Philox rng(seed);
  // stream position: (call counter, global sample index, element) -- see RngPos
  const RngPos pos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = 0x71ULL | (pos.calls << 8);
  const uint64_t base = offset + pos.sample0 * (uint64_t)(per_sample / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const int n = (int)(e / per_sample);
    const int64_t tn = t[n];
    const float a = sqrt_ab[tn], b = sqrt_1mab[tn];
    const float4 x = *reinterpret_cast<const float4*>(x0 + e);
    float4 z;
    if (noise_in) {
      z = *reinterpret_cast<const float4*>(noise_in + e);
    } else {
      const uint4 r = rng(base + i, ctr_hi);
      const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
      z = make_float4(z0.x, z0.y, z1.x, z1.y);
    }
    // same operation order as the reference (two products, one sum), no FMA contraction
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(b, z.x));
    o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(b, z.y));
    o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(b, z.z));
    o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(b, z.w));
    *reinterpret_cast<float4*>(x_t + e) = o;
    if (noise_out) *reinterpret_cast<float4*>(noise_out + e) = z;
  }
}

__global__ void mse_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ noise, float* __restrict__ loss,
                               size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - noise[i];
    loss[i] = d * d;
  }
}
__global__ void mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ noise,
                               const float* __restrict__ gout, float* __restrict__ dpred, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    dpred[i] = 2.f * (pred[i] - noise[i]) * gout[i];
}

// One reverse-diffusion update (utils.py:149-167):
//   eps = (1+w) eps_c - w eps_u;  mean = c1[t] x - c2[t] eps;  x' = mean + sigma[t] z  (z = 0 at t = 0)
// eps holds [2B] images: the conditional batch first, then the unconditional batch.  The step
// index is read from device memory so the same launch can be replayed from a CUDA graph.
__global__ void sampler_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const int* __restrict__ step_ptr,
                                      const float* __restrict__ c1, const float* __restrict__ c2,
                                      const float* __restrict__ sigma, float w, const float* __restrict__ noise_in,
                                      uint64_t seed, float* __restrict__ x_out, int* __restrict__ nan_flag, size_t total4,
                                      int clip_last, int dup, const uint64_t* __restrict__ rng_dev, int n_img) {
  const int step = *step_ptr;
  const float k1 = c1[step], k2 = c2[step], sg = sigma[step];
  const float w1 = 1.f + w;
  const Philox rng(seed);
  const RngPos pos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = ((uint64_t)step + 1) | (pos.calls << 16);  // fresh stream per step and per sampler call
  const uint64_t base = pos.sample0 * (uint64_t)(total4 / (size_t)(n_img > 0 ? n_img : 1));
  bool bad = false;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 4;
    const float4 xv = *reinterpret_cast<const float4*>(x + e);
    const float4 ec = *reinterpret_cast<const float4*>(eps + e);
    const float4 eu = *reinterpret_cast<const float4*>(eps + total4 * 4 + e);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (step > 0) {
      if (noise_in) {
        z = *reinterpret_cast<const float4*>(noise_in + e);
      } else {
        const uint4 r = rng(base + i, ctr_hi);
        const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
        z = make_float4(z0.x, z0.y, z1.x, z1.y);
      }
    }
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, cs[4] = {ec.x, ec.y, ec.z, ec.w}, us[4] = {eu.x, eu.y, eu.z, eu.w},
                zs[4] = {z.x, z.y, z.z, z.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ep = __fsub_rn(__fmul_rn(w1, cs[j]), __fmul_rn(w, us[j]));
      const float mean = __fsub_rn(__fmul_rn(k1, xs[j]), __fmul_rn(k2, ep));
      float v = __fadd_rn(mean, __fmul_rn(sg, zs[j]));
      bad |= (v != v);
      if (clip_last && step == 0) v = fminf(fmaxf(v, -1.f), 1.f);
      o[j] = v;
    }
    *reinterpret_cast<float4*>(x_out + e) = make_float4(o[0], o[1], o[2], o[3]);
    if (dup) *reinterpret_cast<float4*>(x_out + total4 * 4 + e) = make_float4(o[0], o[1], o[2], o[3]);
  }
  if (bad) atomicOr(nan_flag, 1);
}


// ------------------------------------------------------------------------------------------ tail conv on tensor cores
// The same convolution as an implicit GEMM on warp-level MMA (m16n8k16, bf16 in, fp32 accumulate): M = 128 output
// pixels per CTA (TH = 128 / W rows of one image), N = 8 (CO <= 4 used), K = 9 taps x 128 channels.  The (TH + 2) x
// (W + 2) halo tile of the bf16 NHWC input is staged once in shared memory (cp.async, zero fill = the padding; 272-byte
// pixel pitch => conflict-free ldmatrix) and every tap is the same tile read at a shifted address; the weights are
// converted to bf16 [n][tap * 128 + c] in shared memory by each CTA (14 KB of L2 reads).  SAMPLE = 1 runs the
// conditional and the unconditional image of the pair through the same staging buffer and applies the classifier-free
// guidance + posterior update + Philox noise + NaN flag per pixel and channel exactly like tail_conv_sample_kernel.
// The input is read ~2x (halo rows) instead of 9x per output, and the 3 456 MACs per pixel run on the tensor pipe.
namespace tcv {
constexpr int C = 128, PITCH = 272, WK = 9 * C, WPITCH = WK + 8;  // bf16 elements per weight row (+16 B pad)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace tcv

template <int CO, int SAMPLE>
__global__ void __launch_bounds__(256, 2)
tail_conv_mma_kernel(const bf16* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ out /* SAMPLE: x (updated in place, both halves); else eps */,
                     const int* __restrict__ step_ptr, const float* __restrict__ c1, const float* __restrict__ c2,
                     const float* __restrict__ sigma, float wcfg, const float* __restrict__ noise_in, uint64_t seed,
                     int* __restrict__ nan_flag, float* __restrict__ eps_out, int B, int H, int W, int clip_last,
                     const uint64_t* __restrict__ rng_dev) {
  using namespace tcv;
  extern __shared__ __align__(16) uint8_t smem_tc[];
  const int TH = 128 / W;                 // output rows per CTA
  const int HWP = W + 2;                  // halo row length in pixels
  bf16* sW = reinterpret_cast<bf16*>(smem_tc);                         // [8][WPITCH]
  uint8_t* sA = smem_tc + 8 * WPITCH * 2;                              // [(TH + 2)][W + 2][PITCH bytes]
  const uint32_t sA_u = smem_addr(sA), sW_u = smem_addr(sW);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int HW = H * W;
  const int tiles_per_img = HW / 128;
  const int n = blockIdx.x / tiles_per_img;          // image index inside the first half (SAMPLE) / the batch
  const int y0 = (blockIdx.x - n * tiles_per_img) * TH;
  // ---- weights: fp32 OIHW [CO][128][3][3] -> bf16 [n][tap * 128 + c], rows >= CO zero
  for (int i = threadIdx.x; i < CO * WK; i += blockDim.x) {  // coalesced read of OIHW, scattered 2-byte smem writes
    const int o = i / WK, r = i - o * WK, c = r / 9, tap = r - c * 9;
    sW[o * WPITCH + tap * C + c] = __float2bfloat16(__ldg(w + i));
  }
  for (int i = threadIdx.x; i < (8 - CO) * WK; i += blockDim.x) sW[(CO + i / WK) * WPITCH + i % WK] = __float2bfloat16(0.f);
  float acc[SAMPLE ? 2 : 1][4];
#pragma unroll
  for (int hf = 0; hf < (SAMPLE ? 2 : 1); ++hf) {
    acc[hf][0] = acc[hf][1] = acc[hf][2] = acc[hf][3] = 0.f;
    if (hf == 1) __syncthreads();  // everyone is done reading the first image's tile
    // ---- stage the halo tile of image n (+ B for the unconditional half)
    const bf16* img = a + (size_t)(n + hf * B) * HW * C;
    const int chunks = (TH + 2) * HWP * 16;  // 16-byte chunks
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
      const int ck = i & 15, px = i >> 4;
      const int hy = px / HWP, hx = px - hy * HWP;
      const int yy = y0 + hy - 1, xx = hx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      cp16(sA_u + px * PITCH + ck * 16, img + ((size_t)(ok ? yy : 0) * W + (ok ? xx : 0)) * C + ck * 8, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- this warp's 16 pixels: tile pixel p = warp * 16 + row, (ty, tx) = (p / W, p % W)
    const int prow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix row of this lane
    const int ty = prow / W, tx = prow - ty * W;
    const uint32_t a_lane = sA_u + (ty * HWP + tx) * PITCH + (lane >> 4) * 16;
    const uint32_t b_lane = sW_u + (g * WPITCH + 2 * t) * 2;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;  // halo coordinates already include the -1 shift
      const uint32_t a_tap = a_lane + (dy * HWP + dx) * PITCH;
      const uint32_t b_tap = b_lane + tap * C * 2;
#pragma unroll
      for (int kc = 0; kc < C / 16; ++kc) {
        uint32_t af[4];
        ldsm4(af, a_tap + kc * 32);
        uint32_t b0, b1;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b0) : "r"(b_tap + kc * 32));
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b1) : "r"(b_tap + kc * 32 + 16));
        mma(acc[hf], af, b0, b1);
      }
    }
  }
  // ---- epilogue: lane (g, t) holds pixels g, g + 8 of the warp's 16 and channels 2t, 2t + 1
  int step = 0;
  float k1 = 0.f, k2 = 0.f, sg = 0.f;
  if (SAMPLE) {
    step = *step_ptr;
    k1 = c1[step]; k2 = c2[step]; sg = sigma[step];
  }
  const float w1 = 1.f + wcfg;
  const Philox rng(seed);
  const RngPos pos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = ((uint64_t)step + 1) | (pos.calls << 16);
  const uint64_t ebase = pos.sample0 * (uint64_t)(CO * HW);  // z depends on the GLOBAL sample index
  bool bad = false;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int p = warp * 16 + g + 8 * r;                 // tile pixel
    const int rem = y0 * W + p;                          // pixel index inside the image (tile rows are contiguous)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int o = 2 * t + cc;
      if (o >= CO) continue;
      const size_t e = ((size_t)n * CO + o) * HW + rem;
      const float ec = acc[0][2 * r + cc] + bias[o];
      if (!SAMPLE) {
        out[e] = ec;
      } else {
        const float eu = acc[SAMPLE ? 1 : 0][2 * r + cc] + bias[o];
        if (eps_out) {
          eps_out[e] = ec;
          eps_out[(size_t)B * CO * HW + e] = eu;
        }
        float z = 0.f;
        if (step > 0) {
          if (noise_in) {
            z = noise_in[e];
          } else {
            const uint4 rr = rng(ebase + e, ctr_hi);
            z = box_muller(rr.x, rr.y).x;
          }
        }
        const float ep = __fsub_rn(__fmul_rn(w1, ec), __fmul_rn(wcfg, eu));
        const float mean = __fsub_rn(__fmul_rn(k1, out[e]), __fmul_rn(k2, ep));
        float v = __fadd_rn(mean, __fmul_rn(sg, z));
        bad |= (v != v);
        if (clip_last && step == 0) v = fminf(fmaxf(v, -1.f), 1.f);
        out[e] = v;
        out[(size_t)B * CO * HW + e] = v;  // the unconditional copy of the 2B batch
      }
    }
  }
  if (SAMPLE && bad) atomicOr(nan_flag, 1);
}

// ------------------------------------------------------------------------------------------ tail conv, per-pixel form
// The same convolution (and the same fused sampling epilogue) with every input pixel read from shared memory ONCE:
//   Y[p][tap * CO + o] = sum_c a[p][c] * w[o][c][tap]        one [pixels x 128] x [128 x 9 CO] product per halo pixel
//   eps[q][o]          = bias[o] + sum_tap Y[q + off(tap)][tap * CO + o]
// N = 27 (36) columns fill four (five) n8 tiles instead of 3 of 8 per tap, the weights live in registers as B
// fragments for the CTA's life, and the nine taps no longer re-read the activation tile with ldmatrix (the kernel
// above reads 9 x 256 B per pixel from shared memory and was bound by that, not by HBM).  Persistent CTAs walk over
// 256-pixel tiles (TH = 256 / W rows + halo); the next tile's activations are in flight while the current one is
// multiplied, summed and written: two TMA boxes [TH + 2][W + 2][64 channels] per tile (out-of-image pixels arrive as
// zeros, 128-byte swizzle = the XOR pattern ldmatrix wants).  A first version staged the tile with cp.async: its
// 25 copies per thread queued in front of the ldmatrix reads in the LSU and the load time added to the compute time
// (TSD_TAIL_Y_DBG experiments: 60 us of loads + 42 us of compute = 102 us).  Y overlays the activation buffer it was
// computed from.
// V = 0: 256 threads, two activation buffers, one CTA per SM.  V = 1: 128 threads and ONE buffer, two CTAs per SM: the
// phases of a tile (tensor pipe, shared-memory traffic, Philox / update arithmetic, global latency) are serial inside a
// CTA, two independent CTAs overlap them.
template <int CO, int SAMPLE, int W, int V>
__global__ void __launch_bounds__(V ? 128 : 256, V ? 2 : 1)
tail_conv_y_kernel(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, const int* __restrict__ step_ptr, const float* __restrict__ c1,
                   const float* __restrict__ c2, const float* __restrict__ sigma, float wcfg,
                   const float* __restrict__ noise_in, uint64_t seed, int* __restrict__ nan_flag,
                   float* __restrict__ eps_out, int B, int H, int clip_last, const uint64_t* __restrict__ rng_dev,
                   int n_tiles, int dbg) {
  using namespace tcv;
  constexpr int NTHR = V ? 128 : 256, NW = NTHR / 32, NBUF = V ? 1 : 2, PPT = 256 / NTHR;
  constexpr int TH = 256 / W, HWP = W + 2, NP = (TH + 2) * HWP, MT = (NP + 15) / 16, MTW = (MT + NW - 1) / NW;
  constexpr int NJ = 9 * CO, NT = (NJ + 7) / 8, YP = NJ | 1;  // odd pitch: conflict-free gathers
  constexpr int HALFB = MT * 16 * 128;                        // one 64-channel half of a tile: [pixel][128 B], swizzled
  constexpr int BUF = 2 * HALFB;                              // bytes per activation buffer
  constexpr int NH = SAMPLE ? 2 : 1;
  static_assert(NP * YP * 4 <= BUF, "Y must fit the buffer it overlays");
  static_assert(HALFB % 1024 == 0, "swizzled TMA destinations need 1024-byte alignment");
  extern __shared__ uint8_t smem_ty_raw[];
  const uint32_t s0 = (smem_addr(smem_ty_raw) + 1023u) & ~1023u;
  uint8_t* smem_ty = smem_ty_raw + (s0 - smem_addr(smem_ty_raw));
  const uint32_t bar0 = s0 + NBUF * BUF;  // "tile landed" mbarriers, one per buffer
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int HW = H * W, tiles_per_img = HW / 256;
  // ---- weights as B fragments: B[k = c][j = tap * CO + o] = w[o][c][tap] (OIHW fp32 -> bf16), zero for j >= 9 CO
  uint32_t bfr[8][NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int j = nt * 8 + g;
    const float* wj = w + (size_t)(j % CO) * C * 9 + j / CO;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      const int k0 = kc * 16 + 2 * t;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (j < NJ) {
        v0 = __ldg(wj + (k0) * 9);
        v1 = __ldg(wj + (k0 + 1) * 9);
        v2 = __ldg(wj + (k0 + 8) * 9);
        v3 = __ldg(wj + (k0 + 9) * 9);
      }
      bfr[kc][nt][0] = pack_bf16(v0, v1);
      bfr[kc][nt][1] = pack_bf16(v2, v3);
    }
  }
  float bv[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) bv[o] = bias[o];
  const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_items = my_tiles * NH;
  auto issue = [&](int it, int buf) {  // one thread
    const int tile = blockIdx.x + (it / NH) * gridDim.x, hf = it % NH;
    const int n = tile / tiles_per_img;
    const int y0 = (tile - n * tiles_per_img) * TH;
    const uint32_t bar = bar0 + 8u * buf, dst = s0 + buf * BUF;
    if (dbg & 8) {
      mbar_arrive(bar);
      return;
    }
    fence_proxy_async_smem();  // the buffer was last touched through the generic proxy (ldmatrix, Y)
    mbar_arrive_expect_tx(bar, 2 * NP * 128);
    tma_load_4d(dst, &tmA, bar, 0, -1, y0 - 1, n + hf * B);
    tma_load_4d(dst + HALFB, &tmA, bar, 64, -1, y0 - 1, n + hf * B);
  };
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    mbar_init(bar0, 1);
    if (NBUF == 2) mbar_init(bar0 + 8u, 1);
    fence_mbar_init();
  }
  __syncthreads();
  int step = 0;
  float k1 = 0.f, k2 = 0.f, sg = 0.f;
  if (SAMPLE) {
    step = *step_ptr;
    k1 = c1[step]; k2 = c2[step]; sg = sigma[step];
  }
  const float w1 = 1.f + wcfg;
  const Philox rng(seed);
  const RngPos pos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = ((uint64_t)step + 1) | (pos.calls << 16);
  const uint64_t ebase = pos.sample0 * (uint64_t)(CO * HW);  // z depends on the GLOBAL sample index
  bool bad = false;
  float ec[PPT][CO];
#pragma unroll
  for (int r = 0; r < PPT; ++r)
#pragma unroll
    for (int o = 0; o < CO; ++o) ec[r][o] = 0.f;

  if (n_items > 0 && tid == 0) issue(0, 0);
  for (int it = 0; it < n_items; ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    if (NBUF == 2) {
      __syncthreads();  // nobody reads the other buffer (the previous tile's Y) any more
      if (it + 1 < n_items && tid == 0) issue(it + 1, buf ^ 1);
    }
    mbar_wait(bar0 + 8u * buf, NBUF == 2 ? ((it >> 1) & 1) : (it & 1));  // tile `it` has landed
    const uint32_t base = s0 + buf * BUF;
    // ---- Y for this warp's m-tiles (16 halo pixels each)
    float acc[MTW][NT][4];
#pragma unroll
    for (int m = 0; m < MTW; ++m) {
      const int mt = warp + NW * m;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
      if (mt < MT && !(dbg & 1)) {
        const int px = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix row of this lane
        const uint32_t row = base + px * 128;
        const uint32_t sw = static_cast<uint32_t>(px & 7), hi = static_cast<uint32_t>(lane >> 4);
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          uint32_t af[4];
          ldsm4(af, row + (kc >> 2) * HALFB + (((2 * (kc & 3) + hi) ^ sw) << 4));
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma(acc[m][nt], af, bfr[kc][nt][0], bfr[kc][nt][1]);
        }
      }
    }
    __syncthreads();  // every warp is done with the activations: Y may overwrite them
    float* Y = reinterpret_cast<float*>(smem_ty + (size_t)buf * BUF);
#pragma unroll
    for (int m = 0; m < MTW; ++m) {
      const int mt = warp + NW * m;
      if (mt < MT && !(dbg & 2)) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int px = mt * 16 + g + 8 * r;
          if (px < NP) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const int j = nt * 8 + 2 * t;
              if (j < NJ) Y[px * YP + j] = acc[m][nt][2 * r];
              if (j + 1 < NJ) Y[px * YP + j + 1] = acc[m][nt][2 * r + 1];
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- thread = PPT output pixels: nine neighbours x CO columns each
    const int tile = blockIdx.x + (it / NH) * gridDim.x, hf = it % NH;
    const int n = tile / tiles_per_img;
    const int rem0 = (tile - n * tiles_per_img) * 256;  // pixel index inside the image (tile rows are contiguous)
    float ev[PPT][CO];
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      const int q = tid + NTHR * r;
      const int ty = q / W, tx = q - ty * W;
      const int hp0 = ty * HWP + tx;
#pragma unroll
      for (int o = 0; o < CO; ++o) ev[r][o] = 0.f;
      if (!(dbg & 2))
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float* yp = Y + (hp0 + (tap / 3) * HWP + (tap % 3)) * YP + tap * CO;
#pragma unroll
        for (int o = 0; o < CO; ++o) ev[r][o] += yp[o];
      }
#pragma unroll
      for (int o = 0; o < CO; ++o) ev[r][o] += bv[o];
    }
    if (NBUF == 1) {
      __syncthreads();  // Y is dead: the next tile may land while this one is written out
      if (it + 1 < n_items && tid == 0) issue(it + 1, 0);
    }
    if (!SAMPLE) {
#pragma unroll
      for (int r = 0; r < PPT; ++r)
#pragma unroll
        for (int o = 0; o < CO; ++o) out[((size_t)n * CO + o) * HW + rem0 + tid + NTHR * r] = ev[r][o];
    } else if (hf == 0) {
#pragma unroll
      for (int r = 0; r < PPT; ++r)
#pragma unroll
        for (int o = 0; o < CO; ++o) ec[r][o] = ev[r][o];
    } else if (!(dbg & 4)) {
#pragma unroll
      for (int r = 0; r < PPT; ++r)
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const size_t e = ((size_t)n * CO + o) * HW + rem0 + tid + NTHR * r;
        const float eu = ev[r][o];
        if (eps_out) {
          eps_out[e] = ec[r][o];
          eps_out[(size_t)B * CO * HW + e] = eu;
        }
        float z = 0.f;
        if (step > 0) {
          if (noise_in) {
            z = noise_in[e];
          } else {
            const uint4 rr = rng(ebase + e, ctr_hi);
            z = box_muller(rr.x, rr.y).x;
          }
        }
        const float ep = __fsub_rn(__fmul_rn(w1, ec[r][o]), __fmul_rn(wcfg, eu));
        const float mean = __fsub_rn(__fmul_rn(k1, out[e]), __fmul_rn(k2, ep));
        float v = __fadd_rn(mean, __fmul_rn(sg, z));
        bad |= (v != v);
        if (clip_last && step == 0) v = fminf(fmaxf(v, -1.f), 1.f);
        out[e] = v;
        out[(size_t)B * CO * HW + e] = v;  // the unconditional copy of the 2B batch
      }
    }
  }
  if (SAMPLE && bad) atomicOr(nan_flag, 1);
}

static int tail_y_variant() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TSD_TAIL_Y");  // 0: the tap-by-tap kernel above; 1: one 256-thread CTA per SM, two buffers; 2: two 128-thread CTAs
    on = e ? atoi(e) : 2;
  }
  return on;
}
static bool tail_y_ok(int H, int W) {
  return tail_y_variant() && (W == 16 || W == 32 || W == 64) && (H * W) % 256 == 0 && H % (256 / W) == 0;
}
template <int CO, int SAMPLE, int W, int V>
static int launch_tail_y_v(cudaStream_t st, const bf16* a, const float* w, const float* bias, float* out,
                           const int* step_ptr, const float* c1, const float* c2, const float* sigma, float wcfg,
                           const float* noise_in, uint64_t seed, int* nan_flag, float* eps_out, int B, int H, int clip_last,
                           const uint64_t* rng_dev) {
  constexpr int NP = (256 / W + 2) * (W + 2), MT = (NP + 15) / 16;
  constexpr int smem = (V ? 1 : 2) * MT * 16 * 256 + 1024 + 64;
  CUtensorMap tmA;
  if (make_tmap_nhwc(&tmA, a, (uint64_t)(SAMPLE ? 2 * B : B), H, W, 128, 64, W + 2, 256 / W + 2, 1, 1)) return 1;
  static tsd::PerDeviceFlag cfgd;
  if (!cfgd.cur()) {
    TSD_CUDA(cudaFuncSetAttribute(tail_conv_y_kernel<CO, SAMPLE, W, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cfgd.cur() = true;
  }
  const int n_tiles = B * (H * W / 256);
  const int slots = (V ? 2 : 1) * num_sms();
  const int grid = n_tiles < slots ? n_tiles : slots;
  static int dbg = -1;  // TSD_TAIL_Y_DBG bit mask, timing experiments with WRONG results: 1 no MMAs, 2 no Y / gather, 4 no update, 8 no loads
  if (dbg < 0) {
    const char* e = getenv("TSD_TAIL_Y_DBG");
    dbg = e ? atoi(e) : 0;
  }
  tail_conv_y_kernel<CO, SAMPLE, W, V><<<grid, V ? 128 : 256, smem, st>>>(tmA, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in,
                                                                          seed, nan_flag, eps_out, B, H, clip_last, rng_dev, n_tiles, dbg);
  TSD_LAUNCH_CHECK();
  return 0;
}
template <int CO, int SAMPLE, int W>
static int launch_tail_y(cudaStream_t st, const bf16* a, const float* w, const float* bias, float* out,
                         const int* step_ptr, const float* c1, const float* c2, const float* sigma, float wcfg,
                         const float* noise_in, uint64_t seed, int* nan_flag, float* eps_out, int B, int H, int clip_last,
                         const uint64_t* rng_dev) {
  if (tail_y_variant() == 1)
    return launch_tail_y_v<CO, SAMPLE, W, 0>(st, a, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
  return launch_tail_y_v<CO, SAMPLE, W, 1>(st, a, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
}
template <int CO, int SAMPLE>
static int launch_tail_y_w(cudaStream_t st, int W, const bf16* a, const float* w, const float* bias, float* out,
                           const int* step_ptr, const float* c1, const float* c2, const float* sigma, float wcfg,
                           const float* noise_in, uint64_t seed, int* nan_flag, float* eps_out, int B, int H,
                           int clip_last, const uint64_t* rng_dev) {
  if (W == 64)
    return launch_tail_y<CO, SAMPLE, 64>(st, a, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
  if (W == 32)
    return launch_tail_y<CO, SAMPLE, 32>(st, a, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
  return launch_tail_y<CO, SAMPLE, 16>(st, a, w, bias, out, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
}

// the tail conv tiles an image into 128-pixel row blocks
static bool tail_mma_ok(int H, int W) {
  return (W == 16 || W == 32 || W == 64) && (H * W) % 128 == 0 && H % (128 / W) == 0;
}
static int tail_mma_smem(int W) { return 8 * tcv::WPITCH * 2 + (128 / W + 2) * (W + 2) * tcv::PITCH; }

// im2col of the fp32 NCHW image for the head-conv weight gradient: patch[p][k], k = c*9 + tap (< ci*9), zero-padded
// to KP columns, bf16 -> the reduction over all pixels runs on the tensor cores (tsd_gemm_wgrad).
__global__ void im2col_head_kernel(const float* __restrict__ x, bf16* __restrict__ patch, int n_img, int ci, int H, int W,
                                   int KP) {
  const int vec = KP / 8;
  const int total = n_img * H * W * vec;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kv = (i % vec) * 8;
    int p = i / vec;
    const int xx = p % W; p /= W;
    const int yy = p % H;
    const int n = p / H;
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kv + j;
      float v = 0.f;
      if (k < ci * 9) {
        const int c = k / 9, tap = k - c * 9;
        const int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) v = __ldg(x + (((size_t)n * ci + c) * H + y2) * W + x2);
      }
      e[j] = v;
    }
    *reinterpret_cast<uint4*>(patch + (size_t)i * 8) =
        make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
  }
}
// fp32 NCHW [n][co][H][W] -> bf16 NHWC padded to CP channels (zeros beyond co): operand of the tail-conv weight gradient
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int n_img, int co, int HW,
                                        int CP) {
  const int vec = CP / 8;
  const int total = n_img * HW * vec;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = (i % vec) * 8;
    const int p = i / vec;
    const int n = p / HW, r = p - n * HW;
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = (cv + j) < co ? __ldg(src + ((size_t)n * co + cv + j) * HW + r) : 0.f;
    *reinterpret_cast<uint4*>(dst + (size_t)i * 8) =
        make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
  }
}

__global__ void step_counter_kernel(int* step_ptr, int delta) { *step_ptr += delta; }
__global__ void counter_add_u64_kernel(uint64_t* p, uint64_t delta) { *p += delta; }
// t[n] ~ U{0..T-1} (utils.py:112), a function of (seed, call counter, GLOBAL sample index) only
__global__ void draw_timesteps_kernel(int64_t* __restrict__ t, int n, int T, uint64_t seed,
                                      const uint64_t* __restrict__ rng_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Philox rng(seed);
  const RngPos pos = load_rng_pos(rng_dev);
  const uint4 r = rng(pos.sample0 + (uint64_t)i, 0x72ULL | (pos.calls << 8));
  t[i] = (int64_t)(((uint64_t)r.x * (uint64_t)T) >> 32);
}
// out[0:len] = table[*step][0:len]  (per-step conditioning rows, indexed on the device so a CUDA graph can replay)
__global__ void gather_row_kernel(const float* __restrict__ table, const int* __restrict__ step_ptr, int len,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = table[(size_t)(*step_ptr) * len + i];
}

inline int ew_grid(size_t items) {
  size_t g = (items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int tsd_head_conv_fwd(void* stream, const float* x, const float* w, const float* bias, void* out, int n_img,
                                 int ci, int H, int W, int co) {
  TSD_CHECK(ci <= MAX_CI && co % 8 == 0 && W % 4 == 0, "head_conv_fwd: unsupported shape ci=%d co=%d W=%d", ci, co, W);
  if (in_conv_tf32_ok(ci, H, W, co)) {
    if (ci == 3) return launch_in_conv_tf32<3, 0>((cudaStream_t)stream, x, w, bias, (bf16*)out, n_img, H, W);
    return launch_in_conv_tf32<4, 0>((cudaStream_t)stream, x, w, bias, (bf16*)out, n_img, H, W);
  }
  const size_t smem = (size_t)(ci * 9 * co + co) * sizeof(float);
  const size_t total = (size_t)n_img * H * (W / 4) * (co / 8);
  head_conv_fwd_kernel<<<ew_grid(total), 256, smem, (cudaStream_t)stream>>>(x, w, bias, (bf16*)out, n_img, ci, H, W, co);
  TSD_LAUNCH_CHECK();
  return 0;
}
#define TSD_TAIL_SMEM_ATTR(KERN)                                                                              \
  do {                                                                                                        \
    static tsd::PerDeviceFlag cfgd;                                                                           \
    if (!cfgd.cur()) {                                                                                        \
      TSD_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));          \
      cfgd.cur() = true;                                                                                      \
    }                                                                                                         \
  } while (0)
extern "C" int tsd_tail_conv_fwd(void* stream, const void* a, const float* w, const float* bias, float* out, int n_img,
                                 int H, int W, int c_in, int co) {
  TSD_CHECK(c_in == 128 && (co == 3 || co == 4), "tail_conv_fwd: unsupported channels c_in=%d co=%d", c_in, co);
  TSD_CHECK(tail_mma_ok(H, W), "tail_conv_fwd: unsupported image size %dx%d (W in {16,32,64}, H*W %% 128 == 0)", H, W);
  const int smem = tail_mma_smem(W);
  const int grid_tc = n_img * (H * W / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (tail_y_ok(H, W)) {
    if (co == 3)
      return launch_tail_y_w<3, 0>(st, W, (const bf16*)a, w, bias, out, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, n_img, H, 0, nullptr);
    return launch_tail_y_w<4, 0>(st, W, (const bf16*)a, w, bias, out, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, n_img, H, 0, nullptr);
  }
  if (co == 3) {
    TSD_TAIL_SMEM_ATTR((tail_conv_mma_kernel<3, 0>));
    tail_conv_mma_kernel<3, 0><<<grid_tc, 256, smem, st>>>((const bf16*)a, w, bias, out, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, 0, H, W, 0, nullptr);
  } else {
    TSD_TAIL_SMEM_ATTR((tail_conv_mma_kernel<4, 0>));
    tail_conv_mma_kernel<4, 0><<<grid_tc, 256, smem, st>>>((const bf16*)a, w, bias, out, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, 0, H, W, 0, nullptr);
  }
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_q_sample(void* stream, const float* x0, const int64_t* t, const float* sqrt_ab, const float* sqrt_1mab,
                            const float* noise_in, uint64_t seed, uint64_t offset, float* x_t, float* noise_out,
                            int n_img, int64_t per_sample, const uint64_t* rng_dev) {
  TSD_CHECK(per_sample % 4 == 0, "q_sample: per-sample element count must be a multiple of 4");
  const size_t total4 = (size_t)n_img * per_sample / 4;
  q_sample_kernel<<<ew_grid(total4), 256, 0, (cudaStream_t)stream>>>(x0, t, sqrt_ab, sqrt_1mab, noise_in, seed, offset, x_t, noise_out, per_sample, total4, rng_dev);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_draw_timesteps(void* stream, int64_t* t, int n, int T, uint64_t seed, const uint64_t* rng_dev) {
  TSD_CHECK(n > 0 && T > 0, "draw_timesteps: n=%d T=%d", n, T);
  draw_timesteps_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(t, n, T, seed, rng_dev);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_counter_add_u64(void* stream, uint64_t* counter, uint64_t delta) {
  counter_add_u64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_mse_fwd(void* stream, const float* pred, const float* noise, float* loss, int64_t total) {
  mse_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(pred, noise, loss, total);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_mse_bwd(void* stream, const float* pred, const float* noise, const float* gout, float* dpred,
                           int64_t total) {
  mse_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(pred, noise, gout, dpred, total);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_sampler_update(void* stream, const float* x, const float* eps, const int* step_ptr, const float* c1,
                                  const float* c2, const float* sigma, float w, const float* noise_in, uint64_t seed,
                                  float* x_out, int* nan_flag, int64_t total, int clip_last, int dup,
                                  const uint64_t* rng_dev, int n_img) {
  TSD_CHECK(total % 4 == 0, "sampler_update: element count must be a multiple of 4");
  TSD_CHECK(rng_dev == nullptr || (n_img > 0 && (total / 4) % n_img == 0), "sampler_update: n_img=%d does not divide the batch", n_img);
  sampler_update_kernel<<<ew_grid(total / 4), 256, 0, (cudaStream_t)stream>>>(x, eps, step_ptr, c1, c2, sigma, w, noise_in, seed, x_out, nan_flag, total / 4, clip_last, dup, rng_dev, n_img);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_step_add(void* stream, int* step_ptr, int delta) {
  step_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr, delta);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_gather_row_f32(void* stream, const float* table, const int* step_ptr, int len, float* out) {
  gather_row_kernel<<<ceil_div(len, 256), 256, 0, (cudaStream_t)stream>>>(table, step_ptr, len, out);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_tail_conv_sample(void* stream, const void* a, const float* w, const float* bias, float* x,
                                    const int* step_ptr, const float* c1, const float* c2, const float* sigma, float wcfg,
                                    const float* noise_in, uint64_t seed, int* nan_flag, float* eps_out, int B, int H, int W,
                                    int c_in, int co, int clip_last, const uint64_t* rng_dev) {
  TSD_CHECK(c_in == 128 && (co == 3 || co == 4), "tail_conv_sample: unsupported channels c_in=%d co=%d", c_in, co);
  TSD_CHECK(tail_mma_ok(H, W), "tail_conv_sample: unsupported image size %dx%d (W in {16,32,64}, H*W %% 128 == 0)", H, W);
  // tensor-core implicit GEMM, both halves of the CFG pair per CTA
  const int smem = tail_mma_smem(W);
  const int grid_tc = B * (H * W / 128);
  cudaStream_t stt = (cudaStream_t)stream;
  if (tail_y_ok(H, W)) {
    if (co == 3)
      return launch_tail_y_w<3, 1>(stt, W, (const bf16*)a, w, bias, x, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
    return launch_tail_y_w<4, 1>(stt, W, (const bf16*)a, w, bias, x, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, clip_last, rng_dev);
  }
  if (co == 3) {
    TSD_TAIL_SMEM_ATTR((tail_conv_mma_kernel<3, 1>));
    tail_conv_mma_kernel<3, 1><<<grid_tc, 256, smem, stt>>>((const bf16*)a, w, bias, x, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, W, clip_last, rng_dev);
  } else {
    TSD_TAIL_SMEM_ATTR((tail_conv_mma_kernel<4, 1>));
    tail_conv_mma_kernel<4, 1><<<grid_tc, 256, smem, stt>>>((const bf16*)a, w, bias, x, step_ptr, c1, c2, sigma, wcfg, noise_in, seed, nan_flag, eps_out, B, H, W, clip_last, rng_dev);
  }
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_im2col_head(void* stream, const float* x, void* patch, int n_img, int ci, int H, int W, int KP) {
  TSD_CHECK(KP % 8 == 0 && ci * 9 <= KP, "im2col_head: KP=%d too small for ci=%d", KP, ci);
  im2col_head_kernel<<<ew_grid((size_t)n_img * H * W * (KP / 8)), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)patch, n_img, ci, H, W, KP);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_nchw_to_nhwc_pad(void* stream, const float* src, void* dst, int n_img, int co, int HW, int CP) {
  TSD_CHECK(CP % 8 == 0 && co <= CP, "nchw_to_nhwc_pad: CP=%d too small for co=%d", CP, co);
  nchw_to_nhwc_pad_kernel<<<ew_grid((size_t)n_img * HW * (CP / 8)), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n_img, co, HW, CP);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_tail_conv_dgrad(void* stream, const float* dy, const float* w, void* da, int n_img, int H, int W,
                                   int c_in, int co) {
  TSD_CHECK(c_in == 128 && (co == 3 || co == 4), "tail_conv_dgrad: unsupported channels c_in=%d co=%d", c_in, co);
  TSD_CHECK(W % 4 == 0, "tail_conv_dgrad: W=%d must be a multiple of 4", W);
  if (in_conv_tf32_ok(co, H, W, c_in)) {
    if (co == 3) return launch_in_conv_tf32<3, 1>((cudaStream_t)stream, dy, w, nullptr, (bf16*)da, n_img, H, W);
    return launch_in_conv_tf32<4, 1>((cudaStream_t)stream, dy, w, nullptr, (bf16*)da, n_img, H, W);
  }
  const size_t total = (size_t)n_img * H * W;
  if (co == 3) tail_conv_dgrad4_kernel<3><<<ew_grid(total * 4), 256, 0, (cudaStream_t)stream>>>(dy, w, (bf16*)da, n_img, H, W);
  else tail_conv_dgrad4_kernel<4><<<ew_grid(total * 4), 256, 0, (cudaStream_t)stream>>>(dy, w, (bf16*)da, n_img, H, W);
  TSD_LAUNCH_CHECK();
  return 0;
}
