// GroupNorm(+SiLU,+dropout) and LayerNorm, forward and backward, on bf16 channels-last tensors.
// These are the HBM-bound kernels of the path: 16-byte vector accesses, fp32 statistics,
// warp-shuffle reductions, one atomic per (warp, statistic).
//
// Reference: nn.GroupNorm(32, C) + nn.SiLU (+ nn.Dropout) at diffusion.py:90-91, 95-97, 122, 258-259;
// nn.LayerNorm(C) at diffusion.py:127, 130, 132.
#include "../../include/tinysd_b200.h"
#include "common.cuh"

using namespace tsd;

namespace {

constexpr int GROUPS = 32;

// two warp reductions interleaved (independent shuffles in flight together)
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a += ta;
    b += tb;
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm statistics.  grid = (chunks, n_img); each CTA reduces a slab of pixels of one image
// over all channels of the (optionally concatenated) input and adds per-group partial sums.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x0, const bf16* __restrict__ x1,
                                                         int c0, int c1, int hw, int pix_per_cta,
                                                         float* __restrict__ part /* [n][chunks][32][2] */) {
  const int C = c0 + c1;
  const int cpg = C / GROUPS;
  const int vec_per_pix = C / 8;
  const int n = blockIdx.y;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  // blockDim (256) is a multiple of vec_per_pix (C/8 in {8,16,32,64,128,256}), so every thread owns a
  // fixed 8-channel slot: accumulate in registers over its pixels, then reduce in a FIXED order (no atomics:
  // the statistics, hence the whole forward pass, are bit-reproducible run to run).
  const int slots = blockDim.x / vec_per_pix;
  const int cv = (threadIdx.x % vec_per_pix) * 8;
  const int my_slot = threadIdx.x / vec_per_pix;
  float ls[4] = {0.f, 0.f, 0.f, 0.f}, lq[4] = {0.f, 0.f, 0.f, 0.f};
  const bool from0 = cv < c0;
  const bf16* src_base = from0 ? x0 + (size_t)n * hw * c0 + cv : x1 + (size_t)n * hw * c1 + (cv - c0);
  const int src_ld = from0 ? c0 : c1;
  constexpr int UNROLL = 8;  // independent 16-byte loads in flight per thread
  for (int pix0 = p_begin + my_slot; pix0 < p_end; pix0 += slots * UNROLL) {
    uint4 u[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      u[k] = pix < p_end ? *reinterpret_cast<const uint4*>(src_base + (size_t)pix * src_ld) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
      const float e[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
      if (cpg >= 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { ls[0] += e[j]; lq[0] += e[j] * e[j]; }
      } else if (cpg == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { ls[j >> 2] += e[j]; lq[j >> 2] += e[j] * e[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { ls[j >> 1] += e[j]; lq[j >> 1] += e[j] * e[j]; }
      }
    }
  }
  __shared__ float s_ps[256][4], s_pq[256][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { s_ps[threadIdx.x][k] = ls[k]; s_pq[threadIdx.x][k] = lq[k]; }
  __syncthreads();
  if (threadIdx.x < GROUPS) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    if (cpg >= 8) {  // the group spans cpg/8 whole vectors, partial index 0
      const int v0 = g * cpg / 8, nv = cpg / 8;
      for (int sl = 0; sl < slots; ++sl)
        for (int v = 0; v < nv; ++v) { s += s_ps[sl * vec_per_pix + v0 + v][0]; q += s_pq[sl * vec_per_pix + v0 + v][0]; }
    } else {         // the group is the k-th sub-block of one vector
      const int v0 = g * cpg / 8, k = (g * cpg % 8) / cpg;
      for (int sl = 0; sl < slots; ++sl) { s += s_ps[sl * vec_per_pix + v0][k]; q += s_pq[sl * vec_per_pix + v0][k]; }
    }
    float* dst = part + (((size_t)n * gridDim.x + blockIdx.x) * GROUPS + g) * 2;
    dst[0] = s;
    dst[1] = q;
  }
}

// per-CTA partials (fixed order over chunks) -> (mean, rstd)
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int count, int chunks,
                                   float inv_cnt, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // n * GROUPS + g
  if (i >= count) return;
  const int n = i / GROUPS, g = i - n * GROUPS;
  float s = 0.f, q = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float* src = part + (((size_t)n * chunks + c) * GROUPS + g) * 2;
    s += src[0];
    q += src[1];
  }
  const float mean = s * inv_cnt;
  const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
  stats[2 * i] = mean;
  stats[2 * i + 1] = rsqrtf(var + eps);
}

// (mean, rstd) from the per-channel partials that the producing GEMM / convolution left behind (tsd_*_fwd_gn):
// part_k [n_img * hw / 32][c_k][2] = (sum, sum of squares) of every 32-row quarter-tile.  One warp per (image, group):
// lanes walk the (half-tile, channel) pairs in a fixed order, then a shuffle tree -- reproducible run to run.
__global__ void __launch_bounds__(256) gn_finalize_parts_kernel(const float* __restrict__ part0, const float* __restrict__ part1,
                                                                int c0, int c1, int n_img, int halves, float inv_cnt,
                                                                float eps, float* __restrict__ stats) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_img * GROUPS) return;
  const int n = wid / GROUPS, g = wid - n * GROUPS;
  const int C = c0 + c1, cpg = C / GROUPS;
  const int ch0 = g * cpg;
  const bool from0 = ch0 < c0;  // a group never straddles the two sources (c0 is a multiple of the group width)
  const float* part = from0 ? part0 : part1;
  const int cs = from0 ? c0 : c1, cb = from0 ? ch0 : ch0 - c0;
  float s = 0.f, q = 0.f;
  for (int i = lane; i < halves * cpg; i += 32) {
    const int h = i / cpg, c = i - h * cpg;
    const float2 v = *reinterpret_cast<const float2*>(part + ((size_t)(n * halves + h) * cs + cb + c) * 2);
    s += v.x;
    q += v.y;
  }
  warp_sum2(s, q);
  if (lane == 0) {
    const float mean = s * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    stats[2 * wid] = mean;
    stats[2 * wid + 1] = rsqrtf(var + eps);
  }
}

// Keep/scale factors for the 8 consecutive elements starting at flat index ebase (ebase % 8 == 0): one Philox call,
// one 16-bit uniform per element (keep probability quantised to 1/65536).  The backward kernels regenerate the mask.
__device__ __forceinline__ void dropout_scales8(const Philox& rng, uint64_t ebase, uint64_t ctr_hi, float p,
                                                float inv_keep, float* sc) {
  const uint4 r = rng.rounds<7>(ebase >> 3, ctr_hi);
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[2 * j] = (w[j] & 0xFFFFu) >= thr ? inv_keep : 0.f;
    sc[2 * j + 1] = (w[j] >> 16) >= thr ? inv_keep : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm apply: out = dropout(silu(gamma * (x - mean) * rstd + beta)).
// grid = (chunks, n_img); thread = one 8-channel vector.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x0, const bf16* __restrict__ x1, int c0,
                                                       int c1, int hw, int pix_per_cta, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int act_silu, float drop_p, uint64_t seed,
                                                       bf16* __restrict__ out, const uint64_t* __restrict__ rng_dev) {
  const int C = c0 + c1;
  const int cpg = C / GROUPS;
  const int vec_per_pix = C / 8;
  const int n = blockIdx.y;
  // per-channel affine of this image folded with the statistics: z = x * a[c] + b[c]
  extern __shared__ float s_ab[];  // a[C] then b[C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = stats[((size_t)n * GROUPS + g) * 2], rstd = stats[((size_t)n * GROUPS + g) * 2 + 1];
    const float a = rstd * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = beta[c] - mean * a;
  }
  __syncthreads();
  // blockDim is a multiple of vec_per_pix: each thread owns one 8-channel slot, keeps its affine in registers
  const int slots = blockDim.x / vec_per_pix;
  const int cv = (threadIdx.x % vec_per_pix) * 8;
  const int my_slot = threadIdx.x / vec_per_pix;
  float fa[8], fb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { fa[j] = s_ab[cv + j]; fb[j] = s_ab[C + cv + j]; }
  const Philox rng(seed);
  // mask = f(seed, call counter, GLOBAL sample index, element): identical for any sharding of the batch (RngPos)
  const RngPos rpos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = 0x5eedULL | (rpos.calls << 16);
  const size_t ng = (size_t)n + (size_t)rpos.sample0;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  const bool from0 = cv < c0;
  const bf16* src_base = from0 ? x0 + (size_t)n * hw * c0 + cv : x1 + (size_t)n * hw * c1 + (cv - c0);
  const int src_ld = from0 ? c0 : c1;
  bf16* dst_base = out + (size_t)n * hw * C + cv;
  constexpr int UNROLL = 4;
  for (int pix0 = p_begin + my_slot; pix0 < p_end; pix0 += slots * UNROLL) {
    uint4 u[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix < p_end) u[k] = *reinterpret_cast<const uint4*>(src_base + (size_t)pix * src_ld);
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix >= p_end) break;
      float e[8];
      { const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = d.x; e[7] = d.y; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = fmaf(e[j], fa[j], fb[j]);
        if (act_silu) z = silu_f(z);
        e[j] = z;
      }
      if (drop_p > 0.f) {
        float sc[8];
        dropout_scales8(rng, (ng * hw + pix) * C + cv, ctr_hi, drop_p, inv_keep, sc);
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] *= sc[j];
      }
      *reinterpret_cast<uint4*>(dst_base + (size_t)pix * C) =
          make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm backward, pass 1: per-(image, channel) sums  A = sum dz * xhat,  B = sum dz
// where dz = dy * dropout_scale * silu'(z).  grid = (chunks, n_img).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) gn_bwd_sums_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x0,
                                                          const bf16* __restrict__ x1, int c0, int c1, int hw,
                                                          int pix_per_cta, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int act_silu, float drop_p, uint64_t seed,
                                                          float* __restrict__ ab /* [n][C][2] */,
                                                          const uint64_t* __restrict__ rng_dev) {
  const int C = c0 + c1;
  const int cpg = C / GROUPS;
  const int vec_per_pix = C / 8;
  const int n = blockIdx.y;
  extern __shared__ float s_ab[];  // [C][2]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_ab[i] = 0.f;
  __syncthreads();
  const Philox rng(seed);
  // mask = f(seed, call counter, GLOBAL sample index, element): identical for any sharding of the batch (RngPos)
  const RngPos rpos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = 0x5eedULL | (rpos.calls << 16);
  const size_t ng = (size_t)n + (size_t)rpos.sample0;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  // thread -> fixed 8-channel slot (blockDim % vec_per_pix == 0): all per-channel constants live in registers
  const int slots = blockDim.x / vec_per_pix;
  const int cv = (threadIdx.x % vec_per_pix) * 8;
  const int my_slot = threadIdx.x / vec_per_pix;
  // cpg >= 4 here (C >= 128): an 8-channel vector touches at most two groups -> per-group constants in 2-entry arrays
  float fa[8], fb[8], rs[2], ms[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int g = (cv + 4 * h2) / cpg;
    const float mean = stats[((size_t)n * GROUPS + g) * 2], rstd = stats[((size_t)n * GROUPS + g) * 2 + 1];
    rs[h2] = rstd;
    ms[h2] = -mean * rstd;
#pragma unroll
    for (int j = 4 * h2; j < 4 * h2 + 4; ++j) {
      fa[j] = rstd * gamma[cv + j];
      fb[j] = beta[cv + j] - mean * fa[j];
    }
  }
  float accA[8], accB[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { accA[j] = 0.f; accB[j] = 0.f; }
  const bool from0 = cv < c0;
  const bf16* src_base = from0 ? x0 + (size_t)n * hw * c0 + cv : x1 + (size_t)n * hw * c1 + (cv - c0);
  const int src_ld = from0 ? c0 : c1;
  const bf16* dy_base = dy + (size_t)n * hw * C + cv;
  constexpr int UNROLL = 4;
  for (int pix0 = p_begin + my_slot; pix0 < p_end; pix0 += slots * UNROLL) {
    uint4 u[UNROLL], gq[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix < p_end) {
        u[k] = *reinterpret_cast<const uint4*>(src_base + (size_t)pix * src_ld);
        gq[k] = *reinterpret_cast<const uint4*>(dy_base + (size_t)pix * C);
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix >= p_end) break;
      float e[8], d[8];
      { const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), dd = unpack_bf16(u[k].w);
        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = dd.x; e[7] = dd.y; }
      { const float2 a = unpack_bf16(gq[k].x), b = unpack_bf16(gq[k].y), c = unpack_bf16(gq[k].z), dd = unpack_bf16(gq[k].w);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y; d[4] = c.x; d[5] = c.y; d[6] = dd.x; d[7] = dd.y; }
      if (drop_p > 0.f) {
        float sc[8];
        dropout_scales8(rng, (ng * hw + pix) * C + cv, ctr_hi, drop_p, inv_keep, sc);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= sc[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = d[j];
        if (act_silu) dz *= silu_grad_f(fmaf(e[j], fa[j], fb[j]));
        accA[j] = fmaf(dz, fmaf(e[j], rs[j >> 2], ms[j >> 2]), accA[j]);
        accB[j] += dz;
      }
    }
  }
  // s_ab is laid out [j][A|B][8-channel vector]: consecutive lanes hit consecutive banks (the [C][2] layout put the 32
  // lanes of a warp on two banks: ncu counted a 7.7-way conflict on these atomics)
  const int vv = threadIdx.x % vec_per_pix;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_ab[(2 * j) * vec_per_pix + vv], accA[j]);
    atomicAdd(&s_ab[(2 * j + 1) * vec_per_pix + vv], accB[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int c = i >> 1, h = i & 1;
    atomicAdd(&ab[(size_t)n * 2 * C + i], s_ab[(2 * (c & 7) + h) * vec_per_pix + (c >> 3)]);
  }
}

// pass 2: dx = rstd * (gamma*dz - mean_g(gamma*dz) - xhat * mean_g(gamma*dz*xhat)) (+ radd), written to the
// (optionally split) destinations dx0 [.., c0] and dx1 [.., c1].
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x0,
                                                           const bf16* __restrict__ x1, int c0, int c1, int hw,
                                                           int pix_per_cta, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int act_silu, float drop_p, uint64_t seed,
                                                           const float* __restrict__ ab, const bf16* __restrict__ radd,
                                                           bf16* __restrict__ dx0, bf16* __restrict__ dx1,
                                                           const uint64_t* __restrict__ rng_dev,
                                                           float* __restrict__ cs_out, float* __restrict__ cs_total) {
  const int C = c0 + c1;
  const int cpg = C / GROUPS;
  const int vec_per_pix = C / 8;
  const int n = blockIdx.y;
  __shared__ float s_m1[GROUPS], s_m2[GROUPS];
  if (threadIdx.x < GROUPS) {
    const int g = threadIdx.x;
    float m1 = 0.f, m2 = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const int c = g * cpg + j;
      m2 += gamma[c] * ab[((size_t)n * C + c) * 2];      // sum gamma*dz*xhat
      m1 += gamma[c] * ab[((size_t)n * C + c) * 2 + 1];  // sum gamma*dz
    }
    const float inv = 1.f / (float)(cpg * hw);
    s_m1[g] = m1 * inv;
    s_m2[g] = m2 * inv;
  }
  __syncthreads();
  const Philox rng(seed);
  // mask = f(seed, call counter, GLOBAL sample index, element): identical for any sharding of the batch (RngPos)
  const RngPos rpos = load_rng_pos(rng_dev);
  const uint64_t ctr_hi = 0x5eedULL | (rpos.calls << 16);
  const size_t ng = (size_t)n + (size_t)rpos.sample0;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int slots = blockDim.x / vec_per_pix;
  const int cv = (threadIdx.x % vec_per_pix) * 8;
  const int my_slot = threadIdx.x / vec_per_pix;
  // dx = g1*dz - (k1 + xhat*k2), z = x*fa + fb, xhat = x*rs + ms
  float fa[8], fb[8], rs[2], ms[2], k1[2], k2[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int g = (cv + 4 * h2) / cpg;
    const float mean = stats[((size_t)n * GROUPS + g) * 2], rstd = stats[((size_t)n * GROUPS + g) * 2 + 1];
    rs[h2] = rstd;
    ms[h2] = -mean * rstd;
    k1[h2] = rstd * s_m1[g];
    k2[h2] = rstd * s_m2[g];
#pragma unroll
    for (int j = 4 * h2; j < 4 * h2 + 4; ++j) {
      fa[j] = rstd * gamma[cv + j];
      fb[j] = beta[cv + j] - mean * fa[j];
    }
  }
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  const bool from0 = cv < c0;
  const bf16* src_base = from0 ? x0 + (size_t)n * hw * c0 + cv : x1 + (size_t)n * hw * c1 + (cv - c0);
  bf16* dst_base = from0 ? dx0 + (size_t)n * hw * c0 + cv : dx1 + (size_t)n * hw * c1 + (cv - c0);
  const int src_ld = from0 ? c0 : c1;
  const size_t full_base = (size_t)n * hw * C + cv;
  // optional by-product: column sums of dx per image (cs_out [n][C]) and over all images (cs_total [C]) -- the
  // time-embedding / bias gradients that a separate pass over dx would otherwise compute (diffusion.py:112-113)
  extern __shared__ float s_cs[];  // [C], only when cs_out != null
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cs_out) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.f;
    __syncthreads();
  }
  constexpr int UNROLL = 4;
  for (int pix0 = p_begin + my_slot; pix0 < p_end; pix0 += slots * UNROLL) {
    uint4 u[UNROLL], gq[UNROLL], rq[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix < p_end) {
        u[k] = *reinterpret_cast<const uint4*>(src_base + (size_t)pix * src_ld);
        gq[k] = *reinterpret_cast<const uint4*>(dy + full_base + (size_t)pix * C);
        rq[k] = radd ? *reinterpret_cast<const uint4*>(radd + full_base + (size_t)pix * C) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int pix = pix0 + k * slots;
      if (pix >= p_end) break;
      float e[8], d[8], r[8];
      { const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), dd = unpack_bf16(u[k].w);
        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y; e[6] = dd.x; e[7] = dd.y; }
      { const float2 a = unpack_bf16(gq[k].x), b = unpack_bf16(gq[k].y), c = unpack_bf16(gq[k].z), dd = unpack_bf16(gq[k].w);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y; d[4] = c.x; d[5] = c.y; d[6] = dd.x; d[7] = dd.y; }
      { const float2 a = unpack_bf16(rq[k].x), b = unpack_bf16(rq[k].y), c = unpack_bf16(rq[k].z), dd = unpack_bf16(rq[k].w);
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = c.x; r[5] = c.y; r[6] = dd.x; r[7] = dd.y; }
      if (drop_p > 0.f) {
        float sc[8];
        dropout_scales8(rng, (ng * hw + pix) * C + cv, ctr_hi, drop_p, inv_keep, sc);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= sc[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = d[j];
        if (act_silu) dz *= silu_grad_f(fmaf(e[j], fa[j], fb[j]));
        const float xhat = fmaf(e[j], rs[j >> 2], ms[j >> 2]);
        e[j] = fmaf(fa[j], dz, r[j]) - fmaf(xhat, k2[j >> 2], k1[j >> 2]);
        cs[j] += e[j];
      }
      *reinterpret_cast<uint4*>(dst_base + (size_t)pix * src_ld) =
          make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
    }
  }
  if (cs_out) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_cs[cv + j], cs[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&cs_out[(size_t)n * C + i], s_cs[i]);
      if (cs_total) atomicAdd(&cs_total[i], s_cs[i]);
    }
  }
}

// dgamma[c] += sum_n A[n][c];  dbeta[c] += sum_n B[n][c].  CTA = 8 channels x 32 image lanes, all loads of a lane in
// flight at once (a single thread per channel walking all images serially made this 30 us of pure load latency, 39
// times per step; 32 channels x 8 lanes still 10 us).
__global__ void __launch_bounds__(256) gn_bwd_params_kernel(const float* __restrict__ ab, int n_img, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int cx = threadIdx.x & 7, ny = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cx;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int n0 = ny; n0 < n_img; n0 += 32 * 8) {
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int n = n0 + 32 * k;
        v[k] = n < n_img ? *reinterpret_cast<const float2*>(ab + ((size_t)n * C + c) * 2) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { a += v[k].x; b += v[k].y; }
    }
  }
  __shared__ float sa[32][8], sb[32][8];
  sa[ny][cx] = a;
  sb[ny][cx] = b;
  __syncthreads();
  if (ny == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 32; ++k) { a += sa[k][cx]; b += sb[k][cx]; }  // fixed order: reproducible
    dgamma[c] += a;
    dbeta[c] += b;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over C (128 / 256 / 512): one warp per token row.
// ------------------------------------------------------------------------------------------
template <int VEC>  // C = 128 * VEC: each lane handles VEC chunks of 4 channels (8 bytes)
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, int M, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, bf16* __restrict__ out) {
  constexpr int C = 128 * VEC;
  constexpr int RPW = 8;  // rows per warp per pass: RPW independent row loads in flight
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  float4 g[VEC], bt[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
    bt[i] = *reinterpret_cast<const float4*>(beta + (i * 32 + lane) * 4);
  }
  for (int row0 = warp_global * RPW; row0 < M; row0 += nwarps * RPW) {
    uint2 u[RPW][VEC];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        u[r][i] = (row0 + r < M) ? *reinterpret_cast<const uint2*>(x + (size_t)(row0 + r) * C + (i * 32 + lane) * 4)
                                 : make_uint2(0, 0);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row0 + r >= M) break;
      float v[4 * VEC];
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float2 a = unpack_bf16(u[r][i].x), b = unpack_bf16(u[r][i].y);
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b.x; v[4 * i + 3] = b.y;
        s += a.x + a.y + b.x + b.y;
        q += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
      }
      // sum and sum of squares reduced side by side: one dependent shuffle chain per row instead of two (the kernel is
      // bound by that latency, not by HBM); fp32 E[x^2] - mean^2 over C <= 512 bf16 values, as in the GroupNorm kernels
      warp_sum2(s, q);
      const float mean = s * (1.f / C);
      const float rstd = rsqrtf(fmaxf(q * (1.f / C) - mean * mean, 0.f) + eps);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float o0 = (v[4 * i] - mean) * rstd * g[i].x + bt[i].x, o1 = (v[4 * i + 1] - mean) * rstd * g[i].y + bt[i].y;
        const float o2 = (v[4 * i + 2] - mean) * rstd * g[i].z + bt[i].z, o3 = (v[4 * i + 3] - mean) * rstd * g[i].w + bt[i].w;
        *reinterpret_cast<uint2*>(out + (size_t)(row0 + r) * C + (i * 32 + lane) * 4) =
            make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
      }
    }
  }
}

// LayerNorm backward: dx = rstd * (g*dy - mean(g*dy) - xhat*mean(g*dy*xhat)) (+ radd); dgamma/dbeta are
// accumulated per CTA in shared memory then atomically added (fp32).
template <int VEC>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, int M,
                                                     const float* __restrict__ gamma, float eps,
                                                     const bf16* __restrict__ radd, bf16* __restrict__ dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     int rows_per_cta, int rows_per_sample, float* __restrict__ cs_out,
                                                     float* __restrict__ cs_total) {
  constexpr int C = 128 * VEC;
  __shared__ float s_dg[C], s_db[C], s_cs[C];
  // optional by-product (cs_out / cs_total != null): column sums of dx per sample and over all rows -- the gradients
  // of the bias / per-sample vector added in front of this LayerNorm (diffusion.py:141-148); rows_per_cta then divides
  // rows_per_sample, so a CTA's rows belong to one sample
  const bool want_cs = cs_out != nullptr || cs_total != nullptr;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; s_cs[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float adg[4 * VEC], adb[4 * VEC], gm[4 * VEC], acs[4 * VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
    gm[4 * i] = g.x; gm[4 * i + 1] = g.y; gm[4 * i + 2] = g.z; gm[4 * i + 3] = g.w;
  }
#pragma unroll
  for (int i = 0; i < 4 * VEC; ++i) { adg[i] = 0.f; adb[i] = 0.f; acs[i] = 0.f; }
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  constexpr int RPW = VEC == 1 ? 4 : 2;  // rows per warp per pass: all loads of these rows are issued before the first is used
  for (int row0 = r_begin + warp * RPW; row0 < r_end; row0 += nwarps * RPW) {
    uint2 ux[RPW][VEC], ug[RPW][VEC], ur[RPW][VEC];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int row = row0 + r < r_end ? row0 + r : row0;  // a duplicate load for the ragged last pass
        const size_t off = (size_t)row * C + (i * 32 + lane) * 4;
        ux[r][i] = *reinterpret_cast<const uint2*>(x + off);
        ug[r][i] = *reinterpret_cast<const uint2*>(dy + off);
        if (radd) ur[r][i] = *reinterpret_cast<const uint2*>(radd + off);
      }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = row0 + r;
      if (row >= r_end) break;
      float v[4 * VEC], d[4 * VEC];
      float s = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float2 a = unpack_bf16(ux[r][i].x), b = unpack_bf16(ux[r][i].y), c = unpack_bf16(ug[r][i].x),
                     e = unpack_bf16(ug[r][i].y);
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = b.x; v[4 * i + 3] = b.y;
        d[4 * i] = c.x; d[4 * i + 1] = c.y; d[4 * i + 2] = e.x; d[4 * i + 3] = e.y;
        s += a.x + a.y + b.x + b.y;
        sq += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
      }
      warp_sum2(s, sq);
      const float mean = s * (1.f / C);
      const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + eps);
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4 * VEC; ++i) {
        v[i] = (v[i] - mean) * rstd;  // xhat
        adg[i] += d[i] * v[i];
        adb[i] += d[i];
        d[i] *= gm[i];
        m1 += d[i];
        m2 += d[i] * v[i];
      }
      warp_sum2(m1, m2);
      m1 *= (1.f / C);
      m2 *= (1.f / C);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const size_t off = (size_t)row * C + (i * 32 + lane) * 4;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rstd * (d[4 * i + j] - m1 - v[4 * i + j] * m2);
        if (radd) {
          const float2 a = unpack_bf16(ur[r][i].x), b = unpack_bf16(ur[r][i].y);
          o[0] += a.x; o[1] += a.y; o[2] += b.x; o[3] += b.y;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acs[4 * i + j] += o[j];
        *reinterpret_cast<uint2*>(dx + off) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_dg[(i * 32 + lane) * 4 + j], adg[4 * i + j]);
      atomicAdd(&s_db[(i * 32 + lane) * 4 + j], adb[4 * i + j]);
      if (want_cs) atomicAdd(&s_cs[(i * 32 + lane) * 4 + j], acs[4 * i + j]);
    }
  __syncthreads();
  const int sample = want_cs ? r_begin / rows_per_sample : 0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_dg[i]);
    atomicAdd(&dbeta[i], s_db[i]);
    if (cs_out) atomicAdd(&cs_out[(size_t)sample * C + i], s_cs[i]);
    if (cs_total) atomicAdd(&cs_total[i], s_cs[i]);
  }
}

int gn_grid_x(int hw, int C, int n_img, int* pix_per_cta) {
  // enough CTAs to fill the machine ~4x, at least 32 pixels per CTA
  int want = ceil_div(8 * num_sms(), n_img);
  if (want < 1) want = 1;
  int ppc = ceil_div(hw, want);
  const int min_pix = ceil_div(16384, C);  // >= 16K elements per CTA
  if (ppc < min_pix) ppc = min_pix;
  if (ppc > hw) ppc = hw;
  *pix_per_cta = ppc;
  return ceil_div(hw, ppc);
}

}  // namespace

extern "C" int64_t tsd_gn_scratch_floats(int n_img) { return ((int64_t)8 * num_sms() + n_img) * 64 + 64; }

extern "C" int tsd_gn_stats(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int hw,
                            float eps, float* scratch, float* stats) {
  const int C = c0 + c1;
  TSD_CHECK(C % 64 == 0 && c0 % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "gn_stats: unsupported channels c0=%d c1=%d", c0, c1);
  int ppc;
  const int gx = gn_grid_x(hw, C, n_img, &ppc);
  gn_partial_kernel<<<dim3(gx, n_img), 256, 0, (cudaStream_t)stream>>>((const bf16*)x0, (const bf16*)x1, c0, c1, hw, ppc,
                                                                      scratch);
  TSD_LAUNCH_CHECK();
  const int count = n_img * GROUPS;
  gn_finalize_kernel<<<ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(scratch, stats, count, gx,
                                                                            1.f / ((float)hw * (C / GROUPS)), eps);
  TSD_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsd_gn_stats_from_parts(void* stream, const float* part0, const float* part1, int c0, int c1, int n_img,
                                       int hw, float eps, float* stats) {
  const int C = c0 + c1;
  TSD_CHECK(C % GROUPS == 0 && hw % 64 == 0 && (c1 == 0 || c0 % (C / GROUPS) == 0) && (c1 == 0 || part1 != nullptr),
            "gn_stats_from_parts: unsupported shape c0=%d c1=%d hw=%d", c0, c1, hw);
  const int warps = n_img * GROUPS;
  gn_finalize_parts_kernel<<<ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      part0, part1, c0, c1, n_img, hw / 32, 1.f / ((float)hw * (C / GROUPS)), eps, stats);
  TSD_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsd_gn_apply(void* stream, const void* x0, const void* x1, int c0, int c1, int n_img, int hw,
                            const float* stats, const float* gamma, const float* beta, int act_silu, float drop_p,
                            uint64_t seed, void* out, const uint64_t* rng_dev) {
  const int C = c0 + c1;
  TSD_CHECK(C % 64 == 0 && c0 % 8 == 0, "gn_apply: unsupported channels c0=%d c1=%d", c0, c1);
  TSD_CHECK(C <= 2048 && 256 % (C / 8) == 0, "gn_apply: unsupported channel count %d", C);
  int ppc;
  const int gx = gn_grid_x(hw, C, n_img, &ppc);
  gn_apply_kernel<<<dim3(gx, n_img), 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)x0, (const bf16*)x1, c0, c1, hw, ppc, stats, gamma, beta, act_silu, drop_p, seed, (bf16*)out, rng_dev);
  TSD_LAUNCH_CHECK();
  return 0;
}

// ab: scratch [n_img][C][2] fp32, must be zero on entry (the call leaves it dirty; zero it with tsd_zero_f32).
extern "C" int tsd_gn_bwd(void* stream, const void* dy, const void* x0, const void* x1, int c0, int c1, int n_img,
                          int hw, const float* stats, const float* gamma, const float* beta, int act_silu,
                          float drop_p, uint64_t seed, float* ab, const void* radd, void* dx0, void* dx1,
                          float* dgamma, float* dbeta, const uint64_t* rng_dev, float* colsum_out, float* colsum_total) {
  const int C = c0 + c1;
  TSD_CHECK(C % 128 == 0 && c0 % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "gn_bwd: unsupported channels c0=%d c1=%d", c0, c1);
  cudaStream_t st = (cudaStream_t)stream;
  TSD_CUDA(cudaMemsetAsync(ab, 0, (size_t)n_img * C * 2 * sizeof(float), st));
  int ppc;
  const int gx = gn_grid_x(hw, C, n_img, &ppc);
  gn_bwd_sums_kernel<<<dim3(gx, n_img), 256, 2 * C * sizeof(float), st>>>(
      (const bf16*)dy, (const bf16*)x0, (const bf16*)x1, c0, c1, hw, ppc, stats, gamma, beta, act_silu, drop_p, seed, ab, rng_dev);
  TSD_LAUNCH_CHECK();
  gn_bwd_apply_kernel<<<dim3(gx, n_img), 256, colsum_out ? C * sizeof(float) : 0, st>>>(
      (const bf16*)dy, (const bf16*)x0, (const bf16*)x1, c0, c1, hw, ppc, stats, gamma, beta, act_silu, drop_p, seed, ab,
      (const bf16*)radd, (bf16*)dx0, (bf16*)dx1, rng_dev, colsum_out, colsum_total);
  TSD_LAUNCH_CHECK();
  if (dgamma) {
    gn_bwd_params_kernel<<<ceil_div(C, 8), 256, 0, st>>>(ab, n_img, C, dgamma, dbeta);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int tsd_ln_fwd(void* stream, const void* x, int M, int C, const float* gamma, const float* beta, float eps,
                          void* out) {
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(M, 8 * 8);  // 8 warps x 8 rows per pass
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  if (C == 128) ln_fwd_kernel<1><<<grid, 256, 0, st>>>((const bf16*)x, M, gamma, beta, eps, (bf16*)out);
  else if (C == 256) ln_fwd_kernel<2><<<grid, 256, 0, st>>>((const bf16*)x, M, gamma, beta, eps, (bf16*)out);
  else if (C == 512) ln_fwd_kernel<4><<<grid, 256, 0, st>>>((const bf16*)x, M, gamma, beta, eps, (bf16*)out);
  else TSD_CHECK(false, "ln_fwd: C=%d not in {128,256,512}", C);
  TSD_LAUNCH_CHECK();
  return 0;
}

extern "C" int tsd_ln_bwd(void* stream, const void* dy, const void* x, int M, int C, const float* gamma, float eps,
                          const void* radd, void* dx, float* dgamma, float* dbeta, int rows_per_sample, float* colsum_out,
                          float* colsum_total) {
  cudaStream_t st = (cudaStream_t)stream;
  int rows_per_cta = ceil_div(M, 8 * num_sms());  // 8 CTAs (64 warps) per SM: the per-row shuffle chains are latency bound
  if (rows_per_cta < 8) rows_per_cta = 8;
  if (colsum_out || colsum_total) {
    TSD_CHECK(rows_per_sample > 0 && M % rows_per_sample == 0, "ln_bwd: rows_per_sample=%d does not divide M=%d", rows_per_sample, M);
    if (rows_per_cta > rows_per_sample) rows_per_cta = rows_per_sample;
    while (rows_per_sample % rows_per_cta != 0) --rows_per_cta;  // a CTA never straddles two samples
  }
  const int grid = ceil_div(M, rows_per_cta);
  if (C == 128)
    ln_bwd_kernel<1><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, M, gamma, eps, (const bf16*)radd, (bf16*)dx, dgamma, dbeta, rows_per_cta, rows_per_sample, colsum_out, colsum_total);
  else if (C == 256)
    ln_bwd_kernel<2><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, M, gamma, eps, (const bf16*)radd, (bf16*)dx, dgamma, dbeta, rows_per_cta, rows_per_sample, colsum_out, colsum_total);
  else if (C == 512)
    ln_bwd_kernel<4><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, M, gamma, eps, (const bf16*)radd, (bf16*)dx, dgamma, dbeta, rows_per_cta, rows_per_sample, colsum_out, colsum_total);
  else TSD_CHECK(false, "ln_bwd: C=%d not in {128,256,512}", C);
  TSD_LAUNCH_CHECK();
  return 0;
}
