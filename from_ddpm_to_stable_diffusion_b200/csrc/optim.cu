// Caller-side step body (02_train_direct.py:72-73): clip_grad_norm_(params, max_norm) + AdamW.step(),
// as two passes over flat fp32 buffers: (1) global sum of squares, (2) clip + AdamW in one sweep.
#include "../../include/tinysd_b200.h"
#include "common.cuh"

using namespace tsd;

namespace {

// Global sum of squares, bit-reproducible: every CTA writes its partial to scratch[1 + blockIdx.x]; the CTA that finishes
// last (ticket counter in scratch[0]) adds the partials in index order.  Data-parallel ranks hold identical all-reduced
// gradients and must derive the identical clip coefficient from them, or their weights drift apart ulp by ulp (a float
// atomicAdd across CTAs sums in arrival order).
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n4, size_t n, float* __restrict__ out,
                                                    float* __restrict__ scratch) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ bool s_last;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s[w];
    scratch[1 + blockIdx.x] = v;
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u);
    s_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // fixed-order reduction of the per-CTA partials by one warp (lane-strided partial sums, then a shuffle tree)
    if (threadIdx.x < 32) {
      float t = 0.f;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) t += __ldcg(scratch + 1 + i);
      t = warp_sum(t);
      if (threadIdx.x == 0) {
        out[0] += t;
        *reinterpret_cast<unsigned*>(scratch) = 0u;  // ready for the next call
      }
    }
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias-corrected), gradient pre-scaled by the
// clip coefficient min(1, max_norm / (||g|| + 1e-6)) of torch.nn.utils.clip_grad_norm_.
// lr_dev / step_dev (optional): learning rate and step count read from DEVICE memory, so that a captured CUDA graph
// of the training iteration follows the LR schedule and the bias correction on every replay.
__global__ void __launch_bounds__(256) adamw_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, size_t n, float lr, float beta1, float beta2,
                                                         float eps, float wd, float bc1, float bc2_sqrt, float max_norm,
                                                         const float* __restrict__ sumsq, int write_clipped_grad,
                                                         const float* __restrict__ lr_dev, const int* __restrict__ step_dev) {
  if (lr_dev) lr = *lr_dev;
  if (step_dev) {
    // same expressions as the host path below (fp32 powf), evaluated once per thread: a handful of instructions
    const float st = (float)(*step_dev);
    bc1 = 1.f - powf(beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(beta2, st));
  }
  float coef = 1.f;
  if (max_norm > 0.f) {
    coef = max_norm / (sqrtf(*sumsq) + 1e-6f);
    coef = coef > 1.f ? 1.f : coef;
  }
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * wd, ib1 = 1.f - beta1, ib2 = 1.f - beta2, inv_bc2 = 1.f / bc2_sqrt;
  // flat buffers are 16-byte aligned and padded to a multiple of 4 elements: 128-bit accesses throughout
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = gp[j] * coef;
      mp[j] = beta1 * mp[j] + ib1 * gi;
      vp[j] = beta2 * vp[j] + ib2 * gi * gi;
      pp[j] = pp[j] * decay - step_size * (mp[j] / (sqrtf(vp[j]) * inv_bc2 + eps));
      gp[j] = gi;
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (write_clipped_grad) reinterpret_cast<float4*>(g)[i] = gv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (size_t i = n4 * 4; i < n; ++i) {
      const float gi = g[i] * coef;
      const float mi = beta1 * m[i] + ib1 * gi, vi = beta2 * v[i] + ib2 * gi * gi;
      p[i] = p[i] * decay - step_size * (mi / (sqrtf(vi) * inv_bc2 + eps));
      m[i] = mi; v[i] = vi;
      if (write_clipped_grad) g[i] = gi;
    }
  }
}

// dst[r][c] += src[r][c] for c < cols (dst row pitch ldd, src row pitch lds)
__global__ void add_cols_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols, int ldd, int lds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  dst[(size_t)r * ldd + c] += src[(size_t)r * lds + c];
}

// shadow = decay * shadow + (1 - decay) * p   (EMA.update, utils.py:54-58)
__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, size_t n4, float decay, float w) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 e = reinterpret_cast<float4*>(ema)[i];
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    // (1 - decay) * p + decay * shadow with each product and the sum rounded separately, as torch evaluates it
    e.x = __fadd_rn(__fmul_rn(w, v.x), __fmul_rn(decay, e.x));
    e.y = __fadd_rn(__fmul_rn(w, v.y), __fmul_rn(decay, e.y));
    e.z = __fadd_rn(__fmul_rn(w, v.z), __fmul_rn(decay, e.z));
    e.w = __fadd_rn(__fmul_rn(w, v.w), __fmul_rn(decay, e.w));
    reinterpret_cast<float4*>(ema)[i] = e;
  }
}

__global__ void scale_kernel(float* __restrict__ x, size_t n, float s) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}

inline int ew_grid(size_t items) {
  size_t g = (items + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// out[0] += sum g^2   (out zeroed by the caller; scratch: tsd_sumsq_scratch_floats() floats, zero before the FIRST call)
extern "C" int64_t tsd_sumsq_scratch_floats(void) { return 1 + 16 * 1024; }
extern "C" int tsd_sumsq_f32(void* stream, const float* g, int64_t n, float* out, float* scratch) {
  TSD_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: buffer must be 16-byte aligned");
  TSD_CHECK(scratch != nullptr, "sumsq: scratch buffer required (deterministic reduction)");
  const int grid = ew_grid(n / 4);
  TSD_CHECK(grid <= 16 * 1024, "sumsq: grid %d exceeds the scratch size", grid);
  sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, n / 4, n, out, scratch);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_adamw_clip(void* stream, float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, float wd, int step, float max_norm, const float* sumsq,
                              int write_clipped_grad) {
  TSD_CHECK(step >= 1, "adamw: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_clip_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt,
                                                                 max_norm, sumsq, write_clipped_grad, nullptr, nullptr);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_adamw_clip_dev(void* stream, float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                                  const int* step_dev, float beta1, float beta2, float eps, float wd, float max_norm,
                                  const float* sumsq, int write_clipped_grad) {
  TSD_CHECK(lr_dev != nullptr && step_dev != nullptr, "adamw_clip_dev: lr / step device pointers are required");
  adamw_clip_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, wd, 1.f, 1.f,
                                                                 max_norm, sumsq, write_clipped_grad, lr_dev, step_dev);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_scale_f32(void* stream, float* x, int64_t n, float s) {
  scale_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, s);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_add_cols_f32(void* stream, float* dst, const float* src, int rows, int cols, int ldd, int lds) {
  add_cols_kernel<<<ceil_div(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, rows, cols, ldd, lds);
  TSD_LAUNCH_CHECK();
  return 0;
}
extern "C" int tsd_ema_update(void* stream, float* ema, const float* p, int64_t n, float decay, float one_minus_decay) {
  TSD_CHECK(n % 4 == 0, "ema_update: flat buffers are padded to multiples of 4 elements");
  ema_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(ema, p, n / 4, decay, one_minus_decay);
  TSD_LAUNCH_CHECK();
  return 0;
}
