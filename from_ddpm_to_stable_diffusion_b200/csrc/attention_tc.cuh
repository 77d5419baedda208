// tcgen05 / TMEM attention kernels (attention_tc.cu); dispatched from tsd_attn_fwd / tsd_attn_bwd (attention.cu).
#pragma once
#include "common.cuh"

namespace tsd {

// head_dim 16 or 32 and L a multiple of 256 (two 128-row query tiles per CTA)
bool attn_tc_supported(int L, int C, int heads);
// poly: every poly-th pair of exponentials is evaluated on the FMA pipe (0 = none; 2, 3, 4)
// ws: optional scratch of B * heads floats (max |k|^2 per sample and head); enables the bound mode (no max pass)
int launch_attn_fwd_tc(cudaStream_t st, const void* qkv, void* out, float* lse2, float* ws, int B, int L, int C,
                       int heads, int poly);

// one-pass backward, head_dim 16 and L a multiple of 256 (attention_bwd_tc.cu); ws = zeroed fp32 [B][heads][L][16] for dQ
bool attn_bwd_tc_supported(int L, int C, int heads);
int launch_attn_bwd_tc(cudaStream_t st, const void* qkv, const void* dout, const float* lse2, const float* delta,
                       void* dqkv, float* ws, int B, int L, int C, int heads);

// the same for head_dim 32 and L a multiple of 128, L >= 256 (attention_bwd_tc32.cu); ws = zeroed fp32 [B][heads][L][32]
bool attn_bwd_tc32_supported(int L, int C, int heads);
int launch_attn_bwd_tc32(cudaStream_t st, const void* qkv, const void* dout, const float* lse2, const float* delta,
                         void* dqkv, float* ws, int B, int L, int C, int heads);

}  // namespace tsd
