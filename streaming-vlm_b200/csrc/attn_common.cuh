// Declarations shared by the attention translation units (attn_fwd.cu: prefill / ViT kernel and the split merge;
// attn_decode.cu: swap-AB decode kernel).
#pragma once
#include "common.cuh"

namespace svlm {

__global__ void attn_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, uint16_t* __restrict__ out,
                                    int rows, int D, int num_splits);

int attn_decode_launch(const svlm_arena_t* arena, int buffer, int layer, const void* q, void* out, int q_heads, int kv_len,
                       const int32_t* kv_len_dev, int num_splits, void* workspace, float scale, int flags, cudaStream_t st);

}  // namespace svlm
