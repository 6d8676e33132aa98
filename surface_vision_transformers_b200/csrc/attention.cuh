// attention.cuh -- fused multi-head self-attention over the whole (<= 384 token) sequence, sm_100a.
//
// Replaces, for one encoder block, the reference's (vit_pytorch Attention.forward, constructed at
// /root/reference/models/sit.py:57):
//     q,k,v = chunk(to_qkv(x)) ; dots = q k^T * dim_head^-0.5 ; attn = softmax(dots) ; out = attn v ; 'b h n d -> b n (h d)'
// and its autograd backward.  The (B,H,T,T) score tensor never exists in HBM: K and V of one (sample, head)
// stay resident in shared memory, S/P/O live in TMEM, softmax is exact (single pass over the full row).
#pragma once
#include <cuda_runtime.h>

namespace svit {

struct AttnDesc {
    const void* qkv;  // bf16 [B, T, 3*H*64]   (q | k | v, head h = columns [h*64, h*64+64) of each third)
    void* out;        // bf16 [B, T, H*64]      head-merged attention output
    float* lse;       // fp32 [B, H, T]         log-sum-exp of the scaled scores (saved for backward)
    int B, H, T;
    float scale;  // dim_head ** -0.5
};
int launch_attn_fwd(const AttnDesc& d, cudaStream_t stream);

struct AttnBwdDesc {
    const void* qkv;   // bf16 [B, T, 3*H*64]
    const void* out;   // bf16 [B, T, H*64]
    const void* dout;  // bf16 [B, T, H*64]
    const float* lse;  // fp32 [B, H, T]
    void* dqkv;        // bf16 [B, T, 3*H*64]
    int B, H, T;
    float scale;
};
int launch_attn_bwd(const AttnBwdDesc& d, cudaStream_t stream);

// The same attention for the query row of token 0 only (attention_cls.cu): the last encoder block under cls pooling.
struct AttnClsDesc {
    const void* qkv;  // bf16 [B, T, 3*H*64]
    void* out;        // bf16 [B, H*64]     attention output of token 0, head-merged
    float* prob;      // fp32 [B, H, T]     softmax row of token 0 (saved for backward)
    int B, H, T;
    float scale;
};
int launch_attn_cls_fwd(const AttnClsDesc& d, cudaStream_t stream);
struct AttnClsBwdDesc {
    const void* qkv;    // bf16 [B, T, 3*H*64]
    const float* prob;  // fp32 [B, H, T]
    const void* dout;   // bf16 [B, H*64]
    void* dqkv;         // bf16 [B, T, 3*H*64]: dk, dv of every key, dq of token 0, zeros in the other dq rows
    int B, H, T;
    float scale;
};
int launch_attn_cls_bwd(const AttnClsBwdDesc& d, cudaStream_t stream);

}  // namespace svit
