// attention_cls.cu -- self-attention of ONE query row per (sample, head): the cls token of the last encoder block.
//
// With pool = 'cls' (models/sit.py:78, the reference default) the head reads token 0 of the encoder output only, and
// everything after the last block's key / value projection is row-wise (to_out, residual, LayerNorm, FeedForward).  The
// other 320 query rows of the LAST block therefore influence neither the prediction nor any gradient: the engine computes
// that block's attention for the cls query alone (all keys, all values) and runs the row-wise rest on B rows instead of
// B * T (engine.cu, "cls_last").  Results and every parameter gradient are those of the full computation -- the dropped
// rows have exactly zero gradient in the reference too.
//
// One query against <= 384 keys of 64 dims is ~100 kFLOP per (sample, head) and 2 x 41 KB of K / V: HBM-bound work for
// the CUDA cores (no tensor-core tile has a single live row), one CTA per (sample, head):
//   forward : s_j = scale q0.k_j ; p = softmax(s) ; o = sum_j p_j v_j            -> o bf16 [B, H*64], p fp32 [B, H, T]
//   backward: dp_j = do.v_j ; delta = sum_j p_j dp_j ; ds_j = scale p_j (dp_j - delta)
//             dq0 = sum_j ds_j k_j ; dk_j = ds_j q0 ; dv_j = p_j do              -> dqkv bf16 [B, T, 3*H*64] (dq rows > 0 = 0)
#include "attention.cuh"

#include <cuda_bf16.h>

#include "gemm.cuh"  // set_error, count_launch, launch_pdl
#include "ptx.cuh"

namespace svit {
namespace {

constexpr int CLS_THREADS = 128;  // small CTAs: all B * H of them resident at once (11 per SM at the benchmark shape)
constexpr int CLS_MAX_T = 384;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide reduction over CLS_THREADS threads (red: 8 floats of scratch)
template <bool MAX>
__device__ __forceinline__ float block_reduce(float v, float* red) {
    v = MAX ? warp_max_f(v) : warp_sum_f(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int w = 1; w < CLS_THREADS / 32; ++w) t = MAX ? fmaxf(t, red[w]) : t + red[w];
    return t;
}
// dot product of a 64-element bf16 row in global memory with a 64-float vector in shared memory
__device__ __forceinline__ float dot64(const __nv_bfloat16* row, const float* vec) {
    const uint4* r = reinterpret_cast<const uint4*>(row);
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 x = __ldg(r + c);
        const float* v = vec + c * 8;
        acc += bf16_lo(x.x) * v[0] + bf16_hi(x.x) * v[1] + bf16_lo(x.y) * v[2] + bf16_hi(x.y) * v[3] + bf16_lo(x.z) * v[4] +
               bf16_hi(x.z) * v[5] + bf16_lo(x.w) * v[6] + bf16_hi(x.w) * v[7];
    }
    return acc;
}

__global__ void __launch_bounds__(CLS_THREADS) attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   __nv_bfloat16* __restrict__ out, float* __restrict__ prob,
                                                                   int H, int T, float scale) {
    __shared__ float sq[64];
    __shared__ float sp[CLS_MAX_T];
    __shared__ float sacc[CLS_THREADS / 32][64];
    __shared__ float red[CLS_THREADS / 32];
    griddep_launch();
    griddep_wait();
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int inner = H * 64;
    const size_t pitch = static_cast<size_t>(3) * inner;
    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * pitch + h * 64;
    if (threadIdx.x < 64) sq[threadIdx.x] = __bfloat162float(base[threadIdx.x]);  // q of token 0
    __syncthreads();
    // scores: one key per thread
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < T; j += CLS_THREADS) {
        const float s = dot64(base + j * pitch + inner, sq) * scale;
        sp[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce<true>(mx, red);
    float sum = 0.0f;
    for (int j = threadIdx.x; j < T; j += CLS_THREADS) {
        const float p = __expf(sp[j] - mx);
        sp[j] = p;
        sum += p;
    }
    sum = block_reduce<false>(sum, red);
    const float inv = 1.0f / sum;
    float* pout = prob + static_cast<size_t>(blockIdx.x) * T;
    for (int j = threadIdx.x; j < T; j += CLS_THREADS) {
        const float p = sp[j] * inv;
        sp[j] = p;
        pout[j] = p;
    }
    __syncthreads();
    // o = sum_j p_j v_j: a warp load covers four key rows (8 lanes x 16 bytes each); lane (g, c) owns dims [8 c, 8 c + 8)
    // of the keys j = 4 (warp + 8 it) + g
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, c = lane & 7;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
#pragma unroll 4
    for (int jb = warp * 4; jb < T; jb += 4 * (CLS_THREADS / 32)) {
        const int j = jb + g;
        if (j < T) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + j * pitch + 2 * inner) + c);
            const float p = sp[j];
            acc[0] = fmaf(p, bf16_lo(v.x), acc[0]); acc[1] = fmaf(p, bf16_hi(v.x), acc[1]);
            acc[2] = fmaf(p, bf16_lo(v.y), acc[2]); acc[3] = fmaf(p, bf16_hi(v.y), acc[3]);
            acc[4] = fmaf(p, bf16_lo(v.z), acc[4]); acc[5] = fmaf(p, bf16_hi(v.z), acc[5]);
            acc[6] = fmaf(p, bf16_lo(v.w), acc[6]); acc[7] = fmaf(p, bf16_hi(v.w), acc[7]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (g == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sacc[warp][c * 8 + k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float o = 0.0f;
#pragma unroll
        for (int w = 0; w < CLS_THREADS / 32; ++w) o += sacc[w][threadIdx.x];
        out[static_cast<size_t>(b) * inner + h * 64 + threadIdx.x] = __float2bfloat16(o);
    }
}

__global__ void __launch_bounds__(CLS_THREADS) attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   const float* __restrict__ prob,
                                                                   const __nv_bfloat16* __restrict__ dout,
                                                                   __nv_bfloat16* __restrict__ dqkv, int H, int T, float scale) {
    __shared__ float sq[64];
    __shared__ float sdo[64];
    __shared__ float sp[CLS_MAX_T];
    __shared__ float sds[CLS_MAX_T];
    __shared__ float sacc[CLS_THREADS / 32][64];
    __shared__ float red[CLS_THREADS / 32];
    griddep_launch();
    griddep_wait();
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int inner = H * 64;
    const size_t pitch = static_cast<size_t>(3) * inner;
    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * pitch + h * 64;
    __nv_bfloat16* dbase = dqkv + static_cast<size_t>(b) * T * pitch + h * 64;
    if (threadIdx.x < 64) {
        sq[threadIdx.x] = __bfloat162float(base[threadIdx.x]);
        sdo[threadIdx.x] = __bfloat162float(dout[static_cast<size_t>(b) * inner + h * 64 + threadIdx.x]);
    }
    __syncthreads();
    const float* pin = prob + static_cast<size_t>(blockIdx.x) * T;
    float part = 0.0f;
    for (int j = threadIdx.x; j < T; j += CLS_THREADS) {
        const float dp = dot64(base + j * pitch + 2 * inner, sdo);
        const float p = pin[j];
        sp[j] = p;
        sds[j] = dp;
        part = fmaf(p, dp, part);
    }
    const float delta = block_reduce<false>(part, red);
    for (int j = threadIdx.x; j < T; j += CLS_THREADS) sds[j] = scale * sp[j] * (sds[j] - delta);
    __syncthreads();
    // a warp access covers four key rows (8 lanes x 16 bytes each); lane (g, c) owns dims [8 c, 8 c + 8) of the keys
    // j = 4 (warp + 8 it) + g: dk_j, dv_j (and the zero dq_j) rows out, dq0 accumulated
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, c = lane & 7;
    float q8[8], g8[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        q8[k] = sq[c * 8 + k];
        g8[k] = sdo[c * 8 + k];
        acc[k] = 0.0f;
    }
#pragma unroll 2
    for (int jb = warp * 4; jb < T; jb += 4 * (CLS_THREADS / 32)) {
        const int j = jb + g;
        if (j < T) {
            const float ds = sds[j], p = sp[j];
            const uint4 kk = __ldg(reinterpret_cast<const uint4*>(base + j * pitch + inner) + c);
            acc[0] = fmaf(ds, bf16_lo(kk.x), acc[0]); acc[1] = fmaf(ds, bf16_hi(kk.x), acc[1]);
            acc[2] = fmaf(ds, bf16_lo(kk.y), acc[2]); acc[3] = fmaf(ds, bf16_hi(kk.y), acc[3]);
            acc[4] = fmaf(ds, bf16_lo(kk.z), acc[4]); acc[5] = fmaf(ds, bf16_hi(kk.z), acc[5]);
            acc[6] = fmaf(ds, bf16_lo(kk.w), acc[6]); acc[7] = fmaf(ds, bf16_hi(kk.w), acc[7]);
            uint4 o;
            o.x = pack_bf16(ds * q8[0], ds * q8[1]); o.y = pack_bf16(ds * q8[2], ds * q8[3]);
            o.z = pack_bf16(ds * q8[4], ds * q8[5]); o.w = pack_bf16(ds * q8[6], ds * q8[7]);
            reinterpret_cast<uint4*>(dbase + j * pitch + inner)[c] = o;
            o.x = pack_bf16(p * g8[0], p * g8[1]); o.y = pack_bf16(p * g8[2], p * g8[3]);
            o.z = pack_bf16(p * g8[4], p * g8[5]); o.w = pack_bf16(p * g8[6], p * g8[7]);
            reinterpret_cast<uint4*>(dbase + j * pitch + 2 * inner)[c] = o;
            if (j > 0) reinterpret_cast<uint4*>(dbase + j * pitch)[c] = make_uint4(0u, 0u, 0u, 0u);  // rows that were never queries
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    }
    if (g == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sacc[warp][c * 8 + k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float o = 0.0f;
#pragma unroll
        for (int w = 0; w < CLS_THREADS / 32; ++w) o += sacc[w][threadIdx.x];
        dbase[threadIdx.x] = __float2bfloat16(o);
    }
}

}  // namespace

int launch_attn_cls_fwd(const AttnClsDesc& d, cudaStream_t stream) {
    if (d.B <= 0) return 0;
    if (d.T < 1 || d.T > CLS_MAX_T) {
        set_error("attn_cls_fwd: T=%d unsupported (1..%d)", d.T, CLS_MAX_T);
        return -2;
    }
    cudaError_t e = launch_pdl(attn_cls_fwd_kernel, dim3(d.B * d.H), dim3(CLS_THREADS), 0, stream,
                               reinterpret_cast<const __nv_bfloat16*>(d.qkv), reinterpret_cast<__nv_bfloat16*>(d.out), d.prob,
                               d.H, d.T, d.scale);
    if (e != cudaSuccess) {
        set_error("attn_cls_fwd launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

int launch_attn_cls_bwd(const AttnClsBwdDesc& d, cudaStream_t stream) {
    if (d.B <= 0) return 0;
    if (d.T < 1 || d.T > CLS_MAX_T) {
        set_error("attn_cls_bwd: T=%d unsupported (1..%d)", d.T, CLS_MAX_T);
        return -2;
    }
    cudaError_t e = launch_pdl(attn_cls_bwd_kernel, dim3(d.B * d.H), dim3(CLS_THREADS), 0, stream,
                               reinterpret_cast<const __nv_bfloat16*>(d.qkv), d.prob,
                               reinterpret_cast<const __nv_bfloat16*>(d.dout), reinterpret_cast<__nv_bfloat16*>(d.dqkv), d.H,
                               d.T, d.scale);
    if (e != cudaSuccess) {
        set_error("attn_cls_bwd launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

}  // namespace svit
