// dropout.cu -- see dropout.cuh. Four elements per thread and Philox call, 16-byte (fp32) / 8-byte (bf16) accesses.
#include "dropout.cuh"

#include <cuda_bf16.h>

#include "gemm.cuh"  // set_error, count_launch

namespace svit {

#define SVIT_CHECK_LAUNCH(name)                                                  \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            set_error("%s launch failed: %s", name, cudaGetErrorString(e__));    \
            return -11;                                                          \
        }                                                                        \
        count_launch();                                                          \
    } while (0)

struct DropParams {
    uint32_t k0, k1, c1, c2, c3, thresh;
    float scale;
};

// Philox4x32-10 (Salmon et al., SC'11): the published round function and Weyl key schedule.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// multipliers of elements 4q .. 4q+3
__device__ __forceinline__ void drop_mult(const DropParams& d, uint32_t q, float m[4]) {
    const uint4 r = philox4x32_10(q, d.c1, d.c2, d.c3, d.k0, d.k1);
    m[0] = r.x >= d.thresh ? d.scale : 0.0f;
    m[1] = r.y >= d.thresh ? d.scale : 0.0f;
    m[2] = r.z >= d.thresh ? d.scale : 0.0f;
    m[3] = r.w >= d.thresh ? d.scale : 0.0f;
}

static int make_params(const DropoutSite& s, size_t n, DropParams* d, const char* who) {
    if (!(s.p >= 0.0f) || s.p >= 1.0f) {
        set_error("%s: dropout probability must be in [0, 1) (got %f)", who, s.p);
        return -1;
    }
    if ((n >> 2) > 0xFFFFFFFFull) {
        set_error("%s: tensor too large for the 32-bit element-group counter", who);
        return -1;
    }
    d->k0 = static_cast<uint32_t>(s.seed);
    d->k1 = static_cast<uint32_t>(s.seed >> 32);
    d->c1 = s.site;
    d->c2 = static_cast<uint32_t>(s.offset);
    d->c3 = static_cast<uint32_t>(s.offset >> 32);
    d->thresh = static_cast<uint32_t>(static_cast<double>(s.p) * 4294967296.0);
    d->scale = static_cast<float>(1.0 / (1.0 - static_cast<double>(s.p)));
    return 0;
}
static inline int grid_for(size_t groups) {
    size_t blocks = (groups + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return static_cast<int>(blocks < 1 ? 1 : blocks);
}

__device__ __forceinline__ void load4(const float* p, size_t i, size_t n, bool vec, float v[4]) {
    if (vec) {
        const float4 t = *reinterpret_cast<const float4*>(p + i);
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (i + e < n) ? p[i + e] : 0.0f;
    }
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, size_t i, size_t n, bool vec, float v[4]) {
    if (vec) {
        const uint2 t = *reinterpret_cast<const uint2*>(p + i);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (i + e < n) ? __bfloat162float(p[i + e]) : 0.0f;
    }
}
__device__ __forceinline__ void store4(float* p, size_t i, size_t n, bool vec, const float v[4]) {
    if (vec) {
        *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (i + e < n) p[i + e] = v[e];
    }
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, size_t i, size_t n, bool vec, const float v[4]) {
    if (vec) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<const uint32_t*>(&a);
        t.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p + i) = t;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (i + e < n) p[i + e] = __float2bfloat16(v[e]);
    }
}

template <typename T>
__global__ void dropout_scale_kernel(T* __restrict__ a, T* __restrict__ b, size_t n, DropParams d) {
    const size_t groups = (n + 3) >> 2, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < groups; q += stride) {
        const size_t i = q << 2;
        const bool vec = i + 4 <= n;
        float m[4], v[4];
        drop_mult(d, static_cast<uint32_t>(q), m);
        load4(a, i, n, vec, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] *= m[e];
        store4(a, i, n, vec, v);
        if (b != nullptr) {
            load4(b, i, n, vec, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] *= m[e];
            store4(b, i, n, vec, v);
        }
    }
}
__global__ void dropout_residual_kernel(float* __restrict__ out, const float* __restrict__ resid, size_t n, DropParams d) {
    const size_t groups = (n + 3) >> 2, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < groups; q += stride) {
        const size_t i = q << 2;
        const bool vec = i + 4 <= n;
        float m[4], v[4], r[4];
        drop_mult(d, static_cast<uint32_t>(q), m);
        load4(out, i, n, vec, v);
        load4(resid, i, n, vec, r);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fmaf(v[e] - r[e], m[e], r[e]);
        store4(out, i, n, vec, v);
    }
}
template <typename T>
__global__ void dropout_grad_kernel(const float* __restrict__ g, T* __restrict__ out, size_t n, DropParams d) {
    const size_t groups = (n + 3) >> 2, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < groups; q += stride) {
        const size_t i = q << 2;
        const bool vec = i + 4 <= n;
        float m[4], v[4];
        drop_mult(d, static_cast<uint32_t>(q), m);
        load4(g, i, n, vec, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] *= m[e];
        store4(out, i, n, vec, v);
    }
}
__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, size_t n, DropParams d) {
    const size_t groups = (n + 3) >> 2, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < groups; q += stride) {
        float m[4];
        drop_mult(d, static_cast<uint32_t>(q), m);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if ((q << 2) + e < n) keep[(q << 2) + e] = m[e] != 0.0f;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_dropout_scale(void* a, void* b, size_t n, int is_bf16, const DropoutSite& s, cudaStream_t st) {
    if (n == 0) return 0;
    DropParams d;
    if (make_params(s, n, &d, "dropout_scale")) return -1;
    if (!aligned16(a) || !aligned16(b)) {
        set_error("dropout_scale: buffers must be 16-byte aligned");
        return -1;
    }
    const int grid = grid_for((n + 3) >> 2);
    if (is_bf16)
        dropout_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(a),
                                                                   reinterpret_cast<__nv_bfloat16*>(b), n, d);
    else
        dropout_scale_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(a), reinterpret_cast<float*>(b), n, d);
    SVIT_CHECK_LAUNCH("dropout_scale");
    return 0;
}
int launch_dropout_residual(float* out, const float* resid, size_t n, const DropoutSite& s, cudaStream_t st) {
    if (n == 0) return 0;
    DropParams d;
    if (make_params(s, n, &d, "dropout_residual")) return -1;
    if (!aligned16(out) || !aligned16(resid)) {
        set_error("dropout_residual: buffers must be 16-byte aligned");
        return -1;
    }
    dropout_residual_kernel<<<grid_for((n + 3) >> 2), 256, 0, st>>>(out, resid, n, d);
    SVIT_CHECK_LAUNCH("dropout_residual");
    return 0;
}
int launch_dropout_grad(const float* g, void* out, size_t n, int out_bf16, const DropoutSite& s, cudaStream_t st) {
    if (n == 0) return 0;
    DropParams d;
    if (make_params(s, n, &d, "dropout_grad")) return -1;
    if (!aligned16(g) || !aligned16(out)) {
        set_error("dropout_grad: buffers must be 16-byte aligned");
        return -1;
    }
    const int grid = grid_for((n + 3) >> 2);
    if (out_bf16)
        dropout_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, reinterpret_cast<__nv_bfloat16*>(out), n, d);
    else
        dropout_grad_kernel<float><<<grid, 256, 0, st>>>(g, reinterpret_cast<float*>(out), n, d);
    SVIT_CHECK_LAUNCH("dropout_grad");
    return 0;
}
int launch_dropout_mask(uint8_t* keep, size_t n, const DropoutSite& s, cudaStream_t st) {
    if (n == 0) return 0;
    DropParams d;
    if (make_params(s, n, &d, "dropout_mask")) return -1;
    dropout_mask_kernel<<<grid_for((n + 3) >> 2), 256, 0, st>>>(keep, n, d);
    SVIT_CHECK_LAUNCH("dropout_mask");
    return 0;
}

}  // namespace svit
