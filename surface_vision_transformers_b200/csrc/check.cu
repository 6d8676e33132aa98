// check.cu -- the fp32 CHECK MODE of the hot path (north star: "... tightening to 1e-4 in an fp32-accumulate check mode").
//
// Same operator sequence, same parameter / gradient layout and the same C ABI as the bf16 tensor-core path, but every
// operand stays fp32 and every product is an fp32 FMA on the CUDA cores: plain tiled SGEMMs, one-warp-per-row
// LayerNorm, one-block-per-query attention.  Nothing here is tuned -- the point is an independent, high-precision
// execution of the engine's orchestration (offsets, residual wiring, gradient routing, reductions) on the GPU, so that
// the bf16 path's 1e-2 budget can be told apart from logic errors.  Selected per engine with svit_set_check_mode().
//
// Reference being restated: models/sit.py:45-82 and the vit_pytorch Transformer it builds (sit.py:57), in fp32 as the
// reference itself runs it.
#include "check.cuh"

#include <cuda_runtime.h>

#include <cstdint>

namespace svit {
void set_error(const char* fmt, ...);
void count_launch(int n);
namespace ck {

#define CK_CHECK_LAUNCH(name)                                              \
    do {                                                                   \
        count_launch(1);                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) {                                          \
            set_error("%s launch failed: %s", name, cudaGetErrorString(e__)); \
            return -11;                                                    \
        }                                                                  \
    } while (0)

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_exact(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// C[i, j] (+)= sum_k A(i, k) * B(k, j)   with arbitrary element strides -> covers X W^T, dY W and dY^T X
// 64 x 64 tile, 16-deep K slices, 256 threads x (4 x 4) outputs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const Sgemm d) {
    __shared__ float As[16][65];
    __shared__ float Bs[16][65];
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d.K; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            // A tile: pick the faster-varying index along the smaller stride so that the loads coalesce
            int kk, ii;
            if (d.sa_k <= d.sa_i) {
                kk = e & 15;
                ii = e >> 4;
            } else {
                ii = e & 63;
                kk = e >> 6;
            }
            const int gi = i0 + ii, gk = k0 + kk;
            As[kk][ii] = (gi < d.M && gk < d.K) ? d.A[gi * d.sa_i + gk * d.sa_k] : 0.0f;
            int kb, jj;
            if (d.sb_k <= d.sb_j) {
                kb = e & 15;
                jj = e >> 4;
            } else {
                jj = e & 63;
                kb = e >> 6;
            }
            const int gj = j0 + jj, gkb = k0 + kb;
            Bs[kb][jj] = (gj < d.N && gkb < d.K) ? d.B[gkb * d.sb_k + gj * d.sb_j] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gi = i0 + ty * 4 + r;
        if (gi >= d.M) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gj = j0 + tx * 4 + c;
            if (gj >= d.N) continue;
            float v = acc[r][c];
            if (d.bias != nullptr) v += d.bias[gj];
            if (d.resid != nullptr) v += d.resid[gi * d.ldc + gj];
            float* dst = d.C + gi * d.ldc + gj;
            if (d.accumulate) v += *dst;
            *dst = v;
            if (d.gelu_out != nullptr) d.gelu_out[gi * d.ldc + gj] = gelu_exact(v);
        }
    }
}
int sgemm(const Sgemm& d, cudaStream_t st) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) return 0;
    dim3 grid((d.N + 63) / 64, (d.M + 63) / 64);
    sgemm_kernel<<<grid, 256, 0, st>>>(d);
    CK_CHECK_LAUNCH("check sgemm");
    return 0;
}

// y[i] *= gelu'(u[i])
__global__ void mul_dgelu_kernel(float* __restrict__ y, const float* __restrict__ u, size_t n) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] *= dgelu_exact(u[i]);
}
int mul_dgelu(float* y, const float* u, size_t n, cudaStream_t st) {
    mul_dgelu_kernel<<<148 * 8, 256, 0, st>>>(y, u, n);
    CK_CHECK_LAUNCH("check mul_dgelu");
    return 0;
}

// out[j] += sum_i Y[i, j]
__global__ void colsum_kernel(const float* __restrict__ Y, float* __restrict__ out, int M, int N, int rows_per_block) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.0f;
    for (int i = r0; i < r1; ++i) s += Y[static_cast<size_t>(i) * N + j];
    atomicAdd(&out[j], s);
}
int colsum(const float* Y, float* out, int M, int N, cudaStream_t st) {
    const int rpb = 256;
    dim3 grid((N + 127) / 128, (M + rpb - 1) / rpb);
    colsum_kernel<<<grid, 128, 0, st>>>(Y, out, M, N, rpb);
    CK_CHECK_LAUNCH("check colsum");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (one warp per row)
// ------------------------------------------------------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ a, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
                              float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + static_cast<size_t>(row) * D;
    float s = 0.0f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mean = warp_sum(s) / D;
    float q = 0.0f;
    for (int i = lane; i < D; i += 32) {
        const float dx = xr[i] - mean;
        q += dx * dx;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
    for (int i = lane; i < D; i += 32) a[static_cast<size_t>(row) * D + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}
int ln_fwd(const float* x, const float* gamma, const float* beta, float* a, float* mean, float* rstd, int M, int D, float eps,
           cudaStream_t st) {
    ln_fwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, gamma, beta, a, mean, rstd, M, D, eps);
    CK_CHECK_LAUNCH("check ln_fwd");
    return 0;
}
// g_out = g_in + dLN(da) ; dgamma += sum da * xhat ; dbeta += sum da
__global__ void ln_bwd_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ g_in,
                              float* __restrict__ g_out, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const size_t off = static_cast<size_t>(row) * D;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = lane; i < D; i += 32) {
        const float xh = (x[off + i] - mu) * rs;
        const float dy = da[off + i] * gamma[i];
        s1 += dy;
        s2 += dy * xh;
        atomicAdd(&dgamma[i], da[off + i] * xh);
        atomicAdd(&dbeta[i], da[off + i]);
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    for (int i = lane; i < D; i += 32) {
        const float xh = (x[off + i] - mu) * rs;
        const float dy = da[off + i] * gamma[i];
        g_out[off + i] = g_in[off + i] + rs * (dy - s1 - xh * s2);
    }
}
int ln_bwd(const float* da, const float* x, const float* mean, const float* rstd, const float* gamma, const float* g_in,
           float* g_out, float* dgamma, float* dbeta, int M, int D, cudaStream_t st) {
    ln_bwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(da, x, mean, rstd, gamma, g_in, g_out, dgamma, dbeta, M, D);
    CK_CHECK_LAUNCH("check ln_bwd");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// attention, head dimension 64: one block (128 threads) per (sample, head, query)
// ------------------------------------------------------------------------------------------------
constexpr int CK_MAX_T = 384;
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                       float* __restrict__ lse, int B, int H, int T, float scale) {
    __shared__ float sq[64];
    __shared__ float sp[CK_MAX_T];
    __shared__ float red[4];
    const int t = blockIdx.x % T, h = (blockIdx.x / T) % H, b = blockIdx.x / (T * H);
    const int inner = H * 64;
    const size_t rowq = (static_cast<size_t>(b) * T + t) * 3 * inner;
    if (threadIdx.x < 64) sq[threadIdx.x] = qkv[rowq + h * 64 + threadIdx.x];
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < T; j += 128) {
        const float* kr = qkv + (static_cast<size_t>(b) * T + j) * 3 * inner + inner + h * 64;
        float s = 0.0f;
#pragma unroll 8
        for (int d = 0; d < 64; ++d) s = fmaf(sq[d], kr[d], s);
        s *= scale;
        sp[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.0f;
    for (int j = threadIdx.x; j < T; j += 128) {
        const float p = expf(sp[j] - mx);
        sp[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = red[0] + red[1] + red[2] + red[3];
    if (threadIdx.x < 64) {
        float o = 0.0f;
        for (int j = 0; j < T; ++j)
            o = fmaf(sp[j], qkv[(static_cast<size_t>(b) * T + j) * 3 * inner + 2 * inner + h * 64 + threadIdx.x], o);
        out[(static_cast<size_t>(b) * T + t) * inner + h * 64 + threadIdx.x] = o / sum;
    }
    if (threadIdx.x == 0) lse[(static_cast<size_t>(b) * H + h) * T + t] = mx + logf(sum);
}
int attn_fwd(const float* qkv, float* out, float* lse, int B, int H, int T, float scale, cudaStream_t st) {
    if (T > CK_MAX_T) {
        set_error("check attention: T=%d > %d", T, CK_MAX_T);
        return -1;
    }
    attn_fwd_kernel<<<B * H * T, 128, 0, st>>>(qkv, out, lse, B, H, T, scale);
    CK_CHECK_LAUNCH("check attn_fwd");
    return 0;
}
// dqkv must be zeroed by the caller: dK / dV rows receive atomic adds from every query
__global__ void __launch_bounds__(128) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                                       const float* __restrict__ dout, const float* __restrict__ lse,
                                                       float* __restrict__ dqkv, int B, int H, int T, float scale) {
    __shared__ float sq[64], sdo[64];
    __shared__ float sds[CK_MAX_T], spp[CK_MAX_T];
    __shared__ float red[4];
    const int t = blockIdx.x % T, h = (blockIdx.x / T) % H, b = blockIdx.x / (T * H);
    const int inner = H * 64;
    const size_t rowq = (static_cast<size_t>(b) * T + t) * 3 * inner;
    const size_t rowo = (static_cast<size_t>(b) * T + t) * inner + h * 64;
    float dl = 0.0f;
    if (threadIdx.x < 64) {
        sq[threadIdx.x] = qkv[rowq + h * 64 + threadIdx.x];
        sdo[threadIdx.x] = dout[rowo + threadIdx.x];
        dl = dout[rowo + threadIdx.x] * out[rowo + threadIdx.x];
    }
    dl = warp_sum(dl);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dl;
    __syncthreads();
    const float delta = red[0] + red[1];  // (warps 2, 3 contributed zeros)
    const float l = lse[(static_cast<size_t>(b) * H + h) * T + t];
    for (int j = threadIdx.x; j < T; j += 128) {
        const float* kr = qkv + (static_cast<size_t>(b) * T + j) * 3 * inner + inner + h * 64;
        const float* vr = kr + inner;
        float s = 0.0f, dp = 0.0f;
#pragma unroll 8
        for (int d = 0; d < 64; ++d) {
            s = fmaf(sq[d], kr[d], s);
            dp = fmaf(sdo[d], vr[d], dp);
        }
        const float p = expf(s * scale - l);
        spp[j] = p;
        sds[j] = p * (dp - delta) * scale;
    }
    __syncthreads();
    // dQ_t = sum_j dS_j K_j ; dK_j += dS_j Q_t ; dV_j += P_j dO_t
    if (threadIdx.x < 64) {
        float acc = 0.0f;
        for (int j = 0; j < T; ++j) acc = fmaf(sds[j], qkv[(static_cast<size_t>(b) * T + j) * 3 * inner + inner + h * 64 + threadIdx.x], acc);
        dqkv[rowq + h * 64 + threadIdx.x] = acc;
    }
    for (int e = threadIdx.x; e < T * 64; e += 128) {
        const int j = e >> 6, d = e & 63;
        float* dk = dqkv + (static_cast<size_t>(b) * T + j) * 3 * inner + inner + h * 64 + d;
        atomicAdd(dk, sds[j] * sq[d]);
        atomicAdd(dk + inner, spp[j] * sdo[d]);
    }
}
int attn_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int B, int H, int T,
             float scale, cudaStream_t st) {
    if (T > CK_MAX_T) {
        set_error("check attention: T=%d > %d", T, CK_MAX_T);
        return -1;
    }
    cudaMemsetAsync(dqkv, 0, static_cast<size_t>(B) * T * 3 * H * 64 * sizeof(float), st);
    attn_bwd_kernel<<<B * H * T, 128, 0, st>>>(qkv, out, dout, lse, dqkv, B, H, T, scale);
    CK_CHECK_LAUNCH("check attn_bwd");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// patch rows in the reference's own order: A[b*T + 1 + n, v*C + c] = x[b, c, n', v] ('b c n v -> b n (v c)',
// models/sit.py:46), with the MPP corruption / raw-mesh gather options of PackDesc; row b*T (cls slot) = 0
// ------------------------------------------------------------------------------------------------
__global__ void patches_kernel(const PackDesc d, float* __restrict__ A) {
    const int T = d.N + 1;
    const int row = blockIdx.x;
    const int b = row / T, t = row % T;
    const int K = d.C * d.V;
    float* dst = A + static_cast<size_t>(row) * K;
    if (t == 0) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] = 0.0f;
        return;
    }
    int n = t - 1;
    const size_t bn = static_cast<size_t>(b) * d.N + n;
    const bool replace = d.replace_sel != nullptr && d.replace_sel[bn] != 0;
    if (!replace && d.swap_sel != nullptr && d.swap_sel[bn] != 0) n = static_cast<int>(d.swap_src[bn]);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int v = k / d.C, c = k - v * d.C;
        float val;
        if (replace) {
            val = d.mask_token[k];
        } else if (d.table != nullptr) {
            val = d.x[(static_cast<size_t>(b) * d.C + c) * d.n_mesh + d.table[static_cast<size_t>(v) * d.N + n]];
            if (d.ch_mean != nullptr) val = (val - d.ch_mean[c]) / d.ch_std[c];
        } else {
            val = d.x[((static_cast<size_t>(b) * d.C + c) * d.N + n) * d.V + v];
        }
        dst[k] = val;
    }
}
int patches(const PackDesc& d, float* A, cudaStream_t st) {
    patches_kernel<<<d.B * (d.N + 1), 256, 0, st>>>(d, A);
    CK_CHECK_LAUNCH("check patches");
    return 0;
}
// x0[b*T + t, :] = (t == 0 ? cls : x0 + bias) + pos[t]        (models/sit.py:68-73)
__global__ void embed_finish_kernel(float* __restrict__ x0, const float* __restrict__ pos, const float* __restrict__ cls,
                                    const float* __restrict__ bias, int T, int D) {
    const int row = blockIdx.x, t = row % T;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        float* p = x0 + static_cast<size_t>(row) * D + i;
        *p = (t == 0 ? cls[i] : *p + bias[i]) + pos[static_cast<size_t>(t) * D + i];
    }
}
int embed_finish(float* x0, const float* pos, const float* cls, const float* bias, int B, int T, int D, cudaStream_t st) {
    embed_finish_kernel<<<B * T, 128, 0, st>>>(x0, pos, cls, bias, T, D);
    CK_CHECK_LAUNCH("check embed_finish");
    return 0;
}

// dy[b*T + 1 + n, k] = coef * (y - target) on masked rows, 0 elsewhere (target = input rearranged 'b c n v -> b n (v c)')
__global__ void mpp_loss_bwd_kernel(const float* __restrict__ y, const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                    const float* __restrict__ coef_dev, float* __restrict__ dy, int C, int N, int V) {
    const int T = N + 1, K = C * V;
    const int row = blockIdx.x;
    const int b = row / T, t = row % T, n = t - 1;
    const bool on = t > 0 && mask[static_cast<size_t>(b) * N + n] != 0;
    const float coef = *coef_dev;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float val = 0.0f;
        if (on) {
            const int v = k / C, c = k - v * C;
            val = coef * (y[static_cast<size_t>(row) * K + k] - x[((static_cast<size_t>(b) * C + c) * N + n) * V + v]);
        }
        dy[static_cast<size_t>(row) * K + k] = val;
    }
}
int mpp_loss_bwd(const float* y, const float* x, const uint8_t* mask, const float* coef_dev, float* dy, int B, int C, int N,
                 int V, cudaStream_t st) {
    mpp_loss_bwd_kernel<<<B * (N + 1), 128, 0, st>>>(y, x, mask, coef_dev, dy, C, N, V);
    CK_CHECK_LAUNCH("check mpp_loss_bwd");
    return 0;
}

}  // namespace ck
}  // namespace svit
