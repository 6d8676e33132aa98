// elementwise.cu -- HBM-bound kernels of the SiT hot path. See elementwise.cuh.
#include "elementwise.cuh"

#include <cuda_bf16.h>

#include "gemm.cuh"  // set_error, launch_pdl
#include "ptx.cuh"   // griddep_launch / griddep_wait

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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// =================================================================================================
// a1: patch gather (bit-exact)
// =================================================================================================
// One block per (sample, channel) plane.  The reference's table is (V, N) -- column j lists the vertices of patch j -- while
// the output of a plane is (N, V) contiguous, so a tiny kernel first writes the table transposed ((N, V): entry e of it is
// the vertex of output element e) into a stream-ordered scratch allocation (196 KB, L2-resident for every plane).  The plane
// kernel then stages the whole ico-6 plane (40,962 floats = 160 KB) in shared memory with 8-byte loads, ten in flight per
// thread, and streams the output: 16 bytes of table in, four shared-memory reads, 16 bytes out per thread and step -- no
// tile staging, no barrier after the plane has landed.
constexpr int GATHER_T_THREADS = 1024;
__global__ void transpose_table_kernel(const int32_t* __restrict__ table, int32_t* __restrict__ tt, int N, int V) {
    __shared__ int32_t tile[32][33];
    const int j0 = blockIdx.x * 32, v0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
    for (int r = ty; r < 32; r += 8)
        if (v0 + r < V && j0 + tx < N) tile[r][tx] = table[static_cast<size_t>(v0 + r) * N + j0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < N && v0 + tx < V) tt[static_cast<size_t>(j0 + r) * V + v0 + tx] = tile[tx][r];
}
__global__ void __launch_bounds__(GATHER_T_THREADS) gather_patches_plane_kernel(const float* __restrict__ mesh,
                                                                                const int32_t* __restrict__ tt,
                                                                                float* __restrict__ out, int n_mesh, int NV) {
    extern __shared__ float gather_smem[];
    float* plane_s = gather_smem;  // [n_mesh] (n_mesh even: 8-byte loads; a plane starts 8-byte aligned)
    const int sc = blockIdx.x;
    const float2* plane = reinterpret_cast<const float2*>(mesh + static_cast<size_t>(sc) * n_mesh);
    float2* ps2 = reinterpret_cast<float2*>(plane_s);
    const int n2 = n_mesh >> 1;
    int i = threadIdx.x;
    for (; i + 9 * GATHER_T_THREADS < n2; i += 10 * GATHER_T_THREADS) {
        float2 v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = __ldcs(plane + i + k * GATHER_T_THREADS);
#pragma unroll
        for (int k = 0; k < 10; ++k) ps2[i + k * GATHER_T_THREADS] = v[k];
    }
    for (; i < n2; i += GATHER_T_THREADS) ps2[i] = __ldcs(plane + i);
    __syncthreads();
    float* dst = out + static_cast<size_t>(sc) * NV;
    const int n4 = NV >> 2;  // NV % 4 == 0 (checked by the launcher)
    const int4* t4 = reinterpret_cast<const int4*>(tt);
    float4* d4 = reinterpret_cast<float4*>(dst);
    int e = threadIdx.x;
    for (; e + 3 * GATHER_T_THREADS < n4; e += 4 * GATHER_T_THREADS) {
        int4 id[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = __ldg(t4 + e + k * GATHER_T_THREADS);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            __stcs(d4 + e + k * GATHER_T_THREADS, make_float4(plane_s[id[k].x], plane_s[id[k].y], plane_s[id[k].z], plane_s[id[k].w]));
    }
    for (; e < n4; e += GATHER_T_THREADS) {
        const int4 id = __ldg(t4 + e);
        __stcs(d4 + e, make_float4(plane_s[id.x], plane_s[id.y], plane_s[id.z], plane_s[id.w]));
    }
}

// (round-1 kernel, kept for shapes the plane kernel does not take: n_mesh odd or N * V not a multiple of 4)
// One block per (sample, channel) plane: the whole ico-6 plane (40,962 floats = 160 KB) is staged in shared memory
// with coalesced loads; then, JT patches at a time, the table entries are read coalesced along the patch index, the
// vertices are gathered from the staged plane into a [JT][V] tile, and the tile -- one contiguous run of the output --
// is written out fully coalesced.
constexpr int GATHER_THREADS = 1024;
constexpr int GATHER_TILE_FLOATS = 12 * 1024;  // 48 KB next to the 160 KB plane
__global__ void __launch_bounds__(GATHER_THREADS) gather_patches_smem_kernel(const float* __restrict__ mesh,
                                                                             const int32_t* __restrict__ table,
                                                                             float* __restrict__ out, int n_mesh, int N, int V,
                                                                             int JT) {
    extern __shared__ float gather_smem[];
    float* plane_s = gather_smem;                       // [n_mesh]
    float* tile = gather_smem + ((n_mesh + 31) & ~31);  // [JT][V]
    const int sc = blockIdx.x;
    const float* plane = mesh + static_cast<size_t>(sc) * n_mesh;
    for (int i = threadIdx.x; i < n_mesh; i += GATHER_THREADS) plane_s[i] = __ldcs(plane + i);
    __syncthreads();
    float* dst = out + static_cast<size_t>(sc) * N * V;
    for (int j0 = 0; j0 < N; j0 += JT) {
        const int nj = min(JT, N - j0);
        // lane -> patch, warp -> vertex slot: table reads coalesced along the patch index, no integer division
        const int jj = threadIdx.x & 31;
        if (jj < nj) {
            for (int v = threadIdx.x >> 5; v < V; v += GATHER_THREADS / 32)
                tile[jj * V + v] = plane_s[__ldg(table + static_cast<size_t>(v) * N + j0 + jj)];
        }
        __syncthreads();
        float* d0 = dst + static_cast<size_t>(j0) * V;
        for (int e = threadIdx.x; e < nj * V; e += GATHER_THREADS) d0[e] = tile[e];
        __syncthreads();
    }
}
// fallback for meshes that do not fit in shared memory: one block per (patch, plane)
__global__ void gather_patches_kernel(const float* __restrict__ mesh, const int32_t* __restrict__ table,
                                      float* __restrict__ out, int SC, int n_mesh, int N, int V) {
    const int j = blockIdx.x;
    const int sc = blockIdx.y;
    const float* plane = mesh + static_cast<size_t>(sc) * n_mesh;
    float* dst = out + (static_cast<size_t>(sc) * N + j) * V;
    for (int v = threadIdx.x; v < V; v += blockDim.x) dst[v] = __ldg(plane + table[static_cast<size_t>(v) * N + j]);
}

int launch_gather_patches(const float* mesh, const int32_t* table, float* out, int S, int C, int n_mesh, int N, int V,
                          cudaStream_t st) {
    if (S <= 0) return 0;
    const long long NV = static_cast<long long>(N) * V;
    const size_t plane_bytes = static_cast<size_t>(n_mesh) * sizeof(float);
    if ((n_mesh % 2) == 0 && (NV % 4) == 0 && NV > 0 && NV < (1LL << 31) && plane_bytes <= 227 * 1024 &&
        (reinterpret_cast<uintptr_t>(mesh) % 8) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0) {
        static bool plane_configured = false;
        if (!plane_configured) {
            cudaFuncSetAttribute(gather_patches_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            // keep the scratch of the transposed table in the device's default pool between calls
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = 64ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            plane_configured = true;
        }
        int32_t* tt = nullptr;
        if (cudaMallocAsync(reinterpret_cast<void**>(&tt), static_cast<size_t>(NV) * sizeof(int32_t), st) != cudaSuccess) {
            set_error("gather_patches: scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return -12;
        }
        transpose_table_kernel<<<dim3((N + 31) / 32, (V + 31) / 32), 256, 0, st>>>(table, tt, N, V);
        gather_patches_plane_kernel<<<S * C, GATHER_T_THREADS, plane_bytes, st>>>(mesh, tt, out, n_mesh, static_cast<int>(NV));
        const cudaError_t le = cudaGetLastError();
        cudaFreeAsync(tt, st);
        if (le != cudaSuccess) {
            set_error("gather_patches launch failed: %s", cudaGetErrorString(le));
            return -11;
        }
        count_launch(2);
        return 0;
    }
    int JT = GATHER_TILE_FLOATS / (V > 0 ? V : 1);
    if (JT > 32) JT = 32;
    const size_t smem = (static_cast<size_t>((n_mesh + 31) & ~31) + static_cast<size_t>(JT > 0 ? JT : 1) * V) * sizeof(float);
    if (JT >= 1 && smem <= 227 * 1024) {
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(gather_patches_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            configured = true;
        }
        gather_patches_smem_kernel<<<S * C, GATHER_THREADS, smem, st>>>(mesh, table, out, n_mesh, N, V, JT);
        SVIT_CHECK_LAUNCH("gather_patches");
        return 0;
    }
    dim3 grid(N, S * C);
    gather_patches_kernel<<<grid, 128, 0, st>>>(mesh, table, out, S * C, n_mesh, N, V);
    SVIT_CHECK_LAUNCH("gather_patches");
    return 0;
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// =================================================================================================
// a2/a8: pack patches (+ MPP corruption, + optional fused gather / z-score)
// =================================================================================================
// One thread per 8 consecutive output elements (one 16-byte store); the operand row is channel-major (k = c * V + v),
// so the 8 inputs of a thread are consecutive floats of one channel (or straddle one channel boundary).
// IdxT = uint32_t whenever the vector count fits (always, in practice): the two divisions per thread are 32-bit then.
// input element `off`: fp32, or bf16 when the caller staged the batch as bf16 (the operand is rounded to bf16 here
// anyway, so a batch rounded once on the host gives bit-identical results at half the host-to-device bytes)
__device__ __forceinline__ float pack_ldx(const PackDesc& d, size_t off) {
    return d.x_bf16 ? __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(d.x) + off)) : __ldg(d.x + off);
}
template <typename IdxT>
__global__ void __launch_bounds__(256) pack_patches_kernel(const PackDesc d) {
    const int T = d.N + 1;
    const int nvec = d.Kp >> 3;
    const IdxT total = static_cast<IdxT>(d.B) * T * nvec;
    const int CV = d.C * d.V;
    for (IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<IdxT>(gridDim.x) * blockDim.x) {
        const IdxT row = idx / static_cast<IdxT>(nvec);  // b*T + t
        const int i = static_cast<int>(idx - row * nvec);
        const int b = static_cast<int>(row / static_cast<IdxT>(T)), t = static_cast<int>(row - static_cast<IdxT>(b) * T);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.A) + static_cast<size_t>(row) * d.Kp) + i;
        if (t == 0) {
            *dst = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        int n = t - 1;
        const size_t bn = static_cast<size_t>(b) * d.N + n;
        const bool replace = d.replace_sel != nullptr && d.replace_sel[bn] != 0;
        if (!replace && d.swap_sel != nullptr && d.swap_sel[bn] != 0) n = static_cast<int>(d.swap_src[bn]);
        float val[8];
        const int k0 = i * 8;
        int c = k0 / d.V, v = k0 - c * d.V;
        if (!replace && d.table == nullptr && k0 + 8 <= CV) {
            // plain pre-patched input, no padding column in this vector: 8 consecutive (c, v) of one patch
            size_t src = ((static_cast<size_t>(b) * d.C + c) * d.N + n) * d.V + v;
            const size_t cstep = static_cast<size_t>(d.N - 1) * d.V;  // extra offset once the channel wraps
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                val[e] = pack_ldx(d, src + e);
                if (++v == d.V) {
                    v = 0;
                    src += cstep;
                }
            }
            uint4 o;
            o.x = pack2_bf16(val[0], val[1]);
            o.y = pack2_bf16(val[2], val[3]);
            o.z = pack2_bf16(val[4], val[5]);
            o.w = pack2_bf16(val[6], val[7]);
            *dst = o;
            continue;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = 0.0f;
            if (k0 + e < CV) {
                if (replace) {
                    x = d.mask_token[v * d.C + c];
                } else if (d.table != nullptr) {
                    x = pack_ldx(d, (static_cast<size_t>(b) * d.C + c) * d.n_mesh + d.table[static_cast<size_t>(v) * d.N + n]);
                    if (d.ch_mean != nullptr) x = (x - d.ch_mean[c]) / d.ch_std[c];
                } else {
                    x = pack_ldx(d, ((static_cast<size_t>(b) * d.C + c) * d.N + n) * d.V + v);
                }
            }
            val[e] = x;
            if (++v == d.V) {
                v = 0;
                ++c;
            }
        }
        uint4 o;
        o.x = pack2_bf16(val[0], val[1]);
        o.y = pack2_bf16(val[2], val[3]);
        o.z = pack2_bf16(val[4], val[5]);
        o.w = pack2_bf16(val[6], val[7]);
        *dst = o;
    }
}

int launch_pack_patches(const PackDesc& d, cudaStream_t st) {
    if (d.B <= 0) return 0;
    if (d.Kp < d.C * d.V || (d.Kp % 8) != 0) {
        set_error("pack_patches: Kp=%d must be a multiple of 8 and >= C*V=%d", d.Kp, d.C * d.V);
        return -2;
    }
    const size_t total = static_cast<size_t>(d.B) * (d.N + 1) * (d.Kp / 8);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (total + static_cast<size_t>(148) * 64 * 256 < (static_cast<size_t>(1) << 32))
        pack_patches_kernel<uint32_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(d);
    else
        pack_patches_kernel<size_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(d);
    SVIT_CHECK_LAUNCH("pack_patches");
    return 0;
}

// =================================================================================================
// LayerNorm forward (one warp per row; row cached in registers, two-pass variance).
// NVEC = float4 per lane (compile time, so only the registers a given D needs are allocated).
// =================================================================================================
constexpr int LN_MAX_VEC = 8;  // D <= 1024

template <int NVEC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ a,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int M,
                                                     int D, float eps) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    const float inv_d = 1.0f / D;
    griddep_launch();
    griddep_wait();
    float4 g[NVEC], bb[NVEC];
#pragma unroll
    for (int k = 0; k < NVEC; ++k) {
        const int i = lane + k * 32;
        if (i < nvec) {
            g[k] = reinterpret_cast<const float4*>(gamma)[i];
            bb[k] = reinterpret_cast<const float4*>(beta)[i];
        }
    }
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
        float4 v[NVEC];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                v[k] = __ldcs(xr + i);
                s += v[k].x + v[k].y + v[k].z + v[k].w;
            }
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.0f;
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
                q += dx * dx + dy * dy + dz * dz + dw * dw;
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
        uint2* ar = reinterpret_cast<uint2*>(a + static_cast<size_t>(row) * D);
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                __nv_bfloat162 lo = __floats2bfloat162_rn((v[k].x - mean) * rstd * g[k].x + bb[k].x,
                                                          (v[k].y - mean) * rstd * g[k].y + bb[k].y);
                __nv_bfloat162 hi = __floats2bfloat162_rn((v[k].z - mean) * rstd * g[k].z + bb[k].z,
                                                          (v[k].w - mean) * rstd * g[k].w + bb[k].w);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t*>(&lo);
                o.y = *reinterpret_cast<uint32_t*>(&hi);
                ar[i] = o;
            }
        }
    }
}

static int ln_shape_ok(int D) {
    if (D <= 0 || (D % 4) != 0 || D > LN_MAX_VEC * 128) {
        set_error("layernorm: D=%d must be a multiple of 4 and <= %d", D, LN_MAX_VEC * 128);
        return 0;
    }
    return 1;
}

#define SVIT_LN_DISPATCH(nv, CALL)                 \
    switch (nv) {                                  \
        case 1: { constexpr int NV = 1; CALL; break; } \
        case 2: { constexpr int NV = 2; CALL; break; } \
        case 3: { constexpr int NV = 3; CALL; break; } \
        case 4: { constexpr int NV = 4; CALL; break; } \
        case 5: case 6: { constexpr int NV = 6; CALL; break; } \
        default: { constexpr int NV = 8; CALL; break; } \
    }

int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* a_bf16, float* mean, float* rstd, int M,
                  int D, float eps, cudaStream_t st) {
    if (M <= 0) return 0;
    if (!ln_shape_ok(D)) return -2;
    const int wpb = 8;
    int blocks = (M + wpb - 1) / wpb;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const int nv = (D + 127) / 128;
    SVIT_LN_DISPATCH(nv, (launch_pdl(ln_fwd_kernel<NV>, dim3(blocks), dim3(wpb * 32), 0, st, x, gamma, beta,
                                     reinterpret_cast<__nv_bfloat16*>(a_bf16), mean, rstd, M, D, eps)));
    SVIT_CHECK_LAUNCH("ln_fwd");
    return 0;
}

// =================================================================================================
// LayerNorm backward + residual-gradient add + column reductions
// =================================================================================================
// Register budget: 112 per thread, so that THREE 128-thread CTAs (12 warps) fit an SM next to a resident weight-gradient
// GEMM CTA (20 k of the 64 k registers; engine.cu runs those GEMMs on a side stream under this kernel) and four when the
// kernel has the SM to itself.  gamma therefore lives in shared memory, not in registers.
template <int NVEC, bool CLSG = false>
__global__ void __maxnreg__(NVEC <= 3 ? 112 : 232) ln_bwd_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* g_in, float* g_out,
                                                     __nv_bfloat16* __restrict__ g_out_bf16, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, float* __restrict__ colsum_out, int M,
                                                     int D, int period) {
    extern __shared__ float red[];  // [3][D] block-level partial column sums, then [D] gamma
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    const float inv_d = 1.0f / D;
    const float4* sgm = reinterpret_cast<const float4*>(red + 3 * D);
    griddep_launch();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.0f;
    griddep_wait();
    for (int i = threadIdx.x; i < D; i += blockDim.x) red[3 * D + i] = gamma[i];
    __syncthreads();
    float4 acc_g[NVEC], acc_b[NVEC], acc_c[NVEC];
#pragma unroll
    for (int k = 0; k < NVEC; ++k) {
        acc_g[k] = make_float4(0, 0, 0, 0);
        acc_b[k] = make_float4(0, 0, 0, 0);
        acc_c[k] = make_float4(0, 0, 0, 0);
    }
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
        const float mu = mean[row], rs = rstd[row];
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
        const uint2* dar = reinterpret_cast<const uint2*>(da + static_cast<size_t>(row) * D);
        // CLSG: the incoming gradient exists for rows 0, period, 2 period, ... only (compact matrix), zero elsewhere
        const bool has_g = !CLSG || (row % period) == 0;
        const float4* gir = reinterpret_cast<const float4*>(g_in + static_cast<size_t>(CLSG ? row / period : row) * D);
        float4 xh[NVEC], dy[NVEC], gi[NVEC];
        float s1 = 0.0f, s2 = 0.0f;
        // issue every load of the row first (memory-level parallelism), then reduce
        uint2 dv[NVEC];
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                xh[k] = __ldcs(xr + i);
                dv[k] = __ldcs(dar + i);
                gi[k] = has_g ? __ldcs(gir + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv[k].x);
                const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv[k].y);
                const float4 d = make_float4(__bfloat162float(d01.x), __bfloat162float(d01.y), __bfloat162float(d23.x),
                                             __bfloat162float(d23.y));
                xh[k] = make_float4((xh[k].x - mu) * rs, (xh[k].y - mu) * rs, (xh[k].z - mu) * rs, (xh[k].w - mu) * rs);
                acc_g[k].x += d.x * xh[k].x; acc_g[k].y += d.y * xh[k].y; acc_g[k].z += d.z * xh[k].z; acc_g[k].w += d.w * xh[k].w;
                acc_b[k].x += d.x; acc_b[k].y += d.y; acc_b[k].z += d.z; acc_b[k].w += d.w;
                const float4 gm = sgm[i];
                dy[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
                s1 += dy[k].x + dy[k].y + dy[k].z + dy[k].w;
                s2 += dy[k].x * xh[k].x + dy[k].y * xh[k].y + dy[k].z * xh[k].z + dy[k].w * xh[k].w;
            }
        }
        // two independent butterfly reductions interleaved
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 * inv_d, m2 = s2 * inv_d;
        float4* gor = reinterpret_cast<float4*>(g_out + static_cast<size_t>(row) * D);
        uint2* gbr = reinterpret_cast<uint2*>(g_out_bf16 + static_cast<size_t>(row) * D);
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            const int i = lane + k * 32;
            if (i < nvec) {
                float4 o;
                o.x = gi[k].x + rs * (dy[k].x - m1 - xh[k].x * m2);
                o.y = gi[k].y + rs * (dy[k].y - m1 - xh[k].y * m2);
                o.z = gi[k].z + rs * (dy[k].z - m1 - xh[k].z * m2);
                o.w = gi[k].w + rs * (dy[k].w - m1 - xh[k].w * m2);
                gor[i] = o;
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                uint2 ob;
                ob.x = *reinterpret_cast<uint32_t*>(&lo);
                ob.y = *reinterpret_cast<uint32_t*>(&hi);
                gbr[i] = ob;
                acc_c[k].x += o.x; acc_c[k].y += o.y; acc_c[k].z += o.z; acc_c[k].w += o.w;
            }
        }
    }
    // block reduction through shared memory, then one atomic per column per block
#pragma unroll
    for (int k = 0; k < NVEC; ++k) {
        const int i = lane + k * 32;
        if (i < nvec) {
            const int c = i * 4;
            atomicAdd(&red[c + 0], acc_g[k].x); atomicAdd(&red[c + 1], acc_g[k].y);
            atomicAdd(&red[c + 2], acc_g[k].z); atomicAdd(&red[c + 3], acc_g[k].w);
            atomicAdd(&red[D + c + 0], acc_b[k].x); atomicAdd(&red[D + c + 1], acc_b[k].y);
            atomicAdd(&red[D + c + 2], acc_b[k].z); atomicAdd(&red[D + c + 3], acc_b[k].w);
            atomicAdd(&red[2 * D + c + 0], acc_c[k].x); atomicAdd(&red[2 * D + c + 1], acc_c[k].y);
            atomicAdd(&red[2 * D + c + 2], acc_c[k].z); atomicAdd(&red[2 * D + c + 3], acc_c[k].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        atomicAdd(&dgamma[i], red[i]);
        atomicAdd(&dbeta[i], red[D + i]);
        if (colsum_out != nullptr) atomicAdd(&colsum_out[i], red[2 * D + i]);
    }
}

int launch_ln_bwd(const void* da_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum_out,
                  int M, int D, cudaStream_t st, int g_in_period) {
    if (M <= 0) return 0;
    if (!ln_shape_ok(D)) return -2;
    // 128-thread CTAs: three of them fit next to a weight-gradient GEMM CTA (see the kernel); SVIT_LN_BWD_WARPS=8 brings the
    // former 256-thread CTAs back for A/B timing.  The same number of warps (and rows per warp) either way.
    static const int wpb_env = getenv("SVIT_LN_BWD_WARPS") != nullptr ? atoi(getenv("SVIT_LN_BWD_WARPS")) : 4;
    const int wpb = (wpb_env == 8) ? 8 : 4;
    int blocks = (M + wpb - 1) / wpb;
    if (blocks > 148 * 32 / wpb) blocks = 148 * 32 / wpb;
    const int nv = (D + 127) / 128;
    if (g_in_period > 0) {
        if (g_in == g_out) {
            set_error("ln_bwd: a compact incoming gradient cannot alias the full-size output");
            return -2;
        }
        SVIT_LN_DISPATCH(nv, (launch_pdl(ln_bwd_kernel<NV, true>, dim3(blocks), dim3(wpb * 32), 4 * D * sizeof(float), st,
                                         reinterpret_cast<const __nv_bfloat16*>(da_bf16), x, mean, rstd, gamma, g_in, g_out,
                                         reinterpret_cast<__nv_bfloat16*>(g_out_bf16), dgamma, dbeta, colsum_out, M, D,
                                         g_in_period)));
        SVIT_CHECK_LAUNCH("ln_bwd");
        return 0;
    }
    SVIT_LN_DISPATCH(nv, (launch_pdl(ln_bwd_kernel<NV>, dim3(blocks), dim3(wpb * 32), 4 * D * sizeof(float), st,
                                     reinterpret_cast<const __nv_bfloat16*>(da_bf16), x, mean, rstd, gamma, g_in, g_out,
                                     reinterpret_cast<__nv_bfloat16*>(g_out_bf16), dgamma, dbeta, colsum_out, M, D, 0)));
    SVIT_CHECK_LAUNCH("ln_bwd");
    return 0;
}

// =================================================================================================
// a5: head forward / backward   (one block per sample)
// =================================================================================================
__device__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0f;
    if (w == 0) t = warp_sum(t);
    if (threadIdx.x == 0) scratch[0] = t;
    __syncthreads();
    return scratch[0];
}

// pooled[d] into shared memory p[] (D floats); returns (mean, rstd) of the pooled vector
__device__ void head_pool_stats(const float* x, int T, int D, int pool_mean, float eps, float* p, float* scratch,
                                float& mu, float& rs) {
    const float* xb = x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float v;
        if (pool_mean) {
            v = 0.0f;
            for (int t = 0; t < T; ++t) v += xb[static_cast<size_t>(t) * D + d];
            v /= T;
        } else {
            v = xb[d];
        }
        p[d] = v;
    }
    __syncthreads();
    float s = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) s += p[d];
    mu = block_sum(s, scratch) / D;
    float q = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) q += (p[d] - mu) * (p[d] - mu);
    rs = rsqrtf(block_sum(q, scratch) / D + eps);
}

__global__ void head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out, int T,
                                int D, int C, int pool_mean, float eps) {
    extern __shared__ float sm[];
    float* p = sm;
    float* scratch = sm + D;
    const int b = blockIdx.x;
    float mu, rs;
    head_pool_stats(x + static_cast<size_t>(b) * T * D, T, D, pool_mean, eps, p, scratch, mu, rs);
    for (int d = threadIdx.x; d < D; d += blockDim.x) p[d] = (p[d] - mu) * rs * gamma[d] + beta[d];
    __syncthreads();
    for (int c = 0; c < C; ++c) {
        float s = 0.0f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) s += p[d] * W[static_cast<size_t>(c) * D + d];
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) out[static_cast<size_t>(b) * C + c] = s + bias[c];
    }
}

__global__ void head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ W, const float* __restrict__ dout, float* __restrict__ g,
                                __nv_bfloat16* __restrict__ g_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                float* __restrict__ dW, float* __restrict__ dbias, float* __restrict__ colsum_out, int T,
                                int D, int C, int pool_mean, float eps) {
    extern __shared__ float sm[];
    float* p = sm;           // pooled -> xhat
    float* dz = sm + D;      // grad wrt LN output
    float* scratch = sm + 2 * D;
    const int b = blockIdx.x;
    float mu, rs;
    head_pool_stats(x + static_cast<size_t>(b) * T * D, T, D, pool_mean, eps, p, scratch, mu, rs);
    float s1 = 0.0f, s2 = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float xh = (p[d] - mu) * rs;
        const float z = xh * gamma[d] + beta[d];
        float dzd = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float go = dout[static_cast<size_t>(b) * C + c];
            dzd += go * W[static_cast<size_t>(c) * D + d];
            atomicAdd(&dW[static_cast<size_t>(c) * D + d], go * z);
        }
        atomicAdd(&dgamma[d], dzd * xh);
        atomicAdd(&dbeta[d], dzd);
        const float dy = dzd * gamma[d];
        p[d] = xh;
        dz[d] = dy;
        s1 += dy;
        s2 += dy * xh;
    }
    if (threadIdx.x < C) atomicAdd(&dbias[threadIdx.x], dout[static_cast<size_t>(b) * C + threadIdx.x]);
    const float m1 = block_sum(s1, scratch) / D;
    const float m2 = block_sum(s2, scratch) / D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float dp = rs * (dz[d] - m1 - p[d] * m2);  // grad wrt pooled vector
        dz[d] = dp;
        if (colsum_out != nullptr) atomicAdd(&colsum_out[d], dp);  // sum over rows of g equals dp for both pool modes
    }
    __syncthreads();
    float* gb = g + static_cast<size_t>(b) * T * D;
    __nv_bfloat16* gbb = g_bf16 + static_cast<size_t>(b) * T * D;
    const float inv_t = 1.0f / T;
    // g of this sample: every row (mean pooling) or row 0 only (cls pooling) carries dz, the rest is zero -- 6 bytes per
    // element of pure store traffic: one warp per row, 16-byte fp32 and 8-byte bf16 stores (D % 4 == 0)
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int t = threadIdx.x >> 5; t < T; t += nw) {
        const bool live = pool_mean || t == 0;
        const float sc = pool_mean ? inv_t : 1.0f;
        for (int d = lane * 4; d < D; d += 128) {
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (live) v = make_float4(dz[d] * sc, dz[d + 1] * sc, dz[d + 2] * sc, dz[d + 3] * sc);
            *reinterpret_cast<float4*>(gb + static_cast<size_t>(t) * D + d) = v;
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&lo);
            o.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(gbb + static_cast<size_t>(t) * D + d) = o;
        }
    }
}

int launch_head_fwd(const float* x, const float* gamma, const float* beta, const float* W, const float* bias, float* out,
                    int B, int T, int D, int C, int pool_mean, float eps, cudaStream_t st) {
    if (B <= 0) return 0;
    head_fwd_kernel<<<B, 128, (D + 32) * sizeof(float), st>>>(x, gamma, beta, W, bias, out, T, D, C, pool_mean, eps);
    SVIT_CHECK_LAUNCH("head_fwd");
    return 0;
}

int launch_head_bwd(const float* x, const float* gamma, const float* beta, const float* W, const float* dout, float* g,
                    void* g_bf16, float* dgamma, float* dbeta, float* dW, float* dbias, float* colsum_out, int B, int T,
                    int D, int C, int pool_mean, float eps, cudaStream_t st) {
    if (B <= 0) return 0;
    if (C > 256) {
        set_error("head_bwd: num_classes=%d > 256 unsupported", C);
        return -2;
    }
    head_bwd_kernel<<<B, 512, (2 * D + 32) * sizeof(float), st>>>(x, gamma, beta, W, dout, g,
                                                                  reinterpret_cast<__nv_bfloat16*>(g_bf16), dgamma, dbeta,
                                                                  dW, dbias, colsum_out, T, D, C, pool_mean, eps);
    SVIT_CHECK_LAUNCH("head_bwd");
    return 0;
}

// =================================================================================================
// column sums, casts, weight shadows
// =================================================================================================
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ Y, float* __restrict__ out, int M, int N, int ld,
                                   int rows_per_block) {
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    for (int c = threadIdx.x * 2; c < N; c += blockDim.x * 2) {
        float s0 = 0.0f, s1 = 0.0f;
        for (int r = r0; r < r1; ++r) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(Y + static_cast<size_t>(r) * ld + c);
            s0 += __bfloat162float(v.x);
            s1 += __bfloat162float(v.y);
        }
        atomicAdd(&out[c], s0);
        if (c + 1 < N) atomicAdd(&out[c + 1], s1);
    }
}

int launch_colsum_bf16(const void* Y, float* out, int M, int N, int ld, cudaStream_t st) {
    if (M <= 0) return 0;
    if ((N & 1) || (ld & 1)) {
        set_error("colsum_bf16: N and ld must be even");
        return -2;
    }
    const int rpb = 128;
    colsum_bf16_kernel<<<(M + rpb - 1) / rpb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(Y), out, M, N, ld, rpb);
    SVIT_CHECK_LAUNCH("colsum_bf16");
    return 0;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16(src[i]);
}
// 8 elements per thread and step: two 16-byte loads, one 16-byte store, two steps in flight (n8 = n / 8 vectors)
__global__ void __launch_bounds__(256) cast_bf16_vec_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, size_t n8) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) {
        const float4 a0 = __ldcs(src + 2 * i), a1 = __ldcs(src + 2 * i + 1);
        const float4 b0 = __ldcs(src + 2 * (i + stride)), b1 = __ldcs(src + 2 * (i + stride) + 1);
        dst[i] = make_uint4(pack2_bf16(a0.x, a0.y), pack2_bf16(a0.z, a0.w), pack2_bf16(a1.x, a1.y), pack2_bf16(a1.z, a1.w));
        dst[i + stride] = make_uint4(pack2_bf16(b0.x, b0.y), pack2_bf16(b0.z, b0.w), pack2_bf16(b1.x, b1.y), pack2_bf16(b1.z, b1.w));
    }
    for (; i < n8; i += stride) {
        const float4 a0 = __ldcs(src + 2 * i), a1 = __ldcs(src + 2 * i + 1);
        dst[i] = make_uint4(pack2_bf16(a0.x, a0.y), pack2_bf16(a0.z, a0.w), pack2_bf16(a1.x, a1.y), pack2_bf16(a1.z, a1.w));
    }
}
int launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    const size_t n8 = ((reinterpret_cast<uintptr_t>(src) % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0) ? n / 8 : 0;
    if (n8 > 0) {
        size_t blocks = (n8 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        cast_bf16_vec_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<const float4*>(src),
                                                                      reinterpret_cast<uint4*>(dst), n8);
        SVIT_CHECK_LAUNCH("cast_bf16");
    }
    const size_t done = n8 * 8;
    if (done < n) {   // tail (or everything, for unaligned pointers)
        const size_t rest = n - done;
        size_t blocks = (rest + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        cast_bf16_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(src + done, reinterpret_cast<__nv_bfloat16*>(dst) + done, rest);
        SVIT_CHECK_LAUNCH("cast_bf16");
    }
    return 0;
}

__global__ void cast_transpose_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      __nv_bfloat16* __restrict__ dstT, int rows, int cols, int ld_direct, int ld_t,
                                      size_t src_stride, size_t dst_stride, size_t dstT_stride) {
    __shared__ float tile[32][33];
    const int z = blockIdx.z;
    src += z * src_stride;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        float v = 0.0f;
        if (r < rows && c < cols) {
            v = src[static_cast<size_t>(r) * cols + c];
            if (dst != nullptr) dst[z * dst_stride + static_cast<size_t>(r) * ld_direct + c] = __float2bfloat16(v);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    if (dstT != nullptr) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int c = c0 + i, r = r0 + threadIdx.x;
            if (c < cols && r < rows) dstT[z * dstT_stride + static_cast<size_t>(c) * ld_t + r] = __float2bfloat16(tile[threadIdx.x][i]);
        }
    }
}

int launch_cast_transpose(const float* src, void* dst, void* dstT, int rows, int cols, int ld_direct, int ld_t, int count,
                          size_t src_stride, size_t dst_stride, size_t dstT_stride, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || count <= 0) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, count);
    dim3 block(32, 8);
    cast_transpose_kernel<<<grid, block, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                  reinterpret_cast<__nv_bfloat16*>(dstT), rows, cols, ld_direct, ld_t,
                                                  src_stride, dst_stride, dstT_stride);
    SVIT_CHECK_LAUNCH("cast_transpose");
    return 0;
}

__global__ void prepare_patch_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wp, int D, int C, int V,
                                            int Kp) {
    const int d = blockIdx.x;
    const int CV = C * V;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        float v = 0.0f;
        if (k < CV) {
            const int c = k / V, vv = k - c * V;
            v = W[static_cast<size_t>(d) * CV + vv * C + c];
        }
        Wp[static_cast<size_t>(d) * Kp + k] = __float2bfloat16(v);
    }
}
int launch_prepare_patch_weight(const float* W, void* Wp, int D, int C, int V, int Kp, cudaStream_t st) {
    prepare_patch_weight_kernel<<<D, 256, 0, st>>>(W, reinterpret_cast<__nv_bfloat16*>(Wp), D, C, V, Kp);
    SVIT_CHECK_LAUNCH("prepare_patch_weight");
    return 0;
}

__global__ void prepare_rowtab_kernel(const float* __restrict__ pos, const float* __restrict__ cls,
                                      const float* __restrict__ bias, float* __restrict__ E, int T, int D) {
    const int t = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        E[static_cast<size_t>(t) * D + d] = pos[static_cast<size_t>(t) * D + d] + (t == 0 ? cls[d] : bias[d]);
}
int launch_prepare_rowtab(const float* pos, const float* cls, const float* bias, float* E, int T, int D, cudaStream_t st) {
    prepare_rowtab_kernel<<<T, 128, 0, st>>>(pos, cls, bias, E, T, D);
    SVIT_CHECK_LAUNCH("prepare_rowtab");
    return 0;
}

// =================================================================================================
// patch-embedding backward reductions
// =================================================================================================
// dpos[t, :] += sum_b g0[b, t, :] ; dcls += row t = 0 ; dbias += rows t >= 1.
// grid (T, batch slices): every block sums its slice of the batch for one token position (float4 loads, 4 independent
// accumulator chains per thread) and adds the partial sums atomically.
__global__ void embed_bwd_kernel(const float* __restrict__ g0, float* __restrict__ dpos, float* __restrict__ dcls,
                                 float* __restrict__ dbias, int B, int T, int D) {
    const int t = blockIdx.x;
    const int per = (B + gridDim.y - 1) / gridDim.y;
    const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
    if ((D & 3) == 0) {
        for (int d4 = threadIdx.x; d4 < (D >> 2); d4 += blockDim.x) {
            float4 acc[4] = {};
            int b = b0;
            for (; b + 4 <= b1; b += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 v = __ldcs(reinterpret_cast<const float4*>(g0 + (static_cast<size_t>(b + u) * T + t) * D) + d4);
                    acc[u].x += v.x;
                    acc[u].y += v.y;
                    acc[u].z += v.z;
                    acc[u].w += v.w;
                }
            }
            for (; b < b1; ++b) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(g0 + (static_cast<size_t>(b) * T + t) * D) + d4);
                acc[0].x += v.x;
                acc[0].y += v.y;
                acc[0].z += v.z;
                acc[0].w += v.w;
            }
            const float sx[4] = {acc[0].x + acc[1].x + acc[2].x + acc[3].x, acc[0].y + acc[1].y + acc[2].y + acc[3].y,
                                 acc[0].z + acc[1].z + acc[2].z + acc[3].z, acc[0].w + acc[1].w + acc[2].w + acc[3].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = d4 * 4 + e;
                atomicAdd(&dpos[static_cast<size_t>(t) * D + d], sx[e]);
                atomicAdd(t == 0 ? &dcls[d] : &dbias[d], sx[e]);
            }
        }
    } else {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float s = 0.0f;
            for (int b = b0; b < b1; ++b) s += g0[(static_cast<size_t>(b) * T + t) * D + d];
            atomicAdd(&dpos[static_cast<size_t>(t) * D + d], s);
            atomicAdd(t == 0 ? &dcls[d] : &dbias[d], s);
        }
    }
}
int launch_embed_bwd(const float* g0, float* dpos, float* dcls, float* dbias, int B, int T, int D, cudaStream_t st) {
    int slices = (4 * 148 + T - 1) / T;  // ~4 blocks per SM
    if (slices > B) slices = B;
    if (slices < 1) slices = 1;
    embed_bwd_kernel<<<dim3(T, slices), 128, 0, st>>>(g0, dpos, dcls, dbias, B, T, D);
    SVIT_CHECK_LAUNCH("embed_bwd");
    return 0;
}

__global__ void unpermute_patch_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int C, int V, int Kp) {
    const int d = blockIdx.x;
    const int CV = C * V;
    for (int k = threadIdx.x; k < CV; k += blockDim.x) {
        const int c = k / V, vv = k - c * V;
        dW[static_cast<size_t>(d) * CV + vv * C + c] += dWp[static_cast<size_t>(d) * Kp + k];
    }
}
int launch_unpermute_patch_wgrad(const float* dWp, float* dW, int D, int C, int V, int Kp, cudaStream_t st) {
    unpermute_patch_wgrad_kernel<<<D, 256, 0, st>>>(dWp, dW, C, V, Kp);
    SVIT_CHECK_LAUNCH("unpermute_patch_wgrad");
    return 0;
}

// =================================================================================================
// a9: MPP loss
// =================================================================================================
// A block of 128 threads per group of patch rows (grid-stride; unmasked rows cost one byte): the 612 elements of a row are
// spread over the block (every load of the row in flight at once), one block reduction and ONE atomic per block at the
// end.  (89 us for 217 MB, like round 1's block per row with 40,960 atomics on one address, 92 us: the 612-byte runs of x and
// the 2,448-byte rows of y at random positions bound it, not the block structure.  A warp per row, 19 dependent steps per
// lane, was slower: 165 us.)
constexpr int MPP_LOSS_MAXIT = 8;   // C * V <= 128 * 8 takes the unrolled path
__global__ void __launch_bounds__(128) mpp_loss_fwd_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ x,
                                                           const uint8_t* __restrict__ mask, float* __restrict__ loss_sum,
                                                           int BN, int C, int N, int V) {
    __shared__ float scratch[32];
    const int T = N + 1;
    const int CV = C * V;
    float s = 0.0f;
    for (int bn = blockIdx.x; bn < BN; bn += gridDim.x) {
        if (mask[bn] == 0) continue;
        const int b = bn / N, n = bn - b * N;
        const float* yr = y + (static_cast<size_t>(b) * T + 1 + n) * ldy;
        const float* xb = x + (static_cast<size_t>(b) * C * N + n) * V;   // + c * N * V + v
        if (CV <= 128 * MPP_LOSS_MAXIT) {
            float yv[MPP_LOSS_MAXIT], xv[MPP_LOSS_MAXIT];
#pragma unroll
            for (int i = 0; i < MPP_LOSS_MAXIT; ++i) {
                const int k = threadIdx.x + i * 128;
                yv[i] = 0.0f;
                xv[i] = 0.0f;
                if (k < CV) {
                    const int v = k / C, c = k - v * C;
                    yv[i] = __ldcs(yr + k);
                    xv[i] = __ldg(xb + static_cast<size_t>(c) * N * V + v);
                }
            }
#pragma unroll
            for (int i = 0; i < MPP_LOSS_MAXIT; ++i) {
                const float dlt = yv[i] - xv[i];
                s = fmaf(dlt, dlt, s);
            }
        } else {
            for (int k = threadIdx.x; k < CV; k += 128) {
                const int v = k / C, c = k - v * C;
                const float dlt = yr[k] - xb[static_cast<size_t>(c) * N * V + v];
                s = fmaf(dlt, dlt, s);
            }
        }
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0 && s != 0.0f) atomicAdd(loss_sum, s);
}
int launch_mpp_loss_fwd(const float* y, int ldy, const float* x, const uint8_t* mask, float* loss_sum, int B, int C, int N,
                        int V, cudaStream_t st) {
    if (B <= 0) return 0;
    const int BN = B * N;
    int blocks = BN < 148 * 16 ? BN : 148 * 16;
    mpp_loss_fwd_kernel<<<blocks, 128, 0, st>>>(y, ldy, x, mask, loss_sum, BN, C, N, V);
    SVIT_CHECK_LAUNCH("mpp_loss_fwd");
    return 0;
}

// dy[row, :] = coef * (y - x) for masked patch rows, zeros elsewhere (cls rows, unmasked rows, pad columns), bf16: a block per
// row, a thread writes 8 consecutive outputs as one 16-byte store (113 -> 95 us at the benchmark shape; a grid-stride variant
// with several rows per block was slower, 132 us).
__global__ void mpp_loss_bwd_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ x,
                                    const uint8_t* __restrict__ mask, const float* __restrict__ coef_dev,
                                    __nv_bfloat16* __restrict__ dy, int lddy, int C, int N, int V) {
    const int T = N + 1;
    const int row = blockIdx.x;  // b*T + t
    const int b = row / T, t = row % T;
    __nv_bfloat16* dr = dy + static_cast<size_t>(row) * lddy;
    const bool on = t > 0 && mask[static_cast<size_t>(b) * N + (t - 1)] != 0;
    const float coef = *coef_dev;
    const int n = t - 1;
    const float* yr = y + static_cast<size_t>(row) * ldy;
    if ((lddy & 7) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
        for (int k0 = threadIdx.x * 8; k0 < lddy; k0 += blockDim.x * 8) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (on && k0 < C * V) {
                float val[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = k0 + e;
                    val[e] = 0.0f;
                    if (k < C * V) {
                        const int v = k / C, c = k - v * C;
                        val[e] = coef * (__ldcs(yr + k) - __ldg(x + ((static_cast<size_t>(b) * C + c) * N + n) * V + v));
                    }
                }
                o = make_uint4(pack2_bf16(val[0], val[1]), pack2_bf16(val[2], val[3]), pack2_bf16(val[4], val[5]),
                               pack2_bf16(val[6], val[7]));
            }
            *reinterpret_cast<uint4*>(dr + k0) = o;
        }
        return;
    }
    for (int k = threadIdx.x; k < lddy; k += blockDim.x) {
        float val = 0.0f;
        if (on && k < C * V) {
            const int v = k / C, c = k - v * C;
            val = coef * (yr[k] - x[((static_cast<size_t>(b) * C + c) * N + n) * V + v]);
        }
        dr[k] = __float2bfloat16(val);
    }
}
int launch_mpp_loss_bwd(const float* y, int ldy, const float* x, const uint8_t* mask, const float* coef_dev, void* dy,
                        int lddy, int B, int C, int N, int V, cudaStream_t st) {
    if (B <= 0) return 0;
    mpp_loss_bwd_kernel<<<B * (N + 1), 128, 0, st>>>(y, ldy, x, mask, coef_dev, reinterpret_cast<__nv_bfloat16*>(dy), lddy,
                                                     C, N, V);
    SVIT_CHECK_LAUNCH("mpp_loss_bwd");
    return 0;
}

// r[d] += sum over the selected patch rows of g0: a warp per row (the selection byte is read first, unselected rows
// cost nothing), the whole row in flight per warp, block-level partial sums in shared memory, one atomic per column and
// block.  blockDim = 256 (8 warps), dynamic shared memory D floats.
__global__ void __launch_bounds__(256) masked_rowsum_kernel(const float* __restrict__ g0, const uint8_t* __restrict__ sel,
                                                            float* __restrict__ r, int B, int T, int D, int rows_per_block) {
    extern __shared__ float mr_red[];  // [D]
    const int N = T - 1;
    const int i0 = blockIdx.x * rows_per_block;
    const int i1 = min(B * N, i0 + rows_per_block);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = threadIdx.x; d < D; d += blockDim.x) mr_red[d] = 0.0f;
    __syncthreads();
    for (int d0 = 0; d0 < D; d0 += 32 * 4 * 4) {   // 512 columns per pass: 4 float4 per lane
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int i = i0 + warp; i < i1; i += 8) {
            if (!sel[i]) continue;
            const int b = i / N, n = i - b * N;
            const float4* row = reinterpret_cast<const float4*>(g0 + (static_cast<size_t>(b) * T + 1 + n) * D + d0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = (k * 32 + lane) * 4;
                if (d0 + c < D) {
                    const float4 v = __ldcs(row + k * 32 + lane);
                    acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = d0 + (k * 32 + lane) * 4;
            if (c < D) {
                atomicAdd(&mr_red[c], acc[k].x); atomicAdd(&mr_red[c + 1], acc[k].y);
                atomicAdd(&mr_red[c + 2], acc[k].z); atomicAdd(&mr_red[c + 3], acc[k].w);
            }
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(&r[d], mr_red[d]);
}
// dmt[k] += sum_d r[d] W[d, k]: grid (K / 128, slices of d), partial sums added atomically
__global__ void mask_token_gemv_kernel(const float* __restrict__ r, const float* __restrict__ W, float* __restrict__ dmt,
                                       int D, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int per = (D + gridDim.y - 1) / gridDim.y;
    const int da = blockIdx.y * per, db = min(D, da + per);
    float s0 = 0.0f, s1 = 0.0f;
    int d = da;
    for (; d + 2 <= db; d += 2) {
        s0 = fmaf(r[d], W[static_cast<size_t>(d) * K + k], s0);
        s1 = fmaf(r[d + 1], W[static_cast<size_t>(d + 1) * K + k], s1);
    }
    if (d < db) s0 = fmaf(r[d], W[static_cast<size_t>(d) * K + k], s0);
    atomicAdd(&dmt[k], s0 + s1);
}
int launch_mask_token_grad(const float* g0, const uint8_t* replace_sel, const float* W, float* scratch_r, float* dmt,
                           int B, int T, int D, int K, cudaStream_t st) {
    if (B <= 0) return 0;
    cudaMemsetAsync(scratch_r, 0, D * sizeof(float), st);
    const int rpb = 64;
    if ((D % 4) != 0) {
        set_error("mask_token_grad: D=%d must be a multiple of 4", D);
        return -2;
    }
    masked_rowsum_kernel<<<(B * (T - 1) + rpb - 1) / rpb, 256, D * sizeof(float), st>>>(g0, replace_sel, scratch_r, B, T, D, rpb);
    mask_token_gemv_kernel<<<dim3((K + 127) / 128, 16), 128, 0, st>>>(scratch_r, W, dmt, D, K);
    count_launch();
    SVIT_CHECK_LAUNCH("mask_token_grad");
    return 0;
}

// =================================================================================================
// a10: optimizers over flat buffers
// =================================================================================================
constexpr int ADAM_BLOCK_ELEMS = 4096;

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, const AdamSegment* __restrict__ segs,
                             const int* __restrict__ block_map, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int decoupled, float grad_scale) {
    // block_map[2*blockIdx.x] = segment index, [2*blockIdx.x+1] = chunk index inside the segment
    const int si = block_map[2 * blockIdx.x];
    const AdamSegment sg = segs[si];
    if (!sg.active) return;
    const long long c0 = static_cast<long long>(block_map[2 * blockIdx.x + 1]) * ADAM_BLOCK_ELEMS;
    const long long c1 = min(sg.numel, c0 + ADAM_BLOCK_ELEMS);
    const float step_size = lr / sg.bias_corr1;
    const float inv_sqrt_bc2 = rsqrtf(sg.bias_corr2);
    auto update = [&](float& pj, float gj, float& mj, float& vj) {
        gj *= grad_scale;
        if (decoupled) pj *= (1.0f - lr * weight_decay);
        else gj += weight_decay * pj;
        mj = beta1 * mj + (1.0f - beta1) * gj;
        vj = beta2 * vj + (1.0f - beta2) * gj * gj;
        const float denom = sqrtf(vj) * inv_sqrt_bc2 + eps;
        pj = pj - step_size * (mj / denom);
    };
    // segment offsets and chunk starts are multiples of 64 elements: 16-byte accesses for the body, scalars for a ragged tail
    const long long c4 = c0 + ((c1 - c0) & ~3LL);
    for (long long i = c0 + 4 * threadIdx.x; i < c4; i += 4 * blockDim.x) {
        const long long j = sg.offset + i;
        float4 p4 = *reinterpret_cast<const float4*>(p + j);
        const float4 g4 = *reinterpret_cast<const float4*>(g + j);
        float4 m4 = *reinterpret_cast<const float4*>(m + j);
        float4 v4 = *reinterpret_cast<const float4*>(v + j);
        update(p4.x, g4.x, m4.x, v4.x);
        update(p4.y, g4.y, m4.y, v4.y);
        update(p4.z, g4.z, m4.z, v4.z);
        update(p4.w, g4.w, m4.w, v4.w);
        *reinterpret_cast<float4*>(m + j) = m4;
        *reinterpret_cast<float4*>(v + j) = v4;
        *reinterpret_cast<float4*>(p + j) = p4;
    }
    for (long long i = c4 + threadIdx.x; i < c1; i += blockDim.x) {
        const long long j = sg.offset + i;
        float pj = p[j], mj = m[j], vj = v[j];
        update(pj, g[j], mj, vj);
        m[j] = mj;
        v[j] = vj;
        p[j] = pj;
    }
}

__global__ void adamw_advance_kernel(AdamSegment* __restrict__ segs, int nsegs, float beta1, float beta2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsegs || !segs[i].active) return;
    const int k = segs[i].step + 1;
    segs[i].step = k;
    segs[i].bias_corr1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(k)));
    segs[i].bias_corr2 = static_cast<float>(1.0 - pow(static_cast<double>(beta2), static_cast<double>(k)));
}
int launch_adamw_advance(AdamSegment* segs_dev, int nsegs, float beta1, float beta2, cudaStream_t st) {
    if (nsegs <= 0) return 0;
    adamw_advance_kernel<<<(nsegs + 127) / 128, 128, 0, st>>>(segs_dev, nsegs, beta1, beta2);
    SVIT_CHECK_LAUNCH("adamw_advance");
    return 0;
}

int launch_adamw(float* p, const float* g, float* m, float* v, const AdamSegment* segs_dev, int nsegs,
                 const int* block_map_dev, int nblocks, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int decoupled, float grad_scale, cudaStream_t st) {
    if (nblocks <= 0 || nsegs <= 0) return 0;
    adamw_kernel<<<nblocks, 256, 0, st>>>(p, g, m, v, segs_dev, block_map_dev, lr, beta1, beta2, eps, weight_decay,
                                          decoupled, grad_scale);
    SVIT_CHECK_LAUNCH("adamw");
    return 0;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, long long n,
                           float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                           float grad_scale) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * grad_scale + weight_decay * p[i];
        if (momentum != 0.0f) {
            const float b = first_step ? gi : momentum * mom[i] + (1.0f - dampening) * gi;
            mom[i] = b;
            gi = nesterov ? gi + momentum * b : b;
        }
        p[i] -= lr * gi;
    }
}
int launch_sgd(float* p, const float* g, float* mom, long long n, float lr, float momentum, float dampening,
               float weight_decay, int nesterov, int first_step, float grad_scale, cudaStream_t st) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    sgd_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(p, g, mom, n, lr, momentum, dampening, weight_decay, nesterov,
                                                         first_step, grad_scale);
    SVIT_CHECK_LAUNCH("sgd");
    return 0;
}


// ---------------------------------------------------------------------------------------------
// regression criterion of tools/train.py:245-248, 288: nn.MSELoss(reduction='mean') or nn.L1Loss() on outputs.squeeze()
// vs targets.  ONE launch produces the scalar loss and d loss / d out (so the backward pass of the criterion is free);
// one block, fixed-order tree reduction: deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void regression_loss_kernel(const float* __restrict__ out, const float* __restrict__ target, int n, int l1,
                                       float* __restrict__ loss, float* __restrict__ dout) {
    __shared__ float red[256];
    const float inv_n = 1.0f / static_cast<float>(n);
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = out[i] - target[i];
        if (l1) {
            acc += fabsf(d);
            dout[i] = (d > 0.0f ? inv_n : (d < 0.0f ? -inv_n : 0.0f));   // torch: sign(0) = 0
        } else {
            acc += d * d;
            dout[i] = 2.0f * d * inv_n;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = red[0] * inv_n;
}
int launch_regression_loss(const float* out, const float* target, int n, int l1, float* loss, float* dout, cudaStream_t st) {
    if (n <= 0) {
        set_error("regression_loss: empty batch");
        return -1;
    }
    regression_loss_kernel<<<1, 256, 0, st>>>(out, target, n, l1, loss, dout);
    SVIT_CHECK_LAUNCH("regression_loss");
    return 0;
}

}  // namespace svit
