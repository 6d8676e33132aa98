// gemm.cu -- warp-specialised TMA -> smem -> tcgen05.mma -> TMEM -> epilogue GEMM kernels for sm_100a.
//
// Hot path reference (what these kernels replace): every nn.Linear of the SiT encoder
// (/root/reference/models/sit.py:45-64 and the vit_pytorch Transformer it constructs at sit.py:57),
// forward and backward.  See gemm.cuh for the two entry points.
#include "gemm.cuh"

#include <cstdarg>
#include <cstdio>

#include <cuda_bf16.h>
#include <cstring>
#include <cstdlib>
#include "ptx.cuh"
#include "tma.h"

namespace svit {

static thread_local char g_last_error[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_last_error; }
static unsigned long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
unsigned long long launch_count() { return g_launches; }

// ------------------------------------------------------------------------------------------------
// exact (erf) GELU, as nn.GELU() in the reference FeedForward.
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 resolution and below the fp32
// check-mode tolerance): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1/(1 + p z), z >= 0.
// One MUFU.RCP + one MUFU.EX2 per element; the exponential exp(-u^2/2) is shared with the Gaussian pdf
// needed by the derivative.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// returns Phi(u) = 0.5 (1 + erf(u / sqrt 2)) and E = exp(-u^2 / 2)
__device__ __forceinline__ float gauss_cdf(float u, float& E) {
    const float au = fabsf(u);
    const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752f, au, 1.0f));
    E = fast_ex2(u * u * -0.72134752044448170f);  // -0.5 * log2(e)
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_tail = 0.5f * poly * t * E;          // 0.5 * (1 - erf(|u|/sqrt2))
    return u >= 0.0f ? 1.0f - half_tail : half_tail;
}
__device__ __forceinline__ float gelu_f(float x) {
    float E;
    return x * gauss_cdf(x, E);
}
__device__ __forceinline__ float dgelu_f(float x) {
    float E;
    const float cdf = gauss_cdf(x, E);
    return fmaf(x * 0.3989422804014327f, E, cdf);
}

// ------------------------------------------------------------------------------------------------
// TN GEMM
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int WBUF_BYTES = 32 * 128;        // one epilogue-warp staging box: 32 rows x 128 B (128B-swizzled)
constexpr int TN_THREADS = 384;             // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int SMEM_LIMIT = 232448;          // 227 KB

struct TnArgs {
    CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
    const float* bias;
    const float* rowtab;
    int rowtab_period;
    int M, N, K;
};

template <int MODE>
struct EpiTraits {
    static constexpr bool HAS_AUX = (MODE == EPI_RESID || MODE == EPI_DGELU);
    // staging boxes per epilogue warp: in-place aux/out rotation of 3, two outputs double-buffered, or one output x2
    static constexpr int NBUF = HAS_AUX ? 3 : (MODE == EPI_GELU ? 4 : 2);
};

template <int BN, int CG, int MODE>
struct TnCfg {
    static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;  // with a CTA pair every CTA holds half of the B rows
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * EpiTraits<MODE>::NBUF * WBUF_BYTES;
    static constexpr int FIXED_BYTES = 1024 /*align slack*/ + EPI_BYTES + EPI_WARPS * 256 /*bias*/ + 1024 /*barriers*/;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static_assert(STAGES >= 2, "not enough shared memory for the operand pipeline");
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = FIXED_BYTES + STAGES * STAGE_BYTES;
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BN tile --
// each CTA stages its own 128 rows of A and half of the B rows, the leader issues M=256 MMAs that read both CTAs'
// shared memory and write both CTAs' TMEM, and every CTA runs the epilogue for its own 128 rows (halves the B
// traffic per MAC; the L2 -> SM path bounds these small-K GEMMs).
//
// Epilogue: 8 warps, fully independent of each other (no block-level barrier).  Warp (q, p) owns TMEM lane
// quadrant q (32 rows) and every second 128-byte output unit (parity p); per unit it reads its accumulator slice
// (tcgen05.ld), applies the fused epilogue, writes a 32 x 128 B swizzled box into its private staging ring and
// issues its own TMA store.  Residual / pre-activation tiles are prefetched two units ahead into the same ring by
// the aux-loader warp and overwritten in place by the result.
template <typename OutT, int MODE, int BN, int CG>
__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tn_kernel(const __grid_constant__ TnArgs args) {
    using Cfg = TnCfg<BN, CG, MODE>;
    using ET = EpiTraits<MODE>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int UC = 128 / (int)sizeof(OutT);  // columns per epilogue unit (one 128-byte row of the staging box)
    constexpr int UNITS = BN / UC;
    static_assert(BN % UC == 0, "BN must be a multiple of the epilogue unit");
    constexpr bool HAS_AUX = ET::HAS_AUX;
    constexpr int NBUF = ET::NBUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES;                       // [EPI_WARPS][NBUF][4 KB]
    float* sBias = reinterpret_cast<float*>(sEpi + Cfg::EPI_BYTES);         // [EPI_WARPS][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + EPI_WARPS * 64);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]
    uint64_t* afull_bar = tempty_bar + 2;       // [EPI_WARPS][3]
    uint64_t* aempty_bar = afull_bar + EPI_WARPS * 3;  // [EPI_WARPS][3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + EPI_WARPS * 3);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int M = args.M, N = args.N, K = args.K;
    constexpr int TM = BM * CG;  // rows of the (pair) tile
    const int tiles_m = (M + TM - 1) / TM;
    const int tiles_n = (N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (K + BK - 1) / BK;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0;
    const int first_tile = blockIdx.x / CG;      // cluster index
    const int tile_stride = gridDim.x / CG;      // number of clusters
    const int row_off = static_cast<int>(cta_rank) * BM;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmA);
        tma_prefetch_desc(&args.tmB);
        tma_prefetch_desc(&args.tmOut);
        if (MODE == EPI_GELU) tma_prefetch_desc(&args.tmOut2);
        if (HAS_AUX) tma_prefetch_desc(&args.tmAux);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], EPI_WARPS * CG);  // the leader's barrier also collects the peer's epilogue warps
        }
        for (int i = 0; i < EPI_WARPS * 3; ++i) {
            mbar_init(&afull_bar[i], 1);
            mbar_init(&aempty_bar[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 3) {
        if constexpr (CG == 2) {
            tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
                const int m0 = (tile / tiles_n) * TM + row_off;
                const int n0 = (tile % tiles_n) * BN + static_cast<int>(cta_rank) * (BN / CG);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        // both CTAs' bytes are credited to the leader's barrier, which the leader arms for the pair
                        if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_2d_2cta(sA + stage * A_STAGE_BYTES, &args.tmA, &full_bar[stage], kb * BK, m0);
                        tma_load_2d_2cta(sB + stage * Cfg::B_STAGE_BYTES, &args.tmB, &full_bar[stage], kb * BK, n0);
                    } else {
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(sA + stage * A_STAGE_BYTES, &args.tmA, &full_bar[stage], kb * BK, m0);
                        tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &args.tmB, &full_bar[stage], kb * BK, n0);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
                        const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
                        if constexpr (CG == 2) umma_ss_2cta(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (CG == 2) umma_commit_2cta(&empty_bar[stage], 3);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if constexpr (CG == 2) umma_commit_2cta(&tfull_bar[as], 3);
                else umma_commit(&tfull_bar[as]);
            }
        }
    } else if (warp == 2) {
        // ===================== aux (residual / pre-activation) loader =====================
        // serves the per-warp rings in job order; job k of a warp lands in ring slot k % 3
        if (HAS_AUX && elect_one()) {
            int kcnt[2] = {0, 0};  // jobs issued so far per parity group
            int it = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
                const int m0 = (tile / tiles_n) * TM + row_off;
                const int n0 = (tile % tiles_n) * BN;
                for (int u = 0; u < UNITS; ++u) {
                    const int p = (it * UNITS + u) & 1;
                    const int k = kcnt[p]++;
                    const int slot = k % 3;
                    const uint32_t ph = (k / 3) & 1;
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        const int w = p * 4 + q;
                        mbar_wait(&aempty_bar[w * 3 + slot], ph ^ 1);
                        mbar_expect_tx(&afull_bar[w * 3 + slot], WBUF_BYTES);
                        tma_load_2d(sEpi + (w * NBUF + slot) * WBUF_BYTES, &args.tmAux, &afull_bar[w * 3 + slot], n0 + u * UC,
                                    m0 + q * 32);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 independent warps =====================
        const int ew = warp - 4;       // 0..7
        const int q = ew & 3;          // TMEM lane quadrant == 32-row slice of the tile
        const int p = ew >> 2;         // unit parity owned by this warp
        uint8_t* wbuf = sEpi + ew * NBUF * WBUF_BYTES;
        float* wbias = sBias + ew * 64;
        const int sw = lane & 7;       // swizzle key of this thread's row inside a 32-row box
        int it = 0, k = 0;             // k = jobs done by this warp
        for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
            const int m0 = (tile / tiles_n) * TM + row_off + q * 32;  // first row of this warp's slice
            const int n0 = (tile % tiles_n) * BN;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const uint32_t t_addr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
            const int grow = m0 + lane;
            // units of this tile owned by this warp: u with ((it*UNITS + u) & 1) == p
            const int u_first = ((it * UNITS) & 1) == p ? 0 : 1;
            int u_last = -1;
            for (int u = u_first; u < UNITS; u += 2) u_last = u;
            bool waited = false;
#pragma unroll 1
            for (int u = u_first; u < UNITS; u += 2, ++k) {
                const int c0 = n0 + u * UC;
                // bias slice of this unit -> warp-private shared memory (broadcast reads below)
                __syncwarp();
                for (int i = lane; i < UC; i += 32) wbias[i] = (args.bias != nullptr && c0 + i < N) ? args.bias[c0 + i] : 0.0f;
                __syncwarp();
                if (!waited) {
                    mbar_wait(&tfull_bar[as], aphase);
                    tc_fence_after();
                    waited = true;
                }
                float v[UC];
#pragma unroll
                for (int hh = 0; hh < UC / 32; ++hh) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + u * UC + hh * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[hh * 32 + j] = __uint_as_float(r[j]) + wbias[hh * 32 + j];
                }
                if (u == u_last) {
                    // this warp's slice of the accumulator is fully read: hand the TMEM stage back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(&tempty_bar[as], 0);
                        else mbar_arrive(&tempty_bar[as]);
                    }
                }
                if (MODE == EPI_STORE && args.rowtab != nullptr) {
                    const float* tr = args.rowtab + static_cast<size_t>(grow % args.rowtab_period) * N + c0;
#pragma unroll
                    for (int j = 0; j < UC; j += 4) {
                        if (c0 + j < N) {
                            const float4 t4 = *reinterpret_cast<const float4*>(tr + j);
                            v[j] += t4.x;
                            v[j + 1] += t4.y;
                            v[j + 2] += t4.z;
                            v[j + 3] += t4.w;
                        }
                    }
                }
                // ---- pick the staging box(es) of this job ----
                uint8_t* obuf;
                uint8_t* obuf2 = nullptr;
                if constexpr (HAS_AUX) {
                    const int slot = k % 3;
                    obuf = wbuf + slot * WBUF_BYTES;
                    mbar_wait(&afull_bar[ew * 3 + slot], (k / 3) & 1);
                    const uint8_t* arow = obuf + lane * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 a4 = *reinterpret_cast<const uint4*>(arow + ((c ^ sw) << 4));
                        if (sizeof(OutT) == 4) {
                            v[(c * 4 + 0) % UC] += __uint_as_float(a4.x);
                            v[(c * 4 + 1) % UC] += __uint_as_float(a4.y);
                            v[(c * 4 + 2) % UC] += __uint_as_float(a4.z);
                            v[(c * 4 + 3) % UC] += __uint_as_float(a4.w);
                        } else {
                            const uint32_t w4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float lo = bf16_lo(w4[e]), hi = bf16_hi(w4[e]);
                                if (MODE == EPI_DGELU) {
                                    v[(c * 8 + e * 2) % UC] *= dgelu_f(lo);
                                    v[(c * 8 + e * 2 + 1) % UC] *= dgelu_f(hi);
                                } else {
                                    v[(c * 8 + e * 2) % UC] += lo;
                                    v[(c * 8 + e * 2 + 1) % UC] += hi;
                                }
                            }
                        }
                    }
                    __syncwarp();  // every lane has consumed the aux box before it is overwritten in place
                } else if constexpr (MODE == EPI_GELU) {
                    obuf = wbuf + (k & 1) * 2 * WBUF_BYTES;
                    obuf2 = obuf + WBUF_BYTES;
                    if (lane == 0) tma_store_wait_read<1>();  // the store that used this pair two jobs ago has drained
                    __syncwarp();
                } else {
                    obuf = wbuf + (k & 1) * WBUF_BYTES;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
                uint8_t* orow = obuf + lane * 128;
                if (sizeof(OutT) == 4) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4 o;
                        o.x = __float_as_uint(v[(c * 4 + 0) % UC]);
                        o.y = __float_as_uint(v[(c * 4 + 1) % UC]);
                        o.z = __float_as_uint(v[(c * 4 + 2) % UC]);
                        o.w = __float_as_uint(v[(c * 4 + 3) % UC]);
                        *reinterpret_cast<uint4*>(orow + ((c ^ sw) << 4)) = o;
                    }
                } else {
                    if (MODE != EPI_GELU_ONLY) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint4 o;
                            o.x = pack_bf16(v[(c * 8 + 0) % UC], v[(c * 8 + 1) % UC]);
                            o.y = pack_bf16(v[(c * 8 + 2) % UC], v[(c * 8 + 3) % UC]);
                            o.z = pack_bf16(v[(c * 8 + 4) % UC], v[(c * 8 + 5) % UC]);
                            o.w = pack_bf16(v[(c * 8 + 6) % UC], v[(c * 8 + 7) % UC]);
                            *reinterpret_cast<uint4*>(orow + ((c ^ sw) << 4)) = o;
                        }
                    }
                    if (MODE == EPI_GELU || MODE == EPI_GELU_ONLY) {
                        uint8_t* orow2 = (MODE == EPI_GELU) ? obuf2 + lane * 128 : orow;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float g[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) g[e] = gelu_f(v[(c * 8 + e) % UC]);
                            uint4 o;
                            o.x = pack_bf16(g[0], g[1]);
                            o.y = pack_bf16(g[2], g[3]);
                            o.z = pack_bf16(g[4], g[5]);
                            o.w = pack_bf16(g[6], g[7]);
                            *reinterpret_cast<uint4*>(orow2 + ((c ^ sw) << 4)) = o;
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&args.tmOut, obuf, c0, m0);
                    if (MODE == EPI_GELU) tma_store_2d(&args.tmOut2, obuf2, c0, m0);
                    tma_store_commit();
                    if constexpr (HAS_AUX) {
                        // the store of the previous job has finished reading its box: give that slot back to the
                        // aux loader (it then prefetches the tile of job k+2 into it)
                        if (k > 0) {
                            tma_store_wait_read<1>();
                            mbar_arrive(&aempty_bar[ew * 3 + (k - 1) % 3]);
                        }
                    }
                }
            }
            if (u_last < 0) {
                // no unit of this tile belongs to this warp (UNITS == 1): still release the accumulator stage
                if (!waited) mbar_wait(&tfull_bar[as], aphase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(&tempty_bar[as], 0);
                    else mbar_arrive(&tempty_bar[as]);
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-gradient GEMM: dW[N,K] += dY[M,N]^T X[M,K]
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;
constexpr int WG_BN = 192;  // tile over the K (input-feature) axis of dW
constexpr int WG_STAGES = 5;
constexpr int WG_A_BYTES = 64 * 128 * 2;    // 64 tokens x 128 out-features
constexpr int WG_B_BYTES = 64 * WG_BN * 2;  // 64 tokens x 192 in-features
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;

struct WgArgs {
    CUtensorMap tmY, tmX;
    float* dW;
    int M, N, K, ldw;
    int kb_per_split;
};

__global__ void __launch_bounds__(WG_THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* done_bar = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_k = (args.K + WG_BN - 1) / WG_BN;
    const int n0 = (blockIdx.x / tiles_k) * 128;
    const int k0 = (blockIdx.x % tiles_k) * WG_BN;
    const int total_kb = (args.M + 63) / 64;
    const int kb_begin = blockIdx.y * args.kb_per_split;
    const int kb_end = min(total_kb, kb_begin + args.kb_per_split);
    const int num_kb = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmY);
        tma_prefetch_desc(&args.tmX);
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            if (elect_one()) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
                    uint8_t* a = smem + stage * WG_STAGE_BYTES;
                    uint8_t* b = a + WG_A_BYTES;
#pragma unroll
                    for (int j = 0; j < 2; ++j) tma_load_2d(a + j * 8192, &args.tmY, &full_bar[stage], n0 + j * 64, kb * 64);
#pragma unroll
                    for (int j = 0; j < WG_BN / 64; ++j)
                        tma_load_2d(b + j * 8192, &args.tmX, &full_bar[stage], k0 + j * 64, kb * 64);
                    if (++stage == WG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                constexpr uint32_t idesc = umma_idesc_bf16(128, WG_BN, 1, 1);
                int stage = 0;
                uint32_t phase = 0;
                for (int i = 0; i < num_kb; ++i) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * WG_STAGE_BYTES);
                    const uint32_t b_addr = a_addr + WG_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // MN-major, 128B swizzle: 64-element MN blocks 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO)
                        const uint64_t ad = umma_smem_desc(a_addr + k * 2048, 8192, 1024);
                        const uint64_t bd = umma_smem_desc(b_addr + k * 2048, 8192, 1024);
                        umma_ss(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == WG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(done_bar);
            }
        } else {
            // epilogue: warps 2..5 -> TMEM lane quadrants 2,3,0,1
            const int q = warp & 3;
            const int row = q * 32 + lane;
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const int n = n0 + row;
            float* dst = args.dW + static_cast<size_t>(n) * args.ldw + k0;
#pragma unroll 1
            for (int c0 = 0; c0 < WG_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + c0 + (static_cast<uint32_t>(q * 32) << 16), r);
                tmem_ld_wait();
                if (n < args.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (k0 + c0 + j < args.K) atomicAdd(dst + c0 + j, __uint_as_float(r[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename OutT, int MODE, int BN, int CG>
static int launch_tn_inst(const TnArgs& a, int num_sms, cudaStream_t stream) {
    using Cfg = TnCfg<BN, CG, MODE>;
    auto kfn = gemm_tn_kernel<OutT, MODE, BN, CG>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_tn) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    const int tiles = ((a.M + BM * CG - 1) / (BM * CG)) * ((a.N + BN - 1) / BN);
    const int max_groups = num_sms / CG;
    const int groups = tiles < max_groups ? tiles : max_groups;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(groups * CG);
    cfg.blockDim = dim3(TN_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
    if (e != cudaSuccess) {
        set_error("gemm_tn launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

template <int CG>
static int dispatch_tn(const GemmTnDesc& d, const TnArgs& a, int num_sms, cudaStream_t stream) {
    constexpr int BN = 192;
    if (d.out_f32) {
        if (d.mode == EPI_STORE) return launch_tn_inst<float, EPI_STORE, BN, CG>(a, num_sms, stream);
        if (d.mode == EPI_RESID) return launch_tn_inst<float, EPI_RESID, BN, CG>(a, num_sms, stream);
    } else {
        if (d.mode == EPI_STORE) return launch_tn_inst<__nv_bfloat16, EPI_STORE, BN, CG>(a, num_sms, stream);
        if (d.mode == EPI_GELU) return launch_tn_inst<__nv_bfloat16, EPI_GELU, BN, CG>(a, num_sms, stream);
        if (d.mode == EPI_RESID) return launch_tn_inst<__nv_bfloat16, EPI_RESID, BN, CG>(a, num_sms, stream);
        if (d.mode == EPI_DGELU) return launch_tn_inst<__nv_bfloat16, EPI_DGELU, BN, CG>(a, num_sms, stream);
        if (d.mode == EPI_GELU_ONLY) return launch_tn_inst<__nv_bfloat16, EPI_GELU_ONLY, BN, CG>(a, num_sms, stream);
    }
    set_error("gemm_tn: unsupported mode %d for out_f32=%d", d.mode, d.out_f32);
    return -4;
}

int launch_gemm_tn(const GemmTnDesc& d, int num_sms, cudaStream_t stream) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) {
        set_error("gemm_tn: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
        return -1;
    }
    if ((d.lda % 8) || (d.ldb % 8) || (d.ldo % (d.out_f32 ? 4 : 8)) || (d.N % 4)) {
        set_error("gemm_tn: pitches must be 16-byte multiples (lda=%d ldb=%d ldo=%d N=%d)", d.lda, d.ldb, d.ldo, d.N);
        return -2;
    }
    constexpr int BN = 192;
    // CTA pairs (256-row tiles) once the problem has at least one full wave of pair tiles
    static const int force_cg = getenv("SVIT_GEMM_CG") ? atoi(getenv("SVIT_GEMM_CG")) : 0;
    const int pair_tiles = ((d.M + 255) / 256) * ((d.N + BN - 1) / BN);
    int cg = (pair_tiles >= num_sms / 2) ? 2 : 1;
    if (force_cg == 1 || force_cg == 2) cg = force_cg;
    TnArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = d.bias;
    a.rowtab = d.rowtab;
    a.rowtab_period = d.rowtab_period > 0 ? d.rowtab_period : 1;
    a.M = d.M;
    a.N = d.N;
    a.K = d.K;
    int rc = 0;
    rc |= make_tmap_2d(&a.tmA, d.A, TmapDtype::BF16, d.K, d.M, (uint64_t)d.lda * 2, BK, BM);
    rc |= make_tmap_2d(&a.tmB, d.B, TmapDtype::BF16, d.K, d.N, (uint64_t)d.ldb * 2, BK, BN / cg);
    const TmapDtype odt = d.out_f32 ? TmapDtype::F32 : TmapDtype::BF16;
    const int osz = d.out_f32 ? 4 : 2;
    rc |= make_tmap_2d(&a.tmOut, d.out, odt, d.N, d.M, (uint64_t)d.ldo * osz, 128 / osz, 32);
    if (d.mode == EPI_GELU) rc |= make_tmap_2d(&a.tmOut2, d.out2, odt, d.N, d.M, (uint64_t)d.ldo * osz, 128 / osz, 32);
    if (d.mode == EPI_RESID || d.mode == EPI_DGELU)
        rc |= make_tmap_2d(&a.tmAux, d.aux, odt, d.N, d.M, (uint64_t)d.ldo * osz, 128 / osz, 32);
    if (rc != 0) {
        set_error("gemm_tn: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    return cg == 2 ? dispatch_tn<2>(d, a, num_sms, stream) : dispatch_tn<1>(d, a, num_sms, stream);
}

int launch_gemm_wgrad(const GemmWgradDesc& d, int num_sms, cudaStream_t stream) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) {
        set_error("gemm_wgrad: empty problem");
        return -1;
    }
    if ((d.ldy % 8) || (d.ldx % 8)) {
        set_error("gemm_wgrad: pitches must be 16-byte multiples (ldy=%d ldx=%d)", d.ldy, d.ldx);
        return -2;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_wgrad) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    WgArgs a;
    memset(&a, 0, sizeof(a));
    int rc = 0;
    rc |= make_tmap_2d(&a.tmY, d.dY, TmapDtype::BF16, d.N, d.M, (uint64_t)d.ldy * 2, 64, 64);
    rc |= make_tmap_2d(&a.tmX, d.X, TmapDtype::BF16, d.K, d.M, (uint64_t)d.ldx * 2, 64, 64);
    if (rc != 0) {
        set_error("gemm_wgrad: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.dW = d.dW;
    a.M = d.M;
    a.N = d.N;
    a.K = d.K;
    a.ldw = d.ldw;
    const int tiles = ((d.N + 127) / 128) * ((d.K + WG_BN - 1) / WG_BN);
    const int total_kb = (d.M + 63) / 64;
    int splits = (2 * num_sms + tiles - 1) / tiles;
    if (splits > total_kb) splits = total_kb;
    if (splits < 1) splits = 1;
    a.kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + a.kb_per_split - 1) / a.kb_per_split;
    dim3 grid(tiles, splits);
    gemm_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, stream>>>(a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("gemm_wgrad launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    return 0;
}

}  // namespace svit
