// gemm.cu -- warp-specialised TMA -> smem -> tcgen05.mma -> TMEM -> epilogue GEMM kernels for sm_100a.
//
// Hot path reference (what these kernels replace): every nn.Linear of the SiT encoder
// (/root/reference/models/sit.py:45-64 and the vit_pytorch Transformer it constructs at sit.py:57),
// forward and backward.  See gemm.cuh for the two entry points.
#include "gemm.cuh"

#include <atomic>
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
// incremented from the forward and the autograd (backward) host threads
bool pdl_enabled() {
    static const bool on = !(getenv("SVIT_NO_PDL") != nullptr && atoi(getenv("SVIT_NO_PDL")) != 0);
    return on;
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// erf GELU, as nn.GELU() in the reference FeedForward.
// erf via Abramowitz-Stegun 7.1.25: erf(z) = 1 - (a1 t + a2 t^2 + a3 t^3) exp(-z^2), t = 1/(1 + p z), z >= 0, |abs err| <=
// 2.5e-5 -- 1.3e-5 in Phi, two orders of magnitude below the bf16 resolution of the outputs this epilogue writes (the fp32
// check mode has its own erff() path).  The epilogue is bound by its instruction count (74 % of the issue slots with 16
// warps), so the three-term form replaces the five-term 7.1.26 of round 1: 16 instead of 19 instructions per element.
// One MUFU.RCP + one MUFU.EX2 per element; the exponential exp(-u^2/2) is shared with the Gaussian pdf needed by the
// derivative; the factor 0.5 of the tail is folded into the coefficients.
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
    const float t = fast_rcp(fmaf(0.47047f * 0.70710678118654752f, au, 1.0f));
    E = fast_ex2(u * u * -0.72134752044448170f);  // -0.5 * log2(e)
    float poly = fmaf(t, 0.5f * 0.7478556f, 0.5f * -0.0958798f);
    poly = fmaf(poly, t, 0.5f * 0.3480242f);
    const float half_tail = (poly * t) * E;               // 0.5 * (1 - erf(|u|/sqrt2))
    return u >= 0.0f ? 1.0f - half_tail : half_tail;
}
__device__ __forceinline__ float gelu_f(float x) {
    float E;
    return x * gauss_cdf(x, E);
}
// gelu(x) and gelu'(x) from one evaluation of Phi and the Gaussian
__device__ __forceinline__ float gelu_and_grad(float x, float& grad) {
    float E;
    const float cdf = gauss_cdf(x, E);
    grad = fmaf(x * 0.3989422804014327f, E, cdf);
    return x * cdf;
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
#ifndef SVIT_EPI_WARPS
#define SVIT_EPI_WARPS 8
#endif
constexpr int SMEM_LIMIT = 232448;          // 227 KB

struct TnArgs {
    CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
    const float* bias;
    const float* rowtab;
    int rowtab_period;
    int M, N, K;
};

constexpr int ARES_KB = 6;  // A-resident mode: up to 6 K blocks (K <= 384) of the A row block stay in shared memory

// Epilogue geometry.  EW epilogue warps (a multiple of 4: EW / 4 warps per TMEM lane quadrant); a warp's staging box is
// 32 rows x BOXB bytes.  The GELU epilogues (~20 instructions per element, two MUFU among them) are bound by the issue
// rate of their warps: with 8 warps (2 per SM sub-partition) dependent-issue latency caps them at ~57 % of the issue
// slots, so those modes run 16 warps (4 per sub-partition) on 64-byte boxes -- half-width units, the same 64 KB of
// staging, the same operand ring depth.
template <int MODE, bool ARES>
struct EpiTraits {
    static constexpr bool HAS_AUX = (MODE == EPI_RESID || MODE == EPI_DGELU || MODE == EPI_MUL);
    static constexpr bool TWO_OUT = (MODE == EPI_GELU || MODE == EPI_GELU_GRAD);
    static constexpr bool HEAVY = (MODE == EPI_GELU || MODE == EPI_GELU_GRAD || MODE == EPI_GELU_ONLY);
    // 64-byte staging boxes for the GELU epilogues (16 warps).  (Measured and rejected: the `* stored gelu'` epilogue of
    // the MLP input gradient on 64-byte boxes with a resident A block, 118.7 -> 137.8 us at M = 82176.)
    static constexpr bool SMALLBOX = HEAVY;
    static constexpr int EW = HEAVY ? 16 : SVIT_EPI_WARPS;   // epilogue warps
    static constexpr int NP = EW / 4;                          // warps per TMEM lane quadrant
    static constexpr int BOXB = SMALLBOX ? 64 : 128;           // bytes per staging-box row (= TMA swizzle span)
    static constexpr int WBUF = 32 * BOXB;                     // one staging box
    static constexpr int THREADS = 128 + 32 * EW;              // 4 control warps + the epilogue warps
    // staging boxes per epilogue warp: in-place aux/out rotation of 3 (2 for the A-resident variant with 4 KB boxes, which
    // gives 96 KB to the A row block), two outputs single-buffered, or one output x2
    static constexpr int NBUF = HAS_AUX ? ((ARES && BOXB == 128) ? 2 : 3) : 2;
};

template <int BN, int CG, int MODE, bool ARES>
struct TnCfg {
    using ET = EpiTraits<MODE, ARES>;
    static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;  // with a CTA pair every CTA holds half of the B rows
    static constexpr int EPI_BYTES = ET::EW * ET::NBUF * ET::WBUF;
    static constexpr int BAR_BYTES = ET::EW > 8 ? 2048 : 1024;
    static constexpr int FIXED_BYTES = 1024 /*align slack*/ + EPI_BYTES + ET::EW * 256 /*bias*/ + BAR_BYTES;
    // A and B share one ring (stage = A block + B block) -- or A keeps ARES_KB fixed slots and only B is a ring
    static constexpr int STAGE_BYTES = ARES ? B_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int A_SLOTS_FIXED = ARES ? ARES_KB : 0;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED_BYTES - A_SLOTS_FIXED * A_STAGE_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static_assert(STAGES >= 2, "not enough shared memory for the operand pipeline");
    static constexpr int A_SLOTS = ARES ? ARES_KB : STAGES;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = FIXED_BYTES + A_SLOTS * A_STAGE_BYTES + STAGES * B_STAGE_BYTES;
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
//
// ARES (A-resident, K <= 384): the kernel is bound by the L2 -> SM feed (~45 B/clk/SM), so when the whole K extent
// of a 128-row A block fits in shared memory (96 KB) it is loaded ONCE and every N tile of that row block streams
// only its B tile past it.  Tiles are then handed out as contiguous m-major ranges (one range per cluster) so that a
// cluster re-loads A only when it crosses into the next row block.
template <typename OutT, int MODE, int BN, int CG, bool ARES>
__global__ void __launch_bounds__((EpiTraits<MODE, ARES>::THREADS), 1) gemm_tn_kernel(const __grid_constant__ TnArgs args) {
    using Cfg = TnCfg<BN, CG, MODE, ARES>;
    using ET = EpiTraits<MODE, ARES>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_WARPS = ET::EW, EPI_NP = ET::NP, BOXB = ET::BOXB, WBUF_BYTES = ET::WBUF;
    constexpr int UC = BOXB / (int)sizeof(OutT);  // columns per epilogue unit (one row of the staging box)
    constexpr int UNITS = BN / UC;
    static_assert(BN % UC == 0, "BN must be a multiple of the epilogue unit");
    constexpr bool HAS_AUX = ET::HAS_AUX;
    constexpr int NBUF = ET::NBUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::A_SLOTS * A_STAGE_BYTES;
    uint8_t* sEpi = sB + STAGES * Cfg::B_STAGE_BYTES;                       // [EPI_WARPS][NBUF][4 KB]
    float* sBias = reinterpret_cast<float*>(sEpi + Cfg::EPI_BYTES);         // [EPI_WARPS][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + EPI_WARPS * 64);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]
    uint64_t* afull_bar = tempty_bar + 2;       // [EPI_WARPS][3]
    uint64_t* aempty_bar = afull_bar + EPI_WARPS * 3;  // [EPI_WARPS][3]
    uint64_t* ares_full = aempty_bar + EPI_WARPS * 3;  // [ARES_KB] resident A block kb landed
    uint64_t* ares_empty = ares_full + ARES_KB;        // [ARES_KB] last MMA of the row block that reads it retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ares_empty + ARES_KB);

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
    // tile schedule: round-robin over clusters, or (ARES) one contiguous m-major range per cluster
    const int num_clusters = gridDim.x / CG, cluster_id = blockIdx.x / CG;
    const int first_tile = ARES ? static_cast<int>(static_cast<long long>(cluster_id) * num_tiles / num_clusters) : cluster_id;
    const int end_tile = ARES ? static_cast<int>(static_cast<long long>(cluster_id + 1) * num_tiles / num_clusters) : num_tiles;
    const int tile_stride = ARES ? 1 : num_clusters;
    const int row_off = static_cast<int>(cta_rank) * BM;

    griddep_launch();  // the next kernel of the stream may set itself up under this one (ptx.cuh)
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmA);
        tma_prefetch_desc(&args.tmB);
        tma_prefetch_desc(&args.tmOut);
        if (ET::TWO_OUT) tma_prefetch_desc(&args.tmOut2);
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
        for (int i = 0; i < ARES_KB; ++i) {
            mbar_init(&ares_full[i], 1);
            mbar_init(&ares_empty[i], 1);
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
    griddep_wait();  // everything above overlapped the previous kernel; its results are visible from here on

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int cur_m = -1, visits = 0;  // ARES: row block whose A is resident, number of row blocks loaded so far
            for (int tile = first_tile; tile < end_tile; tile += tile_stride) {
                const int m0 = (tile / tiles_n) * TM + row_off;
                const int n0 = (tile % tiles_n) * BN + static_cast<int>(cta_rank) * (BN / CG);
                if constexpr (ARES) {
                    const bool new_m = (tile / tiles_n) != cur_m;
                    if (new_m) {
                        cur_m = tile / tiles_n;
                        ++visits;
                    }
                    for (int kb = 0; kb < num_kb; ++kb) {
                        if (new_m) {
                            mbar_wait(&ares_empty[kb], (visits & 1));  // visit v waits for completion v-2 (parity v & 1)
                            if constexpr (CG == 2) {
                                if (is_leader) mbar_expect_tx(&ares_full[kb], 2 * A_STAGE_BYTES);
                                tma_load_2d_2cta(sA + kb * A_STAGE_BYTES, &args.tmA, &ares_full[kb], kb * BK, m0);
                            } else {
                                mbar_expect_tx(&ares_full[kb], A_STAGE_BYTES);
                                tma_load_2d(sA + kb * A_STAGE_BYTES, &args.tmA, &ares_full[kb], kb * BK, m0);
                            }
                        }
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if constexpr (CG == 2) {
                            if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::B_STAGE_BYTES);
                            tma_load_2d_2cta(sB + stage * Cfg::B_STAGE_BYTES, &args.tmB, &full_bar[stage], kb * BK, n0);
                        } else {
                            mbar_expect_tx(&full_bar[stage], Cfg::B_STAGE_BYTES);
                            tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &args.tmB, &full_bar[stage], kb * BK, n0);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                } else {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        // both CTAs' bytes are credited to the leader's barrier, which the leader arms for the pair
                        if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
                        tma_load_2d_2cta(sA + stage * A_STAGE_BYTES, &args.tmA, &full_bar[stage], kb * BK, m0);
                        tma_load_2d_2cta(sB + stage * Cfg::B_STAGE_BYTES, &args.tmB, &full_bar[stage], kb * BK, n0);
                    } else {
                        mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + Cfg::B_STAGE_BYTES);
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
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            int cur_m = -1, visits = 0;
            for (int tile = first_tile; tile < end_tile; tile += tile_stride, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                bool new_m = false, last_of_m = false;
                if constexpr (ARES) {
                    new_m = (tile / tiles_n) != cur_m;
                    if (new_m) {
                        cur_m = tile / tiles_n;
                        ++visits;
                    }
                    last_of_m = (tile + 1 >= end_tile) || ((tile + 1) / tiles_n != cur_m);
                }
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (ARES && new_m) mbar_wait(&ares_full[kb], (visits & 1) ^ 1);
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + (ARES ? kb : stage) * A_STAGE_BYTES);
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
                    if (ARES && last_of_m) {  // this row block's A slot kb may be refilled
                        if constexpr (CG == 2) umma_commit_2cta(&ares_empty[kb], 3);
                        else umma_commit(&ares_empty[kb]);
                    }
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
            // jobs issued so far per warp group -- scalars, not an indexed array: a dynamically indexed local array lives
            // in local memory, and with ~227 KB of the unified L1 / shared memory taken by the tiles every access of this
            // (single, latency-critical) thread would be an L2 round trip
            static_assert(EPI_NP <= 4, "job counters are kept in four scalars");
            int kc0 = 0, kc1 = 0, kc2 = 0, kc3 = 0;
            int it = 0;
            for (int tile = first_tile; tile < end_tile; tile += tile_stride, ++it) {
                const int m0 = (tile / tiles_n) * TM + row_off;
                const int n0 = (tile % tiles_n) * BN;
                for (int u = 0; u < UNITS; ++u) {
                    const int p = (it * UNITS + u) % EPI_NP;
                    const int k = p == 0 ? kc0 : p == 1 ? kc1 : p == 2 ? kc2 : kc3;
                    if (p == 0) ++kc0; else if (p == 1) ++kc1; else if (p == 2) ++kc2; else ++kc3;
                    const int slot = k % NBUF;
                    const uint32_t ph = (k / NBUF) & 1;
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
        const int ew = warp - 4;       // 0 .. EPI_WARPS-1
        const int q = ew & 3;          // TMEM lane quadrant == 32-row slice of the tile
        const int p = ew >> 2;         // this warp owns the units u with (it * UNITS + u) % EPI_NP == p
        uint8_t* wbuf = sEpi + ew * NBUF * WBUF_BYTES;
        float* wbias = sBias + ew * 64;
        // swizzle key of this thread's row inside a 32-row box: 128-byte rows XOR the 16-byte chunk index with row % 8,
        // 64-byte rows (SWIZZLE_64B) with (row / 2) % 4
        const int sw = BOXB == 128 ? (lane & 7) : ((lane >> 1) & 3);
        constexpr int CPH = 32 * (int)sizeof(OutT) / 16;  // 16-byte chunks per 32-column half of a unit (4 bf16, 8 fp32)
        int it = 0, k = 0;             // k = jobs done by this warp
        for (int tile = first_tile; tile < end_tile; tile += tile_stride, ++it) {
            const int m0 = (tile / tiles_n) * TM + row_off + q * 32;  // first row of this warp's slice
            const int n0 = (tile % tiles_n) * BN;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const uint32_t t_addr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
            const int grow = m0 + lane;
            const int u_first = ((p - it * UNITS) % EPI_NP + EPI_NP) % EPI_NP;
            int u_last = -1;
            for (int u = u_first; u < UNITS; u += EPI_NP) u_last = u;
            bool waited = false;
#pragma unroll 1
            for (int u = u_first; u < UNITS; u += EPI_NP, ++k) {
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
                // ---- pick the staging box(es) of this job ----
                uint8_t* obuf;
                uint8_t* obuf2 = nullptr;
                if constexpr (HAS_AUX) {
                    const int slot = k % NBUF;
                    obuf = wbuf + slot * WBUF_BYTES;
                    mbar_wait(&afull_bar[ew * 3 + slot], (k / NBUF) & 1);
                } else if constexpr (ET::TWO_OUT) {
                    if constexpr (NBUF == 4) {
                        obuf = wbuf + (k & 1) * 2 * WBUF_BYTES;
                        if (lane == 0) tma_store_wait_read<1>();  // the store that used this pair two jobs ago has drained
                    } else {
                        obuf = wbuf;
                        if (lane == 0) tma_store_wait_read<0>();  // single pair of boxes: the previous store has drained
                    }
                    obuf2 = obuf + WBUF_BYTES;
                    __syncwarp();
                } else {
                    obuf = wbuf + (k & 1) * WBUF_BYTES;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
                uint8_t* orow = obuf + lane * BOXB;
                // ---- the unit in 32-column halves: TMEM -> registers -> fused epilogue -> staging box ----
#pragma unroll 1
                for (int hh = 0; hh < UC / 32; ++hh) {
                    float v[32];
                    {
                        uint32_t r[32];
                        tmem_ld_32x32(t_addr + u * UC + hh * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + wbias[hh * 32 + j];
                    }
                    if (u == u_last && hh == UC / 32 - 1) {
                        // this warp's slice of the accumulator is fully read: hand the TMEM stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (CG == 2) mbar_arrive_cluster(&tempty_bar[as], 0);
                            else mbar_arrive(&tempty_bar[as]);
                        }
                    }
                    if (MODE == EPI_STORE && args.rowtab != nullptr) {
                        const float* tr = args.rowtab + static_cast<size_t>(grow % args.rowtab_period) * N + c0 + hh * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (c0 + hh * 32 + j < N) {
                                const float4 t4 = *reinterpret_cast<const float4*>(tr + j);
                                v[j] += t4.x;
                                v[j + 1] += t4.y;
                                v[j + 2] += t4.z;
                                v[j + 3] += t4.w;
                            }
                        }
                    }
                    if constexpr (HAS_AUX) {
                        // residual / pre-activation: read this half of the aux box (each lane only touches its own row,
                        // so the result can overwrite it in place)
#pragma unroll
                        for (int c = 0; c < CPH; ++c) {
                            const uint4 a4 = *reinterpret_cast<const uint4*>(orow + (((hh * CPH + c) ^ sw) << 4));
                            if (sizeof(OutT) == 4) {
                                v[c * 4 + 0] += __uint_as_float(a4.x);
                                v[c * 4 + 1] += __uint_as_float(a4.y);
                                v[c * 4 + 2] += __uint_as_float(a4.z);
                                v[c * 4 + 3] += __uint_as_float(a4.w);
                            } else {
                                const uint32_t w4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float lo = bf16_lo(w4[e]), hi = bf16_hi(w4[e]);
                                    if (MODE == EPI_DGELU) {
                                        v[c * 8 + e * 2] *= dgelu_f(lo);
                                        v[c * 8 + e * 2 + 1] *= dgelu_f(hi);
                                    } else if (MODE == EPI_MUL) {
                                        v[c * 8 + e * 2] *= lo;
                                        v[c * 8 + e * 2 + 1] *= hi;
                                    } else {
                                        v[c * 8 + e * 2] += lo;
                                        v[c * 8 + e * 2 + 1] += hi;
                                    }
                                }
                            }
                        }
                    }
                    if (sizeof(OutT) == 4) {
#pragma unroll
                        for (int c = 0; c < CPH; ++c) {
                            uint4 o;
                            o.x = __float_as_uint(v[c * 4 + 0]);
                            o.y = __float_as_uint(v[c * 4 + 1]);
                            o.z = __float_as_uint(v[c * 4 + 2]);
                            o.w = __float_as_uint(v[c * 4 + 3]);
                            *reinterpret_cast<uint4*>(orow + (((hh * CPH + c) ^ sw) << 4)) = o;
                        }
                    } else {
                        if constexpr (MODE == EPI_GELU_GRAD) {
                            uint8_t* orow2 = obuf2 + lane * BOXB;
#pragma unroll
                            for (int c = 0; c < CPH; ++c) {
                                float g[8], dg[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) g[e] = gelu_and_grad(v[c * 8 + e], dg[e]);
                                uint4 o;
                                o.x = pack_bf16(dg[0], dg[1]);
                                o.y = pack_bf16(dg[2], dg[3]);
                                o.z = pack_bf16(dg[4], dg[5]);
                                o.w = pack_bf16(dg[6], dg[7]);
                                *reinterpret_cast<uint4*>(orow + (((hh * CPH + c) ^ sw) << 4)) = o;
                                o.x = pack_bf16(g[0], g[1]);
                                o.y = pack_bf16(g[2], g[3]);
                                o.z = pack_bf16(g[4], g[5]);
                                o.w = pack_bf16(g[6], g[7]);
                                *reinterpret_cast<uint4*>(orow2 + (((hh * CPH + c) ^ sw) << 4)) = o;
                            }
                        } else {
                        if (MODE != EPI_GELU_ONLY) {
#pragma unroll
                            for (int c = 0; c < CPH; ++c) {
                                uint4 o;
                                o.x = pack_bf16(v[c * 8 + 0], v[c * 8 + 1]);
                                o.y = pack_bf16(v[c * 8 + 2], v[c * 8 + 3]);
                                o.z = pack_bf16(v[c * 8 + 4], v[c * 8 + 5]);
                                o.w = pack_bf16(v[c * 8 + 6], v[c * 8 + 7]);
                                *reinterpret_cast<uint4*>(orow + (((hh * CPH + c) ^ sw) << 4)) = o;
                            }
                        }
                        if (MODE == EPI_GELU || MODE == EPI_GELU_ONLY) {
                            uint8_t* orow2 = (MODE == EPI_GELU) ? obuf2 + lane * BOXB : orow;
#pragma unroll
                            for (int c = 0; c < CPH; ++c) {
                                float g[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) g[e] = gelu_f(v[c * 8 + e]);
                                uint4 o;
                                o.x = pack_bf16(g[0], g[1]);
                                o.y = pack_bf16(g[2], g[3]);
                                o.z = pack_bf16(g[4], g[5]);
                                o.w = pack_bf16(g[6], g[7]);
                                *reinterpret_cast<uint4*>(orow2 + (((hh * CPH + c) ^ sw) << 4)) = o;
                            }
                        }
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&args.tmOut, obuf, c0, m0);
                    if (ET::TWO_OUT) tma_store_2d(&args.tmOut2, obuf2, c0, m0);
                    tma_store_commit();
                    if constexpr (HAS_AUX) {
                        // the store of the previous job has finished reading its box: give that slot back to the
                        // aux loader (it then prefetches the tile of a later job into it)
                        if (k > 0) {
                            tma_store_wait_read<1>();
                            mbar_arrive(&aempty_bar[ew * 3 + (k - 1) % NBUF]);
                        }
                    }
                }
            }
            if (u_last < 0) {
                // no unit of this tile belongs to this warp: still release the accumulator stage
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
// weight-gradient GEMM: R[p, q] += sum_t P[t, p] * Q[t, q]      (reduction over the token axis t)
//
// Both operands are read MN-major straight from the row-major activations (no transposes).  A CTA pair
// (cta_group::2) owns a 256 (p) x 384 (q) tile of R for one slice of the token axis: every CTA stages its own 128
// p-columns and half of the q-columns per 64-token block, the leader issues an N=256 and an N=128 MMA per 16 tokens
// (384 fp32 accumulator columns in each CTA's TMEM), and the partial tile is added to global memory with fp32
// reductions.  The big tile matters more than anything else here: the kernel is bound by the L2 -> SM feed
// (~45 B/clk/SM), and 256x384 needs 52 B/clk at full tensor rate where the former 128x192 tile needed 107.
// The bias gradient (column sums of P) rides along as one extra N=16 MMA against a constant tile of ones.
// The host picks which of dY / X plays P so that tiles are full; strides (ld_p, ld_q) make either orientation
// land in dW[N, K].
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int WG_STAGES = 5;
constexpr int WG_BOX_BYTES = 64 * 64 * 2;           // one TMA box: 64 tokens x 64 features
constexpr int WG_P_BYTES = 2 * WG_BOX_BYTES;        // 64 tokens x 128 p-features per CTA
constexpr int WG_Q_BYTES = 3 * WG_BOX_BYTES;        // 64 tokens x up to 192 q-features per CTA
constexpr int WG_STAGE_BYTES = WG_P_BYTES + WG_Q_BYTES;
constexpr int WG_ONES_BYTES = 4096;
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 256;
constexpr int WG_TP = 256;        // tile rows (p), pair-wide
constexpr int WG_TQ = 384;        // tile columns (q)
constexpr int WG_ONES_COL = 384;  // TMEM column of the bias accumulator

struct WgArgs {
    CUtensorMap tmP, tmQ;
    float* R;
    float* dbias;         // [Pdim] or nullptr: += column sums of P (added by the q-tile 0 CTAs only)
    long long ld_p, ld_q;  // R[p, q] lives at R + p * ld_p + q * ld_q
    int M, Pdim, Qdim;
    int tiles_q;
    int kb_per_split;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sOnes = smem + WG_STAGES * WG_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + WG_ONES_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* done_bar = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int tile = blockIdx.x >> 1;
    const int p0 = (tile / args.tiles_q) * WG_TP;
    const int q0 = (tile % args.tiles_q) * WG_TQ;
    const int nq = min(args.Qdim - q0, WG_TQ);  // valid q columns of this tile
    const int n_mma0 = nq > 128 ? 256 : 128;    // N of the first MMA; the second one (N = 128) covers q >= 256
    const bool has_mma1 = nq > 256;
    const int q_boxes = (n_mma0 >> 7) + (has_mma1 ? 1 : 0);
    const bool with_bias = args.dbias != nullptr && q0 == 0;
    const int total_kb = (args.M + 63) / 64;
    const int kb_begin = blockIdx.y * args.kb_per_split;
    const int kb_end = min(total_kb, kb_begin + args.kb_per_split);
    const int num_kb = kb_end - kb_begin;

    griddep_launch();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmP);
        tma_prefetch_desc(&args.tmQ);
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    // constant tile of bf16 ones (0x3F80): B operand of the bias MMA; every layout of it reads the same
    for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_proxy_async_smem();
    if (warp == 1) {
        tmem_alloc_2cta(tmem_slot, 512);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (num_kb > 0) {
        if (warp == 0) {
            if (elect_one()) {
                const int pc = p0 + static_cast<int>(cta_rank) * 128;
                // q columns staged by this CTA: its half of the first MMA's range, then its half of the second's
                const int qc0 = q0 + static_cast<int>(cta_rank) * (n_mma0 >> 1);
                const int qc1 = q0 + 256 + static_cast<int>(cta_rank) * 64;
                const uint32_t stage_tx = static_cast<uint32_t>(WG_P_BYTES + q_boxes * WG_BOX_BYTES);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * stage_tx);
                    uint8_t* a = smem + stage * WG_STAGE_BYTES;
                    uint8_t* b = a + WG_P_BYTES;
                    tma_load_2d_2cta(a, &args.tmP, &full_bar[stage], pc, kb * 64);
                    tma_load_2d_2cta(a + WG_BOX_BYTES, &args.tmP, &full_bar[stage], pc + 64, kb * 64);
                    tma_load_2d_2cta(b, &args.tmQ, &full_bar[stage], qc0, kb * 64);
                    if (n_mma0 == 256) tma_load_2d_2cta(b + WG_BOX_BYTES, &args.tmQ, &full_bar[stage], qc0 + 64, kb * 64);
                    if (has_mma1) tma_load_2d_2cta(b + 2 * WG_BOX_BYTES, &args.tmQ, &full_bar[stage], qc1, kb * 64);
                    if (++stage == WG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp == 1) {
            if (is_leader && elect_one()) {
                const uint32_t idesc0 = umma_idesc_bf16(256, n_mma0, 1, 1);
                constexpr uint32_t idesc1 = umma_idesc_bf16(256, 128, 1, 1);
                constexpr uint32_t idesc_ones = umma_idesc_bf16(256, 16, 1, 1);
                const uint64_t ones_desc = umma_smem_desc(smem_u32(sOnes), 8192, 1024);
                int stage = 0;
                uint32_t phase = 0;
                for (int i = 0; i < num_kb; ++i) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * WG_STAGE_BYTES);
                    const uint32_t b_addr = a_addr + WG_P_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // MN-major, 128B swizzle: 64-feature blocks 8192 B apart (LBO), 8-token groups 1024 B apart (SBO)
                        const uint32_t acc = (i | k) != 0 ? 1u : 0u;
                        const uint64_t ad = umma_smem_desc(a_addr + k * 2048, 8192, 1024);
                        umma_ss_2cta(tmem_base, ad, umma_smem_desc(b_addr + k * 2048, 8192, 1024), idesc0, acc);
                        if (has_mma1)
                            umma_ss_2cta(tmem_base + 256, ad, umma_smem_desc(b_addr + 2 * WG_BOX_BYTES + k * 2048, 8192, 1024),
                                         idesc1, acc);
                        if (with_bias) umma_ss_2cta(tmem_base + WG_ONES_COL, ad, ones_desc, idesc_ones, acc);
                    }
                    umma_commit_2cta(&empty_bar[stage], 3);
                    if (++stage == WG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(done_bar, 3);
            }
        } else {
            // epilogue: warp w owns TMEM lane quadrant w % 4 (32 p-rows) and one half of the q columns
            const int quad = warp & 3;
            const int half = (warp - 2) >> 2;
            const int p = p0 + static_cast<int>(cta_rank) * 128 + quad * 32 + lane;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const bool p_ok = p < args.Pdim;
            float* dst = args.R + static_cast<long long>(p) * args.ld_p + static_cast<long long>(q0) * args.ld_q;
            const bool vec = args.ld_q == 1 && (args.ld_p & 3) == 0 && (args.Qdim & 3) == 0 &&
                             (reinterpret_cast<uintptr_t>(args.R) & 15) == 0;
            const int c_begin = half * (WG_TQ / 2), c_end = min(nq, c_begin + WG_TQ / 2);
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + c0, r);
                tmem_ld_wait();
                if (!p_ok) continue;
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (c0 + j < nq)
                            red_add_v4(dst + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < nq) atomicAdd(dst + static_cast<long long>(c0 + j) * args.ld_q, __uint_as_float(r[j]));
                }
            }
            if (with_bias && half == 0) {
                const float s = __uint_as_float(tmem_ld_32x1(t_row + WG_ONES_COL));
                tmem_ld_wait();
                if (p_ok) atomicAdd(args.dbias + p, s);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename OutT, int MODE, int BN, int CG, bool ARES>
static int launch_tn_inst(const TnArgs& a, int num_sms, cudaStream_t stream) {
    using Cfg = TnCfg<BN, CG, MODE, ARES>;
    auto kfn = gemm_tn_kernel<OutT, MODE, BN, CG, ARES>;
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
    cfg.blockDim = dim3(EpiTraits<MODE, ARES>::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(attr, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
    if (e != cudaSuccess) {
        set_error("gemm_tn launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

template <int CG, bool ARES>
static int dispatch_tn(const GemmTnDesc& d, const TnArgs& a, int num_sms, cudaStream_t stream) {
    constexpr int BN = 192;
    if (d.out_f32) {
        if (d.mode == EPI_STORE) return launch_tn_inst<float, EPI_STORE, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_RESID) return launch_tn_inst<float, EPI_RESID, BN, CG, ARES>(a, num_sms, stream);
    } else {
        if (d.mode == EPI_STORE) return launch_tn_inst<__nv_bfloat16, EPI_STORE, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_GELU) return launch_tn_inst<__nv_bfloat16, EPI_GELU, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_RESID) return launch_tn_inst<__nv_bfloat16, EPI_RESID, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_DGELU) return launch_tn_inst<__nv_bfloat16, EPI_DGELU, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_GELU_ONLY) return launch_tn_inst<__nv_bfloat16, EPI_GELU_ONLY, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_GELU_GRAD) return launch_tn_inst<__nv_bfloat16, EPI_GELU_GRAD, BN, CG, ARES>(a, num_sms, stream);
        if (d.mode == EPI_MUL) return launch_tn_inst<__nv_bfloat16, EPI_MUL, BN, CG, ARES>(a, num_sms, stream);
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
    // staging-box row of the mode's epilogue (EpiTraits::BOXB): the GELU epilogues use 64-byte boxes
    const bool heavy = d.mode == EPI_GELU || d.mode == EPI_GELU_GRAD || d.mode == EPI_GELU_ONLY;
    const int box_cols = (heavy ? 64 : 128) / osz;
    rc |= make_tmap_2d(&a.tmOut, d.out, odt, d.N, d.M, (uint64_t)d.ldo * osz, box_cols, 32);
    if (d.mode == EPI_GELU || d.mode == EPI_GELU_GRAD) rc |= make_tmap_2d(&a.tmOut2, d.out2, odt, d.N, d.M, (uint64_t)d.ldo * osz, box_cols, 32);
    if (d.mode == EPI_RESID || d.mode == EPI_DGELU || d.mode == EPI_MUL)
        rc |= make_tmap_2d(&a.tmAux, d.aux, odt, d.N, d.M, (uint64_t)d.ldo * osz, box_cols, 32);
    if (rc != 0) {
        set_error("gemm_tn: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    // A-resident variant: the whole K extent of a row block fits in 6 K blocks and there are several N tiles to reuse it
    static const int force_ares = getenv("SVIT_GEMM_ARES") ? atoi(getenv("SVIT_GEMM_ARES")) : -1;
    // (measured at M = 82176: plain store 69.7 -> 68.9 us (N = 1152); fc1 + GELU (+ derivative) 149.8 -> 143.1 us with the
    // 64-byte-box epilogue, whose staging leaves room for the resident A block; the aux epilogues (residual, * gelu') LOSE
    // 10-20 % with their rings cut to two boxes and keep the streaming schedule)
    const bool heavy_mode = d.mode == EPI_GELU || d.mode == EPI_GELU_GRAD || d.mode == EPI_GELU_ONLY;
    bool ares = cg == 2 && d.K <= ARES_KB * BK && (d.N + BN - 1) / BN >= 3 && (d.mode == EPI_STORE || heavy_mode);
    if (force_ares == 0) ares = false;
    if (force_ares == 1 && d.K <= ARES_KB * BK && cg == 2) ares = true;
    if (cg == 2) return ares ? dispatch_tn<2, true>(d, a, num_sms, stream) : dispatch_tn<2, false>(d, a, num_sms, stream);
    return dispatch_tn<1, false>(d, a, num_sms, stream);  // (the A-resident variant exists for CTA pairs only)
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
    // orientation: which operand supplies the 256-row side of the tile.  MMA work of a tile = 256 x (128 | 256 | 384).
    auto mma_area = [](long long pdim, long long qdim) {
        const long long tp = (pdim + WG_TP - 1) / WG_TP;
        const long long full_q = qdim / WG_TQ, rem = qdim % WG_TQ;
        const long long cols = full_q * WG_TQ + (rem == 0 ? 0 : rem <= 128 ? 128 : rem <= 256 ? 256 : 384);
        return tp * WG_TP * cols;
    };
    const bool transposed = d.dbias == nullptr && mma_area(d.K, d.N) < mma_area(d.N, d.K);
    WgArgs a;
    memset(&a, 0, sizeof(a));
    int rc = 0;
    if (!transposed) {
        rc |= make_tmap_2d(&a.tmP, d.dY, TmapDtype::BF16, d.N, d.M, (uint64_t)d.ldy * 2, 64, 64);
        rc |= make_tmap_2d(&a.tmQ, d.X, TmapDtype::BF16, d.K, d.M, (uint64_t)d.ldx * 2, 64, 64);
        a.Pdim = d.N;
        a.Qdim = d.K;
        a.ld_p = d.ldw;
        a.ld_q = 1;
    } else {
        rc |= make_tmap_2d(&a.tmP, d.X, TmapDtype::BF16, d.K, d.M, (uint64_t)d.ldx * 2, 64, 64);
        rc |= make_tmap_2d(&a.tmQ, d.dY, TmapDtype::BF16, d.N, d.M, (uint64_t)d.ldy * 2, 64, 64);
        a.Pdim = d.K;
        a.Qdim = d.N;
        a.ld_p = 1;
        a.ld_q = d.ldw;
    }
    if (rc != 0) {
        set_error("gemm_wgrad: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.R = d.dW;
    a.dbias = d.dbias;
    a.M = d.M;
    a.tiles_q = (a.Qdim + WG_TQ - 1) / WG_TQ;
    const int tiles = ((a.Pdim + WG_TP - 1) / WG_TP) * a.tiles_q;
    const int total_kb = (d.M + 63) / 64;
    int splits = (num_sms / 2) / tiles;  // one CTA pair per SM pair and wave
    if (splits > total_kb) splits = total_kb;
    if (splits < 1) splits = 1;
    a.kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + a.kb_per_split - 1) / a.kb_per_split;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles * 2, splits);
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = WG_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(attr, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_wgrad_kernel, a);
    if (e != cudaSuccess) {
        set_error("gemm_wgrad launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

}  // namespace svit
