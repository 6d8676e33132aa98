// gemm_ln.cu -- residual Linear + the FOLLOWING LayerNorm in one kernel (sm_100a, tcgen05 / TMEM / TMA).
//
//   x_out[M, D] (fp32) = A[M, K] (bf16) * W[D, K]^T (bf16) + bias + x_in[M, D] (fp32)      <- to_out / FeedForward.net[3]
//   a[M, D]     (bf16) = LayerNorm(x_out) * gamma + beta ;  mean[M], rstd[M]               <- the next PreNorm
//
// What it replaces in the reference: `x = attn(x) + x` / `x = ff(x) + x` of vit_pytorch's Transformer.forward followed
// by the `norm` of the next PreNorm block (parameter layout pinned by /root/reference/utils/utils.py:18-23,28-31;
// constructed at /root/reference/models/sit.py:57).  A stand-alone LayerNorm kernel costs a full read of the fp32
// residual stream for zero flops; here the row statistics come out of the accumulator tile while it is still on chip.
//
// That needs the WHOLE row in one CTA, so the tile is 256 x D per CTA pair (cta_group::2; D = 384 accumulator columns
// per CTA -- SiT-small; other widths keep the separate kernels): an N = 256 and an N = 128 MMA per 16-wide K step.  Besides making the fusion possible the
// wide tile reads A once instead of once per 192-column tile -- these GEMMs are bound by the L2 -> SM operand feed.
// One accumulator stage (2 x 384 columns would not fit TMEM): the MMAs of the next tile wait for the epilogue, the
// operand ring keeps prefetching meanwhile.  (The same tile as a plain bf16-store GEMM for the 384-wide input-gradient
// GEMMs was measured and is no faster than gemm_tn's double-buffered 256 x 192 tiles: d fc1 90.4 vs 88.1 us, d qkv 71.6 vs
// 68.5 us -- what the wide tile saves in operand traffic it loses to the exposed epilogue and to 4.34 -> 5 waves.  Starting every other cluster half a tile period late, so that the HBM-bound
// epilogues of one half of the chip run under the MMA phases of the other, was measured and does not help: 85.7 -> 87.9 us
// for the out-projection shape, 139.7 -> 147.4 us for fc2 at M = 82176.)
//
// Epilogue, 8 warps = TMEM lane quadrant q (32 rows, one per thread) x column half p:
//   pass 1  acc + bias + residual -> x (fp32): staged in place over the TMA-prefetched residual box, TMA-stored, and
//           written back over the accumulator in TMEM; row sums
//   pass 2  (x - mean)^2 from TMEM (exact two-pass variance, like the stand-alone kernel)
//   pass 3  (x - mean) * rstd * gamma + beta -> bf16 -> staging -> TMA store
// The two warps of a quadrant exchange their half-row sums through shared memory.
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>

#include "gemm.cuh"
#include "ptx.cuh"
#include "tma.h"

namespace svit {

namespace {

constexpr int LBM = 128;            // rows per CTA
constexpr int LBK = 64;
constexpr int L_A_BYTES = LBM * LBK * 2;  // 16 KB
constexpr int L_BOX = 4096;         // staging box: 32 rows x 128 B
constexpr int L_NBUF = 3;           // boxes per epilogue warp (in-place residual / output rotation)
constexpr int L_EW = 8;             // epilogue warps
constexpr int L_THREADS = 128 + 32 * L_EW;
constexpr int L_STAGES = 3;
constexpr int L_SMEM_LIMIT = 232448;

template <int BN>
struct LnCfg {
    static constexpr int B_BYTES = (BN / 2) * LBK * 2;            // every CTA of the pair stages half of the weight rows
    static constexpr int STAGE_BYTES = L_A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = L_EW * L_NBUF * L_BOX;        // 96 KB
    static constexpr int VEC_BYTES = 3 * BN * 4;                   // bias, gamma, beta
    static constexpr int RED_BYTES = 2 * 128 * 4;                  // half-row partial sums
    static constexpr int BAR_BYTES = 1024;
    static constexpr int SMEM_BYTES = 1024 + L_STAGES * STAGE_BYTES + EPI_BYTES + VEC_BYTES + RED_BYTES + BAR_BYTES;
    static_assert(SMEM_BYTES <= L_SMEM_LIMIT, "gemm_ln shared memory");
    static constexpr int TMEM_COLS = BN <= 256 ? 256 : 512;
    static constexpr int N0 = BN < 256 ? BN : 256;                 // first MMA
    static constexpr int N1 = BN - N0;                             // second MMA (0 or 128)
    static constexpr int HALF = BN / 2;                            // columns per epilogue warp
    static constexpr int U1 = HALF / 32;                           // pass-1 units (32 fp32 columns = one 128-byte box row)
    static constexpr int U3 = HALF / 64;                           // pass-3 units (64 bf16 columns)
    static constexpr int JOBS = U1 + U3;                           // staging-ring jobs per warp and tile
    static_assert(HALF % 64 == 0 && JOBS % L_NBUF == 0, "the slot pattern must repeat every tile");
};

struct LnArgs {
    CUtensorMap tmA, tmB, tmX, tmAux, tmAn;
    const float* bias;
    const float* gamma;
    const float* beta;
    float* mean;
    float* rstd;
    int M, K;
    float eps;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int BN>
__global__ void __launch_bounds__(L_THREADS, 1) gemm_ln_kernel(const __grid_constant__ LnArgs args) {
    using Cfg = LnCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                         // [STAGES][16 KB]
    uint8_t* sB = sA + L_STAGES * L_A_BYTES;                    // [STAGES][B_BYTES]
    uint8_t* sEpi = sB + L_STAGES * Cfg::B_BYTES;               // [EW][NBUF][4 KB]
    float* sBias = reinterpret_cast<float*>(sEpi + Cfg::EPI_BYTES);
    float* sGamma = sBias + BN;
    float* sBeta = sGamma + BN;
    float* sRed = sBeta + BN;                                   // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 256);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + L_STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * L_STAGES;    // accumulator complete       (MMA -> epilogue)
    uint64_t* tempty_bar = tfull_bar + 1;         // accumulator read out        (epilogue of both CTAs -> MMA)
    uint64_t* afull_bar = tempty_bar + 1;         // [EW][NBUF] residual box landed
    uint64_t* aempty_bar = afull_bar + L_EW * L_NBUF;  // [EW][NBUF] box free for the next residual
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + L_EW * L_NBUF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int M = args.M, K = args.K;
    const int num_tiles = (M + 2 * LBM - 1) / (2 * LBM);
    const int num_kb = (K + LBK - 1) / LBK;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int num_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
    const int row_off = static_cast<int>(cta_rank) * LBM;

    griddep_launch();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmA);
        tma_prefetch_desc(&args.tmB);
        tma_prefetch_desc(&args.tmX);
        tma_prefetch_desc(&args.tmAux);
        tma_prefetch_desc(&args.tmAn);
        for (int i = 0; i < L_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, L_EW * 2);  // the leader's barrier also collects the peer's epilogue warps
        for (int i = 0; i < L_EW * L_NBUF; ++i) {
            mbar_init(&afull_bar[i], 1);
            mbar_init(&aempty_bar[i], 1);
        }
        fence_mbar_init();
    }
    griddep_wait();  // before the first global read (the parameters are written by the optimiser kernels)
    for (int i = threadIdx.x; i < BN; i += L_THREADS) {
        sBias[i] = args.bias != nullptr ? args.bias[i] : 0.0f;
        sGamma[i] = args.gamma[i];
        sBeta[i] = args.beta[i];
    }
    if (warp == 3) {
        tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = tile * 2 * LBM + row_off;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    // both CTAs' bytes are credited to the leader's barrier, which the leader arms for the pair
                    if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    tma_load_2d_2cta(sA + stage * L_A_BYTES, &args.tmA, &full_bar[stage], kb * LBK, m0);
                    // weight rows of this CTA: its half of the first MMA's N range, then its half of the second's
                    uint8_t* b = sB + stage * Cfg::B_BYTES;
                    const int n_first = static_cast<int>(cta_rank) * (Cfg::N0 / 2);
#pragma unroll
                    for (int r = 0; r < Cfg::N0 / 2; r += 32)
                        tma_load_2d_2cta(b + r * 128, &args.tmB, &full_bar[stage], kb * LBK, n_first + r);
                    if constexpr (Cfg::N1 > 0) {
                        const int n_second = Cfg::N0 + static_cast<int>(cta_rank) * (Cfg::N1 / 2);
#pragma unroll
                        for (int r = 0; r < Cfg::N1 / 2; r += 32)
                            tma_load_2d_2cta(b + (Cfg::N0 / 2 + r) * 128, &args.tmB, &full_bar[stage], kb * LBK, n_second + r);
                    }
                    if (++stage == L_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc0 = umma_idesc_bf16(256, Cfg::N0, 0, 0);
            constexpr uint32_t idesc1 = umma_idesc_bf16(256, Cfg::N1 > 0 ? Cfg::N1 : 16, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                mbar_wait(tempty_bar, (it & 1) ^ 1);  // the epilogues of the previous tile have read the accumulator out
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * L_A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < LBK / 16; ++k) {
                        const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        umma_ss_2cta(tmem_base, ad, umma_smem_desc(b_addr + k * 32, 16, 1024), idesc0, acc);
                        if constexpr (Cfg::N1 > 0)
                            umma_ss_2cta(tmem_base + Cfg::N0, ad, umma_smem_desc(b_addr + (Cfg::N0 / 2) * 128 + k * 32, 16, 1024),
                                         idesc1, acc);
                    }
                    umma_commit_2cta(&empty_bar[stage], 3);
                    if (++stage == L_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(tfull_bar, 3);
            }
        }
    } else if (warp == 2) {
        // ===================== residual loader =====================
        // Per epilogue warp and tile: U1 pass-1 jobs (residual box in, x out, in place) and U3 pass-3 jobs (bf16 out only);
        // job n of a warp uses box n % NBUF, and JOBS % NBUF == 0, so the pattern is the same for every tile.
        if (elect_one()) {
            uint32_t loads = 0;  // residual loads issued per warp so far (the same for all warps)
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = tile * 2 * LBM + row_off;
                for (int u = 0; u < Cfg::U1; ++u, ++loads) {
                    const int slot = u % L_NBUF;
                    // this is load number loads / ... into `slot`: every slot takes U1 / NBUF loads per tile
                    const uint32_t nth = (loads / Cfg::U1) * (Cfg::U1 / L_NBUF) + u / L_NBUF;  // loads into this slot before this one
#pragma unroll 1
                    for (int w = 0; w < L_EW; ++w) {
                        const int q = w & 3, p = w >> 2;
                        mbar_wait(&aempty_bar[w * L_NBUF + slot], (nth & 1) ^ 1);
                        mbar_expect_tx(&afull_bar[w * L_NBUF + slot], L_BOX);
                        tma_load_2d(sEpi + (w * L_NBUF + slot) * L_BOX, &args.tmAux, &afull_bar[w * L_NBUF + slot],
                                    p * Cfg::HALF + u * 32, m0 + q * 32);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const int q = ew & 3;   // TMEM lane quadrant == 32-row slice of the tile
        const int p = ew >> 2;  // column half
        uint8_t* wbuf = sEpi + ew * L_NBUF * L_BOX;
        const int sw = lane & 7;
        const int row = q * 32 + lane;
        const int cbase = p * Cfg::HALF;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cbase;
        const float inv_d = 1.0f / static_cast<float>(BN);
        // the two warps of a quadrant hold the two halves of the same rows (constant barrier ids)
        auto pair_sum = [&](float mine) {
            sRed[p * 128 + row] = mine;
            if (q == 0) named_bar_sync(1, 64);
            else if (q == 1) named_bar_sync(2, 64);
            else if (q == 2) named_bar_sync(3, 64);
            else named_bar_sync(4, 64);
            const float tot = sRed[row] + sRed[128 + row];
            if (q == 0) named_bar_sync(1, 64);
            else if (q == 1) named_bar_sync(2, 64);
            else if (q == 2) named_bar_sync(3, 64);
            else named_bar_sync(4, 64);
            return tot;
        };
        uint32_t nload = 0;   // residual loads consumed from each slot pattern (pass-1 jobs done), for the afull parity
        bool have_prev = false;
        int prev_slot = 0;
        bool prev_release = false;
        // after committing a job's store: the PREVIOUS job's store has been read out of its box; hand that box back to the
        // residual loader if its next occupant is a pass-1 job
        auto job_done = [&](int slot, bool release_for_aux) {
            if (lane == 0) {
                tma_store_commit();
                if (have_prev) {
                    tma_store_wait_read<1>();
                    if (prev_release) mbar_arrive(&aempty_bar[ew * L_NBUF + prev_slot]);
                }
            }
            have_prev = true;
            prev_slot = slot;
            prev_release = release_for_aux;
        };
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m0 = tile * 2 * LBM + row_off + q * 32;  // first row of this warp's slice
            const int grow = m0 + lane;
            mbar_wait(tfull_bar, it & 1);
            tc_fence_after();
            // ---------------- pass 1: x = acc + bias + residual ----------------
            float s = 0.0f;
#pragma unroll 1
            for (int u = 0; u < Cfg::U1; ++u, ++nload) {
                const int slot = u % L_NBUF;
                const uint32_t nth = (nload / Cfg::U1) * (Cfg::U1 / L_NBUF) + u / L_NBUF;
                uint8_t* orow = wbuf + slot * L_BOX + lane * 128;
                uint32_t r[32];
                tmem_ld_32x32(t_row + u * 32, r);
                mbar_wait(&afull_bar[ew * L_NBUF + slot], nth & 1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4* cell = reinterpret_cast<uint4*>(orow + ((c ^ sw) << 4));
                    const uint4 a4 = *cell;
                    const float* bs = sBias + cbase + u * 32 + c * 4;
                    float v0 = __uint_as_float(r[c * 4 + 0]) + bs[0] + __uint_as_float(a4.x);
                    float v1 = __uint_as_float(r[c * 4 + 1]) + bs[1] + __uint_as_float(a4.y);
                    float v2 = __uint_as_float(r[c * 4 + 2]) + bs[2] + __uint_as_float(a4.z);
                    float v3 = __uint_as_float(r[c * 4 + 3]) + bs[3] + __uint_as_float(a4.w);
                    s += (v0 + v1) + (v2 + v3);
                    r[c * 4 + 0] = __float_as_uint(v0);
                    r[c * 4 + 1] = __float_as_uint(v1);
                    r[c * 4 + 2] = __float_as_uint(v2);
                    r[c * 4 + 3] = __float_as_uint(v3);
                    *cell = make_uint4(r[c * 4 + 0], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]);
                }
                tmem_st_32x32(t_row + u * 32, r);  // x back over the accumulator: passes 2 and 3 read it from there
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tma_store_2d(&args.tmX, wbuf + slot * L_BOX, cbase + u * 32, m0);
                // next occupant of this box: job u + NBUF -- a pass-1 job (residual load) iff u + NBUF < U1
                job_done(slot, u + L_NBUF < Cfg::U1);
            }
            tmem_st_wait();
            const float mean = pair_sum(s) * inv_d;
            // ---------------- pass 2: exact variance ----------------
            float qsum = 0.0f;
#pragma unroll 1
            for (int u = 0; u < Cfg::U1; ++u) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + u * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float d = __uint_as_float(r[j]) - mean;
                    qsum = fmaf(d, d, qsum);
                }
            }
            const float rstd = rsqrtf(pair_sum(qsum) * inv_d + args.eps);
            if (p == 0 && grow < M) {
                args.mean[grow] = mean;
                args.rstd[grow] = rstd;
            }
            // ---------------- pass 3: normalised bf16 operand of the next GEMM ----------------
#pragma unroll 1
            for (int u = 0; u < Cfg::U3; ++u) {
                const int slot = (Cfg::U1 + u) % L_NBUF;
                uint8_t* orow = wbuf + slot * L_BOX + lane * 128;
                // the store that last used this box (three jobs ago) was waited for when the previous job committed
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + u * 64 + hh * 32, r);
                    tmem_ld_wait();
                    if (u == Cfg::U3 - 1 && hh == 1) {
                        // this warp's slice of the accumulator is fully read: hand TMEM back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_bar, 0);
                    }
                    const float* gm = sGamma + cbase + u * 64 + hh * 32;
                    const float* bt = sBeta + cbase + u * 64 + hh * 32;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            y[e] = fmaf((__uint_as_float(r[c * 8 + e]) - mean) * rstd, gm[c * 8 + e], bt[c * 8 + e]);
                        uint4 o;
                        o.x = pack_bf16(y[0], y[1]);
                        o.y = pack_bf16(y[2], y[3]);
                        o.z = pack_bf16(y[4], y[5]);
                        o.w = pack_bf16(y[6], y[7]);
                        *reinterpret_cast<uint4*>(orow + (((hh * 4 + c) ^ sw) << 4)) = o;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tma_store_2d(&args.tmAn, wbuf + slot * L_BOX, cbase + u * 64, m0);
                // next occupant of this box: job U1 + u + NBUF of this tile (pass 3) or a pass-1 job of the next tile
                job_done(slot, Cfg::U1 + u + L_NBUF >= Cfg::JOBS);
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN>
int launch_ln_inst(const LnArgs& a, int num_sms, cudaStream_t stream) {
    using Cfg = LnCfg<BN>;
    auto kfn = gemm_ln_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_ln) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    const int tiles = (a.M + 2 * LBM - 1) / (2 * LBM);
    const int max_groups = num_sms / 2;
    const int groups = tiles < max_groups ? tiles : max_groups;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(groups * 2);
    cfg.blockDim = dim3(L_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(attr, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
    if (e != cudaSuccess) {
        set_error("gemm_ln launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    count_launch();
    return 0;
}

}  // namespace

bool gemm_ln_supported(int D) { return D == 384; }

int launch_gemm_ln(const GemmLnDesc& d, int num_sms, cudaStream_t stream) {
    if (!gemm_ln_supported(d.D)) {
        set_error("gemm_ln: the fused Linear + LayerNorm tile is built for dim 384 (got %d)", d.D);
        return -4;
    }
    if (d.M <= 0 || d.K <= 0 || (d.lda % 8) || (d.ldb % 8)) {
        set_error("gemm_ln: bad problem M=%d K=%d lda=%d ldb=%d", d.M, d.K, d.lda, d.ldb);
        return -1;
    }
    LnArgs a;
    memset(&a, 0, sizeof(a));
    int rc = 0;
    rc |= make_tmap_2d(&a.tmA, d.A, TmapDtype::BF16, d.K, d.M, (uint64_t)d.lda * 2, LBK, LBM);
    rc |= make_tmap_2d(&a.tmB, d.W, TmapDtype::BF16, d.K, d.D, (uint64_t)d.ldb * 2, LBK, 32);
    rc |= make_tmap_2d(&a.tmX, d.x_out, TmapDtype::F32, d.D, d.M, (uint64_t)d.D * 4, 32, 32);
    rc |= make_tmap_2d(&a.tmAux, d.x_in, TmapDtype::F32, d.D, d.M, (uint64_t)d.D * 4, 32, 32);
    rc |= make_tmap_2d(&a.tmAn, d.a_out, TmapDtype::BF16, d.D, d.M, (uint64_t)d.D * 2, 64, 32);
    if (rc != 0) {
        set_error("gemm_ln: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.bias = d.bias;
    a.gamma = d.gamma;
    a.beta = d.beta;
    a.mean = d.mean;
    a.rstd = d.rstd;
    a.M = d.M;
    a.K = d.K;
    a.eps = d.eps;
    return launch_ln_inst<384>(a, num_sms, stream);
}

}  // namespace svit
