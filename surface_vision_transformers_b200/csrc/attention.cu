// attention.cu -- tcgen05 fused attention forward / backward for sequences of at most 384 tokens.
// See attention.cuh for what it replaces in the reference.
#include "attention.cuh"

#include <cuda_bf16.h>
#include <cstring>
#include <cstdlib>

#include "gemm.cuh"  // set_error
#include "ptx.cuh"
#include "tma.h"

namespace svit {

constexpr int ATT_MAX_T = 384;
constexpr int TILE_BYTES = 128 * 128;  // one [128 rows x 64 bf16] swizzled tile

// =================================================================================================
// forward
// =================================================================================================
// One CTA per (sample, head); K and V (<= 384 keys) stay in shared memory, Q blocks of 128 rows stream through.
// TMEM reads cost 64 B/clk/SM, as much as the exponentials themselves, so the scores are read ONCE: the softmax
// shift is the row's score against key 0 (any shift is exact for softmax; the log-sum-exp is reported with the same
// shift), and a guard on the row sum falls back to the classic max-shift pass in the (never observed) overflow case.
// 8 softmax warps: warp (q, hf) owns TMEM lane quadrant q and one half of the key columns.
// shared memory map (bytes):  sQ 16K | sK 48K | sV 48K | sP 96K | sO 16K | row sums / maxima | barriers
constexpr int FWD_SQ = 0;
constexpr int FWD_SK = FWD_SQ + TILE_BYTES;
constexpr int FWD_SV = FWD_SK + 3 * TILE_BYTES;
constexpr int FWD_SP = FWD_SV + 3 * TILE_BYTES;
constexpr int FWD_SO = FWD_SP + 6 * TILE_BYTES;
constexpr int FWD_RED = FWD_SO + TILE_BYTES;      // float [2][128]
constexpr int FWD_BAR = FWD_RED + 2 * 128 * 4;
constexpr int FWD_SMEM = 1024 + FWD_BAR + 128;
constexpr int FWD_THREADS = 288;
constexpr int FWD_TMEM_O = 384;  // O accumulator columns [384, 448)

struct AttnFwdArgs {
    CUtensorMap tmQKV;  // (3*inner, T, B) bf16, box 64 x 128 x 1
    CUtensorMap tmO;    // (inner, T, B) bf16, box 64 x 128 x 1
    float* lse;
    int B, H, T;
    float scale, scale_log2e;
};

__device__ __forceinline__ float fwd_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(FWD_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnFwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + FWD_SQ;
    uint8_t* sK = smem + FWD_SK;
    uint8_t* sV = smem + FWD_SV;
    uint8_t* sP = smem + FWD_SP;
    uint8_t* sO = smem + FWD_SO;
    float* sRed = reinterpret_cast<float*>(smem + FWD_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_BAR);
    uint64_t* bar_kv = bars + 0;
    uint64_t* bar_q = bars + 1;
    uint64_t* bar_s = bars + 2;
    uint64_t* bar_p = bars + 3;
    uint64_t* bar_o = bars + 4;
    uint64_t* bar_of = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = args.T, H = args.H;
    const int inner = H * 64;
    const int b = blockIdx.x / H;
    const int h = blockIdx.x % H;
    const int tk = (T + 15) & ~15;       // keys padded to the UMMA K step
    const int nkb = (T + 127) / 128;     // 128-row K/V boxes
    const int nqb = (T + 127) / 128;     // query blocks
    const int n1 = tk < 256 ? tk : 256;  // first S chunk (UMMA N <= 256)
    const int n2 = tk - n1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmQKV);
        tma_prefetch_desc(&args.tmO);
        mbar_init(bar_kv, 1);
        mbar_init(bar_q, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 256);
        mbar_init(bar_o, 1);
        mbar_init(bar_of, 256);
        fence_mbar_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ============================ control: TMA + MMA issue ============================
        if (elect_one()) {
            mbar_expect_tx(bar_kv, nkb * 2 * TILE_BYTES);
            for (int r = 0; r < nkb; ++r) {
                tma_load_3d(sK + r * TILE_BYTES, &args.tmQKV, bar_kv, inner + h * 64, r * 128, b);
                tma_load_3d(sV + r * TILE_BYTES, &args.tmQKV, bar_kv, 2 * inner + h * 64, r * 128, b);
            }
            mbar_expect_tx(bar_q, TILE_BYTES);
            tma_load_3d(sQ, &args.tmQKV, bar_q, h * 64, 0, b);
            const uint32_t idesc_s1 = umma_idesc_bf16(128, n1, 0, 0);
            const uint32_t idesc_s2 = umma_idesc_bf16(128, n2 > 0 ? n2 : 16, 0, 0);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
            mbar_wait(bar_kv, 0);
            for (int i = 0; i < nqb; ++i) {
                const uint32_t ph = i & 1;
                mbar_wait(bar_q, ph);
                tc_fence_after();
                // S = Q K^T  (K-major A and B)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = umma_smem_desc(q_addr + k * 32, 16, 1024);
                    umma_ss(tmem_base, ad, umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s1, k != 0);
                    if (n2 > 0)
                        umma_ss(tmem_base + n1, ad, umma_smem_desc(k_addr + n1 * 128 + k * 32, 16, 1024), idesc_s2, k != 0);
                }
                umma_commit(bar_s);
                mbar_wait(bar_s, ph);  // S done -> sQ reusable
                if (i + 1 < nqb) {
                    mbar_expect_tx(bar_q, TILE_BYTES);
                    tma_load_3d(sQ, &args.tmQKV, bar_q, h * 64, (i + 1) * 128, b);
                }
                mbar_wait(bar_p, ph);  // P written to smem, S columns free
                if (i > 0) mbar_wait(bar_of, (i - 1) & 1);  // previous O drained from TMEM
                tc_fence_after();
                // O = P V  (A = P K-major from smem, B = V MN-major)
                const int ksteps = tk / 16;
                for (int s = 0; s < ksteps; ++s) {
                    const uint64_t ad = umma_smem_desc(p_addr + (s >> 2) * TILE_BYTES + (s & 3) * 32, 16, 1024);
                    const uint64_t bd = umma_smem_desc(v_addr + s * 2048, 8192, 1024);
                    umma_ss(tmem_base + FWD_TMEM_O, ad, bd, idesc_pv, s != 0);
                }
                umma_commit(bar_o);
            }
        }
    } else {
        // ============================ softmax + epilogue warps ============================
        const int q = warp & 3;    // TMEM lane quadrant
        const int hf = warp >> 2;  // key-column half
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const bool leader = threadIdx.x == 0;
        const float c = args.scale_log2e;
        const int split = (tk / 2) & ~31;                 // half 0: columns [0, split), half 1: [split, tk)
        const int col_lo = hf == 0 ? 0 : split;
        const int col_hi = hf == 0 ? split : tk;

        // exp2((s - shift) * c) of this thread's columns -> bf16 P tile in smem; returns the partial row sum
        auto softmax_pass = [&](float shift_c) {
            float sum = 0.0f;
            for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(t_row + c0, r);
                tmem_ld_wait();
                uint8_t* prow = sP + (c0 >> 6) * TILE_BYTES + row * 128;
                const int cb = (c0 & 63) >> 3;  // first 16-byte chunk of this 32-column group (0 or 4)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = c0 + g * 8 + e;
                        const float ex = fwd_ex2(fmaf(__uint_as_float(r[g * 8 + e]), c, -shift_c));
                        pv[e] = (col < T && col < col_hi) ? ex : 0.0f;
                    }
                    uint4 o;
                    o.x = pack_bf16(pv[0], pv[1]);
                    o.y = pack_bf16(pv[2], pv[3]);
                    o.z = pack_bf16(pv[4], pv[5]);
                    o.w = pack_bf16(pv[6], pv[7]);
                    // the row sum uses the bf16-rounded probabilities that the PV product will see
                    sum += bf16_lo(o.x) + bf16_hi(o.x) + bf16_lo(o.y) + bf16_hi(o.y) + bf16_lo(o.z) + bf16_hi(o.z) +
                           bf16_lo(o.w) + bf16_hi(o.w);
                    // a 32-column group that straddles col_hi belongs to this half only up to col_hi; the other
                    // half never writes these chunks (its range starts at a multiple of 32)
                    *reinterpret_cast<uint4*>(prow + (((cb + g) ^ (row & 7)) << 4)) = o;
                }
            }
            return sum;
        };

        for (int i = 0; i < nqb; ++i) {
            const uint32_t ph = i & 1;
            mbar_wait(bar_s, ph);
            tc_fence_after();
            // ---- single pass with the score against key 0 as the softmax shift ----
            float shift = __uint_as_float(tmem_ld_32x1(t_row));
            tmem_ld_wait();
            float part = softmax_pass(shift * c);
            sRed[hf * 128 + row] = part;
            bool bad = named_bar_or(1, 256, false);  // (barrier only: make both halves' partial sums visible)
            float total = sRed[row] + sRed[128 + row];
            bad = !(total > 0.0f && total < 1e30f);
            if (named_bar_or(2, 256, bad)) {
                // ---- fallback (uniform for the CTA): classic max-shifted softmax ----
                float mx = -INFINITY;
                for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < T && c0 + j < col_hi) mx = fmaxf(mx, __uint_as_float(r[j]));
                }
                sRed[hf * 128 + row] = mx;
                named_bar_sync(1, 256);
                shift = fmaxf(sRed[row], sRed[128 + row]);
                named_bar_sync(2, 256);
                part = softmax_pass(shift * c);
                sRed[hf * 128 + row] = part;
                named_bar_sync(1, 256);
                total = sRed[row] + sRed[128 + row];
                named_bar_sync(2, 256);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p);
            // ---- epilogue: O / sum -> bf16 -> staging -> TMA store (each half converts 32 of the 64 columns) ----
            mbar_wait(bar_o, ph);
            tc_fence_after();
            const float inv = 1.0f / total;
            uint32_t o0[32];
            tmem_ld_32x32(t_row + FWD_TMEM_O + hf * 32, o0);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_of);
            if (i > 0) {
                if (leader) tma_store_wait_read<0>();  // previous O tile left the staging buffer
                named_bar_sync(1, 256);
            }
            uint8_t* orow = sO + row * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t* src = &o0[g * 8];
                uint4 o;
                o.x = pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
                o.y = pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
                o.z = pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
                o.w = pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
                *reinterpret_cast<uint4*>(orow + (((hf * 4 + g) ^ (row & 7)) << 4)) = o;
            }
            const int t = i * 128 + row;
            if (hf == 0 && t < T) args.lse[(static_cast<size_t>(b) * H + h) * T + t] = shift * args.scale + logf(total);
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (leader) {
                tma_store_3d(&args.tmO, sO, h * 64, i * 128, b);
                tma_store_commit();
            }
        }
        if (leader) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =================================================================================================
// backward
// =================================================================================================
// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   -- one warp per (b,t) row
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int B, int H, int T) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= B * T) return;
    const int b = warp_global / T, t = warp_global % T;
    const size_t base = static_cast<size_t>(warp_global) * H * 64;
    for (int h = 0; h < H; ++h) {
        const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(out + base + h * 64 + lane * 2);
        const __nv_bfloat162 d2 = *reinterpret_cast<const __nv_bfloat162*>(dout + base + h * 64 + lane * 2);
        float s = __bfloat162float(o2.x) * __bfloat162float(d2.x) + __bfloat162float(o2.y) * __bfloat162float(d2.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) delta[(static_cast<size_t>(b) * H + h) * T + t] = s;
    }
}

// One kernel, two roles (template DQ):
//   DQ = false : CTA owns key block j -> accumulates dK_j, dV_j over all query blocks
//   DQ = true  : CTA owns query block i -> accumulates dQ_i over all key blocks
// Per step: S = Q K^T, dP = dO V^T (TMEM); threads (one query row each) form P = exp(S*scale - lse) and
// dS = P * (dP - delta) in bf16 shared memory tiles [query][key]; then the accumulate MMAs read them.
// smem: sQ | sdO | sK | sV (16K each) | sP 32K | sdS 32K | barriers
constexpr int BWD_SQ = 0;
constexpr int BWD_SDO = BWD_SQ + TILE_BYTES;
constexpr int BWD_SK = BWD_SDO + TILE_BYTES;
constexpr int BWD_SV = BWD_SK + TILE_BYTES;
constexpr int BWD_SP = BWD_SV + TILE_BYTES;
constexpr int BWD_SDS = BWD_SP + 2 * TILE_BYTES;
constexpr int BWD_BAR = BWD_SDS + 2 * TILE_BYTES;
constexpr int BWD_SMEM = 1024 + BWD_BAR + 128;
constexpr int BWD_THREADS = 128;
// TMEM columns: S [0,128) dP [128,256) acc0 [256,320) acc1 [320,384)

struct AttnBwdArgs {
    CUtensorMap tmQKV;   // (3*inner, T, B) bf16 box 64x128x1 (loads)
    CUtensorMap tmDO;    // (inner, T, B) bf16 box 64x128x1
    CUtensorMap tmDQKV;  // (3*inner, T, B) bf16 box 64x128x1 (stores)
    const float* lse;
    const float* delta;
    int B, H, T;
    float scale, scale_log2e;
};

template <bool DQ>
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + BWD_SQ;
    uint8_t* sdO = smem + BWD_SDO;
    uint8_t* sK = smem + BWD_SK;
    uint8_t* sV = smem + BWD_SV;
    uint8_t* sP = smem + BWD_SP;
    uint8_t* sdS = smem + BWD_SDS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BWD_BAR);
    uint64_t* bar_fix = bars + 0;  // resident operands landed
    uint64_t* bar_ld = bars + 1;   // streamed operands landed
    uint64_t* bar_s = bars + 2;    // S and dP ready
    uint64_t* bar_acc = bars + 3;  // accumulate MMAs retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = args.T, H = args.H;
    const int inner = H * 64;
    const int nblk = (T + 127) / 128;
    const int fixed = blockIdx.x % nblk;  // key block (DQ=false) or query block (DQ=true)
    const int bh = blockIdx.x / nblk;
    const int b = bh / H, h = bh % H;
    const bool t0 = threadIdx.x == 0;

    if (t0) {
        tma_prefetch_desc(&args.tmQKV);
        tma_prefetch_desc(&args.tmDO);
        tma_prefetch_desc(&args.tmDQKV);
        mbar_init(bar_fix, 1);
        mbar_init(bar_ld, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int row = warp * 32 + lane;

    const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);     // S / dP : K-major x K-major
    const uint32_t idesc_tn = umma_idesc_bf16(128, 64, 1, 1);     // dK / dV: A = P^T/dS^T (MN-major), B MN-major
    const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);     // dQ    : A = dS (K-major),      B = K MN-major

    if (t0) {
        mbar_expect_tx(bar_fix, 2 * TILE_BYTES);
        if (DQ) {
            tma_load_3d(sQ, &args.tmQKV, bar_fix, h * 64, fixed * 128, b);
            tma_load_3d(sdO, &args.tmDO, bar_fix, h * 64, fixed * 128, b);
        } else {
            tma_load_3d(sK, &args.tmQKV, bar_fix, inner + h * 64, fixed * 128, b);
            tma_load_3d(sV, &args.tmQKV, bar_fix, 2 * inner + h * 64, fixed * 128, b);
        }
    }
    mbar_wait(bar_fix, 0);

    float lse_r = 0.0f, delta_r = 0.0f;
    bool qvalid = false;
    if (DQ) {
        const int t = fixed * 128 + row;
        qvalid = t < T;
        if (qvalid) {
            lse_r = args.lse[(static_cast<size_t>(b) * H + h) * T + t];
            delta_r = args.delta[(static_cast<size_t>(b) * H + h) * T + t];
        }
    }

    for (int step = 0; step < nblk; ++step) {
        const uint32_t ph = step & 1;
        const int qb = DQ ? fixed : step;
        const int kb = DQ ? step : fixed;
        if (t0) {
            mbar_expect_tx(bar_ld, 2 * TILE_BYTES);
            if (DQ) {
                tma_load_3d(sK, &args.tmQKV, bar_ld, inner + h * 64, kb * 128, b);
                tma_load_3d(sV, &args.tmQKV, bar_ld, 2 * inner + h * 64, kb * 128, b);
            } else {
                tma_load_3d(sQ, &args.tmQKV, bar_ld, h * 64, qb * 128, b);
                tma_load_3d(sdO, &args.tmDO, bar_ld, h * 64, qb * 128, b);
            }
        }
        if (!DQ) {
            const int t = qb * 128 + row;
            qvalid = t < T;
            lse_r = qvalid ? args.lse[(static_cast<size_t>(b) * H + h) * T + t] : 0.0f;
            delta_r = qvalid ? args.delta[(static_cast<size_t>(b) * H + h) * T + t] : 0.0f;
        }
        mbar_wait(bar_ld, ph);
        if (t0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_ss(tmem_base, umma_smem_desc(q_addr + k * 32, 16, 1024), umma_smem_desc(k_addr + k * 32, 16, 1024),
                        idesc_s, k != 0);
                umma_ss(tmem_base + 128, umma_smem_desc(do_addr + k * 32, 16, 1024),
                        umma_smem_desc(v_addr + k * 32, 16, 1024), idesc_s, k != 0);
            }
            umma_commit(bar_s);
        }
        mbar_wait(bar_s, ph);
        tc_fence_after();
        // ---- P and dS for this (query block, key block) pair ----
        const float lse2 = lse_r * 1.4426950408889634f;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t s[32], dp[32];
            tmem_ld_32x32(t_row + c0, s);
            tmem_ld_32x32(t_row + 128 + c0, dp);
            tmem_ld_wait();
            float p[32], ds[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool valid = qvalid && (kb * 128 + c0 + j < T);
                const float e = exp2f(__uint_as_float(s[j]) * args.scale_log2e - lse2);
                p[j] = valid ? e : 0.0f;
                ds[j] = p[j] * (__uint_as_float(dp[j]) - delta_r);
            }
            const int tile = c0 >> 6;
            const int cb = (c0 & 63) >> 3;
            uint8_t* prow = sP + tile * TILE_BYTES + row * 128;
            uint8_t* dsrow = sdS + tile * TILE_BYTES + row * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                const int sw = ((cb + g) ^ (row & 7)) << 4;
                if (!DQ) {
                    o.x = pack_bf16(p[g * 8 + 0], p[g * 8 + 1]);
                    o.y = pack_bf16(p[g * 8 + 2], p[g * 8 + 3]);
                    o.z = pack_bf16(p[g * 8 + 4], p[g * 8 + 5]);
                    o.w = pack_bf16(p[g * 8 + 6], p[g * 8 + 7]);
                    *reinterpret_cast<uint4*>(prow + sw) = o;
                }
                o.x = pack_bf16(ds[g * 8 + 0], ds[g * 8 + 1]);
                o.y = pack_bf16(ds[g * 8 + 2], ds[g * 8 + 3]);
                o.z = pack_bf16(ds[g * 8 + 4], ds[g * 8 + 5]);
                o.w = pack_bf16(ds[g * 8 + 6], ds[g * 8 + 7]);
                *reinterpret_cast<uint4*>(dsrow + sw) = o;
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t0) {
            tc_fence_after();
            if (DQ) {
                // dQ += dS K : A = dS [q][key] K-major, B = K [key][d] MN-major, 8 K-steps of 16 keys
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    umma_ss(tmem_base + 256, umma_smem_desc(ds_addr + (s >> 2) * TILE_BYTES + (s & 3) * 32, 16, 1024),
                            umma_smem_desc(k_addr + s * 2048, 8192, 1024), idesc_dq, (step | s) != 0);
            } else {
                // dV += P^T dO ; dK += dS^T Q : A = [q][key] tiles read MN-major (M = key), B = dO / Q MN-major
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    umma_ss(tmem_base + 320, umma_smem_desc(p_addr + s * 2048, TILE_BYTES, 1024),
                            umma_smem_desc(do_addr + s * 2048, 8192, 1024), idesc_tn, (step | s) != 0);
                    umma_ss(tmem_base + 256, umma_smem_desc(ds_addr + s * 2048, TILE_BYTES, 1024),
                            umma_smem_desc(q_addr + s * 2048, 8192, 1024), idesc_tn, (step | s) != 0);
                }
            }
            umma_commit(bar_acc);
        }
        mbar_wait(bar_acc, ph);
        tc_fence_after();
    }

    // ---- epilogue: accumulators -> bf16 -> staging (sP) -> TMA store into dqkv ----
    const int nacc = DQ ? 1 : 2;
    for (int a = 0; a < nacc; ++a) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_row + 256 + a * 64, o0);
        tmem_ld_32x32(t_row + 256 + a * 64 + 32, o1);
        tmem_ld_wait();
        const float sc = (a == 0) ? args.scale : 1.0f;  // dQ and dK carry the softmax scale, dV does not
        uint8_t* orow = sP + a * TILE_BYTES + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t* src = g < 4 ? &o0[g * 8] : &o1[(g - 4) * 8];
            uint4 o;
            o.x = pack_bf16(__uint_as_float(src[0]) * sc, __uint_as_float(src[1]) * sc);
            o.y = pack_bf16(__uint_as_float(src[2]) * sc, __uint_as_float(src[3]) * sc);
            o.z = pack_bf16(__uint_as_float(src[4]) * sc, __uint_as_float(src[5]) * sc);
            o.w = pack_bf16(__uint_as_float(src[6]) * sc, __uint_as_float(src[7]) * sc);
            *reinterpret_cast<uint4*>(orow + ((g ^ (row & 7)) << 4)) = o;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t0) {
        if (DQ) {
            tma_store_3d(&args.tmDQKV, sP, h * 64, fixed * 128, b);
        } else {
            tma_store_3d(&args.tmDQKV, sP, inner + h * 64, fixed * 128, b);                    // dK
            tma_store_3d(&args.tmDQKV, sP + TILE_BYTES, 2 * inner + h * 64, fixed * 128, b);   // dV
        }
        tma_store_commit();
        tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =================================================================================================
// backward, fused and software-pipelined: one CTA per (sample, head, key block j).
//   tensor core (one issuing thread)            compute warps (8 warps, one query row x 64 key columns per thread)
//   a(i): S  = Q_i K_j^T                         P(i)  = exp2(S*scale*log2e - lse*log2e)      -> smem tile sP
//   b(i): dP = dO_i V_j^T                        dS(i) = P * (dP*scale - delta*scale)          -> smem tile sdS
//   c(i): dV_j += P^T dO_i                       dQ(i) : TMEM partial -> fp32 red.add into dq_accum
//   d(i): dK_j += dS^T Q_i ; dQ_part = dS K_j
// Issue order  a(0) b(0) | c(i) a(i+1) d(i) b(i+1) | ...  so the MMAs of pair i+1 / the accumulate MMAs of pair i run
// while the compute warps are still busy with pair i; Q_i / dO_i are double-buffered by a TMA producer warp.
// smem: sK | sV | sQ[2] | sdO[2] (16K each) | sP 32K | sdS 32K | barriers  (161 KB, 1 CTA / SM)
// TMEM: S [0,128) dP [128,256) dK [256,320) dV [320,384) dQ_part [384,448)
// =================================================================================================
constexpr int BF_SK = 0;
constexpr int BF_SV = BF_SK + TILE_BYTES;
constexpr int BF_SQ = BF_SV + TILE_BYTES;        // 2 stages
constexpr int BF_SDO = BF_SQ + 2 * TILE_BYTES;   // 2 stages
constexpr int BF_SP = BF_SDO + 2 * TILE_BYTES;
constexpr int BF_SDS = BF_SP + 2 * TILE_BYTES;
constexpr int BF_BAR = BF_SDS + 2 * TILE_BYTES;
constexpr int BF_SMEM = 1024 + BF_BAR + 256;
constexpr int BF_THREADS = 320;  // 8 compute warps + TMA warp + MMA warp
constexpr int BF_T_S = 0, BF_T_DP = 128, BF_T_DK = 256, BF_T_DV = 320, BF_T_DQ = 384;

struct AttnBwdFusedArgs {
    CUtensorMap tmQKV;   // (3*inner, T, B) bf16 box 64x128x1 (loads)
    CUtensorMap tmDO;    // (inner, T, B) bf16 box 64x128x1
    CUtensorMap tmDQKV;  // (3*inner, T, B) bf16 box 64x128x1 (stores of dK, dV)
    const float* lse;
    const float* delta;
    float* dq_accum;     // fp32 [B, T, inner], zeroed by the launcher
    int B, H, T;
    float scale, scale_log2e;
    int debug;  // experiment switches (SVIT_ATTN_DEBUG): 1 = skip dQ atomics, 2 = plain stores
};

__device__ long long g_attn_prof[256];
#define PROF(slot) do { if ((args.debug & 4) && blockIdx.x == 148 * 3) g_attn_prof[slot] = clock64(); } while (0)

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(BF_THREADS, 1) attn_bwd_fused_kernel(const __grid_constant__ AttnBwdFusedArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem + BF_SK;
    uint8_t* sV = smem + BF_SV;
    uint8_t* sQ = smem + BF_SQ;
    uint8_t* sdO = smem + BF_SDO;
    uint8_t* sP = smem + BF_SP;
    uint8_t* sdS = smem + BF_SDS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BF_BAR);
    uint64_t* bar_kv = bars + 0;
    uint64_t* qdo_full = bars + 1;    // [2]
    uint64_t* qdo_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_free = bars + 6;
    uint64_t* dp_full = bars + 7;
    uint64_t* dp_free = bars + 8;
    uint64_t* p_full = bars + 9;
    uint64_t* p_free = bars + 10;
    uint64_t* ds_full = bars + 11;
    uint64_t* d_done = bars + 12;
    uint64_t* dq_free = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = args.T, H = args.H;
    const int inner = H * 64;
    const int nblk = (T + 127) / 128;
    const int j = blockIdx.x % nblk;  // key block
    const int bh = blockIdx.x / nblk;
    const int b = bh / H, h = bh % H;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmQKV);
        tma_prefetch_desc(&args.tmDO);
        tma_prefetch_desc(&args.tmDQKV);
        mbar_init(bar_kv, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 256);
        mbar_init(dp_full, 1);
        mbar_init(dp_free, 256);
        mbar_init(p_full, 256);
        mbar_init(p_free, 1);
        mbar_init(ds_full, 256);
        mbar_init(d_done, 1);
        mbar_init(dq_free, 256);
        fence_mbar_init();
    }
    if (warp == 9) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            mbar_expect_tx(bar_kv, 2 * TILE_BYTES);
            tma_load_3d(sK, &args.tmQKV, bar_kv, inner + h * 64, j * 128, b);
            tma_load_3d(sV, &args.tmQKV, bar_kv, 2 * inner + h * 64, j * 128, b);
            for (int i = 0; i < nblk; ++i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&qdo_empty[s], ph ^ 1);
                mbar_expect_tx(&qdo_full[s], 2 * TILE_BYTES);
                tma_load_3d(sQ + s * TILE_BYTES, &args.tmQKV, &qdo_full[s], h * 64, i * 128, b);
                tma_load_3d(sdO + s * TILE_BYTES, &args.tmDO, &qdo_full[s], h * 64, i * 128, b);
            }
        }
    } else if (warp == 9) {
        // ================================ MMA issuer ================================
        if (elect_one()) {
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
            const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S / dP : K-major x K-major
            const uint32_t idesc_tn = umma_idesc_bf16(128, 64, 1, 1);   // dK / dV: A = P^T / dS^T (MN-major), B MN-major
            const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);   // dQ    : A = dS (K-major), B = K (MN-major)
            auto issue_a = [&](int i) {  // S = Q_i K^T
                const uint32_t q_addr = smem_u32(sQ + (i & 1) * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ss(tmem_base + BF_T_S, umma_smem_desc(q_addr + k * 32, 16, 1024),
                            umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
                umma_commit(s_full);
            };
            auto issue_b = [&](int i) {  // dP = dO_i V^T
                const uint32_t do_addr = smem_u32(sdO + (i & 1) * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ss(tmem_base + BF_T_DP, umma_smem_desc(do_addr + k * 32, 16, 1024),
                            umma_smem_desc(v_addr + k * 32, 16, 1024), idesc_s, k != 0);
                umma_commit(dp_full);
            };
            PROF(0);
            mbar_wait(bar_kv, 0);
            mbar_wait(&qdo_full[0], 0);
            PROF(1);
            tc_fence_after();
            issue_a(0);
            issue_b(0);
            PROF(2);
            for (int i = 0; i < nblk; ++i) {
                const uint32_t pi = i & 1;
                const uint32_t q_addr = smem_u32(sQ + (i & 1) * TILE_BYTES);
                const uint32_t do_addr = smem_u32(sdO + (i & 1) * TILE_BYTES);
                // c(i): dV += P^T dO_i
                PROF(10 + i * 10 + 0);
                mbar_wait(p_full, pi);
                PROF(10 + i * 10 + 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    umma_ss(tmem_base + BF_T_DV, umma_smem_desc(p_addr + s * 2048, TILE_BYTES, 1024),
                            umma_smem_desc(do_addr + s * 2048, 8192, 1024), idesc_tn, (i | s) != 0);
                umma_commit(p_free);
                PROF(10 + i * 10 + 2);
                // a(i+1)
                if (i + 1 < nblk) {
                    mbar_wait(&qdo_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
                    mbar_wait(s_free, pi);
                    tc_fence_after();
                    issue_a(i + 1);
                }
                // d(i): dK += dS^T Q_i ; dQ_part = dS K
                PROF(10 + i * 10 + 3);
                mbar_wait(ds_full, pi);
                PROF(10 + i * 10 + 4);
                if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    umma_ss(tmem_base + BF_T_DK, umma_smem_desc(ds_addr + s * 2048, TILE_BYTES, 1024),
                            umma_smem_desc(q_addr + s * 2048, 8192, 1024), idesc_tn, (i | s) != 0);
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    umma_ss(tmem_base + BF_T_DQ, umma_smem_desc(ds_addr + (s >> 2) * TILE_BYTES + (s & 3) * 32, 16, 1024),
                            umma_smem_desc(k_addr + s * 2048, 8192, 1024), idesc_dq, s != 0);
                umma_commit(d_done);
                umma_commit(&qdo_empty[i & 1]);
                PROF(10 + i * 10 + 5);
                // b(i+1)
                if (i + 1 < nblk) {
                    mbar_wait(dp_free, pi);
                    tc_fence_after();
                    issue_b(i + 1);
                }
            }
        }
    } else {
        // ================================ compute warps ================================
        const int q = warp & 3;      // TMEM lane quadrant
        const int half = warp >> 2;  // which 64 key columns of the S / dP tile (== which 64-key smem tile)
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int kvalid = T - j * 128 - half * 64;  // columns [0, kvalid) of this thread's 64 are real keys
        uint8_t* p_row = sP + half * TILE_BYTES + row * 128;
        uint8_t* ds_row = sdS + half * TILE_BYTES + row * 128;
        const float c = args.scale_log2e;

        auto dq_readout = [&](int i) {
            mbar_wait(d_done, i & 1);
            tc_fence_after();
            uint32_t dq[32];
            tmem_ld_32x32(t_row + BF_T_DQ + half * 32, dq);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(dq_free);
            const int t = i * 128 + row;
            if (t < T && !(args.debug & 1)) {
                float* dst = args.dq_accum + (static_cast<size_t>(b) * T + t) * inner + h * 64 + half * 32;
                if (args.debug & 2) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(dq[e]), __uint_as_float(dq[e + 1]),
                                                                          __uint_as_float(dq[e + 2]), __uint_as_float(dq[e + 3]));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        red_add_v4(dst + e, __uint_as_float(dq[e]), __uint_as_float(dq[e + 1]), __uint_as_float(dq[e + 2]),
                                   __uint_as_float(dq[e + 3]));
                }
            }
        };

        for (int i = 0; i < nblk; ++i) {
            const uint32_t pi = i & 1;
            const int t = i * 128 + row;
            const bool qvalid = t < T;
            const size_t sidx = (static_cast<size_t>(b) * H + h) * T + (qvalid ? t : 0);
            const float lse2 = args.lse[sidx] * 1.4426950408889634f;
            const float sdelta = args.delta[sidx] * args.scale;
            const bool full_tile = qvalid && kvalid >= 64;
            uint32_t pk[32];  // P of this thread's 64 columns, packed bf16
            // ---- P(i) = exp2(S*c - lse2) ----
            if (threadIdx.x == 0) PROF(100 + i * 10 + 0);
            mbar_wait(s_full, pi);
            if (threadIdx.x == 0) PROF(100 + i * 10 + 1);
            tc_fence_after();
            {
                uint32_t s0[32], s1[32];
                tmem_ld_32x32(t_row + BF_T_S + half * 64, s0);
                tmem_ld_32x32(t_row + BF_T_S + half * 64 + 32, s1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(s_free);
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(s0[e]), c, -lse2));
                    float p1 = ex2_approx(fmaf(__uint_as_float(s0[e + 1]), c, -lse2));
                    float p2 = ex2_approx(fmaf(__uint_as_float(s1[e]), c, -lse2));
                    float p3 = ex2_approx(fmaf(__uint_as_float(s1[e + 1]), c, -lse2));
                    if (!full_tile) {
                        if (!qvalid || e >= kvalid) p0 = 0.0f;
                        if (!qvalid || e + 1 >= kvalid) p1 = 0.0f;
                        if (!qvalid || 32 + e >= kvalid) p2 = 0.0f;
                        if (!qvalid || 32 + e + 1 >= kvalid) p3 = 0.0f;
                    }
                    pk[e / 2] = pack_bf16(p0, p1);
                    pk[16 + e / 2] = pack_bf16(p2, p3);
                }
            }
            if (i > 0) mbar_wait(p_free, (i - 1) & 1);  // c(i-1) finished reading the P tile
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint4 o = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
                *reinterpret_cast<uint4*>(p_row + ((g ^ (row & 7)) << 4)) = o;
            }
            fence_proxy_async_smem();
            mbar_arrive(p_full);
            if (threadIdx.x == 0) PROF(100 + i * 10 + 2);
            // ---- dQ read-out of the previous pair (its MMAs had the whole P(i) phase to finish) ----
            if (i > 0) dq_readout(i - 1);
            // ---- dS(i) = P * (dP*scale - delta*scale) ----
            if (threadIdx.x == 0) PROF(100 + i * 10 + 3);
            mbar_wait(dp_full, pi);
            if (threadIdx.x == 0) PROF(100 + i * 10 + 4);
            tc_fence_after();
            {
                uint32_t d0[32], d1[32];
                tmem_ld_32x32(t_row + BF_T_DP + half * 64, d0);
                tmem_ld_32x32(t_row + BF_T_DP + half * 64 + 32, d1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(dp_free);
                uint32_t dsk[32];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const uint32_t pa = pk[e / 2], pb = pk[16 + e / 2];
                    dsk[e / 2] = pack_bf16(bf16_lo(pa) * fmaf(__uint_as_float(d0[e]), args.scale, -sdelta),
                                           bf16_hi(pa) * fmaf(__uint_as_float(d0[e + 1]), args.scale, -sdelta));
                    dsk[16 + e / 2] = pack_bf16(bf16_lo(pb) * fmaf(__uint_as_float(d1[e]), args.scale, -sdelta),
                                                bf16_hi(pb) * fmaf(__uint_as_float(d1[e + 1]), args.scale, -sdelta));
                }
                // the dS tile is free: d(i-1) completion was observed in dq_readout(i-1)
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint4 o = make_uint4(dsk[g * 4], dsk[g * 4 + 1], dsk[g * 4 + 2], dsk[g * 4 + 3]);
                    *reinterpret_cast<uint4*>(ds_row + ((g ^ (row & 7)) << 4)) = o;
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(ds_full);
            if (threadIdx.x == 0) PROF(100 + i * 10 + 5);
        }
        dq_readout(nblk - 1);
        if (threadIdx.x == 0) PROF(150);
        // ---- epilogue: half 0 stores dK, half 1 stores dV (bf16, staged in the P tiles) ----
        // d_done of the last pair (observed above) implies every MMA of this CTA has retired
        {
            uint32_t o0[32], o1[32];
            const uint32_t acc = t_row + BF_T_DK + half * 64;
            tmem_ld_32x32(acc, o0);
            tmem_ld_32x32(acc + 32, o1);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t* src = g < 4 ? &o0[g * 8] : &o1[(g - 4) * 8];
                uint4 o;
                o.x = pack_bf16(__uint_as_float(src[0]), __uint_as_float(src[1]));
                o.y = pack_bf16(__uint_as_float(src[2]), __uint_as_float(src[3]));
                o.z = pack_bf16(__uint_as_float(src[4]), __uint_as_float(src[5]));
                o.w = pack_bf16(__uint_as_float(src[6]), __uint_as_float(src[7]));
                *reinterpret_cast<uint4*>(p_row + ((g ^ (row & 7)) << 4)) = o;
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 256);
        if (threadIdx.x == 0) {
            tma_store_3d(&args.tmDQKV, sP, inner + h * 64, j * 128, b);                    // dK
            tma_store_3d(&args.tmDQKV, sP + TILE_BYTES, 2 * inner + h * 64, j * 128, b);   // dV
            tma_store_commit();
            tma_store_wait_all<0>();
            PROF(151);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// dqkv[b, t, 0:inner] = bf16(dq_accum[b, t, :])
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int inner) {
    const int per_row = inner >> 2;
    const size_t total = rows * per_row;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = idx / per_row;
        const int cidx = static_cast<int>(idx - r * per_row);
        const float4 v = reinterpret_cast<const float4*>(acc)[idx];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dqkv + r * 3 * inner + cidx * 4) = o;
    }
}

// =================================================================================================
// host
// =================================================================================================
static int check_attn_shape(int B, int H, int T) {
    if (B <= 0 || H <= 0 || T <= 0 || T > ATT_MAX_T) {
        set_error("attention: unsupported shape B=%d H=%d T=%d (T must be in [1,%d])", B, H, T, ATT_MAX_T);
        return -1;
    }
    return 0;
}

int launch_attn_fwd(const AttnDesc& d, cudaStream_t stream) {
    if (check_attn_shape(d.B, d.H, d.T)) return -1;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_fwd) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    const int inner = d.H * 64;
    AttnFwdArgs a;
    memset(&a, 0, sizeof(a));
    int rc = 0;
    rc |= make_tmap_3d(&a.tmQKV, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, (uint64_t)3 * inner * 2,
                       (uint64_t)d.T * 3 * inner * 2, 64, 128);
    rc |= make_tmap_3d(&a.tmO, d.out, TmapDtype::BF16, inner, d.T, d.B, (uint64_t)inner * 2, (uint64_t)d.T * inner * 2, 64,
                       128);
    if (rc) {
        set_error("attn_fwd: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.lse = d.lse;
    a.B = d.B;
    a.H = d.H;
    a.T = d.T;
    a.scale = d.scale;
    a.scale_log2e = d.scale * 1.4426950408889634f;
    attn_fwd_kernel<<<d.B * d.H, FWD_THREADS, FWD_SMEM, stream>>>(a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_fwd launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    return 0;
}

int launch_attn_bwd(const AttnBwdDesc& d, cudaStream_t stream) {
    if (check_attn_shape(d.B, d.H, d.T)) return -1;
    static bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        cudaError_t e3 = cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_bwd) failed");
            return -10;
        }
        configured = true;
    }
    const int inner = d.H * 64;
    {
        const int rows = d.B * d.T;
        const int threads = 256;
        const int blocks = (rows * 32 + threads - 1) / threads;
        attn_delta_kernel<<<blocks, threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(d.out),
                                                          reinterpret_cast<const __nv_bfloat16*>(d.dout), d.delta, d.B,
                                                          d.H, d.T);
        count_launch();
    }
    const int nblk = (d.T + 127) / 128;
    if (d.dq_accum != nullptr) {
        AttnBwdFusedArgs a;
        memset(&a, 0, sizeof(a));
        int rc = 0;
        rc |= make_tmap_3d(&a.tmQKV, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, (uint64_t)3 * inner * 2,
                           (uint64_t)d.T * 3 * inner * 2, 64, 128);
        rc |= make_tmap_3d(&a.tmDO, d.dout, TmapDtype::BF16, inner, d.T, d.B, (uint64_t)inner * 2,
                           (uint64_t)d.T * inner * 2, 64, 128);
        rc |= make_tmap_3d(&a.tmDQKV, d.dqkv, TmapDtype::BF16, 3 * inner, d.T, d.B, (uint64_t)3 * inner * 2,
                           (uint64_t)d.T * 3 * inner * 2, 64, 128);
        if (rc) {
            set_error("attn_bwd: tensor map creation failed: %s", tmap_last_error());
            return -3;
        }
        a.lse = d.lse;
        a.delta = d.delta;
        a.dq_accum = d.dq_accum;
        a.B = d.B;
        a.H = d.H;
        a.T = d.T;
        a.scale = d.scale;
        a.scale_log2e = d.scale * 1.4426950408889634f;
        {
            const char* dbg = getenv("SVIT_ATTN_DEBUG");
            a.debug = dbg ? atoi(dbg) : 0;
        }
        const size_t rows = static_cast<size_t>(d.B) * d.T;
        cudaMemsetAsync(d.dq_accum, 0, rows * inner * sizeof(float), stream);
        attn_bwd_fused_kernel<<<d.B * d.H * nblk, BF_THREADS, BF_SMEM, stream>>>(a);
        size_t cblocks = (rows * (inner / 4) + 255) / 256;
        if (cblocks > 148 * 16) cblocks = 148 * 16;
        attn_dq_convert_kernel<<<static_cast<int>(cblocks), 256, 0, stream>>>(d.dq_accum,
                                                                             reinterpret_cast<__nv_bfloat16*>(d.dqkv), rows, inner);
        count_launch(2);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("attn_bwd (fused) launch failed: %s", cudaGetErrorString(e));
            return -11;
        }
        return 0;
    }
    AttnBwdArgs a;
    memset(&a, 0, sizeof(a));
    int rc = 0;
    rc |= make_tmap_3d(&a.tmQKV, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, (uint64_t)3 * inner * 2,
                       (uint64_t)d.T * 3 * inner * 2, 64, 128);
    rc |= make_tmap_3d(&a.tmDO, d.dout, TmapDtype::BF16, inner, d.T, d.B, (uint64_t)inner * 2, (uint64_t)d.T * inner * 2,
                       64, 128);
    rc |= make_tmap_3d(&a.tmDQKV, d.dqkv, TmapDtype::BF16, 3 * inner, d.T, d.B, (uint64_t)3 * inner * 2,
                       (uint64_t)d.T * 3 * inner * 2, 64, 128);
    if (rc) {
        set_error("attn_bwd: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.lse = d.lse;
    a.delta = d.delta;
    a.B = d.B;
    a.H = d.H;
    a.T = d.T;
    a.scale = d.scale;
    a.scale_log2e = d.scale * 1.4426950408889634f;
    attn_bwd_kernel<false><<<d.B * d.H * nblk, BWD_THREADS, BWD_SMEM, stream>>>(a);
    attn_bwd_kernel<true><<<d.B * d.H * nblk, BWD_THREADS, BWD_SMEM, stream>>>(a);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("attn_bwd launch failed: %s", cudaGetErrorString(e));
        return -11;
    }
    return 0;
}

int debug_read_attn_prof(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, g_attn_prof, sizeof(long long) * (n < 256 ? n : 256)) == cudaSuccess ? 0 : -1;
}

}  // namespace svit
