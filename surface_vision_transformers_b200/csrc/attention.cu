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
// One CTA per (sample, head), TWO CTAs resident per SM.  K and V (<= 384 keys) stay in shared memory, Q blocks of 128
// rows stream through.  Inside a CTA the keys of a query block are processed in chunks of 96 that ping-pong between two
// TMEM score buffers and two groups of four softmax warps:
//     tensor pipe :  S(0) S(1) | PV(0) S(2) | PV(1) S(3) | PV(2) PV(3)             S = Q K_c^T, O += P_c V_c
//     group 0     :       exp(0) ........ exp(2) ........
//     group 1     :            exp(1) ........ exp(3) ....
// so the exponentials of one chunk run while the tensor core produces the next scores and consumes the previous
// probabilities, and what a CTA still cannot overlap (the load of the next Q tile, the block epilogue) the second CTA on
// the SM fills.  Two CTAs per SM need half the TMEM and half the shared memory each:
//   * TMEM: 2 x 96 score columns + 64 O columns = 256;
//   * shared memory holds exactly the padded keys (tk = T rounded up to 16 rows) and ONE Q tile; O leaves straight from
//     registers (64 contiguous bytes per thread and block -- tiny next to the operand streams).
// The scores are read from TMEM ONCE: the softmax shift is the row's score against key 0 (any shift is exact for
// softmax; the log-sum-exp is reported with the same shift), which makes the chunks independent -- no running maximum,
// no rescaling of O.  A guard on the row sum (it cannot underflow: key 0 contributes exp2(0) = 1) detects overflow at
// the end of the block; the block is then redone in "safe mode" with the classic max shift (a pass over all chunks for
// the maxima, then the exponentials).
// A warp owns one TMEM lane quadrant (32 query rows, one per thread) and ALL columns of its chunk, so the bf16
// probabilities go back into TMEM over the scores just read (tcgen05.st, two per 32-bit column) without any cross-warp
// hazard, and feed the PV product as the A operand FROM TMEM: P never touches shared memory.
// shared memory map (bytes):  sQ 16K | sK tk*128 | sV tk*128 | shift / row sums 1.5K | barriers   (T = 321: 103 KB)
constexpr int FWD_THREADS = 288;
constexpr int FWD_TMEM_COLS = 256;
constexpr int FWD_CHUNK = 96;    // keys per chunk = score columns per buffer
constexpr int FWD_TMEM_O = 192;  // O accumulator columns [192, 256)

__host__ __device__ constexpr int fwd_smem_bytes(int tk) { return 1024 + TILE_BYTES + 2 * tk * 128 + 3 * 128 * 4 + 256; }

struct AttnFwdArgs {
    CUtensorMap tmQ;   // qkv (3*inner, T, B) bf16, box 64 x 128 x 1: Q_i loads
    CUtensorMap tmKV;  // qkv                      box 64 x kv_box x 1: K / V loads (kv_box divides the padded key count)
    __nv_bfloat16* out;  // [B, T, inner]
    float* lse;
    int B, H, T;
    int kv_box;
    float scale, scale_log2e;
};

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial for 2^f on [-0.5, 0.5], max relative error
// 7.6e-5 -- far below the bf16 rounding of the probabilities): every fourth exponential of the forward pass takes this
// path and relieves the MUFU (16 ex2 / clk / SM).
__device__ __forceinline__ float fwd_ex2_poly(float x) {
    x = fminf(fmaxf(x, -126.0f), 126.0f);
    const float t = x + 12582912.0f;              // 1.5 * 2^23: round(x) now sits in the low mantissa bits
    const float f = x - (t - 12582912.0f);        // in [-0.5, 0.5]
    float p = fmaf(f, 0.05522261559963226f, 0.24261537194252014f);
    p = fmaf(p, f, 0.6932516098022461f);
    p = fmaf(p, f, 0.9999275803565979f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // * 2^round(x)
}
__device__ __forceinline__ float fwd_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(384, 2) attn_fwd_kernel(const __grid_constant__ AttnFwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int T = args.T, H = args.H;
    const int inner = H * 64;
    const int tk = (T + 15) & ~15;  // keys padded to the UMMA K step
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;
    uint8_t* sV = sK + tk * 128;
    float* sShift = reinterpret_cast<float*>(sV + tk * 128);  // [128]   softmax shift of every query row of the block
    float* sRed = sShift + 128;                                // [2][128] per-group partial row sums / maxima
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 256);
    uint64_t* bar_k = bars + 0;     // K landed                                      (TMA -> control)
    uint64_t* bar_vv = bars + 1;    // V landed                                      (TMA -> control)
    uint64_t* bar_q = bars + 2;     // Q_i landed                                    (TMA -> control)
    uint64_t* s_full = bars + 3;    // [2] scores of a chunk in buffer b             (MMA -> softmax group b, control)
    uint64_t* p_full = bars + 5;    // [2] P of a chunk in buffer b / scores consumed (softmax group b -> control)
    uint64_t* pv_done = bars + 7;   // [2] the PV product that read buffer b retired (MMA -> control)
    uint64_t* bar_o = bars + 9;     // every MMA of the block retired                (MMA -> softmax, control)
    uint64_t* bar_of = bars + 10;   // O copied to registers                         (softmax -> control)
    uint64_t* bar_v = bars + 11;    // verdict of the block published in *flag       (softmax -> control)
    uint32_t* flag = reinterpret_cast<uint32_t*>(bars + 12);
    uint32_t* tmem_slot = flag + 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x / H;
    const int h = blockIdx.x % H;
    const int nqb = (T + 127) / 128;                  // query blocks
    const int nch = (tk + FWD_CHUNK - 1) / FWD_CHUNK;  // key chunks (<= 4); chunk c lives in buffer c & 1, group c & 1

    griddep_launch();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmQ);
        tma_prefetch_desc(&args.tmKV);
        mbar_init(bar_k, 1);
        mbar_init(bar_vv, 1);
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        mbar_init(bar_o, 1);
        mbar_init(bar_of, 8);
        mbar_init(bar_v, 8);
        *flag = 0;
        fence_mbar_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, FWD_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (warp == 8) {
        // ============================ control: TMA + MMA issue ============================
        if (elect_one()) {
            const int nbox = tk / args.kv_box;
            mbar_expect_tx(bar_k, tk * 128);
            for (int r = 0; r < nbox; ++r)
                tma_load_3d(sK + r * args.kv_box * 128, &args.tmKV, bar_k, inner + h * 64, r * args.kv_box, b);
            mbar_expect_tx(bar_q, TILE_BYTES);
            tma_load_3d(sQ, &args.tmQ, bar_q, h * 64, 0, b);
            mbar_expect_tx(bar_vv, tk * 128);
            for (int r = 0; r < nbox; ++r)
                tma_load_3d(sV + r * args.kv_box * 128, &args.tmKV, bar_vv, 2 * inner + h * 64, r * args.kv_box, b);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            uint32_t ph_q = 0, ph_o = 0, ph_of = 0, ph_v = 0;
            // per-buffer state as scalars (an indexed local array would live in local memory: with ~210 KB of the SM's
            // unified L1 / shared memory given to the two CTAs' tiles, every such access is an L2 round trip)
            uint32_t ph_p0 = 0, ph_p1 = 0, ph_pv0 = 0, ph_pv1 = 0;
            uint32_t s_cnt0 = 0, s_cnt1 = 0;  // score batches issued into buffer b so far: batch n completes parity n & 1 of s_full[b]
            auto chunk_keys = [&](int c) { return min(FWD_CHUNK, tk - c * FWD_CHUNK); };
            // S(c) = Q K_c^T (K-major A and B) into score buffer c & 1
            auto issue_s = [&](int c) {
                const int bf = c & 1;
                const uint32_t idesc = umma_idesc_bf16(128, chunk_keys(c), 0, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ss(tmem_base + bf * FWD_CHUNK, umma_smem_desc(q_addr + k * 32, 16, 1024),
                            umma_smem_desc(k_addr + c * FWD_CHUNK * 128 + k * 32, 16, 1024), idesc, k != 0);
                umma_commit(&s_full[bf]);
                if (bf) ++s_cnt1; else ++s_cnt0;
            };
            // O (+)= P_c V_c  (A = P from TMEM: lanes = query rows, 8 columns of packed bf16 pairs per 16-key step; B = V MN-major)
            auto issue_pv = [&](int c) {
                const int bf = c & 1, ks = chunk_keys(c) >> 4;
                for (int s = 0; s < ks; ++s)
                    umma_ts(tmem_base + FWD_TMEM_O, tmem_base + bf * FWD_CHUNK + s * 8,
                            umma_smem_desc(v_addr + (c * (FWD_CHUNK >> 4) + s) * 2048, 8192, 1024), idesc_pv,
                            (c > 0 || s != 0) ? 1u : 0u);
            };
            auto load_q = [&](int i) {
                mbar_expect_tx(bar_q, TILE_BYTES);
                tma_load_3d(sQ, &args.tmQ, bar_q, h * 64, i * 128, b);
            };
            auto wait_p = [&](int bf) {
                mbar_wait(&p_full[bf], bf ? ph_p1 : ph_p0);
                if (bf) ph_p1 ^= 1; else ph_p0 ^= 1;
            };
            // the last score batch of the block has retired -> sQ is dead: the next block's queries may land
            auto prefetch_q = [&](int i) {
                const int bf = (nch - 1) & 1;
                mbar_wait(&s_full[bf], ((bf ? s_cnt1 : s_cnt0) - 1) & 1);
                if (i + 1 < nqb) load_q(i + 1);
            };
            // one pass over the chunks of block i: scores for everybody, and (exp pass) the PV products behind them
            auto run_pass = [&](int i, bool exp_pass, bool first_attempt) {
                auto do_pv = [&](int c) {
                    const int bf = c & 1;
                    wait_p(bf);
                    if (!exp_pass) return;
                    if (c == 0 && first_attempt) {
                        if (i > 0) {
                            mbar_wait(bar_of, ph_of);  // the previous block's O left TMEM
                            ph_of ^= 1;
                        } else {
                            mbar_wait(bar_vv, 0);
                        }
                    }
                    tc_fence_after();
                    issue_pv(c);
                    if (c + 2 < nch) umma_commit(&pv_done[bf]);
                };
                for (int c = 0; c < nch; ++c) {
                    const int bf = c & 1;
                    if (c >= 2 && exp_pass) {
                        mbar_wait(&pv_done[bf], bf ? ph_pv1 : ph_pv0);  // PV(c-2) consumed P(c-2): the buffer is free
                        if (bf) ph_pv1 ^= 1; else ph_pv0 ^= 1;
                    }
                    tc_fence_after();
                    issue_s(c);
                    if (c >= 1) do_pv(c - 1);   // (max pass, c >= 2: this is also the "buffer consumed" wait for S(c+1))
                    if (c == nch - 1 && exp_pass) prefetch_q(i);
                }
                do_pv(nch - 1);
            };
            mbar_wait(bar_k, 0);
            for (int i = 0; i < nqb; ++i) {
                mbar_wait(bar_q, ph_q);
                ph_q ^= 1;
                run_pass(i, true, true);
                umma_commit(bar_o);
                mbar_wait(bar_o, ph_o);
                ph_o ^= 1;
                mbar_wait(bar_v, ph_v);
                ph_v ^= 1;
                if (*reinterpret_cast<volatile uint32_t*>(flag) != 0) {
                    // ---- safe mode: redo block i with the max shift (see the softmax warps) ----
                    if (i + 1 < nqb) {
                        mbar_wait(bar_q, ph_q);  // the prefetched Q_{i+1} has landed; Q_i comes back first
                        ph_q ^= 1;
                    }
                    if (nqb > 1) {
                        load_q(i);
                        mbar_wait(bar_q, ph_q);
                        ph_q ^= 1;
                    }
                    run_pass(i, false, false);
                    run_pass(i, true, false);
                    umma_commit(bar_o);
                    mbar_wait(bar_o, ph_o);
                    ph_o ^= 1;
                }
            }
        }
    } else {
        // ============================ softmax + epilogue warps ============================
        const int q = warp & 3;     // TMEM lane quadrant
        const int grp = warp >> 2;  // softmax group = score buffer = parity of the chunks this warp processes
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t t_buf = t_row + grp * FWD_CHUNK;
        const float c = args.scale_log2e;
        uint32_t ph_s = 0, ph_o = 0;
        // the two warps of a quadrant (one per group) hold the same query rows
        // (constant barrier ids: with a run-time id ptxas reserves all 16 named barriers and only one CTA fits an SM)
        auto pairbar = [&]() {
            if (q == 0) named_bar_sync(1, 64);
            else if (q == 1) named_bar_sync(2, 64);
            else if (q == 2) named_bar_sync(3, 64);
            else named_bar_sync(4, 64);
        };
        auto pair_exchange = [&](float mine, bool is_max) {
            sRed[grp * 128 + row] = mine;
            pairbar();
            const float a = sRed[row], bb = sRed[128 + row];
            pairbar();
            return is_max ? fmaxf(a, bb) : a + bb;
        };
        auto wait_scores = [&]() {
            mbar_wait(&s_full[grp], ph_s);
            ph_s ^= 1;
            tc_fence_after();
        };
        // exp2((s - shift) * c) of chunk ch, 32 columns at a time; the packed bf16 pairs go straight back into TMEM over
        // the scores just read (pair (2k, 2k+1) of the chunk's columns lands in column k).  Returns the partial row sum.
        auto softmax_chunk = [&](int ch, float shift_c) {
            const int nk = min(FWD_CHUNK, tk - ch * FWD_CHUNK);  // columns of this chunk
            const int nreal = T - ch * FWD_CHUNK;                // real (un-padded) keys from column 0 on
            float sm0 = 0.0f, sm1 = 0.0f;
#pragma unroll
            for (int g = 0; g < FWD_CHUNK / 32; ++g) {
                const int c0 = g * 32;
                if (c0 < nk) {
                    uint32_t r[32], pk[16];
                    tmem_ld_32x32(t_buf + c0, r);
                    tmem_ld_wait();
                    const int nv = min(nk, nreal) - c0;  // real keys among the 32 columns of this group
                    if (nv >= 32) {
                        // The row sum adds the fp32 probabilities (their bf16 rounding in the PV product is unbiased; the
                        // difference is ~1e-4 relative, far below bf16 resolution)
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            const float p0 = fwd_ex2(fmaf(__uint_as_float(r[e]), c, -shift_c));
                            const float x1 = fmaf(__uint_as_float(r[e + 1]), c, -shift_c);
                            const float p1 = ((e >> 1) & 1) ? fwd_ex2_poly(x1) : fwd_ex2(x1);  // every 4th on the FMA pipe
                            pk[e / 2] = pack_bf16(p0, p1);
                            sm0 += p0;
                            sm1 += p1;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            float p0 = fwd_ex2(fmaf(__uint_as_float(r[e]), c, -shift_c));
                            float p1 = fwd_ex2(fmaf(__uint_as_float(r[e + 1]), c, -shift_c));
                            if (e >= nv) p0 = 0.0f;
                            if (e + 1 >= nv) p1 = 0.0f;
                            pk[e / 2] = pack_bf16(p0, p1);
                            sm0 += p0;
                            sm1 += p1;
                        }
                    }
                    tmem_st_32x16(t_buf + (c0 >> 1), pk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(&p_full[grp]);
            return sm0 + sm1;
        };
        auto max_chunk = [&](int ch) {
            const int nk = min(FWD_CHUNK, tk - ch * FWD_CHUNK);
            const int nreal = T - ch * FWD_CHUNK;
            float mx = -INFINITY;
            for (int c0 = 0; c0 < nk; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(t_buf + c0, r);
                tmem_ld_wait();
                const int nv = min(nk, nreal) - c0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
            tc_fence_before();
            mbar_arrive_warp(&p_full[grp]);  // scores consumed
            return mx;
        };

        for (int i = 0; i < nqb; ++i) {
            float shift = 0.0f, total = 0.0f;
            bool safe = false;
#pragma unroll 1
            for (;;) {
                bool have_scores = false;  // group 0 has already waited for chunk 0 (to read the shift)
                if (!safe) {
                    // shift = the row's score against key 0 (column 0 of chunk 0, which group 0 owns)
                    if (grp == 0) {
                        wait_scores();
                        have_scores = true;
                        shift = __uint_as_float(tmem_ld_32x1(t_buf));
                        tmem_ld_wait();
                        sShift[row] = shift;
                    }
                    pairbar();
                    if (grp == 1) shift = sShift[row];
                } else {
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int ch = grp; ch < nch; ch += 2) {
                        wait_scores();
                        mx = fmaxf(mx, max_chunk(ch));
                    }
                    shift = pair_exchange(mx, true);
                }
                float part = 0.0f;
#pragma unroll 1
                for (int ch = grp; ch < nch; ch += 2) {
                    if (!have_scores) wait_scores();
                    have_scores = false;
                    part += softmax_chunk(ch, shift * c);
                }
                total = pair_exchange(part, false);
                if (safe) break;
                // The row sum cannot underflow (key 0 contributes exp2(0) = 1), so "not < 1e30" (overflow, inf or NaN) is
                // the whole test; the verdict is uniform for the CTA and published to the control warp.
                const bool bad = named_bar_or(5, 256, !(total < 1e30f));
                if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(flag) = bad ? 1u : 0u;
                mbar_arrive_warp(bar_v);
                if (!bad) break;
                mbar_wait(bar_o, ph_o);  // the failed attempt's MMAs have retired
                ph_o ^= 1;
                safe = true;
            }
            mbar_wait(bar_o, ph_o);
            ph_o ^= 1;
            // ---- epilogue: O / sum -> bf16 -> global, straight from registers (each group converts 32 of the 64 columns) ----
            tc_fence_after();
            uint32_t o0[32];
            tmem_ld_32x32(t_row + FWD_TMEM_O + grp * 32, o0);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive_warp(bar_of);
            const float inv = 1.0f / total;
            const int t = i * 128 + row;
            if (t < T) {
                uint4* dst = reinterpret_cast<uint4*>(args.out + (static_cast<size_t>(b) * T + t) * inner + h * 64 + grp * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t* src = &o0[g * 8];
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
                    o.y = pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
                    o.z = pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
                    o.w = pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
                    dst[g] = o;
                }
                if (grp == 0) args.lse[(static_cast<size_t>(b) * H + h) * T + t] = shift * args.scale + logf(total);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FWD_TMEM_COLS);
    }
}

// clock64 timeline of one CTA for scripts/prof_attn_bwd.py; costs nothing unless enabled (SVIT_ATTN_DEBUG=4)
__device__ long long g_attn_prof[256];

// =================================================================================================
// backward: one CTA per (sample, head), every operand loaded exactly once, no atomics
// =================================================================================================
// Keys are processed in blocks of 96 (j, outer loop, K_j / V_j double-buffered by TMA), queries in blocks of 128
// (i, inner loop, all Q_i / dO_i resident in shared memory).  With 96-column score tiles everything fits in TMEM:
//   S [0,96)  dP [96,192)  dK_j [192,256)  dV_j [256,320)  dQ_0..2 [320,512)
// so dQ accumulates on chip across the key blocks (no atomics, no fp32 scratch, no separate delta / convert kernels).
//
//   tensor core (two issuing warps X, Y)     16 compute warps: warp (q, slab) = TMEM lane quadrant q (one query row per
//   a(p): S  = Q_i K_j^T          (X)         thread) x key columns [24 slab, 24 slab + 24) of the 96-key block
//   b(p): dP = dO_i V_j^T         (Y)           P  = exp2(S*scale*log2e - lse*log2e)    -> bf16 smem tile sP
//   c(p): dV_j += P^T dO_i        (X)           dS = P * (dP*scale - delta*scale)       -> bf16 smem tile sdS
//   d(p): dK_j += dS^T Q_i ; dQ_i += dS K_j (Y)
// Step p of the compute warps: dS(p) from dP(p) and the P(p) they kept in registers (packed bf16) since the previous step,
// hand it to the tensor core, then P(p+1) from S(p+1).  S(p+1) and dP(p) were issued one step earlier, as soon as their
// TMEM buffers had been read out, so a step never waits for the tensor pipe; c and d of a step run underneath the next.
// Row masking costs nothing (lse = +inf for rows past T), column masking only touches the one slab that straddles T.
// What the profile (SVIT_ATTN_DEBUG=4, scripts/prof_attn_bwd.py) taught, in clocks per CTA at T = 321 (63 k at first):
//   * delta = rowsum(dO * O): per-thread global loads of the O / dO rows took 11 k; O_i now arrives by TMA in the not yet
//     used dS tiles and delta is read from shared memory (1.5 k);
//   * dK_j / dV_j leave through the block's own K / V stage and two TMA stores (the stage is refilled one step later):
//     per-thread row stores took 3 k per key block, and the LSU starves while the tensor core streams smem operands;
//   * the timeout printf of mbar_wait cost every inlined wait a call site; spills inside the step loop are fatal
//     (with 227 KB of the unified L1 / shared memory taken, local memory is an L2 round trip);
//   * round 2: 12 warps on 32-column slabs running dS(p) and P(p+1) as one hand-merged stream (152 registers via
//     setmaxnreg) took 287 us; 16 warps on 24-column slabs doing dS then P (90 registers, four warps per sub-partition
//     hide the TMEM-load / MUFU latencies) take 275 us.  A step is still ~2,300 clocks against ~1,500 of tensor work: the
//     phases of a step (dP load + dS 650, dS hand-off 500, S load + exponentials + P hand-off 1,150) are latency chains
//     through ONE P tile and ONE S / dP buffer each; letting half of the warps do the exponentials first (so that MUFU
//     and FMA phases overlap across warps) did not shorten them (291 us).  The next step would be a second S / dP / P
//     set, for which neither TMEM (512 columns used) nor shared memory (227 KB used) has room at 96-key blocks.
// A [128 x 96] bf16 tile occupies one full 64-column swizzled tile plus half of a second one; the two dS buffers
// share that second tile (buffer 1 lives in its columns 32..63, i.e. 64 bytes into every row).
// smem: sQ[3] sdO[3] (16K each) | sK,sV x2 stages (12K each) | sP 32K | sdS 16K + 16K + 16K shared | barriers, delta (~227 KB)
constexpr int BK_KEYS = 96;
constexpr int BK_KV_TILE = BK_KEYS * 128;  // [96 keys x 64 bf16] swizzled tile
constexpr int BK_SQ = 0;
constexpr int BK_SDO = BK_SQ + 3 * TILE_BYTES;
constexpr int BK_SKV = BK_SDO + 3 * TILE_BYTES;  // stage s: K at + s * 2 * BK_KV_TILE, V right behind it
constexpr int BK_SP = BK_SKV + 4 * BK_KV_TILE;
constexpr int BK_SDS = BK_SP + 2 * TILE_BYTES;   // buffer b: keys 0..63 in tile b, keys 64..95 in tile 2 at byte 64 * b of each row
constexpr int BK_BAR = BK_SDS + 3 * TILE_BYTES;
constexpr int BK_DELTA = BK_BAR + 256;          // [3][128] fp32: delta * scale of every query row
constexpr int BK_SMEM = 1024 + BK_DELTA + 3 * 128 * 4;
static_assert(BK_SMEM <= 232448, "attn_bwd shared memory");
constexpr int BK_COMPUTE_WARPS = 16;
constexpr int BK_COMPUTE_THREADS = 32 * BK_COMPUTE_WARPS;
constexpr int BK_THREADS = 128 + BK_COMPUTE_THREADS;  // warp group 0: MMA warps X and Y (+ 2 idle warps), then 16 compute warps
constexpr int BK_SLAB = BK_KEYS / (BK_COMPUTE_WARPS / 4);  // 24 key columns per compute warp
constexpr int BK_T_S = 0, BK_T_DP = 96, BK_T_DK = 192, BK_T_DV = 256, BK_T_DQ = 320;

struct AttnBwdArgs {
    CUtensorMap tmQ;    // qkv  (3*inner, T, B) bf16 box 64 x 128: Q_i loads
    CUtensorMap tmDO;   // dout (inner, T, B)   bf16 box 64 x 128
    CUtensorMap tmKV;   // qkv                  bf16 box 64 x 96 : K_j, V_j loads
    CUtensorMap tmDQ;   // dqkv (3*inner, T, B) bf16 box 64 x 128: dQ_i stores
    CUtensorMap tmO;    // out  (inner, T, B)   bf16 box 64 x 128: O_i loads (delta only)
    CUtensorMap tmDKV;  // dqkv                 bf16 box 64 x 96 : dK_j, dV_j stores
    const float* lse;
    int B, H, T;
    float scale, scale_log2e;
    int debug;  // SVIT_ATTN_DEBUG: 4 = record a clock64 timeline of one CTA (svit_debug_attn_prof); 8 = dead warps do the full
                // arithmetic like round 1 (A/B of the dead-pair shortcut)
};

#define PROF(slot) do { if ((args.debug & 4) && blockIdx.x == 148 * 3) g_attn_prof[slot] = clock64(); } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
    return bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
           bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
}

__global__ void __launch_bounds__(BK_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + BK_SQ;
    uint8_t* sdO = smem + BK_SDO;
    uint8_t* sKV = smem + BK_SKV;
    uint8_t* sP = smem + BK_SP;
    uint8_t* sdS = smem + BK_SDS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BK_BAR);
    uint64_t* q_full = bars + 0;     // [3] Q_i and dO_i landed (once per CTA)
    uint64_t* kv_full = bars + 4;    // [2]
    uint64_t* s_full = bars + 8;     // S in TMEM                      (MMA -> compute warps)
    uint64_t* s_free = bars + 9;     // S copied to registers          (compute warps -> MMA)
    uint64_t* dp_full = bars + 10;   // dP in TMEM                     (MMA -> compute warps)
    uint64_t* dp_free = bars + 11;   // dP copied to registers         (compute warps -> MMA)
    uint64_t* p_full = bars + 12;    // P tile written                 (compute warps -> MMA)
    uint64_t* p_free = bars + 13;    // c(p) retired                   (MMA X -> compute warps)
    uint64_t* ds_full = bars + 14;   // [2] dS tile p & 1 written      (compute warps -> MMA)
    uint64_t* ds_free = bars + 16;   // [2] d(p) retired               (MMA -> compute warps)
    uint64_t* dv_full = bars + 18;   // dV_j final (c(j, last) retired)  (MMA X -> compute warps)
    uint64_t* dv_free = bars + 19;   // dV_j copied to registers         (compute warps -> MMA X)
    uint64_t* dk_full = bars + 20;   // dK_j final (d(j, last) retired)  (MMA Y -> compute warps)
    uint64_t* dk_free = bars + 21;   // dK_j copied to registers         (compute warps -> MMA Y)
    uint64_t* dq_full = bars + 22;   // every MMA of the CTA retired     (MMA -> compute warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = args.T, H = args.H;
    const int inner = H * 64;
    const int nqb = (T + 127) / 128;
    const int nkb = (T + BK_KEYS - 1) / BK_KEYS;
    const int total = nqb * nkb;
    const int b = blockIdx.x / H, h = blockIdx.x % H;

    griddep_launch();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&args.tmQ);
        tma_prefetch_desc(&args.tmDO);
        tma_prefetch_desc(&args.tmKV);
        tma_prefetch_desc(&args.tmDQ);
        tma_prefetch_desc(&args.tmO);
        tma_prefetch_desc(&args.tmDKV);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&ds_full[s], BK_COMPUTE_WARPS);
            mbar_init(&ds_free[s], 1);
        }
        for (int i = 0; i < 3; ++i) mbar_init(&q_full[i], 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, BK_COMPUTE_WARPS);
        mbar_init(dp_full, 1);
        mbar_init(dp_free, BK_COMPUTE_WARPS);
        mbar_init(p_full, BK_COMPUTE_WARPS);
        mbar_init(p_free, 1);
        mbar_init(dv_full, 1);
        mbar_init(dv_free, 8);
        mbar_init(dk_full, 1);
        mbar_init(dk_free, 8);
        mbar_init(dq_full, 1);
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
    griddep_wait();
    // The issuing warps need few registers, the compute warps many: move them (per warp group of 4 warps).  Each
    // setmaxnreg sits at the top of the role branch it governs: ptxas allocates a region by the setmaxnreg that dominates
    // it (after a common if / else it may fall back to the smaller limit).

    auto load_kv = [&](int j) {  // K_j, V_j -> stage j & 1 (called by one thread)
        const int s = j & 1;
        uint8_t* dst = sKV + s * 2 * BK_KV_TILE;
        mbar_expect_tx(&kv_full[s], 2 * BK_KV_TILE);
        tma_load_3d(dst, &args.tmKV, &kv_full[s], inner + h * 64, j * BK_KEYS, b);
        tma_load_3d(dst + BK_KV_TILE, &args.tmKV, &kv_full[s], 2 * inner + h * 64, j * BK_KEYS, b);
    };

    if (warp < 4) {
      if (warp == 0) {
        // ---- initial loads: all Q_i / dO_i and the first two K/V blocks ----
        if (elect_one()) {
            auto load_q = [&](int i) {
                mbar_expect_tx(&q_full[i], 3 * TILE_BYTES);
                tma_load_3d(sQ + i * TILE_BYTES, &args.tmQ, &q_full[i], h * 64, i * 128, b);
                tma_load_3d(sdO + i * TILE_BYTES, &args.tmDO, &q_full[i], h * 64, i * 128, b);
                tma_load_3d(sdS + i * TILE_BYTES, &args.tmO, &q_full[i], h * 64, i * 128, b);  // O_i: only for delta, before any dS
            };
            load_q(0);
            load_kv(0);
            for (int i = 1; i < nqb; ++i) load_q(i);
            if (nkb > 1) load_kv(1);
        }
      }
      if (warp < 2) {
        // ================================ MMA issuers ================================
        // warp 0 (X): a(p+1) = S, c(p) = dV.   warp 1 (Y): b(p+1) = dP, d(p) = dQ, dK.
        // Descriptors are (lo, hi) 32-bit pairs; a K-step adds a constant to lo (see ptx.cuh).
        if (elect_one()) {
            const bool X = warp == 0;
            constexpr uint32_t hi = umma_desc_hi(1024);
            const uint32_t idesc_tn = umma_idesc_bf16(128, 64, 1, 1);   // dK / dV: A = P^T / dS^T (MN-major), B MN-major
            const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);   // dQ    : A = dS (K-major), B = K (MN-major)
            // K-major operand tile (Q, dO, K, V): K-step = 32 B;  MN-major operand: K-step = 16 rows = 2048 B
            const uint32_t kmaj0 = umma_desc_lo(smem_u32(X ? sQ : sdO), 16);          // A of a / b, + slot * TILE
            const uint32_t kv0 = umma_desc_lo(smem_u32(sKV + (X ? 0 : BK_KV_TILE)), 16);  // B of a (K_j) / b (V_j), + stage
            const uint32_t tr0 = umma_desc_lo(smem_u32(X ? sP : sdS), TILE_BYTES);    // A of c / dK: P^T / dS^T (+ dS buffer)
            const uint32_t mn0 = umma_desc_lo(smem_u32(X ? sdO : sQ), 8192);          // B of c / dK: dO_i / Q_i, + slot * TILE
            const uint32_t dsk0 = umma_desc_lo(smem_u32(sdS), 16);                    // A of dQ: dS, K-major (+ dS buffer)
            const uint32_t kmn0 = umma_desc_lo(smem_u32(sKV), 8192);                  // B of dQ: K_j MN-major, + stage
            auto keys_in = [&](int j) { return min(BK_KEYS, T - j * BK_KEYS); };
            auto rows_in = [&](int i) { return min(128, T - i * 128); };
            auto issue_ab = [&](int j, int i) {  // X: S = Q_i K_j^T      Y: dP = dO_i V_j^T
                const uint32_t a_lo = kmaj0 + i * (TILE_BYTES >> 4);
                const uint32_t b_lo = kv0 + (j & 1) * (2 * BK_KV_TILE >> 4);
                const uint32_t idesc = umma_idesc_bf16(128, (keys_in(j) + 15) & ~15, 0, 0);
                const uint32_t d = tmem_base + (X ? BK_T_S : BK_T_DP);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss_lohi(d, a_lo + k * 2, b_lo + k * 2, hi, idesc, k != 0);
                umma_commit(X ? s_full : dp_full);
            };
            if (X) PROF(0);
            mbar_wait(&q_full[0], 0);
            mbar_wait(&kv_full[0], 0);
            if (X) PROF(1);
            tc_fence_after();
            issue_ab(0, 0);
            int p = 0;
            for (int j = 0; j < nkb; ++j) {
                const int ks_k = (keys_in(j) + 15) >> 4;  // K steps over the keys of this block
                for (int i = 0; i < nqb; ++i, ++p) {
                    const uint32_t pp = p & 1;
                    const bool has_next = p + 1 < total;
                    const int jn = (i + 1 < nqb) ? j : j + 1;
                    const int in = (i + 1 < nqb) ? i + 1 : 0;
                    const int ks_q = (rows_in(i) + 15) >> 4;  // K steps over the queries of this block
                    if (has_next) {
                        mbar_wait(X ? s_free : dp_free, pp);
                        if (in == 0) mbar_wait(&kv_full[jn & 1], (jn >> 1) & 1);
                        mbar_wait(&q_full[in], 0);
                        tc_fence_after();
                        issue_ab(jn, in);
                    }
                    if (p < 8) PROF((X ? 10 : 12) + p * 4);
                    if (X) mbar_wait(p_full, pp);
                    else mbar_wait(&ds_full[pp], (p >> 1) & 1);
                    if (p < 8) PROF((X ? 11 : 13) + p * 4);
                    if (i == 0 && j > 0) mbar_wait(X ? dv_free : dk_free, (j - 1) & 1);  // dV_{j-1} / dK_{j-1} were copied out
                    tc_fence_after();
                    const uint32_t b_lo = mn0 + i * (TILE_BYTES >> 4);
                    if (X) {
                        // c(p): dV_j += P^T dO_i
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            if (s < ks_q) umma_ss_lohi(tmem_base + BK_T_DV, tr0 + s * 128, b_lo + s * 128, hi, idesc_tn, (i | s) != 0);
                        umma_commit(p_free);
                    } else {
                        // d(p): dQ_i += dS K_j ; dK_j += dS^T Q_i
                        const uint32_t k_lo = kmn0 + (j & 1) * (2 * BK_KV_TILE >> 4);
#pragma unroll
                        for (int s = 0; s < 6; ++s)
                            if (s < ks_k)
                                umma_ss_lohi(tmem_base + BK_T_DQ + i * 64,
                                             (s < 4 ? dsk0 + pp * (TILE_BYTES >> 4) + s * 2
                                                    : dsk0 + (2 * TILE_BYTES >> 4) + pp * 4 + (s - 4) * 2),
                                             k_lo + s * 128, hi, idesc_dq, (j | s) != 0);
                        // dS^T as MN-major A: keys 0..63 from tile pp, keys 64..127 from the shared tile (LBO spans the gap)
                        const uint32_t dst_lo = umma_desc_lo(smem_u32(sdS) + pp * TILE_BYTES, (2 - pp) * TILE_BYTES + pp * 64);
#pragma unroll
                        for (int s = 0; s < 8; ++s)
                            if (s < ks_q) umma_ss_lohi(tmem_base + BK_T_DK, dst_lo + s * 128, b_lo + s * 128, hi, idesc_tn, (i | s) != 0);
                        umma_commit(&ds_free[pp]);
                    }
                    if (i == nqb - 1) umma_commit(X ? dv_full : dk_full);
                }
            }
            if (!X) umma_commit(dq_full);
            if (X) PROF(2);
        }
      }
    } else {
        // ================================ compute warps ================================
        // 16 warps (four per SM sub-partition): warp (q, slab) owns TMEM lane quadrant q (one query row per thread) and
        // the 24 key columns [24 slab, 24 slab + 24) of every 96-key block -- for the probabilities AND for dS, so P(p)
        // stays in registers (packed bf16) from the step that computes it to the step that needs it.
        const int q = warp & 3;
        const int slab = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int c0 = slab * BK_SLAB;
        const float c = args.scale_log2e;
        const int sw = row & 7;
        const bool prof_thread = threadIdx.x == 128;
        // ---- prologue: delta * scale = rowsum(dO * O) * scale of query block `slab` -> smem, then everybody reads its rows.
        // O_i arrives by TMA in the (still unused) dS tiles next to dO_i: per-thread global loads of these rows took
        // 11,000 clocks here, the TMA tiles land in about 1,500.
        float* sdl = reinterpret_cast<float*>(smem + BK_DELTA);
        if (prof_thread) PROF(90);
        if (slab < nqb) {
            mbar_wait(&q_full[slab], 0);
            const uint8_t* orow = sdS + slab * TILE_BYTES + row * 128;
            const uint8_t* drow = sdO + slab * TILE_BYTES + row * 128;
            float acc = 0.0f;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                acc += dot8_bf16(*reinterpret_cast<const uint4*>(orow + ((g ^ sw) << 4)), *reinterpret_cast<const uint4*>(drow + ((g ^ sw) << 4)));
            sdl[slab * 128 + row] = acc * args.scale;  // rows past T were zero-filled by the TMA
        }
        // Query rows past T: lse = +inf makes their probabilities exact zeros (S = 0 there: the TMA zero-fills Q), and with
        // P = 0, dP = 0 (dO zero-filled) and delta = 0 their dS is an exact zero too -- no masking needed for rows.
        float lse2[3], sdelta[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int t = i * 128 + row;
            lse2[i] = (i < nqb && t < T) ? args.lse[(static_cast<size_t>(b) * H + h) * T + t] * 1.4426950408889634f : __int_as_float(0x7f800000);
        }
        if (prof_thread) PROF(91);
        named_bar_sync(1, BK_COMPUTE_THREADS);
        if (prof_thread) PROF(92);
#pragma unroll
        for (int i = 0; i < 3; ++i) sdelta[i] = sdl[i * 128 + row];

        // Key block jb is complete: dK | dV (TMEM columns [192, 320), lane = key, eight 16-column parts, two per slab: slabs
        // 0, 1 take dK, slabs 2, 3 take dV) -> bf16 -> the block's own K / V stage (every MMA that read it has retired) ->
        // two TMA stores; then the stage is refilled with block jb+2.  Per-thread global stores from here crawl (32 rows
        // per request, and the LSU starves while the tensor core streams operands from shared memory).
        auto store_kv = [&](int jb) {
            const bool key_warp = q * 32 < min(BK_KEYS, T - jb * BK_KEYS);
            uint8_t* stage = sKV + (jb & 1) * 2 * BK_KV_TILE;
            if (slab < 2) mbar_wait(dk_full, jb & 1);
            else mbar_wait(dv_full, jb & 1);
            if (prof_thread && jb == 0) PROF(70);
            tc_fence_after();
            if (key_warp) {
#pragma unroll 1
                for (int a = slab * 2; a < slab * 2 + 2; ++a) {
                    uint32_t o[16];
                    tmem_ld_32x16(t_row + BK_T_DK + a * 16, o);
                    tmem_ld_wait();
                    uint8_t* trow = stage + (a >> 2) * BK_KV_TILE + row * 128;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint4 v;
                        v.x = pack_bf16(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
                        v.y = pack_bf16(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
                        v.z = pack_bf16(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
                        v.w = pack_bf16(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
                        *reinterpret_cast<uint4*>(trow + ((((a & 3) * 2 + g) ^ sw) << 4)) = v;
                    }
                }
            }
            tc_fence_before();
            if (slab < 2) mbar_arrive_warp(dk_free);
            else mbar_arrive_warp(dv_free);
            if (prof_thread && jb == 0) PROF(71);
            fence_proxy_async_smem();
            named_bar_sync(2, BK_COMPUTE_THREADS);
            if (prof_thread && jb == 0) PROF(72);
            if (warp == 4 && lane == 0) {
                tma_store_3d(&args.tmDKV, stage, inner + h * 64, jb * BK_KEYS, b);
                tma_store_3d(&args.tmDKV, stage + BK_KV_TILE, 2 * inner + h * 64, jb * BK_KEYS, b);
                tma_store_commit();
            }
        };
        // one step later the stores have long read the stage: refill it with block jb+2 (waiting right away would stall
        // this warp, and with it every hand-off of the step, for the ~1,500 clocks the TMA needs to drain 24 KB)
        auto reload_kv = [&](int jb) {
            if (warp == 4 && lane == 0 && jb + 2 < nkb) {
                tma_store_wait_read<0>();
                load_kv(jb + 2);
            }
        };

        // Step p = pair (j, i) first turns P(p) (in registers since the previous step) and dP(p) into dS(p) and hands it
        // to the tensor core (dQ_i, dK_j), then computes P(p+1) from S(p+1) (dV_j needs it, and so does the next step).
        // S(p+1) and dP(p) were issued one step earlier, as soon as their TMEM buffers had been read out, so neither wait
        // sees the tensor pipe's latency.  The two halves of a step are NOT interleaved by hand: with four warps per
        // sub-partition the FMA-heavy dS half of one warp overlaps the MUFU-heavy P half of another, and keeping only one
        // 24-column operand live at a time leaves the registers that let four warps fit.
        // Both halves compute all 24 columns unconditionally; slabs that contain keys past the block's end (the last key
        // block only) then clear those entries with integer masks, which also kills whatever stale TMEM contents (columns
        // beyond the MMA's N extent) may have produced.
        uint32_t pk[BK_SLAB / 2];  // P(p), then dS(p), of this thread's 24 columns, packed bf16 (exact zeros where masked)
        auto clear_unreal = [&](int jj) {
            const int nv = min(BK_KEYS, T - jj * BK_KEYS) - c0;  // real keys among this slab's columns
            if (nv >= BK_SLAB) return;                            // warp-uniform
#pragma unroll
            for (int k = 0; k < BK_SLAB / 2; ++k) pk[k] &= (2 * k + 1 < nv ? 0xffffffffu : (2 * k < nv ? 0x0000ffffu : 0u));
        };
        auto load_slab = [&](uint32_t col, uint32_t (&v)[BK_SLAB]) {
#pragma unroll
            for (int g = 0; g < BK_SLAB / 8; ++g)   // 8-column loads: every address is aligned to the width of its load
                tmem_ld_32x8(t_row + col + c0 + g * 8, *reinterpret_cast<uint32_t(*)[8]>(&v[g * 8]));
            tmem_ld_wait();
        };
        auto p_math = [&](const uint32_t (&sv)[BK_SLAB], int ii) {
            const float l2 = ii == 0 ? lse2[0] : (ii == 1 ? lse2[1] : lse2[2]);
#pragma unroll
            for (int e = 0; e < BK_SLAB; e += 2)
                pk[e / 2] = pack_bf16(ex2_approx(fmaf(__uint_as_float(sv[e]), c, -l2)),
                                      ex2_approx(fmaf(__uint_as_float(sv[e + 1]), c, -l2)));
        };
        auto ds_math = [&](const uint32_t (&dv)[BK_SLAB], int ii) {
            const float sd = ii == 0 ? sdelta[0] : (ii == 1 ? sdelta[1] : sdelta[2]);
#pragma unroll
            for (int e = 0; e < BK_SLAB; e += 2) {
                const uint32_t pa = pk[e / 2];
                pk[e / 2] = pack_bf16(bf16_lo(pa) * fmaf(__uint_as_float(dv[e]), args.scale, -sd),
                                      bf16_hi(pa) * fmaf(__uint_as_float(dv[e + 1]), args.scale, -sd));
            }
        };
        // this slab's three 16-byte chunks of a [128 x 96] tile row: chunk gc = 3 slab + g of the 12 chunks of the row
        auto p_store = [&]() {
            uint8_t* prow = sP + row * 128;
#pragma unroll
            for (int g = 0; g < BK_SLAB / 8; ++g) {
                const int gc = slab * (BK_SLAB / 8) + g;
                *reinterpret_cast<uint4*>(prow + (gc >> 3) * TILE_BYTES + (((gc & 7) ^ sw) << 4)) =
                    make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive_warp(p_full);
        };
        auto ds_store = [&](int p) {
            const uint32_t pp = p & 1;
            if (p > 1) mbar_wait(&ds_free[pp], ((p >> 1) - 1) & 1);  // d(p-2) retired: this dS buffer is free
            uint8_t* dsrow0 = sdS + pp * TILE_BYTES + row * 128;  // keys 0..63
            uint8_t* dsrow1 = sdS + 2 * TILE_BYTES + row * 128;   // keys 64..95: chunks 4*pp .. 4*pp+3 of the shared tile
#pragma unroll
            for (int g = 0; g < BK_SLAB / 8; ++g) {
                const int gc = slab * (BK_SLAB / 8) + g;
                const uint4 o = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
                if (gc < 8) *reinterpret_cast<uint4*>(dsrow0 + ((gc ^ sw) << 4)) = o;
                else *reinterpret_cast<uint4*>(dsrow1 + ((((gc - 8) + 4 * pp) ^ sw) << 4)) = o;
            }
            fence_proxy_async_smem();
            mbar_arrive_warp(&ds_full[pp]);
        };
        // P(pair) from S in TMEM -> registers (and the S buffer back to the tensor core)
        // A (warp, pair) is DEAD when none of its 32 x 24 entries is real: the slab lies past the last key of the block
        // (T = 321: slabs 2 and 3 of the fourth key block) or the lane quadrant past the last query (quadrant 3 of the third
        // query block).  Its P and dS are exact zeros either way; a dead warp keeps the barrier protocol (and stores the
        // zeros) but skips the TMEM loads and the arithmetic, which leaves its sub-partition's issue slots to the live slabs.
        auto dead_pair = [&](int jj, int ii) {
            return (min(BK_KEYS, T - jj * BK_KEYS) <= c0 || ii * 128 + q * 32 >= T) && !(args.debug & 8);   // warp-uniform
        };
        auto zero_pk = [&]() {
#pragma unroll
            for (int k = 0; k < BK_SLAB / 2; ++k) pk[k] = 0u;
        };
        auto make_p = [&](uint32_t parity, int jj, int ii) {
            uint32_t sv[BK_SLAB];
            mbar_wait(s_full, parity);
            if (dead_pair(jj, ii)) {
                mbar_arrive_warp(s_free);
                zero_pk();
                return;
            }
            tc_fence_after();
            load_slab(BK_T_S, sv);
            tc_fence_before();
            mbar_arrive_warp(s_free);
            p_math(sv, ii);
            clear_unreal(jj);
        };
        // dS(pair) from P (registers) and dP in TMEM -> registers (and the dP buffer back to the tensor core)
        auto make_ds = [&](uint32_t parity, int jj, int ii) {
            uint32_t dv[BK_SLAB];
            mbar_wait(dp_full, parity);
            if (dead_pair(jj, ii)) {   // pk already holds the zeros of P(pair)
                mbar_arrive_warp(dp_free);
                return;
            }
            tc_fence_after();
            load_slab(BK_T_DP, dv);
            tc_fence_before();
            mbar_arrive_warp(dp_free);
            ds_math(dv, ii);
            clear_unreal(jj);
        };

        // ---- step -1: P(0)
        make_p(0, 0, 0);
        if (prof_thread) PROF(93);
        p_store();
        if (prof_thread) PROF(94);
        // ---- steps 0 .. total-2: dS(p), then P(p+1);  pair p = (j, i), pair p+1 = (jn, in)
        int j = 0, i = 0, jn = nqb > 1 ? 0 : 1, in = nqb > 1 ? 1 : 0;
#pragma unroll 1
        for (int p = 0; p + 1 < total; ++p) {
            if (prof_thread && p < 8) PROF(100 + p * 4);
            make_ds(p & 1, j, i);
            if (prof_thread && p < 8) PROF(101 + p * 4);
            // Key block j-1 is final: copy dK / dV out.  Here rather than at the top of the step: d(p-1) needs ~1,000 clocks
            // after the previous step's dS hand-off to retire, and the arithmetic above has just covered them.
            if (i == 0 && j > 0) {
                if (prof_thread) PROF(80 + j * 2);
                store_kv(j - 1);
                if (prof_thread) PROF(81 + j * 2);
            }
            ds_store(p);
            if (prof_thread && p < 8) PROF(102 + p * 4);
            make_p((p + 1) & 1, jn, in);
            mbar_wait(p_free, p & 1);  // c(p) retired: the P tile may be overwritten
            p_store();
            if (prof_thread && p < 8) PROF(103 + p * 4);
            if (i == 0 && j > 0) reload_kv(j - 1);
            j = jn, i = in;
            if (++in == nqb) in = 0, ++jn;
        }
        // ---- last step: dS(total-1)
        {
            const int p = total - 1;
            if (i == 0 && j > 0) store_kv(j - 1);
            make_ds(p & 1, j, i);
            ds_store(p);
        }
        if (prof_thread) PROF(95);
        store_kv(nkb - 1);
        if (prof_thread) PROF(96);
        // ---- epilogue: slab i converts dQ_i -> bf16 -> staging (sP tiles, then dS buffer 0) -> TMA store ----
        mbar_wait(dq_full, 0);
        if (prof_thread) PROF(97);
        tc_fence_after();
        if (slab < nqb && slab * 128 + q * 32 < T) {  // rows past T are clipped by the store anyway
            const int i = slab;
            uint8_t* orow = (i < 2 ? sP + i * TILE_BYTES : sdS) + row * 128;
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                uint32_t o0[32];
                tmem_ld_32x32(t_row + BK_T_DQ + i * 64 + part * 32, o0);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t* src = &o0[g * 8];
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(src[0]), __uint_as_float(src[1]));
                    o.y = pack_bf16(__uint_as_float(src[2]), __uint_as_float(src[3]));
                    o.z = pack_bf16(__uint_as_float(src[4]), __uint_as_float(src[5]));
                    o.w = pack_bf16(__uint_as_float(src[6]), __uint_as_float(src[7]));
                    *reinterpret_cast<uint4*>(orow + (((part * 4 + g) ^ sw) << 4)) = o;
                }
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, BK_COMPUTE_THREADS);
        if (warp == 4 && lane == 0) {
            for (int i = 0; i < nqb; ++i)
                tma_store_3d(&args.tmDQ, i < 2 ? sP + i * TILE_BYTES : sdS, h * 64, i * 128, b);
            tma_store_commit();
            tma_store_wait_all<0>();
            PROF(3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
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
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem_bytes(ATT_MAX_T));
        if (e == cudaSuccess)  // two CTAs per SM need (nearly) the whole shared-memory carve-out
            e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_fwd) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    const int inner = d.H * 64;
    const int tk = (d.T + 15) & ~15;
    // K / V arrive in equal boxes of kv_box rows (<= 256, a multiple of 16) that tile the padded key count exactly
    int nbox = (tk + 255) / 256;
    while ((tk / 16) % nbox != 0) ++nbox;
    AttnFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.kv_box = tk / nbox;
    const uint64_t p_row = (uint64_t)3 * inner * 2;
    int rc = 0;
    rc |= make_tmap_3d(&a.tmQ, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_row, (uint64_t)d.T * p_row, 64, 128);
    rc |= make_tmap_3d(&a.tmKV, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_row, (uint64_t)d.T * p_row, 64, a.kv_box);
    if (rc) {
        set_error("attn_fwd: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    if (reinterpret_cast<uintptr_t>(d.out) & 15) {
        set_error("attn_fwd: the output must be 16-byte aligned");
        return -2;
    }
    a.out = reinterpret_cast<__nv_bfloat16*>(d.out);
    a.lse = d.lse;
    a.B = d.B;
    a.H = d.H;
    a.T = d.T;
    a.scale = d.scale;
    a.scale_log2e = d.scale * 1.4426950408889634f;
    launch_pdl(attn_fwd_kernel, dim3(d.B * d.H), dim3(FWD_THREADS), fwd_smem_bytes(tk), stream, a);
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
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attn_bwd) failed: %s", cudaGetErrorString(e));
            return -10;
        }
        configured = true;
    }
    const int inner = d.H * 64;
    AttnBwdArgs a;
    memset(&a, 0, sizeof(a));
    const uint64_t p_qkv = (uint64_t)3 * inner * 2, p_o = (uint64_t)inner * 2;
    int rc = 0;
    rc |= make_tmap_3d(&a.tmQ, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_qkv, (uint64_t)d.T * p_qkv, 64, 128);
    rc |= make_tmap_3d(&a.tmDO, d.dout, TmapDtype::BF16, inner, d.T, d.B, p_o, (uint64_t)d.T * p_o, 64, 128);
    rc |= make_tmap_3d(&a.tmKV, d.qkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_qkv, (uint64_t)d.T * p_qkv, 64, BK_KEYS);
    rc |= make_tmap_3d(&a.tmDQ, d.dqkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_qkv, (uint64_t)d.T * p_qkv, 64, 128);
    rc |= make_tmap_3d(&a.tmO, d.out, TmapDtype::BF16, inner, d.T, d.B, p_o, (uint64_t)d.T * p_o, 64, 128);
    rc |= make_tmap_3d(&a.tmDKV, d.dqkv, TmapDtype::BF16, 3 * inner, d.T, d.B, p_qkv, (uint64_t)d.T * p_qkv, 64, BK_KEYS);
    if (rc) {
        set_error("attn_bwd: tensor map creation failed: %s", tmap_last_error());
        return -3;
    }
    a.lse = d.lse;
    a.B = d.B;
    a.H = d.H;
    a.T = d.T;
    a.scale = d.scale;
    a.scale_log2e = d.scale * 1.4426950408889634f;
    {
        static const char* dbg = getenv("SVIT_ATTN_DEBUG");
        a.debug = dbg ? atoi(dbg) : 0;
    }
    launch_pdl(attn_bwd_kernel, dim3(d.B * d.H), dim3(BK_THREADS), BK_SMEM, stream, a);
    count_launch();
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
