// engine.cu -- C-ABI implementation: orchestrates the sm_100a kernels into SiT / MPP forward and backward.
// Host-side only bookkeeping lives here (offsets, workspace carving, launch order); see include/svit_b200.h.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "attention.cuh"
#include "check.cuh"
#include "dropout.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "svit_b200.h"

namespace svit {
const char* last_error();
unsigned long long launch_count();
int debug_read_attn_prof(long long* out, int n);
}
using namespace svit;

typedef __nv_bfloat16 bf16;

struct svit_engine {
    svit_config cfg;
    int D, depth, H, I, mlp, N, V, C, NC, T, K, Kp, Kd;
    int num_sms;
    std::vector<long long> poff, pnum;
    long long flat_numel;
    long long layer_stride;  // elements between the same tensor of consecutive layers
    // shadow offsets in bytes
    size_t sh_wpe, sh_rowtab, sh_qkv, sh_qkvT, sh_o, sh_oT, sh_w1, sh_w1T, sh_w2, sh_w2T, sh_total;
    size_t msh_wdec, msh_wdecT, msh_total;
    int check;  // fp32 check mode (svit_set_check_mode)
    const float* wdec_f32;  // fp32 master of the MPP decoder weight (recorded by svit_mpp_prepare_weights, check mode)
    // dropout state of the next forward / backward (svit_set_dropout); p = 0 disables a site
    float drop_p, drop_emb_p;
    unsigned long long drop_seed, drop_offset;
    int no_fuse_ln;  // SVIT_NO_FUSE_LN=1: keep the stand-alone LayerNorm kernels (A/B timing)
    int full_last;   // SVIT_FULL_LAST_LAYER=1: compute every row of the last block even under cls pooling (A/B timing)
    // Weight-gradient GEMMs on a second stream, next to the LayerNorm backward kernels (encoder_bwd).  SVIT_WGRAD_OVERLAP:
    // 0 = everything on the caller's stream, 1 = only the wgrad that directly precedes a LayerNorm backward, 2 (default) =
    // both wgrads of a sub-layer (needs the second bf16 gradient buffer Ws::g16b).
    int wgrad_overlap;
    int side_dev;
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join[2];
};

// The side stream and its events, created on first use for the device the engine is called on.
static bool ensure_side(svit_engine* e) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (e->side != nullptr && e->side_dev == dev) return true;
    if (e->side != nullptr) return false;  // created for another device: this call runs on the caller's stream only
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStream_t s = nullptr;
    // high priority: the weight-gradient CTAs claim their SM before the LayerNorm CTAs fill it (SVIT_SIDE_PRIO=0: lowest, A/B)
    const bool low = getenv("SVIT_SIDE_PRIO") != nullptr && atoi(getenv("SVIT_SIDE_PRIO")) == 0;
    if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, low ? lo : hi) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; ++i)
        if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < i; ++j) cudaEventDestroy(ev[j]);
            cudaStreamDestroy(s);
            return false;
        }
    e->side = s;
    e->side_dev = dev;
    e->ev_fork = ev[0];
    e->ev_join[0] = ev[1];
    e->ev_join[1] = ev[2];
    return true;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

enum LayerParam { LN1_W = 0, LN1_B, QKV_W, OUT_W, OUT_B, LN2_W, LN2_B, FC1_W, FC1_B, FC2_W, FC2_B };
enum { P_POS = 0, P_CLS = 1, P_PE_W = 2, P_PE_B = 3, P_LAYER0 = 4 };

static inline int pidx_layer(int l, int which) { return P_LAYER0 + 11 * l + which; }
static inline int pidx_head(const svit_engine* e, int which) { return P_LAYER0 + 11 * e->depth + which; }

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct LayerWs {
    float *xmid, *xout, *mean1, *rstd1, *mean2, *rstd2, *lse;
    bf16 *a1, *a2, *qkv, *O, *u, *h;
};
struct Ws {
    int B, M, training, mpp;
    bf16* Apatch;
    float* x0;
    std::vector<LayerWs> L;
    // backward temporaries
    float *g, *dWp, *rvec;
    bf16 *g16, *g16b, *du, *da, *dO, *dqkv;
    // mpp
    bf16 *xL16, *dy;
    size_t total;
};

struct Bump {
    uint8_t* base;
    size_t off;
    template <typename Tp>
    Tp* take(size_t n) {
        off = align_up(off, 256);
        Tp* p = reinterpret_cast<Tp*>(base + off);
        off += n * sizeof(Tp);
        return p;
    }
};

static void carve(const svit_engine* e, int B, int training, int mpp, int with_embed, void* base, Ws* w) {
    Bump bp{reinterpret_cast<uint8_t*>(base), 0};
    const size_t M = static_cast<size_t>(B) * e->T;
    w->B = B;
    w->M = static_cast<int>(M);
    w->training = training;
    w->mpp = mpp;
    w->Apatch = with_embed ? bp.take<bf16>(M * e->Kp) : nullptr;
    w->x0 = with_embed ? bp.take<float>(M * e->D) : nullptr;
    w->L.resize(e->depth);
    const size_t BHT = static_cast<size_t>(B) * e->H * e->T;
    if (training) {
        for (int l = 0; l < e->depth; ++l) {
            LayerWs& L = w->L[l];
            L.xmid = bp.take<float>(M * e->D);
            L.xout = bp.take<float>(M * e->D);
            L.mean1 = bp.take<float>(M);
            L.rstd1 = bp.take<float>(M);
            L.mean2 = bp.take<float>(M);
            L.rstd2 = bp.take<float>(M);
            L.lse = bp.take<float>(BHT);
            L.a1 = bp.take<bf16>(M * e->D);
            L.a2 = bp.take<bf16>(M * e->D);
            L.qkv = bp.take<bf16>(M * 3 * e->I);
            L.O = bp.take<bf16>(M * e->I);
            L.u = bp.take<bf16>(M * e->mlp);
            L.h = bp.take<bf16>(M * e->mlp);
        }
        w->g = bp.take<float>(M * e->D);
        w->g16 = bp.take<bf16>(M * e->D);
        w->g16b = bp.take<bf16>(M * e->D);  // LN2' writes its bf16 gradient here while the side stream still reads g16
        w->du = bp.take<bf16>(M * e->mlp);
        w->da = bp.take<bf16>(M * e->D);
        w->dO = bp.take<bf16>(M * e->I);
        w->dqkv = bp.take<bf16>(M * 3 * e->I);
        w->dWp = bp.take<float>(static_cast<size_t>(e->D) * e->Kp);
        w->rvec = bp.take<float>(e->D);
    } else {
        // inference: all layers share one set of buffers, the residual stream ping-pongs
        LayerWs S;
        S.xmid = bp.take<float>(M * e->D);
        float* xa = bp.take<float>(M * e->D);
        float* xb = bp.take<float>(M * e->D);
        S.mean1 = bp.take<float>(M);
        S.rstd1 = bp.take<float>(M);
        S.mean2 = S.mean1;
        S.rstd2 = S.rstd1;
        S.lse = bp.take<float>(BHT);
        S.a1 = bp.take<bf16>(M * e->D);
        S.a2 = S.a1;
        S.qkv = bp.take<bf16>(M * 3 * e->I);
        S.O = bp.take<bf16>(M * e->I);
        S.h = bp.take<bf16>(M * e->mlp);
        S.u = nullptr;
        for (int l = 0; l < e->depth; ++l) {
            w->L[l] = S;
            w->L[l].xout = (l & 1) ? xb : xa;
        }
        w->g = nullptr;
        w->g16 = w->g16b = w->du = w->da = w->dO = w->dqkv = nullptr;
        w->dWp = w->rvec = nullptr;
    }
    if (mpp) {
        w->xL16 = bp.take<bf16>(M * e->D);
        w->dy = training ? bp.take<bf16>(M * e->Kd) : nullptr;
    } else {
        w->xL16 = w->dy = nullptr;
    }
    w->total = align_up(bp.off, 256);
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
#define RET_IF(x)            \
    do {                     \
        int rc__ = (x);      \
        if (rc__) return rc__; \
    } while (0)

static int gemm(const svit_engine* e, cudaStream_t st, const void* A, int lda, const void* Bm, int ldb, void* out, int ldo,
                int M, int N, int K, int mode, int out_f32, const float* bias = nullptr, const void* aux = nullptr,
                void* out2 = nullptr, const float* rowtab = nullptr, int period = 1) {
    GemmTnDesc d{A, Bm, out, out2, aux, bias, rowtab, period, M, N, K, lda, ldb, ldo, mode, out_f32};
    return launch_gemm_tn(d, e->num_sms, st);
}
static int wgrad(const svit_engine* e, cudaStream_t st, const void* dY, int ldy, const void* X, int ldx, float* dW, int ldw,
                 int M, int N, int K, float* dbias = nullptr) {
    GemmWgradDesc d{dY, X, dW, M, N, K, ldy, ldx, ldw, dbias};
    return launch_gemm_wgrad(d, e->num_sms, st);
}

// x_out = A W^T + bias + x_in and, in the same kernel, the LayerNorm that follows (gemm_ln.cu)
static int gemm_ln(const svit_engine* e, cudaStream_t st, const void* A, int lda, const void* W, int ldb, const float* bias,
                   const float* x_in, float* x_out, const float* gamma, const float* beta, void* a_out, float* mean,
                   float* rstd, int M, int K) {
    GemmLnDesc d{A, W, bias, x_in, x_out, a_out, gamma, beta, mean, rstd, M, e->D, K, lda, ldb, 1e-5f};
    return launch_gemm_ln(d, e->num_sms, st);
}

static inline DropoutSite drop_layer(const svit_engine* e, int l, int which) {
    return DropoutSite{e->drop_seed, e->drop_offset, dropout_site_layer(l, which), e->drop_p};
}
static inline DropoutSite drop_emb(const svit_engine* e) {
    return DropoutSite{e->drop_seed, e->drop_offset, DROP_SITE_EMB, e->drop_emb_p};
}

static inline const bf16* shp(const void* shadow, size_t off) {
    return reinterpret_cast<const bf16*>(reinterpret_cast<const uint8_t*>(shadow) + off);
}

static int check_ws(const svit_engine* e, int B, int training, int mpp, int with_embed, void* ws_ptr, size_t ws_bytes,
                    Ws* w) {
    if (B <= 0) {
        set_error("batch must be >= 1 (got %d)", B);
        return -1;
    }
    if (ws_ptr == nullptr || (reinterpret_cast<uintptr_t>(ws_ptr) & 255)) {
        set_error("workspace must be a 256-byte aligned device pointer");
        return -1;
    }
    carve(e, B, training, mpp, with_embed, ws_ptr, w);
    if (ws_bytes != 0 && w->total > ws_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", w->total, ws_bytes);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// cls pooling: the last block on B rows
// ---------------------------------------------------------------------------------------------
// With pool = 'cls' the head reads token 0 of the encoder output only (models/sit.py:78), and after the last block's key /
// value projection everything is row-wise.  The other T - 1 query rows of the LAST block reach neither the prediction nor
// any gradient (their gradient is exactly zero in the reference as well), so svit_forward / svit_backward run that block's
// attention for the cls query alone (attention_cls.cu) and its to_out / LayerNorm / FeedForward on B rows.  The compact
// [B, *] operands live at the start of the block's own (otherwise unused) activation buffers.  Not with dropout (the masks
// are indexed by the position in the full tensors), not for the encoder-only / MPP entry points (they return every token).
static inline bool cls_last(const svit_engine* e) {
    return !e->cfg.pool_mean && !e->full_last && e->drop_p == 0.0f && e->T >= 4 && e->depth >= 1;
}
struct ClsWs {
    float *xin, *xmid, *xout, *g, *mean2, *rstd2, *prob;
    bf16 *a2, *g16, *O, *u, *h;
};
static inline ClsWs cls_views(const svit_engine* e, const LayerWs& L, int B) {
    const size_t BD = static_cast<size_t>(B) * e->D;
    return ClsWs{L.xmid, L.xmid + BD, L.xout, L.xmid + 2 * BD, L.mean2, L.rstd2, L.lse, L.a2, L.a2 + BD, L.O, L.u, L.h};
}
// rows b * T of a [B * T, D] fp32 matrix <-> a compact [B, D] matrix
static int copy_cls_rows(float* dst, size_t dst_pitch, const float* src, size_t src_pitch, int B, int D, cudaStream_t st) {
    if (cudaMemcpy2DAsync(dst, dst_pitch * sizeof(float), src, src_pitch * sizeof(float), D * sizeof(float), B,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("cls-row copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -12;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// encoder forward / backward over the workspace
// ---------------------------------------------------------------------------------------------
static int encoder_fwd(const svit_engine* e, const float* P, const void* sh, Ws& w, const float* x_in, cudaStream_t st,
                       const float** x_final, bool cls = false) {
    const int M = w.M, D = e->D, I = e->I, mlp = e->mlp;
    const float scale = 0.125f;  // dim_head ** -0.5 with dim_head = 64
    const bool drop = e->drop_p > 0.0f;  // extra passes (dropout.cuh); the p = 0 launch sequence is unchanged
    // Without dropout (every shipped config) and for the width the fused tile is built for, each residual Linear also
    // produces the LayerNorm of the NEXT sub-layer (gemm_ln.cu): to_out -> LN2 of the same block, fc2 -> LN1 of the
    // next block.  Only the very first LayerNorm of the encoder is a kernel of its own then.
    const bool fuse_ln = !drop && gemm_ln_supported(D) && !e->no_fuse_ln;
    const size_t nD = static_cast<size_t>(M) * D, nmlp = static_cast<size_t>(M) * mlp;
    const float* xin = x_in;
    for (int l = 0; l < e->depth; ++l) {
        LayerWs& L = w.L[l];
        const float* lp = P;
        auto pp = [&](int which) { return lp + e->poff[pidx_layer(l, which)]; };
        if (l == 0 || !fuse_ln) RET_IF(launch_ln_fwd(xin, pp(LN1_W), pp(LN1_B), L.a1, L.mean1, L.rstd1, M, D, 1e-5f, st));
        RET_IF(gemm(e, st, L.a1, D, shp(sh, e->sh_qkv) + static_cast<size_t>(l) * 3 * I * D, D, L.qkv, 3 * I, M, 3 * I, D,
                    EPI_STORE, 0));
        if (cls && l == e->depth - 1) {
            // ---- last block under cls pooling: token 0 only from here on (B rows, compact operands) ----
            const ClsWs c = cls_views(e, L, w.B);
            const int Mc = w.B;
            AttnClsDesc cd{L.qkv, c.O, c.prob, w.B, e->H, e->T, scale};
            RET_IF(launch_attn_cls_fwd(cd, st));
            RET_IF(copy_cls_rows(c.xin, D, xin, static_cast<size_t>(e->T) * D, w.B, D, st));
            if (fuse_ln) {
                RET_IF(gemm_ln(e, st, c.O, I, shp(sh, e->sh_o) + static_cast<size_t>(l) * D * I, I, pp(OUT_B), c.xin, c.xmid,
                               pp(LN2_W), pp(LN2_B), c.a2, c.mean2, c.rstd2, Mc, I));
            } else {
                RET_IF(gemm(e, st, c.O, I, shp(sh, e->sh_o) + static_cast<size_t>(l) * D * I, I, c.xmid, D, Mc, D, I, EPI_RESID,
                            1, pp(OUT_B), c.xin));
                RET_IF(launch_ln_fwd(c.xmid, pp(LN2_W), pp(LN2_B), c.a2, c.mean2, c.rstd2, Mc, D, 1e-5f, st));
            }
            if (w.training) {
                RET_IF(gemm(e, st, c.a2, D, shp(sh, e->sh_w1) + static_cast<size_t>(l) * mlp * D, D, c.u, mlp, Mc, mlp, D,
                            EPI_GELU_GRAD, 0, pp(FC1_B), nullptr, c.h));
            } else {
                RET_IF(gemm(e, st, c.a2, D, shp(sh, e->sh_w1) + static_cast<size_t>(l) * mlp * D, D, c.h, mlp, Mc, mlp, D,
                            EPI_GELU_ONLY, 0, pp(FC1_B)));
            }
            RET_IF(gemm(e, st, c.h, mlp, shp(sh, e->sh_w2) + static_cast<size_t>(l) * D * mlp, mlp, c.xout, D, Mc, D, mlp,
                        EPI_RESID, 1, pp(FC2_B), c.xmid));
            xin = c.xout;  // [B, D]: the head pools it with T = 1
            break;
        }
        AttnDesc ad{L.qkv, L.O, L.lse, w.B, e->H, e->T, scale};
        RET_IF(launch_attn_fwd(ad, st));
        if (fuse_ln) {
            RET_IF(gemm_ln(e, st, L.O, I, shp(sh, e->sh_o) + static_cast<size_t>(l) * D * I, I, pp(OUT_B), xin, L.xmid,
                           pp(LN2_W), pp(LN2_B), L.a2, L.mean2, L.rstd2, M, I));
        } else {
            RET_IF(gemm(e, st, L.O, I, shp(sh, e->sh_o) + static_cast<size_t>(l) * D * I, I, L.xmid, D, M, D, I, EPI_RESID, 1,
                        pp(OUT_B), xin));
            if (drop) RET_IF(launch_dropout_residual(L.xmid, xin, nD, drop_layer(e, l, DROP_SITE_TO_OUT), st));
            RET_IF(launch_ln_fwd(L.xmid, pp(LN2_W), pp(LN2_B), L.a2, L.mean2, L.rstd2, M, D, 1e-5f, st));
        }
        if (w.training) {
            // training keeps gelu'(u) (in L.u) instead of u: the backward epilogue is then a single multiply
            RET_IF(gemm(e, st, L.a2, D, shp(sh, e->sh_w1) + static_cast<size_t>(l) * mlp * D, D, L.u, mlp, M, mlp, D,
                        EPI_GELU_GRAD, 0, pp(FC1_B), nullptr, L.h));
        } else {
            RET_IF(gemm(e, st, L.a2, D, shp(sh, e->sh_w1) + static_cast<size_t>(l) * mlp * D, D, L.h, mlp, M, mlp, D,
                        EPI_GELU_ONLY, 0, pp(FC1_B)));
        }
        // the same mask on gelu(u) and on the stored gelu'(u): d/du [gelu(u) m / (1-p)] = gelu'(u) m / (1-p)
        if (drop) RET_IF(launch_dropout_scale(L.h, w.training ? L.u : nullptr, nmlp, 1, drop_layer(e, l, DROP_SITE_FF_ACT), st));
        if (fuse_ln && l + 1 < e->depth) {
            LayerWs& Ln = w.L[l + 1];
            auto pn = [&](int which) { return P + e->poff[pidx_layer(l + 1, which)]; };
            RET_IF(gemm_ln(e, st, L.h, mlp, shp(sh, e->sh_w2) + static_cast<size_t>(l) * D * mlp, mlp, pp(FC2_B), L.xmid, L.xout,
                           pn(LN1_W), pn(LN1_B), Ln.a1, Ln.mean1, Ln.rstd1, M, mlp));
        } else {
            RET_IF(gemm(e, st, L.h, mlp, shp(sh, e->sh_w2) + static_cast<size_t>(l) * D * mlp, mlp, L.xout, D, M, D, mlp,
                        EPI_RESID, 1, pp(FC2_B), L.xmid));
            if (drop) RET_IF(launch_dropout_residual(L.xout, L.xmid, nD, drop_layer(e, l, DROP_SITE_FF_OUT), st));
        }
        xin = L.xout;
    }
    *x_final = xin;
    return 0;
}

// On entry w.g / w.g16 hold dL/dx_final (fp32 + bf16) and grads[fc2_b of last layer] already holds colsum(g).
// On exit w.g / w.g16 hold dL/dx_in.
// With dropout the gradient that enters a dropped branch is g * m / (1-p) (a bf16 copy in w.da, which is free at
// both points); the bias gradients of to_out / fc2 are then column sums of that masked copy, so they come from the
// wgrad kernel's bias column instead of the LayerNorm-backward / head column sums (callers skip those too).
//
// Side stream (use_side, set by the C-ABI entry points after ensure_side): the weight-gradient GEMMs leave the critical path.
// LayerNorm backward is HBM-bound and leaves the tensor cores idle (76 us per launch, 10 % of the step), the weight-gradient
// GEMMs are tensor / L2-bound and touch little HBM, and one CTA of each fits an SM together (wgrad: 205 KB shared memory,
// 20 k registers, TMEM; ln_bwd: 5 KB, 32 k registers).  So the wgrads of a sub-layer are enqueued on a high-priority second
// stream at the moment the LayerNorm backward of that sub-layer is enqueued on the caller's stream:
//     main:  dfc2, dfc1 | LN2' (writes the bf16 gradient to g16b)  dto_out, attn', dqkv GEMM | LN1' (writes g16)
//     side:             | wgrad fc2 (reads g16), wgrad fc1 (du)    .........................  | wgrad to_out (g16b), wgrad qkv
// Buffers: a side kernel reads g16 / g16b / du / dqkv; the main stream waits for the side stream (join) before the kernel
// that overwrites the first of them -- LN1' for the FeedForward pair (g16), the next layer's LN2' for the attention pair
// (g16b; its attn' overwrites dqkv later) -- and before every progress callback (gradients of a stage must be final in
// the caller's stream order).  Mode 1 moves only the wgrad that directly precedes a LayerNorm backward and needs no second
// buffer.  Not with dropout (masked gradient copies live in w.da) and not for the B-row last block under cls pooling.
static int encoder_bwd(const svit_engine* e, const float* P, const void* sh, Ws& w, const float* x_in, float* G,
                       cudaStream_t st, svit_progress_fn progress = nullptr, void* user = nullptr, bool cls = false,
                       bool use_side = false) {
    const int M = w.M, D = e->D, I = e->I, mlp = e->mlp;
    const float scale = 0.125f;
    const bool drop = e->drop_p > 0.0f;
    const size_t nD = static_cast<size_t>(M) * D;
    const int ov = (use_side && !drop && e->side != nullptr) ? e->wgrad_overlap : 0;
    cudaStream_t ss = e->side;
    bool pending[2] = {false, false};
    auto fork = [&]() {   // the side stream continues from this point of the caller's stream
        cudaEventRecord(e->ev_fork, st);
        cudaStreamWaitEvent(ss, e->ev_fork, 0);
    };
    auto mark = [&](int k) {   // ... and this is where the caller's stream will pick it up again
        cudaEventRecord(e->ev_join[k], ss);
        pending[k] = true;
    };
    auto join = [&](int k) {
        if (pending[k]) {
            cudaStreamWaitEvent(st, e->ev_join[k], 0);
            pending[k] = false;
        }
    };
    struct Joiner {   // every exit path leaves the side stream joined (stream capture requires it, and so do the callers)
        decltype(join)& j;
        ~Joiner() { j(0), j(1); }
    } joiner{join};
    for (int l = e->depth - 1; l >= 0; --l) {
        LayerWs& L = w.L[l];
        const float* xin = (l == 0) ? x_in : w.L[l - 1].xout;
        auto pp = [&](int which) { return P + e->poff[pidx_layer(l, which)]; };
        auto gp = [&](int which) { return G + e->poff[pidx_layer(l, which)]; };
        if (cls && l == e->depth - 1) {
            // ---- last block under cls pooling: B rows; c.g / c.g16 hold dL/dx_final of token 0 (head backward) ----
            const ClsWs c = cls_views(e, L, w.B);
            const int Mc = w.B;
            // (side stream: the B-row weight-gradient GEMMs are one-tile kernels that would otherwise each hold the whole
            // stream for ~15 us; the full-size QKV one runs under the LayerNorm backward like in the other layers)
            bf16* c_g16mid = ov ? w.g16b : c.g16;   // compact [B, D]: LN2' must not overwrite what wgrad fc2 still reads
            RET_IF(gemm(e, st, c.g16, D, shp(sh, e->sh_w2T) + static_cast<size_t>(l) * mlp * D, D, w.du, mlp, Mc, mlp, D,
                        EPI_MUL, 0, nullptr, c.u));
            cudaStream_t sw = ov ? ss : st;
            if (ov) fork();
            RET_IF(wgrad(e, sw, c.g16, D, c.h, mlp, gp(FC2_W), mlp, Mc, D, mlp));
            RET_IF(wgrad(e, sw, w.du, mlp, c.a2, D, gp(FC1_W), D, Mc, mlp, D, gp(FC1_B)));
            if (ov) mark(1);   // (re-recorded after every group: the side stream is in order, the last record covers all)
            RET_IF(gemm(e, st, w.du, mlp, shp(sh, e->sh_w1T) + static_cast<size_t>(l) * D * mlp, mlp, w.da, D, Mc, D, mlp,
                        EPI_STORE, 0));
            RET_IF(launch_ln_bwd(w.da, c.xmid, c.mean2, c.rstd2, pp(LN2_W), c.g, c.g, c_g16mid, gp(LN2_W), gp(LN2_B), gp(OUT_B), Mc,
                                 D, st));
            RET_IF(gemm(e, st, c_g16mid, D, shp(sh, e->sh_oT) + static_cast<size_t>(l) * I * D, D, w.dO, I, Mc, I, D, EPI_STORE, 0));
            if (ov) fork();
            RET_IF(wgrad(e, sw, c_g16mid, D, c.O, I, gp(OUT_W), I, Mc, D, I));
            if (ov) mark(1);
            if (progress != nullptr) progress(SVIT_STAGE_WINDOW + l, user);
            AttnClsBwdDesc cb{L.qkv, c.prob, w.dO, w.dqkv, w.B, e->H, e->T, scale};
            RET_IF(launch_attn_cls_bwd(cb, st));
            RET_IF(gemm(e, st, w.dqkv, 3 * I, shp(sh, e->sh_qkvT) + static_cast<size_t>(l) * D * 3 * I, 3 * I, w.da, D, M, D,
                        3 * I, EPI_STORE, 0));
            if (ov) fork();
            RET_IF(wgrad(e, sw, w.dqkv, 3 * I, L.a1, D, gp(QKV_W), D, M, 3 * I, D));
            if (ov) mark(1);
            // the residual gradient entering LN1' is g_mid in the cls rows (compact c.g) and zero elsewhere
            float* cs0 = (l > 0) ? G + e->poff[pidx_layer(l - 1, FC2_B)] : nullptr;
            RET_IF(launch_ln_bwd(w.da, xin, L.mean1, L.rstd1, pp(LN1_W), c.g, w.g, w.g16, gp(LN1_W), gp(LN1_B), cs0, M, D, st,
                                 e->T));
            join(1);   // the layer below overwrites du / dqkv / g16b, and the stage is reported final next
            if (progress != nullptr) progress(l, user);
            continue;
        }
        // ---- FeedForward ----
        // du = (g W2) * gelu'(u)      (L.u holds gelu'(u), stored by the forward epilogue)
        const bf16* gb = w.g16;
        bf16* g16_mid = ov == 2 ? w.g16b : w.g16;   // bf16 gradient of the residual stream between the two sub-layers
        if (drop) {
            RET_IF(launch_dropout_grad(w.g, w.da, nD, 1, drop_layer(e, l, DROP_SITE_FF_OUT), st));
            gb = w.da;
        }
        RET_IF(gemm(e, st, gb, D, shp(sh, e->sh_w2T) + static_cast<size_t>(l) * mlp * D, D, w.du, mlp, M, mlp, D, EPI_MUL,
                    0, nullptr, L.u));
        if (ov != 2) RET_IF(wgrad(e, st, gb, D, L.h, mlp, gp(FC2_W), mlp, M, D, mlp, drop ? gp(FC2_B) : nullptr));
        // da2 = du W1
        RET_IF(gemm(e, st, w.du, mlp, shp(sh, e->sh_w1T) + static_cast<size_t>(l) * D * mlp, mlp, w.da, D, M, D, mlp,
                    EPI_STORE, 0));
        if (ov) {
            join(1);   // the attention pair of the layer above has read g16b (LN2' below overwrites it) and dqkv
            fork();
            if (ov == 2) RET_IF(wgrad(e, ss, gb, D, L.h, mlp, gp(FC2_W), mlp, M, D, mlp));
            RET_IF(wgrad(e, ss, w.du, mlp, L.a2, D, gp(FC1_W), D, M, mlp, D, gp(FC1_B)));
            mark(0);
        } else {
            RET_IF(wgrad(e, st, w.du, mlp, L.a2, D, gp(FC1_W), D, M, mlp, D, gp(FC1_B)));  // + d fc1_b = colsum(du)
        }
        // g_mid = g + LN2'(da2) ; colsum(g_mid) = d out_b
        RET_IF(launch_ln_bwd(w.da, L.xmid, L.mean2, L.rstd2, pp(LN2_W), w.g, w.g, g16_mid, gp(LN2_W), gp(LN2_B),
                             drop ? nullptr : gp(OUT_B), M, D, st));
        if (ov == 1) join(0);
        // ---- Attention ----
        gb = g16_mid;
        if (drop) {
            RET_IF(launch_dropout_grad(w.g, w.da, nD, 1, drop_layer(e, l, DROP_SITE_TO_OUT), st));
            gb = w.da;
        }
        RET_IF(gemm(e, st, gb, D, shp(sh, e->sh_oT) + static_cast<size_t>(l) * I * D, D, w.dO, I, M, I, D, EPI_STORE, 0));
        if (ov != 2) RET_IF(wgrad(e, st, gb, D, L.O, I, gp(OUT_W), I, M, D, I, drop ? gp(OUT_B) : nullptr));
        AttnBwdDesc bd{L.qkv, L.O, w.dO, L.lse, w.dqkv, w.B, e->H, e->T, scale};
        // communication window: the next kernel is long and hands SMs out CTA by CTA (svit_b200.h, SVIT_STAGE_WINDOW)
        if (progress != nullptr) progress(SVIT_STAGE_WINDOW + l, user);
        RET_IF(launch_attn_bwd(bd, st));
        RET_IF(gemm(e, st, w.dqkv, 3 * I, shp(sh, e->sh_qkvT) + static_cast<size_t>(l) * D * 3 * I, 3 * I, w.da, D, M, D, 3 * I,
                    EPI_STORE, 0));
        if (ov) {
            join(0);   // the FeedForward pair has read g16 (LN1' below overwrites it) and du
            fork();
            if (ov == 2) RET_IF(wgrad(e, ss, gb, D, L.O, I, gp(OUT_W), I, M, D, I));
            RET_IF(wgrad(e, ss, w.dqkv, 3 * I, L.a1, D, gp(QKV_W), D, M, 3 * I, D));
            mark(1);
        } else {
            RET_IF(wgrad(e, st, w.dqkv, 3 * I, L.a1, D, gp(QKV_W), D, M, 3 * I, D));
        }
        // g_in = g_mid + LN1'(da1) ; colsum(g_in) = d fc2_b of the previous layer
        float* cs = (l > 0 && !drop) ? G + e->poff[pidx_layer(l - 1, FC2_B)] : nullptr;
        RET_IF(launch_ln_bwd(w.da, xin, L.mean1, L.rstd1, pp(LN1_W), w.g, w.g, w.g16, gp(LN1_W), gp(LN1_B), cs, M, D, st));
        if (ov == 1 || progress != nullptr) join(1);
        // every gradient of layer l is final now (d fc2_b[l] was added by the layer above / the head)
        if (progress != nullptr) progress(l, user);
    }
    return 0;
}

static int embed_fwd(const svit_engine* e, const void* sh, Ws& w, const PackDesc& pd, cudaStream_t st) {
    RET_IF(launch_pack_patches(pd, st));
    RET_IF(gemm(e, st, w.Apatch, e->Kp, shp(sh, e->sh_wpe), e->Kp, w.x0, e->D, w.M, e->D, e->Kp, EPI_STORE, 1, nullptr,
                nullptr, nullptr, reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(sh) + e->sh_rowtab),
                e->T));
    // emb_dropout over all T rows, after the position add (models/sit.py:73-74)
    if (e->drop_emb_p > 0.0f) RET_IF(launch_dropout_scale(w.x0, nullptr, static_cast<size_t>(w.M) * e->D, 0, drop_emb(e), st));
    return 0;
}

// g / g16 hold dL/dx0
// use_side: the weight-gradient GEMM runs on the side stream under the HBM-bound position / cls / bias reduction (see encoder_bwd)
static int embed_bwd(const svit_engine* e, Ws& w, float* G, cudaStream_t st, bool use_side = false) {
    if (e->drop_emb_p > 0.0f) {
        const size_t n = static_cast<size_t>(w.M) * e->D;
        RET_IF(launch_dropout_scale(w.g, nullptr, n, 0, drop_emb(e), st));
        RET_IF(launch_cast_bf16(w.g, w.g16, n, st));
    }
    cudaMemsetAsync(w.dWp, 0, static_cast<size_t>(e->D) * e->Kp * sizeof(float), st);
    if (use_side && e->side != nullptr && e->wgrad_overlap != 0) {
        cudaEventRecord(e->ev_fork, st);
        cudaStreamWaitEvent(e->side, e->ev_fork, 0);
        int rc = wgrad(e, e->side, w.g16, e->D, w.Apatch, e->Kp, w.dWp, e->Kp, w.M, e->D, e->Kp);
        cudaEventRecord(e->ev_join[0], e->side);
        int rc2 = launch_embed_bwd(w.g, G + e->poff[P_POS], G + e->poff[P_CLS], G + e->poff[P_PE_B], w.B, e->T, e->D, st);
        cudaStreamWaitEvent(st, e->ev_join[0], 0);   // joined on every path
        RET_IF(rc);
        RET_IF(rc2);
    } else {
        RET_IF(launch_embed_bwd(w.g, G + e->poff[P_POS], G + e->poff[P_CLS], G + e->poff[P_PE_B], w.B, e->T, e->D, st));
        RET_IF(wgrad(e, st, w.g16, e->D, w.Apatch, e->Kp, w.dWp, e->Kp, w.M, e->D, e->Kp));
    }
    return launch_unpermute_patch_wgrad(w.dWp, G + e->poff[P_PE_W], e->D, e->C, e->V, e->Kp, st);
}

// ---------------------------------------------------------------------------------------------
// fp32 check mode: same orchestration, every operand fp32 (check.cu)
// ---------------------------------------------------------------------------------------------
struct CkLayer {
    float *xmid, *xout, *a1, *a2, *qkv, *O, *u, *h, *mean1, *rstd1, *mean2, *rstd2, *lse;
};
struct CkWs {
    int B, M;
    float *Ap, *x0;
    std::vector<CkLayer> L;
    float *g, *da, *dO, *dqkv, *du, *dy, *rvec;
    bf16* g16;
    size_t total;
};
static void ck_carve(const svit_engine* e, int B, int mpp, void* base, CkWs* w) {
    Bump bp{reinterpret_cast<uint8_t*>(base), 0};
    const size_t M = static_cast<size_t>(B) * e->T;
    const size_t D = e->D, I = e->I, mlp = e->mlp;
    w->B = B;
    w->M = static_cast<int>(M);
    w->Ap = bp.take<float>(M * e->K);
    w->x0 = bp.take<float>(M * D);
    w->L.resize(e->depth);
    const size_t BHT = static_cast<size_t>(B) * e->H * e->T;
    for (int l = 0; l < e->depth; ++l) {
        CkLayer& L = w->L[l];
        L.xmid = bp.take<float>(M * D);
        L.xout = bp.take<float>(M * D);
        L.a1 = bp.take<float>(M * D);
        L.a2 = bp.take<float>(M * D);
        L.qkv = bp.take<float>(M * 3 * I);
        L.O = bp.take<float>(M * I);
        L.u = bp.take<float>(M * mlp);
        L.h = bp.take<float>(M * mlp);
        L.mean1 = bp.take<float>(M);
        L.rstd1 = bp.take<float>(M);
        L.mean2 = bp.take<float>(M);
        L.rstd2 = bp.take<float>(M);
        L.lse = bp.take<float>(BHT);
    }
    w->g = bp.take<float>(M * D);
    w->da = bp.take<float>(M * D);
    w->dO = bp.take<float>(M * I);
    w->dqkv = bp.take<float>(M * 3 * I);
    w->du = bp.take<float>(M * mlp);
    w->g16 = bp.take<bf16>(M * D);
    w->dy = mpp ? bp.take<float>(M * e->K) : nullptr;
    w->rvec = mpp ? bp.take<float>(D) : nullptr;
    w->total = align_up(bp.off, 256);
}
static int ck_check_ws(const svit_engine* e, int B, int mpp, void* ws_ptr, size_t ws_bytes, CkWs* w) {
    if (B <= 0) {
        set_error("batch must be >= 1 (got %d)", B);
        return -1;
    }
    if (ws_ptr == nullptr || (reinterpret_cast<uintptr_t>(ws_ptr) & 255)) {
        set_error("workspace must be a 256-byte aligned device pointer");
        return -1;
    }
    ck_carve(e, B, mpp, ws_ptr, w);
    if (ws_bytes != 0 && w->total > ws_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", w->total, ws_bytes);
        return -1;
    }
    return 0;
}
// Y[M, N] = X[M, K] W[N, K]^T (+ bias) (+ resid) ; optional gelu copy
static int ck_linear(cudaStream_t st, const float* X, const float* W, float* Y, int M, int N, int K, const float* bias = nullptr,
                     const float* resid = nullptr, float* gelu_out = nullptr) {
    ck::Sgemm d{X, K, 1, W, 1, K, Y, N, M, N, K, bias, resid, gelu_out, 0};
    return ck::sgemm(d, st);
}
// dX[M, K] = dY[M, N] W[N, K]
static int ck_dgrad(cudaStream_t st, const float* dY, const float* W, float* dX, int M, int N, int K) {
    ck::Sgemm d{dY, N, 1, W, K, 1, dX, K, M, K, N, nullptr, nullptr, nullptr, 0};
    return ck::sgemm(d, st);
}
// dW[N, K] += dY[M, N]^T X[M, K]
static int ck_wgrad(cudaStream_t st, const float* dY, const float* X, float* dW, int M, int N, int K) {
    ck::Sgemm d{dY, 1, N, X, K, 1, dW, K, N, K, M, nullptr, nullptr, nullptr, 1};
    return ck::sgemm(d, st);
}

static int ck_encoder_fwd(const svit_engine* e, const float* P, CkWs& w, const float* x_in, cudaStream_t st,
                          const float** x_final) {
    const int M = w.M, D = e->D, I = e->I, mlp = e->mlp;
    const bool drop = e->drop_p > 0.0f;
    const size_t nD = static_cast<size_t>(M) * D, nmlp = static_cast<size_t>(M) * mlp;
    const float* xin = x_in;
    for (int l = 0; l < e->depth; ++l) {
        CkLayer& L = w.L[l];
        auto pp = [&](int which) { return P + e->poff[pidx_layer(l, which)]; };
        RET_IF(ck::ln_fwd(xin, pp(LN1_W), pp(LN1_B), L.a1, L.mean1, L.rstd1, M, D, 1e-5f, st));
        RET_IF(ck_linear(st, L.a1, pp(QKV_W), L.qkv, M, 3 * I, D));
        RET_IF(ck::attn_fwd(L.qkv, L.O, L.lse, w.B, e->H, e->T, 0.125f, st));
        RET_IF(ck_linear(st, L.O, pp(OUT_W), L.xmid, M, D, I, pp(OUT_B), xin));
        if (drop) RET_IF(launch_dropout_residual(L.xmid, xin, nD, drop_layer(e, l, DROP_SITE_TO_OUT), st));
        RET_IF(ck::ln_fwd(L.xmid, pp(LN2_W), pp(LN2_B), L.a2, L.mean2, L.rstd2, M, D, 1e-5f, st));
        RET_IF(ck_linear(st, L.a2, pp(FC1_W), L.u, M, mlp, D, pp(FC1_B), nullptr, L.h));
        if (drop) RET_IF(launch_dropout_scale(L.h, nullptr, nmlp, 0, drop_layer(e, l, DROP_SITE_FF_ACT), st));
        RET_IF(ck_linear(st, L.h, pp(FC2_W), L.xout, M, D, mlp, pp(FC2_B), L.xmid));
        if (drop) RET_IF(launch_dropout_residual(L.xout, L.xmid, nD, drop_layer(e, l, DROP_SITE_FF_OUT), st));
        xin = L.xout;
    }
    *x_final = xin;
    return 0;
}
// on entry w.g = dL/dx_final and grads[fc2_b of the last layer] already holds colsum(g); on exit w.g = dL/dx_in
static int ck_encoder_bwd(const svit_engine* e, const float* P, CkWs& w, const float* x_in, float* G, cudaStream_t st,
                          svit_progress_fn progress = nullptr, void* user = nullptr) {
    const int M = w.M, D = e->D, I = e->I, mlp = e->mlp;
    const bool drop = e->drop_p > 0.0f;
    const size_t nD = static_cast<size_t>(M) * D, nmlp = static_cast<size_t>(M) * mlp;
    for (int l = e->depth - 1; l >= 0; --l) {
        CkLayer& L = w.L[l];
        const float* xin = (l == 0) ? x_in : w.L[l - 1].xout;
        auto pp = [&](int which) { return P + e->poff[pidx_layer(l, which)]; };
        auto gp = [&](int which) { return G + e->poff[pidx_layer(l, which)]; };
        const float* gs = w.g;  // gradient entering the (possibly dropped) branch; w.da is free here
        if (drop) {
            RET_IF(launch_dropout_grad(w.g, w.da, nD, 0, drop_layer(e, l, DROP_SITE_FF_OUT), st));
            gs = w.da;
        }
        RET_IF(ck_dgrad(st, gs, pp(FC2_W), w.du, M, D, mlp));           // dh = g W2
        if (drop) RET_IF(launch_dropout_scale(w.du, nullptr, nmlp, 0, drop_layer(e, l, DROP_SITE_FF_ACT), st));
        RET_IF(ck::mul_dgelu(w.du, L.u, nmlp, st));  // du = dh * gelu'(u)
        RET_IF(ck_wgrad(st, gs, L.h, gp(FC2_W), M, D, mlp));
        if (drop) RET_IF(ck::colsum(gs, gp(FC2_B), M, D, st));
        RET_IF(ck::colsum(w.du, gp(FC1_B), M, mlp, st));
        RET_IF(ck_dgrad(st, w.du, pp(FC1_W), w.da, M, mlp, D));
        RET_IF(ck_wgrad(st, w.du, L.a2, gp(FC1_W), M, mlp, D));
        RET_IF(ck::ln_bwd(w.da, L.xmid, L.mean2, L.rstd2, pp(LN2_W), w.g, w.g, gp(LN2_W), gp(LN2_B), M, D, st));
        gs = w.g;
        if (drop) {
            RET_IF(launch_dropout_grad(w.g, w.da, nD, 0, drop_layer(e, l, DROP_SITE_TO_OUT), st));
            gs = w.da;
        }
        RET_IF(ck::colsum(gs, gp(OUT_B), M, D, st));
        RET_IF(ck_dgrad(st, gs, pp(OUT_W), w.dO, M, D, I));
        RET_IF(ck_wgrad(st, gs, L.O, gp(OUT_W), M, D, I));
        if (progress != nullptr) progress(SVIT_STAGE_WINDOW + l, user);  // same callback sequence as the tensor-core path
        RET_IF(ck::attn_bwd(L.qkv, L.O, w.dO, L.lse, w.dqkv, w.B, e->H, e->T, 0.125f, st));
        RET_IF(ck_dgrad(st, w.dqkv, pp(QKV_W), w.da, M, 3 * I, D));
        RET_IF(ck_wgrad(st, w.dqkv, L.a1, gp(QKV_W), M, 3 * I, D));
        RET_IF(ck::ln_bwd(w.da, xin, L.mean1, L.rstd1, pp(LN1_W), w.g, w.g, gp(LN1_W), gp(LN1_B), M, D, st));
        if (l > 0 && !drop) RET_IF(ck::colsum(w.g, G + e->poff[pidx_layer(l - 1, FC2_B)], M, D, st));
        if (progress != nullptr) progress(l, user);
    }
    return 0;
}
static int ck_forward(svit_engine* e, const float* P, void* ws_ptr, size_t ws_bytes, const PackDesc& pd, int B, float* out,
                      cudaStream_t st) {
    CkWs w;
    RET_IF(ck_check_ws(e, B, 0, ws_ptr, ws_bytes, &w));
    RET_IF(ck::patches(pd, w.Ap, st));
    RET_IF(ck_linear(st, w.Ap, P + e->poff[P_PE_W], w.x0, w.M, e->D, e->K));
    RET_IF(ck::embed_finish(w.x0, P + e->poff[P_POS], P + e->poff[P_CLS], P + e->poff[P_PE_B], B, e->T, e->D, st));
    if (e->drop_emb_p > 0.0f) RET_IF(launch_dropout_scale(w.x0, nullptr, static_cast<size_t>(w.M) * e->D, 0, drop_emb(e), st));
    const float* xf = nullptr;
    RET_IF(ck_encoder_fwd(e, P, w, w.x0, st, &xf));
    return launch_head_fwd(xf, P + e->poff[pidx_head(e, 0)], P + e->poff[pidx_head(e, 1)], P + e->poff[pidx_head(e, 2)],
                           P + e->poff[pidx_head(e, 3)], out, B, e->T, e->D, e->NC, e->cfg.pool_mean, 1e-5f, st);
}
static int ck_backward(svit_engine* e, const float* P, void* ws_ptr, int B, const float* dout, float* G,
                       svit_progress_fn progress, void* user, cudaStream_t st) {
    CkWs w;
    RET_IF(ck_check_ws(e, B, 0, ws_ptr, 0, &w));
    const float* xf = w.L[e->depth - 1].xout;
    RET_IF(launch_head_bwd(xf, P + e->poff[pidx_head(e, 0)], P + e->poff[pidx_head(e, 1)], P + e->poff[pidx_head(e, 2)], dout,
                           w.g, w.g16, G + e->poff[pidx_head(e, 0)], G + e->poff[pidx_head(e, 1)],
                           G + e->poff[pidx_head(e, 2)], G + e->poff[pidx_head(e, 3)],
                           e->drop_p > 0.0f ? nullptr : G + e->poff[pidx_layer(e->depth - 1, FC2_B)], B, e->T, e->D, e->NC,
                           e->cfg.pool_mean, 1e-5f, st));
    if (progress != nullptr) progress(e->depth, user);
    RET_IF(ck_encoder_bwd(e, P, w, w.x0, G, st, progress, user));
    if (e->drop_emb_p > 0.0f) RET_IF(launch_dropout_scale(w.g, nullptr, static_cast<size_t>(w.M) * e->D, 0, drop_emb(e), st));
    RET_IF(launch_embed_bwd(w.g, G + e->poff[P_POS], G + e->poff[P_CLS], G + e->poff[P_PE_B], B, e->T, e->D, st));
    RET_IF(ck_wgrad(st, w.g, w.Ap, G + e->poff[P_PE_W], w.M, e->D, e->K));
    if (progress != nullptr) progress(-1, user);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* svit_last_error(void) { return svit::last_error(); }
int svit_version(void) { return 100; }
unsigned long long svit_launch_count(void) { return svit::launch_count(); }
int svit_debug_attn_prof(long long* host_out, int n) { return svit::debug_read_attn_prof(host_out, n); }

svit_engine* svit_create(const svit_config* cfg) {
    if (cfg == nullptr) {
        set_error("svit_create: null config");
        return nullptr;
    }
    if (cfg->dim_head != 64) {
        set_error("svit_create: dim_head must be 64 (got %d)", cfg->dim_head);
        return nullptr;
    }
    if (cfg->dim <= 0 || cfg->dim % 8 != 0 || cfg->dim > 1024 || cfg->depth <= 0 || cfg->heads <= 0 || cfg->mlp_dim <= 0 ||
        cfg->mlp_dim % 8 != 0 || cfg->num_patches <= 0 || cfg->num_vertices <= 0 || cfg->num_channels <= 0 ||
        cfg->num_classes <= 0) {
        set_error("svit_create: unsupported configuration (dim %% 8, mlp_dim %% 8, dim <= 1024 required)");
        return nullptr;
    }
    if (cfg->num_patches + 1 > 384) {
        set_error("svit_create: sequence length %d > 384 unsupported by the fused attention kernel", cfg->num_patches + 1);
        return nullptr;
    }
    if ((cfg->num_channels * cfg->num_vertices) % 4 != 0) {
        set_error("svit_create: num_channels*num_vertices must be a multiple of 4");
        return nullptr;
    }
    svit_engine* e = new svit_engine();
    e->cfg = *cfg;
    e->D = cfg->dim;
    e->depth = cfg->depth;
    e->H = cfg->heads;
    e->I = cfg->heads * 64;
    e->mlp = cfg->mlp_dim;
    e->N = cfg->num_patches;
    e->V = cfg->num_vertices;
    e->C = cfg->num_channels;
    e->NC = cfg->num_classes;
    e->T = e->N + 1;
    e->K = e->C * e->V;
    e->Kp = static_cast<int>(align_up(e->K, 64));
    e->Kd = static_cast<int>(align_up(e->K, 8));
    e->num_sms = 148;
    e->check = 0;
    e->wdec_f32 = nullptr;
    e->drop_p = e->drop_emb_p = 0.0f;
    e->drop_seed = e->drop_offset = 0;
    e->no_fuse_ln = getenv("SVIT_NO_FUSE_LN") != nullptr && atoi(getenv("SVIT_NO_FUSE_LN")) != 0;
    e->full_last = getenv("SVIT_FULL_LAST_LAYER") != nullptr && atoi(getenv("SVIT_FULL_LAST_LAYER")) != 0;
    e->wgrad_overlap = getenv("SVIT_WGRAD_OVERLAP") != nullptr ? atoi(getenv("SVIT_WGRAD_OVERLAP")) : 2;
    if (e->wgrad_overlap < 0 || e->wgrad_overlap > 2) e->wgrad_overlap = 2;
    e->side = nullptr;
    e->side_dev = -1;
    e->ev_fork = e->ev_join[0] = e->ev_join[1] = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) e->num_sms = sms;
    } else {
        cudaGetLastError();
    }
    // parameter table
    const long long D = e->D, I = e->I, mlp = e->mlp;
    std::vector<long long> sizes;
    sizes.push_back(static_cast<long long>(e->T) * D);
    sizes.push_back(D);
    sizes.push_back(D * e->K);
    sizes.push_back(D);
    for (int l = 0; l < e->depth; ++l) {
        const long long ls[11] = {D, D, 3 * I * D, D * I, D, D, D, mlp * D, mlp, D * mlp, D};
        for (int i = 0; i < 11; ++i) sizes.push_back(ls[i]);
    }
    sizes.push_back(D);
    sizes.push_back(D);
    sizes.push_back(static_cast<long long>(e->NC) * D);
    sizes.push_back(e->NC);
    long long off = 0;
    for (size_t i = 0; i < sizes.size(); ++i) {
        e->poff.push_back(off);
        e->pnum.push_back(sizes[i]);
        off += static_cast<long long>(align_up(static_cast<size_t>(sizes[i]), 64));
    }
    e->flat_numel = off;
    e->layer_stride = e->depth > 1 ? e->poff[pidx_layer(1, 0)] - e->poff[pidx_layer(0, 0)] : 0;
    // shadow layout
    size_t so = 0;
    auto take = [&](size_t bytes) {
        so = align_up(so, 256);
        size_t r = so;
        so += bytes;
        return r;
    };
    e->sh_wpe = take(static_cast<size_t>(D) * e->Kp * 2);
    e->sh_rowtab = take(static_cast<size_t>(e->T) * D * 4);
    const size_t dl = e->depth;
    e->sh_qkv = take(dl * 3 * I * D * 2);
    e->sh_qkvT = take(dl * 3 * I * D * 2);
    e->sh_o = take(dl * D * I * 2);
    e->sh_oT = take(dl * D * I * 2);
    e->sh_w1 = take(dl * mlp * D * 2);
    e->sh_w1T = take(dl * mlp * D * 2);
    e->sh_w2 = take(dl * mlp * D * 2);
    e->sh_w2T = take(dl * mlp * D * 2);
    e->sh_total = align_up(so, 256);
    e->msh_wdec = 0;
    e->msh_wdecT = align_up(static_cast<size_t>(e->K) * D * 2, 256);
    e->msh_total = e->msh_wdecT + align_up(static_cast<size_t>(D) * e->Kd * 2, 256);
    return e;
}

void svit_destroy(svit_engine* e) {
    if (e == nullptr) return;
    if (e->side != nullptr) {   // best effort: the context may already be gone at interpreter exit
        cudaEventDestroy(e->ev_fork);
        cudaEventDestroy(e->ev_join[0]);
        cudaEventDestroy(e->ev_join[1]);
        cudaStreamDestroy(e->side);
        cudaGetLastError();
    }
    delete e;
}
int svit_num_params(const svit_engine* e) { return static_cast<int>(e->poff.size()); }
long long svit_param_offset(const svit_engine* e, int i) { return (i >= 0 && i < (int)e->poff.size()) ? e->poff[i] : -1; }
long long svit_param_numel(const svit_engine* e, int i) { return (i >= 0 && i < (int)e->pnum.size()) ? e->pnum[i] : -1; }
long long svit_flat_numel(const svit_engine* e) { return e->flat_numel; }
size_t svit_shadow_bytes(const svit_engine* e) { return e->sh_total; }
size_t svit_mpp_shadow_bytes(const svit_engine* e) { return e->msh_total; }
int svit_set_check_mode(svit_engine* e, int on) {
    if (e == nullptr) return -1;
    e->check = on ? 1 : 0;
    return 0;
}
int svit_get_check_mode(const svit_engine* e) { return e != nullptr ? e->check : 0; }
int svit_set_dropout(svit_engine* e, float p, float emb_p, unsigned long long seed, unsigned long long offset) {
    if (e == nullptr) return -1;
    if (!(p >= 0.0f) || p >= 1.0f || !(emb_p >= 0.0f) || emb_p >= 1.0f) {
        set_error("svit_set_dropout: probabilities must be in [0, 1) (got %f, %f)", p, emb_p);
        return -1;
    }
    e->drop_p = p;
    e->drop_emb_p = emb_p;
    e->drop_seed = seed;
    e->drop_offset = offset;
    return 0;
}
int svit_dropout_mask(uint8_t* keep, size_t n, float p, unsigned long long seed, unsigned long long offset, unsigned site,
                      void* stream) {
    return launch_dropout_mask(keep, n, DropoutSite{seed, offset, site, p}, reinterpret_cast<cudaStream_t>(stream));
}

size_t svit_workspace_bytes(const svit_engine* e, int batch, int training, int mpp) {
    if (batch <= 0) return 0;
    if (e->check) {
        CkWs cw;
        ck_carve(e, batch, mpp, nullptr, &cw);
        return cw.total;
    }
    Ws w;
    carve(e, batch, training, mpp, 1, nullptr, &w);
    return w.total;
}

int svit_prepare_weights(svit_engine* e, const float* P, void* shadow, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* sh = reinterpret_cast<uint8_t*>(shadow);
    const int D = e->D, I = e->I, mlp = e->mlp, dl = e->depth;
    const size_t ls = static_cast<size_t>(e->layer_stride);
    RET_IF(launch_prepare_patch_weight(P + e->poff[P_PE_W], sh + e->sh_wpe, D, e->C, e->V, e->Kp, st));
    RET_IF(launch_prepare_rowtab(P + e->poff[P_POS], P + e->poff[P_CLS], P + e->poff[P_PE_B],
                                 reinterpret_cast<float*>(sh + e->sh_rowtab), e->T, D, st));
    RET_IF(launch_cast_transpose(P + e->poff[pidx_layer(0, QKV_W)], sh + e->sh_qkv, sh + e->sh_qkvT, 3 * I, D, D, 3 * I, dl, ls,
                                 static_cast<size_t>(3) * I * D, static_cast<size_t>(3) * I * D, st));
    RET_IF(launch_cast_transpose(P + e->poff[pidx_layer(0, OUT_W)], sh + e->sh_o, sh + e->sh_oT, D, I, I, D, dl, ls,
                                 static_cast<size_t>(D) * I, static_cast<size_t>(D) * I, st));
    RET_IF(launch_cast_transpose(P + e->poff[pidx_layer(0, FC1_W)], sh + e->sh_w1, sh + e->sh_w1T, mlp, D, D, mlp, dl, ls,
                                 static_cast<size_t>(mlp) * D, static_cast<size_t>(mlp) * D, st));
    RET_IF(launch_cast_transpose(P + e->poff[pidx_layer(0, FC2_W)], sh + e->sh_w2, sh + e->sh_w2T, D, mlp, mlp, D, dl, ls,
                                 static_cast<size_t>(mlp) * D, static_cast<size_t>(mlp) * D, st));
    return 0;
}

int svit_mpp_prepare_weights(svit_engine* e, const float* Wdec, void* mpp_shadow, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* sh = reinterpret_cast<uint8_t*>(mpp_shadow);
    e->wdec_f32 = Wdec;  // the fp32 check mode reads the master weight itself
    // Wdec (K, D): direct [K, D] pitch D; transposed [D, K] pitch Kd
    return launch_cast_transpose(Wdec, sh + e->msh_wdec, sh + e->msh_wdecT, e->K, e->D, e->D, e->Kd, 1, 0, 0, 0, st);
}

int svit_forward(svit_engine* e, const float* P, const void* sh, void* ws_ptr, size_t ws_bytes, const float* input, int B,
                 const int32_t* table, int n_mesh, const float* ch_mean, const float* ch_std, float* out, int training,
                 void* stream) {
    return svit_forward_ex(e, P, sh, ws_ptr, ws_bytes, input, 0, B, table, n_mesh, ch_mean, ch_std, out, training, stream);
}

int svit_forward_ex(svit_engine* e, const float* P, const void* sh, void* ws_ptr, size_t ws_bytes, const void* input,
                    int input_bf16, int B, const int32_t* table, int n_mesh, const float* ch_mean, const float* ch_std,
                    float* out, int training, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float* in_f = reinterpret_cast<const float*>(input);
    if (e->check) {
        if (input_bf16) {
            set_error("fp32 check mode takes fp32 input");
            return -1;
        }
        PackDesc cpd{in_f, nullptr, B, e->C, e->N, e->V, e->K, nullptr, nullptr, nullptr, nullptr, table, n_mesh, ch_mean, ch_std};
        return ck_forward(e, P, ws_ptr, ws_bytes, cpd, B, out, st);
    }
    Ws w;
    RET_IF(check_ws(e, B, training, 0, 1, ws_ptr, ws_bytes, &w));
    PackDesc pd{in_f, w.Apatch, B, e->C, e->N, e->V, e->Kp, nullptr, nullptr, nullptr, nullptr, table, n_mesh, ch_mean, ch_std,
                input_bf16 ? 1 : 0};
    RET_IF(embed_fwd(e, sh, w, pd, st));
    const float* xf = nullptr;
    const bool cls = cls_last(e);  // then xf is the compact [B, D] encoder output of token 0
    RET_IF(encoder_fwd(e, P, sh, w, w.x0, st, &xf, cls));
    return launch_head_fwd(xf, P + e->poff[pidx_head(e, 0)], P + e->poff[pidx_head(e, 1)], P + e->poff[pidx_head(e, 2)],
                           P + e->poff[pidx_head(e, 3)], out, B, cls ? 1 : e->T, e->D, e->NC, e->cfg.pool_mean, 1e-5f, st);
}

int svit_backward(svit_engine* e, const float* P, const void* sh, void* ws_ptr, int B, const float* dout, float* G,
                  svit_progress_fn progress, void* user, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (e->check) return ck_backward(e, P, ws_ptr, B, dout, G, progress, user, st);
    Ws w;
    RET_IF(check_ws(e, B, 1, 0, 1, ws_ptr, 0, &w));
    const float* xf = w.L[e->depth - 1].xout;
    const bool cls = cls_last(e);  // xf, g and g16 are then the compact [B, D] token-0 tensors of the last block
    const ClsWs c = cls_views(e, w.L[e->depth - 1], B);
    RET_IF(launch_head_bwd(xf, P + e->poff[pidx_head(e, 0)], P + e->poff[pidx_head(e, 1)], P + e->poff[pidx_head(e, 2)], dout,
                           cls ? c.g : w.g, cls ? c.g16 : w.g16, G + e->poff[pidx_head(e, 0)], G + e->poff[pidx_head(e, 1)],
                           G + e->poff[pidx_head(e, 2)], G + e->poff[pidx_head(e, 3)],
                           e->drop_p > 0.0f ? nullptr : G + e->poff[pidx_layer(e->depth - 1, FC2_B)], B, cls ? 1 : e->T, e->D,
                           e->NC, e->cfg.pool_mean, 1e-5f, st));
    if (progress != nullptr) progress(e->depth, user);
    const bool side = e->wgrad_overlap != 0 && ensure_side(e);
    RET_IF(encoder_bwd(e, P, sh, w, w.x0, G, st, progress, user, cls, side));
    RET_IF(embed_bwd(e, w, G, st, side));
    if (progress != nullptr) progress(-1, user);
    return 0;
}

int svit_encoder_forward(svit_engine* e, const float* P, const void* sh, void* ws_ptr, size_t ws_bytes, const float* x, int B,
                         float* y, int training, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (e->check) {
        CkWs cw;
        RET_IF(ck_check_ws(e, B, 0, ws_ptr, ws_bytes, &cw));
        const float* cxf = nullptr;
        RET_IF(ck_encoder_fwd(e, P, cw, x, st, &cxf));
        if (cudaMemcpyAsync(y, cxf, static_cast<size_t>(cw.M) * e->D * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("encoder_forward (check mode): copy failed");
            return -12;
        }
        return 0;
    }
    Ws w;
    RET_IF(check_ws(e, B, training, 0, 1, ws_ptr, ws_bytes, &w));
    const float* xf = nullptr;
    RET_IF(encoder_fwd(e, P, sh, w, x, st, &xf));
    cudaError_t ce = cudaMemcpyAsync(y, xf, static_cast<size_t>(w.M) * e->D * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) {
        set_error("encoder_forward: copy failed: %s", cudaGetErrorString(ce));
        return -12;
    }
    return 0;
}

int svit_encoder_backward(svit_engine* e, const float* P, const void* sh, void* ws_ptr, int B, const float* x,
                            const float* dy, float* dx, float* G, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (e->check) {
        CkWs cw;
        RET_IF(ck_check_ws(e, B, 0, ws_ptr, 0, &cw));
        const size_t cn = static_cast<size_t>(cw.M) * e->D;
        cudaMemcpyAsync(cw.g, dy, cn * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (!(e->drop_p > 0.0f)) RET_IF(ck::colsum(cw.g, G + e->poff[pidx_layer(e->depth - 1, FC2_B)], cw.M, e->D, st));
        RET_IF(ck_encoder_bwd(e, P, cw, x, G, st));
        if (dx != nullptr) cudaMemcpyAsync(dx, cw.g, cn * sizeof(float), cudaMemcpyDeviceToDevice, st);
        return 0;
    }
    Ws w;
    RET_IF(check_ws(e, B, 1, 0, 1, ws_ptr, 0, &w));
    const size_t n = static_cast<size_t>(w.M) * e->D;
    cudaMemcpyAsync(w.g, dy, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    RET_IF(launch_cast_bf16(dy, w.g16, n, st));
    if (!(e->drop_p > 0.0f)) RET_IF(launch_colsum_bf16(w.g16, G + e->poff[pidx_layer(e->depth - 1, FC2_B)], w.M, e->D, e->D, st));
    RET_IF(encoder_bwd(e, P, sh, w, x, G, st, nullptr, nullptr, false, e->wgrad_overlap != 0 && ensure_side(e)));
    if (dx != nullptr) cudaMemcpyAsync(dx, w.g, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return 0;
}

int svit_mpp_forward(svit_engine* e, const float* P, const void* sh, const void* msh, const float* bdec,
                     const float* mask_token, void* ws_ptr, size_t ws_bytes, const float* input, int B, const uint8_t* mask,
                     const uint8_t* swap_sel, const int64_t* swap_src, const uint8_t* replace_sel, float* loss_sum,
                     float* batch_out, int training, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (e->check) {
        if (e->wdec_f32 == nullptr) {
            set_error("svit_mpp_forward (check mode): call svit_mpp_prepare_weights first");
            return -20;
        }
        CkWs cw;
        RET_IF(ck_check_ws(e, B, 1, ws_ptr, ws_bytes, &cw));
        PackDesc cpd{input, nullptr, B, e->C, e->N, e->V, e->K, swap_sel, swap_src, replace_sel, mask_token, nullptr, 0, nullptr, nullptr};
        RET_IF(ck::patches(cpd, cw.Ap, st));
        RET_IF(ck_linear(st, cw.Ap, P + e->poff[P_PE_W], cw.x0, cw.M, e->D, e->K));
        RET_IF(ck::embed_finish(cw.x0, P + e->poff[P_POS], P + e->poff[P_CLS], P + e->poff[P_PE_B], B, e->T, e->D, st));
        if (e->drop_emb_p > 0.0f)
            RET_IF(launch_dropout_scale(cw.x0, nullptr, static_cast<size_t>(cw.M) * e->D, 0, drop_emb(e), st));
        const float* cxf = nullptr;
        RET_IF(ck_encoder_fwd(e, P, cw, cw.x0, st, &cxf));
        RET_IF(ck_linear(st, cxf, e->wdec_f32, batch_out, cw.M, e->K, e->D, bdec));
        return launch_mpp_loss_fwd(batch_out, e->K, input, mask, loss_sum, B, e->C, e->N, e->V, st);
    }
    Ws w;
    RET_IF(check_ws(e, B, training, 1, 1, ws_ptr, ws_bytes, &w));
    PackDesc pd{input, w.Apatch, B, e->C, e->N, e->V, e->Kp, swap_sel, swap_src, replace_sel, mask_token, nullptr, 0, nullptr, nullptr};
    RET_IF(embed_fwd(e, sh, w, pd, st));
    const float* xf = nullptr;
    RET_IF(encoder_fwd(e, P, sh, w, w.x0, st, &xf));
    RET_IF(launch_cast_bf16(xf, w.xL16, static_cast<size_t>(w.M) * e->D, st));
    // decoder: batch_out[B*T, K] = xL Wdec^T + b
    RET_IF(gemm(e, st, w.xL16, e->D, shp(msh, e->msh_wdec), e->D, batch_out, e->K, w.M, e->K, e->D, EPI_STORE, 1, bdec));
    return launch_mpp_loss_fwd(batch_out, e->K, input, mask, loss_sum, B, e->C, e->N, e->V, st);
}

int svit_mpp_backward(svit_engine* e, const float* P, const void* sh, const void* msh, void* ws_ptr, int B,
                        const float* input, const float* batch_out, const uint8_t* mask, const uint8_t* replace_sel,
                        const float* coef, float* G, float* MG, svit_progress_fn progress, void* user, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (e->check) {
        CkWs cw;
        RET_IF(ck_check_ws(e, B, 1, ws_ptr, 0, &cw));
        const int cM = cw.M, cD = e->D, cK = e->K;
        float* cgW = MG;
        float* cgb = MG + static_cast<size_t>(cK) * cD;
        float* cgmt = cgb + cK;
        const float* cxf = cw.L[e->depth - 1].xout;
        RET_IF(ck::mpp_loss_bwd(batch_out, input, mask, coef, cw.dy, B, e->C, e->N, e->V, st));
        RET_IF(ck_wgrad(st, cw.dy, cxf, cgW, cM, cK, cD));
        RET_IF(ck::colsum(cw.dy, cgb, cM, cK, st));
        RET_IF(ck_dgrad(st, cw.dy, e->wdec_f32, cw.g, cM, cK, cD));
        if (!(e->drop_p > 0.0f)) RET_IF(ck::colsum(cw.g, G + e->poff[pidx_layer(e->depth - 1, FC2_B)], cM, cD, st));
        RET_IF(ck_encoder_bwd(e, P, cw, cw.x0, G, st, progress, user));
        if (e->drop_emb_p > 0.0f)
            RET_IF(launch_dropout_scale(cw.g, nullptr, static_cast<size_t>(cM) * cD, 0, drop_emb(e), st));
        RET_IF(launch_embed_bwd(cw.g, G + e->poff[P_POS], G + e->poff[P_CLS], G + e->poff[P_PE_B], B, e->T, cD, st));
        RET_IF(ck_wgrad(st, cw.g, cw.Ap, G + e->poff[P_PE_W], cM, cD, cK));
        if (progress != nullptr) progress(-1, user);
        if (replace_sel != nullptr)
            RET_IF(launch_mask_token_grad(cw.g, replace_sel, P + e->poff[P_PE_W], cw.rvec, cgmt, B, e->T, cD, cK, st));
        return 0;
    }
    Ws w;
    RET_IF(check_ws(e, B, 1, 1, 1, ws_ptr, 0, &w));
    const int M = w.M, D = e->D, K = e->K, Kd = e->Kd;
    float* gWdec = MG;
    float* gbdec = MG + static_cast<size_t>(K) * D;
    float* gmt = gbdec + K;
    RET_IF(launch_mpp_loss_bwd(batch_out, K, input, mask, coef, w.dy, Kd, B, e->C, e->N, e->V, st));
    // decoder grads
    RET_IF(wgrad(e, st, w.dy, Kd, w.xL16, D, gWdec, D, M, K, D, gbdec));  // + d to_original.bias = colsum(dy)
    // g = dy Wdec   (A = dy [M, K] pitch Kd, B = WdecT [D, K] pitch Kd)
    RET_IF(gemm(e, st, w.dy, Kd, shp(msh, e->msh_wdecT), Kd, w.g, D, M, D, K, EPI_STORE, 1));
    RET_IF(launch_cast_bf16(w.g, w.g16, static_cast<size_t>(M) * D, st));
    if (!(e->drop_p > 0.0f)) RET_IF(launch_colsum_bf16(w.g16, G + e->poff[pidx_layer(e->depth - 1, FC2_B)], M, D, D, st));
    const bool side = e->wgrad_overlap != 0 && ensure_side(e);
    RET_IF(encoder_bwd(e, P, sh, w, w.x0, G, st, progress, user, false, side));
    RET_IF(embed_bwd(e, w, G, st, side));
    if (progress != nullptr) progress(-1, user);
    if (replace_sel != nullptr)
        RET_IF(launch_mask_token_grad(w.g, replace_sel, P + e->poff[P_PE_W], w.rvec, gmt, B, e->T, D, K, st));
    return 0;
}

int svit_gather_patches(const float* mesh, const int32_t* table, float* out, int S, int C, int n_mesh, int N, int V,
                        void* stream) {
    return launch_gather_patches(mesh, table, out, S, C, n_mesh, N, V, reinterpret_cast<cudaStream_t>(stream));
}

int svit_adamw_step(float* p, const float* g, float* m, float* v, const svit_adam_segment* segs_dev, int nsegs,
                    const int* block_map_dev, int nblocks, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int decoupled, float grad_scale, void* stream) {
    static_assert(sizeof(svit_adam_segment) == sizeof(AdamSegment), "segment layout mismatch");
    return launch_adamw(p, g, m, v, reinterpret_cast<const AdamSegment*>(segs_dev), nsegs, block_map_dev, nblocks, lr, beta1,
                        beta2, eps, weight_decay, decoupled, grad_scale, reinterpret_cast<cudaStream_t>(stream));
}
int svit_adamw_advance(svit_adam_segment* segs_dev, int nsegs, float beta1, float beta2, void* stream) {
    return launch_adamw_advance(reinterpret_cast<AdamSegment*>(segs_dev), nsegs, beta1, beta2,
                                reinterpret_cast<cudaStream_t>(stream));
}
int svit_sgd_step(float* p, const float* g, float* mom, long long n, float lr, float momentum, float dampening,
                  float weight_decay, int nesterov, int first_step, float grad_scale, void* stream) {
    return launch_sgd(p, g, mom, n, lr, momentum, dampening, weight_decay, nesterov, first_step, grad_scale,
                      reinterpret_cast<cudaStream_t>(stream));
}

int svit_gemm_ln(const void* A, const void* W, const float* bias, const float* x_in, float* x_out, void* a_out,
                 const float* gamma, const float* beta, float* mean, float* rstd, int M, int D, int K, int lda, int ldb, float eps,
                 int num_sms, void* stream) {
    GemmLnDesc d{A, W, bias, x_in, x_out, a_out, gamma, beta, mean, rstd, M, D, K, lda, ldb, eps};
    return launch_gemm_ln(d, num_sms, reinterpret_cast<cudaStream_t>(stream));
}

int svit_regression_loss(const float* out, const float* target, int n, int l1, float* loss, float* dout, void* stream) {
    return launch_regression_loss(out, target, n, l1, loss, dout, reinterpret_cast<cudaStream_t>(stream));
}

int svit_gemm_tn(const void* A, const void* B, void* out, void* out2, const void* aux, const float* bias, const float* rowtab,
                 int rowtab_period, int M, int N, int K, int lda, int ldb, int ldo, int mode, int out_f32, int num_sms,
                 void* stream) {
    GemmTnDesc d{A, B, out, out2, aux, bias, rowtab, rowtab_period, M, N, K, lda, ldb, ldo, mode, out_f32};
    return launch_gemm_tn(d, num_sms, reinterpret_cast<cudaStream_t>(stream));
}
int svit_gemm_wgrad(const void* dY, const void* X, float* dW, int M, int N, int K, int ldy, int ldx, int ldw, int num_sms,
                    void* stream) {
    GemmWgradDesc d{dY, X, dW, M, N, K, ldy, ldx, ldw, nullptr};
    return launch_gemm_wgrad(d, num_sms, reinterpret_cast<cudaStream_t>(stream));
}
int svit_gemm_wgrad_bias(const void* dY, const void* X, float* dW, float* dbias, int M, int N, int K, int ldy, int ldx,
                         int ldw, int num_sms, void* stream) {
    GemmWgradDesc d{dY, X, dW, M, N, K, ldy, ldx, ldw, dbias};
    return launch_gemm_wgrad(d, num_sms, reinterpret_cast<cudaStream_t>(stream));
}
int svit_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int T, float scale, void* stream) {
    AttnDesc d{qkv, out, lse, B, H, T, scale};
    return launch_attn_fwd(d, reinterpret_cast<cudaStream_t>(stream));
}
int svit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int H, int T,
                  float scale, void* stream) {
    AttnBwdDesc d{qkv, out, dout, lse, dqkv, B, H, T, scale};
    return launch_attn_bwd(d, reinterpret_cast<cudaStream_t>(stream));
}
int svit_attn_cls_fwd(const void* qkv, void* out, float* prob, int B, int H, int T, float scale, void* stream) {
    AttnClsDesc d{qkv, out, prob, B, H, T, scale};
    return launch_attn_cls_fwd(d, reinterpret_cast<cudaStream_t>(stream));
}
int svit_attn_cls_bwd(const void* qkv, const float* prob, const void* dout, void* dqkv, int B, int H, int T, float scale,
                      void* stream) {
    AttnClsBwdDesc d{qkv, prob, dout, dqkv, B, H, T, scale};
    return launch_attn_cls_bwd(d, reinterpret_cast<cudaStream_t>(stream));
}
int svit_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* a_bf16, float* mean, float* rstd, int M,
                       int D, float eps, void* stream) {
    return launch_ln_fwd(x, gamma, beta, a_bf16, mean, rstd, M, D, eps, reinterpret_cast<cudaStream_t>(stream));
}
int svit_layernorm_bwd(const void* da_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum_out,
                       int M, int D, void* stream) {
    return launch_ln_bwd(da_bf16, x, mean, rstd, gamma, g_in, g_out, g_out_bf16, dgamma, dbeta, colsum_out, M, D,
                         reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
