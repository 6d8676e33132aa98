// elementwise.cuh -- HBM-bound kernels of the SiT hot path (everything that is not a GEMM / attention):
// patch gather + packing, LayerNorm fwd/bwd, head, MPP loss, weight-shadow preparation, AdamW.
// Each launcher returns 0 or a negative error code (message via svit_last_error()).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace svit {

// ---- a1: patch gather (tools/preprocessing.py:79-84): out[s,c,j,v] = mesh[s,c,table[v*N + j]] (bit-exact copy)
int launch_gather_patches(const float* mesh, const int32_t* table, float* out, int S, int C, int n_mesh, int N, int V,
                          cudaStream_t st);

// ---- a2/a8: pack (B,C,N,V) fp32 patches into the bf16 A operand [B*T, Kp] of the patch-embedding GEMM.
// Row b*T is the (zero) cls slot; row b*T+1+n holds patch n in (c,v) order, columns >= C*V are zero.
// MPP corruption (models/mpp.py:87-112) is folded in: rows with swap_sel take patch swap_src[b,n] of the same
// sample, rows with replace_sel take mask_token (given in the reference's (v c) order); replace wins over swap.
// If table != nullptr the input is a raw mesh (B,C,n_mesh) and the gather (a1) plus an optional per-channel
// z-score ((x-mean)/std, preprocessing.py:72) is fused in.
struct PackDesc {
    const float* x;          // (B,C,N,V) or (B,C,n_mesh) when table != nullptr
    void* A;                 // bf16 [B*T, Kp]
    int B, C, N, V, Kp;
    const uint8_t* swap_sel;     // (B,N) or nullptr
    const int64_t* swap_src;     // (B,N) or nullptr
    const uint8_t* replace_sel;  // (B,N) or nullptr
    const float* mask_token;     // (C*V) in (v c) order, or nullptr
    const int32_t* table;        // (V,N) gather table or nullptr
    int n_mesh;
    const float* ch_mean;        // (C) or nullptr
    const float* ch_std;         // (C) or nullptr
    int x_bf16;                  // 1: x points to bf16 data of the same layout (inputs staged as bf16 on the host)
};
int launch_pack_patches(const PackDesc& d, cudaStream_t st);

// ---- LayerNorm forward: a = LN(x)*gamma+beta (bf16), mean/rstd saved (fp32, eps inside rstd)
int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* a_bf16, float* mean, float* rstd, int M,
                  int D, float eps, cudaStream_t st);

// ---- LayerNorm backward fused with the residual-gradient add:
//   g_out = g_in + dLN(da)            (fp32, may alias g_in)   and a bf16 copy
//   dgamma += sum_rows da * xhat ; dbeta += sum_rows da ; colsum_out += sum_rows g_out   (fp32 atomics)
int launch_ln_bwd(const void* da_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum_out,
                  int M, int D, cudaStream_t st, int g_in_period = 0);
// g_in_period = T > 0: g_in is a compact [M / T, D] matrix that holds the incoming gradient of rows 0, T, 2T, ... only; every
// other row's incoming gradient is zero (the last block under cls pooling, engine.cu) -- no zero-filled [M, D] buffer.

// ---- a5: pooling + mlp_head (LayerNorm + Linear(D, C)) forward:  out[b,c]
int launch_head_fwd(const float* x, const float* gamma, const float* beta, const float* W, const float* bias, float* out,
                    int B, int T, int D, int C, int pool_mean, float eps, cudaStream_t st);
// backward: writes g (fp32 [B*T, D]) and its bf16 copy completely; accumulates head parameter grads and
// colsum(g) (bias gradient of the last residual Linear).
int launch_head_bwd(const float* x, const float* gamma, const float* beta, const float* W, const float* dout, float* g,
                    void* g_bf16, float* dgamma, float* dbeta, float* dW, float* dbias, float* colsum_out, int B, int T,
                    int D, int C, int pool_mean, float eps, cudaStream_t st);

// ---- column sums of a bf16 matrix (bias gradients): out[n] += sum_m Y[m,n]
int launch_colsum_bf16(const void* Y, float* out, int M, int N, int ld, cudaStream_t st);

// ---- fp32 -> bf16 copy
int launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st);

// ---- weight shadows: dst[r, c] = bf16(src[r, c]) (pitch ld_direct) and dstT[c, r] = bf16(src[r, c]) (pitch ld_t);
// batched over `count` matrices with constant element strides between consecutive matrices. Either dst may be null.
int launch_cast_transpose(const float* src, void* dst, void* dstT, int rows, int cols, int ld_direct, int ld_t, int count,
                          size_t src_stride, size_t dst_stride, size_t dstT_stride, cudaStream_t st);
// patch-embedding weight (D, C*V) in the reference's (v c) column order -> bf16 [D, Kp] in (c v) order, zero padded
int launch_prepare_patch_weight(const float* W, void* Wp, int D, int C, int V, int Kp, cudaStream_t st);
// additive table E[T, D]: E[0] = cls + pos[0]; E[t] = bias + pos[t]   (models/sit.py:70-73)
int launch_prepare_rowtab(const float* pos, const float* cls, const float* bias, float* E, int T, int D, cudaStream_t st);

// ---- patch-embedding backward reductions (pos / cls / bias) and weight-gradient un-permutation
int launch_embed_bwd(const float* g0, float* dpos, float* dcls, float* dbias, int B, int T, int D, cudaStream_t st);
int launch_unpermute_patch_wgrad(const float* dWp, float* dW, int D, int C, int V, int Kp, cudaStream_t st);

// ---- a9: MPP reconstruction loss on the decoder output y [B*T, ldy] (row b*T+1+n <-> patch n), target = the
// uncorrupted input rearranged 'b c n v -> b n (v c)'; loss_sum += sum over masked rows of (y - t)^2
int launch_mpp_loss_fwd(const float* y, int ldy, const float* x, const uint8_t* mask, float* loss_sum, int B, int C, int N,
                        int V, cudaStream_t st);
// dy[b*T+1+n, k] = coef * (y - t) on masked rows, 0 elsewhere (bf16, pitch lddy, all Kd columns written)
int launch_mpp_loss_bwd(const float* y, int ldy, const float* x, const uint8_t* mask, const float* coef_dev, void* dy,
                        int lddy, int B, int C, int N, int V, cudaStream_t st);
// mask_token gradient: r[d] = sum over replaced rows of g0[b*T+1+n, d]; dmt[k] += sum_d r[d] * W[d, k]
int launch_mask_token_grad(const float* g0, const uint8_t* replace_sel, const float* W, float* scratch_r, float* dmt,
                           int B, int T, int D, int K, cudaStream_t st);

// ---- a10: fused AdamW / Adam over flat fp32 buffers, by segment (segments without a gradient are skipped,
// exactly like torch.optim which ignores params whose .grad is None)
struct AdamSegment {
    long long offset;   // element offset in the flat buffers
    long long numel;
    float bias_corr1;   // 1 - beta1^step
    float bias_corr2;   // 1 - beta2^step
    int active;         // 0: skip (no gradient)
    int step;           // updates applied so far (device-side counter, see launch_adamw_advance)
};
// step += 1, bias_corr = 1 - beta^step for every active segment (one tiny stream-ordered kernel before launch_adamw)
int launch_adamw_advance(AdamSegment* segs_dev, int nsegs, float beta1, float beta2, cudaStream_t st);
int launch_adamw(float* p, const float* g, float* m, float* v, const AdamSegment* segs_dev, int nsegs,
                 const int* block_map_dev, int nblocks, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int decoupled, float grad_scale, cudaStream_t st);
int launch_sgd(float* p, const float* g, float* mom, long long n, float lr, float momentum, float dampening,
               float weight_decay, int nesterov, int first_step, float grad_scale, cudaStream_t st);

// loss = mean((out - target)^2) (l1 = 0) or mean(|out - target|) (l1 = 1) over n elements; dout[i] = d loss / d out[i]
int launch_regression_loss(const float* out, const float* target, int n, int l1, float* loss, float* dout, cudaStream_t st);

}  // namespace svit
