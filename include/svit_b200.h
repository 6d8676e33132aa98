/* svit_b200.h -- C ABI of the B200-native SiT hot path (libsvit_b200.so).
 *
 * Plain pointers and sizes only; every pointer is a CUDA device pointer unless stated otherwise; every call is
 * asynchronous on the given cudaStream_t (passed as void*).  All functions returning int return 0 on success and
 * a negative value on error; svit_last_error() then holds a message (thread local).
 * There is NO CPU fallback: without a sm_100a device the launches fail and the error is reported.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   svit_forward / svit_backward          models/sit.py:66-82  SiT.forward  (+ autograd backward, tools/train.py:286-290)
 *   svit_encoder_forward / _backward      vit_pytorch.vit.Transformer.forward, constructed at models/sit.py:57, called :76
 *   svit_mpp_forward / svit_mpp_backward  models/mpp.py:77-134 masked_patch_pretraining.forward (+ backward, tools/pretrain.py:316-318)
 *   svit_gather_patches                   tools/preprocessing.py:79-84 (patch gather with utils/triangle_indices_*.csv)
 *   svit_adamw_step / svit_sgd_step       torch.optim.AdamW / Adam / SGD .step() as used at tools/train.py:228-241,291
 *   svit_gemm_tn / svit_gemm_wgrad / svit_attn_fwd / svit_attn_bwd / svit_layernorm_*   kernel-level entry points (tests, benchmarks)
 *
 * Parameter layout: one flat fp32 buffer, tensors in the canonical order below, each starting at a multiple of
 * 64 floats (svit_param_offset).  Shapes are the reference's state_dict shapes (models/sit.py:45-64,
 * utils/utils.py:13-33):
 *   0 pos_embedding (1,N+1,D)   1 cls_token (1,1,D)   2 to_patch_embedding.1.weight (D, C*V)   3 to_patch_embedding.1.bias (D)
 *   4+11*l+{0..10}: layers.l.0.norm.weight, .0.norm.bias, .0.fn.to_qkv.weight (3*H*64, D), .0.fn.to_out.0.weight (D, H*64),
 *                   .0.fn.to_out.0.bias, .1.norm.weight, .1.norm.bias, .1.fn.net.0.weight (mlp, D), .1.fn.net.0.bias,
 *                   .1.fn.net.3.weight (D, mlp), .1.fn.net.3.bias
 *   4+11*depth+{0..3}: mlp_head.0.weight, mlp_head.0.bias, mlp_head.1.weight (num_classes, D), mlp_head.1.bias
 * Gradients use the same layout in a second flat buffer.
 */
#ifndef SVIT_B200_H
#define SVIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svit_config {
    int dim;          /* D */
    int depth;
    int heads;        /* H */
    int dim_head;     /* must be 64 */
    int mlp_dim;
    int num_patches;  /* N */
    int num_vertices; /* V */
    int num_channels; /* C */
    int num_classes;
    int pool_mean;    /* 0: 'cls', 1: 'mean' (models/sit.py:78) */
} svit_config;

typedef struct svit_engine svit_engine;

const char* svit_last_error(void);
int svit_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
unsigned long long svit_launch_count(void);

/* ---- fp32 check mode (north star: "... tightening to 1e-4 in an fp32-accumulate check mode") ----
 * With check mode on, svit_forward / svit_backward / svit_encoder_forward / svit_encoder_backward / svit_mpp_forward /
 * svit_mpp_backward run the same operator sequence with every operand in fp32 on the CUDA cores (csrc/check.cu) instead
 * of bf16 on the tensor cores; parameter and gradient layouts are unchanged, svit_workspace_bytes() reports the (larger)
 * fp32 workspace, the bf16 weight shadows are not read (the MPP decoder weight is taken from the fp32 pointer last given
 * to svit_mpp_prepare_weights).  Measured: outputs and every gradient within 1e-4 relative L2 of the fp32 reference
 * restatement (tests/test_gpu_model.py).  Meant for verification, not for speed. */
int svit_set_check_mode(svit_engine* e, int on);
int svit_get_check_mode(const svit_engine* e);

/* ---- dropout > 0 ----
 * Replaces the reference's four nn.Dropout sites: emb_dropout after the position add (models/sit.py:55,74) and, in
 * every encoder block, the Dropout after to_out, after the GELU and after the second FeedForward Linear (the
 * vit-pytorch layout pinned by utils/utils.py:18-33).  svit_set_dropout() fixes the state used by the NEXT forward
 * and by the backward that belongs to it (call it again with the same values before that backward if another forward
 * ran in between); p = emb_p = 0 (the default, and what nn.Module.eval() means) disables the extra passes.
 * Keep decisions are a pure function of (seed, offset, site, element index) -- element i of a site's row-major tensor
 * is kept iff word (i & 3) of Philox4x32-10(counter = {i >> 2, site, offset_lo, offset_hi}, key = {seed_lo, seed_hi})
 * is >= floor(p * 2^32); kept values are scaled by 1 / (1 - p).  Sites: 4 * layer + {0: to_out, 1: after GELU,
 * 2: after the second Linear}; 0xFFFF0000: emb_dropout.  The host advances `offset` once per training step.
 * The stream differs from torch's nn.Dropout stream (as any two dropout implementations do); the distribution
 * (independent Bernoulli(1 - p) keeps, 1 / (1 - p) scaling) is the reference's.
 * svit_dropout_mask() writes the keep decisions as bytes (test hook; oracle/dropout.py restates the generator). */
int svit_set_dropout(svit_engine* e, float p, float emb_p, unsigned long long seed, unsigned long long offset);
int svit_dropout_mask(uint8_t* keep, size_t n, float p, unsigned long long seed, unsigned long long offset, unsigned site,
                      void* stream);

/* ---- engine life cycle (host-side object; owns no device memory) ---- */
svit_engine* svit_create(const svit_config* cfg);
void svit_destroy(svit_engine* e);
int svit_num_params(const svit_engine* e);                      /* 8 + 11*depth */
long long svit_param_offset(const svit_engine* e, int index);   /* element offset in the flat buffer */
long long svit_param_numel(const svit_engine* e, int index);
long long svit_flat_numel(const svit_engine* e);                /* total elements of the flat buffer */
size_t svit_shadow_bytes(const svit_engine* e);                 /* bf16 weight shadows + fp32 tables (persistent) */
size_t svit_mpp_shadow_bytes(const svit_engine* e);
size_t svit_workspace_bytes(const svit_engine* e, int batch, int training, int mpp);

/* ---- weight shadows: bf16 copies (direct and transposed) of the fp32 master weights; call after any update ---- */
int svit_prepare_weights(svit_engine* e, const float* params, void* shadow, void* stream);
int svit_mpp_prepare_weights(svit_engine* e, const float* to_original_w /* (C*V, D) */, void* mpp_shadow, void* stream);

/* ---- SiT.forward: input (B,C,N,V) fp32 -> out (B,num_classes) fp32.  training=1 keeps activations for backward.
 * If table != NULL the input is a raw mesh (B,C,n_mesh) and the patch gather (+ optional z-score) is fused. ---- */
int svit_forward(svit_engine* e, const float* params, const void* shadow, void* workspace, size_t workspace_bytes,
                 const float* input, int batch, const int32_t* table, int n_mesh, const float* ch_mean,
                 const float* ch_std, float* out, int training, void* stream);
/* The same with the input given as bf16 (input_bf16 = 1; same layout).  The patch-embedding GEMM consumes bf16 either way
 * (the packing kernel rounds fp32 input to bf16), so a batch rounded once on the host -- `inputs.to(device)` of
 * tools/train.py:281-283 with a pinned bf16 tensor -- gives bit-identical results at half the host-to-device bytes. */
int svit_forward_ex(svit_engine* e, const float* params, const void* shadow, void* workspace, size_t workspace_bytes,
                    const void* input, int input_bf16, int batch, const int32_t* table, int n_mesh, const float* ch_mean,
                    const float* ch_std, float* out, int training, void* stream);
/* dout (B,num_classes) -> grads (flat, ACCUMULATED: caller zeroes). Must follow svit_forward(training=1) on the
 * same workspace. */
/* progress (may be NULL) is called on the host while the work is being enqueued: stage = depth after the head
 * gradients are final, stage = l after those of encoder layer l, stage = -1 after patch-embedding / pos / cls.
 * A data-parallel caller launches the all-reduce of that flat-buffer range from it (overlap with backward).
 * stage = SVIT_STAGE_WINDOW + l announces a communication window: the kernel enqueued next (the attention backward of
 * layer l, ~270 us at the benchmark shape) is not a persistent kernel -- its CTAs are handed to the SMs one by one, so a
 * collective launched now shares the chip with it gracefully, whereas under a persistent GEMM (static tile schedule over
 * all SMs) it stalls the CTAs whose SMs it holds.  The gradients that are final at that point are those of the stages
 * reported so far (layers > l and the head).
 *
 * Streams: everything the caller may observe is ordered on `stream`.  Internally the backward entry points (svit_backward,
 * svit_encoder_backward, svit_mpp_backward) enqueue the weight-gradient GEMMs on one engine-owned high-priority side stream
 * that forks from and joins `stream` through events, so that they run next to the HBM-bound LayerNorm backward kernels
 * (engine.cu, encoder_bwd).  The side stream is joined before every progress callback (the gradients of a reported stage are
 * final in `stream` order) and before the entry point returns; it takes part in CUDA-graph capture of `stream` like any forked
 * stream.  SVIT_WGRAD_OVERLAP=0 keeps every kernel on `stream`. */
#define SVIT_STAGE_WINDOW 1000
typedef void (*svit_progress_fn)(int stage, void* user);
int svit_backward(svit_engine* e, const float* params, const void* shadow, void* workspace, int batch,
                  const float* dout, float* grads, svit_progress_fn progress, void* user, void* stream);

/* ---- encoder only: x (B,T,D) fp32 -> y (B,T,D) fp32 ; backward: dy -> dx (and accumulated grads) ---- */
int svit_encoder_forward(svit_engine* e, const float* params, const void* shadow, void* workspace,
                         size_t workspace_bytes, const float* x, int batch, float* y, int training, void* stream);
int svit_encoder_backward(svit_engine* e, const float* params, const void* shadow, void* workspace, int batch,
                          const float* x, const float* dy, float* dx, float* grads, void* stream);

/* ---- masked patch pretraining (models/mpp.py).  Masks are produced by the caller with the reference's own
 * torch RNG calls (mpp.py:25-43, 85-111) and passed in: mask/swap_sel/replace_sel uint8 (B,N), swap_src int64 (B,N).
 * loss_sum (device scalar, caller zeroes) receives sum over masked rows of (y - target)^2; batch_out is
 * (B, N+1, C*V) fp32 whose rows 1..N of each sample are the reference's batch_out (row 0 = decoded cls, unused). ---- */
int svit_mpp_forward(svit_engine* e, const float* params, const void* shadow, const void* mpp_shadow,
                     const float* to_original_b, const float* mask_token, void* workspace, size_t workspace_bytes,
                     const float* input, int batch, const uint8_t* mask, const uint8_t* swap_sel,
                     const int64_t* swap_src, const uint8_t* replace_sel, float* loss_sum, float* batch_out,
                     int training, void* stream);
/* coef (device scalar) = dL/dloss * 2 / (num_masked_rows * C*V).  mpp_grads: to_original.weight (C*V*D),
 * to_original.bias (C*V), mask_token (C*V) consecutively, accumulated. */
int svit_mpp_backward(svit_engine* e, const float* params, const void* shadow, const void* mpp_shadow, void* workspace,
                      int batch, const float* input, const float* batch_out, const uint8_t* mask,
                      const uint8_t* replace_sel, const float* coef, float* grads, float* mpp_grads,
                      svit_progress_fn progress, void* user, void* stream);

/* ---- a1: standalone bit-exact patch gather: out[s,c,j,v] = mesh[s,c,table[v*N+j]] ---- */
int svit_gather_patches(const float* mesh, const int32_t* table, float* out, int S, int C, int n_mesh, int N, int V,
                        void* stream);

/* ---- optimizers over flat buffers ---- */
typedef struct svit_adam_segment {
    long long offset;
    long long numel;
    float bias_corr1; /* 1 - beta1^step */
    float bias_corr2; /* 1 - beta2^step */
    int active;       /* 0 = parameter had no gradient: skipped like torch.optim does */
    int step;         /* updates applied so far (torch's state['step']); maintained on the device by svit_adamw_advance */
} svit_adam_segment;
#define SVIT_ADAM_BLOCK_ELEMS 4096
/* block_map: int pairs (segment index, chunk index), one per thread block, chunks of SVIT_ADAM_BLOCK_ELEMS elements */
int svit_adamw_step(float* p, const float* g, float* m, float* v, const svit_adam_segment* segs_dev, int nsegs,
                    const int* block_map_dev, int nblocks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int decoupled, float grad_scale, void* stream);
/* Stream-ordered bookkeeping for the next svit_adamw_step: step += 1 and bias_corr{1,2} = 1 - beta{1,2}^step for every
 * active segment of the DEVICE table.  The host therefore uploads the table only when offsets or active flags change:
 * nothing step-dependent crosses the host/device boundary asynchronously any more (a pinned staging buffer rewritten by
 * a host that runs several steps ahead of the GPU used to), and a captured CUDA graph of the training step replays
 * correctly. */
int svit_adamw_advance(svit_adam_segment* segs_dev, int nsegs, float beta1, float beta2, void* stream);
int svit_sgd_step(float* p, const float* g, float* momentum_buf, long long n, float lr, float momentum, float dampening,
                  float weight_decay, int nesterov, int first_step, float grad_scale, void* stream);

/* Regression criterion of the reference's training loop (tools/train.py:245-248 constructor, :288 use): nn.MSELoss(mean)
 * (l1 = 0) or nn.L1Loss() (l1 = 1) of out[n] (= outputs.squeeze(), fp32) against target[n].  One launch writes the scalar
 * *loss and dout[n] = d loss / d out, so the criterion's backward needs no kernel of its own. */
int svit_regression_loss(const float* out, const float* target, int n, int l1, float* loss, float* dout, void* stream);

/* ---- kernel-level entry points ---- */
int svit_gemm_tn(const void* A, const void* B, void* out, void* out2, const void* aux, const float* bias,
                 const float* rowtab, int rowtab_period, int M, int N, int K, int lda, int ldb, int ldo, int mode,
                 int out_f32, int num_sms, void* stream);
/* Residual Linear + the LayerNorm that follows it, one kernel (dim 384): x_out = A W^T + bias + x_in (fp32),
 * a_out = LayerNorm(x_out) * gamma + beta (bf16), mean / rstd per row.  Replaces `x = fn(x) + x` of one PreNorm block and
 * the `norm` of the next (vit_pytorch Transformer.forward; keys utils/utils.py:18-23,28-31). */
int svit_gemm_ln(const void* A, const void* W, const float* bias, const float* x_in, float* x_out, void* a_out,
                 const float* gamma, const float* beta, float* mean, float* rstd, int M, int D, int K, int lda, int ldb,
                 float eps, int num_sms, void* stream);
int svit_gemm_wgrad(const void* dY, const void* X, float* dW, int M, int N, int K, int ldy, int ldx, int ldw,
                    int num_sms, void* stream);
/* same, and dbias[N] += sum over the M rows of dY (bias gradient of the Linear, fused as one extra MMA) */
int svit_gemm_wgrad_bias(const void* dY, const void* X, float* dW, float* dbias, int M, int N, int K, int ldy, int ldx,
                         int ldw, int num_sms, void* stream);
int svit_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int T, float scale, void* stream);
/* (rowsum(dout * out) is computed inside the kernel and dQ accumulates on chip: no scratch arguments) */
int svit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int H, int T,
                  float scale, void* stream);
/* The same attention for the query row of token 0 only -- what the last block needs under `pool = 'cls'`
 * (models/sit.py:78: the head reads x[:, 0]); svit_forward / svit_backward use it there and run the block's row-wise
 * rest on B rows.  out [B, H*64] bf16, prob [B, H, T] fp32 (softmax row of token 0, kept for the backward);
 * dout [B, H*64] bf16 -> dqkv [B, T, 3*H*64] bf16 (dk, dv of every key, dq of token 0, zeros in the other dq rows). */
int svit_attn_cls_fwd(const void* qkv, void* out, float* prob, int B, int H, int T, float scale, void* stream);
int svit_attn_cls_bwd(const void* qkv, const float* prob, const void* dout, void* dqkv, int B, int H, int T, float scale,
                      void* stream);
int svit_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* a_bf16, float* mean, float* rstd,
                       int M, int D, float eps, void* stream);
int svit_layernorm_bwd(const void* da_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum_out,
                       int M, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVIT_B200_H */
