// gemm.cuh -- host-visible declarations of the tcgen05 GEMM kernels.
//
//   gemm_tn   : C[M,N] = A[M,K] * B[N,K]^T   (both operands K-major bf16, fp32 accumulate in TMEM)
//               used for every nn.Linear forward (models/sit.py:50,  vit_pytorch Attention.to_qkv /
//               to_out / FeedForward.net in the reference) and, with a transposed bf16 weight shadow,
//               for every input-gradient GEMM of loss.backward() (tools/train.py:290).
//   gemm_wgrad: dW[N,K] += dY[M,N]^T * X[M,K] (both operands MN-major, reduction over the token axis,
//               256x384 CTA-pair tiles split over the token axis, fp32 reductions into the flat gradient
//               buffer; optionally db[N] += column sums of dY from one extra MMA against a tile of ones).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace svit {

enum EpiMode : int {
    EPI_STORE = 0,  // out = acc (+ bias[col]) (+ rowtab[row % period][col])
    EPI_GELU = 1,   // out = acc + bias (pre-activation); out2 = gelu(out)      (both bf16)
    EPI_RESID = 2,  // out = acc + bias + aux                                    (aux/out same dtype)
    EPI_DGELU = 3,  // out = acc * gelu'(aux)                                    (aux/out bf16)
    EPI_GELU_ONLY = 4,  // out = gelu(acc + bias)   (inference: pre-activation not kept)
    EPI_GELU_GRAD = 5,  // out = gelu'(acc + bias); out2 = gelu(acc + bias)  (training: the backward pass only ever needs
                        //       gelu'(u), so the forward epilogue -- which has exp(-u^2/2) and Phi(u) at hand -- stores that
                        //       instead of u and the backward epilogue shrinks to one multiply)
    EPI_MUL = 6,        // out = acc * aux                                   (aux/out bf16)
};

struct GemmTnDesc {
    const void* A;      // bf16 [M, K], row pitch lda elements
    const void* B;      // bf16 [N, K], row pitch ldb elements
    void* out;          // OutT [M, N], row pitch ldo elements
    void* out2;         // EPI_GELU / EPI_GELU_GRAD: bf16 [M, N] (pitch ldo)
    const void* aux;    // EPI_RESID / EPI_DGELU / EPI_MUL: same dtype/pitch as out
    const float* bias;  // [N] or nullptr
    const float* rowtab;  // [period, N] fp32 or nullptr (EPI_STORE only)
    int rowtab_period;
    int M, N, K;
    int lda, ldb, ldo;
    int mode;
    int out_f32;  // 1: out (and aux) fp32, 0: bf16
};

// Returns 0 on success; on failure a negative code and a message retrievable by svit_last_error().
int launch_gemm_tn(const GemmTnDesc& d, int num_sms, cudaStream_t stream);

struct GemmWgradDesc {
    const void* dY;  // bf16 [M, N], pitch ldy
    const void* X;   // bf16 [M, K], pitch ldx
    float* dW;       // fp32 [N, K], pitch ldw; accumulated with atomics (caller zeroes)
    int M, N, K;
    int ldy, ldx, ldw;
    float* dbias = nullptr;  // fp32 [N] or nullptr: += sum over tokens of dY (the Linear's bias gradient)
};
int launch_gemm_wgrad(const GemmWgradDesc& d, int num_sms, cudaStream_t stream);

// Residual Linear + the following LayerNorm in one kernel (gemm_ln.cu):
//   x_out = A W^T + bias + x_in (fp32) ; a_out = LayerNorm(x_out) * gamma + beta (bf16) ; mean, rstd per row
struct GemmLnDesc {
    const void* A;       // bf16 [M, K], row pitch lda
    const void* W;       // bf16 [D, K], row pitch ldb
    const float* bias;   // [D] or nullptr
    const float* x_in;   // fp32 [M, D] residual
    float* x_out;        // fp32 [M, D]
    void* a_out;         // bf16 [M, D]
    const float* gamma;  // [D]
    const float* beta;   // [D]
    float* mean;         // [M]
    float* rstd;         // [M]
    int M, D, K, lda, ldb;
    float eps;
};
bool gemm_ln_supported(int D);
int launch_gemm_ln(const GemmLnDesc& d, int num_sms, cudaStream_t stream);

void set_error(const char* fmt, ...);

// Programmatic dependent launch (ptx.cuh: griddep_launch / griddep_wait): the per-layer kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that the set-up of kernel n + 1 (barriers, TMEM, descriptors) runs
// under the tail of kernel n.  SVIT_NO_PDL=1 turns the attribute off (A/B timing); the instructions are then no-ops.
bool pdl_enabled();
inline int pdl_attr(cudaLaunchAttribute* attr, int n) {  // appends the attribute, returns the new attribute count
    if (!pdl_enabled()) return n;
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    return n + 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(attr, 0);
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
void count_launch(int n = 1);  // kernels launched through this library (svit_launch_count)

}  // namespace svit
