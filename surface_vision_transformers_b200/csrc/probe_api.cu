// temporary probe entry points (will be folded into the real C-ABI)
#include "gemm.cuh"
namespace svit { const char* last_error(); }
extern "C" {
const char* svit_last_error() { return svit::last_error(); }
int svit_gemm_tn(const void* A, const void* B, void* out, void* out2, const void* aux, const float* bias,
                 const float* rowtab, int rowtab_period, int M, int N, int K, int lda, int ldb, int ldo, int mode,
                 int out_f32, int num_sms, void* stream) {
    svit::GemmTnDesc d{A, B, out, out2, aux, bias, rowtab, rowtab_period, M, N, K, lda, ldb, ldo, mode, out_f32};
    return svit::launch_gemm_tn(d, num_sms, (cudaStream_t)stream);
}
int svit_gemm_wgrad(const void* dY, const void* X, float* dW, int M, int N, int K, int ldy, int ldx, int ldw,
                    int num_sms, void* stream) {
    svit::GemmWgradDesc d{dY, X, dW, M, N, K, ldy, ldx, ldw};
    return svit::launch_gemm_wgrad(d, num_sms, (cudaStream_t)stream);
}
}
#include "attention.cuh"
extern "C" {
int svit_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int T, float scale, void* stream) {
    svit::AttnDesc d{qkv, out, lse, B, H, T, scale};
    return svit::launch_attn_fwd(d, (cudaStream_t)stream);
}
int svit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B,
                  int H, int T, float scale, void* stream) {
    svit::AttnBwdDesc d{qkv, out, dout, lse, delta, dqkv, B, H, T, scale};
    return svit::launch_attn_bwd(d, (cudaStream_t)stream);
}
}
