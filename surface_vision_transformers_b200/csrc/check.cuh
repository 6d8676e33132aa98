// check.cuh -- fp32 check-mode kernels (see check.cu).  Launchers return 0 or a negative error code.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

#include "elementwise.cuh"

namespace svit {
namespace ck {

// C[i, j] (+)= sum_k A(i, k) B(k, j) (+ bias[j]) (+ resid[i, j]);  gelu_out[i, j] = gelu(C[i, j]) if requested.
// A(i, k) = A[i * sa_i + k * sa_k], B(k, j) = B[k * sb_k + j * sb_j]; C, resid and gelu_out are row-major with pitch ldc.
struct Sgemm {
    const float* A;
    long long sa_i, sa_k;
    const float* B;
    long long sb_k, sb_j;
    float* C;
    long long ldc;
    int M, N, K;
    const float* bias;
    const float* resid;
    float* gelu_out;
    int accumulate;
};
int sgemm(const Sgemm& d, cudaStream_t st);
int mul_dgelu(float* y, const float* u, size_t n, cudaStream_t st);
int colsum(const float* Y, float* out, int M, int N, cudaStream_t st);
int ln_fwd(const float* x, const float* gamma, const float* beta, float* a, float* mean, float* rstd, int M, int D, float eps,
           cudaStream_t st);
int ln_bwd(const float* da, const float* x, const float* mean, const float* rstd, const float* gamma, const float* g_in,
           float* g_out, float* dgamma, float* dbeta, int M, int D, cudaStream_t st);
int attn_fwd(const float* qkv, float* out, float* lse, int B, int H, int T, float scale, cudaStream_t st);
int attn_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int B, int H, int T,
             float scale, cudaStream_t st);
int patches(const PackDesc& d, float* A, cudaStream_t st);
int mpp_loss_bwd(const float* y, const float* x, const uint8_t* mask, const float* coef_dev, float* dy, int B, int C, int N,
                 int V, cudaStream_t st);
int embed_finish(float* x0, const float* pos, const float* cls, const float* bias, int B, int T, int D, cudaStream_t st);

}  // namespace ck
}  // namespace svit
