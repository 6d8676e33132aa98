// dropout.cuh -- counter-based dropout for the reference's four nn.Dropout sites:
//   models/sit.py:55,74 (emb_dropout after the position add), and per encoder block the Dropout after to_out
//   (Attention.to_out = Sequential(Linear, Dropout)), after the GELU and after the second FeedForward Linear
//   (FeedForward.net = [Linear, GELU, Dropout, Linear, Dropout]; layout pinned by utils/utils.py:18-33).
// Keep decisions are a pure function of (seed, offset, site, element index): the backward pass regenerates them
// instead of storing masks. Element i of a site's tensor is kept iff word (i & 3) of
//   Philox4x32-10(counter = {i >> 2, site, offset_lo, offset_hi}, key = {seed_lo, seed_hi})
// is >= threshold = floor(p * 2^32); kept values are scaled by 1 / (1 - p) (torch.nn.functional.dropout semantics).
// These kernels run only when p > 0 -- no shipped reference configuration does (config/SiT/*/ *.yml: 0.0), so
// they are separate HBM-bound passes that leave the p = 0 hot path untouched.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace svit {

struct DropoutSite {
    unsigned long long seed, offset;
    uint32_t site;
    float p;
};
static inline uint32_t dropout_site_layer(int layer, int which) { return static_cast<uint32_t>(layer) * 4u + which; }
enum { DROP_SITE_TO_OUT = 0, DROP_SITE_FF_ACT = 1, DROP_SITE_FF_OUT = 2 };
static const uint32_t DROP_SITE_EMB = 0xFFFF0000u;

// a[i] = a[i] * m[i] / (1-p) (and the same for b when b != nullptr); T = float (is_bf16 = 0) or bf16 (is_bf16 = 1)
int launch_dropout_scale(void* a, void* b, size_t n, int is_bf16, const DropoutSite& s, cudaStream_t st);
// out[i] = resid[i] + (out[i] - resid[i]) * m[i] / (1-p)      (dropout on the branch of a residual sum, fp32)
int launch_dropout_residual(float* out, const float* resid, size_t n, const DropoutSite& s, cudaStream_t st);
// out[i] = g[i] * m[i] / (1-p), out in bf16 (out_bf16 = 1) or fp32 (the gradient that enters a dropped branch)
int launch_dropout_grad(const float* g, void* out, size_t n, int out_bf16, const DropoutSite& s, cudaStream_t st);
// keep[i] = m[i] as bytes (test hook: pins the generator against the numpy restatement in oracle/)
int launch_dropout_mask(uint8_t* keep, size_t n, const DropoutSite& s, cudaStream_t st);

}  // namespace svit
