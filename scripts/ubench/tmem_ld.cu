// Micro-benchmark: tcgen05.ld throughput per SM (how many bytes per clock can warps pull out of TMEM?) and
// MUFU.EX2 throughput, to size the softmax stages of the attention kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../surface_vision_transformers_b200/csrc tmem_ld.cu -o tmem_ld
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace svit;

__global__ void k_tmem(int iters, int nwarps_active, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps_active) {
        for (int i = 0; i < iters; ++i) {
            uint32_t r[32];
            tmem_ld_32x32(base + ((i * 32 + (warp >> 2) * 64) & 255), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}
// two loads in flight before the wait
__global__ void k_tmem2(int iters, int nwarps_active, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps_active) {
        for (int i = 0; i < iters; i += 2) {
            uint32_t r[32], s[32];
            tmem_ld_32x32(base + ((i * 32) & 255), r);
            tmem_ld_32x32(base + ((i * 32 + 32) & 255), s);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]) * __uint_as_float(s[j]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}
__global__ void k_mufu(int iters, int nwarps_active, long long* out, float* sink) {
    const int warp = threadIdx.x >> 5;
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps_active) {
        for (int i = 0; i < iters; ++i) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (a0 + a1 + a2 + a3 == 123.456f) sink[0] = a0;
}
int main() {
    long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    const int iters = 4096;
    for (int nw : {1, 2, 4, 8, 16}) {
        long long h;
        k_tmem<<<1, 512>>>(iters, nw, d, s); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("tmem_ld x32 (1 in flight)  warps=%2d: %8lld cyc  -> %.1f B/clk/SM\n", nw, h, (double)nw * iters * 4096 / h);
        k_tmem2<<<1, 512>>>(iters, nw, d, s); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("tmem_ld x32 (2 in flight)  warps=%2d: %8lld cyc  -> %.1f B/clk/SM\n", nw, h, (double)nw * iters * 4096 / h);
        k_mufu<<<1, 512>>>(iters, nw, d, s); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("ex2.approx                 warps=%2d: %8lld cyc  -> %.2f ex2/clk/SM\n", nw, h, (double)nw * iters * 4 * 32 / h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
