// Micro-benchmark 2: tensor-pipe time per tcgen05.mma (M=128, K=16, bf16, SS) issued by ONE thread with uniform
// operands (TMEM base asserted to be 0 so ptxas needs no waterfall loop), optionally with "noise" warps that hammer
// shared memory (STS.128 / LDS.128) and TMEM (tcgen05.ld) the way the attention compute warps do.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace svit;

__global__ void k(int n, int noise, int iters, long long* out, float* sink) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
    fence_proxy_async_smem();
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (slot != 0) { if (threadIdx.x == 0) out[2] = -1; return; }
    const uint32_t tb = 0;
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t hi = umma_desc_hi(1024);
            const uint32_t a0 = umma_desc_lo(smem_u32(smem), 16384), b0 = umma_desc_lo(smem_u32(smem + 32768), 8192);
            const uint32_t idesc = umma_idesc_bf16(128, n, 1, 1);
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int s = 0; s < 8; ++s) umma_ss_lohi(tb + 256 + (s & 1) * 64, a0 + s * 128, b0 + s * 128, hi, idesc, 1);
            }
            umma_commit(&bar);
            long long t1 = clock64();
            mbar_wait(&bar, 0);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0; out[2] = 0;
            done = 1;
        }
    } else if (warp >= 4 && noise) {
        // noise: tcgen05.ld of 32 columns + 8 STS.128 + 8 LDS.128 (+ 32 ex2) per round, like a P / dS warp
        const uint32_t t_row = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint8_t* row = smem + 65536 + ((warp - 4) & 3) * 4096 * 4 + lane * 128;
        float acc = 0.f;
        long long tn0 = clock64(); long long rounds = 0;
        while (!done) {
            ++rounds;
            uint32_t r[32];
            if (noise & 1) { tmem_ld_32x32(t_row, r); tmem_ld_wait(); } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = j;
            }
            if (noise & 2) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    *reinterpret_cast<uint4*>(row + ((g ^ (lane & 7)) << 4)) = make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
#pragma unroll
                for (int g = 0; g < 8; ++g) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(row + ((g ^ (lane & 7)) << 4)))); acc += __uint_as_float(v.x); }
            }
            if (noise & 4) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += exp2f(__uint_as_float(r[j]));
            }
        }
        long long tn1 = clock64();
        if (threadIdx.x == 128) { out[3] = tn1 - tn0; out[4] = rounds; }
        if (acc == 123.f) sink[0] = acc;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
    long long* d; float* s; cudaMalloc(&d, 48); cudaMalloc(&s, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    for (int noise : {1, 2, 4})
        for (int n : {64, 96, 128}) {
            long long h[5] = {0,0,0,0,1};
            k<<<1, 384, 140 * 1024>>>(n, noise, 256, d, s);
            cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
            printf("noise=%d (1=tmem_ld 2=sts/lds 4=ex2, 8 warps)  N=%3d: issue %.1f, retire %.1f cyc/MMA (ideal %d) %s | noise warp: %.0f cyc per round\n", noise, n,
                   (double)h[0] / 2048, (double)h[1] / 2048, n / 2, h[2] ? "TMEM base != 0" : "", (double)h[3] / (h[4] ? h[4] : 1));
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
