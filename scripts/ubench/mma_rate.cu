// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) as a function of N, operand majorness and
// A source (smem / TMEM), issued back to back by one thread.  Sizes the attention kernels' MMA budgets.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace svit;

// mode: 0 = SS both K-major, 1 = SS A MN-major + B MN-major, 2 = SS A K-major + B MN-major, 3 = TS (A in TMEM) + B MN-major
__global__ void k_mma(int n, int mode, int iters, int nd, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, nd); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if ((threadIdx.x & 31) == 0 && warp < nd) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
        const uint32_t idesc = umma_idesc_bf16(128, n, (mode == 1) ? 1 : 0, (mode >= 1) ? 1 : 0);
        uint64_t adv[4], bdv[4];
        uint32_t dcol[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (mode == 1) adv[k] = umma_smem_desc(a_addr + k * 2048, 16384, 1024);
            else adv[k] = umma_smem_desc(a_addr + k * 32, 16, 1024);
            if (mode == 0) bdv[k] = umma_smem_desc(b_addr + k * 32, 16, 1024);
            else bdv[k] = umma_smem_desc(b_addr + k * 2048, 8192, 1024);
            dcol[k] = tb + warp * 64;
        }
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 3) umma_ts(dcol[k], tb + 384 + k * 8, bdv[k], idesc, 1);
                else umma_ss(dcol[k], adv[k], bdv[k], idesc, 1);
            }
        }
        umma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 256;
    const char* names[4] = {"SS K-major x K-major", "SS MN-major x MN-major", "SS K-major x MN-major", "TS (A in TMEM) x MN-major"};
    for (int mode = 0; mode < 4; ++mode)
        for (int nd : {1, 2, 4})
            for (int n : {64, 128, 256}) {
                if (n > 64 && nd > 1) continue;
                long long h[2];
                k_mma<<<1, 128, 100 * 1024>>>(n, mode, iters, nd, d);
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%-28s N=%3d issuing warps=%d: issue %.1f cyc/MMA, retire %.1f cyc/MMA  (ideal %d)\n", names[mode], n, nd,
                       (double)h[0] / (iters * 4 * nd), (double)h[1] / (iters * 4 * nd), n / 2);
            }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
