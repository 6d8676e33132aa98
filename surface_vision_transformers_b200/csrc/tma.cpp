#include "tma.h"

#include <cstdio>
#include <mutex>

namespace svit {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;
static char g_err[256] = "";

const char* tmap_last_error() { return g_err; }

static void resolve() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        snprintf(g_err, sizeof(g_err), "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                 cudaGetErrorString(e));
        g_encode = nullptr;
        return;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

int make_tmap_2d(CUtensorMap* out, const void* gptr, TmapDtype dt, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
    std::call_once(g_once, resolve);
    if (!g_encode) return -1;
    const uint32_t esz = dt == TmapDtype::BF16 ? 2 : 4;
    if (box_inner * esz != 128 || box_outer == 0 || box_outer > 256 || (row_pitch_bytes & 15) ||
        (reinterpret_cast<uintptr_t>(gptr) & 15)) {
        snprintf(g_err, sizeof(g_err),
                 "make_tmap_2d: bad geometry ptr=%p inner=%llu outer=%llu pitch=%llu box=%ux%u", gptr,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_pitch_bytes, box_inner,
                 box_outer);
        return -2;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, dt == TmapDtype::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          2, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err),
                 "cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r, gptr,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_pitch_bytes, box_inner,
                 box_outer);
        return (int)r;
    }
    return 0;
}

}  // namespace svit
