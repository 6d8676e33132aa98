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
    const uint32_t span = box_inner * esz;  // bytes per box row = swizzle span
    if ((span != 128 && span != 64) || box_outer == 0 || box_outer > 256 || (row_pitch_bytes & 15) ||
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
                          span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

int make_tmap_3d(CUtensorMap* out, const void* gptr, TmapDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0, uint32_t box1) {
    std::call_once(g_once, resolve);
    if (!g_encode) return -1;
    const uint32_t esz = dt == TmapDtype::BF16 ? 2 : 4;
    if (box0 * esz != 128 || box1 == 0 || box1 > 256 || (pitch1_bytes & 15) || (pitch2_bytes & 15) ||
        (reinterpret_cast<uintptr_t>(gptr) & 15)) {
        snprintf(g_err, sizeof(g_err), "make_tmap_3d: bad geometry ptr=%p dims=%llu,%llu,%llu pitch=%llu,%llu box=%ux%u",
                 gptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                 (unsigned long long)pitch1_bytes, (unsigned long long)pitch2_bytes, box0, box1);
        return -2;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, dt == TmapDtype::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          3, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p dims=%llu,%llu,%llu", (int)r, gptr,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
        return (int)r;
    }
    return 0;
}

}  // namespace svit
