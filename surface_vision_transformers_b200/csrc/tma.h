// tma.h -- host-side CUtensorMap construction (driver entry point resolved at run time,
// so the shared library has no link-time dependency on libcuda and loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace svit {

enum class TmapDtype { BF16, F32 };

// 2-D row-major tensor [outer][inner] with a row pitch in bytes, swizzled box.
// box_inner * elem_size must be 128 bytes (SWIZZLE_128B) or 64 bytes (SWIZZLE_64B); box_outer <= 256.
// Returns 0 on success, a CUresult otherwise.
int make_tmap_2d(CUtensorMap* out, const void* gptr, TmapDtype dt, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer);

// 3-D tensor [d2][d1][d0] (d0 innermost), 128-byte swizzled box (box0 x box1 x 1).
int make_tmap_3d(CUtensorMap* out, const void* gptr, TmapDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0, uint32_t box1);

const char* tmap_last_error();

}  // namespace svit
