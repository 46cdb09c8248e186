// Host-side helpers shared by every translation unit of libcreamfl_b200: error channel, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

// Status codes of the C ABI (include/creamfl_b200.h).
#define CFL_OK 0
#define CFL_EINVAL (-1)
#define CFL_EWORKSPACE (-2)
#define CFL_ECUDA (-3)

namespace cfl {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

// 2-D row-major bf16/fp32 array [rows, cols] (cols contiguous, `ld` elements between rows) as a TMA map
// with a (box_cols x box_rows) box and 128-byte swizzle.  Returns CFL_OK or a negative status.
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_cols, uint32_t box_rows);

// NHWC bf16 activation [n, h, w, c] as a 4-D map (c, w, h, n) with box (box_c, box_w, box_h, box_n), element
// traversal stride `stride` along w and h (strided convolutions); out-of-bound box elements read as zero.
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c,
                   uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n, uint32_t stride);

}  // namespace cfl
