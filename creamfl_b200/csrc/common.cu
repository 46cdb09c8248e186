#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

namespace cfl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return CFL_ECUDA;
  }
  return CFL_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// The driver-API encoder needs a current context on the calling thread; PyTorch's autograd worker threads may
// reach us before any runtime call has bound the primary context to them.
static void bind_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  bind_context();
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_cols, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CFL_ECUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (base=%p ld=%llu)", base,
              (unsigned long long)ld);
    return CFL_EINVAL;
  }
  if (box_cols * elem_bytes > 128 || box_rows > 256) {
    set_error("TMA box %ux%u too large for 128B swizzle", box_cols, box_rows);
    return CFL_EINVAL;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u) failed: %d",
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows,
              (int)r);
    return CFL_ECUDA;
  }
  return CFL_OK;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c,
                   uint32_t box_c, uint32_t box_w, uint32_t box_h, uint32_t box_n, uint32_t stride) {
  auto fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CFL_ECUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((c * 2) & 15)) {
    set_error("NHWC TMA operand needs 16-byte aligned base and C %% 8 == 0");
    return CFL_EINVAL;
  }
  cuuint64_t dims[4] = {c, w, h, n};
  cuuint64_t strides[3] = {c * 2, w * c * 2, h * w * c * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, box_n};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(nhwc) failed: %d", (int)r);
    return CFL_ECUDA;
  }
  return CFL_OK;
}

}  // namespace cfl

extern "C" const char* creamfl_last_error(void) { return cfl::last_error(); }
