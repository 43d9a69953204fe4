#include "common.h"

#include <mutex>

namespace utx {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime's driver entry point, so the library
// links (and its symbols can be checked) on a box without a driver.
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  UTX_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  UTX_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  UTX_CHECK((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  UTX_CHECK(box_rows <= 256 && box_cols <= 256, "TMA box dims must be <= 256");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UTX_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

int make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t box_w,
                        uint32_t box_h) {
  EncodeTiledFn fn = encode_fn();
  UTX_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  UTX_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  UTX_CHECK(C % 64 == 0 && box_w <= 256 && box_h <= 256, "NHWC TMA map: C must be a multiple of 64, box dims <= 256");
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UTX_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (NHWC) failed with CUresult " + std::to_string((int)r));
  return 0;
}

int num_sms() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

}  // namespace utx
