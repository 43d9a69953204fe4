// Host-side shared helpers: error reporting for the C ABI and TMA descriptor creation.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace utx {

// Last-error slot behind utx_last_error() (thread-local: the ABI has no global state).
void set_error(const std::string& msg);
const char* get_error();

#define UTX_CHECK(cond, msg)                                                           \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::utx::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define UTX_CUDA(expr)                                                                 \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::utx::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                       cudaGetErrorString(_e));                                        \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define UTX_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

// 2-D bf16 tensor map: global [rows, cols] with row stride ld (elements), box [box_rows, box_cols],
// 128B swizzle when box_cols*2 == 128, no swizzle otherwise.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// 4-D NHWC bf16 activation map for implicit-GEMM 3x3 convolutions: dims (C, W, H, N), box (64 channels, box_w, box_h, 1),
// 128B swizzle; out-of-range coordinates (the padding ring, the tail past the last image) read as zeros.
int make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, uint64_t N, uint64_t H, uint64_t W, uint64_t C, uint32_t box_w,
                        uint32_t box_h);

int num_sms();   // of the CURRENT device

// Opt-in kernel attributes (dynamic shared memory size) are per device: `static PerDeviceOnce once; if (once.first()) {...}`
// runs the block once for every device a process drives, not once per process.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

}  // namespace utx
