// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace lvd {

// status codes of the C-ABI (include/lvd_b200.h)
enum { LVD_OK = 0, LVD_ERR_INVALID = -1, LVD_ERR_UNSUPPORTED = -2, LVD_ERR_CUDA = -3 };

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define LVD_CHECK_CUDA(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::lvd::cuda_fail(_e, #expr); \
  } while (0)

#define LVD_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::lvd::set_error(__VA_ARGS__);    \
      return ::lvd::LVD_ERR_INVALID;    \
    }                                   \
  } while (0)

// Encode a tiled fp16 tensor map with 128-byte swizzle. dims/strides are in ELEMENTS, dim 0 is contiguous
// (strides[0] must be 1); box[0] must be 64 (=128 bytes). Out-of-bounds elements read as zero.
int encode_tmap_f16_sw128(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                          const int* box);
// Same with 64-byte swizzle and box[0] == 32 (64-byte rows): the output / residual slabs of the GEMM epilogue.
int encode_tmap_f16_sw64(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                         const int* box);

int num_sms();

}  // namespace lvd
