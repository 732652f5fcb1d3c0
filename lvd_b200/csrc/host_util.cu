#include "host_util.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace lvd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return LVD_ERR_CUDA;
}

const char* last_error() { return g_err; }

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so that this library does not
// link libcuda directly (it must dlopen on a CPU-only box for the symbol-export test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_tmap_impl(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                            const int* box, int box0, CUtensorMapSwizzle swz);

int encode_tmap_f16_sw128(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                          const int* box) {
  return encode_tmap_impl(out, base, rank, dims, strides, box, 64, CU_TENSOR_MAP_SWIZZLE_128B);
}
int encode_tmap_f16_sw64(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                         const int* box) {
  return encode_tmap_impl(out, base, rank, dims, strides, box, 32, CU_TENSOR_MAP_SWIZZLE_64B);
}

static int encode_tmap_impl(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                            const int* box, int box0, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return LVD_ERR_CUDA;
  }
  if (rank < 2 || rank > 5 || strides[0] != 1 || box[0] != box0) {
    set_error("encode_tmap: bad rank/stride0/box0 (%d, %lld, %d)", rank, (long long)strides[0], box[0]);
    return LVD_ERR_INVALID;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = 1;
    if (dims[i] <= 0 || box[i] <= 0 || box[i] > 256) {
      set_error("encode_tmap: bad dim/box at %d: dim=%lld box=%d", i, (long long)dims[i], box[i]);
      return LVD_ERR_INVALID;
    }
    if (i > 0) {
      gstr[i - 1] = (cuuint64_t)strides[i] * 2;
      if ((gstr[i - 1] & 15) != 0) {
        set_error("encode_tmap: stride %d (%lld elements) is not a multiple of 16 bytes", i, (long long)strides[i]);
        return LVD_ERR_INVALID;
      }
    }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("encode_tmap: base pointer not 16-byte aligned");
    return LVD_ERR_INVALID;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %lld,%lld,.. box %d,%d,..)", (int)r, rank,
              (long long)dims[0], (long long)dims[1], box[0], box[1]);
    return LVD_ERR_CUDA;
  }
  return LVD_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace lvd

extern "C" const char* lvd_last_error(void) { return lvd::last_error(); }
