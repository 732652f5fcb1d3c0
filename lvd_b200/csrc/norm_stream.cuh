// norm_stream.cuh -- the shared-memory ring the streaming GroupNorm passes (norm.cu, backward.cu) read their rows through.
#pragma once
#include "ptx.cuh"

namespace lvd {

constexpr int kGnStages = 5;

struct GnStream {
  uint64_t* full;      // [kGnStages]
  uint4* ring;         // [kGnStages][T][vpr] (two-source streams: [kGnStages][2][T][vpr])
  const __half* src;   // first row of this CTA's range
  long long ld;
  const __half* src2 = nullptr;  // optional second tensor streamed in lock step (same rows, own stride)
  long long ld2 = 0;
  int n, T, vpr, nit;
  bool dense;
  __device__ __forceinline__ void issue(int it) const {  // whole warp 0
    const int stage = it % kGnStages;
    const int rows = n - it * T < T ? n - it * T : T;
    const uint32_t row_bytes = static_cast<uint32_t>(vpr) * 16u;
    const int lane = threadIdx.x & 31;
    const int nsrc = src2 ? 2 : 1;
    if (lane == 0) mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(nsrc * rows) * row_bytes);
    __syncwarp();
    uint4* dst = ring + static_cast<long long>(stage) * nsrc * T * vpr;
    const __half* from = src + static_cast<long long>(it) * T * ld;
    if (dense) {
      if (lane == 0) bulk_load_1d(dst, from, static_cast<uint32_t>(rows) * row_bytes, &full[stage]);
    } else {
      for (int r = lane; r < rows; r += 32)
        bulk_load_1d(dst + static_cast<long long>(r) * vpr, from + r * ld, row_bytes, &full[stage]);
    }
    if (src2) {
      uint4* dst2 = dst + static_cast<long long>(T) * vpr;
      const __half* from2 = src2 + static_cast<long long>(it) * T * ld2;
      if (ld2 == static_cast<long long>(vpr) * 8) {
        if (lane == 0) bulk_load_1d(dst2, from2, static_cast<uint32_t>(rows) * row_bytes, &full[stage]);
      } else {
        for (int r = lane; r < rows; r += 32)
          bulk_load_1d(dst2 + static_cast<long long>(r) * vpr, from2 + r * ld2, row_bytes, &full[stage]);
      }
    }
  }
};


}  // namespace lvd
