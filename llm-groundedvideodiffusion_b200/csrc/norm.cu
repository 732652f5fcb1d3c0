// norm.cu -- GroupNorm (4-D per-frame and 5-D cross-frame statistics) and LayerNorm on channels-last fp16.
// HBM/L2-bound: 16-byte vector loads, fp32 statistics, warp-shuffle / shared-memory reductions.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

__device__ __forceinline__ void ld8(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// x is viewed as [samples][rows_per_sample][C] (row stride ld).  4-D GroupNorm: sample = one frame image;
// 5-D GroupNorm (TemporalConvLayer / TransformerTemporalModel): sample = one video, rows span all frames.
// Pass 1 writes per-(sample, chunk, group) partial {sum, sumsq}; pass 2 folds the chunks (deterministic order),
// normalises, applies the affine and optional SiLU.  Thread -> 8-channel vector mapping is fixed so the
// per-channel accumulators / affine live in registers.
__global__ void gn_stats_kernel(const __half* __restrict__ x, long long ld, long long rows_per_sample, int C, int G,
                                int chunks, float* __restrict__ partial) {
  extern __shared__ float s_part[];  // [rows per pass][2*C]: per-thread partial sums, folded in a FIXED order below
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;  // rows per pass
  const int v = threadIdx.x % vpr;
  const int rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  if (rsub < rpp) {
    const __half* base = x + (s * rows_per_sample) * ld + v * 8;
    long long r = r0 + rsub;
    for (; r + 3LL * rpp < r1; r += 4LL * rpp) {  // 4 independent 16-byte loads in flight per thread
      float t0[8], t1[8], t2[8], t3[8];
      ld8(base + r * ld, t0);
      ld8(base + (r + rpp) * ld, t1);
      ld8(base + (r + 2LL * rpp) * ld, t2);
      ld8(base + (r + 3LL * rpp) * ld, t3);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum[i] += (t0[i] + t1[i]) + (t2[i] + t3[i]);
        sq[i] += (t0[i] * t0[i] + t1[i] * t1[i]) + (t2[i] * t2[i] + t3[i] * t3[i]);
      }
    }
    for (; r < r1; r += rpp) {
      float t[8];
      ld8(base + r * ld, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum[i] += t[i];
        sq[i] += t[i] * t[i];
      }
    }
    // no atomics: the summation order (and therefore every output bit) is the same on every run
    float* mine = s_part + static_cast<long long>(rsub) * 2 * C + v * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = sum[i];
      mine[C + i] = sq[i];
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rs = 0; rs < rpp; ++rs) {
      const float* row = s_part + static_cast<long long>(rs) * 2 * C;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        a += row[c];
        b += row[C + c];
      }
    }
    float* o = partial + ((static_cast<long long>(s) * chunks + chunk) * G + g) * 2;
    o[0] = a;
    o[1] = b;
  }
}

// Folds the per-chunk partial sums of one sample into {mean, rstd} per group (deterministic order).  A separate
// tiny launch: with 5-D GroupNorm (2 samples x ~600 chunks) doing this in every apply CTA serialised ~600 dependent
// loads in front of the streaming loop.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, int G, float n, float eps,
                                   float* __restrict__ stats, int raw_means) {
  __shared__ float s_a[256], s_b[256];
  const int s = blockIdx.x;
  const int g = threadIdx.x % G, part = threadIdx.x / G, nparts = blockDim.x / G;
  float a = 0.f, b = 0.f;
  if (part < nparts)
    for (int k = part; k < chunks; k += nparts) {
      const float* p = partial + ((static_cast<long long>(s) * chunks + k) * G + g) * 2;
      a += p[0];
      b += p[1];
    }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x < G) {
    a = 0.f;
    b = 0.f;
    for (int q = 0; q < nparts; ++q) {
      a += s_a[q * G + threadIdx.x];
      b += s_b[q * G + threadIdx.x];
    }
    const float mean = a / n;
    float var = b / n - mean * mean;
    var = var > 0.f ? var : 0.f;
    stats[(static_cast<long long>(s) * G + threadIdx.x) * 2] = mean;
    stats[(static_cast<long long>(s) * G + threadIdx.x) * 2 + 1] = raw_means ? b / n : rsqrtf(var + eps);
  }
}

__global__ void gn_apply_kernel(const __half* __restrict__ x, long long ld_x, __half* __restrict__ y, long long ld_y,
                                long long rows_per_sample, int C, int G, const float* __restrict__ stats,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, int silu,
                                int chunks) {
  __shared__ float s_mean[64], s_rstd[64];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr;
  const int rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_mean[g] = stats[(static_cast<long long>(s) * G + g) * 2];
    s_rstd[g] = stats[(static_cast<long long>(s) * G + g) * 2 + 1];
  }
  __syncthreads();
  if (rsub >= rpp) return;
  float ga[8], be[8], mu[8], rs[8];
  {
    float t[8];
    ld8(gamma + v * 8, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) ga[i] = t[i];
    ld8(beta + v * 8, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      be[i] = t[i];
      const int g = (v * 8 + i) / cpg;
      mu[i] = s_mean[g];
      rs[i] = s_rstd[g];
    }
  }
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  const __half* xb = x + (s * rows_per_sample) * ld_x + v * 8;
  __half* yb = y + (s * rows_per_sample) * ld_y + v * 8;
  long long r = r0 + rsub;
  for (; r + rpp < r1; r += 2LL * rpp) {  // two rows per iteration: loads issued before the dependent math
    float t[8], u[8];
    ld8(xb + r * ld_x, t);
    ld8(xb + (r + rpp) * ld_x, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float o = round_h((t[i] - mu[i]) * rs[i] * ga[i] + be[i]);
      float o2 = round_h((u[i] - mu[i]) * rs[i] * ga[i] + be[i]);
      if (silu) {
        o = __fdividef(o, 1.0f + __expf(-o));
        o2 = __fdividef(o2, 1.0f + __expf(-o2));
      }
      t[i] = o;
      u[i] = o2;
    }
    st8(yb + r * ld_y, t);
    st8(yb + (r + rpp) * ld_y, u);
  }
  for (; r < r1; r += rpp) {
    float t[8];
    ld8(xb + r * ld_x, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float o = round_h((t[i] - mu[i]) * rs[i] * ga[i] + be[i]);
      if (silu) o = __fdividef(o, 1.0f + __expf(-o));
      t[i] = o;
    }
    st8(yb + r * ld_y, t);
  }
}

static int gn_threads(int C) {
  const int vpr = C / 8;
  int rpp = 256 / vpr;
  if (rpp < 1) rpp = 1;
  return rpp * vpr;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row stays in registers (C <= 2048), exact two-pass mean / variance in fp32.
constexpr int kLnMaxVec = 8;  // 8 vectors of 8 halves per lane -> C <= 2048
// NV = ceil(C / 256) vectors per lane: a template parameter so that narrow rows (C = 320: NV = 2) hold 16 instead of
// 64 registers of row data and many more rows are in flight per SM.
template <int NV>
__global__ void layernorm_kernel(const __half* __restrict__ x, long long ld_x, __half* __restrict__ y, long long ld_y,
                                 long long rows, int C, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, float eps) {
  extern __shared__ float s_affine[];  // gamma[C], beta[C] as fp32, loaded once per block
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_affine[i] = __half2float(gamma[i]);
    s_affine[C + i] = __half2float(beta[i]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int vpr = C >> 3;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_total) {
    float vals[NV][8];
    float sum = 0.f;
    const __half* xr = x + row * ld_x;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vpr) {
        ld8(xr + v * 8, vals[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += vals[k][i];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vpr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = vals[k][i] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / C + eps);
    __half* yr = y + row * ld_y;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + k * 32;
      if (v < vpr) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = (vals[k][i] - mean) * rstd * s_affine[v * 8 + i] + s_affine[C + v * 8 + i];
        st8(yr + v * 8, o);
      }
    }
  }
}

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_groupnorm(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t samples,
                             int64_t rows_per_sample, int32_t channels, int32_t groups, const void* gamma,
                             const void* beta, float eps, int32_t silu, void* stats_out, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  LVD_REQUIRE(x && y && gamma && beta && workspace, "lvd_groupnorm: null pointer");
  LVD_REQUIRE(channels % 8 == 0 && channels % groups == 0 && groups <= 64 && channels <= 4096,
              "lvd_groupnorm: C=%d G=%d unsupported", channels, groups);
  LVD_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0, "lvd_groupnorm: ld must be a multiple of 8");
  LVD_REQUIRE(samples >= 1 && rows_per_sample >= 1, "lvd_groupnorm: empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // enough CTAs to fill 148 SMs a few times over, but at least ~64 rows per chunk
  long long chunks = (8LL * num_sms() + samples - 1) / samples;
  const long long max_chunks = (rows_per_sample + 15) / 16;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  const long long need = static_cast<long long>(samples) * chunks * groups * 2 * sizeof(float);
  LVD_REQUIRE(workspace_bytes >= need + static_cast<long long>(samples) * groups * 2 * sizeof(float),
              "lvd_groupnorm: workspace %lld < %lld bytes", (long long)workspace_bytes, need);
  const int threads = gn_threads(channels);
  LVD_REQUIRE(threads <= 1024, "lvd_groupnorm: C=%d too wide", channels);
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(samples));
  gn_stats_kernel<<<grid, threads, static_cast<size_t>(threads / (channels / 8)) * 2 * channels * sizeof(float), st>>>(
      static_cast<const __half*>(x), ld_x, rows_per_sample, channels, groups, static_cast<int>(chunks),
      static_cast<float*>(workspace));
  float* stats = stats_out ? static_cast<float*>(stats_out) : static_cast<float*>(workspace) + need / sizeof(float);
  gn_finalize_kernel<<<samples, 256 / groups * groups, 0, st>>>(static_cast<const float*>(workspace),
                                                               static_cast<int>(chunks), groups,
                                                               static_cast<float>(rows_per_sample) * (channels / groups),
                                                               eps, stats, 0);
  gn_apply_kernel<<<grid, threads, 0, st>>>(static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y,
                                            rows_per_sample, channels, groups, stats,
                                            static_cast<const __half*>(gamma), static_cast<const __half*>(beta), silu,
                                            static_cast<int>(chunks));
  g_launches.fetch_add(3, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_layernorm(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels,
                             const void* gamma, const void* beta, float eps, void* stream) {
  LVD_REQUIRE(x && y && gamma && beta, "lvd_layernorm: null pointer");
  LVD_REQUIRE(channels % 8 == 0 && channels <= kLnMaxVec * 256, "lvd_layernorm: C=%d unsupported", channels);
  LVD_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0, "lvd_layernorm: ld must be a multiple of 8");
  if (rows <= 0) return LVD_OK;
  const int warps = 8;
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  const int nv = (channels / 8 + 31) / 32;
#define LVD_LN_LAUNCH(NV)                                                                                         \
  layernorm_kernel<NV><<<static_cast<unsigned>(blocks), warps * 32, 2 * channels * sizeof(float),                \
                         static_cast<cudaStream_t>(stream)>>>(                                                   \
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows, channels,                        \
      static_cast<const __half*>(gamma), static_cast<const __half*>(beta), eps)
  if (nv <= 2) {
    LVD_LN_LAUNCH(2);
  } else if (nv <= 3) {
    LVD_LN_LAUNCH(3);
  } else if (nv <= 5) {
    LVD_LN_LAUNCH(5);
  } else {
    LVD_LN_LAUNCH(8);
  }
#undef LVD_LN_LAUNCH
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
