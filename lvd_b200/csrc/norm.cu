// norm.cu -- GroupNorm (4-D per-frame and 5-D cross-frame statistics) and LayerNorm on channels-last fp16.
// HBM/L2-bound: 16-byte vector loads, fp32 statistics, warp-shuffle / shared-memory reductions.
#include <atomic>
#include <cstdlib>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"
#include "norm_stream.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;
constexpr int kGnApplyCtasPerSm = 3;

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
__device__ __forceinline__ void unpack8v(const uint4& u, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
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
// Sum of the `count` partial {sum, sumsq} pairs of (sample s, group g) that thread `part` of `nparts` owns: four
// independent loads in flight, fixed order (deterministic).
__device__ __forceinline__ void fold_partials(const float* __restrict__ partial, int s, int count, int G, int g,
                                              int part, int nparts, float& a, float& b) {
  const float2* p = reinterpret_cast<const float2*>(partial) + static_cast<long long>(s) * count * G + g;
  int k = part;
  for (; k + 3 * nparts < count; k += 4 * nparts) {
    const float2 v0 = __ldcg(p + static_cast<long long>(k) * G);
    const float2 v1 = __ldcg(p + static_cast<long long>(k + nparts) * G);
    const float2 v2 = __ldcg(p + static_cast<long long>(k + 2 * nparts) * G);
    const float2 v3 = __ldcg(p + static_cast<long long>(k + 3 * nparts) * G);
    a += (v0.x + v1.x) + (v2.x + v3.x);
    b += (v0.y + v1.y) + (v2.y + v3.y);
  }
  for (; k < count; k += nparts) {
    const float2 v0 = __ldcg(p + static_cast<long long>(k) * G);
    a += v0.x;
    b += v0.y;
  }
}

// Pass 1 writes per-(sample, chunk, group) partial {sum, sumsq}; pass 2 folds the chunks (deterministic order),
// normalises, applies the affine and optional SiLU.  Thread -> 8-channel vector mapping is fixed so the
// per-channel accumulators / affine live in registers.
// Per-sample arrival tickets of gn_stats_kernel (zero at module load, reset by the last CTA of each sample): the
// library serialises GroupNorm launches on one stream per handle (include/lvd_b200.h, threading), so one set suffices.
constexpr int kGnMaxSamples = 8192;
__device__ unsigned int g_gn_tickets[kGnMaxSamples];

__global__ void gn_stats_kernel(const __half* __restrict__ x, long long ld, long long rows_per_sample, int C, int G,
                                int chunks, float* __restrict__ partial, float n, float eps, float* __restrict__ stats) {
  extern __shared__ float s_part[];  // [rows per pass][2*C] per-thread partial sums, then [2*C] column sums
  __shared__ int s_last;
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
    for (; r < r1; r += 8LL * rpp) {  // 8 independent 16-byte loads in flight per thread (predicated at the chunk end)
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long rk = r + static_cast<long long>(k) * rpp;
        u[k] = rk < r1 ? *reinterpret_cast<const uint4*>(base + rk * ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        float t0[8], t1[8];
        unpack8v(u[k], t0);
        unpack8v(u[k + 1], t1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sum[i] += t0[i] + t1[i];
          sq[i] += t0[i] * t0[i] + t1[i] * t1[i];
        }
      }
    }
    // no floating-point atomics anywhere: the summation order (and therefore every output bit) is the same on every run
    float* mine = s_part + static_cast<long long>(rsub) * 2 * C + v * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = sum[i];
      mine[C + i] = sq[i];
    }
  }
  __syncthreads();
  // fold rows-per-pass -> one value per column (all threads), then columns -> groups
  float* s_col = s_part + static_cast<long long>(rpp) * 2 * C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int rs = 0; rs < rpp; ++rs) a += s_part[static_cast<long long>(rs) * 2 * C + c];
    s_col[c] = a;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int t = threadIdx.x; t < 2 * G; t += blockDim.x) {
    const int g = t >> 1, which = t & 1;
    const float* col = s_col + which * C + g * cpg;
    float a = 0.f;
    for (int c = 0; c < cpg; ++c) a += col[c];
    partial[((static_cast<long long>(s) * chunks + chunk) * G + g) * 2 + which] = a;
  }
  // the last CTA of the sample to arrive folds the chunk partials into {mean, rstd} (fixed order: deterministic)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&g_gn_tickets[s], 1u) == static_cast<unsigned>(chunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* s_a = s_part;  // [blockDim.x] x 2
  float* s_b = s_part + blockDim.x;
  const int nparts = blockDim.x / G;
  {
    const int g = threadIdx.x % G, part = threadIdx.x / G;
    float a = 0.f, b = 0.f;
    if (part < nparts) fold_partials(partial, s, chunks, G, g, part, nparts, a, b);
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float a = 0.f, b = 0.f;
    for (int q = 0; q < nparts; ++q) {
      a += s_a[q * G + threadIdx.x];
      b += s_b[q * G + threadIdx.x];
    }
    const float mean = a / n;
    float var = b / n - mean * mean;
    var = var > 0.f ? var : 0.f;
    stats[(static_cast<long long>(s) * G + threadIdx.x) * 2] = mean;
    stats[(static_cast<long long>(s) * G + threadIdx.x) * 2 + 1] = rsqrtf(var + eps);
  }
  if (threadIdx.x == 0) g_gn_tickets[s] = 0u;  // ready for the next launch on this stream
}

// Folds per-chunk partial sums of one sample into per-group means (used by the backward pass, backward.cu).
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
  // y = x * sc + sh with sc = rstd * gamma, sh = beta - mean * rstd * gamma
  float sc[8], sh[8];
  {
    float ga[8], be[8];
    ld8(gamma + v * 8, ga);
    ld8(beta + v * 8, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (v * 8 + i) / cpg;
      sc[i] = s_rstd[g] * ga[i];
      sh[i] = be[i] - s_mean[g] * sc[i];
    }
  }
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  const __half* xb = x + (s * rows_per_sample) * ld_x + v * 8;
  __half* yb = y + (s * rows_per_sample) * ld_y + v * 8;
  long long r = r0 + rsub;
  for (; r < r1; r += 8LL * rpp) {  // eight rows per iteration (predicated at the chunk end): loads before the math
    uint4 u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long rk = r + static_cast<long long>(k) * rpp;
      if (rk < r1) u[k] = *reinterpret_cast<const uint4*>(xb + rk * ld_x);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long rk = r + static_cast<long long>(k) * rpp;
      if (rk < r1) {
        float t[8];
        unpack8v(u[k], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float o = round_h(fmaf(t[i], sc[i], sh[i]));
          if (silu) o = __fdividef(o, 1.0f + __expf(-o));
          t[i] = o;
        }
        st8(yb + rk * ld_y, t);
      }
    }
  }
}

// ---- streaming GroupNorm passes.  Two persistent CTAs per SM; each streams its row range through a ring of kGnStages
// shared-memory stages filled by 1-D bulk copies (cp.async.bulk, one per row, or one per stage when rows are dense):
// the bytes in flight (~75 KB per CTA) no longer depend on registers or occupancy, which is what held the
// register-staged kernels at half of the HBM rate.  Pass 1 = statistics (+ ticketed finalize), pass 2 = normalise.
// (kGnStages, GnStream: norm_stream.cuh)

__global__ void __launch_bounds__(320, 2) gn_stream_stats_kernel(const __half* __restrict__ x, long long ld,
                                                                  long long rows_per_sample, int C, int G, int chunks,
                                                                  int T, float* __restrict__ partial) {
  extern __shared__ uint4 s_dyn[];
  __shared__ uint64_t s_full[kGnStages];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr;
  const int rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  GnStream st;
  st.full = s_full;
  st.ring = s_dyn;
  st.src = x + (s * rows_per_sample + r0) * ld;
  st.ld = ld;
  st.n = r1 > r0 ? static_cast<int>(r1 - r0) : 0;
  st.T = T;
  st.vpr = vpr;
  st.nit = (st.n + T - 1) / T;
  st.dense = ld == C;
  float* s_part = reinterpret_cast<float*>(s_dyn + static_cast<long long>(kGnStages) * T * vpr);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGnStages; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int it = 0; it < kGnStages && it < st.nit; ++it) st.issue(it);
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  for (int it = 0; it < st.nit; ++it) {
    const int stage = it % kGnStages;
    mbar_wait(&s_full[stage], static_cast<uint32_t>(it / kGnStages) & 1u);
    const int rows = st.n - it * T < T ? st.n - it * T : T;
    const uint4* buf = s_dyn + static_cast<long long>(stage) * T * vpr;
    if (rsub < rpp) {
      for (int r = rsub; r < rows; r += rpp) {
        float t[8];
        unpack8v(buf[static_cast<long long>(r) * vpr + v], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sum[i] += t[i];
          sq[i] += t[i] * t[i];
        }
      }
    }
    __syncthreads();  // everyone is done with the stage before it is refilled
    if (threadIdx.x < 32 && it + kGnStages < st.nit) st.issue(it + kGnStages);
  }
  if (rsub < rpp) {
    float* mine = s_part + static_cast<long long>(rsub) * 2 * C + v * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = sum[i];
      mine[C + i] = sq[i];
    }
  }
  __syncthreads();
  float* s_col = s_part + static_cast<long long>(rpp) * 2 * C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int rs = 0; rs < rpp; ++rs) a += s_part[static_cast<long long>(rs) * 2 * C + c];
    s_col[c] = a;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int t = threadIdx.x; t < 2 * G; t += blockDim.x) {
    const int g = t >> 1, which = t & 1;
    const float* col = s_col + which * C + g * cpg;
    float a = 0.f;
    for (int c = 0; c < cpg; ++c) a += col[c];
    partial[((static_cast<long long>(s) * chunks + chunk) * G + g) * 2 + which] = a;
  }
}

template <bool kSilu>
__global__ void __launch_bounds__(320, kGnApplyCtasPerSm) gn_stream_apply_kernel(const __half* __restrict__ x, long long ld_x,
                                                                  __half* __restrict__ y, long long ld_y,
                                                                  long long rows_per_sample, int C, int G, int chunks,
                                                                  int pchunks, int T, const float* __restrict__ partial,
                                                                  float n,
                                                                  float eps, float* __restrict__ stats,
                                                                  const __half* __restrict__ gamma,
                                                                  const __half* __restrict__ beta) {
  extern __shared__ uint4 s_dyn[];
  __shared__ uint64_t s_full[kGnStages];
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float s_a[320], s_b[320];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr;
  const int rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  GnStream st;
  st.full = s_full;
  st.ring = s_dyn;
  st.src = x + (s * rows_per_sample + r0) * ld_x;
  st.ld = ld_x;
  st.n = r1 > r0 ? static_cast<int>(r1 - r0) : 0;
  st.T = T;
  st.vpr = vpr;
  st.nit = (st.n + T - 1) / T;
  st.dense = ld_x == C;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGnStages; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int it = 0; it < kGnStages && it < st.nit; ++it) st.issue(it);
  // Every CTA folds the sample's chunk partials itself (fixed order: deterministic, identical in all CTAs) while its
  // first stages are in flight -- a finalising "last CTA" in the statistics pass left the GPU idle for its whole tail.
  {
    const int nparts = blockDim.x / G;
    const int g = threadIdx.x % G, part = threadIdx.x / G;
    float a = 0.f, b = 0.f;
    if (part < nparts) fold_partials(partial, s, pchunks, G, g, part, nparts, a, b);
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
      const float rstd = rsqrtf(var + eps);
      s_mean[threadIdx.x] = mean;
      s_rstd[threadIdx.x] = rstd;
      if (chunk == 0) {  // kept for lvd_groupnorm_bwd (stats_out)
        stats[(static_cast<long long>(s) * G + threadIdx.x) * 2] = mean;
        stats[(static_cast<long long>(s) * G + threadIdx.x) * 2 + 1] = rstd;
      }
    }
    __syncthreads();
  }
  const int cpg = C / G;
  float sc[8], sh[8];
  if (rsub < rpp) {
    float ga[8], be[8];
    ld8(gamma + v * 8, ga);
    ld8(beta + v * 8, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (v * 8 + i) / cpg;
      sc[i] = s_rstd[g] * ga[i];
      sh[i] = be[i] - s_mean[g] * sc[i];
    }
  }
  __half* yb = y + (s * rows_per_sample + r0) * ld_y + v * 8;
  for (int it = 0; it < st.nit; ++it) {
    const int stage = it % kGnStages;
    mbar_wait(&s_full[stage], static_cast<uint32_t>(it / kGnStages) & 1u);
    const int rows = st.n - it * T < T ? st.n - it * T : T;
    const uint4* buf = s_dyn + static_cast<long long>(stage) * T * vpr;
    if (rsub < rpp) {
      for (int r = rsub; r < rows; r += rpp) {
        // The normalised value is rounded to fp16 (the reference's GroupNorm output tensor) with the PACKED conversion
        // and SiLU then works on that value: two MUFU operations per element (ex2, rcp) and no scalar F2F -- the
        // single-value conversions share the MUFU pipe, which bounds this pass.
        const uint4 u = buf[static_cast<long long>(r) * vpr + v];
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        uint4 o4;
        __half2* oh = reinterpret_cast<__half2*>(&o4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]);
          __half2 yh = __floats2half2_rn(fmaf(f.x, sc[2 * i], sh[2 * i]), fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]));
          if (kSilu) {
            const float2 o = __half22float2(yh);
            yh = __floats2half2_rn(o.x * fast_rcp(1.0f + fast_exp2(o.x * -1.4426950408889634f)),
                                   o.y * fast_rcp(1.0f + fast_exp2(o.y * -1.4426950408889634f)));
          }
          oh[i] = yh;
        }
        *reinterpret_cast<uint4*>(yb + (static_cast<long long>(it) * T + r) * ld_y) = o4;
      }
    }
    __syncthreads();
    if (threadIdx.x < 32 && it + kGnStages < st.nit) st.issue(it + kGnStages);
  }
}

// Geometry of the normalisation pass: kGnApplyCtasPerSm co-resident CTAs per SM (the pass is bound by instruction issue
// and the MUFU pipe, not by bytes in flight: a third CTA per SM buys more than deeper rings), stages of ~12 KB.
struct GnApplyPlan {
  int T;
  long long chunks;
  size_t ring_bytes;
};
static GnApplyPlan gn_apply_plan(int channels, int threads, int samples, long long rows_per_sample) {
  const long long row_bytes = 2LL * channels;
  const int rpp_ = threads / (channels / 8);
  GnApplyPlan p;
  p.T = static_cast<int>((12 * 1024) / row_bytes) / rpp_ * rpp_;
  if (p.T < rpp_) p.T = rpp_;
  p.chunks = (static_cast<long long>(kGnApplyCtasPerSm) * num_sms()) / samples;  // one wave of co-resident CTAs
  const long long most = (rows_per_sample + p.T - 1) / p.T;
  if (p.chunks > most) p.chunks = most;
  if (p.chunks < 1) p.chunks = 1;
  p.ring_bytes = static_cast<size_t>(kGnStages) * p.T * row_bytes;
  return p;
}

static int gn_threads(int C) {
  const int vpr = C / 8;
  int rpp = 256 / vpr;
  if (rpp < 1) rpp = 1;
  return rpp * vpr;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// L lanes per row (8 / 16 / 32, so that a row is NV <= 8 16-byte vectors per lane: C = 320 -> 8 lanes x 5, 640 -> 16 x 5,
// 1280 -> 32 x 5), 32 / L rows side by side in a warp and R such row sets in flight.  Rows are held PACKED (fp16) in
// registers, so a warp keeps (32 / L) * R rows -- 7.7 KB at C = 320 -- of loads outstanding before the first dependent
// shuffle: with one row per warp in flight the kernel sat at 1/3 of the HBM rate.  Exact two-pass mean / variance, fp32.
__device__ __forceinline__ void unpack8u(const uint4& u, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// Makes the compiler forget what it knows about a packed vector: the fp32 conversions are redone in every pass instead of
// all converted values being kept live (which doubled the register count and halved the rows in flight).
__device__ __forceinline__ void opaque(uint4& u) { asm volatile("" : "+r"(u.x), "+r"(u.y), "+r"(u.z), "+r"(u.w)); }
constexpr int kLnMaxVec = 8;
template <int NV, int L, int R>
__global__ void __launch_bounds__(256, 2) layernorm_kernel(const __half* __restrict__ x, long long ld_x,
                                                        __half* __restrict__ y, long long ld_y, long long rows, int C,
                                                        const __half* __restrict__ gamma,
                                                        const __half* __restrict__ beta, float eps) {
  extern __shared__ uint4 s_affine_v[];  // gamma[C], beta[C] (fp16), loaded once per block
  __half* s_affine = reinterpret_cast<__half*>(s_affine_v);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_affine[i] = gamma[i];
    s_affine[C + i] = beta[i];
  }
  __syncthreads();
  constexpr int kRowsPerSet = 32 / L;
  const int lane = threadIdx.x & 31;
  const int j = lane % L;       // lane within the row
  const int sub = lane / L;     // row within the set
  const int vpr = C >> 3;
  const float inv_c = 1.0f / static_cast<float>(C);
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const long long rows_per_iter = static_cast<long long>(kRowsPerSet) * R;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rows_per_iter;
       row0 < rows; row0 += warps_total * rows_per_iter) {
    uint4 vals[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r * kRowsPerSet + sub;
      const __half* xr = x + row * ld_x;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int v = j + k * L;
        vals[r][k] = (row < rows && v < vpr) ? *reinterpret_cast<const uint4*>(xr + v * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float f[8];
        unpack8u(vals[r][k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a += f[i];
      }
      mean[r] = a;
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < NV; ++k) opaque(vals[r][k]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] *= inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (j + k * L < vpr) {
          float f[8];
          unpack8u(vals[r][k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d = f[i] - mean[r];
            q += d * d;
          }
        }
      }
      rstd[r] = q;
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < NV; ++k) opaque(vals[r][k]);
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(rstd[r] * inv_c + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = j + k * L;
      if (v < vpr) {
        float ga[8], be[8];
        unpack8u(s_affine_v[v], ga);
        unpack8u(s_affine_v[vpr + v], be);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const long long row = row0 + r * kRowsPerSet + sub;
          if (row < rows) {
            float f[8], o[8];
            unpack8u(vals[r][k], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (f[i] - mean[r]) * rstd[r] * ga[i] + be[i];
            st8(y + row * ld_y + v * 8, o);
          }
        }
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
  LVD_REQUIRE(samples <= kGnMaxSamples, "lvd_groupnorm: more than %d samples", kGnMaxSamples);
  float* stats = stats_out ? static_cast<float*>(stats_out) : static_cast<float*>(workspace) + need / sizeof(float);
  {
    // streaming passes: ~2 persistent CTAs per SM, each with a ring of kGnStages stages of ~15 KB
    const long long row_bytes = 2LL * channels;
    const int rpp_ = threads / (channels / 8);
    int T = static_cast<int>((15 * 1024) / row_bytes) / rpp_ * rpp_;
    if (T < rpp_) T = rpp_;
    long long schunks = (2LL * num_sms()) / samples;  // one wave: never more CTAs than are co-resident (2 per SM)
    const long long most = (rows_per_sample + T - 1) / T;
    if (schunks > most) schunks = most;
    if (schunks > chunks) schunks = chunks;  // the workspace was sized for `chunks`
    if (schunks < 1) schunks = 1;
    const size_t ring_bytes = static_cast<size_t>(kGnStages) * T * row_bytes;
    const size_t scratch = (static_cast<size_t>(rpp_) + 1) * 2 * channels * sizeof(float);
    static bool attr2 = false;
    if (!attr2) {
      LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
      LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr2 = true;
    }
    if (threads <= 320 && ring_bytes + scratch <= 220 * 1024) {
    dim3 sgrid(static_cast<unsigned>(schunks), static_cast<unsigned>(samples));
    gn_stream_stats_kernel<<<sgrid, threads, ring_bytes + scratch, st>>>(
        static_cast<const __half*>(x), ld_x, rows_per_sample, channels, groups, static_cast<int>(schunks), T,
        static_cast<float*>(workspace));
    const GnApplyPlan ap = gn_apply_plan(channels, threads, samples, rows_per_sample);
    dim3 agrid(static_cast<unsigned>(ap.chunks), static_cast<unsigned>(samples));
    (silu ? gn_stream_apply_kernel<true> : gn_stream_apply_kernel<false>)<<<agrid, threads, ap.ring_bytes, st>>>(
        static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows_per_sample, channels, groups,
        static_cast<int>(ap.chunks), static_cast<int>(schunks), ap.T, static_cast<const float*>(workspace),
        static_cast<float>(rows_per_sample) * (channels / groups), eps, stats, static_cast<const __half*>(gamma),
        static_cast<const __half*>(beta));
    } else {  // very wide rows: register-staged passes
      dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(samples));
      const size_t smem_stats = (static_cast<size_t>(rpp_) + 1) * 2 * channels * sizeof(float);
      gn_stats_kernel<<<grid, threads, smem_stats, st>>>(
          static_cast<const __half*>(x), ld_x, rows_per_sample, channels, groups, static_cast<int>(chunks),
          static_cast<float*>(workspace), static_cast<float>(rows_per_sample) * (channels / groups), eps, stats);
      gn_apply_kernel<<<grid, threads, 0, st>>>(static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y,
                                                rows_per_sample, channels, groups, stats,
                                                static_cast<const __half*>(gamma), static_cast<const __half*>(beta),
                                                silu, static_cast<int>(chunks));
    }
  }
  g_launches.fetch_add(2, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

namespace lvd {
// Folds the per-warp partial sums written by gemm_tc_kernel's epilogue (gn_part, `slots` per sample) into kGnFoldParts
// partial pairs per (sample, group) that gn_stream_apply_kernel finishes, and clears them for the next producer.  CTA
// (c, sample) owns a contiguous range of slots; thread = (group, part) walks it with four loads in flight; every order
// is fixed (deterministic).
constexpr int kGnFoldParts = 16;
__global__ void __launch_bounds__(1024) gn_tiles_fold_kernel(float* __restrict__ part, int slots, int G,
                                                             float* __restrict__ partial) {
  __shared__ float s_a[1024], s_b[1024];
  const int c = blockIdx.x, s = blockIdx.y;
  const int g = threadIdx.x % G, p = threadIdx.x / G, nparts = blockDim.x / G;
  const int per = (slots + kGnFoldParts - 1) / kGnFoldParts;
  const int k0 = c * per, k1 = k0 + per < slots ? k0 + per : slots;
  float2* base = reinterpret_cast<float2*>(part) + static_cast<long long>(s) * slots * G + g;
  const float2 zero = make_float2(0.f, 0.f);
  float a = 0.f, b = 0.f;
  int k = k0 + p;
  for (; k + 3 * nparts < k1; k += 4 * nparts) {
    float2* q0 = base + static_cast<long long>(k) * G;
    float2* q1 = base + static_cast<long long>(k + nparts) * G;
    float2* q2 = base + static_cast<long long>(k + 2 * nparts) * G;
    float2* q3 = base + static_cast<long long>(k + 3 * nparts) * G;
    const float2 v0 = __ldcg(q0), v1 = __ldcg(q1), v2 = __ldcg(q2), v3 = __ldcg(q3);
    *q0 = zero;
    *q1 = zero;
    *q2 = zero;
    *q3 = zero;
    a += (v0.x + v1.x) + (v2.x + v3.x);
    b += (v0.y + v1.y) + (v2.y + v3.y);
  }
  for (; k < k1; k += nparts) {
    float2* q0 = base + static_cast<long long>(k) * G;
    const float2 v0 = __ldcg(q0);
    *q0 = zero;
    a += v0.x;
    b += v0.y;
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x < G) {
    a = 0.f;
    b = 0.f;
    for (int i = 0; i < nparts; ++i) {
      a += s_a[i * G + threadIdx.x];
      b += s_b[i * G + threadIdx.x];
    }
    float* o = partial + ((static_cast<long long>(s) * kGnFoldParts + c) * G + threadIdx.x) * 2;
    o[0] = a;
    o[1] = b;
  }
}
}  // namespace lvd

extern "C" int lvd_groupnorm_apply(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t samples,
                                   int64_t rows_per_sample, int32_t channels, int32_t groups, const void* gamma,
                                   const void* beta, float eps, int32_t silu, void* gn_partial, int32_t slots,
                                   void* workspace, void* stats_out, void* stream) {
  LVD_REQUIRE(x && y && gamma && beta && gn_partial && workspace && stats_out, "lvd_groupnorm_apply: null pointer");
  LVD_REQUIRE(channels % 8 == 0 && channels % groups == 0 && groups <= 64, "lvd_groupnorm_apply: C=%d G=%d unsupported",
              channels, groups);
  LVD_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0 && samples >= 1 && rows_per_sample >= 1 && slots >= 1,
              "lvd_groupnorm_apply: bad shape");
  const int threads = gn_threads(channels);
  const GnApplyPlan ap = gn_apply_plan(channels, threads, samples, rows_per_sample);
  const int T = ap.T;
  const long long schunks = ap.chunks;
  const size_t ring_bytes = ap.ring_bytes;
  LVD_REQUIRE(threads <= 320 && ring_bytes <= 220 * 1024, "lvd_groupnorm_apply: C=%d too wide for the streaming pass",
              channels);
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 sgrid(static_cast<unsigned>(schunks), static_cast<unsigned>(samples));
  gn_tiles_fold_kernel<<<dim3(kGnFoldParts, static_cast<unsigned>(samples)), 1024 / groups * groups, 0, st>>>(
      static_cast<float*>(gn_partial), slots, groups, static_cast<float*>(workspace));
  (silu ? gn_stream_apply_kernel<true> : gn_stream_apply_kernel<false>)<<<sgrid, threads, ring_bytes, st>>>(
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows_per_sample, channels, groups,
      static_cast<int>(schunks), kGnFoldParts, T, static_cast<const float*>(workspace),
      static_cast<float>(rows_per_sample) * (channels / groups), eps, static_cast<float*>(stats_out),
      static_cast<const __half*>(gamma), static_cast<const __half*>(beta));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

namespace lvd {
// {sum, sum of squares} per (slab, row) left by gemm_tc_kernel's epilogue -> {mean, rstd} per row (fixed order).
__global__ void __launch_bounds__(256) ln_rowstats_kernel(const float2* __restrict__ part, int slabs, long long rows,
                                                          float inv_c, float eps, float2* __restrict__ stats) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float a = 0.f, b = 0.f;
  int k = 0;
  for (; k + 3 < slabs; k += 4) {
    const float2 v0 = __ldcg(part + k * rows + r), v1 = __ldcg(part + (k + 1) * rows + r);
    const float2 v2 = __ldcg(part + (k + 2) * rows + r), v3 = __ldcg(part + (k + 3) * rows + r);
    a += (v0.x + v1.x) + (v2.x + v3.x);
    b += (v0.y + v1.y) + (v2.y + v3.y);
  }
  for (; k < slabs; ++k) {
    const float2 v0 = __ldcg(part + k * rows + r);
    a += v0.x;
    b += v0.y;
  }
  const float mean = a * inv_c;
  float var = b * inv_c - mean * mean;
  var = var > 0.f ? var : 0.f;
  stats[r] = make_float2(mean, rsqrtf(var + eps));
}
}  // namespace lvd

extern "C" int lvd_ln_rowstats(const void* lnstat_partial, int32_t slabs, int64_t rows, int32_t channels, float eps,
                               void* stats_f32, void* stream) {
  LVD_REQUIRE(lnstat_partial && stats_f32 && slabs >= 1 && channels == slabs * 32, "lvd_ln_rowstats: bad arguments");
  if (rows <= 0) return LVD_OK;
  ln_rowstats_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(lnstat_partial), slabs, rows, 1.0f / channels, eps, static_cast<float2*>(stats_f32));
  g_launches.fetch_add(1, std::memory_order_relaxed);
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
  const int vpr = channels / 8;
  const int lanes = vpr <= 8 * kLnMaxVec ? 8 : (vpr <= 16 * kLnMaxVec ? 16 : 32);  // lanes per row
  const int nv = (vpr + lanes - 1) / lanes;
  // R row sets in flight per warp; small problems (fewer than ~2 warps per scheduler slot) take R = 1 to fill the SMs
  int rsets = nv <= 2 ? 4 : (nv <= 5 ? 3 : 2);
  if (rows / ((32 / lanes) * rsets) < 16LL * num_sms()) rsets = 1;
  const int rows_per_warp = (32 / lanes) * rsets;
  long long blocks = (rows + warps * rows_per_warp - 1) / (warps * rows_per_warp);
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
#define LVD_LN_LAUNCH(NV, L, R)                                                                                   \
  layernorm_kernel<NV, L, R><<<static_cast<unsigned>(blocks), warps * 32, 2 * channels * sizeof(__half),         \
                               static_cast<cudaStream_t>(stream)>>>(                                             \
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows, channels,                        \
      static_cast<const __half*>(gamma), static_cast<const __half*>(beta), eps)
#define LVD_LN_LANES(L)                                          \
  if (nv <= 2) {                                                 \
    if (rsets == 1) LVD_LN_LAUNCH(2, L, 1); else LVD_LN_LAUNCH(2, L, 4); \
  } else if (nv <= 5) {                                          \
    if (rsets == 1) LVD_LN_LAUNCH(5, L, 1); else LVD_LN_LAUNCH(5, L, 3); \
  } else {                                                       \
    if (rsets == 1) LVD_LN_LAUNCH(8, L, 1); else LVD_LN_LAUNCH(8, L, 2); \
  }
  if (lanes == 8) {
    LVD_LN_LANES(8)
  } else if (lanes == 16) {
    LVD_LN_LANES(16)
  } else {
    LVD_LN_LANES(32)
  }
#undef LVD_LN_LANES
#undef LVD_LN_LAUNCH
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
