// backward.cu -- hand-written input-gradient ("dgrad") kernels for the guidance path: no autograd runs anywhere.
// Only d(energy)/d(latents) is needed (models/pipelines.py:120 asks autograd for [latents] only), so weight
// gradients do not exist here.  Linear / Conv dgrads run on the tensor-core kernel in gemm.cu (MN-major weights);
// this file holds the HBM-bound pieces: GroupNorm(+SiLU), LayerNorm, GEGLU, nearest-upsample, conv_in adjoints,
// the small attention adjoints that are not worth a tensor-core tile (temporal S=F, cross S_k=77) and helpers.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"
#include "norm_stream.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

__device__ __forceinline__ void ld8b(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack8b(const uint4& u, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void opaque_b(uint4& u) { asm volatile("" : "+r"(u.x), "+r"(u.y), "+r"(u.z), "+r"(u.w)); }
__device__ __forceinline__ void st8b(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
static unsigned bgrid(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 32LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

// ------------------------------------------------------------------------------------------------ GroupNorm backward
// y = xhat*gamma + beta, z = silu(y) (optional).  g = dz * silu'(y) * gamma;
// dx = rstd * (g - mean_grp(g) - xhat * mean_grp(g * xhat)).
// Pass A accumulates per-(sample, chunk, group) {sum g, sum g*xhat}; pass B applies.  Forward statistics come from
// the forward stats kernel's partials (fwd_partial: {sum x, sum x^2}).
__device__ __forceinline__ void gn_finalize_fwd(const float* fwd_partial, int s, int chunks, int G, int g, float n,
                                                float eps, float& mean, float& rstd) {
  float a = 0.f, b = 0.f;
  for (int k = 0; k < chunks; ++k) {
    const float* p = fwd_partial + ((static_cast<long long>(s) * chunks + k) * G + g) * 2;
    a += p[0];
    b += p[1];
  }
  mean = a / n;
  float var = b / n - mean * mean;
  var = var > 0.f ? var : 0.f;
  rstd = rsqrtf(var + eps);
}
__device__ __forceinline__ float silu_grad(float y) {
  const float s = __fdividef(1.f, 1.f + __expf(-y));
  return s * fmaf(y, 1.f - s, 1.f);
}

__global__ void gn_bwd_stats_kernel(const __half* __restrict__ x, long long ld_x, const __half* __restrict__ dz,
                                    long long ld_dz, long long rows_per_sample, int C, int G, int chunks,
                                    const float* __restrict__ stats, const __half* __restrict__ gamma,
                                    const __half* __restrict__ beta, float eps, int silu,
                                    float* __restrict__ bwd_partial) {
  extern __shared__ float s_part[];  // [rows per pass][2*C] partial sums (folded in a fixed order), then [2*G] mean/rstd
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  float* s_mean = s_part + static_cast<long long>(rpp) * 2 * C;
  float* s_rstd = s_mean + G;
  const int v = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_mean[g] = stats[(static_cast<long long>(s) * G + g) * 2];
    s_rstd[g] = stats[(static_cast<long long>(s) * G + g) * 2 + 1];
  }
  __syncthreads();
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  if (rsub < rpp) {
    float ga[8], be[8], mu[8], rs[8], a1[8], a2[8];
    ld8b(gamma + v * 8, ga);
    ld8b(beta + v * 8, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (v * 8 + i) / cpg;
      mu[i] = s_mean[g];
      rs[i] = s_rstd[g];
      a1[i] = a2[i] = 0.f;
    }
    const __half* xb = x + (s * rows_per_sample) * ld_x + v * 8;
    const __half* db = dz + (s * rows_per_sample) * ld_dz + v * 8;
    long long r = r0 + rsub;
    for (; r + rpp < r1; r += 2LL * rpp) {
      float t[8], d[8], t2[8], d2[8];
      ld8b(xb + r * ld_x, t);
      ld8b(db + r * ld_dz, d);
      ld8b(xb + (r + rpp) * ld_x, t2);
      ld8b(db + (r + rpp) * ld_dz, d2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (t[i] - mu[i]) * rs[i], xh2 = (t2[i] - mu[i]) * rs[i];
        float gg = d[i], gg2 = d2[i];
        if (silu) {
          gg *= silu_grad(xh * ga[i] + be[i]);
          gg2 *= silu_grad(xh2 * ga[i] + be[i]);
        }
        gg *= ga[i];
        gg2 *= ga[i];
        a1[i] += gg + gg2;
        a2[i] += gg * xh + gg2 * xh2;
      }
    }
    for (; r < r1; r += rpp) {
      float t[8], d[8];
      ld8b(xb + r * ld_x, t);
      ld8b(db + r * ld_dz, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (t[i] - mu[i]) * rs[i];
        float gg = d[i];
        if (silu) gg *= silu_grad(xh * ga[i] + be[i]);
        gg *= ga[i];
        a1[i] += gg;
        a2[i] += gg * xh;
      }
    }
    float* mine = s_part + static_cast<long long>(rsub) * 2 * C + v * 8;  // no atomics: run-to-run reproducible
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = a1[i];
      mine[C + i] = a2[i];
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rs = 0; rs < rpp; ++rs) {
      const float* row = s_part + static_cast<long long>(rs) * 2 * C;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        a += row[c];
        b += row[C + c];
      }
    }
    float* o = bwd_partial + ((static_cast<long long>(s) * chunks + chunk) * G + g) * 2;
    o[0] = a;
    o[1] = b;
  }
}

__global__ void gn_bwd_apply_kernel(const __half* __restrict__ x, long long ld_x, const __half* __restrict__ dz,
                                    long long ld_dz, __half* __restrict__ dx, long long ld_dx,
                                    long long rows_per_sample, int C, int G, int chunks,
                                    const float* __restrict__ stats, const float* __restrict__ bstats,
                                    const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                                    int silu, int accumulate) {
  __shared__ float s_mean[64], s_rstd[64], s_m1[64], s_m2[64];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_mean[g] = stats[(static_cast<long long>(s) * G + g) * 2];
    s_rstd[g] = stats[(static_cast<long long>(s) * G + g) * 2 + 1];
    s_m1[g] = bstats[(static_cast<long long>(s) * G + g) * 2];
    s_m2[g] = bstats[(static_cast<long long>(s) * G + g) * 2 + 1];
  }
  __syncthreads();
  if (rsub >= rpp) return;
  float ga[8], be[8], mu[8], rs[8], m1[8], m2[8];
  ld8b(gamma + v * 8, ga);
  ld8b(beta + v * 8, be);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (v * 8 + i) / cpg;
    mu[i] = s_mean[g];
    rs[i] = s_rstd[g];
    m1[i] = s_m1[g];
    m2[i] = s_m2[g];
  }
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  const __half* xb = x + (s * rows_per_sample) * ld_x + v * 8;
  const __half* db = dz + (s * rows_per_sample) * ld_dz + v * 8;
  __half* ob = dx + (s * rows_per_sample) * ld_dx + v * 8;
  for (long long r = r0 + rsub; r < r1; r += rpp) {
    float t[8], d[8], o[8];
    ld8b(xb + r * ld_x, t);
    ld8b(db + r * ld_dz, d);
    if (accumulate) ld8b(ob + r * ld_dx, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (t[i] - mu[i]) * rs[i];
      float gg = d[i];
      if (silu) gg *= silu_grad(xh * ga[i] + be[i]);
      gg *= ga[i];
      const float res = rs[i] * (gg - m1[i] - xh * m2[i]);
      o[i] = accumulate ? o[i] + res : res;
    }
    st8b(ob + r * ld_dx, o);
  }
}

// ---- streaming versions (same ring as the forward passes, norm_stream.cuh): x and dz travel through the ring in lock
// step; pass A's last CTA per sample (ticket) folds the chunk partials into {mean g, mean g*xhat}.
__device__ unsigned int g_gnb_tickets[8192];

__global__ void __launch_bounds__(320, 2) gn_stream_bwd_stats_kernel(
    const __half* __restrict__ x, long long ld_x, const __half* __restrict__ dz, long long ld_dz,
    long long rows_per_sample, int C, int G, int chunks, int T, const float* __restrict__ stats,
    const __half* __restrict__ gamma, const __half* __restrict__ beta, int silu, float* __restrict__ bwd_partial,
    float n, float* __restrict__ bstats) {
  extern __shared__ uint4 s_dyn[];
  __shared__ uint64_t s_full[kGnStages];
  __shared__ int s_last;
  __shared__ float s_mean[64], s_rstd[64];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const int cpg = C / G;
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  GnStream st;
  st.full = s_full;
  st.ring = s_dyn;
  st.src = x + (s * rows_per_sample + r0) * ld_x;
  st.ld = ld_x;
  st.src2 = dz + (s * rows_per_sample + r0) * ld_dz;
  st.ld2 = ld_dz;
  st.n = r1 > r0 ? static_cast<int>(r1 - r0) : 0;
  st.T = T;
  st.vpr = vpr;
  st.nit = (st.n + T - 1) / T;
  st.dense = ld_x == C;
  float* s_part = reinterpret_cast<float*>(s_dyn + static_cast<long long>(kGnStages) * 2 * T * vpr);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGnStages; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
  }
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_mean[g] = stats[(static_cast<long long>(s) * G + g) * 2];
    s_rstd[g] = stats[(static_cast<long long>(s) * G + g) * 2 + 1];
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int it = 0; it < kGnStages && it < st.nit; ++it) st.issue(it);
  float ga[8], be[8], mu[8], rs[8], a1[8], a2[8];
  if (rsub < rpp) {
    ld8b(gamma + v * 8, ga);
    ld8b(beta + v * 8, be);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = rsub < rpp ? (v * 8 + i) / cpg : 0;
    mu[i] = s_mean[g];
    rs[i] = s_rstd[g];
    a1[i] = a2[i] = 0.f;
  }
  for (int it = 0; it < st.nit; ++it) {
    const int stage = it % kGnStages;
    mbar_wait(&s_full[stage], static_cast<uint32_t>(it / kGnStages) & 1u);
    const int rows = st.n - it * T < T ? st.n - it * T : T;
    const uint4* bx = s_dyn + static_cast<long long>(stage) * 2 * T * vpr;
    const uint4* bd = bx + static_cast<long long>(T) * vpr;
    if (rsub < rpp) {
      for (int r = rsub; r < rows; r += rpp) {
        float t[8], d[8];
        unpack8b(bx[static_cast<long long>(r) * vpr + v], t);
        unpack8b(bd[static_cast<long long>(r) * vpr + v], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (t[i] - mu[i]) * rs[i];
          float gg = d[i];
          if (silu) gg *= silu_grad(xh * ga[i] + be[i]);
          gg *= ga[i];
          a1[i] += gg;
          a2[i] += gg * xh;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 32 && it + kGnStages < st.nit) st.issue(it + kGnStages);
  }
  if (rsub < rpp) {
    float* mine = s_part + static_cast<long long>(rsub) * 2 * C + v * 8;  // no atomics: run-to-run reproducible
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[i] = a1[i];
      mine[C + i] = a2[i];
    }
  }
  __syncthreads();
  float* s_col = s_part + static_cast<long long>(rpp) * 2 * C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int q = 0; q < rpp; ++q) a += s_part[static_cast<long long>(q) * 2 * C + c];
    s_col[c] = a;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * G; t += blockDim.x) {
    const int g = t >> 1, which = t & 1;
    const float* col = s_col + which * C + g * cpg;
    float a = 0.f;
    for (int c = 0; c < cpg; ++c) a += col[c];
    bwd_partial[((static_cast<long long>(s) * chunks + chunk) * G + g) * 2 + which] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&g_gnb_tickets[s], 1u) == static_cast<unsigned>(chunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* s_a = s_part;
  float* s_b = s_part + blockDim.x;
  const int nparts = blockDim.x / G;
  {
    const int g = threadIdx.x % G, part = threadIdx.x / G;
    float a = 0.f, b = 0.f;
    if (part < nparts)
      for (int k = part; k < chunks; k += nparts) {
        const float* p = bwd_partial + ((static_cast<long long>(s) * chunks + k) * G + g) * 2;
        a += __ldcg(p);
        b += __ldcg(p + 1);
      }
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
    bstats[(static_cast<long long>(s) * G + threadIdx.x) * 2] = a / n;
    bstats[(static_cast<long long>(s) * G + threadIdx.x) * 2 + 1] = b / n;
  }
  if (threadIdx.x == 0) g_gnb_tickets[s] = 0u;
}

__global__ void __launch_bounds__(320, 2) gn_stream_bwd_apply_kernel(
    const __half* __restrict__ x, long long ld_x, const __half* __restrict__ dz, long long ld_dz,
    __half* __restrict__ dx, long long ld_dx, long long rows_per_sample, int C, int G, int chunks, int T,
    const float* __restrict__ stats, const float* __restrict__ bstats, const __half* __restrict__ gamma,
    const __half* __restrict__ beta, int silu, int accumulate) {
  extern __shared__ uint4 s_dyn[];
  __shared__ uint64_t s_full[kGnStages];
  __shared__ float s_mean[64], s_rstd[64], s_m1[64], s_m2[64];
  const int vpr = C >> 3;
  const int rpp = blockDim.x / vpr;
  const int v = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
  const int chunk = blockIdx.x, s = blockIdx.y;
  const int cpg = C / G;
  const long long rows_per_chunk = (rows_per_sample + chunks - 1) / chunks;
  const long long r0 = chunk * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows_per_sample) r1 = rows_per_sample;
  GnStream st;
  st.full = s_full;
  st.ring = s_dyn;
  st.src = x + (s * rows_per_sample + r0) * ld_x;
  st.ld = ld_x;
  st.src2 = dz + (s * rows_per_sample + r0) * ld_dz;
  st.ld2 = ld_dz;
  st.n = r1 > r0 ? static_cast<int>(r1 - r0) : 0;
  st.T = T;
  st.vpr = vpr;
  st.nit = (st.n + T - 1) / T;
  st.dense = ld_x == C;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGnStages; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
  }
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_mean[g] = stats[(static_cast<long long>(s) * G + g) * 2];
    s_rstd[g] = stats[(static_cast<long long>(s) * G + g) * 2 + 1];
    s_m1[g] = bstats[(static_cast<long long>(s) * G + g) * 2];
    s_m2[g] = bstats[(static_cast<long long>(s) * G + g) * 2 + 1];
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int it = 0; it < kGnStages && it < st.nit; ++it) st.issue(it);
  float ga[8], be[8], mu[8], rs[8], m1[8], m2[8];
  if (rsub < rpp) {
    ld8b(gamma + v * 8, ga);
    ld8b(beta + v * 8, be);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = rsub < rpp ? (v * 8 + i) / cpg : 0;
    mu[i] = s_mean[g];
    rs[i] = s_rstd[g];
    m1[i] = s_m1[g];
    m2[i] = s_m2[g];
  }
  __half* ob = dx + (s * rows_per_sample + r0) * ld_dx + v * 8;
  for (int it = 0; it < st.nit; ++it) {
    const int stage = it % kGnStages;
    mbar_wait(&s_full[stage], static_cast<uint32_t>(it / kGnStages) & 1u);
    const int rows = st.n - it * T < T ? st.n - it * T : T;
    const uint4* bx = s_dyn + static_cast<long long>(stage) * 2 * T * vpr;
    const uint4* bd = bx + static_cast<long long>(T) * vpr;
    if (rsub < rpp) {
      for (int r = rsub; r < rows; r += rpp) {
        float t[8], d[8], o[8];
        __half* op = ob + (static_cast<long long>(it) * T + r) * ld_dx;
        if (accumulate) ld8b(op, o);
        unpack8b(bx[static_cast<long long>(r) * vpr + v], t);
        unpack8b(bd[static_cast<long long>(r) * vpr + v], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (t[i] - mu[i]) * rs[i];
          float gg = d[i];
          if (silu) gg *= silu_grad(xh * ga[i] + be[i]);
          gg *= ga[i];
          const float res = rs[i] * (gg - m1[i] - xh * m2[i]);
          o[i] = accumulate ? o[i] + res : res;
        }
        st8b(op, o);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32 && it + kGnStages < st.nit) st.issue(it + kGnStages);
  }
}

// forward statistics kernel lives in norm.cu
__global__ void gn_stats_kernel(const __half* __restrict__ x, long long ld, long long rows_per_sample, int C, int G,
                                int chunks, float* __restrict__ partial, float n, float eps, float* __restrict__ stats);
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, int G, float n, float eps,
                                   float* __restrict__ stats, int raw_means);

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.  Same thread layout as the forward kernel
// (norm.cu): L lanes per row, 32 / L rows side by side in a warp, R row sets in flight, x and dy held packed (fp16) in
// registers so that every load of the rows in flight is issued before the first dependent shuffle.
constexpr int kLnbMaxVec = 8;

template <int NV, int L, int R>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const __half* __restrict__ x, long long ld_x,
                                                               const __half* __restrict__ dy, long long ld_dy,
                                                               __half* __restrict__ dx, long long ld_dx,
                                                               long long rows, int C, const __half* __restrict__ gamma,
                                                               float eps, int accumulate) {
  extern __shared__ uint4 s_gamma_v[];  // gamma[C] (fp16), loaded once per block
  {
    __half* sg = reinterpret_cast<__half*>(s_gamma_v);
    for (int i = threadIdx.x; i < C; i += blockDim.x) sg[i] = gamma[i];
  }
  __syncthreads();
  constexpr int kRowsPerSet = 32 / L;
  const int lane = threadIdx.x & 31;
  const int j = lane % L, sub = lane / L;
  const int vpr = C >> 3;
  const float inv_c = 1.0f / static_cast<float>(C);
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const long long rows_per_iter = static_cast<long long>(kRowsPerSet) * R;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rows_per_iter;
       row0 < rows; row0 += warps_total * rows_per_iter) {
    uint4 xv[R][NV], dv[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r * kRowsPerSet + sub;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int v = j + k * L;
        const bool ok = row < rows && v < vpr;
        xv[r][k] = ok ? *reinterpret_cast<const uint4*>(x + row * ld_x + v * 8) : make_uint4(0, 0, 0, 0);
        dv[r][k] = ok ? *reinterpret_cast<const uint4*>(dy + row * ld_dy + v * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    float mean[R], rstd[R], s1[R], s2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float f[8];
        unpack8b(xv[r][k], f);
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
      for (int k = 0; k < NV; ++k) opaque_b(xv[r][k]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] *= inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (j + k * L < vpr) {
          float f[8];
          unpack8b(xv[r][k], f);
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
      for (int k = 0; k < NV; ++k) opaque_b(xv[r][k]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rstd[r] = rsqrtf(rstd[r] * inv_c + eps);
      s1[r] = 0.f;
      s2[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = j + k * L;
      if (v < vpr) {
        float ga[8];
        unpack8b(s_gamma_v[v], ga);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float f[8], d[8];
          unpack8b(xv[r][k], f);
          unpack8b(dv[r][k], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g = d[i] * ga[i];
            s1[r] += g;
            s2[r] += g * ((f[i] - mean[r]) * rstd[r]);
          }
        }
      }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
        s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        opaque_b(xv[r][k]);
        opaque_b(dv[r][k]);
      }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = j + k * L;
      if (v < vpr) {
        float ga[8];
        unpack8b(s_gamma_v[v], ga);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const long long row = row0 + r * kRowsPerSet + sub;
          if (row < rows) {
            float f[8], d[8], o[8];
            unpack8b(xv[r][k], f);
            unpack8b(dv[r][k], d);
            if (accumulate) ld8b(dx + row * ld_dx + v * 8, o);
            const float m1 = s1[r] * inv_c, m2 = s2[r] * inv_c;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xh = (f[i] - mean[r]) * rstd[r];
              const float res = rstd[r] * (d[i] * ga[i] - m1 - xh * m2);
              o[i] = accumulate ? o[i] + res : res;
            }
            st8b(dx + row * ld_dx + v * 8, o);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU (un-fused) fwd/bwd
// pre [rows, 2N] in the interleaved layout of repack.geglu_interleave: every 256-column block = [128 value | 128 gate].
constexpr int kGegluHalf = 128;
__global__ void geglu_fwd_kernel(const __half* __restrict__ pre, long long ld_pre, __half* __restrict__ out,
                                 long long ld_out, long long rows, int n_out) {
  const int vec = n_out / 8;
  const long long total = rows * vec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec;
    const int c = static_cast<int>(idx % vec) * 8;  // output column
    const int pc = (c / kGegluHalf) * (2 * kGegluHalf) + (c % kGegluHalf);
    float a[8], g[8];
    ld8b(pre + r * ld_pre + pc, a);
    ld8b(pre + r * ld_pre + pc + kGegluHalf, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = round_h(a[i] * round_h(gelu_erf_f(g[i])));
    st8b(out + r * ld_out + c, a);
  }
}
__global__ void geglu_bwd_kernel(const __half* __restrict__ pre, long long ld_pre, const __half* __restrict__ dout,
                                 long long ld_dout, __half* __restrict__ dpre, long long ld_dpre, long long rows,
                                 int n_out) {
  const int vec = n_out / 8;
  const long long total = rows * vec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec;
    const int c = static_cast<int>(idx % vec) * 8;
    const int pc = (c / kGegluHalf) * (2 * kGegluHalf) + (c % kGegluHalf);
    float a[8], g[8], d[8], da[8], dg[8];
    ld8b(pre + r * ld_pre + pc, a);
    ld8b(pre + r * ld_pre + pc + kGegluHalf, g);
    ld8b(dout + r * ld_dout + c, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = g[i];
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
      da[i] = d[i] * (x * cdf);
      dg[i] = d[i] * a[i] * (cdf + x * pdf);
    }
    st8b(dpre + r * ld_dpre + pc, da);
    st8b(dpre + r * ld_dpre + pc + kGegluHalf, dg);
  }
}

// ------------------------------------------------------------------------------------------------ small adjoints
// out = a + b (fp16 rows)
__global__ void add_rows_kernel(const __half* __restrict__ a, long long ld_a, const __half* __restrict__ b,
                                long long ld_b, __half* __restrict__ out, long long ld_o, long long rows, int C) {
  const int vec = C / 8;
  const long long total = rows * vec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vec;
    const int c = static_cast<int>(idx % vec) * 8;
    float x[8], y[8];
    ld8b(a + r * ld_a + c, x);
    ld8b(b + r * ld_b + c, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += y[i];
    st8b(out + r * ld_o + c, x);
  }
}
// adjoint of nearest 2x upsample: dx[n,y,x] = sum of the 2x2 block of dy
__global__ void upsample2x_bwd_kernel(const __half* __restrict__ dy, long long ld_dy, __half* __restrict__ dx,
                                      long long ld_dx, int n_img, int H, int W, int C) {
  const int vec = C / 8;
  const long long total = static_cast<long long>(n_img) * H * W * vec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % vec) * 8;
    long long p = idx / vec;
    const int x = static_cast<int>(p % W);
    p /= W;
    const int y = static_cast<int>(p % H);
    const long long n = p / H;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float t[8];
        ld8b(dy + ((n * 2 * H + 2 * y + a) * 2 * W + 2 * x + b) * ld_dy + c, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += t[i];
      }
    st8b(dx + ((n * H + y) * W + x) * ld_dx + c, acc);
  }
}
// adjoint of conv_in: d_sample (B, Cin, F, H, W) fp32 = sum_{co,ky,kx} d_rows[(y-ky+1, x-kx+1), co] * w[co, ci, ky, kx]
// one warp per latent pixel; lanes split the output channels of the forward conv.
__global__ void conv_in_bwd_kernel(const __half* __restrict__ drows, long long ld, const __half* __restrict__ w,
                                   float* __restrict__ dsample, int B, int Cin, int F, int H, int W, int Cout) {
  extern __shared__ __half s_wi[];  // [Cout][Cin][9]
  const int wn = Cout * Cin * 9;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) s_wi[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long npix = static_cast<long long>(B) * F * H * W;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); pix < npix;
       pix += warps_total) {
    long long p = pix;
    const int xx = static_cast<int>(p % W);
    p /= W;
    const int yy = static_cast<int>(p % H);
    p /= H;  // b*F + f
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int oy = yy - ky + 1;  // output pixel that read this input with tap ky
      if (oy < 0 || oy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ox = xx - kx + 1;
        if (ox < 0 || ox >= W) continue;
        const __half* dr = drows + ((p * H + oy) * W + ox) * ld;
        for (int co = lane; co < Cout; co += 32) {
          const float d = __half2float(dr[co]);
          for (int ci = 0; ci < Cin; ++ci) acc[ci] += d * __half2float(s_wi[(co * Cin + ci) * 9 + ky * 3 + kx]);
        }
      }
    }
    for (int ci = 0; ci < Cin; ++ci) {
      float a = acc[ci];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[ci] = a;
    }
    if (lane < Cin) {
      const int b = static_cast<int>(p / F), f = static_cast<int>(p % F);
      float a = 0.f;
      for (int ci = 0; ci < Cin; ++ci)
        if (ci == lane) a = acc[ci];
      dsample[(((static_cast<long long>(b) * Cin + lane) * F + f) * H + yy) * W + xx] = a;
    }
  }
}
// latents_out = latents - coef * d_sample (fp32 gradient, un-scaled by coef = sqrt(1-abar_t) / grad_scale), fp16 result
__global__ void latent_update_kernel(const __half* __restrict__ lat, const float* __restrict__ grad,
                                     __half* __restrict__ out, __half* __restrict__ grad_h, long long n, float coef,
                                     float unscale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gh = round_h(grad[i] * unscale);  // grad_cond in the model dtype (models/pipelines.py:120)
    if (grad_h) grad_h[i] = __float2half_rn(gh);
    out[i] = __float2half_rn(__half2float(lat[i]) - round_h(coef * gh));
  }
}
// D[row, head] = sum_d O[row, head*64+d] * dO[row, head*64+d]   (softmax-backward row term)
__global__ void attn_rowdot_kernel(const __half* __restrict__ o, long long ld_o, const __half* __restrict__ d_o,
                                   long long ld_do, float* __restrict__ out, long long rows, int heads) {
  const long long total = rows * heads;
  const long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3;  // 8 lanes per (row, head)
  const int sub = threadIdx.x & 7;
  if (w >= total) return;
  const long long r = w / heads;
  const int h = static_cast<int>(w % heads);
  float a[8], b[8];
  ld8b(o + r * ld_o + h * 64 + sub * 8, a);
  ld8b(d_o + r * ld_do + h * 64 + sub * 8, b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] * b[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (sub == 0) out[w] = s;
}

}  // namespace lvd

using namespace lvd;
#define ST(s) static_cast<cudaStream_t>(s)
#define HC(p) static_cast<const __half*>(p)
#define HM(p) static_cast<__half*>(p)

extern "C" int lvd_groupnorm_bwd(const void* x, int64_t ld_x, const void* dz, int64_t ld_dz, void* dx, int64_t ld_dx,
                                 int32_t samples, int64_t rows_per_sample, int32_t channels, int32_t groups,
                                 const void* gamma, const void* beta, float eps, int32_t silu, int32_t accumulate,
                                 const void* fwd_stats, void* workspace, int64_t workspace_bytes, void* stream) {
  LVD_REQUIRE(x && dz && dx && gamma && beta && workspace, "lvd_groupnorm_bwd: null pointer");
  LVD_REQUIRE(channels % 8 == 0 && channels % groups == 0 && groups <= 64 && channels <= 4096,
              "lvd_groupnorm_bwd: C=%d G=%d unsupported", channels, groups);
  LVD_REQUIRE(ld_x % 8 == 0 && ld_dz % 8 == 0 && ld_dx % 8 == 0, "lvd_groupnorm_bwd: ld must be a multiple of 8");
  long long chunks = (8LL * num_sms() + samples - 1) / samples;
  const long long max_chunks = (rows_per_sample + 15) / 16;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  const long long one = static_cast<long long>(samples) * chunks * groups * 2 * sizeof(float);
  const long long small = static_cast<long long>(samples) * groups * 2 * sizeof(float);
  LVD_REQUIRE(workspace_bytes >= 2 * one + 2 * small, "lvd_groupnorm_bwd: workspace too small");
  float* fwd = static_cast<float*>(workspace);
  float* bwd = fwd + one / sizeof(float);
  float* stats = bwd + one / sizeof(float);
  float* bstats = stats + small / sizeof(float);
  const int vpr = channels / 8;
  int rpp = 256 / vpr;
  if (rpp < 1) rpp = 1;
  const int threads = rpp * vpr;
  LVD_REQUIRE(threads <= 1024, "lvd_groupnorm_bwd: C too wide");
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(samples));
  const float n = static_cast<float>(rows_per_sample) * (channels / groups);
  const int fin_threads = 256 / groups * groups;
  if (fwd_stats) {
    stats = const_cast<float*>(static_cast<const float*>(fwd_stats));  // {mean, rstd} kept from the forward launch
  } else {
    LVD_REQUIRE(samples <= 8192, "lvd_groupnorm_bwd: too many samples");
    gn_stats_kernel<<<grid, threads, (static_cast<size_t>(threads / (channels / 8)) + 1) * 2 * channels * sizeof(float),
                      ST(stream)>>>(
        HC(x), ld_x, rows_per_sample, channels, groups, static_cast<int>(chunks), fwd, n, eps, stats);
  }
  {
    const long long row_bytes = 2LL * channels;
    int T = static_cast<int>((8 * 1024) / row_bytes) / rpp * rpp;  // ~8 KB of x + ~8 KB of dz per stage
    if (T < rpp) T = rpp;
    long long schunks = (2LL * num_sms()) / samples;  // one wave of 2 CTAs per SM
    const long long most = (rows_per_sample + T - 1) / T;
    if (schunks > most) schunks = most;
    if (schunks > chunks) schunks = chunks;  // the workspace is sized for `chunks`
    if (schunks < 1) schunks = 1;
    const size_t ring_bytes = static_cast<size_t>(kGnStages) * 2 * T * row_bytes;
    const size_t scratch = (static_cast<size_t>(rpp) + 1) * 2 * channels * sizeof(float);
    if (threads <= 320 && ring_bytes + scratch <= 225 * 1024 && samples <= 8192) {
      static bool attr = false;
      if (!attr) {
        LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_bwd_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            225 * 1024));
        LVD_CHECK_CUDA(cudaFuncSetAttribute(gn_stream_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            225 * 1024));
        attr = true;
      }
      dim3 sgrid(static_cast<unsigned>(schunks), static_cast<unsigned>(samples));
      gn_stream_bwd_stats_kernel<<<sgrid, threads, ring_bytes + scratch, ST(stream)>>>(
          HC(x), ld_x, HC(dz), ld_dz, rows_per_sample, channels, groups, static_cast<int>(schunks), T, stats,
          HC(gamma), HC(beta), silu, bwd, n, bstats);
      gn_stream_bwd_apply_kernel<<<sgrid, threads, ring_bytes, ST(stream)>>>(
          HC(x), ld_x, HC(dz), ld_dz, HM(dx), ld_dx, rows_per_sample, channels, groups, static_cast<int>(schunks), T,
          stats, bstats, HC(gamma), HC(beta), silu, accumulate);
    } else {
    gn_bwd_stats_kernel<<<grid, threads,
                          (static_cast<size_t>(threads / (channels / 8)) * 2 * channels + 2 * groups) * sizeof(float),
                          ST(stream)>>>(
        HC(x), ld_x, HC(dz), ld_dz, rows_per_sample, channels, groups, static_cast<int>(chunks), stats, HC(gamma),
        HC(beta), eps, silu, bwd);
    gn_finalize_kernel<<<samples, fin_threads, 0, ST(stream)>>>(bwd, static_cast<int>(chunks), groups, n, eps, bstats, 1);
    gn_bwd_apply_kernel<<<grid, threads, 0, ST(stream)>>>(HC(x), ld_x, HC(dz), ld_dz, HM(dx), ld_dx, rows_per_sample,
                                                          channels, groups, static_cast<int>(chunks), stats, bstats,
                                                          HC(gamma), HC(beta), eps, silu, accumulate);
    }
  }
  g_launches.fetch_add(5, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_layernorm_bwd(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx,
                                 int64_t rows, int32_t channels, const void* gamma, float eps, int32_t accumulate,
                                 void* stream) {
  LVD_REQUIRE(x && dy && dx && gamma, "lvd_layernorm_bwd: null pointer");
  LVD_REQUIRE(channels % 8 == 0 && channels <= kLnbMaxVec * 256, "lvd_layernorm_bwd: C=%d unsupported", channels);
  LVD_REQUIRE(ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0, "lvd_layernorm_bwd: ld must be a multiple of 8");
  if (rows <= 0) return LVD_OK;
  const int warps = 8;
  const int vpr = channels / 8;
  const int lanes = vpr <= 8 * kLnbMaxVec ? 8 : (vpr <= 16 * kLnbMaxVec ? 16 : 32);  // lanes per row
  const int nv = (vpr + lanes - 1) / lanes;
  int rsets = nv <= 5 ? 2 : 1;
  if (rows / ((32 / lanes) * rsets) < 16LL * num_sms()) rsets = 1;
  const int rows_per_warp = (32 / lanes) * rsets;
  long long blocks = (rows + warps * rows_per_warp - 1) / (warps * rows_per_warp);
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
#define LVD_LNB_LAUNCH(NV, L, R)                                                                              \
  layernorm_bwd_kernel<NV, L, R><<<static_cast<unsigned>(blocks), warps * 32, channels * sizeof(__half),     \
                                   ST(stream)>>>(                                                             \
      HC(x), ld_x, HC(dy), ld_dy, HM(dx), ld_dx, rows, channels, HC(gamma), eps, accumulate)
#define LVD_LNB_LANES(L)                                                   \
  if (nv <= 2) {                                                           \
    if (rsets == 1) LVD_LNB_LAUNCH(2, L, 1); else LVD_LNB_LAUNCH(2, L, 2); \
  } else if (nv <= 5) {                                                    \
    if (rsets == 1) LVD_LNB_LAUNCH(5, L, 1); else LVD_LNB_LAUNCH(5, L, 2); \
  } else {                                                                 \
    LVD_LNB_LAUNCH(8, L, 1);                                               \
  }
  if (lanes == 8) {
    LVD_LNB_LANES(8)
  } else if (lanes == 16) {
    LVD_LNB_LANES(16)
  } else {
    LVD_LNB_LANES(32)
  }
#undef LVD_LNB_LANES
#undef LVD_LNB_LAUNCH
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_geglu(const void* pre, int64_t ld_pre, void* out, int64_t ld_out, int64_t rows, int32_t n_out,
                         void* stream) {
  LVD_REQUIRE(pre && out && n_out % kGegluHalf == 0 && ld_pre % 8 == 0 && ld_out % 8 == 0, "lvd_geglu: bad arguments");
  geglu_fwd_kernel<<<bgrid(rows * (n_out / 8), 256), 256, 0, ST(stream)>>>(HC(pre), ld_pre, HM(out), ld_out, rows,
                                                                          n_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_geglu_bwd(const void* pre, int64_t ld_pre, const void* dout, int64_t ld_dout, void* dpre,
                             int64_t ld_dpre, int64_t rows, int32_t n_out, void* stream) {
  LVD_REQUIRE(pre && dout && dpre && n_out % kGegluHalf == 0, "lvd_geglu_bwd: bad arguments");
  LVD_REQUIRE(ld_pre % 8 == 0 && ld_dout % 8 == 0 && ld_dpre % 8 == 0, "lvd_geglu_bwd: ld must be a multiple of 8");
  geglu_bwd_kernel<<<bgrid(rows * (n_out / 8), 256), 256, 0, ST(stream)>>>(HC(pre), ld_pre, HC(dout), ld_dout,
                                                                          HM(dpre), ld_dpre, rows, n_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_add_rows(const void* a, int64_t ld_a, const void* b, int64_t ld_b, void* out, int64_t ld_out,
                            int64_t rows, int32_t channels, void* stream) {
  LVD_REQUIRE(a && b && out && channels % 8 == 0 && ld_a % 8 == 0 && ld_b % 8 == 0 && ld_out % 8 == 0,
              "lvd_add_rows: bad arguments");
  add_rows_kernel<<<bgrid(rows * (channels / 8), 256), 256, 0, ST(stream)>>>(HC(a), ld_a, HC(b), ld_b, HM(out), ld_out,
                                                                            rows, channels);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_upsample2x_bwd(const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int32_t n_img, int32_t height,
                                  int32_t width, int32_t channels, void* stream) {
  LVD_REQUIRE(dy && dx && channels % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0, "lvd_upsample2x_bwd: bad arguments");
  const long long total = static_cast<long long>(n_img) * height * width * (channels / 8);
  upsample2x_bwd_kernel<<<bgrid(total, 256), 256, 0, ST(stream)>>>(HC(dy), ld_dy, HM(dx), ld_dx, n_img, height, width,
                                                                  channels);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_conv_in_bwd(const void* drows, int64_t ld, const void* weight, void* dsample_f32, int32_t batch,
                               int32_t cin, int32_t frames, int32_t height, int32_t width, int32_t cout, void* stream) {
  LVD_REQUIRE(drows && weight && dsample_f32 && cin >= 1 && cin <= 8, "lvd_conv_in_bwd: bad arguments");
  const size_t smem = static_cast<size_t>(cout) * cin * 9 * sizeof(__half);
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(conv_in_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  const long long npix = static_cast<long long>(batch) * frames * height * width;
  conv_in_bwd_kernel<<<bgrid(npix * 32, 256), 256, smem, ST(stream)>>>(HC(drows), ld, HC(weight),
                                                                      static_cast<float*>(dsample_f32), batch, cin,
                                                                      frames, height, width, cout);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_latent_update(const void* latents, const void* grad_f32, void* out, void* grad_h_out, int64_t n,
                                 float coef, float unscale, void* stream) {
  LVD_REQUIRE(latents && grad_f32 && out, "lvd_latent_update: null pointer");
  latent_update_kernel<<<bgrid(n, 256), 256, 0, ST(stream)>>>(HC(latents), static_cast<const float*>(grad_f32), HM(out),
                                                             HM(grad_h_out), n, coef, unscale);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
extern "C" int lvd_attn_rowdot(const void* o, int64_t ld_o, const void* d_o, int64_t ld_do, void* out_f32, int64_t rows,
                               int32_t heads, void* stream) {
  LVD_REQUIRE(o && d_o && out_f32 && ld_o % 8 == 0 && ld_do % 8 == 0, "lvd_attn_rowdot: bad arguments");
  const long long threads = rows * heads * 8;
  attn_rowdot_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ST(stream)>>>(
      HC(o), ld_o, HC(d_o), ld_do, static_cast<float*>(out_f32), rows, heads);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
