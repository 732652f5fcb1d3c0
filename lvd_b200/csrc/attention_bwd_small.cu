// attention_bwd_small.cu -- attention adjoints that are too small for a tensor-core tile:
//   temporal self-attention (sequence = F <= 32 frames, packed rows)  -> dQ, dK, dV
//   cross-attention over the 77 text tokens                            -> dQ only (K, V come from the text
//     encoder, which is no function of the latents), with the box-energy gradient dE/dP injected on the object-token
//     columns: this is where utils/guidance.py's loss enters the backward pass.
// Both recompute the probabilities from Q, K (nothing but Q/K/V/O is kept from the forward).  fp32 math.
#include <atomic>
#include <cstdlib>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;
int launch_cross_attn_bwd_mma(const void* q, long long ld_q, const void* kv, long long ld_kv, int k_col0, int v_col0,
                              const void* d_o, long long ld_do, const void* dp_tok, const int* slot_col, int n_slots,
                              void* dq, long long ld_dq, long long s_q, int s_k, int heads, int nz, int kv_z_div,
                              float scale, cudaStream_t stream, const int* slot_col_dev = nullptr);  // attention_temporal.cu

constexpr int kPad = 66;  // padded row length (halves) of 64-wide tiles in shared memory: conflict-free column reads

__device__ __forceinline__ void load_rows_to_smem(__half* dst, const __half* src, long long ld, int nrows, int lane) {
  // nrows x 64 halves; 8 lanes move one 128-byte row per step
  for (int r = lane >> 3; r < nrows; r += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * ld + (lane & 7) * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    __half2* d = reinterpret_cast<__half2*>(dst + r * kPad + (lane & 7) * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = h[i];
  }
}
__device__ __forceinline__ float dot64(const __half* a, const __half* b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const float2 x = __half22float2(*reinterpret_cast<const __half2*>(a + 2 * d));
    const float2 y = __half22float2(*reinterpret_cast<const __half2*>(b + 2 * d));
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
  }
  return s;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void store_row64(__half* dst, const float (&acc)[64], float scale) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(acc[c * 8 + 2 * i] * scale, acc[c * 8 + 2 * i + 1] * scale);
    *reinterpret_cast<uint4*>(dst + c * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------ temporal (packed)
// rows ((z*hw + p)*F + f); one warp per (sequence, head); lane j plays key/value row j and, for dQ, query row j.
constexpr int kTbWarps = 4;
__global__ void __launch_bounds__(kTbWarps * 32) temporal_attn_bwd_kernel(
    const __half* __restrict__ qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const __half* __restrict__ d_o,
    long long ld_do, __half* __restrict__ d_qkv, long long ld_dqkv, long long n_seq, int heads, int F, float scale) {
  extern __shared__ __half sm_h[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tile = 32 * kPad;
  __half* sQ = sm_h + warp * 4 * tile;
  __half* sK = sQ + tile;
  __half* sV = sK + tile;
  __half* sD = sV + tile;
  float* sS = reinterpret_cast<float*>(sm_h + kTbWarps * 4 * tile) + warp * 32 * 33;
  const long long w = static_cast<long long>(blockIdx.x) * kTbWarps + warp;
  if (w >= n_seq * heads) return;
  const long long seq = w / heads;
  const int h = static_cast<int>(w % heads);
  const long long row0 = seq * F;
  load_rows_to_smem(sQ, qkv + row0 * ld_qkv + q_col0 + h * 64, ld_qkv, F, lane);
  load_rows_to_smem(sK, qkv + row0 * ld_qkv + k_col0 + h * 64, ld_qkv, F, lane);
  load_rows_to_smem(sV, qkv + row0 * ld_qkv + v_col0 + h * 64, ld_qkv, F, lane);
  load_rows_to_smem(sD, d_o + row0 * ld_do + h * 64, ld_do, F, lane);
  __syncwarp();
  const bool act = lane < F;
  const int j = act ? lane : 0;
  float p[32], ds[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    p[i] = 0.f;
    ds[i] = 0.f;
    if (i < F) {
      const float s = act ? dot64(sQ + i * kPad, sK + j * kPad) * scale : -INFINITY;
      const float m = warp_max(s);
      const float e = act ? __expf(s - m) : 0.f;
      const float l = warp_sum(e);
      p[i] = e / l;
      const float dp = act ? dot64(sD + i * kPad, sV + j * kPad) : 0.f;
      const float dsum = warp_sum(p[i] * dp);
      ds[i] = p[i] * (dp - dsum);
      sS[i * 33 + lane] = ds[i];
    }
  }
  __syncwarp();
  float acc[64];
  __half* out_row = d_qkv + (row0 + j) * ld_dqkv + h * 64;
  // dV[j] = sum_i P[i][j] dO[i]
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < F) {
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(sD + i * kPad + 2 * d));
        acc[2 * d] = fmaf(p[i], x.x, acc[2 * d]);
        acc[2 * d + 1] = fmaf(p[i], x.y, acc[2 * d + 1]);
      }
    }
  }
  if (act) store_row64(out_row + v_col0, acc, 1.f);
  // dK[j] = scale * sum_i dS[i][j] Q[i]
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < F) {
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(sQ + i * kPad + 2 * d));
        acc[2 * d] = fmaf(ds[i], x.x, acc[2 * d]);
        acc[2 * d + 1] = fmaf(ds[i], x.y, acc[2 * d + 1]);
      }
    }
  }
  if (act) store_row64(out_row + k_col0, acc, scale);
  // dQ[i=lane] = scale * sum_j dS[i][j] K[j]
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  for (int jj = 0; jj < F; ++jj) {
    const float w_ = sS[j * 33 + jj];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float2 x = __half22float2(*reinterpret_cast<const __half2*>(sK + jj * kPad + 2 * d));
      acc[2 * d] = fmaf(w_, x.x, acc[2 * d]);
      acc[2 * d + 1] = fmaf(w_, x.y, acc[2 * d + 1]);
    }
  }
  if (act) store_row64(out_row + q_col0, acc, scale);
}

// ------------------------------------------------------------------------------------------------ cross (dQ only)
struct CrossBwdArgs {
  const __half* q;
  const __half* kv;
  const __half* d_o;
  const __half* dp_tok;
  __half* dq;
  long long ld_q, ld_kv, ld_do, ld_dq;
  int k_col0, v_col0, s_q, s_k, heads, nz, kv_z_div, n_slots;
  int slot_col[16];
  float scale;
};
constexpr int kCbWarps = 4;
constexpr int kCbRows = 64;  // query rows per block
__global__ void __launch_bounds__(kCbWarps * 32) cross_attn_bwd_kernel(const CrossBwdArgs g) {
  extern __shared__ __half sm_c[];
  const int sk_pad = (g.s_k + 31) / 32 * 32;
  __half* sK = sm_c;
  __half* sV = sK + sk_pad * kPad;
  float* wbuf = reinterpret_cast<float*>(sV + sk_pad * kPad);  // per warp: q[64], do[64], ds[sk_pad]
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x, z = blockIdx.z;
  const int zk = z / g.kv_z_div;
  for (int r = threadIdx.x >> 3; r < sk_pad; r += (kCbWarps * 32) >> 3) {
    uint4 kq = make_uint4(0, 0, 0, 0), vq = make_uint4(0, 0, 0, 0);
    if (r < g.s_k) {
      const __half* base = g.kv + (static_cast<long long>(zk) * g.s_k + r) * g.ld_kv + h * 64 + (threadIdx.x & 7) * 8;
      kq = *reinterpret_cast<const uint4*>(base + g.k_col0);
      vq = *reinterpret_cast<const uint4*>(base + g.v_col0);
    }
    const __half2* kh = reinterpret_cast<const __half2*>(&kq);
    const __half2* vh = reinterpret_cast<const __half2*>(&vq);
    __half2* dk = reinterpret_cast<__half2*>(sK + r * kPad + (threadIdx.x & 7) * 8);
    __half2* dv = reinterpret_cast<__half2*>(sV + r * kPad + (threadIdx.x & 7) * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dk[i] = kh[i];
      dv[i] = vh[i];
    }
  }
  __syncthreads();
  float* wq = wbuf + warp * (128 + sk_pad);
  float* wdo = wq + 64;
  float* wds = wdo + 64;
  const int nk = sk_pad / 32;  // keys per lane (<= 4 supported)
  for (int rr = warp; rr < kCbRows; rr += kCbWarps) {
    const int qi = blockIdx.y * kCbRows + rr;
    if (qi >= g.s_q) break;
    const long long row = static_cast<long long>(z) * g.s_q + qi;
    {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(g.q + row * g.ld_q + h * 64 + lane * 2));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(g.d_o + row * g.ld_do + h * 64 + lane * 2));
      wq[lane * 2] = a.x;
      wq[lane * 2 + 1] = a.y;
      wdo[lane * 2] = b.x;
      wdo[lane * 2 + 1] = b.y;
    }
    __syncwarp();
    float s[4], dp[4];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      s[t] = -INFINITY;
      dp[t] = 0.f;
      if (t < nk) {
        const int j = lane + 32 * t;
        if (j < g.s_k) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            const float2 kx = __half22float2(*reinterpret_cast<const __half2*>(sK + j * kPad + 2 * d));
            const float2 vx = __half22float2(*reinterpret_cast<const __half2*>(sV + j * kPad + 2 * d));
            a = fmaf(wq[2 * d], kx.x, a);
            a = fmaf(wq[2 * d + 1], kx.y, a);
            b = fmaf(wdo[2 * d], vx.x, b);
            b = fmaf(wdo[2 * d + 1], vx.y, b);
          }
          s[t] = a * g.scale;
          dp[t] = b;
          for (int sl = 0; sl < g.n_slots; ++sl)
            if (g.slot_col[sl] == j)
              dp[t] += __half2float(
                  g.dp_tok[((static_cast<long long>(sl) * g.nz + z) * g.heads + h) * g.s_q + qi]);
        }
        m = fmaxf(m, s[t]);
      }
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      s[t] = (s[t] > -INFINITY) ? __expf(s[t] - m) : 0.f;
      l += s[t];
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    float dsum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // like the forward, the probabilities that multiply V are the fp16-rounded ones
      s[t] = round_h(s[t] * inv);
      dsum += s[t] * dp[t];
    }
    dsum = warp_sum(dsum);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < nk) wds[lane + 32 * t] = s[t] * (dp[t] - dsum);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < g.s_k; ++j) {
      const float w_ = wds[j];
      const float2 kx = __half22float2(*reinterpret_cast<const __half2*>(sK + j * kPad + 2 * lane));
      a0 = fmaf(w_, kx.x, a0);
      a1 = fmaf(w_, kx.y, a1);
    }
    *reinterpret_cast<__half2*>(g.dq + row * g.ld_dq + h * 64 + lane * 2) =
        __floats2half2_rn(a0 * g.scale, a1 * g.scale);
    __syncwarp();
  }
}

int launch_attn_bwd_small(int mode, const void* q, long long ld_q, long long q_width, long long q_rows, const void* kv,
                          long long ld_kv, long long kv_width, long long kv_rows, const void* d_o, long long ld_do,
                          long long do_width, const void* dp_tok, const int* slot_col, int n_slots, void* d_q,
                          long long ld_dq, int q_col0, int k_col0, int v_col0, int dq_col0, int dk_col0, int dv_col0,
                          int frames, long long s_q, int s_k, int heads, int nz, int kv_z_div, float scale,
                          cudaStream_t st);

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_temporal_attn_bwd(const void* qkv, int64_t ld_qkv, int64_t width_qkv, int32_t q_col0, int32_t k_col0,
                                     int32_t v_col0, const void* d_o, int64_t ld_do, void* d_qkv, int64_t ld_dqkv,
                                     int64_t n_seq, int32_t heads, int32_t frames, float scale, void* stream) {
  LVD_REQUIRE(qkv && d_o && d_qkv, "lvd_temporal_attn_bwd: null pointer");
  if (frames % 8 == 0 && frames >= 8 && frames <= 64 && width_qkv > 0) {
    // tensor-core path: packed tiles of floor(128/frames) sequences (attn_bwd_small_kernel, mode 1)
    return launch_attn_bwd_small(1, qkv, ld_qkv, width_qkv, n_seq * frames, qkv, ld_qkv, width_qkv, n_seq * frames, d_o,
                                 ld_do, static_cast<long long>(heads) * 64, nullptr, nullptr, 0, d_qkv, ld_dqkv, q_col0,
                                 k_col0, v_col0, q_col0, k_col0, v_col0, frames, 0, 0, heads, 1, 1, scale,
                                 static_cast<cudaStream_t>(stream));
  }
  LVD_REQUIRE(frames >= 1 && frames <= 32, "lvd_temporal_attn_bwd: frames=%d unsupported (<= 32)", frames);
  LVD_REQUIRE(ld_qkv % 8 == 0 && ld_do % 8 == 0 && ld_dqkv % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 &&
                  v_col0 % 8 == 0,
              "lvd_temporal_attn_bwd: alignment");
  const size_t smem = static_cast<size_t>(kTbWarps) * (4 * 32 * kPad * sizeof(__half) + 32 * 33 * sizeof(float));
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  const long long total = n_seq * heads;
  const unsigned blocks = static_cast<unsigned>((total + kTbWarps - 1) / kTbWarps);
  temporal_attn_bwd_kernel<<<blocks, kTbWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(qkv), ld_qkv, q_col0, k_col0, v_col0, static_cast<const __half*>(d_o), ld_do,
      static_cast<__half*>(d_qkv), ld_dqkv, n_seq, heads, frames, scale);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_cross_attn_bwd(const void* q, int64_t ld_q, int64_t width_q, const void* kv, int64_t ld_kv,
                                  int64_t width_kv, int32_t k_col0, int32_t v_col0, const void* d_o, int64_t ld_do,
                                  const void* dp_tok,
                                  const int32_t* slot_col_host, int32_t n_slots, void* dq, int64_t ld_dq, int64_t s_q,
                                  int32_t s_k, int32_t heads, int32_t nz, int32_t kv_z_div, float scale,
                                  const int32_t* slot_col_dev, void* stream) {
  LVD_REQUIRE(q && kv && d_o && dq, "lvd_cross_attn_bwd: null pointer");
  LVD_REQUIRE(s_k >= 1 && s_k <= 128 && n_slots >= -1 && n_slots <= 16 &&
                  (n_slots == 0 || dp_tok) && (n_slots <= 0 || slot_col_host),
              "lvd_cross_attn_bwd: s_k=%d n_slots=%d unsupported", s_k, n_slots);
  if (n_slots == -1) {  // dense dE/dP map: warp-level kernel only
    LVD_REQUIRE(width_q > 0 && width_kv > 0, "lvd_cross_attn_bwd: the dense dE/dP form needs the tensor-core path");
    const int r = launch_cross_attn_bwd_mma(q, ld_q, kv, ld_kv, k_col0, v_col0, d_o, ld_do, dp_tok, nullptr, -1, dq,
                                            ld_dq, s_q, s_k, heads, nz, kv_z_div, scale,
                                            static_cast<cudaStream_t>(stream));
    LVD_REQUIRE(r != 1, "lvd_cross_attn_bwd: dense dE/dP form: unsupported shape");
    return r;
  }
  if (width_q > 0 && width_kv > 0) {
    static const bool force_tile = getenv("LVD_CROSS_BWD_TILE") != nullptr;  // A/B switch for tools/
    if (!force_tile) {  // warp-level mma.sync kernel (attention_temporal.cu)
      const int r = launch_cross_attn_bwd_mma(q, ld_q, kv, ld_kv, k_col0, v_col0, d_o, ld_do, dp_tok, slot_col_host,
                                              n_slots, dq, ld_dq, s_q, s_k, heads, nz, kv_z_div, scale,
                                              static_cast<cudaStream_t>(stream), slot_col_dev);
      if (r != 1) return r;
    }
    LVD_REQUIRE(!slot_col_dev, "lvd_cross_attn_bwd: device-side column lists need the warp-level kernel");
    // tcgen05 tile path (attn_bwd_small_kernel, mode 0)
    return launch_attn_bwd_small(0, q, ld_q, width_q, static_cast<long long>(nz) * s_q, kv, ld_kv, width_kv,
                                 static_cast<long long>((nz + kv_z_div - 1) / kv_z_div) * s_k, d_o, ld_do,
                                 static_cast<long long>(heads) * 64, dp_tok, slot_col_host, n_slots, dq, ld_dq, 0, k_col0,
                                 v_col0, 0, 0, 0, 0, s_q, s_k, heads, nz, kv_z_div, scale,
                                 static_cast<cudaStream_t>(stream));
  }
  CrossBwdArgs g;
  g.q = static_cast<const __half*>(q);
  g.kv = static_cast<const __half*>(kv);
  g.d_o = static_cast<const __half*>(d_o);
  g.dp_tok = static_cast<const __half*>(dp_tok);
  g.dq = static_cast<__half*>(dq);
  g.ld_q = ld_q;
  g.ld_kv = ld_kv;
  g.ld_do = ld_do;
  g.ld_dq = ld_dq;
  g.k_col0 = k_col0;
  g.v_col0 = v_col0;
  g.s_q = static_cast<int>(s_q);
  g.s_k = s_k;
  g.heads = heads;
  g.nz = nz;
  g.kv_z_div = kv_z_div;
  g.n_slots = n_slots;
  for (int i = 0; i < n_slots; ++i) g.slot_col[i] = slot_col_host[i];
  g.scale = scale;
  const int sk_pad = (s_k + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(2) * sk_pad * kPad * sizeof(__half) + kCbWarps * (128 + sk_pad) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid(heads, static_cast<unsigned>((s_q + kCbRows - 1) / kCbRows), nz);
  cross_attn_bwd_kernel<<<grid, kCbWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
