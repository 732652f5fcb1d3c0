// energy_prep.cu -- the optional pre-processing of a saved cross-attention map ahead of the box energy, and the
// adjoints that carry dE/d(column) back to dE/dP (utils/guidance.py:209-226, 281-311; utils/attn.py:91-157):
//   smooth   : GaussianSmoothing -- a depthwise 3x3 blur of the reflect-padded (hw x n_text) plane of every
//              (frame, head); it mixes neighbouring FLATTENED pixel indices and neighbouring text tokens, in the map dtype
//   renorm   : softmax(P[..., 1:num_tokens-1] * renorm_scale) over the kept text tokens
//   columns  : gather of the object-token columns, bilinearly interpolated to (H_out, W_out) when upsample_scale != 1
//              (F.interpolate(mode="bilinear", align_corners=False))
// All are off by default in the runners; one thread per output element, gather-form adjoints (no atomics, deterministic).
// Gradients travel as fp32 between the adjoint stages; the last stage writes the fp16 dense dE/dP the cross-attention
// adjoint consumes (lvd_cross_attn_bwd, n_slots == -1).
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

struct Blur9 {
  float w[9];  // [di + 1][dj + 1], already rounded to fp16 like `self.weight.to(input.dtype)` (utils/attn.py:157)
};

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void smooth_kernel(const __half* __restrict__ p, __half* __restrict__ out, long long planes, int hw,
                              int n_text, Blur9 k) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(hw) * n_text;
  if (idx >= planes * per) return;
  const long long pl = idx / per;
  const int i = static_cast<int>((idx % per) / n_text), j = static_cast<int>(idx % n_text);
  const __half* base = p + pl * per;
  float acc = 0.f;
#pragma unroll
  for (int di = -1; di <= 1; ++di) {
    const int ii = reflect1(i + di, hw);
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj)
      acc += k.w[(di + 1) * 3 + dj + 1] * __half2float(base[static_cast<long long>(ii) * n_text + reflect1(j + dj, n_text)]);
  }
  out[idx] = __float2half_rn(acc);
}

// adjoint of smooth_kernel: dP[i][j] = sum over padded coordinates (a, b) that reflect onto (i, j) and taps (di, dj)
// of w[di][dj] * dS[a - di][b - dj]
template <typename TOut>
__global__ void smooth_bwd_kernel(const float* __restrict__ ds, TOut* __restrict__ dp, long long planes, int hw,
                                  int n_text, Blur9 k) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(hw) * n_text;
  if (idx >= planes * per) return;
  const long long pl = idx / per;
  const int i = static_cast<int>((idx % per) / n_text), j = static_cast<int>(idx % n_text);
  const float* base = ds + pl * per;
  int ai[3], bj[3], na = 0, nb = 0;
  ai[na++] = i;
  if (i == 1) ai[na++] = -1;
  if (i == hw - 2) ai[na++] = hw;
  bj[nb++] = j;
  if (j == 1) bj[nb++] = -1;
  if (j == n_text - 2) bj[nb++] = n_text;
  float acc = 0.f;
  for (int x = 0; x < na; ++x)
    for (int y = 0; y < nb; ++y)
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        const int oi = ai[x] - di;
        if (oi < 0 || oi >= hw) continue;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
          const int oj = bj[y] - dj;
          if (oj < 0 || oj >= n_text) continue;
          acc += k.w[(di + 1) * 3 + dj + 1] * base[static_cast<long long>(oi) * n_text + oj];
        }
      }
  if constexpr (sizeof(TOut) == 2) {
    dp[idx] = __float2half_rn(acc);
  } else {
    dp[idx] = acc;
  }
}

// one thread per row: out[row][0 .. n_keep) = softmax(p[row][1 .. 1 + n_keep) * scale)
__global__ void renorm_kernel(const __half* __restrict__ p, __half* __restrict__ out, long long rows, int n_text,
                              int n_keep, float scale) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const __half* x = p + r * n_text + 1;
  float mx = -3.0e38f;
  for (int j = 0; j < n_keep; ++j) mx = fmaxf(mx, __half2float(x[j]) * scale);
  float sum = 0.f;
  for (int j = 0; j < n_keep; ++j) sum += __expf(__half2float(x[j]) * scale - mx);
  const float inv = 1.f / sum;
  for (int j = 0; j < n_keep; ++j) out[r * n_keep + j] = __float2half_rn(__expf(__half2float(x[j]) * scale - mx) * inv);
}

// dP[row][1 + j] = scale * R[j] * (dR[j] - sum_k R[k] dR[k]); the dropped columns get zero
template <typename TOut>
__global__ void renorm_bwd_kernel(const __half* __restrict__ r_out, const float* __restrict__ dr, TOut* __restrict__ dp,
                                  long long rows, int n_text, int n_keep, float scale) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const __half* R = r_out + r * n_keep;
  const float* d = dr + r * n_keep;
  float dot = 0.f;
  for (int j = 0; j < n_keep; ++j) dot += __half2float(R[j]) * d[j];
  for (int j = 0; j < n_text; ++j) {
    float v = 0.f;
    if (j >= 1 && j < 1 + n_keep) v = scale * __half2float(R[j - 1]) * (d[j - 1] - dot);
    if constexpr (sizeof(TOut) == 2) {
      dp[r * n_text + j] = __float2half_rn(v);
    } else {
      dp[r * n_text + j] = v;
    }
  }
}

// PyTorch's bilinear source index (align_corners = False): src = (dst + 0.5) * in / out - 0.5, clamped at 0
__device__ __forceinline__ void bilinear_src(int dst, int n_in, int n_out, int& i0, int& i1, float& lam) {
  const float scale = static_cast<float>(n_in) / static_cast<float>(n_out);
  float src = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  i0 = i0 > n_in - 1 ? n_in - 1 : i0;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  lam = src - static_cast<float>(i0);
}

// cols[slot][plane][Y * W_out + X] = interp(map[plane][., slot_col[slot]])
__global__ void columns_kernel(const __half* __restrict__ map, const int* __restrict__ slot_col, __half* __restrict__ cols,
                               int n_slots, long long planes, int n_cols, int h_in, int w_in, int h_out, int w_out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw_out = static_cast<long long>(h_out) * w_out;
  if (idx >= static_cast<long long>(n_slots) * planes * hw_out) return;
  const int slot = static_cast<int>(idx / (planes * hw_out));
  const long long pl = (idx / hw_out) % planes;
  const int Y = static_cast<int>((idx % hw_out) / w_out), X = static_cast<int>(idx % w_out);
  const __half* base = map + pl * h_in * w_in * n_cols + slot_col[slot];
  if (h_in == h_out && w_in == w_out) {
    cols[idx] = base[static_cast<long long>(Y * w_in + X) * n_cols];
    return;
  }
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(Y, h_in, h_out, y0, y1, ly);
  bilinear_src(X, w_in, w_out, x0, x1, lx);
  const float a = __half2float(base[static_cast<long long>(y0 * w_in + x0) * n_cols]);
  const float b = __half2float(base[static_cast<long long>(y0 * w_in + x1) * n_cols]);
  const float c = __half2float(base[static_cast<long long>(y1 * w_in + x0) * n_cols]);
  const float d = __half2float(base[static_cast<long long>(y1 * w_in + x1) * n_cols]);
  cols[idx] = __float2half_rn((1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d));
}

// adjoint: dmap[plane][pixel][j] = sum over slots with slot_col == j of the interpolation-weighted d(cols); gather form
// over the output pixels whose footprint can include the input pixel
template <typename TOut>
__global__ void columns_bwd_kernel(const __half* __restrict__ dcols, const int* __restrict__ slot_col,
                                   TOut* __restrict__ dmap, int n_slots, long long planes, int n_cols, int h_in, int w_in,
                                   int h_out, int w_out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw_in = static_cast<long long>(h_in) * w_in;
  if (idx >= planes * hw_in * n_cols) return;
  const int j = static_cast<int>(idx % n_cols);
  const int pix = static_cast<int>((idx / n_cols) % hw_in);
  const long long pl = idx / (n_cols * hw_in);
  const int y = pix / w_in, x = pix % w_in;
  const long long hw_out = static_cast<long long>(h_out) * w_out;
  float acc = 0.f;
  for (int s = 0; s < n_slots; ++s) {
    if (slot_col[s] != j) continue;
    const __half* src = dcols + (static_cast<long long>(s) * planes + pl) * hw_out;
    if (h_in == h_out && w_in == w_out) {
      acc += __half2float(src[pix]);
      continue;
    }
    const int ry = (h_out + h_in - 1) / h_in, rx = (w_out + w_in - 1) / w_in;
    const int Y_lo = max(0, (y - 1) * ry - ry), Y_hi = min(h_out, (y + 2) * ry + ry);
    const int X_lo = max(0, (x - 1) * rx - rx), X_hi = min(w_out, (x + 2) * rx + rx);
    for (int Y = Y_lo; Y < Y_hi; ++Y) {
      int y0, y1;
      float ly;
      bilinear_src(Y, h_in, h_out, y0, y1, ly);
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int X = X_lo; X < X_hi; ++X) {
        int x0, x1;
        float lx;
        bilinear_src(X, w_in, w_out, x0, x1, lx);
        const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
        if (wx != 0.f) acc += wy * wx * __half2float(src[static_cast<long long>(Y) * w_out + X]);
      }
    }
  }
  if constexpr (sizeof(TOut) == 2) {
    dmap[idx] = __float2half_rn(acc);
  } else {
    dmap[idx] = acc;
  }
}

static Blur9 make_blur(const float* w9) {
  Blur9 k;
  for (int i = 0; i < 9; ++i) k.w[i] = __half2float(__float2half_rn(w9[i]));
  return k;
}

static inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_attn_smooth(const void* p, void* out, int64_t planes, int32_t hw, int32_t n_text, const float* w9_host,
                               void* stream) {
  LVD_REQUIRE(p && out && w9_host && planes >= 1 && hw >= 2 && n_text >= 2, "lvd_attn_smooth: bad arguments");
  const long long n = planes * hw * n_text;
  LVD_REQUIRE(n < (1ll << 39), "lvd_attn_smooth: map too large");
  smooth_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(p), static_cast<__half*>(out), planes, hw, n_text, make_blur(w9_host));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_attn_smooth_bwd(const void* d_out_f32, void* d_p, int32_t d_p_is_half, int64_t planes, int32_t hw,
                                   int32_t n_text, const float* w9_host, void* stream) {
  LVD_REQUIRE(d_out_f32 && d_p && w9_host && planes >= 1 && hw >= 2 && n_text >= 2, "lvd_attn_smooth_bwd: bad arguments");
  const long long n = planes * hw * n_text;
  const Blur9 k = make_blur(w9_host);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_p_is_half) {
    smooth_bwd_kernel<__half><<<blocks_for(n), 256, 0, st>>>(static_cast<const float*>(d_out_f32),
                                                             static_cast<__half*>(d_p), planes, hw, n_text, k);
  } else {
    smooth_bwd_kernel<float><<<blocks_for(n), 256, 0, st>>>(static_cast<const float*>(d_out_f32),
                                                            static_cast<float*>(d_p), planes, hw, n_text, k);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_attn_renorm(const void* p, void* out, int64_t rows, int32_t n_text, int32_t num_tokens, float scale,
                               void* stream) {
  LVD_REQUIRE(p && out && rows >= 1 && num_tokens >= 3 && num_tokens <= n_text, "lvd_attn_renorm: bad arguments");
  renorm_kernel<<<blocks_for(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(p), static_cast<__half*>(out), rows, n_text, num_tokens - 2, scale);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_attn_renorm_bwd(const void* r_out, const void* d_r_f32, void* d_p, int32_t d_p_is_half, int64_t rows,
                                   int32_t n_text, int32_t num_tokens, float scale, void* stream) {
  LVD_REQUIRE(r_out && d_r_f32 && d_p && rows >= 1 && num_tokens >= 3 && num_tokens <= n_text,
              "lvd_attn_renorm_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_p_is_half) {
    renorm_bwd_kernel<__half><<<blocks_for(rows), 256, 0, st>>>(static_cast<const __half*>(r_out),
                                                                static_cast<const float*>(d_r_f32),
                                                                static_cast<__half*>(d_p), rows, n_text, num_tokens - 2,
                                                                scale);
  } else {
    renorm_bwd_kernel<float><<<blocks_for(rows), 256, 0, st>>>(static_cast<const __half*>(r_out),
                                                               static_cast<const float*>(d_r_f32),
                                                               static_cast<float*>(d_p), rows, n_text, num_tokens - 2,
                                                               scale);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_attn_columns(const void* map, const int32_t* slot_col_dev, void* cols, int32_t n_slots, int64_t planes,
                                int32_t n_cols, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out, void* stream) {
  LVD_REQUIRE(map && slot_col_dev && cols && n_slots >= 1 && planes >= 1 && n_cols >= 1 && h_in >= 1 && w_in >= 1 &&
                  h_out >= h_in && w_out >= w_in,
              "lvd_attn_columns: bad arguments");
  const long long n = static_cast<long long>(n_slots) * planes * h_out * w_out;
  columns_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(map), slot_col_dev, static_cast<__half*>(cols), n_slots, planes, n_cols, h_in, w_in,
      h_out, w_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_attn_columns_bwd(const void* d_cols, const int32_t* slot_col_dev, void* d_map, int32_t d_map_is_half,
                                    int32_t n_slots, int64_t planes, int32_t n_cols, int32_t h_in, int32_t w_in,
                                    int32_t h_out, int32_t w_out, void* stream) {
  LVD_REQUIRE(d_cols && slot_col_dev && d_map && n_slots >= 1 && planes >= 1 && n_cols >= 1 && h_out >= h_in &&
                  w_out >= w_in,
              "lvd_attn_columns_bwd: bad arguments");
  const long long n = planes * h_in * w_in * n_cols;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_map_is_half) {
    columns_bwd_kernel<__half><<<blocks_for(n), 256, 0, st>>>(static_cast<const __half*>(d_cols), slot_col_dev,
                                                              static_cast<__half*>(d_map), n_slots, planes, n_cols, h_in,
                                                              w_in, h_out, w_out);
  } else {
    columns_bwd_kernel<float><<<blocks_for(n), 256, 0, st>>>(static_cast<const __half*>(d_cols), slot_col_dev,
                                                             static_cast<float*>(d_map), n_slots, planes, n_cols, h_in,
                                                             w_in, h_out, w_out);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
