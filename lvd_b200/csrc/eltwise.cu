// eltwise.cu -- small HBM-bound kernels around the tensor-core path: conv_in / conv_out (4-channel ends of the
// UNet, too thin for a tensor-core tile), nearest-2x upsample, timestep sinusoid, SiLU, latent/CFG/scheduler math.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

// ---------------------------------------------------------------------------------------------- conv_in
// sample (B, Cin<=8, F, H, W) fp16 NCFHW -> rows [(b*F+f)*H*W + p, Cout] ; 3x3 pad 1 (unet_3d_condition.py:297,726-729)
// One thread = one pixel x 8 output channels; weights [Cout][Cin][3][3] + bias staged in shared memory.
__global__ void conv_in_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                               const __half* __restrict__ bias, __half* __restrict__ out, long long ld_out, int B,
                               int Cin, int F, int H, int W, int Cout) {
  // weights transposed to [ci][ky][kx][Cout] in shared memory so that the 8 channels of a thread are one 16-byte,
  // bank-conflict-free read (neighbouring threads own neighbouring channel vectors)
  extern __shared__ __align__(16) __half s_w[];  // [Cin*9][Cout] + [Cout]
  const int wn = Cout * Cin * 9;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) {
    const int co = i / (Cin * 9), rem = i % (Cin * 9);  // source layout [co][ci][ky][kx]
    s_w[rem * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_w[wn + i] = bias[i];
  __syncthreads();
  const int cvec = Cout / 8;
  const long long total = static_cast<long long>(B) * F * H * W * cvec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    long long p = idx / cvec;
    const int xx = static_cast<int>(p % W);
    p /= W;
    const int yy = static_cast<int>(p % H);
    p /= H;
    const int f = static_cast<int>(p % F);
    const int b = static_cast<int>(p / F);
    float acc[8];
    {
      const uint4 bq = *reinterpret_cast<const uint4*>(&s_w[wn + cv * 8]);
      const __half2* bh = reinterpret_cast<const __half2*>(&bq);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(bh[i]);
        acc[2 * i] = t.x;
        acc[2 * i + 1] = t.y;
      }
    }
    for (int ci = 0; ci < Cin; ++ci) {
      const __half* xp = x + ((static_cast<long long>(b) * Cin + ci) * F + f) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xx + kx - 1;
          if (ix < 0 || ix >= W) continue;
          const float v = __half2float(xp[iy * W + ix]);
          const uint4 wq = *reinterpret_cast<const uint4*>(&s_w[((ci * 3 + ky) * 3 + kx) * Cout + cv * 8]);
          const __half2* wh = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(wh[i]);
            acc[2 * i] = fmaf(v, t.x, acc[2 * i]);
            acc[2 * i + 1] = fmaf(v, t.y, acc[2 * i + 1]);
          }
        }
      }
    }
    const long long row = ((static_cast<long long>(b) * F + f) * H + yy) * W + xx;
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(out + row * ld_out + cv * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------------- conv_out
// rows [(b*F+f)*H*W + p, Cin] -> (B, Cout<=8, F, H, W) NCFHW ; 3x3 pad 1 (unet_3d_condition.py:424,847-854).
// One warp per output pixel: lanes split the input channels, shuffle-reduce the Cout sums.
__global__ void conv_out_kernel(const __half* __restrict__ x, long long ld_x, const __half* __restrict__ wk,
                                const __half* __restrict__ bias, __half* __restrict__ out, int B, int Cin, int F, int H,
                                int W, int Cout) {
  extern __shared__ __half s_wo[];  // [Cout][9][Cin]  (K ordered ky,kx,ci like every other conv here)
  const int wn = Cout * 9 * Cin;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) s_wo[i] = wk[i];
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
    p /= H;  // p = b*F + f
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yy + ky - 1;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = xx + kx - 1;
        if (ix < 0 || ix >= W) continue;
        const __half* xr = x + ((p * H + iy) * W + ix) * ld_x;
        for (int c = lane * 2; c < Cin; c += 64) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(xr + c));
          for (int co = 0; co < Cout; ++co) {
            const float2 ww =
                __half22float2(*reinterpret_cast<const __half2*>(&s_wo[(co * 9 + ky * 3 + kx) * Cin + c]));
            acc[co] += v.x * ww.x + v.y * ww.y;
          }
        }
      }
    }
    for (int co = 0; co < Cout; ++co) {
      float a = acc[co];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[co] = a;
    }
    if (lane < Cout) {
      const int b = static_cast<int>(p / F), f = static_cast<int>(p % F);
      float a = 0.f;
      for (int co = 0; co < Cout; ++co)
        if (co == lane) a = acc[co];
      out[(((static_cast<long long>(b) * Cout + lane) * F + f) * H + yy) * W + xx] =
          __float2half_rn(a + __half2float(bias[lane]));
    }
  }
}

// ---------------------------------------------------------------------------------------------- rows -> NCFHW
// The wide path of conv_out: the 3x3 conv runs on the tensor-core kernel with the 4 output channels zero-padded to one
// 64-wide tile (gemm.cu), and this kernel gathers columns [0, Cout) of the [pixels, ld] result into (B, Cout, F, H, W).
__global__ void rows_to_ncfhw_kernel(const __half* __restrict__ rows, long long ld, __half* __restrict__ out, int B,
                                     int Cout, int F, long long hw) {
  const long long total = static_cast<long long>(B) * Cout * F * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = idx % hw;
    long long t = idx / hw;
    const int f = static_cast<int>(t % F);
    t /= F;
    const int co = static_cast<int>(t % Cout);
    const long long b = t / Cout;
    out[idx] = rows[((b * F + f) * hw + p) * ld + co];
  }
}

// ---------------------------------------------------------------------------------------------- VAE helpers
// Row softmax (fp32 math, in place) for the single-head 512-wide attention of the VAE mid block, whose two products run
// as plain GEMMs (diffusers Attention with upcast_softmax: softmax(scale * Q K^T) in fp32, cast to fp16).  One CTA per
// row, the row lives in registers (cols <= 8192).
constexpr int kSmxThreads = 256;
constexpr int kSmxVec = 4;  // 16-byte vectors per thread -> 8 * 4 * 256 = 8192 columns
__global__ void __launch_bounds__(kSmxThreads) softmax_rows_kernel(__half* __restrict__ x, long long ld, int cols,
                                                                   float scale_log2) {
  __shared__ float s_red[kSmxThreads / 32];
  __half* row = x + static_cast<long long>(blockIdx.x) * ld;
  const int nvec = cols >> 3;
  float v[kSmxVec][8];
  float mx = -3.0e38f;
#pragma unroll
  for (int k = 0; k < kSmxVec; ++k) {
    const int i = threadIdx.x + k * kSmxThreads;
    if (i < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[k][2 * j] = f.x;
        v[k][2 * j + 1] = f.y;
        mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < kSmxThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kSmxVec; ++k) {
    if (threadIdx.x + k * kSmxThreads < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] = exp2f((v[k][j] - mx) * scale_log2);
        sum += v[k][j];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;  // fixed order: deterministic
#pragma unroll
  for (int w = 0; w < kSmxThreads / 32; ++w) sum += s_red[w];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < kSmxVec; ++k) {
    const int i = threadIdx.x + k * kSmxThreads;
    if (i < nvec) {
      uint4 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[k][2 * j] * inv, v[k][2 * j + 1] * inv);
      *reinterpret_cast<uint4*>(row + i * 8) = u;
    }
  }
}

// im2col for the 4-channel conv_in (3x3, pad 1): sample (B, Cin<=7, F, H, W) -> rows [(b*F+f)*H*W + p][64] with column
// ci*9 + ky*3 + kx (the OIHW weight flattened), zero beyond Cin*9 and outside the image.  The conv itself then runs on the
// tensor-core kernel as a K = 64 GEMM; one thread per (pixel, 8-column vector).
__global__ void im2col_small_kernel(const __half* __restrict__ x, __half* __restrict__ out, int B, int Cin, int F, int H,
                                    int W) {
  const long long total = static_cast<long long>(B) * F * H * W * 8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx & 7);
    long long p = idx >> 3;
    const long long row = p;
    const int xx = static_cast<int>(p % W);
    p /= W;
    const int yy = static_cast<int>(p % H);
    p /= H;
    const int f = static_cast<int>(p % F);
    const int b = static_cast<int>(p / F);
    __align__(16) __half vals[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = v * 8 + i;
      const int ci = col / 9, tap = col % 9;
      const int iy = yy + tap / 3 - 1, ix = xx + tap % 3 - 1;
      const bool ok = ci < Cin && iy >= 0 && iy < H && ix >= 0 && ix < W;
      vals[i] = ok ? x[(((static_cast<long long>(b) * Cin + ci) * F + f) * H + iy) * W + ix] : __float2half(0.f);
    }
    *reinterpret_cast<uint4*>(out + row * 64 + v * 8) = *reinterpret_cast<const uint4*>(vals);
  }
}

// CLIPTextEmbeddings: token embedding gather + position embedding (fp16 add), one 16-byte vector per thread.
__global__ void embed_tokens_kernel(const long long* __restrict__ ids, const __half* __restrict__ tok,
                                    const __half* __restrict__ pos, __half* __restrict__ out, long long n_tokens, int seq,
                                    int dim, int vocab) {
  const int vpr = dim >> 3;
  const long long total = n_tokens * vpr;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = idx / vpr;
    const int v = static_cast<int>(idx % vpr);
    long long id = ids[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const uint4 a = *reinterpret_cast<const uint4*>(tok + id * dim + v * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(pos + (t % seq) * dim + v * 8);
    uint4 o;
    const __half2* ha = reinterpret_cast<const __half2*>(&a);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __hadd2(ha[i], hb[i]);
    *reinterpret_cast<uint4*>(out + t * dim + v * 8) = o;
  }
}

// 1x1 convolution between <= 8 channels on planar (N, C, HW) fp16 (AutoencoderKL.post_quant_conv): thread per pixel.
__global__ void pointwise_small_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                       const __half* __restrict__ bias, __half* __restrict__ out, int N, int Cin,
                                       int Cout, long long hw) {
  const long long total = static_cast<long long>(N) * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = idx / hw, p = idx % hw;
    float in[8];
    for (int ci = 0; ci < Cin; ++ci) in[ci] = __half2float(x[(n * Cin + ci) * hw + p]);
    for (int co = 0; co < Cout; ++co) {
      float a = bias ? __half2float(bias[co]) : 0.f;
      for (int ci = 0; ci < Cin; ++ci) a = fmaf(__half2float(w[co * Cin + ci]), in[ci], a);
      out[(n * Cout + co) * hw + p] = __float2half_rn(a);
    }
  }
}

// (N, C, HW) fp16 in [-1, 1]  ->  (N, HW, C) fp32 in [0, 1]: VaeImageProcessor.postprocess(output_type="np") =
// denormalize (x / 2 + 0.5).clamp(0, 1) + channels-last, as tensor2vid applies it per video
// (controllable_pipeline_text_to_video_synth.py:66-88).
__global__ void frames_postprocess_kernel(const __half* __restrict__ x, float* __restrict__ out, long long n, int C,
                                          long long hw) {
  const long long total = n * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = idx / hw, p = idx % hw;
    for (int c = 0; c < C; ++c) {
      float v = __half2float(x[(i * C + c) * hw + p]) * 0.5f + 0.5f;
      v = fminf(fmaxf(v, 0.f), 1.f);
      out[idx * C + c] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- upsample 2x nearest
__global__ void upsample2x_kernel(const __half* __restrict__ x, long long ld_x, __half* __restrict__ y, long long ld_y,
                                  int n_img, int H, int W, int C) {
  const int cvec = C / 8;
  const long long total = static_cast<long long>(n_img) * (2 * H) * (2 * W) * cvec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    long long p = idx / cvec;
    const int ox = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int oy = static_cast<int>(p % (2 * H));
    const long long n = p / (2 * H);
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((n * H + (oy >> 1)) * W + (ox >> 1)) * ld_x + cv * 8);
    *reinterpret_cast<uint4*>(y + ((n * 2 * H + oy) * 2 * W + ox) * ld_y + cv * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------- timestep sinusoid
// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) (unet_3d_condition.py:304,712-713):
// out[b, i] = cos(t_b * f_i), out[b, half + i] = sin(t_b * f_i), f_i = exp(-ln(10000) * i / half); fp32 -> fp16.
__global__ void timestep_embed_kernel(const float* __restrict__ t, __half* __restrict__ out, int B, int dim) {
  const int half = dim / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * half; idx += gridDim.x * blockDim.x) {
    const int b = idx / half, i = idx % half;
    const float freq = expf(-9.210340371976184f * static_cast<float>(i) / static_cast<float>(half));
    const float a = t[b] * freq;
    out[b * dim + i] = __float2half_rn(cosf(a));
    out[b * dim + half + i] = __float2half_rn(sinf(a));
  }
}

// ---------------------------------------------------------------------------------------------- y = silu(x) (fp16)
__global__ void silu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __half2float(x[i]);
    y[i] = __float2half_rn(v / (1.0f + expf(-v)));
  }
}

// ---------------------------------------------------------------------------------------------- strided row copy
__global__ void copy_rows_kernel(const __half* __restrict__ x, long long ld_x, __half* __restrict__ y, long long ld_y,
                                 long long rows, int C) {
  const int cvec = C / 8;
  const long long total = rows * cvec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / cvec;
    const int cv = static_cast<int>(idx % cvec);
    *reinterpret_cast<uint4*>(y + r * ld_y + cv * 8) = *reinterpret_cast<const uint4*>(x + r * ld_x + cv * 8);
  }
}

static unsigned grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 32LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_conv_in(const void* sample, const void* weight, const void* bias, void* out, int64_t ld_out,
                           int32_t batch, int32_t cin, int32_t frames, int32_t height, int32_t width, int32_t cout,
                           void* stream) {
  LVD_REQUIRE(sample && weight && bias && out, "lvd_conv_in: null pointer");
  LVD_REQUIRE(cin >= 1 && cin <= 8 && cout % 8 == 0 && cout <= 1024 && ld_out % 8 == 0, "lvd_conv_in: bad channels");
  const long long total = static_cast<long long>(batch) * frames * height * width * (cout / 8);
  const size_t smem = (static_cast<size_t>(cout) * cin * 9 + cout) * sizeof(__half);
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  conv_in_kernel<<<grid_for(total, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(sample), static_cast<const __half*>(weight), static_cast<const __half*>(bias),
      static_cast<__half*>(out), ld_out, batch, cin, frames, height, width, cout);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_conv_out(const void* x, int64_t ld_x, const void* weight_krsc, const void* bias, void* out,
                            int32_t batch, int32_t cin, int32_t frames, int32_t height, int32_t width, int32_t cout,
                            void* stream) {
  LVD_REQUIRE(x && weight_krsc && bias && out, "lvd_conv_out: null pointer");
  LVD_REQUIRE(cout >= 1 && cout <= 8 && cin % 2 == 0 && ld_x % 8 == 0, "lvd_conv_out: bad channels");
  const size_t smem = static_cast<size_t>(cout) * 9 * cin * sizeof(__half);
  LVD_REQUIRE(smem <= 160 * 1024, "lvd_conv_out: weights do not fit shared memory");
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  const long long npix = static_cast<long long>(batch) * frames * height * width;
  conv_out_kernel<<<grid_for(npix * 32, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), ld_x, static_cast<const __half*>(weight_krsc), static_cast<const __half*>(bias),
      static_cast<__half*>(out), batch, cin, frames, height, width, cout);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_rows_to_ncfhw(const void* rows, int64_t ld, void* out, int32_t batch, int32_t cout, int32_t frames,
                                 int64_t hw, void* stream) {
  LVD_REQUIRE(rows && out && batch >= 1 && cout >= 1 && cout <= ld && frames >= 1 && hw >= 1,
              "lvd_rows_to_ncfhw: bad arguments");
  const long long total = static_cast<long long>(batch) * cout * frames * hw;
  rows_to_ncfhw_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(rows), ld, static_cast<__half*>(out), batch, cout, frames, hw);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_softmax_rows(void* x, int64_t ld, int64_t rows, int32_t cols, float scale, void* stream) {
  LVD_REQUIRE(x && rows >= 1 && cols >= 8 && cols % 8 == 0 && cols <= 8 * kSmxVec * kSmxThreads && ld % 8 == 0 &&
                  rows < (1ll << 31),
              "lvd_softmax_rows: bad arguments (cols=%d)", cols);
  softmax_rows_kernel<<<static_cast<unsigned>(rows), kSmxThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__half*>(x), ld, cols, scale * 1.4426950408889634f);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_frames_postprocess(const void* x, void* out, int64_t n, int32_t channels, int64_t hw, void* stream) {
  LVD_REQUIRE(x && out && n >= 1 && channels >= 1 && channels <= 8 && hw >= 1, "lvd_frames_postprocess: bad arguments");
  frames_postprocess_kernel<<<grid_for(n * hw, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<float*>(out), n, channels, hw);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_im2col_small(const void* sample, void* out, int32_t batch, int32_t cin, int32_t frames, int32_t height,
                                int32_t width, void* stream) {
  LVD_REQUIRE(sample && out && batch >= 1 && cin >= 1 && cin <= 7 && frames >= 1 && height >= 1 && width >= 1,
              "lvd_im2col_small: bad arguments (cin * 9 must fit 64 columns)");
  const long long total = static_cast<long long>(batch) * frames * height * width * 8;
  im2col_small_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(sample), static_cast<__half*>(out), batch, cin, frames, height, width);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_embed_tokens(const int64_t* ids, const void* token_embedding, const void* position_embedding,
                                void* out, int64_t n_tokens, int32_t seq, int32_t dim, int32_t vocab, void* stream) {
  LVD_REQUIRE(ids && token_embedding && position_embedding && out && n_tokens >= 1 && seq >= 1 && dim % 8 == 0 &&
                  vocab >= 1,
              "lvd_embed_tokens: bad arguments");
  embed_tokens_kernel<<<grid_for(n_tokens * (dim / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), static_cast<const __half*>(token_embedding),
      static_cast<const __half*>(position_embedding), static_cast<__half*>(out), n_tokens, seq, dim, vocab);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_pointwise_small(const void* x, const void* weight, const void* bias, void* out, int32_t n,
                                   int32_t cin, int32_t cout, int64_t hw, void* stream) {
  LVD_REQUIRE(x && weight && out && n >= 1 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8 && hw >= 1,
              "lvd_pointwise_small: bad arguments");
  pointwise_small_kernel<<<grid_for(static_cast<long long>(n) * hw, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<const __half*>(weight), static_cast<const __half*>(bias),
      static_cast<__half*>(out), n, cin, cout, hw);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_upsample2x(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t n_img, int32_t height,
                              int32_t width, int32_t channels, void* stream) {
  LVD_REQUIRE(x && y && channels % 8 == 0 && ld_x % 8 == 0 && ld_y % 8 == 0, "lvd_upsample2x: bad arguments");
  const long long total = static_cast<long long>(n_img) * 4 * height * width * (channels / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, n_img, height, width, channels);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_timestep_embed(const void* timesteps_f32, void* out, int32_t batch, int32_t dim, void* stream) {
  LVD_REQUIRE(timesteps_f32 && out && dim % 2 == 0, "lvd_timestep_embed: bad arguments");
  timestep_embed_kernel<<<grid_for(batch * (dim / 2), 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(timesteps_f32), static_cast<__half*>(out), batch, dim);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_silu(const void* x, void* y, int64_t n, void* stream) {
  LVD_REQUIRE(x && y, "lvd_silu: null pointer");
  silu_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), n);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_copy_rows(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels,
                             void* stream) {
  LVD_REQUIRE(x && y && channels % 8 == 0 && ld_x % 8 == 0 && ld_y % 8 == 0, "lvd_copy_rows: bad arguments");
  copy_rows_kernel<<<grid_for(rows * (channels / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows, channels);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

// ---------------------------------------------------------------------------------------------- CFG + DPM-Solver++ step
// Fuses, per latent element (controllable_pipeline_text_to_video_synth.py:926-950 + diffusers 0.27.2
// DPMSolverMultistepScheduler.step):
//   eps  = eps_u + gs * (eps_c - eps_u)                        (fp16 ops, as the reference evaluates them)
//   x0   = (x - sigma_t * eps) / alpha_t                       (fp16: convert_model_output, epsilon prediction)
//   x'   = c_x * float(x) - half(c_d0 * x0) - half(c_d1 * D1),  D1 = half(inv_r0 * half(x0 - x0_prev))   (2nd order)
// x0 is written to x0_out (the scheduler's model_outputs ring).  order == 1 drops the D1 term.
namespace lvd {
__device__ __forceinline__ float hmul(float a, float b) { return round_h(a * b); }
__global__ void cfg_dpm_step_kernel(const __half* __restrict__ eps_u, const __half* __restrict__ eps_c,
                                    const __half* __restrict__ x, const __half* __restrict__ x0_prev,
                                    __half* __restrict__ x_out, __half* __restrict__ x0_out, long long n, float gs,
                                    float sigma_t, float alpha_t, float c_x, float c_d0, float c_d1, float inv_r0,
                                    int order, int use_cfg) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float u = __half2float(eps_u[i]);
    float e = u;
    if (use_cfg) {
      const float c = __half2float(eps_c[i]);
      e = round_h(u + hmul(gs, round_h(c - u)));
    }
    const float xs = __half2float(x[i]);
    const float x0 = hmul(round_h(xs - hmul(sigma_t, e)), 1.0f / alpha_t);  // torch divides by a scalar as a multiply
    float xn = c_x * xs - hmul(c_d0, x0);
    if (order == 2) {
      const float d1 = hmul(inv_r0, round_h(x0 - __half2float(x0_prev[i])));
      xn -= hmul(c_d1, d1);
    }
    x_out[i] = __float2half_rn(xn);
    x0_out[i] = __float2half_rn(x0);
  }
}
// x <- x - scale * grad   (models/pipelines.py:124-132), fp16 ops
__global__ void axpy_h_kernel(const __half* __restrict__ x, const __half* __restrict__ gvec, __half* __restrict__ out,
                              long long n, float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2half_rn(__half2float(x[i]) - hmul(scale, __half2float(gvec[i])));
}
}  // namespace lvd

extern "C" int lvd_cfg_dpm_step(const void* eps_uncond, const void* eps_cond, const void* x, const void* x0_prev,
                                void* x_out, void* x0_out, int64_t n, float guidance_scale, float sigma_t,
                                float alpha_t, float c_x, float c_d0, float c_d1, float inv_r0, int32_t order,
                                void* stream) {
  LVD_REQUIRE(eps_uncond && x && x_out && x0_out, "lvd_cfg_dpm_step: null pointer");
  LVD_REQUIRE(order == 1 || (order == 2 && x0_prev), "lvd_cfg_dpm_step: order %d", order);
  cfg_dpm_step_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(eps_uncond), static_cast<const __half*>(eps_cond), static_cast<const __half*>(x),
      static_cast<const __half*>(x0_prev), static_cast<__half*>(x_out), static_cast<__half*>(x0_out), n,
      guidance_scale, sigma_t, alpha_t, c_x, c_d0, c_d1, inv_r0, order, eps_cond != nullptr);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_axpy_h(const void* x, const void* g, void* out, int64_t n, float scale, void* stream) {
  LVD_REQUIRE(x && g && out, "lvd_axpy_h: null pointer");
  axpy_h_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<const __half*>(g), static_cast<__half*>(out), n, scale);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

// ---------------------------------------------------------------------------------------------- sequence-strided row copy
// dst[(r / n)*dst_seq + dst_off + r % n] = src[(r / n)*src_seq + src_off + r % n]   for r in [0, groups*n)
// (builds / splits the [visual tokens ; grounding tokens] sequences of GatedSelfAttentionDense, models/attention.py:50-57)
namespace lvd {
__global__ void copy_rows_seq_kernel(const __half* __restrict__ x, long long ld_x, __half* __restrict__ y,
                                     long long ld_y, long long rows, int C, int n, int src_seq, int src_off,
                                     int dst_seq, int dst_off) {
  const int cvec = C / 8;
  const long long total = rows * cvec;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / cvec;
    const int cv = static_cast<int>(idx % cvec);
    const long long grp = r / n, j = r % n;
    *reinterpret_cast<uint4*>(y + (grp * dst_seq + dst_off + j) * ld_y + cv * 8) =
        *reinterpret_cast<const uint4*>(x + (grp * src_seq + src_off + j) * ld_x + cv * 8);
  }
}

// GLIGEN PositionNet input (models/unet_3d_condition.py:47-64,119-150): Fourier embedding of the xyxy box
// (8 frequencies 100^(j/8), layout j*8 + {sin,cos}*4 + coord), masked blend with the learned null features, and
// concatenation [positive_embedding (D) | position embedding (64)].  fp32 math on fp16 inputs.
__global__ void gligen_prep_kernel(const __half* __restrict__ boxes, const __half* __restrict__ masks,
                                   const __half* __restrict__ pos, const __half* __restrict__ null_pos,
                                   const __half* __restrict__ null_xyxy, __half* __restrict__ out, long long n_obj,
                                   int D) {
  const int width = D + 64;
  const long long total = n_obj * width;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long o = idx / width;
    const int c = static_cast<int>(idx % width);
    const float m = __half2float(masks[o]);
    float v;
    if (c < D) {
      v = __half2float(pos[o * D + c]) * m + (1.f - m) * __half2float(null_pos[c]);
    } else {
      const int e = c - D, j = e >> 3, s = (e >> 2) & 1, k = e & 3;
      const float freq = powf(100.f, static_cast<float>(j) / 8.f);
      const float a = freq * __half2float(boxes[o * 4 + k]);
      v = (s == 0 ? sinf(a) : cosf(a)) * m + (1.f - m) * __half2float(null_xyxy[e]);
    }
    out[idx] = __float2half_rn(v);
  }
}
}  // namespace lvd

extern "C" int lvd_copy_rows_seq(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels,
                                 int32_t n, int32_t src_seq, int32_t src_off, int32_t dst_seq, int32_t dst_off,
                                 void* stream) {
  LVD_REQUIRE(x && y && channels % 8 == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 && n >= 1, "lvd_copy_rows_seq: bad arguments");
  copy_rows_seq_kernel<<<grid_for(rows * (channels / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), ld_x, static_cast<__half*>(y), ld_y, rows, channels, n, src_seq, src_off, dst_seq,
      dst_off);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

extern "C" int lvd_gligen_prep(const void* boxes, const void* masks, const void* positive_embeddings,
                               const void* null_positive, const void* null_position, void* out, int64_t n_obj,
                               int32_t dim, void* stream) {
  LVD_REQUIRE(boxes && masks && positive_embeddings && null_positive && null_position && out, "lvd_gligen_prep: null");
  gligen_prep_kernel<<<grid_for(n_obj * (dim + 64), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(boxes), static_cast<const __half*>(masks),
      static_cast<const __half*>(positive_embeddings), static_cast<const __half*>(null_positive),
      static_cast<const __half*>(null_position), static_cast<__half*>(out), n_obj, dim);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
