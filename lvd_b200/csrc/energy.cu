// energy.cu -- cross-attention box energy of the LVD backward guidance and its gradient w.r.t. the attention
// probabilities (utils/guidance.py:160-526 add_ca_loss_per_attn_map_to_loss; :529-574 compute_ca_lossv3 normalisation):
//   loss_kind 0  "max-based" top-k fg/bg loss (:341-353, the default)
//   loss_kind 1  ratio-based loss (:312-323, CLI --use_ratio_based_loss)
//   loss_kind 2  CE / NLL top-k loss (:363-396)
//   + attn_sync (:401-426), BoxDiff corner constraint (:433-465, CLI --boxdiff_loss_scale / --boxdiff_normed),
//   + centre-of-mass position / velocity terms (:468-522, CLI --com_loss_scale).
//
// One CTA = one (slot = (object, text token), frame, head): the token's attention column over the hw pixels
// (<= 4096) is staged in shared memory, both top-k selections run as an in-CTA bitonic sort of packed
// (fp16 value bits, index) keys, and the CTA emits its loss contribution (one atomicAdd) and dE/dP for its column.
// The reference does this with ~10^4 tiny kernel launches and a host sync per top-k.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

constexpr int kEnergyThreads = 256;

struct EnergyArgs {
  const __half* p;
  const int* slot_col;
  const int* boxes;  // [n_slots][frames][4] = x0,y0,x1,y1 at this map's resolution
  const float* slot_weight;
  float* loss;
  __half* dp;
  int layout, n_text, n_slots, frames, heads, H, W, npad;
  float fg_top_p, bg_top_p, fg_weight, bg_weight, com_scale, out_scale, grad_scale;
  int loss_kind, k_min, boxdiff_normed, boxdiff_L;
  float eps, eps16, hi16, sync_w, boxdiff_scale;
  // several prompts in one map (batched guidance): slot -> first frame of its prompt inside p's frame axis, and the
  // index of its prompt's loss accumulator; p_frames = extent of p's frame axis (== frames for one prompt)
  const int* slot_frame0;
  const int* slot_loss;
  int p_frames;
};

// One element of a token column; `clamped`: the CE branch works on clamp(P, eps, 1 - eps) evaluated in fp16 (:368)
__device__ __forceinline__ float col_at(const EnergyArgs& g, int slot, int f, int h, int i, bool clamped);

__device__ __forceinline__ long long p_index(const EnergyArgs& g, int slot, int f, int h, int i) {
  const long long hw = static_cast<long long>(g.H) * g.W;
  const int z = f + (g.slot_frame0 ? g.slot_frame0[slot] : 0);
  if (g.layout == 0) return ((static_cast<long long>(slot) * g.p_frames + z) * g.heads + h) * hw + i;
  return ((static_cast<long long>(z) * g.heads + h) * hw + i) * g.n_text + g.slot_col[slot];
}

__device__ __forceinline__ float col_at(const EnergyArgs& g, int slot, int f, int h, int i, bool clamped) {
  float a = __half2float(g.p[p_index(g, slot, f, h, i)]);
  if (clamped) a = fminf(fmaxf(a, g.eps16), g.hi16);
  return a;
}

// descending bitonic sort of n (power of two) uint32 keys in shared memory
__device__ void bitonic_desc(uint32_t* k, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint32_t a = k[lo], b = k[hi];
        if ((a < b) == desc) {
          k[lo] = b;
          k[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

__device__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

struct Com {
  float s, ch, cw;
};

// centre of mass of one frame's column (center_of_mass, utils/guidance.py:154-157), fp32
__device__ Com com_of(const EnergyArgs& g, int slot, int f, int h, float* red) {
  const int hw = g.H * g.W;
  float s = 0.f, sy = 0.f, sx = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float a = __half2float(g.p[p_index(g, slot, f, h, i)]);
    s += a;
    sy += a * static_cast<float>(i / g.W);
    sx += a * static_cast<float>(i % g.W);
  }
  Com c;
  c.s = block_sum(s, red);
  c.ch = block_sum(sy, red) / c.s;
  c.cw = block_sum(sx, red) / c.s;
  return c;
}

__global__ void __launch_bounds__(kEnergyThreads) box_energy_kernel(const EnergyArgs g) {
  extern __shared__ uint32_t sm[];
  uint32_t* kf = sm;               // [npad] fg keys
  uint32_t* kb = sm + g.npad;      // [npad] bg keys
  float* grad = reinterpret_cast<float*>(sm + 2 * g.npad);  // [hw]
  __half* col = reinterpret_cast<__half*>(grad + g.H * g.W);  // [hw] the working column (clamped in the CE branch)
  __shared__ float red[kEnergyThreads / 32];
  const int slot = blockIdx.y;
  const int f = blockIdx.x / g.heads, h = blockIdx.x % g.heads;
  const int hw = g.H * g.W;
  if (g.slot_weight[slot] == 0.f) {  // padding slot of a graph-replayed table: no energy, zero gradient
    if (g.dp && g.layout == 0)
      for (int i = threadIdx.x; i < hw; i += blockDim.x) g.dp[p_index(g, slot, f, h, i)] = __float2half_rn(0.f);
    return;
  }
  const int* bx = g.boxes + (static_cast<long long>(slot) * g.frames + f) * 4;
  const int x0 = bx[0], y0 = bx[1], x1 = bx[2], y1 = bx[3];
  const int area = (x1 > x0 && y1 > y0) ? (x1 - x0) * (y1 - y0) : 0;
  const bool ce = g.loss_kind == 2;
  // k = max(k_min, trunc(fp32(count) * fp32(top_p)))   (guidance.py:328-337)
  int k_fg = static_cast<int>(static_cast<float>(area) * g.fg_top_p);
  int k_bg = static_cast<int>(static_cast<float>(hw - area) * g.bg_top_p);
  k_fg = k_fg < g.k_min ? g.k_min : k_fg;
  k_bg = k_bg < g.k_min ? g.k_min : k_bg;
  for (int i = threadIdx.x; i < g.npad; i += blockDim.x) {
    uint32_t a = 0, b = 0;
    if (i < hw) {
      const int y = i / g.W, x = i % g.W;
      const bool in = (x >= x0 && x < x1 && y >= y0 && y < y1);
      const __half v = __float2half_rn(col_at(g, slot, f, h, i, ce));
      col[i] = v;
      const uint32_t bits = __half_as_ushort(v);
      const uint32_t key = (bits << 16) | static_cast<uint32_t>(0xFFFF - i);  // ties -> lower index first
      a = in ? key : 0u;
      b = in ? 0u : key;
      grad[i] = 0.f;
    }
    kf[i] = a;
    kb[i] = b;
  }
  const float wslot = g.slot_weight[slot] * g.out_scale;
  float loss = 0.f;
  if (g.loss_kind == 1) {
    // ratio-based: activation = sum(P * mask) / (sum(P) + eps) per head, loss = mean_heads (1 - activation)^2
    __syncthreads();
    float sn = 0.f, sd = 0.f;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const float a = __half2float(col[i]);
      sd += a;
      const int y = i / g.W, x = i % g.W;
      if (x >= x0 && x < x1 && y >= y0 && y < y1) sn += a;
    }
    sn = block_sum(sn, red);
    sd = block_sum(sd, red) + g.eps;
    const float act = sn / sd;
    const float ih = 1.f / static_cast<float>(g.heads);
    loss = (1.f - act) * (1.f - act) * ih;
    const float c0 = -2.f * (1.f - act) * ih / (sd * sd);
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int y = i / g.W, x = i % g.W;
      const bool in = (x >= x0 && x < x1 && y >= y0 && y < y1);
      grad[i] = c0 * ((in ? sd : 0.f) - sn);
    }
  } else {
    bitonic_desc(kf, g.npad);
    bitonic_desc(kb, g.npad);
    const float ikf = 1.f / static_cast<float>(k_fg), ikb = 1.f / static_cast<float>(k_bg);
    if (!ce) {
      float sfg = 0.f, sbg = 0.f;
      for (int j = threadIdx.x; j < k_fg; j += blockDim.x) {
        const uint32_t key = kf[j];
        if (key != 0u) {
          sfg += __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16)));
          grad[0xFFFF - (key & 0xFFFFu)] += -g.fg_weight * ikf;
        }
      }
      for (int j = threadIdx.x; j < k_bg; j += blockDim.x) {
        const uint32_t key = kb[j];
        if (key != 0u) {
          sbg += __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16)));
          grad[0xFFFF - (key & 0xFFFFu)] += g.bg_weight * ikb;
        }
      }
      sfg = block_sum(sfg, red);
      sbg = block_sum(sbg, red);
      loss = (1.f - sfg * ikf) * g.fg_weight + (sbg * ikb) * g.bg_weight;
    } else {
      // CE / NLL: fg = mean_k -log(max(topk(mask * Pc), eps)); bg = -log(1 - mean_k topk((1 - mask) * Pc));
      // Pc = clamp(P, eps, 1 - eps); the clamps pass gradient only where the raw value lies inside the bounds
      float sfg = 0.f, sbg = 0.f;
      for (int j = threadIdx.x; j < k_fg; j += blockDim.x) {
        const uint32_t key = kf[j];
        const float v = key != 0u ? __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16))) : 0.f;
        sfg += -__logf(fmaxf(v, g.eps));
      }
      for (int j = threadIdx.x; j < k_bg; j += blockDim.x) {
        const uint32_t key = kb[j];
        if (key != 0u) sbg += __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16)));
      }
      sfg = block_sum(sfg, red);
      sbg = block_sum(sbg, red);
      const float mean_bg = sbg * ikb;
      loss = sfg * ikf * g.fg_weight - __logf(1.f - mean_bg) * g.bg_weight;
      for (int j = threadIdx.x; j < k_fg; j += blockDim.x) {
        const uint32_t key = kf[j];
        if (key == 0u) continue;
        const int i = 0xFFFF - (key & 0xFFFFu);
        const float v = __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16)));
        const float raw = col_at(g, slot, f, h, i, false);
        if (raw == v && v >= g.eps) grad[i] += -g.fg_weight * ikf / v;
      }
      for (int j = threadIdx.x; j < k_bg; j += blockDim.x) {
        const uint32_t key = kb[j];
        if (key == 0u) continue;
        const int i = 0xFFFF - (key & 0xFFFFu);
        const float v = __half2float(__ushort_as_half(static_cast<unsigned short>(key >> 16)));
        const float raw = col_at(g, slot, f, h, i, false);
        if (raw == v) grad[i] += g.bg_weight * ikb / (1.f - mean_bg);
      }
    }
  }
  __syncthreads();

  // ---- attention sync between consecutive frames (guidance.py:401-426; fp32 here, the reference evaluates the
  // squared difference in the map dtype).  The crop is the NEXT frame's rectangle: the reference's x_min..y_max were
  // last assigned by its loop over frame_boxes_t1 (:268-271).  An empty crop makes the reference's mean() NaN.
  if (g.sync_w != 0.f) {
    if (f + 1 < g.frames) {
      const int* bn = g.boxes + (static_cast<long long>(slot) * g.frames + (f + 1)) * 4;
      const int nx0 = bn[0], ny0 = bn[1], nx1 = bn[2], ny1 = bn[3];
      const int narea = (nx1 > nx0 && ny1 > ny0) ? (nx1 - nx0) * (ny1 - ny0) : 0;
      float acc = 0.f;
      if (narea > 0) {
        const int bw = nx1 - nx0;
        const float c2 = 2.f * g.sync_w / static_cast<float>(narea);
        for (int j = threadIdx.x; j < narea; j += blockDim.x) {
          const int i = (ny0 + j / bw) * g.W + nx0 + j % bw;
          const float d = __half2float(col[i]) - col_at(g, slot, f + 1, h, i, false);
          acc += d * d;
          grad[i] += c2 * d;
        }
      }
      acc = block_sum(acc, red);
      loss += narea > 0 ? g.sync_w * acc / static_cast<float>(narea) : __int_as_float(0x7fc00000);
    }
    if (f >= 1 && area > 0) {  // this frame is "frame 2" of the previous frame's term, cropped to THIS frame's box
      const int bw = x1 - x0;
      const float c2 = 2.f * g.sync_w / static_cast<float>(area);
      __syncthreads();
      for (int j = threadIdx.x; j < area; j += blockDim.x) {
        const int i = (y0 + j / bw) * g.W + x0 + j % bw;
        const float d = col_at(g, slot, f - 1, h, i, ce) - col_at(g, slot, f, h, i, false);
        grad[i] += -c2 * d;
      }
    }
    __syncthreads();
  }

  // ---- BoxDiff corner constraint (guidance.py:433-465): |max over rows/columns - projected mask| near the box edges
  if (g.boxdiff_scale > 0.f) {
    const int L = g.boxdiff_L;
    float acc = 0.f;
    const float nx = g.boxdiff_normed ? 1.f / static_cast<float>(g.heads * g.W) : 1.f;
    const float ny = g.boxdiff_normed ? 1.f / static_cast<float>(g.heads * g.H) : 1.f;
    for (int line = threadIdx.x; line < g.W + g.H; line += blockDim.x) {
      const bool is_x = line < g.W;
      const int c = is_x ? line : line - g.W;
      const int n = is_x ? g.H : g.W;
      const int stride = is_x ? g.W : 1, base = is_x ? c : c * g.W;
      float mx = -1.f;
      int arg = 0;
      for (int j = 0; j < n; ++j) {
        const float a = __half2float(col[base + j * stride]);
        if (a > mx) {
          mx = a;
          arg = base + j * stride;
        }
      }
      const int lo = is_x ? x0 : y0, hi = is_x ? x1 : y1, ext = is_x ? g.W : g.H;
      const float proj = (area > 0 && c >= lo && c < hi) ? 1.f : 0.f;
      const bool near_lo = c >= (lo - L > 0 ? lo - L : 0) && c < (lo + L + 1 < ext ? lo + L + 1 : ext);
      const bool near_hi = c >= (hi - L > 0 ? hi - L : 0) && c < (hi + L + 1 < ext ? hi + L + 1 : ext);
      if (near_lo || near_hi) {
        const float d = mx - proj;
        const float wn = g.boxdiff_scale * (is_x ? nx : ny);
        acc += fabsf(d) * wn;
        if (d != 0.f) atomicAdd(&grad[arg], d > 0.f ? wn : -wn);
      }
    }
    loss += block_sum(acc, red);
    __syncthreads();
  }

  if (g.com_scale > 0.f && area > 0) {
    const float cs = g.com_scale / static_cast<float>(g.heads);
    const float mh = 0.5f * static_cast<float>(y0 + y1 - 1), mw = 0.5f * static_cast<float>(x0 + x1 - 1);
    const Com c = com_of(g, slot, f, h, red);
    float gh = 2.f * (c.ch - mh) * cs, gw = 2.f * (c.cw - mw) * cs;  // dL/dch, dL/dcw (position term)
    loss += ((c.ch - mh) * (c.ch - mh) + (c.cw - mw) * (c.cw - mw)) * cs;
    // velocity term where this frame is the current one (next = min(f+1, F-1))
    const int f1 = f + 1 < g.frames ? f + 1 : g.frames - 1;
    if (f1 != f) {
      const int* b1 = g.boxes + (static_cast<long long>(slot) * g.frames + f1) * 4;
      if (b1[2] > b1[0] && b1[3] > b1[1]) {
        const float mh1 = 0.5f * static_cast<float>(b1[1] + b1[3] - 1), mw1 = 0.5f * static_cast<float>(b1[0] + b1[2] - 1);
        const Com c1 = com_of(g, slot, f1, h, red);
        const float vh = (c1.ch - c.ch) - (mh1 - mh), vw = (c1.cw - c.cw) - (mw1 - mw);
        loss += (vh * vh + vw * vw) * cs;
        gh += -2.f * vh * cs;
        gw += -2.f * vw * cs;
      }
    }
    // velocity term of the previous frame, where this frame is the "next" one
    if (f >= 1) {
      const int* bp = g.boxes + (static_cast<long long>(slot) * g.frames + (f - 1)) * 4;
      if (bp[2] > bp[0] && bp[3] > bp[1]) {
        const float mhp = 0.5f * static_cast<float>(bp[1] + bp[3] - 1), mwp = 0.5f * static_cast<float>(bp[0] + bp[2] - 1);
        const Com cp = com_of(g, slot, f - 1, h, red);
        const float vh = (c.ch - cp.ch) - (mh - mhp), vw = (c.cw - cp.cw) - (mw - mwp);
        gh += 2.f * vh * cs;
        gw += 2.f * vw * cs;
      }
    }
    const float inv_s = 1.f / c.s;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const float y = static_cast<float>(i / g.W), x = static_cast<float>(i % g.W);
      grad[i] += (gh * (y - c.ch) + gw * (x - c.cw)) * inv_s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(g.loss + (g.slot_loss ? g.slot_loss[slot] : 0), loss * wslot);
  if (g.dp) {
    const float gs = wslot * g.grad_scale;
    for (int i = threadIdx.x; i < hw; i += blockDim.x)
      g.dp[p_index(g, slot, f, h, i)] = __float2half_rn(grad[i] * gs);
  }
}

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_box_energy(const lvd_energy_desc* d, void* stream) {
  LVD_REQUIRE(d && d->p && d->boxes && d->slot_weight && d->loss, "lvd_box_energy: null pointer");
  LVD_REQUIRE(d->layout == 0 || (d->layout == 1 && d->slot_col && d->n_text > 0), "lvd_box_energy: bad layout");
  LVD_REQUIRE(d->n_slots >= 1 && d->frames >= 1 && d->heads >= 1 && d->H >= 1 && d->W >= 1, "lvd_box_energy: empty");
  const int hw = d->H * d->W;
  LVD_REQUIRE(hw <= 4096, "lvd_box_energy: attention map of %d pixels exceeds 4096", hw);
  int npad = 2;
  while (npad < hw) npad <<= 1;
  EnergyArgs g;
  g.p = static_cast<const __half*>(d->p);
  g.slot_col = d->slot_col;
  g.boxes = d->boxes;
  g.slot_weight = d->slot_weight;
  g.loss = d->loss;
  g.dp = static_cast<__half*>(d->dp);
  g.layout = d->layout;
  g.n_text = d->n_text;
  g.n_slots = d->n_slots;
  g.frames = d->frames;
  g.heads = d->heads;
  g.H = d->H;
  g.W = d->W;
  g.npad = npad;
  g.fg_top_p = d->fg_top_p;
  g.bg_top_p = d->bg_top_p;
  g.fg_weight = d->fg_weight;
  g.bg_weight = d->bg_weight;
  g.com_scale = d->com_loss_scale;
  g.out_scale = d->out_scale;
  g.grad_scale = d->grad_scale;
  LVD_REQUIRE(d->loss_kind >= 0 && d->loss_kind <= 2, "lvd_box_energy: loss_kind %d", d->loss_kind);
  LVD_REQUIRE(d->loss_kind != 2 || (d->attn_sync_weight == 0.f && d->boxdiff_loss_scale == 0.f &&
                                     d->com_loss_scale == 0.f),
              "lvd_box_energy: the CE loss combined with sync / BoxDiff / centre-of-mass terms is not implemented");
  g.loss_kind = d->loss_kind;
  g.k_min = d->k_min < 1 ? 1 : d->k_min;
  g.boxdiff_normed = d->boxdiff_normed;
  g.boxdiff_L = d->boxdiff_L;
  g.eps = d->eps;
  g.eps16 = __half2float(__float2half_rn(d->eps));
  g.hi16 = __half2float(__float2half_rn(1.f - d->eps));
  g.sync_w = d->attn_sync_weight;
  g.boxdiff_scale = d->boxdiff_loss_scale;
  g.slot_frame0 = d->slot_frame0;
  g.slot_loss = d->slot_loss;
  g.p_frames = d->p_frames > 0 ? d->p_frames : d->frames;
  LVD_REQUIRE(g.p_frames >= d->frames, "lvd_box_energy: p_frames < frames");
  const size_t smem = static_cast<size_t>(2 * npad + hw) * 4 + static_cast<size_t>(hw) * 2;
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(box_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  dim3 grid(d->frames * d->heads, d->n_slots);
  box_energy_kernel<<<grid, kEnergyThreads, smem, static_cast<cudaStream_t>(stream)>>>(g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
