// attention_temporal.cu -- frame-axis self-attention (TransformerTemporalModel, models/transformer_temporal.py:158-170;
// BasicTransformerBlock attn1/attn2 with S = F <= 32 frames, 64-wide heads) as ONE WARP per (pixel sequence, head).
//
// Why not the tcgen05 tile kernel (attention.cu, packed mode): a sequence is 24 rows, so a 128-row tile holds 5 of them
// and the CTA's life is prologue (barriers, TMEM allocation, 112 KB of zero fill) rather than math; the op moves 353 MB
// for 4 GFLOP at the Zeroscope shape and is bound by how many bytes are in flight.  Here every lane issues all of its
// 16-byte loads of Q, K and V (the warp's whole 9 KB working set) before the first dependent instruction, the two small
// products run on mma.sync m16n8k16 straight out of those registers, and nothing touches shared memory, TMEM or a
// barrier:
//   * S = Q K^T: the contraction index may be permuted freely as long as Q and K agree, so lane (g, t) takes its
//     A / B fragment halves from the two 16-byte pieces of its own row (columns 8t.. and 32 + 8t..);
//   * O = P V: P is the S accumulator re-packed (the usual register-level C -> A reuse); the B fragment needs two keys
//     of one feature column, which __byte_perm builds from the lane's own four V rows; the output feature permutation
//     is chosen so that every lane ends up with two contiguous 16-byte pieces of each output row.
// Softmax is exact per row (fp32, exp2 of the max-subtracted scaled scores), probabilities are rounded to fp16 before
// P V like the reference's attention_probs.to(value.dtype) (models/attention_processor.py:253-256).
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t word_of(const uint4& u, int i) {
  return i == 0 ? u.x : (i == 1 ? u.y : (i == 2 ? u.z : u.w));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint4 ldg16(const __half* p) { return *reinterpret_cast<const uint4*>(p); }

struct TemporalArgs {
  const __half* q;
  const __half* kv;
  __half* out;
  long long ld_q, ld_kv, ld_o;
  long long n_warps;  // sequences * heads
  int q_col0, k_col0, v_col0, o_col0;
  int frames, heads;
  float scale_log2;
};

// MT = 16-row query tiles, NT = 8-key tiles (= frames / 8), KS = 16-key contraction steps of P V
template <int MT, int NT, int KS>
__global__ void __launch_bounds__(128) temporal_attn_mma_kernel(const TemporalArgs a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long w = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (w >= a.n_warps) return;
  const int h = static_cast<int>(w % a.heads);  // heads of one sequence sit in neighbouring warps: same DRAM rows
  const long long seq = w / a.heads;
  const int F = a.frames;
  const __half* qb = a.q + seq * F * a.ld_q + a.q_col0 + h * 64 + t * 8;
  const __half* kb = a.kv + seq * F * a.ld_kv + a.k_col0 + h * 64 + t * 8;
  const int vchunk = (g >> 1) + 4 * (g & 1);  // feature permutation: logical column g lives in 16-byte piece pi(g)
  const __half* vb = a.kv + seq * F * a.ld_kv + a.v_col0 + h * 64 + vchunk * 8;
  const uint4 zero = make_uint4(0, 0, 0, 0);

  // ---- every load of the warp's working set, back to back
  uint4 qf[2 * MT][2], kf[NT][2], vf[KS][4];
#pragma unroll
  for (int i = 0; i < 2 * MT; ++i) {
    const int r = i * 8 + g;  // i = 2 * mt + half
    qf[i][0] = r < F ? ldg16(qb + r * a.ld_q) : zero;
    qf[i][1] = r < F ? ldg16(qb + r * a.ld_q + 32) : zero;
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int n = nt * 8 + g;
    kf[nt][0] = ldg16(kb + n * a.ld_kv);
    kf[nt][1] = ldg16(kb + n * a.ld_kv + 32);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ks * 16 + 2 * t + (i & 1) + (i >> 1) * 8;  // keys 2t, 2t+1, 2t+8, 2t+9 of the step
      vf[ks][i] = j < F ? ldg16(vb + j * a.ld_kv) : zero;
    }
  }

  // ---- S = Q K^T (fp32 accumulators).  Step s of the contraction uses words 2s, 2s+1 of the lane's 8 words per row.
  float s[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const int piece = st >> 1, w0 = (st & 1) * 2;
        mma16816(s[mt][nt], word_of(qf[2 * mt][piece], w0), word_of(qf[2 * mt + 1][piece], w0),
                 word_of(qf[2 * mt][piece], w0 + 1), word_of(qf[2 * mt + 1][piece], w0 + 1),
                 word_of(kf[nt][piece], w0), word_of(kf[nt][piece], w0 + 1));
      }
    }

  // ---- exact softmax per query row (rows g and g + 8 of every 16-row tile; a row's 8 * NT scores live in 4 lanes)
  uint32_t p[MT][NT][2];  // fp16 pairs: [0] = row g, [1] = row g + 8
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float mx = -3.0e38f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mx = fmaxf(mx, fmaxf(s[mt][nt][2 * half], s[mt][nt][2 * half + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float e0 = fast_exp2((s[mt][nt][2 * half] - mx) * a.scale_log2);
        const float e1 = fast_exp2((s[mt][nt][2 * half + 1] - mx) * a.scale_log2);
        s[mt][nt][2 * half] = e0;
        s[mt][nt][2 * half + 1] = e1;
        sum += e0 + e1;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        p[mt][nt][half] = pack_h2(s[mt][nt][2 * half] * inv, s[mt][nt][2 * half + 1] * inv);
    }
  }

  // ---- O = P V: 8 feature tiles; tile nv, logical column n  <->  physical column 8 * pi(n) + nv
  float o[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nv = 0; nv < 8; ++nv)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][nv][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int nv = 0; nv < 8; ++nv) {
      const uint32_t sel = (nv & 1) ? 0x7632u : 0x5410u;  // (key 2t, key 2t+1) halves of feature nv
      const uint32_t b0 = __byte_perm(word_of(vf[ks][0], nv >> 1), word_of(vf[ks][1], nv >> 1), sel);
      const uint32_t b1 = __byte_perm(word_of(vf[ks][2], nv >> 1), word_of(vf[ks][3], nv >> 1), sel);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t a0 = p[mt][2 * ks][0], a1 = p[mt][2 * ks][1];
        const uint32_t a2 = 2 * ks + 1 < NT ? p[mt][2 * ks + 1 < NT ? 2 * ks + 1 : 0][0] : 0u;
        const uint32_t a3 = 2 * ks + 1 < NT ? p[mt][2 * ks + 1 < NT ? 2 * ks + 1 : 0][1] : 0u;
        mma16816(o[mt][nv], a0, a1, a2, a3, b0, b1);
      }
    }
  }

  // ---- store: lane (g, t) holds physical columns 8t .. 8t+7 (logical 2t) and 32 + 8t .. (logical 2t+1) of its rows
  __half* ob = a.out + seq * F * a.ld_o + a.o_col0 + h * 64 + t * 8;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = mt * 16 + half * 8 + g;
      if (r < F) {
        uint4 u0, u1;
        u0.x = pack_h2(o[mt][0][2 * half], o[mt][1][2 * half]);
        u0.y = pack_h2(o[mt][2][2 * half], o[mt][3][2 * half]);
        u0.z = pack_h2(o[mt][4][2 * half], o[mt][5][2 * half]);
        u0.w = pack_h2(o[mt][6][2 * half], o[mt][7][2 * half]);
        u1.x = pack_h2(o[mt][0][2 * half + 1], o[mt][1][2 * half + 1]);
        u1.y = pack_h2(o[mt][2][2 * half + 1], o[mt][3][2 * half + 1]);
        u1.z = pack_h2(o[mt][4][2 * half + 1], o[mt][5][2 * half + 1]);
        u1.w = pack_h2(o[mt][6][2 * half + 1], o[mt][7][2 * half + 1]);
        *reinterpret_cast<uint4*>(ob + r * a.ld_o) = u0;
        *reinterpret_cast<uint4*>(ob + r * a.ld_o + 32) = u1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ cross-attention
// softmax(Q K^T) V against <= 128 keys (the 77 text tokens) without probability outputs: one warp per (16 query rows,
// head).  Same register-level scheme as above; K and V fragments of the (image-batch, head) are re-read by every warp
// and come out of L1 / L2 (20 KB per head), Q and O are the only HBM traffic.  Keys >= s_k score -inf.
struct CrossArgs {
  const __half* q;
  const __half* kv;
  __half* out;
  long long ld_q, ld_kv, ld_o;
  long long n_warps;  // nz * q_tiles * heads
  int q_col0, k_col0, v_col0, o_col0;
  int s_q, s_k, heads, q_tiles, kv_z_div, causal;
  float scale_log2;
};

template <int NT, int KS>
__global__ void __launch_bounds__(128) cross_attn_mma_kernel(const CrossArgs a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long w = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (w >= a.n_warps) return;
  const int h = static_cast<int>(w % a.heads);
  const long long zq = w / a.heads;
  const int qt = static_cast<int>(zq % a.q_tiles);
  const long long z = zq / a.q_tiles;
  const long long zk = z / a.kv_z_div;
  const int r0 = qt * 16;
  const __half* qb = a.q + (z * a.s_q) * a.ld_q + a.q_col0 + h * 64 + t * 8;
  const __half* kb = a.kv + (zk * a.s_k) * a.ld_kv + a.k_col0 + h * 64 + t * 8;
  const int vchunk = (g >> 1) + 4 * (g & 1);
  const __half* vb = a.kv + (zk * a.s_k) * a.ld_kv + a.v_col0 + h * 64 + vchunk * 8;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 qf[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + i * 8 + g;
    qf[i][0] = r < a.s_q ? ldg16(qb + r * a.ld_q) : zero;
    qf[i][1] = r < a.s_q ? ldg16(qb + r * a.ld_q + 32) : zero;
  }
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int n = nt * 8 + g;
    const uint4 k0 = n < a.s_k ? ldg16(kb + n * a.ld_kv) : zero;
    const uint4 k1 = n < a.s_k ? ldg16(kb + n * a.ld_kv + 32) : zero;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int w0 = (st & 1) * 2;
      const uint4& kq = (st >> 1) ? k1 : k0;
      mma16816(s[nt], word_of(qf[0][st >> 1], w0), word_of(qf[1][st >> 1], w0), word_of(qf[0][st >> 1], w0 + 1),
               word_of(qf[1][st >> 1], w0 + 1), word_of(kq, w0), word_of(kq, w0 + 1));
    }
  }
  uint32_t p[NT][2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float mx = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = nt * 8 + 2 * t;
      const int lim = a.causal ? min(a.s_k, r0 + half * 8 + g + 1) : a.s_k;  // causal: keys 0..query row
      if (c >= lim) s[nt][2 * half] = -INFINITY;
      if (c + 1 >= lim) s[nt][2 * half + 1] = -INFINITY;
      mx = fmaxf(mx, fmaxf(s[nt][2 * half], s[nt][2 * half + 1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float e0 = fast_exp2((s[nt][2 * half] - mx) * a.scale_log2);
      const float e1 = fast_exp2((s[nt][2 * half + 1] - mx) * a.scale_log2);
      s[nt][2 * half] = e0;
      s[nt][2 * half + 1] = e1;
      sum += e0 + e1;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) p[nt][half] = pack_h2(s[nt][2 * half] * inv, s[nt][2 * half + 1] * inv);
  }
  float o[8][4];
#pragma unroll
  for (int nv = 0; nv < 8; ++nv)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nv][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint4 vf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ks * 16 + 2 * t + (i & 1) + (i >> 1) * 8;
      vf[i] = j < a.s_k ? ldg16(vb + j * a.ld_kv) : zero;
    }
    const uint32_t a0 = p[2 * ks][0], a1 = p[2 * ks][1];
    const uint32_t a2 = 2 * ks + 1 < NT ? p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][0] : 0u;
    const uint32_t a3 = 2 * ks + 1 < NT ? p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][1] : 0u;
#pragma unroll
    for (int nv = 0; nv < 8; ++nv) {
      const uint32_t sel = (nv & 1) ? 0x7632u : 0x5410u;
      const uint32_t b0 = __byte_perm(word_of(vf[0], nv >> 1), word_of(vf[1], nv >> 1), sel);
      const uint32_t b1 = __byte_perm(word_of(vf[2], nv >> 1), word_of(vf[3], nv >> 1), sel);
      mma16816(o[nv], a0, a1, a2, a3, b0, b1);
    }
  }
  __half* ob = a.out + (z * a.s_q) * a.ld_o + a.o_col0 + h * 64 + t * 8;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8 + g;
    if (r < a.s_q) {
      uint4 u0, u1;
      u0.x = pack_h2(o[0][2 * half], o[1][2 * half]);
      u0.y = pack_h2(o[2][2 * half], o[3][2 * half]);
      u0.z = pack_h2(o[4][2 * half], o[5][2 * half]);
      u0.w = pack_h2(o[6][2 * half], o[7][2 * half]);
      u1.x = pack_h2(o[0][2 * half + 1], o[1][2 * half + 1]);
      u1.y = pack_h2(o[2][2 * half + 1], o[3][2 * half + 1]);
      u1.z = pack_h2(o[4][2 * half + 1], o[5][2 * half + 1]);
      u1.w = pack_h2(o[6][2 * half + 1], o[7][2 * half + 1]);
      *reinterpret_cast<uint4*>(ob + r * a.ld_o) = u0;
      *reinterpret_cast<uint4*>(ob + r * a.ld_o + 32) = u1;
    }
  }
}

// Cross-attention without probability outputs and <= 128 keys; returns 1 when the shape is not covered.
int launch_cross_attn_mma(const lvd_attn_desc* d, cudaStream_t stream) {
  if (d->p_full || d->p_tok || d->lse || d->s_k < 1 || d->s_k > 128) return 1;
  CrossArgs a;
  a.q = static_cast<const __half*>(d->q);
  a.kv = static_cast<const __half*>(d->kv);
  a.out = static_cast<__half*>(d->out);
  a.ld_q = d->ld_q;
  a.ld_kv = d->ld_kv;
  a.ld_o = d->ld_o;
  a.q_col0 = d->q_col0;
  a.k_col0 = d->k_col0;
  a.v_col0 = d->v_col0;
  a.o_col0 = d->o_col0;
  a.s_q = static_cast<int>(d->s_q);
  a.s_k = static_cast<int>(d->s_k);
  a.heads = d->heads;
  a.q_tiles = (a.s_q + 15) / 16;
  a.kv_z_div = d->kv_z_div;
  a.causal = d->causal ? 1 : 0;
  a.n_warps = static_cast<long long>(d->nz) * a.q_tiles * d->heads;
  a.scale_log2 = d->scale * 1.4426950408889634f;
  const long long blocks = (a.n_warps + 3) / 4;
  LVD_REQUIRE(blocks < (1ll << 31), "lvd_attention: too many query tiles");
  const unsigned nb = static_cast<unsigned>(blocks);
  const int nt = (a.s_k + 7) / 8;
  if (nt <= 4) {
    cross_attn_mma_kernel<4, 2><<<nb, 128, 0, stream>>>(a);
  } else if (nt <= 10) {
    cross_attn_mma_kernel<10, 5><<<nb, 128, 0, stream>>>(a);
  } else {
    cross_attn_mma_kernel<16, 8><<<nb, 128, 0, stream>>>(a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

// ------------------------------------------------------------------------------------------------ cross-attention dQ
// Adjoint of the map-less cross-attention above w.r.t. Q only (K, V come from the text encoder, no function of the
// latents), with the box-energy gradient dE/dP added on the object-token columns -- the place where utils/guidance.py's
// loss enters the backward pass.  One warp per (16 query rows, head), everything from registers:
//   S = Q K^T and dP = dO V^T share one fragment scheme (V's rows play K's role), P = fp16 softmax as in the forward,
//   dS = P * (dP - rowsum(P * dP)), dQ = scale * dS K with K in the role V has in P V.
struct CrossBwdMmaArgs {
  const __half* q;
  const __half* kv;
  const __half* d_o;
  const __half* dp_tok;  // [n_slots][nz][heads][s_q]
  __half* dq;
  long long ld_q, ld_kv, ld_do, ld_dq;
  long long n_warps;
  int k_col0, v_col0;
  int s_q, s_k, heads, nz, q_tiles, kv_z_div, n_slots;
  int slot_col[16];
  const int* slot_col_dev;  // optional device copy, read at run time (graph replays with other token positions)
  float scale, scale_log2;
};

template <int NT, int KS>
__global__ void __launch_bounds__(128) cross_attn_bwd_mma_kernel(const CrossBwdMmaArgs a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long w = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (w >= a.n_warps) return;
  const int h = static_cast<int>(w % a.heads);
  const long long zq = w / a.heads;
  const int qt = static_cast<int>(zq % a.q_tiles);
  const long long z = zq / a.q_tiles;
  const long long zk = z / a.kv_z_div;
  const int r0 = qt * 16;
  const __half* qb = a.q + (z * a.s_q) * a.ld_q + h * 64 + t * 8;
  const __half* db = a.d_o + (z * a.s_q) * a.ld_do + h * 64 + t * 8;
  const __half* kb = a.kv + (zk * a.s_k) * a.ld_kv + a.k_col0 + h * 64;
  const __half* vb = a.kv + (zk * a.s_k) * a.ld_kv + a.v_col0 + h * 64;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 qf[2][2], df[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + i * 8 + g;
    qf[i][0] = r < a.s_q ? ldg16(qb + r * a.ld_q) : zero;
    qf[i][1] = r < a.s_q ? ldg16(qb + r * a.ld_q + 32) : zero;
    df[i][0] = r < a.s_q ? ldg16(db + r * a.ld_do) : zero;
    df[i][1] = r < a.s_q ? ldg16(db + r * a.ld_do + 32) : zero;
  }
  float s[NT][4], dp[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int n = nt * 8 + g;
    const bool ok = n < a.s_k;
    const uint4 k0 = ok ? ldg16(kb + n * a.ld_kv + t * 8) : zero, k1 = ok ? ldg16(kb + n * a.ld_kv + t * 8 + 32) : zero;
    const uint4 v0 = ok ? ldg16(vb + n * a.ld_kv + t * 8) : zero, v1 = ok ? ldg16(vb + n * a.ld_kv + t * 8 + 32) : zero;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int w0 = (st & 1) * 2, pc = st >> 1;
      const uint4& kq = pc ? k1 : k0;
      const uint4& vq = pc ? v1 : v0;
      mma16816(s[nt], word_of(qf[0][pc], w0), word_of(qf[1][pc], w0), word_of(qf[0][pc], w0 + 1),
               word_of(qf[1][pc], w0 + 1), word_of(kq, w0), word_of(kq, w0 + 1));
      mma16816(dp[nt], word_of(df[0][pc], w0), word_of(df[1][pc], w0), word_of(df[0][pc], w0 + 1),
               word_of(df[1][pc], w0 + 1), word_of(vq, w0), word_of(vq, w0 + 1));
    }
  }
  // dE/dP of the guided token columns (this lane owns columns 8 nt + 2t, + 1 of rows g and g + 8); the register
  // arrays keep static indices: the slot list is searched per owned column
  if (a.n_slots < 0) {
    // dense dE/dP [nz][heads][s_q][s_k] (map pre-processing variants of the energy, or more than 16 token slots)
    const int ra = r0 + g, rb = r0 + 8 + g;
    const __half* src = a.dp_tok + ((z * a.heads + h) * static_cast<long long>(a.s_q)) * a.s_k;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = nt * 8 + 2 * t + e;
        if (c < a.s_k) {
          if (ra < a.s_q) dp[nt][e] += __half2float(src[static_cast<long long>(ra) * a.s_k + c]);
          if (rb < a.s_q) dp[nt][2 + e] += __half2float(src[static_cast<long long>(rb) * a.s_k + c]);
        }
      }
    }
  } else if (a.n_slots > 0) {
    const int ra = r0 + g, rb = r0 + 8 + g;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = nt * 8 + 2 * t + e;
        float va = 0.f, vb_ = 0.f;
        for (int sl = 0; sl < a.n_slots; ++sl) {
          if ((a.slot_col_dev ? a.slot_col_dev[sl] : a.slot_col[sl]) == c) {
            const __half* src = a.dp_tok + ((static_cast<long long>(sl) * a.nz + z) * a.heads + h) * a.s_q;
            if (ra < a.s_q) va += __half2float(src[ra]);
            if (rb < a.s_q) vb_ += __half2float(src[rb]);
          }
        }
        dp[nt][e] += va;
        dp[nt][2 + e] += vb_;
      }
    }
  }
  uint32_t ds[NT][2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float mx = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= a.s_k) s[nt][2 * half] = -INFINITY;
      if (c + 1 >= a.s_k) s[nt][2 * half + 1] = -INFINITY;
      mx = fmaxf(mx, fmaxf(s[nt][2 * half], s[nt][2 * half + 1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float e0 = fast_exp2((s[nt][2 * half] - mx) * a.scale_log2);
      const float e1 = fast_exp2((s[nt][2 * half + 1] - mx) * a.scale_log2);
      s[nt][2 * half] = e0;
      s[nt][2 * half + 1] = e1;
      sum += e0 + e1;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    float dsum = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {  // the probabilities that multiplied V in the forward are the fp16-rounded ones
      s[nt][2 * half] = round_h(s[nt][2 * half] * inv);
      s[nt][2 * half + 1] = round_h(s[nt][2 * half + 1] * inv);
      dsum += s[nt][2 * half] * dp[nt][2 * half] + s[nt][2 * half + 1] * dp[nt][2 * half + 1];
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      ds[nt][half] = pack_h2(s[nt][2 * half] * (dp[nt][2 * half] - dsum),
                             s[nt][2 * half + 1] * (dp[nt][2 * half + 1] - dsum));
  }
  // dQ = scale * dS K  (K rows in the role of V rows in P V)
  float o[8][4];
#pragma unroll
  for (int nv = 0; nv < 8; ++nv)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nv][i] = 0.f;
  const int vchunk = (g >> 1) + 4 * (g & 1);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint4 kf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = ks * 16 + 2 * t + (i & 1) + (i >> 1) * 8;
      kf[i] = j < a.s_k ? ldg16(kb + j * a.ld_kv + vchunk * 8) : zero;
    }
    const uint32_t a0 = ds[2 * ks][0], a1 = ds[2 * ks][1];
    const uint32_t a2 = 2 * ks + 1 < NT ? ds[2 * ks + 1 < NT ? 2 * ks + 1 : 0][0] : 0u;
    const uint32_t a3 = 2 * ks + 1 < NT ? ds[2 * ks + 1 < NT ? 2 * ks + 1 : 0][1] : 0u;
#pragma unroll
    for (int nv = 0; nv < 8; ++nv) {
      const uint32_t sel = (nv & 1) ? 0x7632u : 0x5410u;
      const uint32_t b0 = __byte_perm(word_of(kf[0], nv >> 1), word_of(kf[1], nv >> 1), sel);
      const uint32_t b1 = __byte_perm(word_of(kf[2], nv >> 1), word_of(kf[3], nv >> 1), sel);
      mma16816(o[nv], a0, a1, a2, a3, b0, b1);
    }
  }
  __half* ob = a.dq + (z * a.s_q) * a.ld_dq + h * 64 + t * 8;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8 + g;
    if (r < a.s_q) {
      uint4 u0, u1;
      u0.x = pack_h2(o[0][2 * half] * a.scale, o[1][2 * half] * a.scale);
      u0.y = pack_h2(o[2][2 * half] * a.scale, o[3][2 * half] * a.scale);
      u0.z = pack_h2(o[4][2 * half] * a.scale, o[5][2 * half] * a.scale);
      u0.w = pack_h2(o[6][2 * half] * a.scale, o[7][2 * half] * a.scale);
      u1.x = pack_h2(o[0][2 * half + 1] * a.scale, o[1][2 * half + 1] * a.scale);
      u1.y = pack_h2(o[2][2 * half + 1] * a.scale, o[3][2 * half + 1] * a.scale);
      u1.z = pack_h2(o[4][2 * half + 1] * a.scale, o[5][2 * half + 1] * a.scale);
      u1.w = pack_h2(o[6][2 * half + 1] * a.scale, o[7][2 * half + 1] * a.scale);
      *reinterpret_cast<uint4*>(ob + r * a.ld_dq) = u0;
      *reinterpret_cast<uint4*>(ob + r * a.ld_dq + 32) = u1;
    }
  }
}

int launch_cross_attn_bwd_mma(const void* q, long long ld_q, const void* kv, long long ld_kv, int k_col0, int v_col0,
                              const void* d_o, long long ld_do, const void* dp_tok, const int* slot_col, int n_slots,
                              void* dq, long long ld_dq, long long s_q, int s_k, int heads, int nz, int kv_z_div,
                              float scale, cudaStream_t stream, const int* slot_col_dev) {
  if (s_k < 1 || s_k > 128 || n_slots > 16) return 1;  // n_slots == -1: dp_tok is the dense map
  CrossBwdMmaArgs a;
  a.q = static_cast<const __half*>(q);
  a.kv = static_cast<const __half*>(kv);
  a.d_o = static_cast<const __half*>(d_o);
  a.dp_tok = static_cast<const __half*>(dp_tok);
  a.dq = static_cast<__half*>(dq);
  a.ld_q = ld_q;
  a.ld_kv = ld_kv;
  a.ld_do = ld_do;
  a.ld_dq = ld_dq;
  a.k_col0 = k_col0;
  a.v_col0 = v_col0;
  a.s_q = static_cast<int>(s_q);
  a.s_k = s_k;
  a.heads = heads;
  a.nz = nz;
  a.q_tiles = (a.s_q + 15) / 16;
  a.kv_z_div = kv_z_div;
  a.n_slots = dp_tok ? n_slots : 0;
  for (int i = 0; i < 16; ++i) a.slot_col[i] = (i < a.n_slots && slot_col) ? slot_col[i] : -1;
  a.slot_col_dev = slot_col_dev;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.n_warps = static_cast<long long>(nz) * a.q_tiles * heads;
  const long long blocks = (a.n_warps + 3) / 4;
  LVD_REQUIRE(blocks < (1ll << 31), "lvd_cross_attn_bwd: too many query tiles");
  const unsigned nb = static_cast<unsigned>(blocks);
  const int nt = (s_k + 7) / 8;
  if (nt <= 4) {
    cross_attn_bwd_mma_kernel<4, 2><<<nb, 128, 0, stream>>>(a);
  } else if (nt <= 10) {
    cross_attn_bwd_mma_kernel<10, 5><<<nb, 128, 0, stream>>>(a);
  } else {
    cross_attn_bwd_mma_kernel<16, 8><<<nb, 128, 0, stream>>>(a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

// Packed-sequence temporal attention for frames in {8, 16, 24, 32}; returns 1 when the shape is not covered (the caller
// then uses the tcgen05 tile kernel), LVD_OK after a launch.
int launch_temporal_attn_mma(const lvd_attn_desc* d, cudaStream_t stream) {
  const int F = d->frames;
  if (F != 8 && F != 16 && F != 24 && F != 32) return 1;
  if (d->p_full || d->p_tok || d->lse || d->kv_z_div != 1) return 1;
  TemporalArgs a;
  a.q = static_cast<const __half*>(d->q);
  a.kv = static_cast<const __half*>(d->kv);
  a.out = static_cast<__half*>(d->out);
  a.ld_q = d->ld_q;
  a.ld_kv = d->ld_kv;
  a.ld_o = d->ld_o;
  a.n_warps = static_cast<long long>(d->nz) * d->hw * d->heads;
  a.q_col0 = d->q_col0;
  a.k_col0 = d->k_col0;
  a.v_col0 = d->v_col0;
  a.o_col0 = d->o_col0;
  a.frames = F;
  a.heads = d->heads;
  a.scale_log2 = d->scale * 1.4426950408889634f;
  const long long blocks = (a.n_warps + 3) / 4;
  LVD_REQUIRE(blocks < (1ll << 31), "lvd_attention: too many sequences");
  const unsigned nb = static_cast<unsigned>(blocks);
  if (F == 8) {
    temporal_attn_mma_kernel<1, 1, 1><<<nb, 128, 0, stream>>>(a);
  } else if (F == 16) {
    temporal_attn_mma_kernel<1, 2, 1><<<nb, 128, 0, stream>>>(a);
  } else if (F == 24) {
    temporal_attn_mma_kernel<2, 3, 2><<<nb, 128, 0, stream>>>(a);
  } else {
    temporal_attn_mma_kernel<2, 4, 2><<<nb, 128, 0, stream>>>(a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

}  // namespace lvd
