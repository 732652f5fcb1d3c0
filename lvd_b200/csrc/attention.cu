// attention.cu -- softmax(Q K^T * scale) V on tcgen05 tensor cores with TMA-staged operands.
//
// One CTA = one (query tile of <=128 rows, head, batch index z).  Warp roles:
//   warp 0     : TMA producer (Q once; K/V 128-row units through a 2-stage ring)
//   warp 1     : MMA issuer   (S = Q K^T into TMEM cols [0,128); O_unit = P V into TMEM cols [128,192))
//   warps 2..5 : softmax      (one query row per thread: tcgen05.ld S -> exp2 -> fp16 P into swizzled smem ->
//                              accumulate / rescale O in registers -> fp16 store)
// Three shapes of the reference run through it (models/attention_processor.py:344-430 fast path, :432-589 slow path):
//   spatial self-attention  (S up to 2880+30, online softmax over several 128-column units)
//   cross-attention         (77 text tokens: ONE unit -> exact softmax, fp16 probabilities feed P V exactly like
//                            get_attention_scores()+bmm (:222-258,:536); optional emission of the probability map
//                            or of selected token columns for the box-energy guidance)
//   temporal self-attention (S = F frames; floor(128/F) pixels' sequences are packed in one tile with a
//                            block-diagonal mask; rows are gathered straight from the channels-last activation)
#include <atomic>
#include <cstdlib>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

constexpr int kTile = 128 * 128;  // bytes of a 128-row x 64-half operand tile
constexpr int kAttnThreads = 192;
constexpr int kAttnSmem = kTile /*Q*/ + 4 * kTile /*K,V x2*/ + 2 * kTile /*P*/ + 128;

struct AttnArgs {
  CUtensorMap map_q;
  CUtensorMap map_kv;
  long long o_d1_stride, o_d2_stride, o_z_stride, ld_o;
  long long p_rows;
  __half* out;
  __half* p_full;
  __half* p_tok;
  int q_col0, k_col0, v_col0, o_col0;
  int qb1, qb2, q_dim1;
  int q_bytes, kv_bytes;
  int kv_same_tile, n_units, kv_total, diag_mod, kv_z_div, heads, nz;
  int n_tok;
  int tok_cols[16];
  const int* tok_cols_dev;  // optional: the same list in device memory (read at run time: CUDA-graph replays)
  int last_cols16;
  int seq_len;  // > 0: packed-sequence (temporal) mode
  float* lse;   // optional [nz, heads, p_rows] log2-domain logsumexp (multi-unit mode), for the backward kernels
  float scale_log2;
  uint32_t idesc_qk_full, idesc_qk_last, idesc_pv;
};

__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;      // 2 stages
  uint8_t* sV = smem + 3 * kTile;  // 2 stages
  uint8_t* sP = smem + 5 * kTile;  // 2 column blocks of 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  // heads vary fastest across concurrently scheduled CTAs so that neighbouring CTAs read neighbouring 128-byte
  // pieces of the same activation rows (DRAM page locality; matters most for the temporal row gather)
  const int h = blockIdx.x, qt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (g.q_bytes < kTile || g.kv_bytes < kTile || g.seq_len > 0) {
    // partial boxes (temporal packing): rows the TMA never writes must hold finite values (0 * NaN = NaN in P V)
    uint4 zero = make_uint4(0, 0, 0, 0);
    const int nfill = (g.seq_len > 0 ? 7 : 5) * kTile / 16;  // packed mode also needs zeros in P outside the windows
    for (int i = threadIdx.x; i < nfill; i += kAttnThreads) reinterpret_cast<uint4*>(smem)[i] = zero;
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_q);
    tma_prefetch_desc(&g.map_kv);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
  const int zk = z / g.kv_z_div;

  // producer and MMA warps: every lane walks the loop, one elected lane issues -- under `if (lane == 0)` each
  // UTMALDG / UTCHMMA is wrapped in an ELECT / R2UR.BROADCAST retry loop (~45 ns per tcgen05.mma; see ptx.cuh)
  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, static_cast<uint32_t>(g.q_bytes));
      tma_load_4d(sQ, &g.map_q, q_full, g.q_col0 + h * 64, qt * g.qb1, 0, z);
    }
    __syncwarp();
    const int n_units = g.n_units;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      mbar_wait(&kv_empty[st], ((u >> 1) & 1) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[st], static_cast<uint32_t>(2 * g.kv_bytes));
        const int c1 = g.kv_same_tile ? qt * g.qb1 : u * 128;
        tma_load_4d(sK + st * kTile, &g.map_kv, &kv_full[st], g.k_col0 + h * 64, c1, 0, zk);
        tma_load_4d(sV + st * kTile, &g.map_kv, &kv_full[st], g.v_col0 + h * 64, c1, 0, zk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mbar_wait(q_full, 0);
    // descriptors: constant upper word (SBO 1024 B, version 1, 128-byte swizzle), lower word = address >> 4 | LBO >> 4 << 16
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t q_lo = ((smem_u32(sQ) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t p_lo = ((smem_u32(sP) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(sK) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t v_lo0 = ((smem_u32(sV) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);  // V is MN-major: [kv][d]
    const int n_units = g.n_units;
    const uint32_t idesc_pv = g.idesc_pv;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      const bool last = (u == n_units - 1);
      const int ksteps = (last ? g.last_cols16 : 128) / 16;
      const uint32_t idesc_qk = last ? g.idesc_qk_last : g.idesc_qk_full;
      mbar_wait(&kv_full[st], (u >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t k_lo = k_lo0 + static_cast<uint32_t>(st) * (kTile >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_s, (static_cast<uint64_t>(desc_hi) << 32) | (q_lo + 2u * k),
                   (static_cast<uint64_t>(desc_hi) << 32) | (k_lo + 2u * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, u & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t v_lo = v_lo0 + static_cast<uint32_t>(st) * (kTile >> 4);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) |
                              (p_lo + static_cast<uint32_t>((ks >> 2) * (kTile >> 4) + (ks & 3) * 2));
          const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (v_lo + static_cast<uint32_t>(ks) * (2048u >> 4));
          umma_f16(tmem_o, da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int j1 = r % g.qb1, j2 = r / g.qb1;
    const int c1 = qt * g.qb1 + j1;
    const bool row_valid = (j2 < g.qb2) && (c1 < g.q_dim1);
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float sc = g.scale_log2;
    uint8_t* p_row = sP + r * 128;
    __half* dst = g.out + (static_cast<long long>(z) * g.o_z_stride + static_cast<long long>(j2) * g.o_d2_stride +
                           static_cast<long long>(c1) * g.o_d1_stride) *
                              g.ld_o +
                  g.o_col0 + h * 64;

    if (g.seq_len > 0) {
      // ---------------- packed sequences (temporal attention): rows [i*F, (i+1)*F) of the tile form sequence i.
      // Only the F columns of the row's own sequence matter, so each warp loads the union window of its 32 rows
      // (<= 64 columns) once, keeps it in registers, and writes exact fp16 probabilities; everything outside the
      // window stays zero from the initial fill of sP.
      const int F = g.seq_len;
      const int wbase = (q * 32) / F * F;
      int wend = ((q * 32 + 31) / F + 1) * F;
      if (wend > 128) wend = 128;
      const int nch = (wend - wbase + 7) >> 3;
      const int mystart = (r / F) * F;
      mbar_wait(s_full, 0);
      tc_fence_after();
      float sv[8][8];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < nch) {
          uint32_t rr[8];
          tmem_ld8(tmem_s + lane_off + wbase + k * 8, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = wbase + k * 8 + i;
            const bool ok = (c >= mystart) && (c < mystart + F) && (c < g.kv_total);
            sv[k][i] = ok ? __uint_as_float(rr[i]) * sc : -INFINITY;
            mx = fmaxf(mx, sv[k][i]);
          }
        }
      }
      if (!(mx > -INFINITY)) mx = 0.f;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < nch) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sv[k][i] = fast_exp2(sv[k][i] - mx);
            sum += sv[k][i];
          }
        }
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < nch) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(sv[k][2 * i] * inv, sv[k][2 * i + 1] * inv);
          const int col = wbase + k * 8;
          *reinterpret_cast<uint4*>(p_row + (col >> 6) * kTile + ((((col & 63) >> 3) ^ (r & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, 0);
      tc_fence_after();
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t rr[32];
        tmem_ld32(tmem_o + lane_off + hlf * 32, rr);
        tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint4 v;
            __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              hh[i] = __floats2half2_rn(__uint_as_float(rr[s8 * 8 + 2 * i]), __uint_as_float(rr[s8 * 8 + 2 * i + 1]));
            *reinterpret_cast<uint4*>(dst + hlf * 32 + s8 * 8) = v;
          }
        }
      }
      tc_fence_before();
    } else {
      const int diag = g.diag_mod;
      const int rmod = diag ? (r % diag) : 0;
      const bool single = (g.n_units == 1);
      float m_run = -INFINITY, l_run = 0.f, l_single = 1.f;
      float o[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) o[i] = 0.f;

      for (int u = 0; u < g.n_units; ++u) {
        const bool last = (u == g.n_units - 1);
        const int ncols = last ? g.last_cols16 : 128;
        const int nchunks = (ncols + 31) / 32;
        int nvalid = g.kv_total - u * 128;
        if (nvalid > 128) nvalid = 128;
        const bool full = (nvalid == 128) && (diag == 0) && !single;  // no masking needed anywhere in this unit
        mbar_wait(s_full, u & 1);
        tc_fence_after();
        // ---- pass 1: row maximum
        float mx = -INFINITY;
        for (int ch = 0; ch < nchunks; ++ch) {
          uint32_t rr[32];
          tmem_ld32(tmem_s + lane_off + ch * 32, rr);
          tmem_ld_wait();
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rr[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = ch * 32 + i;
              const bool ok = (c < nvalid) && (diag == 0 || (c % diag) == rmod);
              mx = fmaxf(mx, ok ? __uint_as_float(rr[i]) : -INFINITY);
            }
          }
        }
        mx *= sc;  // sc > 0
        float m_new = fmaxf(m_run, mx);
        if (!(m_new > -INFINITY)) m_new = 0.f;  // fully masked / garbage row: keep arithmetic finite
        float inv = 1.f;
        if (single) {
          // ---- pass 2 (single-unit modes): exact denominator, so that the fp16 probabilities equal the reference's
          float sum = 0.f;
          for (int ch = 0; ch < nchunks; ++ch) {
            uint32_t rr[32];
            tmem_ld32(tmem_s + lane_off + ch * 32, rr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = ch * 32 + i;
              const bool ok = (c < nvalid) && (diag == 0 || (c % diag) == rmod);
              sum += ok ? fast_exp2(fmaf(__uint_as_float(rr[i]), sc, -m_new)) : 0.f;
            }
          }
          inv = 1.f / sum;
          l_single = sum;
        }
        const float alpha = fast_exp2(m_run - m_new);
        float psum = 0.f;
        // ---- pass 3: probabilities -> fp16 -> swizzled smem (A operand of P V)
        for (int ch = 0; ch < nchunks; ++ch) {
          uint32_t rr[32];
          tmem_ld32(tmem_s + lane_off + ch * 32, rr);
          tmem_ld_wait();
          float p[32];
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              p[i] = fast_exp2(fmaf(__uint_as_float(rr[i]), sc, -m_new));
              psum += p[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = ch * 32 + i;
              const bool ok = (c < nvalid) && (diag == 0 || (c % diag) == rmod);
              const float e = ok ? fast_exp2(fmaf(__uint_as_float(rr[i]), sc, -m_new)) * inv : 0.f;
              p[i] = e;
              psum += e;
            }
          }
          uint8_t* blk = p_row + (ch >> 1) * kTile;
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint4 v;
            __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(p[s8 * 8 + 2 * i], p[s8 * 8 + 2 * i + 1]);
            const int chunk16 = ((ch & 1) * 4 + s8) ^ (r & 7);
            *reinterpret_cast<uint4*>(blk + chunk16 * 16) = v;
          }
          if (single && row_valid) {
            if (g.p_full) {
              __half* pd = g.p_full + ((static_cast<long long>(z) * g.heads + h) * g.p_rows + c1) * g.kv_total;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (ch * 32 + i < g.kv_total) pd[ch * 32 + i] = __float2half_rn(p[i]);
            }
            if (g.p_tok) {
              for (int t = 0; t < g.n_tok; ++t) {
                const int col = (g.tok_cols_dev ? g.tok_cols_dev[t] : g.tok_cols[t]) - ch * 32;
                if (col >= 0 && col < 32) {
                  float pv = 0.f;
#pragma unroll
                  for (int i = 0; i < 32; ++i) pv = (i == col) ? p[i] : pv;
                  g.p_tok[((static_cast<long long>(t) * g.nz + z) * g.heads + h) * g.p_rows + c1] =
                      __float2half_rn(pv);
                }
              }
            }
          }
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
        // ---- O accumulation
        mbar_wait(o_full, u & 1);
        tc_fence_after();
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t rr[32];
          tmem_ld32(tmem_o + lane_off + hlf * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[hlf * 32 + i] = fmaf(o[hlf * 32 + i], alpha, __uint_as_float(rr[i]));
        }
        tc_fence_before();
      }
      if (row_valid) {
        const float fin = single ? 1.f : 1.f / l_run;
        if (g.lse)
          g.lse[(static_cast<long long>(z) * g.heads + h) * g.p_rows + c1] = m_run + log2f(single ? l_single : l_run);
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hh[i] = __floats2half2_rn(o[s8 * 8 + 2 * i] * fin, o[s8 * 8 + 2 * i + 1] * fin);
          *reinterpret_cast<uint4*>(dst + s8 * 8) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ spatial self-attention
// Multi-unit (s_k > 128) softmax(Q K^T) V, the UNet's spatial self-attention (hw = 2880 / 720 / 180 keys).  Same
// TMA producer / single-thread MMA issuer as attention_kernel, but the softmax side is restructured around what bounded
// that kernel (4 softmax warps per CTA, a serial S -> softmax -> P V -> read O chain per 128-key unit, O accumulated
// in registers: tensor pipe 17 % busy):
//   * EIGHT softmax warps: two threads per score row, each owning one 64-key half of the unit (TMEM lanes are tied to
//     warp_id % 4, columns are free) -> 4 softmax warps per SM sub-partition with two resident CTAs;
//   * each half runs its OWN online softmax (m, l) and its own accumulator O_a / O_b in TMEM (keys 0-63 / 64-127 of every
//     unit: a split over the key axis, combined exactly at the end), so the two threads of a row never exchange
//     anything inside the loop;
//   * O stays in TMEM: the P V product ACCUMULATES into it across units.  The reference point m of the exponentials is
//     only moved when the running maximum grows by more than 2^8 (probabilities then stay <= 256, exact in fp16 to the
//     same 11 bits); only then is the accumulator rescaled in place (tcgen05.ld / st).  After the first unit this is
//     rare, so the per-unit chain is S -> two TMEM passes (max, exp2 -> fp16 -> smem) -> P V, with no accumulator
//     read-back and no register-resident O.
// Ordering notes: the tensor pipe executes in issue order and tcgen05.commit covers everything issued before it, so
// s_full(u) also tells the softmax warps that P V of unit u-1 has finished reading the P tile and updating O.
constexpr int kSelfThreads = 320;
constexpr float kRescaleLog2 = 8.0f;

__global__ void __launch_bounds__(kSelfThreads, 2) attention_self_kernel(const __grid_constant__ AttnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;      // 2 stages
  uint8_t* sV = smem + 3 * kTile;  // 2 stages
  uint8_t* sP = smem + 5 * kTile;  // 2 column blocks of 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const int h = blockIdx.x, qt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (g.q_bytes < kTile || g.kv_bytes < kTile) {
    uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 5 * kTile / 16; i += kSelfThreads) reinterpret_cast<uint4*>(smem)[i] = zero;
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_q);
    tma_prefetch_desc(&g.map_kv);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;  // O_a at +128, O_b at +192
  const int zk = z / g.kv_z_div;
  const int n_units = g.n_units;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, static_cast<uint32_t>(g.q_bytes));
      tma_load_4d(sQ, &g.map_q, q_full, g.q_col0 + h * 64, qt * g.qb1, 0, z);
    }
    __syncwarp();
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      mbar_wait(&kv_empty[st], ((u >> 1) & 1) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[st], static_cast<uint32_t>(2 * g.kv_bytes));
        tma_load_4d(sK + st * kTile, &g.map_kv, &kv_full[st], g.k_col0 + h * 64, u * 128, 0, zk);
        tma_load_4d(sV + st * kTile, &g.map_kv, &kv_full[st], g.v_col0 + h * 64, u * 128, 0, zk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mbar_wait(q_full, 0);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t q_lo = ((smem_u32(sQ) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t p_lo = ((smem_u32(sP) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(sK) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t v_lo0 = ((smem_u32(sV) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);  // V is MN-major: [kv][d]
    const uint32_t idesc_pv = g.idesc_pv;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      const bool last = (u == n_units - 1);
      const int ksteps = (last ? g.last_cols16 : 128) / 16;
      const uint32_t idesc_qk = last ? g.idesc_qk_last : g.idesc_qk_full;
      mbar_wait(&kv_full[st], (u >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t k_lo = k_lo0 + static_cast<uint32_t>(st) * (kTile >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_s, (static_cast<uint64_t>(desc_hi) << 32) | (q_lo + 2u * k),
                   (static_cast<uint64_t>(desc_hi) << 32) | (k_lo + 2u * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, u & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t v_lo = v_lo0 + static_cast<uint32_t>(st) * (kTile >> 4);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) |
                              (p_lo + static_cast<uint32_t>((ks >> 2) * (kTile >> 4) + (ks & 3) * 2));
          const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (v_lo + static_cast<uint32_t>(ks) * (2048u >> 4));
          // keys 0-63 of the unit accumulate into O_a, keys 64-127 into O_b; the first product of each overwrites
          umma_f16(tmem_o + static_cast<uint32_t>(ks >> 2) * 64u, da, db, idesc_pv,
                   (u != 0 || (ks & 3) != 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // key half of every unit / accumulator owned by this thread
    const int r = q * 32 + lane;
    const int c1 = qt * g.qb1 + r;
    const bool row_valid = (r < g.qb1) && (c1 < g.q_dim1);
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t my_s = tmem_s + lane_off + static_cast<uint32_t>(half) * 64u;
    const uint32_t my_o = tmem_o + lane_off + static_cast<uint32_t>(half) * 64u;
    const float sc = g.scale_log2;
    uint8_t* p_blk = sP + half * kTile + r * 128;
    float m_used = -INFINITY, l_run = 0.f;

    for (int u = 0; u < n_units; ++u) {
      int nvalid = g.kv_total - u * 128 - half * 64;  // valid keys among this thread's 64
      nvalid = nvalid > 64 ? 64 : (nvalid < 0 ? 0 : nvalid);
      mbar_wait(s_full, u & 1);
      tc_fence_after();
      // ---- the thread's 64 scores: both TMEM loads in flight at once, kept in registers for the exponentials (one
      // TMEM read per score; the accumulator no longer lives in registers, so there is room)
      uint32_t ra[32], rb[32];
      tmem_ld32(my_s, ra);
      tmem_ld32(my_s + 32, rb);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (nvalid != 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= nvalid) ra[i] = __float_as_uint(-INFINITY);
          if (32 + i >= nvalid) rb[i] = __float_as_uint(-INFINITY);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmax3(mx, __uint_as_float(ra[i]), __uint_as_float(rb[i]));
      const float m_cand = fmaxf(m_used, mx * sc);  // sc > 0
      const bool move = (m_cand - m_used) > kRescaleLog2;  // first unit: m_used = -inf
      if (u > 0 && __any_sync(0xffffffffu, move)) {
        // rare: move the reference point of the rows that need it and rescale their accumulator rows in TMEM
        // (s_full(u) implies P V of unit u-1 is complete, see the header); 8 columns at a time: few live registers
        const float f = move ? fast_exp2(m_used - m_cand) : 1.f;
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t rr[8];
          tmem_ld8(my_o + c8 * 8, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * f);
          tmem_st8(my_o + c8 * 8, rr);
        }
        tmem_st_wait();
        l_run *= f;
      }
      if (move) m_used = m_cand;
      // ---- probabilities relative to m_used -> fp16 -> swizzled smem (A operand of P V); masked scores are -inf -> 0
      float psum = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = s8 * 8 + 2 * i;
            const float p0 = fast_exp2(fmaf(__uint_as_float(ch ? rb[c] : ra[c]), sc, -m_used));
            const float p1 = fast_exp2(fmaf(__uint_as_float(ch ? rb[c + 1] : ra[c + 1]), sc, -m_used));
            psum += p0 + p1;
            hh[i] = __floats2half2_rn(p0, p1);
          }
          const int chunk16 = (ch * 4 + s8) ^ (r & 7);
          *reinterpret_cast<uint4*>(p_blk + chunk16 * 16) = v;
        }
      }
      l_run += psum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- combine the two key halves: O = (O_a 2^(m_a - m) + O_b 2^(m_b - m)) / (l_a 2^(m_a - m) + l_b 2^(m_b - m))
    mbar_wait(o_full, 0);
    tc_fence_after();
    float2* xch = reinterpret_cast<float2*>(sQ);  // Q / K / V / P tiles are dead: every MMA has completed
    xch[half * 128 + r] = make_float2(m_used, l_run);
    named_bar_sync(1 + q, 64);
    const float2 other = xch[(half ^ 1) * 128 + r];
    const float m_a = half ? other.x : m_used, l_a = half ? other.y : l_run;
    const float m_b = half ? m_used : other.x, l_b = half ? l_run : other.y;
    const float m_all = fmaxf(m_a, m_b);
    const float fa = fast_exp2(m_a - m_all), fb = fast_exp2(m_b - m_all);
    const float l_all = l_a * fa + l_b * fb;
    const float inv = 1.f / l_all;
    uint32_t ra[32], rb[32];
    tmem_ld32(tmem_o + lane_off + half * 32, ra);       // O_a, output columns [32 half, 32 half + 32)
    tmem_ld32(tmem_o + lane_off + 64 + half * 32, rb);  // O_b, same columns
    tmem_ld_wait();
    if (row_valid) {
      __half* dst = g.out + (static_cast<long long>(z) * g.o_z_stride + static_cast<long long>(c1) * g.o_d1_stride) * g.ld_o +
                    g.o_col0 + h * 64 + half * 32;
      const float wa = fa * inv, wb = fb * inv;
#pragma unroll
      for (int s8 = 0; s8 < 4; ++s8) {
        uint4 v;
        __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = s8 * 8 + 2 * i;
          hh[i] = __floats2half2_rn(fmaf(__uint_as_float(ra[c]), wa, __uint_as_float(rb[c]) * wb),
                                    fmaf(__uint_as_float(ra[c + 1]), wa, __uint_as_float(rb[c + 1]) * wb));
        }
        *reinterpret_cast<uint4*>(dst + s8 * 8) = v;
      }
      if (g.lse && half == 0) g.lse[(static_cast<long long>(z) * g.heads + h) * g.p_rows + c1] = m_all + log2f(l_all);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ ... with two S buffers
// attention_self_kernel leaves the softmax warps idle while the tensor pipe computes P V and the next S (one S buffer: the
// per-CTA chain S -> softmax -> P V -> S is serial; ncu: 5.5 long-scoreboard stall cycles per issue, MUFU pipe 57 %).
// Here a unit is 64 keys, so that TWO score buffers fit next to the two accumulators in the CTA's 256 TMEM columns
// (S0 | S1 | O_a | O_b, 64 columns each): the MMA warp issues S(u+1) before it waits for the probabilities of unit u, the
// softmax warps go from unit to unit without ever waiting for a P V product, and P is double-buffered in shared memory.
// Two threads per row as before, now 32 keys each.  Hazards, all covered by issue order + tcgen05.commit semantics:
//   * S(u+1) overwrites the buffer softmax(u-1) read: it is issued after the MMA warp saw p_full(u-1);
//   * softmax(u+1) overwrites the P buffer P V(u-1) read: it starts after s_full(u+1), and S(u+1) was issued after
//     P V(u-1);
//   * the (rare) accumulator rescale in softmax(u+1) must not overlap P V(u): it first waits for pv_done(u), which the
//     MMA warp commits after every P V; P V(u+1) cannot start before the rescaling thread arrives at p_full(u+1).
constexpr int kKvStages = 4;   // 64-key K and V tiles (8 KB each)
constexpr int kTile64 = 64 * 128;

__global__ void __launch_bounds__(kSelfThreads, 2) attention_self2_kernel(const __grid_constant__ AttnArgs g,
                                                                           const __grid_constant__ CUtensorMap map_kv64) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sQ = smem;                              // 128 x 64 halves
  uint8_t* sK = smem + kTile;                      // kKvStages x (64 keys x 64 halves)
  uint8_t* sV = smem + kTile + kKvStages * kTile64;
  uint8_t* sP = smem + kTile + 2 * kKvStages * kTile64;  // 2 buffers x (128 rows x 64 keys)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [4]
  uint64_t* kv_empty = bars + 5;   // [4]
  uint64_t* s_full = bars + 9;     // [2]
  uint64_t* p_full = bars + 11;    // [2]
  uint64_t* pv_done = bars + 13;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 15);
  const int h = blockIdx.x, qt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_q);
    tma_prefetch_desc(&map_kv64);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;  // S0 +0, S1 +64, O_a +128, O_b +192
  const int zk = z / g.kv_z_div;
  const int n_units = (g.kv_total + 63) / 64;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, static_cast<uint32_t>(kTile));
      tma_load_4d(sQ, &g.map_q, q_full, g.q_col0 + h * 64, qt * 128, 0, z);
    }
    __syncwarp();
    for (int u = 0; u < n_units; ++u) {
      const int st = u & (kKvStages - 1);
      mbar_wait(&kv_empty[st], ((u / kKvStages) & 1) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[st], static_cast<uint32_t>(2 * kTile64));
        tma_load_4d(sK + st * kTile64, &map_kv64, &kv_full[st], g.k_col0 + h * 64, u * 64, 0, zk);
        tma_load_4d(sV + st * kTile64, &map_kv64, &kv_full[st], g.v_col0 + h * 64, u * 64, 0, zk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mbar_wait(q_full, 0);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t q_lo = ((smem_u32(sQ) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t p_lo0 = ((smem_u32(sP) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(sK) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t v_lo0 = ((smem_u32(sV) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);  // V is MN-major: [kv][d]
    const uint32_t idesc_qk = make_idesc_f16(128, 64, 0, 0);
    const uint32_t idesc_pv = g.idesc_pv;
    auto issue_s = [&](int u) {  // S(u) = Q K(u)^T into score buffer u & 1
      const int st = u & (kKvStages - 1);
      mbar_wait(&kv_full[st], (u / kKvStages) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t k_lo = k_lo0 + static_cast<uint32_t>(st) * (kTile64 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_s + static_cast<uint32_t>(u & 1) * 64u, (static_cast<uint64_t>(desc_hi) << 32) | (q_lo + 2u * k),
                   (static_cast<uint64_t>(desc_hi) << 32) | (k_lo + 2u * k), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[u & 1]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int u = 0; u < n_units; ++u) {
      if (u + 1 < n_units) issue_s(u + 1);
      const int st = u & (kKvStages - 1);
      int nkeys = g.kv_total - u * 64;
      nkeys = nkeys > 64 ? 64 : nkeys;
      const int ksteps = (nkeys + 15) / 16;
      mbar_wait(&p_full[u & 1], (u >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t v_lo = v_lo0 + static_cast<uint32_t>(st) * (kTile64 >> 4);
        const uint32_t p_lo = p_lo0 + static_cast<uint32_t>(u & 1) * (kTile >> 4);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (p_lo + static_cast<uint32_t>(ks) * 2u);
          const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (v_lo + static_cast<uint32_t>(ks) * (2048u >> 4));
          // keys 0-31 of the unit accumulate into O_a, keys 32-63 into O_b; the first product of each overwrites
          umma_f16(tmem_o + static_cast<uint32_t>(ks >> 1) * 64u, da, db, idesc_pv,
                   (u != 0 || (ks & 1) != 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&pv_done[u & 1]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // 32-key half of every unit / accumulator owned by this thread
    const int r = q * 32 + lane;
    const int c1 = qt * 128 + r;
    const bool row_valid = c1 < g.q_dim1;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t my_o = tmem_o + lane_off + static_cast<uint32_t>(half) * 64u;
    const float sc = g.scale_log2;
    float m_used = -INFINITY, l_run = 0.f;

    for (int u = 0; u < n_units; ++u) {
      int nvalid = g.kv_total - u * 64 - half * 32;  // valid keys among this thread's 32
      nvalid = nvalid > 32 ? 32 : (nvalid < 0 ? 0 : nvalid);
      mbar_wait(&s_full[u & 1], (u >> 1) & 1);
      tc_fence_after();
      uint32_t rr[32];
      tmem_ld32(tmem_s + lane_off + static_cast<uint32_t>(u & 1) * 64u + static_cast<uint32_t>(half) * 32u, rr);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (nvalid != 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= nvalid) rr[i] = __float_as_uint(-INFINITY);
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(rr[i]), __uint_as_float(rr[i + 1]));
      const float m_cand = fmaxf(m_used, mx * sc);  // sc > 0
      const bool move = (m_cand - m_used) > kRescaleLog2;  // first unit: m_used = -inf
      if (u > 0 && __any_sync(0xffffffffu, move)) {
        // rare: P V(u-1) may still be running (or not yet issued) -- wait for its commit before touching the accumulator
        mbar_wait(&pv_done[(u - 1) & 1], ((u - 1) >> 1) & 1);
        tc_fence_after();
        const float f = move ? fast_exp2(m_used - m_cand) : 1.f;
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t t8[8];
          tmem_ld8(my_o + c8 * 8, t8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = __float_as_uint(__uint_as_float(t8[i]) * f);
          tmem_st8(my_o + c8 * 8, t8);
        }
        tmem_st_wait();
        l_run *= f;
      }
      if (move) m_used = m_cand;
      float psum = 0.f;
      uint8_t* p_row = sP + (u & 1) * kTile + r * 128;
#pragma unroll
      for (int s8 = 0; s8 < 4; ++s8) {
        uint4 v;
        __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = s8 * 8 + 2 * i;
          const float p0 = fast_exp2(fmaf(__uint_as_float(rr[c]), sc, -m_used));
          const float p1 = fast_exp2(fmaf(__uint_as_float(rr[c + 1]), sc, -m_used));
          psum += p0 + p1;
          hh[i] = __floats2half2_rn(p0, p1);
        }
        const int chunk16 = (half * 4 + s8) ^ (r & 7);
        *reinterpret_cast<uint4*>(p_row + chunk16 * 16) = v;
      }
      l_run += psum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[u & 1]);
    }
    // ---- combine the two key halves: O = (O_a 2^(m_a - m) + O_b 2^(m_b - m)) / (l_a 2^(m_a - m) + l_b 2^(m_b - m))
    mbar_wait(&pv_done[(n_units - 1) & 1], ((n_units - 1) >> 1) & 1);
    tc_fence_after();
    float2* xch = reinterpret_cast<float2*>(sQ);  // Q / K / V / P tiles are dead: every MMA has completed
    xch[half * 128 + r] = make_float2(m_used, l_run);
    named_bar_sync(1 + q, 64);
    const float2 other = xch[(half ^ 1) * 128 + r];
    const float m_a = half ? other.x : m_used, l_a = half ? other.y : l_run;
    const float m_b = half ? m_used : other.x, l_b = half ? l_run : other.y;
    const float m_all = fmaxf(m_a, m_b);
    const float fa = fast_exp2(m_a - m_all), fb = fast_exp2(m_b - m_all);
    const float l_all = l_a * fa + l_b * fb;
    const float inv = 1.f / l_all;
    uint32_t ra[32], rb[32];
    tmem_ld32(tmem_o + lane_off + half * 32, ra);       // O_a, output columns [32 half, 32 half + 32)
    tmem_ld32(tmem_o + lane_off + 64 + half * 32, rb);  // O_b, same columns
    tmem_ld_wait();
    if (row_valid) {
      __half* dst = g.out + (static_cast<long long>(z) * g.o_z_stride + static_cast<long long>(c1) * g.o_d1_stride) * g.ld_o +
                    g.o_col0 + h * 64 + half * 32;
      const float wa = fa * inv, wb = fb * inv;
#pragma unroll
      for (int s8 = 0; s8 < 4; ++s8) {
        uint4 v;
        __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = s8 * 8 + 2 * i;
          hh[i] = __floats2half2_rn(fmaf(__uint_as_float(ra[c]), wa, __uint_as_float(rb[c]) * wb),
                                    fmaf(__uint_as_float(ra[c + 1]), wa, __uint_as_float(rb[c + 1]) * wb));
        }
        *reinterpret_cast<uint4*>(dst + s8 * 8) = v;
      }
      if (g.lse && half == 0) g.lse[(static_cast<long long>(z) * g.heads + h) * g.p_rows + c1] = m_all + log2f(l_all);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_temporal_attn_mma(const lvd_attn_desc* d, cudaStream_t stream);  // attention_temporal.cu
int launch_cross_attn_mma(const lvd_attn_desc* d, cudaStream_t stream);     // attention_temporal.cu

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_attention(const lvd_attn_desc* d, void* stream) {
  LVD_REQUIRE(d && d->q && d->kv && d->out, "lvd_attention: null pointer");
  LVD_REQUIRE(d->heads >= 1 && d->nz >= 1 && d->kv_z_div >= 1, "lvd_attention: bad heads/nz");
  LVD_REQUIRE(d->ld_q % 8 == 0 && d->ld_kv % 8 == 0 && d->ld_o % 8 == 0, "lvd_attention: ld must be multiples of 8");
  LVD_REQUIRE(d->q_col0 % 8 == 0 && d->k_col0 % 8 == 0 && d->v_col0 % 8 == 0 && d->o_col0 % 8 == 0,
              "lvd_attention: column offsets must be multiples of 8");
  LVD_REQUIRE(d->n_tok >= 0 && d->n_tok <= 16, "lvd_attention: n_tok=%d", d->n_tok);
  AttnArgs g;
  memset(&g, 0, sizeof(g));
  int rc;
  const int64_t width_q = d->q_width, width_kv = d->kv_width;
  if (d->mode == 0) {
    // spatial / cross: Q [nz][s_q] rows, KV [nz/kv_z_div][s_k] rows
    LVD_REQUIRE(d->s_q >= 1 && d->s_k >= 1, "lvd_attention: empty sequence");
    if (d->s_k <= 128 && !d->p_full && !d->p_tok && !d->lse) {
      // <= 128 keys (the 77 text tokens; the 45-pixel mid block) and no map requested: warp-level kernel
      // (attention_temporal.cu)
      static const bool force_tile = getenv("LVD_CROSS_TILE") != nullptr;  // A/B switch for tools/
      if (!force_tile) {
        const int r = launch_cross_attn_mma(d, static_cast<cudaStream_t>(stream));
        if (r != 1) return r;
      }
    }
    LVD_REQUIRE(!d->causal, "lvd_attention: causal masking needs s_k <= 128 and no probability / lse outputs");
    int64_t qd[4] = {width_q, d->s_q, 1, d->nz};
    int64_t qs[4] = {1, d->ld_q, d->ld_q * d->s_q, d->ld_q * d->s_q};
    int qb[4] = {64, d->s_q < 128 ? (int)d->s_q : 128, 1, 1};
    rc = encode_tmap_f16_sw128(&g.map_q, d->q, 4, qd, qs, qb);
    if (rc) return rc;
    const int64_t nzk = (d->nz + d->kv_z_div - 1) / d->kv_z_div;
    int64_t kd[4] = {width_kv, d->s_k, 1, nzk};
    int64_t ks[4] = {1, d->ld_kv, d->ld_kv * d->s_k, d->ld_kv * d->s_k};
    int kb[4] = {64, d->s_k < 128 ? (int)d->s_k : 128, 1, 1};
    rc = encode_tmap_f16_sw128(&g.map_kv, d->kv, 4, kd, ks, kb);
    if (rc) return rc;
    g.qb1 = qb[1];
    g.qb2 = 1;
    g.q_dim1 = static_cast<int>(d->s_q);
    g.q_bytes = qb[1] * 128;
    g.kv_bytes = kb[1] * 128;
    g.kv_same_tile = 0;
    g.n_units = static_cast<int>((d->s_k + 127) / 128);
    g.kv_total = static_cast<int>(d->s_k);
    g.diag_mod = 0;
    g.o_d1_stride = 1;
    g.o_d2_stride = 0;
    g.o_z_stride = d->s_q;
    g.p_rows = d->s_q;
  } else if (d->mode == 2) {
    LVD_REQUIRE(!d->causal, "lvd_attention: causal masking is a mode-0 option");
    // packed sequences: rows (z*hw + p)*F + f -- every F consecutive rows are one sequence; bp sequences per tile
    const int F = d->frames;
    LVD_REQUIRE(F >= 8 && F <= 32 && F % 8 == 0 && d->hw >= 1,
                "lvd_attention: packed mode covers frames in {8, 16, 24, 32} (got %d); use mode 1", F);
    {  // F <= 32: one warp per (sequence, head) on mma.sync (attention_temporal.cu); the tile kernel covers the rest
      static const bool force_tile = getenv("LVD_TEMPORAL_TILE") != nullptr;  // A/B switch for tools/
      if (!force_tile) {
        const int r = launch_temporal_attn_mma(d, static_cast<cudaStream_t>(stream));
        if (r != 1) return r;
      }
    }
    int bp = 128 / F;
    if (bp > d->hw) bp = static_cast<int>(d->hw);
    const int64_t rows_z = d->hw * F;
    int64_t qd[4] = {width_q, rows_z, 1, d->nz};
    int64_t qs[4] = {1, d->ld_q, d->ld_q * rows_z, d->ld_q * rows_z};
    int qb[4] = {64, bp * F, 1, 1};
    rc = encode_tmap_f16_sw128(&g.map_q, d->q, 4, qd, qs, qb);
    if (rc) return rc;
    int64_t kd[4] = {width_kv, rows_z, 1, d->nz};
    int64_t ks[4] = {1, d->ld_kv, d->ld_kv * rows_z, d->ld_kv * rows_z};
    rc = encode_tmap_f16_sw128(&g.map_kv, d->kv, 4, kd, ks, qb);
    if (rc) return rc;
    g.qb1 = bp * F;
    g.qb2 = 1;
    g.q_dim1 = static_cast<int>(rows_z);
    g.q_bytes = bp * F * 128;
    g.kv_bytes = bp * F * 128;
    g.kv_same_tile = 1;
    g.n_units = 1;
    g.kv_total = bp * F;
    g.diag_mod = 0;
    g.seq_len = F;
    g.o_d1_stride = 1;
    g.o_d2_stride = 0;
    g.o_z_stride = rows_z;
    g.p_rows = rows_z;
    LVD_REQUIRE(d->kv_z_div == 1 && !d->p_full && !d->p_tok, "lvd_attention: packed mode has no kv broadcast / maps");
  } else {
    // temporal: rows ((b*F + f)*hw + p); sequence = the F frames of pixel p; bp pixels per tile
    const int F = d->frames;
    LVD_REQUIRE(F >= 1 && F <= 128 && d->hw >= 1, "lvd_attention: frames=%d unsupported", F);
    int bp = 128 / F;
    if (bp > d->hw) bp = static_cast<int>(d->hw);
    int64_t qd[4] = {width_q, d->hw, F, d->nz};
    int64_t qs[4] = {1, d->ld_q, d->ld_q * d->hw, d->ld_q * d->hw * F};
    int qb[4] = {64, bp, F, 1};
    rc = encode_tmap_f16_sw128(&g.map_q, d->q, 4, qd, qs, qb);
    if (rc) return rc;
    int64_t kd[4] = {width_kv, d->hw, F, d->nz};
    int64_t ks[4] = {1, d->ld_kv, d->ld_kv * d->hw, d->ld_kv * d->hw * F};
    rc = encode_tmap_f16_sw128(&g.map_kv, d->kv, 4, kd, ks, qb);
    if (rc) return rc;
    g.qb1 = bp;
    g.qb2 = F;
    g.q_dim1 = static_cast<int>(d->hw);
    g.q_bytes = bp * F * 128;
    g.kv_bytes = bp * F * 128;
    g.kv_same_tile = 1;
    g.n_units = 1;
    g.kv_total = bp * F;
    g.diag_mod = bp;
    g.o_d1_stride = 1;
    g.o_d2_stride = d->hw;
    g.o_z_stride = static_cast<long long>(F) * d->hw;
    g.p_rows = d->hw;
    LVD_REQUIRE(d->kv_z_div == 1, "lvd_attention: temporal mode has no kv broadcast");
  }
  LVD_REQUIRE((!d->p_full && !d->p_tok) || g.n_units == 1, "lvd_attention: probability output needs s_k <= 128");
  const int last_valid = g.kv_total - (g.n_units - 1) * 128;
  g.last_cols16 = (last_valid + 15) / 16 * 16;
  g.q_col0 = d->q_col0;
  g.k_col0 = d->k_col0;
  g.v_col0 = d->v_col0;
  g.o_col0 = d->o_col0;
  g.kv_z_div = d->kv_z_div;
  g.heads = d->heads;
  g.nz = d->nz;
  g.ld_o = d->ld_o;
  g.out = static_cast<__half*>(d->out);
  g.p_full = static_cast<__half*>(d->p_full);
  g.p_tok = static_cast<__half*>(d->p_tok);
  g.lse = static_cast<float*>(d->lse);
  LVD_REQUIRE(!d->lse || g.n_units >= 1, "lvd_attention: lse");
  g.n_tok = d->n_tok;
  for (int i = 0; i < d->n_tok; ++i) g.tok_cols[i] = d->tok_cols[i];
  g.tok_cols_dev = d->tok_cols_dev;
  g.scale_log2 = d->scale * 1.4426950408889634f;
  g.idesc_qk_full = make_idesc_f16(128, 128, 0, 0);
  g.idesc_qk_last = make_idesc_f16(128, static_cast<uint32_t>(g.last_cols16), 0, 0);
  g.idesc_pv = make_idesc_f16(128, 64, 0, 1);
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr = true;
  }
  const int q_tiles = (g.q_dim1 + g.qb1 - 1) / g.qb1;
  dim3 grid(d->heads, q_tiles, d->nz);
  LVD_REQUIRE(q_tiles <= 65535 && d->nz <= 65535, "lvd_attention: grid too large");
  static const bool old_self = getenv("LVD_ATTN_SELF_OLD") != nullptr;  // A/B switches for tools/ and tests
  static const bool one_buf = getenv("LVD_ATTN_SELF_1BUF") != nullptr;
  if (d->mode == 0 && g.n_units >= 2 && !d->p_full && !d->p_tok && !old_self && !one_buf && d->s_q >= 128) {
    // two score buffers, 64-key units (attention_self2_kernel): needs its own K/V tensor map with 64-row boxes
    CUtensorMap map64;
    const int64_t nzk = (d->nz + d->kv_z_div - 1) / d->kv_z_div;
    int64_t kd[4] = {width_kv, d->s_k, 1, nzk};
    int64_t ks[4] = {1, d->ld_kv, d->ld_kv * d->s_k, d->ld_kv * d->s_k};
    int kb[4] = {64, 64, 1, 1};
    rc = encode_tmap_f16_sw128(&map64, d->kv, 4, kd, ks, kb);
    if (rc) return rc;
    static bool attr3 = false;
    if (!attr3) {
      LVD_CHECK_CUDA(cudaFuncSetAttribute(attention_self2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
      attr3 = true;
    }
    attention_self2_kernel<<<grid, kSelfThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(g, map64);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LVD_CHECK_CUDA(cudaGetLastError());
    return LVD_OK;
  }
  if (d->mode == 0 && g.n_units >= 2 && !d->p_full && !d->p_tok && !old_self) {
    static bool attr2 = false;
    if (!attr2) {
      LVD_CHECK_CUDA(cudaFuncSetAttribute(attention_self_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
      attr2 = true;
    }
    attention_self_kernel<<<grid, kSelfThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(g);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LVD_CHECK_CUDA(cudaGetLastError());
    return LVD_OK;
  }
  attention_kernel<<<grid, kAttnThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
