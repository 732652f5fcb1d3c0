// attention_bwd.cu -- adjoint of spatial self-attention on tcgen05 tensor cores (flash-style: probabilities are
// recomputed from Q, K and the saved log-sum-exp; nothing of size S x S is ever stored).
//
//   dP = dO V^T ;  dS = P o (dP - D),  D = rowsum(dO o O) ;  dQ = scale dS K ;  dK = scale dS^T Q ;  dV = P^T dO
//
// Two kernels, both one CTA per (128-row tile, head, image): warp 0 TMA, warp 1 MMA issue, warps 2..9 compute -- two
// threads per score row, each owning 64 of the unit's 128 key columns (TMEM lanes are tied to warp_id % 4, columns are
// free; the log-sum-exp is known, so the halves are independent):
//   attn_bwd_dq_kernel   : fixed query tile, loops over key units;   dQ accumulates in TMEM across units.
//   attn_bwd_dkdv_kernel : fixed key tile,   loops over query units; dK, dV accumulate in TMEM across units.  The
//                          P / dS tiles are written once in shared memory as [query][key] and consumed by the
//                          tensor core as MN-major A operands (P^T, dS^T) -- no transposition pass.
// Gradients are carried scaled (grad_scale, see lvd_box_energy) so that fp16 tiles do not underflow.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {
extern std::atomic<long long> g_launches;

constexpr int kT = 128 * 128;  // bytes of a 128 x 64 fp16 tile
constexpr int kBwdThreads = 320;  // 2 control warps + 8 compute warps: two threads per score row, 64 key columns each

struct AttnBwdArgs {
  CUtensorMap map_qkv;  // [rows, width] of the q|k|v buffer, box 64 x 128
  CUtensorMap map_do;   // [rows, C] of dO, box 64 x 128
  const float* lse;     // [nz, heads, S] log2-domain
  const float* dvec;    // [nz*S, heads]
  __half* d_qkv;
  long long ld_dqkv;
  int q_col0, k_col0, v_col0;
  int S, heads, n_units;
  float scale, scale_log2;
  uint32_t idesc_s_full, idesc_s_last;  // M128 x N(keys) K-major/K-major
  uint32_t idesc_dq;                    // M128 x N64, A K-major, B MN-major
  uint32_t idesc_dkv;                   // M128 x N64, A MN-major, B MN-major
  int last_cols16;
};

__device__ __forceinline__ void store_chunk16(uint8_t* tile_row, int r, int col, const float (&v)[8]) {
  uint4 u;
  __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(tile_row + (col >> 6) * kT + ((((col & 63) >> 3) ^ (r & 7)) << 4)) = u;
}

// ------------------------------------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + kT;
  uint8_t* sK = smem + 2 * kT;   // 2 stages
  uint8_t* sV = smem + 4 * kT;   // 2 stages
  uint8_t* sDS = smem + 6 * kT;  // 2 column blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * kT);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const int h = blockIdx.x, qt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 256);
    mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_qkv);
    tma_prefetch_desc(&g.map_do);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dq = tmem_base + 256;
  const int row0 = z * g.S;  // first row of this image in the row buffers

  // producer / MMA warps: all lanes walk the loops, one elected lane issues (no ELECT/R2UR retry loops, see ptx.cuh)
  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, 2 * kT);
      tma_load_2d(sQ, &g.map_qkv, q_full, g.q_col0 + h * 64, row0 + qt * 128);
      tma_load_2d(sDO, &g.map_do, q_full, h * 64, row0 + qt * 128);
    }
    __syncwarp();
    const int n_units = g.n_units;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      mbar_wait(&kv_empty[st], ((u >> 1) & 1) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * kT);
        tma_load_2d(sK + st * kT, &g.map_qkv, &kv_full[st], g.k_col0 + h * 64, row0 + u * 128);
        tma_load_2d(sV + st * kT, &g.map_qkv, &kv_full[st], g.v_col0 + h * 64, row0 + u * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mbar_wait(q_full, 0);
    const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), ds_addr = smem_u32(sDS);
    const int n_units = g.n_units;
    const uint32_t idesc_dq = g.idesc_dq;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      const bool last = (u == n_units - 1);
      const int ncols = last ? g.last_cols16 : 128;
      const uint32_t idesc_s = last ? g.idesc_s_last : g.idesc_s_full;
      mbar_wait(&kv_full[st], (u >> 1) & 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + st * kT), v_addr = smem_u32(sV + st * kT);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_s, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                   make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_dp, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                   make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(ds_full, u & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        for (int ks = 0; ks < ncols / 16; ++ks)
          umma_f16(tm_dq, make_smem_desc_sw128(ds_addr + (ks >> 2) * kT + (ks & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(k_addr + ks * 2048, 8192, 1024), idesc_dq, (u | ks) != 0 ? 1u : 0u);
        umma_commit(dq_done);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;  // which 64 of the unit's 128 key columns (and which 32 dQ columns) this thread owns
    const int r = q * 32 + lane;
    const int qi = qt * 128 + r;
    const bool row_valid = qi < g.S;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float L = row_valid ? g.lse[(static_cast<long long>(z) * g.heads + h) * g.S + qi] : 0.f;
    const float Dv = row_valid ? g.dvec[(static_cast<long long>(row0) + qi) * g.heads + h] : 0.f;
    uint8_t* ds_row = sDS + r * 128;
    for (int u = 0; u < g.n_units; ++u) {
      const bool last = (u == g.n_units - 1);
      const int ncols = last ? g.last_cols16 : 128;
      int nvalid = g.S - u * 128;
      if (nvalid > 128) nvalid = 128;
      mbar_wait(sdp_full, u & 1);
      tc_fence_after();
      if (u > 0) mbar_wait(dq_done, (u - 1) & 1);  // the previous dQ MMA has finished reading sDS
      for (int ch = half * 2; ch < half * 2 + 2 && ch * 32 < ncols; ++ch) {
        uint32_t rs[32], rp[32];
        tmem_ld32(tm_s + lane_off + ch * 32, rs);
        tmem_ld32(tm_dp + lane_off + ch * 32, rp);
        tmem_ld_wait();
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = ch * 32 + s8 * 8 + i;
            const bool ok = row_valid && (c < nvalid);
            const float p = ok ? fast_exp2(fmaf(__uint_as_float(rs[s8 * 8 + i]), g.scale_log2, -L)) : 0.f;
            v[i] = p * (__uint_as_float(rp[s8 * 8 + i]) - Dv);
          }
          store_chunk16(ds_row, r, ch * 32 + s8 * 8, v);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, (g.n_units - 1) & 1);
    tc_fence_after();
    __half* dst = g.d_qkv + (static_cast<long long>(row0) + qi) * g.ld_dqkv + g.q_col0 + h * 64;
    {
      const int hlf = half;
      uint32_t rr[32];
      tmem_ld32(tm_dq + lane_off + hlf * 32, rr);
      tmem_ld_wait();
      if (row_valid) {
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hh[i] = __floats2half2_rn(__uint_as_float(rr[s8 * 8 + 2 * i]) * g.scale,
                                      __uint_as_float(rr[s8 * 8 + 2 * i + 1]) * g.scale);
          *reinterpret_cast<uint4*>(dst + hlf * 32 + s8 * 8) = v;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_dkdv_kernel(const __grid_constant__ AttnBwdArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sK = smem;
  uint8_t* sV = smem + kT;
  uint8_t* sQ = smem + 2 * kT;   // 2 stages
  uint8_t* sDO = smem + 4 * kT;  // 2 stages
  uint8_t* sP = smem + 6 * kT;   // [query][key], 2 key blocks
  uint8_t* sDS = smem + 8 * kT;  // [query][key], 2 key blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 10 * kT);
  uint64_t* k_full = bars + 0;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const int h = blockIdx.x, kt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    mbar_init(k_full, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1);
    mbar_init(&q_empty[1], 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_qkv);
    tma_prefetch_desc(&g.map_do);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320;
  const int row0 = z * g.S;
  int kvalid = g.S - kt * 128;  // valid keys in this tile
  if (kvalid > 128) kvalid = 128;
  const int kcols16 = (kvalid + 15) / 16 * 16;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(k_full, 2 * kT);
      tma_load_2d(sK, &g.map_qkv, k_full, g.k_col0 + h * 64, row0 + kt * 128);
      tma_load_2d(sV, &g.map_qkv, k_full, g.v_col0 + h * 64, row0 + kt * 128);
    }
    __syncwarp();
    const int n_units = g.n_units;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      mbar_wait(&q_empty[st], ((u >> 1) & 1) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&q_full[st], 2 * kT);
        tma_load_2d(sQ + st * kT, &g.map_qkv, &q_full[st], g.q_col0 + h * 64, row0 + u * 128);
        tma_load_2d(sDO + st * kT, &g.map_do, &q_full[st], h * 64, row0 + u * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    mbar_wait(k_full, 0);
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
    const uint32_t idesc_s = make_idesc_f16(128, static_cast<uint32_t>(kcols16), 0, 0);
    const uint32_t idesc_dkv = g.idesc_dkv;
    const int n_units = g.n_units;
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;
      mbar_wait(&q_full[st], (u >> 1) & 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ + st * kT), do_addr = smem_u32(sDO + st * kT);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_s, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                   make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm_dp, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                   make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, u & 1);
      tc_fence_after();
        // dV[key, d] += sum_q P[q, key] dO[q, d] ;  dK[key, d] += sum_q dS[q, key] Q[q, d]
        // A = P^T / dS^T read MN-major straight from the [query][key] tiles (M = keys: 2 blocks of 64, LBO = kT);
        // B = dO / Q tiles MN-major ([query rows][d]); one k-step = 16 query rows = 2048 bytes in both.
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_f16(tm_dv, make_smem_desc_sw128(p_addr + ks * 2048, kT, 1024),
                   make_smem_desc_sw128(do_addr + ks * 2048, 8192, 1024), idesc_dkv, (u | ks) != 0 ? 1u : 0u);
          umma_f16(tm_dk, make_smem_desc_sw128(ds_addr + ks * 2048, kT, 1024),
                   make_smem_desc_sw128(q_addr + ks * 2048, 8192, 1024), idesc_dkv, (u | ks) != 0 ? 1u : 0u);
        }
        umma_commit(acc_done);
        umma_commit(&q_empty[st]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // key columns [64 half, 64 half + 64) of every unit; dV / dK columns [32 half, +32)
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* p_row = sP + r * 128;
    uint8_t* ds_row = sDS + r * 128;
    for (int u = 0; u < g.n_units; ++u) {
      const int qi = u * 128 + r;
      const bool row_valid = qi < g.S;
      const float L = row_valid ? g.lse[(static_cast<long long>(z) * g.heads + h) * g.S + qi] : 0.f;
      const float Dv = row_valid ? g.dvec[(static_cast<long long>(row0) + qi) * g.heads + h] : 0.f;
      mbar_wait(sdp_full, u & 1);
      tc_fence_after();
      if (u > 0) mbar_wait(acc_done, (u - 1) & 1);  // previous dV/dK MMAs have finished reading sP / sDS
      for (int ch = half * 2; ch < half * 2 + 2; ++ch) {
        float pv[32], dv_[32];
        if (ch * 32 < kcols16) {
          uint32_t rs[32], rp[32];
          tmem_ld32(tm_s + lane_off + ch * 32, rs);
          tmem_ld32(tm_dp + lane_off + ch * 32, rp);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = row_valid && (ch * 32 + i < kvalid);
            const float p = ok ? fast_exp2(fmaf(__uint_as_float(rs[i]), g.scale_log2, -L)) : 0.f;
            pv[i] = p;
            dv_[i] = p * (__uint_as_float(rp[i]) - Dv);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = dv_[i] = 0.f;  // keys beyond the tile: the M=128 MMA still reads them
        }
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          float a[8], b[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a[i] = pv[s8 * 8 + i];
            b[i] = dv_[s8 * 8 + i];
          }
          store_chunk16(p_row, r, ch * 32 + s8 * 8, a);
          store_chunk16(ds_row, r, ch * 32 + s8 * 8, b);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
    }
    mbar_wait(acc_done, (g.n_units - 1) & 1);
    tc_fence_after();
    const int ki = kt * 128 + r;  // this thread's key row
    const bool key_valid = r < kvalid;
    __half* dst = g.d_qkv + (static_cast<long long>(row0) + ki) * g.ld_dqkv + h * 64;
#pragma unroll
    for (int part = half; part < 4; part += 2) {  // 0,1: dV halves; 2,3: dK halves -- this thread's column half of both
      uint32_t rr[32];
      tmem_ld32((part < 2 ? tm_dv : tm_dk) + lane_off + (part & 1) * 32, rr);
      tmem_ld_wait();
      if (key_valid) {
        const float sc = part < 2 ? 1.f : g.scale;
        __half* d = dst + (part < 2 ? g.v_col0 : g.k_col0) + (part & 1) * 32;
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hh[i] = __floats2half2_rn(__uint_as_float(rr[s8 * 8 + 2 * i]) * sc, __uint_as_float(rr[s8 * 8 + 2 * i + 1]) * sc);
          *reinterpret_cast<uint4*>(d + s8 * 8) = v;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------ single-unit adjoint
// All keys of a query tile fit one 128-column unit, so the probabilities are recomputed exactly in-kernel (row max,
// row sum) and D = sum_j P dP needs no forward output.  Two users:
//   mode 1, temporal packed attention: the tile's 128 rows are bp sequences of F consecutive rows; query tile ==
//           key tile; block-diagonal window mask; emits dQ, dK, dV.
//   mode 0, cross-attention: keys/values are the text tokens of batch z / kv_z_div; emits dQ only and injects the
//           box-energy gradient dE/dP on the object-token columns.
struct AttnBwdSmallArgs {
  CUtensorMap map_q;   // query-side buffer (q | and for mode 1 also k | v), box 64 x 128
  CUtensorMap map_kv;  // key/value buffer
  CUtensorMap map_do;
  const __half* dp_tok;
  __half* d_q;
  long long ld_dq;
  long long total_q_rows;  // mode 1: all packed rows; mode 0: s_q
  int mode, q_col0, k_col0, v_col0, dq_col0, dk_col0, dv_col0;
  int tile_rows;  // row stride between tiles (mode 1: bp*F, mode 0: 128)
  int seq_len, s_q, s_k, heads, nz, kv_z_div, n_slots;
  int slot_col[16];
  int ncols16;
  float scale, scale_log2;
  uint32_t idesc_s, idesc_dq, idesc_dkv;
};

constexpr int kBwdSmallThreads = 192;  // 2 control warps + 4 compute warps (one score row per thread)

__global__ void __launch_bounds__(kBwdSmallThreads, 1) attn_bwd_small_kernel(const __grid_constant__ AttnBwdSmallArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // (warp index through a shuffle: ptxas then knows it is warp-uniform and keeps the role loops in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + kT;
  uint8_t* sK = smem + 2 * kT;
  uint8_t* sV = smem + 3 * kT;
  uint8_t* sP = smem + 4 * kT;   // [query][key], 2 key blocks
  uint8_t* sDS = smem + 6 * kT;  // [query][key], 2 key blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * kT);
  uint64_t* in_full = bars + 0;
  uint64_t* sdp_full = bars + 1;
  uint64_t* pds_full = bars + 2;
  uint64_t* acc_done = bars + 3;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 4);
  const int h = blockIdx.x, qt = blockIdx.y, z = blockIdx.z;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  {
    uint4 zero = make_uint4(0, 0, 0, 0);  // P / dS must be zero outside the columns the threads write
    for (int i = threadIdx.x; i < 4 * kT / 16; i += kBwdSmallThreads) reinterpret_cast<uint4*>(sP)[i] = zero;
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    mbar_init(in_full, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 128);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dq = tmem_base + 256, tm_dk = tmem_base + 320,
                 tm_dv = tmem_base + 384;
  const bool temporal = g.mode == 1;
  const long long q_row0 = temporal ? static_cast<long long>(qt) * g.tile_rows
                                    : static_cast<long long>(z) * g.s_q + static_cast<long long>(qt) * 128;
  const long long kv_row0 = temporal ? q_row0 : static_cast<long long>(z / g.kv_z_div) * g.s_k;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(in_full, 4 * kT);
      tma_load_2d(sQ, &g.map_q, in_full, g.q_col0 + h * 64, static_cast<int>(q_row0));
      tma_load_2d(sDO, &g.map_do, in_full, h * 64, static_cast<int>(q_row0));
      tma_load_2d(sK, &g.map_kv, in_full, g.k_col0 + h * 64, static_cast<int>(kv_row0));
      tma_load_2d(sV, &g.map_kv, in_full, g.v_col0 + h * 64, static_cast<int>(kv_row0));
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(in_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tm_s, make_smem_desc_sw128(q_addr + k * 32, 16, 1024), make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                 g.idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tm_dp, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                 make_smem_desc_sw128(v_addr + k * 32, 16, 1024), g.idesc_s, k != 0 ? 1u : 0u);
      umma_commit(sdp_full);
    }
    __syncwarp();
    mbar_wait(pds_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      for (int ks = 0; ks < g.ncols16 / 16; ++ks)  // dQ = dS K        (K tile as MN-major B)
        umma_f16(tm_dq, make_smem_desc_sw128(ds_addr + (ks >> 2) * kT + (ks & 3) * 32, 16, 1024),
                 make_smem_desc_sw128(k_addr + ks * 2048, 8192, 1024), g.idesc_dq, ks != 0 ? 1u : 0u);
      if (temporal) {
        for (int ks = 0; ks < 8; ++ks) {  // dV = P^T dO, dK = dS^T Q   (P / dS tiles as MN-major A)
          umma_f16(tm_dv, make_smem_desc_sw128(p_addr + ks * 2048, kT, 1024),
                   make_smem_desc_sw128(do_addr + ks * 2048, 8192, 1024), g.idesc_dkv, ks != 0 ? 1u : 0u);
          umma_f16(tm_dk, make_smem_desc_sw128(ds_addr + ks * 2048, kT, 1024),
                   make_smem_desc_sw128(q_addr + ks * 2048, 8192, 1024), g.idesc_dkv, ks != 0 ? 1u : 0u);
        }
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const long long qrow = q_row0 + r;  // global query row (mode 1: also this thread's key row)
    const int qi = qt * 128 + r;        // mode 0: query index inside image z
    const bool row_valid = temporal ? (r < g.tile_rows && qrow < g.total_q_rows) : (qi < g.s_q);
    // columns this warp has to look at, in chunks of 8
    int cbeg = 0, cend = g.ncols16, mystart = 0, myend = g.s_k;
    if (temporal) {
      const int F = g.seq_len;
      cbeg = (q * 32) / F * F;
      cend = ((q * 32 + 31) / F + 1) * F;
      if (cend > 128) cend = 128;
      mystart = (r / F) * F;
      myend = mystart + F;
      if (myend > g.tile_rows) myend = g.tile_rows;
    }
    const int nch = (cend - cbeg + 7) >> 3;
    mbar_wait(sdp_full, 0);
    tc_fence_after();
    // energy gradient on this row's object-token columns (mode 0)
    float tokg[16];
#pragma unroll
    for (int t = 0; t < 16; ++t)
      tokg[t] = (t < g.n_slots && row_valid)
                    ? __half2float(g.dp_tok[((static_cast<long long>(t) * g.nz + z) * g.heads + h) * g.s_q + qi])
                    : 0.f;
    float mx = -INFINITY;
    for (int k = 0; k < nch; ++k) {
      uint32_t rs[8];
      tmem_ld8(tm_s + lane_off + cbeg + k * 8, rs);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cbeg + k * 8 + i;
        if (c >= mystart && c < myend) mx = fmaxf(mx, __uint_as_float(rs[i]));
      }
    }
    mx *= g.scale_log2;
    if (!(mx > -INFINITY)) mx = 0.f;
    float l = 0.f, td = 0.f;
    for (int k = 0; k < nch; ++k) {
      uint32_t rs[8], rp[8];
      tmem_ld8(tm_s + lane_off + cbeg + k * 8, rs);
      tmem_ld8(tm_dp + lane_off + cbeg + k * 8, rp);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cbeg + k * 8 + i;
        if (c >= mystart && c < myend) {
          const float e = fast_exp2(fmaf(__uint_as_float(rs[i]), g.scale_log2, -mx));
          float dp = __uint_as_float(rp[i]);
          for (int t = 0; t < g.n_slots; ++t)
            if (g.slot_col[t] == c) dp += tokg[t];
          l += e;
          td += e * dp;
        }
      }
    }
    const float inv = row_valid ? 1.f / l : 0.f;
    const float Dv = td * inv;
    uint8_t* p_row = sP + r * 128;
    uint8_t* ds_row = sDS + r * 128;
    for (int k = 0; k < nch; ++k) {
      uint32_t rs[8], rp[8];
      tmem_ld8(tm_s + lane_off + cbeg + k * 8, rs);
      tmem_ld8(tm_dp + lane_off + cbeg + k * 8, rp);
      tmem_ld_wait();
      float pv[8], dsv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cbeg + k * 8 + i;
        float p = 0.f, ds = 0.f;
        if (c >= mystart && c < myend) {
          p = round_h(fast_exp2(fmaf(__uint_as_float(rs[i]), g.scale_log2, -mx)) * inv);
          float dp = __uint_as_float(rp[i]);
          for (int t = 0; t < g.n_slots; ++t)
            if (g.slot_col[t] == c) dp += tokg[t];
          ds = p * (dp - Dv);
        }
        pv[i] = p;
        dsv[i] = ds;
      }
      if (temporal) store_chunk16(p_row, r, cbeg + k * 8, pv);
      store_chunk16(ds_row, r, cbeg + k * 8, dsv);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(pds_full);
    mbar_wait(acc_done, 0);
    tc_fence_after();
    __half* dst = g.d_q + qrow * g.ld_dq + h * 64;
    const int nparts = temporal ? 6 : 2;
#pragma unroll 1
    for (int part = 0; part < nparts; ++part) {  // 0,1: dQ halves; 2,3: dK halves; 4,5: dV halves
      uint32_t rr[32];
      const uint32_t base = part < 2 ? tm_dq : (part < 4 ? tm_dk : tm_dv);
      tmem_ld32(base + lane_off + (part & 1) * 32, rr);
      tmem_ld_wait();
      if (row_valid) {
        const float sc = part < 4 ? g.scale : 1.f;
        __half* d = dst + (part < 2 ? g.dq_col0 : (part < 4 ? g.dk_col0 : g.dv_col0)) + (part & 1) * 32;
#pragma unroll
        for (int s8 = 0; s8 < 4; ++s8) {
          uint4 v;
          __half2* hh = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hh[i] = __floats2half2_rn(__uint_as_float(rr[s8 * 8 + 2 * i]) * sc, __uint_as_float(rr[s8 * 8 + 2 * i + 1]) * sc);
          *reinterpret_cast<uint4*>(d + s8 * 8) = v;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lvd

using namespace lvd;

extern "C" int lvd_self_attn_bwd(const void* qkv, int64_t ld_qkv, int64_t width_qkv, int32_t q_col0, int32_t k_col0,
                                 int32_t v_col0, const void* d_o, int64_t ld_do, const void* lse, const void* dvec,
                                 void* d_qkv, int64_t ld_dqkv, int64_t s, int32_t heads, int32_t nz, float scale,
                                 void* stream) {
  LVD_REQUIRE(qkv && d_o && lse && dvec && d_qkv, "lvd_self_attn_bwd: null pointer");
  LVD_REQUIRE(s >= 1 && heads >= 1 && nz >= 1 && nz <= 65535, "lvd_self_attn_bwd: bad sizes");
  LVD_REQUIRE(ld_qkv % 8 == 0 && ld_do % 8 == 0 && ld_dqkv % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 &&
                  v_col0 % 8 == 0,
              "lvd_self_attn_bwd: alignment");
  AttnBwdArgs g;
  memset(&g, 0, sizeof(g));
  const int64_t rows = s * nz;
  // plain 2-D maps; tiles that cross an image boundary read the neighbour's rows, which the per-row / per-key
  // validity masks discard (they are finite values), rows past the end of the buffer read as zero.
  int64_t d1[2] = {width_qkv, rows};
  int64_t s1[2] = {1, ld_qkv};
  int b1[2] = {64, 128};
  int rc = encode_tmap_f16_sw128(&g.map_qkv, qkv, 2, d1, s1, b1);
  if (rc) return rc;
  int64_t d2[2] = {static_cast<int64_t>(heads) * 64, rows};
  int64_t s2[2] = {1, ld_do};
  rc = encode_tmap_f16_sw128(&g.map_do, d_o, 2, d2, s2, b1);
  if (rc) return rc;
  g.lse = static_cast<const float*>(lse);
  g.dvec = static_cast<const float*>(dvec);
  g.d_qkv = static_cast<__half*>(d_qkv);
  g.ld_dqkv = ld_dqkv;
  g.q_col0 = q_col0;
  g.k_col0 = k_col0;
  g.v_col0 = v_col0;
  g.S = static_cast<int>(s);
  g.heads = heads;
  g.n_units = static_cast<int>((s + 127) / 128);
  g.scale = scale;
  g.scale_log2 = scale * 1.4426950408889634f;
  const int last_valid = g.S - (g.n_units - 1) * 128;
  g.last_cols16 = (last_valid + 15) / 16 * 16;
  g.idesc_s_full = make_idesc_f16(128, 128, 0, 0);
  g.idesc_s_last = make_idesc_f16(128, static_cast<uint32_t>(g.last_cols16), 0, 0);
  g.idesc_dq = make_idesc_f16(128, 64, 0, 1);
  g.idesc_dkv = make_idesc_f16(128, 64, 1, 1);
  const int smem_dq = 8 * kT + 128, smem_dkv = 10 * kT + 128;
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
    LVD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv));
    attr = true;
  }
  dim3 grid(heads, g.n_units, nz);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  attn_bwd_dq_kernel<<<grid, kBwdThreads, smem_dq, st>>>(g);
  attn_bwd_dkdv_kernel<<<grid, kBwdThreads, smem_dkv, st>>>(g);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

// tensor-core single-unit adjoints; declared in attention_bwd_small.cu's entry points
namespace lvd {
int launch_attn_bwd_small(int mode, const void* q, long long ld_q, long long q_width, long long q_rows, const void* kv,
                          long long ld_kv, long long kv_width, long long kv_rows, const void* d_o, long long ld_do,
                          long long do_width, const void* dp_tok, const int* slot_col, int n_slots, void* d_q,
                          long long ld_dq, int q_col0, int k_col0, int v_col0, int dq_col0, int dk_col0, int dv_col0,
                          int frames, long long s_q, int s_k, int heads, int nz, int kv_z_div, float scale,
                          cudaStream_t st) {
  AttnBwdSmallArgs g;
  memset(&g, 0, sizeof(g));
  int b[2] = {64, 128};
  int64_t d1[2] = {q_width, q_rows}, s1[2] = {1, ld_q};
  int rc = encode_tmap_f16_sw128(&g.map_q, q, 2, d1, s1, b);
  if (rc) return rc;
  int64_t d2[2] = {kv_width, kv_rows}, s2[2] = {1, ld_kv};
  rc = encode_tmap_f16_sw128(&g.map_kv, kv, 2, d2, s2, b);
  if (rc) return rc;
  int64_t d3[2] = {do_width, q_rows}, s3[2] = {1, ld_do};
  rc = encode_tmap_f16_sw128(&g.map_do, d_o, 2, d3, s3, b);
  if (rc) return rc;
  g.dp_tok = static_cast<const __half*>(dp_tok);
  g.d_q = static_cast<__half*>(d_q);
  g.ld_dq = ld_dq;
  g.mode = mode;
  g.q_col0 = q_col0;
  g.k_col0 = k_col0;
  g.v_col0 = v_col0;
  g.dq_col0 = dq_col0;
  g.dk_col0 = dk_col0;
  g.dv_col0 = dv_col0;
  g.heads = heads;
  g.nz = nz;
  g.kv_z_div = kv_z_div;
  g.n_slots = n_slots;
  for (int i = 0; i < n_slots; ++i) g.slot_col[i] = slot_col[i];
  g.scale = scale;
  g.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid;
  if (mode == 1) {
    const int bp = 128 / frames;
    g.tile_rows = bp * frames;
    g.seq_len = frames;
    g.total_q_rows = q_rows;
    g.s_q = 0;
    g.s_k = g.tile_rows;
    g.ncols16 = (g.tile_rows + 15) / 16 * 16;
    grid = dim3(heads, static_cast<unsigned>((q_rows + g.tile_rows - 1) / g.tile_rows), 1);
  } else {
    g.tile_rows = 128;
    g.s_q = static_cast<int>(s_q);
    g.s_k = s_k;
    g.total_q_rows = s_q;
    g.ncols16 = (s_k + 15) / 16 * 16;
    grid = dim3(heads, static_cast<unsigned>((s_q + 127) / 128), nz);
  }
  g.idesc_s = make_idesc_f16(128, static_cast<uint32_t>(g.ncols16), 0, 0);
  g.idesc_dq = make_idesc_f16(128, 64, 0, 1);
  g.idesc_dkv = make_idesc_f16(128, 64, 1, 1);
  const int smem_b = 8 * kT + 128;
  static bool attr = false;
  if (!attr) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    attr = true;
  }
  attn_bwd_small_kernel<<<grid, kBwdSmallThreads, smem_b, st>>>(g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}
}  // namespace lvd
