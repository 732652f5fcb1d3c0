// gemm.cu -- persistent, warp-specialised tcgen05 contraction with fused epilogue (see include/lvd_b200.h).
//
//   warp 0      : TMA producer   (one elected lane; cp.async.bulk.tensor into a multi-stage smem ring)
//   warp 1      : MMA issuer     (one elected lane; tcgen05.mma, fp32 accumulators double-buffered in TMEM)
//   warps 2..17 : epilogue       (four groups of four warps -- one warp per TMEM lane quarter; a group turns one
//                 128-row x 32-column slab of the accumulator into fp16 in shared memory and TMA-stores it)
//
// A tiles are rank-5 TMA boxes of the channels-last activation; a convolution tap is a coordinate offset and the
// zero padding is TMA out-of-bounds fill, so Linear, Conv2d (3x3, 1x1, stride 2) and Conv3d (3,1,1) all run here.
#include <atomic>
#include <cstdlib>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {

std::atomic<long long> g_launches{0};

constexpr int kAStageBytes = 128 * 128;  // 128 rows x 64 halves
constexpr int kThreads = 576;            // 2 control warps + 16 epilogue warps
constexpr int kSmemLimit = 232448;       // 227 KB
constexpr int kSlabBytes = 128 * 64;     // epilogue staging slab: 128 rows x 32 halves
constexpr int kSlabBufs = 8;             // 4 epilogue groups x 2 buffers
constexpr int kBarrierBytes = 512;       // mbarriers (8+8 ring, 4+4 accumulator, 8 residual, 1 resident B) + the TMEM address

// epilogue specialisations (template parameter): the common cases carry no feature branches in the slab loop
enum { kEpiGeneric = 0, kEpiBias = 1, kEpiResidual = 2, kEpiGeglu = 3, kEpiLnBias = 4, kEpiLnGeglu = 5 };

struct GemmArgs {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_out;  // 32-column slabs of the output (64-byte swizzle), rows mapped like the A box
  CUtensorMap map_res;  // same geometry over the residual
  long long ld_rowvec;
  const __half* bias;
  const float* bias32;  // optional fp32 copy of bias (same values): the lean epilogues skip the conversions
  const __half* rowvec;
  const __half* residual;
  int box[4];
  int ntile[4];
  int out_dims[4];
  int rv_stride[4];
  int tap_off[LVD_MAX_TAPS][5];
  int n_taps, kb_per_tap;
  int n_tiles_n, block_n, b_mn_major, act;
  int num_tiles, stages, a_bytes, b_bytes;
  int b_k_tap_stride, b_n_tap_stride;
  int slab_bytes;  // rows of the box * 64
  int n_acc;       // accumulator stages in TMEM (512 columns / block_n, 2..4), block_n columns apart
  // persistent tile walk: tile += tile_step; its n-tile / row-tile digits advance by precomputed mixed-radix steps
  int tile_step, step_nt;
  int step_lo[4], step_hi[4];
  // weight-resident mode: the persistent walk advances by a multiple of n_tiles_n, so a CTA (pair) keeps ONE n-tile; its B
  // tile (all k-blocks) is loaded once into b_slots = kblocks slots and only A streams through the ring
  int b_resident, b_slots;
  int debug;  // tools/perf_gemm_sweep.py: bit 0 = skip the TMA loads, bit 1 = skip the MMAs, bit 2 = skip the output
              // stores, bit 3 = skip the residual loads (results are garbage)
  float gamma;
  uint32_t idesc;
  // GroupNorm statistics of the OUTPUT for the GroupNorm that consumes it (NULL = off): every epilogue warp adds the
  // column sums {sum, sum of squares} of its 32 staged rows, folded per channel group, into
  // gn_part[sample][row tile of the sample][row quarter][slab residue][group][2]
  float* gn_part;
  int gn_gw, gn_groups, gn_tiles_per_sample, gn_sub;
  // LayerNorm folded into this linear (kEpiLn*): A is the raw LayerNorm input, B carries the LayerNorm weight
  // (B' = gamma o W), and the epilogue finishes  out[r][n] = rstd_r * (acc[r][n] - mean_r * ln_cs[n]) + ln_bias[n]
  // with ln_cs[n] = sum_k B'[n][k], ln_bias[n] = sum_k beta_k W[n][k] + bias[n]; ln_stats[row] = {mean, rstd}, row =
  // sum coord_i * ln_stride[i] (the A row).
  const float2* ln_stats;
  const float* ln_cs;
  const float* ln_bias;
  int ln_stride[4];
  // Row statistics of the OUTPUT for the LayerNorm that consumes it (NULL = off): every epilogue thread stores
  // {sum, sum of squares} of the 32 fp16 values of its row in the slab to lnstat_part[slab][row] (one writer each).
  float2* lnstat_part;
  long long lnstat_rows;
  int lnstat_stride[4];
};

// Tuning aids (tools/diag_gemm_k320.py, perf_gemm_sweep.py, trace_gemm.py) exist only in builds with
// LVD_EXTRA_NVCC_FLAGS=-DLVD_GEMM_DEBUG_BITS: the LVD_GEMM_DEBUG environment bits switch parts of the kernel off, bit 6
// makes CTA 0 append (event << 48 | globaltimer ns) to g_trace.  Release builds carry none of these branches.
__device__ unsigned long long g_trace[8192];
__device__ unsigned int g_trace_n;
#ifdef LVD_GEMM_DEBUG_BITS
#define LVD_DBG(g, bit) (((g).debug & (bit)) != 0)
__device__ __forceinline__ void trace_event(const GemmArgs& g, unsigned ev) {
  if ((g.debug & 64) && blockIdx.x == 0) {
    const unsigned i = atomicAdd(&g_trace_n, 1u);
    if (i < 8192) g_trace[i] = (static_cast<unsigned long long>(ev) << 48) | (globaltimer_ns() & 0xFFFFFFFFFFFFull);
  }
}
#else
#define LVD_DBG(g, bit) false
__device__ __forceinline__ void trace_event(const GemmArgs&, unsigned) {}
#endif

// Per-warp phase profile of the epilogue (build with LVD_EXTRA_NVCC_FLAGS=-DLVD_GEMM_PHASE_PROF; tools/phase_gemm.py):
// clock64 deltas accumulated in registers, dumped into g_trace[warp * 16 + phase] by CTA 0 at kernel end.
#ifdef LVD_GEMM_PHASE_PROF
#define LVD_PH(i)                 \
  do {                            \
    const long long t_ = clock64(); \
    ph[i] += t_ - ph_last;        \
    ph_last = t_;                 \
  } while (0)
#else
#define LVD_PH(i)
#endif

// (n-tile, row-tile digits) of the current tile.  Integer division by run-time extents costs ~100 dependent
// instructions each; doing six of them per tile in the single producer thread was ~0.8 us per tile (measured).
struct TileIter {
  int tile, nt, t0, t1, t2, t3;
};
template <int kCtas>
__device__ __forceinline__ void tile_init(const GemmArgs& g, TileIter& it, int tile, int cta_rank) {
  it.tile = tile;
  it.nt = tile % g.n_tiles_n;
  int r = (tile / g.n_tiles_n) * kCtas + cta_rank;
  it.t0 = r % g.ntile[0];
  r /= g.ntile[0];
  it.t1 = r % g.ntile[1];
  r /= g.ntile[1];
  it.t2 = r % g.ntile[2];
  it.t3 = r / g.ntile[2];
}
__device__ __forceinline__ void tile_advance(const GemmArgs& g, TileIter& it) {
  it.tile += g.tile_step;
  it.nt += g.step_nt;
  const bool carry = it.nt >= g.n_tiles_n;
  if (carry) it.nt -= g.n_tiles_n;
  it.t0 += carry ? g.step_hi[0] : g.step_lo[0];
  int c = it.t0 >= g.ntile[0] ? 1 : 0;
  it.t0 -= c ? g.ntile[0] : 0;
  it.t1 += (carry ? g.step_hi[1] : g.step_lo[1]) + c;
  c = it.t1 >= g.ntile[1] ? 1 : 0;
  it.t1 -= c ? g.ntile[1] : 0;
  it.t2 += (carry ? g.step_hi[2] : g.step_lo[2]) + c;
  c = it.t2 >= g.ntile[2] ? 1 : 0;
  it.t2 -= c ? g.ntile[2] : 0;
  it.t3 += (carry ? g.step_hi[3] : g.step_lo[3]) + c;  // past ntile[3]: out of bounds for every TMA box
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint32_t pack2h(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// erf-GELU as in gelu_erf_f (Abramowitz-Stegun 7.1.26 tail h(|v|)), finished as max(v, 0) - |v| h: two instructions
// instead of select + subtract + multiply
__device__ __forceinline__ float gelu_erf_lean(float v) {
  const float a = fabsf(v);
  const float t = fast_rcp(fmaf(0.23164189f, a, 1.0f));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float h = (p * t) * fast_exp2((a * a) * -0.72134752f);
  return fmaf(-a, h, fmaxf(v, 0.f));
}

// kCtas == 2: the CTA pair of a 2-CTA cluster works on one 256-row tile (tcgen05 cta_group::2).  Each CTA stages its
// own 128-row A tile and HALF of the B tile, the leader's MMA warp issues M=256 instructions that read both CTAs'
// shared memory, and each CTA's TMEM receives (and its epilogue warps drain) its own 128 accumulator rows.
// kFeat: which output statistics the epilogue emits -- a compile-time choice for the specialised epilogues (bit 0 =
// LayerNorm row statistics, bit 1 = GroupNorm partial sums), -1 = decided at run time (kEpiGeneric).
template <int kCtas, int kEpi, int kFeat>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte aligned stage bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // warp index as a value ptxas knows to be warp-uniform: the role loops then live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int stages = g.stages;
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const int b_rows = g.block_n / kCtas;  // B rows (N) staged by this CTA
  const uint32_t b_stage_bytes = static_cast<uint32_t>(b_rows) * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + stages * kAStageBytes;
  uint8_t* sOut = sB + g.b_slots * b_stage_bytes;  // output slabs [128 rows][32 cols] fp16: 2 per epilogue group
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + kSlabBufs * kSlabBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + 4;
  uint64_t* res_full = tempty + 4;  // [group][buffer]
  uint64_t* bres_full = res_full + kSlabBufs;  // weight-resident mode: the CTA's whole B tile has landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bres_full + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16 * kCtas);  // one arrival per epilogue warp of every CTA of the pair
    }
    for (int i = 0; i < kSlabBufs; ++i) mbar_init(&res_full[i], 1);
    mbar_init(bres_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_a);
    tma_prefetch_desc(&g.map_b);
    tma_prefetch_desc(&g.map_out);
  }
  if (warp == 1) {
    if (kCtas == 2) {
      tmem_alloc_pair(tmem_holder, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_holder, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int tile0 = static_cast<int>(blockIdx.x) / kCtas;  // first (pair-)tile of this CTA
  const int num_tiles = g.num_tiles;

  if (warp == 0) {
    // ---- TMA producer: the whole warp walks the loop, one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    const int n_taps = g.n_taps, kb_per_tap = g.kb_per_tap, block_n = g.block_n;
    const int k_tap_stride = g.b_k_tap_stride, n_tap_stride = g.b_n_tap_stride;
    const bool mn_major = g.b_mn_major != 0;
    const bool resident = g.b_resident != 0;
    const uint32_t tx_bytes =
        static_cast<uint32_t>(kCtas) * static_cast<uint32_t>(g.a_bytes + (resident ? 0 : g.b_bytes));
    const bool skip_loads = LVD_DBG(g, 1);
    TileIter it;
    tile_init<kCtas>(g, it, tile0, static_cast<int>(cta_rank));
    if (resident && it.tile < num_tiles) {
      // this CTA's n-tile never changes: stage its B tile once (K-major B, one tap)
      if (elect_one_sync()) {
        const int ncoord = it.nt * block_n + static_cast<int>(cta_rank) * b_rows;
        const uint32_t total = static_cast<uint32_t>(kCtas) * static_cast<uint32_t>(g.b_bytes) *
                               static_cast<uint32_t>(kb_per_tap);
        if (kCtas == 2) {
          if (cta_rank == 0) mbar_arrive_expect_tx(bres_full, total);
          const uint32_t fbar = mapa_shared(smem_u32(bres_full), 0);
          for (int kb = 0; kb < kb_per_tap; ++kb) tma_load_2d_pair(sB + kb * b_stage_bytes, &g.map_b, fbar, kb * 64, ncoord);
        } else {
          mbar_arrive_expect_tx(bres_full, total);
          for (int kb = 0; kb < kb_per_tap; ++kb) tma_load_2d(sB + kb * b_stage_bytes, &g.map_b, bres_full, kb * 64, ncoord);
        }
      }
      __syncwarp();
    }
    for (; it.tile < num_tiles; tile_advance(g, it)) {
      if (lane == 0) trace_event(g, 1);  // producer: tile start
      const int c0 = it.t0 * g.box[0], c1_ = it.t1 * g.box[1], c2_ = it.t2 * g.box[2], c3_ = it.t3 * g.box[3];
      const int ncoord0 = it.nt * block_n + static_cast<int>(cta_rank) * b_rows;
      for (int tap = 0; tap < n_taps; ++tap) {
        const int o0 = g.tap_off[tap][0];
        const int c1 = c0 + g.tap_off[tap][1], c2 = c1_ + g.tap_off[tap][2];
        const int c3 = c2_ + g.tap_off[tap][3], c4 = c3_ + g.tap_off[tap][4];
        const int ncoord = ncoord0 + tap * n_tap_stride;
        for (int kb = 0; kb < kb_per_tap; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (elect_one_sync()) {
            const int kcoord = tap * k_tap_stride + kb * 64;
            uint8_t* adst = sA + stage * kAStageBytes;
            uint8_t* bdst = sB + stage * b_stage_bytes;
            if (skip_loads) {
              if (cta_rank == 0) mbar_arrive(&full[stage]);
            } else if (kCtas == 2) {
              // both CTAs' bytes are counted on the leader's barrier (the leader's MMA warp is the only waiter)
              if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], tx_bytes);
              const uint32_t fbar = mapa_shared(smem_u32(&full[stage]), 0);
              tma_load_5d_pair(adst, &g.map_a, fbar, o0 + kb * 64, c1, c2, c3, c4);
              if (resident) {
              } else if (!mn_major) {
                tma_load_2d_pair(bdst, &g.map_b, fbar, kcoord, ncoord);
              } else {
                for (int ch = 0; ch < b_rows / 64; ++ch)
                  tma_load_2d_pair(bdst + ch * 8192, &g.map_b, fbar, ncoord + ch * 64, kcoord);
              }
            } else {
              mbar_arrive_expect_tx(&full[stage], tx_bytes);
              tma_load_5d(adst, &g.map_a, &full[stage], o0 + kb * 64, c1, c2, c3, c4);
              if (resident) {
              } else if (!mn_major) {
                tma_load_2d(bdst, &g.map_b, &full[stage], kcoord, ncoord);
              } else {
                for (int ch = 0; ch < b_rows / 64; ++ch)
                  tma_load_2d(bdst + ch * 8192, &g.map_b, &full[stage], ncoord + ch * 64, kcoord);
              }
            }
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA of a pair only).  Shared-memory descriptors: the upper word is constant, the lower
    // word is (address >> 4) | (LBO >> 4) << 16, so stage / k-step changes are plain adds on the lower word.
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const bool mn_major = g.b_mn_major != 0;
      const uint32_t idesc = g.idesc;
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);  // LBO (unused) = 16 B
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((mn_major ? (8192u >> 4) : 1u) << 16);
      const uint32_t b_kstep = mn_major ? (2048u >> 4) : (32u >> 4);  // 16 k: 16 rows of 128 B / 32 B along the row
      const uint32_t b_stage16 = b_stage_bytes >> 4;
      const bool skip_mma = LVD_DBG(g, 2);
      const int kblocks = g.n_taps * g.kb_per_tap;
      const bool resident = g.b_resident != 0;
      if (resident && tile0 < num_tiles) {
        mbar_wait(bres_full, 0);
        tc_fence_after();
      }
      for (int tile = tile0; tile < num_tiles; tile += g.tile_step) {
        if (lane == 0) trace_event(g, 2);  // MMA: waiting for the accumulator
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0) trace_event(g, 3);  // MMA: accumulator free, first k-block wait
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * g.block_n);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * (kAStageBytes >> 4);
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(resident ? kb : stage) * b_stage16;
            if (!skip_mma) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k);
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + b_kstep * k);
                if (kCtas == 2)
                  umma_f16_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                else
                  umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            if (kCtas == 2) umma_commit_pair(&empty[stage], 3); else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one_sync()) {
          if (kCtas == 2) umma_commit_pair(&tfull[acc], 3); else umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (lane == 0) trace_event(g, 4);  // MMA: all k-blocks issued
        if (++acc == g.n_acc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ---- epilogue: 16 warps = 4 groups x 4 warps (one per TMEM lane quarter).  The tile's output is emitted as
    // 32-column slabs, dealt round-robin to the groups ACROSS tiles (slab number tile_seq*nslabs + sl goes to group
    // number mod 4), so 5-slab tiles (block_n = 160) still load the groups evenly.  Per slab a thread reads its row's
    // 32 accumulator columns from TMEM, applies the epilogue and writes four 16-byte chunks into the group's staging
    // slab (64-byte swizzle); one elected lane then issues ONE TMA store for the slab -- global traffic is full-line
    // bulk copies with the row mapping and edge clipping done by the tensor map.  The residual is TMA-prefetched INTO
    // the staging slab one slab ahead and added in place.  Two staging buffers per group: the store of slab s-1 drains
    // while slab s is computed.
    const int q = warp & 3;        // this warp may only touch TMEM lanes [32q, 32q+32)
    const int grp = (warp - 2) >> 2;
    const bool leader_warp = ((warp - 2) & 3) == 0;  // one elected lane of this warp issues the group's TMA traffic
    const int r = q * 32 + lane;
    const bool geglu = kEpi == kEpiGeglu || kEpi == kEpiLnGeglu || (kEpi == kEpiGeneric && g.act == 2);
    const bool has_ln = kEpi == kEpiLnBias || kEpi == kEpiLnGeglu;
    const bool has_lnstat = kFeat < 0 ? g.lnstat_part != nullptr : (kFeat & 1) != 0;
    const bool has_res = kEpi == kEpiResidual || (kEpi == kEpiGeneric && g.residual != nullptr);
    const bool has_rowvec = kEpi == kEpiGeneric && g.rowvec != nullptr;
    const bool has_gamma = kEpi == kEpiGeneric && g.gamma != 1.0f;
    const bool has_silu = kEpi == kEpiGeneric && g.act == 1;
    const bool has_gelu = kEpi == kEpiGeneric && g.act == 3;
    const __half* const bias = g.bias;
    const bool has_gn = kFeat < 0 ? g.gn_part != nullptr : (kFeat & 2) != 0;
    constexpr bool kLean = kEpi != kEpiGeneric;  // no run-time feature switches inside the slab loop
    int j[4];
    if (has_rowvec || has_gn || has_ln || has_lnstat) {
      int t = r;
      j[0] = t % g.box[0];
      t /= g.box[0];
      j[1] = t % g.box[1];
      t /= g.box[1];
      j[2] = t % g.box[2];
      t /= g.box[2];
      j[3] = t;  // >= box[3] when the box has fewer than 128 rows
    }
    uint8_t* my_out = sOut + grp * 2 * kSlabBytes;  // two buffers, alternated per slab
    uint64_t* my_res_full = res_full + grp * 2;
    // 64-byte swizzle: 16-byte chunk c of row r lives at r*64 + ((c ^ ((r >> 1) & 3)) << 4)
    const uint32_t row_off = static_cast<uint32_t>(r * 64);
    const uint32_t sw = static_cast<uint32_t>((r >> 1) & 3);
    const int block_n = g.block_n;
    const int n_half = block_n >> 1;                  // GEGLU: tile = [n_half value columns | n_half gate columns]
    const int out_cols = geglu ? n_half : block_n;    // output columns one tile produces
    const int nslabs = out_cols / 32;
    const int bar_id = 1 + grp;
    // the accumulator-drained signal goes to the leader CTA's barrier (the only MMA issuer of the pair)
    const uint32_t tempty_addr0 = kCtas == 2 ? mapa_shared(smem_u32(&tempty[0]), 0) : smem_u32(&tempty[0]);
    uint32_t seq = 0;   // slabs processed by this group: buffer = seq & 1, residual barrier parity = (seq >> 1) & 1
    int acc = 0;        // accumulator stage of the current tile and its barrier parity
    uint32_t acc_phase = 0;
    const int n_acc = g.n_acc;
    int first = grp;    // this group's first slab of the current tile
    bool res_issued = false;  // the residual of the group's next slab is already on its way
    TileIter it;
    tile_init<kCtas>(g, it, tile0, static_cast<int>(cta_rank));
#ifdef LVD_GEMM_PHASE_PROF
    long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_last = clock64();
#endif
    auto issue_res = [&](const TileIter& ti, int sl_, uint32_t seq_) {  // elected lane: TMA-load a residual slab
      if (LVD_DBG(g, 8)) {  // tuning aid: no residual traffic
        mbar_arrive(&my_res_full[seq_ & 1]);
        return;
      }
      mbar_arrive_expect_tx(&my_res_full[seq_ & 1], static_cast<uint32_t>(g.slab_bytes));
      // the residual lands in the slab's own output buffer: threads add to it in place
      tma_load_5d(my_out + (seq_ & 1) * kSlabBytes, &g.map_res, &my_res_full[seq_ & 1], ti.nt * out_cols + sl_ * 32,
                  ti.t0 * g.box[0], ti.t1 * g.box[1], ti.t2 * g.box[2], ti.t3 * g.box[3]);
    };
    // The accumulator goes back to the MMA warp as soon as this warp has pulled its last columns of the tile out of
    // TMEM -- before the slab is staged, fenced and stored (measured: releasing after the store made the MMA warp wait
    // for the accumulator ~2/3 of the time on 5-k-block tiles).
    auto release_acc = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_cluster(tempty_addr0 + 8u * acc); else mbar_arrive(&tempty[acc]);
      }
    };
    for (; it.tile < num_tiles; tile_advance(g, it), first = (first - nslabs) & 3) {
      bool valid = false;
      long long rvrow = 0;
      long long gn_slot0 = 0;  // first float of this warp's partial-sum block of the current tile
      if (has_gn) {
        const int rt = it.t0 + g.ntile[0] * (it.t1 + g.ntile[1] * (it.t2 + g.ntile[2] * it.t3));
        const int smp = rt / g.gn_tiles_per_sample, til = rt - smp * g.gn_tiles_per_sample;
        gn_slot0 = ((static_cast<long long>(smp) * g.gn_tiles_per_sample + til) * 4 + q) * g.gn_sub * g.gn_groups * 2;
      }
      float ln_mean = 0.f, ln_rstd = 1.f;
      long long lnstat_row = 0;
      if (has_rowvec || has_gn || has_ln || has_lnstat) {
        const int x0 = it.t0 * g.box[0] + j[0], x1 = it.t1 * g.box[1] + j[1];
        const int x2 = it.t2 * g.box[2] + j[2], x3 = it.t3 * g.box[3] + j[3];
        valid = j[3] < g.box[3] && x0 < g.out_dims[0] && x1 < g.out_dims[1] && x2 < g.out_dims[2] &&
                x3 < g.out_dims[3];
        rvrow = static_cast<long long>(x0) * g.rv_stride[0] + static_cast<long long>(x1) * g.rv_stride[1] +
                static_cast<long long>(x2) * g.rv_stride[2] + static_cast<long long>(x3) * g.rv_stride[3];
        if (has_ln && valid) {
          const float2 st = __ldg(g.ln_stats + static_cast<long long>(x0) * g.ln_stride[0] +
                                  static_cast<long long>(x1) * g.ln_stride[1] +
                                  static_cast<long long>(x2) * g.ln_stride[2] +
                                  static_cast<long long>(x3) * g.ln_stride[3]);
          ln_mean = st.x;
          ln_rstd = st.y;
        }
        if (has_lnstat)
          lnstat_row = static_cast<long long>(x0) * g.lnstat_stride[0] + static_cast<long long>(x1) * g.lnstat_stride[1] +
                       static_cast<long long>(x2) * g.lnstat_stride[2] + static_cast<long long>(x3) * g.lnstat_stride[3];
      }
      if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 0);  // epilogue group: waiting for the tile
      LVD_PH(7);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      LVD_PH(0);
      if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 1);  // tile's accumulator complete
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * block_n) + (static_cast<uint32_t>(q * 32) << 16);
      if (first >= nslabs) release_acc();  // no slab of this tile falls to this group
      for (int sl = first; sl < nslabs; sl += 4, ++seq) {
        const int gcol = it.nt * out_cols + sl * 32;  // first global output column of the slab
        uint8_t* obuf = my_out + (seq & 1) * kSlabBytes;
        if (has_res && !res_issued && leader_warp) {  // no prefetch happened (first slab, or a slab-less tile between)
          if (elect_one_sync()) issue_res(it, sl, seq);
          __syncwarp();
        }
        res_issued = false;
        LVD_PH(1);
        uint4 packed[4];
        float ls1 = 0.f, ls2 = 0.f;  // has_lnstat: sum / sum of squares of this row's 32 staged values
        if constexpr (kLean) {
          // ---- specialised epilogues: straight-line code, no feature switches.  The accumulator goes back to the MMA
          // warp right after this warp's last TMEM read of the tile has landed in registers.
          uint32_t pk[16];  // the row's 32 outputs as packed fp16 pairs
          [[maybe_unused]] float fv[32];  // ... and in fp32 (LayerNorm row statistics of a residual-free output)
          if constexpr (kEpi == kEpiGeglu || kEpi == kEpiLnGeglu) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // two 16-column halves: value and gate columns are both live
              const int ucol = sl * 32 + h * 16;          // tile-relative value column; its gate is n_half further
              const int tcol = it.nt * block_n + ucol;    // the same column of the interleaved weight / bias
              uint32_t ra[16], rg[16];
              tmem_ld16(taddr + ucol, ra);
              tmem_ld16(taddr + n_half + ucol, rg);
              float val[16], gate[16];
              if constexpr (kEpi == kEpiLnGeglu) {
                const float4* cv4 = reinterpret_cast<const float4*>(g.ln_cs + tcol);
                const float4* bv4 = reinterpret_cast<const float4*>(g.ln_bias + tcol);
                const float4* cg4 = reinterpret_cast<const float4*>(g.ln_cs + tcol + n_half);
                const float4* bg4 = reinterpret_cast<const float4*>(g.ln_bias + tcol + n_half);
                // the first four columns' constants travel under the TMEM read
                const float4 cv0 = __ldg(cv4), bv0 = __ldg(bv4), cg0 = __ldg(cg4), bg0 = __ldg(bg4);
                tmem_ld_wait();
                if (h == 1 && sl + 4 >= nslabs) release_acc();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 c1 = i == 0 ? cv0 : __ldg(cv4 + i), b1 = i == 0 ? bv0 : __ldg(bv4 + i);
                  const float4 c2 = i == 0 ? cg0 : __ldg(cg4 + i), b2 = i == 0 ? bg0 : __ldg(bg4 + i);
                  const float c1a[4] = {c1.x, c1.y, c1.z, c1.w}, b1a[4] = {b1.x, b1.y, b1.z, b1.w};
                  const float c2a[4] = {c2.x, c2.y, c2.z, c2.w}, b2a[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    val[4 * i + e] = fmaf(ln_rstd, fmaf(-ln_mean, c1a[e], __uint_as_float(ra[4 * i + e])), b1a[e]);
                    gate[4 * i + e] = fmaf(ln_rstd, fmaf(-ln_mean, c2a[e], __uint_as_float(rg[4 * i + e])), b2a[e]);
                  }
                }
              } else if (g.bias32) {
                const float4* bv4 = reinterpret_cast<const float4*>(g.bias32 + tcol);
                const float4* bg4 = reinterpret_cast<const float4*>(g.bias32 + tcol + n_half);
                float4 bv[4], bg[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  bv[i] = __ldg(bv4 + i);
                  bg[i] = __ldg(bg4 + i);
                }
                tmem_ld_wait();
                if (h == 1 && sl + 4 >= nslabs) release_acc();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  val[4 * i] = __uint_as_float(ra[4 * i]) + bv[i].x;
                  val[4 * i + 1] = __uint_as_float(ra[4 * i + 1]) + bv[i].y;
                  val[4 * i + 2] = __uint_as_float(ra[4 * i + 2]) + bv[i].z;
                  val[4 * i + 3] = __uint_as_float(ra[4 * i + 3]) + bv[i].w;
                  gate[4 * i] = __uint_as_float(rg[4 * i]) + bg[i].x;
                  gate[4 * i + 1] = __uint_as_float(rg[4 * i + 1]) + bg[i].y;
                  gate[4 * i + 2] = __uint_as_float(rg[4 * i + 2]) + bg[i].z;
                  gate[4 * i + 3] = __uint_as_float(rg[4 * i + 3]) + bg[i].w;
                }
              } else {
                float bv[16], bg[16];
                if (bias) {
                  unpack8(*reinterpret_cast<const uint4*>(bias + tcol), *reinterpret_cast<float(*)[8]>(&bv[0]));
                  unpack8(*reinterpret_cast<const uint4*>(bias + tcol + 8), *reinterpret_cast<float(*)[8]>(&bv[8]));
                  unpack8(*reinterpret_cast<const uint4*>(bias + tcol + n_half), *reinterpret_cast<float(*)[8]>(&bg[0]));
                  unpack8(*reinterpret_cast<const uint4*>(bias + tcol + n_half + 8),
                          *reinterpret_cast<float(*)[8]>(&bg[8]));
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) bv[i] = bg[i] = 0.f;
                }
                tmem_ld_wait();
                if (h == 1 && sl + 4 >= nslabs) release_acc();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  val[i] = __uint_as_float(ra[i]) + bv[i];
                  gate[i] = __uint_as_float(rg[i]) + bg[i];
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                pk[8 * h + i] = pack2h(val[2 * i] * gelu_erf_lean(gate[2 * i]), val[2 * i + 1] * gelu_erf_lean(gate[2 * i + 1]));
            }
          } else {
            uint32_t rv[32];
            tmem_ld32(taddr + sl * 32, rv);
            if constexpr (kEpi == kEpiLnBias) {
              const float4* cs4 = reinterpret_cast<const float4*>(g.ln_cs + gcol);
              const float4* lb4 = reinterpret_cast<const float4*>(g.ln_bias + gcol);
              float4 c0[2], b0[2];  // the first eight columns' constants travel under the TMEM read
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                c0[i] = __ldg(cs4 + i);
                b0[i] = __ldg(lb4 + i);
              }
              tmem_ld_wait();
              if (sl + 4 >= nslabs) release_acc();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 c = i < 2 ? c0[i & 1] : __ldg(cs4 + i), b = i < 2 ? b0[i & 1] : __ldg(lb4 + i);
                fv[4 * i] = fmaf(ln_rstd, fmaf(-ln_mean, c.x, __uint_as_float(rv[4 * i])), b.x);
                fv[4 * i + 1] = fmaf(ln_rstd, fmaf(-ln_mean, c.y, __uint_as_float(rv[4 * i + 1])), b.y);
                fv[4 * i + 2] = fmaf(ln_rstd, fmaf(-ln_mean, c.z, __uint_as_float(rv[4 * i + 2])), b.z);
                fv[4 * i + 3] = fmaf(ln_rstd, fmaf(-ln_mean, c.w, __uint_as_float(rv[4 * i + 3])), b.w);
              }
            } else if (g.bias32) {
              const float4* b4 = reinterpret_cast<const float4*>(g.bias32 + gcol);
              float4 b[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) b[i] = __ldg(b4 + i);
              tmem_ld_wait();
              if (sl + 4 >= nslabs) release_acc();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                fv[4 * i] = __uint_as_float(rv[4 * i]) + b[i].x;
                fv[4 * i + 1] = __uint_as_float(rv[4 * i + 1]) + b[i].y;
                fv[4 * i + 2] = __uint_as_float(rv[4 * i + 2]) + b[i].z;
                fv[4 * i + 3] = __uint_as_float(rv[4 * i + 3]) + b[i].w;
              }
            } else if (bias) {
              uint4 bh[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) bh[i] = *reinterpret_cast<const uint4*>(bias + gcol + 8 * i);
              tmem_ld_wait();
              if (sl + 4 >= nslabs) release_acc();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float t[8];
                unpack8(bh[i], t);
#pragma unroll
                for (int e = 0; e < 8; ++e) fv[8 * i + e] = __uint_as_float(rv[8 * i + e]) + t[e];
              }
            } else {
              tmem_ld_wait();
              if (sl + 4 >= nslabs) release_acc();
#pragma unroll
              for (int i = 0; i < 32; ++i) fv[i] = __uint_as_float(rv[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack2h(fv[2 * i], fv[2 * i + 1]);
            if constexpr (kEpi == kEpiResidual) {
              // out = fp16(linear) + residual, one fp16 addition like the reference's `hidden_states + residual`; the
              // residual slab was TMA-loaded into this slab's own staging buffer
              mbar_wait(&my_res_full[seq & 1], (seq >> 1) & 1u);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 t = *reinterpret_cast<const uint4*>(obuf + row_off + ((static_cast<uint32_t>(c) ^ sw) << 4));
                pk[4 * c] = hadd2_u32(pk[4 * c], t.x);
                pk[4 * c + 1] = hadd2_u32(pk[4 * c + 1], t.y);
                pk[4 * c + 2] = hadd2_u32(pk[4 * c + 2], t.z);
                pk[4 * c + 3] = hadd2_u32(pk[4 * c + 3], t.w);
              }
              if (has_lnstat) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pk[i]));
                  fv[2 * i] = f.x;
                  fv[2 * i + 1] = f.y;
                }
              }
            }
            if (has_lnstat) {  // (of the unrounded values when there is no residual: a 1e-5 relative change of rstd)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                ls1 += fv[i];
                ls2 = fmaf(fv[i], fv[i], ls2);
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) packed[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two 16-column halves: bounds the live registers
          const int ucol = sl * 32 + h * 16;  // tile-relative accumulator column
          const int ocol = gcol + h * 16;     // global output column
          uint32_t rv[16];
          float v[16];
          tmem_ld16(taddr + ucol, rv);
          if (geglu && has_ln) {
            uint32_t rg[16];
            tmem_ld16(taddr + n_half + ucol, rg);
            const int tcol = it.nt * block_n + ucol;  // value at tcol, gate at tcol + n_half of the interleaved weight
            tmem_ld_wait();
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 cv = __ldg(reinterpret_cast<const float4*>(g.ln_cs + tcol) + i4);
              const float4 bv = __ldg(reinterpret_cast<const float4*>(g.ln_bias + tcol) + i4);
              const float4 cg = __ldg(reinterpret_cast<const float4*>(g.ln_cs + tcol + n_half) + i4);
              const float4 bg = __ldg(reinterpret_cast<const float4*>(g.ln_bias + tcol + n_half) + i4);
              const float cva[4] = {cv.x, cv.y, cv.z, cv.w}, bva[4] = {bv.x, bv.y, bv.z, bv.w};
              const float cga[4] = {cg.x, cg.y, cg.z, cg.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * i4 + e;
                const float val = fmaf(ln_rstd, fmaf(-ln_mean, cva[e], __uint_as_float(rv[i])), bva[e]);
                const float gate = fmaf(ln_rstd, fmaf(-ln_mean, cga[e], __uint_as_float(rg[i])), bga[e]);
                v[i] = val * gelu_erf_f(gate);
              }
            }
          } else if (geglu) {
            uint32_t rg[16];
            tmem_ld16(taddr + n_half + ucol, rg);
            // value * gelu(gate); biases of the interleaved weight: value at tcol, gate at tcol + n_half
            const int tcol = it.nt * block_n + ucol;
            float bv[16], bg[16];
            if (bias) {
              unpack8(*reinterpret_cast<const uint4*>(bias + tcol), *reinterpret_cast<float(*)[8]>(&bv[0]));
              unpack8(*reinterpret_cast<const uint4*>(bias + tcol + 8), *reinterpret_cast<float(*)[8]>(&bv[8]));
              unpack8(*reinterpret_cast<const uint4*>(bias + tcol + n_half), *reinterpret_cast<float(*)[8]>(&bg[0]));
              unpack8(*reinterpret_cast<const uint4*>(bias + tcol + n_half + 8),
                      *reinterpret_cast<float(*)[8]>(&bg[8]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) bv[i] = bg[i] = 0.f;
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (__uint_as_float(rv[i]) + bv[i]) * gelu_erf_f(__uint_as_float(rg[i]) + bg[i]);
          } else if (has_ln) {
            tmem_ld_wait();
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 c = __ldg(reinterpret_cast<const float4*>(g.ln_cs + ocol) + i4);
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.ln_bias + ocol) + i4);
              v[4 * i4] = fmaf(ln_rstd, fmaf(-ln_mean, c.x, __uint_as_float(rv[4 * i4])), b.x);
              v[4 * i4 + 1] = fmaf(ln_rstd, fmaf(-ln_mean, c.y, __uint_as_float(rv[4 * i4 + 1])), b.y);
              v[4 * i4 + 2] = fmaf(ln_rstd, fmaf(-ln_mean, c.z, __uint_as_float(rv[4 * i4 + 2])), b.z);
              v[4 * i4 + 3] = fmaf(ln_rstd, fmaf(-ln_mean, c.w, __uint_as_float(rv[4 * i4 + 3])), b.w);
            }
          } else {
            if (bias) {
              float b[16];
              unpack8(*reinterpret_cast<const uint4*>(bias + ocol), *reinterpret_cast<float(*)[8]>(&b[0]));
              unpack8(*reinterpret_cast<const uint4*>(bias + ocol + 8), *reinterpret_cast<float(*)[8]>(&b[8]));
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rv[i]) + b[i];
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rv[i]);
            }
            if (has_silu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = silu_f(round_h(v[i]));
            }
            if (has_gelu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = gelu_erf_f(round_h(v[i]));
            }
          }
          if (has_rowvec && valid) {
            float t[16];
            unpack8(*reinterpret_cast<const uint4*>(g.rowvec + rvrow * g.ld_rowvec + ocol),
                    *reinterpret_cast<float(*)[8]>(&t[0]));
            unpack8(*reinterpret_cast<const uint4*>(g.rowvec + rvrow * g.ld_rowvec + ocol + 8),
                    *reinterpret_cast<float(*)[8]>(&t[8]));
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = round_h(v[i]) + t[i];
          }
          if (has_gamma) {
            const float gm = g.gamma;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = round_h(v[i]) * gm;
          }
          if (has_res) {
            if (h == 0) mbar_wait(&my_res_full[seq & 1], (seq >> 1) & 1u);
            float t[16];
            unpack8(*reinterpret_cast<const uint4*>(obuf + row_off + (((2u * h) ^ sw) << 4)),
                    *reinterpret_cast<float(*)[8]>(&t[0]));
            unpack8(*reinterpret_cast<const uint4*>(obuf + row_off + (((2u * h + 1u) ^ sw) << 4)),
                    *reinterpret_cast<float(*)[8]>(&t[8]));
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = round_h(v[i]) + t[i];
          }
          packed[2 * h] = pack8(&v[0]);
          packed[2 * h + 1] = pack8(&v[8]);
          if (has_lnstat) {  // statistics of the fp16 values the LayerNorm will read
            unpack8(packed[2 * h], *reinterpret_cast<float(*)[8]>(&v[0]));
            unpack8(packed[2 * h + 1], *reinterpret_cast<float(*)[8]>(&v[8]));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              ls1 += v[i];
              ls2 = fmaf(v[i], v[i], ls2);
            }
          }
        }
        }
        if (has_lnstat && valid)
          g.lnstat_part[static_cast<long long>(gcol >> 5) * g.lnstat_rows + lnstat_row] = make_float2(ls1, ls2);
        LVD_PH(2);
        if (!kLean && sl + 4 >= nslabs) release_acc();  // this warp's last TMEM reads of the tile have completed
        if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 2);  // slab math done
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx)
          *reinterpret_cast<uint4*>(obuf + row_off + ((static_cast<uint32_t>(cidx) ^ sw) << 4)) = packed[cidx];
        if (has_gn) {
          // GroupNorm statistics of the values just staged (the fp16 the consumer will read): lane c sums column c of the
          // slab over this warp's 32 rows, a segmented shuffle reduction folds the columns of one channel group, and the
          // head lane of each group adds the pair into the warp's own partial slot.  A group spans at most gn_sub
          // consecutive slabs and slabs go to slots by (slab index mod gn_sub), so every float has exactly one writer
          // per launch: the result does not depend on scheduling (bit-reproducible), the adds only serve zero-init.
          __syncwarp();
          const unsigned vmask = __ballot_sync(0xffffffffu, valid);
          const uint32_t chunk = static_cast<uint32_t>(lane) >> 3, within = (static_cast<uint32_t>(lane) & 7u) * 2u;
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            if ((vmask >> rr) & 1u) {
              const uint32_t row = static_cast<uint32_t>(q * 32 + rr);
              const float f = __half2float(*reinterpret_cast<const __half*>(
                  obuf + row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4) + within));
              s1 += f;
              s2 = fmaf(f, f, s2);
            }
          }
          const int gid = (gcol + lane) / g.gn_gw;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const float t1 = __shfl_down_sync(0xffffffffu, s1, off), t2 = __shfl_down_sync(0xffffffffu, s2, off);
            const int g2 = __shfl_down_sync(0xffffffffu, gid, off);
            if (lane + off < 32 && g2 == gid) {
              s1 += t1;
              s2 += t2;
            }
          }
          const int gprev = __shfl_up_sync(0xffffffffu, gid, 1);
          if (lane == 0 || gprev != gid) {
            float* dst = g.gn_part + gn_slot0 + (static_cast<long long>((gcol >> 5) % g.gn_sub) * g.gn_groups + gid) * 2;
            atomicAdd(dst, s1);
            atomicAdd(dst + 1, s2);
          }
        }
        fence_proxy_async_smem();
        LVD_PH(3);
        // the group's other buffer must have been drained (store seq-1) before anyone writes it in the next slab
        if (leader_warp) {
          if (elect_one_sync()) bulk_wait_read_all();
          __syncwarp();
        }
        if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 3);  // fence + previous store drained
        LVD_PH(4);
        named_bar_sync(bar_id, 128);
        LVD_PH(5);
        if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 4);  // group barrier passed
        if (leader_warp) {
          if (elect_one_sync()) {  // elect.sync is deterministic: the same lane owns the bulk async-groups throughout
            if (!LVD_DBG(g, 4))  // tuning aid: bit 2 drops the output stores
              tma_store_5d(&g.map_out, obuf, gcol, it.t0 * g.box[0], it.t1 * g.box[1], it.t2 * g.box[2],
                           it.t3 * g.box[3]);
            bulk_commit_group();
          }
          __syncwarp();
        }
        if (has_res) {  // prefetch the residual of this group's next slab: same tile, or the next tile's first slab
          if (sl + 4 < nslabs) {
            if (leader_warp) {
              if (elect_one_sync()) issue_res(it, sl + 4, seq + 1);
              __syncwarp();
            }
            res_issued = true;
          } else {
            const int nf = (first - nslabs) & 3;
            if (nf < nslabs && it.tile + g.tile_step < num_tiles) {
              if (leader_warp) {
                if (elect_one_sync()) {
                  TileIter nx = it;
                  tile_advance(g, nx);
                  issue_res(nx, nf, seq + 1);
                }
                __syncwarp();
              }
              res_issued = true;
            }
          }
        }
      }
      if (leader_warp && lane == 0) trace_event(g, 16 + grp * 16 + 5);  // tile done
      LVD_PH(6);
      if (++acc == n_acc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (leader_warp) {  // the staging slabs must outlive the last store's reads
      if (elect_one_sync()) bulk_wait_all();
      __syncwarp();
    }
#ifdef LVD_GEMM_PHASE_PROF
    LVD_PH(8);
    if (blockIdx.x == 0 && lane == 0)
      for (int i = 0; i < 10; ++i) g_trace[warp * 16 + i] = static_cast<unsigned long long>(ph[i]);
    if (blockIdx.x == 0 && threadIdx.x == 64) g_trace_n = 18 * 16;
#endif
  }
  tc_fence_before();
  // the peer's shared memory / TMEM / barriers must stay alive until the pair's last MMA and commit have landed
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static int launch_gemm(const lvd_gemm_desc* d, cudaStream_t stream) {
  LVD_REQUIRE(d && d->a && d->b && d->out, "lvd_gemm: null pointer");
  LVD_REQUIRE(d->block_n == 64 || d->block_n == 128 || d->block_n == 160 || d->block_n == 256,
              "lvd_gemm: block_n %d unsupported", d->block_n);
  LVD_REQUIRE(d->n > 0 && d->n % d->block_n == 0, "lvd_gemm: n=%d not a multiple of block_n=%d", d->n, d->block_n);
  LVD_REQUIRE(d->k_per_tap > 0 && d->k_per_tap % 64 == 0, "lvd_gemm: k_per_tap=%d not a multiple of 64",
              d->k_per_tap);
  LVD_REQUIRE(d->n_taps >= 1 && d->n_taps <= LVD_MAX_TAPS, "lvd_gemm: n_taps=%d", d->n_taps);
  LVD_REQUIRE(d->act >= 0 && d->act <= 3, "lvd_gemm: act=%d", d->act);
  LVD_REQUIRE(d->act != 2 || d->block_n == 128 || d->block_n == 256, "lvd_gemm: GEGLU needs block_n 128 or 256");
  LVD_REQUIRE(!d->b_mn_major || d->block_n % 64 == 0, "lvd_gemm: MN-major B needs block_n %% 64 == 0");
  LVD_REQUIRE(d->a_box[0] == 64, "lvd_gemm: a_box[0] must be 64");
  LVD_REQUIRE(d->cta_pair >= 0 && d->cta_pair <= 2, "lvd_gemm: cta_pair=%d", d->cta_pair);
  long long rows = 1;
  for (int i = 1; i < 5; ++i) rows *= d->a_box[i];
  LVD_REQUIRE(rows >= 1 && rows <= 128, "lvd_gemm: tile has %lld rows (max 128)", rows);
  LVD_REQUIRE((d->ld_out % 8) == 0 && (!d->residual || d->ld_res % 8 == 0) && (!d->rowvec || d->ld_rowvec % 8 == 0),
              "lvd_gemm: leading dimensions must be multiples of 8");

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  long long m_tiles = 1;
  for (int i = 0; i < 4; ++i) {
    g.box[i] = d->a_box[i + 1];
    g.out_dims[i] = d->out_dims[i];
    LVD_REQUIRE(d->out_dims[i] >= 1, "lvd_gemm: out_dims[%d]=%d", i, d->out_dims[i]);
    g.ntile[i] = (d->out_dims[i] + g.box[i] - 1) / g.box[i];
    m_tiles *= g.ntile[i];
    g.rv_stride[i] = d->rv_stride[i];
  }
  // CTA pairs (2-CTA clusters, cta_group::2 MMA with M = 256): each SM stages only half of every B tile.  Measured on
  // the Zeroscope shapes (tools/perf_gemm_shapes.py): +4..15% from K = 640 up, but -5..15% on the 5-k-block GEMMs
  // (K = 320), which are bound by the output write and lose from coupling two CTAs' epilogues.  MN-major B is staged
  // in 64-column chunks, so its half tile must be a multiple of 64.
  const long long k_contract = static_cast<long long>(d->n_taps) * d->k_per_tap;
  const bool pair_ok = m_tiles >= 2 && (!d->b_mn_major || d->block_n % 128 == 0);
  LVD_REQUIRE(d->cta_pair != 2 || pair_ok, "lvd_gemm: cta_pair=2 needs >= 2 row tiles (MN-major B: block_n %% 128 == 0)");
  const bool pair = d->cta_pair == 2 || (d->cta_pair == 0 && pair_ok && k_contract >= 640);
  const int ctas = pair ? 2 : 1;
  const int b_rows = d->block_n / ctas;

  int rc = encode_tmap_f16_sw128(&g.map_a, d->a, 5, d->a_dims, d->a_strides, d->a_box);
  if (rc) return rc;
  const int k_tap_stride = (d->b_k_tap_stride == 0 && d->b_n_tap_stride == 0) ? d->k_per_tap : d->b_k_tap_stride;
  const long long k_total = static_cast<long long>(d->n_taps - 1) * k_tap_stride + d->k_per_tap;
  const long long n_total = static_cast<long long>(d->n_taps - 1) * d->b_n_tap_stride + d->n;
  if (!d->b_mn_major) {
    int64_t dims[2] = {k_total, n_total};
    int64_t str[2] = {1, d->ldb};
    int box[2] = {64, b_rows};
    rc = encode_tmap_f16_sw128(&g.map_b, d->b, 2, dims, str, box);
  } else {
    int64_t dims[2] = {n_total, k_total};
    int64_t str[2] = {1, d->ldb};
    int box[2] = {64, 64};
    rc = encode_tmap_f16_sw128(&g.map_b, d->b, 2, dims, str, box);
  }
  if (rc) return rc;
  for (int t = 0; t < d->n_taps; ++t)
    for (int i = 0; i < 5; ++i) g.tap_off[t][i] = d->tap_off[t][i];
  g.n_taps = d->n_taps;
  g.kb_per_tap = d->k_per_tap / 64;
  g.block_n = d->block_n;
  g.n_tiles_n = d->n / d->block_n;
  g.b_mn_major = d->b_mn_major;
  g.b_k_tap_stride = k_tap_stride;
  g.b_n_tap_stride = d->b_n_tap_stride;
  g.act = d->act;
  const long long row_tiles = pair ? (m_tiles + 1) / 2 : m_tiles;  // an odd tail tile is paired with an all-OOB tile
  LVD_REQUIRE(row_tiles * g.n_tiles_n < (1ll << 30), "lvd_gemm: too many tiles");
  g.num_tiles = static_cast<int>(row_tiles * g.n_tiles_n);
  g.a_bytes = static_cast<int>(rows) * 128;
  g.b_bytes = b_rows * 128;
  g.slab_bytes = static_cast<int>(rows) * 64;
  g.n_acc = 512 / d->block_n > 4 ? 4 : 512 / d->block_n;
  // grid and the mixed-radix step of the persistent tile walk
  const int max_units = pair ? num_sms() / 2 : num_sms();
  int units = g.num_tiles < max_units ? g.num_tiles : max_units;  // CTAs, or CTA pairs
  // Weight-resident mode (linears with a short K): with a walk step that is a multiple of n_tiles_n every CTA (pair) stays
  // on one n-tile, keeps that B tile in shared memory for the whole launch and streams only A -- the operand traffic from
  // L2 (the bound of the K = 320 / 640 linears: ~10.7 TB/s chip-wide) drops from (A + B) to A per tile.
  const int kblocks_total = d->n_taps * (d->k_per_tap / 64);
  const int b_res_bytes = kblocks_total * b_rows * 128;
  int stages_res = (kSmemLimit - 1024 - kBarrierBytes - kSlabBufs * kSlabBytes - b_res_bytes) / kAStageBytes;
  if (stages_res > 8) stages_res = 8;
  const bool resident = d->weight_resident == 1 && d->n_taps == 1 && !d->b_mn_major && stages_res >= 3 &&
                        g.n_tiles_n <= units;
  if (resident) units = units / g.n_tiles_n * g.n_tiles_n;
  g.b_resident = resident ? 1 : 0;
  g.tile_step = units;
  g.step_nt = units % g.n_tiles_n;
  for (int hi = 0; hi < 2; ++hi) {
    long long delta = static_cast<long long>(ctas) * (units / g.n_tiles_n + hi);
    int* e = hi ? g.step_hi : g.step_lo;
    for (int i = 0; i < 3; ++i) {
      e[i] = static_cast<int>(delta % g.ntile[i]);
      delta /= g.ntile[i];
    }
    e[3] = static_cast<int>(delta);
  }
  {
    // output / residual as rank-5 tensors whose axes 1..4 are the tile axes: stride_i = out_row_stride[i] * ld
    const int n_out = d->act == 2 ? d->n / 2 : d->n;
    int64_t od[5] = {n_out, d->out_dims[0], d->out_dims[1], d->out_dims[2], d->out_dims[3]};
    int ob[5] = {32, d->a_box[1], d->a_box[2], d->a_box[3], d->a_box[4]};
    int64_t os[5], rs[5];
    os[0] = rs[0] = 1;
    for (int i = 0; i < 4; ++i) {
      LVD_REQUIRE(d->out_row_stride[i] > 0 || d->out_dims[i] == 1, "lvd_gemm: out_row_stride[%d] == 0 on an axis of extent %d",
                  i, d->out_dims[i]);
      const int64_t st = d->out_row_stride[i] > 0 ? d->out_row_stride[i] : 1;
      os[i + 1] = st * d->ld_out;
      rs[i + 1] = st * d->ld_res;
    }
    rc = encode_tmap_f16_sw64(&g.map_out, d->out, 5, od, os, ob);
    if (rc) return rc;
    if (d->residual) {
      rc = encode_tmap_f16_sw64(&g.map_res, d->residual, 5, od, rs, ob);
      if (rc) return rc;
    }
  }
  const int stage_bytes = kAStageBytes + g.b_bytes;
  const int n_slabs_smem = kSlabBufs;
  int stages = (kSmemLimit - 1024 - kBarrierBytes - n_slabs_smem * kSlabBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (resident) stages = stages_res;
  {  // tuning aid (tools/perf_gemm_sweep.py): cap the ring depth
    static const char* env = getenv("LVD_GEMM_STAGES");
    if (env && atoi(env) >= 2 && atoi(env) < stages) stages = atoi(env);
  }
  g.stages = stages;
  g.b_slots = resident ? kblocks_total : stages;
  {
    static const char* env = getenv("LVD_GEMM_DEBUG");
    g.debug = env ? atoi(env) : 0;
  }
  g.bias = static_cast<const __half*>(d->bias);
  g.bias32 = static_cast<const float*>(d->bias_f32);
  g.rowvec = static_cast<const __half*>(d->rowvec);
  g.residual = static_cast<const __half*>(d->residual);
  g.ld_rowvec = d->ld_rowvec;
  g.gamma = d->gamma;
  g.gn_part = static_cast<float*>(d->gn_partial);
  g.gn_gw = d->gn_group_width;
  g.gn_groups = d->gn_groups;
  g.gn_tiles_per_sample = d->gn_tiles_per_sample;
  g.gn_sub = d->gn_subslots;
  if (g.gn_part) {
    LVD_REQUIRE(d->act != 2 && g.gn_gw >= 1 && g.gn_groups >= 1 && d->n == g.gn_gw * g.gn_groups,
                "lvd_gemm: gn statistics need n == groups * group_width and no GEGLU");
    LVD_REQUIRE(g.gn_tiles_per_sample >= 1 && m_tiles % g.gn_tiles_per_sample == 0,
                "lvd_gemm: gn_tiles_per_sample=%d does not divide the %lld row tiles", g.gn_tiles_per_sample, m_tiles);
    LVD_REQUIRE(g.gn_sub >= (g.gn_gw + 30) / 32 + 1, "lvd_gemm: gn_subslots=%d too small for groups of %d channels",
                g.gn_sub, g.gn_gw);
  }
  g.ln_stats = static_cast<const float2*>(d->ln_stats);
  g.ln_cs = static_cast<const float*>(d->ln_colsum);
  g.ln_bias = static_cast<const float*>(d->ln_bias);
  g.lnstat_part = static_cast<float2*>(d->lnstat_partial);
  g.lnstat_rows = d->lnstat_rows;
  for (int i = 0; i < 4; ++i) {
    g.ln_stride[i] = d->ln_row_stride[i];
    g.lnstat_stride[i] = d->lnstat_row_stride[i];
  }
  if (g.ln_stats) {
    LVD_REQUIRE(g.ln_cs && g.ln_bias, "lvd_gemm: ln_stats needs ln_colsum and ln_bias");
    LVD_REQUIRE(d->rowvec == nullptr && d->gamma == 1.0f && !d->residual && !d->bias && (d->act == 0 || d->act == 2) &&
                    d->n_taps == 1 && !g.gn_part,
                "lvd_gemm: a folded LayerNorm needs a plain linear (act none / GEGLU; bias inside ln_bias)");
  }
  if (g.lnstat_part)
    LVD_REQUIRE(d->act != 2 && g.lnstat_rows >= 1, "lvd_gemm: lnstat_partial needs lnstat_rows and no GEGLU");
  g.idesc = make_idesc_f16(128u * ctas, static_cast<uint32_t>(d->block_n), 0, d->b_mn_major ? 1u : 0u);

  const int smem_bytes = 1024 + stages * kAStageBytes + g.b_slots * g.b_bytes + n_slabs_smem * kSlabBytes + kBarrierBytes;
  const bool plain = d->rowvec == nullptr && d->gamma == 1.0f;
  int epi = kEpiGeneric;
  if (plain && d->act == 2 && !d->residual) epi = kEpiGeglu;
  if (plain && d->act == 0) epi = d->residual ? kEpiResidual : kEpiBias;
  if (g.ln_stats) epi = d->act == 2 ? kEpiLnGeglu : kEpiLnBias;
  // statistics emitted by the epilogue: compile-time for the specialised epilogues (only Bias / Residual carry them)
  int feat = 0;
  if (g.lnstat_part) feat = 1;
  if (g.gn_part) feat = 2;
  LVD_REQUIRE(!(g.lnstat_part && g.gn_part), "lvd_gemm: lnstat_partial and gn_partial are mutually exclusive");
  if (feat != 0 && epi != kEpiBias && epi != kEpiResidual) {
    LVD_REQUIRE(!g.ln_stats, "lvd_gemm: a folded LayerNorm cannot emit output statistics");
    epi = kEpiGeneric;
  }
  typedef void (*KernelFn)(const GemmArgs);
  // [ctas - 1][slot]: 0 generic | 1..3 bias (none, lnstat, gn) | 4..6 residual (none, lnstat, gn) | 7 geglu |
  // 8 ln+bias | 9 ln+geglu
  static const KernelFn kernels[2][10] = {
      {gemm_tc_kernel<1, kEpiGeneric, -1>, gemm_tc_kernel<1, kEpiBias, 0>, gemm_tc_kernel<1, kEpiBias, 1>,
       gemm_tc_kernel<1, kEpiBias, 2>, gemm_tc_kernel<1, kEpiResidual, 0>, gemm_tc_kernel<1, kEpiResidual, 1>,
       gemm_tc_kernel<1, kEpiResidual, 2>, gemm_tc_kernel<1, kEpiGeglu, 0>, gemm_tc_kernel<1, kEpiLnBias, 0>,
       gemm_tc_kernel<1, kEpiLnGeglu, 0>},
      {gemm_tc_kernel<2, kEpiGeneric, -1>, gemm_tc_kernel<2, kEpiBias, 0>, gemm_tc_kernel<2, kEpiBias, 1>,
       gemm_tc_kernel<2, kEpiBias, 2>, gemm_tc_kernel<2, kEpiResidual, 0>, gemm_tc_kernel<2, kEpiResidual, 1>,
       gemm_tc_kernel<2, kEpiResidual, 2>, gemm_tc_kernel<2, kEpiGeglu, 0>, gemm_tc_kernel<2, kEpiLnBias, 0>,
       gemm_tc_kernel<2, kEpiLnGeglu, 0>}};
  int slot = 0;
  if (epi == kEpiBias) slot = 1 + feat;
  if (epi == kEpiResidual) slot = 4 + feat;
  if (epi == kEpiGeglu) slot = 7;
  if (epi == kEpiLnBias) slot = 8;
  if (epi == kEpiLnGeglu) slot = 9;
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 10; ++j)
        LVD_CHECK_CUDA(cudaFuncSetAttribute(kernels[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(ctas * units));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(ctas);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LVD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernels[ctas - 1][slot], g));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

}  // namespace lvd

extern "C" int lvd_gemm(const lvd_gemm_desc* d, void* stream) {
  return lvd::launch_gemm(d, static_cast<cudaStream_t>(stream));
}
// tools/trace_gemm.py: copy out (and reset) the timeline CTA 0 recorded under LVD_GEMM_DEBUG=64
extern "C" int lvd_debug_trace_read(unsigned long long* host_buf, int max_events) {
  unsigned int n = 0;
  if (cudaMemcpyFromSymbol(&n, lvd::g_trace_n, sizeof(n)) != cudaSuccess) return -1;
  if (n > 8192u) n = 8192u;
  if (static_cast<int>(n) > max_events) n = static_cast<unsigned>(max_events);
  if (n && cudaMemcpyFromSymbol(host_buf, lvd::g_trace, n * sizeof(unsigned long long)) != cudaSuccess) return -1;
  unsigned int zero = 0;
  cudaMemcpyToSymbol(lvd::g_trace_n, &zero, sizeof(zero));
  return static_cast<int>(n);
}
extern "C" int lvd_version(void) { return 1; }
extern "C" int64_t lvd_launch_count(void) { return lvd::g_launches.load(); }
