// gemm.cu -- persistent, warp-specialised tcgen05 contraction with fused epilogue (see include/lvd_b200.h).
//
//   warp 0      : TMA producer   (one elected lane; cp.async.bulk.tensor into a multi-stage smem ring)
//   warp 1      : MMA issuer     (one elected lane; tcgen05.mma, fp32 accumulators double-buffered in TMEM)
//   warps 2..17 : epilogue       (tcgen05.ld -> bias / SiLU / GEGLU / time-embedding / gate / residual -> fp16 store;
//                 four warps per TMEM lane quarter, interleaved over 16-column units)
//
// A tiles are rank-5 TMA boxes of the channels-last activation; a convolution tap is a coordinate offset and the
// zero padding is TMA out-of-bounds fill, so Linear, Conv2d (3x3, 1x1, stride 2) and Conv3d (3,1,1) all run here.
#include <atomic>

#include "../../include/lvd_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace lvd {

std::atomic<long long> g_launches{0};

constexpr int kAStageBytes = 128 * 128;  // 128 rows x 64 halves
constexpr int kThreads = 576;  // 2 control warps + 16 epilogue warps
constexpr int kSmemLimit = 232448;  // 227 KB
constexpr int kSlabBytes = 128 * 64;  // epilogue staging slab: 128 rows x 32 halves

struct GemmArgs {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_out;  // 32-column slabs of the output (64-byte swizzle), rows mapped like the A box
  CUtensorMap map_res;  // same geometry over the residual
  long long out_row_stride[4];
  long long ld_rowvec, ld_res, ld_out;
  const __half* bias;
  const __half* rowvec;
  const __half* residual;
  __half* out;
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
  float gamma;
  uint32_t idesc;
};

__device__ __forceinline__ void tile_coords(const GemmArgs& g, int mt, int (&c)[4]) {
  int t0 = mt % g.ntile[0];
  int r = mt / g.ntile[0];
  int t1 = r % g.ntile[1];
  r /= g.ntile[1];
  int t2 = r % g.ntile[2];
  int t3 = r / g.ntile[2];
  c[0] = t0 * g.box[0];
  c[1] = t1 * g.box[1];
  c[2] = t2 * g.box[2];
  c[3] = t3 * g.box[3];
}

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte aligned stage bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = g.stages;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(g.block_n) * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + stages * kAStageBytes;
  uint8_t* sOut = sB + stages * b_stage_bytes;  // output slabs [128 rows][32 cols] fp16: 2 per epilogue group
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + 4 * kSlabBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_full = tempty + 2;  // [group][buffer]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_full + 4);

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 512);
    mbar_init(&tempty[1], 512);
    for (int i = 0; i < 4; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&g.map_a);
    tma_prefetch_desc(&g.map_b);
    tma_prefetch_desc(&g.map_out);
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int kblocks = g.n_taps * g.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int nt = tile % g.n_tiles_n;
        int c[4];
        tile_coords(g, tile / g.n_tiles_n, c);
        for (int tap = 0; tap < g.n_taps; ++tap) {
          const int o0 = g.tap_off[tap][0];
          const int c1 = c[0] + g.tap_off[tap][1], c2 = c[1] + g.tap_off[tap][2];
          const int c3 = c[2] + g.tap_off[tap][3], c4 = c[3] + g.tap_off[tap][4];
          for (int kb = 0; kb < g.kb_per_tap; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(g.a_bytes + g.b_bytes));
            tma_load_5d(sA + stage * kAStageBytes, &g.map_a, &full[stage], o0 + kb * 64, c1, c2, c3, c4);
            const int kcoord = tap * g.b_k_tap_stride + kb * 64;
            const int ncoord = nt * g.block_n + tap * g.b_n_tap_stride;
            uint8_t* bdst = sB + stage * b_stage_bytes;
            if (!g.b_mn_major) {
              tma_load_2d(bdst, &g.map_b, &full[stage], kcoord, ncoord);
            } else {
              for (int ch = 0; ch < g.block_n / 64; ++ch)
                tma_load_2d(bdst + ch * 8192, &g.map_b, &full[stage], ncoord + ch * 64, kcoord);
            }
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kAStageBytes);
          const uint32_t b_addr = smem_u32(sB + stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = g.b_mn_major ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                             : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, da, db, g.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // 16 epilogue warps in two groups of 8 (256 threads).  The tile's output is emitted as 32-column slabs: group 0
    // takes the even slabs, group 1 the odd ones.  Per slab: [leader] wait until the group's staging buffer has been
    // drained by the previous TMA store and kick the residual slab's TMA load -> named barrier -> every thread reads
    // its 16 accumulator columns from TMEM, applies the epilogue and writes two 16-byte chunks of its row into the
    // swizzled staging slab -> named barrier -> [leader] one TMA store of the slab.  Global traffic is therefore
    // full-line bulk copies issued by the TMA engine (row mapping and edge clipping come from the tensor map), not
    // 32 scattered 16-byte accesses per warp instruction.
    const int q = warp & 3;            // this warp may only touch TMEM lanes [32q, 32q+32)
    const int wi = warp - 2;           // 0..15
    const int grp = wi >> 3;           // 0 / 1
    const int usub = (wi >> 2) & 1;    // which 16-column half of the slab
    const bool leader = (wi & 7) == 0 && lane == 0;
    const int r = q * 32 + lane;
    int j[4];
    {
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
    // 64-byte swizzle: physical = logical ^ (((logical >> 7) & 3) << 4), logical = r*64 + chunk*16
    const uint32_t sw = static_cast<uint32_t>((r >> 1) & 3);
    const uint32_t off0 = static_cast<uint32_t>(r * 64) + (((static_cast<uint32_t>(usub * 2)) ^ sw) << 4);
    const uint32_t off1 = static_cast<uint32_t>(r * 64) + (((static_cast<uint32_t>(usub * 2 + 1)) ^ sw) << 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool geglu = (g.act == 2);
    const int nslabs = geglu ? 2 : g.block_n / 32;
    const int bar_id = 1 + grp;
    uint32_t seq = 0;  // slabs processed by this group so far: buffer = seq & 1, residual parity = (seq >> 1) & 1
    auto issue_res = [&](int tile_, int sl_, uint32_t seq_) {  // leader only: TMA-load the residual slab (tile_, sl_)
      int cc[4];
      tile_coords(g, tile_ / g.n_tiles_n, cc);
      const int gc = (geglu ? (tile_ % g.n_tiles_n) * 64 : (tile_ % g.n_tiles_n) * g.block_n) + sl_ * 32;
      mbar_arrive_expect_tx(&my_res_full[seq_ & 1], static_cast<uint32_t>(g.slab_bytes));
      // the residual lands in the slab's own output buffer: threads add to it in place
      tma_load_5d(my_out + (seq_ & 1) * kSlabBytes, &g.map_res, &my_res_full[seq_ & 1], gc, cc[0], cc[1], cc[2], cc[3]);
    };
    if (leader && g.residual && blockIdx.x < g.num_tiles && grp < nslabs) issue_res(blockIdx.x, grp, 0);
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int nt = tile % g.n_tiles_n;
      int c[4];
      tile_coords(g, tile / g.n_tiles_n, c);
      bool valid = j[3] < g.box[3];
      long long rvrow = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = c[i] + j[i];
        valid = valid && (x < g.out_dims[i]);
        rvrow += static_cast<long long>(x) * g.rv_stride[i];
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc) * 256u + (static_cast<uint32_t>(q * 32) << 16);
      for (int sl = grp; sl < nslabs; sl += 2, ++seq) {
        const int gcol = (geglu ? nt * 64 : nt * g.block_n) + sl * 32;  // first global output column of the slab
        uint8_t* obuf = my_out + (seq & 1) * kSlabBytes;
        const int ucol = sl * 32 + usub * 16;  // tile-relative column of this thread's 16-column unit
        const int ocol = gcol + usub * 16;
        uint32_t rv[16], rg[16];
        tmem_ld16(taddr + ucol, rv);
        if (geglu) tmem_ld16(taddr + 64 + ucol, rg);
        tmem_ld_wait();
        float v[2][8];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] = __uint_as_float(rv[i]);
        if (geglu) {
          // value * gelu(gate); biases of the interleaved weight: value at tcol, gate at tcol + 64
          const int tcol = nt * 128 + ucol;
          float bv[2][8], bg[2][8];
          if (g.bias) {
            load8(g.bias + tcol, bv[0]);
            load8(g.bias + tcol + 8, bv[1]);
            load8(g.bias + tcol + 64, bg[0]);
            load8(g.bias + tcol + 72, bg[1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) bv[i >> 3][i & 7] = bg[i >> 3][i & 7] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i >> 3][i & 7] = (v[i >> 3][i & 7] + bv[i >> 3][i & 7]) *
                               gelu_erf_f(__uint_as_float(rg[i]) + bg[i >> 3][i & 7]);
        } else {
          if (g.bias) {
            float b[2][8];
            load8(g.bias + ocol, b[0]);
            load8(g.bias + ocol + 8, b[1]);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] += b[i >> 3][i & 7];
          }
          if (g.act == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] = silu_f(round_h(v[i >> 3][i & 7]));
          }
        }
        if (g.rowvec && valid) {
          float t[2][8];
          load8(g.rowvec + rvrow * g.ld_rowvec + ocol, t[0]);
          load8(g.rowvec + rvrow * g.ld_rowvec + ocol + 8, t[1]);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] = round_h(v[i >> 3][i & 7]) + t[i >> 3][i & 7];
        }
        if (g.gamma != 1.0f) {
          const float gm = g.gamma;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] = round_h(v[i >> 3][i & 7]) * gm;
        }
        if (g.residual) {
          mbar_wait(&my_res_full[seq & 1], (seq >> 1) & 1u);
          float t[2][8];
          load8(reinterpret_cast<const __half*>(obuf + off0), t[0]);
          load8(reinterpret_cast<const __half*>(obuf + off1), t[1]);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i >> 3][i & 7] = round_h(v[i >> 3][i & 7]) + t[i >> 3][i & 7];
        }
        store8(reinterpret_cast<__half*>(obuf + off0), v[0]);
        store8(reinterpret_cast<__half*>(obuf + off1), v[1]);
        fence_proxy_async_smem();
        // the other output buffer must have been drained (store seq-1) before anyone writes it in the next slab
        if (leader) bulk_wait_read_all();
        named_bar_sync(bar_id, 256);
        if (leader) {
          tma_store_5d(&g.map_out, obuf, gcol, c[0], c[1], c[2], c[3]);
          bulk_commit_group();
          if (g.residual) {  // prefetch the residual of this group's next slab (same tile, or the next tile's first)
            if (sl + 2 < nslabs) {
              issue_res(tile, sl + 2, seq + 1);
            } else if (tile + static_cast<int>(gridDim.x) < g.num_tiles) {
              issue_res(tile + static_cast<int>(gridDim.x), grp, seq + 1);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (leader) bulk_wait_all();  // the staging slabs must outlive the last store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
  LVD_REQUIRE(d->act >= 0 && d->act <= 2, "lvd_gemm: act=%d", d->act);
  LVD_REQUIRE(d->act != 2 || d->block_n == 128, "lvd_gemm: GEGLU needs block_n == 128");
  LVD_REQUIRE(!d->b_mn_major || d->block_n % 64 == 0, "lvd_gemm: MN-major B needs block_n %% 64 == 0");
  LVD_REQUIRE(d->a_box[0] == 64, "lvd_gemm: a_box[0] must be 64");
  long long rows = 1;
  for (int i = 1; i < 5; ++i) rows *= d->a_box[i];
  LVD_REQUIRE(rows >= 1 && rows <= 128, "lvd_gemm: tile has %lld rows (max 128)", rows);
  LVD_REQUIRE((d->ld_out % 8) == 0 && (!d->residual || d->ld_res % 8 == 0) && (!d->rowvec || d->ld_rowvec % 8 == 0),
              "lvd_gemm: leading dimensions must be multiples of 8");

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  int rc = encode_tmap_f16_sw128(&g.map_a, d->a, 5, d->a_dims, d->a_strides, d->a_box);
  if (rc) return rc;
  const int k_tap_stride = (d->b_k_tap_stride == 0 && d->b_n_tap_stride == 0) ? d->k_per_tap : d->b_k_tap_stride;
  const long long k_total = static_cast<long long>(d->n_taps - 1) * k_tap_stride + d->k_per_tap;
  const long long n_total = static_cast<long long>(d->n_taps - 1) * d->b_n_tap_stride + d->n;
  if (!d->b_mn_major) {
    int64_t dims[2] = {k_total, n_total};
    int64_t str[2] = {1, d->ldb};
    int box[2] = {64, d->block_n};
    rc = encode_tmap_f16_sw128(&g.map_b, d->b, 2, dims, str, box);
  } else {
    int64_t dims[2] = {n_total, k_total};
    int64_t str[2] = {1, d->ldb};
    int box[2] = {64, 64};
    rc = encode_tmap_f16_sw128(&g.map_b, d->b, 2, dims, str, box);
  }
  if (rc) return rc;
  long long m_tiles = 1;
  for (int i = 0; i < 4; ++i) {
    g.box[i] = d->a_box[i + 1];
    g.out_dims[i] = d->out_dims[i];
    LVD_REQUIRE(d->out_dims[i] >= 1, "lvd_gemm: out_dims[%d]=%d", i, d->out_dims[i]);
    g.ntile[i] = (d->out_dims[i] + g.box[i] - 1) / g.box[i];
    m_tiles *= g.ntile[i];
    g.rv_stride[i] = d->rv_stride[i];
    g.out_row_stride[i] = d->out_row_stride[i];
  }
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
  LVD_REQUIRE(m_tiles * g.n_tiles_n < (1ll << 30), "lvd_gemm: too many tiles");
  g.num_tiles = static_cast<int>(m_tiles * g.n_tiles_n);
  g.a_bytes = static_cast<int>(rows) * 128;
  g.b_bytes = d->block_n * 128;
  g.slab_bytes = static_cast<int>(rows) * 64;
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
  const int n_slabs_smem = 4;
  int stages = (kSmemLimit - 1024 - 256 - n_slabs_smem * kSlabBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  g.stages = stages;
  g.bias = static_cast<const __half*>(d->bias);
  g.rowvec = static_cast<const __half*>(d->rowvec);
  g.residual = static_cast<const __half*>(d->residual);
  g.out = static_cast<__half*>(d->out);
  g.ld_rowvec = d->ld_rowvec;
  g.ld_res = d->ld_res;
  g.ld_out = d->ld_out;
  g.gamma = d->gamma;
  g.idesc = make_idesc_f16(128, static_cast<uint32_t>(d->block_n), 0, d->b_mn_major ? 1u : 0u);

  const int smem_bytes = 1024 + stages * stage_bytes + n_slabs_smem * kSlabBytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    LVD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  int grid = g.num_tiles < num_sms() ? g.num_tiles : num_sms();
  gemm_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(g);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  LVD_CHECK_CUDA(cudaGetLastError());
  return LVD_OK;
}

}  // namespace lvd

extern "C" int lvd_gemm(const lvd_gemm_desc* d, void* stream) {
  return lvd::launch_gemm(d, static_cast<cudaStream_t>(stream));
}
extern "C" int lvd_version(void) { return 1; }
extern "C" int64_t lvd_launch_count(void) { return lvd::g_launches.load(); }
