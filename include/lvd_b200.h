/* lvd_b200 -- C-ABI of the B200-native (sm_100a) denoising hot path of LLM-grounded Video Diffusion.
 *
 * Every entry point takes plain pointers / sizes / a CUDA stream handle (void* = cudaStream_t); no C++ or
 * torch types cross this boundary.  Device pointers are caller-owned (`tensor.data_ptr()` on the Python side);
 * the library never frees or retains them.  All calls are asynchronous on `stream`.
 *
 * Return convention: 0 = ok; <0 = error (-1 invalid argument, -2 unsupported shape, -3 CUDA error) with a
 * human-readable message in lvd_last_error().  There is no CPU fallback anywhere behind this interface.
 *
 * "Replaces" notes cite the reference implementation (TonyLianLong/LLM-groundedVideoDiffusion @22d90d2) or
 * the un-vendored diffusers==0.27.2 module the reference imports at that line.
 *
 * Activation layout everywhere: fp16, channels-last.  A (B*F, C, h, w) reference tensor is stored as
 * rows = ((b*F + f)*h + y)*w + x, C contiguous, with an explicit row stride `ld` (elements) so that channel
 * concatenation (skip connections) is a zero-copy column slice.
 */
#ifndef LVD_B200_H_
#define LVD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVD_MAX_TAPS 27

const char* lvd_last_error(void);
int lvd_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
int64_t lvd_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core contraction with a fused epilogue (tcgen05.mma + TMA, fp16 in, fp32 accumulate in TMEM).
 *
 *   acc[row, n] = sum_taps sum_c  A[coord(row) + tap_off[tap]][c] * W[n][tap*k_per_tap + c]
 *   v = acc (+ bias[n]) -> fp16; (SiLU | GEGLU) -> fp16; (+ rowvec[rv(row), n]) -> fp16; (* gamma) -> fp16;
 *       (+ residual[row, n]) -> fp16;   out[row, n] = v
 *
 * One kernel covers: nn.Linear (attention_processor.py:137-152, attention.py:355-376, transformer_2d.py:162,230,
 * transformer_temporal.py:87,107), nn.Conv2d 3x3 / 1x1 / stride-2 of diffusers ResnetBlock2D / Downsample2D /
 * Upsample2D (used at unet_3d_blocks.py:330,379-390,610-613) and nn.Conv3d (3,1,1) of diffusers
 * TemporalConvLayer (unet_3d_blocks.py:343) -- convolutions as implicit GEMM: each filter tap is a shifted TMA box
 * of the channels-last activation, zero padding comes from TMA out-of-bounds fill.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lvd_gemm_desc {
  /* A operand: logical rank-5 fp16 tensor, dim 0 = channels (contiguous); unused trailing dims have extent 1 */
  const void* a;
  int64_t a_dims[5];
  int64_t a_strides[5]; /* elements; a_strides[0] == 1, others multiples of 8 */
  int32_t a_box[5];     /* a_box[0] == 64; prod(a_box[1..4]) <= 128 = rows of one M tile */
  int32_t n_taps;       /* 1 linear / 1x1; 9 for 3x3; 3 temporal */
  int32_t k_per_tap;    /* channels contracted per tap; multiple of 64 */
  int32_t tap_off[LVD_MAX_TAPS][5];
  /* B operand (weights), fp16 */
  const void* b;
  int32_t b_mn_major; /* 0: B = [n][K_total], K contiguous.  1: B = [K_total][n], n contiguous (dgrad) */
  int32_t block_n;    /* 64 | 128 | 160 | 256; n % block_n == 0 (GEGLU: 128) */
  int64_t ldb;        /* elements */
  int32_t n;          /* GEMM N (for GEGLU: 2x the output width, value/gate interleaved per 64) */
  int32_t act;        /* 0 none, 1 SiLU, 2 GEGLU */
  /* output geometry: tile axis i (=A dims 1..4) covers out_dims[i] positions in steps of a_box[i+1] */
  int32_t out_dims[4];
  int32_t rv_stride[4];      /* rowvec row = sum coord_i * rv_stride[i] */
  int64_t out_row_stride[4]; /* output row = sum coord_i * out_row_stride[i] */
  const void* bias;          /* fp16 [n] or NULL */
  const void* rowvec;        /* fp16 [*, ld_rowvec] or NULL (time-embedding projection per image) */
  const void* residual;      /* fp16 [rows, ld_res] or NULL */
  void* out;                 /* fp16 [rows, ld_out] */
  int64_t ld_rowvec, ld_res, ld_out;
  float gamma; /* scalar multiplier (GLIGEN tanh gate), 1.0f = off */
  int32_t reserved;
} lvd_gemm_desc;

int lvd_gemm(const lvd_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LVD_B200_H_ */
