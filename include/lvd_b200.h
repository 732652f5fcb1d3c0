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
/* tuning aid: with LVD_GEMM_DEBUG=64 in the environment CTA 0 of every lvd_gemm launch records a timeline
 * (event << 48 | %globaltimer ns); this copies out and resets it.  Returns the number of events, < 0 on error. */
int lvd_debug_trace_read(unsigned long long* host_buf, int max_events);

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
  int32_t block_n;    /* 64 | 128 | 160 | 256; n % block_n == 0 (GEGLU: 128 | 256) */
  int64_t ldb;        /* elements */
  int32_t n;          /* GEMM N (for GEGLU: 2x the output width; every block_n columns = [block_n/2 value | block_n/2 gate]) */
  int32_t act;        /* 0 none, 1 SiLU, 2 GEGLU, 3 GELU (erf; CLIP text encoder MLP) */
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
  /* B coordinate of tap t: k0 = t * b_k_tap_stride, n0 += t * b_n_tap_stride.  Forward conv: (k_per_tap, 0).
   * Conv dgrad reads the SAME [Cout][taps*Cin] weight as an MN-major B: (0, Cin), with negated tap offsets on A. */
  int32_t b_k_tap_stride;
  int32_t b_n_tap_stride;
  /* 0: automatic; 1: one CTA per 128-row tile; 2: CTA pairs (2-CTA clusters, tcgen05 cta_group::2, 256-row tiles --
   * each SM stages half of every B tile; needs >= 2 row tiles, MN-major B: block_n % 128 == 0) */
  int32_t cta_pair;
  /* GroupNorm statistics of the output, produced by the epilogue for the GroupNorm that consumes it (diffusers
   * ResnetBlock2D.norm2 after conv1, TemporalConvLayer's norms after the resnet / the temporal convs:
   * unet_3d_blocks.py:330-349) -- the normalisation then needs no statistics pass over the tensor.  NULL = off.
   * gn_partial: fp32 [samples][gn_tiles_per_sample][4][gn_subslots][gn_groups][2], ZERO on entry; every epilogue warp
   * adds {sum, sum of squares} of its 32 rows per channel group (one writer per float: bit-reproducible).  A row tile
   * must lie inside one sample (row tiles of a sample are consecutive in tile order); gn_group_width = n / gn_groups;
   * gn_subslots >= (gn_group_width + 30) / 32 + 1.  lvd_groupnorm_apply folds the partials into the statistics, clears
   * them and normalises. */
  void* gn_partial;
  int32_t gn_group_width, gn_groups, gn_tiles_per_sample, gn_subslots;
  /* LayerNorm folded into the linear that consumes it (models/attention.py:126-170: norm1 -> attn1.to_q/k/v,
   * norm2 -> attn2.to_q, norm3 -> ff.net.0): A is the RAW LayerNorm input, B the weight with the LayerNorm scale
   * folded in (B'[n][k] = gamma_k W[n][k], fp16), and the epilogue finishes
   *     out[r][n] = rstd_r * (acc[r][n] - mean_r * ln_colsum[n]) + ln_bias[n]
   * ln_colsum[n] = sum_k B'[n][k] (fp32, of the rounded B'), ln_bias[n] = sum_k beta_k W[n][k] + bias[n] (fp32);
   * ln_stats = fp32 {mean, rstd} per row, row = sum coord_i * ln_row_stride[i] (lvd_ln_rowstats).  Needs act none or
   * GEGLU, no bias / residual / rowvec / gamma.  NULL = off. */
  const void* ln_stats;
  const void* ln_colsum;
  const void* ln_bias;
  int32_t ln_row_stride[4];
  /* Row statistics of the OUTPUT for the LayerNorm that consumes it: lnstat_partial = fp32 {sum, sum of squares}
   * [n / 32][lnstat_rows] of the 32 fp16 outputs of a row in each 32-column slab (every element written: no initial
   * state); row = sum coord_i * lnstat_row_stride[i].  NULL = off. */
  void* lnstat_partial;
  int64_t lnstat_rows;
  int32_t lnstat_row_stride[4];
  /* Optional fp32 copy of `bias` (the same n values, exactly converted): the specialised epilogues then add it without
   * converting 32 halves per row and slab -- these epilogues bound the short-K linears (K = 320 / 640).  NULL = read
   * the fp16 bias. */
  const void* bias_f32;
  /* 1: weight-resident walk (one tap, K-major B whose block_n x K tile fits next to >= 3 A stages; silently the
   * streaming walk otherwise): the persistent walk keeps every CTA (pair) on one n-tile, whose B tile is staged once
   * for the whole launch, and only A streams -- for the short-K linears, which are bound by operand delivery from L2.
   * 0: both operands stream through the ring. */
  int32_t weight_resident;
} lvd_gemm_desc;

int lvd_gemm(const lvd_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention: softmax(Q K^T * scale) V, 64-wide heads, tcgen05 QK^T / PV tiles staged by TMA.
 * Replaces F.scaled_dot_product_attention (models/attention_processor.py:406-408) and the materialising slow path
 * baddbmm -> softmax -> bmm (models/attention_processor.py:222-258, :515-536, :553-586).
 *
 * mode 0 (spatial self / cross): Q rows [z][s_q] at q + column q_col0 + head*64 (row stride ld_q); K and V rows
 *   [z / kv_z_div][s_k] of the same buffer `kv` at columns k_col0 / v_col0 (+ head*64).  kv_z_div = F lets all
 *   frames of a video share one text K/V (the reference repeat_interleaves encoder_hidden_states instead,
 *   unet_3d_condition.py:721-723).
 * mode 1 (temporal, gathered): rows are ((z*frames + f)*hw + p); the sequence of pixel p is its `frames` rows
 *   (transformer_temporal.py:142-156 permutes instead).  Q, K, V in one buffer.
 * mode 2 (temporal, packed): rows are ((z*hw + p)*frames + f), i.e. every `frames` consecutive rows are one
 *   sequence (the QKV projection writes this order through lvd_gemm's output row mapping, the output projection
 *   reads it back); frames in {8, 16, 24, 32}.  This is the fast path: one warp per (sequence, head), every operand
 *   straight from registers into mma.sync (attention_temporal.cu).
 * Kernel selection inside mode 0: s_k <= 128 with no map / lse output (the 77 text tokens, the 45-pixel mid block, the
 *   CLIP text encoder's causal self-attention) runs the same warp-level scheme; everything else (spatial
 *   self-attention over thousands of keys, map outputs, lse) runs the tcgen05 tile kernel (attention.cu).
 * lse (optional, fp32 [nz, heads, rows]) receives log2-domain logsumexp of every query row for the backward kernels.
 * Output row order == Q row order; column o_col0 + head*64.
 * Optional (s_k <= 128 only): p_full [nz, heads, s_q, s_k] fp16 = the reference's saved attention map
 *   (attention_processor.py:553-586); p_tok [n_tok, nz, heads, s_q] = selected text-token columns only
 *   (what utils/guidance.py:280-290 reads).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lvd_attn_desc {
  const void* q;
  const void* kv;
  void* out;
  void* p_full;
  void* p_tok;
  float* lse;
  int64_t ld_q, ld_kv, ld_o;
  int64_t q_width, kv_width; /* logical column counts of the q / kv buffers (TMA bounds) */
  int64_t s_q, s_k;          /* mode 0 */
  int64_t hw;                /* mode 1 */
  int32_t frames;            /* mode 1 */
  int32_t mode;
  int32_t q_col0, k_col0, v_col0, o_col0;
  int32_t heads, nz, kv_z_div;
  int32_t n_tok;
  int32_t tok_cols[16];
  float scale;
  int32_t causal; /* mode 0, s_k <= 128, no map outputs: key j is visible to query i iff j <= i (CLIP text encoder) */
  const int32_t* tok_cols_dev; /* optional device copy of tok_cols[n_tok], read by the kernel at run time (so a captured
                                  CUDA graph serves layouts with other token positions); NULL = use tok_cols */
} lvd_attn_desc;

int lvd_attention(const lvd_attn_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Normalisation (fp32 statistics, fp16 in/out, channels-last).
 * lvd_groupnorm: x viewed [samples][rows_per_sample][channels]; statistics per (sample, group) over
 *   rows_per_sample x channels/groups.  4-D GroupNorm of ResnetBlock2D / Transformer2DModel
 *   (transformer_2d.py:155-160,311) -> sample = one frame; 5-D GroupNorm of diffusers TemporalConvLayer and
 *   TransformerTemporalModel.norm (transformer_temporal.py:84-86,150) -> sample = one video (rows span frames).
 *   silu != 0 fuses the following SiLU.  workspace: >= samples*1024*groups*8 bytes is always enough.
 * lvd_layernorm: nn.LayerNorm over channels (models/attention.py:126,145,163).
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_groupnorm(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t samples, int64_t rows_per_sample,
                  int32_t channels, int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu,
                  void* stats_out /* optional fp32 [samples, groups, 2] = {mean, rstd}, kept for lvd_groupnorm_bwd */,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* GroupNorm of a tensor written by an lvd_gemm with gn_partial set: the statistics pass is replaced by a fold of the
 * `slots` = tiles_per_sample * 4 * subslots partial pairs per (sample, group) that GEMM's epilogue left in gn_partial
 * (fixed order; the partials are reset to zero for the next producer), then the normalisation pass of lvd_groupnorm
 * runs.  workspace as for lvd_groupnorm; stats_out (required) receives {mean, rstd}. */
int lvd_groupnorm_apply(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t samples, int64_t rows_per_sample,
                        int32_t channels, int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu,
                        void* gn_partial, int32_t slots, void* workspace, void* stats_out, void* stream);
int lvd_layernorm(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels,
                  const void* gamma, const void* beta, float eps, void* stream);
/* Folds lvd_gemm's lnstat_partial ([slabs][rows] {sum, sum of squares}, slabs = channels / 32) into stats[row] =
 * {mean, rstd} for a LayerNorm folded into the next lvd_gemm (ln_stats). */
int lvd_ln_rowstats(const void* lnstat_partial, int32_t slabs, int64_t rows, int32_t channels, float eps,
                    void* stats_f32, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Thin ends of the UNet and small element-wise pieces.
 * lvd_conv_in : sample (B, cin, F, H, W) NCFHW fp16 -> rows, 3x3 conv (unet_3d_condition.py:297,726-729);
 *               weight in the reference's OIHW layout.
 * lvd_conv_out: rows -> (B, cout, F, H, W) NCFHW (unet_3d_condition.py:424,847-854); weight [cout][ky][kx][cin].
 * lvd_rows_to_ncfhw: columns [0, cout) of a [B*F*hw, ld] row matrix -> (B, cout, F, H, W); the layout change of
 *               conv_out when the conv itself runs on lvd_gemm with cout zero-padded to one 64-wide tile.
 * lvd_upsample2x: nearest 2x (diffusers Upsample2D's F.interpolate, unet_3d_blocks.py:610-613).
 * lvd_timestep_embed: diffusers Timesteps(dim, flip_sin_to_cos=True, shift 0) (unet_3d_condition.py:304,712).
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_conv_in(const void* sample, const void* weight, const void* bias, void* out, int64_t ld_out, int32_t batch,
                int32_t cin, int32_t frames, int32_t height, int32_t width, int32_t cout, void* stream);
/* lvd_im2col_small: sample (B, cin <= 7, F, H, W) -> rows [B*F*H*W, 64] = the 3x3 / pad 1 patches, column ci*9 + ky*3 + kx
 *               (zero padded): the wide path of conv_in, which then is one K = 64 lvd_gemm against the OIHW weight
 *               flattened and zero-padded to 64 columns. */
int lvd_im2col_small(const void* sample, void* out, int32_t batch, int32_t cin, int32_t frames, int32_t height,
                     int32_t width, void* stream);
int lvd_conv_out(const void* x, int64_t ld_x, const void* weight_krsc, const void* bias, void* out, int32_t batch,
                 int32_t cin, int32_t frames, int32_t height, int32_t width, int32_t cout, void* stream);
int lvd_rows_to_ncfhw(const void* rows, int64_t ld, void* out, int32_t batch, int32_t cout, int32_t frames, int64_t hw,
                      void* stream);
int lvd_upsample2x(const void* x, int64_t ld_x, void* y, int64_t ld_y, int32_t n_img, int32_t height, int32_t width,
                   int32_t channels, void* stream);
int lvd_timestep_embed(const void* timesteps_f32, void* out, int32_t batch, int32_t dim, void* stream);
int lvd_silu(const void* x, void* y, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * VAE decode (TextToVideoSDPipeline.decode_latents,
 * controllable_pipeline_text_to_video_synth.py:374-400, over diffusers 0.27.2 AutoencoderKL -- not vendored by the
 * reference).  The decoder's convolutions / GroupNorms / up-sampling run on the entry points above; two small
 * pieces are specific to it:
 * lvd_softmax_rows   : in-place fp32 row softmax of scale * x (fp16 storage) -- the single-head, 512-wide attention
 *                      of the decoder's mid block runs its two products as lvd_gemm launches.
 * lvd_pointwise_small: 1x1 convolution between <= 8 channels on planar (N, C, HW) fp16 (post_quant_conv).
 * lvd_frames_postprocess: decoded frames (N, C, HW) fp16 in [-1, 1] -> (N, HW, C) fp32 in [0, 1] -- tensor2vid /
 *                      VaeImageProcessor.postprocess("np") (controllable_pipeline_text_to_video_synth.py:66-88).
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_softmax_rows(void* x, int64_t ld, int64_t rows, int32_t cols, float scale, void* stream);
int lvd_frames_postprocess(const void* x, void* out, int64_t n, int32_t channels, int64_t hw, void* stream);
/* lvd_embed_tokens: out[(b*seq + s), :] = token_embedding[ids[b, s], :] + position_embedding[s, :]  (fp16; CLIPTextEmbeddings,
 * the caller-supplied text encoder of _encode_prompt, controllable_pipeline_text_to_video_synth.py:197-372). */
int lvd_embed_tokens(const int64_t* ids, const void* token_embedding, const void* position_embedding, void* out,
                     int64_t n_tokens, int32_t seq, int32_t dim, int32_t vocab, void* stream);
int lvd_pointwise_small(const void* x, const void* weight, const void* bias, void* out, int32_t n, int32_t cin,
                        int32_t cout, int64_t hw, void* stream);
int lvd_copy_rows(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-step latent arithmetic (fp16 tensors of n elements, evaluated with the reference's fp16 op order).
 * lvd_cfg_dpm_step: classifier-free-guidance combine (controllable_pipeline_text_to_video_synth.py:926-930) fused
 *   with diffusers 0.27.2 DPMSolverMultistepScheduler.step (dpmsolver++, midpoint, epsilon prediction; used at
 *   :941-943).  eps_cond == NULL skips the CFG combine.  Host passes the scalar coefficients:
 *     x0 = (x - sigma_t*eps)/alpha_t ;  x_out = c_x*x - c_d0*x0 - c_d1*D1 ,  D1 = inv_r0*(x0 - x0_prev) (order 2)
 * lvd_axpy_h: latents - scale*grad (models/pipelines.py:124-132).
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_cfg_dpm_step(const void* eps_uncond, const void* eps_cond, const void* x, const void* x0_prev, void* x_out,
                     void* x0_out, int64_t n, float guidance_scale, float sigma_t, float alpha_t, float c_x,
                     float c_d0, float c_d1, float inv_r0, int32_t order, void* stream);
int lvd_axpy_h(const void* x, const void* g, void* out, int64_t n, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Box energy of the LVD backward guidance on one saved cross-attention map, and dE/dP.
 * Replaces utils/guidance.py:160-526 (add_ca_loss_per_attn_map_to_loss: top-k "max-based" / ratio / CE losses, attention
 * sync, BoxDiff corner constraint, centre-of-mass position/velocity terms) and the normalisation of compute_ca_lossv3
 * (:529-574).
 *   slot  = one (object, text token) pair; boxes[slot][frame] = (x0,y0,x1,y1) integer rectangle at this map's
 *           resolution (host applies utils/utils.py:82-103 scale_proportion -- Python banker's rounding);
 *   slot_weight[slot] = 1 / len(object_positions[object])  (guidance.py:524)
 *   out_scale = loss_scale / (n_objects * n_keys)           (guidance.py:569-572, models/pipelines.py:101-112)
 * layout 0: p = token-column maps [n_slots][frames][heads][H*W] (what lvd_attention's p_tok emits);
 * layout 1: p = the reference's saved map [frames][heads][H*W][n_text], slot_col[slot] = text token index.
 * *loss is atomically incremented (zero it first).  dp (same layout as p, may be NULL) receives
 * grad_scale * dE/dP for the slot's column (fp16; grad_scale guards against underflow).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct lvd_energy_desc {
  const void* p;
  const int32_t* slot_col;
  const int32_t* boxes;
  const float* slot_weight;
  float* loss;
  void* dp;
  int32_t layout, n_text, n_slots, frames, heads, H, W;
  float fg_top_p, bg_top_p, fg_weight, bg_weight, com_loss_scale, out_scale, grad_scale;
  /* loss_kind: 0 max-based top-k (guidance.py:341-353), 1 ratio-based (:312-323), 2 CE / NLL (:363-396);
   * k_min = int(upsample_scale^2) (:328-337); eps (:179, 1e-2) enters kinds 1 and 2;
   * attn_sync_weight (:401-426); boxdiff_* (:433-465).  Kind 2 excludes the sync / BoxDiff / COM terms. */
  int32_t loss_kind, k_min, boxdiff_normed, boxdiff_L;
  float eps, attn_sync_weight, boxdiff_loss_scale;
  /* several prompts in one map (batched guidance pass; an extension, the reference guides one prompt at a time):
   * p's frame axis has p_frames (0 = frames) entries, slot s covers frames [slot_frame0[s], slot_frame0[s] + frames)
   * and adds its energy to loss[slot_loss[s]].  NULL tables = one prompt. */
  int32_t p_frames;
  const int32_t* slot_frame0;
  const int32_t* slot_loss;
} lvd_energy_desc;

int lvd_box_energy(const lvd_energy_desc* d, void* stream);

/* Optional pre-processing of a saved map ahead of lvd_box_energy, and the adjoints back to dE/dP
 * (utils/guidance.py:209-226,281-311; utils/attn.py:91-157 GaussianSmoothing).  Maps are fp16
 * [planes = frames*heads][hw][n_text]; gradients travel as fp32 between adjoint stages, the stage that produces dE/dP
 * itself writes fp16 (`*_is_half` = 1).
 *   lvd_attn_smooth      : depthwise 3x3 blur (weights w9_host[9], row-major (d_pixel, d_token)) of the reflect-padded
 *                          (hw x n_text) plane -- smooth_attn=True
 *   lvd_attn_renorm      : out[rows][num_tokens-2] = softmax(p[rows][1:num_tokens-1] * scale) -- attn_renorm=True
 *   lvd_attn_columns     : cols[n_slots][planes][h_out*w_out] = column slot_col[slot] of the map, bilinearly
 *                          interpolated (align_corners = False) when (h_out, w_out) != (h_in, w_in) -- upsample_scale
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_attn_smooth(const void* p, void* out, int64_t planes, int32_t hw, int32_t n_text, const float* w9_host,
                    void* stream);
int lvd_attn_smooth_bwd(const void* d_out_f32, void* d_p, int32_t d_p_is_half, int64_t planes, int32_t hw,
                        int32_t n_text, const float* w9_host, void* stream);
int lvd_attn_renorm(const void* p, void* out, int64_t rows, int32_t n_text, int32_t num_tokens, float scale,
                    void* stream);
int lvd_attn_renorm_bwd(const void* r_out, const void* d_r_f32, void* d_p, int32_t d_p_is_half, int64_t rows,
                        int32_t n_text, int32_t num_tokens, float scale, void* stream);
int lvd_attn_columns(const void* map, const int32_t* slot_col_dev, void* cols, int32_t n_slots, int64_t planes,
                     int32_t n_cols, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out, void* stream);
int lvd_attn_columns_bwd(const void* d_cols, const int32_t* slot_col_dev, void* d_map, int32_t d_map_is_half,
                         int32_t n_slots, int64_t planes, int32_t n_cols, int32_t h_in, int32_t w_in, int32_t h_out,
                         int32_t w_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * GLIGEN helpers.
 * lvd_copy_rows_seq: dst[(r/n)*dst_seq + dst_off + r%n] = src[(r/n)*src_seq + src_off + r%n]; builds / splits the
 *   [visual ; grounding-token] sequence of GatedSelfAttentionDense.forward (models/attention.py:50-57).
 * lvd_gligen_prep: FourierEmbedder + masked null blending + concat of PositionNet.forward
 *   (models/unet_3d_condition.py:47-64,119-150); out [n_obj, dim + 64] feeds position_net.linears via lvd_gemm.
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_copy_rows_seq(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t rows, int32_t channels, int32_t n,
                      int32_t src_seq, int32_t src_off, int32_t dst_seq, int32_t dst_off, void* stream);
int lvd_gligen_prep(const void* boxes, const void* masks, const void* positive_embeddings, const void* null_positive,
                    const void* null_position, void* out, int64_t n_obj, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Input-gradient kernels of the guidance path -- replace torch.autograd.grad(loss, [latents])
 * (models/pipelines.py:120).  Only d/d(latents) exists; weight gradients are never formed.  Linear / Conv dgrads run
 * on lvd_gemm (b_mn_major = 1, per-tap weight column offsets); these are the HBM-bound adjoints around them.
 *   lvd_groupnorm_bwd : GroupNorm (+ fused SiLU) adjoint; same sample/row view as lvd_groupnorm; recomputes the
 *                       forward statistics; accumulate != 0 adds into dx (gradient fan-in of a residual branch).
 *                       workspace >= 2 x the forward requirement.
 *   lvd_layernorm_bwd : LayerNorm adjoint (row statistics recomputed).
 *   lvd_geglu / lvd_geglu_bwd : un-fused GEGLU on the value/gate-interleaved projection (the guidance forward keeps
 *                       the pre-activation so that the adjoint needs no recomputation), models/attention.py:355-376.
 *   lvd_add_rows, lvd_upsample2x_bwd, lvd_conv_in_bwd (fp32 d_sample in NCFHW), lvd_attn_rowdot
 *                       (D[row, head] = <O, dO>, the softmax-adjoint row term),
 *   lvd_latent_update : latents - sqrt(1 - alpha_bar_t) * grad (models/pipelines.py:124-132) from the fp32 gradient.
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_groupnorm_bwd(const void* x, int64_t ld_x, const void* dz, int64_t ld_dz, void* dx, int64_t ld_dx,
                      int32_t samples, int64_t rows_per_sample, int32_t channels, int32_t groups, const void* gamma,
                      const void* beta, float eps, int32_t silu, int32_t accumulate,
                      const void* fwd_stats /* optional: stats_out of the forward launch (skips the recomputation) */,
                      void* workspace, int64_t workspace_bytes, void* stream);
int lvd_layernorm_bwd(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int64_t rows,
                      int32_t channels, const void* gamma, float eps, int32_t accumulate, void* stream);
int lvd_geglu(const void* pre, int64_t ld_pre, void* out, int64_t ld_out, int64_t rows, int32_t n_out, void* stream);
int lvd_geglu_bwd(const void* pre, int64_t ld_pre, const void* dout, int64_t ld_dout, void* dpre, int64_t ld_dpre,
                  int64_t rows, int32_t n_out, void* stream);
int lvd_add_rows(const void* a, int64_t ld_a, const void* b, int64_t ld_b, void* out, int64_t ld_out, int64_t rows,
                 int32_t channels, void* stream);
int lvd_upsample2x_bwd(const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int32_t n_img, int32_t height,
                       int32_t width, int32_t channels, void* stream);
int lvd_conv_in_bwd(const void* drows, int64_t ld, const void* weight, void* dsample_f32, int32_t batch, int32_t cin,
                    int32_t frames, int32_t height, int32_t width, int32_t cout, void* stream);
int lvd_latent_update(const void* latents, const void* grad_f32, void* out, void* grad_h_out, int64_t n, float coef,
                      float unscale, void* stream);
int lvd_attn_rowdot(const void* o, int64_t ld_o, const void* d_o, int64_t ld_do, void* out_f32, int64_t rows,
                    int32_t heads, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention adjoints (probabilities recomputed from Q, K; fp32 math; gradients w.r.t. the q|k|v row buffer).
 *   lvd_self_attn_bwd     : spatial self-attention, tcgen05 (two kernels: dQ; dK+dV).  qkv rows [nz*s]; `lse` is
 *                           lvd_attention's log2-domain log-sum-exp [nz, heads, s]; `dvec` = lvd_attn_rowdot(O, dO).
 *   lvd_temporal_attn_bwd : packed temporal attention (mode 2 rows, frames <= 32), one warp per (sequence, head).
 *   lvd_cross_attn_bwd    : cross-attention, dQ only (text K/V carry no gradient to the latents).  dp_tok
 *                           [n_slots, nz, heads, s_q] is lvd_box_energy's dE/dP on the object-token columns
 *                           slot_col_host[n_slots]: the point where the guidance energy enters the backward pass
 *                           (replaces autograd through attention_processor.py:515-536 and utils/guidance.py:346-353).
 * ---------------------------------------------------------------------------------------------------------- */
int lvd_self_attn_bwd(const void* qkv, int64_t ld_qkv, int64_t width_qkv, int32_t q_col0, int32_t k_col0,
                      int32_t v_col0, const void* d_o, int64_t ld_do, const void* lse, const void* dvec, void* d_qkv,
                      int64_t ld_dqkv, int64_t s, int32_t heads, int32_t nz, float scale, void* stream);
/* width_* = logical column count of the buffer (TMA bounds); > 0 selects the tcgen05 single-unit kernel
 * (attn_bwd_small_kernel: probabilities recomputed exactly in-kernel), 0 the scalar fallback for odd frame counts. */
int lvd_temporal_attn_bwd(const void* qkv, int64_t ld_qkv, int64_t width_qkv, int32_t q_col0, int32_t k_col0,
                          int32_t v_col0, const void* d_o, int64_t ld_do, void* d_qkv, int64_t ld_dqkv, int64_t n_seq,
                          int32_t heads, int32_t frames, float scale, void* stream);
/* dp_tok: n_slots in 1..16 -> dE/dP of the listed token columns [n_slots][nz][heads][s_q]; n_slots == -1 -> the dense
 * map [nz][heads][s_q][s_k] (map pre-processing variants of the energy, more than 16 slots); 0 -> none.
 * slot_col_dev (optional): device copy of the column list, read at run time instead of slot_col_host (graph replays). */
int lvd_cross_attn_bwd(const void* q, int64_t ld_q, int64_t width_q, const void* kv, int64_t ld_kv, int64_t width_kv,
                       int32_t k_col0, int32_t v_col0, const void* d_o, int64_t ld_do, const void* dp_tok,
                       const int32_t* slot_col_host,
                       int32_t n_slots, void* dq, int64_t ld_dq, int64_t s_q, int32_t s_k, int32_t heads, int32_t nz,
                       int32_t kv_z_div, float scale, const int32_t* slot_col_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LVD_B200_H_ */
