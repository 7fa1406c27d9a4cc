/*
 * t2v_b200.h — C ABI of libt2v_b200.so: the sm_100a kernels behind the T2V-Turbo (VideoCrafter2)
 * latent-video UNet denoising hot path, the LCM scheduler step and the AutoencoderKL decode.
 *
 * The reference (Ji4chenLi/t2v-turbo) has NO FFI of its own: its "operator API" is the torch.nn
 * call surface.  Each entry point below names the reference call sites whose
 * arithmetic it replaces.  Conventions (all entry points):
 *   - extern "C", POD structs, raw device pointers, sizes/strides in ELEMENTS;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is retained past the call;
 *   - work is enqueued on the caller's stream only; no synchronisation, no allocation
 *     => every call is CUDA-graph capturable;
 *   - return 0 on success; <0 = argument / shape / alignment error (nothing launched);
 *     >0 = cudaError_t / CUresult from the launch.  t2v_last_error() describes the last failure
 *     on the calling thread.  There is no CPU or library fallback.
 *   - activations are channels-last bf16: a frame batch is [B*T, H, W, C] (C contiguous).
 */
#ifndef T2V_B200_H_
#define T2V_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* t2v_stream_t; /* cudaStream_t */

#define T2V_MAX_DIMS 4
#define T2V_MAX_TAPS 9

/* epilogue flags for T2VGemmDesc.flags */
#define T2V_EPI_GEGLU 1u   /* out[m,j] = (acc[m,v_j]+b) * gelu_erf(acc[m,g_j]+b); weights packed by t2v_pack_geglu_rows */
#define T2V_EPI_OUT_F32 2u /* write fp32 instead of bf16 */
#define T2V_EPI_GELU 4u    /* out = gelu_erf(acc) (unused by VC2; kept for FeedForward(glu=False)) */
#define T2V_WS_CLEAN 8u    /* the caller guarantees `workspace` is all-zero on entry (and gets it back zeroed): split-K then
                              skips its zero kernel — the finalize kernel clears the partial sums as it reads them */

int t2v_version(void);
const char* t2v_last_error(void);

/*
 * Implicit GEMM on tcgen05 tensor cores:  out[m, n] = epi( alpha * sum_k A[m, k] * W[n, k] ).
 *
 * A is never materialised: row m is a point (x1..x4) of up to two channels-last activation tensors
 * (concatenated along channels), and the K axis is (tap, channel): k = tap * c_in + c, where each
 * tap shifts the point by tap_off[tap][*] (out-of-range points read as zero = conv padding) and
 * may add a channel offset tap_ch_off[tap] (used by the stride-2 parity view).  One CTA tile is the
 * box[0] x box[1] x box[2] x box[3] block of points (<= 128 rows) times block_n output columns.
 *
 * Replaces, with the matching geometry: nn.Linear (attention.py:70-75,348,370,520,537;
 * openaimodel3d.py:172-178), Conv2d 3x3 / 1x1 (openaimodel3d.py:155-159,179-193,65-72,104-111,666-670),
 * Conv3d (3,1,1) (openaimodel3d.py:274-296), Conv1d k=1 (attention.py:422-424), and the VAE
 * decoder convs (ae_modules.py:146-203,108-122,560-600).  Epilogue fusions: bias (per row group,
 * which carries the ResBlock timestep-embedding add openaimodel3d.py:237-246), residual add
 * (attention.py:300-311,389,513; openaimodel3d.py:247,309), GEGLU (attention.py:516-523).
 */
typedef struct T2VGemmDesc {
  /* A operand */
  const void* a[2];                     /* bf16 activation sources; a[1] may be NULL */
  int32_t a_ch[2];                      /* channels consumed per tap from each source (multiples of 64) */
  int32_t a_ch_total[2];                /* extent of the channel axis of each source tensor (>= a_ch + max tap_ch_off) */
  int64_t a_size[T2V_MAX_DIMS];         /* extents of x1..x4 in the INPUT tensors (x1 fastest); unused = 1 */
  int64_t a_stride[2][T2V_MAX_DIMS];    /* element strides of x1..x4 per source (channel stride = 1) */
  int32_t box[T2V_MAX_DIMS];            /* tile extents; product in [8,128] and a multiple of 8 */
  /* K loop */
  int32_t n_taps;                       /* 1..T2V_MAX_TAPS */
  int32_t tap_off[T2V_MAX_TAPS][T2V_MAX_DIMS];
  int32_t tap_ch_off[T2V_MAX_TAPS];
  /* B operand: weights [b_batches][b_rows][K] bf16, K = n_taps * (a_ch[0] + a_ch[1]) contiguous */
  const void* b;
  int64_t b_rows;                       /* N of the GEMM (rows of W); any N >= 1 */
  int64_t b_batches;                    /* 1 if shared */
  int64_t b_batch_stride;               /* elements between batches */
  int64_t b_row_stride;                 /* elements between rows of W; 0 = K (dense) */
  int32_t b_batch_dim;                  /* index (0..3) of the A dim whose coordinate selects the batch; -1 = none */
  /* output: row (x1..x4) at sum_j x_j * o_stride[j], columns contiguous */
  void* out;
  int64_t o_size[T2V_MAX_DIMS];         /* extents of the OUTPUT point grid (rows outside are dropped) */
  int64_t o_stride[T2V_MAX_DIMS];
  int32_t n_out;                        /* columns written: b_rows, or b_rows/2 with T2V_EPI_GEGLU */
  /* epilogue */
  const float* bias;                    /* fp32 [bias_rows][b_rows] or NULL */
  int64_t bias_row_stride;
  int32_t bias_dim;                     /* A dim whose coordinate / bias_div selects the bias row; -1 = row 0 */
  int32_t bias_div;
  const void* residual;                 /* bf16, indexed like out with r_stride; NULL = none */
  int64_t r_stride[T2V_MAX_DIMS];
  float alpha;
  uint32_t flags;
  int32_t block_n;                      /* 0 = choose; else one of 32,64,128,160,256 */
  /* split-K for small-M layers: K is cut into split_k slices accumulated in fp32 into `workspace`
   * ([points][b_rows] floats, zeroed by the call) and a finalize kernel applies the epilogue.
   * 0 = automatic (only when a large-enough workspace is given and the tile grid cannot fill the SMs),
   * 1 = off.  Not available with GEGLU or batched B. */
  int32_t split_k;
  int32_t tune;                         /* 0 = defaults. bits 0-7: pipeline stages to use; bit 11: disable CTA pairs (cta_group::2); bits 12-15: timing experiments (wrong results) */
  void* workspace;
  int64_t workspace_bytes;
  /* LayerNorm folded into the GEMM that consumes it (attention.py:279-281 feeding to_q/to_k/to_v and GEGLU.proj):
   * with W' = W * gamma (per input channel), col_sum[n] = sum_k W'[n,k], bias' = W beta + bias, and per input row
   * row_stats[m] = (rstd_m, -rstd_m * mean_m) from t2v_layernorm_stats,
   *   LN(x) W^T + bias  ==  rstd_m * (x W'^T)[m,n] + (-rstd_m mean_m) * col_sum[n] + bias'[n],
   * so the normalised activation is never written.  Plain [M,K] x [N,K] GEMMs only (one tap, a_size[1..3] = 1),
   * alpha = 1, bf16 output, bias (if any) a single [N] vector.  NULL = off. */
  const float* row_stats;               /* fp32 [M][2] */
  const float* col_sum;                 /* fp32 [N] (GEGLU: packed like the bias) */
  /* The statistics can also be accumulated by the GEMM that PRODUCES the LayerNorm input (the out-projection +
   * residual / proj_in GEMMs of a transformer block): with row_accum set, the epilogue adds (sum, sum of squares)
   * of every output row into row_accum[m] (fp32 [M][2], zeroed by the caller; plain [M,K] x [N,K], N % 32 == 0).
   * The consumer then passes that buffer as row_stats with ln_raw = 1, ln_channels = N_producer, ln_eps:
   * no LayerNorm kernel runs at all.  row_stats and row_accum are mutually exclusive in one call. */
  int32_t ln_raw; int32_t ln_channels; float ln_eps;
  float* row_accum;
  /* GroupNorm statistics from the producing GEMM (basics.py:78-89 after every conv / proj_out): with col_accum set
   * the epilogue adds, for every output point, its bf16-rounded channel values and their squares into
   * col_accum[sample][channel][2] (fp32, zeroed by the caller), sample = sum_j coord[j] * cs_mult[j] over the output
   * point grid (o_size); t2v_groupnorm then takes these per-channel sums (T2VGroupNormDesc.chan_sums) instead of
   * running its statistics pass.  Needs N % 32 == 0, box[0] % 8 == 0, cs_mult[0] == 0, bf16 output. */
  float* col_accum;
  int32_t cs_mult[4];
} T2VGemmDesc;

int t2v_gemm(const T2VGemmDesc* desc, t2v_stream_t stream);

/*
 * Fused scaled-dot-product attention, head_dim 64, no mask, on tcgen05:
 *   O[b, i, h, :] = softmax_j( scale * Q[b,i,h,:] . K[b,j,h,:] ) V[b,j,h,:]
 * Q/K/V/O are bf16 with arbitrary batch/token/head strides (head_dim contiguous), so the kernel
 * reads the projection outputs in place (no "b n (h d) -> (b h) n d" copy).  kv_batch_div lets
 * several query batches share one K/V batch (cross-attention: one text context per video, 16 frames).
 * Replaces CrossAttention.forward / efficient_forward core (attention.py:121-149,198-224) for the
 * spatial self- and cross-attention layers.
 */
typedef struct T2VAttnDesc {
  const void* q; const void* k; const void* v; void* o;
  int32_t batch, heads, len_q, len_k;
  int64_t q_stride_b, q_stride_t, q_stride_h;
  int64_t k_stride_b, k_stride_t, k_stride_h;
  int64_t v_stride_b, v_stride_t, v_stride_h;
  int64_t o_stride_b, o_stride_t, o_stride_h;
  int32_t kv_batch_div;                 /* kv batch index = q batch index / kv_batch_div (>= 1) */
  float scale;
  int32_t causal;                       /* 1: key j is visible to query i iff j <= i (the CLIP text tower's attn_mask,
                                           lvdm/modules/encoders/condition.py:262-266); 0 = no mask (the UNet) */
  float* lse2;                          /* optional fp32 [batch][heads][len_q]: log2(sum_j exp2(scale*log2e * q.k_j)) per row, kept
                                           for t2v_attn_bwd; NULL = not written */
} T2VAttnDesc;

int t2v_attn_fwd(const T2VAttnDesc* desc, t2v_stream_t stream);

/*
 * Short-sequence attention (len <= 32, head_dim 64): one warp per (sequence, head); token stride
 * is arbitrary so the temporal sequences of a [B*T, H*W, C] activation are read in place
 * (token stride = H*W*C).  Replaces the temporal CrossAttention core (attention.py:121-149 under
 * TemporalTransformer, attention.py:471-513).
 */
typedef struct T2VShortAttnDesc {
  const void* q; const void* k; const void* v; void* o;
  int32_t n_seq_outer, n_seq_inner, heads, len;  /* sequence id = (outer, inner) */
  int64_t q_stride_outer, q_stride_inner, q_stride_t, q_stride_h;
  int64_t k_stride_outer, k_stride_inner, k_stride_t, k_stride_h;
  int64_t v_stride_outer, v_stride_inner, v_stride_t, v_stride_h;
  int64_t o_stride_outer, o_stride_inner, o_stride_t, o_stride_h;
  float scale;
  /* optional export of the attention probabilities (attention.py:124-126 `record_attn_probs`, consumed by the motion-prior
   * code: motion_prior_sample.py:40-56): probs[(seq * heads + head)][query][key], seq = outer * n_seq_inner + inner — the
   * reference's "(b h) i j" layout; dtype 0 bf16 / 1 fp16 / 2 fp32; NULL = off */
  void* probs;
  int32_t probs_dtype;
} T2VShortAttnDesc;

int t2v_attn_short_fwd(const T2VShortAttnDesc* desc, t2v_stream_t stream);

/*
 * GroupNorm (+ optional SiLU) over channels-last activations.
 * A sample = rows_per_sample consecutive rows of [rows, C]; statistics over (rows_per_sample x C/groups).
 * The input may be the channel-concatenation of two tensors (UNet skip connections:
 * openaimodel3d.py:732-734) — the output is the normalised concatenation.
 * Replaces GroupNorm32/normalization (basics.py:78-89), nn.GroupNorm (attention.py:340-342,
 * openaimodel3d.py:275-295, ae_modules.py:16-19) and the following SiLU / swish.
 * Samples that fit a thread-block cluster's shared memory run as ONE kernel (cluster per sample, statistics
 * exchanged through distributed shared memory, workspace untouched); larger samples use a statistics + apply pair.
 * workspace: fp32 [n_samples * groups * 2 + 1]; must be ZERO on entry and is left zeroed on return (the
 * kernels clean it themselves, so a buffer zeroed once at allocation can be reused by every call on a stream).
 */
typedef struct T2VGroupNormDesc {
  const void* x[2]; int32_t ch[2];      /* bf16 sources, channels per source (ch[1] = 0 if single) */
  int64_t x_row_stride[2];
  void* out; int64_t out_row_stride;    /* bf16 [rows, ch0+ch1] */
  const float* gamma; const float* beta;/* fp32 [C] */
  int64_t rows; int64_t rows_per_sample;
  int32_t groups; float eps; int32_t silu;
  float* workspace;
  int32_t mode;                         /* 0 = automatic; 1 = force the two-kernel path; 2 = require the single-kernel cluster path */
  /* Optional: per-channel (sum, sum of squares) accumulated by the GEMMs that produced x[0] / x[1]
   * (T2VGemmDesc.col_accum): fp32 [n_samples * chan_group][ch[i]][2].  When chan_sums[0] is set only the apply kernel
   * runs (no statistics pass, workspace unused); chan_group producer samples are summed per GroupNorm sample (a
   * GroupNorm over (t, h, w) fed with per-frame sums: chan_group = t). */
  const float* chan_sums[2];
  int32_t chan_group;
} T2VGroupNormDesc;

int t2v_groupnorm(const T2VGroupNormDesc* desc, t2v_stream_t stream);

/* LayerNorm over the last dim of bf16 [rows, C]; fp32 affine. Replaces nn.LayerNorm (attention.py:279-281). */
typedef struct T2VLayerNormDesc {
  const void* x; int64_t x_row_stride;
  void* out; int64_t out_row_stride;
  const float* gamma; const float* beta;
  int64_t rows; int32_t channels; float eps;
} T2VLayerNormDesc;

int t2v_layernorm(const T2VLayerNormDesc* desc, t2v_stream_t stream);

/* Per-row LayerNorm statistics only: stats[row] = (rstd, -rstd * mean), fp32 [rows][2] (desc->out, gamma, beta are
 * ignored).  Feeds T2VGemmDesc.row_stats. */
int t2v_layernorm_stats(const T2VLayerNormDesc* desc, float* stats, t2v_stream_t stream);

/*
 * Small-M linear on CUDA cores: out[m, n] = act_out( sum_k act_in(x[m,k]) W[n,k] + bias[n] + add[m,n] ).
 * fp32 in/out, bf16 weights; for the timestep / fps / guidance embedding MLPs and the 22 ResBlock
 * emb projections (openaimodel3d.py:403-430,683-711,172-178) where M = batch size.
 */
typedef struct T2VSmallLinearDesc {
  const float* x; int64_t x_row_stride;
  const void* w;                        /* bf16 [N][K] */
  const float* bias;                    /* fp32 [N] or NULL */
  const float* add; int64_t add_row_stride; /* fp32 [M][N] or NULL */
  float* out; int64_t out_row_stride;
  int32_t m, n, k;
  int32_t silu_in;                      /* apply SiLU to x on load */
  int32_t silu_out;
  int32_t round_bf16;                   /* round the result through bf16 (mimics a bf16 module) */
} T2VSmallLinearDesc;

int t2v_small_linear(const T2VSmallLinearDesc* desc, t2v_stream_t stream);

/* Sinusoidal embedding: out[m, :] = concat(cos(t*f), sin(t*f)) (timestep_embedding,
 * utils_diffusion.py:8-32) or concat(sin, cos) when sin_first != 0 (get_w_embedding,
 * t2v_turbo_vc2_pipeline.py:99-120).  freqs: fp32 [half] (built on the host with the reference's
 * own formula); out: fp32 [m, 2*half]; round_bf16 mimics the `.to(self.dtype)` cast. */
int t2v_sinusoidal_embedding(const float* t, const float* freqs, float* out, int32_t m,
                             int32_t half, int32_t sin_first, int32_t round_bf16,
                             t2v_stream_t stream);

/* Direct 3x3 convolution for tiny channel counts on the input side (C_in <= 8): latent -> features.
 * in: channels-last bf16 [N,H,W,Cin]; w: bf16 [Cout][3][3][Cin]; out bf16 [N,H,W,Cout].
 * Replaces input_blocks.0.0 (openaimodel3d.py:433-437) and the VAE conv_in (ae_modules.py:545-547). */
int t2v_conv3x3_small_cin(const void* in, const void* w, const float* bias, void* out, int32_t n,
                          int32_t h, int32_t wdt, int32_t cin, int32_t cout, t2v_stream_t stream);

/* Layout / resampling helpers (all bf16 unless stated). */
/* [B,C,T,H,W] (any float dtype given by in_dtype: 0 bf16, 1 fp16, 2 fp32) -> [B*T,H,W,C] bf16, times scale */
int t2v_bcthw_to_frames(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c,
                        int32_t t, int32_t h, int32_t w, float scale, t2v_stream_t stream);
/* same with the channel count padded to c_pad (<= 64) by zero channels: RGB video -> 4-channel frames for the
 * encoder's direct small-Cin conv (ae_modules.py:411-413); 4-channel latents -> 64 channels for the training path's conv_in */
int t2v_bcthw_to_frames_pad(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c, int32_t c_pad,
                            int32_t t, int32_t h, int32_t w, float scale, t2v_stream_t stream);
/* same, followed by a per-pixel channel mix out[o] = sum_c mix[o][c] * scale * in[c] + bias[o]  (c <= 8):
 * `1/scale_factor * z` + post_quant_conv 1x1 (ddpm3d.py:669, autoencoder.py:111). mix/bias: fp32 device arrays */
int t2v_bcthw_to_frames_mix(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c,
                            int32_t t, int32_t h, int32_t w, float scale, const float* mix,
                            const float* bias, t2v_stream_t stream);
/* [B*T,H,W,C_pad] bf16 (first c channels) -> [B,C,T,H,W] in out_dtype */
int t2v_frames_to_bcthw(const void* in, int32_t c_pad, void* out, int32_t out_dtype, int32_t b,
                        int32_t c, int32_t t, int32_t h, int32_t w, t2v_stream_t stream);
/* nearest 2x upsample of [N,H,W,C] -> [N,2H,2W,C] (F.interpolate nearest: openaimodel3d.py:104-109, ae_modules.py:118-120) */
int t2v_upsample_nearest2x(const void* in, void* out, int32_t n, int32_t h, int32_t w, int32_t c,
                           t2v_stream_t stream);
/* out[r, :] = concat(a[r, :ca], b[r, :cb]) (torch.cat dim=1: openaimodel3d.py:733) */
int t2v_concat_channels(const void* a, int32_t ca, const void* b, int32_t cb, void* out,
                        int64_t rows, t2v_stream_t stream);
/* row softmax over bf16 [rows, cols] in place, with pre-scale (VAE AttnBlock ae_modules.py:60-62) */
int t2v_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t row_stride, float scale,
                     t2v_stream_t stream);

/*
 * Fused LCM scheduler step (T2VTurboScheduler.step, scheduler/t2v_turbo_scheduler.py:367-467):
 *   x0 = (x - sqrt(1-a_t) * eps) / sqrt(a_t);  den = c_out * x0 + c_skip * x;
 *   prev = sqrt(a_prev) * den + sqrt(1 - a_prev) * noise
 * Elementwise over n values; x/eps/noise/prev/den share dtype (0 bf16, 1 fp16, 2 fp32); the
 * intermediate roundings of the reference's tensor-dtype arithmetic are reproduced (torch divides
 * by a CPU scalar as a multiply by its fp32 reciprocal, hence inv_sqrt_alpha_t).  noise may be
 * NULL (single-step sampling: prev = denoised).
 */
int t2v_lcm_step(const void* x, const void* eps, const void* noise, void* prev, void* denoised,
                 int64_t n, int32_t dtype, float inv_sqrt_alpha_t, float sqrt_beta_t, float c_skip,
                 float c_out, float sqrt_alpha_prev, float sqrt_beta_prev, t2v_stream_t stream);

/*
 * out[r, i] = rnd( rnd(a[r] * x[r, i]) + rnd(b[r] * y[r, i]) ) over [rows, row_len] tensors of one dtype (0 bf16,
 * 1 fp16, 2 fp32), a / b fp32 per-row scalars on the device, rnd = round to the tensor dtype after every tensor op as
 * torch does.  T2VTurboScheduler.add_noise (scheduler/t2v_turbo_scheduler.py:470-495, rows = samples), and the DDIM
 * solver / predicted-x0 lines of the distillation step (ode_solver/ddim_solver.py:67-87, utils/common_utils.py:87-133).
 * y (and b) may be NULL: out = rnd(a[r] * x).
 */
int t2v_scale_add_rows(const void* x, const void* y, const float* a, const float* b, void* out, int64_t rows,
                       int64_t row_len, int32_t dtype, t2v_stream_t stream);

/* Steps either side of the denoising path.
 * out[i, :] = table[ids[i], :] + pos[i % ctx, :] as bf16: token + positional embedding of the OpenCLIP text tower
 * (condition.py:262-263); table / pos dtype 0 bf16 / 1 fp16 / 2 fp32; ids int64. */
int t2v_embedding_gather(const void* table, const void* pos, int32_t dtype, const int64_t* ids, void* out, int64_t n, int32_t width,
                         int32_t ctx, int32_t vocab, t2v_stream_t stream);
/* video [B, 3, T, H, W] (dtype 0 / 1 / 2) -> uint8 [B, T, H, W, 3] = trunc((clamp(v, -1, 1) + 1) / 2 * 255): the tensor
 * post-processing of app.py:90-94 in one pass (the h264 encode that follows is host code, out of scope). */
int t2v_video_to_uint8(const void* video, int32_t dtype, uint8_t* out, int32_t b, int32_t t, int32_t h, int32_t w, t2v_stream_t stream);

/* KL-VAE posterior (lvdm/distributions.py:24-42 + ddpm3d.py:558-567): moments fp32 channels-last
 * [B*T, H, W, 2*zc] = (mean | logvar) -> out [B, zc, T, H, W] (out_dtype 0 bf16 / 1 fp16 / 2 fp32)
 * = scale * (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise); noise fp32 [B*T, zc, H, W] or NULL (posterior mode). */
int t2v_gaussian_sample(const float* moments, const float* noise, void* out, int32_t out_dtype, int32_t b,
                        int32_t t, int32_t h, int32_t w, int32_t zc, float scale, t2v_stream_t stream);

/* ------------------------------------------------------------------ consistency-distillation step (LoRA training)
 * Backward twins of the LoRA-injected layers (utils/lora.py:19-230: y = W x + b + scale * dropout(U (D x))):
 *   dgrad  — dx = dy W (+ d(Dx) D): the forward t2v_gemm on pre-transposed / tap-rotated weights (host packing);
 *   wgrad  — the LoRA weight gradients, accumulated into the caller's fp32 gradient ARENA by t2v_wgrad:
 *            out[j, c, tap] += alpha * sum_points A[point + off(tap), c] * B[point, j],  c < a_ch, j < b_cols <= 64
 *            (lora_down: A = layer input, B = d(Dx), taps of the base kernel; lora_up: A = scale * mask * dy, B = Dx).
 * A / B are channels-last bf16 read in place as MN-major tensor-core operands (no transposed copies, no im2col; a tap
 * is a coordinate offset, conv padding is TMA zero fill).  out is ACCUMULATED into (zero the arena once per step):
 * element (j, c, tap) lives at out[j * out_j_stride + c * out_c_stride + tap * out_tap_stride]. */
typedef struct T2VWgradDesc {
  const void* a;                        /* bf16 [points..][a_ch] channels-last (point grid a_size, strides a_stride) */
  int32_t a_ch;                         /* multiple of 8 */
  int64_t a_size[T2V_MAX_DIMS];
  int64_t a_stride[T2V_MAX_DIMS];
  const void* b;                        /* bf16 [points..][b_cols] over the OUTPUT point grid o_size, strides b_stride */
  int32_t b_cols;                       /* multiple of 8, <= 64 (the LoRA rank) */
  int64_t o_size[T2V_MAX_DIMS];
  int64_t b_stride[T2V_MAX_DIMS];
  int32_t box[T2V_MAX_DIMS];            /* point tile: product a multiple of 16, <= 128 */
  int32_t n_taps;
  int32_t tap_off[T2V_MAX_TAPS][T2V_MAX_DIMS];
  float* out;                           /* fp32 gradient arena slice (accumulated into) */
  int64_t out_j_stride, out_c_stride, out_tap_stride;
  float alpha;
} T2VWgradDesc;

int t2v_wgrad(const T2VWgradDesc* desc, t2v_stream_t stream);

/* out[i] = x[i] * scale * (mask ? mask[i] : 1)  (bf16; mask is a uint8 keep-mask): the `dropout(...) * scale` of the LoRA
 * branch in the forward (applied to U(Dx) before the add) and its adjoint on dy in the backward (utils/lora.py:45-50). */
int t2v_scale_mask(const void* x, const uint8_t* mask, void* out, int64_t n, float scale, t2v_stream_t stream);

/* Training-mode dropout fused with its scale: out[i] = x[i] * scale * keep[i] (+ addend[i]), keep[i] ~ Bernoulli(keep_prob) drawn in the
 * kernel (Philox4x32-10 keyed by *seed, stream (call_id, i / 8)); mask_out[i] = keep[i] (uint8) for the adjoint
 * (t2v_scale_mask on dy).  Replaces nn.Dropout after lora_up (utils/lora.py:37,45-50: `self.dropout(self.lora_up(...)) *
 * self.scale`) and the TemporalConvBlock dropouts (lvdm/modules/networks/openaimodel3d.py:280-296): the same distribution,
 * not torch's random stream.  addend (bf16, may be NULL) folds the residual add that follows the layer
 * (`x + attn(...)`, `h + out_layers(...)`: attention.py:276-281, openaimodel3d.py:246-250) into the same pass: the layer's base GEMM
 * takes `out` as its residual operand.  seed is read on the device, so a captured CUDA graph draws a fresh mask on every replay
 * once the caller advances *seed between steps; n % 8 == 0. */
int t2v_dropout_scale(const void* x, const void* addend, void* out, uint8_t* mask_out, int64_t n, float keep_prob, float scale,
                      const uint64_t* seed, uint32_t call_id, t2v_stream_t stream);

/* Fused AdamW over the flat fp32 LoRA parameter / gradient arenas (torch.optim.AdamW semantics, one launch for all 575
 * layers; train_t2v_turbo_v1_lora.py:897-906,1193): grad is multiplied by grad_scale first (1/world for the DDP mean
 * and the gradient-clipping factor folded in). step >= 1. */
int t2v_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int32_t step, float grad_scale, t2v_stream_t stream);

/* out[0] += sum_i x[i]^2 over n fp32 values (gradient-norm clipping, accelerator.clip_grad_norm_, :1191); out is
 * accumulated into (zero it first). */
int t2v_sum_squares(const float* x, int64_t n, float* out, t2v_stream_t stream);

/* mean-squared-error loss and its gradient (F.mse_loss(model_pred.float(), target.float()), :1183-1186):
 * loss[0] += sum (a - b)^2 / n ; grad[i] = 2 (a[i] - b[i]) / n * grad_scale in the dtype of a (0 bf16, 1 fp16, 2 fp32). */
int t2v_mse_loss_grad(const void* a, const void* b, void* grad, float* loss, int64_t n, int32_t dtype, float grad_scale,
                      t2v_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Student backward (train_t2v_turbo_v1_lora.py:1190 `accelerator.backward(distill_loss)` through the LoRA-injected UNet):
 * adjoints of the non-GEMM layers.  Only the LoRA weights train, so none of these produces gamma / beta / bias gradients.
 * Gradients are channels-last bf16 like the activations; arithmetic and statistics fp32. */

/* GroupNorm (+SiLU) backward (adjoint of t2v_groupnorm; basics.py:78-89, openaimodel3d.py:155-159,179-184,275-295):
 * dx = rstd * (dyh - mean_g(dyh) - xh * mean_g(dyh * xh)) (+ dx_add), dyh = dy * act'(xh * gamma + beta) * gamma.
 * workspace: fp32 [n_samples * groups * 4], ZERO on entry (left holding the statistics). */
typedef struct T2VGroupNormBwdDesc {
  const void* x; int64_t x_row_stride;       /* forward input, bf16 [rows, channels] */
  const void* dy; int64_t dy_row_stride;     /* gradient of the forward output */
  const void* dx_add; int64_t dx_add_row_stride; /* optional: added to dx (a second gradient path into x); may be NULL */
  void* dx; int64_t dx_row_stride;
  const float* gamma; const float* beta;
  int64_t rows, rows_per_sample;
  int32_t channels, groups; float eps; int32_t silu;
  float* workspace;
} T2VGroupNormBwdDesc;
int t2v_groupnorm_bwd(const T2VGroupNormBwdDesc* desc, t2v_stream_t stream);

/* LayerNorm backward (attention.py:279-281,300-311), one warp per row, statistics recomputed: channels in
 * {64, 128, 256, 320, 512, 640, 1024, 1280}; dx_add (optional) is the residual-path gradient summed into dx. */
int t2v_layernorm_bwd(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, const void* dx_add,
                      int64_t dx_add_row_stride, void* dx, int64_t dx_row_stride, const float* gamma, int64_t rows,
                      int32_t channels, float eps, t2v_stream_t stream);

/* GEGLU (attention.py:516-523), unfused for training: pre = [a | gate] bf16 [rows, 2*inner].
 * dout == NULL: out[rows, inner] = a * gelu_erf(gate);  else: out[rows, 2*inner] = d(pre) given dout[rows, inner]. */
int t2v_geglu(const void* pre, int64_t pre_row_stride, const void* dout, int64_t dout_row_stride, void* out,
              int64_t out_row_stride, int64_t rows, int32_t inner, t2v_stream_t stream);

/* Row-strided elementwise bf16 [rows, channels]: op 0 out = a + b (gradient accumulation at a fan-out);
 * op 1 out = silu(a); op 2 out = b * silu'(a) (the SiLU of emb_layers / time_embed, openaimodel3d.py:172-178,403-411). */
int t2v_ew2d(int32_t op, const void* a, int64_t a_row_stride, const void* b, int64_t b_row_stride, void* out,
             int64_t out_row_stride, int64_t rows, int32_t channels, t2v_stream_t stream);

/* out[s, c] += sum of x[row, c] over the rows of sample s (fp32 [rows / rows_per_sample, channels], accumulated into):
 * adjoint of the broadcast timestep-embedding add h + emb_out[:, :, None, None] (openaimodel3d.py:237-246). */
int t2v_colsum_samples(const void* x, int64_t x_row_stride, float* out, int64_t rows, int64_t rows_per_sample,
                       int32_t channels, t2v_stream_t stream);

/* 2x resampling of [n, h, w, channels] frames; (h_out, w_out) are the OUTPUT extents.
 * mode 0: out[y, x] = in[2y, 2x] (a stride-2 / pad-1 conv is the stride-1 conv subsampled: Downsample, openaimodel3d.py:104-111);
 * mode 1: zero stuffing out[2y, 2x] = in[y, x], 0 elsewhere (adjoint of mode 0);
 * mode 2: out[y, x] = sum of the 2x2 block of in (adjoint of the nearest 2x upsampling, openaimodel3d.py:65-72). */
int t2v_resample2x(int32_t mode, const void* in, void* out, int64_t n, int32_t h_out, int32_t w_out, int32_t channels,
                   t2v_stream_t stream);

/* delta[b, h, i] = sum_d dO[b,i,h,d] * O[b,i,h,d] (fp32 [batch][heads][len]): the row term of the softmax Jacobian. */
int t2v_attn_delta(const void* o, int64_t o_stride_b, int64_t o_stride_t, int64_t o_stride_h, const void* d_o,
                   int64_t do_stride_b, int64_t do_stride_t, int64_t do_stride_h, float* delta, int32_t batch, int32_t heads,
                   int32_t len, t2v_stream_t stream);

/* Flash-attention backward on tcgen05 (adjoint of t2v_attn_fwd, head_dim 64, no mask; attention.py:121-149,198-224):
 *   P = exp2(scale*log2e * Q K^T - lse2),  dP = dO V^T,  dS = P * (dP - delta) * scale
 *   dQ = dS K,   dK = dS^T Q,   dV = P^T dO
 * lse2[b][h][i] = log2-domain log-sum-exp of row i (written by t2v_attn_fwd when T2VAttnDesc.lse2 is set), delta from
 * t2v_attn_delta.  Two launches of one kernel: row tiles over queries (dQ) and row tiles over keys (dK, dV; the query
 * batches sharing a K/V batch are looped inside the CTA, no atomics).  Any of dq / (dk, dv) may be NULL to skip that half. */
typedef struct T2VAttnBwdDesc {
  const void* q; const void* k; const void* v; const void* d_o;
  const float* lse2; const float* delta;
  void* dq; void* dk; void* dv;
  int32_t batch, heads, len_q, len_k;
  int64_t q_stride_b, q_stride_t, q_stride_h;       /* q, d_o, dq share the query geometry but have their own strides */
  int64_t k_stride_b, k_stride_t, k_stride_h;
  int64_t v_stride_b, v_stride_t, v_stride_h;
  int64_t do_stride_b, do_stride_t, do_stride_h;
  int64_t dq_stride_b, dq_stride_t, dq_stride_h;
  int64_t dk_stride_b, dk_stride_t, dk_stride_h;
  int64_t dv_stride_b, dv_stride_t, dv_stride_h;
  int32_t kv_batch_div;
  float scale;
} T2VAttnBwdDesc;
int t2v_attn_bwd(const T2VAttnBwdDesc* desc, t2v_stream_t stream);

/* Short-sequence attention backward (adjoint of t2v_attn_short_fwd; temporal self-attention, attention.py:471-513):
 * one warp per (sequence, head), probabilities recomputed; d_o / dq / dk / dv use the strides of o / q / k / v. */
typedef struct T2VShortAttnBwdDesc {
  T2VShortAttnDesc fwd;                 /* q, k, v and their strides, o strides (o itself unused), scale; probs ignored */
  const void* d_o;                      /* gradient of o, o's strides */
  void* dq; void* dk; void* dv;         /* q's / k's / v's strides */
} T2VShortAttnBwdDesc;
int t2v_attn_short_bwd(const T2VShortAttnBwdDesc* desc, t2v_stream_t stream);

/* pseudo-Huber distillation loss and its gradient (huber_loss, utils/common_utils.py:302-304; --loss_type huber is the training
 * script's default): loss[0] += mean(sqrt((a-b)^2 + c^2) - c); grad[i] = (a-b) / sqrt((a-b)^2 + c^2) / n * grad_scale. */
int t2v_huber_loss_grad(const void* a, const void* b, void* grad, float* loss, int64_t n, int32_t dtype, float huber_c,
                        float grad_scale, t2v_stream_t stream);

/* Workspace sizes for callers that are not the Python host mirror (the caller owns every buffer, see the header comment):
 * t2v_gemm_workspace_bytes: the fp32 split-K workspace that ALLOWS t2v_gemm to split K for `desc` (o_size and b_rows are read;
 *   a smaller or NULL workspace is always legal — the GEMM then runs unsplit); zero it once and pass T2V_WS_CLEAN afterwards.
 * t2v_groupnorm_workspace_bytes: T2VGroupNormDesc.workspace (backward = 0) / T2VGroupNormBwdDesc.workspace (backward = 1).
 * Both return -1 on a malformed argument; neither touches the device. */
int64_t t2v_gemm_workspace_bytes(const T2VGemmDesc* desc);
int64_t t2v_groupnorm_workspace_bytes(int64_t n_samples, int32_t groups, int32_t backward);

/* ---- full fine-tune step only (train_latent_t2v_turbo_v2.py:945-1276: every UNet parameter trains) ----
 * Affine gradients of GroupNorm(+SiLU) (lvdm/basics.py:78-89 GroupNormSpecific; the `normalization(ch)` + SiLU pairs of
 * openaimodel3d.py:155-159,179-184,275-295):  dgamma[c] += sum dpre * xh,  dbeta[c] += sum dpre,  dpre = dy * act'(xh * gamma + beta),
 * over all rows of all samples (fp32, accumulated into — the gradient arena).  stats_ws is the workspace t2v_groupnorm_bwd filled
 * for the same (x, rows_per_sample, groups): fp32 [n_samples][groups][4], slots 0 / 1 = sum x / sum x^2.  channels <= 2560. */
int t2v_groupnorm_affine_grad(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, const float* gamma,
                              const float* beta, const float* stats_ws, float* dgamma, float* dbeta, int64_t rows,
                              int64_t rows_per_sample, int32_t channels, int32_t groups, float eps, int32_t silu,
                              t2v_stream_t stream);

/* Affine gradients of LayerNorm (attention.py:279-281 norm1..3): dgamma[c] += sum_rows dy * xh, dbeta[c] += sum_rows dy
 * (fp32, accumulated into); channels as t2v_layernorm_bwd. */
int t2v_layernorm_affine_grad(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, float* dgamma,
                              float* dbeta, int64_t rows, int32_t channels, float eps, t2v_stream_t stream);

/* Backward of t2v_softmax_rows (P = softmax(scale * S) per row), in place on dp:
 *   dS[r, c] = scale * P[r, c] * (dP[r, c] - sum_j dP[r, j] * P[r, j])        bf16 [rows, cols], row strides in elements.
 * The KL-VAE decoder's AttnBlock (ae_modules.py:48-73) under `vae.decode` WITH grad — the reward terms of the training scripts
 * (train_t2v_turbo_v1_lora.py:1055-1098, train_latent_t2v_turbo_v2.py:1062-1166). */
int t2v_softmax_bwd_rows(void* dp, int64_t dp_row_stride, const void* p, int64_t p_row_stride, int64_t rows, int32_t cols,
                         float scale, t2v_stream_t stream);

/* Adjoint of the PROBABILITIES export of t2v_attn_short_fwd w.r.t. q and k (temporal self-attention, head dim 64, len <= 16):
 *   P = softmax(scale q k^T),  dS = P * (dP - rowsum(dP * P)),  dq = scale dS k,  dk = scale dS^T q.
 * q / k: bf16 token matrices [(outer, len, inner), heads * 64] with row strides in elements (views of a fused projection are fine);
 * d_probs: fp32 [(seq * heads + head)][query][key], seq = outer * n_inner + inner (the export's layout); dq / dk: bf16, contiguous
 * [rows, heads * 64].  The motion-prior score of the v2 preprocessing (motion_prior_sample.py:59-84: the gradient w.r.t. the latents
 * of a loss on the temporal attention probabilities of output_blocks.3-11). */
int t2v_attn_short_probs_bwd(const void* q, int64_t q_row_stride, const void* k, int64_t k_row_stride, const float* d_probs,
                             void* dq, void* dk, int32_t n_outer, int32_t n_inner, int32_t heads, int32_t len, float scale,
                             t2v_stream_t stream);

/* EMA of the target network's parameters over the flat fp32 arenas (update_ema, utils/common_utils.py:308-319;
 * train_latent_t2v_turbo_v2.py:1273-1276): target = target * rate + src * (1 - rate). */
int t2v_ema_update(float* target, const float* src, int64_t n, float rate, t2v_stream_t stream);

/* Weight packing helpers (device-side, run once at load). */
/* conv weight [Cout][Cin][kh*kw] (torch OIHW / OIDHW flattened taps) -> [Cout][taps][Cin] bf16 */
int t2v_pack_conv_weight(const void* w, int32_t w_dtype, void* out, int32_t cout, int32_t cin,
                         int32_t taps, t2v_stream_t stream);
/* GEGLU projection rows [2*inner][K]: interleave value/gate rows in blocks of 16 */
int t2v_pack_geglu_rows(const void* w, int32_t w_dtype, void* out, const void* bias,
                        int32_t bias_dtype, float* bias_out, int32_t inner, int32_t k,
                        t2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* T2V_B200_H_ */
