/* gradtts_b200 -- C ABI of the B200-native Grad-TTS hot path (reverse-diffusion mel decoder + MAS).
 *
 * Plain pointers and sizes only; no torch types.  Every pointer is a DEVICE pointer on the handle's
 * GPU unless the name says `_host`.  Every function that launches work takes the CUDA stream as a
 * `void*` (cudaStream_t); nothing synchronises unless stated.  Return value 0 = success; non-zero =
 * failure with a message available from gtts_last_error() (the Python host raises RuntimeError).
 * There is no CPU fallback: on a machine without an sm_100 GPU gtts_decoder_create fails.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   gtts_mas_maximum_path          model/monotonic_align/__init__.py:8-23  maximum_path(value, mask)
 *   gtts_mas_maximum_path_c        model/monotonic_align/core.pyx:40-45    maximum_path_c(paths, values, t_xs, t_ys, max_neg_val)
 *   gtts_align_log_prior / _outputs model/tts.py:143-149,155,184-185       log-prior ahead of MAS, durations and mu_y after it
 *   gtts_forward_diffusion / gtts_score_loss model/diffusion.py:244-252,274-281  forward_diffusion, the scalar of loss_t (forward value)
 *   gtts_decoder_reverse_diffusion model/diffusion.py:254-272  Diffusion.forward / reverse_diffusion(z, mask, mu, n_timesteps, stoc, spk)
 *   gtts_decoder_estimator         model/diffusion.py:174-216  GradLogPEstimator2d.forward(x, mask, mu, t, spk)
 *   gtts_decoder_create/set_param  model/diffusion.py:128-172,227-242  module construction + load_state_dict
 *   gtts_encoder_create/set_param/forward  model/text_encoder.py:285-335  TextEncoder(...), load_state_dict, forward(x, x_lengths, spk)
 *   gtts_vocoder_create/set_param/forward  hifi-gan/models.py:77-118  Generator(h), load_state_dict + remove_weight_norm, forward(mel)
 */
#ifndef GRADTTS_B200_H
#define GRADTTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gtts_decoder gtts_decoder;

#define GTTS_FLAG_FP32 1   /* true-fp32 arithmetic (FFMA convs, precise libm); default is bf16 tensor-core mode */
#define GTTS_FLAG_SDE  2   /* upstream Grad-TTS stochastic update with caller-supplied noise (extension; the reference fork is ODE-only):
                              x <- (x - ((0.5*(mu-x) - score)*beta*h + noise*sqrt(beta*h))) * mask -- the noise is subtracted, as upstream
                              does; BASELINE.json's "+ sqrt(beta*h)*z" form is the same update with -noise */

int gtts_version(void);
const char* gtts_last_error(void);
/* number of visible CUDA devices with compute capability 10.x (0 if none / no driver) */
int gtts_sm100_device_count(void);

/* ---- Monotonic Alignment Search ------------------------------------------------------------------
 * value, mask: [B][t_x][t_y] fp32.  path: same shape, fp32, entries {0,1} (the reference returns
 * value.dtype).  Lengths are taken from the mask exactly as the reference does:
 * t_x = sum_x mask[b,x,0], t_y = sum_y mask[b,0,y]; the DP runs on value*mask.
 * bits_ws: scratch for the direction bits when t_x*t_y is too large for shared memory
 * (gtts_mas_workspace_bytes() bytes, may be NULL when that returns 0).
 * status: one int32 on the device, set to 1 if any item had t_x > t_y (the reference's undefined case;
 * that item's path is left all-zero).  The caller zero-initialises nothing: path and status are cleared here. */
size_t gtts_mas_workspace_bytes(int B, int t_x, int t_y);
int gtts_mas_maximum_path(const float* value, const float* mask, float* path, int B, int t_x, int t_y,
                          void* bits_ws, size_t bits_ws_bytes, int32_t* status, void* stream);
/* Same contract as the Cython entry point: int32 paths (cleared here), values already masked (NOT modified,
 * unlike the reference which mutates its private copy), explicit lengths. */
int gtts_mas_maximum_path_c(int32_t* paths, const float* values, const int32_t* t_xs, const int32_t* t_ys, int B,
                            int t_x, int t_y, float max_neg_val, void* bits_ws, size_t bits_ws_bytes,
                            int32_t* status, void* stream);
/* Host-buffer convenience (pinned or pageable): H2D, kernel, D2H, stream sync. Returns status via *status_host. */
int gtts_mas_maximum_path_host(const float* value_host, const float* mask_host, float* path_host, int B, int t_x,
                               int t_y, int32_t* status_host, int device);

/* ---- Alignment stage around MAS (training / scoring path, model/tts.py:139-185) ---------------------
 * gtts_align_log_prior  model/tts.py:143-149  log_prior[b,i,j] = log N(y[b,:,j]; mu_x[b,:,i], I): the value MAS maximises.
 *   mu_x: [B][n_feats][t_x], y: [B][n_feats][t_y], log_prior: [B][t_x][t_y], fp32 device pointers.
 * gtts_align_outputs    model/tts.py:155      logw_[b,i] = log(1e-8 + sum_j attn[b,i,j]) * x_mask[b,i]   (logw may be NULL)
 *                       model/tts.py:184-185  mu_y[b,:,j] = sum_i attn[b,i,j] mu_x[b,:,i]                 (mu_y may be NULL)
 *   attn: [B][t_x][t_y] fp32 path from gtts_mas_maximum_path, x_mask: [B][t_x]. */
int gtts_align_log_prior(const float* mu_x, const float* y, float* log_prior, int B, int n_feats, int t_x, int t_y, void* stream);
int gtts_align_outputs(const float* attn, const float* mu_x, const float* x_mask, float* logw, float* mu_y, int B, int n_feats,
                       int t_x, int t_y, void* stream);

/* ---- Forward value of the training objective (model/diffusion.py:244-252, :274-281) -----------------
 * gtts_forward_diffusion  Diffusion.forward_diffusion with caller-supplied N(0,1) noise (the reference draws it with
 *   torch.randn inside, :249): xt = (x0 e^{-c/2} + mu (1 - e^{-c/2}) + noise sqrt(1 - e^{-c})) mask, z_masked = noise mask,
 *   c = beta_min t + 0.5 (beta_max - beta_min) t^2 per sample.  x0, mu, noise, xt, z_masked: [B][n_feats][T]; mask: [B][T];
 *   t: [B]; T % 4 == 0.
 * gtts_score_loss         the scalar of Diffusion.loss_t (:278-280) given the estimator output:
 *   loss = sum (noise_estimation sqrt(1 - e^{-c}) + z_masked)^2 / (sum(mask) n_feats), one fp32 on the device.
 *   ws: gtts_score_loss_workspace_bytes() bytes of device scratch.  Forward value only: there is no backward. */
size_t gtts_score_loss_workspace_bytes(void);
int gtts_forward_diffusion(const float* x0, const float* mask, const float* mu, const float* t, const float* noise, float* xt,
                           float* z_masked, int B, int n_feats, int T, double beta_min, double beta_max, void* stream);
int gtts_score_loss(const float* noise_estimation, const float* z_masked, const float* mask, const float* t, void* ws,
                    size_t ws_bytes, float* loss, int B, int n_feats, int T, double beta_min, double beta_max, void* stream);

/* ---- Decoder --------------------------------------------------------------------------------------
 * n_spks follows the reference constructor: 1 (or <2) = two input channels; >1 = speaker channel, spk
 * required; -1 = spk_mlp exists but is unused (params_tedlium.py). */
int gtts_decoder_create(gtts_decoder** out, int n_spks, int n_feats, int dim, double beta_min, double beta_max,
                        double pe_scale, int device);
void gtts_decoder_destroy(gtts_decoder* d);
/* Upload one tensor of Diffusion.state_dict() (key e.g. "estimator.downs.0.0.block1.block.0.weight"),
 * fp32, contiguous, PyTorch layout; `data` may be a host or a device pointer.  Invalidates packed weights/plans. */
int gtts_decoder_set_param(gtts_decoder* d, const char* name, const float* data, size_t numel);
/* options: "max_chunk" (samples per workspace chunk, <= 64), "max_plans" (LRU bound on cached per-(B,T) plans; default 24),
 * "trim" (any value: free every cached plan and the pooled workspace), "use_graph" (0/1), "conv_impl_bf16" (1 tcgen05, 0 FFMA cross-check),
 * "halo_mode" (3x3 convs: 0 per-tap TMA boxes, 1 / 2 halo box 18x16 / 18x10 + shifted descriptor views),
 * "fused_attn" (1: fused k-projection + context kernel for C <= 128, 0: 1x1 kv conv + context kernel),
 * "fuse_epi" (Block convs finish GroupNorm+Mish(+time bias / residual) in their own epilogue -- no raw tensor, no gn_apply pass:
 *  0 never, 1 (default) for plans of at most "fuse_epi_max_b" (default 2) samples, where it wins; 2 always),
 * "fuse_async" (1, default: in larger plans the same fusion is done by asynchronous apply warps inside the conv, raw tile via L2; 0: gn_apply pass),
 * "fp32_tc" (1, default: fp32 mode runs its convolutions on the tensor cores as six bf16 partial products; 0: CUDA-core FFMA),
 * "fuse_gn" (1: block2 convs apply block1's GroupNorm+Mish on their operand tiles; only used with fuse_epi = 0) */
int gtts_decoder_set_option(gtts_decoder* d, const char* key, int value);

/* z, mu, out: [B][80][T] fp32; mask: [B][1][T] fp32 with entries in {0,1}; spk: [B][64] or NULL;
 * noise: [n_timesteps][B][80][T] (only with GTTS_FLAG_SDE).  T % 4 == 0.  out may alias z. */
int gtts_decoder_reverse_diffusion(gtts_decoder* d, const float* z, const float* mask, const float* mu,
                                   const float* spk, float* out, int B, int T, int n_timesteps, int flags,
                                   const float* noise, void* stream);
/* one score-network evaluation with per-sample times t[B] */
int gtts_decoder_estimator(gtts_decoder* d, const float* x, const float* mask, const float* mu, const float* t,
                           const float* spk, float* out, int B, int T, int flags, void* stream);
/* Score AND vector-Jacobian product w.r.t. the input in one pass (the backward of GradLogPEstimator2d.forward for its `x` argument):
 *   out_score = estimator(x, mask, mu, t, spk)  (may be NULL),  out_gx = (d out_score / d x)^T v,  all [B][80][T] fp32.
 * This is the gradient torch.autograd.grad(sum(fn(x, t) * eps), x) takes through the score network in the Hutchinson divergence of the
 * probability-flow likelihood (reference n_best/likelihood/likelihood.py:27-38).  Parameter gradients are not computed. */
int gtts_decoder_estimator_vjp(gtts_decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                               const float* v, float* out_score, float* out_gx, int B, int T, int flags, void* stream);
/* Full backward of one estimator call (what loss.backward() does to GradLogPEstimator2d in the reference's training step,
 * model/diffusion.py:274-281 reached from train*.py): given the cotangent v of the score,
 *   out_gx, out_gmu  [B][80][T]   gradients w.r.t. the x and mu input planes
 *   out_gs_pix       [B][80][T]   per-pixel gradient of the speaker plane (sum it over T for d/d spk_mlp output; zeros unless n_spks > 1)
 *   out_gtb          [B][1792]    gradient of the 12 per-ResnetBlock time biases (rows in the order of the reference's ModuleList walk:
 *                                 downs.0.0, downs.0.1, downs.1.0, ..., mid_block1, mid_block2, ups.0.0, ..., ups.1.1)
 * and, kept in the handle until the next call, the gradient of every Conv2d / ConvTranspose2d / GroupNorm / Rezero parameter of the
 * estimator in its PyTorch layout, summed over the batch: fetch each with gtts_decoder_get_param_grad (name = state_dict key, e.g.
 * "estimator.downs.0.1.block2.block.0.weight").  The time MLP, the per-block Linear(64, C) and spk_mlp are differentiated by the host
 * (PyTorch) from out_gtb / out_gs_pix.  out_score, out_gmu, out_gs_pix, out_gtb may be NULL. */
int gtts_decoder_estimator_backward(gtts_decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                                    const float* v, float* out_score, float* out_gx, float* out_gmu, float* out_gs_pix, float* out_gtb,
                                    int B, int T, int flags, void* stream);
int gtts_decoder_get_param_grad(gtts_decoder* d, const char* name, float* dst, size_t numel, void* stream);
/* The same gradients in ONE copy: gtts_decoder_param_grad_slot gives a parameter's offset and size (floats) in the flat buffer
 * (name == NULL: offset 0 and the total size; valid after the first estimator_backward call on the handle), and
 * gtts_decoder_get_param_grads_flat copies the whole buffer -- an optimizer step then needs one device copy, not 172. */
int gtts_decoder_param_grad_slot(const gtts_decoder* d, const char* name, size_t* offset, size_t* numel);
int gtts_decoder_get_param_grads_flat(gtts_decoder* d, float* dst, size_t numel, void* stream);
/* Host-buffer variant of reverse_diffusion: copies inputs H2D, runs, copies the mel D2H, synchronises. */
int gtts_decoder_reverse_diffusion_host(gtts_decoder* d, const float* z_host, const float* mask_host,
                                        const float* mu_host, const float* spk_host, float* out_host, int B, int T,
                                        int n_timesteps, int flags);
/* Measurement hook: replays one Euler step of the (min(B,max_chunk), T) sampler plan eagerly `reps` times with a CUDA
 * event pair around every launch (on `stream`) and writes a JSON report
 * {"B":..,"T":..,"ops":[{"name","is_conv","flops","bytes","ms"},...]} (algorithmic flops/bytes per launch) into buf. */
int gtts_decoder_profile_step(gtts_decoder* d, int B, int T, int flags, int reps, char* buf, size_t buflen, void* stream);
/* kernels launched by the last reverse_diffusion / estimator call on this handle */
long gtts_decoder_launches_last_call(const gtts_decoder* d);
/* Workspace bookkeeping (n >= 5): out[0] plans cached, [1] plans created so far, [2] bytes of pooled activation workspace shared by
 * all plans of the handle, [3] live workspace of the largest cached plan after liveness-based buffer reuse, [4] the same plan's
 * footprint with one buffer per tensor.  Calls on one handle are ordered by the library even across streams (plans share the pool). */
int gtts_decoder_cache_info(const gtts_decoder* d, long long* out, int n);

/* ---- Kernel-level test hooks (used by tests/ only) --------------------------------------------------
 * conv: kind 0 = 3x3 s1, 1 = 1x1, 2 = 3x3 s2, 3 = convT 4x4 s2 p1.  impl 0 = FFMA, 1 = tcgen05 per-tap boxes,
 * 2 / 3 = tcgen05 halo box 18x16 / 18x10 pixels with shifted descriptor views (3x3 s1 only; bf16 only).
 * act 0 = fp32, 1 = bf16 activations and packed weights.  weight_pt is the fp32 PyTorch-layout weight.
 * Inputs/outputs are NHWC in the activation type.  gn_stats (B*16 floats, mean/rstd) may be NULL. */
int gtts_test_conv(int impl, int act, int kind, int B, int H, int W, int Cin0, int Cin1, int Cout,
                   const void* src0, const void* src1, const float* weight_pt, const float* bias,
                   const void* residual, const float* mask, void* out, float* gn_stats, int per_sample_weights,
                   void* stream);

/* Test hook: 3x3 stride-1 bf16 conv with the GroupNorm-apply epilogue (the Block of model/diffusion.py:49-58 plus one of the two adds
 * of ResnetBlock.forward, :75-78, in ONE kernel): out = (Mish(GroupNorm8(conv(src) + bias)) [+ tbias[b]] [+ residual]) * mask[b][w].
 * Exactly one of tbias ([B or 1][Cout], stride tb_bstride) / residual (NHWC bf16) is given.  gn_stats: [B][8][2] mean, rstd out.
 * |reps| > 1 re-launches and prints the time per launch to stderr; reps < 0 selects the asynchronous apply-warp variant (the raw tile
 * goes to a scratch tensor, eight extra warps finish the activation behind the per-sample counter) instead of the TMEM-resident one. */
int gtts_test_conv_apply(int B, int H, int W, int Cin0, int Cin1, int Cout, const void* src0, const void* src1, const float* weight_pt,
                         const float* bias, const float* gamma, const float* beta, const float* tbias, int tb_bstride,
                         const void* residual, const float* mask, void* out, float* gn_stats, int reps, void* stream);

/* Test hook: fused k-projection + softmax + context partials of the linear attention (model/diffusion.py:90-100).
 * x: [B][n][C] bf16, wkv: [256][C] bf16 (k rows then v rows), partials: [B][4][chunks][1088] floats
 * (m[32], l[32], ctx[32][32] per head and chunk).  use_tc = 1: tcgen05 kernel (C = 64), 0: mma.sync kernel. */
int gtts_test_attn_xk(const void* x_bf16, const void* wkv_bf16, float* partials, int B, int n, int C, int chunks, int chunk_len,
                      int use_tc, void* stream);

/* Test hook: per-sample folded attention weights M_b = g * Wout . blockdiag(ctxn_b^T) . Wq (to_out after the context einsum and
 * to_qkv's q rows, model/diffusion.py:95-104, as one C x C matrix per sample).  ctxn: [B][4][32][32] floats (normalised contexts),
 * wout: [C][128], wq: [128][C] floats, m_out: [B][C][C] floats (out_bf16 = 0) or bf16 (1).  variant: 0 = 16-row tiles (small
 * batches), 1 = full-row tiles (large batches), -1 = the product path's choice by batch size; the variants agree bit for bit. */
int gtts_test_attn_fold(const float* ctxn, const float* wout, const float* wq, float g, void* m_out, int B, int C, int out_bf16,
                        int variant, void* stream);

/* Test hook: tcgen05 issue-path micro-benchmark (cycles per CTA, averaged over `grid` CTAs): `iters` rounds of
 * {n_mma tcgen05.mma M128xNx16, n_commit tcgen05.commit}, optionally waiting on the last commit every round. */
int gtts_test_issue_microbench(int N, int n_mma, int n_commit, int iters, int wait_each, int grid, double* issue_cycles,
                               double* total_cycles);

/* ---- Text encoder (the step before the decoder, model/tts.py:84) ---------------------------------------
 * Constructor arguments as model/text_encoder.py:286-288 (window_size < 0 = None); parameters by their reference state_dict key
 * ("emb.weight", "prenet.conv_layers.0.weight", "encoder.attn_layers.3.emb_rel_k", "proj_w.norm_1.gamma", ...), fp32, PyTorch
 * layout, host or device pointer.  forward = TextEncoder.forward in eval mode: tokens (B, T) int64, lengths (B) int64, spk
 * (B, spk_emb_dim) fp32 or NULL -> mu (B, n_feats, T), logw (B, 1, T), x_mask (B, 1, T), all fp32.  fp32 arithmetic throughout
 * (the output decides integer durations).  gtts_encoder_check_tokens reports a token id outside [0, n_vocab) seen by the last
 * forward (synchronises the stream). */
typedef struct gtts_encoder gtts_encoder;
int gtts_encoder_create(gtts_encoder** out, int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp,
                        int n_heads, int n_layers, int kernel_size, int window_size, int spk_emb_dim, int n_spks, int device);
void gtts_encoder_destroy(gtts_encoder* e);
int gtts_encoder_set_param(gtts_encoder* e, const char* name, const float* data, size_t numel);
int gtts_encoder_forward(gtts_encoder* e, const int64_t* tokens, const int64_t* lengths, const float* spk, float* mu, float* logw,
                         float* x_mask, int B, int T, void* stream);
int gtts_encoder_check_tokens(gtts_encoder* e, void* stream);
long gtts_encoder_launches_last_call(const gtts_encoder* e);

/* ---- HiFi-GAN generator (vocoder; the step after the decoder, inference.py:73-76,97) ---------------------
 * Configuration = the fields of checkpts/hifigan-config.json that hifi-gan/models.py:77-99 reads: resblock ("1" / "2"),
 * upsample_rates / upsample_kernel_sizes (n_ups entries), upsample_initial_channel, resblock_kernel_sizes (n_rb <= 3 entries) and
 * resblock_dilation_sizes (n_rb rows of n_dil entries, row-major).  Parameters are set by their reference state_dict key AFTER
 * remove_weight_norm ("conv_pre.weight", "ups.0.bias", "resblocks.4.convs1.2.weight", "conv_post.bias", ...), fp32, PyTorch layout,
 * host or device pointer.  forward: mel (B, num_mels, T) fp32 -> audio (B, 1, T * prod(upsample_rates)) fp32, as Generator.forward.
 * flags: GTTS_FLAG_FP32 = fp32 activations and CUDA-core FFMA convs; default bf16 activations, tcgen05 convs, fp32 accumulation.
 * Options: "max_chunk" (utterances per workspace chunk, default 32), "workspace_mb", "use_graph", "force_ffma", "max_plans". */
typedef struct gtts_vocoder gtts_vocoder;
int gtts_vocoder_create(gtts_vocoder** out, int resblock, int n_ups, const int* upsample_rates, const int* upsample_kernel_sizes,
                        int upsample_initial_channel, int n_rb, const int* resblock_kernel_sizes, const int* resblock_dilation_sizes,
                        int n_dil, int num_mels, int device);
void gtts_vocoder_destroy(gtts_vocoder* v);
int gtts_vocoder_set_param(gtts_vocoder* v, const char* name, const float* data, size_t numel);
int gtts_vocoder_set_option(gtts_vocoder* v, const char* key, long long value);
int gtts_vocoder_hop(const gtts_vocoder* v);                     /* output samples per mel frame */
int gtts_vocoder_forward(gtts_vocoder* v, const float* mel, float* audio, int B, int T, int flags, void* stream);
/* same with HOST buffers: H2D, forward, D2H, synchronised */
int gtts_vocoder_forward_host(gtts_vocoder* v, const float* mel_host, float* audio_host, int B, int T, int flags);
/* per-launch CUDA-event times of one forward of (min(B, max_chunk), T) as a text table */
int gtts_vocoder_profile(gtts_vocoder* v, int B, int T, int flags, char* buf, size_t buflen, void* stream);
long gtts_vocoder_launches_last_call(const gtts_vocoder* v);
/* n >= 3: out[0] plans cached (descriptors + graph per (B, T); LRU, option "max_plans", default 8), [1] bytes of the workspace pool
 * all plans of the handle share, [2] peak live workspace of the most recent plan.  Calls on one handle are ordered across streams. */
int gtts_vocoder_cache_info(const gtts_vocoder* v, long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* GRADTTS_B200_H */
