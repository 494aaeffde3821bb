// Internal launcher interface between the decoder scheduler (decoder.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace gtts {

enum ActKind { ACT_F32 = 0, ACT_BF16 = 1 };

// ------------------------------------------------------------------------------------------------
// Generic convolution descriptor (implicit GEMM view: M = output-grid pixels, N = Cout, K = taps*Cin)
//
// Activations are NHWC.  Output-grid position (j, i) of phase p reads, for tap t, input pixel
// (j*stride + dy[p][t], i*stride + dx[p][t]) of each source (zero outside [0,Hin)x[0,Win)), multiplies
// with weight rows  wrow[p][t] + b*w_batch_rows + co  of the packed K-major weight matrix
// [rows][Cin0+Cin1], and writes output pixel (j*out_step + oy[p], i*out_step + ox[p]).
//   3x3 s1 : stride 1, 9 taps dy,dx in {-1,0,1}, 1 phase
//   3x3 s2 : stride 2, same taps                                   (Downsample, diffusion.py:30-36)
//   1x1    : 1 tap
//   convT  : 4 phases x 4 taps, out_step 2                         (Upsample, diffusion.py:21-27)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 24;   // taps per phase: 16 for the 4x4 stride-2 conv (Upsample dgrad), up to 17 for the position-packed k = 11 Conv1d (vocoder.cu)

struct ConvGeom {
    int B, Hin, Win;          // input spatial size (both sources)
    int Hg, Wg;               // output grid per phase
    int Hout, Wout;           // output tensor spatial size
    int Cin0, Cin1, Cout;
    int ntaps, nphase, stride, out_step;
    int8_t dy[4][kMaxTaps], dx[4][kMaxTaps];   // up to kMaxTaps taps per phase
    int wrow[4][kMaxTaps];
    int oy[4], ox[4];
    int w_batch_rows;         // 0, or Cout for per-sample weights
    // fp32 arithmetic on the tensor cores: the sources are [hi | mid | lo] bf16 planes of fp32 activations (3*Cin channels per
    // pixel, split_f32_planes) and the weight matrix has 6*Cin columns [wl wm wm wh wh wh] (split_pack_weights): the six partial
    // products hl, mm, hm, lh, mh, hh accumulate, smallest first, in the fp32 TMEM accumulator (the three products below 2^-24
    // relative are dropped, like fp32 rounding would)
    int split;
};

struct ConvEpilogue {
    const float* bias;        // [Cout] or null
    const void* residual;     // NHWC at output resolution, activation type, or null
    const float* mask;        // [B][Wout] or null (output multiplied by mask[b][w])
    void* out;                // NHWC activation type
    int out_f32;              // tcgen05 kernels: out / residual are fp32 NHWC instead of bf16 (the fp32 mode on tensor cores)
    // GroupNorm statistics of (acc + bias) over the whole sample (null partials = off)
    float* gn_partials;
    float* gn_stats;          // [B][8][2] mean, rstd
    unsigned int* gn_counters;
    float gn_eps;
    // Input transform fused into the operand path (CTA-pair halo kernel only; null in_stats = off): the conv reads the RAW
    // output of the previous conv and applies  (Mish(GroupNorm(raw)) + tbias) * mask  while the tile sits in shared memory
    // (reference: Block.forward + the time-embedding add of ResnetBlock, model/diffusion.py:52-58, 75-77).
    const float* in_stats;    // [B][8][2] mean, rstd of the previous conv
    const float* in_gamma;    // [Cin]
    const float* in_beta;     // [Cin]
    const float* in_tbias;    // [B or 1][Cin] (stride in_tb_bstride floats per sample; 0 = shared)
    int in_tb_bstride;
    const float* in_mask;     // [B][W]
    // GroupNorm apply fused into THIS conv's epilogue (CTA-pair halo kernel only; apply = 0: off).  The kernel keeps every
    // accumulator tile of a sample in TMEM until the sample's statistics are complete (per-sample grid barrier), then writes
    //     out = (Mish(GroupNorm(acc + bias)) [+ ap_tbias] [+ residual]) * mask
    // so the raw conv output never goes to HBM and the separate gn_apply pass disappears (reference: Block.forward and the two
    // adds of ResnetBlock.forward, model/diffusion.py:52-58, 75-78).  Needs gn_partials / gn_stats / gn_counters and mask.
    int apply;                // 1: out of TMEM (above); 2: asynchronous apply warps -- the raw bf16 tile goes to `out` as in a plain conv,
                              // eight extra warps write the finished activation to `ap_out` once the sample's statistics are complete
                              // (no TMEM limit, the MMA pipeline never waits; conv_tc_halo2.cu kApply 3/4)
    void* ap_out;
    const float* ap_gamma;    // [Cout]
    const float* ap_beta;     // [Cout]
    const float* ap_tbias;    // [B or 1][Cout] added after Mish (stride ap_tb_bstride floats per sample; null = off)
    int ap_tb_bstride;
    // Leaky-ReLU outputs of the 1-D conv stacks (HiFi-GAN generator, hifi-gan/models.py:38-45,104-113; per-tap tcgen05 kernel and
    // conv_ffma only): with v = acc + bias (+ residual),  out = act_out ? lrelu(v) : v  and, when out2 is set,  out2 = lrelu(v)
    // -- the next conv's input and the next residual come out of one epilogue.
    float act_slope;
    int act_out;
    void* out2;
};

// CUDA-core implicit GEMM (fp32 accumulate, FFMA).  Strict-fp32 path and debugging cross-check.
int conv_ffma(ActKind act, const ConvGeom& g, const void* src0, const void* src1, const void* weight,
              const ConvEpilogue& e, cudaStream_t stream);
size_t conv_ffma_partials_slots(const ConvGeom& g);

// tcgen05 / TMEM / TMA implicit GEMM, bf16 operands, fp32 accumulate.
struct TcConvPlan;   // opaque, owns the tensor maps for one (geometry, buffers) binding
// halo_mode: 0 = one TMA box per tap; 1/2 = one halo box per tile (18 x 16 / 18 x 10 pixels), the nine taps as
// shifted shared-memory descriptor views of it; falls back to 0 when the geometry is not eligible.
TcConvPlan* conv_tc_plan_create(const ConvGeom& g, const void* src0, const void* src1, const void* weight,
                                int weight_rows, const ConvEpilogue& e, int num_sms, int halo_mode);
bool conv_tc_halo_eligible(const ConvGeom& g);
size_t conv_tc_halo_partials_slots(const ConvGeom& g);
void conv_tc_plan_destroy(TcConvPlan* p);
void conv_tc_plan_set_debug(TcConvPlan* p, unsigned long long* dbg_out);   // per-CTA cycle counters (experiments)
int conv_tc_plan_grid(const TcConvPlan* p);
bool conv_tc_cta2_enabled();
bool conv_tc_convT_halo_eligible(const ConvGeom& g);
// true when a 3x3 stride-1 conv of this geometry can run with the GroupNorm-apply epilogue (ConvEpilogue::apply): CTA-pair halo
// kernel, and no CTA owns more tiles of one sample than there are TMEM accumulator buffers
bool conv_tc_apply_eligible(const ConvGeom& g, int num_sms);
bool conv_tc_apply_async_eligible(const ConvGeom& g, int num_sms);   // the asynchronous variant: any run length
size_t conv_tc_counter_words(int B);   // unsigned ints in ConvEpilogue::gn_counters for a batch of B samples
int microbench_issue(int N, int n_mma, int n_commit, int iters, int wait_each, int grid, unsigned long long* out_dev,
                     cudaStream_t stream);                                  // tcgen05 issue-path micro-benchmark
int conv_tc_launch(const TcConvPlan* p, cudaStream_t stream);
size_t conv_tc_partials_slots(const ConvGeom& g);

// 1-D conv weights (vocoder.cu, text_encoder.cu): PyTorch Conv1d (Cout, Cin, k) -- or ConvTranspose1d (Cin, Cout, k) when
// `transposed` -- to the K-major packed rows [tap][Cout_p] x [Cin_p] the implicit-GEMM kernels read (zero padding)
int pack_conv1d_weight(ActKind act, const float* src, void* dst, int Cout, int Cin, int k, int Cout_p, int Cin_p, bool transposed,
                       cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Point-wise / small kernels
// ------------------------------------------------------------------------------------------------
// level masks: m0 = mask (B,T), m1 = m0[::2], m2 = m1[::2]   (diffusion.py:185-196)
int build_level_masks(const float* mask, float* m0, float* m1, float* m2, int B, int T, cudaStream_t s);

// speaker plane s[b][80] = Linear(Mish(Linear(spk)))            (diffusion.py:139-141,176)
int spk_mlp(const float* spk, const float* w0t, const float* b0, const float* w2t, const float* b2, float* s_out,
            int B, int n_feats, bool strict, cudaStream_t s);

// time embedding + 12 per-block biases                           (diffusion.py:113-125,143-144,64-65)
struct TembWeights {
    const float* w0t; const float* b0;   // [64][256] transposed, [256]
    const float* w2t; const float* b2;   // [256][64] transposed, [64]
    const float* wbt; const float* bb;   // [64][1792] transposed (12 blocks concatenated), [1792]
};
// t: per-sample times (stride 1) or t_table indexed by *step (t_is_table): out tb[nb][1792]
int temb_bias(const TembWeights& w, const float* t, const int* step, int t_is_table, float pe_scale, float* tb,
              int nb, bool strict, cudaStream_t s);

// first block: 3x3 conv from the fp32 planes [mu, x, (s)] * mask -> raw NHWC(64) + GN stats
struct FirstConvArgs {
    const float* mu; const float* x; const float* splane;   // (B,80,T), (B,80,T), (B,80) or null
    const float* mask;                                      // (B,T)
    const float* w;                                         // [27|18][64]  (k = (ci*3+ky)*3+kx, transposed)
    const float* bias;                                      // [64]
    int B, H, W, cin;
    int use_mma;                                            // bf16 activations: 1 = tensor-core kernel (bf16 weights, split inputs)
    void* raw;                                              // NHWC act
    float* gn_partials; float* gn_stats; unsigned int* gn_counters; float gn_eps;
};
int first_conv(ActKind act, const FirstConvArgs& a, cudaStream_t s);
size_t first_conv_partials_slots(int H, int W);

// GroupNorm apply + Mish + mask (+ time bias) (+ residual)       (diffusion.py:53-58,76-78)
struct GnApplyArgs {
    const void* raw; const float* stats;                    // NHWC act, [B][8][2]
    const float* gamma; const float* beta;                  // [C]
    const float* mask;                                      // [B][W]
    const float* tbias; int tbias_bstride;                  // [.. + c] added after Mish (null = off)
    const void* residual;                                   // NHWC act added after Mish (null = off)
    // first-block residual computed inline: res = Wres[c][:] . [mu,x,s]*mask + bres[c]
    const float* fr_mu; const float* fr_x; const float* fr_s; const float* fr_w; const float* fr_b; int fr_cin;
    void* out;                                              // NHWC act
    int B, H, W, C;
};
int gn_apply(ActKind act, const GnApplyArgs& a, bool strict, cudaStream_t s);

// SDE extension: base pointer and per-step stride of the caller's noise tensor [n_timesteps][B][80][T], kept in DEVICE memory so
// that the captured Euler node does not bake either in (the same graph serves every call and every chunk offset).
struct NoiseSlot { const float* base; unsigned long long step_stride; };

// final Block's norm/act + final 1x1 conv + masks + Euler update  (diffusion.py:212-216, 259-267)
struct EulerArgs {
    const void* raw; const float* stats; const float* gamma; const float* beta;   // final_block conv output
    const float* wf; float bf;                                                    // final_conv (64 -> 1)
    const float* mask;                                                            // (B,T)
    const float* mu; float* xt;                                                   // (B,80,T) fp32, xt updated
    float* score_out;                                                             // (B,80,T) or null
    const float* beta_tab; const int* step; const float* h_ptr;                   // beta_t per step, step index, h
    const NoiseSlot* noise_slot; int sde;                                         // SDE extension (noise base + stride via slot)
    int update;                                                                   // 0: only write score
    int B, H, W;
};
int euler_step(ActKind act, const EulerArgs& a, bool strict, cudaStream_t s);
int advance_step(int* step, cudaStream_t s);
int init_xt(const float* z, const float* mask, float* xt, int B, int H, int W, cudaStream_t s);

// LinearAttention core                                            (diffusion.py:93-99)
struct AttnCtxArgs {
    const void* kv;               // NHWC act, 256 channels: k = [0,128), v = [128,256), head-major
    int B, n;                     // n = H*W positions
    float* partials;              // [B][4][chunks][1088]
    float* ctxn;                  // [B][4][32][32]  ctx[d][e] / l[d]
    int chunks, chunk_len;
    float* ml;                    // optional [B][4][64]: global max m[32] and sum l[32] of the key softmax (kept for the backward pass)
};
int attn_ctx(ActKind act, const AttnCtxArgs& a, bool strict, cudaStream_t s);      // per-chunk partials
int attn_merge(const AttnCtxArgs& a, bool strict, cudaStream_t s);                   // partials -> ctxn
void attn_ctx_plan(int n, int* chunks, int* chunk_len);
// fused k-projection + context for bf16 activations with C = 64 / 128 (attention.cu): reads x only
int attn_xk(const void* x, const void* wkv_bf16 /*[256][C]*/, float* partials, int B, int n, int C, int chunks,
            int chunk_len, cudaStream_t s);                  // writes attn_merge-format partials
// tcgen05 version (attention_tc.cu), C = 64 / 128
int attn_xk_tc(const void* x, const void* wkv_bf16, float* partials, int B, int n, int C, int chunks, int chunk_len, cudaStream_t s);
// per-sample folded weights M_b = g * Wout * blockdiag(ctxn^T) * Wq  -> [B*C][C] in weight type
int attn_fold(ActKind wkind, const float* ctxn, const float* wout /*[C][128]*/, const float* wq /*[128][C]*/,
              float g, void* mb_out, int B, int C, cudaStream_t s, int variant = -1 /* test hook: 0 / 1 force a kernel */);

// alignment stage around MAS (align.cu): log-prior (tts.py:143-149), durations and aligned means (tts.py:155,184-185)
int align_log_prior(const float* mu_x, const float* y, float* log_prior, int B, int C, int tx, int ty, cudaStream_t s);
int align_outputs(const float* attn, const float* mu_x, const float* x_mask, float* logw, float* mu_y, int B, int C, int tx,
                  int ty, cudaStream_t s);

// forward half of the training objective (loss.cu): model/diffusion.py:244-252 and :274-281, forward value only
size_t score_loss_workspace_bytes();
int forward_diffusion(const float* x0, const float* mask, const float* mu, const float* t, const float* z, float* xt, float* zm,
                      int B, int C, int T, float beta_min, float beta_max, cudaStream_t s);
int score_loss(const float* est, const float* zm, const float* mask, const float* t, void* ws, size_t ws_bytes, float* loss, int B,
               int C, int T, float beta_min, float beta_max, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Backward w.r.t. the network input (backward.cu): point-wise / reduction kernels and dgrad weight packing
// ------------------------------------------------------------------------------------------------
struct GnBwdArgs {
    const void* raw; const float* stats;        // saved conv output (NHWC act) and its [B][8][2] mean, rstd
    const float* gamma; const float* beta;      // [C]
    const float* mask;                          // [B][W]
    const void* gy;                             // gradient w.r.t. the Block output (NHWC act)
    void* graw;                                 // out: gradient w.r.t. the conv output (NHWC act)
    float* partials;                            // scratch [B][gn_bwd_blocks][16]
    float* param_partials;                      // optional scratch [B][gn_bwd_blocks][2C]: per-channel d gamma | d beta partial sums,
                                                // produced by the same pass over (raw, gy) as the group sums (training plans)
    int B, H, W, C;
};
// fixed-order sum of the param_partials rows of gn_bwd into d gamma [C], d beta [C]
int gn_param_reduce(const float* param_partials, float* dgamma, float* dbeta, int B, int H, int W, int C, cudaStream_t s);
int gn_bwd_blocks(int H, int W);
int gn_bwd(ActKind act, const GnBwdArgs& a, cudaStream_t s);
int final_bwd(ActKind act, const float* v, const float* wf, const float* mask, void* ghf, int B, int H, int W, cudaStream_t s);
// gradient planes of the network input: gx (x channel) always; gmu (mu channel) and gs (speaker channel, per pixel; the caller sums it
// over frames) when non-null
int first_bwd(ActKind act, const void* graw1, const void* gres, const float* w1t, const float* wres, const float* mask, float* gx,
              float* gmu, float* gs, int B, int H, int W, int cin, cudaStream_t s);
int mask_mul(ActKind act, const void* in, const float* mask, void* out, int B, int H, int W, int C, cudaStream_t s);
int add_tensors(ActKind act, const void* a, const void* b, void* out, size_t numel, cudaStream_t s);
// g_ctx = sum_n q go^T per (sample, head) (partials [B][4][chunks][1024]) and sdot[b][h][d] = sum_e g_ctx[d][e] ctxn[d][e]
int attn_outer(ActKind act, const void* q, const void* go, float* partials, const float* ctxn, float* gctx, float* sdot, int B, int n,
               int chunks, int chunk_len, cudaStream_t s);
// q / ao optional: ao[b][n][128] = attention output before to_out (for its weight gradient)
int attn_pos_bwd(ActKind act, const void* kv, const void* go, const float* ctxn, const float* gctx, const float* ml, const float* sdot,
                 void* gq, void* gkv, int B, int n, const void* q, void* ao, cudaStream_t s);
int pack_dgrad3(ActKind wkind, const float* w_oihw, void* out, int Cout, int Cin, int ci_off, int Cn, cudaStream_t s);
int pack_t1(ActKind wkind, const float* w_oi, void* out, int Cout, int Cin, int ci_off, int Cn, float scale, cudaStream_t s);
int pack_down_dgrad(ActKind wkind, const float* w_oihw, void* out, int C, cudaStream_t s);
int pack_up_dgrad(ActKind wkind, const float* w_iohw, void* out, int C, cudaStream_t s);

// fp32 -> three bf16 planes with x = hi + mid + lo exactly (24 = 8 + 8 + 8 mantissa bits): in [P][C] fp32 -> out [P][3C] bf16
int split_f32_planes(const float* in, void* out, size_t npix, int C, cudaStream_t s);
// packed fp32 weight rows [rows][K] -> bf16 [rows][6K] = [wl | wm | wm | wh | wh | wh]
int split_pack_weights(const float* w, void* out, size_t rows, int K, cudaStream_t s);

// ---- parameter gradients (backward_params.cu): reductions of the activation gradients against the saved forward tensors
size_t wgrad_partial_floats(const ConvGeom& g, int* slices_out);
// the same on tcgen05 (wgrad_tc.cu): stride-1 convs, bf16 activations; operands read MN-major straight from the NHWC tensors
struct WgradTcPlan;   // opaque, owns the tensor maps of one (geometry, buffers) binding
bool wgrad_tc_eligible(const ConvGeom& g);
size_t wgrad_tc_partial_floats(const ConvGeom& g, int num_sms, int* slices_out);
WgradTcPlan* wgrad_tc_plan_create(const ConvGeom& g, const void* gout, const void* x0, const void* x1, float* partial, int num_sms);
void wgrad_tc_plan_destroy(WgradTcPlan* p);
int wgrad_tc_slices(const WgradTcPlan* p);
int wgrad_tc_launch(const WgradTcPlan* p, cudaStream_t s);
// fixed-order sum of the partial slices, scattered into the PyTorch weight layout (kind as conv_wgrad)
int wgrad_reduce(const float* partial, float* dst, int slices, int n_pt, int Cout, int Cin, int kind, float scale, int accumulate,
                 cudaStream_t s);
// dW of any conv geometry; dst in the PyTorch layout: kind 0 = Conv2d (Cout, Cin, kh, kw) (also 1x1), 1 = ConvTranspose2d (Cin, Cout, 4, 4),
// 2 = packed rows [n_pt * Cout][Cin] as is
int conv_wgrad(ActKind act, const ConvGeom& g, const void* gout, const void* x0, const void* x1, float* partial, float* dst, int kind,
               float scale, int accumulate, cudaStream_t s);
int col_sums(ActKind act, const void* gsrc, float* partial, float* dst, long npix, int C, float scale, int accumulate, cudaStream_t s);
int col_sums_per_sample(ActKind act, const void* gsrc, const float* mask, float* partial, float* dst, int B, int H, int W, int C, int dst_ld,
                        cudaStream_t s);
int attn_out_grads(const float* A, const float* sv, const float* wout, const float* bout, float g, float* dwout, float* dbout, float* dg,
                   int C, cudaStream_t s);
int accumulate_floats(float* dst, const float* src, size_t n, int accumulate, cudaStream_t s);
int final_param_grad(ActKind act, const void* rawf, const float* stats, const float* gamma, const float* beta, const float* v,
                     const float* mask, float* partial, float* dwf_dbf, int B, int H, int W, int accumulate, cudaStream_t s);
int first_param_blocks(int H, int W);
int first_param_grad(ActKind act, const void* graw1, const void* gres, const float* mu, const float* x, const float* splane,
                     const float* mask, float* partial, float* dst, int B, int H, int W, int cin, int accumulate, cudaStream_t s);

// weight packing helpers (device side, fp32 source in PyTorch layout)
int pack_conv_weight(ActKind wkind, const float* w_oihw, void* packed, int Cout, int Cin, int kh, int kw,
                     cudaStream_t s);                       // -> [(ky*kw+kx)*Cout + co][ci]
int pack_convT_weight(ActKind wkind, const float* w_iohw, void* packed, int C, cudaStream_t s);   // 4x4 s2 p1
int transpose_2d(const float* src, float* dst, int rows, int cols, cudaStream_t s);               // dst[c][r]

// MAS (mas.cu)
size_t mas_bits_workspace_bytes(int B, int tx, int ty);
int mas_forward_f32(const float* value, const float* mask, const int* t_xs, const int* t_ys, float* path, int B,
                    int tx, int ty, float max_neg, uint32_t* bits_ws, size_t bits_ws_bytes, int* status,
                    cudaStream_t stream);
int mas_forward_i32(const float* value, const float* mask, const int* t_xs, const int* t_ys, int32_t* path, int B,
                    int tx, int ty, float max_neg, uint32_t* bits_ws, size_t bits_ws_bytes, int* status,
                    cudaStream_t stream);

}  // namespace gtts
