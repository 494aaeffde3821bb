// extern "C" surface declared in include/gradtts_b200.h.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/gradtts_b200.h"
#include "common.cuh"
#include "decoder_api.h"
#include "ops.h"
#include "text_encoder_api.h"
#include "vocoder_api.h"
#include <vector>

namespace gtts {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static thread_local int g_pdl_override = 0;
void pdl_set_override(int on) { g_pdl_override = on; }
bool pdl_enabled() {
    static int v = -1;
    // measured on B200: -2 % on the throughput workloads, so off unless GTTS_PDL=1 -- except while a small-batch plan is being
    // captured (pdl_set_override, decoder.cu): there most SMs are idle, the successor's CTAs start on them while the predecessor
    // still runs, and its prologue and resident-weight loads come off the critical path
    if (v < 0) { const char* e = getenv("GTTS_PDL"); v = e ? (atoi(e) != 0) : 0; }
    return v != 0 || g_pdl_override != 0;
}
}  // namespace gtts

using namespace gtts;

struct gtts_decoder {
    Decoder* impl;
    // staging for the host-buffer entry point
    float *z = nullptr, *mask = nullptr, *mu = nullptr, *spk = nullptr, *out = nullptr;
    size_t cap_plane = 0, cap_mask = 0, cap_spk = 0;
    cudaStream_t stream = nullptr;
};

struct gtts_encoder {
    TextEncoder* impl;
};

struct gtts_vocoder {
    Vocoder* impl;
    float *mel = nullptr, *audio = nullptr;          // staging for the host-buffer entry point
    size_t cap_mel = 0, cap_audio = 0;
    cudaStream_t stream = nullptr;
};

extern "C" {

int gtts_version(void) { return 100; }

const char* gtts_last_error(void) { return g_last_error.c_str(); }

int gtts_sm100_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int c = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++c;
    }
    return c;
}

// ------------------------------------------------------------------------------------------------ MAS
size_t gtts_mas_workspace_bytes(int B, int t_x, int t_y) { return mas_bits_workspace_bytes(B, t_x, t_y); }

int gtts_mas_maximum_path(const float* value, const float* mask, float* path, int B, int t_x, int t_y, void* bits_ws,
                          size_t bits_ws_bytes, int32_t* status, void* stream) {
    GTTS_REQUIRE(value && mask && path && status, "maximum_path: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    GTTS_CHECK_CUDA(cudaMemsetAsync(path, 0, (size_t)B * t_x * t_y * sizeof(float), s));
    GTTS_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    return mas_forward_f32(value, mask, nullptr, nullptr, path, B, t_x, t_y, -1e9f, (uint32_t*)bits_ws, bits_ws_bytes,
                           status, s);
}

int gtts_mas_maximum_path_c(int32_t* paths, const float* values, const int32_t* t_xs, const int32_t* t_ys, int B,
                            int t_x, int t_y, float max_neg_val, void* bits_ws, size_t bits_ws_bytes,
                            int32_t* status, void* stream) {
    GTTS_REQUIRE(paths && values && t_xs && t_ys && status, "maximum_path_c: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    GTTS_CHECK_CUDA(cudaMemsetAsync(paths, 0, (size_t)B * t_x * t_y * sizeof(int32_t), s));
    GTTS_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    return mas_forward_i32(values, nullptr, t_xs, t_ys, paths, B, t_x, t_y, max_neg_val, (uint32_t*)bits_ws,
                           bits_ws_bytes, status, s);
}

int gtts_mas_maximum_path_host(const float* value_host, const float* mask_host, float* path_host, int B, int t_x,
                               int t_y, int32_t* status_host, int device) {
    GTTS_REQUIRE(value_host && mask_host && path_host && status_host, "maximum_path_host: null pointer");
    GTTS_CHECK_CUDA(cudaSetDevice(device));
    const size_t n = (size_t)B * t_x * t_y;
    float *dv = nullptr, *dm = nullptr, *dp = nullptr;
    int32_t* ds = nullptr;
    void* ws = nullptr;
    const size_t wsb = mas_bits_workspace_bytes(B, t_x, t_y);
    cudaStream_t s;
    GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(dv); cudaFree(dm); cudaFree(dp); cudaFree(ds); cudaFree(ws);
        cudaStreamDestroy(s);
    };
#define GTTS_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); cleanup(); return 1; } } while (0)
    GTTS_TRY(cudaMalloc(&dv, n * 4));
    GTTS_TRY(cudaMalloc(&dm, n * 4));
    GTTS_TRY(cudaMalloc(&dp, n * 4));
    GTTS_TRY(cudaMalloc(&ds, 4));
    if (wsb) GTTS_TRY(cudaMalloc(&ws, wsb));
    GTTS_TRY(cudaMemcpyAsync(dv, value_host, n * 4, cudaMemcpyHostToDevice, s));
    GTTS_TRY(cudaMemcpyAsync(dm, mask_host, n * 4, cudaMemcpyHostToDevice, s));
    rc = gtts_mas_maximum_path(dv, dm, dp, B, t_x, t_y, ws, wsb, ds, s);
    if (rc) { cleanup(); return rc; }
    GTTS_TRY(cudaMemcpyAsync(path_host, dp, n * 4, cudaMemcpyDeviceToHost, s));
    GTTS_TRY(cudaMemcpyAsync(status_host, ds, 4, cudaMemcpyDeviceToHost, s));
    GTTS_TRY(cudaStreamSynchronize(s));
#undef GTTS_TRY
    cleanup();
    return 0;
}

// ------------------------------------------------------------------------------------------------ decoder
int gtts_decoder_create(gtts_decoder** out, int n_spks, int n_feats, int dim, double beta_min, double beta_max,
                        double pe_scale, int device) {
    GTTS_REQUIRE(out != nullptr, "null out pointer");
    Decoder* d = decoder_new(n_spks, n_feats, dim, beta_min, beta_max, pe_scale, device);
    if (!d) return 1;
    gtts_decoder* h = new gtts_decoder();
    h->impl = d;
    *out = h;
    return 0;
}

void gtts_decoder_destroy(gtts_decoder* h) {
    if (!h) return;
    cudaSetDevice(decoder_device(h->impl));
    cudaFree(h->z); cudaFree(h->mask); cudaFree(h->mu); cudaFree(h->spk); cudaFree(h->out);
    if (h->stream) cudaStreamDestroy(h->stream);
    decoder_delete(h->impl);
    delete h;
}

int gtts_decoder_set_param(gtts_decoder* h, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_set_param(h->impl, name, data, numel);
}

int gtts_decoder_set_option(gtts_decoder* h, const char* key, int value) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_set_option(h->impl, key, value);
}

int gtts_decoder_reverse_diffusion(gtts_decoder* h, const float* z, const float* mask, const float* mu,
                                   const float* spk, float* out, int B, int T, int n_timesteps, int flags,
                                   const float* noise, void* stream) {
    GTTS_REQUIRE(h && z && mask && mu && out, "reverse_diffusion: null pointer");
    return decoder_reverse_diffusion(h->impl, z, mask, mu, spk, out, B, T, n_timesteps, flags, noise,
                                     (cudaStream_t)stream);
}

int gtts_decoder_estimator(gtts_decoder* h, const float* x, const float* mask, const float* mu, const float* t,
                           const float* spk, float* out, int B, int T, int flags, void* stream) {
    GTTS_REQUIRE(h && x && mask && mu && t && out, "estimator: null pointer");
    return decoder_estimator(h->impl, x, mask, mu, t, spk, out, B, T, flags, (cudaStream_t)stream);
}

int gtts_decoder_estimator_vjp(gtts_decoder* h, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                               const float* v, float* out_score, float* out_gx, int B, int T, int flags, void* stream) {
    GTTS_REQUIRE(h && x && mask && mu && t && v && out_gx, "estimator_vjp: null pointer");
    return decoder_estimator_vjp(h->impl, x, mask, mu, t, spk, v, out_score, out_gx, B, T, flags, (cudaStream_t)stream);
}

int gtts_decoder_estimator_backward(gtts_decoder* h, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                                    const float* v, float* out_score, float* out_gx, float* out_gmu, float* out_gs_pix, float* out_gtb,
                                    int B, int T, int flags, void* stream) {
    GTTS_REQUIRE(h && x && mask && mu && t && v && out_gx, "estimator_backward: null pointer");
    return decoder_estimator_backward(h->impl, x, mask, mu, t, spk, v, out_score, out_gx, out_gmu, out_gs_pix, out_gtb, B, T, flags,
                                      (cudaStream_t)stream);
}

int gtts_decoder_get_param_grad(gtts_decoder* h, const char* name, float* dst, size_t numel, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_get_param_grad(h->impl, name, dst, numel, (cudaStream_t)stream);
}

int gtts_decoder_get_param_grads_flat(gtts_decoder* h, float* dst, size_t numel, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_get_param_grads_flat(h->impl, dst, numel, (cudaStream_t)stream);
}

int gtts_decoder_param_grad_slot(const gtts_decoder* h, const char* name, size_t* offset, size_t* numel) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_param_grad_slot(h->impl, name, offset, numel);
}

int gtts_decoder_reverse_diffusion_host(gtts_decoder* h, const float* z_host, const float* mask_host,
                                        const float* mu_host, const float* spk_host, float* out_host, int B, int T,
                                        int n_timesteps, int flags) {
    GTTS_REQUIRE(h && z_host && mask_host && mu_host && out_host, "reverse_diffusion_host: null pointer");
    GTTS_REQUIRE((flags & GTTS_FLAG_SDE) == 0, "reverse_diffusion_host: SDE noise is not supported on this entry");
    GTTS_CHECK_CUDA(cudaSetDevice(decoder_device(h->impl)));
    const size_t plane = (size_t)B * 80 * T, nmask = (size_t)B * T, nspk = (size_t)B * 64;
    if (!h->stream) GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (plane > h->cap_plane) {
        cudaFree(h->z); cudaFree(h->mu); cudaFree(h->out);
        h->z = h->mu = h->out = nullptr; h->cap_plane = 0;
        GTTS_CHECK_CUDA(cudaMalloc(&h->z, plane * 4));
        GTTS_CHECK_CUDA(cudaMalloc(&h->mu, plane * 4));
        GTTS_CHECK_CUDA(cudaMalloc(&h->out, plane * 4));
        h->cap_plane = plane;
    }
    if (nmask > h->cap_mask) {
        cudaFree(h->mask); h->mask = nullptr; h->cap_mask = 0;
        GTTS_CHECK_CUDA(cudaMalloc(&h->mask, nmask * 4));
        h->cap_mask = nmask;
    }
    if (spk_host && nspk > h->cap_spk) {
        cudaFree(h->spk); h->spk = nullptr; h->cap_spk = 0;
        GTTS_CHECK_CUDA(cudaMalloc(&h->spk, nspk * 4));
        h->cap_spk = nspk;
    }
    cudaStream_t s = h->stream;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(h->z, z_host, plane * 4, cudaMemcpyHostToDevice, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(h->mu, mu_host, plane * 4, cudaMemcpyHostToDevice, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(h->mask, mask_host, nmask * 4, cudaMemcpyHostToDevice, s));
    if (spk_host) GTTS_CHECK_CUDA(cudaMemcpyAsync(h->spk, spk_host, nspk * 4, cudaMemcpyHostToDevice, s));
    if (int rc = decoder_reverse_diffusion(h->impl, h->z, h->mask, h->mu, spk_host ? h->spk : nullptr, h->out, B, T,
                                           n_timesteps, flags, nullptr, s))
        return rc;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(out_host, h->out, plane * 4, cudaMemcpyDeviceToHost, s));
    GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int gtts_decoder_profile_step(gtts_decoder* h, int B, int T, int flags, int reps, char* buf, size_t buflen, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_profile_step(h->impl, B, T, flags, reps, buf, buflen, (cudaStream_t)stream);
}

long gtts_decoder_launches_last_call(const gtts_decoder* h) { return h ? decoder_launches_last_call(h->impl) : 0; }

int gtts_decoder_cache_info(const gtts_decoder* h, long long* out, int n) {
    GTTS_REQUIRE(h != nullptr, "null decoder handle");
    return decoder_cache_info(h->impl, out, n);
}

// ------------------------------------------------------------------------------------------------ alignment stage
int gtts_align_log_prior(const float* mu_x, const float* y, float* log_prior, int B, int n_feats, int t_x, int t_y, void* stream) {
    GTTS_REQUIRE(mu_x && y && log_prior, "null argument");
    return align_log_prior(mu_x, y, log_prior, B, n_feats, t_x, t_y, (cudaStream_t)stream);
}

int gtts_align_outputs(const float* attn, const float* mu_x, const float* x_mask, float* logw, float* mu_y, int B, int n_feats,
                       int t_x, int t_y, void* stream) {
    GTTS_REQUIRE(attn && mu_x && (logw == nullptr || x_mask != nullptr), "null argument");
    return align_outputs(attn, mu_x, x_mask, logw, mu_y, B, n_feats, t_x, t_y, (cudaStream_t)stream);
}

size_t gtts_score_loss_workspace_bytes(void) { return score_loss_workspace_bytes(); }

int gtts_forward_diffusion(const float* x0, const float* mask, const float* mu, const float* t, const float* noise, float* xt,
                           float* z_masked, int B, int n_feats, int T, double beta_min, double beta_max, void* stream) {
    GTTS_REQUIRE(x0 && mask && mu && t && noise && xt && z_masked, "null argument");
    return forward_diffusion(x0, mask, mu, t, noise, xt, z_masked, B, n_feats, T, (float)beta_min, (float)beta_max,
                             (cudaStream_t)stream);
}

int gtts_score_loss(const float* noise_estimation, const float* z_masked, const float* mask, const float* t, void* ws,
                    size_t ws_bytes, float* loss, int B, int n_feats, int T, double beta_min, double beta_max, void* stream) {
    GTTS_REQUIRE(noise_estimation && z_masked && mask && t && loss, "null argument");
    return score_loss(noise_estimation, z_masked, mask, t, ws, ws_bytes, loss, B, n_feats, T, (float)beta_min, (float)beta_max,
                      (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ test hooks
int gtts_test_attn_xk(const void* x_bf16, const void* wkv_bf16, float* partials, int B, int n, int C, int chunks, int chunk_len,
                      int use_tc, void* stream) {
    GTTS_REQUIRE(x_bf16 && wkv_bf16 && partials, "null argument");
    if (use_tc) {
        return attn_xk_tc(x_bf16, wkv_bf16, partials, B, n, C, chunks, chunk_len, (cudaStream_t)stream);
    }
    setenv("GTTS_ATTN_TC", "0", 1);
    int rc = attn_xk(x_bf16, wkv_bf16, partials, B, n, C, chunks, chunk_len, (cudaStream_t)stream);
    unsetenv("GTTS_ATTN_TC");
    return rc;
}

int gtts_test_attn_fold(const float* ctxn, const float* wout, const float* wq, float g, void* m_out, int B, int C, int out_bf16,
                        int variant, void* stream) {
    GTTS_REQUIRE(ctxn && wout && wq && m_out && B >= 1, "null argument");
    GTTS_REQUIRE(variant >= -1 && variant <= 1, "gtts_test_attn_fold: variant must be -1, 0 or 1");
    return attn_fold(out_bf16 ? ACT_BF16 : ACT_F32, ctxn, wout, wq, g, m_out, B, C, (cudaStream_t)stream, variant);
}

int gtts_test_issue_microbench(int N, int n_mma, int n_commit, int iters, int wait_each, int grid, double* issue_cycles,
                               double* total_cycles) {
    GTTS_REQUIRE((N == 64 || N == 128 || N == 256) && grid >= 1 && grid <= 1024 && iters >= 1,
                 "gtts_test_issue_microbench: bad arguments");
    unsigned long long* d = nullptr;
    GTTS_CHECK_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * 2 * grid));
    int rc = microbench_issue(N, n_mma, n_commit, iters, wait_each, grid, d, nullptr);
    if (rc == 0) rc = microbench_issue(N, n_mma, n_commit, iters, wait_each, grid, d, nullptr);   // second run is the warm one
    std::vector<unsigned long long> h(2 * grid);
    cudaError_t e = cudaMemcpy(h.data(), d, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc) return rc;
    GTTS_CHECK_CUDA(e);
    double a = 0, b = 0;
    for (int i = 0; i < grid; ++i) { a += (double)h[2 * i]; b += (double)h[2 * i + 1]; }
    *issue_cycles = a / grid;
    *total_cycles = b / grid;
    return 0;
}

// 3x3 stride-1 conv with the GroupNorm-apply epilogue (conv_tc_halo2.cu, kApply): out = (Mish(GN(conv + bias)) [+ tbias] [+ residual]) * mask
int gtts_test_conv_apply(int B, int H, int W, int Cin0, int Cin1, int Cout, const void* src0, const void* src1, const float* weight_pt,
                         const float* bias, const float* gamma, const float* beta, const float* tbias, int tb_bstride,
                         const void* residual, const float* mask, void* out, float* gn_stats, int reps, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    GTTS_REQUIRE(src0 && weight_pt && gamma && beta && mask && out && gn_stats, "null argument");
    const int Cin = Cin0 + Cin1;
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = H; g.Win = W; g.Hg = H; g.Wg = W; g.Hout = H; g.Wout = W;
    g.Cin0 = Cin0; g.Cin1 = Cin1; g.Cout = Cout; g.ntaps = 9; g.nphase = 1; g.stride = 1; g.out_step = 1;
    for (int t = 0; t < 9; ++t) { g.dy[0][t] = (int8_t)(t / 3 - 1); g.dx[0][t] = (int8_t)(t % 3 - 1); g.wrow[0][t] = t * Cout; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool async_apply = reps < 0;                   // negative reps: the asynchronous apply-warp variant (raw tile to a scratch tensor)
    if (reps < 0) reps = -reps;
    GTTS_REQUIRE(async_apply ? conv_tc_apply_async_eligible(g, sms) : conv_tc_apply_eligible(g, sms),
                 "gtts_test_conv_apply: geometry not eligible for the apply epilogue");
    void* raw_scratch = nullptr;
    if (async_apply) GTTS_CHECK_CUDA(cudaMalloc(&raw_scratch, (size_t)B * H * W * Cout * 2));
    void* wpk = nullptr;
    float* partials = nullptr;
    unsigned int* counters = nullptr;
    GTTS_CHECK_CUDA(cudaMalloc(&wpk, (size_t)9 * Cout * Cin * 2));
    if (int rc = pack_conv_weight(ACT_BF16, weight_pt, wpk, Cout, Cin, 3, 3, s)) return rc;
    GTTS_CHECK_CUDA(cudaMalloc(&partials, (size_t)B * conv_tc_halo_partials_slots(g) * 16 * 4));
    GTTS_CHECK_CUDA(cudaMalloc(&counters, conv_tc_counter_words(B) * 4));
    GTTS_CHECK_CUDA(cudaMemsetAsync(counters, 0, conv_tc_counter_words(B) * 4, s));
    ConvEpilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.residual = residual; e.mask = mask; e.out = out;
    e.gn_partials = partials; e.gn_stats = gn_stats; e.gn_counters = counters; e.gn_eps = 1e-5f;
    e.apply = 1; e.ap_gamma = gamma; e.ap_beta = beta; e.ap_tbias = tbias; e.ap_tb_bstride = tb_bstride;
    if (async_apply) { e.apply = 2; e.ap_out = out; e.out = raw_scratch; }
    int rc = 0;
    TcConvPlan* tp = conv_tc_plan_create(g, src0, src1, wpk, 9 * Cout, e, sms, 2);
    if (!tp) rc = 1;
    else {
        rc = conv_tc_launch(tp, s);                       // the counters return to zero by themselves: launch again and again
        if (reps > 1 && rc == 0) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, s);
            for (int i = 1; i < reps && rc == 0; ++i) rc = conv_tc_launch(tp, s);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "[gtts_test_conv_apply] B=%d H=%d W=%d Cin=%d Cout=%d grid=%d: %.1f us/launch\n", B, H, W, Cin, Cout,
                    conv_tc_plan_grid(tp), 1e3f * ms / (reps - 1));
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        cudaStreamSynchronize(s);
        conv_tc_plan_destroy(tp);
    }
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(wpk); cudaFree(partials); cudaFree(counters); cudaFree(raw_scratch);
    if (rc) return rc;
    GTTS_CHECK_CUDA(ce);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int gtts_test_conv(int impl, int act, int kind, int B, int H, int W, int Cin0, int Cin1, int Cout, const void* src0,
                   const void* src1, const float* weight_pt, const float* bias, const void* residual,
                   const float* mask, void* out, float* gn_stats, int per_sample_weights, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const ActKind ak = act ? ACT_BF16 : ACT_F32;
    const bool split = impl == 4;                              // fp32 tensors through the six-term bf16 split on the tensor cores
    GTTS_REQUIRE(impl == 0 || split || ak == ACT_BF16, "tcgen05 conv needs bf16 activations");
    GTTS_REQUIRE(!split || (ak == ACT_F32 && !per_sample_weights), "split conv: fp32 activations, shared weights");
    const int halo_mode = split ? 2 : (impl >= 2 ? impl - 1 : 0);   // impl 2 -> halo box 18x16, 3 -> halo box 18x10
    const int Cin = Cin0 + Cin1;
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = H; g.Win = W; g.Hg = H; g.Wg = W; g.Hout = H; g.Wout = W;
    g.Cin0 = Cin0; g.Cin1 = Cin1; g.Cout = Cout; g.ntaps = 1; g.nphase = 1; g.stride = 1; g.out_step = 1;
    size_t wrows = Cout;
    const size_t es = ak == ACT_F32 ? 4 : 2;
    void* wpk = nullptr;
    if (kind == 0 || kind == 2) {
        g.ntaps = 9;
        for (int t = 0; t < 9; ++t) { g.dy[0][t] = (int8_t)(t / 3 - 1); g.dx[0][t] = (int8_t)(t % 3 - 1); g.wrow[0][t] = t * Cout; }
        if (kind == 2) { g.stride = 2; g.Hg = H / 2; g.Wg = W / 2; g.Hout = H / 2; g.Wout = W / 2; }
        wrows = (size_t)9 * Cout;
        GTTS_CHECK_CUDA(cudaMalloc(&wpk, wrows * Cin * es));
        if (int rc = pack_conv_weight(ak, weight_pt, wpk, Cout, Cin, 3, 3, s)) return rc;
    } else if (kind == 1) {
        if (per_sample_weights) { g.w_batch_rows = Cout; wrows = (size_t)B * Cout; }
        GTTS_CHECK_CUDA(cudaMalloc(&wpk, wrows * Cin * es));
        // (rows, Cin, 1, 1) is already row-major [rows][Cin]
        if (int rc = pack_conv_weight(ak, weight_pt, wpk, (int)wrows, Cin, 1, 1, s)) return rc;
    } else if (kind == 3) {
        GTTS_REQUIRE(Cin1 == 0 && Cin0 == Cout, "convT test: Cin must equal Cout");
        g.ntaps = 4; g.nphase = 4; g.out_step = 2; g.Hout = 2 * H; g.Wout = 2 * W;
        const int dd[2][2] = {{0, -1}, {0, 1}};
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                const int ph = py * 2 + px;
                g.oy[ph] = py; g.ox[ph] = px;
                for (int ty = 0; ty < 2; ++ty)
                    for (int tx = 0; tx < 2; ++tx) {
                        const int t = ty * 2 + tx;
                        g.dy[ph][t] = (int8_t)dd[py][ty]; g.dx[ph][t] = (int8_t)dd[px][tx];
                        g.wrow[ph][t] = (ph * 4 + t) * Cout;
                    }
            }
        wrows = (size_t)16 * Cout;
        GTTS_CHECK_CUDA(cudaMalloc(&wpk, wrows * Cin * es));
        if (int rc = pack_convT_weight(ak, weight_pt, wpk, Cout, s)) return rc;
    } else {
        set_error("gtts_test_conv: unknown kind");
        return 2;
    }
    ConvEpilogue e;
    memset(&e, 0, sizeof(e));
    e.bias = bias; e.residual = residual; e.mask = mask; e.out = out;
    float* partials = nullptr;
    unsigned int* counters = nullptr;
    if (gn_stats) {
        size_t slots = impl >= 1 ? conv_tc_partials_slots(g) : conv_ffma_partials_slots(g);
        if (halo_mode && conv_tc_halo_eligible(g)) slots = conv_tc_halo_partials_slots(g);
        GTTS_CHECK_CUDA(cudaMalloc(&partials, (size_t)B * slots * 16 * 4));
        GTTS_CHECK_CUDA(cudaMalloc(&counters, conv_tc_counter_words(B) * 4));
        GTTS_CHECK_CUDA(cudaMemsetAsync(counters, 0, conv_tc_counter_words(B) * 4, s));
        e.gn_partials = partials; e.gn_stats = gn_stats; e.gn_counters = counters; e.gn_eps = 1e-5f;
    }
    int rc = 0;
    void *p0 = nullptr, *p1 = nullptr, *w3 = nullptr;
    if (split) {
        const size_t npix = (size_t)B * H * W;
        GTTS_CHECK_CUDA(cudaMalloc(&p0, npix * 3 * Cin0 * 2));
        if (Cin1) GTTS_CHECK_CUDA(cudaMalloc(&p1, npix * 3 * Cin1 * 2));
        GTTS_CHECK_CUDA(cudaMalloc(&w3, wrows * 6 * Cin * 2));
        if (int r = split_f32_planes((const float*)src0, p0, npix, Cin0, s)) return r;
        if (Cin1) if (int r = split_f32_planes((const float*)src1, p1, npix, Cin1, s)) return r;
        if (int r = split_pack_weights((const float*)wpk, w3, wrows, Cin, s)) return r;
        g.split = 1;
        e.out_f32 = 1;
    }
    if (impl >= 1) {
        GTTS_REQUIRE(split || !halo_mode || (conv_tc_halo_eligible(g) && !residual && !mask) || (halo_mode == 2 && conv_tc_convT_halo_eligible(g) && mask && !residual),
                     "halo test: geometry not eligible");
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        TcConvPlan* tp = split ? conv_tc_plan_create(g, p0, p1, w3, (int)wrows, e, sms, halo_mode)
                               : conv_tc_plan_create(g, src0, src1, wpk, (int)wrows, e, sms, halo_mode);
        if (!tp) rc = 1;
        else {
            rc = conv_tc_launch(tp, s);
            cudaStreamSynchronize(s);
            if (getenv("GTTS_CONV_TIMING")) {
                const int L = 4, G = conv_tc_plan_grid(tp);                       // L back-to-back instrumented launches
                unsigned long long* dbg = nullptr;
                cudaMalloc(&dbg, (size_t)L * 256 * 32 * 8);
                cudaMemset(dbg, 0, (size_t)L * 256 * 32 * 8);
                cudaStreamSynchronize(s);
                for (int l = 0; l < L && rc == 0; ++l) {
                    conv_tc_plan_set_debug(tp, dbg + (size_t)l * 256 * 32);
                    rc = conv_tc_launch(tp, s);
                }
                cudaStreamSynchronize(s);
                std::vector<unsigned long long> hv((size_t)L * 256 * 32);
                cudaMemcpy(hv.data(), dbg, hv.size() * 8, cudaMemcpyDeviceToHost);
                unsigned long long base = ~0ull;
                for (int c = 0; c < G; ++c) if (hv[c * 32 + 7] < base) base = hv[c * 32 + 7];
                for (int l = 0; l < L; ++l) {
                    unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
                    for (int c = 0; c < G; ++c) {
                        const unsigned long long* o = hv.data() + ((size_t)l * 256 + c) * 32;
                        if (o[7] < s0) s0 = o[7]; if (o[7] > s1) s1 = o[7];
                        if (o[15] < e0) e0 = o[15]; if (o[15] > e1) e1 = o[15];
                    }
                    fprintf(stderr, "[conv timing] launch %d: CTA start %.1f..%.1f us, CTA end %.1f..%.1f us (globaltimer, rel.)\n", l,
                            (s0 - base) * 1e-3, (s1 - base) * 1e-3, (e0 - base) * 1e-3, (e1 - base) * 1e-3);
                }
                const unsigned long long* h = hv.data() + (size_t)(L - 1) * 256 * 32;
                {
                    double mean = 0;
                    for (int c = 0; c < G; ++c) mean += (double)(h[c * 32 + 15] - h[c * 32 + 7]);
                    fprintf(stderr, "[conv timing] mean CTA lifetime %.1f us; per CTA (cta:smid:us:tiles:issue_cyc:full_wait_cyc):", mean / G * 1e-3);
                    for (int c = 0; c < G; ++c)
                        fprintf(stderr, " %d:%llu:%.0f:%llu:%llu:%llu", c, h[c * 32 + 4] >> 32, (h[c * 32 + 15] - h[c * 32 + 7]) * 1e-3,
                                h[c * 32 + 4] & 0xffffffffull, h[c * 32 + 2], h[c * 32 + 1]);
                    fprintf(stderr, "\n");
                    {
                        unsigned long long mn = ~0ull, mx = 0, sm = 0, fmn = ~0ull, fmx = 0;
                        for (int c = 0; c < G; ++c) {
                            const unsigned long long v = h[c * 32 + 17], f = h[c * 32 + 16];
                            if (v < mn) mn = v; if (v > mx) mx = v; sm += v;
                            if (f < fmn) fmn = f; if (f > fmx) fmx = f;
                        }
                        fprintf(stderr, "[conv timing] per-CTA ticket phase: min %llu mean %llu max %llu cycles; fence+sync min %llu max %llu\n",
                                mn, sm / G, mx, fmn, fmx);
                    }
                    fprintf(stderr, "[conv timing] CTA 0 arrival at 2nd barrier (cycles after 1st), warps 0..7: %llu %llu %llu %llu %llu %llu %llu %llu\n",
                            h[24], h[25], h[26], h[27], h[28], h[29], h[30], h[31]);
                    fprintf(stderr, "[conv timing] CTA 0 epilogue group 0 warp 0: math %llu store %llu stats-reduce %llu cycles\n", h[23], h[24], h[25]);
                    int slow = 0;
                    for (int c = 0; c < G; ++c) if (h[c * 32 + 15] > h[slow * 32 + 15]) slow = c;
                    fprintf(stderr, "[conv timing] last CTA %d: teardown cycles: fence+sync %llu, tickets %llu, fence2 %llu, finalize %llu (nfin %llu) atomic %llu dealloc %llu\n",
                            slow, h[slow * 32 + 16], h[slow * 32 + 17], h[slow * 32 + 18], h[slow * 32 + 19], h[slow * 32 + 20], h[slow * 32 + 21], h[slow * 32 + 22]);
                }
                for (int c : {0, 73, 147})
                    fprintf(stderr, "[conv timing] cta %d: mma{tempty %llu full %llu issue %llu commit %llu n %llu loop %llu} "
                            "producer{wait_empty %llu} epi{wait_tfull %llu tmem_ld %llu stats_ring %llu loop %llu} "
                            "prologue %llu main %llu total %llu cycles = %llu ns\n", c,
                            h[c * 32 + 0], h[c * 32 + 1], h[c * 32 + 2], h[c * 32 + 3], h[c * 32 + 4] & 0xffffffffull, h[c * 32 + 5], h[c * 32 + 6],
                            h[c * 32 + 8], h[c * 32 + 9], h[c * 32 + 13], h[c * 32 + 14], h[c * 32 + 10], h[c * 32 + 11], h[c * 32 + 12],
                            h[c * 32 + 15] - h[c * 32 + 7]);
                conv_tc_plan_set_debug(tp, nullptr);
                cudaFree(dbg);
            }
            if (const char* reps_env = getenv("GTTS_CONV_REPS")) {
                const int reps = atoi(reps_env);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0, s);
                for (int i = 0; i < reps && rc == 0; ++i) rc = conv_tc_launch(tp, s);
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                fprintf(stderr, "[gtts_test_conv] impl=%d kind=%d B=%d H=%d W=%d Cin=%d Cout=%d: %.1f us/launch\n", impl, kind, B, H, W,
                        Cin, Cout, 1e3f * ms / (reps > 0 ? reps : 1));
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
            conv_tc_plan_destroy(tp);
        }
    } else {
        rc = conv_ffma(ak, g, src0, src1, wpk, e, s);
    }
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(wpk); cudaFree(partials); cudaFree(counters); cudaFree(p0); cudaFree(p1); cudaFree(w3);
    if (rc) return rc;
    GTTS_CHECK_CUDA(ce);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ text encoder
int gtts_encoder_create(gtts_encoder** out, int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp,
                        int n_heads, int n_layers, int kernel_size, int window_size, int spk_emb_dim, int n_spks, int device) {
    GTTS_REQUIRE(out != nullptr, "null out pointer");
    TextEncoder* e = text_encoder_new(n_vocab, n_feats, n_channels, filter_channels, filter_channels_dp, n_heads, n_layers, kernel_size,
                                      window_size, spk_emb_dim, n_spks, device);
    if (!e) return 1;
    gtts_encoder* h = new gtts_encoder();
    h->impl = e;
    *out = h;
    return 0;
}

void gtts_encoder_destroy(gtts_encoder* h) {
    if (!h) return;
    text_encoder_delete(h->impl);
    delete h;
}

int gtts_encoder_set_param(gtts_encoder* h, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(h != nullptr, "null encoder handle");
    return text_encoder_set_param(h->impl, name, data, numel);
}

int gtts_encoder_forward(gtts_encoder* h, const int64_t* tokens, const int64_t* lengths, const float* spk, float* mu, float* logw,
                         float* x_mask, int B, int T, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null encoder handle");
    return text_encoder_forward(h->impl, (const long long*)tokens, (const long long*)lengths, spk, mu, logw, x_mask, B, T,
                                (cudaStream_t)stream);
}

int gtts_encoder_check_tokens(gtts_encoder* h, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null encoder handle");
    return text_encoder_check_tokens(h->impl, (cudaStream_t)stream);
}

long gtts_encoder_launches_last_call(const gtts_encoder* h) { return h ? text_encoder_launches_last_call(h->impl) : 0; }

// ------------------------------------------------------------------------------------------------ vocoder
int gtts_vocoder_create(gtts_vocoder** out, int resblock, int n_ups, const int* upsample_rates, const int* upsample_kernel_sizes,
                        int upsample_initial_channel, int n_rb, const int* resblock_kernel_sizes, const int* resblock_dilation_sizes,
                        int n_dil, int num_mels, int device) {
    GTTS_REQUIRE(out != nullptr, "null out pointer");
    Vocoder* v = vocoder_new(resblock, n_ups, upsample_rates, upsample_kernel_sizes, upsample_initial_channel, n_rb, resblock_kernel_sizes,
                             resblock_dilation_sizes, n_dil, num_mels, device);
    if (!v) return 1;
    gtts_vocoder* h = new gtts_vocoder();
    h->impl = v;
    *out = h;
    return 0;
}

void gtts_vocoder_destroy(gtts_vocoder* h) {
    if (!h) return;
    cudaSetDevice(vocoder_device(h->impl));
    cudaDeviceSynchronize();
    cudaFree(h->mel); cudaFree(h->audio);
    if (h->stream) cudaStreamDestroy(h->stream);
    vocoder_delete(h->impl);
    delete h;
}

int gtts_vocoder_set_param(gtts_vocoder* h, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(h != nullptr, "null vocoder handle");
    return vocoder_set_param(h->impl, name, data, numel);
}

int gtts_vocoder_set_option(gtts_vocoder* h, const char* key, long long value) {
    GTTS_REQUIRE(h != nullptr, "null vocoder handle");
    return vocoder_set_option(h->impl, key, value);
}

int gtts_vocoder_hop(const gtts_vocoder* h) { return h ? vocoder_total_upsampling(h->impl) : 0; }

int gtts_vocoder_forward(gtts_vocoder* h, const float* mel, float* audio, int B, int T, int flags, void* stream) {
    GTTS_REQUIRE(h && mel && audio, "vocoder_forward: null pointer");
    return vocoder_forward(h->impl, mel, audio, B, T, flags, (cudaStream_t)stream);
}

int gtts_vocoder_forward_host(gtts_vocoder* h, const float* mel_host, float* audio_host, int B, int T, int flags) {
    GTTS_REQUIRE(h && mel_host && audio_host, "vocoder_forward_host: null pointer");
    GTTS_REQUIRE(B >= 1 && T >= 1, "vocoder_forward_host: bad batch or length");
    GTTS_CHECK_CUDA(cudaSetDevice(vocoder_device(h->impl)));
    if (!h->stream) GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t n_audio = (size_t)B * T * vocoder_total_upsampling(h->impl);
    // the mel staging size depends on num_mels, which the caller's buffer layout fixes: (B, num_mels, T); num_mels = n_mel / (B*T)
    // is not known here, so the handle keeps what vocoder_new was given
    const size_t n_mel = (size_t)B * T * (size_t)vocoder_num_mels(h->impl);
    if (n_mel > h->cap_mel) {
        cudaFree(h->mel); h->mel = nullptr; h->cap_mel = 0;
        GTTS_CHECK_CUDA(cudaMalloc(&h->mel, n_mel * 4));
        h->cap_mel = n_mel;
    }
    if (n_audio > h->cap_audio) {
        cudaFree(h->audio); h->audio = nullptr; h->cap_audio = 0;
        GTTS_CHECK_CUDA(cudaMalloc(&h->audio, n_audio * 4));
        h->cap_audio = n_audio;
    }
    cudaStream_t s = h->stream;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(h->mel, mel_host, n_mel * 4, cudaMemcpyHostToDevice, s));
    if (int rc = vocoder_forward(h->impl, h->mel, h->audio, B, T, flags, s)) return rc;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(audio_host, h->audio, n_audio * 4, cudaMemcpyDeviceToHost, s));
    GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int gtts_vocoder_profile(gtts_vocoder* h, int B, int T, int flags, char* buf, size_t buflen, void* stream) {
    GTTS_REQUIRE(h != nullptr, "null vocoder handle");
    return vocoder_profile(h->impl, B, T, flags, buf, buflen, (cudaStream_t)stream);
}

int gtts_vocoder_cache_info(const gtts_vocoder* h, long long* out, int n) {
    GTTS_REQUIRE(h != nullptr, "null vocoder handle");
    return vocoder_cache_info(h->impl, out, n);
}

long gtts_vocoder_launches_last_call(const gtts_vocoder* h) { return h ? vocoder_launches_last_call(h->impl) : 0; }

}  // extern "C"
