// Internal C++ interface of decoder.cu used by capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gtts {
struct Decoder;
Decoder* decoder_new(int n_spks, int n_feats, int dim, double beta_min, double beta_max, double pe_scale, int device);
void decoder_delete(Decoder* d);
int decoder_set_param(Decoder* d, const char* name, const float* data, size_t numel);
int decoder_set_option(Decoder* d, const char* key, int value);
int decoder_reverse_diffusion(Decoder* d, const float* z, const float* mask, const float* mu, const float* spk,
                              float* out, int B, int T, int n_timesteps, int flags, const float* noise,
                              cudaStream_t stream);
int decoder_estimator(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                      float* out, int B, int T, int flags, cudaStream_t stream);
int decoder_estimator_vjp(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                          const float* v, float* out_score, float* out_gx, int B, int T, int flags, cudaStream_t stream);
int decoder_estimator_backward(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                               const float* v, float* out_score, float* out_gx, float* out_gmu, float* out_gs_pix, float* out_gtb, int B,
                               int T, int flags, cudaStream_t stream);
int decoder_get_param_grad(Decoder* d, const char* name, float* dst, size_t numel, cudaStream_t stream);
int decoder_get_param_grads_flat(Decoder* d, float* dst, size_t numel, cudaStream_t stream);
int decoder_param_grad_slot(const Decoder* d, const char* name, size_t* offset, size_t* numel);
int decoder_profile_step(Decoder* d, int B, int T, int flags, int reps, char* buf, size_t buflen, cudaStream_t stream);
long decoder_launches_last_call(const Decoder* d);
int decoder_cache_info(const Decoder* d, long long* out, int n);
int decoder_device(const Decoder* d);
}  // namespace gtts
