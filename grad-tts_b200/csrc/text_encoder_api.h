// Internal C++ interface of text_encoder.cu used by capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gtts {
struct TextEncoder;
TextEncoder* text_encoder_new(int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp, int n_heads,
                              int n_layers, int kernel_size, int window_size, int spk_emb_dim, int n_spks, int device);
void text_encoder_delete(TextEncoder* e);
int text_encoder_device(const TextEncoder* e);
long text_encoder_launches_last_call(const TextEncoder* e);
int text_encoder_set_param(TextEncoder* e, const char* name, const float* data, size_t numel);
int text_encoder_forward(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu, float* logw,
                         float* x_mask, int B, int T, cudaStream_t s);
int text_encoder_check_tokens(TextEncoder* e, cudaStream_t s);
}  // namespace gtts
