// Internal C++ interface of vocoder.cu used by capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gtts {
struct Vocoder;
Vocoder* vocoder_new(int resblock, int n_ups, const int* rates, const int* up_kernels, int initial_channel, int n_rb,
                     const int* rb_kernels, const int* rb_dilations, int n_dil, int num_mels, int device);
void vocoder_delete(Vocoder* v);
int vocoder_device(const Vocoder* v);
int vocoder_total_upsampling(const Vocoder* v);
int vocoder_num_mels(const Vocoder* v);
long vocoder_launches_last_call(const Vocoder* v);
int vocoder_set_param(Vocoder* v, const char* name, const float* data, size_t numel);
int vocoder_set_option(Vocoder* v, const char* key, long long value);
int vocoder_forward(Vocoder* v, const float* mel, float* audio, int B, int T, int flags, cudaStream_t stream);
int vocoder_cache_info(const Vocoder* v, long long* out, int n);
int vocoder_profile(Vocoder* v, int B, int T, int flags, char* buf, size_t buflen, cudaStream_t stream);
}  // namespace gtts
