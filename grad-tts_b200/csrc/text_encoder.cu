// Glow-TTS text encoder (tokens -> mu_x, logw, x_mask), the step before the decoder (reference model/tts.py:84).
//
// Reference: model/text_encoder.py -- TextEncoder.forward :321-335, ConvReluNorm :30-63 (prenet), Encoder :244-282,
// MultiHeadAttention with windowed relative-position embeddings :96-216, FFN :219-241, DurationPredictor :67-93,
// LayerNorm over channels :11-27 (eps 1e-4).  Inference only (dropout is the identity in eval mode).
//
// Design.  The encoder is ~7 M parameters applied once per utterance to <= a few hundred tokens: 0.01 % of the FLOPs of the
// 100-step decoder, and its output decides DISCRETE durations (ceil(exp(logw))), so it runs in plain fp32 on the CUDA cores and
// the work goes into launch count, not tensor-core tiles: activations are channels-last (B, T, C) so every Conv1d (k = 1, 3, 5)
// is the H = 1 case of the implicit-GEMM kernel in conv_ffma.cu with bias / ReLU / residual / mask in its epilogue; q, k and v are
// ONE 1x1 conv with the three weight matrices stacked; LayerNorm (+ReLU, +mask) is one warp per position; the attention kernel
// does scores, the relative-key logits, softmax, the value sum and the relative-value sum for one (sample, head, query) per warp.
// Padded positions are zeroed after every LayerNorm: they never reach a valid position (convs read masked input, attention
// masks the keys) and every output is masked, so results at valid positions are those of the reference.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "ops.h"
#include "text_encoder_api.h"

namespace gtts {

namespace {

// x[b,t,:] = emb[token] * sqrt(C) * mask; mask[b,t] = t < len[b]   (text_encoder.py:322-324; sequence_mask model/utils.py:6-10)
__global__ void embed_kernel(const long long* __restrict__ tok, const long long* __restrict__ len, const float* __restrict__ emb,
                             float* __restrict__ x, float* __restrict__ mask, float* __restrict__ mask_out, int B, int T, int C,
                             int n_vocab, float scale, int* __restrict__ status) {
    const int bt = blockIdx.x;
    const int b = bt / T, t = bt % T;
    const bool valid = (long long)t < len[b];
    long long id = tok[bt];
    if (id < 0 || id >= n_vocab) { if (threadIdx.x == 0) atomicExch(status, 1); id = 0; }
    if (threadIdx.x == 0) { mask[bt] = valid ? 1.f : 0.f; if (mask_out) mask_out[bt] = valid ? 1.f : 0.f; }
    for (int c = threadIdx.x; c < C; c += blockDim.x) x[(size_t)bt * C + c] = valid ? emb[(size_t)id * C + c] * scale : 0.f;
}

// xe[b,t,:C0] = x[b,t,:], xe[b,t,C0:] = spk[b,:] * mask   (text_encoder.py:327-328; the encoder masks its input, :273)
__global__ void concat_spk_kernel(const float* __restrict__ x, const float* __restrict__ spk, const float* __restrict__ mask,
                                  float* __restrict__ xe, int T, int C0, int Cs) {
    const int bt = blockIdx.x, b = bt / T;
    const float m = mask[bt];
    const int C = C0 + Cs;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        xe[(size_t)bt * C + c] = c < C0 ? x[(size_t)bt * C0 + c] : spk[(size_t)b * Cs + (c - C0)] * m;
}

// LayerNorm over channels, one warp per position (text_encoder.py:20-27): two-pass mean / variance, eps inside the rsqrt.
// relu_after: prenet (norm -> ReLU, :58-60); the result is multiplied by the position's mask.
template <int kMaxPerLane>
__global__ void layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mask, float* __restrict__ out, long n_pos, int C, float eps, int relu_after) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long pos = (long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (pos >= n_pos) return;
    const float m = mask[pos];
    float* op = out + pos * C;
    if (m == 0.f) {
        for (int c = lane; c < C; c += 32) op[c] = 0.f;
        return;
    }
    const float* ip = in + pos * C;
    float v[kMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int c = lane + 32 * k;
        v[k] = c < C ? ip[c] : 0.f;
        s += v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int c = lane + 32 * k;
        const float d = c < C ? v[k] - mean : 0.f;
        q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / (float)C + eps);
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
            float y = (v[k] - mean) * rstd * gamma[c] + beta[c];
            if (relu_after) y = fmaxf(y, 0.f);
            op[c] = y;
        }
    }
}

// Multi-head self-attention with windowed relative-position embeddings (text_encoder.py:140-182).
//   qkv: (B, T, 3C), q | k | v, head h = channels [h*kc, (h+1)*kc);  out: (B, T, C)
//   score(i,j) = (q_i . k_j + [|j-i| <= w] q_i . Ek[j-i+w]) / sqrt(kc);  masked keys -> -1e4;  p = softmax_j
//   out_i = sum_j p(i,j) v_j + sum_{|d| <= w} p(i,i+d) Ev[d+w]
// One warp per (sample, head, query); lanes run over keys for the scores and over channels for the output.
template <int kKC>
__global__ void rel_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ mask, const float* __restrict__ emb_k,
                                     const float* __restrict__ emb_v, float* __restrict__ out, int T, int C, int window,
                                     int heads_rel) {
    extern __shared__ float smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int i = blockIdx.x * nwarps + warp;
    float* s_q = smem + (size_t)warp * (T + kKC);
    float* s_p = s_q + kKC;
    if (i >= T) return;
    float* op = out + ((size_t)b * T + i) * C + h * kKC;
    if (mask[(size_t)b * T + i] == 0.f) {
        for (int c = lane; c < kKC; c += 32) op[c] = 0.f;
        return;
    }
    const float scale = 1.0f / sqrtf((float)kKC);
    const float* qp = qkv + ((size_t)b * T + i) * 3 * C + h * kKC;
    for (int c = lane; c < kKC; c += 32) s_q[c] = qp[c];
    __syncwarp();
    const float* ek = emb_k + (size_t)(heads_rel > 1 ? h : 0) * (2 * window + 1) * kKC;
    const float* ev = emb_v + (size_t)(heads_rel > 1 ? h : 0) * (2 * window + 1) * kKC;
    float mx = -3.0e38f;
    for (int j = lane; j < T; j += 32) {
        float s;
        if (mask[(size_t)b * T + j] == 0.f) {
            s = -1e4f;
        } else {
            const float4* kp = reinterpret_cast<const float4*>(qkv + ((size_t)b * T + j) * 3 * C + C + h * kKC);
            float acc = 0.f;
#pragma unroll 8
            for (int c4 = 0; c4 < kKC / 4; ++c4) {
                const float4 kv = __ldg(kp + c4);
                acc = fmaf(s_q[c4 * 4 + 0], kv.x, acc); acc = fmaf(s_q[c4 * 4 + 1], kv.y, acc);
                acc = fmaf(s_q[c4 * 4 + 2], kv.z, acc); acc = fmaf(s_q[c4 * 4 + 3], kv.w, acc);
            }
            s = acc * scale;
            const int d = j - i;
            if (window >= 0 && d >= -window && d <= window) {
                const float* ep = ek + (size_t)(d + window) * kKC;
                float r = 0.f;
                for (int c = 0; c < kKC; ++c) r = fmaf(s_q[c], ep[c], r);
                s += r * scale;
            }
        }
        s_p[j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
        const float e = expf(s_p[j] - mx);
        s_p[j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    __syncwarp();
    constexpr int kPer = kKC / 32;
    float acc[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) acc[k] = 0.f;
    const float* vbase = qkv + (size_t)b * T * 3 * C + 2 * C + h * kKC;
    // masked keys have p = 0 exactly (exp(-1e4 - max) underflows) and v is finite, so every key can take the same path: the loop
    // unrolls and its loads overlap
#pragma unroll 4
    for (int j = 0; j < T; ++j) {
        const float p = s_p[j];
        const float* vp = vbase + (size_t)j * 3 * C;
#pragma unroll
        for (int k = 0; k < kPer; ++k) acc[k] = fmaf(p, __ldg(vp + lane + 32 * k), acc[k]);
    }
    if (window >= 0) {
        for (int d = -window; d <= window; ++d) {
            const int j = i + d;
            if (j < 0 || j >= T) continue;
            const float p = s_p[j];
            const float* ep = ev + (size_t)(d + window) * kKC;
#pragma unroll
            for (int k = 0; k < kPer; ++k) acc[k] = fmaf(p, ep[lane + 32 * k], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) op[lane + 32 * k] = acc[k] * inv;
}

// Latency-shaped Conv1d for short token sequences (M = B*T positions of a few hundred): a 64 x 64 implicit-GEMM tile with a serial
// K loop leaves most SMs idle and pays one block barrier pair per 32 channels.  Here a CTA owns 16 positions x 32 output channels
// and its 8 warps split K = k * Cin between them in chunks of 32 channels: a warp fetches a chunk with twelve independent 16-byte
// loads per lane (its own weight row: one full 128-byte line; the 16 x 128-byte activation tile: coalesced, staged through a
// per-warp shared-memory tile), one chunk ahead of the FMAs, and never meets a block barrier until the 8 partial tiles are summed
// in shared memory in a fixed order (deterministic).  Same epilogue as conv_ffma: (+bias) (+residual) (*mask) (ReLU).
// (ncu on the first version, which loaded 16 bytes per row per step: L1 hit rate 56 %, 7.9 long-scoreboard stalls per issue,
// 70 us for the 768 -> 192 k = 3 conv at 100 tokens.)
__global__ void __launch_bounds__(256)
conv1d_splitk_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ residual, const float* __restrict__ mask, float* __restrict__ out, int M, int T, int Cin,
                     int Cout, int k, int relu) {
    __shared__ float part[8][16][33];
    __shared__ float4 a_s[8][16][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * 16, co = blockIdx.y * 32 + lane;
    const int c32n = Cin / 32;                              // 32-channel chunks per tap
    const int chunks = k * c32n;
    const int per = (chunks + 7) / 8;
    const int ch_begin = warp * per, ch_end = min(chunks, ch_begin + per);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    // this lane stages rows lane/8 + 4q (q = 0..3), 16 bytes at column lane%8 of the chunk
    const int f4 = lane & 7;
    int row_p[4], row_t[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { row_p[q] = p0 + (lane >> 3) + 4 * q; row_t[q] = row_p[q] % T; }
    float4 a_n[4], w_n[8], w_c[8];
    auto prefetch = [&](int ch) {
        const int tap = ch / c32n, cc = ch - tap * c32n;
        const int dt = tap - k / 2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ti = row_t[q] + dt;
            a_n[q] = (row_p[q] < M && ti >= 0 && ti < T)
                         ? __ldg(reinterpret_cast<const float4*>(x + (size_t)(row_p[q] + dt) * Cin) + cc * 8 + f4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4* wr = reinterpret_cast<const float4*>(w + ((size_t)tap * Cout + co) * Cin) + cc * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) w_n[u] = __ldg(wr + u);
    };
    if (ch_begin < ch_end) prefetch(ch_begin);
    for (int ch = ch_begin; ch < ch_end; ++ch) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a_s[warp][(lane >> 3) + 4 * q][f4] = a_n[q];
#pragma unroll
        for (int u = 0; u < 8; ++u) w_c[u] = w_n[u];
        __syncwarp();
        if (ch + 1 < ch_end) prefetch(ch + 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 a = a_s[warp][i][u];
                acc[i] = fmaf(a.x, w_c[u].x, acc[i]); acc[i] = fmaf(a.y, w_c[u].y, acc[i]);
                acc[i] = fmaf(a.z, w_c[u].z, acc[i]); acc[i] = fmaf(a.w, w_c[u].w, acc[i]);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) part[warp][i][lane] = acc[i];
    __syncthreads();
    for (int o = threadIdx.x; o < 16 * 32; o += 256) {
        const int i = o >> 5, c = o & 31;
        const int p = p0 + i;
        if (p >= M) continue;
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += part[q][i][c];
        const int cc = blockIdx.y * 32 + c;
        if (bias) v += bias[cc];
        if (residual) v += residual[(size_t)p * Cout + cc];
        if (mask) v *= mask[p];
        if (relu) v = fmaxf(v, 0.f);
        out[(size_t)p * Cout + cc] = v;
    }
}

// Attention for sequences whose keys and values fit in shared memory (T * kc * 8 bytes <= ~200 KB: 270 tokens at kc = 96).  A CTA
// stages K and V of one (sample, head) once -- coalesced, rows padded to kc + 1 floats so that "lane = key" reads are conflict-free
// -- plus the two relative-embedding tables, and its warps take queries round-robin: scores with lane = key, the relative-key
// logits with lanes across channels and a shuffle reduction, softmax, then the value sum with lane = channel.  The query range is
// split over gridDim.x CTAs so that a single utterance still fills a few dozen SMs.  Same arithmetic as rel_attention_kernel.
// (ncu on rel_attention_kernel at 1 x 100 tokens: 58 us, 29 long-scoreboard stalls per issue -- every key row came from L1/L2.)
template <int kKC>
__global__ void __launch_bounds__(256)
rel_attention_smem_kernel(const float* __restrict__ qkv, const float* __restrict__ mask, const float* __restrict__ emb_k,
                          const float* __restrict__ emb_v, float* __restrict__ out, int T, int C, int window, int q_per_cta) {
    extern __shared__ float smem[];
    constexpr int kPitch = kKC + 4;                     // 16-byte aligned rows; 8 lanes x 16 bytes of consecutive keys hit 32 distinct banks
    const int nrel = window >= 0 ? 2 * window + 1 : 0;
    float* s_k = smem;                                   // [T][kPitch]
    float* s_v = s_k + (size_t)T * kPitch;               // [T][kPitch]
    float* s_ek = s_v + (size_t)T * kPitch;              // [nrel][kKC]
    float* s_ev = s_ek + (size_t)nrel * kKC;
    const int Tp = (T + 3) / 4 * 4;
    float* s_m = s_ev + (size_t)nrel * kKC;              // [Tp] key mask
    float* s_w = s_m + Tp;                               // per warp: q[kKC] | p[Tp]  (16-byte aligned)
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const float* base = qkv + (size_t)b * T * 3 * C + h * kKC;
    for (int idx = threadIdx.x; idx < T * (kKC / 4); idx += blockDim.x) {
        const int j = idx / (kKC / 4), c4 = idx - j * (kKC / 4);
        const float4 kv = __ldg(reinterpret_cast<const float4*>(base + (size_t)j * 3 * C + C) + c4);
        const float4 vv = __ldg(reinterpret_cast<const float4*>(base + (size_t)j * 3 * C + 2 * C) + c4);
        *reinterpret_cast<float4*>(s_k + (size_t)j * kPitch + c4 * 4) = kv;
        *reinterpret_cast<float4*>(s_v + (size_t)j * kPitch + c4 * 4) = vv;
    }
    for (int idx = threadIdx.x; idx < nrel * kKC; idx += blockDim.x) { s_ek[idx] = emb_k[idx]; s_ev[idx] = emb_v[idx]; }
    for (int j = threadIdx.x; j < T; j += blockDim.x) s_m[j] = mask[(size_t)b * T + j];
    __syncthreads();
    float* s_q = s_w + (size_t)warp * (kKC + Tp);
    float* s_p = s_q + kKC;
    const float scale = 1.0f / sqrtf((float)kKC);
    constexpr int kPer = kKC / 32;
    const int q_begin = blockIdx.x * q_per_cta, q_end = min(T, q_begin + q_per_cta);
    for (int i = q_begin + warp; i < q_end; i += nwarps) {
        float* op = out + ((size_t)b * T + i) * C + h * kKC;
        if (s_m[i] == 0.f) {
#pragma unroll
            for (int k = 0; k < kPer; ++k) op[lane + 32 * k] = 0.f;
            continue;
        }
        float qreg[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) { qreg[k] = __ldg(base + (size_t)i * 3 * C + lane + 32 * k); s_q[lane + 32 * k] = qreg[k]; }
        __syncwarp();
        float mx = -3.0e38f;
        for (int j = lane; j < T; j += 32) {
            float sc;
            if (s_m[j] == 0.f) {
                sc = -1e4f;
            } else {
                // one 16-byte read of the key row and one 16-byte broadcast of q per four FMAs (the first version read both
                // operands word by word: two shared-memory instructions per FMA, LDS-bound)
                const float4* kp = reinterpret_cast<const float4*>(s_k + (size_t)j * kPitch);
                const float4* qp4 = reinterpret_cast<const float4*>(s_q);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
                for (int c4 = 0; c4 < kKC / 4; ++c4) {
                    const float4 kv = kp[c4], qv = qp4[c4];
                    a0 = fmaf(qv.x, kv.x, a0); a1 = fmaf(qv.y, kv.y, a1);
                    a2 = fmaf(qv.z, kv.z, a2); a3 = fmaf(qv.w, kv.w, a3);
                }
                sc = ((a0 + a1) + (a2 + a3)) * scale;
            }
            s_p[j] = sc;
        }
        __syncwarp();
        // relative-key logits: lanes across channels, one shuffle reduction per offset
        for (int d = -window; d <= window && window >= 0; ++d) {
            const int j = i + d;
            if (j < 0 || j >= T) continue;                  // warp-uniform
            const float* ep = s_ek + (size_t)(d + window) * kKC;
            float r = 0.f;
#pragma unroll
            for (int k = 0; k < kPer; ++k) r = fmaf(qreg[k], ep[lane + 32 * k], r);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0 && s_m[j] != 0.f) s_p[j] += r * scale;
        }
        __syncwarp();
        for (int j = lane; j < T; j += 32) mx = fmaxf(mx, s_p[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) {
            const float e = expf(s_p[j] - mx);
            s_p[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        __syncwarp();
        float acc[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) acc[k] = 0.f;
#pragma unroll 4
        for (int j = 0; j < T; ++j) {
            const float p = s_p[j];
            const float* vp = s_v + (size_t)j * kPitch;
#pragma unroll
            for (int k = 0; k < kPer; ++k) acc[k] = fmaf(p, vp[lane + 32 * k], acc[k]);
        }
        for (int d = -window; d <= window && window >= 0; ++d) {
            const int j = i + d;
            if (j < 0 || j >= T) continue;
            const float p = s_p[j];
            const float* ep = s_ev + (size_t)(d + window) * kKC;
#pragma unroll
            for (int k = 0; k < kPer; ++k) acc[k] = fmaf(p, ep[lane + 32 * k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) op[lane + 32 * k] = acc[k] * inv;
        __syncwarp();
    }
}

// 1x1 projection to the reference's (B, Cout, T) layout with the mask: out[b,co,t] = (x[b,t,:] . w[co,:] + bias[co]) * mask[b,t]
// (proj_m :329).  Block = 8 positions; thread = output channel, 8 accumulators; wt is the weight transposed to [C][Cout] so that
// the lanes of a warp read consecutive addresses.
__global__ void proj_nct_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                                const float* __restrict__ mask, float* __restrict__ out, int T, int C, int Cout) {
    extern __shared__ float xs[];                     // [8][C]
    const int b = blockIdx.y, t0 = blockIdx.x * 8;
    for (int idx = threadIdx.x; idx < 8 * C; idx += blockDim.x) {
        const int p = idx / C;
        xs[idx] = (t0 + p < T) ? x[((size_t)b * T + t0) * C + idx] : 0.f;
    }
    __syncthreads();
    const int co = threadIdx.x;
    if (co >= Cout) return;
    float acc[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) acc[p] = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        const float wv = __ldg(wt + (size_t)c * Cout + co);
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[p] = fmaf(xs[p * C + c], wv, acc[p]);
    }
    const float bv = bias[co];
#pragma unroll
    for (int p = 0; p < 8; ++p)
        if (t0 + p < T) out[((size_t)b * Cout + co) * T + t0 + p] = (acc[p] + bv) * mask[(size_t)b * T + t0 + p];
}

// the duration predictor's final projection (Cout = 1, :91-92): one warp per position
__global__ void proj1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                             const float* __restrict__ mask, float* __restrict__ out, long n_pos, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long pos = (long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (pos >= n_pos) return;
    const float* xp = x + pos * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(xp[c], __ldg(w + c), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[pos] = (acc + bias[0]) * mask[pos];
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
    const int n = R * Cc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / Cc, c = i % Cc;
        dst[(size_t)c * R + r] = src[i];
    }
}

struct Param { float* p = nullptr; size_t numel = 0; };

}  // namespace

struct TextEncoder {
    int device = 0, num_sms = 148;
    int n_vocab, n_feats, C0, F, Fdp, heads, layers, ksize, window, spk_dim, n_spks;
    int C;                                              // encoder width: C0 (+ spk_dim when n_spks > 1)
    std::map<std::string, Param> params;
    std::map<std::string, float*> packed;               // conv weights in the [tap][Cout] x [Cin] layout of conv_ffma
    bool packed_valid = false;
    float* ws = nullptr;
    size_t ws_bytes = 0;
    int* status = nullptr;
    long launches_last_call = 0;
    long small_m = 1024;                                // B*T up to which the latency-shaped conv kernel is used
    struct Graph {
        cudaGraphExec_t exec = nullptr;
        void* stage = nullptr;
        long long *tok = nullptr, *len = nullptr;
        float *spk = nullptr, *mu = nullptr, *logw = nullptr, *mask = nullptr;
        long launches = 0;
        unsigned long long last_use = 0;
    };
    std::map<std::pair<int, int>, Graph> graphs;        // per (B, T)
    unsigned long long use_clock = 0;
    int use_graph = 1;
    cudaEvent_t done_ev = nullptr;                      // calls on one handle are ordered across streams (one workspace)
    cudaStream_t last_stream = nullptr;
    bool has_done = false;
    long graph_max_pos = 8192;                          // larger batches are not launch-bound
    ~TextEncoder() {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        for (auto& kv : graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); cudaFree(kv.second.stage); }
        if (done_ev) cudaEventDestroy(done_ev);
        for (auto& kv : params) cudaFree(kv.second.p);
        for (auto& kv : packed) cudaFree(kv.second);
        cudaFree(ws); cudaFree(status);
    }
};

namespace {

ConvGeom geom1d(int B, int T, int Cin, int Cout, int k) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = 1; g.Win = T; g.Hg = 1; g.Wg = T; g.Hout = 1; g.Wout = T;
    g.Cin0 = Cin; g.Cout = Cout; g.ntaps = k; g.nphase = 1; g.stride = 1; g.out_step = 1;
    for (int t = 0; t < k; ++t) { g.dx[0][t] = (int8_t)(t - k / 2); g.wrow[0][t] = t * Cout; }
    return g;
}

struct Run {
    TextEncoder* e;
    int B, T;
    cudaStream_t s;
    long launches = 0;
    int rc = 0;
    // GTTS_ENC_PROFILE=1: an event after every launch, per-launch times on stderr at the end of the call (measurement aid)
    bool prof = false;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void mark(const std::string& name) {
        if (!prof) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        marks.emplace_back(name, ev);
    }
    void report() {
        if (!prof || marks.empty()) return;
        cudaEventSynchronize(marks.back().second);
        std::map<std::string, float> agg;
        float total = 0.f;
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            agg[marks[i].first] += ms;
            total += ms;
        }
        fprintf(stderr, "[text encoder profile] B=%d T=%d: %.1f us over %zu launches\n", B, T, total * 1e3f, marks.size() - 1);
        for (auto& kv : agg) fprintf(stderr, "   %-28s %9.1f us\n", kv.first.c_str(), kv.second * 1e3f);
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }

    const float* P(const std::string& name, size_t numel) {
        auto it = e->params.find(name);
        if (it == e->params.end() || !it->second.p) { if (!rc) { set_error("text encoder: parameter " + name + " was not set"); rc = 3; } return nullptr; }
        if (it->second.numel != numel) { if (!rc) { set_error("text encoder: parameter " + name + " has the wrong number of elements"); rc = 3; } return nullptr; }
        return it->second.p;
    }
    const float* W(const std::string& name) {
        auto it = e->packed.find(name);
        if (it == e->packed.end()) { if (!rc) { set_error("text encoder: conv weight " + name + " was not packed"); rc = 3; } return nullptr; }
        return it->second;
    }
    // out = [relu](conv(x) + bias) [+ residual] [* mask]
    void conv(const std::string& name, int Cin, int Cout, int k, const float* x, const float* residual, const float* mask, float* out,
              bool relu) {
        if (rc) return;
        const float* w = W(name + ".weight");
        const float* b = P(name + ".bias", (size_t)Cout);
        if (rc) return;
        conv_raw(Cin, Cout, k, x, w, b, residual, mask, out, relu);
    }
    void conv_raw(int Cin, int Cout, int k, const float* x, const float* w, const float* b, const float* residual, const float* mask,
                  float* out, bool relu) {
        if (rc) return;
        const long M = (long)B * T;
        if (M <= e->small_m) {
            dim3 grid((unsigned)((M + 15) / 16), Cout / 32);
            conv1d_splitk_kernel<<<grid, 256, 0, s>>>(x, w, b, residual, mask, out, (int)M, T, Cin, Cout, k, relu ? 1 : 0);
            if (cudaGetLastError() != cudaSuccess) { set_error("text encoder: conv launch failed"); rc = 1; }
        } else {
            ConvEpilogue ep;
            memset(&ep, 0, sizeof(ep));
            ep.bias = b; ep.residual = residual; ep.mask = mask; ep.out = out;
            if (relu) { ep.act_out = 1; ep.act_slope = 0.f; }
            rc = conv_ffma(ACT_F32, geom1d(B, T, Cin, Cout, k), x, nullptr, w, ep, s);
        }
        ++launches;
        mark("conv_k" + std::to_string(k) + "_" + std::to_string(Cin) + "_" + std::to_string(Cout));
    }
    void layernorm(const std::string& name, int C, const float* in, const float* mask, float* out, bool relu_after) {
        if (rc) return;
        const float* g = P(name + ".gamma", (size_t)C);
        const float* bt = P(name + ".beta", (size_t)C);
        if (rc) return;
        const long n_pos = (long)B * T;
        const int wpb = 8;
        const int blocks = (int)((n_pos + wpb - 1) / wpb);
        if (C <= 256) layernorm_kernel<8><<<blocks, wpb * 32, 0, s>>>(in, g, bt, mask, out, n_pos, C, 1e-4f, relu_after ? 1 : 0);
        else layernorm_kernel<32><<<blocks, wpb * 32, 0, s>>>(in, g, bt, mask, out, n_pos, C, 1e-4f, relu_after ? 1 : 0);
        if (cudaGetLastError() != cudaSuccess) { set_error("text encoder: layernorm launch failed"); rc = 1; }
        ++launches;
        mark("layernorm_" + std::to_string(C));
    }
};

int pack_all(TextEncoder* e) {
    if (e->packed_valid) return 0;
    auto pack = [&](const std::string& name, int Cout, int Cin, int k) -> int {
        auto it = e->params.find(name);
        if (it == e->params.end() || !it->second.p) { set_error("text encoder: parameter " + name + " was not set"); return 3; }
        GTTS_REQUIRE(it->second.numel == (size_t)Cout * Cin * k, "text encoder: conv weight has the wrong number of elements");
        float*& dst = e->packed[name];
        if (!dst) GTTS_CHECK_CUDA(cudaMalloc((void**)&dst, (size_t)Cout * Cin * k * 4));
        return pack_conv1d_weight(ACT_F32, it->second.p, dst, Cout, Cin, k, Cout, Cin, false, 0);
    };
    const int C0 = e->C0, C = e->C;
    for (int i = 0; i < 3; ++i)
        if (int rc = pack("prenet.conv_layers." + std::to_string(i) + ".weight", C0, C0, 5)) return rc;
    if (int rc = pack("prenet.proj.weight", C0, C0, 1)) return rc;
    for (int l = 0; l < e->layers; ++l) {
        const std::string a = "encoder.attn_layers." + std::to_string(l), f = "encoder.ffn_layers." + std::to_string(l);
        // q, k, v stacked into one (3C x C) 1x1 conv
        {
            float*& dst = e->packed[a + ".qkv.weight"];
            float*& bdst = e->packed[a + ".qkv.bias"];
            if (!dst) GTTS_CHECK_CUDA(cudaMalloc((void**)&dst, (size_t)3 * C * C * 4));
            if (!bdst) GTTS_CHECK_CUDA(cudaMalloc((void**)&bdst, (size_t)3 * C * 4));
            const char* names[3] = {".conv_q", ".conv_k", ".conv_v"};
            for (int q = 0; q < 3; ++q) {
                auto iw = e->params.find(a + names[q] + ".weight");
                auto ib = e->params.find(a + names[q] + ".bias");
                if (iw == e->params.end() || ib == e->params.end()) { set_error("text encoder: parameter " + a + names[q] + " was not set"); return 3; }
                GTTS_REQUIRE(iw->second.numel == (size_t)C * C && ib->second.numel == (size_t)C, "text encoder: attention conv has the wrong size");
                // (C, C, 1) PyTorch layout is already [Cout][Cin] rows
                GTTS_CHECK_CUDA(cudaMemcpy(dst + (size_t)q * C * C, iw->second.p, (size_t)C * C * 4, cudaMemcpyDeviceToDevice));
                GTTS_CHECK_CUDA(cudaMemcpy(bdst + (size_t)q * C, ib->second.p, (size_t)C * 4, cudaMemcpyDeviceToDevice));
            }
        }
        if (int rc = pack(a + ".conv_o.weight", C, C, 1)) return rc;
        if (int rc = pack(f + ".conv_1.weight", e->F, C, e->ksize)) return rc;
        if (int rc = pack(f + ".conv_2.weight", C, e->F, e->ksize)) return rc;
    }
    {
        auto it = e->params.find("proj_m.weight");
        if (it == e->params.end() || !it->second.p) { set_error("text encoder: parameter proj_m.weight was not set"); return 3; }
        GTTS_REQUIRE(it->second.numel == (size_t)e->n_feats * C, "text encoder: proj_m.weight has the wrong number of elements");
        float*& dst = e->packed["proj_m.wt"];
        if (!dst) GTTS_CHECK_CUDA(cudaMalloc((void**)&dst, (size_t)e->n_feats * C * 4));
        transpose_kernel<<<64, 256>>>(it->second.p, dst, e->n_feats, C);
        GTTS_CHECK_CUDA(cudaGetLastError());
    }
    if (int rc = pack("proj_w.conv_1.weight", e->Fdp, C, e->ksize)) return rc;
    if (int rc = pack("proj_w.conv_2.weight", e->Fdp, e->Fdp, e->ksize)) return rc;
    GTTS_CHECK_CUDA(cudaDeviceSynchronize());
    e->packed_valid = true;
    return 0;
}

template <int kKC>
int launch_attention(const float* qkv, const float* mask, const float* ek, const float* ev, float* out, int B, int T, int C, int heads,
                     int window, int num_sms, cudaStream_t s) {
    const int nrel = window >= 0 ? 2 * window + 1 : 0;
    {
        // keys and values of one (sample, head) resident in shared memory
        const int nwarps = 8;
        const size_t smem = ((size_t)2 * T * (kKC + 4) + (size_t)2 * nrel * kKC + (size_t)(T + 3) / 4 * 4 + (size_t)nwarps * (kKC + (T + 3) / 4 * 4)) * 4;
        if (smem <= 220 * 1024 && !getenv("GTTS_ENC_ATTN_GLOBAL")) {
            static size_t attr = 0;
            if (smem > 48 * 1024 && smem > attr) {
                GTTS_CHECK_CUDA(cudaFuncSetAttribute(rel_attention_smem_kernel<kKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                attr = 220 * 1024;
            }
            // split the queries of a (sample, head) over enough CTAs to fill the GPU once, at least 8 queries (one per warp) each
            int split = std::max(1, num_sms / std::max(1, B * heads));
            split = std::min(split, (T + nwarps - 1) / nwarps);
            const int q_per_cta = (T + split - 1) / split;
            dim3 grid((T + q_per_cta - 1) / q_per_cta, heads, B);
            rel_attention_smem_kernel<kKC><<<grid, nwarps * 32, smem, s>>>(qkv, mask, ek, ev, out, T, C, window, q_per_cta);
            GTTS_CHECK_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const int nwarps = 8;
    const size_t smem = (size_t)nwarps * (T + kKC) * 4;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        GTTS_REQUIRE(smem <= 200 * 1024, "text encoder: sequence too long for the attention kernel");
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(rel_attention_kernel<kKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid((T + nwarps - 1) / nwarps, heads, B);
    rel_attention_kernel<kKC><<<grid, nwarps * 32, smem, s>>>(qkv, mask, ek, ev, out, T, C, window, 1);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

TextEncoder* text_encoder_new(int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp, int n_heads,
                              int n_layers, int kernel_size, int window_size, int spk_emb_dim, int n_spks, int device) {
    const int C = n_channels + (n_spks > 1 ? spk_emb_dim : 0);
    if (n_vocab < 1 || n_feats < 1 || n_layers < 0 || n_heads < 1 || C % n_heads || (kernel_size != 1 && kernel_size != 3 && kernel_size != 5)) {
        set_error("text_encoder_new: bad configuration");
        return nullptr;
    }
    const int kc = C / n_heads;
    if (n_channels % 64 || C % 64 || filter_channels % 64 || filter_channels_dp % 64 || !(kc == 32 || kc == 64 || kc == 96 || kc == 128)) {
        set_error("text_encoder_new: channel counts must be multiples of 64 and channels / n_heads one of 32, 64, 96, 128");
        return nullptr;
    }
    if (window_size > 64) { set_error("text_encoder_new: window_size out of range"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("text_encoder_new: cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        set_error("text_encoder_new: this library is built for sm_100a (B200) only");
        return nullptr;
    }
    TextEncoder* e = new TextEncoder();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    e->n_vocab = n_vocab; e->n_feats = n_feats; e->C0 = n_channels; e->F = filter_channels; e->Fdp = filter_channels_dp;
    e->heads = n_heads; e->layers = n_layers; e->ksize = kernel_size; e->window = window_size; e->spk_dim = spk_emb_dim; e->n_spks = n_spks;
    e->C = C;
    if (const char* sm = getenv("GTTS_ENC_SMALL_M")) e->small_m = atol(sm);
    if (const char* ug = getenv("GTTS_ENC_GRAPH")) e->use_graph = atoi(ug) != 0;
    if (cudaMalloc((void**)&e->status, 4) != cudaSuccess) { set_error("text_encoder_new: out of device memory"); delete e; return nullptr; }
    cudaMemset(e->status, 0, 4);
    return e;
}

void text_encoder_delete(TextEncoder* e) { delete e; }
int text_encoder_device(const TextEncoder* e) { return e->device; }
long text_encoder_launches_last_call(const TextEncoder* e) { return e->launches_last_call; }

int text_encoder_set_param(TextEncoder* e, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(e && name && data && numel > 0, "text_encoder_set_param: null argument");
    GTTS_CHECK_CUDA(cudaSetDevice(e->device));
    GTTS_CHECK_CUDA(cudaDeviceSynchronize());
    Param& p = e->params[name];
    if (p.p && p.numel != numel) {
        // captured graphs hold the old pointer
        for (auto& kv : e->graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); cudaFree(kv.second.stage); }
        e->graphs.clear();
        cudaFree(p.p); p.p = nullptr;
    }
    if (!p.p) GTTS_CHECK_CUDA(cudaMalloc((void**)&p.p, numel * 4));
    p.numel = numel;
    GTTS_CHECK_CUDA(cudaMemcpy(p.p, data, numel * 4, cudaMemcpyDefault));
    e->packed_valid = false;
    return 0;
}

// tokens (B, T) int64, lengths (B) int64, spk (B, spk_emb_dim) or null -> mu (B, n_feats, T), logw (B, 1, T), x_mask (B, 1, T)
namespace {

size_t workspace_bytes(const TextEncoder* e, int B, int T) {
    const size_t n_pos = (size_t)B * T, wide = (size_t)std::max(e->F, e->Fdp);
    return (n_pos * (1 + 3 * (size_t)e->C0 + 3 * (size_t)e->C + 3 * (size_t)e->C + 2 * wide) + 1024) * 4;
}

void drop_graphs(TextEncoder* e) {
    for (auto& kv : e->graphs) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        cudaFree(kv.second.stage);
    }
    e->graphs.clear();
}

int forward_body(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu, float* logw,
                 float* x_mask, int B, int T, cudaStream_t s);

}  // namespace

// One forward = ~56 small launches; for a given (B, T) they are captured once as a CUDA graph over fixed staging buffers (tokens,
// lengths, speaker in; mu, logw, mask out -- a few KB copied on either side), so a call costs one graph launch instead of 56
// kernel launches from the host.  At most 16 shapes are kept (least recently used goes first); growing the workspace drops them all.
static int text_encoder_forward_impl(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu,
                                     float* logw, float* x_mask, int B, int T, cudaStream_t s);

int text_encoder_forward(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu, float* logw,
                         float* x_mask, int B, int T, cudaStream_t s) {
    GTTS_REQUIRE(e != nullptr, "text_encoder_forward: null handle");
    GTTS_CHECK_CUDA(cudaSetDevice(e->device));
    if (!e->done_ev) GTTS_CHECK_CUDA(cudaEventCreateWithFlags(&e->done_ev, cudaEventDisableTiming));
    if (e->has_done && e->last_stream != s) GTTS_CHECK_CUDA(cudaStreamWaitEvent(s, e->done_ev, 0));
    const int rc = text_encoder_forward_impl(e, tokens, lengths, spk, mu, logw, x_mask, B, T, s);
    if (rc) return rc;
    GTTS_CHECK_CUDA(cudaEventRecord(e->done_ev, s));
    e->last_stream = s;
    e->has_done = true;
    return 0;
}

static int text_encoder_forward_impl(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu,
                                     float* logw, float* x_mask, int B, int T, cudaStream_t s) {
    GTTS_REQUIRE(e && tokens && lengths && mu && logw && x_mask, "text_encoder_forward: null pointer");
    GTTS_REQUIRE(B >= 1 && T >= 1, "text_encoder_forward: bad batch or length");
    GTTS_REQUIRE(e->n_spks <= 1 || spk != nullptr, "text_encoder_forward: this encoder was built with n_spks > 1: spk is required");
    GTTS_CHECK_CUDA(cudaSetDevice(e->device));
    if (int rc = pack_all(e)) return rc;
    const size_t need = workspace_bytes(e, B, T);
    if (need > e->ws_bytes) {
        GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
        GTTS_CHECK_CUDA(cudaDeviceSynchronize());
        drop_graphs(e);
        cudaFree(e->ws); e->ws = nullptr; e->ws_bytes = 0;
        GTTS_CHECK_CUDA(cudaMalloc((void**)&e->ws, need));
        e->ws_bytes = need;
    }
    const bool want_graph = e->use_graph && (long)B * T <= e->graph_max_pos && !getenv("GTTS_ENC_PROFILE");
    if (!want_graph) return forward_body(e, tokens, lengths, spk, mu, logw, x_mask, B, T, s);

    const size_t n_pos = (size_t)B * T;
    const size_t n_spk = e->n_spks > 1 ? (size_t)B * e->spk_dim : 0;
    const std::pair<int, int> key(B, T);
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        GTTS_CHECK_CUDA(cudaStreamSynchronize(s));            // the dry run below shares the workspace with whatever was queued
        if (e->graphs.size() >= 16) {
            auto lru = e->graphs.begin();
            for (auto j = e->graphs.begin(); j != e->graphs.end(); ++j)
                if (j->second.last_use < lru->second.last_use) lru = j;
            GTTS_CHECK_CUDA(cudaDeviceSynchronize());
            if (lru->second.exec) cudaGraphExecDestroy(lru->second.exec);
            cudaFree(lru->second.stage);
            e->graphs.erase(lru);
        }
        TextEncoder::Graph g;
        // staging: tokens (n_pos int64) | lengths (B int64) | spk | mu | logw | mask
        const size_t bytes = (n_pos + B) * 8 + (n_spk + n_pos * e->n_feats + 2 * n_pos) * 4 + 256;
        GTTS_CHECK_CUDA(cudaMalloc(&g.stage, bytes));
        char* q = (char*)g.stage;
        g.tok = (long long*)q; q += n_pos * 8;
        g.len = (long long*)q; q += (size_t)B * 8;
        g.spk = (float*)q; q += n_spk * 4;
        g.mu = (float*)q; q += n_pos * e->n_feats * 4;
        g.logw = (float*)q; q += n_pos * 4;
        g.mask = (float*)q;
        cudaStream_t cs;
        GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(g.tok, tokens, n_pos * 8, cudaMemcpyDeviceToDevice, cs));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(g.len, lengths, (size_t)B * 8, cudaMemcpyDeviceToDevice, cs));
        if (n_spk) GTTS_CHECK_CUDA(cudaMemcpyAsync(g.spk, spk, n_spk * 4, cudaMemcpyDeviceToDevice, cs));
        int rc = forward_body(e, g.tok, g.len, n_spk ? g.spk : nullptr, g.mu, g.logw, g.mask, B, T, cs);   // dry run (validates every launch)
        if (!rc && cudaStreamSynchronize(cs) != cudaSuccess) { set_error("text encoder: dry run failed"); rc = 1; }
        cudaGraph_t graph = nullptr;
        if (!rc) {
            cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                rc = forward_body(e, g.tok, g.len, n_spk ? g.spk : nullptr, g.mu, g.logw, g.mask, B, T, cs);
                ce = cudaStreamEndCapture(cs, &graph);
            }
            if (!rc && ce != cudaSuccess) { set_error(std::string("text encoder: graph capture failed: ") + cudaGetErrorString(ce)); rc = 1; }
            if (!rc && cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) { set_error("text encoder: cudaGraphInstantiate failed"); rc = 1; }
            if (graph) cudaGraphDestroy(graph);
        }
        cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
        if (rc) { cudaFree(g.stage); cudaGetLastError(); return rc; }
        g.launches = e->launches_last_call;
        it = e->graphs.emplace(key, g).first;
    }
    TextEncoder::Graph& g = it->second;
    g.last_use = ++e->use_clock;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(g.tok, tokens, n_pos * 8, cudaMemcpyDeviceToDevice, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(g.len, lengths, (size_t)B * 8, cudaMemcpyDeviceToDevice, s));
    if (n_spk) GTTS_CHECK_CUDA(cudaMemcpyAsync(g.spk, spk, n_spk * 4, cudaMemcpyDeviceToDevice, s));
    GTTS_CHECK_CUDA(cudaGraphLaunch(g.exec, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(mu, g.mu, n_pos * e->n_feats * 4, cudaMemcpyDeviceToDevice, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(logw, g.logw, n_pos * 4, cudaMemcpyDeviceToDevice, s));
    GTTS_CHECK_CUDA(cudaMemcpyAsync(x_mask, g.mask, n_pos * 4, cudaMemcpyDeviceToDevice, s));
    e->launches_last_call = g.launches;
    return 0;
}

namespace {

int forward_body(TextEncoder* e, const long long* tokens, const long long* lengths, const float* spk, float* mu, float* logw,
                 float* x_mask, int B, int T, cudaStream_t s) {
    GTTS_REQUIRE((size_t)e->C * 8 * 4 <= 48 * 1024 && e->n_feats <= 1024, "text_encoder_forward: channel count too large for the projection kernel");
    const int C0 = e->C0, C = e->C, F = e->F, Fdp = e->Fdp;
    const size_t n_pos = (size_t)B * T;
    // workspace: mask | x0 (C0) | a (C0) | b (C0) | xe (C) | t1 (C) | t2 (C) | qkv (3C) | wide (max(F, Fdp)) x 2
    const size_t wide = (size_t)std::max(F, Fdp);
    float* p = e->ws;
    auto take = [&](size_t n) { float* r = p; p += (n + 15) / 16 * 16; return r; };
    float* mask = take(n_pos);
    float* x0 = take(n_pos * C0); float* ha = take(n_pos * C0); float* hb = take(n_pos * C0);
    float* xe = take(n_pos * C); float* t1 = take(n_pos * C); float* t2 = take(n_pos * C);
    float* qkv = take(n_pos * 3 * C);
    float* w1 = take(n_pos * wide); float* w2 = take(n_pos * wide);

    Run r{e, B, T, s};
    r.prof = getenv("GTTS_ENC_PROFILE") != nullptr;
    r.mark("start");
    // ---- embedding (:322-324)
    {
        const float* emb = r.P("emb.weight", (size_t)e->n_vocab * C0);
        if (r.rc) return r.rc;
        GTTS_CHECK_CUDA(cudaMemsetAsync(e->status, 0, 4, s));
        embed_kernel<<<(unsigned)n_pos, 64, 0, s>>>(tokens, lengths, emb, x0, mask, x_mask, B, T, C0, e->n_vocab, sqrtf((float)C0), e->status);
        GTTS_CHECK_CUDA(cudaGetLastError());
        ++r.launches;
    }
    // ---- prenet: 3 x (conv k5 -> LayerNorm -> ReLU), 1x1 proj, residual, mask (:53-63)
    {
        const float* cur = x0;
        float* bufs[2] = {ha, hb};
        for (int i = 0; i < 3; ++i) {
            const std::string n = std::to_string(i);
            r.conv("prenet.conv_layers." + n, C0, C0, 5, cur, nullptr, nullptr, t1 /* raw, C0 <= C */, false);
            r.layernorm("prenet.norm_layers." + n, C0, t1, mask, bufs[i & 1], true);
            cur = bufs[i & 1];
        }
        r.conv("prenet.proj", C0, C0, 1, cur, x0, mask, hb == cur ? ha : hb, false);
        if (r.rc) return r.rc;
        const float* pre = hb == cur ? ha : hb;
        if (e->n_spks > 1) {
            concat_spk_kernel<<<(unsigned)n_pos, 128, 0, s>>>(pre, spk, mask, xe, T, C0, e->spk_dim);
            GTTS_CHECK_CUDA(cudaGetLastError());
            ++r.launches;
        } else {
            GTTS_CHECK_CUDA(cudaMemcpyAsync(xe, pre, n_pos * C0 * 4, cudaMemcpyDeviceToDevice, s));
        }
    }
    // ---- encoder layers (:271-282)
    float* x = xe;
    float* xn = t2;
    const int kc = C / e->heads;
    for (int l = 0; l < e->layers && !r.rc; ++l) {
        const std::string a = "encoder.attn_layers." + std::to_string(l), f = "encoder.ffn_layers." + std::to_string(l);
        // q | k | v in one 1x1 conv
        {
            const float* qb = r.W(a + ".qkv.bias");
            const float* w = r.W(a + ".qkv.weight");
            if (r.rc) break;
            r.conv_raw(C, 3 * C, 1, x, w, qb, nullptr, nullptr, qkv, false);
            if (r.rc) break;
        }
        {
            const float* ek = nullptr; const float* ev = nullptr;
            if (e->window >= 0) {
                ek = r.P(a + ".emb_rel_k", (size_t)(2 * e->window + 1) * kc);
                ev = r.P(a + ".emb_rel_v", (size_t)(2 * e->window + 1) * kc);
                if (r.rc) break;
            }
            int rc = 0;
            if (kc == 32) rc = launch_attention<32>(qkv, mask, ek, ev, t1, B, T, C, e->heads, e->window, e->num_sms, s);
            else if (kc == 64) rc = launch_attention<64>(qkv, mask, ek, ev, t1, B, T, C, e->heads, e->window, e->num_sms, s);
            else if (kc == 96) rc = launch_attention<96>(qkv, mask, ek, ev, t1, B, T, C, e->heads, e->window, e->num_sms, s);
            else rc = launch_attention<128>(qkv, mask, ek, ev, t1, B, T, C, e->heads, e->window, e->num_sms, s);
            ++r.launches;
            r.mark("rel_attention");
            if (rc) return rc;
        }
        r.conv(a + ".conv_o", C, C, 1, t1, x, nullptr, xn, false);                        // x + attn(x)
        r.layernorm("encoder.norm_layers_1." + std::to_string(l), C, xn, mask, x, false);  // x <- LN1(x + y), masked
        r.conv(f + ".conv_1", C, F, e->ksize, x, nullptr, mask, w1, true);                 // relu(conv_1(x * mask)) * mask
        r.conv(f + ".conv_2", F, C, e->ksize, w1, x, nullptr, xn, false);                  // x + conv_2(.)
        r.layernorm("encoder.norm_layers_2." + std::to_string(l), C, xn, mask, x, false);  // x <- LN2(x + y), masked
    }
    if (r.rc) return r.rc;
    // ---- mu = proj_m(x) * mask (:329), in the reference's (B, n_feats, T) layout
    {
        const float* w = r.W("proj_m.wt");
        const float* b = r.P("proj_m.bias", (size_t)e->n_feats);
        if (r.rc) return r.rc;
        dim3 grid((T + 7) / 8, B);
        proj_nct_kernel<<<grid, (e->n_feats + 31) / 32 * 32, (size_t)C * 8 * 4, s>>>(x, w, b, mask, mu, T, C, e->n_feats);
        GTTS_CHECK_CUDA(cudaGetLastError());
        ++r.launches;
        r.mark("proj_mu");
    }
    // ---- duration predictor (:84-93): conv -> ReLU -> LayerNorm, twice, 1x1 proj, mask
    r.conv("proj_w.conv_1", C, Fdp, e->ksize, x, nullptr, nullptr, w1, true);
    r.layernorm("proj_w.norm_1", Fdp, w1, mask, w2, false);
    r.conv("proj_w.conv_2", Fdp, Fdp, e->ksize, w2, nullptr, nullptr, w1, true);
    r.layernorm("proj_w.norm_2", Fdp, w1, mask, w2, false);
    if (r.rc) return r.rc;
    {
        const float* w = r.P("proj_w.proj.weight", (size_t)Fdp);
        const float* b = r.P("proj_w.proj.bias", 1);
        if (r.rc) return r.rc;
        proj1_kernel<<<(unsigned)((n_pos + 7) / 8), 256, 0, s>>>(w2, w, b, mask, logw, (long)n_pos, Fdp);
        GTTS_CHECK_CUDA(cudaGetLastError());
        ++r.launches;
        r.mark("proj_logw");
    }
    r.report();
    e->launches_last_call = r.launches;
    return 0;
}

}  // namespace

// 1 if the last forward saw a token id outside [0, n_vocab) (synchronises the stream)
int text_encoder_check_tokens(TextEncoder* e, cudaStream_t s) {
    int h = 0;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(&h, e->status, 4, cudaMemcpyDeviceToHost, s));
    GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
    if (h) { set_error("text encoder: token id outside [0, n_vocab)"); return 2; }
    return 0;
}

}  // namespace gtts
