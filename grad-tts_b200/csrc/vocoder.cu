// HiFi-GAN generator (mel -> waveform), the step after the decoder in the reference's inference.py:97.
//
// Reference: hifi-gan/models.py:77-118 (Generator.forward), :14-51 (ResBlock1), :54-74 (ResBlock2); config
// checkpts/hifigan-config.json (V1: 512 initial channels, rates 8,8,2,2, kernels 16,16,4,4, resblock kernels 3,7,11 with
// dilations 1,3,5).  Weight norm is removed on the host side (the effective weights are what this file sees).
//
// B200 design.  Activations are channels-last (B, L, C) so that a 1-D conv is the H = 1 case of the implicit GEMM the decoder
// already has: every conv with >= 64 (padded) channels runs on the per-tap tcgen05 kernel of conv_tc.cu -- one TMA box of 128
// positions x 64 channels per (tap, channel chunk), the dilation is just the tap's box offset, the zero padding is TMA
// out-of-bounds fill -- with bf16 operands and fp32 accumulation in TMEM.  A transposed conv of stride u is u output phases of
// k/u taps each (four phases per launch).  The leaky ReLUs never get their own pass: a conv epilogue writes lrelu(v) instead of v
// (convs1 -> convs2), or v AND lrelu(v) (the residual stream and the next conv's input) -- ConvEpilogue::act_out / out2.
// Channel counts below 64 (the last stage, 32) are zero-padded to 64.  conv_pre (80 -> 512, 0.2 % of the FLOPs) and every
// layer in strict fp32 mode run on the CUDA-core implicit GEMM (conv_ffma.cu); conv_post (32 -> 1) + tanh is a small
// dedicated kernel.  One forward over a chunk of utterances is captured as a CUDA graph per (B, T).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "ops.h"
#include "vocoder_api.h"

namespace gtts {

namespace {

// ------------------------------------------------------------------------------------------------ small kernels
// (B, C, T) fp32 -> (B, T, Cp) activation, channels >= C zero
template <typename T>
__global__ void mel_to_nwc_kernel(const float* __restrict__ mel, T* __restrict__ out, int B, int C, int L, int Cp) {
    const long n = (long)B * L * Cp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long bl = i / Cp;
        const int l = (int)(bl % L), b = (int)(bl / L);
        const float v = c < C ? mel[((size_t)b * C + c) * L + l] : 0.f;
        Act<T>::st(out + i, v);
    }
}

// out = lrelu(scale * (a + b + c), slope); b, c optional.  8 elements per thread step.
template <typename T>
__global__ void sum_lrelu_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, T* __restrict__ out,
                                 long n8, float scale, float slope) {
    // four vectors of 8 per thread and round, their (up to twelve) loads issued before the first use: a streaming pass over
    // gigabytes with one load in flight per thread ran at 40 % of the HBM rate
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += 4 * stride) {
        float va[4][8], vb[4][8], vc[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long i = i0 + u * stride;
            if (i < n8) {
                Act<T>::load8(a + i * 8, va[u]);
                if (b) Act<T>::load8(b + i * 8, vb[u]);
                if (c) Act<T>::load8(c + i * 8, vc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long i = i0 + u * stride;
            if (i < n8) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float v = va[u][k];
                    if (b) v += vb[u][k];
                    if (c) v += vc[u][k];
                    v *= scale;
                    o[k] = v > 0.f ? v : v * slope;
                }
                Act<T>::store8(out + i * 8, o);
            }
        }
    }
}

// conv_post: (B, L, Cp) -> (B, 1, L) fp32, k taps, Cout = 1, + bias, tanh.  w: [k][C] fp32.  kC8 > 0: C = 8 * kC8 and k = 7 known at
// compile time (the reference configuration: every load of a position's window is issued before the first FMA).
template <typename T, int kC8>
__global__ void conv_post_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ out, int B, int L, int C, int Cp, int k) {
    extern __shared__ float s_w[];
    for (int i = threadIdx.x; i < k * C; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    const float b0 = bias[0];
    const long n = (long)B * L;
    const int pad = (k - 1) / 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const long b = i / L;
        float acc = b0;
        if (kC8 > 0) {
            float v[7][kC8 > 0 ? kC8 : 1][8];
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                const int il = l + t - 3;
                const bool ok = il >= 0 && il < L;
                const T* xp = x + ((size_t)b * L + (ok ? il : l)) * Cp;
#pragma unroll
                for (int c8 = 0; c8 < kC8; ++c8) {
                    Act<T>::load8(xp + c8 * 8, v[t][c8]);
                    if (!ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[t][c8][q] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 7; ++t)
#pragma unroll
                for (int c8 = 0; c8 < kC8; ++c8)
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc = fmaf(v[t][c8][q], s_w[t * (kC8 * 8) + c8 * 8 + q], acc);
        } else {
            for (int t = 0; t < k; ++t) {
                const int il = l + t - pad;
                if (il < 0 || il >= L) continue;
                const T* xp = x + ((size_t)b * L + il) * Cp;
                const float* wp = s_w + t * C;
                for (int c = 0; c < C; c += 8) {
                    float v[8];
                    Act<T>::load8(xp + c, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc = fmaf(v[q], wp[c + q], acc);
                }
            }
        }
        out[i] = tanhf(acc);
    }
}

// PyTorch conv weight -> K-major packed rows [tap][Cout_p] x [Cin_p]:  dst[(tap*Cout_p + co)*Cin_p + ci] = src[co*s_co + ci*s_ci + tap]
template <typename T>
__global__ void pack_w1d_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cout, int Cin, int k, int Cout_p, int Cin_p,
                                long s_co, long s_ci) {
    const long n = (long)k * Cout_p * Cin_p;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin_p);
        const long r = i / Cin_p;
        const int co = (int)(r % Cout_p), tap = (int)(r / Cout_p);
        const float v = (co < Cout && ci < Cin) ? src[co * s_co + ci * s_ci + tap] : 0.f;
        Act<T>::st(dst + i, v);
    }
}

// ---- position packing for stages with fewer than 64 channels ---------------------------------------------------------------
// A (B, L, C) tensor with C = 32 is, byte for byte, a (B, L/2, 64) tensor whose row j holds positions 2j and 2j+1.  Instead of
// zero-padding the channels to 64 (twice the HBM bytes of the longest, HBM-bound stage, and 4x its MMAs) the conv is run on that
// view: output row j, slot a (position P*j + a) reads input position P*j + a + s = row j + floor((a+s)/P), slot (a+s) mod P, so a
// k-tap conv with shifts s_t becomes a conv over rows with taps delta in {floor((a+s_t)/P)} and 64 x 64 block weights
//   W'[delta][(a, co)][(b, ci)] = W_t[co][ci]  where  a + s_t = delta*P + b.
// dst rows [delta index][P*C], cols [P*C]; deltas[] lists the row offsets in ascending order.
template <typename T>
__global__ void pack_w1d_packed_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int k, int dil, int P, int ndelta,
                                       const int* __restrict__ deltas) {
    const int PC = P * C;
    const long n = (long)ndelta * PC * PC;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int ri = (int)(i % PC);
        const long r = i / PC;
        const int ro = (int)(r % PC), di = (int)(r / PC);
        const int a = ro / C, co = ro % C, b = ri / C, ci = ri % C;
        const int sft = deltas[di] * P + b - a;                       // shift in positions
        float v = 0.f;
        if (sft % dil == 0) {
            const int t = sft / dil + (k - 1) / 2;
            if (t >= 0 && t < k) v = src[((size_t)co * C + ci) * k + t];
        }
        Act<T>::st(dst + i, v);
    }
}
// Transposed conv (stride u = m * P, kernel k, padding (k-u)/2) from an unpacked (B, Lin, Cin_p) input into the packed output:
// output position o = u*i + p lies in row m*i + p/P, slot p mod P; row phase pr = p / P.  dst rows [pr][delta index][P*C], cols
// [Cin_p]: W'[pr][delta][(a, co)][ci] = w[ci][co][r] with r = (pr*P + a) + pad - delta*u (ConvTranspose1d weight (Cin, Cout, k)).
template <typename T>
__global__ void pack_wT_packed_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cin, int Cin_p, int C, int k, int u, int P,
                                      int nphase, int ndelta, const int* __restrict__ deltas /* [nphase][ndelta] */) {
    const int PC = P * C, pad = (k - u) / 2;
    const long n = (long)nphase * ndelta * PC * Cin_p;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin_p);
        long r = i / Cin_p;
        const int ro = (int)(r % PC); r /= PC;
        const int di = (int)(r % ndelta), pr = (int)(r / ndelta);
        const int a = ro / C, co = ro % C;
        const int rr = pr * P + a + pad - deltas[pr * ndelta + di] * u;
        float v = 0.f;
        if (ci < Cin && rr >= 0 && rr < k) v = src[((size_t)ci * C + co) * k + rr];
        Act<T>::st(dst + i, v);
    }
}

int pad_ch(int c, int q) { return (c + q - 1) / q * q; }

// positions per row of a stage with C channels (1 = not packed)
int pack_factor(int C) { return (C < 64 && C >= 8 && 64 % C == 0) ? 64 / C : 1; }

// row offsets of a k-tap, dilation-d conv in the packed domain (ascending)
std::vector<int> packed_deltas(int k, int d, int P) {
    std::vector<int> out;
    for (int t = 0; t < k; ++t)
        for (int a = 0; a < P; ++a) {
            const int sft = (t - (k - 1) / 2) * d + a;
            const int dl = sft >= 0 ? sft / P : -((-sft + P - 1) / P);
            if (std::find(out.begin(), out.end(), dl) == out.end()) out.push_back(dl);
        }
    std::sort(out.begin(), out.end());
    return out;
}
// row offsets of row phase pr of the transposed conv into a packed stage; every phase is padded to the same count with a
// harmless duplicate... no: phases with fewer offsets get offsets whose weights are all zero (listed for a uniform tap count)
std::vector<int> packed_convT_deltas(int k, int u, int P, int pr) {
    std::vector<int> out;
    const int pad = (k - u) / 2;
    for (int a = 0; a < P; ++a) {
        const int p = pr * P + a;
        for (int r = (p + pad) % u; r < k; r += u) {
            const int dl = (p + pad - r) / u;
            if (std::find(out.begin(), out.end(), dl) == out.end()) out.push_back(dl);
        }
    }
    std::sort(out.begin(), out.end());
    return out;
}


struct ConvW {
    // effective PyTorch-layout fp32 weights as uploaded (device), and the packed forms (built lazily per activation type)
    float* w = nullptr; float* b = nullptr;
    size_t w_numel = 0, b_numel = 0;
    void* packed[2] = {nullptr, nullptr};     // [ACT_F32], [ACT_BF16]
    float* bias_p = nullptr;                  // bias zero-padded to Cout_p (position-packed stages: repeated per slot)
    size_t packed_n = 0;                      // elements of the packed buffers (cached plans hold pointers into them)
    std::vector<int> deltas;                  // position-packed layers: row offsets of the taps ([phase][ndelta] for the transposed conv)
    int ndelta = 0;
};

}  // namespace

struct VocoderPlan;

struct Vocoder {
    int device = 0, num_sms = 148;
    int resblock = 1, num_mels = 80, initial = 512;
    std::vector<int> rates, up_k, rb_k;
    std::vector<std::vector<int>> rb_d;
    std::map<std::string, ConvW> params;      // by reference key without the weight-norm suffix ("ups.0", "resblocks.3.convs1.2", ...)
    bool packed_valid = false;
    int max_chunk = 32;
    size_t workspace_budget = (size_t)24 << 30;
    int use_graph = 1;
    int force_ffma = 0;
    std::map<std::string, std::unique_ptr<VocoderPlan>> plans;
    std::vector<std::string> plan_order;
    long launches_last_call = 0;
    // Workspace pool shared by every plan of the handle.  Calls on one handle are ordered (same stream, or through done_ev), so
    // only one plan runs at a time and a new (B, T) reuses the blocks of earlier shapes: a server that sees a different mel length
    // on every call rebuilds descriptors and the graph, not gigabytes of cudaMalloc.  `busy` marks liveness during ONE plan build.
    struct Block { void* p; size_t bytes; bool busy; };
    std::vector<Block> pool;
    size_t pool_bytes = 0;
    int max_plans = 8;
    cudaEvent_t done_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_done = false;
    static size_t round_size(size_t bytes) {               // 1/8 of the power of two below: neighbouring lengths share size classes
        if (bytes < 4096) return 4096;
        size_t p2 = 1;
        while (p2 * 2 <= bytes) p2 *= 2;
        const size_t gran = p2 / 8;
        return (bytes + gran - 1) / gran * gran;
    }
    void* pool_alloc(size_t bytes) {
        bytes = round_size(bytes);
        int best = -1;
        for (int i = 0; i < (int)pool.size(); ++i)
            if (!pool[i].busy && pool[i].bytes >= bytes && (best < 0 || pool[i].bytes < pool[best].bytes)) best = i;
        if (best >= 0) { pool[best].busy = true; return pool[best].p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        pool.push_back(Block{p, bytes, true});
        pool_bytes += bytes;
        return p;
    }
    void pool_release(const void* p) { for (auto& b : pool) if (b.p == p) b.busy = false; }
    void pool_release_all() { for (auto& b : pool) b.busy = false; }
    void pool_free() { for (auto& b : pool) cudaFree(b.p); pool.clear(); pool_bytes = 0; }
    int total_up() const { int u = 1; for (int r : rates) u *= r; return u; }
    ~Vocoder();
};

struct VocoderPlan {
    Vocoder* v = nullptr;
    int B = 0, T = 0;
    ActKind kind = ACT_BF16;
    std::vector<TcConvPlan*> tc_plans;
    std::vector<std::function<int(cudaStream_t)>> ops;
    struct OpInfo { std::string name; double flops; };
    std::vector<OpInfo> info;
    float* mel_in = nullptr;                  // (B, num_mels, T) fp32, staging (pool)
    float* audio = nullptr;                   // (B, 1, T*up) fp32 (pool)
    cudaGraphExec_t graph_exec = nullptr;
    bool oom = false;
    size_t live_bytes = 0, peak_bytes = 0;
    std::map<const void*, size_t> sizes;
    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        void* p = v->pool_alloc(bytes);
        if (!p) { oom = true; return nullptr; }
        sizes[p] = bytes;
        live_bytes += bytes;
        peak_bytes = std::max(peak_bytes, live_bytes);
        return p;
    }
    void release(const void* p) {
        auto it = sizes.find(p);
        if (it != sizes.end()) { live_bytes -= it->second; sizes.erase(it); }
        v->pool_release(p);
    }
    ~VocoderPlan() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        for (auto* t : tc_plans) conv_tc_plan_destroy(t);
    }
};

Vocoder::~Vocoder() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    plans.clear();
    pool_free();
    if (done_ev) cudaEventDestroy(done_ev);
    for (auto& kv : params) {
        cudaFree(kv.second.w); cudaFree(kv.second.b); cudaFree(kv.second.packed[0]); cudaFree(kv.second.packed[1]); cudaFree(kv.second.bias_p);
    }
}

namespace {

size_t esz(ActKind k) { return k == ACT_F32 ? 4 : 2; }

ConvGeom geom_conv1d(int B, int L, int Cin_p, int Cout_p, int k, int dil) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = 1; g.Win = L; g.Hg = 1; g.Wg = L; g.Hout = 1; g.Wout = L;
    g.Cin0 = Cin_p; g.Cin1 = 0; g.Cout = Cout_p;
    g.ntaps = k; g.nphase = 1; g.stride = 1; g.out_step = 1;
    for (int t = 0; t < k; ++t) { g.dx[0][t] = (int8_t)((t - (k - 1) / 2) * dil); g.wrow[0][t] = t * Cout_p; }
    return g;
}

// conv over rows with explicit tap offsets (position-packed stages): tap t reads row j + deltas[t], weight rows t*Cout_p ..
ConvGeom geom_rows(int B, int L, int Cin_p, int Cout_p, const std::vector<int>& deltas) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = 1; g.Win = L; g.Hg = 1; g.Wg = L; g.Hout = 1; g.Wout = L;
    g.Cin0 = Cin_p; g.Cin1 = 0; g.Cout = Cout_p;
    g.ntaps = (int)deltas.size(); g.nphase = 1; g.stride = 1; g.out_step = 1;
    for (int t = 0; t < g.ntaps; ++t) { g.dx[0][t] = (int8_t)deltas[t]; g.wrow[0][t] = t * Cout_p; }
    return g;
}

// Transposed conv, stride u, kernel k, padding (k-u)/2, output phases p0 .. p0+np-1 (np <= 4): output o = j*u + p reads input
// j + d with weight tap r = p + pad - d*u for every r in [0, k) of that residue (hifi-gan/models.py:88-91)
bool geom_convT1d(ConvGeom* out, int B, int Lin, int Cin_p, int Cout_p, int k, int u, int p0, int np) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    const int pad = (k - u) / 2;
    g.B = B; g.Hin = 1; g.Win = Lin; g.Hg = 1; g.Wg = Lin; g.Hout = 1; g.Wout = Lin * u;
    g.Cin0 = Cin_p; g.Cin1 = 0; g.Cout = Cout_p;
    g.nphase = np; g.stride = 1; g.out_step = u;
    int ntaps = -1;
    for (int q = 0; q < np; ++q) {
        const int p = p0 + q;
        int n = 0;
        for (int r = (p + pad) % u; r < k; r += u) {
            if (n >= 16) return false;
            const int d = (p + pad - r) / u;             // exact: r = p + pad (mod u)
            if (d < -127 || d > 127) return false;
            g.dx[q][n] = (int8_t)d;
            g.wrow[q][n] = r * Cout_p;
            ++n;
        }
        if (ntaps >= 0 && n != ntaps) return false;      // every phase of one launch must have the same tap count
        ntaps = n;
        g.ox[q] = p;
    }
    if (ntaps <= 0) return false;
    g.ntaps = ntaps;
    *out = g;
    return true;
}

// Position packing is used for the last stage only (its consumers are the point-wise kernels and conv_post, which read the dense
// (B, L, C) tensor as it is) and when the stage's upsampling rate equals the packing factor, so that input position i of the
// transposed conv produces exactly row i.
bool stage_packed(const Vocoder* v, int i, int Co) {
    if (getenv("GTTS_VOC_NOPACK")) return false;
    const int P = pack_factor(Co);
    return P > 1 && i == (int)v->rates.size() - 1 && v->rates[i] == P;
}

struct Builder {
    Vocoder* v;
    VocoderPlan* pl;
    int B, T;
    ActKind kind;
    bool failed = false;

    void* act(long L, int Cp) {
        void* p = pl->alloc((size_t)B * L * Cp * esz(kind));
        if (!p) { failed = true; set_error("vocoder: out of device memory for the activation workspace"); }
        return p;
    }
    const ConvW* get(const std::string& name) {
        auto it = v->params.find(name);
        if (it == v->params.end() || !it->second.w || !it->second.b) {
            set_error("vocoder: parameter " + name + " (weight / bias) was not set");
            failed = true;
            return nullptr;
        }
        return &it->second;
    }
    bool tc_ok(const ConvGeom& g) const {
        return kind == ACT_BF16 && !v->force_ffma && (g.Cout == 64 || g.Cout == 128 || g.Cout == 256) && g.Cin0 % 64 == 0;
    }
    // one conv launch: v = conv(src) + bias (+ residual); out = act_out ? lrelu(v) : v; out2 = lrelu(v) if set
    void add_conv(const std::string& name, const ConvGeom& g, const void* src, const ConvW* w, int weight_rows, const void* residual,
                  void* out, int act_out, void* out2, float slope) {
        if (failed) return;
        ConvEpilogue e;
        memset(&e, 0, sizeof(e));
        e.bias = w->bias_p; e.residual = residual; e.out = out; e.act_out = act_out; e.out2 = out2; e.act_slope = slope;
        const double flops = 2.0 * (double)g.B * g.nphase * g.Wg * g.Cout * g.ntaps * g.Cin0;
        const void* wp = w->packed[kind];
        if (tc_ok(g)) {
            TcConvPlan* tp = conv_tc_plan_create(g, src, nullptr, wp, weight_rows, e, v->num_sms, 0);
            if (!tp) { failed = true; return; }
            pl->tc_plans.push_back(tp);
            pl->ops.push_back([tp](cudaStream_t s) { return conv_tc_launch(tp, s); });
        } else {
            ActKind k = kind;
            pl->ops.push_back([k, g, src, wp, e](cudaStream_t s) { return conv_ffma(k, g, src, nullptr, wp, e, s); });
        }
        pl->info.push_back({name, flops});
    }
    void add_sum(const void* a, const void* b, const void* c, void* out, long numel, float scale, float slope) {
        if (failed) return;
        ActKind k = kind;
        const long n8 = numel / 8;
        const int blocks = (int)std::min<long>((n8 + 255) / 256, 148L * 16);
        pl->ops.push_back([=](cudaStream_t s) {
            if (k == ACT_F32)
                sum_lrelu_kernel<float><<<blocks, 256, 0, s>>>((const float*)a, (const float*)b, (const float*)c, (float*)out, n8, scale, slope);
            else
                sum_lrelu_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                       (const __nv_bfloat16*)c, (__nv_bfloat16*)out, n8, scale, slope);
            GTTS_CHECK_CUDA(cudaGetLastError());
            return 0;
        });
        pl->info.push_back({"sum_lrelu", 0.0});
    }

    int build() {
        const int n_up = (int)v->rates.size(), n_rb = (int)v->rb_k.size();
        const int melp = pad_ch(v->num_mels, 32);
        // ---- input: (B, 80, T) fp32 -> (B, T, 96)
        void* x0 = act(T, melp);
        if (failed) return 4;
        {
            ActKind k = kind;
            const float* mel = pl->mel_in;
            const int Bb = B, C = v->num_mels, L = T;
            const long n = (long)B * T * melp;
            const int blocks = (int)std::min<long>((n + 255) / 256, 148L * 16);
            pl->ops.push_back([=](cudaStream_t s) {
                if (k == ACT_F32) mel_to_nwc_kernel<float><<<blocks, 256, 0, s>>>(mel, (float*)x0, Bb, C, L, melp);
                else mel_to_nwc_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(mel, (__nv_bfloat16*)x0, Bb, C, L, melp);
                GTTS_CHECK_CUDA(cudaGetLastError());
                return 0;
            });
            pl->info.push_back({"mel_to_nwc", 0.0});
        }
        // ---- conv_pre (k = 7) with the first leaky ReLU in its epilogue (models.py:101,103)
        int C = v->initial, Cp = std::max(64, pad_ch(C, 64));
        long L = T;
        void* lx = act(L, Cp);
        if (failed) return 4;
        {
            const ConvW* w = get("conv_pre");
            if (failed) return 3;
            add_conv("conv_pre", geom_conv1d(B, (int)L, melp, Cp, 7, 1), x0, w, 7 * Cp, nullptr, lx, 1, nullptr, 0.1f);
        }
        pl->release(x0);
        for (int i = 0; i < n_up; ++i) {
            const int u = v->rates[i], k = v->up_k[i];
            const int Co = C / 2;
            const bool packed = stage_packed(v, i, Co);           // < 64 channels: several positions per 64-channel row
            const int Cop = packed ? 64 : std::max(64, pad_ch(Co, 64));
            const long Lpos = L * u;                              // output positions
            const long Lo = packed ? Lpos / pack_factor(Co) : Lpos;   // rows of the stage's tensors
            // ---- ups[i]: x = convT(lrelu(x)); the epilogue also writes lrelu(x) for the resblocks' first convs
            void* x = act(Lo, Cop);
            void* lxo = act(Lo, Cop);
            if (failed) return 4;
            const ConvW* wu = get("ups." + std::to_string(i));
            if (failed) return 3;
            if (packed) {
                add_conv("ups" + std::to_string(i), geom_rows(B, (int)L, Cp, Cop, wu->deltas), lx, wu, wu->ndelta * Cop, nullptr, x, 0, lxo, 0.1f);
            }
            // phases grouped so that every launch has one tap count
            for (int p0 = packed ? u : 0; p0 < u;) {
                int np = std::min(4, u - p0);
                ConvGeom g;
                while (np > 0 && !geom_convT1d(&g, B, (int)L, Cp, Cop, k, u, p0, np)) --np;
                if (np == 0) { set_error("vocoder: unsupported transposed-conv geometry (kernel " + std::to_string(k) + ", stride " + std::to_string(u) + ")"); return 3; }
                add_conv("ups" + std::to_string(i), g, lx, wu, k * Cop, nullptr, x, 0, lxo, 0.1f);
                p0 += np;
            }
            pl->release(lx);
            // ---- resblocks (models.py:105-111): xs = sum_j resblock_j(x); x = xs / num_kernels
            std::vector<void*> ys;
            for (int j = 0; j < n_rb; ++j) {
                const int rk = v->rb_k[j];
                const std::vector<int>& dil = v->rb_d[j];
                const std::string base = "resblocks." + std::to_string(i * n_rb + j);
                const int nd = (int)dil.size();
                const void* cur = x;
                const void* lcur = lxo;
                void* pp[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // ping-pong (x, lrelu(x)) pairs of the residual stream
                for (int m = 0; m < nd; ++m) {
                    const bool last = m == nd - 1;
                    void* nx;
                    void* lnx = nullptr;
                    if (last) { nx = act(Lo, Cop); }
                    else {
                        void** q = pp[m & 1];
                        if (!q[0]) { q[0] = act(Lo, Cop); q[1] = act(Lo, Cop); }
                        nx = q[0]; lnx = q[1];
                    }
                    if (failed) return 4;
                    if (v->resblock == 1) {
                        // xt = c2(lrelu(c1(lrelu(x)))); x = xt + x   (ResBlock1.forward, models.py:38-45)
                        const ConvW* w1 = get(base + ".convs1." + std::to_string(m));
                        const ConvW* w2 = get(base + ".convs2." + std::to_string(m));
                        if (failed) return 3;
                        void* t = act(Lo, Cop);
                        if (failed) return 4;
                        if (packed) {
                            add_conv(base + ".c1", geom_rows(B, (int)Lo, Cop, Cop, w1->deltas), lcur, w1, w1->ndelta * Cop, nullptr, t, 1, nullptr, 0.1f);
                            add_conv(base + ".c2", geom_rows(B, (int)Lo, Cop, Cop, w2->deltas), t, w2, w2->ndelta * Cop, cur, nx, 0, lnx, 0.1f);
                        } else {
                            add_conv(base + ".c1", geom_conv1d(B, (int)Lo, Cop, Cop, rk, dil[m]), lcur, w1, rk * Cop, nullptr, t, 1, nullptr, 0.1f);
                            add_conv(base + ".c2", geom_conv1d(B, (int)Lo, Cop, Cop, rk, 1), t, w2, rk * Cop, cur, nx, 0, lnx, 0.1f);
                        }
                        pl->release(t);
                    } else {
                        // xt = c(lrelu(x)); x = xt + x   (ResBlock2.forward, models.py:66-70)
                        const ConvW* w1 = get(base + ".convs." + std::to_string(m));
                        if (failed) return 3;
                        if (packed) add_conv(base + ".c", geom_rows(B, (int)Lo, Cop, Cop, w1->deltas), lcur, w1, w1->ndelta * Cop, cur, nx, 0, lnx, 0.1f);
                        else add_conv(base + ".c", geom_conv1d(B, (int)Lo, Cop, Cop, rk, dil[m]), lcur, w1, rk * Cop, cur, nx, 0, lnx, 0.1f);
                    }
                    cur = nx; lcur = lnx;
                }
                for (auto& q : pp) if (q[0]) { pl->release(q[0]); pl->release(q[1]); }
                ys.push_back((void*)cur);
            }
            pl->release(lxo);
            // x = lrelu(xs / num_kernels): slope 0.1 before the next ups (models.py:103), PyTorch's default 0.01 before conv_post (:114)
            void* nlx = act(Lo, Cop);
            if (failed) return 4;
            const float slope = (i == n_up - 1) ? 0.01f : 0.1f;
            if (n_rb > 3) { set_error("vocoder: at most 3 resblock kernels"); return 3; }
            add_sum(ys[0], n_rb > 1 ? ys[1] : nullptr, n_rb > 2 ? ys[2] : nullptr, nlx, (long)B * Lo * Cop, 1.0f / (float)n_rb, slope);
            pl->release(x);
            for (void* y : ys) pl->release(y);
            lx = nlx; C = Co; L = Lpos;
            Cp = packed ? Co : Cop;                               // channels per POSITION of the stage's tensors (packed = dense)
        }
        // ---- conv_post (k = 7, Cout = 1) + tanh (models.py:115-116)
        {
            const ConvW* w = get("conv_post");
            if (failed) return 3;
            ActKind k = kind;
            const float* wp = (const float*)w->packed[ACT_F32];    // [7][C] fp32 (packed with Cout_p = 1)
            const float* bp = w->b;
            float* out = pl->audio;
            const int Bb = B, Ll = (int)L, Cc = pad_ch(C, 8), Cpp = Cp;
            const long n = (long)B * L;
            const int blocks = (int)std::min<long>((n + 127) / 128, 148L * 32);
            const void* src = lx;
            pl->ops.push_back([=](cudaStream_t s) {
                const size_t sm = (size_t)7 * Cc * 4;
                if (k == ACT_F32) conv_post_kernel<float, 0><<<blocks, 128, sm, s>>>((const float*)src, wp, bp, out, Bb, Ll, Cc, Cpp, 7);
                else if (Cc == 32) conv_post_kernel<__nv_bfloat16, 4><<<blocks, 128, sm, s>>>((const __nv_bfloat16*)src, wp, bp, out, Bb, Ll, Cc, Cpp, 7);
                else conv_post_kernel<__nv_bfloat16, 0><<<blocks, 128, sm, s>>>((const __nv_bfloat16*)src, wp, bp, out, Bb, Ll, Cc, Cpp, 7);
                GTTS_CHECK_CUDA(cudaGetLastError());
                return 0;
            });
            pl->info.push_back({"conv_post", 2.0 * (double)n * 7 * C});
        }
        pl->release(lx);
        return failed ? 3 : 0;
    }
};

// (re)build the packed weights of every layer (both activation types: the strict mode and conv_pre / conv_post use fp32)
int pack_all(Vocoder* v) {
    if (v->packed_valid) return 0;
    // packed buffers are reused across re-packs (cached plans hold pointers into them); a size change drops the plans
    auto ensure_packed = [&](ConvW& w, size_t n) -> int {
        if (w.packed_n == n && w.packed[ACT_F32] && w.packed[ACT_BF16]) return 0;
        GTTS_CHECK_CUDA(cudaDeviceSynchronize());
        v->plans.clear(); v->plan_order.clear();
        cudaFree(w.packed[ACT_F32]); cudaFree(w.packed[ACT_BF16]);
        w.packed[ACT_F32] = w.packed[ACT_BF16] = nullptr;
        GTTS_CHECK_CUDA(cudaMalloc(&w.packed[ACT_F32], n * 4));
        GTTS_CHECK_CUDA(cudaMalloc(&w.packed[ACT_BF16], n * 2));
        w.packed_n = n;
        return 0;
    };
    auto pack_conv = [&](const std::string& name, int Cout, int Cin, int k, int Cout_p, int Cin_p, bool transposed) -> int {
        auto it = v->params.find(name);
        if (it == v->params.end() || !it->second.w || !it->second.b) { set_error("vocoder: parameter " + name + " was not set"); return 3; }
        ConvW& w = it->second;
        GTTS_REQUIRE(w.w_numel == (size_t)Cout * Cin * k, "vocoder: weight has the wrong number of elements");
        GTTS_REQUIRE(w.b_numel == (size_t)Cout, "vocoder: bias has the wrong number of elements");
        const size_t n = (size_t)k * Cout_p * Cin_p;
        w.deltas.clear(); w.ndelta = 0;
        if (int rc = ensure_packed(w, n)) return rc;
        if (int rc = pack_conv1d_weight(ACT_F32, w.w, w.packed[ACT_F32], Cout, Cin, k, Cout_p, Cin_p, transposed, 0)) return rc;
        if (int rc = pack_conv1d_weight(ACT_BF16, w.w, w.packed[ACT_BF16], Cout, Cin, k, Cout_p, Cin_p, transposed, 0)) return rc;
        if (!w.bias_p) GTTS_CHECK_CUDA(cudaMalloc((void**)&w.bias_p, (size_t)std::max(Cout_p, 1) * 4));
        GTTS_CHECK_CUDA(cudaMemset(w.bias_p, 0, (size_t)std::max(Cout_p, 1) * 4));
        GTTS_CHECK_CUDA(cudaMemcpy(w.bias_p, w.b, (size_t)Cout * 4, cudaMemcpyDeviceToDevice));
        return 0;
    };
    const int melp = pad_ch(v->num_mels, 32);
    int C = v->initial, Cp = std::max(64, pad_ch(C, 64));
    if (int rc = pack_conv("conv_pre", C, v->num_mels, 7, Cp, melp, false)) return rc;
    const int n_rb = (int)v->rb_k.size();
    int* d_deltas = nullptr;
    GTTS_CHECK_CUDA(cudaMalloc((void**)&d_deltas, 64 * sizeof(int)));
    // position-packed layer: block weights over row offsets (see pack_w1d_packed_kernel)
    auto pack_packed = [&](const std::string& name, int Cc, int k, int dil, int P, int Cin_t, int Cin_p_t, int u_t) -> int {
        auto it = v->params.find(name);
        if (it == v->params.end() || !it->second.w || !it->second.b) { set_error("vocoder: parameter " + name + " was not set"); return 3; }
        ConvW& w = it->second;
        const bool transposed = u_t > 0;
        GTTS_REQUIRE(w.w_numel == (size_t)Cc * (transposed ? Cin_t : Cc) * k, "vocoder: weight has the wrong number of elements");
        GTTS_REQUIRE(w.b_numel == (size_t)Cc, "vocoder: bias has the wrong number of elements");
        w.deltas = transposed ? packed_convT_deltas(k, u_t, P, 0) : packed_deltas(k, dil, P);
        w.ndelta = (int)w.deltas.size();
        GTTS_REQUIRE(w.ndelta >= 1 && w.ndelta <= kMaxTaps, "vocoder: too many row offsets for a position-packed layer");
        for (int dl : w.deltas) GTTS_REQUIRE(dl >= -127 && dl <= 127, "vocoder: row offset out of range");
        GTTS_CHECK_CUDA(cudaMemcpy(d_deltas, w.deltas.data(), w.ndelta * sizeof(int), cudaMemcpyHostToDevice));
        const int PC = P * Cc, cols = transposed ? Cin_p_t : PC;
        const size_t n = (size_t)w.ndelta * PC * cols;
        if (int rc = ensure_packed(w, n)) return rc;
        const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
        if (transposed) {
            pack_wT_packed_kernel<float><<<blocks, 256>>>(w.w, (float*)w.packed[ACT_F32], Cin_t, Cin_p_t, Cc, k, u_t, P, 1, w.ndelta, d_deltas);
            pack_wT_packed_kernel<__nv_bfloat16><<<blocks, 256>>>(w.w, (__nv_bfloat16*)w.packed[ACT_BF16], Cin_t, Cin_p_t, Cc, k, u_t, P, 1, w.ndelta, d_deltas);
        } else {
            pack_w1d_packed_kernel<float><<<blocks, 256>>>(w.w, (float*)w.packed[ACT_F32], Cc, k, dil, P, w.ndelta, d_deltas);
            pack_w1d_packed_kernel<__nv_bfloat16><<<blocks, 256>>>(w.w, (__nv_bfloat16*)w.packed[ACT_BF16], Cc, k, dil, P, w.ndelta, d_deltas);
        }
        GTTS_CHECK_CUDA(cudaGetLastError());
        GTTS_CHECK_CUDA(cudaDeviceSynchronize());               // d_deltas is reused by the next layer
        if (!w.bias_p) GTTS_CHECK_CUDA(cudaMalloc((void**)&w.bias_p, (size_t)PC * 4));
        for (int a = 0; a < P; ++a) GTTS_CHECK_CUDA(cudaMemcpy(w.bias_p + (size_t)a * Cc, w.b, (size_t)Cc * 4, cudaMemcpyDeviceToDevice));
        return 0;
    };
    for (int i = 0; i < (int)v->rates.size(); ++i) {
        const int Co = C / 2;
        if (stage_packed(v, i, Co)) {
            const int P = pack_factor(Co);
            if (int rc = pack_packed("ups." + std::to_string(i), Co, v->up_k[i], 1, P, C, Cp, v->rates[i])) return rc;
            for (int j = 0; j < n_rb; ++j) {
                const std::string base = "resblocks." + std::to_string(i * n_rb + j);
                for (int m = 0; m < (int)v->rb_d[j].size(); ++m) {
                    if (v->resblock == 1) {
                        if (int rc = pack_packed(base + ".convs1." + std::to_string(m), Co, v->rb_k[j], v->rb_d[j][m], P, 0, 0, 0)) return rc;
                        if (int rc = pack_packed(base + ".convs2." + std::to_string(m), Co, v->rb_k[j], 1, P, 0, 0, 0)) return rc;
                    } else {
                        if (int rc = pack_packed(base + ".convs." + std::to_string(m), Co, v->rb_k[j], v->rb_d[j][m], P, 0, 0, 0)) return rc;
                    }
                }
            }
            C = Co; Cp = Co;
            continue;
        }
        const int Cop = std::max(64, pad_ch(Co, 64));
        if (int rc = pack_conv("ups." + std::to_string(i), Co, C, v->up_k[i], Cop, Cp, true)) return rc;
        for (int j = 0; j < n_rb; ++j) {
            const std::string base = "resblocks." + std::to_string(i * n_rb + j);
            for (int m = 0; m < (int)v->rb_d[j].size(); ++m) {
                if (v->resblock == 1) {
                    if (int rc = pack_conv(base + ".convs1." + std::to_string(m), Co, Co, v->rb_k[j], Cop, Cop, false)) return rc;
                    if (int rc = pack_conv(base + ".convs2." + std::to_string(m), Co, Co, v->rb_k[j], Cop, Cop, false)) return rc;
                } else {
                    if (int rc = pack_conv(base + ".convs." + std::to_string(m), Co, Co, v->rb_k[j], Cop, Cop, false)) return rc;
                }
            }
        }
        C = Co; Cp = Cop;
    }
    cudaFree(d_deltas);
    // conv_post: [7][pad8(C)] fp32 rows (Cout_p = 1)
    if (int rc = pack_conv("conv_post", 1, C, 7, 1, pad_ch(C, 8), false)) return rc;
    GTTS_CHECK_CUDA(cudaDeviceSynchronize());
    v->packed_valid = true;
    return 0;
}

int plan_create(Vocoder* v, int B, int T, ActKind kind, VocoderPlan** out) {
    std::unique_ptr<VocoderPlan> pl(new VocoderPlan());
    pl->v = v; pl->B = B; pl->T = T; pl->kind = kind;
    const size_t n_mel = (size_t)B * v->num_mels * T, n_audio = (size_t)B * T * v->total_up();
    v->pool_release_all();                               // nothing of an earlier plan is live while this one is being laid out
    pl->mel_in = (float*)pl->alloc(n_mel * 4);
    pl->audio = (float*)pl->alloc(n_audio * 4);
    if (!pl->mel_in || !pl->audio) { set_error("vocoder: out of device memory"); return 4; }
    Builder b{v, pl.get(), B, T, kind};
    int rc = b.build();
    if (rc) return pl->oom ? 4 : rc;
    cudaStream_t cs;
    GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    auto run_ops = [&](cudaStream_t s) -> int {
        for (auto& op : pl->ops)
            if (int r = op(s)) return r;
        return 0;
    };
    // dry run on a zero input (validates every launch), then capture
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->mel_in, 0, n_mel * 4, cs));
    rc = run_ops(cs);
    if (rc) { cudaStreamDestroy(cs); return rc; }
    GTTS_CHECK_CUDA(cudaStreamSynchronize(cs));
    GTTS_CHECK_CUDA(cudaGetLastError());
    if (v->use_graph) {
        cudaGraph_t graph = nullptr;
        // short inputs leave most SMs idle: programmatic dependent launch lets a conv's prologue and resident-weight loads run
        // under its predecessor (pdl_enabled, capi.cu); off for the throughput shapes, where it costs 2 %
        pdl_set_override((long)B * T <= 4096 && !getenv("GTTS_VOC_NOPDL") ? 1 : 0);
        GTTS_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = run_ops(cs);
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        pdl_set_override(0);
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(cs);
            if (!rc) { set_error(std::string("vocoder: graph capture failed: ") + cudaGetErrorString(ce)); rc = 1; }
            return rc;
        }
        ce = cudaGraphInstantiate(&pl->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            cudaStreamDestroy(cs);
            set_error(std::string("vocoder: cudaGraphInstantiate failed: ") + cudaGetErrorString(ce));
            return 1;
        }
    }
    GTTS_CHECK_CUDA(cudaStreamSynchronize(cs));
    cudaStreamDestroy(cs);
    *out = pl.release();
    return 0;
}

int get_plan(Vocoder* v, int B, int T, ActKind kind, cudaStream_t stream, VocoderPlan** out) {
    const std::string key = std::to_string(B) + ":" + std::to_string(T) + ":" + std::to_string((int)kind) + ":" + std::to_string(v->use_graph) +
                            std::to_string(v->force_ffma);
    auto it = v->plans.find(key);
    if (it != v->plans.end()) {
        v->plan_order.erase(std::find(v->plan_order.begin(), v->plan_order.end(), key));
        v->plan_order.push_back(key);
        *out = it->second.get();
        return 0;
    }
    // slow path: the dry run writes into pool blocks that queued work of this handle may still be using
    GTTS_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (v->has_done) GTTS_CHECK_CUDA(cudaEventSynchronize(v->done_ev));
    while ((int)v->plans.size() >= std::max(1, v->max_plans)) {          // descriptors + graph only; the memory is the shared pool
        v->plans.erase(v->plan_order.front());
        v->plan_order.erase(v->plan_order.begin());
    }
    VocoderPlan* pl = nullptr;
    int rc = plan_create(v, B, T, kind, &pl);
    if (rc == 4) {
        // out of device memory: give back everything this handle caches (plans and pool) and try once more
        GTTS_CHECK_CUDA(cudaDeviceSynchronize());
        v->plans.clear(); v->plan_order.clear();
        v->pool_free();
        rc = plan_create(v, B, T, kind, &pl);
    }
    if (rc) return rc;
    v->plans[key].reset(pl);
    v->plan_order.push_back(key);
    *out = pl;
    return 0;
}

// bytes of activation workspace one utterance of T frames needs (eleven live buffers at the last, longest stage)
size_t workspace_per_sample(const Vocoder* v, int T, ActKind kind) {
    int C = v->initial;
    long L = T;
    size_t peak = 0;
    for (size_t i = 0; i < v->rates.size(); ++i) {
        C /= 2; L *= v->rates[i];
        const size_t one = (size_t)L * (stage_packed(v, (int)i, C) ? C : std::max(64, pad_ch(C, 64))) * esz(kind);
        peak = std::max(peak, one * 11);
    }
    return peak;
}

}  // namespace

int pack_conv1d_weight(ActKind act, const float* src, void* dst, int Cout, int Cin, int k, int Cout_p, int Cin_p, bool transposed,
                       cudaStream_t s) {
    // Conv1d weight (Cout, Cin, k); ConvTranspose1d weight (Cin, Cout, k)
    const long s_co = transposed ? k : (long)Cin * k, s_ci = transposed ? (long)Cout * k : k;
    const size_t n = (size_t)k * Cout_p * Cin_p;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
    if (act == ACT_F32) pack_w1d_kernel<float><<<blocks, 256, 0, s>>>(src, (float*)dst, Cout, Cin, k, Cout_p, Cin_p, s_co, s_ci);
    else pack_w1d_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(src, (__nv_bfloat16*)dst, Cout, Cin, k, Cout_p, Cin_p, s_co, s_ci);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

Vocoder* vocoder_new(int resblock, int n_ups, const int* rates, const int* up_kernels, int initial_channel, int n_rb,
                     const int* rb_kernels, const int* rb_dilations, int n_dil, int num_mels, int device) {
    if (!(resblock == 1 || resblock == 2) || n_ups < 1 || n_ups > 8 || n_rb < 1 || n_rb > 3 || n_dil < 1 || n_dil > 8 || num_mels < 1 ||
        initial_channel < (1 << n_ups) || !rates || !up_kernels || !rb_kernels || !rb_dilations) {
        set_error("vocoder_new: bad configuration");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("vocoder_new: cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("vocoder_new: cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major != 10) { set_error("vocoder_new: this library is built for sm_100a (B200) only"); return nullptr; }
    Vocoder* v = new Vocoder();
    v->device = device; v->num_sms = prop.multiProcessorCount;
    v->resblock = resblock; v->num_mels = num_mels; v->initial = initial_channel;
    for (int i = 0; i < n_ups; ++i) {
        if (rates[i] < 1 || up_kernels[i] < rates[i] || ((up_kernels[i] - rates[i]) & 1)) {
            set_error("vocoder_new: upsample kernel must be >= its rate and differ from it by an even number");
            delete v;
            return nullptr;
        }
        v->rates.push_back(rates[i]); v->up_k.push_back(up_kernels[i]);
    }
    for (int j = 0; j < n_rb; ++j) {
        if (rb_kernels[j] < 1 || rb_kernels[j] > 15 || !(rb_kernels[j] & 1)) { set_error("vocoder_new: resblock kernels must be odd and <= 15"); delete v; return nullptr; }
        v->rb_k.push_back(rb_kernels[j]);
        std::vector<int> d;
        for (int m = 0; m < n_dil; ++m) {
            const int dd = rb_dilations[j * n_dil + m];
            if (dd < 1 || dd * (rb_kernels[j] - 1) / 2 > 127) { set_error("vocoder_new: dilation out of range"); delete v; return nullptr; }
            d.push_back(dd);
        }
        v->rb_d.push_back(d);
    }
    return v;
}

void vocoder_delete(Vocoder* v) { delete v; }
int vocoder_device(const Vocoder* v) { return v->device; }
int vocoder_total_upsampling(const Vocoder* v) { return v->total_up(); }
int vocoder_num_mels(const Vocoder* v) { return v->num_mels; }
long vocoder_launches_last_call(const Vocoder* v) { return v->launches_last_call; }

// name: reference state_dict key after remove_weight_norm ("conv_pre.weight", "ups.0.bias", "resblocks.4.convs1.2.weight", ...)
int vocoder_set_param(Vocoder* v, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(v && name && data && numel > 0, "vocoder_set_param: null argument");
    GTTS_CHECK_CUDA(cudaSetDevice(v->device));
    std::string n(name);
    const size_t dot = n.rfind('.');
    GTTS_REQUIRE(dot != std::string::npos, "vocoder_set_param: bad parameter name");
    const std::string leaf = n.substr(dot + 1), base = n.substr(0, dot);
    GTTS_REQUIRE(leaf == "weight" || leaf == "bias", "vocoder_set_param: parameter names end in .weight or .bias");
    ConvW& w = v->params[base];
    float** dst = leaf == "weight" ? &w.w : &w.b;
    size_t* cnt = leaf == "weight" ? &w.w_numel : &w.b_numel;
    GTTS_CHECK_CUDA(cudaDeviceSynchronize());                  // a running forward may still read the old packed weights
    if (*dst && *cnt != numel) { cudaFree(*dst); *dst = nullptr; }
    if (!*dst) GTTS_CHECK_CUDA(cudaMalloc((void**)dst, numel * 4));
    *cnt = numel;
    GTTS_CHECK_CUDA(cudaMemcpy(*dst, data, numel * 4, cudaMemcpyDefault));
    v->packed_valid = false;
    return 0;
}

int vocoder_set_option(Vocoder* v, const char* key, long long value) {
    GTTS_REQUIRE(v && key, "vocoder_set_option: null argument");
    const std::string k(key);
    if (k == "max_chunk") { GTTS_REQUIRE(value >= 1 && value <= 4096, "max_chunk out of range"); v->max_chunk = (int)value; }
    else if (k == "workspace_mb") { GTTS_REQUIRE(value >= 64, "workspace_mb out of range"); v->workspace_budget = (size_t)value << 20; }
    else if (k == "use_graph") v->use_graph = value != 0;
    else if (k == "max_plans") { GTTS_REQUIRE(value >= 1 && value <= 256, "max_plans out of range"); v->max_plans = (int)value; }
    else if (k == "force_ffma") v->force_ffma = value != 0;
    else { set_error("vocoder_set_option: unknown option " + k); return 2; }
    return 0;
}

// flags: bit0 = strict fp32 (CUDA-core FFMA convs), else bf16 activations with tcgen05 convs
int vocoder_forward(Vocoder* v, const float* mel, float* audio, int B, int T, int flags, cudaStream_t stream) {
    GTTS_REQUIRE(v && mel && audio, "vocoder_forward: null pointer");
    GTTS_REQUIRE(B >= 1 && T >= 1, "vocoder_forward: bad batch or length");
    GTTS_REQUIRE((long)T * v->total_up() < (1L << 30), "vocoder_forward: utterance too long");
    GTTS_CHECK_CUDA(cudaSetDevice(v->device));
    if (int rc = pack_all(v)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const size_t per = workspace_per_sample(v, T, kind);
    long chunk = (long)(v->workspace_budget / std::max<size_t>(per, 1));
    chunk = std::max(1L, std::min<long>(chunk, v->max_chunk));
    chunk = std::min<long>(chunk, B);
    const size_t up = (size_t)v->total_up();
    v->launches_last_call = 0;
    // calls on one handle are ordered even across streams (the plans share the pool)
    if (!v->done_ev) GTTS_CHECK_CUDA(cudaEventCreateWithFlags(&v->done_ev, cudaEventDisableTiming));
    if (v->has_done && v->last_stream != stream) GTTS_CHECK_CUDA(cudaStreamWaitEvent(stream, v->done_ev, 0));
    for (int b0 = 0; b0 < B; b0 += (int)chunk) {
        const int nb = std::min<int>((int)chunk, B - b0);
        VocoderPlan* pl = nullptr;
        if (int rc = get_plan(v, nb, T, kind, stream, &pl)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->mel_in, mel + (size_t)b0 * v->num_mels * T, (size_t)nb * v->num_mels * T * 4,
                                        cudaMemcpyDeviceToDevice, stream));
        if (pl->graph_exec) GTTS_CHECK_CUDA(cudaGraphLaunch(pl->graph_exec, stream));
        else
            for (auto& op : pl->ops)
                if (int r = op(stream)) return r;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(audio + (size_t)b0 * T * up, pl->audio, (size_t)nb * T * up * 4, cudaMemcpyDeviceToDevice, stream));
        v->launches_last_call += (long)pl->ops.size();
    }
    GTTS_CHECK_CUDA(cudaEventRecord(v->done_ev, stream));
    v->last_stream = stream;
    v->has_done = true;
    return 0;
}

// out[0] plans cached, [1] bytes of the shared workspace pool, [2] peak live bytes of the most recent plan
int vocoder_cache_info(const Vocoder* v, long long* out, int n) {
    GTTS_REQUIRE(v && out && n >= 3, "vocoder_cache_info: bad arguments");
    out[0] = (long long)v->plans.size();
    out[1] = (long long)v->pool_bytes;
    out[2] = 0;
    if (!v->plan_order.empty()) out[2] = (long long)v->plans.at(v->plan_order.back())->peak_bytes;
    return 0;
}

// per-launch CUDA-event times of one forward (text table into buf)
int vocoder_profile(Vocoder* v, int B, int T, int flags, char* buf, size_t buflen, cudaStream_t stream) {
    GTTS_REQUIRE(v && buf && buflen > 0, "vocoder_profile: null pointer");
    GTTS_CHECK_CUDA(cudaSetDevice(v->device));
    if (int rc = pack_all(v)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    VocoderPlan* pl = nullptr;
    if (int rc = get_plan(v, B, T, kind, stream, &pl)) return rc;
    cudaEvent_t e0, e1;
    GTTS_CHECK_CUDA(cudaEventCreate(&e0));
    GTTS_CHECK_CUDA(cudaEventCreate(&e1));
    std::string text;
    double total = 0.0;
    std::map<std::string, std::pair<double, double>> agg;
    std::vector<std::string> order;
    for (size_t i = 0; i < pl->ops.size(); ++i) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            GTTS_CHECK_CUDA(cudaEventRecord(e0, stream));
            if (int r = pl->ops[i](stream)) return r;
            GTTS_CHECK_CUDA(cudaEventRecord(e1, stream));
            GTTS_CHECK_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            GTTS_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        total += best;
        char line[256];
        snprintf(line, sizeof(line), "   %-34s %9.1f us %9.1f TFLOP/s\n", pl->info[i].name.c_str(), best * 1e3,
                 pl->info[i].flops / (best * 1e-3) / 1e12);
        text += line;
    }
    char head[128];
    snprintf(head, sizeof(head), "--- vocoder profile B=%d T=%d: total %.3f ms over %zu launches\n", B, T, total, pl->ops.size());
    text = head + text;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    snprintf(buf, buflen, "%s", text.c_str());
    return 0;
}

}  // namespace gtts
