// Decoder scheduler: owns the weights (reference state_dict names), packs them per precision mode, builds a
// per-(batch, T) plan = workspace + ordered kernel list for one GradLogPEstimator2d evaluation fused with the
// Euler update, captures it in a CUDA graph and replays it n_timesteps times.
//
// Reference: Diffusion.reverse_diffusion (model/diffusion.py:254-268), GradLogPEstimator2d (:128-216).
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <algorithm>

#include "common.cuh"
#include "decoder_api.h"
#include "ops.h"

namespace gtts {

namespace {

__global__ void schedule_kernel(float* t_tab, float* beta_tab, float* h_out, int n, double beta_min,
                                double beta_max) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *h_out = (float)(1.0 / (double)n);          // dxt * noise_t * h: python float -> fp32 (:266)
    if (i >= n) return;
    // h = 1.0 / n ; t = float(1.0 - (i + 0.5) * h)                               (:256,259)
    double h = 1.0 / (double)n;
    float t = (float)(1.0 - ((double)i + 0.5) * h);
    // noise_t = beta_min + (beta_max - beta_min) * t   (python scalars cast to fp32) (:219-224,262)
    float bt = __fadd_rn((float)beta_min, __fmul_rn((float)(beta_max - beta_min), t));
    t_tab[i] = t;
    beta_tab[i] = bt;
}

__global__ void scale_vec_kernel(const float* src, float* dst, float s, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = s * src[i];
}

__global__ void copy_cols_kernel(const float* src, float* dst, int rows, int cols, int dst_ld, int dst_off) {
    // dst[c][dst_off + r] = src[r][c]   (transpose a (rows, cols) Linear weight into a [cols][dst_ld] table)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int r = i / cols, c = i % cols;
    dst[(size_t)c * dst_ld + dst_off + r] = src[i];
}

struct DevMem {
    std::vector<void*> ptrs;
    size_t total = 0;
    // zeroing runs on the legacy default stream: whoever uses the buffer on another stream synchronises stream 0 first
    // (plan_create does, before its dry run)
    void* alloc(size_t bytes, bool zero = false) {
        void* p = nullptr;
        if (bytes == 0) bytes = 16;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (zero) cudaMemsetAsync(p, 0, bytes, 0);
        ptrs.push_back(p);
        total += bytes;
        return p;
    }
    ~DevMem() { for (void* p : ptrs) cudaFree(p); }
};

// Workspace pool of one decoder.  Plans never run concurrently (calls on one handle are ordered, see enter_call), and nothing a
// plan keeps in pooled memory outlives a call, so ALL plans draw their activation buffers from the same blocks: a new (B, T)
// shape reuses the blocks of earlier shapes wherever they are large enough, and only descriptors + graph are rebuilt.
// Sizes are rounded up to 1/8 of their power of two (<= 12.5 % slack) so that neighbouring T land in the same size classes.
struct Pool {
    struct Block { void* p; size_t bytes; };
    std::vector<Block> blocks;
    size_t total = 0;
    static size_t round_size(size_t bytes) {
        if (bytes < 4096) return 4096;
        size_t p2 = 1;
        while (p2 * 2 <= bytes) p2 *= 2;
        const size_t gran = p2 / 8;
        return (bytes + gran - 1) / gran * gran;
    }
    void clear() {
        for (auto& b : blocks) cudaFree(b.p);
        blocks.clear();
        total = 0;
    }
    ~Pool() { clear(); }
};

const char* kResnetNames[12] = {"downs.0.0", "downs.0.1", "downs.1.0", "downs.1.1", "downs.2.0", "downs.2.1",
                                "mid_block1", "mid_block2", "ups.0.0", "ups.0.1", "ups.1.0", "ups.1.1"};
const int kResnetCout[12] = {64, 64, 128, 128, 256, 256, 256, 256, 128, 128, 64, 64};

struct BlockW { void* w; const float* bias; const float* gamma; const float* beta; int cin, cout; };
struct ResnetW { BlockW b1, b2; void* wres; const float* bres; int cin, cout, tb_off; };
struct AttnW { void* wkv; const float* wq; const float* wout; float* gb; float g; int C; };
struct SampW { void* w; const float* bias; int C; };

struct Packed {
    ActKind kind;
    DevMem mem;
    TembWeights temb;
    const float *spk_w0t = nullptr, *spk_b0 = nullptr, *spk_w2t = nullptr, *spk_b2 = nullptr;
    const float *first_wT = nullptr, *first_b = nullptr, *fr_w = nullptr, *fr_b = nullptr;
    ResnetW res[12];
    AttnW attn[6];          // downs.0.2, downs.1.2, downs.2.2, mid_attn, ups.0.2, ups.1.2
    SampW down[2], up[2];
    BlockW final_block;
    const float* wf = nullptr;
    float bf = 0.f;
    // ---- data-gradient weights (built on the first VJP plan): the same convolutions transposed
    struct ResnetB { void* d1[2]; void* d2; void* rT[2]; };   // block1 dgrad per input source, block2 dgrad, res_conv^T per source
    struct AttnB { void* wq; void* woT; void* wqT; void* wkvT; };   // Wq [128][C], g * Wout^T [128][C], Wq^T [C][128], Wkv^T [C][256]
    std::map<const void*, void*> w3cache;      // fp32 packed weight -> its [wh wh wh wm wm wl] bf16 expansion (fp32_tc)
    bool bwd_ready = false;
    ResnetB bres[12];
    AttnB battn[6];
    void* bdown[2] = {nullptr, nullptr};
    void* bup[2] = {nullptr, nullptr};
    void* bfinal = nullptr;
};

}  // namespace

struct Plan;

struct Decoder {
    int n_spks_mode = 0;      // 0: 2 input channels; 1: speaker channel (n_spks > 1); 2: n_spks == -1 (spk_mlp unused)
    int cin_first = 2;
    int n_feats = 80, dim = 64;
    double beta_min = 0.05, beta_max = 20.0, pe_scale = 1000.0;
    int device = 0, num_sms = 148;
    int max_chunk = 64;
    bool use_graph = true;
    int conv_impl_bf16 = 1;   // 1: tcgen05, 0: FFMA (debug cross-check)
    int fused_attn = 1;       // bf16, C <= 128: fused k-projection + context kernel (no kv tensor)
    int fuse_gn = 0;          // 1: block2 convs apply block1's GroupNorm+Mish(+time bias, mask) on their operand tiles (no gn_apply
                              // pass).  Bitwise identical, but measured SLOWER (64->64 @L0: 435 us vs 134 + 103 us): GN+Mish at
                              // 5 TB/s already needs a whole SM's issue/MUFU capacity, four transform warps cannot supply it.
    int fuse_epi = 1;         // Block convs apply GroupNorm+Mish(+time bias / residual, mask) in their own epilogue (accumulators wait in
                              // TMEM for the sample's statistics, per-sample grid barrier): no raw tensor, no gn_apply pass.
                              // 0 = never, 1 = for plans of at most fuse_epi_max_b samples, 2 = always.  Measured (profiles/r02_gn_fusion.md):
                              // wins where launches are latency-bound (B = 1, T = 400: 21 us vs 17 + 8.6 us per Block) and loses on big
                              // batches (chunk 16 x 1720: 128->128 @h40 249 vs 109 + 57 us) because every sample costs one grid barrier.
    int fuse_epi_max_b = 2;
    int first_conv_mma = 1;   // bf16 mode: the first conv (2 | 3 fp32 planes -> 64 channels) on mma.sync with hi + lo split inputs and bf16
                              // weights (pointwise.cu: first_conv_mma_kernel); 0 = the FFMA kernel with fp32 weights (GTTS_FIRST_CONV_MMA=0)
    int pdl_small = 1;        // small-batch sampler plans are captured with programmatic dependent launch (GTTS_PDL_SMALL=0: off)
    int wgrad_tc = 1;         // bf16 training plan: weight gradients of the stride-1 convs on tcgen05 (wgrad_tc.cu); 0 = mma.sync kernel
    int side_lanes = 1;       // small batches: time-embedding MLP and res_conv on a parallel branch of the step graph (GTTS_SIDE=0: off)
    int fuse_async = 0;       // 1: GroupNorm+Mish(+time bias / residual) by asynchronous apply warps inside the Block conv (the raw tile goes
                              // through L2, the MMA pipeline never waits).  Correct and bitwise equal to the separate pass, but OFF: with the
                              // registers the conv leaves (72 per thread at 896 threads) eight apply warps keep only ~8 KB of loads in flight
                              // per SM and the kernel becomes bound by them (chunk 16 x 1720: 128->128 @h40 380 us vs 109 + 57 us;
                              // profiles/r02_gn_fusion.md).  0 = separate gn_apply pass
    int fp32_tc = 1;          // fp32 mode: 1 = convolutions on the tensor cores as six bf16 partial products of [hi|mid|lo] splits (fp32
                              // accuracy, ~1/6 of the bf16 MMA rate), 0 = CUDA-core FFMA convolutions
    int halo_mode = 2;        // 3x3 convs: 0 = per-tap boxes, 1/2 = halo box (18x16 / 18x10) + shifted descriptor views
    std::map<std::string, float*> params;
    std::map<std::string, size_t> param_numel;
    DevMem param_mem;
    std::unique_ptr<Packed> packed[2];
    struct CachedPlan { Plan* plan; unsigned long long last_use; };
    std::map<std::string, CachedPlan> plans;
    unsigned long long use_clock = 0;
    int max_plans = 24;       // LRU bound on cached (descriptors + graph) plans; their activation memory is the shared pool
    Pool pool;
    // calls on one handle are ordered across streams: the event is recorded at the end of every call, and a call on a different
    // stream waits for it (plans share the pool, and each plan's state buffers are reused from call to call)
    cudaEvent_t done_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_done = false;
    long launches_last_call = 0;
    long plans_created = 0;
    // parameter gradients of the last estimator_backward call (flat, fp32, PyTorch layouts; offsets in pg_layout)
    std::map<std::string, std::pair<size_t, size_t>> pg_layout;   // full name -> (offset, numel)
    size_t pg_floats = 0;
    float* pg_accum = nullptr;
    ~Decoder();
};

namespace {

size_t esize(ActKind k) { return k == ACT_F32 ? 4 : 2; }

int pack_weights(Decoder* d, ActKind kind) {
    if (d->packed[kind]) return 0;
    std::unique_ptr<Packed> P(new Packed());
    P->kind = kind;
    cudaStream_t s = 0;
    auto get = [&](const std::string& n) -> const float* {
        auto it = d->params.find("estimator." + n);
        return it == d->params.end() ? nullptr : it->second;
    };
    auto numel = [&](const std::string& n) -> size_t {
        auto it = d->param_numel.find("estimator." + n);
        return it == d->param_numel.end() ? 0 : it->second;
    };
#define NEED(ptr, name)                                                                 \
    do {                                                                                \
        if ((ptr) == nullptr) { set_error(std::string("missing parameter estimator.") + (name)); return 3; } \
    } while (0)
    // ---- time MLP
    {
        const float *w0 = get("mlp.0.weight"), *b0 = get("mlp.0.bias"), *w2 = get("mlp.2.weight"), *b2 = get("mlp.2.bias");
        NEED(w0, "mlp.0.weight"); NEED(b0, "mlp.0.bias"); NEED(w2, "mlp.2.weight"); NEED(b2, "mlp.2.bias");
        float* w0t = (float*)P->mem.alloc(64 * 256 * 4);
        float* w2t = (float*)P->mem.alloc(256 * 64 * 4);
        float* wbt = (float*)P->mem.alloc(64 * 1792 * 4);
        float* bb = (float*)P->mem.alloc(1792 * 4);
        if (!w0t || !w2t || !wbt || !bb) { set_error("out of device memory packing weights"); return 4; }
        if (transpose_2d(w0, w0t, 256, 64, s)) return 1;
        if (transpose_2d(w2, w2t, 64, 256, s)) return 1;
        int off = 0;
        for (int r = 0; r < 12; ++r) {
            std::string n = std::string(kResnetNames[r]) + ".mlp.1";
            const float *w = get(n + ".weight"), *b = get(n + ".bias");
            NEED(w, n + ".weight"); NEED(b, n + ".bias");
            const int co = kResnetCout[r];
            copy_cols_kernel<<<(co * 64 + 255) / 256, 256, 0, s>>>(w, wbt, co, 64, 1792, off);
            GTTS_CHECK_CUDA(cudaMemcpyAsync(bb + off, b, co * 4, cudaMemcpyDeviceToDevice, s));
            P->res[r].tb_off = off;
            off += co;
        }
        P->temb = TembWeights{w0t, b0, w2t, b2, wbt, bb};
    }
    // ---- speaker MLP
    if (d->n_spks_mode == 1) {
        const float *w0 = get("spk_mlp.0.weight"), *b0 = get("spk_mlp.0.bias"), *w2 = get("spk_mlp.2.weight"),
                    *b2 = get("spk_mlp.2.bias");
        NEED(w0, "spk_mlp.0.weight"); NEED(b0, "spk_mlp.0.bias"); NEED(w2, "spk_mlp.2.weight"); NEED(b2, "spk_mlp.2.bias");
        float* w0t = (float*)P->mem.alloc(64 * 256 * 4);
        float* w2t = (float*)P->mem.alloc(256 * d->n_feats * 4);
        if (transpose_2d(w0, w0t, 256, 64, s)) return 1;
        if (transpose_2d(w2, w2t, d->n_feats, 256, s)) return 1;
        P->spk_w0t = w0t; P->spk_b0 = b0; P->spk_w2t = w2t; P->spk_b2 = b2;
    }
    // ---- blocks
    auto pack_block = [&](const std::string& n, int cin, int cout, BlockW* bw, bool first) -> int {
        const float *w = get(n + ".block.0.weight"), *b = get(n + ".block.0.bias"), *ga = get(n + ".block.1.weight"),
                    *be = get(n + ".block.1.bias");
        NEED(w, n + ".block.0.weight"); NEED(b, n + ".block.0.bias"); NEED(ga, n + ".block.1.weight"); NEED(be, n + ".block.1.bias");
        if (numel(n + ".block.0.weight") != (size_t)cout * cin * 9) { set_error("bad shape for estimator." + n + ".block.0.weight"); return 3; }
        bw->bias = b; bw->gamma = ga; bw->beta = be; bw->cin = cin; bw->cout = cout;
        if (first) {
            float* wt = (float*)P->mem.alloc((size_t)cin * 9 * 64 * 4);
            if (transpose_2d(w, wt, 64, cin * 9, s)) return 1;
            bw->w = wt;
        } else {
            bw->w = P->mem.alloc((size_t)9 * cout * cin * esize(kind));
            if (!bw->w) { set_error("out of device memory packing weights"); return 4; }
            if (pack_conv_weight(kind, w, bw->w, cout, cin, 3, 3, s)) return 1;
        }
        return 0;
    };
    const int res_cin[12] = {d->cin_first, 64, 64, 128, 128, 256, 256, 256, 512, 128, 256, 64};
    for (int r = 0; r < 12; ++r) {
        ResnetW& R = P->res[r];
        const std::string n = kResnetNames[r];
        R.cin = res_cin[r]; R.cout = kResnetCout[r];
        if (int rc = pack_block(n + ".block1", R.cin, R.cout, &R.b1, r == 0)) return rc;
        if (int rc = pack_block(n + ".block2", R.cout, R.cout, &R.b2, false)) return rc;
        R.wres = nullptr; R.bres = nullptr;
        if (R.cin != R.cout) {
            const float *w = get(n + ".res_conv.weight"), *b = get(n + ".res_conv.bias");
            NEED(w, n + ".res_conv.weight"); NEED(b, n + ".res_conv.bias");
            R.bres = b;
            if (r == 0) {
                P->fr_w = w; P->fr_b = b;            // (64, cin) row-major, used inline by gn_apply
            } else {
                R.wres = P->mem.alloc((size_t)R.cout * R.cin * esize(kind));
                if (pack_conv_weight(kind, w, R.wres, R.cout, R.cin, 1, 1, s)) return 1;
            }
        }
    }
    P->first_wT = (const float*)P->res[0].b1.w;
    P->first_b = P->res[0].b1.bias;
    // ---- attention
    const char* attn_names[6] = {"downs.0.2", "downs.1.2", "downs.2.2", "mid_attn", "ups.0.2", "ups.1.2"};
    const int attn_c[6] = {64, 128, 256, 256, 128, 64};
    for (int a = 0; a < 6; ++a) {
        AttnW& A = P->attn[a];
        const std::string n = attn_names[a];
        const float *g = get(n + ".fn.g"), *wqkv = get(n + ".fn.fn.to_qkv.weight"), *wo = get(n + ".fn.fn.to_out.weight"),
                    *bo = get(n + ".fn.fn.to_out.bias");
        NEED(g, n + ".fn.g"); NEED(wqkv, n + ".fn.fn.to_qkv.weight"); NEED(wo, n + ".fn.fn.to_out.weight"); NEED(bo, n + ".fn.fn.to_out.bias");
        A.C = attn_c[a];
        if (numel(n + ".fn.fn.to_qkv.weight") != (size_t)384 * A.C) { set_error("bad shape for estimator." + n + ".fn.fn.to_qkv.weight"); return 3; }
        GTTS_CHECK_CUDA(cudaMemcpy(&A.g, g, 4, cudaMemcpyDeviceToHost));
        A.wq = wqkv;                                   // rows [0,128): q = 'b (qkv heads c) h w' (:91-92)
        A.wout = wo;                                   // (C, 128)
        A.wkv = P->mem.alloc((size_t)256 * A.C * esize(kind));
        if (pack_conv_weight(kind, wqkv + (size_t)128 * A.C, A.wkv, 256, A.C, 1, 1, s)) return 1;
        A.gb = (float*)P->mem.alloc(A.C * 4);
        scale_vec_kernel<<<1, 256, 0, s>>>(bo, A.gb, A.g, A.C);
    }
    // ---- down / up sampling
    for (int i = 0; i < 2; ++i) {
        const int C = i == 0 ? 64 : 128;
        std::string n = "downs." + std::to_string(i) + ".3.conv";
        const float *w = get(n + ".weight"), *b = get(n + ".bias");
        NEED(w, n + ".weight"); NEED(b, n + ".bias");
        P->down[i].C = C; P->down[i].bias = b;
        P->down[i].w = P->mem.alloc((size_t)9 * C * C * esize(kind));
        if (pack_conv_weight(kind, w, P->down[i].w, C, C, 3, 3, s)) return 1;
    }
    for (int i = 0; i < 2; ++i) {
        const int C = i == 0 ? 128 : 64;
        std::string n = "ups." + std::to_string(i) + ".3.conv";
        const float *w = get(n + ".weight"), *b = get(n + ".bias");
        NEED(w, n + ".weight"); NEED(b, n + ".bias");
        if (numel(n + ".weight") != (size_t)16 * C * C) { set_error("bad shape for estimator." + n + ".weight"); return 3; }
        P->up[i].C = C; P->up[i].bias = b;
        P->up[i].w = P->mem.alloc((size_t)16 * C * C * esize(kind));
        if (pack_convT_weight(kind, w, P->up[i].w, C, s)) return 1;
    }
    // ---- final
    if (int rc = pack_block("final_block", 64, 64, &P->final_block, false)) return rc;
    {
        const float *w = get("final_conv.weight"), *b = get("final_conv.bias");
        NEED(w, "final_conv.weight"); NEED(b, "final_conv.bias");
        P->wf = w;
        GTTS_CHECK_CUDA(cudaMemcpy(&P->bf, b, 4, cudaMemcpyDeviceToHost));
    }
#undef NEED
    GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
    GTTS_CHECK_CUDA(cudaGetLastError());
    d->packed[kind] = std::move(P);
    return 0;
}

// Transposed / flipped copies of every convolution weight for the backward pass w.r.t. the input (backward.cu).
int pack_backward_weights(Decoder* d, ActKind kind) {
    Packed* P = d->packed[kind].get();
    if (P->bwd_ready) return 0;
    cudaStream_t s = 0;
    auto get = [&](const std::string& n) -> const float* {
        auto it = d->params.find("estimator." + n);
        return it == d->params.end() ? nullptr : it->second;
    };
    const int res_c0[12] = {0, 64, 64, 128, 128, 256, 256, 256, 256, 128, 128, 64};      // channels of the first / second input source
    const int res_c1[12] = {0, 0, 0, 0, 0, 0, 0, 0, 256, 0, 128, 0};
    const size_t es = esize(kind);
    for (int r = 0; r < 12; ++r) {
        Packed::ResnetB& Bw = P->bres[r];
        Bw.d1[0] = Bw.d1[1] = Bw.d2 = Bw.rT[0] = Bw.rT[1] = nullptr;
        const ResnetW& R = P->res[r];
        const std::string n = kResnetNames[r];
        const float* w1 = get(n + ".block1.block.0.weight");
        const float* w2 = get(n + ".block2.block.0.weight");
        const float* wr = get(n + ".res_conv.weight");
        Bw.d2 = P->mem.alloc((size_t)9 * R.cout * R.cout * es);
        if (!Bw.d2) { set_error("out of device memory packing backward weights"); return 4; }
        if (pack_dgrad3(kind, w2, Bw.d2, R.cout, R.cout, 0, R.cout, s)) return 1;
        if (r == 0) continue;                                  // first conv / first res_conv: first_bwd reads the forward weights
        const int cs[2] = {res_c0[r], res_c1[r]};
        int off = 0;
        for (int k = 0; k < 2; ++k) {
            if (cs[k] == 0) continue;
            Bw.d1[k] = P->mem.alloc((size_t)9 * cs[k] * R.cout * es);
            if (!Bw.d1[k]) { set_error("out of device memory packing backward weights"); return 4; }
            if (pack_dgrad3(kind, w1, Bw.d1[k], R.cout, R.cin, off, cs[k], s)) return 1;
            if (wr) {
                Bw.rT[k] = P->mem.alloc((size_t)cs[k] * R.cout * es);
                if (pack_t1(kind, wr, Bw.rT[k], R.cout, R.cin, off, cs[k], 1.0f, s)) return 1;
            }
            off += cs[k];
        }
    }
    const char* attn_names[6] = {"downs.0.2", "downs.1.2", "downs.2.2", "mid_attn", "ups.0.2", "ups.1.2"};
    for (int a = 0; a < 6; ++a) {
        const AttnW& A = P->attn[a];
        Packed::AttnB& Bw = P->battn[a];
        const std::string n = attn_names[a];
        const float* wqkv = get(n + ".fn.fn.to_qkv.weight");
        Bw.wq = P->mem.alloc((size_t)128 * A.C * es);
        if (!Bw.wq || pack_conv_weight(kind, wqkv, Bw.wq, 128, A.C, 1, 1, s)) return 1;
        Bw.woT = P->mem.alloc((size_t)128 * A.C * es);
        Bw.wqT = P->mem.alloc((size_t)A.C * 128 * es);
        Bw.wkvT = P->mem.alloc((size_t)A.C * 256 * es);
        if (!Bw.woT || !Bw.wqT || !Bw.wkvT) { set_error("out of device memory packing backward weights"); return 4; }
        if (pack_t1(kind, A.wout, Bw.woT, A.C, 128, 0, 128, A.g, s)) return 1;                       // rows e (128), cols c: g * Wout[c][e]
        if (pack_t1(kind, wqkv, Bw.wqT, 128, A.C, 0, A.C, 1.0f, s)) return 1;                        // rows c, cols hd (q rows of to_qkv)
        if (pack_t1(kind, wqkv + (size_t)128 * A.C, Bw.wkvT, 256, A.C, 0, A.C, 1.0f, s)) return 1;   // rows c, cols k|v
    }
    for (int i = 0; i < 2; ++i) {
        const int Cd = i == 0 ? 64 : 128, Cu = i == 0 ? 128 : 64;
        P->bdown[i] = P->mem.alloc((size_t)16 * Cd * Cd * es);
        P->bup[i] = P->mem.alloc((size_t)16 * Cu * Cu * es);
        if (!P->bdown[i] || !P->bup[i]) { set_error("out of device memory packing backward weights"); return 4; }
        if (pack_down_dgrad(kind, get("downs." + std::to_string(i) + ".3.conv.weight"), P->bdown[i], Cd, s)) return 1;
        if (pack_up_dgrad(kind, get("ups." + std::to_string(i) + ".3.conv.weight"), P->bup[i], Cu, s)) return 1;
    }
    P->bfinal = P->mem.alloc((size_t)9 * 64 * 64 * es);
    if (!P->bfinal) { set_error("out of device memory packing backward weights"); return 4; }
    if (pack_dgrad3(kind, get("final_block.block.0.weight"), P->bfinal, 64, 64, 0, 64, s)) return 1;
    GTTS_CHECK_CUDA(cudaStreamSynchronize(s));
    GTTS_CHECK_CUDA(cudaGetLastError());
    P->bwd_ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ geometry helpers
ConvGeom geom_base(int B, int Hin, int Win, int Cin0, int Cin1, int Cout) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Hin = Hin; g.Win = Win; g.Hg = Hin; g.Wg = Win; g.Hout = Hin; g.Wout = Win;
    g.Cin0 = Cin0; g.Cin1 = Cin1; g.Cout = Cout;
    g.ntaps = 1; g.nphase = 1; g.stride = 1; g.out_step = 1;
    return g;
}
ConvGeom geom_3x3(int B, int H, int W, int Cin0, int Cin1, int Cout, int stride) {
    ConvGeom g = geom_base(B, H, W, Cin0, Cin1, Cout);
    g.ntaps = 9; g.stride = stride;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
            g.dy[0][ky * 3 + kx] = (int8_t)(ky - 1);
            g.dx[0][ky * 3 + kx] = (int8_t)(kx - 1);
            g.wrow[0][ky * 3 + kx] = (ky * 3 + kx) * Cout;
        }
    if (stride == 2) { g.Hg = H / 2; g.Wg = W / 2; g.Hout = H / 2; g.Wout = W / 2; }
    return g;
}
ConvGeom geom_1x1(int B, int H, int W, int Cin0, int Cin1, int Cout, int w_batch_rows) {
    ConvGeom g = geom_base(B, H, W, Cin0, Cin1, Cout);
    g.w_batch_rows = w_batch_rows;
    return g;
}
ConvGeom geom_convT(int B, int H, int W, int C) {
    ConvGeom g = geom_base(B, H, W, C, 0, C);
    g.ntaps = 4; g.nphase = 4; g.out_step = 2; g.Hout = 2 * H; g.Wout = 2 * W;
    const int dd[2][2] = {{0, -1}, {0, 1}};           // parity 0: in j, j-1 ; parity 1: in j, j+1
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int ph = py * 2 + px;
            g.oy[ph] = py; g.ox[ph] = px;
            for (int ty = 0; ty < 2; ++ty)
                for (int tx = 0; tx < 2; ++tx) {
                    const int t = ty * 2 + tx;
                    g.dy[ph][t] = (int8_t)dd[py][ty];
                    g.dx[ph][t] = (int8_t)dd[px][tx];
                    g.wrow[ph][t] = (ph * 4 + t) * C;
                }
        }
    return g;
}

// data gradient of the 3x3 stride-2 Downsample: gradient at (Hc, Wc) -> gradient at (2Hc, 2Wc), four output parities
ConvGeom geom_down_dgrad(int B, int Hc, int Wc, int C) {
    ConvGeom g = geom_base(B, Hc, Wc, C, 0, C);
    g.ntaps = 4; g.nphase = 4; g.out_step = 2; g.Hout = 2 * Hc; g.Wout = 2 * Wc;
    const int dd[2][2] = {{0, 0}, {1, 0}};            // parity 0: tap 0 reads j (tap 1 has zero weights); parity 1: j+1, j
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const int ph = py * 2 + px;
            g.oy[ph] = py; g.ox[ph] = px;
            for (int ty = 0; ty < 2; ++ty)
                for (int tx = 0; tx < 2; ++tx) {
                    const int t = ty * 2 + tx;
                    g.dy[ph][t] = (int8_t)dd[py][ty];
                    g.dx[ph][t] = (int8_t)dd[px][tx];
                    g.wrow[ph][t] = (ph * 4 + t) * C;
                }
        }
    return g;
}
// data gradient of the 4x4 stride-2 transposed conv (Upsample): a 16-tap stride-2 conv from (2Hc, 2Wc) down to (Hc, Wc)
ConvGeom geom_up_dgrad(int B, int Hc, int Wc, int C) {
    ConvGeom g = geom_base(B, 2 * Hc, 2 * Wc, C, 0, C);
    g.Hg = Hc; g.Wg = Wc; g.Hout = Hc; g.Wout = Wc;
    g.ntaps = 16; g.stride = 2;
    for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
            const int t = ky * 4 + kx;
            g.dy[0][t] = (int8_t)(ky - 1);
            g.dx[0][t] = (int8_t)(kx - 1);
            g.wrow[0][t] = t * C;
        }
    return g;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ plan
struct Plan {
    Decoder* d;
    ActKind kind;
    bool strict;
    int B, T;
    bool est_mode, sde;
    DevMem mem;
    std::vector<TcConvPlan*> tc_plans;
    std::vector<WgradTcPlan*> wg_plans;
    std::vector<std::function<int(cudaStream_t)>> ops;
    struct OpInfo { std::string name; int is_conv; double flops; double bytes; };
    std::vector<OpInfo> info;
    // Small batches leave most SMs idle, so ops that are off the critical path (the time-embedding MLP, a ResnetBlock's 1x1
    // res_conv) are captured on a side stream -- a parallel branch of the step graph -- and joined before their first consumer.
    // lane[i] = 1: op i forks from the main branch at its position; join[i] = 1: the main branch waits for every open side op first.
    // Eager replays (dry run, profile, no-graph mode) just run the list in order.
    std::vector<uint8_t> lane, join;
    int cur_lane = 0;
    bool cur_join = false, side_open = false;
    void push(const std::string& name, int is_conv, double flops, double bytes, std::function<int(cudaStream_t)> fn) {
        ops.push_back(std::move(fn));
        info.push_back(OpInfo{name, is_conv, flops, bytes});
        lane.push_back((uint8_t)cur_lane);
        join.push_back(cur_join ? 1 : 0);
        if (cur_lane) side_open = true;
        if (cur_join) side_open = false;
        cur_lane = 0; cur_join = false;
        kernels_per_step++;
    }
    void on_side() { cur_lane = 1; }                       // the next push goes to the side branch
    void join_side() { if (side_open) cur_join = true; }   // the next push waits for the side branch (no-op if nothing is open)
    // plan-owned I/O
    float *xt = nullptr, *mu = nullptr, *m0 = nullptr, *m1 = nullptr, *m2 = nullptr, *splane = nullptr, *spk = nullptr;
    float *tb = nullptr, *t_tab = nullptr, *beta_tab = nullptr, *t_per_sample = nullptr, *score = nullptr;
    int* step = nullptr;
    NoiseSlot* noise_slot = nullptr;
    float* h_dev = nullptr;
    int max_steps = 4096;
    bool vjp = false;             // forward (est mode, every intermediate kept) followed by the backward pass w.r.t. x
    bool pgrads = false;          // ... and the gradients of every parameter, of mu, of the speaker plane and of the time biases
    float *vin = nullptr, *gx = nullptr;      // (B,80,T) cotangent in, J^T v out
    float *gmu = nullptr, *gs_pix = nullptr, *gtb = nullptr, *pg_flat = nullptr;   // (B,80,T), (B,80,T), (B,1792), flat parameter grads
    size_t pooled_bytes = 0;      // peak of live pooled activation memory of this plan (after liveness reuse)
    size_t unpooled_bytes = 0;    // what one-buffer-per-tensor allocation would have needed (reported for comparison)
    float* partials = nullptr;
    unsigned int* counters = nullptr;
    size_t partial_slots = 0;
    cudaGraphExec_t graph_exec = nullptr;
    bool warmed = false;
    long kernels_per_step = 0;
    ~Plan() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        for (auto* p : tc_plans) conv_tc_plan_destroy(p);
        for (auto* p : wg_plans) wgrad_tc_plan_destroy(p);
    }
};

Decoder::~Decoder() {
    for (auto& kv : plans) delete kv.second.plan;
    if (done_ev) cudaEventDestroy(done_ev);
    if (pg_accum) cudaFree(pg_accum);
}

namespace {

struct PlanBuilder {
    Plan* pl;
    Decoder* d;
    Packed* P;
    ActKind kind;
    bool strict;
    int B, T;
    int H[3], W[3];
    const float* lmask[3];
    bool failed = false, oom = false;
    std::vector<char> in_use;     // per pool block: live in THIS plan (other plans alias the same blocks)
    size_t live_bytes = 0;
    // ---- VJP plans (forward + backward w.r.t. x): nothing is released during the forward half (the backward half reads the raw
    // conv outputs, the statistics, the attention inputs / kv / contexts); every forward module pushes its backward onto `tape`
    bool vjp = false, keep_all = false;
    std::vector<std::function<void()>> tape;
    std::map<const void*, void*> grad;          // activation -> gradient accumulated so far
    float* gn_bwd_partials = nullptr;
    const void* a1_of[12] = {nullptr};          // block2's input of every ResnetBlock (for its weight gradient)

    void fail_oom(size_t bytes) {
        failed = true; oom = true;
        set_error("out of device memory building the decoder plan for B=" + std::to_string(B) + ", T=" + std::to_string(T) +
                  " (request of " + std::to_string(bytes >> 20) + " MiB on top of " + std::to_string(d->pool.total >> 20) +
                  " MiB of pooled workspace); lower max_chunk or free device memory");
    }
    // Liveness-based reuse: a buffer goes back to the free list (release) once its last consumer has been enqueued; all
    // launches of a plan are ordered on one stream, so the next producer cannot overtake that consumer.
    void* pooled(size_t bytes) {
        const size_t want = Pool::round_size(bytes);
        pl->unpooled_bytes += want;
        auto& blocks = d->pool.blocks;
        in_use.resize(blocks.size(), 0);
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); ++i)
            if (!in_use[i] && blocks[i].bytes >= want && blocks[i].bytes <= want + want / 2 &&
                (best < 0 || blocks[i].bytes < blocks[best].bytes)) best = i;
        if (best < 0) {
            void* p = nullptr;
            if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); fail_oom(want); return nullptr; }
            blocks.push_back(Pool::Block{p, want});
            d->pool.total += want;
            in_use.push_back(0);
            best = (int)blocks.size() - 1;
        }
        in_use[best] = 1;
        live_bytes += blocks[best].bytes;
        pl->pooled_bytes = std::max(pl->pooled_bytes, live_bytes);
        return blocks[best].p;
    }
    void release(const void* p, bool temporary = false) {    // temporary: scratch that even a VJP plan does not keep
        if (!p || (keep_all && !temporary)) return;
        auto& blocks = d->pool.blocks;
        for (int i = 0; i < (int)blocks.size(); ++i)
            if (blocks[i].p == p && in_use[i]) { in_use[i] = 0; live_bytes -= blocks[i].bytes; return; }
    }
    void* act(int lvl, int C) { return pooled((size_t)B * H[lvl] * W[lvl] * C * esize(kind)); }
    float* stats() {
        float* p = (float*)pl->mem.alloc((size_t)B * 16 * 4, true);
        if (!p) fail_oom((size_t)B * 64);
        return p;
    }
    bool use_tc() const { return kind == ACT_BF16 && d->conv_impl_bf16 == 1; }
    // fp32 mode on the tensor cores (six bf16 partial products per MAC, fp32 TMEM accumulation); per-sample weights stay on FFMA
    bool use_split(const ConvGeom& g) const {
        return kind == ACT_F32 && d->fp32_tc == 1 && g.w_batch_rows == 0 && g.Cin0 > 0 && g.Cin0 % 64 == 0 && g.Cin1 % 64 == 0 &&
               (g.Cout == 64 || g.Cout == 128 || g.Cout == 256) && 6 * (g.Cin0 + g.Cin1) / 64 <= 48;
    }
    // fp32 activations -> [hi | mid | lo] bf16 planes, fp32 packed weights -> 6-term bf16 rows, then the tcgen05 conv with fp32 output
    bool add_split_conv(const std::string& name, const ConvGeom& g, const void* src0, const void* src1, const void* w, int wrows,
                        const ConvEpilogue& e, double flops, double bytes) {
        const size_t npix = (size_t)g.B * g.Hin * g.Win;
        void* p0 = pooled(npix * 3 * g.Cin0 * 2);
        void* p1 = g.Cin1 ? pooled(npix * 3 * g.Cin1 * 2) : nullptr;
        if (failed) return false;
        void* w3 = nullptr;
        auto it = P->w3cache.find(w);
        if (it != P->w3cache.end()) w3 = it->second;
        else {
            const int K = g.Cin0 + g.Cin1;
            w3 = P->mem.alloc((size_t)wrows * 6 * K * 2);
            if (!w3) { fail_oom((size_t)wrows * 12 * K); return false; }
            if (split_pack_weights((const float*)w, w3, (size_t)wrows, K, 0)) { failed = true; return false; }   // stream 0: synchronised before the dry run
            P->w3cache[w] = w3;
        }
        const int C0 = g.Cin0, C1 = g.Cin1;
        pl->push("split_planes", 0, 0.0, 0.0, [src0, p0, npix, C0](cudaStream_t s) { return split_f32_planes((const float*)src0, p0, npix, C0, s); });
        if (p1) pl->push("split_planes", 0, 0.0, 0.0, [src1, p1, npix, C1](cudaStream_t s) { return split_f32_planes((const float*)src1, p1, npix, C1, s); });
        ConvGeom gs = g;
        gs.split = 1;
        ConvEpilogue es = e;
        es.out_f32 = 1;
        TcConvPlan* tp = conv_tc_plan_create(gs, p0, p1, w3, wrows, es, d->num_sms, d->halo_mode);
        if (!tp) { failed = true; return false; }
        pl->tc_plans.push_back(tp);
        pl->push(name, 1, flops, bytes, [tp](cudaStream_t s) { return conv_tc_launch(tp, s); });
        release(p0, true); release(p1, true);
        return true;
    }
    size_t slots_for(const ConvGeom& g) const {
        if (use_split(g)) return d->halo_mode && conv_tc_halo_eligible(g) ? conv_tc_halo_partials_slots(g) : conv_tc_partials_slots(g);
        if (!use_tc()) return conv_ffma_partials_slots(g);
        if (d->halo_mode && conv_tc_halo_eligible(g)) return conv_tc_halo_partials_slots(g);
        return conv_tc_partials_slots(g);
    }

    // Registers one convolution.  GN statistics are requested by passing a stats buffer.
    struct InFuse { const float* stats; const float* gamma; const float* beta; const float* tbias; int tb_bstride; const float* mask; };

    // true when a 3x3 conv with this geometry would run on the CTA-pair halo kernel (the only one with the fused input path)
    bool can_fuse_input(const ConvGeom& g) const {
        if (vjp) return false;
        if (!use_tc() || !d->fuse_gn || d->halo_mode != 2 || !conv_tc_cta2_enabled() || !conv_tc_halo_eligible(g)) return false;
        if (getenv("GTTS_FUSE_GN") && atoi(getenv("GTTS_FUSE_GN")) == 0) return false;
        if (const char* e = getenv("GTTS_FUSE_GN_MINC")) { if (g.Cout < atoi(e)) return false; }
        const long tiles = (long)g.B * ((g.Hg + 7) / 8) * ((g.Wg + 15) / 16);
        return g.Cin1 == 0 && tiles >= 2 && d->num_sms >= 2;
    }

    struct Apply { const float* gamma; const float* beta; const float* tbias; int tb_bstride; void* async_out; };

    // small batches: off-critical-path ops go to a parallel branch of the step graph (Plan::lane)
    bool side_lanes() const { return d->side_lanes && B <= d->fuse_epi_max_b && !vjp && !strict; }   // (fp32 convs are several launches)

    // true when a Block conv of this geometry can finish GroupNorm+Mish in its own epilogue (ConvEpilogue::apply)
    bool can_apply(const ConvGeom& g) const {
        if (vjp || !use_tc() || d->halo_mode != 2 || d->fuse_epi == 0 || (d->fuse_epi == 1 && B > d->fuse_epi_max_b)) return false;
        return conv_tc_apply_eligible(g, d->num_sms);
    }
    // ... or by the asynchronous apply warps (any batch; used where the TMEM-resident variant is not)
    bool can_apply_async(const ConvGeom& g) const {
        if (vjp || !use_tc() || d->halo_mode != 2 || !d->fuse_async) return false;
        return conv_tc_apply_async_eligible(g, d->num_sms);
    }

    void add_conv(const ConvGeom& g, const void* src0, const void* src1, const void* w, int wrows, const float* bias,
                  const void* residual, const float* mask, void* out, float* gn_stats, const InFuse* fuse = nullptr,
                  const Apply* apply = nullptr) {
        ConvEpilogue e;
        memset(&e, 0, sizeof(e));
        e.bias = bias; e.residual = residual; e.mask = mask; e.out = out;
        if (apply) {
            e.apply = apply->async_out ? 2 : 1; e.ap_out = apply->async_out;
            e.ap_gamma = apply->gamma; e.ap_beta = apply->beta; e.ap_tbias = apply->tbias; e.ap_tb_bstride = apply->tb_bstride;
        }
        if (fuse) {
            e.in_stats = fuse->stats; e.in_gamma = fuse->gamma; e.in_beta = fuse->beta; e.in_tbias = fuse->tbias;
            e.in_tb_bstride = fuse->tb_bstride; e.in_mask = fuse->mask;
        }
        if (gn_stats) {
            e.gn_partials = pl->partials; e.gn_stats = gn_stats; e.gn_counters = pl->counters; e.gn_eps = 1e-5f;
            if (slots_for(g) > pl->partial_slots) { set_error("internal: GN partial buffer too small"); failed = true; return; }
        }
        const double cin = g.Cin0 + g.Cin1;
        const double npx = (double)g.B * g.nphase * g.Hg * g.Wg;
        const double flops = 2.0 * npx * g.Cout * g.ntaps * cin;
        const double es = (double)esize(kind);
        // algorithmic bytes: read every input once, write every output once, weights once
        const double bytes = ((double)g.B * g.Hin * g.Win * cin + (double)g.B * g.Hout * g.Wout * g.Cout * (residual ? 2 : 1)) * es +
                             (double)wrows * cin * es;
        std::string name = std::string(g.nphase == 4 ? "convT4x4" : (g.ntaps == 9 ? (g.stride == 2 ? "conv3x3s2" : (apply ? (apply->async_out ? "conv3x3ga" : "conv3x3gn") : "conv3x3")) : "conv1x1")) +
                           "_" + std::to_string((int)cin) + "_" + std::to_string(g.Cout) + "_h" + std::to_string(g.Hin);
        if (use_split(g)) {
            add_split_conv(name + "_x3", g, src0, src1, w, wrows, e, flops, bytes);
        } else if (use_tc()) {
            TcConvPlan* tp = conv_tc_plan_create(g, src0, src1, w, wrows, e, d->num_sms, d->halo_mode);   // the plan drops the halo path where the epilogue or geometry rules it out
            if (!tp) { failed = true; return; }
            pl->tc_plans.push_back(tp);
            pl->push(name, 1, flops, bytes, [tp](cudaStream_t s) { return conv_tc_launch(tp, s); });
        } else {
            ActKind k = kind;
            pl->push(name, 1, flops, bytes, [k, g, src0, src1, w, e](cudaStream_t s) { return conv_ffma(k, g, src0, src1, w, e, s); });
        }
    }

    void add_gn_apply(int lvl, int C, const void* raw, const float* st, const BlockW& bw, const float* tbias,
                      const void* residual, bool first_res, void* out) {
        GnApplyArgs a;
        memset(&a, 0, sizeof(a));
        a.raw = raw; a.stats = st; a.gamma = bw.gamma; a.beta = bw.beta; a.mask = lmask[lvl];
        a.tbias = tbias; a.tbias_bstride = pl->est_mode ? 1792 : 0;
        a.residual = residual;
        if (first_res) {
            a.fr_mu = pl->mu; a.fr_x = pl->xt; a.fr_s = pl->splane; a.fr_w = P->fr_w; a.fr_b = P->fr_b;
            a.fr_cin = d->cin_first;
        }
        a.out = out; a.B = B; a.H = H[lvl]; a.W = W[lvl]; a.C = C;
        ActKind k = kind;
        bool st_ = strict;
        const double el = (double)B * H[lvl] * W[lvl] * C;
        pl->push("gn_apply_" + std::to_string(C) + "_h" + std::to_string(H[lvl]), 0, 0.0,
                 el * esize(kind) * (residual ? 3 : 2), [k, a, st_](cudaStream_t s) { return gn_apply(k, a, st_, s); });
    }

    // ResnetBlock (:61-79).  x0/x1 are the (masked) input sources; returns the masked output.
    void* resnet(int r, int lvl, const void* x0, int c0, const void* x1, int c1) {
        const ResnetW& R = P->res[r];
        const int Co = R.cout;
        const float* tb = pl->tb + R.tb_off;
        const int tb_bstride = pl->est_mode ? 1792 : 0;
        const ConvGeom g1 = geom_3x3(B, H[lvl], W[lvl], r == 0 ? 64 : c0, r == 0 ? 0 : c1, Co, 1);
        const ConvGeom g2 = geom_3x3(B, H[lvl], W[lvl], Co, 0, Co, 1);
        const bool ap1 = r != 0 && can_apply(g1);        // block1 writes Mish(GN(conv)) + time bias itself (r == 0: the first conv is not a tcgen05 kernel)
        const bool ap2 = r != 0 && can_apply(g2);        // block2 writes Mish(GN(conv)) + residual itself (r == 0: residual computed inline by gn_apply)
        const bool as1 = r != 0 && !ap1 && can_apply_async(g1);   // the same two fusions done by asynchronous apply warps (raw tile via L2)
        const bool as2 = r != 0 && !ap2 && can_apply_async(g2);
        const bool fuse2 = !ap1 && !ap2 && !as1 && !as2 && can_fuse_input(g2);   // older variant: block2's conv applies block1's GroupNorm+Mish on its operand tiles
        void* raw1 = ap1 ? nullptr : act(lvl, Co);
        void* a1 = fuse2 ? nullptr : act(lvl, Co);
        void* raw2 = ap2 ? nullptr : act(lvl, Co);
        void* out = act(lvl, Co);
        float* st1 = stats(); float* st2 = stats();
        if (failed) return nullptr;
        // the residual branch first (the fused block2 epilogue adds it); at small batch it runs beside block1's conv
        const void* resid = nullptr;
        bool first_res = false;
        if (r == 0) {
            first_res = true;
        } else if (R.wres) {
            void* rb = act(lvl, Co);
            if (failed) return nullptr;
            if (side_lanes()) pl->on_side();
            add_conv(geom_1x1(B, H[lvl], W[lvl], c0, c1, Co, 0), x0, x1, R.wres, Co, R.bres, nullptr, nullptr, rb, nullptr);
            resid = rb;
        } else {
            resid = x0;                                   // Identity(x * mask): x is stored masked
        }
        if (failed) return nullptr;
        if (r == 0) {
            FirstConvArgs f;
            memset(&f, 0, sizeof(f));
            f.mu = pl->mu; f.x = pl->xt; f.splane = pl->splane; f.mask = lmask[0];
            f.w = P->first_wT; f.bias = P->first_b; f.B = B; f.H = H[0]; f.W = W[0]; f.cin = d->cin_first;
            f.use_mma = d->first_conv_mma;
            f.raw = raw1; f.gn_partials = pl->partials; f.gn_stats = st1; f.gn_counters = pl->counters; f.gn_eps = 1e-5f;
            if (first_conv_partials_slots(H[0], W[0]) > pl->partial_slots) { set_error("internal: GN partial buffer too small"); failed = true; return nullptr; }
            ActKind k = kind;
            const double px = (double)B * H[0] * W[0];
            pl->push("first_conv", 0, 2.0 * px * 64 * 9 * d->cin_first, px * (8 + 64 * esize(kind)),
                     [k, f](cudaStream_t s) { return first_conv(k, f, s); });
        } else if (ap1) {
            Apply ap{R.b1.gamma, R.b1.beta, tb, tb_bstride, nullptr};
            add_conv(g1, x0, x1, R.b1.w, 9 * Co, R.b1.bias, nullptr, lmask[lvl], a1, st1, nullptr, &ap);
        } else if (as1) {
            Apply ap{R.b1.gamma, R.b1.beta, tb, tb_bstride, a1};
            add_conv(g1, x0, x1, R.b1.w, 9 * Co, R.b1.bias, nullptr, lmask[lvl], raw1, st1, nullptr, &ap);
        } else {
            add_conv(g1, x0, x1, R.b1.w, 9 * Co, R.b1.bias, nullptr, nullptr, raw1, st1);
        }
        if (fuse2) {
            // block1's GroupNorm + Mish + time bias + mask are applied by block2's conv on its operand tiles
            InFuse fz{st1, R.b1.gamma, R.b1.beta, tb, tb_bstride, lmask[lvl]};
            pl->join_side();
            add_conv(g2, raw1, nullptr, R.b2.w, 9 * Co, R.b2.bias, nullptr, nullptr, raw2, st2, &fz);
        } else {
            if (!ap1 && !as1) {
                if (r == 0) pl->join_side();                 // the time-embedding biases (side branch) are first used here
                add_gn_apply(lvl, Co, raw1, st1, R.b1, tb, nullptr, false, a1);
            }
            if (ap2 || as2) pl->join_side();                  // this conv's epilogue adds the res_conv output
            if (ap2) {
                Apply ap{R.b2.gamma, R.b2.beta, nullptr, 0, nullptr};
                add_conv(g2, a1, nullptr, R.b2.w, 9 * Co, R.b2.bias, resid, lmask[lvl], out, st2, nullptr, &ap);
            } else if (as2) {
                Apply ap{R.b2.gamma, R.b2.beta, nullptr, 0, out};
                add_conv(g2, a1, nullptr, R.b2.w, 9 * Co, R.b2.bias, resid, lmask[lvl], raw2, st2, nullptr, &ap);
            } else {
                add_conv(g2, a1, nullptr, R.b2.w, 9 * Co, R.b2.bias, nullptr, nullptr, raw2, st2);
            }
        }
        if (!ap2 && !as2) {
            pl->join_side();
            add_gn_apply(lvl, Co, raw2, st2, R.b2, nullptr, resid, first_res, out);
        }
        release(raw1); release(a1); release(raw2);
        if (resid != x0) release(resid);
        if (vjp) {
            const bool identity = (r != 0 && !R.wres);
            a1_of[r] = a1;
            tape.push_back([=]() { resnet_backward(r, lvl, x0, c0, x1, c1, raw1, st1, raw2, st2, out, identity); });
        }
        return out;
    }

    // ------------------------------------------------------------------------------------------------ backward (VJP plans)
    void* take_grad(const void* act_ptr) {
        auto it = grad.find(act_ptr);
        if (it == grad.end()) { set_error("internal: backward reached a tensor without a gradient"); failed = true; return nullptr; }
        void* g = it->second;
        grad.erase(it);
        return g;
    }
    void* peek_grad(const void* act_ptr) {
        auto it = grad.find(act_ptr);
        return it == grad.end() ? nullptr : it->second;
    }
    // conv on the gradient path: CUDA-core implicit GEMM in the plan's activation type (generic over every geometry)
    void add_bconv(const char* what, const ConvGeom& g, const void* src, const void* w, const void* residual, void* out) {
        ConvEpilogue e;
        memset(&e, 0, sizeof(e));
        e.residual = residual; e.out = out;
        const double npx = (double)g.B * g.nphase * g.Hg * g.Wg;
        const double flops = 2.0 * npx * g.Cout * g.ntaps * (g.Cin0 + g.Cin1);
        ActKind k = kind;
        const std::string name = std::string("bwd_") + what + "_" + std::to_string(g.Cin0) + "_" + std::to_string(g.Cout) + "_h" + std::to_string(g.Hout);
        if (use_split(g)) {
            add_split_conv(name + "_x3", g, src, nullptr, w, g.ntaps * g.nphase * g.Cout, e, flops, 0.0);
            return;
        }
        if (use_tc()) {
            // bf16 mode: the data gradients run on the same tcgen05 kernels as the forward convolutions
            TcConvPlan* tp = conv_tc_plan_create(g, src, nullptr, w, g.ntaps * g.nphase * g.Cout, e, d->num_sms, d->halo_mode);
            if (!tp) { failed = true; return; }
            pl->tc_plans.push_back(tp);
            pl->push(name, 1, flops, 0.0, [tp](cudaStream_t s) { return conv_tc_launch(tp, s); });
            return;
        }
        pl->push(name, 1, flops, 0.0, [k, g, src, w, e](cudaStream_t s) { return conv_ffma(k, g, src, nullptr, w, e, s); });
    }
    void add_mask_mul(int lvl, int C, const void* in, void* out) {
        ActKind k = kind;
        const float* m = lmask[lvl];
        int Bb = B, Hh = H[lvl], Ww = W[lvl];
        pl->push("bwd_mask", 0, 0.0, 0.0, [k, in, m, out, Bb, Hh, Ww, C](cudaStream_t s) { return mask_mul(k, in, m, out, Bb, Hh, Ww, C, s); });
    }
    // GroupNorm + Mish backward; with `pprefix` (training plans) the same pass over (raw, gy) also leaves the per-channel partial
    // sums of d gamma / d beta, reduced right after
    void add_gn_bwd(int lvl, int C, const void* raw, const float* st, const BlockW& bw, const void* gy, void* graw,
                    const std::string* pprefix = nullptr) {
        GnBwdArgs a;
        memset(&a, 0, sizeof(a));
        a.raw = raw; a.stats = st; a.gamma = bw.gamma; a.beta = bw.beta; a.mask = lmask[lvl]; a.gy = gy; a.graw = graw;
        a.partials = gn_bwd_partials; a.B = B; a.H = H[lvl]; a.W = W[lvl]; a.C = C;
        ActKind k = kind;
        float* ppart = nullptr;
        if (pprefix) {
            ppart = (float*)pooled((size_t)B * gn_bwd_blocks(H[lvl], W[lvl]) * 2 * C * 4);
            if (failed) return;
            a.param_partials = ppart;
        }
        pl->push("bwd_gn_" + std::to_string(C) + "_h" + std::to_string(H[lvl]), 0, 0.0, 0.0, [k, a](cudaStream_t s) { return gn_bwd(k, a, s); });
        pl->kernels_per_step++;                               // two launches
        if (pprefix) {
            float* dga = pg(*pprefix + ".block.1.weight");
            float* dbe = pg(*pprefix + ".block.1.bias");
            int Bb = B, Hh = H[lvl], Ww = W[lvl];
            pl->push("bwd_gn_params", 0, 0.0, 0.0, [ppart, dga, dbe, Bb, Hh, Ww, C](cudaStream_t s) { return gn_param_reduce(ppart, dga, dbe, Bb, Hh, Ww, C, s); });
            release(ppart, true);
        }
    }
    void add_add(int lvl, int C, const void* a, const void* b, void* out) {
        ActKind k = kind;
        const size_t numel = (size_t)B * H[lvl] * W[lvl] * C;
        pl->push("bwd_add", 0, 0.0, 0.0, [k, a, b, out, numel](cudaStream_t s) { return add_tensors(k, a, b, out, numel, s); });
    }

    // ---- parameter gradients (plans with pgrads): every destination is a slice of pl->pg_flat in the parameter's PyTorch layout
    bool pgrads = false;
    float* pg(const std::string& name) {
        auto it = d->pg_layout.find("estimator." + name);
        if (it == d->pg_layout.end()) { set_error("internal: no gradient slot for estimator." + name); failed = true; return pl->pg_flat; }
        return pl->pg_flat + it->second.first;
    }
    const float* param(const std::string& name) {
        auto it = d->params.find("estimator." + name);
        return it == d->params.end() ? nullptr : it->second;
    }
    // dW and db of one conv: gout = gradient of its (pre-norm / pre-mask) output, x0 / x1 = its saved input(s)
    void add_conv_param_grads(const std::string& prefix, const ConvGeom& gfwd, const void* gout, const void* x0, const void* x1, int kind_layout,
                              float* dw_override = nullptr, bool want_bias = true) {
        int slices;
        const bool tcw = kind == ACT_BF16 && d->wgrad_tc && wgrad_tc_eligible(gfwd);   // tcgen05 wgrad (wgrad_tc.cu)
        const size_t pf = tcw ? wgrad_tc_partial_floats(gfwd, d->num_sms, &slices) : wgrad_partial_floats(gfwd, &slices);
        float* part = (float*)pooled(pf * 4);
        float* bpart = (float*)pooled((size_t)256 * gfwd.Cout * 4);
        if (failed) return;
        float* dw = dw_override ? dw_override : pg(prefix + ".weight");
        float* db = want_bias ? pg(prefix + ".bias") : nullptr;
        ActKind k = kind;
        ConvGeom g = gfwd;
        const long npix_out = (long)g.B * g.Hout * g.Wout;
        if (tcw) {
            WgradTcPlan* tp = wgrad_tc_plan_create(g, gout, x0, x1, part, d->num_sms);
            if (!tp) { failed = true; return; }
            pl->wg_plans.push_back(tp);
            const int n_pt = g.ntaps, cin = g.Cin0 + g.Cin1;
            pl->push("bwd_wgrad_" + std::to_string(cin) + "_" + std::to_string(g.Cout) + "_h" + std::to_string(g.Hin), 0,
                     2.0 * (double)g.B * g.Hg * g.Wg * g.Cout * g.ntaps * cin, 0.0,
                     [k, g, gout, tp, part, dw, kind_layout, db, bpart, npix_out, slices, n_pt, cin](cudaStream_t s) {
                         if (int rc = wgrad_tc_launch(tp, s)) return rc;
                         if (int rc = wgrad_reduce(part, dw, slices, n_pt, g.Cout, cin, kind_layout, 1.0f, 0, s)) return rc;
                         if (db) return col_sums(k, gout, bpart, db, npix_out, g.Cout, 1.0f, 0, s);
                         return 0;
                     });
            pl->kernels_per_step += db ? 3 : 1;
            release(part, true); release(bpart, true);
            return;
        }
        pl->push("bwd_wgrad_" + std::to_string(g.Cin0 + g.Cin1) + "_" + std::to_string(g.Cout) + "_h" + std::to_string(g.Hin), 0,
                 2.0 * (double)g.B * g.nphase * g.Hg * g.Wg * g.Cout * g.ntaps * (g.Cin0 + g.Cin1), 0.0,
                 [k, g, gout, x0, x1, part, dw, kind_layout, db, bpart, npix_out](cudaStream_t s) {
                     if (int rc = conv_wgrad(k, g, gout, x0, x1, part, dw, kind_layout, 1.0f, 0, s)) return rc;
                     if (db) return col_sums(k, gout, bpart, db, npix_out, g.Cout, 1.0f, 0, s);
                     return 0;
                 });
        pl->kernels_per_step += db ? 3 : 1;
        release(part, true); release(bpart, true);
    }
    // ResnetBlock backward: gradient of `out` -> gradients of its input source(s) (or, for the first block, of the x plane)
    void resnet_backward(int r, int lvl, const void* x0, int c0, const void* x1, int c1, void* raw1, float* st1, void* raw2, float* st2,
                         void* out, bool identity) {
        const ResnetW& R = P->res[r];
        const Packed::ResnetB& Bw = P->bres[r];
        const int Co = R.cout;
        void* g_out = take_grad(out);
        if (failed) return;
        void* gpre = act(lvl, Co);                            // out = (Mish(GN(raw2)) + residual) * mask
        void* g_raw2 = act(lvl, Co);
        void* g_a1 = act(lvl, Co);
        if (failed) return;
        const std::string rn = kResnetNames[r];
        add_mask_mul(lvl, Co, g_out, gpre);
        release(g_out);
        const std::string pb2 = rn + ".block2", pb1 = rn + ".block1";
        add_gn_bwd(lvl, Co, raw2, st2, R.b2, gpre, g_raw2, pgrads ? &pb2 : nullptr);
        if (pgrads) {
            add_conv_param_grads(rn + ".block2.block.0", geom_3x3(B, H[lvl], W[lvl], Co, 0, Co, 1), g_raw2, a1_of[r], nullptr, 0);
        }
        add_bconv("dgrad3", geom_3x3(B, H[lvl], W[lvl], Co, 0, Co, 1), g_raw2, Bw.d2, nullptr, g_a1);
        release(g_raw2);
        void* g_raw1 = act(lvl, Co);                          // a1 = (Mish(GN(raw1)) + time bias) * mask
        if (failed) return;
        add_gn_bwd(lvl, Co, raw1, st1, R.b1, g_a1, g_raw1, pgrads ? &pb1 : nullptr);
        if (pgrads) {
            {   // time bias: d tb[b][c] = sum_p g_a1[b][p][c] * mask
                float* part = (float*)pooled((size_t)B * 64 * Co * 4);
                if (failed) return;
                ActKind k = kind;
                const float* m = lmask[lvl];
                float* dst = pl->gtb + R.tb_off;
                int Bb = B, Hh = H[lvl], Ww = W[lvl];
                pl->push("bwd_tbias", 0, 0.0, 0.0, [k, g_a1, m, part, dst, Bb, Hh, Ww, Co](cudaStream_t s) {
                    return col_sums_per_sample(k, g_a1, m, part, dst, Bb, Hh, Ww, Co, 1792, s);
                });
                pl->kernels_per_step++;
                release(part, true);
            }
        }
        release(g_a1);
        if (r == 0) {
            ActKind k = kind;
            const float *w1t = P->first_wT, *wres = P->fr_w, *m = lmask[0];
            float *gx = pl->gx, *gmu = pgrads ? pl->gmu : nullptr, *gs = pgrads ? pl->gs_pix : nullptr;
            int Bb = B, Hh = H[0], Ww = W[0], cin = d->cin_first;
            pl->push("bwd_first", 0, 0.0, 0.0, [k, g_raw1, gpre, w1t, wres, m, gx, gmu, gs, Bb, Hh, Ww, cin](cudaStream_t s) {
                return first_bwd(k, g_raw1, gpre, w1t, wres, m, gx, gmu, gs, Bb, Hh, Ww, cin, s);
            });
            if (pgrads) {
                // first conv (K = cin*9) and first res_conv (K = cin) weights + biases in one small kernel, then scattered
                const int nk = cin * 9 + cin + 2, blocks = first_param_blocks(Hh, Ww);
                float* part = (float*)pooled((size_t)B * blocks * 64 * nk * 4);
                float* tmp = (float*)pl->mem.alloc((size_t)64 * nk * 4);
                if (failed || !tmp) { if (!failed) fail_oom(64 * nk * 4); return; }
                const float *mu = pl->mu, *xx = pl->xt, *sp = pl->splane;
                float *dw1 = pg(rn + ".block1.block.0.weight"), *db1 = pg(rn + ".block1.block.0.bias");
                float *dwr = pg(rn + ".res_conv.weight"), *dbr = pg(rn + ".res_conv.bias");
                pl->push("bwd_first_params", 0, 0.0, 0.0, [k, g_raw1, gpre, mu, xx, sp, m, part, tmp, dw1, db1, dwr, dbr, Bb, Hh, Ww, cin, nk](cudaStream_t s) {
                    if (int rc = first_param_grad(k, g_raw1, gpre, mu, xx, sp, m, part, tmp, Bb, Hh, Ww, cin, 0, s)) return rc;
                    const size_t ld = (size_t)nk * 4;
                    GTTS_CHECK_CUDA(cudaMemcpy2DAsync(dw1, (size_t)cin * 9 * 4, tmp, ld, (size_t)cin * 9 * 4, 64, cudaMemcpyDeviceToDevice, s));
                    GTTS_CHECK_CUDA(cudaMemcpy2DAsync(dwr, (size_t)cin * 4, tmp + cin * 9, ld, (size_t)cin * 4, 64, cudaMemcpyDeviceToDevice, s));
                    GTTS_CHECK_CUDA(cudaMemcpy2DAsync(db1, 4, tmp + cin * 9 + cin, ld, 4, 64, cudaMemcpyDeviceToDevice, s));
                    GTTS_CHECK_CUDA(cudaMemcpy2DAsync(dbr, 4, tmp + cin * 9 + cin + 1, ld, 4, 64, cudaMemcpyDeviceToDevice, s));
                    return 0;
                });
                pl->kernels_per_step += 5;
                release(part, true);
            }
            release(g_raw1); release(gpre);
            return;
        }
        if (pgrads) {
            add_conv_param_grads(rn + ".block1.block.0", geom_3x3(B, H[lvl], W[lvl], c0, c1, Co, 1), g_raw1, x0, x1, 0);
            if (R.wres) add_conv_param_grads(rn + ".res_conv", geom_1x1(B, H[lvl], W[lvl], c0, c1, Co, 0), gpre, x0, x1, 0);
        }
        const void* xs[2] = {x0, x1};
        const int cs[2] = {c0, c1};
        for (int k = 0; k < 2; ++k) {
            if (cs[k] == 0) continue;
            void* prev = peek_grad(xs[k]);                     // gradient the source already received from later consumers
            if (prev) grad.erase(xs[k]);
            const void* acc = prev;
            void* owned = prev;
            if (k == 0 && identity) {                          // Identity residual: the source also receives gpre itself
                if (prev) {
                    void* t = act(lvl, cs[k]);
                    if (failed) return;
                    add_add(lvl, cs[k], prev, gpre, t);
                    release(prev);
                    acc = t; owned = t;
                } else {
                    acc = gpre; owned = nullptr;
                }
            }
            void* gn = act(lvl, cs[k]);
            if (failed) return;
            add_bconv("dgrad3", geom_3x3(B, H[lvl], W[lvl], Co, 0, cs[k], 1), g_raw1, Bw.d1[k], acc, gn);
            release(owned);
            if (Bw.rT[k]) {                                    // res_conv (1x1) transposed, on top
                void* gn2 = act(lvl, cs[k]);
                if (failed) return;
                add_bconv("dgrad1", geom_1x1(B, H[lvl], W[lvl], Co, 0, cs[k], 0), gpre, Bw.rT[k], gn, gn2);
                release(gn);
                gn = gn2;
            }
            grad[xs[k]] = gn;
        }
        release(g_raw1); release(gpre);
    }

    // Residual(Rezero(LinearAttention)) backward.  Forward kept: x, kv = [k | v] (1x1 conv of x), ctxn (normalised contexts), ml.
    void attention_backward(int a, int lvl, const void* x, void* out, void* kv, float* ctxn, float* ml, int chunks, int chunk_len) {
        const AttnW& A = P->attn[a];
        const Packed::AttnB& Bw = P->battn[a];
        const int C = A.C, n = H[lvl] * W[lvl];
        void* g_out = take_grad(out);
        if (failed) return;
        void* G = act(lvl, C);
        void* q = act(lvl, 128);
        void* go = act(lvl, 128);
        float* opart = (float*)pl->mem.alloc((size_t)B * 4 * chunks * 1024 * 4);
        float* gctx = (float*)pl->mem.alloc((size_t)B * 4 * 1024 * 4);
        float* sdot = (float*)pl->mem.alloc((size_t)B * 4 * 32 * 4);
        if (failed || !opart || !gctx || !sdot) { if (!failed) fail_oom((size_t)B * 4 * chunks * 4096); return; }
        add_mask_mul(lvl, C, g_out, G);                       // out = (M_b x + g b_out + x) * mask
        release(g_out);
        add_bconv("attn_q", geom_1x1(B, H[lvl], W[lvl], C, 0, 128, 0), x, Bw.wq, nullptr, q);
        add_bconv("attn_go", geom_1x1(B, H[lvl], W[lvl], C, 0, 128, 0), G, Bw.woT, nullptr, go);
        {
            ActKind k = kind;
            int Bb = B;
            pl->push("bwd_attn_outer", 0, 0.0, 0.0, [k, q, go, opart, ctxn, gctx, sdot, Bb, n, chunks, chunk_len](cudaStream_t s) {
                return attn_outer(k, q, go, opart, ctxn, gctx, sdot, Bb, n, chunks, chunk_len, s);
            });
            pl->kernels_per_step++;
        }
        if (!pgrads) release(q);
        void* gq = act(lvl, 128);
        void* gkv = act(lvl, 256);
        void* ao = pgrads ? act(lvl, 128) : nullptr;          // attention output before to_out (for dWout)
        if (failed) return;
        {
            ActKind k = kind;
            int Bb = B;
            const void* qp = pgrads ? q : nullptr;
            pl->push("bwd_attn_pos", 0, 0.0, 0.0, [k, kv, go, ctxn, gctx, ml, sdot, gq, gkv, Bb, n, qp, ao](cudaStream_t s) {
                return attn_pos_bwd(k, kv, go, ctxn, gctx, ml, sdot, gq, gkv, Bb, n, qp, ao, s);
            });
        }
        release(go);
        if (pgrads) {
            release(q);
            const char* attn_names[6] = {"downs.0.2", "downs.1.2", "downs.2.2", "mid_attn", "ups.0.2", "ups.1.2"};
            const std::string an = std::string(attn_names[a]) + ".fn";
            // to_qkv (384, C): rows [0,128) from gq, rows [128,384) from gkv; no bias
            float* dqkv = pg(an + ".fn.to_qkv.weight");
            add_conv_param_grads("", geom_1x1(B, H[lvl], W[lvl], C, 0, 128, 0), gq, x, nullptr, 0, dqkv, false);
            add_conv_param_grads("", geom_1x1(B, H[lvl], W[lvl], C, 0, 256, 0), gkv, x, nullptr, 0, dqkv + (size_t)128 * C, false);
            // to_out behind the Rezero gate: raw A = sum G (x) ao and s = sum G, then dWout = g A, dbout = g s, dg = <Wout, A> + <bout, s>
            float* Araw = (float*)pl->mem.alloc((size_t)C * 128 * 4);
            float* sraw = (float*)pl->mem.alloc((size_t)C * 4);
            float* bpart = (float*)pooled((size_t)256 * C * 4);
            if (failed || !Araw || !sraw) { if (!failed) fail_oom((size_t)C * 512); return; }
            add_conv_param_grads("", geom_1x1(B, H[lvl], W[lvl], 128, 0, C, 0), G, ao, nullptr, 0, Araw, false);
            {
                ActKind k = kind;
                const long npix = (long)B * n;
                const float *wout = A.wout, *bout = param(an + ".fn.to_out.bias");
                const float g = A.g;
                float *dwo = pg(an + ".fn.to_out.weight"), *dbo = pg(an + ".fn.to_out.bias"), *dg = pg(an + ".g");
                pl->push("bwd_attn_out_params", 0, 0.0, 0.0, [k, G, bpart, sraw, npix, C, Araw, wout, bout, g, dwo, dbo, dg](cudaStream_t s) {
                    if (int rc = col_sums(k, G, bpart, sraw, npix, C, 1.0f, 0, s)) return rc;
                    return attn_out_grads(Araw, sraw, wout, bout, g, dwo, dbo, dg, C, s);
                });
                pl->kernels_per_step += 2;
            }
            release(bpart, true); release(ao);
        }
        void* prev = peek_grad(x);
        const void* acc = G;
        void* owned = nullptr;
        if (prev) {
            grad.erase(x);
            void* t = act(lvl, C);
            if (failed) return;
            add_add(lvl, C, prev, G, t);
            release(prev);
            acc = t; owned = t;
        }
        void* g1 = act(lvl, C);
        void* g2 = act(lvl, C);
        if (failed) return;
        add_bconv("attn_gq", geom_1x1(B, H[lvl], W[lvl], 128, 0, C, 0), gq, Bw.wqT, acc, g1);
        add_bconv("attn_gkv", geom_1x1(B, H[lvl], W[lvl], 256, 0, C, 0), gkv, Bw.wkvT, g1, g2);
        release(owned); release(G); release(gq); release(gkv); release(g1);
        grad[x] = g2;
    }

    void resample_backward(bool down, int i, int lvl_in, const void* x, void* out) {
        // down: x at lvl_in -> out at lvl_in + 1 ; up: x at lvl_in -> out at lvl_in - 1
        const int lvl_out = down ? lvl_in + 1 : lvl_in - 1;
        const int C = down ? P->down[i].C : P->up[i].C;
        void* g_out = take_grad(out);
        if (failed) return;
        void* gm = act(lvl_out, C);
        if (failed) return;
        add_mask_mul(lvl_out, C, g_out, gm);
        release(g_out);
        void* prev = peek_grad(x);
        if (prev) grad.erase(x);
        void* gn = act(lvl_in, C);
        if (failed) return;
        if (pgrads) {
            if (down) add_conv_param_grads("downs." + std::to_string(i) + ".3.conv", geom_3x3(B, H[lvl_in], W[lvl_in], C, 0, C, 2), gm, x, nullptr, 0);
            else add_conv_param_grads("ups." + std::to_string(i) + ".3.conv", geom_convT(B, H[lvl_in], W[lvl_in], C), gm, x, nullptr, 1);
        }
        if (down) add_bconv("down", geom_down_dgrad(B, H[lvl_out], W[lvl_out], C), gm, P->bdown[i], prev, gn);
        else add_bconv("up", geom_up_dgrad(B, H[lvl_in], W[lvl_in], C), gm, P->bup[i], prev, gn);
        release(prev); release(gm);
        grad[x] = gn;
    }

    // Residual(Rezero(LinearAttention)) (:39-46, 82-110)
    void* attention(int a, int lvl, const void* x) {
        const AttnW& A = P->attn[a];
        const int C = A.C, n = H[lvl] * W[lvl];
        void* out = act(lvl, C);
        int chunks, chunk_len;
        attn_ctx_plan(n, &chunks, &chunk_len);
        float* ctxn = (float*)pl->mem.alloc((size_t)B * 4096 * 4);
        void* mb = pl->mem.alloc((size_t)B * C * C * esize(kind));
        if (!out || !ctxn || !mb) { failed = true; return nullptr; }
        ActKind k = kind;
        bool st_ = strict;
        int Bb = B;
        float* partials = (float*)pl->mem.alloc((size_t)B * 4 * chunks * 1088 * 4);
        if (!partials) { failed = true; return nullptr; }
        AttnCtxArgs ca{nullptr, B, n, partials, ctxn, chunks, chunk_len};
        if (vjp) {
            ca.ml = (float*)pl->mem.alloc((size_t)B * 4 * 64 * 4);
            if (!ca.ml) { fail_oom((size_t)B * 1024); return nullptr; }
        }
        if (kind == ACT_BF16 && C <= 128 && d->fused_attn && !vjp) {
            // fused k-projection + context: reads x only, k and v are never written to HBM
            const void* wk = A.wkv;                                  // packed bf16 [256][C]: k rows, then v rows
            pl->push("attn_xk_" + std::to_string(C) + "_h" + std::to_string(H[lvl]), 0,
                     2.0 * B * (double)n * 128 * C * 2, (double)B * n * C * 2,
                     [x, wk, partials, Bb, n, C, chunks, chunk_len](cudaStream_t s) {
                         return attn_xk(x, wk, partials, Bb, n, C, chunks, chunk_len, s);
                     });
        } else {
            void* kv = act(lvl, 256);
            if (!kv) { failed = true; return nullptr; }
            add_conv(geom_1x1(B, H[lvl], W[lvl], C, 0, 256, 0), x, nullptr, A.wkv, 256, nullptr, nullptr, nullptr, kv, nullptr);
            ca.kv = kv;
            pl->push("attn_ctx_h" + std::to_string(H[lvl]), 0, 2.0 * B * 4 * (double)n * 1024, (double)B * n * 256 * esize(kind),
                     [k, ca, st_](cudaStream_t s) { return attn_ctx(k, ca, st_, s); });
        }
        pl->push("attn_merge", 0, 0.0, 0.0, [ca, st_](cudaStream_t s) { return attn_merge(ca, st_, s); });
        const float *wout = A.wout, *wq = A.wq;
        float g = A.g;
        pl->push("attn_fold_" + std::to_string(C), 0, 2.0 * B * ((double)C * 128 * 32 + (double)C * C * 128), 0.0,
                 [k, ctxn, wout, wq, g, mb, Bb, C](cudaStream_t s) { return attn_fold(k, ctxn, wout, wq, g, mb, Bb, C, s); });
        add_conv(geom_1x1(B, H[lvl], W[lvl], C, 0, C, C), x, nullptr, mb, B * C, A.gb, x, lmask[lvl], out, nullptr);
        release(ca.kv);
        if (vjp) {
            void* kvp = const_cast<void*>(ca.kv);
            float* mlp = ca.ml;
            tape.push_back([=]() { attention_backward(a, lvl, x, out, kvp, ctxn, mlp, chunks, chunk_len); });
        }
        return out;
    }

    void* downsample(int i, int lvl, const void* x) {
        const SampW& S = P->down[i];
        void* out = act(lvl + 1, S.C);
        if (failed) return nullptr;
        add_conv(geom_3x3(B, H[lvl], W[lvl], S.C, 0, S.C, 2), x, nullptr, S.w, 9 * S.C, S.bias, nullptr, lmask[lvl + 1], out, nullptr);
        if (vjp) tape.push_back([=]() { resample_backward(true, i, lvl, x, out); });
        return out;
    }
    void* upsample(int i, int lvl, const void* x) {
        const SampW& S = P->up[i];
        void* out = act(lvl - 1, S.C);
        if (failed) return nullptr;
        add_conv(geom_convT(B, H[lvl], W[lvl], S.C), x, nullptr, S.w, 16 * S.C, S.bias, nullptr, lmask[lvl - 1], out, nullptr);
        if (vjp) tape.push_back([=]() { resample_backward(false, i, lvl, x, out); });
        return out;
    }

    int build() {
        H[0] = d->n_feats; W[0] = T; H[1] = H[0] / 2; W[1] = T / 2; H[2] = H[0] / 4; W[2] = T / 4;
        // plan-owned state
        const size_t plane = (size_t)B * H[0] * T;
        pl->xt = (float*)pooled(plane * 4);              // rewritten at the start of every call, like mu and the masks
        pl->mu = (float*)pooled(plane * 4);
        pl->score = (float*)pooled(plane * 4);
        if (vjp) {
            pl->vin = (float*)pooled(plane * 4);
            pl->gx = (float*)pooled(plane * 4);
            gn_bwd_partials = (float*)pl->mem.alloc((size_t)B * 256 * 16 * 4);
            if (!gn_bwd_partials) { fail_oom((size_t)B * 16384); return 4; }
            keep_all = true;
            if (pgrads) {
                if (d->pg_layout.empty()) {
                    // every estimator parameter except the time / speaker MLPs (their gradients come from the time-bias and
                    // speaker-plane gradients through PyTorch autograd on the host side): flat buffer in map (= name) order
                    size_t off = 0;
                    for (auto& kv : d->params) {
                        const std::string& n = kv.first;
                        if (n.rfind("estimator.", 0) != 0) continue;
                        if (n.find(".mlp.") != std::string::npos || n.rfind("estimator.mlp.", 0) == 0 || n.rfind("estimator.spk_mlp.", 0) == 0) continue;
                        const size_t numel = d->param_numel[n];
                        d->pg_layout[n] = std::make_pair(off, numel);
                        off += (numel + 3) / 4 * 4;
                    }
                    d->pg_floats = off;
                }
                pl->pg_flat = (float*)pl->mem.alloc(d->pg_floats * 4, true);
                pl->gmu = (float*)pooled(plane * 4);
                pl->gs_pix = (float*)pooled(plane * 4);
                pl->gtb = (float*)pl->mem.alloc((size_t)B * 1792 * 4, true);
                if (failed || !pl->pg_flat || !pl->gtb) { if (!failed) fail_oom(d->pg_floats * 4); return 4; }
            }
        }
        pl->m0 = (float*)pl->mem.alloc((size_t)B * T * 4);
        pl->m1 = (float*)pl->mem.alloc((size_t)B * T / 2 * 4);
        pl->m2 = (float*)pl->mem.alloc((size_t)B * T / 4 * 4);
        pl->splane = (float*)pl->mem.alloc((size_t)B * H[0] * 4, true);
        pl->spk = (float*)pl->mem.alloc((size_t)B * 64 * 4, true);
        pl->tb = (float*)pl->mem.alloc((size_t)(pl->est_mode ? B : 1) * 1792 * 4);
        pl->t_tab = (float*)pl->mem.alloc(pl->max_steps * 4, true);
        pl->beta_tab = (float*)pl->mem.alloc(pl->max_steps * 4, true);
        pl->t_per_sample = (float*)pl->mem.alloc((size_t)B * 4, true);
        pl->step = (int*)pl->mem.alloc(16, true);
        pl->noise_slot = (NoiseSlot*)pl->mem.alloc(sizeof(NoiseSlot), true);
        pl->h_dev = (float*)pl->mem.alloc(16, true);
        pl->counters = (unsigned int*)pl->mem.alloc(conv_tc_counter_words(B) * 4, true);   // tickets + per-sample barrier counters
        // GN partial buffer: the largest slot count of any conv (level 0, either tiling)
        {
            ConvGeom g0 = geom_3x3(B, H[0], W[0], 64, 0, 64, 1);
            size_t s = conv_ffma_partials_slots(g0);
            size_t s2 = std::max(conv_tc_partials_slots(g0), conv_tc_halo_partials_slots(g0));
            size_t s3 = first_conv_partials_slots(H[0], W[0]);
            for (int l = 1; l < 3; ++l)
                for (int C = 64; C <= 256; C *= 2) {
                    ConvGeom gl = geom_3x3(B, H[l], W[l], 64, 0, C, 1);
                    s = std::max(s, conv_ffma_partials_slots(gl));
                    s2 = std::max(s2, std::max(conv_tc_partials_slots(gl), conv_tc_halo_partials_slots(gl)));
                }
            pl->partial_slots = std::max(s, std::max(s2, s3));
            pl->partials = (float*)pl->mem.alloc((size_t)B * pl->partial_slots * 16 * 4);
        }
        if (failed) return 4;
        if (!pl->m0 || !pl->m1 || !pl->m2 || !pl->partials || !pl->counters || !pl->tb) {
            fail_oom((size_t)B * pl->partial_slots * 64);
            return 4;
        }
        lmask[0] = pl->m0; lmask[1] = pl->m1; lmask[2] = pl->m2;

        // ---- time-embedding biases
        {
            TembWeights tw = P->temb;
            const float* tsrc = pl->est_mode ? pl->t_per_sample : pl->t_tab;
            const int* step = pl->step;
            int is_table = pl->est_mode ? 0 : 1;
            float pes = (float)d->pe_scale;
            float* tb = pl->tb;
            int nb = pl->est_mode ? B : 1;
            bool st_ = strict;
            if (side_lanes()) pl->on_side();                  // first consumer: the gn_apply of the first block
            pl->push("temb", 0, 0.0, 0.0, [tw, tsrc, step, is_table, pes, tb, nb, st_](cudaStream_t s) {
                return temb_bias(tw, tsrc, step, is_table, pes, tb, nb, st_, s);
            });
        }
        // ---- U-Net (:189-211).  `next(y)` makes y the running activation and frees the previous one unless a skip keeps it.
        void* x = nullptr;
        void* skip[3] = {nullptr, nullptr, nullptr};
        auto next = [&](void* y, bool keep_prev = false) {
            if (!keep_prev) release(x);
            x = y;
        };
        next(resnet(0, 0, nullptr, 0, nullptr, 0));
        if (!failed) next(resnet(1, 0, x, 64, nullptr, 0));
        if (!failed) next(attention(0, 0, x));
        skip[0] = x;                                         // hiddens[0]: stored by the reference but never consumed (two up stages)
        if (!failed) next(downsample(0, 0, x));
        if (!failed) next(resnet(2, 1, x, 64, nullptr, 0));
        if (!failed) next(resnet(3, 1, x, 128, nullptr, 0));
        if (!failed) next(attention(1, 1, x));
        skip[1] = x;
        if (!failed) next(downsample(1, 1, x), true);
        if (!failed) next(resnet(4, 2, x, 128, nullptr, 0));
        if (!failed) next(resnet(5, 2, x, 256, nullptr, 0));
        if (!failed) next(attention(2, 2, x));
        skip[2] = x;                                         // downs.2.3 = Identity (:158)
        if (!failed) next(resnet(6, 2, x, 256, nullptr, 0), true);
        if (!failed) next(attention(3, 2, x));
        if (!failed) next(resnet(7, 2, x, 256, nullptr, 0));
        if (!failed) next(resnet(8, 2, x, 256, skip[2], 256));   // cat((x, hiddens.pop()), 1) (:207)
        release(skip[2]);
        if (!failed) next(resnet(9, 2, x, 128, nullptr, 0));
        if (!failed) next(attention(4, 2, x));
        if (!failed) next(upsample(0, 2, x));
        if (!failed) next(resnet(10, 1, x, 128, skip[1], 128));
        release(skip[1]);
        if (!failed) next(resnet(11, 1, x, 64, nullptr, 0));
        if (!failed) next(attention(5, 1, x));
        if (!failed) next(upsample(1, 1, x));
        // ---- final block + final conv + Euler update (:212-216, 265-267)
        void* rawf = failed ? nullptr : act(0, 64);
        float* stf = failed ? nullptr : stats();
        if (failed) return 5;
        add_conv(geom_3x3(B, H[0], W[0], 64, 0, 64, 1), x, nullptr, P->final_block.w, 9 * 64, P->final_block.bias, nullptr, nullptr, rawf, stf);
        if (failed) return 5;
        {
            EulerArgs e;
            memset(&e, 0, sizeof(e));
            e.raw = rawf; e.stats = stf; e.gamma = P->final_block.gamma; e.beta = P->final_block.beta;
            e.wf = P->wf; e.bf = P->bf; e.mask = pl->m0; e.mu = pl->mu; e.xt = pl->xt;
            e.score_out = pl->est_mode ? pl->score : nullptr;
            e.beta_tab = pl->beta_tab; e.step = pl->step; e.h_ptr = pl->h_dev;
            e.noise_slot = pl->noise_slot; e.sde = pl->sde ? 1 : 0;
            e.update = pl->est_mode ? 0 : 1;
            e.B = B; e.H = H[0]; e.W = W[0];
            ActKind k = kind;
            bool st_ = strict;
            const double px = (double)B * H[0] * W[0];
            pl->push("euler_step", 0, 0.0, px * (64 * esize(kind) + 12), [k, e, st_](cudaStream_t s) { return euler_step(k, e, st_, s); });
            int* step = pl->step;
            pl->push("advance_step", 0, 0.0, 0.0, [step](cudaStream_t s) { return advance_step(step, s); });
        }
        if (vjp) {
            // ---- backward half: cotangent v of the score -> J^T v w.r.t. x, modules in reverse order
            keep_all = false;
            void* ghf = act(0, 64);
            void* g_rawf = act(0, 64);
            void* gin = act(0, 64);
            if (failed) return 5;
            {
                ActKind k = kind;
                const float *v = pl->vin, *wf = P->wf, *m = pl->m0;
                int Bb = B, Hh = H[0], Ww = W[0];
                pl->push("bwd_final", 0, 0.0, 0.0, [k, v, wf, m, ghf, Bb, Hh, Ww](cudaStream_t s) { return final_bwd(k, v, wf, m, ghf, Bb, Hh, Ww, s); });
            }
            const std::string pfb = "final_block";
            add_gn_bwd(0, 64, rawf, stf, P->final_block, ghf, g_rawf, pgrads ? &pfb : nullptr);
            if (pgrads) {
                add_conv_param_grads("final_block.block.0", geom_3x3(B, H[0], W[0], 64, 0, 64, 1), g_rawf, x, nullptr, 0);
                // final_conv (64 -> 1): dwf[c] = sum v * mask * hf[c], dbf = sum v * mask
                float* part = (float*)pooled((size_t)256 * 65 * 4);
                float* tmp = (float*)pl->mem.alloc(65 * 4);
                if (failed || !tmp) { if (!failed) fail_oom(260); return 5; }
                ActKind k = kind;
                const float *st = stf, *ga = P->final_block.gamma, *be = P->final_block.beta, *v = pl->vin, *m = pl->m0;
                float *dwf = pg("final_conv.weight"), *dbf = pg("final_conv.bias");
                int Bb = B, Hh = H[0], Ww = W[0];
                pl->push("bwd_final_params", 0, 0.0, 0.0, [k, rawf, st, ga, be, v, m, part, tmp, dwf, dbf, Bb, Hh, Ww](cudaStream_t s) {
                    if (int rc = final_param_grad(k, rawf, st, ga, be, v, m, part, tmp, Bb, Hh, Ww, 0, s)) return rc;
                    GTTS_CHECK_CUDA(cudaMemcpyAsync(dwf, tmp, 64 * 4, cudaMemcpyDeviceToDevice, s));
                    GTTS_CHECK_CUDA(cudaMemcpyAsync(dbf, tmp + 64, 4, cudaMemcpyDeviceToDevice, s));
                    return 0;
                });
                pl->kernels_per_step += 3;
                release(part, true);
            }
            release(ghf);
            add_bconv("dgrad3", geom_3x3(B, H[0], W[0], 64, 0, 64, 1), g_rawf, P->bfinal, nullptr, gin);
            release(g_rawf);
            grad[x] = gin;
            for (int i = (int)tape.size() - 1; i >= 0 && !failed; --i) tape[i]();
            if (failed) return 5;
            if (!grad.empty()) { set_error("internal: backward left " + std::to_string(grad.size()) + " gradient(s) unconsumed"); return 5; }
        }
        return 0;
    }
};


int plan_create(Decoder* d, ActKind kind, int B, int T, bool est_mode, bool sde, bool vjp, bool pgrads, Plan** out) {
    if (int rc = pack_weights(d, kind)) return rc;
    if (vjp) { if (int rc = pack_backward_weights(d, kind)) return rc; }
    std::unique_ptr<Plan> pl(new Plan());
    pl->d = d; pl->kind = kind; pl->strict = (kind == ACT_F32); pl->B = B; pl->T = T;
    pl->est_mode = est_mode; pl->sde = sde; pl->vjp = vjp; pl->pgrads = pgrads;
    PlanBuilder pb;
    pb.vjp = vjp; pb.pgrads = pgrads;
    pb.pl = pl.get(); pb.d = d; pb.P = d->packed[kind].get(); pb.kind = kind; pb.strict = pl->strict;
    pb.B = B; pb.T = T;
    int rc = pb.build();
    if (rc == 0 && pb.failed) rc = 5;
    if (rc) return pb.oom ? 4 : rc;                  // 4 = out of device memory (get_plan drops cached plans + pool and retries)
    // ---- dry run on zeroed state (validates every launch), then capture the step as a CUDA graph
    GTTS_CHECK_CUDA(cudaStreamSynchronize(0));       // the zero-fills of the plan-owned buffers ran on the legacy stream
    cudaStream_t cs;
    GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    auto run_ops = [&](cudaStream_t s) -> int {
        for (auto& op : pl->ops)
            if (int r = op(s)) return r;
        return 0;
    };
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->xt, 0, (size_t)B * d->n_feats * T * 4, cs));
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->mu, 0, (size_t)B * d->n_feats * T * 4, cs));
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->m0, 0, (size_t)B * T * 4, cs));
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->m1, 0, (size_t)B * T / 2 * 4, cs));
    GTTS_CHECK_CUDA(cudaMemsetAsync(pl->m2, 0, (size_t)B * T / 4 * 4, cs));
    if (vjp) GTTS_CHECK_CUDA(cudaMemsetAsync(pl->vin, 0, (size_t)B * d->n_feats * T * 4, cs));
    schedule_kernel<<<1, 32, 0, cs>>>(pl->t_tab, pl->beta_tab, pl->h_dev, 1, d->beta_min, d->beta_max);
    {
        // point the SDE noise slot at something readable for the dry run
        const NoiseSlot dummy{pl->score, 0ull};
        GTTS_CHECK_CUDA(cudaMemcpyAsync((void*)pl->noise_slot, &dummy, sizeof(dummy), cudaMemcpyHostToDevice, cs));
    }
    rc = run_ops(cs);
    if (rc) { cudaStreamDestroy(cs); return rc; }
    GTTS_CHECK_CUDA(cudaStreamSynchronize(cs));
    GTTS_CHECK_CUDA(cudaGetLastError());
    if (d->use_graph) {
        cudaGraph_t graph = nullptr;
        bool any_side = false;
        for (uint8_t l : pl->lane) any_side = any_side || l;
        cudaStream_t side = nullptr;
        std::vector<cudaEvent_t> evs;
        if (any_side) GTTS_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        auto new_event = [&]() { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); return e; };
        // small-batch sampler plans: programmatic dependent launch between consecutive kernels of the step (see pdl_enabled)
        const bool pdl_small = d->pdl_small && B <= d->fuse_epi_max_b && !vjp;
        pdl_set_override(pdl_small ? 1 : 0);
        GTTS_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        if (!any_side) {
            rc = run_ops(cs);
        } else {
            std::vector<cudaEvent_t> open;                   // completion events of side ops not yet joined
            for (size_t i = 0; i < pl->ops.size() && !rc; ++i) {
                if (pl->lane[i]) {
                    cudaEvent_t f = new_event(), j = new_event();
                    cudaEventRecord(f, cs);
                    cudaStreamWaitEvent(side, f, 0);
                    rc = pl->ops[i](side);
                    cudaEventRecord(j, side);
                    open.push_back(j);
                } else {
                    if (pl->join[i]) { for (cudaEvent_t j : open) cudaStreamWaitEvent(cs, j, 0); open.clear(); }
                    rc = pl->ops[i](cs);
                }
            }
            for (cudaEvent_t j : open) cudaStreamWaitEvent(cs, j, 0);   // every branch rejoins before the capture ends
        }
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        pdl_set_override(0);
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(cs);
            if (!rc) { set_error(std::string("graph capture failed: ") + cudaGetErrorString(ce)); rc = 1; }
            return rc;
        }
        ce = cudaGraphInstantiate(&pl->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            cudaStreamDestroy(cs);
            set_error(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ce));
            return 1;
        }
    }
    GTTS_CHECK_CUDA(cudaStreamSynchronize(cs));
    cudaStreamDestroy(cs);
    *out = pl.release();
    return 0;
}

void drop_plans(Decoder* d, bool drop_pool) {
    cudaDeviceSynchronize();                         // nothing may still be running out of the memory about to be freed
    for (auto& kv : d->plans) delete kv.second.plan;
    d->plans.clear();
    if (drop_pool) d->pool.clear();
}

int get_plan(Decoder* d, ActKind kind, int B, int T, bool est_mode, bool sde, cudaStream_t stream, Plan** out, bool vjp = false,
             bool pgrads = false) {
    std::string key = std::to_string((int)kind) + ":" + std::to_string(B) + ":" + std::to_string(T) + ":" +
                      (est_mode ? "e" : "s") + (sde ? "n" : "o") + (vjp ? (pgrads ? "p" : "v") : "-") + (d->use_graph ? "g" : "x") +
                      std::to_string(d->conv_impl_bf16) + std::to_string(d->halo_mode) + std::to_string(d->fused_attn) + std::to_string(d->fuse_gn) +
                      std::to_string(d->fuse_epi) + "." + std::to_string(d->fuse_epi_max_b) + "." + std::to_string(d->fp32_tc) + std::to_string(d->fuse_async) +
                      std::to_string(d->side_lanes) + std::to_string(d->wgrad_tc) + std::to_string(d->pdl_small) + std::to_string(d->first_conv_mma);
    auto it = d->plans.find(key);
    if (it != d->plans.end()) {
        it->second.last_use = ++d->use_clock;
        *out = it->second.plan;
        return 0;
    }
    // Slow path.  The new plan's dry run writes into pool blocks that launches already queued by this call (an earlier chunk)
    // or by an earlier call may still be using: wait for them.
    GTTS_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (d->has_done) GTTS_CHECK_CUDA(cudaEventSynchronize(d->done_ev));
    // LRU bound on the number of cached plans (descriptors + graph; the activation memory is the shared pool)
    while ((int)d->plans.size() >= std::max(1, d->max_plans)) {
        auto lru = d->plans.begin();
        for (auto j = d->plans.begin(); j != d->plans.end(); ++j)
            if (j->second.last_use < lru->second.last_use) lru = j;
        delete lru->second.plan;
        d->plans.erase(lru);
    }
    Plan* pl = nullptr;
    int rc = plan_create(d, kind, B, T, est_mode, sde, vjp, pgrads, &pl);
    if (rc == 4) {
        // out of device memory: give back everything this handle caches (all plans and the whole pool) and try once more
        drop_plans(d, true);
        rc = plan_create(d, kind, B, T, est_mode, sde, vjp, pgrads, &pl);
    }
    if (rc) return rc;
    d->plans_created++;
    d->plans[key] = Decoder::CachedPlan{pl, ++d->use_clock};
    *out = pl;
    return 0;
}

// Calls on one handle are ordered even when they arrive on different streams (plans share pooled workspace and keep per-call
// state): a call waits for the event the previous call recorded, unless it runs on the same stream.
int enter_call(Decoder* d, cudaStream_t stream) {
    GTTS_CHECK_CUDA(cudaSetDevice(d->device));
    if (d->has_done && d->last_stream != stream) GTTS_CHECK_CUDA(cudaStreamWaitEvent(stream, d->done_ev, 0));
    return 0;
}
int leave_call(Decoder* d, cudaStream_t stream) {
    GTTS_CHECK_CUDA(cudaEventRecord(d->done_ev, stream));
    d->last_stream = stream;
    d->has_done = true;
    return 0;
}

int plan_step(Plan* pl, cudaStream_t s) {
    if (pl->graph_exec) {
        GTTS_CHECK_CUDA(cudaGraphLaunch(pl->graph_exec, s));
        return 0;
    }
    for (auto& op : pl->ops)
        if (int r = op(s)) return r;
    return 0;
}

int common_checks(Decoder* d, int B, int T, const float* spk) {
    GTTS_REQUIRE(d != nullptr, "null decoder handle");
    GTTS_REQUIRE(B >= 1 && T >= 4, "bad batch or length");
    GTTS_REQUIRE(T % 4 == 0, "T must be a multiple of 4 (fix_len_compatibility, model/utils.py:13-17)");
    GTTS_REQUIRE(d->n_spks_mode != 1 || spk != nullptr, "this decoder was built with n_spks > 1: spk is required");
    return 0;
}

}  // namespace

// flags: bit0 = fp32 (strict) precision, else bf16; bit1 = SDE update with injected noise
int decoder_reverse_diffusion(Decoder* d, const float* z, const float* mask, const float* mu, const float* spk,
                              float* out, int B, int T, int n_timesteps, int flags, const float* noise,
                              cudaStream_t stream) {
    if (int rc = common_checks(d, B, T, spk)) return rc;
    GTTS_REQUIRE(n_timesteps >= 1 && n_timesteps <= 4096, "n_timesteps must be in [1, 4096]");
    const bool sde = (flags & 2) != 0;
    GTTS_REQUIRE(!sde || noise != nullptr, "SDE update requested without a noise tensor");
    if (int rc = enter_call(d, stream)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const size_t plane = (size_t)d->n_feats * T;
    d->launches_last_call = 0;
    for (int b0 = 0; b0 < B; b0 += d->max_chunk) {
        const int Bc = std::min(d->max_chunk, B - b0);
        Plan* pl = nullptr;
        if (int rc = get_plan(d, kind, Bc, T, false, sde, stream, &pl)) return rc;
        Packed* P = d->packed[kind].get();
        if (int rc = build_level_masks(mask + (size_t)b0 * T, pl->m0, pl->m1, pl->m2, Bc, T, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->mu, mu + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        if (int rc = init_xt(z + b0 * plane, pl->m0, pl->xt, Bc, d->n_feats, T, stream)) return rc;
        if (d->n_spks_mode == 1)
            if (int rc = spk_mlp(spk + (size_t)b0 * 64, P->spk_w0t, P->spk_b0, P->spk_w2t, P->spk_b2, pl->splane, Bc,
                                 d->n_feats, pl->strict, stream)) return rc;
        schedule_kernel<<<(n_timesteps + 255) / 256, 256, 0, stream>>>(pl->t_tab, pl->beta_tab, pl->h_dev, n_timesteps,
                                                                      d->beta_min, d->beta_max);
        GTTS_CHECK_CUDA(cudaMemsetAsync(pl->step, 0, 4, stream));
        if (sde) {
            // base of this chunk's noise and the stride between Euler steps go to device memory: the captured graph reads them
            const NoiseSlot ns{noise + b0 * plane, (unsigned long long)((size_t)B * plane)};
            GTTS_CHECK_CUDA(cudaMemcpyAsync((void*)pl->noise_slot, &ns, sizeof(ns), cudaMemcpyHostToDevice, stream));
        }
        for (int i = 0; i < n_timesteps; ++i)
            if (int rc = plan_step(pl, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(out + b0 * plane, pl->xt, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        d->launches_last_call += pl->kernels_per_step * n_timesteps + 5;
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return leave_call(d, stream);
}

int decoder_estimator(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                      float* out, int B, int T, int flags, cudaStream_t stream) {
    if (int rc = common_checks(d, B, T, spk)) return rc;
    if (int rc = enter_call(d, stream)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const size_t plane = (size_t)d->n_feats * T;
    d->launches_last_call = 0;
    for (int b0 = 0; b0 < B; b0 += d->max_chunk) {
        const int Bc = std::min(d->max_chunk, B - b0);
        Plan* pl = nullptr;
        if (int rc = get_plan(d, kind, Bc, T, true, false, stream, &pl)) return rc;
        Packed* P = d->packed[kind].get();
        if (int rc = build_level_masks(mask + (size_t)b0 * T, pl->m0, pl->m1, pl->m2, Bc, T, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->mu, mu + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->xt, x + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->t_per_sample, t + b0, Bc * 4, cudaMemcpyDeviceToDevice, stream));
        if (d->n_spks_mode == 1)
            if (int rc = spk_mlp(spk + (size_t)b0 * 64, P->spk_w0t, P->spk_b0, P->spk_w2t, P->spk_b2, pl->splane, Bc,
                                 d->n_feats, pl->strict, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemsetAsync(pl->step, 0, 4, stream));
        if (int rc = plan_step(pl, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(out + b0 * plane, pl->score, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        d->launches_last_call += pl->kernels_per_step + 3;
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return leave_call(d, stream);
}

// Score and vector-Jacobian product in one pass: out_score = estimator(x, mask, mu, t, spk), out_gx = (d out_score / d x)^T v.
// This is what the Hutchinson divergence of the probability-flow ODE needs (reference n_best/likelihood/likelihood.py:27-38).
int decoder_estimator_vjp(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                          const float* v, float* out_score, float* out_gx, int B, int T, int flags, cudaStream_t stream) {
    if (int rc = common_checks(d, B, T, spk)) return rc;
    if (int rc = enter_call(d, stream)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const size_t plane = (size_t)d->n_feats * T;
    const int chunk = std::min(d->max_chunk, 16);             // every intermediate of the U-Net is kept for the backward half
    d->launches_last_call = 0;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int Bc = std::min(chunk, B - b0);
        Plan* pl = nullptr;
        if (int rc = get_plan(d, kind, Bc, T, true, false, stream, &pl, true)) return rc;
        Packed* P = d->packed[kind].get();
        if (int rc = build_level_masks(mask + (size_t)b0 * T, pl->m0, pl->m1, pl->m2, Bc, T, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->mu, mu + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->xt, x + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->vin, v + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->t_per_sample, t + b0, Bc * 4, cudaMemcpyDeviceToDevice, stream));
        if (d->n_spks_mode == 1)
            if (int rc = spk_mlp(spk + (size_t)b0 * 64, P->spk_w0t, P->spk_b0, P->spk_w2t, P->spk_b2, pl->splane, Bc,
                                 d->n_feats, pl->strict, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemsetAsync(pl->step, 0, 4, stream));
        if (int rc = plan_step(pl, stream)) return rc;
        if (out_score) GTTS_CHECK_CUDA(cudaMemcpyAsync(out_score + b0 * plane, pl->score, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(out_gx + b0 * plane, pl->gx, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        d->launches_last_call += pl->kernels_per_step + 4;
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return leave_call(d, stream);
}

// Full backward of one estimator call (training): score, J^T v w.r.t. x AND mu, the per-pixel gradient of the speaker plane, the gradient
// of the 1792 per-block time biases, and the gradient of every convolution / GroupNorm / attention parameter (kept in the handle, fetched
// with decoder_get_param_grad).  Chunks of at most 16 samples; the parameter gradients are summed over the chunks.
int decoder_estimator_backward(Decoder* d, const float* x, const float* mask, const float* mu, const float* t, const float* spk,
                               const float* v, float* out_score, float* out_gx, float* out_gmu, float* out_gs_pix, float* out_gtb, int B,
                               int T, int flags, cudaStream_t stream) {
    if (int rc = common_checks(d, B, T, spk)) return rc;
    if (int rc = enter_call(d, stream)) return rc;
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const size_t plane = (size_t)d->n_feats * T;
    const int chunk = std::min(d->max_chunk, 16);
    d->launches_last_call = 0;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int Bc = std::min(chunk, B - b0);
        Plan* pl = nullptr;
        if (int rc = get_plan(d, kind, Bc, T, true, false, stream, &pl, true, true)) return rc;
        if (!d->pg_accum) {
            GTTS_CHECK_CUDA(cudaMalloc(&d->pg_accum, d->pg_floats * 4));
        }
        Packed* P = d->packed[kind].get();
        if (int rc = build_level_masks(mask + (size_t)b0 * T, pl->m0, pl->m1, pl->m2, Bc, T, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->mu, mu + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->xt, x + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->vin, v + b0 * plane, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(pl->t_per_sample, t + b0, Bc * 4, cudaMemcpyDeviceToDevice, stream));
        if (d->n_spks_mode == 1)
            if (int rc = spk_mlp(spk + (size_t)b0 * 64, P->spk_w0t, P->spk_b0, P->spk_w2t, P->spk_b2, pl->splane, Bc,
                                 d->n_feats, pl->strict, stream)) return rc;
        GTTS_CHECK_CUDA(cudaMemsetAsync(pl->step, 0, 4, stream));
        if (int rc = plan_step(pl, stream)) return rc;
        if (out_score) GTTS_CHECK_CUDA(cudaMemcpyAsync(out_score + b0 * plane, pl->score, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        GTTS_CHECK_CUDA(cudaMemcpyAsync(out_gx + b0 * plane, pl->gx, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        if (out_gmu) GTTS_CHECK_CUDA(cudaMemcpyAsync(out_gmu + b0 * plane, pl->gmu, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        if (out_gs_pix) GTTS_CHECK_CUDA(cudaMemcpyAsync(out_gs_pix + b0 * plane, pl->gs_pix, Bc * plane * 4, cudaMemcpyDeviceToDevice, stream));
        if (out_gtb) GTTS_CHECK_CUDA(cudaMemcpyAsync(out_gtb + (size_t)b0 * 1792, pl->gtb, (size_t)Bc * 1792 * 4, cudaMemcpyDeviceToDevice, stream));
        if (int rc = accumulate_floats(d->pg_accum, pl->pg_flat, d->pg_floats, b0 > 0 ? 1 : 0, stream)) return rc;
        d->launches_last_call += pl->kernels_per_step + 8;
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return leave_call(d, stream);
}

int decoder_get_param_grad(Decoder* d, const char* name, float* dst, size_t numel, cudaStream_t stream) {
    GTTS_REQUIRE(d != nullptr && name != nullptr && dst != nullptr, "null argument");
    auto it = d->pg_layout.find(name);
    GTTS_REQUIRE(it != d->pg_layout.end(), "get_param_grad: this parameter has no device-side gradient (time / speaker MLPs go through the time-bias and speaker-plane gradients)");
    GTTS_REQUIRE(it->second.second == numel && d->pg_accum != nullptr, "get_param_grad: wrong size, or no backward call yet");
    if (int rc = enter_call(d, stream)) return rc;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(dst, d->pg_accum + it->second.first, numel * 4, cudaMemcpyDeviceToDevice, stream));
    return leave_call(d, stream);
}

// all parameter gradients in one copy: dst[offset(name) .. offset(name) + numel(name)) in the layout decoder_param_grad_slot reports
int decoder_get_param_grads_flat(Decoder* d, float* dst, size_t numel, cudaStream_t stream) {
    GTTS_REQUIRE(d != nullptr && dst != nullptr, "null argument");
    GTTS_REQUIRE(d->pg_accum != nullptr && numel == d->pg_floats, "get_param_grads_flat: wrong size, or no backward call yet");
    if (int rc = enter_call(d, stream)) return rc;
    GTTS_CHECK_CUDA(cudaMemcpyAsync(dst, d->pg_accum, numel * 4, cudaMemcpyDeviceToDevice, stream));
    return leave_call(d, stream);
}
// offset and size (floats) of one parameter's gradient in the flat buffer; total size with name == nullptr
int decoder_param_grad_slot(const Decoder* d, const char* name, size_t* offset, size_t* numel) {
    GTTS_REQUIRE(d != nullptr && offset != nullptr && numel != nullptr, "null argument");
    if (name == nullptr) { *offset = 0; *numel = d->pg_floats; return 0; }
    auto it = d->pg_layout.find(name);
    GTTS_REQUIRE(it != d->pg_layout.end(), "param_grad_slot: this parameter has no device-side gradient");
    *offset = it->second.first; *numel = it->second.second;
    return 0;
}

// Runs one step of the cached sampler plan for (B<=max_chunk, T) eagerly, `reps` times, with a CUDA event pair
// around every launch; writes a JSON report {ops:[{name,is_conv,flops,bytes,ms}...]} into buf.
int decoder_profile_step(Decoder* d, int B, int T, int flags, int reps, char* buf, size_t buflen, cudaStream_t stream) {
    if (int rc = common_checks(d, B, T, (const float*)1)) return rc;
    GTTS_REQUIRE(buf != nullptr && buflen > 64 && reps >= 1, "profile: bad arguments");
    const ActKind kind = (flags & 1) ? ACT_F32 : ACT_BF16;
    const bool pgr = (flags & 8) != 0;                       // bit 3: ... with the parameter gradients (the training plan)
    const bool vjp = (flags & 4) != 0 || pgr;                // bit 2: the forward + backward (VJP) plan instead of the sampler step
    const int Bc = std::min(vjp ? std::min(d->max_chunk, 16) : d->max_chunk, B);
    if (int rc = enter_call(d, stream)) return rc;
    Plan* pl = nullptr;
    if (int rc = get_plan(d, kind, Bc, T, vjp, false, stream, &pl, vjp, pgr)) return rc;
    const size_t nops = pl->ops.size();
    std::vector<cudaEvent_t> ev(2 * nops);
    for (auto& e : ev) GTTS_CHECK_CUDA(cudaEventCreate(&e));
    std::vector<double> ms(nops, 0.0);
    for (int r = 0; r < reps; ++r) {
        for (size_t i = 0; i < nops; ++i) {
            GTTS_CHECK_CUDA(cudaEventRecord(ev[2 * i], stream));
            if (int rc = pl->ops[i](stream)) return rc;
            GTTS_CHECK_CUDA(cudaEventRecord(ev[2 * i + 1], stream));
        }
        GTTS_CHECK_CUDA(cudaStreamSynchronize(stream));
        for (size_t i = 0; i < nops; ++i) {
            float t = 0.f;
            GTTS_CHECK_CUDA(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
            ms[i] += t / reps;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    std::string js = "{\"B\":" + std::to_string(Bc) + ",\"T\":" + std::to_string(T) +
                     ",\"workspace_bytes\":" + std::to_string(pl->pooled_bytes + pl->mem.total) +
                     ",\"workspace_bytes_without_reuse\":" + std::to_string(pl->unpooled_bytes + pl->mem.total) +
                     ",\"pool_bytes\":" + std::to_string(d->pool.total) + ",\"plans_cached\":" + std::to_string(d->plans.size()) +
                     ",\"plans_created\":" + std::to_string(d->plans_created) + ",\"ops\":[";
    for (size_t i = 0; i < nops; ++i) {
        char tmp[512];
        snprintf(tmp, sizeof(tmp), "%s{\"name\":\"%s\",\"is_conv\":%d,\"flops\":%.6e,\"bytes\":%.6e,\"ms\":%.6f}",
                 i ? "," : "", pl->info[i].name.c_str(), pl->info[i].is_conv, pl->info[i].flops, pl->info[i].bytes, ms[i]);
        js += tmp;
    }
    js += "]}";
    GTTS_REQUIRE(js.size() + 1 <= buflen, "profile: report buffer too small");
    memcpy(buf, js.c_str(), js.size() + 1);
    return leave_call(d, stream);
}

Decoder* decoder_new(int n_spks, int n_feats, int dim, double beta_min, double beta_max, double pe_scale, int device) {
    if (n_feats != 80 || dim != 64) { set_error("only n_feats=80, dim=64 (the reference configuration) is supported"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major != 10) { set_error("this library contains sm_100a code only; found compute capability " +
                                      std::to_string(prop.major) + "." + std::to_string(prop.minor)); return nullptr; }
    Decoder* d = new Decoder();
    d->n_spks_mode = n_spks > 1 ? 1 : (n_spks == -1 ? 2 : 0);
    d->cin_first = n_spks > 1 ? 3 : 2;
    d->n_feats = n_feats; d->dim = dim;
    d->beta_min = beta_min; d->beta_max = beta_max; d->pe_scale = pe_scale;
    d->device = device; d->num_sms = prop.multiProcessorCount;
    // experiment switches (profiling runs): defaults of the two GroupNorm fusion options
    if (const char* e = getenv("GTTS_FUSE_EPI")) d->fuse_epi = atoi(e);
    if (const char* e = getenv("GTTS_FUSE_GN")) d->fuse_gn = atoi(e);
    if (const char* e = getenv("GTTS_FP32_TC")) d->fp32_tc = atoi(e);
    if (const char* e = getenv("GTTS_FUSE_ASYNC")) d->fuse_async = atoi(e);
    if (const char* e = getenv("GTTS_SIDE")) d->side_lanes = atoi(e);
    if (const char* e = getenv("GTTS_WGRAD_TC")) d->wgrad_tc = atoi(e);
    if (const char* e = getenv("GTTS_PDL_SMALL")) d->pdl_small = atoi(e);
    if (const char* e = getenv("GTTS_FIRST_CONV_MMA")) d->first_conv_mma = atoi(e);
    if (cudaEventCreateWithFlags(&d->done_ev, cudaEventDisableTiming) != cudaSuccess) {
        set_error("cudaEventCreate failed"); cudaGetLastError(); delete d; return nullptr;
    }
    return d;
}

int decoder_set_param(Decoder* d, const char* name, const float* data, size_t numel) {
    GTTS_REQUIRE(d != nullptr && name != nullptr && data != nullptr, "null argument");
    GTTS_CHECK_CUDA(cudaSetDevice(d->device));
    std::string n(name);
    float* dst = nullptr;
    auto it = d->params.find(n);
    if (it != d->params.end() && d->param_numel[n] == numel) {
        dst = it->second;
    } else {
        dst = (float*)d->param_mem.alloc(numel * 4);
        GTTS_REQUIRE(dst != nullptr, "out of device memory for parameters");
        d->params[n] = dst;
        d->param_numel[n] = numel;
    }
    // any packed weights / plans built from the old values are stale; nothing may still be running from them
    if (!d->plans.empty() || d->packed[0] || d->packed[1]) {
        drop_plans(d, false);
        d->packed[0].reset();
        d->packed[1].reset();
    }
    // Ordered on the legacy stream, which pack_weights synchronises before anything reads the parameters.  The SOURCE must be
    // complete when this is called: the Python host synchronises the producing stream before uploading (model/diffusion.py).
    GTTS_CHECK_CUDA(cudaMemcpyAsync(dst, data, numel * 4, cudaMemcpyDefault, 0));
    GTTS_CHECK_CUDA(cudaStreamSynchronize(0));        // `data` (host or device) may be released by the caller on return
    return 0;
}

void decoder_delete(Decoder* d) { delete d; }

int decoder_set_option(Decoder* d, const char* key, int value) {
    GTTS_REQUIRE(d != nullptr && key != nullptr, "null argument");
    std::string k(key);
    if (k == "max_chunk") { GTTS_REQUIRE(value >= 1 && value <= 64, "max_chunk must be in [1, 64]"); d->max_chunk = value; }
    else if (k == "max_plans") { GTTS_REQUIRE(value >= 1, "max_plans must be >= 1"); d->max_plans = value; }
    else if (k == "trim") drop_plans(d, true);                  // give back every cached plan and the pooled workspace
    else if (k == "use_graph") d->use_graph = value != 0;
    else if (k == "conv_impl_bf16") d->conv_impl_bf16 = value;
    else if (k == "halo_mode") d->halo_mode = value;
    else if (k == "fused_attn") d->fused_attn = value;
    else if (k == "fuse_gn") d->fuse_gn = value;
    else if (k == "fuse_epi") d->fuse_epi = value;
    else if (k == "fuse_epi_max_b") d->fuse_epi_max_b = value;
    else if (k == "fp32_tc") d->fp32_tc = value;
    else if (k == "fuse_async") d->fuse_async = value;
    else if (k == "side_lanes") d->side_lanes = value;
    else if (k == "wgrad_tc") d->wgrad_tc = value;
    else if (k == "pdl_small") d->pdl_small = value;
    else if (k == "first_conv_mma") d->first_conv_mma = value;
    else { set_error("unknown option " + k); return 2; }
    return 0;
}

long decoder_launches_last_call(const Decoder* d) { return d ? d->launches_last_call : 0; }

// out[0] plans cached, [1] plans created since the handle exists, [2] pooled workspace bytes, [3] largest per-plan live
// workspace (pooled peak + plan-owned), [4] what that plan would need with one buffer per tensor
int decoder_cache_info(const Decoder* d, long long* out, int n) {
    GTTS_REQUIRE(d != nullptr && out != nullptr && n >= 5, "cache_info: bad arguments");
    long long live = 0, flat = 0;
    for (auto& kv : d->plans) {
        const Plan* p = kv.second.plan;
        const long long l = (long long)(p->pooled_bytes + p->mem.total);
        if (l > live) { live = l; flat = (long long)(p->unpooled_bytes + p->mem.total); }
    }
    out[0] = (long long)d->plans.size(); out[1] = d->plans_created; out[2] = (long long)d->pool.total; out[3] = live; out[4] = flat;
    return 0;
}
int decoder_device(const Decoder* d) { return d ? d->device : -1; }

}  // namespace gtts
