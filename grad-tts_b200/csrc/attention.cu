// K6 -- LinearAttention core (reference model/diffusion.py:82-100).
//
//   k = softmax(k, over all n = H*W positions)      (:95, padded positions have k = 0 and still count)
//   ctx[d,e] = sum_n k[d,n] v[e,n]                  (:96)
//   out[e,n] = sum_d ctx[d,e] q[d,n]                (:97)    then to_out, *g, +x   (:46,100,109)
//
// B200 restructuring: q is never materialised.  out = ctx^T (Wq x), so
//   to_out(out)*g = (g * Wout * blockdiag_h(ctx_h^T) * Wq) x + g*b_out = M_b x + g*b_out
// with a per-sample CxC matrix M_b.  attn_ctx computes the normalised 32x32 contexts with an online
// softmax (split over n, merged deterministically by the last CTA of each (b, head)); attn_fold builds
// M_b; the product M_b x runs as a 1x1 conv with per-sample weights on the tensor cores (conv_tc.cu).
#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

constexpr int kSub = 128;        // pixels per sub-tile
constexpr int kKPitch = 36;      // floats per k row in smem (16-byte aligned rows)

template <typename T, bool kStrict>
__global__ void __launch_bounds__(256)
attn_ctx_kernel(AttnCtxArgs a) {
    __shared__ __align__(16) float ks[kSub * kKPitch];
    __shared__ __align__(16) float vs[kSub * 32];
    __shared__ float s_m[32], s_scale[32];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31;
    const int chunk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int n0 = chunk * a.chunk_len;
    const int n1 = min(a.n, n0 + a.chunk_len);
    const T* kv = reinterpret_cast<const T*>(a.kv) + (size_t)b * a.n * 256;

    const int ps = tid >> 6, q = tid & 63, d0 = (q >> 3) * 4, e0 = (q & 7) * 4;
    float acc[4][4], lsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    if (tid < 32) s_m[tid] = -INFINITY;
    __syncthreads();

    for (int p0 = n0; p0 < n1; p0 += kSub) {
        // ---- stage k and v of this head for 128 positions (fp32 in smem)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int item = it * 256 + tid;             // 0..1023: (pixel, which, vec)
            const int pix = item >> 3, which = (item >> 2) & 1, vec = item & 3;
            const int n = p0 + pix;
            float v[8];
            if (n < n1) {
                Act<T>::load8(kv + (size_t)n * 256 + which * 128 + head * 32 + vec * 8, v);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = which ? 0.f : -INFINITY;
            }
            float* dst = which ? &vs[pix * 32 + vec * 8] : &ks[pix * kKPitch + vec * 8];
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
        // ---- running max per d (8 threads per column, then a shuffle reduce)
        {
            const int d = tid >> 3, part = tid & 7;
            float mx = -INFINITY;
            for (int pix = part; pix < kSub; pix += 8) mx = fmaxf(mx, ks[pix * kKPitch + d]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            if (part == 0) {
                const float mo = s_m[d], mn = fmaxf(mo, mx);
                s_scale[d] = kStrict ? expf(mo - mn) : __expf(mo - mn);   // exp(-inf) = 0 on the first tile
                s_m[d] = mn;
            }
        }
        __syncthreads();
        // ---- p = exp(k - m) in place
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int item = it * 256 + tid;             // 0..4095: (pixel, d)
            const int pix = item >> 5, d = item & 31;
            const float x = ks[pix * kKPitch + d] - s_m[d];
            ks[pix * kKPitch + d] = kStrict ? expf(x) : __expf(x);
        }
        __syncthreads();
        // ---- rescale and accumulate ctx[d][e] += p[n][d] * v[n][e]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sc = s_scale[d0 + i];
            lsum[i] *= sc;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] *= sc;
        }
#pragma unroll 4
        for (int pp = 0; pp < 32; ++pp) {
            const int pix = ps * 32 + pp;
            const float4 p4 = *reinterpret_cast<const float4*>(&ks[pix * kKPitch + d0]);
            const float4 v4 = *reinterpret_cast<const float4*>(&vs[pix * 32 + e0]);
            const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                lsum[i] += pv[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pv[i], vv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // ---- reduce the four pixel-splits in a fixed order, write this chunk's partial (m, l, ctx)
    float* red = ks;                                      // 4 x 1056 floats fit in ks+vs (contiguous? no: use ks only)
    // ks holds 128*36 = 4608 floats >= 4*1056 = 4224
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[ps * 1056 + 32 + (d0 + i) * 32 + e0 + j] = acc[i][j];
        if ((q & 7) == 0) red[ps * 1056 + d0 + i] = lsum[i];
    }
    __syncthreads();
    float* part = a.partials + (((size_t)b * 4 + head) * a.chunks + chunk) * 1088;
    for (int i = tid; i < 1056; i += 256) {
        float s = (red[i] + red[1056 + i]) + (red[2112 + i] + red[3168 + i]);
        part[32 + i] = s;                                 // [32..64) = l, [64..1088) = ctx
    }
    if (tid < 32) part[tid] = s_m[tid];
    __threadfence();
    __syncthreads();

    // ---- ticket: the last chunk of this (b, head) merges all partials
    if (tid == 0) {
        __threadfence();
        unsigned int old = atomicAdd(&a.counters[b * 4 + head], 1u);
        s_flag = (old == (unsigned int)(a.chunks - 1));
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    const float* pbase = a.partials + ((size_t)b * 4 + head) * a.chunks * 1088;
    if (tid < 32) {
        float M = -INFINITY;
        for (int c = 0; c < a.chunks; ++c) M = fmaxf(M, __ldcg(pbase + (size_t)c * 1088 + tid));
        float l = 0.f;
        for (int c = 0; c < a.chunks; ++c) {
            const float mc = __ldcg(pbase + (size_t)c * 1088 + tid);
            const float wgt = kStrict ? expf(mc - M) : __expf(mc - M);
            l += wgt * __ldcg(pbase + (size_t)c * 1088 + 32 + tid);
        }
        s_m[tid] = M;
        s_scale[tid] = 1.0f / l;
    }
    __syncthreads();
    for (int i = tid; i < 1024; i += 256) {
        const int d = i >> 5;
        const float M = s_m[d];
        float s = 0.f;
        for (int c = 0; c < a.chunks; ++c) {
            const float mc = __ldcg(pbase + (size_t)c * 1088 + d);
            const float wgt = kStrict ? expf(mc - M) : __expf(mc - M);
            s += wgt * __ldcg(pbase + (size_t)c * 1088 + 64 + i);
        }
        a.ctxn[((size_t)b * 4 + head) * 1024 + i] = s * s_scale[d];
    }
    if (tid == 0) a.counters[b * 4 + head] = 0u;
    (void)lane;
}

// grid (C/16, B), 256 threads
template <typename WT>
__global__ void __launch_bounds__(256)
attn_fold_kernel(const float* __restrict__ ctxn, const float* __restrict__ wout, const float* __restrict__ wq,
                 float g, WT* __restrict__ mb, int C) {
    __shared__ float cs[4 * 32 * 33];
    __shared__ float ws[16 * 128];
    __shared__ float P[16 * 128];
    const int tid = threadIdx.x, b = blockIdx.y, co0 = blockIdx.x * 16;
    for (int i = tid; i < 4096; i += 256) {
        const int h = i >> 10, d = (i >> 5) & 31, e = i & 31;
        cs[(h * 32 + d) * 33 + e] = ctxn[(size_t)b * 4096 + i];
    }
    for (int i = tid; i < 16 * 128; i += 256) ws[i] = wout[(size_t)(co0 + (i >> 7)) * 128 + (i & 127)];
    __syncthreads();
    {   // P[cl][h*32+d] = sum_e Wout[co][h*32+e] * ctxn[h][d][e]
        const int cl = tid >> 4, part = tid & 15;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int hd = part * 8 + k, h = hd >> 5;
            float s = 0.f;
#pragma unroll 8
            for (int e = 0; e < 32; ++e) s = fmaf(ws[cl * 128 + h * 32 + e], cs[hd * 33 + e], s);
            P[cl * 128 + hd] = s;
        }
    }
    __syncthreads();
    {   // M[co][ci] = g * sum_hd P[cl][hd] * Wq[hd][ci]
        const int cl = tid >> 4, cg = tid & 15;
        const int nk = C >> 4;
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
        for (int hd = 0; hd < 128; ++hd) {
            const float pv = P[cl * 128 + hd];
            const float* wr = wq + (size_t)hd * C + cg;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k < nk) acc[k] = fmaf(pv, __ldg(wr + 16 * k), acc[k]);
        }
        WT* o = mb + ((size_t)b * C + co0 + cl) * C + cg;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < nk) Act<WT>::st(o + 16 * k, g * acc[k]);
    }
}

}  // namespace

void attn_ctx_plan(int n, int* chunks, int* chunk_len) {
    int len = ((n + 31) / 32 + kSub - 1) / kSub * kSub;
    if (len < kSub) len = kSub;
    *chunk_len = len;
    *chunks = (n + len - 1) / len;
}

int attn_ctx(ActKind act, const AttnCtxArgs& a, bool strict, cudaStream_t s) {
    GTTS_REQUIRE(a.chunk_len % kSub == 0 && a.chunks >= 1, "attn_ctx: bad chunk plan");
    dim3 grid(a.chunks, 4, a.B);
    if (act == ACT_F32) {
        if (strict) attn_ctx_kernel<float, true><<<grid, 256, 0, s>>>(a);
        else        attn_ctx_kernel<float, false><<<grid, 256, 0, s>>>(a);
    } else {
        if (strict) attn_ctx_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(a);
        else        attn_ctx_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>(a);
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attn_fold(ActKind wkind, const float* ctxn, const float* wout, const float* wq, float g, void* mb_out, int B,
              int C, cudaStream_t s) {
    GTTS_REQUIRE(C % 16 == 0 && C <= 256, "attn_fold: C must be a multiple of 16 and <= 256");
    dim3 grid(C / 16, B);
    if (wkind == ACT_F32) attn_fold_kernel<float><<<grid, 256, 0, s>>>(ctxn, wout, wq, g, (float*)mb_out, C);
    else attn_fold_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(ctxn, wout, wq, g, (__nv_bfloat16*)mb_out, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
