// Backward of the score network with respect to its INPUT x (vector-Jacobian product J^T v), the piece the probability-flow
// likelihood needs (reference n_best/likelihood/likelihood.py:27-38: grad of sum(fn(x, t) * eps) w.r.t. x) -- SURVEY 8(f) ranks 2/3.
//
// The convolution data-gradients run on the same implicit-GEMM kernels as the forward pass (a dgrad of a 3x3 conv is a 3x3
// conv with transposed, flipped weights; of a strided conv a 4-phase transposed conv; of the 4x4 transposed conv a 16-tap strided
// conv): only the weight packing is new (pack_*_dgrad below).  This file holds the point-wise / reduction kernels:
//   GroupNorm + Mish backward      (reference Block, model/diffusion.py:52-58; two per-(sample, group) sums, then the apply pass)
//   final 1x1 conv backward        (:213-216)
//   first conv / first res_conv backward onto the x plane (:181-184, 52, 70)
//   LinearAttention backward       (:87-100): outer-product reduction g_ctx = sum_n q g_o^T and the per-position kernel
// Everything is per-sample, deterministic (fixed-order reductions) and fp32.
#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

// d/dn [ n * tanh(softplus(n)) ], torch softplus threshold 20 (above it softplus(n) = n and its derivative is 1)
__device__ __forceinline__ float mish_grad(float n) {
    if (n > 20.0f) {
        const float t = tanhf(n);
        return t + n * (1.0f - t * t);
    }
    const float sp = log1pf(expf(n));
    const float t = tanhf(sp);
    const float sg = 1.0f / (1.0f + expf(-n));
    return t + n * (1.0f - t * t) * sg;
}

constexpr int kGbPix = 32;            // pixels per CTA iteration group in the GroupNorm backward kernels

// ------------------------------------------------------------------------------------------------ GroupNorm + Mish backward
// y = (Mish(n) [+ ...]) * mask with n = gamma * xhat + beta, xhat = (raw - mean) * rstd over the sample's group.
// Pass 1: per (sample, group) S1 = sum dxhat, S2 = sum dxhat * xhat with dxhat = g_y * mask * Mish'(n) * gamma.
// grid (blocks, B), 256 threads; thread = (pixel slot, 8-channel vector).  Partials [B][blocks][16] (8 x S1, 8 x S2).
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_stats_kernel(GnBwdArgs a, int blocks) {
    const int C8 = a.C >> 3, b = blockIdx.y;
    const int vec = threadIdx.x % C8, pslot = threadIdx.x / C8, pstep = 256 / C8;
    const int c0 = vec * 8, gsz = a.C >> 3, g = c0 / gsz;             // a vector of 8 channels never straddles a group (gsz >= 8)
    const float mean = a.stats[(b * 8 + g) * 2], rstd = a.stats[(b * 8 + g) * 2 + 1];
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = a.gamma[c0 + j]; be[j] = a.beta[c0 + j]; }
    const int HW = a.H * a.W;
    const T* raw = reinterpret_cast<const T*>(a.raw) + (size_t)b * HW * a.C;
    const T* gy = reinterpret_cast<const T*>(a.gy) + (size_t)b * HW * a.C;
    float s1 = 0.f, s2 = 0.f;
    float t1[8], t2[8];                                             // per channel: d gamma, d beta (training plans)
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1[j] = 0.f; t2[j] = 0.f; }
    const bool want_params = a.param_partials != nullptr;
    const int per_block = (HW + blocks - 1) / blocks;
    const int p_lo = blockIdx.x * per_block, p_hi = min(HW, p_lo + per_block);
    for (int p = p_lo + pslot; p < p_hi; p += pstep) {
        const float m = a.mask[(size_t)b * a.W + p % a.W];
        if (m == 0.f) continue;
        float r[8], gg[8];
        Act<T>::load8(raw + (size_t)p * a.C + c0, r);
        Act<T>::load8(gy + (size_t)p * a.C + c0, gg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (r[j] - mean) * rstd;
            const float n = fmaf(ga[j], xh, be[j]);
            const float gn = gg[j] * m * mish_grad(n);              // gradient w.r.t. the GroupNorm output
            const float dxh = gn * ga[j];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            if (want_params) { t1[j] = fmaf(gn, xh, t1[j]); t2[j] += gn; }
        }
    }
    if (want_params) {
        // per-channel partial row of this CTA: [d gamma (C) | d beta (C)], position slots added in order
        __shared__ float s_t[256 * 16];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_t[threadIdx.x * 16 + j] = t1[j]; s_t[threadIdx.x * 16 + 8 + j] = t2[j]; }
        __syncthreads();
        for (int o = threadIdx.x; o < 2 * a.C; o += 256) {
            const int which = o / a.C, c = o % a.C, v = c >> 3, j = c & 7;
            float t = 0.f;
            for (int ps = 0; ps < pstep; ++ps) t += s_t[(ps * C8 + v) * 16 + which * 8 + j];
            a.param_partials[((size_t)b * blocks + blockIdx.x) * 2 * a.C + o] = t;
        }
    }
    // fixed-order CTA reduction: thread values -> shared, 16 threads sum their (group, which) column
    __shared__ float s_v[256 * 2];
    s_v[threadIdx.x * 2] = s1;
    s_v[threadIdx.x * 2 + 1] = s2;
    __syncthreads();
    if (threadIdx.x < 16) {
        const int grp = threadIdx.x & 7, which = threadIdx.x >> 3;
        const int vpg = gsz / 8;                                       // vectors per group
        float s = 0.f;
        for (int ps = 0; ps < pstep; ++ps)
            for (int v = grp * vpg; v < (grp + 1) * vpg; ++v) s += s_v[(ps * C8 + v) * 2 + which];
        a.partials[((size_t)b * blocks + blockIdx.x) * 16 + threadIdx.x] = s;
    }
}

// Pass 2: g_raw = rstd * (dxhat - S1/N - xhat * S2/N), N = (C/8) * H * W.  Every CTA first reduces its sample's partial rows
// in the same fixed order (double).
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(GnBwdArgs a, int blocks, int stat_blocks) {
    __shared__ double s_red[16];
    __shared__ double s_part[16][16];
    const int b = blockIdx.y;
    {   // fixed-order reduction of the sample's partial rows: 16 row slices in parallel, then the slices in order
        const int col = threadIdx.x & 15, slice = threadIdx.x >> 4;
        double s = 0.0;
        const float* pp = a.partials + (size_t)b * stat_blocks * 16 + col;
        for (int i = slice; i < stat_blocks; i += 16) s += (double)__ldcg(pp + (size_t)i * 16);
        s_part[slice][col] = s;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double s = 0.0;
        for (int i = 0; i < 16; ++i) s += s_part[i][threadIdx.x];
        s_red[threadIdx.x] = s;
    }
    __syncthreads();
    const int C8 = a.C >> 3;
    const int vec = threadIdx.x % C8, pslot = threadIdx.x / C8, pstep = 256 / C8;
    const int c0 = vec * 8, gsz = a.C >> 3, g = c0 / gsz;
    const float mean = a.stats[(b * 8 + g) * 2], rstd = a.stats[(b * 8 + g) * 2 + 1];
    const int HW = a.H * a.W;
    const double inv_n = 1.0 / ((double)gsz * (double)HW);
    const float m1 = (float)(s_red[g] * inv_n), m2 = (float)(s_red[8 + g] * inv_n);
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = a.gamma[c0 + j]; be[j] = a.beta[c0 + j]; }
    const T* raw = reinterpret_cast<const T*>(a.raw) + (size_t)b * HW * a.C;
    const T* gy = reinterpret_cast<const T*>(a.gy) + (size_t)b * HW * a.C;
    T* gr = reinterpret_cast<T*>(a.graw) + (size_t)b * HW * a.C;
    const int per_block = (HW + blocks - 1) / blocks;
    const int p_lo = blockIdx.x * per_block, p_hi = min(HW, p_lo + per_block);
    for (int p = p_lo + pslot; p < p_hi; p += pstep) {
        const float m = a.mask[(size_t)b * a.W + p % a.W];
        float r[8], gg[8], o[8];
        Act<T>::load8(raw + (size_t)p * a.C + c0, r);
        if (m != 0.f) Act<T>::load8(gy + (size_t)p * a.C + c0, gg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (r[j] - mean) * rstd;
            float dxh = 0.f;
            if (m != 0.f) {
                const float n = fmaf(ga[j], xh, be[j]);
                dxh = gg[j] * m * mish_grad(n) * ga[j];
            }
            // the statistics cover padded frames too, so the mean terms reach them even where dxhat is zero
            o[j] = rstd * (dxh - m1 - xh * m2);
        }
        Act<T>::store8(gr + (size_t)p * a.C + c0, o);
    }
}

// ------------------------------------------------------------------------------------------------ final conv backward
// score = (sum_c wf[c] * hf[c] * mask + bf) * mask  =>  g_hf[p][c] = wf[c] * v[p] * mask
template <typename T>
__global__ void __launch_bounds__(256)
final_bwd_kernel(const float* __restrict__ v, const float* __restrict__ wf, const float* __restrict__ mask, T* __restrict__ ghf,
                 int B, int H, int W) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;           // (pixel, 8-channel vector)
    const size_t npix = (size_t)B * H * W;
    if (i >= npix * 8) return;
    const size_t pix = i >> 3;
    const int c0 = (int)(i & 7) * 8;
    const int b = (int)(pix / ((size_t)H * W)), w = (int)(pix % W);
    const float s = v[pix] * mask[(size_t)b * W + w];
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = wf[c0 + j] * s;
    Act<T>::store8(ghf + pix * 64 + c0, o);
}

// ------------------------------------------------------------------------------------------------ first conv backward
// The U-Net input is stack([mu, x, (s)]) * mask; only the x plane (channel 1) needs a gradient:
//   gx[b,h,w] = mask * ( sum_{co,ky,kx} W1[co][1][ky][kx] * g_raw1[b, h-ky+1, w-kx+1, co]  +  sum_co Wres[co][1] * g_res[b,h,w,co] )
// w1: transposed first-conv weight [cin*9][64] (k = (ci*3+ky)*3+kx), wres: (64, cin) row-major.  One thread per pixel.
template <typename T>
__global__ void __launch_bounds__(128)
first_bwd_kernel(const T* __restrict__ graw1, const T* __restrict__ gres, const float* __restrict__ w1t, const float* __restrict__ wres,
                 const float* __restrict__ mask, float* __restrict__ gx, float* __restrict__ gmu, float* __restrict__ gs, int B, int H, int W,
                 int cin) {
    __shared__ float s_w[3][9 * 64 + 64];                              // per input channel: 9 taps x 64 co of W1, then 64 co of Wres
    for (int i = threadIdx.x; i < 3 * 9 * 64; i += 128) {
        const int ci = i / 576, r = i % 576;
        s_w[ci][r] = ci < cin ? w1t[(size_t)(ci * 9 + r / 64) * 64 + (r % 64)] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * 64; i += 128) {
        const int ci = i / 64, c = i % 64;
        s_w[ci][576 + c] = ci < cin ? wres[(size_t)c * cin + ci] : 0.f;
    }
    __syncthreads();
    const size_t pix = (size_t)blockIdx.x * 128 + threadIdx.x;
    const size_t npix = (size_t)B * H * W;
    if (pix >= npix) return;
    const int b = (int)(pix / ((size_t)H * W));
    const int rem = (int)(pix % ((size_t)H * W)), h = rem / W, w = rem % W;
    const float m = mask[(size_t)b * W + w];
    if (m == 0.f) {
        gx[pix] = 0.f;
        if (gmu) gmu[pix] = 0.f;
        if (gs) gs[pix] = 0.f;
        return;
    }
    const bool want0 = gmu != nullptr, want2 = gs != nullptr && cin == 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        const int hh = h - ky + 1;
        if (hh < 0 || hh >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ww = w - kx + 1;
            if (ww < 0 || ww >= W) continue;
            const T* gp = graw1 + (((size_t)b * H + hh) * W + ww) * 64;
            const int t = (ky * 3 + kx) * 64;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                float gv[8];
                Act<T>::load8(gp + c8 * 8, gv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a1 = fmaf(s_w[1][t + c8 * 8 + j], gv[j], a1);
                    if (want0) a0 = fmaf(s_w[0][t + c8 * 8 + j], gv[j], a0);
                    if (want2) a2 = fmaf(s_w[2][t + c8 * 8 + j], gv[j], a2);
                }
            }
        }
    }
    {
        const T* gp = gres + pix * 64;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float gv[8];
            Act<T>::load8(gp + c8 * 8, gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a1 = fmaf(s_w[1][576 + c8 * 8 + j], gv[j], a1);
                if (want0) a0 = fmaf(s_w[0][576 + c8 * 8 + j], gv[j], a0);
                if (want2) a2 = fmaf(s_w[2][576 + c8 * 8 + j], gv[j], a2);
            }
        }
    }
    gx[pix] = a1 * m;
    if (gmu) gmu[pix] = a0 * m;
    if (gs) gs[pix] = want2 ? a2 * m : 0.f;
}

// ------------------------------------------------------------------------------------------------ element-wise helpers
template <typename T>
__global__ void mask_mul_kernel(const T* __restrict__ in, const float* __restrict__ mask, T* __restrict__ out, int B, int HW, int W, int C) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;           // (pixel, 8-channel vector)
    const int C8 = C >> 3;
    const size_t nvec = (size_t)B * HW * C8;
    if (i >= nvec) return;
    const size_t pix = i / C8;
    const int b = (int)(pix / HW), w = (int)((pix % HW) % W);
    const float m = mask[(size_t)b * W + w];
    float v[8];
    Act<T>::load8(in + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m;
    Act<T>::store8(out + i * 8, v);
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, size_t nvec) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nvec) return;
    float x[8], y[8];
    Act<T>::load8(a + i * 8, x);
    Act<T>::load8(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    Act<T>::store8(out + i * 8, x);
}

// ------------------------------------------------------------------------------------------------ attention backward
// g_ctx[b,h,d,e] = sum_n q[b,n,h*32+d] * go[b,n,h*32+e]: grid (chunks, 4, B), 256 threads (4 pixel splits x 64 4x4 blocks).
template <typename T>
__global__ void __launch_bounds__(256)
attn_outer_kernel(const T* __restrict__ q, const T* __restrict__ go, float* __restrict__ partials, int n, int chunks, int chunk_len) {
    __shared__ __align__(16) float qs[128 * 36];
    __shared__ __align__(16) float gs[128 * 32];
    const int tid = threadIdx.x, chunk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int n0 = chunk * chunk_len, n1 = min(n, n0 + chunk_len);
    const T* qb = q + (size_t)b * n * 128;
    const T* gb = go + (size_t)b * n * 128;
    const int ps = tid >> 6, qq = tid & 63, d0 = (qq >> 3) * 4, e0 = (qq & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = n0; p0 < n1; p0 += 128) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int item = it * 256 + tid, pix = item >> 3, which = (item >> 2) & 1, vec = item & 3;
            const int nn = p0 + pix;
            float v[8];
            if (nn < n1) Act<T>::load8((which ? gb : qb) + (size_t)nn * 128 + head * 32 + vec * 8, v);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            float* dst = which ? &gs[pix * 32 + vec * 8] : &qs[pix * 36 + vec * 8];
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < 32; ++pp) {
            const int pix = ps * 32 + pp;
            const float4 a4 = *reinterpret_cast<const float4*>(&qs[pix * 36 + d0]);
            const float4 b4 = *reinterpret_cast<const float4*>(&gs[pix * 32 + e0]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* red = qs;                                                   // 4 x 1024 floats fit in qs (4608)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[ps * 1024 + (d0 + i) * 32 + e0 + j] = acc[i][j];
    __syncthreads();
    float* part = partials + (((size_t)b * 4 + head) * chunks + chunk) * 1024;
    for (int i = tid; i < 1024; i += 256) part[i] = (red[i] + red[1024 + i]) + (red[2048 + i] + red[3072 + i]);
}

// sums the chunk partials in a fixed order -> gctx[b][h][32][32]; s[b][h][d] = sum_e gctx[d][e] * ctxn[d][e]
__global__ void __launch_bounds__(256)
attn_outer_merge_kernel(const float* __restrict__ partials, const float* __restrict__ ctxn, float* __restrict__ gctx, float* __restrict__ sdot,
                        int chunks) {
    __shared__ float s_p[1024];
    const int head = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* pb = partials + ((size_t)b * 4 + head) * chunks * 1024;
    for (int i = tid; i < 1024; i += 256) {
        float s = 0.f;
        for (int c = 0; c < chunks; ++c) s += pb[(size_t)c * 1024 + i];
        gctx[((size_t)b * 4 + head) * 1024 + i] = s;
        s_p[i] = s * ctxn[((size_t)b * 4 + head) * 1024 + i];
    }
    __syncthreads();
    if (tid < 32) {
        float s = 0.f;
        for (int e = 0; e < 32; ++e) s += s_p[tid * 32 + e];
        sdot[((size_t)b * 4 + head) * 32 + tid] = s;
    }
}

// Per position and head (ctxn = normalised context sum_n p v, ml = global max m[32] and sum l[32] of the key softmax):
//   p[d]   = exp(k[d] - m[d]) / l[d]
//   g_q[d] = sum_e ctxn[d][e] * go[e]          g_v[e] = sum_d p[d] * gctx[d][e]
//   g_p[d] = sum_e gctx[d][e] * v[e]           g_k[d] = p[d] * (g_p[d] - s[d])
// grid (ceil(n/64), 4, B), 256 threads: 64 positions x 4 threads; thread j of a position owns d (and e) = j*8 .. j*8+7.
template <typename T>
__global__ void __launch_bounds__(256)
attn_pos_bwd_kernel(const T* __restrict__ kv, const T* __restrict__ go, const float* __restrict__ ctxn, const float* __restrict__ gctx,
                    const float* __restrict__ ml, const float* __restrict__ sdot, T* __restrict__ gq, T* __restrict__ gkv, int n,
                    const T* __restrict__ q, T* __restrict__ ao) {
    __shared__ float s_ctx[32 * 33], s_g[32 * 33], s_m[32], s_il[32], s_s[32];
    const int tid = threadIdx.x, head = blockIdx.y, b = blockIdx.z;
    const size_t bh = (size_t)b * 4 + head;
    for (int i = tid; i < 1024; i += 256) {
        s_ctx[(i >> 5) * 33 + (i & 31)] = ctxn[bh * 1024 + i];
        s_g[(i >> 5) * 33 + (i & 31)] = gctx[bh * 1024 + i];
    }
    if (tid < 32) { s_m[tid] = ml[bh * 64 + tid]; s_il[tid] = 1.0f / ml[bh * 64 + 32 + tid]; s_s[tid] = sdot[bh * 32 + tid]; }
    __syncthreads();
    const int pos = blockIdx.x * 64 + (tid >> 2), j = tid & 3;
    if (pos >= n) return;                                              // the 4 threads of a position leave together
    const size_t base = (size_t)b * n + pos;
    float kk[32], vv[32], gg[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        Act<T>::load8(kv + base * 256 + head * 32 + c * 8, *reinterpret_cast<float(*)[8]>(&kk[c * 8]));
        Act<T>::load8(kv + base * 256 + 128 + head * 32 + c * 8, *reinterpret_cast<float(*)[8]>(&vv[c * 8]));
        Act<T>::load8(go + base * 128 + head * 32 + c * 8, *reinterpret_cast<float(*)[8]>(&gg[c * 8]));
    }
    float p[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) p[d] = expf(kk[d] - s_m[d]) * s_il[d];
    float oq[8], ok[8], ov[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = j * 8 + i;                                       // also used as e for g_v
        float aq = 0.f, ap = 0.f, av = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            aq = fmaf(s_ctx[d * 33 + e], gg[e], aq);
            ap = fmaf(s_g[d * 33 + e], vv[e], ap);
            av = fmaf(p[e], s_g[e * 33 + d], av);                      // sum over d' = e of p[d'] * gctx[d'][d]
        }
        oq[i] = aq;
        ok[i] = p[d] * (ap - s_s[d]);
        ov[i] = av;
    }
    Act<T>::store8(gq + base * 128 + head * 32 + j * 8, oq);
    Act<T>::store8(gkv + base * 256 + head * 32 + j * 8, ok);
    Act<T>::store8(gkv + base * 256 + 128 + head * 32 + j * 8, ov);
    if (ao) {                                                          // attention output before to_out (needed for dWout): ao[e] = sum_d ctxn[d][e] q[d]
        float qq[32], oa[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) Act<T>::load8(q + base * 128 + head * 32 + c * 8, *reinterpret_cast<float(*)[8]>(&qq[c * 8]));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ee = j * 8 + i;
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 32; ++d) a = fmaf(s_ctx[d * 33 + ee], qq[d], a);
            oa[i] = a;
        }
        Act<T>::store8(ao + base * 128 + head * 32 + j * 8, oa);
    }
}

// ------------------------------------------------------------------------------------------------ dgrad weight packing
// 3x3 stride-1 dgrad: rows (tap' * Cn + ci), cols co, value w[co][ci_off + ci][2 - ky'][2 - kx']   (Cn input channels of this source)
template <typename WT>
__global__ void pack_dgrad3_kernel(const float* __restrict__ w, WT* __restrict__ out, int Cout, int Cin, int ci_off, int Cn) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)9 * Cn * Cout) return;
    const int co = (int)(i % Cout);
    const size_t r = i / Cout;
    const int ci = (int)(r % Cn), tap = (int)(r / Cn);
    const int ky = 2 - tap / 3, kx = 2 - tap % 3;
    Act<WT>::st(out + i, w[(((size_t)co * Cin + ci_off + ci) * 3 + ky) * 3 + kx]);
}
// 1x1 transposed (optionally scaled): rows ci (ci_off + ..Cn), cols co
template <typename WT>
__global__ void pack_t1_kernel(const float* __restrict__ w, WT* __restrict__ out, int Cout, int Cin, int ci_off, int Cn, float scale) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)Cn * Cout) return;
    const int co = (int)(i % Cout), ci = (int)(i / Cout);
    Act<WT>::st(out + i, scale * w[(size_t)co * Cin + ci_off + ci]);
}
// Downsample (3x3 s2 p1) dgrad as a 4-phase transposed conv: phase (py,px), tap t = ty*2+tx; along each axis parity 0 has the
// single tap (d = 0, k = 1), parity 1 the taps (d = +1, k = 0) and (d = 0, k = 2).  rows ((ph*4 + t) * C + ci), cols co; unused
// taps are zero rows.
template <typename WT>
__global__ void pack_down_dgrad_kernel(const float* __restrict__ w, WT* __restrict__ out, int C) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)16 * C * C) return;
    const int co = (int)(i % C);
    const size_t r = i / C;
    const int ci = (int)(r % C), pt = (int)(r / C), ph = pt >> 2, t = pt & 3;
    const int py = ph >> 1, px = ph & 1, ty = t >> 1, tx = t & 1;
    // axis helper: parity 0 -> tap 0: k = 1, tap 1: none; parity 1 -> tap 0: k = 0 (d = +1), tap 1: k = 2 (d = 0)
    const int ky = py == 0 ? (ty == 0 ? 1 : -1) : (ty == 0 ? 0 : 2);
    const int kx = px == 0 ? (tx == 0 ? 1 : -1) : (tx == 0 ? 0 : 2);
    float v = 0.f;
    if (ky >= 0 && kx >= 0) v = w[(((size_t)co * C + ci) * 3 + ky) * 3 + kx];
    Act<WT>::st(out + i, v);
}
// Upsample (ConvTranspose 4x4 s2 p1, weight (Cin, Cout, 4, 4)) dgrad = 16-tap stride-2 conv: rows ((ky*4+kx) * C + ci), cols co
template <typename WT>
__global__ void pack_up_dgrad_kernel(const float* __restrict__ w, WT* __restrict__ out, int C) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)16 * C * C) return;
    const int co = (int)(i % C);
    const size_t r = i / C;
    const int ci = (int)(r % C), tap = (int)(r / C);
    Act<WT>::st(out + i, w[((size_t)ci * C + co) * 16 + tap]);
}

inline unsigned int nblk(size_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

}  // namespace

int gn_bwd_blocks(int H, int W) {
    const int hw = H * W;
    int blocks = (hw + 127) / 128;                    // ~128 pixels per CTA: a 20 x 100 level still gives 16 CTAs per sample
    if (blocks < 1) blocks = 1;
    if (blocks > 256) blocks = 256;
    return blocks;
}

int gn_bwd(ActKind act, const GnBwdArgs& a, cudaStream_t s) {
    GTTS_REQUIRE(a.C == 64 || a.C == 128 || a.C == 256, "gn_bwd: C must be 64, 128 or 256");
    const int blocks = gn_bwd_blocks(a.H, a.W);
    dim3 grid(blocks, a.B);
    if (act == ACT_F32) {
        gn_bwd_stats_kernel<float><<<grid, 256, 0, s>>>(a, blocks);
        gn_bwd_apply_kernel<float><<<grid, 256, 0, s>>>(a, blocks, blocks);
    } else {
        gn_bwd_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a, blocks);
        gn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a, blocks, blocks);
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int final_bwd(ActKind act, const float* v, const float* wf, const float* mask, void* ghf, int B, int H, int W, cudaStream_t s) {
    const size_t n = (size_t)B * H * W * 8;
    if (act == ACT_F32) final_bwd_kernel<float><<<nblk(n, 256), 256, 0, s>>>(v, wf, mask, (float*)ghf, B, H, W);
    else final_bwd_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(v, wf, mask, (__nv_bfloat16*)ghf, B, H, W);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int first_bwd(ActKind act, const void* graw1, const void* gres, const float* w1t, const float* wres, const float* mask, float* gx,
              float* gmu, float* gs, int B, int H, int W, int cin, cudaStream_t s) {
    const size_t n = (size_t)B * H * W;
    if (act == ACT_F32) first_bwd_kernel<float><<<nblk(n, 128), 128, 0, s>>>((const float*)graw1, (const float*)gres, w1t, wres, mask, gx, gmu, gs, B, H, W, cin);
    else first_bwd_kernel<__nv_bfloat16><<<nblk(n, 128), 128, 0, s>>>((const __nv_bfloat16*)graw1, (const __nv_bfloat16*)gres, w1t, wres, mask, gx, gmu, gs, B, H, W, cin);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int mask_mul(ActKind act, const void* in, const float* mask, void* out, int B, int H, int W, int C, cudaStream_t s) {
    const size_t n = (size_t)B * H * W * (C / 8);
    if (act == ACT_F32) mask_mul_kernel<float><<<nblk(n, 256), 256, 0, s>>>((const float*)in, mask, (float*)out, B, H * W, W, C);
    else mask_mul_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>((const __nv_bfloat16*)in, mask, (__nv_bfloat16*)out, B, H * W, W, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int add_tensors(ActKind act, const void* a, const void* b, void* out, size_t numel, cudaStream_t s) {
    const size_t n = numel / 8;
    if (act == ACT_F32) add_kernel<float><<<nblk(n, 256), 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, n);
    else add_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, n);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attn_outer(ActKind act, const void* q, const void* go, float* partials, const float* ctxn, float* gctx, float* sdot, int B, int n,
               int chunks, int chunk_len, cudaStream_t s) {
    dim3 grid(chunks, 4, B);
    if (act == ACT_F32) attn_outer_kernel<float><<<grid, 256, 0, s>>>((const float*)q, (const float*)go, partials, n, chunks, chunk_len);
    else attn_outer_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)go, partials, n, chunks, chunk_len);
    attn_outer_merge_kernel<<<dim3(4, B), 256, 0, s>>>(partials, ctxn, gctx, sdot, chunks);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attn_pos_bwd(ActKind act, const void* kv, const void* go, const float* ctxn, const float* gctx, const float* ml, const float* sdot,
                 void* gq, void* gkv, int B, int n, const void* q, void* ao, cudaStream_t s) {
    dim3 grid((n + 63) / 64, 4, B);
    if (act == ACT_F32)
        attn_pos_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)kv, (const float*)go, ctxn, gctx, ml, sdot, (float*)gq, (float*)gkv, n,
                                                       (const float*)q, (float*)ao);
    else
        attn_pos_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)kv, (const __nv_bfloat16*)go, ctxn, gctx, ml, sdot,
                                                               (__nv_bfloat16*)gq, (__nv_bfloat16*)gkv, n, (const __nv_bfloat16*)q, (__nv_bfloat16*)ao);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int pack_dgrad3(ActKind wkind, const float* w, void* out, int Cout, int Cin, int ci_off, int Cn, cudaStream_t s) {
    const size_t n = (size_t)9 * Cn * Cout;
    if (wkind == ACT_F32) pack_dgrad3_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)out, Cout, Cin, ci_off, Cn);
    else pack_dgrad3_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)out, Cout, Cin, ci_off, Cn);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_t1(ActKind wkind, const float* w, void* out, int Cout, int Cin, int ci_off, int Cn, float scale, cudaStream_t s) {
    const size_t n = (size_t)Cn * Cout;
    if (wkind == ACT_F32) pack_t1_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)out, Cout, Cin, ci_off, Cn, scale);
    else pack_t1_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)out, Cout, Cin, ci_off, Cn, scale);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_down_dgrad(ActKind wkind, const float* w, void* out, int C, cudaStream_t s) {
    const size_t n = (size_t)16 * C * C;
    if (wkind == ACT_F32) pack_down_dgrad_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)out, C);
    else pack_down_dgrad_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)out, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int pack_up_dgrad(ActKind wkind, const float* w, void* out, int C, cudaStream_t s) {
    const size_t n = (size_t)16 * C * C;
    if (wkind == ACT_F32) pack_up_dgrad_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)out, C);
    else pack_up_dgrad_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)out, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
