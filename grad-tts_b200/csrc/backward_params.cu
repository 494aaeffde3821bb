// Parameter gradients of the score network (training backward; reference: loss.backward() through model/diffusion.py, reached from
// Diffusion.loss_t :274-281 / compute_loss :283-287) -- SURVEY 8(f) rank 2, second half.  backward.cu provides the gradients of every
// activation (d loss / d conv output, d loss / d Block output, ...); the kernels here reduce them against the saved forward tensors:
//   conv weights   dW[tap][co][ci] = sum_pixels g_out[p][co] * x[p + tap][ci]      (wgrad: an implicit GEMM with K = pixels, generic over
//                                                                                   ConvGeom: 3x3, 1x1, stride 2, transposed phases)
//   biases         db[co]          = sum_pixels g_out[p][co]                        (col_sums)
//   GroupNorm      dgamma[c], dbeta[c] = sum g_y * mask * Mish'(n) * {xhat, 1}      (gn_bwd statistics pass + gn_param_reduce)
//   final conv, first conv / first res_conv (tiny K): dedicated kernels
// All reductions are two-stage and run in a fixed order (deterministic); accumulation in fp32, inputs in the activation type.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

__device__ __forceinline__ float mish_grad_p(float n) {          // as in backward.cu
    if (n > 20.0f) {
        const float t = tanhf(n);
        return t + n * (1.0f - t * t);
    }
    const float sp = log1pf(expf(n));
    const float t = tanhf(sp);
    const float sg = 1.0f / (1.0f + expf(-n));
    return t + n * (1.0f - t * t) * sg;
}

// ------------------------------------------------------------------------------------------------ wgrad
// One CTA: (phase-tap pt, 64 co, 64 ci, pixel slice).  Output-grid pixels q = (b, j, i) of phase ph; g_out pixel (j*out_step + oy,
// i*out_step + ox); x pixel (j*stride + dy, i*stride + dx) (zero outside the image).  256 threads, 4 co x 4 ci per thread.
// partial[(slice * n_pt + pt) * Cout * Cin + co * Cin + ci]
constexpr int kWgPix = 32;

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_kernel(ConvGeom g, const T* __restrict__ gout, const T* __restrict__ x0, const T* __restrict__ x1, float* __restrict__ partial,
             int slices, long pix_per_slice) {
    __shared__ __align__(16) float gs[kWgPix][68];
    __shared__ __align__(16) float xs[kWgPix][68];
    const int pt = blockIdx.x, ph = pt / g.ntaps, tap = pt % g.ntaps;
    const int co0 = blockIdx.y * 64;
    const int cin_tot = g.Cin0 + g.Cin1;
    const int nci = cin_tot / 64;
    const int ci0 = (blockIdx.z % nci) * 64, slice = blockIdx.z / nci;
    const long npix = (long)g.B * g.Hg * g.Wg;
    const long p_lo = slice * pix_per_slice, p_hi = min(npix, p_lo + pix_per_slice);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lp = tid >> 3, lv = (tid & 7) * 8;                      // loader: pixel lp of the group, channels lv..lv+7
    const T* xsrc = ci0 < g.Cin0 ? x0 : x1;
    const int xc = ci0 < g.Cin0 ? g.Cin0 : g.Cin1, xoff = ci0 < g.Cin0 ? ci0 : ci0 - g.Cin0;
    for (long p0 = p_lo; p0 < p_hi; p0 += kWgPix) {
        const long q = p0 + lp;
        float gv[8], xv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { gv[k] = 0.f; xv[k] = 0.f; }
        if (q < p_hi) {
            const int b = (int)(q / ((long)g.Hg * g.Wg));
            const int rem = (int)(q % ((long)g.Hg * g.Wg)), j = rem / g.Wg, i = rem % g.Wg;
            const int oh = j * g.out_step + g.oy[ph], ow = i * g.out_step + g.ox[ph];
            const int ih = j * g.stride + g.dy[ph][tap], iw = i * g.stride + g.dx[ph][tap];
            if (ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win) {
                Act<T>::load8(gout + (((size_t)b * g.Hout + oh) * g.Wout + ow) * g.Cout + co0 + lv, gv);
                Act<T>::load8(xsrc + (((size_t)b * g.Hin + ih) * g.Win + iw) * xc + xoff + lv, xv);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) { gs[lp][lv + k] = gv[k]; xs[lp][lv + k] = xv[k]; }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kWgPix; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&gs[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&xs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
    float* o = partial + ((size_t)slice * (g.nphase * g.ntaps) + pt) * g.Cout * cin_tot;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[(size_t)(co0 + ty * 4 + i) * cin_tot + ci0 + tx * 4 + j] = acc[i][j];
}

// bf16 activations: the same implicit GEMM on the tensor cores (mma.sync m16n8k16, fp32 accumulate).  Both operands are
// [pixel][channel] rows, i.e. K-outermost, so both fragments come from ldmatrix.trans.  Same grid, pixel slicing and partial layout as
// wgrad_kernel (the CUDA-core kernel above, which stays for fp32 activations): CTA = (tap, 64 output channels, 64 input channels,
// pixel slice), 8 warps x (16 co x 32 ci), 32 pixels per step with the next step's global loads issued before the MMAs.
// Measured before this kernel: wgrad was 71 % of the training step at 26 TFLOP/s (profiles/r02_profile_train_v1.txt).
__device__ __forceinline__ void wg_ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void wg_mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kWgMmaPix = 64;                                         // pixels per step of the tensor-core kernel

__global__ void __launch_bounds__(256)
wgrad_mma_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ gout, const __nv_bfloat16* __restrict__ x0,
                 const __nv_bfloat16* __restrict__ x1, float* __restrict__ partial, int slices, long pix_per_slice) {
    constexpr int kPitch = 72;                                        // bf16 per pixel row: 64 + 8 pad (144 B, conflict-free ldmatrix)
    __shared__ __align__(16) __nv_bfloat16 gs[2][kWgMmaPix * kPitch];  // double-buffered: one block barrier per step
    __shared__ __align__(16) __nv_bfloat16 xs[2][kWgMmaPix * kPitch];
    const int pt = blockIdx.x, ph = pt / g.ntaps, tap = pt % g.ntaps;
    const int co0 = blockIdx.y * 64;
    const int cin_tot = g.Cin0 + g.Cin1;
    const int nci = cin_tot / 64;
    const int ci0 = (blockIdx.z % nci) * 64, slice = blockIdx.z / nci;
    const long npix = (long)g.B * g.Hg * g.Wg;
    const long p_lo = slice * pix_per_slice, p_hi = min(npix, p_lo + pix_per_slice);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
    const int j = lane >> 3, r = lane & 7;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lp = tid >> 3, lv = (tid & 7) * 8;                      // loader: pixels lp and lp + 32 of the step, channels lv..lv+7
    const __nv_bfloat16* xsrc = ci0 < g.Cin0 ? x0 : x1;
    const int xc = ci0 < g.Cin0 ? g.Cin0 : g.Cin1, xoff = ci0 < g.Cin0 ? ci0 : ci0 - g.Cin0;
    const long hw = (long)g.Hg * g.Wg;
    uint4 gv[2], xv[2];
    // (sample, row, column) of this thread's two pixels, advanced by 64 pixels per step without divisions (the first version
    // recomputed them with a 64-bit and a 32-bit division per pixel per step: more integer work than tensor work)
    int pb[2], pj[2], pi[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const long q = p_lo + lp + 32 * u;
        pb[u] = (int)(q / hw);
        const int rem = (int)(q - (long)pb[u] * hw);
        pj[u] = rem / g.Wg;
        pi[u] = rem - pj[u] * g.Wg;
    }
    auto fetch = [&](long p0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long q = p0 + lp + 32 * u;
            gv[u] = make_uint4(0u, 0u, 0u, 0u);
            xv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (q < p_hi) {
                const int b = pb[u], jj = pj[u], ii = pi[u];
                const int oh = jj * g.out_step + g.oy[ph], ow = ii * g.out_step + g.ox[ph];
                const int ih = jj * g.stride + g.dy[ph][tap], iw = ii * g.stride + g.dx[ph][tap];
                if (ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win) {
                    gv[u] = __ldg(reinterpret_cast<const uint4*>(gout + (((size_t)b * g.Hout + oh) * g.Wout + ow) * g.Cout + co0 + lv));
                    xv[u] = __ldg(reinterpret_cast<const uint4*>(xsrc + (((size_t)b * g.Hin + ih) * g.Win + iw) * xc + xoff + lv));
                }
            }
            pi[u] += kWgMmaPix;
            while (pi[u] >= g.Wg) { pi[u] -= g.Wg; ++pj[u]; }
            while (pj[u] >= g.Hg) { pj[u] -= g.Hg; ++pb[u]; }
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            *reinterpret_cast<uint4*>(&gs[buf][(lp + 32 * u) * kPitch + lv]) = gv[u];
            *reinterpret_cast<uint4*>(&xs[buf][(lp + 32 * u) * kPitch + lv]) = xv[u];
        }
    };
    if (p_lo < p_hi) { fetch(p_lo); stage(0); }
    __syncthreads();
    int buf = 0;
    for (long p0 = p_lo; p0 < p_hi; p0 += kWgMmaPix, buf ^= 1) {
        const bool more = p0 + kWgMmaPix < p_hi;
        if (more) fetch(p0 + kWgMmaPix);                              // in flight during the MMAs
#pragma unroll
        for (int ks = 0; ks < kWgMmaPix / 16; ++ks) {
            const int k0 = ks * 16;
            uint32_t af[4], bf[4];
            wg_ldmatrix_x4_trans(af, &gs[buf][(k0 + (j >> 1) * 8 + r) * kPitch + m0 + (j & 1) * 8]);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                wg_ldmatrix_x4_trans(bf, &xs[buf][(k0 + (j & 1) * 8 + r) * kPitch + n0 + (np * 2 + (j >> 1)) * 8]);
                wg_mma_bf16(acc[2 * np], af, bf[0], bf[1]);
                wg_mma_bf16(acc[2 * np + 1], af, bf[2], bf[3]);
            }
        }
        if (more) stage(buf ^ 1);                                     // nobody reads that buffer in this step
        __syncthreads();
    }
    float* o = partial + ((size_t)slice * (g.nphase * g.ntaps) + pt) * g.Cout * cin_tot;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        const int n = ci0 + n0 + nb * 8 + 2 * tq;
        *reinterpret_cast<float2*>(&o[(size_t)(co0 + m0 + gq) * cin_tot + n]) = make_float2(acc[nb][0], acc[nb][1]);
        *reinterpret_cast<float2*>(&o[(size_t)(co0 + m0 + gq + 8) * cin_tot + n]) = make_float2(acc[nb][2], acc[nb][3]);
    }
}

// sums the slices in order and scatters into the PyTorch layout.  kind 0: conv (Cout, Cin, kh, kw) with packed tap index = ky*kw+kx
// (pt = tap, one phase); kind 1: ConvTranspose2d (Cin, Cout, 4, 4), pt = phase*4 + tap with the (py,px)/(ty,tx) -> (ky,kx) table of
// pack_convT_kernel; kind 2: rows-only matrix [n_pt * Cout][Cin] (1x1 and internal uses), written as is at row offset.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, int slices, int n_pt, int Cout, int Cin,
                                    int kind, int kw, float scale, int accumulate) {
    const size_t n = (size_t)n_pt * Cout * Cin;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += partial[(size_t)k * n + i];
    s *= scale;
    const int ci = (int)(i % Cin);
    const size_t r = i / Cin;
    const int co = (int)(r % Cout), pt = (int)(r / Cout);
    size_t d;
    if (kind == 0) {
        d = ((size_t)co * Cin + ci) * n_pt + pt;                      // (co, ci, ky, kx), pt = ky*kw + kx
    } else if (kind == 1) {
        const int phase = pt >> 2, tap = pt & 3, py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
        const int kk[2][2] = {{1, 3}, {2, 0}};
        d = (((size_t)ci * Cout + co) * 4 + kk[py][ty]) * 4 + kk[px][tx];
    } else {
        d = i;
    }
    (void)kw;
    dst[d] = accumulate ? dst[d] + s : s;
}

// ------------------------------------------------------------------------------------------------ column sums (bias gradients)
// partial[(b * blocks + block)][C] = sum over the block's pixels of sample b of g[p][c] (* mask[b][w]); then reduce per group of rows.
// grid (blocks, nb), 256 threads.  nb = 1 with pix = all pixels gives whole-batch sums.
template <typename T>
__global__ void __launch_bounds__(256)
col_sums_kernel(const T* __restrict__ gsrc, const float* __restrict__ mask, float* __restrict__ partial, long pix_per_sample, int W, int C,
                long pix_per_block) {
    const int C8 = C >> 3, vec = threadIdx.x % C8, pslot = threadIdx.x / C8, pstep = 256 / C8;
    const int b = blockIdx.y;
    const long p_lo = blockIdx.x * pix_per_block, p_hi = min(pix_per_sample, p_lo + pix_per_block);
    const T* base = gsrc + (size_t)b * pix_per_sample * C;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // four pixels per round, their loads issued together (one load per round left the kernel latency-bound: 20 us for 28 MB)
    long p = p_lo + pslot;
    for (; p + 3 * pstep < p_hi; p += 4 * pstep) {
        float v[4][8], m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long q = p + u * pstep;
            m[u] = mask ? mask[(size_t)b * W + (int)(q % W)] : 1.0f;
            Act<T>::load8(base + (size_t)q * C + vec * 8, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = fmaf(v[u][j], m[u], s[j]);
    }
    for (; p < p_hi; p += pstep) {
        float v[8];
        const float m = mask ? mask[(size_t)b * W + (int)(p % W)] : 1.0f;
        Act<T>::load8(base + (size_t)p * C + vec * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = fmaf(v[j], m, s[j]);
    }
    __shared__ float sm[256 * 8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = s[j];
    __syncthreads();
    if (threadIdx.x < C) {
        const int c = threadIdx.x, v = c >> 3, j = c & 7;
        float t = 0.f;
        for (int ps = 0; ps < pstep; ++ps) t += sm[(ps * C8 + v) * 8 + j];
        partial[((size_t)b * gridDim.x + blockIdx.x) * C + c] = t;
    }
}
// dst[g * n + i] (+)= scale * sum_{k < blocks} partial[(g * blocks + k) * n + i]   (dst row stride dst_ld)
// dst[gq * dst_ld + c] = scale * sum_k partial[(gq * blocks + k) * n + c]  (c < n; columns >= split go to dst2[c - split]).
// Block = 32 columns x 32 row slots: a thread adds every 32nd row with four independent accumulators, then the 32 slots are added
// in slot order in shared memory (fixed order: deterministic).  The first version gave every column ONE thread that walked all rows
// serially: 7-40 us per call for a few megabytes, more than the tensor-core weight-gradient kernels they follow.
// grid (ceil(n / 32), groups), 1024 threads.
__global__ void __launch_bounds__(1024)
rows_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, float* __restrict__ dst2, int split, int blocks, int n,
                   int dst_ld, float scale, int accumulate) {
    __shared__ float sm[32][33];
    const int cx = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx, gq = blockIdx.y;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < n) {
        const float* base = partial + (size_t)gq * blocks * n + c;
        int k = slot;
        for (; k + 96 < blocks; k += 128) {
            a0 += base[(size_t)k * n]; a1 += base[(size_t)(k + 32) * n];
            a2 += base[(size_t)(k + 64) * n]; a3 += base[(size_t)(k + 96) * n];
        }
        for (; k < blocks; k += 32) a0 += base[(size_t)k * n];
    }
    sm[slot][cx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (slot == 0 && c < n) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += sm[q][cx];
        t *= scale;
        float* d = (dst2 && c >= split) ? dst2 + (size_t)gq * dst_ld + (c - split) : dst + (size_t)gq * dst_ld + c;
        *d = accumulate ? *d + t : t;
    }
}
// (GroupNorm affine gradients: the per-channel partial sums come out of gn_bwd's statistics pass, backward.cu; gn_param_reduce below
// adds the rows)
// ------------------------------------------------------------------------------------------------ final conv (64 -> 1) gradients
// score = (sum_c wf[c] * hf[c] + bf) * mask with hf = Mish(GN(rawf)) * mask:  dwf[c] = sum v * mask * hf[c],  dbf = sum v * mask.
// partial[block][65]
template <typename T>
__global__ void __launch_bounds__(256)
final_param_kernel(const T* __restrict__ rawf, const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ v, const float* __restrict__ mask, float* __restrict__ partial, int B, int H, int W,
                   long pix_per_block) {
    const int vec = threadIdx.x & 7, pslot = threadIdx.x >> 3;
    const long npix = (long)B * H * W;
    const long p_lo = blockIdx.x * pix_per_block, p_hi = min(npix, p_lo + pix_per_block);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sb = 0.f;
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = gamma[vec * 8 + j]; be[j] = beta[vec * 8 + j]; }
    for (long p = p_lo + pslot; p < p_hi; p += 32) {
        const int b = (int)(p / ((long)H * W)), w = (int)(p % W);
        const float m = mask[(size_t)b * W + w];
        if (m == 0.f) continue;
        const float vv = v[p] * m;
        const float mean = stats[(b * 8 + vec) * 2], rstd = stats[(b * 8 + vec) * 2 + 1];     // 8 channels per group at C = 64
        float r[8];
        Act<T>::load8(rawf + (size_t)p * 64 + vec * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float n = fmaf(ga[j], (r[j] - mean) * rstd, be[j]);
            const float sp = n > 20.f ? n : log1pf(expf(n));
            s[j] = fmaf(vv, n * tanhf(sp), s[j]);
        }
        if (vec == 0) sb += vv;
    }
    __shared__ float sm[256 * 9];
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[threadIdx.x * 9 + j] = s[j];
    sm[threadIdx.x * 9 + 8] = sb;
    __syncthreads();
    if (threadIdx.x < 65) {
        float t = 0.f;
        if (threadIdx.x < 64) {
            const int vv = threadIdx.x >> 3, j = threadIdx.x & 7;
            for (int ps = 0; ps < 32; ++ps) t += sm[(ps * 8 + vv) * 9 + j];
        } else {
            for (int ps = 0; ps < 32; ++ps) t += sm[(ps * 8) * 9 + 8];
        }
        partial[(size_t)blockIdx.x * 65 + threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------ first conv / first res_conv gradients
// in[ci] = plane ci of [mu, x, (s)] * mask.   dW1[co][ci][ky][kx] = sum_p g_raw1[p][co] * in[ci][p + (ky-1, kx-1)]
//                                             dWres[co][ci]       = sum_p g_res[p][co]  * in[ci][p];   db1 / dbres = column sums.
// One CTA per pixel block of one sample; thread = (co, k-group); partial[block][64 * (cin*9 + cin + 2)] laid out
// [co][ cin*9 (W1) | cin (Wres) | b1 | bres ].
template <typename T>
__global__ void __launch_bounds__(256)
first_param_kernel(const T* __restrict__ graw1, const T* __restrict__ gres, const float* __restrict__ mu, const float* __restrict__ x,
                   const float* __restrict__ splane, const float* __restrict__ mask, float* __restrict__ partial, int B, int H, int W, int cin,
                   int blocks_per_sample) {
    constexpr int kPix = 64;
    __shared__ float s_g1[kPix][65], s_gr[kPix][65];
    __shared__ float s_in[kPix][3 * 9];                               // [pixel][ci*9 + tap]: masked input at the tap position
    const int b = blockIdx.y, HW = H * W;
    const int per_block = (HW + blocks_per_sample - 1) / blocks_per_sample;
    const int p_lo = blockIdx.x * per_block, p_hi = min(HW, p_lo + per_block);
    const int nk = cin * 9 + cin + 2;                                 // outputs per co
    const int tid = threadIdx.x, co = tid & 63, kq = tid >> 6;        // 4 threads per co split the k range
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // k indices of this thread: kq, kq + 4, ... (< nk <= 32)
    for (int p0 = p_lo; p0 < p_hi; p0 += kPix) {
        __syncthreads();
        for (int i = tid; i < kPix * 64; i += 256) {
            const int pp = i >> 6, c = i & 63, p = p0 + pp;
            float a = 0.f, r = 0.f;
            if (p < p_hi) {
                a = Act<T>::ld(graw1 + ((size_t)b * HW + p) * 64 + c);
                r = Act<T>::ld(gres + ((size_t)b * HW + p) * 64 + c) * mask[(size_t)b * W + p % W];
            }
            s_g1[pp][c] = a; s_gr[pp][c] = r;
        }
        for (int i = tid; i < kPix * 27; i += 256) {
            const int pp = i / 27, k = i % 27, ci = k / 9, tap = k % 9, p = p0 + pp;
            float v = 0.f;
            if (p < p_hi && ci < cin) {
                const int h = p / W + tap / 3 - 1, w = p % W + tap % 3 - 1;
                if (h >= 0 && h < H && w >= 0 && w < W) {
                    const float m = mask[(size_t)b * W + w];
                    const size_t q = ((size_t)b * H + h) * W + w;
                    v = (ci == 0 ? mu[q] : (ci == 1 ? x[q] : splane[(size_t)b * H + h])) * m;
                }
            }
            s_in[pp][k] = v;
        }
        __syncthreads();
        const int npx = min(kPix, p_hi - p0);
        for (int pp = 0; pp < npx; ++pp) {
            const float g1 = s_g1[pp][co], gr = s_gr[pp][co];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = kq + 4 * i;
                if (k >= nk) break;
                float operand;
                if (k < cin * 9) operand = g1 * s_in[pp][k];
                else if (k < cin * 9 + cin) operand = gr * s_in[pp][(k - cin * 9) * 9 + 4];      // centre tap = the pixel itself
                else if (k == cin * 9 + cin) operand = g1;
                else operand = gr;
                acc[i] += operand;
            }
        }
    }
    float* o = partial + ((size_t)b * blocks_per_sample + blockIdx.x) * 64 * nk + (size_t)co * nk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = kq + 4 * i;
        if (k < nk) o[k] = acc[i];
    }
}

// to_out (1x1, C x 128, bias) behind the Rezero gate g:  y = g * (Wout ao + bout) + x.  From A = sum_n G (x) ao and s = sum_n G:
// dWout = g A, dbout = g s, dg = <Wout, A> + <bout, s>   (fixed-order reduction by one CTA)
__global__ void __launch_bounds__(256)
attn_out_grads_kernel(const float* __restrict__ A, const float* __restrict__ sv, const float* __restrict__ wout, const float* __restrict__ bout,
                      float g, float* __restrict__ dwout, float* __restrict__ dbout, float* __restrict__ dg, int C) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < C * 128; i += 256) {
        const float a = A[i];
        dwout[i] = g * a;
        acc += (double)wout[i] * (double)a;
    }
    for (int i = threadIdx.x; i < C; i += 256) {
        const float v = sv[i];
        dbout[i] = g * v;
        acc += (double)bout[i] * (double)v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 256; ++i) t += red[i];
        dg[0] = (float)t;
    }
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n, int accumulate) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = accumulate ? dst[i] + src[i] : src[i];
}

inline unsigned int nblk(size_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

}  // namespace

int attn_out_grads(const float* A, const float* sv, const float* wout, const float* bout, float g, float* dwout, float* dbout, float* dg,
                   int C, cudaStream_t s) {
    attn_out_grads_kernel<<<1, 256, 0, s>>>(A, sv, wout, bout, g, dwout, dbout, dg, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int accumulate_floats(float* dst, const float* src, size_t n, int accumulate, cudaStream_t s) {
    axpy_kernel<<<nblk(n, 256), 256, 0, s>>>(dst, src, n, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---- host wrappers ---------------------------------------------------------------------------------------------------------------
size_t wgrad_partial_floats(const ConvGeom& g, int* slices_out) {
    const long npix = (long)g.B * g.Hg * g.Wg;
    int slices = (int)((npix + 4095) / 4096);
    if (slices > 64) slices = 64;
    if (slices < 1) slices = 1;
    if (slices_out) *slices_out = slices;
    return (size_t)slices * g.nphase * g.ntaps * g.Cout * (g.Cin0 + g.Cin1);
}

int conv_wgrad(ActKind act, const ConvGeom& g, const void* gout, const void* x0, const void* x1, float* partial, float* dst, int kind,
               float scale, int accumulate, cudaStream_t s) {
    GTTS_REQUIRE(g.Cout % 64 == 0 && g.Cin0 % 64 == 0 && g.Cin1 % 64 == 0, "conv_wgrad: channels must be multiples of 64");
    int slices;
    wgrad_partial_floats(g, &slices);
    const long npix = (long)g.B * g.Hg * g.Wg;
    const long pps = ((npix + slices - 1) / slices + kWgMmaPix - 1) / kWgMmaPix * kWgMmaPix;   // multiple of both kernels' step
    const int n_pt = g.nphase * g.ntaps, cin = g.Cin0 + g.Cin1;
    dim3 grid(n_pt, g.Cout / 64, (cin / 64) * slices);
    if (act == ACT_F32) wgrad_kernel<float><<<grid, 256, 0, s>>>(g, (const float*)gout, (const float*)x0, (const float*)x1, partial, slices, pps);
    else if (getenv("GTTS_WGRAD_FFMA"))                              // CUDA-core kernel on bf16 (cross-check / measurement)
        wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(g, (const __nv_bfloat16*)gout, (const __nv_bfloat16*)x0, (const __nv_bfloat16*)x1, partial, slices, pps);
    else
        wgrad_mma_kernel<<<grid, 256, 0, s>>>(g, (const __nv_bfloat16*)gout, (const __nv_bfloat16*)x0, (const __nv_bfloat16*)x1, partial, slices, pps);
    const size_t n = (size_t)n_pt * g.Cout * cin;
    wgrad_reduce_kernel<<<nblk(n, 256), 256, 0, s>>>(partial, dst, slices, n_pt, g.Cout, cin, kind, 3, scale, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int wgrad_reduce(const float* partial, float* dst, int slices, int n_pt, int Cout, int Cin, int kind, float scale, int accumulate,
                 cudaStream_t s) {
    const size_t n = (size_t)n_pt * Cout * Cin;
    wgrad_reduce_kernel<<<nblk(n, 256), 256, 0, s>>>(partial, dst, slices, n_pt, Cout, Cin, kind, 3, scale, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// whole-batch column sums (bias gradients): dst[C]
int col_sums(ActKind act, const void* gsrc, float* partial, float* dst, long npix, int C, float scale, int accumulate, cudaStream_t s) {
    GTTS_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0, "col_sums: C must be 64, 128 or 256");
    int blocks = (int)((npix + 1023) / 1024);
    if (blocks > 256) blocks = 256;
    if (blocks < 1) blocks = 1;
    const long ppb = (npix + blocks - 1) / blocks;
    if (act == ACT_F32) col_sums_kernel<float><<<dim3(blocks, 1), 256, 0, s>>>((const float*)gsrc, nullptr, partial, npix, 1, C, ppb);
    else col_sums_kernel<__nv_bfloat16><<<dim3(blocks, 1), 256, 0, s>>>((const __nv_bfloat16*)gsrc, nullptr, partial, npix, 1, C, ppb);
    rows_reduce_kernel<<<dim3((C + 31) / 32, 1), 1024, 0, s>>>(partial, dst, nullptr, 0, blocks, C, C, scale, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// per-sample masked column sums: dst[b * dst_ld + c] = sum_p g[b][p][c] * mask[b][w]   (the time-bias gradient of a ResnetBlock)
int col_sums_per_sample(ActKind act, const void* gsrc, const float* mask, float* partial, float* dst, int B, int H, int W, int C, int dst_ld,
                        cudaStream_t s) {
    GTTS_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0, "col_sums: C must be 64, 128 or 256");
    const long hw = (long)H * W;
    int blocks = (int)((hw + 1023) / 1024);
    if (blocks > 64) blocks = 64;
    if (blocks < 1) blocks = 1;
    const long ppb = (hw + blocks - 1) / blocks;
    if (act == ACT_F32) col_sums_kernel<float><<<dim3(blocks, B), 256, 0, s>>>((const float*)gsrc, mask, partial, hw, W, C, ppb);
    else col_sums_kernel<__nv_bfloat16><<<dim3(blocks, B), 256, 0, s>>>((const __nv_bfloat16*)gsrc, mask, partial, hw, W, C, ppb);
    rows_reduce_kernel<<<dim3((C + 31) / 32, B), 1024, 0, s>>>(partial, dst, nullptr, 0, blocks, C, dst_ld, 1.0f, 0);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int gn_param_reduce(const float* param_partials, float* dgamma, float* dbeta, int B, int H, int W, int C, cudaStream_t s) {
    const int blocks = gn_bwd_blocks(H, W);
    rows_reduce_kernel<<<dim3((2 * C + 31) / 32, 1), 1024, 0, s>>>(param_partials, dgamma, dbeta, C, blocks * B, 2 * C, 2 * C, 1.0f, 0);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int final_param_grad(ActKind act, const void* rawf, const float* stats, const float* gamma, const float* beta, const float* v,
                     const float* mask, float* partial, float* dwf_dbf /*[65]*/, int B, int H, int W, int accumulate, cudaStream_t s) {
    const long npix = (long)B * H * W;
    int blocks = (int)((npix + 2047) / 2048);
    if (blocks > 256) blocks = 256;
    if (blocks < 1) blocks = 1;
    const long ppb = (npix + blocks - 1) / blocks;
    if (act == ACT_F32) final_param_kernel<float><<<blocks, 256, 0, s>>>((const float*)rawf, stats, gamma, beta, v, mask, partial, B, H, W, ppb);
    else final_param_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)rawf, stats, gamma, beta, v, mask, partial, B, H, W, ppb);
    rows_reduce_kernel<<<dim3(3, 1), 1024, 0, s>>>(partial, dwf_dbf, nullptr, 0, blocks, 65, 65, 1.0f, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int first_param_blocks(int H, int W) {
    int blocks = (H * W + 255) / 256;                                 // 256 pixels (four 64-pixel rounds) per CTA: the rounds are serial
    if (blocks > 64) blocks = 64;
    return blocks < 1 ? 1 : blocks;
}

int first_param_grad(ActKind act, const void* graw1, const void* gres, const float* mu, const float* x, const float* splane,
                     const float* mask, float* partial, float* dst /*[64][cin*9 + cin + 2]*/, int B, int H, int W, int cin, int accumulate,
                     cudaStream_t s) {
    const int blocks = first_param_blocks(H, W);
    const int nk = cin * 9 + cin + 2;
    dim3 grid(blocks, B);
    if (act == ACT_F32) first_param_kernel<float><<<grid, 256, 0, s>>>((const float*)graw1, (const float*)gres, mu, x, splane, mask, partial, B, H, W, cin, blocks);
    else first_param_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)graw1, (const __nv_bfloat16*)gres, mu, x, splane, mask, partial, B, H, W, cin, blocks);
    rows_reduce_kernel<<<dim3((64 * nk + 31) / 32, 1), 1024, 0, s>>>(partial, dst, nullptr, 0, blocks * B, 64 * nk, 64 * nk, 1.0f, accumulate);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
