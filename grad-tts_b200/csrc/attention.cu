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
#include <vector>
#include <cmath>
#include <cstring>

namespace gtts {

namespace {

constexpr int kSub = 128;        // pixels per sub-tile
constexpr int kKPitch = 36;      // floats per k row in smem (16-byte aligned rows)

template <typename T, bool kStrict>
__global__ void __launch_bounds__(256)
attn_ctx_kernel(AttnCtxArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float ks[kSub * kKPitch];
    __shared__ __align__(16) float vs[kSub * 32];
    __shared__ float s_m[32], s_scale[32];
    const int tid = threadIdx.x;
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
}

// ---------------------------------------------------------------------------------------------------------
// bf16 path: the context contraction ctx[d][e] = sum_n p[n][d] v[n][e] on the (legacy-path) tensor cores.
// One CTA = 4 warps = the 4 heads of one pixel chunk; per 64-pixel sub-tile the kv rows (512 B/pixel) are staged
// once in shared memory, each warp turns its 32 k-columns into p = exp(k - running max) in place (bf16), and
// runs P^T V as mma.sync.m16n8k16 (A and B both fetched with ldmatrix.trans from the [pixel][channel] rows).
// HBM-bound by design: kv is read exactly once.
constexpr int kTcSub = 64;
constexpr int kTcPitch = 264;        // bf16 per smem pixel row: 256 + 8 pad (528 B, conflict-free ldmatrix)

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

__global__ void __launch_bounds__(128)
attn_ctx_tc_kernel(AttnCtxArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) __nv_bfloat16 sm_all[];          // two sub-tile buffers (cp.async double buffering)
    const int tid = threadIdx.x, head = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int n0 = chunk * a.chunk_len, n1 = min(a.n, n0 + a.chunk_len);
    const __nv_bfloat16* kv = reinterpret_cast<const __nv_bfloat16*>(a.kv) + (size_t)b * a.n * 256;

    const int cp = lane & 15, par = lane >> 4;          // element-wise phase: columns 2cp, 2cp+1 ; pixel parity
    const int g = lane >> 2, t = lane & 3;              // mma fragment coordinates
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    constexpr int kWordPitch = kTcPitch / 2;
    auto fetch = [&](int p0, __nv_bfloat16* dst) {
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const int idx = it * 128 + tid, px = idx >> 5, c16 = idx & 31, n = p0 + px;
            __nv_bfloat16* d = &dst[px * kTcPitch + c16 * 8];
            if (n < n1) cp_async16(d, kv + (size_t)n * 256 + c16 * 8);
            else if (c16 < 16) *reinterpret_cast<uint4*>(d) = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // k = -inf
            else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);                                                   // v = 0
        }
        cp_async_commit();
    };
    fetch(n0, sm_all);
    int ib = 0;
    for (int p0 = n0; p0 < n1; p0 += kTcSub, ib ^= 1) {
        __nv_bfloat16* sm = sm_all + ib * (kTcSub * kTcPitch);
        if (p0 + kTcSub < n1) { fetch(p0 + kTcSub, sm_all + (ib ^ 1) * (kTcSub * kTcPitch)); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                                 // this sub-tile has landed for every thread
        uint32_t* kw = reinterpret_cast<uint32_t*>(sm) + head * 16 + cp;   // word (2 bf16) of my k columns, pixel 0
        // ---- running max of my two k columns
        float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll 8
        for (int i = 0; i < kTcSub / 2; ++i) {
            const uint32_t w = kw[(2 * i + par) * kWordPitch];
            x0 = fmaxf(x0, __uint_as_float(w << 16));
            x1 = fmaxf(x1, __uint_as_float(w & 0xffff0000u));
        }
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 16));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 16));
        const float mn0 = fmaxf(m0, x0), mn1 = fmaxf(m1, x1);
        const float sc0 = __expf(m0 - mn0), sc1 = __expf(m1 - mn1);       // exp(-inf) = 0 on the first sub-tile
        m0 = mn0; m1 = mn1;
        // ---- p = exp(k - m) in place (bf16); the row sums use the rounded values the MMA will see
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int i = 0; i < kTcSub / 2; ++i) {
            uint32_t* wp = kw + (2 * i + par) * kWordPitch;
            const uint32_t w = *wp;
            const float e0 = __expf(__uint_as_float(w << 16) - mn0), e1 = __expf(__uint_as_float(w & 0xffff0000u) - mn1);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
            const uint32_t pw = *reinterpret_cast<uint32_t*>(&h2);
            *wp = pw;
            s0 += __uint_as_float(pw << 16);
            s1 += __uint_as_float(pw & 0xffff0000u);
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        l0 = fmaf(l0, sc0, s0);
        l1 = fmaf(l1, sc1, s1);
        __syncwarp();
        // ---- rescale the accumulators: row d = mt*16 + hh*8 + g, its factor lives in lane d>>1, element d&1
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int d = mt * 16 + hh * 8 + g;
                const float f0 = __shfl_sync(0xffffffffu, sc0, d >> 1), f1 = __shfl_sync(0xffffffffu, sc1, d >> 1);
                const float f = (d & 1) ? f1 : f0;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { acc[mt][nt][hh * 2] *= f; acc[mt][nt][hh * 2 + 1] *= f; }
            }
        // ---- ctx += P^T V over 4 k16 steps
        const int j = lane >> 3, r = lane & 7;
#pragma unroll
        for (int ks = 0; ks < kTcSub / 16; ++ks) {
            const int pxb = ks * 16;
            uint32_t af[2][4], bf[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldmatrix_x4_trans(af[mt], &sm[(pxb + (j >> 1) * 8 + r) * kTcPitch + head * 32 + mt * 16 + (j & 1) * 8]);
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldmatrix_x4_trans(bf[np], &sm[(pxb + (j & 1) * 8 + r) * kTcPitch + 128 + head * 32 + (np * 2 + (j >> 1)) * 8]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16_16816(acc[mt][nt], af[mt], bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncthreads();                                 // everyone is done with this buffer before it is refilled
    }
    // ---- this chunk's partial (m[32], l[32], ctx[32][32]) for (b, head)
    float* part = a.partials + (((size_t)b * 4 + head) * a.chunks + chunk) * 1088;
    if (par == 0) {
        part[2 * cp] = m0; part[2 * cp + 1] = m1;
        part[32 + 2 * cp] = l0; part[32 + 2 * cp + 1] = l1;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int d = mt * 16 + hh * 8 + g, e = nt * 8 + 2 * t;
                *reinterpret_cast<float2*>(&part[64 + d * 32 + e]) = make_float2(acc[mt][nt][hh * 2], acc[mt][nt][hh * 2 + 1]);
            }
}

// ---------------------------------------------------------------------------------------------------------
// Fused k-projection + context (C = 64 or 128): reads only x.
//   k = Wk x is computed per 64-pixel sub-tile on the fly (never stored to HBM); v is never formed at all because
//   ctx[d,e] = sum_n p[d,n] (Wv x_n)[e] = (S Wv^T)[d,e]  with  S[d,c] = sum_n p[d,n] x[n,c]   (128 x C per sample).
// The kernel accumulates S with an online softmax over the pixels of its chunk and writes (m, l, S) partials;
// attn_merge_s finishes ctx = (S / l) Wv^T per head.  Replaces the 1x1 kv conv (write-bound: 512 B/pixel) and the
// kv read of attn_ctx_tc: HBM traffic per pixel drops from 128+512+512 bytes to 2C bytes.
constexpr int kXkSub = 64;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}

// Structure (FlashAttention-2 style, everything between two x tiles is warp-local): warp w owns the 16 k-channels
// d = 16w..16w+15 for ALL pixels of a 64-pixel tile:  k^T[16 d][64 px] = Wk[16 d][C] . x^T  (Wk fragments live in
// registers for the whole kernel), so the running max / sum of a row never leave the warp (two shuffles), the fp32
// accumulator fragments of k^T are exactly the bf16 A fragments of P for  S[16 d][C] += P[16 d][64 px] . x[64 px][C],
// and the only CTA-wide synchronisation is the hand-over of the cp.async x-tile ring (one __syncthreads per tile).
constexpr int kXkStages = 3;

template <int C>
__global__ void __launch_bounds__(256, C == 64 ? 2 : 1)
attn_xk_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wk /*[256][C]: k rows, then v rows*/,
               float* __restrict__ partials, int n, int chunks, int chunk_len) {
    pdl_trigger();
    pdl_wait();
    constexpr int kXP = C + 8;                    // bf16 pitch of x / Wk rows (16-byte pad: conflict-free ldmatrix)
    constexpr int kTile = kXkSub * kXP;
    extern __shared__ __align__(16) __nv_bfloat16 xs_all[];
    __nv_bfloat16* wks = xs_all;                                   // [128][kXP]  Wk, later Wv
    __nv_bfloat16* xs = wks + 128 * kXP;                           // [kXkStages][64][kXP]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int n0 = chunk * chunk_len, n1 = min(n, n0 + chunk_len);
    const __nv_bfloat16* xb = x + (size_t)b * n * C;
    const int g = lane >> 2, t = lane & 3, j = lane >> 3, r = lane & 7;

    for (int i = tid; i < 128 * (C / 8); i += 256) {               // Wk -> smem (16-byte pieces)
        const int row = i / (C / 8), c8 = i % (C / 8);
        *reinterpret_cast<uint4*>(&wks[row * kXP + c8 * 8]) = __ldg(reinterpret_cast<const uint4*>(wk + (size_t)row * C + c8 * 8));
    }
    auto fetch = [&](int p0, __nv_bfloat16* dst) {                  // always commits a group (possibly empty)
        if (p0 < n1) {
            for (int i = tid; i < kXkSub * (C / 8); i += 256) {
                const int px = i / (C / 8), c8 = i % (C / 8), nn = p0 + px;
                __nv_bfloat16* d = &dst[px * kXP + c8 * 8];
                if (nn < n1) cp_async16(d, xb + (size_t)nn * C + c8 * 8);
                else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);  // x = 0 beyond the chunk
            }
        }
        cp_async_commit();
    };
    fetch(n0, xs);
    fetch(n0 + kXkSub, xs + kTile);

    float acc[C / 8][4];                                            // S rows 16*warp + {g, g+8}, all C columns
#pragma unroll
    for (int i = 0; i < C / 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // rows g and g+8 (replicated over the 4 t-lanes)
    uint32_t afw[C / 16][4];                                        // this warp's 16 Wk rows as A fragments

    int ib = 0, it = 0;
    for (int p0 = n0; p0 < n1; p0 += kXkSub, ++it) {
        const __nv_bfloat16* xt = xs + ib * kTile;
        cp_async_wait<1>();                                         // tile `it` has landed (tile it+1 may be in flight)
        __syncthreads();                                            // ... for everyone; tile it-1's buffer is free
        {
            int nb = ib + 2; if (nb >= kXkStages) nb -= kXkStages;
            fetch(p0 + 2 * kXkSub, xs + nb * kTile);
        }
        if (it == 0) {
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks)
                ldmatrix_x4(afw[ks], &wks[(warp * 16 + (j & 1) * 8 + r) * kXP + ks * 16 + (j >> 1) * 8]);
        }
        // ---- GEMM1: k^T[16 d][64 px]; n-tile nt = pixels 8nt..8nt+7
        float kc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) kc[i][q] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bf[4];                                     // (b0,b1) of n-tile 2np, then of 2np+1
                ldmatrix_x4(bf, &xt[(np * 16 + (j >> 1) * 8 + r) * kXP + ks * 16 + (j & 1) * 8]);
                mma_bf16_16816(kc[2 * np], afw[ks], bf[0], bf[1]);
                mma_bf16_16816(kc[2 * np + 1], afw[ks], bf[2], bf[3]);
            }
        }
        const int nvalid = n1 - p0;
        if (nvalid < kXkSub) {                                      // pixels beyond the chunk: k = -inf -> p = 0
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = i * 8 + 2 * t;
                if (col >= nvalid) { kc[i][0] = -INFINITY; kc[i][2] = -INFINITY; }
                if (col + 1 >= nvalid) { kc[i][1] = -INFINITY; kc[i][3] = -INFINITY; }
            }
        }
        // ---- online softmax over the pixels (rows of k^T), warp-local
        float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x0 = fmaxf(x0, fmaxf(kc[i][0], kc[i][1]));
            x1 = fmaxf(x1, fmaxf(kc[i][2], kc[i][3]));
        }
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
        const float mn0 = fmaxf(m0, x0), mn1 = fmaxf(m1, x1);
        const float f0 = __expf(m0 - mn0), f1 = __expf(m1 - mn1);   // exp(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1;
        uint32_t pa[4][4];                                          // P as A fragments: k16 slice ks2 = pixels 16ks2..
        float s0 = 0.f, s1 = 0.f;                                   // row sums of the ROUNDED values (what GEMM2 sees)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__expf(kc[i][0] - mn0), __expf(kc[i][1] - mn0));
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__expf(kc[i][2] - mn1), __expf(kc[i][3] - mn1));
            const uint32_t w0 = *reinterpret_cast<uint32_t*>(&h0), w1 = *reinterpret_cast<uint32_t*>(&h1);
            s0 += __uint_as_float(w0 << 16) + __uint_as_float(w0 & 0xffff0000u);
            s1 += __uint_as_float(w1 << 16) + __uint_as_float(w1 & 0xffff0000u);
            pa[i >> 1][(i & 1) * 2 + 0] = w0;                       // a0/a2: row g,   k 2t.. (+8 for the odd n-tile)
            pa[i >> 1][(i & 1) * 2 + 1] = w1;                       // a1/a3: row g+8
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        l0 = fmaf(l0, f0, s0);
        l1 = fmaf(l1, f1, s1);
        // ---- GEMM2: S[16 d][C] = S * f + P[16 d][64 px] . x[64 px][C]
#pragma unroll
        for (int i = 0; i < C / 8; ++i) { acc[i][0] *= f0; acc[i][1] *= f0; acc[i][2] *= f1; acc[i][3] *= f1; }
#pragma unroll
        for (int ks = 0; ks < kXkSub / 16; ++ks) {
#pragma unroll
            for (int np = 0; np < C / 16; ++np) {
                uint32_t bf[4];
                ldmatrix_x4_trans(bf, &xt[(ks * 16 + (j & 1) * 8 + r) * kXP + (np * 2 + (j >> 1)) * 8]);
                mma_bf16_16816(acc[2 * np], pa[ks], bf[0], bf[1]);
                mma_bf16_16816(acc[2 * np + 1], pa[ks], bf[2], bf[3]);
            }
        }
        if (++ib == kXkStages) ib = 0;
    }
    cp_async_wait<0>();
    __syncthreads();                                                // every warp has its Wk fragments: reuse wks for Wv
    // ---- this chunk's partial in the compact format of attn_merge: per head (m[32], l[32], ctx[32][32]) with
    //      ctx[d][e] = sum_c S[d][c] Wv[h*32+e][c]   (linear in S, so it can be applied per chunk).
    // S stays in registers: the fp32 accumulator fragments are repacked as bf16 A fragments (two adjacent n8 tiles
    // = one k16 slice) and multiplied with the bf16 Wv rows staged where Wk was.
    for (int i = tid; i < 128 * (C / 8); i += 256) {
        const int row = i / (C / 8), c8 = i % (C / 8);
        *reinterpret_cast<uint4*>(&wks[row * kXP + c8 * 8]) =
            __ldg(reinterpret_cast<const uint4*>(wk + (size_t)(128 + row) * C + c8 * 8));      // rows [128,256) = v
    }
    __syncthreads();
    const int head = warp >> 1;
    float out[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) out[i][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t af[4];
        {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[2 * ks][0], acc[2 * ks][1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[2 * ks][2], acc[2 * ks][3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
            af[0] = *reinterpret_cast<uint32_t*>(&h0); af[1] = *reinterpret_cast<uint32_t*>(&h1);
            af[2] = *reinterpret_cast<uint32_t*>(&h2); af[3] = *reinterpret_cast<uint32_t*>(&h3);
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            uint32_t bf[4];
            ldmatrix_x4(bf, &wks[(head * 32 + np * 16 + (j >> 1) * 8 + r) * kXP + ks * 16 + (j & 1) * 8]);
            mma_bf16_16816(out[2 * np], af, bf[0], bf[1]);
            mma_bf16_16816(out[2 * np + 1], af, bf[2], bf[3]);
        }
    }
    float* part = partials + (((size_t)b * 4 + head) * chunks + chunk) * 1088;
    const int dl = (warp & 1) * 16 + g;                              // row within the head
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int e = nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(&part[64 + dl * 32 + e]) = make_float2(out[nt][0], out[nt][1]);
        *reinterpret_cast<float2*>(&part[64 + (dl + 8) * 32 + e]) = make_float2(out[nt][2], out[nt][3]);
    }
    if (t == 0) {
        part[dl] = m0;       part[dl + 8] = m1;
        part[32 + dl] = l0;  part[32 + dl + 8] = l1;
    }
}

// Deterministic merge of the per-chunk partials: grid (4 heads, B), 256 threads.
template <bool kStrict>
__global__ void __launch_bounds__(256)
attn_merge_kernel(AttnCtxArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_m[32], s_scale[32], s_part[8][32], s_w[64][32];
    const int tid = threadIdx.x, head = blockIdx.x, b = blockIdx.y;
    const int d = tid & 31, cg = tid >> 5;
    const int nch = a.chunks;                                       // <= 64 by construction (attn_ctx_plan: ~32)
    const float* pbase = a.partials + ((size_t)b * 4 + head) * nch * 1088;
    {   // global max per d: 8 chunk groups in parallel, then a fixed-order combine
        float M = -INFINITY;
        for (int c = cg; c < nch; c += 8) M = fmaxf(M, pbase[(size_t)c * 1088 + d]);
        s_part[cg][d] = M;
    }
    __syncthreads();
    if (tid < 32) {
        float M = s_part[0][tid];
#pragma unroll
        for (int k = 1; k < 8; ++k) M = fmaxf(M, s_part[k][tid]);
        s_m[tid] = M;
    }
    __syncthreads();
    for (int c = cg; c < nch; c += 8) {                             // per-chunk weights exp(m_c - M)
        const float mc = pbase[(size_t)c * 1088 + d];
        s_w[c][d] = kStrict ? expf(mc - s_m[d]) : __expf(mc - s_m[d]);
    }
    __syncthreads();
    if (tid < 32) {
        float l = 0.f;
        for (int c = 0; c < nch; ++c) l += s_w[c][tid] * pbase[(size_t)c * 1088 + 32 + tid];
        s_scale[tid] = 1.0f / l;
        if (a.ml && blockIdx.z == 0) {                                  // kept for the backward pass
            a.ml[((size_t)b * 4 + head) * 64 + tid] = s_m[tid];
            a.ml[((size_t)b * 4 + head) * 64 + 32 + tid] = l;
        }
    }
    __syncthreads();
    {   // blockIdx.z selects a quarter of the 32x32 outputs: one output per thread
        const int i = tid + 256 * (int)blockIdx.z, dd = i >> 5;
        float s = 0.f;
#pragma unroll 8
        for (int c = 0; c < nch; ++c) s += s_w[c][dd] * pbase[(size_t)c * 1088 + 64 + i];
        a.ctxn[((size_t)b * 4 + head) * 1024 + i] = s * s_scale[dd];
    }
}

// grid (C/16 output-row tiles, C/64 input-column tiles, B), 256 threads, fp32:
//   phase 1: P[16 co][128 hd] = Wout[co][h*32 + e] . ctxn[h][d][e]      (8 outputs per thread, K = 32 per head)
//   phase 2: M[16 co][64 ci]  = g * P[16][128] . Wq[128][ci0 + 64]      (4 outputs per thread, K = 128)
// Latency-bound by construction (1-17 MFLOP per sample): every global load of the CTA -- the contexts, its Wout rows and its Wq
// slice, 56 KB -- is in flight before the first use, and 16-row tiles give 4x the CTAs of the first version (64-row tiles, loads
// phase by phase: 20-27 us per launch at any size, 16 % of the batch-1 Euler step in the ncu launch list).
template <typename WT>
__global__ void __launch_bounds__(256)
attn_fold_kernel(const float* __restrict__ ctxn, const float* __restrict__ wout, const float* __restrict__ wq,
                 float g, WT* __restrict__ mb, int C) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float fold_smem[];
    float* cs = fold_smem;                      // ctxn[h*32 + d][e], pitch 33            [128][33]
    float* ws = cs + 128 * 33;                  // Wout rows co0.., pitch 129             [16][129]
    float* P = ws + 16 * 129;                   //                                        [16][129]
    float* qs = P + 16 * 129;                   // Wq[hd][ci0..ci0+63], pitch 64          [128][64]
    const int tid = threadIdx.x, b = blockIdx.z, co0 = blockIdx.x * 16, ci0 = blockIdx.y * 64;
    {
        float c_[16], w_[8];
        float4 q_[8];
#pragma unroll
        for (int k = 0; k < 16; ++k) c_[k] = ctxn[(size_t)b * 4096 + tid + k * 256];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = tid + k * 256; w_[k] = wout[(size_t)(co0 + (i >> 7)) * 128 + (i & 127)]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = tid + k * 256; q_[k] = __ldg(reinterpret_cast<const float4*>(wq + (size_t)(i >> 4) * C + ci0 + (i & 15) * 4)); }
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int i = tid + k * 256; cs[(i >> 5) * 33 + (i & 31)] = c_[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = tid + k * 256; ws[(i >> 7) * 129 + (i & 127)] = w_[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = tid + k * 256; *reinterpret_cast<float4*>(&qs[(i >> 4) * 64 + (i & 15) * 4]) = q_[k]; }
    }
    __syncthreads();
    {
        const int r = tid >> 4, c0 = (tid & 15) * 8, h = c0 >> 5;               // 8 consecutive hd columns lie in one head
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
            const float wv = ws[r * 129 + h * 32 + e];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv, cs[(c0 + j) * 33 + e], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) P[r * 129 + c0 + j] = acc[j];
    }
    __syncthreads();
    {
        const int r = tid >> 4, cq = (tid & 15) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int hd = 0; hd < 128; ++hd) {
            const float4 q4 = *reinterpret_cast<const float4*>(&qs[hd * 64 + cq]);
            const float pv = P[r * 129 + hd];
            acc[0] = fmaf(pv, q4.x, acc[0]); acc[1] = fmaf(pv, q4.y, acc[1]);
            acc[2] = fmaf(pv, q4.z, acc[2]); acc[3] = fmaf(pv, q4.w, acc[3]);
        }
        WT* o = mb + ((size_t)b * C + co0 + r) * C + ci0 + cq;
        Act<WT>::st(o + 0, g * acc[0]); Act<WT>::st(o + 1, g * acc[1]);
        Act<WT>::st(o + 2, g * acc[2]); Act<WT>::st(o + 3, g * acc[3]);
    }
}


// Batched variant of attn_fold_kernel (B >= kFoldWideMinB): the 16-row tiles above re-do phase 1 for every 64-column tile and read
// their operands through 4-way bank conflicts, which is free when one sample is in flight (latency-bound) and shared-memory-bound when
// thousands of CTAs queue up (22 us per wave of CTAs at 16 x C = 256: attn_fold was 2.8 % of the Euler step at chunk 64).  Here one CTA
// owns kR full rows of M: phase 1 runs once per row, the contexts are staged transposed ([h][e][d], conflict-free 16-byte reads), P is
// kept transposed so that phase 2 reads it as warp-uniform broadcasts, the whole Wq is staged once with cp.async (requested together
// with everything else the CTA reads) and each thread owns a kTR x kTC register tile.  Every output is the SAME chain of fmaf as in
// attn_fold_kernel (e = 0..31 from zero, then hd = 0..127 from zero, then g * acc), so the two kernels agree bit for bit and the
// choice by batch size does not break batch invariance (tests/test_gpu_kernels.py::test_attn_fold_variants_agree_bitwise).
constexpr int kFoldWideMinB = 8;
constexpr int kFoldHeadPitch = 32 * 36 + 4;     // [e][36] per head, +4 floats so that the four heads start in different banks

template <int C> struct FoldWide {
    static constexpr int kNch = C >= 256 ? C / 128 : 1;            // column chunks of 128 (one per group of warps)
    static constexpr int kTC = C >= 128 ? 4 : C / 32;              // columns per lane
    static constexpr int kTR = C >= 256 ? 8 : (C >= 128 ? 4 : 2);  // rows per thread: fewer where the matrix is small (more CTAs)
    static constexpr int kRowGroups = 8 / kNch;                    // warps along the rows
    static constexpr int kR = kRowGroups * kTR;                    // rows of M per CTA
    static constexpr int kPPitch = kR + 4;
    static constexpr size_t kSmem = (size_t)(4 * kFoldHeadPitch + kR * 132 + 128 * kPPitch + 128 * C) * sizeof(float);
};

template <typename WT, int C>
__global__ void __launch_bounds__(256)
attn_fold_wide_kernel(const float* __restrict__ ctxn, const float* __restrict__ wout, const float* __restrict__ wq,
                      float g, WT* __restrict__ mb) {
    typedef FoldWide<C> F;
    constexpr int kR = F::kR, kTC = F::kTC, kTR = F::kTR, kPP = F::kPPitch;
    pdl_trigger();
    extern __shared__ __align__(16) float fold_smem[];
    float* cs = fold_smem;                      // ctxn transposed: [h][e][d], pitch 36 per e, kFoldHeadPitch per head
    float* ws = cs + 4 * kFoldHeadPitch;        // Wout rows co0.., [kR][h*33 + e], pitch 132
    float* Pt = ws + kR * 132;                  // P transposed: [hd][kR], pitch kPP
    float* qs = Pt + 128 * kPP;                 // Wq [128][C]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, co0 = blockIdx.x * kR;
    // Wq is a parameter, not a result of the step: its copy may start before the predecessor has finished
#pragma unroll
    for (int k = 0; k < (128 * C / 4) / 256; ++k) cp_async16(qs + (size_t)(tid + k * 256) * 4, wq + (size_t)(tid + k * 256) * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
    pdl_wait();
    {
        constexpr int kWl = (kR * 128 + 255) / 256;
        float c_[16], w_[kWl];
#pragma unroll
        for (int k = 0; k < 16; ++k) c_[k] = ctxn[(size_t)b * 4096 + tid + k * 256];
#pragma unroll
        for (int k = 0; k < kWl; ++k) { const int i = tid + k * 256; w_[k] = i < kR * 128 ? wout[(size_t)(co0 + (i >> 7)) * 128 + (i & 127)] : 0.f; }
#pragma unroll
        for (int k = 0; k < 16; ++k) {                                  // i = h*1024 + d*32 + e
            const int i = tid + k * 256, h = i >> 10, d = (i >> 5) & 31, e = i & 31;
            cs[h * kFoldHeadPitch + e * 36 + d] = c_[k];
        }
#pragma unroll
        for (int k = 0; k < kWl; ++k) {
            const int i = tid + k * 256, col = i & 127;
            if (i < kR * 128) ws[(i >> 7) * 132 + (col >> 5) * 33 + (col & 31)] = w_[k];
        }
    }
    __syncthreads();
    {   // phase 1: thread = (row, kR/2 consecutive hd of one head)
        constexpr int kHd = kR / 2, kG = 128 / kHd;
        static_assert(kHd % 4 == 0 && kHd <= 32, "attn_fold_wide: rows per CTA must be 8..64");
        const int r = tid / kG, hd0 = (tid % kG) * kHd, h = hd0 >> 5, d0 = hd0 & 31;
        float acc[kHd];
#pragma unroll
        for (int j = 0; j < kHd; ++j) acc[j] = 0.f;
        const float* wrow = ws + r * 132 + h * 33;
        const float* crow = cs + h * kFoldHeadPitch + d0;
#pragma unroll 4
        for (int e = 0; e < 32; ++e) {
            const float wv = wrow[e];
#pragma unroll
            for (int q = 0; q < kHd / 4; ++q) {
                const float4 c4 = *reinterpret_cast<const float4*>(crow + e * 36 + 4 * q);
                acc[4 * q + 0] = fmaf(wv, c4.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(wv, c4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(wv, c4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wv, c4.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < kHd; ++j) Pt[(hd0 + j) * kPP + r] = acc[j];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    {   // phase 2: warp = (row group of kTR, column chunk of 32 * kTC), lane = kTC consecutive columns
        const int rg = warp % F::kRowGroups, ch = warp / F::kRowGroups;
        const int col = ch * 128 + lane * kTC;
        float acc[kTR][kTC];
#pragma unroll
        for (int i = 0; i < kTR; ++i)
#pragma unroll
            for (int j = 0; j < kTC; ++j) acc[i][j] = 0.f;
        const float* prow = Pt + rg * kTR;
        const float* qcol = qs + col;
#pragma unroll 8
        for (int hd = 0; hd < 128; ++hd) {
            float qv[kTC], pv[kTR];
            if (kTC == 4) {
                const float4 q4 = *reinterpret_cast<const float4*>(qcol + hd * C);
                qv[0] = q4.x; qv[1 % kTC] = q4.y; qv[2 % kTC] = q4.z; qv[3 % kTC] = q4.w;
            } else {
                const float2 q2 = *reinterpret_cast<const float2*>(qcol + hd * C);
                qv[0] = q2.x; qv[1 % kTC] = q2.y;
            }
            if (kTR >= 4) {
#pragma unroll
                for (int i4 = 0; i4 < kTR / 4; ++i4) {
                    const float4 p4 = *reinterpret_cast<const float4*>(prow + hd * kPP + 4 * i4);
                    pv[(4 * i4 + 0) % kTR] = p4.x; pv[(4 * i4 + 1) % kTR] = p4.y; pv[(4 * i4 + 2) % kTR] = p4.z; pv[(4 * i4 + 3) % kTR] = p4.w;
                }
            } else {
                const float2 p2 = *reinterpret_cast<const float2*>(prow + hd * kPP);
                pv[0] = p2.x; pv[1 % kTR] = p2.y;
            }
#pragma unroll
            for (int i = 0; i < kTR; ++i)
#pragma unroll
                for (int j = 0; j < kTC; ++j) acc[i][j] = fmaf(pv[i], qv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < kTR; ++i) {
            WT* o = mb + ((size_t)b * C + co0 + rg * kTR + i) * C + col;
#pragma unroll
            for (int j = 0; j < kTC; ++j) Act<WT>::st(o + j, g * acc[i][j]);
        }
    }
}

template <typename WT, int C>
int attn_fold_wide_launch(const float* ctxn, const float* wout, const float* wq, float g, WT* mb, int B, cudaStream_t s) {
    typedef FoldWide<C> F;
    static bool set = false;
    auto k = attn_fold_wide_kernel<WT, C>;
    if (!set) { GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F::kSmem)); set = true; }
    GTTS_CHECK_CUDA(launch_pdl(k, dim3(C / F::kR, B), dim3(256), F::kSmem, s, 1, ctxn, wout, wq, g, mb));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename WT>
int attn_fold_wide(const float* ctxn, const float* wout, const float* wq, float g, WT* mb, int B, int C, cudaStream_t s) {
    if (C == 64) return attn_fold_wide_launch<WT, 64>(ctxn, wout, wq, g, mb, B, s);
    if (C == 128) return attn_fold_wide_launch<WT, 128>(ctxn, wout, wq, g, mb, B, s);
    return attn_fold_wide_launch<WT, 256>(ctxn, wout, wq, g, mb, B, s);
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
    if (act == ACT_F32) {
        dim3 grid(a.chunks, 4, a.B);
        if (strict) GTTS_CHECK_CUDA(launch_pdl(attn_ctx_kernel<float, true>, grid, dim3(256), 0, s, 1, a));
        else        GTTS_CHECK_CUDA(launch_pdl(attn_ctx_kernel<float, false>, grid, dim3(256), 0, s, 1, a));
    } else {
        dim3 grid(a.chunks, a.B);
        const int smem = 2 * kTcSub * kTcPitch * 2;
        static bool attr_set = false;
        if (!attr_set) {
            GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_ctx_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr_set = true;
        }
        GTTS_CHECK_CUDA(launch_pdl(attn_ctx_tc_kernel, grid, dim3(128), (size_t)smem, s, 1, a));
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Fused path (bf16, C = 64 / 128): x -> per-chunk (m, l, ctx) partials in the layout attn_merge expects
int attn_xk(const void* x, const void* wkv_bf16, float* partials, int B, int n, int C, int chunks, int chunk_len,
            cudaStream_t s) {
    GTTS_REQUIRE(C == 64 || C == 128, "attn_xk: C must be 64 or 128");
    GTTS_REQUIRE(chunks >= 1 && chunks <= 64 && chunk_len % kXkSub == 0, "attn_xk: bad chunk plan");
    {
        const char* tc = getenv("GTTS_ATTN_TC");
        const char* only = getenv("GTTS_ATTN_TC_ONLY");                 // debug: tcgen05 path for one width / one size only
        const bool sel = !only || atoi(only) == C || atoi(only) == n;
        if ((!tc || atoi(tc) != 0) && sel) {
            int rc = attn_xk_tc(x, wkv_bf16, partials, B, n, C, chunks, chunk_len, s);   // tcgen05 path
            if (getenv("GTTS_ATTN_CHECK")) {                          // debug: scan input and partials on the host
                cudaStreamSynchronize(s);
                std::vector<uint16_t> hx((size_t)B * n * C);
                cudaMemcpy(hx.data(), x, hx.size() * 2, cudaMemcpyDeviceToHost);
                size_t bad_x = 0; float amax = 0.f;
                for (uint16_t v : hx) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); if (!(f == f) || fabsf(f) > 3e38f) ++bad_x; else if (fabsf(f) > amax) amax = fabsf(f); }
                std::vector<float> hp((size_t)B * 4 * chunks * 1088);
                cudaMemcpy(hp.data(), partials, hp.size() * 4, cudaMemcpyDeviceToHost);
                size_t bad_p = 0, first = (size_t)-1;
                for (size_t i = 0; i < hp.size(); ++i) { const float f = hp[i]; const bool is_m = (i % 1088) < 32; if (!(f == f) || (!is_m && fabsf(f) > 3e38f)) { if (first == (size_t)-1) first = i; ++bad_p; } }
                fprintf(stderr, "[attn check] n=%d C=%d chunks=%d len=%d: x nonfinite %zu absmax %.4g; partial nonfinite %zu", n, C, chunks, chunk_len, bad_x, amax, bad_p);
                if (bad_p) { size_t slot = first / 1088; fprintf(stderr, " first at b=%zu head=%zu chunk=%zu idx=%zu", slot / ((size_t)4 * chunks), (slot / chunks) % 4, slot % chunks, first % 1088); }
                fprintf(stderr, "\n");
            }
            return rc;
        }
    }
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
    const __nv_bfloat16* wk = reinterpret_cast<const __nv_bfloat16*>(wkv_bf16);
    dim3 grid(chunks, B);
    if (C == 64) {
        size_t smem = (size_t)(128 * 72 + kXkStages * kXkSub * 72) * 2;
        static bool set64 = false;
        if (!set64) { GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_xk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set64 = true; }
        GTTS_CHECK_CUDA(launch_pdl(attn_xk_kernel<64>, grid, dim3(256), smem, s, 1, xb, wk, partials, n, chunks, chunk_len));
    } else {
        size_t smem = (size_t)(128 * 136 + kXkStages * kXkSub * 136) * 2;
        static bool set128 = false;
        if (!set128) { GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_xk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set128 = true; }
        GTTS_CHECK_CUDA(launch_pdl(attn_xk_kernel<128>, grid, dim3(256), smem, s, 1, xb, wk, partials, n, chunks, chunk_len));
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attn_merge(const AttnCtxArgs& a, bool strict, cudaStream_t s) {
    GTTS_REQUIRE(a.chunks <= 64, "attn_merge: too many chunks");
    dim3 grid(4, a.B, 4);
    if (strict) GTTS_CHECK_CUDA(launch_pdl(attn_merge_kernel<true>, grid, dim3(256), 0, s, 1, a));
    else        GTTS_CHECK_CUDA(launch_pdl(attn_merge_kernel<false>, grid, dim3(256), 0, s, 1, a));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// variant: -1 = by batch size (the product path), 0 = 16-row tiles (latency), 1 = full-row tiles (throughput); bitwise equal results
int attn_fold(ActKind wkind, const float* ctxn, const float* wout, const float* wq, float g, void* mb_out, int B,
              int C, cudaStream_t s, int variant) {
    GTTS_REQUIRE(C % 64 == 0 && C <= 256, "attn_fold: C must be a multiple of 64 and <= 256");
    const bool wide = variant < 0 ? (B >= kFoldWideMinB && (C == 64 || C == 128 || C == 256)) : variant == 1;
    if (wide) {
        GTTS_REQUIRE(C == 64 || C == 128 || C == 256, "attn_fold: the full-row variant needs C = 64, 128 or 256");
        return wkind == ACT_F32 ? attn_fold_wide<float>(ctxn, wout, wq, g, (float*)mb_out, B, C, s)
                                : attn_fold_wide<__nv_bfloat16>(ctxn, wout, wq, g, (__nv_bfloat16*)mb_out, B, C, s);
    }
    dim3 grid(C / 16, C / 64, B);
    const size_t smem = (size_t)(128 * 33 + 2 * 16 * 129 + 128 * 64) * sizeof(float);       // ctx, Wout rows, P, Wq slice
    static bool set_f = false, set_h = false;
    if (wkind == ACT_F32) {
        if (!set_f) { GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fold_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set_f = true; }
        GTTS_CHECK_CUDA(launch_pdl(attn_fold_kernel<float>, grid, dim3(256), smem, s, 1, ctxn, wout, wq, g, (float*)mb_out, C));
    } else {
        if (!set_h) { GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fold_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set_h = true; }
        GTTS_CHECK_CUDA(launch_pdl(attn_fold_kernel<__nv_bfloat16>, grid, dim3(256), smem, s, 1, ctxn, wout, wq, g, (__nv_bfloat16*)mb_out, C));
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
