// HBM-bound point-wise kernels of the decoder: masks, MLPs, first conv, GroupNorm-apply, Euler step,
// weight packing.  Reference lines are cited per kernel (all in /root/reference/model/diffusion.py).
#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

// ------------------------------------------------------------------------------------------------ masks
__global__ void level_masks_kernel(const float* __restrict__ mask, float* m0, float* m1, float* m2, int B, int T) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    int b = i / T, w = i % T;
    float v = mask[i];
    m0[i] = v;
    if ((w & 1) == 0) m1[b * (T / 2) + (w >> 1)] = v;      // mask[..., ::2]      (:196)
    if ((w & 3) == 0) m2[b * (T / 4) + (w >> 2)] = v;      // mask[..., ::2][::2]
}

__global__ void init_xt_kernel(const float* __restrict__ z, const float* __restrict__ mask, float* xt, int B, int H,
                               int W) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)B * H * W;
    if (i >= n) return;
    int w = (int)(i % W);
    int b = (int)(i / ((size_t)H * W));
    xt[i] = __fmul_rn(z[i], mask[(size_t)b * W + w]);       // xt = z * mask       (:257)
}

__global__ void advance_step_kernel(int* step) {
    pdl_trigger();
    pdl_wait(); *step += 1; }

// ------------------------------------------------------------------------------------------------ spk MLP
template <bool kStrict>
__global__ void spk_mlp_kernel(const float* __restrict__ spk, const float* __restrict__ w0t,
                               const float* __restrict__ b0, const float* __restrict__ w2t,
                               const float* __restrict__ b2, float* __restrict__ s_out, int n_feats) {
    __shared__ float sx[64], sh[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < 64) sx[tid] = spk[b * 64 + tid];
    __syncthreads();
    float a = b0[tid];
    for (int i = 0; i < 64; ++i) a += w0t[i * 256 + tid] * sx[i];
    sh[tid] = mish<kStrict>(a);
    __syncthreads();
    if (tid < n_feats) {
        float o = b2[tid];
        for (int j = 0; j < 256; ++j) o += w2t[j * n_feats + tid] * sh[j];
        s_out[b * n_feats + tid] = o;
    }
}

// ------------------------------------------------------------------------------------------------ time MLP
// grid (14, nb), 256 threads; every CTA recomputes the 64-d embedding and writes 128 of the 1792 biases.
template <bool kStrict>
__global__ void temb_kernel(TembWeights w, const float* __restrict__ t, const int* __restrict__ step, int t_is_table,
                            float pe_scale, float* __restrict__ tb) {
    pdl_trigger();
    pdl_wait();
    __shared__ float e[64], h1[256], part[256], h2[64];
    const int tid = threadIdx.x, b = blockIdx.y;
    const float tv = t_is_table ? t[*step] : t[b];
    if (tid < 32) {
        // emb = exp(arange(32).float() * -(log(1e4)/31)) ; arg = (scale * t) * emb          (:118-124)
        const float c = (float)(-9.210340371976184 / 31.0);
        float f = expf((float)tid * c);
        float arg = __fmul_rn(__fmul_rn(pe_scale, tv), f);
        e[tid] = sinf(arg);
        e[tid + 32] = cosf(arg);
    }
    __syncthreads();
    {
        float a = w.b0[tid];
        for (int i = 0; i < 64; ++i) a += w.w0t[i * 256 + tid] * e[i];
        h1[tid] = mish<kStrict>(a);
    }
    __syncthreads();
    {
        const int o = tid & 63, p = tid >> 6;
        float a = 0.f;
        for (int j = p * 64; j < p * 64 + 64; ++j) a += w.w2t[j * 64 + o] * h1[j];
        part[tid] = a;
    }
    __syncthreads();
    if (tid < 64) {
        float a = w.b2[tid] + ((part[tid] + part[64 + tid]) + (part[128 + tid] + part[192 + tid]));
        h2[tid] = mish<kStrict>(a);                       // ResnetBlock.mlp = Mish -> Linear   (:64-65)
    }
    __syncthreads();
    {
        const int j = blockIdx.x * 128 + (tid & 127), p = tid >> 7;
        float a = 0.f;
        for (int i = p * 32; i < p * 32 + 32; ++i) a += w.wbt[i * 1792 + j] * h2[i];
        part[tid] = a;
    }
    __syncthreads();
    if (tid < 128) {
        const int j = blockIdx.x * 128 + tid;
        tb[(size_t)b * 1792 + j] = w.bb[j] + (part[tid] + part[128 + tid]);
    }
}

// ------------------------------------------------------------------------------------------------ first conv
// Block.conv of downs.0.0.block1 on stack([mu, x, (s)]) * mask  (:181-184, 52-57), straight from the fp32 planes.
// Thread = 4 consecutive pixels x 16 output channels (the 4 warps of a CTA cover the 64 channels of 32 pixel quads), so every
// 16-byte weight read from shared memory feeds 16 FMAs.  A CTA walks kFcTiles tiles of 128 pixels of one sample and
// publishes ONE GroupNorm partial (one fence + ticket per CTA).
constexpr int kFcTiles = 8;
constexpr int kFcSpan = kFcTiles * 128 + 2;           // staged pixels per row: the CTA's pixels and one neighbour on either side

template <typename T, int CIN>
__global__ void __launch_bounds__(128)
first_conv_kernel(FirstConvArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float wT[CIN * 9 * 64];
    __shared__ __align__(16) float s_in[2][3][kFcSpan + 6];              // element i + 3 <-> pixel P0 - 1 + i: a quad is 16-byte aligned
    __shared__ __align__(16) float s_mk[kFcSpan + 6];
    __shared__ float sb[64];
    __shared__ float s_tile[16];
    __shared__ double s_red[8 * 16];
    __shared__ int s_flag;
    const int tid = threadIdx.x, b = blockIdx.y, lane = tid & 31, warp = tid >> 5;
    {   // weights: all loads in flight before the first store (a rolled copy loop paid 9 global latencies in a row per CTA:
        // 9.5 % of the kernel's stall samples sat on its STS)
        constexpr int kWn = CIN * 9 * 64, kWIt = (kWn + 127) / 128;
        float wreg[kWIt];
#pragma unroll
        for (int k = 0; k < kWIt; ++k) wreg[k] = tid + k * 128 < kWn ? a.w[tid + k * 128] : 0.f;
#pragma unroll
        for (int k = 0; k < kWIt; ++k)
            if (tid + k * 128 < kWn) wT[tid + k * 128] = wreg[k];
    }
    if (tid < 64) sb[tid] = a.bias[tid];              // visible after the barrier that follows the input staging

    const int H = a.H, W = a.W, HW = H * W;
    // warp = channel block (channels q*16..q*16+15), lane = pixel quad: every weight read is warp-uniform (one shared-memory
    // wavefront).  With q = tid & 3 the 128-bit weight reads of a quarter-warp hit 4 addresses on 2 banks groups: 8 wavefronts
    // per read, and ncu showed the shared-memory pipe 71 % busy -- the kernel's limiter (206 -> 175 us at 16 x 80 x 1720).
    // What is left is still LSU-bound: a warp-uniform LDS.128 costs 4 wavefronts, 308 per tile and warp against 1152 FMA-pipe
    // cycles shared by 4 warps; 4 CTAs per SM (128 registers) changes nothing.
    const int q = tid >> 5, quad = tid & 31;
    float st[4] = {0.f, 0.f, 0.f, 0.f};                   // sums of my two groups, then their sums of squares

    // Stage the CTA's input window once: pixels [P0-1, P0+kFcTiles*128] of the rows above / at / below, both planes already
    // multiplied by the frame mask, plus the mask itself.  (Loading them per tile straight from global left the kernel waiting
    // on the loads with 12 warps per SM: ncu showed 41 % issue-active and 2.4 warps per issue in long-scoreboard stalls.)
    const int P0 = blockIdx.x * (kFcTiles * 128);
    {
        const float* mrow = a.mask + (size_t)b * W;
        const float* mup = a.mu + (size_t)b * HW;
        const float* xp = a.x + (size_t)b * HW;
        // all loads of the window are issued before the first use (9 x 7 independent loads per thread); a rolled loop that
        // loaded, multiplied and stored one element at a time took half of the kernel's time (ncu source view)
        constexpr int kIt = (kFcSpan + 127) / 128;
        float mv[kIt], v[kIt][6];
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
            const int i = tid + k * 128, p = P0 - 1 + i;   // linear pixel of the middle row
            const bool okp = i < kFcSpan && p >= 0 && p < HW;
            mv[k] = okp ? mrow[p % W] : 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int qq = p + (r - 1) * W;
                const bool ok = okp && qq >= 0 && qq < HW;
                v[k][r] = ok ? mup[qq] : 0.f;
                v[k][3 + r] = ok ? xp[qq] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
            const int i = tid + k * 128;
            if (i < kFcSpan) {
                s_mk[i + 3] = mv[k];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    s_in[0][r][i + 3] = v[k][r] * mv[k];
                    s_in[1][r][i + 3] = v[k][3 + r] * mv[k];
                }
            }
        }
    }
    __syncthreads();

    for (int t = 0; t < kFcTiles; ++t) {
        const int local = t * 128 + quad * 4;
        const int p0 = P0 + local;                                        // first pixel of my quad (W % 4 == 0)
        if (p0 >= HW) break;                               // uniform per quad; later tiles are out of range too
        const int h = p0 / W, w0 = p0 - h * W;
        float2 acc[4][8];                                  // packed channel pairs: FFMA2 halves the issue slots
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = make_float2(sb[q * 16 + 2 * c], sb[q * 16 + 2 * c + 1]);
        float mk[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int ww = w0 - 1 + i;
            mk[i] = (ww >= 0 && ww < W) ? s_mk[local + 3 + i] : 0.f;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hh = h + ky - 1;
            if (hh < 0 || hh >= H) continue;
            float in[CIN][6];
            const float sv = CIN == 3 ? a.splane[b * H + hh] : 0.f;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int ww = w0 - 1 + i;
                const bool inb = ww >= 0 && ww < W;
                in[0][i] = inb ? s_in[0][ky][local + 3 + i] : 0.f;
                in[1][i] = inb ? s_in[1][ky][local + 3 + i] : 0.f;
                if (CIN == 3) in[CIN - 1][i] = sv * mk[i];
            }
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4* wr = reinterpret_cast<const float4*>(&wT[((ci * 3 + ky) * 3 + kx) * 64 + q * 16]);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 wv = wr[c4];
                        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 xin = make_float2(in[ci][j + kx], in[ci][j + kx]);
                            acc[j][2 * c4 + 0] = ffma2(xin, w01, acc[j][2 * c4 + 0]);
                            acc[j][2 * c4 + 1] = ffma2(xin, w23, acc[j][2 * c4 + 1]);
                        }
                    }
                }
        }
        T* o = reinterpret_cast<T*>(a.raw) + ((size_t)b * HW + p0) * 64 + q * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v0[8], v1[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                v0[2 * c] = acc[j][c].x;     v0[2 * c + 1] = acc[j][c].y;
                v1[2 * c] = acc[j][4 + c].x; v1[2 * c + 1] = acc[j][4 + c].y;
            }
            if (sizeof(T) == 2) {                          // 16 bf16 = one full 32-byte sector in one store
                uint32_t w[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0[2 * c], v0[2 * c + 1]);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(v1[2 * c], v1[2 * c + 1]);
                    w[c] = *reinterpret_cast<uint32_t*>(&h0);
                    w[4 + c] = *reinterpret_cast<uint32_t*>(&h1);
                }
                st_global_256(o + (size_t)j * 64, w);
            } else {
                Act<T>::store8(o + (size_t)j * 64, v0);
                Act<T>::store8(o + (size_t)j * 64 + 8, v1);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                st[0] += v0[c]; st[2] = fmaf(v0[c], v0[c], st[2]);
                st[1] += v1[c]; st[3] = fmaf(v1[c], v1[c], st[3]);
            }
        }
    }
    // GroupNorm statistics (8 groups of 8 channels) over the unmasked conv output   (:53)
    // the whole warp shares q: fixed butterfly over the 32 quads, lane 0 owns groups 2q, 2q+1
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);
    }
    if (lane == 0) {
        s_tile[2 * warp] = st[0];     s_tile[2 * warp + 1] = st[1];         // sums
        s_tile[8 + 2 * warp] = st[2]; s_tile[8 + 2 * warp + 1] = st[3];     // sums of squares
    }
    __syncthreads();
    GnStatsOut go{a.gn_partials, a.gn_stats, a.gn_counters, (int)gridDim.x, 1.0f / (8.0f * H * W), a.gn_eps};
    gn_stats_publish(go, b, blockIdx.x, tid, 128, s_tile, s_red, &s_flag, [] { __syncthreads(); });
}

// ------------------------------------------------------------------------------------------------ first conv, bf16 mode
// The FFMA kernel above is bound by its shared-memory weight reads (168 us at 16 x 80 x 1720 against a 47 us store floor).  In bf16
// mode the weights are bf16 like those of every other conv, so the K = 18 | 27 contraction goes to the tensor cores as
// mma.sync.m16n8k16 (too thin for a tcgen05 tile pipeline: 2-4 k-steps per pixel group).  The fp32 inputs stay fp32-accurate: every
// staged value is split into bf16 hi + lo halves packed in ONE 32-bit shared-memory word, and K runs over (tap, half) pairs with
// the tap's weight in both B rows, so a single LDS.32 is a whole A register and x = hi + lo enters with ~16 mantissa bits.
// The B columns are permuted (column c of n-tile n <-> channel 16 (c >> 1) + 2n + (c & 1)) so that the accumulators of lane t
// are the 16 CONSECUTIVE channels 16t..16t+15 of its two pixels: one 32-byte store per pixel, and GroupNorm groups 2t, 2t+1.
// The weights live in registers for the whole CTA (48 | 64 per thread).  Grid, window staging and the statistics hand-off are those
// of first_conv_kernel.
__device__ __forceinline__ void fc_mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fc_split(float v) {        // low half = bf16(v), high half = bf16(v - bf16(v))
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

template <int CIN>
__global__ void __launch_bounds__(128)
first_conv_mma_kernel(FirstConvArgs a) {
    constexpr int kTaps = CIN * 9, kSteps = (kTaps * 2 + 15) / 16;      // k = 2 * tap + half
    constexpr int kPitch = kFcSpan + 6;
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) uint32_t s_in[CIN * 3 * kPitch + 16];      // [plane][row][i]: element i + 3 <-> pixel P0 - 1 + i
    __shared__ __align__(16) float sb[64];
    __shared__ float s_tile[16];
    __shared__ float s_part[4][16];
    __shared__ double s_red[8 * 16];
    __shared__ int s_flag;
    const int tid = threadIdx.x, b = blockIdx.y, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int H = a.H, W = a.W, HW = H * W;
    const int P0 = blockIdx.x * (kFcTiles * 128);

    // weights: B fragments of all n-tiles and k-steps; rows (2t, 2t+1) <-> tap 8s + t, rows (2t+8, 2t+9) <-> tap 8s + t + 4
    uint32_t bfrag[kSteps][8][2];
#pragma unroll
    for (int s = 0; s < kSteps; ++s)
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int j = 8 * s + t + 4 * hh, ch = 16 * (g >> 1) + 2 * n + (g & 1);
                const float wv = j < kTaps ? __ldg(a.w + j * 64 + ch) : 0.f;
                const uint32_t wb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(wv));
                bfrag[s][n][hh] = wb | (wb << 16);
            }
    if (tid < 64) sb[tid] = a.bias[tid];
    {   // input window, masked, split and packed (see first_conv_kernel for the staging scheme)
        const float* mrow = a.mask + (size_t)b * W;
        const float* mup = a.mu + (size_t)b * HW;
        const float* xp = a.x + (size_t)b * HW;
        constexpr int kIt = (kFcSpan + 127) / 128;
        float mv[kIt], v[kIt][CIN * 3];
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
            const int i = tid + k * 128, p = P0 - 1 + i;
            const bool okp = i < kFcSpan && p >= 0 && p < HW;
            mv[k] = okp ? mrow[p % W] : 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int qq = p + (r - 1) * W;
                const bool ok = okp && qq >= 0 && qq < HW;
                v[k][r] = ok ? mup[qq] : 0.f;
                v[k][3 + r] = ok ? xp[qq] : 0.f;
                if (CIN == 3) v[k][CIN * 3 - 3 + r] = ok ? a.splane[b * H + qq / W] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
            const int i = tid + k * 128;
            if (i < kFcSpan) {
#pragma unroll
                for (int c = 0; c < CIN * 3; ++c) s_in[c * kPitch + i + 3] = fc_split(v[k][c] * mv[k]);
            }
        }
    }
    // my taps: shared-memory offset of (plane, row, kx) relative to the pixel's slot, and kx for the column-border test
    int toff[kSteps][2], tkx[kSteps][2];
#pragma unroll
    for (int s = 0; s < kSteps; ++s)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = 8 * s + t + 4 * hh;
            const bool okj = j < kTaps;
            const int ci = j / 9, ky = (j - ci * 9) / 3, kx = j - ci * 9 - ky * 3;
            toff[s][hh] = okj ? (ci * 3 + ky) * kPitch + 3 + kx : -1;
            tkx[s][hh] = okj ? kx : 1;
        }
    __syncthreads();

    float st[4] = {0.f, 0.f, 0.f, 0.f};                   // sums of groups 2t, 2t+1, then their sums of squares
    const int wbase = P0 % W;
    for (int grp = warp; grp < kFcTiles * 8; grp += 4) {
        const int local = grp * 16;
        if (P0 + local >= HW) break;                       // H * W is a multiple of 16: groups are whole
        const int pl0 = local + g, pl1 = pl0 + 8;
        const int w0 = (wbase + pl0) % W, w1 = (wbase + pl1) % W;
        float acc[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float2 bb = *reinterpret_cast<const float2*>(&sb[16 * t + 2 * n]);
            acc[n][0] = bb.x; acc[n][1] = bb.y; acc[n][2] = bb.x; acc[n][3] = bb.y;
        }
#pragma unroll
        for (int s = 0; s < kSteps; ++s) {
            uint32_t af[4];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int off = toff[s][hh], kx = tkx[s][hh];
                const bool in0 = off >= 0 && !(kx == 0 && w0 == 0) && !(kx == 2 && w0 == W - 1);
                const bool in1 = off >= 0 && !(kx == 0 && w1 == 0) && !(kx == 2 && w1 == W - 1);
                af[2 * hh + 0] = in0 ? s_in[off + pl0] : 0u;
                af[2 * hh + 1] = in1 ? s_in[off + pl1] : 0u;
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) fc_mma_16816(acc[n], af, bfrag[s][n][0], bfrag[s][n][1]);
        }
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.raw) + ((size_t)b * HW + P0) * 64 + 16 * t;
        uint32_t w0w[8], w1w[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[n][0], acc[n][1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[n][2], acc[n][3]);
            w0w[n] = *reinterpret_cast<uint32_t*>(&h0);
            w1w[n] = *reinterpret_cast<uint32_t*>(&h1);
            const int k = n >> 2;                          // channels 16t + 2n, +1: group 2t + (n >> 2)
            st[k] += (acc[n][0] + acc[n][1]) + (acc[n][2] + acc[n][3]);
            st[2 + k] = fmaf(acc[n][0], acc[n][0], st[2 + k]); st[2 + k] = fmaf(acc[n][1], acc[n][1], st[2 + k]);
            st[2 + k] = fmaf(acc[n][2], acc[n][2], st[2 + k]); st[2 + k] = fmaf(acc[n][3], acc[n][3], st[2 + k]);
        }
        st_global_256(o + (size_t)pl0 * 64, w0w);
        st_global_256(o + (size_t)pl1 * 64, w1w);
    }
    // GroupNorm statistics over the unmasked conv output: fixed butterfly over the 8 pixel lanes, then the 4 warps in order
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);
    }
    if (g == 0) {
        s_part[warp][2 * t] = st[0];     s_part[warp][2 * t + 1] = st[1];
        s_part[warp][8 + 2 * t] = st[2]; s_part[warp][8 + 2 * t + 1] = st[3];
    }
    __syncthreads();
    if (tid < 16) s_tile[tid] = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
    __syncthreads();
    GnStatsOut go{a.gn_partials, a.gn_stats, a.gn_counters, (int)gridDim.x, 1.0f / (8.0f * H * W), a.gn_eps};
    gn_stats_publish(go, b, blockIdx.x, tid, 128, s_tile, s_red, &s_flag, [] { __syncthreads(); });
}

// ------------------------------------------------------------------------------------------------ GN apply
// grid (blocks per sample, B).  Each thread owns one 8-channel vector position (fixed channels, so the affine
// constants stay in registers) and walks kGnIter groups of kGnVec vectors; the loads of group i+1 are issued before
// the arithmetic of group i (software pipeline), so every SM keeps ~48 KB of reads in flight all the time.
constexpr int kGnVec = 4;
constexpr int kGnIter = 4;

template <typename T, bool kStrict, bool kHasRes, bool kHasTb, bool kFirstRes>
__global__ void __launch_bounds__(256, 2)
gn_apply_kernel(GnApplyArgs a) {
    pdl_trigger();
    pdl_wait();
    typedef typename Act<T>::Packed Packed;
    const int C8 = a.C >> 3;
    const int b = blockIdx.y;
    const size_t per_sample = (size_t)a.H * a.W * C8;                  // vectors per sample
    const size_t v00 = (size_t)blockIdx.x * (256 * kGnVec * kGnIter) + threadIdx.x;
    const int c8 = (int)(threadIdx.x % C8);                            // 256 % C8 == 0: same channels for all my vectors
    const int c0 = c8 * 8;
    const int g = (c0 * 8) / a.C;
    const float mean = a.stats[(b * 8 + g) * 2], rstd = a.stats[(b * 8 + g) * 2 + 1];
    float sc[8], sh[8], tb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // GroupNorm (:53) as y = x*scale + shift with scale = rstd*gamma, shift = beta - mean*scale (ATen's form)
        sc[j] = rstd * __ldg(a.gamma + c0 + j);
        sh[j] = __ldg(a.beta + c0 + j) - mean * sc[j];
        tb[j] = kHasTb ? __ldg(a.tbias + (size_t)b * a.tbias_bstride + c0 + j) : 0.f;
    }
    const T* raw = reinterpret_cast<const T*>(a.raw) + (size_t)b * per_sample * 8;
    const T* res = kHasRes ? reinterpret_cast<const T*>(a.residual) + (size_t)b * per_sample * 8 : nullptr;
    T* out = reinterpret_cast<T*>(a.out) + (size_t)b * per_sample * 8;
    const float* mrow = a.mask + (size_t)b * a.W;
    float frw[kFirstRes ? 8 : 1][3], frb[kFirstRes ? 8 : 1];
    if (kFirstRes) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            frb[j] = __ldg(a.fr_b + c0 + j);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) frw[j][ci] = ci < a.fr_cin ? __ldg(a.fr_w + (c0 + j) * a.fr_cin + ci) : 0.f;
        }
    }

    Packed pv[2][kGnVec], pr[2][kGnVec];
    float m[2][kGnVec], fin[2][kGnVec][3];
    auto issue = [&](int stage, int it) {
#pragma unroll
        for (int k = 0; k < kGnVec; ++k) {
            const size_t vi = v00 + (size_t)(it * kGnVec + k) * 256;
            if (vi < per_sample) {
                pv[stage][k] = Act<T>::load_packed(raw + vi * 8);
                if (kHasRes) pr[stage][k] = Act<T>::load_packed(res + vi * 8);
                const size_t pin = vi / C8;                            // pixel within the sample
                m[stage][k] = mrow[(int)(pin % a.W)];
                if (kFirstRes) {
                    const size_t pix = (size_t)b * a.H * a.W + pin;
                    fin[stage][k][0] = a.fr_mu[pix];
                    fin[stage][k][1] = a.fr_x[pix];
                    fin[stage][k][2] = a.fr_cin == 3 ? a.fr_s[pix / a.W] : 0.f;
                }
            }
        }
    };
    issue(0, 0);
#pragma unroll
    for (int it = 0; it < kGnIter; ++it) {
        const int cur = it & 1;
        if (it + 1 < kGnIter) issue(cur ^ 1, it + 1);
#pragma unroll
        for (int k = 0; k < kGnVec; ++k) {
            const size_t vi = v00 + (size_t)(it * kGnVec + k) * 256;
            if (vi >= per_sample) continue;
            float v[8], r[8];
            Act<T>::unpack(pv[cur][k], v);
            if (kHasRes) Act<T>::unpack(pr[cur][k], r);
            const float mk = m[cur][k];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float y = fmaf(v[j], sc[j], sh[j]);
                y = mish<kStrict>(y) * mk;                               // Mish, * mask   (:54,58)
                if (kHasTb) y += tb[j];                                  // h += mlp(t)    (:76)
                if (kHasRes) y += r[j];                                  // + res_conv(x)  (:78)
                if (kFirstRes) {                                         // res_conv(x*mask) of the first block, inline
                    float rr = frb[j];
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) rr = fmaf(frw[j][ci], fin[cur][k][ci] * mk, rr);
                    y += rr;
                }
                v[j] = y * mk;         // every consumer masks its input: store masked (binary masks)
            }
            Act<T>::store8(out + vi * 8, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ GN apply (bf16 fast)
// The pass is issue-bound, not bandwidth-bound (measured: a 190 KB-deep bulk-copy ring fed by 8 math warps was
// slower).  This version spends ~10 issue slots per element: packed f32x2 affine / Mish / mask / bias arithmetic,
// a branch-free Mish (two MUFU per element), bf16 data kept packed in registers until used.
constexpr int kGfVec = 4;
// kIter groups of kGfVec vectors per thread: the per-thread setup (affine constants, 20+ loads) is paid once.  Measured:
// 2 groups help the plain variant (120 -> 109 us at 80x1720x16) and hurt the residual variant (134 -> 158 us).

template <bool kHasRes, bool kHasTb, bool kFirstRes, int kGfIter>
__global__ void __launch_bounds__(256, kFirstRes ? 3 : 4)
gn_apply_fast_kernel(GnApplyArgs a) {
    // first-block variant: the kHasTb flag is reused as "the input has a third (speaker) channel" (the two exclude each other)
    constexpr bool kTb = kHasTb && !kFirstRes, kCin3 = kFirstRes && kHasTb;
    constexpr int kCin = kCin3 ? 3 : 2;
    pdl_trigger();
    pdl_wait();
    typedef __nv_bfloat16 T;
    // All index arithmetic in 32 bits (a sample has < 2^31 vectors) and C/8 is a power of two: the first version spent
    // ~40 % of its instructions in 64-bit division subroutines, on a kernel that is issue-bound.
    const uint32_t C8 = (uint32_t)a.C >> 3, c8shift = (uint32_t)__ffs((int)C8) - 1u, W = (uint32_t)a.W;
    const int b = blockIdx.y;
    const uint32_t per_sample = (uint32_t)a.H * W * C8;
    const uint32_t vbase = blockIdx.x * (256u * kGfVec * kGfIter) + threadIdx.x;
    const int c0 = (int)(threadIdx.x & (C8 - 1u)) * 8;
    const int g = (c0 * 8) / a.C;
    const float mean = a.stats[(b * 8 + g) * 2], rstd = a.stats[(b * 8 + g) * 2 + 1];
    constexpr float kLog2e = 1.4426950408889634f;
    float2 sc[4], sh[4], tb[4];
    const float2 l2e = make_float2(kLog2e, kLog2e);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float s0 = rstd * __ldg(a.gamma + c0 + 2 * q), s1 = rstd * __ldg(a.gamma + c0 + 2 * q + 1);
        // explicit fma: the GroupNorm-apply epilogue of the conv (conv_tc_halo2.cu) must reproduce these bits
        const float h0 = fmaf(-mean, s0, __ldg(a.beta + c0 + 2 * q)), h1 = fmaf(-mean, s1, __ldg(a.beta + c0 + 2 * q + 1));
        sc[q] = make_float2(s0, s1);                   sh[q] = make_float2(h0, h1);
        tb[q] = kTb ? make_float2(__ldg(a.tbias + (size_t)b * a.tbias_bstride + c0 + 2 * q),
                                     __ldg(a.tbias + (size_t)b * a.tbias_bstride + c0 + 2 * q + 1))
                       : make_float2(0.f, 0.f);
    }
    const T* raw = reinterpret_cast<const T*>(a.raw) + (size_t)b * per_sample * 8;
    const T* res = kHasRes ? reinterpret_cast<const T*>(a.residual) + (size_t)b * per_sample * 8 : nullptr;
    T* out = reinterpret_cast<T*>(a.out) + (size_t)b * per_sample * 8;
    const float* mrow = a.mask + (size_t)b * W;

    float2 frw[kFirstRes ? 4 : 1][kCin], frb[kFirstRes ? 4 : 1];
    if (kFirstRes) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            frb[q] = make_float2(__ldg(a.fr_b + c0 + 2 * q), __ldg(a.fr_b + c0 + 2 * q + 1));
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci)
                frw[q][ci] = make_float2(__ldg(a.fr_w + (c0 + 2 * q) * kCin + ci), __ldg(a.fr_w + (c0 + 2 * q + 1) * kCin + ci));
        }
    }
    // pixel column of my vector: one division for the first vector, then += 256/C8 pixels per vector with wrap-around
    const uint32_t pstep = 256u >> c8shift;
    uint32_t hrow_n = (vbase >> c8shift) / W, wcol_n = (vbase >> c8shift) - hrow_n * W;
#pragma unroll 1
    for (int itg = 0; itg < kGfIter; ++itg) {
    const uint32_t v0 = vbase + (uint32_t)itg * (256u * kGfVec);
    uint4 pv[kGfVec], pr[kGfVec];
    float m[kGfVec], fin[kFirstRes ? kGfVec : 1][kCin];
    bool ok[kGfVec];
#pragma unroll
    for (int k = 0; k < kGfVec; ++k) {
        const uint32_t vi = v0 + (uint32_t)k * 256u;
        ok[k] = vi < per_sample;
        if (ok[k]) {
            pv[k] = __ldg(reinterpret_cast<const uint4*>(raw + (size_t)vi * 8));
            if (kHasRes) pr[k] = __ldg(reinterpret_cast<const uint4*>(res + (size_t)vi * 8));
            const uint32_t pin = vi >> c8shift, hrow = hrow_n, wcol = wcol_n;
            m[k] = mrow[wcol];
            if (kFirstRes) {
                const size_t pix = (size_t)b * a.H * W + pin;
                fin[k][0] = a.fr_mu[pix] * m[k];
                fin[k][1] = a.fr_x[pix] * m[k];
                if (kCin3) fin[k][kCin - 1] = a.fr_s[(size_t)b * a.H + hrow] * m[k];
            }
        }
        wcol_n += pstep;
        while (wcol_n >= W) { wcol_n -= W; ++hrow_n; }
    }
#pragma unroll
    for (int k = 0; k < kGfVec; ++k) {
        if (!ok[k]) continue;
        const uint32_t w[4] = {pv[k].x, pv[k].y, pv[k].z, pv[k].w};
        const uint32_t rw[4] = {pr[k].x, pr[k].y, pr[k].z, pr[k].w};
        const float2 m2 = make_float2(m[k], m[k]);
        uint32_t ow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 x = make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u));
            const float2 y = ffma2(x, sc[q], sh[q]);                    // GroupNorm affine          (:53)
            float2 o = mish2_fast(y, fmul2(y, l2e));                    // Mish                      (:54)
            if (kTb) o = fadd2(o, tb[q]);                            // (mish*m + tb)*m == (mish + tb)*m, m in {0,1}
            if (kHasRes) o = fadd2(o, make_float2(__uint_as_float(rw[q] << 16), __uint_as_float(rw[q] & 0xffff0000u)));
            if (kFirstRes) {
                float2 rr = frb[q];
#pragma unroll
                for (int ci = 0; ci < kCin; ++ci) rr = ffma2(frw[q][ci], make_float2(fin[k][ci], fin[k][ci]), rr);
                o = fadd2(o, rr);
            }
            o = fmul2(o, m2);                                           // * mask (stored masked)
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o.x, o.y);
            ow[q] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(out + (size_t)(v0 + (uint32_t)k * 256u) * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    }
}

// ------------------------------------------------------------------------------------------------ Euler step
// final_block GN+Mish+mask -> final_conv(64->1)+bias -> *mask = score; then the sampler update.
// 8 lanes per pixel (one 8-channel vector each), kEuPix pixels per lane group with the loads issued up front.
constexpr int kEuPix = 8;

template <typename T, bool kStrict>
__global__ void __launch_bounds__(256, 3)
euler_kernel(EulerArgs a) {
    pdl_trigger();
    pdl_wait();
    const size_t npix = (size_t)a.B * a.H * a.W;
    const int sub = threadIdx.x & 7;
    const size_t pbase = ((size_t)blockIdx.x * 32 + (threadIdx.x >> 3)) * kEuPix;   // 32 lane groups per CTA
    const int HW = a.H * a.W;
    // ---- everything this thread will need from memory is requested up front
    typename Act<T>::Packed pv[kEuPix];
#pragma unroll
    for (int k = 0; k < kEuPix; ++k) {
        const size_t pix = pbase + k;
        if (pix < npix) pv[k] = Act<T>::load_packed(reinterpret_cast<const T*>(a.raw) + pix * 64 + sub * 8);
    }
    // lane k of the 8-lane group holds xt and mu of pixel k and finishes that pixel
    const size_t mypix = pbase + sub;
    float pre_xt = 0.f, pre_mu = 0.f;
    if (a.update && mypix < npix) { pre_xt = a.xt[mypix]; pre_mu = a.mu[mypix]; }
    float mk[kEuPix], mean[kEuPix], rstd[kEuPix];
#pragma unroll
    for (int q4 = 0; q4 < kEuPix / 4; ++q4) {
        // pbase + 4*q4 is a multiple of 4 and so are W and H*W (T % 4 == 0): each quad of pixels shares sample and row.
        // 32-bit arithmetic, one division pair per quad (the first version did eight 64-bit divisions per thread).
        const uint32_t pp = pbase + 4 * q4 < npix ? (uint32_t)pbase + 4u * q4 : 0u;
        const uint32_t b = pp / (uint32_t)HW, rem = pp - b * (uint32_t)HW, w0 = rem % (uint32_t)a.W;
        const float mean_b = a.stats[(b * 8 + sub) * 2], rstd_b = a.stats[(b * 8 + sub) * 2 + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mk[4 * q4 + k] = a.mask[(size_t)b * a.W + w0 + k];
            mean[4 * q4 + k] = mean_b;
            rstd[4 * q4 + k] = rstd_b;
        }
    }
    float sc[8], sh_[8], wf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = __ldg(a.gamma + sub * 8 + j);
        sh_[j] = __ldg(a.beta + sub * 8 + j);
        wf[j] = __ldg(a.wf + sub * 8 + j);
    }
    const float beta_t = a.update ? a.beta_tab[*a.step] : 0.f;
    const float hh = a.update ? *a.h_ptr : 0.f;
    float noise_v = 0.f;
    if (a.update && a.sde && mypix < npix) {
        const NoiseSlot ns = *a.noise_slot;
        noise_v = ns.base[(size_t)(*a.step) * (size_t)ns.step_stride + mypix];
    }
#pragma unroll
    for (int k = 0; k < kEuPix; ++k) {
        const bool valid = pbase + k < npix;
        const float m = mk[k];
        float part = 0.f;
        if (valid) {
            float v[8];
            Act<T>::unpack(pv[k], v);
            if (kStrict) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float y = (v[j] - mean[k]) * rstd[k] * sc[j] + sh_[j];
                    y = mish<true>(y) * m;                     // final_block(x, mask)              (:212)
                    part = fmaf(y * m, wf[j], part);           // final_conv(x * mask)              (:213)
                }
            } else {
                const float2 r2 = make_float2(rstd[k], rstd[k]), nm2 = make_float2(-mean[k] * rstd[k], -mean[k] * rstd[k]);
                const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
                float2 p2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 xn = ffma2(make_float2(v[2 * q], v[2 * q + 1]), r2, nm2);
                    const float2 y = ffma2(xn, make_float2(sc[2 * q], sc[2 * q + 1]), make_float2(sh_[2 * q], sh_[2 * q + 1]));
                    const float2 o = mish2_fast(y, fmul2(y, l2e));
                    p2 = ffma2(o, make_float2(wf[2 * q], wf[2 * q + 1]), p2);
                }
                part = (p2.x + p2.y) * m;                      // binary mask: (mish*m)*m*w summed == m * sum(mish*w)
            }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (!valid || sub != k) continue;                      // lane k finishes pixel k
        const float mu = pre_mu;
        const size_t pix = pbase + k;
        const float score = __fmul_rn(part + a.bf, m);         // (output * mask)                   (:216)
        if (a.score_out) a.score_out[pix] = score;
        if (!a.update) continue;
        const float xt = pre_xt;
        float nx;
        if (!a.sde) {
            // dxt = 0.5*(mu - xt - est); dxt = dxt*noise_t*h; xt = (xt - dxt)*mask          (:265-267)
            float d = __fsub_rn(__fsub_rn(mu, xt), score);
            d = __fmul_rn(0.5f, d);
            d = __fmul_rn(__fmul_rn(d, beta_t), hh);
            nx = __fmul_rn(__fsub_rn(xt, d), m);
        } else {
            // upstream Grad-TTS stochastic branch (huawei-noah, deleted in this fork), same operation order:
            //   dxt_det = (0.5*(mu - xt) - est) * noise_t * h ; dxt_stoc = z * sqrt(noise_t * h) ; xt = (xt - (dxt_det + dxt_stoc)) * mask
            // i.e. the noise is SUBTRACTED.  BASELINE.json's north-star line writes "+ sqrt(beta*h)*z": the same process with -z
            // (negation is exact), so a caller who wants that literal form passes -noise.
            float d = __fsub_rn(__fmul_rn(0.5f, __fsub_rn(mu, xt)), score);
            d = __fmul_rn(__fmul_rn(d, beta_t), hh);
            d = __fadd_rn(d, __fmul_rn(noise_v, sqrtf(__fmul_rn(beta_t, hh))));
            nx = __fmul_rn(__fsub_rn(xt, d), m);
        }
        a.xt[pix] = nx;
    }
}

// ------------------------------------------------------------------------------------------------ packing
template <typename WT>
__global__ void pack_conv_kernel(const float* __restrict__ w, WT* __restrict__ out, int Cout, int Cin, int kh,
                                 int kw) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)Cout * Cin * kh * kw;
    if (i >= n) return;
    int ci = (int)(i % Cin);
    size_t r = i / Cin;
    int co = (int)(r % Cout);
    int tap = (int)(r / Cout);
    int ky = tap / kw, kx = tap % kw;
    Act<WT>::st(out + i, w[(((size_t)co * Cin + ci) * kh + ky) * kw + kx]);
}

// ConvTranspose2d(C, C, 4, 2, 1): weight (Cin, Cout, 4, 4); out[2j+py] taps: py=0 -> (dy=0,k=1),(dy=-1,k=3);
// py=1 -> (dy=0,k=2),(dy=+1,k=0); same along x.  Row = ((phase*4 + tap)*C + co), phase = py*2+px, tap = ty*2+tx.
template <typename WT>
__global__ void pack_convT_kernel(const float* __restrict__ w, WT* __restrict__ out, int C) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)16 * C * C;
    if (i >= n) return;
    int ci = (int)(i % C);
    size_t r = i / C;
    int co = (int)(r % C);
    int pt = (int)(r / C);
    int phase = pt >> 2, tap = pt & 3;
    int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
    const int kk[2][2] = {{1, 3}, {2, 0}};
    int ky = kk[py][ty], kx = kk[px][tx];
    Act<WT>::st(out + i, w[(((size_t)ci * C + co) * 4 + ky) * 4 + kx]);
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    dst[(size_t)c * rows + r] = src[i];
}

// ------------------------------------------------------------------------------------------------ fp32 -> 3 x bf16
// x = hi + mid + lo exactly: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid) (the differences are exact in fp32)
__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(h));
    m = __float2bfloat16_rn(r1);
    const float r2 = __fsub_rn(r1, __bfloat162float(m));
    l = __float2bfloat16_rn(r2);
}

// in [P][C] fp32 -> out [P][3C] bf16 = [hi | mid | lo]; thread = 8 channels of one pixel
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t nvec, int C8) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nvec) return;
    const size_t pix = i / C8;
    const int c0 = (int)(i % C8) * 8, C = C8 * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + i * 8)), b = __ldg(reinterpret_cast<const float4*>(in + i * 8) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3(v[j], h[j], m[j], l[j]);
    __nv_bfloat16* o = out + pix * (size_t)(3 * C) + c0;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(m);
    *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(l);
}

// packed fp32 weights [rows][K] -> bf16 [rows][6K] = [wl | wm | wm | wh | wh | wh] (the term order of conv_tc_plan_create: smallest first)
__global__ void split_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, size_t n, int K) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const size_t row = i / K;
    const int c = (int)(i % K);
    __nv_bfloat16 h, m, l;
    split3(w[i], h, m, l);
    __nv_bfloat16* o = out + row * (size_t)(6 * K) + c;
    o[0] = l; o[K] = m; o[2 * K] = m; o[3 * K] = h; o[4 * K] = h; o[5 * K] = h;
}

inline unsigned int nblk(size_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

}  // namespace

int split_f32_planes(const float* in, void* out, size_t npix, int C, cudaStream_t s) {
    GTTS_REQUIRE(C % 8 == 0, "split_f32_planes: C must be a multiple of 8");
    const size_t nvec = npix * (size_t)(C / 8);
    GTTS_CHECK_CUDA(launch_pdl(split_planes_kernel, dim3(nblk(nvec, 256)), dim3(256), 0, s, 1, in, (__nv_bfloat16*)out, nvec, C / 8));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int split_pack_weights(const float* w, void* out, size_t rows, int K, cudaStream_t s) {
    const size_t n = rows * (size_t)K;
    split_weights_kernel<<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)out, n, K);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int build_level_masks(const float* mask, float* m0, float* m1, float* m2, int B, int T, cudaStream_t s) {
    level_masks_kernel<<<nblk((size_t)B * T, 256), 256, 0, s>>>(mask, m0, m1, m2, B, T);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int init_xt(const float* z, const float* mask, float* xt, int B, int H, int W, cudaStream_t s) {
    init_xt_kernel<<<nblk((size_t)B * H * W, 256), 256, 0, s>>>(z, mask, xt, B, H, W);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int advance_step(int* step, cudaStream_t s) {
    GTTS_CHECK_CUDA(launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, s, 1, step));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int spk_mlp(const float* spk, const float* w0t, const float* b0, const float* w2t, const float* b2, float* s_out,
            int B, int n_feats, bool strict, cudaStream_t s) {
    GTTS_REQUIRE(n_feats <= 256, "spk_mlp: n_feats too large");
    if (strict) spk_mlp_kernel<true><<<B, 256, 0, s>>>(spk, w0t, b0, w2t, b2, s_out, n_feats);
    else        spk_mlp_kernel<false><<<B, 256, 0, s>>>(spk, w0t, b0, w2t, b2, s_out, n_feats);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int temb_bias(const TembWeights& w, const float* t, const int* step, int t_is_table, float pe_scale, float* tb,
              int nb, bool strict, cudaStream_t s) {
    dim3 grid(14, nb);
    if (strict) GTTS_CHECK_CUDA(launch_pdl(temb_kernel<true>, grid, dim3(256), 0, s, 1, w, t, step, t_is_table, pe_scale, tb));
    else        GTTS_CHECK_CUDA(launch_pdl(temb_kernel<false>, grid, dim3(256), 0, s, 1, w, t, step, t_is_table, pe_scale, tb));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

size_t first_conv_partials_slots(int H, int W) { return (size_t)((H * W + 127) / 128 + kFcTiles - 1) / kFcTiles; }

int first_conv(ActKind act, const FirstConvArgs& a, cudaStream_t s) {
    GTTS_REQUIRE(a.cin == 2 || a.cin == 3, "first_conv: cin must be 2 or 3");
    GTTS_REQUIRE(a.W % 4 == 0, "first_conv: W must be a multiple of 4");
    dim3 grid((unsigned int)first_conv_partials_slots(a.H, a.W), a.B);
    if (act == ACT_F32) {
        if (a.cin == 2) GTTS_CHECK_CUDA(launch_pdl(first_conv_kernel<float, 2>, grid, dim3(128), 0, s, 1, a));
        else            GTTS_CHECK_CUDA(launch_pdl(first_conv_kernel<float, 3>, grid, dim3(128), 0, s, 1, a));
    } else {
        if (a.use_mma && (a.H * a.W) % 16 == 0) {
            if (a.cin == 2) GTTS_CHECK_CUDA(launch_pdl(first_conv_mma_kernel<2>, grid, dim3(128), 0, s, 1, a));
            else            GTTS_CHECK_CUDA(launch_pdl(first_conv_mma_kernel<3>, grid, dim3(128), 0, s, 1, a));
        } else {
            if (a.cin == 2) GTTS_CHECK_CUDA(launch_pdl(first_conv_kernel<__nv_bfloat16, 2>, grid, dim3(128), 0, s, 1, a));
            else            GTTS_CHECK_CUDA(launch_pdl(first_conv_kernel<__nv_bfloat16, 3>, grid, dim3(128), 0, s, 1, a));
        }
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

namespace {
template <typename T, bool kStrict>
int gn_apply_dispatch(const GnApplyArgs& a, cudaStream_t s) {
    const size_t per_sample = (size_t)a.H * a.W * (a.C / 8);
    dim3 grid((unsigned int)((per_sample + 256 * kGnVec * kGnIter - 1) / (256 * kGnVec * kGnIter)), a.B);
    const bool res = a.residual != nullptr, tb = a.tbias != nullptr, fr = a.fr_w != nullptr;
    if (fr)             GTTS_CHECK_CUDA(launch_pdl(gn_apply_kernel<T, kStrict, false, false, true>, grid, dim3(256), 0, s, 1, a));
    else if (res && tb) GTTS_CHECK_CUDA(launch_pdl(gn_apply_kernel<T, kStrict, true, true, false>, grid, dim3(256), 0, s, 1, a));
    else if (res)       GTTS_CHECK_CUDA(launch_pdl(gn_apply_kernel<T, kStrict, true, false, false>, grid, dim3(256), 0, s, 1, a));
    else if (tb)        GTTS_CHECK_CUDA(launch_pdl(gn_apply_kernel<T, kStrict, false, true, false>, grid, dim3(256), 0, s, 1, a));
    else                GTTS_CHECK_CUDA(launch_pdl(gn_apply_kernel<T, kStrict, false, false, false>, grid, dim3(256), 0, s, 1, a));
    return 0;
}
}  // namespace

namespace {
template <bool kHasRes, bool kHasTb, bool kFirstRes>
int gn_apply_fast_launch(const GnApplyArgs& a, cudaStream_t s) {
    constexpr int kGfIter = kFirstRes ? 4 : (kHasRes ? 1 : 2);     // first-block variant: 56 scalar constant loads per thread to amortise
    const size_t per_sample = (size_t)a.H * a.W * (a.C / 8);
    dim3 grid((unsigned int)((per_sample + 256 * kGfVec * kGfIter - 1) / (256 * kGfVec * kGfIter)), a.B);
    GTTS_CHECK_CUDA(launch_pdl(gn_apply_fast_kernel<kHasRes, kHasTb, kFirstRes, kGfIter>, grid, dim3(256), 0, s, 1, a));
    return 0;
}
}  // namespace

int gn_apply(ActKind act, const GnApplyArgs& a, bool strict, cudaStream_t s) {
    GTTS_REQUIRE(a.C % 64 == 0 && a.C <= 256, "gn_apply: C must be 64, 128 or 256");
    GTTS_REQUIRE(!(a.fr_w && (a.residual || a.tbias)), "gn_apply: first-block residual excludes the others");
    if (act == ACT_BF16 && !strict) {
        const bool res = a.residual != nullptr, tb = a.tbias != nullptr, fr = a.fr_w != nullptr;
        int rc;
        if (fr)             rc = a.fr_cin == 3 ? gn_apply_fast_launch<false, true, true>(a, s) : gn_apply_fast_launch<false, false, true>(a, s);
        else if (res && tb) rc = gn_apply_fast_launch<true, true, false>(a, s);
        else if (res)       rc = gn_apply_fast_launch<true, false, false>(a, s);
        else if (tb)        rc = gn_apply_fast_launch<false, true, false>(a, s);
        else                rc = gn_apply_fast_launch<false, false, false>(a, s);
        if (rc) return rc;
        GTTS_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (act == ACT_F32) {
        if (strict) gn_apply_dispatch<float, true>(a, s);
        else        gn_apply_dispatch<float, false>(a, s);
    } else {
        if (strict) gn_apply_dispatch<__nv_bfloat16, true>(a, s);
        else        gn_apply_dispatch<__nv_bfloat16, false>(a, s);
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int euler_step(ActKind act, const EulerArgs& a, bool strict, cudaStream_t s) {
    const size_t npix = (size_t)a.B * a.H * a.W;
    unsigned int g = nblk(npix, 32 * kEuPix);
    if (act == ACT_F32) {
        if (strict) GTTS_CHECK_CUDA(launch_pdl(euler_kernel<float, true>, dim3(g), dim3(256), 0, s, 1, a));
        else        GTTS_CHECK_CUDA(launch_pdl(euler_kernel<float, false>, dim3(g), dim3(256), 0, s, 1, a));
    } else {
        if (strict) GTTS_CHECK_CUDA(launch_pdl(euler_kernel<__nv_bfloat16, true>, dim3(g), dim3(256), 0, s, 1, a));
        else        GTTS_CHECK_CUDA(launch_pdl(euler_kernel<__nv_bfloat16, false>, dim3(g), dim3(256), 0, s, 1, a));
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int pack_conv_weight(ActKind wkind, const float* w, void* packed, int Cout, int Cin, int kh, int kw,
                     cudaStream_t s) {
    size_t n = (size_t)Cout * Cin * kh * kw;
    if (wkind == ACT_F32) pack_conv_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)packed, Cout, Cin, kh, kw);
    else pack_conv_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)packed, Cout, Cin, kh, kw);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int pack_convT_weight(ActKind wkind, const float* w, void* packed, int C, cudaStream_t s) {
    size_t n = (size_t)16 * C * C;
    if (wkind == ACT_F32) pack_convT_kernel<float><<<nblk(n, 256), 256, 0, s>>>(w, (float*)packed, C);
    else pack_convT_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, s>>>(w, (__nv_bfloat16*)packed, C);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int transpose_2d(const float* src, float* dst, int rows, int cols, cudaStream_t s) {
    transpose_kernel<<<nblk((size_t)rows * cols, 256), 256, 0, s>>>(src, dst, rows, cols);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
