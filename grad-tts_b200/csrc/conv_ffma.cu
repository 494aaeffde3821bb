// CUDA-core implicit-GEMM convolution (fp32 accumulate with FFMA).
//
// This is the arithmetic of the `fp32` precision mode (true fp32 multiply-add, the mode that has to meet
// max-abs 1e-3 after 10 Euler steps) and the on-device cross-check for the tcgen05 kernel in conv_tc.cu.
// It covers every conv variant of GradLogPEstimator2d through the ConvGeom descriptor (ops.h).
// Tile: 64 output-grid pixels x 64 output channels per CTA, 256 threads, 4x4 outputs per thread.
#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

constexpr int kTP = 64, kTC = 64, kTK = 32, kPitch = 68;

template <typename T, typename WT>
__global__ void __launch_bounds__(256)
conv_ffma_kernel(ConvGeom g, const T* __restrict__ src0, const T* __restrict__ src1, const WT* __restrict__ weight,
                 ConvEpilogue e, int tiles_per_phase) {
    __shared__ __align__(16) float As[kTK * kPitch];
    __shared__ __align__(16) float Ws[kTK * kPitch];
    __shared__ float s_thr[256 * 2];
    __shared__ float s_tile[16];
    __shared__ double s_red[16 * 16];
    __shared__ int s_flag;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x, nb = blockIdx.y;
    const int b = blockIdx.z / g.nphase, ph = blockIdx.z % g.nphase;
    const int co0 = nb * kTC;
    const int npix = g.Hg * g.Wg;
    const int cin_tot = g.Cin0 + g.Cin1;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loader roles
    const int lp = tid >> 2, lv = (tid & 3) * 8;          // A: pixel lp, channels lv..lv+7 ; W: cout lp, k lv..
    const int lpix = tile * kTP + lp;
    const bool lvalid = lpix < npix;
    const int lj = lvalid ? lpix / g.Wg : 0, li = lvalid ? lpix % g.Wg : 0;

    // K loop over (tap, 32-channel slice), software-pipelined: the global loads of slice it+1 are issued before the FMAs of slice
    // it, so their latency hides behind the arithmetic (one CTA per SM at small sizes has nothing else to hide it with)
    const int nslice = cin_tot / kTK, n_it = g.ntaps * nslice;
    float av[8], wv[8];
    auto fetch = [&](int it) {
        const int tap = it / nslice, c0 = (it - tap * nslice) * kTK;
        const int ih = lj * g.stride + g.dy[ph][tap], iw = li * g.stride + g.dx[ph][tap];
        const bool inb = lvalid && ih >= 0 && ih < g.Hin && iw >= 0 && iw < g.Win;
        if (inb) {
            const size_t ipix = ((size_t)b * g.Hin + ih) * g.Win + iw;
            if (c0 < g.Cin0) Act<T>::load8(src0 + ipix * g.Cin0 + c0 + lv, av);
            else             Act<T>::load8(src1 + ipix * g.Cin1 + (c0 - g.Cin0) + lv, av);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) av[j] = 0.f;
        }
        const size_t wrow = (size_t)g.wrow[ph][tap] + (size_t)b * g.w_batch_rows + co0 + lp;
        Act<WT>::load8(weight + wrow * cin_tot + c0 + lv, wv);
    };
    if (n_it > 0) fetch(0);
    for (int it = 0; it < n_it; ++it) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            As[(lv + j) * kPitch + lp] = av[j];
            Ws[(lv + j) * kPitch + lp] = wv[j];
        }
        __syncthreads();
        if (it + 1 < n_it) fetch(it + 1);
#pragma unroll 8
        for (int k = 0; k < kTK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k * kPitch + ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k * kPitch + tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
    }

    // ---- epilogue
    float ssum = 0.f, ssq = 0.f;
    T* out = reinterpret_cast<T*>(e.out);
    const T* res = reinterpret_cast<const T*>(e.residual);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pidx = tile * kTP + ty * 4 + i;
        if (pidx >= npix) continue;
        const int j = pidx / g.Wg, ii = pidx % g.Wg;
        const int oh = j * g.out_step + g.oy[ph], ow = ii * g.out_step + g.ox[ph];
        const size_t opix = ((size_t)b * g.Hout + oh) * g.Wout + ow;
        const float m = e.mask ? e.mask[(size_t)b * g.Wout + ow] : 1.0f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int co = co0 + tx * 4 + jj;
            float v = acc[i][jj];
            if (e.bias) v += e.bias[co];
            ssum += v;
            ssq += v * v;
            if (res) v += Act<T>::ld(res + opix * g.Cout + co);
            if (e.mask) v *= m;
            const float lv = v > 0.f ? v : v * e.act_slope;      // 1-D conv stacks (ConvEpilogue::act_out / out2)
            Act<T>::st(out + opix * g.Cout + co, e.act_out ? lv : v);
            if (e.out2) Act<T>::st(reinterpret_cast<T*>(e.out2) + opix * g.Cout + co, lv);
        }
    }
    if (e.gn_partials == nullptr) return;

    // GroupNorm partial statistics: fixed-order reduction over the CTA, one 16-vector per (tile, nb) slot
    s_thr[tid * 2] = ssum;
    s_thr[tid * 2 + 1] = ssq;
    __syncthreads();
    if (tid < 16) {
        const int grp = tid & 7, which = tid >> 3;
        const int gsz = g.Cout / 8;                       // channels per group: 8, 16 or 32
        const int c_lo = grp * gsz - co0;                 // group's first channel relative to this CTA
        float s = 0.f;
        if (c_lo >= 0 && c_lo < kTC) {
            const int tx_lo = c_lo / 4, tx_n = gsz / 4;
            for (int yy = 0; yy < 16; ++yy)
                for (int xx = tx_lo; xx < tx_lo + tx_n; ++xx) s += s_thr[(yy * 16 + xx) * 2 + which];
        }
        s_tile[tid] = s;
    }
    __syncthreads();
    const int slots = tiles_per_phase * (g.Cout / kTC);
    GnStatsOut go{e.gn_partials, e.gn_stats, e.gn_counters, slots,
                  1.0f / ((float)(g.Cout / 8) * (float)g.Hout * (float)g.Wout), e.gn_eps};
    gn_stats_publish(go, b, tile * (g.Cout / kTC) + nb, tid, 256, s_tile, s_red, &s_flag, [] { __syncthreads(); });
}

}  // namespace

size_t conv_ffma_partials_slots(const ConvGeom& g) {
    return (size_t)((g.Hg * g.Wg + kTP - 1) / kTP) * (g.Cout / kTC);
}

int conv_ffma(ActKind act, const ConvGeom& g, const void* src0, const void* src1, const void* weight,
              const ConvEpilogue& e, cudaStream_t stream) {
    GTTS_REQUIRE(g.Cout % kTC == 0, "conv_ffma: Cout must be a multiple of 64");
    GTTS_REQUIRE(g.Cin0 % kTK == 0 && g.Cin1 % kTK == 0, "conv_ffma: Cin must be a multiple of 32");
    GTTS_REQUIRE(e.gn_partials == nullptr || (g.nphase == 1 && e.residual == nullptr && e.mask == nullptr),
                 "conv_ffma: GN statistics only on plain convs");
    const int tiles = (g.Hg * g.Wg + kTP - 1) / kTP;
    dim3 grid(tiles, g.Cout / kTC, g.B * g.nphase);
    if (act == ACT_F32)
        conv_ffma_kernel<float, float><<<grid, 256, 0, stream>>>(g, (const float*)src0, (const float*)src1,
                                                                 (const float*)weight, e, tiles);
    else
        conv_ffma_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
            g, (const __nv_bfloat16*)src0, (const __nv_bfloat16*)src1, (const __nv_bfloat16*)weight, e, tiles);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
