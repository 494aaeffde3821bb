// 3x3 stride-1 halo convolution on CTA PAIRS (tcgen05 cta_group::2).
//
// Measured on B200 (profiles/r01_conv_experiments.md): the single-CTA halo kernel is bound by the SHARED-MEMORY port --
// every M128 x N x K16 MMA reads 4 KB of A and N*32 bytes of B, 6-12 KB per 50-130 cycles, and the TMA fill of the next
// tiles shares that port.  A CTA pair issues M = 256 MMAs from one thread of the leader CTA: each CTA supplies its own
// 128 pixel rows of A and only HALF of the weight tile (N/2 rows), so per CTA the weight reads, the weight fill and the
// weight footprint all halve -- 128->128 and 256->64 weights become resident, 256->256 / 512->128 stream half as much.
//
// Pairing: cluster (2,1,1); tile = blockIdx.x + it * gridDim.x as everywhere else, so the two CTAs of a pair work on
// tiles 2q and 2q+1 of the same walk step; both run identical iteration counts (tc_num_iters, mc != 0) and a tile index
// past the end is a dummy whose TMA box is out of bounds (zero fill, bytes still counted) and whose rows are never stored.
// Barriers: `full*` live in the LEADER (both CTAs' TMA bytes are counted there), `empty*` / `tfull` are local to each CTA
// and signalled by the leader's multicast commits, `tempty` lives in the leader and collects both CTAs' epilogue warps.
#include <cstring>

#include "conv_tc_common.cuh"

namespace gtts {

using namespace tc;

namespace {

template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc2() {      // as make_idesc, M = 256 across the pair
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
}


// ================================================================================================ fused GroupNorm-apply epilogue
// kApply: the conv does not write its raw output.  Every accumulator tile stays in TMEM while the epilogue warps read it once for
// the GroupNorm sums (pass 1); when the LAST tile of a sample has been summed anywhere in the grid (per-sample grid barrier through
// two global counters), every CTA reduces the sample's per-CTA partial rows itself, in the same fixed order, to (mean, rstd); the
// epilogue warps then read their tiles of that sample a second time (pass 2), apply GroupNorm + Mish (+ time bias) (+ residual) and
// the frame mask, and store the finished activation.  What this removes per Block: one bf16 write of the raw tensor and the whole
// gn_apply pass (a read and a write), i.e. the pass that measured 23 % of the Euler step at 84-99 % of HBM bandwidth.
//
// Tiles of one sample are consecutive in every CTA's walk ("run"); a run must fit the TMEM accumulator ring (host check:
// ceil(tiles per sample / grid) <= acc_bufs<N>()), otherwise the MMA warp would wait for a buffer that pass 2 can only free after
// the barrier.  All CTAs must be co-resident for the barrier to complete: the grid is clamped to what cudaOccupancyMaxActiveClusters
// reports (conv_tc_halo2_max_grid).  Shared memory: [misc+4096, +8192) holds the per-channel affine tables of two runs in flight
// (scale = rstd*gamma, shift = beta - mean*scale), the kApplyExtra bytes behind misc hold the time-bias rows and four mbarriers
// (aff_full / aff_empty per table slot).  Pass 2 first rounds (accumulator + bias) to bf16, i.e. to the value the unfused plan
// stores, and then uses gn_apply's own formulas: the two plans are bitwise identical (the decoder picks the plan by batch size, and
// a sample's result must not depend on that).
constexpr int kApplyExtra = 3072;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct ApplyShared {
    float* s_sc;            // [2][N]
    float* s_sh;            // [2][N]
    float* s_tb;            // [2][N]
    float* s_mr;            // [16] mean[8], rstd[8] of the run being finalised (statistics warp only)
    uint64_t* aff_full;     // [2] statistics warp -> epilogue warps: tables of slot s are valid
    uint64_t* aff_empty;    // [2] 16 epilogue warps -> statistics warp: slot s may be overwritten
};
template <int N>
__device__ __forceinline__ ApplyShared apply_shared(uint8_t* misc) {
    ApplyShared a;
    a.s_sc = reinterpret_cast<float*>(misc + 4096);
    a.s_sh = a.s_sc + 2 * N;
    a.s_tb = reinterpret_cast<float*>(misc + kMiscBytes);
    a.aff_full = reinterpret_cast<uint64_t*>(misc + kMiscBytes + 2048);
    a.aff_empty = a.aff_full + 2;
    a.s_mr = reinterpret_cast<float*>(misc + kMiscBytes + 2048 + 64);
    return a;
}

// iterations [it, it_end) of this CTA that belong to the sample of iteration `it` (tile = blockIdx.x + it * gridDim.x); a dummy
// tile (index past the end, odd tail of a CTA pair) is a run of its own without a sample (b = -1)
__device__ __forceinline__ int apply_run_end(const TcParams& p, int it, int n_it, int tps, int* b_out) {
    const int G = (int)gridDim.x;
    const int tile0 = (int)blockIdx.x + it * G;
    if (tile0 >= p.num_tiles) { *b_out = -1; return it + 1; }
    const int b = tile0 / tps;
    const int last_tile = (b + 1) * tps - 1;
    int it_end = it + (last_tile - tile0) / G + 1;
    if (it_end > n_it) it_end = n_it;
    *b_out = b;
    return it_end;
}

// Statistics warp (warp 3) of the kApply variant.
template <int N>
__device__ __forceinline__ void tc_stats_apply_loop(const TcParams& p, const TcShared& sh, int lane) {
    constexpr int kGsz = N / 8;
    const ConvEpilogue& e = p.e;
    const ApplyShared ap = apply_shared<N>(sh.misc);
    const int G = (int)gridDim.x, bx = (int)blockIdx.x;
    const int tps = p.tiles_h * p.tiles_w;
    const int n_it = tc_num_iters(p);
    const int g = lane & 7, which = (lane >> 3) & 1, half = g >> 2, idx = which * 4 + (g & 3);
    const int nrows = tps < G ? tps : G;                             // CTAs that hold tiles of any one sample
    const double inv_count = 1.0 / ((double)kGsz * (double)p.Hout * (double)p.Wout);
    unsigned int* cnt = e.gn_counters + 32;                          // [2b] arrivals, [2b+1] departures of sample b
    int run = 0;
    for (int it = 0; it < n_it;) {
        int b;
        const int it_end = apply_run_end(p, it, n_it, tps, &b);
        float acc = 0.f;
        for (int j = it; j < it_end; ++j) {
            const int slot = j % kStatSlots;
            mbar_wait(&sh.sfull[slot], (uint32_t)(j / kStatSlots) & 1u);
            if (lane < 16 && b >= 0) {
                const float* r = sh.s_ring + (slot * 8 + half * 4) * 8 + idx;
                acc += (r[0] + r[8]) + (r[16] + r[24]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.sempty[slot]);
        }
        it = it_end;
        if (b < 0) continue;
        // ---- my partial row of sample b, then the per-sample grid barrier
        if (lane < 16) e.gn_partials[((size_t)b * G + bx) * 16 + lane] = acc;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            atomicAdd(&cnt[2 * b], 1u);
            while (ld_acquire_gpu_u32(&cnt[2 * b]) < (unsigned int)nrows) { }
        }
        __syncwarp();
        // ---- every CTA reduces the rows of the contributing CTAs itself, all in the same order (deterministic; double)
        const int start = tps < G ? (int)(((long long)b * tps) % G) : 0;
        {
            const int q = lane & 3, r0 = lane >> 2;
            const float4* pp = reinterpret_cast<const float4*>(e.gn_partials + (size_t)b * G * 16) + q;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            // rows in CTA order 0..G-1, exactly the order (and lane pattern) of tc_teardown's finalisation: a CTA that holds no tile of
            // this sample contributes zero there and is skipped here, so both kernels add the same numbers in the same order
            for (int base = 0; base < G; base += 64) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = base + r0 + 8 * i;
                    int rel = c - start; if (rel < 0) rel += G;
                    v[i] = (c < G && rel < nrows) ? __ldcg(pp + (size_t)c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a0 += (double)v[i].x; a1 += (double)v[i].y; a2 += (double)v[i].z; a3 += (double)v[i].w;
                }
            }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, off);
                a1 += __shfl_xor_sync(0xffffffffu, a1, off);
                a2 += __shfl_xor_sync(0xffffffffu, a2, off);
                a3 += __shfl_xor_sync(0xffffffffu, a3, off);
            }
            // lanes 0,1 hold the sums of groups 4q..4q+3; lanes 2,3 the matching sums of squares
            const double q0 = __shfl_sync(0xffffffffu, a0, (lane + 2) & 31), q1 = __shfl_sync(0xffffffffu, a1, (lane + 2) & 31);
            const double q2 = __shfl_sync(0xffffffffu, a2, (lane + 2) & 31), q3 = __shfl_sync(0xffffffffu, a3, (lane + 2) & 31);
            if (lane < 2) {
                const double su[4] = {a0, a1, a2, a3}, sq[4] = {q0, q1, q2, q3};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double mean = su[j] * inv_count;
                    double var = sq[j] * inv_count - mean * mean;
                    if (var < 0.0) var = 0.0;
                    const float mf = (float)mean, rf = (float)rsqrt(var + (double)e.gn_eps);
                    const int gg = lane * 4 + j;
                    ap.s_mr[gg] = mf; ap.s_mr[8 + gg] = rf;
                    if (bx == start) {                               // one CTA publishes the statistics for later consumers / tests
                        e.gn_stats[((size_t)b * 8 + gg) * 2 + 0] = mf;
                        e.gn_stats[((size_t)b * 8 + gg) * 2 + 1] = rf;
                    }
                }
            }
            __syncwarp();
        }
        // ---- affine tables of this run (slot = run parity); the slot is free once all 16 epilogue warps are done with run - 2
        const int sl = run & 1;
        mbar_wait(&ap.aff_empty[sl], ((uint32_t)(run >> 1) & 1u) ^ 1u);
        for (int c = lane; c < N; c += 32) {
            const int gg = c / kGsz;
            const float sc = ap.s_mr[8 + gg] * __ldg(e.ap_gamma + c);
            ap.s_sc[sl * N + c] = sc;
            ap.s_sh[sl * N + c] = fmaf(-ap.s_mr[gg], sc, __ldg(e.ap_beta + c));       // the same expression as gn_apply (pointwise.cu)
            ap.s_tb[sl * N + c] = e.ap_tbias ? __ldg(e.ap_tbias + (size_t)b * e.ap_tb_bstride + c) : 0.f;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&ap.aff_full[sl]);                           // release: orders this warp's table writes (after __syncwarp)
            // departure: whoever leaves last resets both counters for the next launch (everyone has finished polling by then)
            if (atomicAdd(&cnt[2 * b + 1], 1u) == (unsigned int)(nrows - 1)) { cnt[2 * b] = 0u; cnt[2 * b + 1] = 0u; }
        }
        ++run;
    }
}

// Epilogue warps (warps 4..19) of the kApply variant: two groups of 8 warps take alternate tile iterations, as in tc_epilogue_loop.
template <int N, bool kTb, bool kRes>
__device__ __forceinline__ void tc_epilogue_apply_loop(const TcParams& p, const TcShared& sh, uint32_t tmem_base, int warp, int lane) {
    constexpr int kColsPerWarp = N / 2;
    constexpr int kGsz = N / 8;
    constexpr int kBufs = acc_bufs<N>();
    const ApplyShared ap = apply_shared<N>(sh.misc);
    const int ew16 = warp - 4, grp = ew16 >> 3, ew = ew16 & 7, wq = ew & 3, half = ew >> 2;
    const int row = wq * 32 + lane;
    const ConvEpilogue& e = p.e;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(e.residual);
    const int cbase = half * kColsPerWarp;
    int hl = row / p.bw, wl = row - hl * p.bw;
    if (p.halo_t) { wl = row >> 3; hl = row & 7; }
    const bool row_in_tile = hl < p.bh;
    const float* s_bias = sh.s_bias;
    const int G = (int)gridDim.x, tps = p.tiles_h * p.tiles_w;
    const int n_it = tc_num_iters(p);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)cbase;
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    TileWalk tw;
    tw.init(p, (int)blockIdx.x, G);
    int run = 0;
    for (int it = 0; it < n_it;) {
        int b;
        const int it_end = apply_run_end(p, it, n_it, tps, &b);
        // ------------------------------------------------------------ pass 1: GroupNorm sums of my tiles of this run
        {
            TileWalk w1 = tw;
            for (int j = it; j < it_end; ++j, w1.advance(G)) {
                if ((j & 1) != grp) continue;
                const int buf = j % kBufs;
                const int jj = w1.th * p.bh + hl, ii = w1.tw * p.bw + wl;
                const bool valid = row_in_tile && (jj < p.Hg) && (ii < p.Wg) && (b >= 0);
                mbar_wait(&sh.tfull[buf], (uint32_t)(j / kBufs) & 1u);
                tc_fence_after();
                float2 ssum[4], ssq[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) { ssum[g] = make_float2(0.f, 0.f); ssq[g] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(lane_addr + (uint32_t)(buf * N + c0), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[cbase + c0 + q4 * 4]);
                        float2 f0 = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 0]), __uint_as_float(r[q4 * 4 + 1])), make_float2(b4.x, b4.y));
                        float2 f1 = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 2]), __uint_as_float(r[q4 * 4 + 3])), make_float2(b4.z, b4.w));
                        // select, not multiply: rows outside the image hold whatever the zero-filled box produced (bias only) or garbage
                        if (!valid) { f0 = make_float2(0.f, 0.f); f1 = make_float2(0.f, 0.f); }
                        const int g0 = (c0 + q4 * 4) / kGsz, g1 = (c0 + q4 * 4 + 2) / kGsz;
                        ssum[g0] = fadd2(ssum[g0], f0); ssq[g0] = ffma2(f0, f0, ssq[g0]);
                        ssum[g1] = fadd2(ssum[g1], f1); ssq[g1] = ffma2(f1, f1, ssq[g1]);
                    }
                }
                if (b < 0) {                                         // dummy tile: nothing to apply later, hand the buffer back now
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&sh.tempty[buf]), 0u));
                }
                float st[8];
#pragma unroll
                for (int g = 0; g < 4; ++g) { st[g] = ssum[g].x + ssum[g].y; st[4 + g] = ssq[g].x + ssq[g].y; }
                const float t = warp_reduce8(st, lane);
                const int slot = j % kStatSlots;
                mbar_wait(&sh.sempty[slot], ((uint32_t)(j / kStatSlots) & 1u) ^ 1u);
                if ((lane & 3) == 0) sh.s_ring[(slot * 8 + ew) * 8 + (lane >> 2)] = t;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.sfull[slot]);
            }
        }
        // ------------------------------------------------------------ pass 2: normalise + Mish (+ bias) (+ residual), mask, store
        if (b >= 0) {
            const int sl = run & 1;
            mbar_wait(&ap.aff_full[sl], (uint32_t)(run >> 1) & 1u);
            const float* t_sc = ap.s_sc + sl * N + cbase;
            const float* t_sh = ap.s_sh + sl * N + cbase;
            const float* t_tb = ap.s_tb + sl * N + cbase;
            TileWalk w2 = tw;
            for (int j = it; j < it_end; ++j, w2.advance(G)) {
                if ((j & 1) != grp) continue;
                const int buf = j % kBufs;
                const int jj = w2.th * p.bh + hl, ii = w2.tw * p.bw + wl;
                const bool valid = row_in_tile && (jj < p.Hg) && (ii < p.Wg);
                const size_t opix = valid ? ((size_t)b * p.Hout + jj) * p.Wout + ii : 0;
                const float m = valid ? e.mask[(size_t)b * p.Wout + ii] : 0.f;
                const float2 m2 = make_float2(m, m);
#pragma unroll
                for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(lane_addr + (uint32_t)(buf * N + c0), r);
                    tmem_ld_wait();
                    if (c0 + 32 >= kColsPerWarp) {                   // last TMEM read of this buffer: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&sh.tempty[buf]), 0u));
                    }
                    if (!valid) continue;
                    float2 f[16];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float4 sc4 = *reinterpret_cast<const float4*>(&t_sc[c0 + q4 * 4]);
                        const float4 sh4 = *reinterpret_cast<const float4*>(&t_sh[c0 + q4 * 4]);
                        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[cbase + c0 + q4 * 4]);
                        // conv + bias rounded to bf16 exactly as the unfused plan stores it, so both plans give the same bits
                        // (a sample must not depend on which plan its batch size selects)
                        const float2 a0 = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 0]), __uint_as_float(r[q4 * 4 + 1])), make_float2(b4.x, b4.y));
                        const float2 a1 = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 2]), __uint_as_float(r[q4 * 4 + 3])), make_float2(b4.z, b4.w));
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a0.x, a0.y), h1 = __floats2bfloat162_rn(a1.x, a1.y);
                        const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
                        const float2 y0 = ffma2(make_float2(__uint_as_float(u0 << 16), __uint_as_float(u0 & 0xffff0000u)),
                                                make_float2(sc4.x, sc4.y), make_float2(sh4.x, sh4.y));
                        const float2 y1 = ffma2(make_float2(__uint_as_float(u1 << 16), __uint_as_float(u1 & 0xffff0000u)),
                                                make_float2(sc4.z, sc4.w), make_float2(sh4.z, sh4.w));
                        f[q4 * 2 + 0] = mish2_fast(y0, fmul2(y0, l2e));
                        f[q4 * 2 + 1] = mish2_fast(y1, fmul2(y1, l2e));
                        if (kTb) {
                            const float4 tb4 = *reinterpret_cast<const float4*>(&t_tb[c0 + q4 * 4]);
                            f[q4 * 2 + 0] = fadd2(f[q4 * 2 + 0], make_float2(tb4.x, tb4.y));
                            f[q4 * 2 + 1] = fadd2(f[q4 * 2 + 1], make_float2(tb4.z, tb4.w));
                        }
                    }
                    if (kRes) {
                        const __nv_bfloat16* rp = res + opix * N + cbase + c0;
#pragma unroll
                        for (int v8 = 0; v8 < 2; ++v8) {
                            uint32_t w[8];
                            ld_global_nc_256(rp + v8 * 16, w);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                f[v8 * 8 + k] = fadd2(f[v8 * 8 + k], make_float2(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u)));
                        }
                    }
                    __nv_bfloat16* op = out + opix * N + cbase + c0;
#pragma unroll
                    for (int v8 = 0; v8 < 2; ++v8) {
                        uint32_t w[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float2 o = fmul2(f[v8 * 8 + k], m2);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o.x, o.y);
                            w[k] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        st_global_256(op + v8 * 16, w);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ap.aff_empty[sl]);
            ++run;
        }
        for (int j = it; j < it_end; ++j) tw.advance(G);
        it = it_end;
    }
}


// ================================================================================================ asynchronous apply warps
// kApply = 3 / 4: like 1 / 2 the GroupNorm-apply happens inside the conv kernel, but NOT out of TMEM.  The ordinary epilogue writes
// the raw bf16 tile to global memory (where it stays in the 126 MB L2 for the few microseconds that matter) and feeds the
// statistics ring; the statistics warp publishes the CTA's partial row when its run of a sample ends and bumps the sample's arrival
// counter WITHOUT waiting; eight extra warps (kApplyWarps) follow behind: for every run of this CTA they wait until the sample's
// counter shows that every contributing CTA has arrived (by then all raw tiles of the sample are visible: every CTA's epilogue
// stores precede its statistics warp's __threadfence + atomicAdd), reduce the partial rows in the fixed order, and turn the CTA's
// own tiles of that sample into the finished activation with gn_apply's formulas.  The MMA / epilogue pipeline never waits for a
// sample to complete, so TMEM capacity does not limit anything (any T, any batch); what the separate gn_apply pass cost -- its
// HBM time -- is spent while the tensor pipe is busy with the next tiles.  Bitwise identical to the unfused plan.
constexpr int kApplyWarps = 8;

template <int N>
__device__ __forceinline__ void tc_stats_async_loop(const TcParams& p, const TcShared& sh, int lane) {
    const ConvEpilogue& e = p.e;
    const int G = (int)gridDim.x, bx = (int)blockIdx.x;
    const int tps = p.tiles_h * p.tiles_w;
    const int n_it = tc_num_iters(p);
    const int g = lane & 7, which = (lane >> 3) & 1, half = g >> 2, idx = which * 4 + (g & 3);
    unsigned int* cnt = e.gn_counters + 32;
    for (int it = 0; it < n_it;) {
        int b;
        const int it_end = apply_run_end(p, it, n_it, tps, &b);
        float acc = 0.f;
        for (int j = it; j < it_end; ++j) {
            const int slot = j % kStatSlots;
            mbar_wait(&sh.sfull[slot], (uint32_t)(j / kStatSlots) & 1u);
            if (lane < 16 && b >= 0) {
                const float* r = sh.s_ring + (slot * 8 + half * 4) * 8 + idx;
                acc += (r[0] + r[8]) + (r[16] + r[24]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.sempty[slot]);
        }
        it = it_end;
        if (b < 0) continue;
        if (lane < 16) e.gn_partials[((size_t)b * G + bx) * 16 + lane] = acc;
        // gpu-scope fence, cumulative: it also covers the epilogue warps' raw-tile stores that this warp observed through the ring
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(&cnt[2 * b], 1u);
    }
}

// The eight apply warps.  Warp 0 is the FINALISER: for every run of this CTA it waits for the sample's arrival counter, reduces the
// partial rows and builds the per-channel tables (scale, shift, time bias) in one of two table slots -- it works ahead of the
// seven STREAMING warps, which turn the CTA's tiles of that sample into finished activations with kInFlight rows of loads in
// flight per thread (the first version did both jobs with all eight warps in lock-step, two loads in flight: the kernel was bound
// by them, 380 us against 109 us of convolution at 128->128 @h40).
constexpr int kStreamWarps = kApplyWarps - 1;
constexpr int kStreamThreads = kStreamWarps * 32;

struct AsyncShared {
    float* tab;             // [2 slots][3][N]: scale, shift, time bias
    float* s_mr;            // [16]
    uint64_t* tab_full;     // [2] finaliser -> streaming warps
    uint64_t* tab_empty;    // [2] streaming warps (kStreamWarps arrivals) -> finaliser
};
template <int N>
__device__ __forceinline__ AsyncShared async_shared(uint8_t* misc) {
    AsyncShared a;
    a.tab = reinterpret_cast<float*>(misc + kMiscBytes);             // 2 * 3 * N * 4 <= 6144 bytes
    a.s_mr = reinterpret_cast<float*>(misc + 4096);
    a.tab_full = reinterpret_cast<uint64_t*>(misc + 4096 + 64);
    a.tab_empty = a.tab_full + 2;
    return a;
}
constexpr int kAsyncExtra = 6144;

template <int N>
__device__ __forceinline__ void tc_apply_finaliser(const TcParams& p, const TcShared& sh, int lane) {
    constexpr int kGsz = N / 8;
    const ConvEpilogue& e = p.e;
    const AsyncShared as = async_shared<N>(sh.misc);
    const int G = (int)gridDim.x, bx = (int)blockIdx.x;
    const int tps = p.tiles_h * p.tiles_w;
    const int n_it = tc_num_iters(p);
    const int nrows = tps < G ? tps : G;
    const double inv_count = 1.0 / ((double)kGsz * (double)p.Hout * (double)p.Wout);
    unsigned int* cnt = e.gn_counters + 32;
    int run = 0;
    for (int it = 0; it < n_it;) {
        int b;
        const int it_end = apply_run_end(p, it, n_it, tps, &b);
        it = it_end;
        if (b < 0) continue;
        if (lane == 0) { while (ld_acquire_gpu_u32(&cnt[2 * b]) < (unsigned int)nrows) { __nanosleep(32); } }
        __syncwarp();
        const int start = tps < G ? (int)(((long long)b * tps) % G) : 0;
        const int q = lane & 3, r0 = lane >> 2;
        const float4* pp = reinterpret_cast<const float4*>(e.gn_partials + (size_t)b * G * 16) + q;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int base = 0; base < G; base += 64) {                   // CTA order 0..G-1, the order of tc_teardown's finalisation
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = base + r0 + 8 * i;
                int rel = c - start; if (rel < 0) rel += G;
                v[i] = (c < G && rel < nrows) ? __ldcg(pp + (size_t)c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a0 += (double)v[i].x; a1 += (double)v[i].y; a2 += (double)v[i].z; a3 += (double)v[i].w;
            }
        }
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, off);
            a1 += __shfl_xor_sync(0xffffffffu, a1, off);
            a2 += __shfl_xor_sync(0xffffffffu, a2, off);
            a3 += __shfl_xor_sync(0xffffffffu, a3, off);
        }
        const double q0 = __shfl_sync(0xffffffffu, a0, (lane + 2) & 31), q1 = __shfl_sync(0xffffffffu, a1, (lane + 2) & 31);
        const double q2 = __shfl_sync(0xffffffffu, a2, (lane + 2) & 31), q3 = __shfl_sync(0xffffffffu, a3, (lane + 2) & 31);
        if (lane < 2) {
            const double su[4] = {a0, a1, a2, a3}, sq[4] = {q0, q1, q2, q3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double mean = su[j] * inv_count;
                double var = sq[j] * inv_count - mean * mean;
                if (var < 0.0) var = 0.0;
                const float mf = (float)mean, rf = (float)rsqrt(var + (double)e.gn_eps);
                const int gg = lane * 4 + j;
                as.s_mr[gg] = mf; as.s_mr[8 + gg] = rf;
                if (bx == start) {
                    e.gn_stats[((size_t)b * 8 + gg) * 2 + 0] = mf;
                    e.gn_stats[((size_t)b * 8 + gg) * 2 + 1] = rf;
                }
            }
        }
        __syncwarp();
        const int sl = run & 1;
        mbar_wait(&as.tab_empty[sl], ((uint32_t)(run >> 1) & 1u) ^ 1u);          // the streaming warps are done with run - 2
        float* t_sc = as.tab + (size_t)sl * 3 * N;
        for (int c = lane; c < N; c += 32) {
            const int gg = c / kGsz;
            const float sc = as.s_mr[8 + gg] * __ldg(e.ap_gamma + c);
            t_sc[c] = sc;
            t_sc[N + c] = fmaf(-as.s_mr[gg], sc, __ldg(e.ap_beta + c));          // gn_apply's expression (pointwise.cu)
            t_sc[2 * N + c] = e.ap_tbias ? __ldg(e.ap_tbias + (size_t)b * e.ap_tb_bstride + c) : 0.f;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&as.tab_full[sl]);
            // departure: the last reader resets the counters for the next launch
            if (atomicAdd(&cnt[2 * b + 1], 1u) == (unsigned int)(nrows - 1)) { cnt[2 * b] = 0u; cnt[2 * b + 1] = 0u; }
        }
        ++run;
    }
}

// (row, 16-byte vector) item `item` of tile w: element offset of its 8 channels, first channel, frame index; false outside the image
__device__ __forceinline__ bool stream_item(const TcParams& p, const TileWalk& w, int b, int item, int C8, int N, size_t* off, int* c0, int* ii_out) {
    const int row = item / C8, vec = item - row * C8;
    int hl, wl;
    if (p.halo_t) { wl = row >> 3; hl = row & 7; }
    else { hl = row >> 3; wl = row & 7; }                            // halo tiles are 16 rows x 8 pixels (or 8 x 16 transposed)
    const int jj = w.th * p.bh + hl, ii = w.tw * p.bw + wl;
    *c0 = vec * 8; *ii_out = ii;
    *off = (((size_t)b * p.Hout + jj) * p.Wout + ii) * N + vec * 8;
    return row < 128 && hl < p.bh && jj < p.Hg && ii < p.Wg;
}

template <int N, bool kTb, bool kRes>
__device__ __forceinline__ void tc_apply_stream(const TcParams& p, const TcShared& sh, int tid_s) {
    constexpr int C8 = N / 8;                                        // 16-byte vectors per pixel
    constexpr int kInFlight = kRes ? 5 : 8;                          // items (16-byte loads, x2 with a residual) in flight per thread
    const ConvEpilogue& e = p.e;
    const AsyncShared as = async_shared<N>(sh.misc);
    const int lane = tid_s & 31;
    const int G = (int)gridDim.x, bx = (int)blockIdx.x;
    const int tps = p.tiles_h * p.tiles_w;
    const int n_it = tc_num_iters(p);
    // (row, vector) items of a 128-pixel tile, strided over the streaming threads; a thread's channel vector changes from item to
    // item when kStreamThreads is not a multiple of C8, so the per-channel constants are read from the shared-memory tables
    constexpr int kItems = 128 * C8;
    const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(e.out);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(e.residual);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.ap_out);
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    TileWalk tw;
    tw.init(p, bx, G);
    int run = 0;
    for (int it = 0; it < n_it;) {
        int b;
        const int it_end = apply_run_end(p, it, n_it, tps, &b);
        if (b >= 0) {
            const int sl = run & 1;
            mbar_wait(&as.tab_full[sl], (uint32_t)(run >> 1) & 1u);
            const float* t_sc = as.tab + (size_t)sl * 3 * N;
            TileWalk w2 = tw;
            for (int j = it; j < it_end && !(p.dbg & 256); ++j, w2.advance(G)) {    // dbg 256: finaliser only, no streaming
                for (int i0 = tid_s; i0 < kItems; i0 += kStreamThreads * kInFlight) {
                    // loads of kInFlight items first; only the data stays in registers, coordinates are recomputed afterwards
                    uint4 v[kInFlight], rv[kRes ? kInFlight : 1];
                    float mk[kInFlight];
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        size_t off; int c0, ii;
                        mk[u] = -1.f;
                        if (stream_item(p, w2, b, i0 + u * kStreamThreads, C8, N, &off, &c0, &ii)) {
                            v[u] = __ldcg(reinterpret_cast<const uint4*>(raw + off));
                            if (kRes) rv[u] = __ldg(reinterpret_cast<const uint4*>(res + off));
                            mk[u] = e.mask[(size_t)b * p.Wout + ii];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) {
                        if (mk[u] < 0.f) continue;
                        size_t off; int c0, ii;
                        stream_item(p, w2, b, i0 + u * kStreamThreads, C8, N, &off, &c0, &ii);
                        const float2 m2 = make_float2(mk[u], mk[u]);
                        const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                        const uint4 rr = rv[kRes ? u : 0];
                        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 sc2 = *reinterpret_cast<const float2*>(&t_sc[c0 + 2 * q]);
                            const float2 sh2 = *reinterpret_cast<const float2*>(&t_sc[N + c0 + 2 * q]);
                            const float2 x = make_float2(__uint_as_float(wv[q] << 16), __uint_as_float(wv[q] & 0xffff0000u));
                            const float2 y = ffma2(x, sc2, sh2);
                            float2 o = mish2_fast(y, fmul2(y, l2e));
                            if (kTb) o = fadd2(o, *reinterpret_cast<const float2*>(&t_sc[2 * N + c0 + 2 * q]));
                            if (kRes) o = fadd2(o, make_float2(__uint_as_float(rw[q] << 16), __uint_as_float(rw[q] & 0xffff0000u)));
                            o = fmul2(o, m2);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o.x, o.y);
                            ow[q] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        *reinterpret_cast<uint4*>(out + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&as.tab_empty[sl]);
            ++run;
        }
        for (int j = it; j < it_end; ++j) tw.advance(G);
        it = it_end;
    }
}

constexpr int kFuseWarps = 4;                       // transform warps of the fused-input variant (warps 20..23, one per SMSP)

// kFuse: the A tiles are the RAW output of the previous conv; four extra warps apply (Mish(GroupNorm(raw)) + tbias) * mask
// in place in shared memory between the TMA arrival (local barrier rawfull) and the MMA (the leader's `full`, which then
// counts warp arrivals of both CTAs instead of TMA bytes).  Saves the separate gn_apply pass (one write + one read of the
// activation through HBM) and is bitwise identical to it -- but it is OFF by default (decoder option fuse_gn): measured
// 435 us vs 134 (conv) + 103 (gn_apply) us for 64->64 at 80x1720x16.  GN+Mish at 5 TB/s already needs a whole SM's issue and
// MUFU capacity (32 warps); four latency-bound transform warps, which also redo the 41 % halo overlap, cannot supply it.
// kConvT: transposed 4x4 stride-2 conv (Upsample, model/diffusion.py:21-27) as four output phases of 2x2 taps each: all 16
// (phase, tap) operands are shifted views of ONE halo box of the input tile, so the input crosses L2->SMEM once instead of 16
// times (the per-tap kernel is fill-bound there: 372 TFLOP/s).  Every phase has its own accumulator and epilogue iteration
// (TcParams::ph_inner = 4, phases innermost); the epilogue multiplies by the mask (kMask).
// kApply: 0 = off; 1 / 2 = GroupNorm-apply epilogue out of TMEM with time bias (block1 of a ResnetBlock) / with residual (block2);
// 3 / 4 = the same two flavours done by asynchronous apply warps from the raw tile in global memory
template <int N, bool kStats, bool kFuse, bool kMask = false, bool kConvT = false, int kApply = 0, bool kOutF32 = false>
__global__ void __launch_bounds__(kThreads + (kFuse ? kFuseWarps * 32 : 0) + (kApply >= 3 ? kApplyWarps * 32 : 0), 1)
conv_tc_halo2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                     const __grid_constant__ CUtensorMap mapWh, const TcParams p) {
    constexpr int kBHalf = N * 64;                                  // bytes of this CTA's half of one weight tile
    constexpr uint32_t kIdesc = make_idesc2<N>();
    constexpr int kBufs = acc_bufs<N>();

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    constexpr int pw = 10;                                           // halo box: 18 x 10 pixels (or transposed)
    const int a_stage = p.a_bytes;
    uint8_t* smem_b = smem + (size_t)p.stages * a_stage;
    const TcShared sh = tc_shared(smem_b + (size_t)p.b_slots * kBHalf);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    uint64_t* rawfull = reinterpret_cast<uint64_t*>(sh.misc + 3904);   // [8] local "raw tile landed" barriers (kFuse)
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapWh);
    }
    // ---- prologue (the shared tc_prologue assumes single-CTA barrier counts)
    pdl_trigger();
    const int nslot = p.b_slots, resident = p.b_resident, nstage = p.stages;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&sh.full[s], kFuse ? 2 * kFuseWarps : 1); mbar_init(&sh.empty[s], 1); }
        for (int s = 0; s < 16; ++s) { mbar_init(&sh.fullb[s], 1); mbar_init(&sh.emptyb[s], 1); }
        if (kFuse) for (int s = 0; s < nstage; ++s) mbar_init(&rawfull[s], 1);
        for (int i = 0; i < kBufs; ++i) { mbar_init(&sh.tfull[i], 1); mbar_init(&sh.tempty[i], 16); }   // 8 warps x 2 CTAs
        for (int i = 0; i < kStatSlots; ++i) { mbar_init(&sh.sfull[i], 8); mbar_init(&sh.sempty[i], 1); }
        if (kApply == 1 || kApply == 2) {
            const ApplyShared ap = apply_shared<N>(sh.misc);
            for (int i = 0; i < 2; ++i) { mbar_init(&ap.aff_full[i], 1); mbar_init(&ap.aff_empty[i], 16); }
        }
        if (kApply >= 3) {
            const AsyncShared as = async_shared<N>(sh.misc);
            for (int i = 0; i < 2; ++i) { mbar_init(&as.tab_full[i], 1); mbar_init(&as.tab_empty[i], kStreamWarps); }
        }
        mbar_fence_init();
    } else if (warp == 2) {
        tmem_alloc2(sh.tmem_slot, kBufs * N);
        tmem_relinquish2();
    }
    for (int i = tid; i < N; i += kThreads) sh.s_bias[i] = p.e.bias ? p.e.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                // the peer's barriers exist before anything targets them
    tc_fence_after();
    // Programmatic dependent launch: everything above overlapped the predecessor's tail.  The producer warp goes on to request the
    // resident weight tiles (parameters, not produced by any kernel of the step) BEFORE it waits for the predecessor's results.
    if (warp != 0) pdl_wait();
    const uint32_t tmem_base = *sh.tmem_slot;

    const int nck = p.nchunk0 + p.nchunk1;
    const int n_it = tc_num_iters(p);

    if (warp == 0) {
        // ================================================================ TMA producer (both CTAs)
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            const int G = (int)gridDim.x, ht = p.halo_t;
            const uint32_t a_tx = (uint32_t)(18 * pw * 128);
            const int wrow_off = (int)rank * (N / 2);                // my half of the weight rows
            TileWalk tw;
            tw.init(p, (int)blockIdx.x, G);
            if (resident) {                                          // all my half tiles, counted on the leader's fullb[0]
                if (rank == 0) mbar_expect_tx(&sh.fullb[0], (uint32_t)(2 * nslot * kBHalf));
                const uint32_t fb = mapa_u32(smem_u32(&sh.fullb[0]), 0u);
                if (kConvT) {                                        // slot (ph*4 + t)*nck + ck
                    for (int pt = 0; pt < 16; ++pt)
                        for (int ck = 0; ck < nck; ++ck)
                            tma_load_2d_2sm(&mapWh, fb, smem_b + (size_t)(pt * nck + ck) * kBHalf, ck * 64,
                                            p.wrow[pt >> 2][pt & 3] + wrow_off);
                } else {
                    for (int ck = 0; ck < nck; ++ck)
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_2d_2sm(&mapWh, fb, smem_b + (size_t)(ck * 9 + tap) * kBHalf, ck * 64, p.wrow[0][tap] + wrow_off);
                }
            }
            pdl_wait();                                              // activations, masks, statistics of the predecessors from here on
            for (int it = 0; it < n_it; ++it, tw.advance(G)) {
                if (kStats && (it == n_it - 8 || it == n_it - 1)) prefetch_l2(p.e.gn_counters);
                const int b = tw.b, h0 = tw.th * p.bh, w0 = tw.tw * p.bw;
                for (int ck = 0; ck < nck; ++ck) {
                    mbar_wait(&sh.empty[sa], pha ^ 1u);
                    int which, chan;
                    tc_chunk_src(p, ck, &which, &chan);
                    if (kFuse) {
                        mbar_expect_tx(&rawfull[sa], a_tx);          // my own tile only; the transform warps pass it on
                        tma_load_4d(which ? &mapA1 : &mapA0, &rawfull[sa], smem + (size_t)sa * a_stage,
                                    chan, ht ? h0 - 1 : w0 - 1, ht ? w0 - 1 : h0 - 1, b);
                    } else {
                        if (rank == 0) mbar_expect_tx(&sh.full[sa], 2u * a_tx);
                        tma_load_4d_2sm(which ? &mapA1 : &mapA0, mapa_u32(smem_u32(&sh.full[sa]), 0u), smem + (size_t)sa * a_stage,
                                        chan, ht ? h0 - 1 : w0 - 1, ht ? w0 - 1 : h0 - 1, b);
                    }
                    if (++sa == nstage) { sa = 0; pha ^= 1u; }
                    if (!resident && !kConvT) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&sh.emptyb[sb], phb ^ 1u);
                            if (rank == 0) mbar_expect_tx(&sh.fullb[sb], (uint32_t)(2 * kBHalf));
                            tma_load_2d_2sm(&mapWh, mapa_u32(smem_u32(&sh.fullb[sb]), 0u), smem_b + (size_t)sb * kBHalf, ck * 64,
                                            p.wrow[0][tap] + wrow_off);
                            if (++sb == nslot) { sb = 0; phb ^= 1u; }
                        }
                    }
                }
                if (kConvT && !resident) {                           // streamed weights in the order the MMA warp uses them
                    for (int pt = 0; pt < 16; ++pt)
                        for (int ck = 0; ck < nck; ++ck) {
                            mbar_wait(&sh.emptyb[sb], phb ^ 1u);
                            if (rank == 0) mbar_expect_tx(&sh.fullb[sb], (uint32_t)(2 * kBHalf));
                            tma_load_2d_2sm(&mapWh, mapa_u32(smem_u32(&sh.fullb[sb]), 0u), smem_b + (size_t)sb * kBHalf, ck * 64,
                                            p.wrow[pt >> 2][pt & 3] + wrow_off);
                            if (++sb == nslot) { sb = 0; phb ^= 1u; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (leader CTA only)
        if (rank == 0) {
            const int dbg = p.dbg;
            const uint64_t a_desc0 = make_sw128_kmajor_desc(smem_u32(smem), (uint32_t)(pw * 128), 0u);
            const uint64_t b_desc0 = make_sw128_kmajor_desc(smem_u32(smem_b));
            const uint64_t a_stage_step = (uint64_t)(a_stage >> 4), b_slot_step = (uint64_t)(kBHalf >> 4);
            uint64_t tap_off[9];
#pragma unroll
            for (int t = 0; t < 9; ++t)
                tap_off[t] = (uint64_t)(((p.halo_t ? (t % 3) * pw + (t / 3) : (t / 3) * pw + (t % 3)) * 128) >> 4);
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            if (kConvT) {
                // per spatial tile: wait for its nck halo boxes, then phase by phase 4 taps x nck chunks x 4 MMAs into the
                // phase's own accumulator (epilogue iteration 4*it + ph)
                uint64_t ct_off[16];
#pragma unroll
                for (int pt = 0; pt < 16; ++pt) {
                    const int dy = p.dy[pt >> 2][pt & 3] + 1, dx = p.dx[pt >> 2][pt & 3] + 1;
                    ct_off[pt] = (uint64_t)(((p.halo_t ? dx * pw + dy : dy * pw + dx) * 128) >> 4);
                }
                for (int it = 0; it < n_it; ++it) {
                    const int sa0 = sa;
                    for (int ck = 0; ck < nck; ++ck) {
                        mbar_wait(&sh.full[sa], pha);
                        if (++sa == nstage) { sa = 0; pha ^= 1u; }
                    }
                    if (resident && it == 0) mbar_wait(&sh.fullb[0], 0u);
#pragma unroll
                    for (int ph = 0; ph < 4; ++ph) {
                        const int e = it * 4 + ph, buf = e % kBufs;
                        mbar_wait(&sh.tempty[buf], ((uint32_t)(e / kBufs) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            for (int ck = 0; ck < nck; ++ck) {
                                int st = sa0 + ck; if (st >= nstage) st -= nstage;
                                const uint64_t adesc = a_desc0 + (uint64_t)st * a_stage_step + ct_off[ph * 4 + t];
                                uint64_t bdesc;
                                if (resident) bdesc = b_desc0 + (uint64_t)((ph * 4 + t) * nck + ck) * b_slot_step;
                                else { mbar_wait(&sh.fullb[sb], phb); tc_fence_after(); bdesc = b_desc0 + (uint64_t)sb * b_slot_step; }
                                if (elect_one()) {
                                    if (!(dbg & 1)) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            tc_mma2_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                                        (uint32_t)((t | ck | k) != 0));
                                    }
                                    if (!resident) tc_commit2_mc(&sh.emptyb[sb], (uint16_t)3);
                                    if (t == 3 && ck == nck - 1) {
                                        tc_commit2_mc(&sh.tfull[buf], (uint16_t)3);
                                        if (ph == 3)
                                            for (int c2 = 0; c2 < nck; ++c2) {
                                                int s2 = sa0 + c2; if (s2 >= nstage) s2 -= nstage;
                                                tc_commit2_mc(&sh.empty[s2], (uint16_t)3);
                                            }
                                    }
                                }
                                __syncwarp();
                                if (!resident) { if (++sb == nslot) { sb = 0; phb ^= 1u; } }
                            }
                        }
                    }
                }
            } else
            for (int it = 0; it < n_it; ++it) {
                const int buf = it % kBufs;
                mbar_wait(&sh.tempty[buf], ((uint32_t)(it / kBufs) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                for (int ck = 0; ck < nck; ++ck) {
                    mbar_wait(&sh.full[sa], pha);
                    const uint64_t adesc = a_desc0 + (uint64_t)sa * a_stage_step;
                    if (resident) {
                        if (it == 0 && ck == 0) mbar_wait(&sh.fullb[0], 0u);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t bdesc = b_desc0 + (uint64_t)(ck * 9) * b_slot_step;
                            if (!(dbg & 1)) {
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc_mma2_f16(d_tmem, adesc + tap_off[tap] + (uint64_t)(2 * k),
                                                    bdesc + (uint64_t)tap * b_slot_step + (uint64_t)(2 * k), kIdesc,
                                                    (uint32_t)((ck | tap | k) != 0));
                            }
                            tc_commit2_mc(&sh.empty[sa], (uint16_t)3);
                            if (ck == nck - 1) tc_commit2_mc(&sh.tfull[buf], (uint16_t)3);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&sh.fullb[sb], phb);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t bdesc = b_desc0 + (uint64_t)sb * b_slot_step;
                                if (!(dbg & 1)) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc_mma2_f16(d_tmem, adesc + tap_off[tap] + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                                    (uint32_t)((ck | tap | k) != 0));
                                }
                                tc_commit2_mc(&sh.emptyb[sb], (uint16_t)3);
                                if (tap == 8) {
                                    tc_commit2_mc(&sh.empty[sa], (uint16_t)3);
                                    if (ck == nck - 1) tc_commit2_mc(&sh.tfull[buf], (uint16_t)3);
                                }
                            }
                            __syncwarp();
                            if (++sb == nslot) { sb = 0; phb ^= 1u; }
                        }
                    }
                    if (++sa == nstage) { sa = 0; pha ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        if (kApply >= 3) tc_stats_async_loop<N>(p, sh, lane);
        else if (kApply) tc_stats_apply_loop<N>(p, sh, lane);
        else tc_stats_loop<kStats>(p, sh, lane);
    } else if (warp >= 4 && warp < kThreads / 32) {
        if (kApply == 1 || kApply == 2) tc_epilogue_apply_loop<N, kApply == 1, kApply == 2>(p, sh, tmem_base, warp, lane);
        else tc_epilogue_loop<N, kStats, false, kMask, kOutF32>(p, sh, tmem_base, warp, lane);
    } else if (kApply >= 3 && warp == kThreads / 32) {
        if (!(p.dbg & 128)) tc_apply_finaliser<N>(p, sh, lane);       // dbg 128: no apply work at all (timing experiments)
    } else if (kApply >= 3 && warp > kThreads / 32) {
        if (!(p.dbg & 128)) tc_apply_stream<N, kApply == 3, kApply == 4>(p, sh, tid - kThreads - 32);
    } else if (kFuse && warp >= kThreads / 32) {
        // ================================================================ input transform (both CTAs)
        // thread -> 16-byte chunk j (8 channels) of rows r0, r0+16, ...; the 128-byte swizzle puts chunk j of row r at
        // chunk position j ^ (r & 7) (stages are 1024-byte aligned).
        const ConvEpilogue& e = p.e;
        const int t = tid - kThreads, j = t & 7, r0 = t >> 3;       // 128 threads: 8 chunks x 16 rows per pass
        const int G = (int)gridDim.x, ht = p.halo_t;
        const uint32_t full_leader = mapa_u32(smem_u32(&sh.full[0]), 0u);
        TileWalk tw;
        tw.init(p, (int)blockIdx.x, G);
        int sa = 0;
        uint32_t pha = 0;
        int cur_b = -1, cur_ck = -1;
        float2 sc[4], sh2[4], tb[4];
        const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
        for (int it = 0; it < n_it; ++it, tw.advance(G)) {
            const int b = tw.b, h0 = tw.th * p.bh - 1, w0 = tw.tw * p.bw - 1;     // image coordinates of box row 0
            const bool live = b < p.B;                                              // dummy tile of an odd tail: zeros
            for (int ck = 0; ck < nck; ++ck) {
                if (live && (b != cur_b || ck != cur_ck)) {                         // per-(sample, chunk) affine constants
                    cur_b = b; cur_ck = ck;
                    const int c0 = ck * 64 + j * 8, Cin = nck * 64;
                    const int g = (c0 * 8) / Cin;
                    const float mean = e.in_stats[(b * 8 + g) * 2], rstd = e.in_stats[(b * 8 + g) * 2 + 1];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float s0 = rstd * __ldg(e.in_gamma + c0 + 2 * q), s1 = rstd * __ldg(e.in_gamma + c0 + 2 * q + 1);
                        const float h0f = fmaf(-mean, s0, __ldg(e.in_beta + c0 + 2 * q)), h1f = fmaf(-mean, s1, __ldg(e.in_beta + c0 + 2 * q + 1));
                        sc[q] = make_float2(s0, s1);
                        sh2[q] = make_float2(h0f, h1f);
                        tb[q] = e.in_tbias ? make_float2(__ldg(e.in_tbias + (size_t)b * e.in_tb_bstride + c0 + 2 * q),
                                                         __ldg(e.in_tbias + (size_t)b * e.in_tb_bstride + c0 + 2 * q + 1))
                                           : make_float2(0.f, 0.f);
                    }
                }
                mbar_wait(&rawfull[sa], pha);
                uint8_t* st = smem + (size_t)sa * a_stage;
#pragma unroll 2
                for (int r = r0; r < 18 * pw; r += 16) {
                    const int a0 = r / pw, a1 = r - a0 * pw;                         // (slow, fast) box coordinates
                    const int h = h0 + (ht ? a1 : a0), w = w0 + (ht ? a0 : a1);
                    uint4* cp = reinterpret_cast<uint4*>(st + r * 128 + ((j ^ (r & 7)) << 4));
                    const bool inb = live && h >= 0 && h < p.Hg && w >= 0 && w < p.Wg;
                    uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
                    if (inb) {
                        const float m = __ldg(e.in_mask + (size_t)b * p.Wg + w);
                        const float2 m2 = make_float2(m, m);
                        const uint4 v = *cp;
                        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 x = make_float2(__uint_as_float(wv[q] << 16), __uint_as_float(wv[q] & 0xffff0000u));
                            const float2 y = ffma2(x, sc[q], sh2[q]);
                            float2 o = mish2_fast(y, fmul2(y, l2e));
                            o = fmul2(fadd2(o, tb[q]), m2);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o.x, o.y);
                            ow[q] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        o4 = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                    *cp = o4;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(full_leader + (uint32_t)(sa * 8));
                if (++sa == nstage) { sa = 0; pha ^= 1u; }
            }
        }
    }
    tc_teardown<N, kStats && !kApply, true>(p, sh, smem, tmem_base, tid, warp, lane);   // kApply finalises its statistics per sample, in the loop
}

template <int N, bool kStats, bool kFuse, bool kMask = false, bool kConvT = false, int kApply = 0, bool kOutF32 = false>
int launch_halo2(const TcConvPlan* pl, cudaStream_t stream) {
    static bool attr_set = false;
    auto k = conv_tc_halo2_kernel<N, kStats, kFuse, kMask, kConvT, kApply, kOutF32>;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GTTS_CHECK_CUDA(launch_pdl(k, dim3(pl->grid), dim3(kThreads + (kFuse ? kFuseWarps * 32 : 0) + (kApply >= 3 ? kApplyWarps * 32 : 0)), pl->smem,
                               stream, 2, pl->mapA0, pl->mapA1, pl->mapWh, pl->p));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// Largest grid (CTAs, even) whose 2-CTA clusters are all co-resident for the apply variant with this much dynamic shared memory: the
// per-sample grid barrier needs every CTA of the grid on an SM at the same time.  0 if the query fails.
namespace {
template <int N, int kApply>
int max_grid_of(size_t smem) {
    auto k = conv_tc_halo2_kernel<N, true, false, false, false, kApply>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(kThreads + (kApply >= 3 ? kApplyWarps * 32 : 0)); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return 2 * n;
}
}  // namespace

int conv_tc_halo2_max_grid(int N, bool residual, size_t smem, bool async_apply) {
    if (async_apply) {
        if (residual) return N == 64 ? max_grid_of<64, 4>(smem) : (N == 128 ? max_grid_of<128, 4>(smem) : max_grid_of<256, 4>(smem));
        return N == 64 ? max_grid_of<64, 3>(smem) : (N == 128 ? max_grid_of<128, 3>(smem) : max_grid_of<256, 3>(smem));
    }
    if (residual) return N == 64 ? max_grid_of<64, 2>(smem) : (N == 128 ? max_grid_of<128, 2>(smem) : max_grid_of<256, 2>(smem));
    return N == 64 ? max_grid_of<64, 1>(smem) : (N == 128 ? max_grid_of<128, 1>(smem) : max_grid_of<256, 1>(smem));
}
int conv_tc_halo2_apply_extra_smem(int apply) { return apply == 2 ? kAsyncExtra : kApplyExtra; }

int conv_tc_halo2_launch(const TcConvPlan* pl, cudaStream_t stream) {
    const ConvEpilogue& e = pl->p.e;
    if (pl->p.ph_inner == 4) {                                      // transposed conv: mask epilogue, no statistics
        GTTS_REQUIRE(e.residual == nullptr && e.mask != nullptr && e.gn_partials == nullptr && !e.in_stats,
                     "conv_tc_halo2: ConvT variant is built with the mask epilogue only");
        if (pl->N == 64) return launch_halo2<64, false, false, true, true>(pl, stream);
        if (pl->N == 128) return launch_halo2<128, false, false, true, true>(pl, stream);
        set_error("conv_tc_halo2: ConvT variant supports 64 / 128 channels");
        return 2;
    }
    if (e.apply) {                                                  // GroupNorm-apply epilogue: time-bias or residual flavour
        GTTS_REQUIRE(e.gn_partials && e.gn_stats && e.gn_counters && e.mask && e.ap_gamma && e.ap_beta && !e.in_stats,
                     "conv_tc_halo2: the apply epilogue needs statistics buffers, affine parameters and the mask");
        GTTS_REQUIRE((e.residual != nullptr) != (e.ap_tbias != nullptr), "conv_tc_halo2: the apply epilogue takes a time bias or a residual");
        if (e.apply == 2) {                                         // asynchronous apply warps: raw tile to e.out, finished activation to e.ap_out
            GTTS_REQUIRE(e.ap_out != nullptr, "conv_tc_halo2: the asynchronous apply variant needs ap_out");
            if (e.residual) {
                if (pl->N == 64) return launch_halo2<64, true, false, false, false, 4>(pl, stream);
                if (pl->N == 128) return launch_halo2<128, true, false, false, false, 4>(pl, stream);
                if (pl->N == 256) return launch_halo2<256, true, false, false, false, 4>(pl, stream);
            } else {
                if (pl->N == 64) return launch_halo2<64, true, false, false, false, 3>(pl, stream);
                if (pl->N == 128) return launch_halo2<128, true, false, false, false, 3>(pl, stream);
                if (pl->N == 256) return launch_halo2<256, true, false, false, false, 3>(pl, stream);
            }
            set_error("conv_tc_halo2: unsupported Cout");
            return 2;
        }
        if (e.residual) {
            if (pl->N == 64) return launch_halo2<64, true, false, false, false, 2>(pl, stream);
            if (pl->N == 128) return launch_halo2<128, true, false, false, false, 2>(pl, stream);
            if (pl->N == 256) return launch_halo2<256, true, false, false, false, 2>(pl, stream);
        } else {
            if (pl->N == 64) return launch_halo2<64, true, false, false, false, 1>(pl, stream);
            if (pl->N == 128) return launch_halo2<128, true, false, false, false, 1>(pl, stream);
            if (pl->N == 256) return launch_halo2<256, true, false, false, false, 1>(pl, stream);
        }
        set_error("conv_tc_halo2: unsupported Cout");
        return 2;
    }
    GTTS_REQUIRE(e.residual == nullptr && e.mask == nullptr, "conv_tc_halo2: plain or GN-statistics epilogue only");
    const bool st = e.gn_partials != nullptr;
    if (e.out_f32) {                                                // fp32 activations (fp32 mode on the tensor cores)
        GTTS_REQUIRE(!e.in_stats, "conv_tc_halo2: fp32 output excludes the fused input transform");
        if (pl->N == 64) return st ? launch_halo2<64, true, false, false, false, 0, true>(pl, stream) : launch_halo2<64, false, false, false, false, 0, true>(pl, stream);
        if (pl->N == 128) return st ? launch_halo2<128, true, false, false, false, 0, true>(pl, stream) : launch_halo2<128, false, false, false, false, 0, true>(pl, stream);
        if (pl->N == 256) return st ? launch_halo2<256, true, false, false, false, 0, true>(pl, stream) : launch_halo2<256, false, false, false, false, 0, true>(pl, stream);
    }
    if (e.in_stats) {                                               // fused input transform (block2 convs: always with stats)
        GTTS_REQUIRE(st, "conv_tc_halo2: the fused-input variant is built with GroupNorm statistics only");
        if (pl->N == 64) return launch_halo2<64, true, true>(pl, stream);
        if (pl->N == 128) return launch_halo2<128, true, true>(pl, stream);
        if (pl->N == 256) return launch_halo2<256, true, true>(pl, stream);
    } else {
        if (pl->N == 64) return st ? launch_halo2<64, true, false>(pl, stream) : launch_halo2<64, false, false>(pl, stream);
        if (pl->N == 128) return st ? launch_halo2<128, true, false>(pl, stream) : launch_halo2<128, false, false>(pl, stream);
        if (pl->N == 256) return st ? launch_halo2<256, true, false>(pl, stream) : launch_halo2<256, false, false>(pl, stream);
    }
    set_error("conv_tc_halo2: unsupported Cout");
    return 2;
}

}  // namespace gtts
