// Shared device/host pieces of the tcgen05 convolution kernels (conv_tc.cu: one TMA box per tap;
// conv_tc_halo.cu: one halo box per tile, taps as shifted descriptor views).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ops.h"

// Per-phase cycle counters of the epilogue loop (GTTS_CONV_TIMING=1): compiled in only with -DGTTS_EPI_TIMING=1, because
// six clock reads per tile cost ~8 % of a 64->64 conv.
#ifndef GTTS_EPI_TIMING
#define GTTS_EPI_TIMING 0
#endif
#if GTTS_EPI_TIMING
#define GTTS_EPI_CLK() clock64()
#else
#define GTTS_EPI_CLK() 0ll
#endif

namespace gtts {
namespace tc {

constexpr int kABytes = 128 * 128;                 // 128 pixel rows x 64 bf16
constexpr int kMiscBytes = 8192;                   // barriers + epilogue scratch
constexpr int kHaloABytes = 18 * 16 * 128;         // halo box: 18 rows x 16 pixels x 64 bf16
constexpr int kStatSlots = 8;                      // ring of per-tile GroupNorm partials (epilogue warps -> stats warp)
constexpr int kThreads = 640;                      // warps 0-3: TMA, MMA, TMEM alloc, stats; warps 4-19: two epilogue groups
template <int N> __host__ __device__ constexpr int acc_bufs() { return N <= 64 ? 8 : (N <= 128 ? 4 : 2); }   // TMEM accumulator buffers (N cols each)

struct TcParams {
    int bh, bw, tiles_h, tiles_w, nphase, B;
    int Hg, Wg, Hout, Wout, out_step;
    int ntaps, nchunk0, nchunk1, Cin0;
    int stride2, w_batch_rows, num_tiles, a_bytes, stages;
    int halo_mode, b_slots, b_resident, halo_prefetch;   // conv_tc_halo.cu only
    int ph_inner;                                  // > 1: that many output phases per spatial tile, consecutive epilogue iterations (ConvT on the halo kernel)
    int halo_t;                                    // 1: halo tile is 8 rows x 16 pixels with H as the fast box dimension
    int pass_tiles;                                // streamed weights: A tiles that share one pass of the weight ring (1 or 2)
    int mc;                                        // 1: 2-CTA cluster, weight tiles multicast (conv_tc.cu); 2: CTA pair, cta_group::2 MMAs (conv_tc_halo2.cu)
    unsigned long long* dbg_out;                   // optional per-CTA cycle counters (GTTS_CONV_TIMING), 16 per CTA
    int dbg;                                       // experiment switches (GTTS_CONV_DBG): 1 no MMA issue, 2 no epilogue work, 4 no A loads (halo), 8 no stats ring
    int8_t dy[4][kMaxTaps], dx[4][kMaxTaps];
    int wrow[4][kMaxTaps];
    int oy[4], ox[4];
    // split (fp32-on-tensor-cores) convs: K chunk ck reads source a_map[ck] at channel a_off[ck] of its 3*Cin-channel plane tensor
    int split;
    int16_t a_off[48];
    int8_t a_map[48];
    // 1-D halo mode of the per-tap kernel (conv_tc.cu; H = 1, stride 1, resident weights): ONE box of bw + 2*halo positions per
    // 64-channel chunk, the taps are row-shifted descriptor views of it (tap_row[t] = halo + dx[t] rows of 128 bytes)
    int halo1d, a_stage;
    int16_t tap_row[kMaxTaps];
    ConvEpilogue e;
};

// misc shared-memory block layout (relative to `misc`)
//   [0,   768)  mbarriers + TMEM slot   [768, 1792) bias[256]   [1792, 3840) stats ring   [3840, 4096) tail flags   [4096, 8192) per-sample statistics rows (<= 64 samples)
struct TcShared {
    uint64_t *full, *empty, *tfull, *tempty, *sfull, *sempty, *fullb, *emptyb;
    uint32_t* tmem_slot;
    float* s_bias;
    float* s_ring;
    uint8_t* misc;
};
__device__ __forceinline__ TcShared tc_shared(uint8_t* misc) {
    TcShared s;
    s.misc = misc;
    s.full = reinterpret_cast<uint64_t*>(misc);        // [8]
    s.empty = s.full + 8;                              // [8]
    s.tfull = s.empty + 8;                             // [8]
    s.tempty = s.tfull + 8;                            // [8]
    s.sfull = s.tempty + 8;                            // [kStatSlots]
    s.sempty = s.sfull + kStatSlots;                   // [kStatSlots]
    s.fullb = s.sempty + kStatSlots;                   // [16]
    s.emptyb = s.fullb + 16;                           // [16]
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.emptyb + 16);
    s.s_bias = reinterpret_cast<float*>(misc + 768);
    s.s_ring = reinterpret_cast<float*>(misc + 1792);
    return s;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor.  sbo = byte distance between 8-row groups,
// base_offset = swizzle phase of the start row (bits 7..9 of the start address within the 1024-byte pattern).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr, uint32_t sbo = 1024, uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(sbo >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)(base_offset & 7u) << 49;       // matrix base offset
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    // kind::f16: D=f32 (bit 4), A=bf16 (bit 7), B=bf16 (bit 10), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// (source, channel offset) of K chunk ck
__device__ __forceinline__ void tc_chunk_src(const TcParams& p, int ck, int* which, int* chan) {
    if (p.split) { *which = p.a_map[ck]; *chan = p.a_off[ck]; }
    else if (ck < p.nchunk0) { *which = 0; *chan = ck * 64; }
    else { *which = 1; *chan = (ck - p.nchunk0) * 64; }
}

// Number of tile iterations of this CTA.  With weight multicast both CTAs of a pair must run the same count, so the
// count is taken from the pair's first CTA; the odd one may then get a dummy tile (tile >= num_tiles).
__device__ __forceinline__ int tc_num_iters(const TcParams& p) {
    const int G = (int)gridDim.x, bx = p.mc ? ((int)blockIdx.x & ~1) : (int)blockIdx.x;
    return p.num_tiles > bx ? (p.num_tiles - bx + G - 1) / G : 0;
}

// Tile walker: decodes tile = blockIdx.x + it*gridDim.x into (tw, th, ph, b) incrementally (no per-tile divisions).
struct TileWalk {
    int tile, tw, th, ph, b;
    int d_tw, d_th, d_ph, d_b, tiles_w, tiles_h, nphase;
    __device__ __forceinline__ void init(const TcParams& p, int start, int step) {
        tiles_w = p.tiles_w; tiles_h = p.tiles_h; nphase = p.nphase;
        tile = start;
        tw = start % tiles_w; int r = start / tiles_w;
        th = r % tiles_h; r /= tiles_h;
        ph = r % nphase; b = r / nphase;
        d_tw = step % tiles_w; r = step / tiles_w;
        d_th = r % tiles_h; r /= tiles_h;
        d_ph = r % nphase; d_b = r / nphase;
    }
    __device__ __forceinline__ void advance(int step) {
        tile += step;
        tw += d_tw; if (tw >= tiles_w) { tw -= tiles_w; ++th; }
        th += d_th; if (th >= tiles_h) { th -= tiles_h; ++ph; }
        ph += d_ph; if (ph >= nphase) { ph -= nphase; ++b; }
        b += d_b;
    }
};

// Butterfly transpose-reduce of 8 per-thread values across the warp in 9 shuffles (instead of 40):
// afterwards every lane holds the full 32-lane sum of value index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1).
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
    float w[4], u[2], t;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b4 ? v[4 + i] : v[i], send = b4 ? v[i] : v[4 + i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b3 ? w[2 + i] : w[i], send = b3 ? w[i] : w[2 + i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float keep = b2 ? u[1] : u[0], send = b2 ? u[0] : u[1];
        t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// Common prologue: barrier init, TMEM allocation, bias staging.  `stages` A/B ring slots, `b_slots` extra B ring slots.
template <int N>
__device__ __forceinline__ uint32_t tc_prologue(const TcParams& p, const TcShared& sh, int nfull, int nfullb, int tid,
                                                int warp, int lane, bool producer_waits_later = false) {
    pdl_trigger();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < nfull; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], p.mc ? 2 : 1); }
        for (int s = 0; s < nfullb; ++s) { mbar_init(&sh.fullb[s], 1); mbar_init(&sh.emptyb[s], 1); }
        for (int i = 0; i < acc_bufs<N>(); ++i) { mbar_init(&sh.tfull[i], 1); mbar_init(&sh.tempty[i], 8); }   // one arrive per epilogue warp
        for (int i = 0; i < kStatSlots; ++i) { mbar_init(&sh.sfull[i], 8); mbar_init(&sh.sempty[i], 1); }
        mbar_fence_init();
    } else if (warp == 2) {
        tmem_alloc(sh.tmem_slot, acc_bufs<N>() * N);
        tmem_relinquish();
    }
    for (int i = tid; i < N; i += kThreads) sh.s_bias[i] = p.e.bias ? p.e.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    if (p.mc) cluster_sync();                      // the peer's barriers are initialised before anything targets them
    tc_fence_after();
    // everything above overlapped the predecessor's tail; a producer warp that first requests resident weight tiles (parameters,
    // not produced by any kernel of the step) calls pdl_wait() itself after that
    if (!(producer_waits_later && warp == 0)) pdl_wait();
    return *sh.tmem_slot;
}

// GroupNorm statistics warp (warp 3): sums the 8 epilogue warps' per-tile values, accumulates them per sample in
// registers (a CTA's tiles of one sample are consecutive in its walk) and writes ONE 16-float partial per
// (sample, CTA): partials[(b * gridDim.x + blockIdx.x) * 16 + k].  Fixed tile->CTA assignment and order: deterministic.
template <bool kStats>
__device__ __forceinline__ void tc_stats_loop(const TcParams& p, const TcShared& sh, int lane) {
    uint64_t* sfull = sh.sfull;
    uint64_t* sempty = sh.sempty;
    float* s_ring = sh.s_ring;
    const int tiles_per_phase = p.tiles_h * p.tiles_w;
    if (kStats && !(p.dbg & 8)) {
        const ConvEpilogue& e = p.e;
        const int G = (int)gridDim.x;
        int cur_b = -1;
        float acc = 0.f;
        // value k (0..7 sums, 8..15 sums of squares) of group g = k & 7 lives in column half g >> 2
        const int g = lane & 7, which = (lane >> 3) & 1, half = g >> 2, idx = which * 4 + (g & 3);
        const int n_it = tc_num_iters(p);
        // Per-sample rows are staged in shared memory and written to global memory only at the very end, all samples at
        // once: the finalising CTA then finds every row in L2 (rows written 100+ us earlier had been evicted to DRAM by
        // the kernel's own output stream, and reading them back cost ~5 us of tail).  Untouched samples read as zero.
        float* s_rows = reinterpret_cast<float*>(sh.misc + 4096);    // [B <= 64][16]
        for (int i = lane; i < p.B * 16; i += 32) s_rows[i] = 0.f;
        __syncwarp();
        for (int it = 0; it < n_it; ++it) {
            const int tile = (int)blockIdx.x + it * G;
            const int slot = it % kStatSlots;
            const bool dummy = tile >= p.num_tiles;
            const int b = dummy ? cur_b : tile / tiles_per_phase;
            if (b != cur_b) {
                if (cur_b >= 0 && lane < 16) s_rows[cur_b * 16 + lane] = acc;
                cur_b = b;
                acc = 0.f;
            }
            mbar_wait(&sfull[slot], (uint32_t)(it / kStatSlots) & 1u);
            if (lane < 16 && !dummy) {
                const float* r = s_ring + (slot * 8 + half * 4) * 8 + idx;
                acc += (r[0] + r[8]) + (r[16] + r[24]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[slot]);
        }
        if (cur_b >= 0 && lane < 16) s_rows[cur_b * 16 + lane] = acc;
        __syncwarp();
        for (int i = lane; i < p.B * 16; i += 32)                    // two 64-byte rows per warp-wide store
            e.gn_partials[((size_t)(i >> 4) * G + blockIdx.x) * 16 + (i & 15)] = s_rows[i];
    }
}

// Epilogue warps (warps 4..19): TMEM -> registers -> (+bias, stats, +residual, *mask) -> bf16 NHWC stores.
// Two groups of 8 warps take alternate tiles (group = tile iteration parity), so one group's TMEM reads, arithmetic
// and stores overlap the other's; within a group two warps share each TMEM lane quarter and split the columns.
template <int N, bool kStats, bool kRes, bool kMask, bool kOutF32 = false, bool kAct = false>
__device__ __forceinline__ void tc_epilogue_loop(const TcParams& p, const TcShared& sh, uint32_t tmem_base, int warp,
                                                 int lane) {
    constexpr int kColsPerWarp = N / 2;            // two epilogue warps share each TMEM lane quarter
    constexpr int kGsz = N / 8;                    // channels per GroupNorm group (4 groups per column half)
    constexpr int kBufs = acc_bufs<N>();
    const int ew16 = warp - 4, grp = ew16 >> 3, ew = ew16 & 7, wq = ew & 3, half = ew >> 2;
    const int row = wq * 32 + lane;                                  // TMEM lane = pixel row of the tile
    const ConvEpilogue& e = p.e;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(e.residual);
    const int cbase = half * kColsPerWarp;
    int hl = row / p.bw, wl = row - hl * p.bw;                       // fixed for the whole kernel
    if (p.halo_t) { wl = row >> 3; hl = row & 7; }                   // transposed halo tile: 8-row groups run along H
    const bool row_in_tile = hl < p.bh;
    const float* s_bias = sh.s_bias;
    TileWalk tw;
    const int G = (int)gridDim.x;
    // Epilogue iteration `it` = accumulator index.  Normally one per tile (this group takes every second tile); with
    // ph_inner = P > 1 the walk is over spatial tiles and iteration it = P * tile_iteration + phase.
    const int nph = p.ph_inner > 1 ? p.ph_inner : 1;
    if (nph > 1) tw.init(p, (int)blockIdx.x, G);
    else tw.init(p, (int)blockIdx.x + grp * G, 2 * G);
    const int n_it = tc_num_iters(p) * nph;
    int walk_it = 0;                                                 // spatial iteration tw currently points at (nph > 1)
    long long c_wait = 0, c_ld = 0, c_sring = 0, c_math = 0, c_store = 0, c_red = 0;
    const long long te0 = GTTS_EPI_CLK();
    for (int it = grp; it < n_it; it += 2) {
        if (nph > 1) { while (walk_it < it / nph) { tw.advance(G); ++walk_it; } }
        else if (it != grp) tw.advance(2 * G);
        const int buf = it % kBufs;
        const int b = tw.b, ph = nph > 1 ? it % nph : tw.ph;
        const int j = tw.th * p.bh + hl, i = tw.tw * p.bw + wl;
        const bool valid = row_in_tile && (j < p.Hg) && (i < p.Wg) && (tw.tile < p.num_tiles);
        const bool all_valid = __all_sync(0xffffffffu, valid);
        const int oh = j * p.out_step + p.oy[ph], ow = i * p.out_step + p.ox[ph];
        const size_t opix = valid ? ((size_t)b * p.Hout + oh) * p.Wout + ow : 0;
        // Paired stores (bf16 outputs): lanes 2k and 2k+1 first write the two 32-byte sectors of lane 2k's pixel, then those of
        // lane 2k+1's, so one store instruction touches 16 lines instead of 32.  The row-per-thread pattern made the 1x1 / N = 64
        // kernels L1TEX-bound (ncu: l1tex throughput 77-90 % at 52-66 % of DRAM bandwidth, profiles/r02_ncu_thin_convs.md).
        constexpr bool kPairSt = !kOutF32 && !kAct && !kStats;   // (the statistics variants have no registers to spare: ptxas spills)
        const size_t opix_p = kPairSt ? __shfl_xor_sync(0xffffffffu, (unsigned long long)opix, 1) : 0;
        const bool valid_p = kPairSt ? __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0 : false;
        float m = 1.0f;
        if (kMask) m = valid ? e.mask[(size_t)b * p.Wout + ow] : 0.f;
        const float2 m2 = make_float2(m, m);

        const long long tq0 = GTTS_EPI_CLK();
        mbar_wait(&sh.tfull[buf], (uint32_t)(it / kBufs) & 1u);
        tc_fence_after();
        c_wait += GTTS_EPI_CLK() - tq0;
        if (p.dbg & 2) { tc_fence_before(); __syncwarp(); if (lane == 0) { if (p.mc == 2) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&sh.tempty[buf]), 0u)); else mbar_arrive(&sh.tempty[buf]); } if (kStats && !(p.dbg & 8)) { const int slot = it % kStatSlots; mbar_wait(&sh.sempty[slot], ((uint32_t)(it / kStatSlots) & 1u) ^ 1u); __syncwarp(); if (lane == 0) mbar_arrive(&sh.sfull[slot]); } continue; }
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * N + cbase);

        float2 ssum[4], ssq[4];                                      // per local group: packed (even, odd) columns
#pragma unroll
        for (int g = 0; g < 4; ++g) { ssum[g] = make_float2(0.f, 0.f); ssq[g] = make_float2(0.f, 0.f); }

#pragma unroll
        for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
            uint32_t r[32];
            const long long tl = GTTS_EPI_CLK();
            tmem_ld32(taddr + (uint32_t)c0, r);
            tmem_ld_wait();
            const long long tm0 = GTTS_EPI_CLK();
            c_ld += tm0 - tl;
            if (c0 + 32 >= kColsPerWarp) {
                // last TMEM read of this buffer: hand it back to the MMA warp before the arithmetic and stores
                // (one arrive per warp: 256 threads hammering one mbarrier word cost ~0.5 us per tile)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (p.mc == 2) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&sh.tempty[buf]), 0u));   // CTA pair: the leader issues the MMAs
                    else mbar_arrive(&sh.tempty[buf]);
                }
            }
            float2 f[16];
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[cbase + c0 + q4 * 4]);
                f[q4 * 2 + 0] = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 0]), __uint_as_float(r[q4 * 4 + 1])),
                                      make_float2(b4.x, b4.y));
                f[q4 * 2 + 1] = fadd2(make_float2(__uint_as_float(r[q4 * 4 + 2]), __uint_as_float(r[q4 * 4 + 3])),
                                      make_float2(b4.z, b4.w));
            }
            if (kStats) {
                if (all_valid) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int g = (c0 + 2 * q) / kGsz;           // local group 0..3 (compile time; kGsz is even)
                        ssum[g] = fadd2(ssum[g], f[q]);
                        ssq[g] = ffma2(f[q], f[q], ssq[g]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int g = (c0 + 2 * q) / kGsz;
                        // select, not multiply: rows outside the tile hold whatever was in shared memory (maybe NaN)
                        const float2 x = valid ? f[q] : make_float2(0.f, 0.f);
                        ssum[g] = fadd2(ssum[g], x);
                        ssq[g] = ffma2(x, x, ssq[g]);
                    }
                }
            }
            const long long tm1 = GTTS_EPI_CLK();
            c_math += tm1 - tm0;
            if (valid && kOutF32) {
                // fp32 activations (fp32 mode on the tensor cores): 32 columns = 128 bytes per pixel, four 32-byte stores
                float* op = reinterpret_cast<float*>(e.out) + opix * N + cbase + c0;
                const float* rp = reinterpret_cast<const float*>(e.residual) + opix * N + cbase + c0;
#pragma unroll
                for (int v8 = 0; v8 < 4; ++v8) {
                    uint32_t w[8];
                    if (kRes) {
                        uint32_t rr[8];
                        ld_global_nc_256(rp + v8 * 8, rr);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            f[v8 * 4 + k] = fadd2(f[v8 * 4 + k], make_float2(__uint_as_float(rr[2 * k]), __uint_as_float(rr[2 * k + 1])));
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 o = kMask ? fmul2(f[v8 * 4 + k], m2) : f[v8 * 4 + k];
                        w[2 * k] = __float_as_uint(o.x); w[2 * k + 1] = __float_as_uint(o.y);
                    }
                    st_global_256(op + v8 * 8, w);
                }
            } else if (kPairSt) {
                uint32_t wv[2][8];
                if (valid) {
                    if (kRes) {
                        const __nv_bfloat16* rp = res + opix * N + cbase + c0;
#pragma unroll
                        for (int v8 = 0; v8 < 2; ++v8) {             // 2 x 32 bytes = 32 bf16 residual values
                            uint32_t w[8];
                            ld_global_nc_256(rp + v8 * 16, w);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                f[v8 * 8 + k] = fadd2(f[v8 * 8 + k], make_float2(__uint_as_float(w[k] << 16),
                                                                                 __uint_as_float(w[k] & 0xffff0000u)));
                        }
                    }
                    if (kMask) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) f[q] = fmul2(f[q], m2);
                    }
#pragma unroll
                    for (int v8 = 0; v8 < 2; ++v8)
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[v8 * 8 + k].x, f[v8 * 8 + k].y);
                            wv[v8][k] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                }
                const bool odd = (lane & 1) != 0;
                uint32_t rx[8];                                      // the partner's half of the pixel I store
#pragma unroll
                for (int k = 0; k < 8; ++k) rx[k] = __shfl_xor_sync(0xffffffffu, odd ? wv[0][k] : wv[1][k], 1);
                __nv_bfloat16* op = out + opix * N + cbase + c0;
                __nv_bfloat16* opp = out + opix_p * N + cbase + c0;
                if (!(p.dbg & 64)) {
                    uint32_t w1[8], w2[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { w1[k] = odd ? rx[k] : wv[0][k]; w2[k] = odd ? wv[1][k] : rx[k]; }
                    if (odd ? valid_p : valid) st_global_256(odd ? opp + 16 : op, w1);        // the even lane's pixel
                    if (odd ? valid : valid_p) st_global_256(odd ? op + 16 : opp, w2);        // the odd lane's pixel
                }
            } else if (valid) {
                if (kRes) {
                    const __nv_bfloat16* rp = res + opix * N + cbase + c0;
#pragma unroll
                    for (int v8 = 0; v8 < 2; ++v8) {                 // 2 x 32 bytes = 32 bf16 residual values
                        uint32_t w[8];
                        ld_global_nc_256(rp + v8 * 16, w);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            f[v8 * 8 + k] = fadd2(f[v8 * 8 + k], make_float2(__uint_as_float(w[k] << 16),
                                                                             __uint_as_float(w[k] & 0xffff0000u)));
                    }
                }
                if (kMask) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = fmul2(f[q], m2);
                }
                __nv_bfloat16* op = out + opix * N + cbase + c0;
                if (kAct) {
                    // leaky-ReLU outputs of the 1-D conv stacks (ConvEpilogue::act_out / out2)
                    const float sl = e.act_slope;
                    if (!e.act_out) {
#pragma unroll
                        for (int v8 = 0; v8 < 2; ++v8) {
                            uint32_t w[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[v8 * 8 + k].x, f[v8 * 8 + k].y);
                                w[k] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            st_global_256(op + v8 * 16, w);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        f[q].x = f[q].x > 0.f ? f[q].x : f[q].x * sl;
                        f[q].y = f[q].y > 0.f ? f[q].y : f[q].y * sl;
                    }
                    if (!e.act_out) op = reinterpret_cast<__nv_bfloat16*>(e.out2) + opix * N + cbase + c0;
                }
#pragma unroll
                for (int v8 = 0; v8 < 2; ++v8) {                     // one full 32-byte sector per store instruction
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[v8 * 8 + k].x, f[v8 * 8 + k].y);
                        w[k] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    if (!(p.dbg & 64)) st_global_256(op + v8 * 16, w);
                }
            }
            c_store += GTTS_EPI_CLK() - tm1;
        }

        if (kStats && !(p.dbg & 8)) {
            float st[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) { st[g] = ssum[g].x + ssum[g].y; st[4 + g] = ssq[g].x + ssq[g].y; }
            const long long tr0 = GTTS_EPI_CLK();
            const float t = warp_reduce8(st, lane);
            const int slot = it % kStatSlots;
            const long long ts = GTTS_EPI_CLK();
            c_red += ts - tr0;
            mbar_wait(&sh.sempty[slot], ((uint32_t)(it / kStatSlots) & 1u) ^ 1u);
            c_sring += GTTS_EPI_CLK() - ts;
            if ((lane & 3) == 0) sh.s_ring[(slot * 8 + ew) * 8 + (lane >> 2)] = t;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.sfull[slot]);             // release: orders the ring writes of this warp
        }
    }
    if (p.dbg_out && ew16 == 0 && lane == 0) {
        unsigned long long* o = p.dbg_out + blockIdx.x * 32;
        o[8] = (unsigned long long)c_wait; o[9] = (unsigned long long)c_ld; o[13] = (unsigned long long)c_sring;
        o[14] = (unsigned long long)(GTTS_EPI_CLK() - te0);
        o[23] = (unsigned long long)c_math; o[24] = (unsigned long long)c_store; o[25] = (unsigned long long)c_red;
    }
}

// Kernel tail: one fence + ticket per CTA; whoever completes a sample's tile count reduces its partials (fixed order).
// kCta2 is a compile-time switch: a kernel that merely CONTAINS a cta_group::2 instruction must be launched as a cluster.
template <int N, bool kStats, bool kCta2 = false>
__device__ __forceinline__ void tc_teardown(const TcParams& p, const TcShared& sh, uint8_t* smem, uint32_t tmem_base,
                                            int tid, int warp, int lane) {
    constexpr int kGsz = N / 8;
    uint8_t* misc = sh.misc;
    const long long td0 = clock64();
    // only the statistics warp has written anything the finalising CTA will read (the partial rows); fencing the
    // epilogue warps too would make every CTA wait ~8 us for its streaming output stores to be acknowledged
    if (kStats && warp == 3) __threadfence();
    tc_fence_before();
    __syncthreads();
    const long long td1 = clock64();
    if (p.mc) cluster_sync();                      // nobody leaves while the peer may still multicast into it
    if (warp == 2) {
        const long long tq = clock64();
        tc_fence_after();
        if (kCta2) tmem_dealloc2(tmem_base, acc_bufs<N>() * N);
        else tmem_dealloc(tmem_base, acc_bufs<N>() * N);
        if (p.dbg_out && lane == 0) p.dbg_out[blockIdx.x * 32 + 22] = (unsigned long long)(clock64() - tq);
    }
    if (kStats) {
        // ONE ticket per CTA (not one per sample).  The CTA that draws the last ticket reduces every sample: one warp per sample,
        // all partial rows of a sample requested before the first add (one L2 round trip), double accumulation in a
        // fixed order -> deterministic.  (The previous tail -- per-sample tickets, samples reduced one after another --
        // cost 55 us at 16 samples because the last CTA is last for every sample.)
        const ConvEpilogue& e = p.e;
        int* s_last = reinterpret_cast<int*>(misc + 3840);
        if (tid == 0) {
            // two-level ticket: same-address atomics serialise at ~100 cycles each in L2, and all CTAs arrive together
            // (148 on one word measured 14k cycles); 16 group counters then one root counter cost ~(10 + 16) x 100.
            const unsigned int grp = blockIdx.x & 15u;
            const unsigned int gsize = (gridDim.x - grp + 15u) >> 4;             // CTAs with this residue
            const unsigned int ngroups = gridDim.x < 16u ? gridDim.x : 16u;
            int last = 0;
            if (p.dbg & 16) {
            } else if (atomicAdd(&e.gn_counters[1 + grp], 1u) == gsize - 1u) {
                e.gn_counters[1 + grp] = 0u;                                     // whole group has arrived: reset for the next launch
                __threadfence();
                last = atomicAdd(&e.gn_counters[0], 1u) == ngroups - 1u;
            }
            *s_last = last;
        }
        if (p.dbg_out && lane == 0 && warp < 8) p.dbg_out[blockIdx.x * 32 + 24 + warp] = (unsigned long long)(clock64() - td1);
        __syncthreads();
        const long long td2 = clock64();
        if (p.dbg_out && tid == 0) {
            p.dbg_out[blockIdx.x * 32 + 16] = (unsigned long long)(td1 - td0);
            p.dbg_out[blockIdx.x * 32 + 17] = (unsigned long long)(td2 - td1);
        }
        if (*s_last) {
            __threadfence();
            const long long td3 = clock64();
            const double inv_count = 1.0 / ((double)kGsz * (double)p.Hout * (double)p.Wout);
            const int G = (int)gridDim.x;
            const int q = lane & 3, r0 = lane >> 2;                  // float4 column q of partial rows r0, r0 + 8, ...
            for (int b = warp; b < p.B && warp < kThreads / 32; b += kThreads / 32) {
                const float4* pp = reinterpret_cast<const float4*>(e.gn_partials + (size_t)b * G * 16) + q;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int base = 0; base < G; base += 64) {
                    float4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = base + r0 + 8 * i;
                        v[i] = c < G ? __ldcg(pp + (size_t)c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
                        const int g = lane * 4 + j;
                        e.gn_stats[((size_t)b * 8 + g) * 2 + 0] = (float)mean;
                        e.gn_stats[((size_t)b * 8 + g) * 2 + 1] = (float)rsqrt(var + (double)e.gn_eps);
                    }
                }
            }
            if (tid == 0) e.gn_counters[0] = 0u;                     // ready for the next launch
            if (p.dbg_out) {
                __syncthreads();
                if (tid == 0) {
                    unsigned long long* o = p.dbg_out + blockIdx.x * 32;
                    o[18] = (unsigned long long)(td3 - td2); o[19] = (unsigned long long)(clock64() - td3);
                    o[20] = (unsigned long long)p.B;
                }
            }
        }
    }
}

}  // namespace tc

struct TcConvPlan {
    CUtensorMap mapA0, mapA1, mapW, mapWh;
    tc::TcParams p;
    int N, grid;
    size_t smem;
};
int conv_tc_halo_launch(const TcConvPlan* pl, cudaStream_t stream);      // conv_tc_halo.cu
int conv_tc_halo2_launch(const TcConvPlan* pl, cudaStream_t stream);     // conv_tc_halo2.cu
int conv_tc_halo2_max_grid(int N, bool residual, size_t smem, bool async_apply);   // co-resident CTAs of the GroupNorm-apply variants
int conv_tc_halo2_apply_extra_smem(int apply);

namespace tc {
// ------------------------------------------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline bool encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return false; }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return false;
    }
    return true;
}

}  // namespace tc
}  // namespace gtts
