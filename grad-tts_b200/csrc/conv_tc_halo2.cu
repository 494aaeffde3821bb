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
template <int N, bool kStats, bool kFuse, bool kMask = false, bool kConvT = false>
__global__ void __launch_bounds__(kThreads + (kFuse ? kFuseWarps * 32 : 0), 1)
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
    pdl_wait();
    const uint32_t tmem_base = *sh.tmem_slot;

    const int nck = p.nchunk0 + p.nchunk1;
    const int n_it = tc_num_iters(p);

    if (warp == 0) {
        // ================================================================ TMA producer (both CTAs)
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            const int G = (int)gridDim.x, nck0 = p.nchunk0, ht = p.halo_t;
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
            for (int it = 0; it < n_it; ++it, tw.advance(G)) {
                if (kStats && (it == n_it - 8 || it == n_it - 1)) prefetch_l2(p.e.gn_counters);
                const int b = tw.b, h0 = tw.th * p.bh, w0 = tw.tw * p.bw;
                for (int ck = 0; ck < nck; ++ck) {
                    mbar_wait(&sh.empty[sa], pha ^ 1u);
                    if (kFuse) {
                        mbar_expect_tx(&rawfull[sa], a_tx);          // my own tile only; the transform warps pass it on
                        tma_load_4d(ck < nck0 ? &mapA0 : &mapA1, &rawfull[sa], smem + (size_t)sa * a_stage,
                                    (ck < nck0 ? ck : ck - nck0) * 64, ht ? h0 - 1 : w0 - 1, ht ? w0 - 1 : h0 - 1, b);
                    } else {
                        if (rank == 0) mbar_expect_tx(&sh.full[sa], 2u * a_tx);
                        tma_load_4d_2sm(ck < nck0 ? &mapA0 : &mapA1, mapa_u32(smem_u32(&sh.full[sa]), 0u), smem + (size_t)sa * a_stage,
                                        (ck < nck0 ? ck : ck - nck0) * 64, ht ? h0 - 1 : w0 - 1, ht ? w0 - 1 : h0 - 1, b);
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
        tc_stats_loop<kStats>(p, sh, lane);
    } else if (warp >= 4 && warp < kThreads / 32) {
        tc_epilogue_loop<N, kStats, false, kMask>(p, sh, tmem_base, warp, lane);
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
                        const float h0f = __ldg(e.in_beta + c0 + 2 * q) - mean * s0, h1f = __ldg(e.in_beta + c0 + 2 * q + 1) - mean * s1;
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
    tc_teardown<N, kStats, true>(p, sh, smem, tmem_base, tid, warp, lane);
}

template <int N, bool kStats, bool kFuse, bool kMask = false, bool kConvT = false>
int launch_halo2(const TcConvPlan* pl, cudaStream_t stream) {
    static bool attr_set = false;
    auto k = conv_tc_halo2_kernel<N, kStats, kFuse, kMask, kConvT>;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GTTS_CHECK_CUDA(launch_pdl(k, dim3(pl->grid), dim3(kThreads + (kFuse ? kFuseWarps * 32 : 0)), pl->smem, stream, 2, pl->mapA0,
                               pl->mapA1, pl->mapWh, pl->p));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

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
    GTTS_REQUIRE(e.residual == nullptr && e.mask == nullptr, "conv_tc_halo2: plain or GN-statistics epilogue only");
    const bool st = e.gn_partials != nullptr;
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
