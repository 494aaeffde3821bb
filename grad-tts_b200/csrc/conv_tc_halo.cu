// 3x3 stride-1 convolution on tcgen05 with ONE halo load per (tile, 64-channel chunk).
//
// conv_tc.cu fetches a separate TMA box for each of the 9 taps, so the same activations cross the L2 -> SMEM
// path nine times; measured on B200 that path (~14.5 TB/s chip-wide) is what bounds the 64- and 128-channel
// layers.  Here the tile is 16 rows x 8 pixels (M = 128) and the producer loads one 18 x 16-pixel box
// (64 channels, 128-byte swizzled, 36 KB).  Pixel (hh, ww) of the box sits in shared-memory row hh*16 + ww, so
// for tap (dy, dx) the A operand "row m = hl*8 + wl -> box row (hl+dy)*16 + (wl+dx)" is an ordinary K-major
// SWIZZLE_128B matrix whose 8-row groups are pitch*128 bytes apart (SBO) and whose start is shifted by
// (dy*pitch + dx) rows: nine descriptor views of one buffer.  Measured on B200: the tcgen05 swizzle XOR is a
// function of the absolute shared-memory address bits (descriptor base_offset = 0), so row-shifted starts and
// group strides that are not multiples of the 1024-byte pattern read exactly what TMA wrote
// (tests/test_gpu_kernels.py::test_halo_conv).  halo_mode 1: box 18 x 16 pixels; halo_mode 2: box 18 x 10.
// Weights are loaded once per CTA when all 9*Cin/64 tiles fit in shared memory (64->64, 64->128), otherwise they
// stream through their own ring.
#include <cstring>

#include "conv_tc_common.cuh"

namespace gtts {

using namespace tc;

namespace {

template <int N, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapW, const TcParams p) {
    constexpr int kBBytes = N * 128;
    constexpr uint32_t kIdesc = make_idesc<N>();

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const int pw = p.halo_mode == 2 ? 10 : 16;                     // box width in pixels = row pitch of the halo tile
    const int a_stage = p.a_bytes;                                  // 18 * pw * 128, rounded up to 1024
    uint8_t* smem_b = smem + (size_t)p.stages * a_stage;
    const TcShared sh = tc_shared(smem_b + (size_t)p.b_slots * kBBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapW);
    }
    const long long tk0 = clock64();
    unsigned long long gt0 = 0;
    if (p.dbg_out && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
    const uint32_t tmem_base = tc_prologue<N>(p, sh, p.stages, p.b_slots, tid, warp, lane);
    if (p.dbg_out && tid == 0) p.dbg_out[blockIdx.x * 32 + 10] = (unsigned long long)(clock64() - tk0);

    const int nck = p.nchunk0 + p.nchunk1;
    const int tiles_per_phase = p.tiles_h * p.tiles_w;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            bool first = true;
            const int G = (int)gridDim.x, nstage = p.stages, nslot = p.b_slots, resident = p.b_resident, dbg = p.dbg;
            const int nck0 = p.nchunk0, ht = p.halo_t;
            const uint32_t a_tx = (uint32_t)(18 * pw * 128);
            TileWalk tw, pf;
            long long c_pwait = 0;
            tw.init(p, (int)blockIdx.x, G);
            const int pf_dist = p.halo_prefetch;                     // tiles of look-ahead for the L2 prefetch (0 = off)
            if (pf_dist > 0) {
                pf.init(p, (int)blockIdx.x, G);
                for (int i = 0; i < pf_dist && pf.tile < p.num_tiles; ++i) {
                    pf.advance(G);
                    if (pf.tile < p.num_tiles)
                        for (int ck = 0; ck < nck; ++ck)
                            tma_prefetch_4d(ck < nck0 ? &mapA0 : &mapA1, (ck < nck0 ? ck : ck - nck0) * 64,
                                            ht ? pf.th * p.bh - 1 : pf.tw * p.bw - 1, ht ? pf.tw * p.bw - 1 : pf.th * p.bh - 1, pf.b);
                }
            }
            const int n_it = tc_num_iters(p);
            int it = 0;
            if (!resident) {
                // ---- streamed weights: `pass_tiles` A tiles share every weight slot (fill per tile: A + B / pass_tiles)
                const int TP = p.pass_tiles;
                for (; tw.tile < p.num_tiles; it += TP) {
                    if (kStats && (it >= n_it - 8)) prefetch_l2(p.e.gn_counters);
                    int tb[2], th0[2], tw0[2];
                    int nt = 0;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < TP && tw.tile < p.num_tiles) { tb[k] = tw.b; th0[k] = tw.th * p.bh; tw0[k] = tw.tw * p.bw; tw.advance(G); nt = k + 1; }
                    for (int ck = 0; ck < nck; ++ck) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (k >= nt) continue;
                            const long long tp0 = clock64();
                            mbar_wait(&sh.empty[sa], pha ^ 1u);
                            c_pwait += clock64() - tp0;
                            mbar_expect_tx(&sh.full[sa], a_tx);
                            tma_load_4d(ck < nck0 ? &mapA0 : &mapA1, &sh.full[sa], smem + (size_t)sa * a_stage,
                                        (ck < nck0 ? ck : ck - nck0) * 64, ht ? th0[k] - 1 : tw0[k] - 1,
                                        ht ? tw0[k] - 1 : th0[k] - 1, tb[k]);
                            if (++sa == nstage) { sa = 0; pha ^= 1u; }
                        }
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&sh.emptyb[sb], phb ^ 1u);
                            mbar_expect_tx(&sh.fullb[sb], (uint32_t)kBBytes);
                            tma_load_2d(&mapW, &sh.fullb[sb], smem_b + (size_t)sb * kBBytes, ck * 64, p.wrow[0][tap]);
                            if (++sb == nslot) { sb = 0; phb ^= 1u; }
                        }
                    }
                }
            } else
            for (; tw.tile < p.num_tiles; tw.advance(G), ++it) {
                const int b = tw.b, h0 = tw.th * p.bh, w0 = tw.tw * p.bw;
                // the tail's ticket counters were last touched a launch ago and have been evicted by this kernel's own
                // traffic; an L2 miss under full load costs ~8 us on the critical tail, so fetch the line ahead of time
                if (kStats && (it == n_it - 8 || it == n_it - 1)) prefetch_l2(p.e.gn_counters);
                if (pf_dist > 0 && pf.tile < p.num_tiles) {
                    pf.advance(G);
                    if (pf.tile < p.num_tiles)
                        for (int ck = 0; ck < nck; ++ck)
                            tma_prefetch_4d(ck < nck0 ? &mapA0 : &mapA1, (ck < nck0 ? ck : ck - nck0) * 64,
                                            ht ? pf.th * p.bh - 1 : pf.tw * p.bw - 1, ht ? pf.tw * p.bw - 1 : pf.th * p.bh - 1, pf.b);
                }
                for (int ck = 0; ck < nck; ++ck) {
                    const long long tp0 = clock64();
                    mbar_wait(&sh.empty[sa], pha ^ 1u);
                    c_pwait += clock64() - tp0;
                    if (dbg & 4) {
                        mbar_arrive(&sh.full[sa]);                   // experiment: no A traffic at all
                    } else {
                        mbar_expect_tx(&sh.full[sa], a_tx);
                        tma_load_4d(ck < nck0 ? &mapA0 : &mapA1, &sh.full[sa], smem + (size_t)sa * a_stage,
                                    (ck < nck0 ? ck : ck - nck0) * 64, ht ? h0 - 1 : w0 - 1, ht ? w0 - 1 : h0 - 1, b);
                    }
                    if (++sa == nstage) { sa = 0; pha ^= 1u; }
                    if (first) {
                        for (int tap = 0; tap < 9; ++tap) {
                            const int slot = ck * 9 + tap;
                            mbar_expect_tx(&sh.fullb[slot], (uint32_t)kBBytes);
                            tma_load_2d(&mapW, &sh.fullb[slot], smem_b + (size_t)slot * kBBytes, ck * 64, p.wrow[0][tap]);
                        }
                    }
                }
                first = false;
            }
            if (p.dbg_out) p.dbg_out[blockIdx.x * 32 + 6] = (unsigned long long)c_pwait;
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        // The whole warp runs the loop converged and `elect.sync` sits directly on the branch around the tcgen05
        // instructions: ptxas then knows exactly one thread is active and emits bare UTCHMMA instead of wrapping each one
        // in an ELECT/BRA.U.ANY "for each active lane" loop (which it must do under a data-dependent `if (leader)`).
        // (A second issuing warp on alternate tiles was tried: no gain -- with both feeding the pipe an N=64 MMA still
        // takes ~65 cycles in this kernel vs 52 in isolation: the tensor pipe's operand reads, 6 KB per MMA, share the
        // shared-memory port with the TMA fill of the next halo boxes.  The kernel is SMEM-bandwidth bound.)
        const int nstage = p.stages, nslot = p.b_slots, resident = p.b_resident, dbg = p.dbg;
        const uint64_t a_desc0 = make_sw128_kmajor_desc(smem_u32(smem), (uint32_t)(pw * 128), 0u);
        const uint64_t b_desc0 = make_sw128_kmajor_desc(smem_u32(smem_b));
        const uint64_t a_stage_step = (uint64_t)(a_stage >> 4), b_slot_step = (uint64_t)(kBBytes >> 4);
        uint64_t tap_off[9];
#pragma unroll
        for (int t = 0; t < 9; ++t)                                  // box row of tap (dy, dx): dy*pw + dx, transposed: dx*pw + dy
            tap_off[t] = (uint64_t)(((p.halo_t ? (t % 3) * pw + (t / 3) : (t / 3) * pw + (t % 3)) * 128) >> 4);
        int sa = 0, sb = 0, it = 0;
        uint32_t pha = 0, phb = 0;
        const int n_it = tc_num_iters(p);
        long long c_tempty = 0, c_full = 0, c_issue = 0, c_commit = 0, c_n = 0;
        const long long tl0 = clock64();
        if (!resident) {
            // ---- streamed weights: every weight slot (one tap of one 64-channel chunk) feeds `pass_tiles` accumulators
            // before it is released, so a tile costs A + B / pass_tiles bytes of L2 -> SMEM fill.  (With one tile per
            // pass the 128->128, 256->256, 512->128 layers stream 0.3 - 1.2 MB of weights per 128-pixel tile and are
            // fill-bound at ~45 B/clk/SM; with two they are tensor/SMEM-bound.)
            const int TP = p.pass_tiles;
            constexpr int kBufs = acc_bufs<N>();
            for (it = 0; it < n_it; it += TP) {
                const int nt = n_it - it < TP ? n_it - it : TP;
                uint32_t d_tmem[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k >= nt) continue;
                    const int buf = (it + k) % kBufs;
                    mbar_wait(&sh.tempty[buf], ((uint32_t)((it + k) / kBufs) & 1u) ^ 1u);
                    d_tmem[k] = tmem_base + (uint32_t)(buf * N);
                }
                tc_fence_after();
                for (int ck = 0; ck < nck; ++ck) {
                    uint64_t adesc[2];
                    int sa_k[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (k >= nt) continue;
                        mbar_wait(&sh.full[sa], pha);
                        sa_k[k] = sa;
                        adesc[k] = a_desc0 + (uint64_t)sa * a_stage_step;
                        if (++sa == nstage) { sa = 0; pha ^= 1u; }
                    }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&sh.fullb[sb], phb);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t bdesc = b_desc0 + (uint64_t)sb * b_slot_step;
                            if (!(dbg & 1)) {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        if (k < nt) tc_mma_f16(d_tmem[k], adesc[k] + tap_off[tap] + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk),
                                                   kIdesc, (uint32_t)((ck | tap | kk) != 0));
                            }
                            tc_commit(&sh.emptyb[sb]);
                            if (tap == 8) {
#pragma unroll
                                for (int k = 0; k < 2; ++k) if (k < nt) tc_commit(&sh.empty[sa_k[k]]);
                                if (ck == nck - 1) {
#pragma unroll
                                    for (int k = 0; k < 2; ++k) if (k < nt) tc_commit(&sh.tfull[(it + k) % kBufs]);
                                }
                            }
                        }
                        __syncwarp();
                        if (++sb == nslot) { sb = 0; phb ^= 1u; }
                    }
                }
            }
        } else
        for (it = 0; it < n_it; ++it) {
            const int buf = it % acc_bufs<N>();
            const long long t0 = clock64();
            mbar_wait(&sh.tempty[buf], ((uint32_t)(it / acc_bufs<N>()) & 1u) ^ 1u);
            tc_fence_after();
            const long long t1 = clock64();
            c_tempty += t1 - t0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
            for (int ck = 0; ck < nck; ++ck) {
                const long long t2 = clock64();
                mbar_wait(&sh.full[sa], pha);
                const long long t2b = clock64();
                c_full += t2b - t2;
                const uint64_t adesc = a_desc0 + (uint64_t)sa * a_stage_step;
                if (it == 0) {
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) mbar_wait(&sh.fullb[ck * 9 + tap], 0u);
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bdesc = b_desc0 + (uint64_t)(ck * 9) * b_slot_step;
                    if (!(dbg & 1)) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_f16(d_tmem, adesc + tap_off[tap] + (uint64_t)(2 * k),
                                           bdesc + (uint64_t)tap * b_slot_step + (uint64_t)(2 * k), kIdesc,
                                           (uint32_t)((ck | tap | k) != 0));
                    }
                    const long long t3 = clock64();
                    tc_commit(&sh.empty[sa]);
                    if (ck == nck - 1) tc_commit(&sh.tfull[buf]);
                    c_issue += t3 - t2b;
                    c_commit += clock64() - t3;
                    c_n += 1;
                }
                __syncwarp();
                if (++sa == nstage) { sa = 0; pha ^= 1u; }
            }
        }
        if (p.dbg_out && lane == 0) {                    // lane 0 is the elected lane of a converged warp
            unsigned long long* o = p.dbg_out + blockIdx.x * 32;
            o[0] = (unsigned long long)c_tempty; o[1] = (unsigned long long)c_full; o[2] = (unsigned long long)c_issue;
            unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            o[3] = (unsigned long long)c_commit; o[4] = (unsigned long long)c_n | ((unsigned long long)smid << 32); o[5] = (unsigned long long)(clock64() - tl0);
        }
    } else if (warp == 3) {
        tc_stats_loop<kStats>(p, sh, lane);
    } else if (warp >= 4) {
        tc_epilogue_loop<N, kStats, false, false>(p, sh, tmem_base, warp, lane);
    }
    if (p.dbg_out && tid == 0) p.dbg_out[blockIdx.x * 32 + 11] = (unsigned long long)(clock64() - tk0);
    tc_teardown<N, kStats>(p, sh, smem, tmem_base, tid, warp, lane);
    if (p.dbg_out && tid == 0) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        p.dbg_out[blockIdx.x * 32 + 12] = (unsigned long long)(clock64() - tk0);
        p.dbg_out[blockIdx.x * 32 + 15] = gt1;
        p.dbg_out[blockIdx.x * 32 + 7] = gt0;
    }
}

template <int N, bool kStats>
int launch_halo(const TcConvPlan* pl, cudaStream_t stream) {
    static bool attr_set = false;
    auto k = conv_tc_halo_kernel<N, kStats>;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GTTS_CHECK_CUDA(launch_pdl(k, dim3(pl->grid), dim3(kThreads), pl->smem, stream, 1, pl->mapA0, pl->mapA1, pl->mapW, pl->p));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int conv_tc_halo_launch(const TcConvPlan* pl, cudaStream_t stream) {
    const ConvEpilogue& e = pl->p.e;
    GTTS_REQUIRE(e.residual == nullptr && e.mask == nullptr, "conv_tc_halo: plain or GN-statistics epilogue only");
    const bool st = e.gn_partials != nullptr;
    if (pl->N == 64) return st ? launch_halo<64, true>(pl, stream) : launch_halo<64, false>(pl, stream);
    if (pl->N == 128) return st ? launch_halo<128, true>(pl, stream) : launch_halo<128, false>(pl, stream);
    if (pl->N == 256) return st ? launch_halo<256, true>(pl, stream) : launch_halo<256, false>(pl, stream);
    set_error("conv_tc_halo: unsupported Cout");
    return 2;
}

}  // namespace gtts
