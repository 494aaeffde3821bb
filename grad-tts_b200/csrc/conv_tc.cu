// K1/K3/K4/K5 -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a only).
//
//   D[128 pixels, N = Cout] += A[128 pixels, 64 ch of one tap] * W[N, 64 ch]^T      (bf16 x bf16 -> fp32)
//
// * A tiles are fetched by TMA straight out of the NHWC activation tensor: one 4-D box
//   (64 ch, bw, bh, 1) per (tap, 64-channel chunk), the tap shift applied to the box coordinates and the
//   zero padding supplied by TMA out-of-bounds fill.  Stride-2 convs use a 5-D view of the same tensor
//   (2C, W/2, 2, H/2, B) so that a unit-stride box picks every second pixel; transposed convs run as four
//   output-parity phases of 2x2 taps.  Two sources (the U-Net skip `cat`) are two tensor maps walked in
//   the K loop, so the concatenation is never materialised.
// * W tiles (N rows x 64 ch, K-major) come from the pre-packed weight matrix by a 2-D TMA box; per-sample
//   weights (the folded linear-attention matrix) just offset the row coordinate by b*Cout.
// * Both land in shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma consumes through
//   shared-memory descriptors; accumulators live in TMEM (two buffers of N columns, so the epilogue of
//   tile i overlaps the MMAs of tile i+1).
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue
//   (tcgen05.ld -> +bias -> GroupNorm partial statistics -> (+residual, *mask) -> bf16 NHWC store).
// * Persistent CTAs (one per SM) walk the tile list round-robin.
//
// Reference ops covered: Conv2d 3x3 (diffusion.py:52), 1x1 (:70,87,88), 3x3 s2 (:33), ConvTranspose2d 4x4 s2 (:24).
#include <cstdlib>
#include <cstring>

#include "conv_tc_common.cuh"

namespace gtts {

using namespace tc;

namespace {

// kStats: GroupNorm partial statistics of (acc + bias); kRes: + residual; kMask: * mask[b][w]
template <int N, bool kStats, bool kRes, bool kMask, bool kOutF32 = false, bool kAct = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapWh, const TcParams p) {
    constexpr int kBBytes = N * 128;
    constexpr int kStage = kABytes + kBBytes;
    constexpr uint32_t kIdesc = make_idesc<N>();

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // Weights either travel with every A tile (stage = A | B) or, when all nphase*ntaps*nchunks tiles fit next to a few A
    // stages, are loaded once per CTA into their own region (1x1, stride-2 and transposed 64-channel convs: the kernel is
    // L2->SMEM-fill bound and a weight tile is a third to two thirds of every stage).
    const int resident = p.b_resident;
    const int stage_bytes = p.halo1d ? p.a_stage : (resident ? kABytes : kStage);
    uint8_t* smem_b = smem + (size_t)p.stages * stage_bytes;        // resident weights (b_slots tiles)
    const TcShared sh = tc_shared(smem_b + (size_t)((resident || p.halo1d) ? p.b_slots : 0) * kBBytes);   // (1-D halo mode: B ring)
    uint64_t* full = sh.full;
    uint64_t* empty = sh.empty;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapWh);
    }
    const uint32_t tmem_base = tc_prologue<N>(p, sh, p.stages, (p.halo1d && !p.b_resident) ? p.b_slots : 1, tid, warp, lane, true);
    const int n_it = tc_num_iters(p);
    const int G = (int)gridDim.x;

    const int nkb = p.ntaps * (p.nchunk0 + p.nchunk1);
    const int tiles_per_phase = p.tiles_h * p.tiles_w;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            int stage = 0, sb_ring = 0;
            uint32_t phase = 0, phb_ring = 0;
            const uint32_t rank = p.mc ? cluster_ctarank() : 0u;
            if (resident && n_it > 0) {
                const int nck = p.nchunk0 + p.nchunk1;
                mbar_expect_tx(&sh.fullb[0], (uint32_t)(p.b_slots * kBBytes));
                for (int ph = 0; ph < p.nphase; ++ph)
                    for (int tap = 0; tap < p.ntaps; ++tap)
                        for (int ck = 0; ck < nck; ++ck)
                            tma_load_2d(&mapW, &sh.fullb[0], smem_b + (size_t)((ph * p.ntaps + tap) * nck + ck) * kBBytes, ck * 64,
                                        p.wrow[ph][tap]);
            }
            pdl_wait();                                              // (see tc_prologue) activations of the predecessors from here on
            for (int it = 0; it < n_it; ++it) {
                if (kStats && (it == n_it - 8 || it == n_it - 1)) prefetch_l2(p.e.gn_counters);   // see conv_tc_halo.cu
                const int tile = (int)blockIdx.x + it * G;
                const bool dummy = tile >= p.num_tiles;              // only with multicast: keeps the pair in lockstep
                const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h;
                const int ph = (tile / tiles_per_phase) % p.nphase, b = tile / (tiles_per_phase * p.nphase);
                const int h0 = th * p.bh, w0 = tw * p.bw;
                if (p.halo1d) {
                    // one box of bw + 2*halo positions per chunk; every tap reads it through a shifted descriptor.  Weights that do
                    // not fit shared memory stream through their own ring (one slot per (tap, chunk)), so a tile costs one A box
                    // per chunk instead of one per (tap, chunk)
                    for (int ck = 0; ck < p.nchunk0 + p.nchunk1; ++ck) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full[stage], (uint32_t)(dummy ? 0 : p.a_bytes));
                        if (!dummy) {
                            int which, chan;
                            tc_chunk_src(p, ck, &which, &chan);
                            tma_load_4d(which ? &mapA1 : &mapA0, &full[stage], sa, chan, w0 - p.halo1d, h0, b);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        if (!resident) {
                            for (int tap = 0; tap < p.ntaps; ++tap) {
                                mbar_wait(&sh.emptyb[sb_ring], phb_ring ^ 1u);
                                mbar_expect_tx(&sh.fullb[sb_ring], (uint32_t)kBBytes);
                                tma_load_2d(&mapW, &sh.fullb[sb_ring], smem_b + (size_t)sb_ring * kBBytes, ck * 64, p.wrow[0][tap]);
                                if (++sb_ring == p.b_slots) { sb_ring = 0; phb_ring ^= 1u; }
                            }
                        }
                    }
                    continue;
                }
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int dy = p.dy[ph][tap], dx = p.dx[ph][tap];
                    const int wr = p.wrow[ph][tap] + b * p.w_batch_rows;
                    for (int ck = 0; ck < p.nchunk0 + p.nchunk1; ++ck) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full[stage], (uint32_t)((dummy ? 0 : p.a_bytes) + (resident ? 0 : kBBytes)));
                        if (dummy) {
                        } else if (p.stride2) {
                            // 5-D view (2C, W/2, 2, H/2, B): input pixel 2*o + d = pair (o + floor(d/2)), parity d & 1; d in {-1,0,1} for
                            // the 3x3 Downsample, {-1,0,1,2} for the 4x4 conv that is the Upsample's data gradient
                            const int px = dx & 1, py = dy & 1;
                            int which, chan;
                            tc_chunk_src(p, ck, &which, &chan);
                            tma_load_5d(&mapA0, &full[stage], sa, px * p.Cin0 + chan, w0 + (dx >> 1), py, h0 + (dy >> 1), b);
                        } else {
                            int which, chan;
                            tc_chunk_src(p, ck, &which, &chan);
                            tma_load_4d(which ? &mapA1 : &mapA0, &full[stage], sa, chan, w0 + dx, h0 + dy, b);
                        }
                        if (resident) {
                        } else if (p.mc) {
                            // my half of the weight tile goes to both CTAs of the pair (same smem offset, same barrier)
                            // (multicast is only used with one phase and shared weights: the row is p.wrow[0][tap])
                            tma_load_2d_mc(&mapWh, &full[stage], sa + kABytes + rank * (kBBytes / 2), ck * 64,
                                           p.wrow[0][tap] + (int)rank * (N / 2), (uint16_t)3);
                        } else {
                            tma_load_2d(&mapW, &full[stage], sa + kABytes, ck * 64, wr);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        // Converged warp, elect.sync directly on the tcgen05 branch (bare UTCHMMA in SASS; see conv_tc_halo.cu).
        const int nstage = p.stages;
        const uint64_t a_desc0 = make_sw128_kmajor_desc(smem_u32(smem));
        const uint64_t b_desc0 = make_sw128_kmajor_desc(smem_u32(smem) + kABytes);
        const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
        const uint64_t b_res0 = make_sw128_kmajor_desc(smem_u32(smem_b));
        int stage = 0, it = 0, sb_ring = 0;
        uint32_t phase = 0, phb_ring = 0;
        const int mc = p.mc;
        if (resident && n_it > 0) mbar_wait(&sh.fullb[0], 0u);
        for (it = 0; it < n_it; ++it) {
            const int buf = it % acc_bufs<N>();
            const int ph_tile = resident ? (((int)blockIdx.x + it * G) / tiles_per_phase) % p.nphase : 0;
            mbar_wait(&sh.tempty[buf], ((uint32_t)(it / acc_bufs<N>()) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
            if (p.halo1d && resident) {
                const int nck = p.nchunk0 + p.nchunk1;
                for (int ck = 0; ck < nck; ++ck) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc0 = a_desc0 + (uint64_t)stage * stage_step;
                        for (int tap = 0; tap < p.ntaps; ++tap) {
                            // rows tap_row .. tap_row + 127 of the box: the swizzle is a function of the absolute shared-memory
                            // address, so a start address shifted by whole 128-byte rows reads what TMA wrote (conv_tc_halo.cu)
                            const uint64_t adesc = adesc0 + (uint64_t)((p.tap_row[tap] * 128) >> 4);
                            const uint64_t bdesc = b_res0 + (uint64_t)(tap * nck + ck) * (uint64_t)(kBBytes >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                           (uint32_t)((ck | tap | k) != 0));
                        }
                        tc_commit(&empty[stage]);
                        if (ck == nck - 1) tc_commit(&sh.tfull[buf]);
                    }
                    __syncwarp();
                    if (++stage == nstage) { stage = 0; phase ^= 1u; }
                }
                continue;
            }
            if (p.halo1d) {                                          // streamed weights: one ring slot per (tap, chunk)
                const int nck = p.nchunk0 + p.nchunk1;
                for (int ck = 0; ck < nck; ++ck) {
                    mbar_wait(&full[stage], phase);
                    const uint64_t adesc0 = a_desc0 + (uint64_t)stage * stage_step;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        mbar_wait(&sh.fullb[sb_ring], phb_ring);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t adesc = adesc0 + (uint64_t)((p.tap_row[tap] * 128) >> 4);
                            const uint64_t bdesc = b_res0 + (uint64_t)sb_ring * (uint64_t)(kBBytes >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                           (uint32_t)((ck | tap | k) != 0));
                            tc_commit(&sh.emptyb[sb_ring]);
                            if (tap == p.ntaps - 1) {
                                tc_commit(&empty[stage]);
                                if (ck == nck - 1) tc_commit(&sh.tfull[buf]);
                            }
                        }
                        __syncwarp();
                        if (++sb_ring == p.b_slots) { sb_ring = 0; phb_ring ^= 1u; }
                    }
                    if (++stage == nstage) { stage = 0; phase ^= 1u; }
                }
                continue;
            }
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = a_desc0 + (uint64_t)stage * stage_step;
                    const uint64_t bdesc = resident ? b_res0 + (uint64_t)(ph_tile * nkb + kb) * (uint64_t)(kBBytes >> 4)
                                                    : b_desc0 + (uint64_t)stage * stage_step;
                    if (!(p.dbg & 1)) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)                 // 4 x (K = 16 bf16 = 32 bytes)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                       (uint32_t)((kb | k) != 0));
                    }
                    if (mc) tc_commit_mc(&empty[stage], (uint16_t)3);   // frees the slot in both CTAs of the pair
                    else tc_commit(&empty[stage]);                  // smem slot free when these MMAs retire
                    if (kb == nkb - 1) tc_commit(&sh.tfull[buf]);   // accumulator complete
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 3) {
        tc_stats_loop<kStats>(p, sh, lane);
    } else if (warp >= 4) {
        tc_epilogue_loop<N, kStats, kRes, kMask, kOutF32, kAct>(p, sh, tmem_base, warp, lane);
    }
    tc_teardown<N, kStats>(p, sh, smem, tmem_base, tid, warp, lane);
}

void pick_tile(int Hg, int Wg, int* bh_out, int* bw_out) {
    long best = -1;
    int bbh = 1, bbw = 128;
    for (int bh = 1; bh <= 128; ++bh) {
        int bw = 128 / bh;
        if (bw < 1) break;
        if (bh > Hg && bh != 1) continue;
        if (bw > 256) bw = 256;
        long tiles = (long)((Hg + bh - 1) / bh) * ((Wg + bw - 1) / bw);
        // fewer tiles first; then wider rows (longer contiguous TMA runs)
        long score = tiles * 1024 - bw;
        if (best < 0 || score < best) { best = score; bbh = bh; bbw = bw; }
    }
    *bh_out = bbh;
    *bw_out = bbw;
}

}  // namespace

size_t conv_tc_partials_slots(const ConvGeom&) { return 256; }         // one partial per (sample, CTA); grid <= #SMs

// CTA-pair (cta_group::2) halo kernel, conv_tc_halo2.cu.  GTTS_CTA2=0 falls back to the single-CTA kernels.
bool conv_tc_cta2_enabled() {
    const char* e = getenv("GTTS_CTA2");
    return e ? atoi(e) != 0 : true;
}

bool conv_tc_halo_eligible(const ConvGeom& g) {
    // Cout = 256 also runs (tests cover it; GTTS_HALO256=1) but loses to the per-tap kernel: two 256-column accumulators
    // fill TMEM, so the epilogue no longer overlaps the next tile's MMAs, and the fixed 128-pixel halo tile wastes rows
    // at H = 20.
    const bool halo256 = getenv("GTTS_HALO256") != nullptr || conv_tc_cta2_enabled();
    return g.ntaps == 9 && g.stride == 1 && g.nphase == 1 && g.w_batch_rows == 0 &&
           (g.Cout == 64 || g.Cout == 128 || (g.Cout == 256 && halo256)) &&
           g.Hg >= 16 && g.Wg >= 8;
}

// Transposed 4x4 stride-2 conv on the CTA-pair halo kernel (conv_tc_halo2.cu, kConvT): 64 / 128 channels, mask epilogue.
bool conv_tc_convT_halo_eligible(const ConvGeom& g) {
    if (getenv("GTTS_CONVT_HALO") && atoi(getenv("GTTS_CONVT_HALO")) == 0) return false;
    return conv_tc_cta2_enabled() && g.nphase == 4 && g.ntaps == 4 && g.out_step == 2 && g.stride == 1 && g.w_batch_rows == 0 &&
           g.Cin1 == 0 && g.Cin0 == g.Cout && (g.Cout == 64 || g.Cout == 128) && g.Hg >= 16 && g.Wg >= 18 &&
           (long)g.B * ((g.Hg + 7) / 8) * ((g.Wg + 15) / 16) >= 2;
}

size_t conv_tc_halo_partials_slots(const ConvGeom&) { return 256; }    // one partial per (sample, CTA); grid <= #SMs

// [0] root ticket, [1..16] group tickets (tc_teardown), [32 + 2b] / [33 + 2b] arrivals / departures of sample b at the per-sample
// grid barrier of the GroupNorm-apply epilogue (conv_tc_halo2.cu)
size_t conv_tc_counter_words(int B) { return (size_t)32 + 2 * (size_t)B + 32; }

namespace {
// halo tiling of the CTA-pair kernel (must match conv_tc_plan_create)
void halo2_tiles(const ConvGeom& g, int* tiles_h, int* tiles_w) {
    int bh = 16, bw = 8;
    if (g.Hg >= 10 && g.Wg >= 18 && !getenv("GTTS_HALO_NO_T")) {
        const long t0 = (long)((g.Hg + 15) / 16) * ((g.Wg + 7) / 8), t1 = (long)((g.Hg + 7) / 8) * ((g.Wg + 15) / 16);
        if (t1 < t0) { bh = 8; bw = 16; }
    }
    *tiles_h = (g.Hg + bh - 1) / bh;
    *tiles_w = (g.Wg + bw - 1) / bw;
}
// A sample's tiles must fit the TMEM accumulator ring of every CTA pair: a pair works on tiles 2q, 2q+1 of each walk step, so
// when a sample can start on an odd tile index (odd tile count per sample) the two CTAs' runs of that sample are offset by one step.
bool apply_run_fits(int tps, int grid, int N) {
    const int bufs = N <= 64 ? 8 : (N <= 128 ? 4 : 2);
    const int run = (tps + grid - 1) / grid + ((tps & 1) ? 1 : 0);
    return run <= bufs;
}
int apply_grid(long num_tiles, int num_sms) {
    int gmax = num_sms & ~1;
    if (gmax > 256) gmax = 256;
    const long need = (num_tiles + 1) & ~1L;
    return (int)(need < gmax ? need : gmax);
}
}  // namespace

bool conv_tc_apply_eligible(const ConvGeom& g, int num_sms) {
    if (const char* e = getenv("GTTS_APPLY")) { if (atoi(e) == 0) return false; }
    if (!conv_tc_cta2_enabled() || !conv_tc_halo_eligible(g) || num_sms < 2) return false;
    int th, tw;
    halo2_tiles(g, &th, &tw);
    const int tps = th * tw;
    const long num_tiles = (long)g.B * tps;
    if (num_tiles < 2) return false;
    int min_h = 0;                                                   // GTTS_APPLY_MAXH: only levels with H <= this (experiments)
    if (const char* e = getenv("GTTS_APPLY_MAXH")) min_h = atoi(e);
    if (min_h && g.Hg > min_h) return false;
    // 16 CTAs of margin: the grid is clamped to the clusters that can be co-resident (conv_tc_halo2_max_grid), known only on the device
    const int grid = apply_grid(num_tiles, num_sms);
    return apply_run_fits(tps, grid > 32 ? grid - 16 : grid, g.Cout);
}

bool conv_tc_apply_async_eligible(const ConvGeom& g, int num_sms) {
    if (const char* e = getenv("GTTS_APPLY_ASYNC")) { if (atoi(e) == 0) return false; }
    if (!conv_tc_cta2_enabled() || !conv_tc_halo_eligible(g) || num_sms < 2 || g.split) return false;
    int th, tw;
    halo2_tiles(g, &th, &tw);
    if (const char* e = getenv("GTTS_APPLY_ASYNC_MAXH")) { if (g.Hg > atoi(e)) return false; }
    if (const char* e = getenv("GTTS_APPLY_ASYNC_MINH")) { if (g.Hg < atoi(e)) return false; }
    return (long)g.B * th * tw >= 2;
}

TcConvPlan* conv_tc_plan_create(const ConvGeom& g, const void* src0, const void* src1, const void* weight,
                                int weight_rows, const ConvEpilogue& e, int num_sms, int halo_mode) {
    if (!(g.Cout == 64 || g.Cout == 128 || g.Cout == 256)) { set_error("conv_tc: Cout must be 64/128/256"); return nullptr; }
    if (g.Cin0 % 64 || g.Cin1 % 64 || g.Cin0 <= 0) { set_error("conv_tc: Cin must be a multiple of 64"); return nullptr; }
    if (g.stride == 2 && (g.Cin1 != 0 || (g.Hin & 1) || (g.Win & 1))) { set_error("conv_tc: bad stride-2 geometry"); return nullptr; }
    if (e.gn_partials && !e.apply && (g.nphase != 1 || e.residual || e.mask)) { set_error("conv_tc: GN statistics only on plain convs"); return nullptr; }
    if (e.apply && !(halo_mode == 2 && conv_tc_halo_eligible(g) && conv_tc_cta2_enabled() && e.gn_partials && e.mask)) {
        set_error("conv_tc: the GroupNorm-apply epilogue needs the CTA-pair halo kernel, statistics buffers and a mask");
        return nullptr;
    }
    TcConvPlan* pl = new TcConvPlan();
    memset(pl, 0, sizeof(*pl));
    TcParams& p = pl->p;
    const bool convT_halo = halo_mode == 2 && conv_tc_convT_halo_eligible(g) && e.mask && !e.residual && !e.gn_partials &&
                            num_sms >= 2 && !g.split;    // split convs: the ConvT halo variant keeps 2*nck A stages resident -- too many chunks
    if (g.split && (e.apply || e.in_stats)) { set_error("conv_tc: split (fp32) convs take the plain epilogues only"); return nullptr; }
    if ((e.act_out || e.out2) && (e.out_f32 || e.mask || e.gn_partials || e.apply || e.in_stats || halo_mode && conv_tc_halo_eligible(g))) {
        set_error("conv_tc: leaky-ReLU outputs are for the per-tap kernel with bf16 activations only");
        return nullptr;
    }
    if (e.act_out && e.out2) { set_error("conv_tc: act_out and out2 are exclusive"); return nullptr; }
    if (g.split && 6 * (g.Cin0 + g.Cin1) / 64 > 48) { set_error("conv_tc: split conv with more than 48 K chunks"); return nullptr; }
    if (halo_mode && !convT_halo && !e.apply && (!conv_tc_halo_eligible(g) || e.residual || e.mask)) halo_mode = 0;
    // split convs know the CTA-pair halo kernel and the per-tap kernel only (the single-CTA halo kernel has no chunk table)
    if (g.split && halo_mode && (halo_mode != 2 || !conv_tc_cta2_enabled() || num_sms < 2 ||
                                 (long)g.B * ((g.Hg + 15) / 16) * ((g.Wg + 7) / 8) < 2)) halo_mode = 0;
    p.halo_mode = halo_mode;
    if (halo_mode) {
        // 128-pixel halo tile: 16 rows x 8 pixels (8-row UMMA groups run along W), or transposed 8 rows x 16 pixels with
        // H as the fast box dimension (groups run along H) when that covers the image with fewer tiles (H = 40: 270 vs
        // 324 tiles per sample, H = 20: 81 vs 108).
        p.bh = 16; p.bw = 8; p.halo_t = 0;
        if (halo_mode == 2 && g.Hg >= 10 && g.Wg >= 18 && !getenv("GTTS_HALO_NO_T")) {
            const long t0 = (long)((g.Hg + 15) / 16) * ((g.Wg + 7) / 8), t1 = (long)((g.Hg + 7) / 8) * ((g.Wg + 15) / 16);
            if (t1 < t0) { p.halo_t = 1; p.bh = 8; p.bw = 16; }
        }
    }
    else pick_tile(g.Hg, g.Wg, &p.bh, &p.bw);
    p.tiles_h = (g.Hg + p.bh - 1) / p.bh;
    p.tiles_w = (g.Wg + p.bw - 1) / p.bw;
    p.nphase = convT_halo ? 1 : g.nphase; p.B = g.B;               // ConvT on the halo kernel walks spatial tiles, 4 phases each
    p.ph_inner = convT_halo ? 4 : 0;
    p.Hg = g.Hg; p.Wg = g.Wg; p.Hout = g.Hout; p.Wout = g.Wout; p.out_step = g.out_step;
    p.ntaps = g.ntaps; p.nchunk0 = g.Cin0 / 64; p.nchunk1 = g.Cin1 / 64; p.Cin0 = g.Cin0;
    if (g.split) {
        // K = 6 * Cin: terms (plane of x, part of w) = (h,l) (m,m) (h,m) (l,h) (m,h) (h,h); within a term source 0 then source 1.
        // Smallest terms first: the tensor core's fp32 accumulation truncates (measured: error grows linearly with the number of
        // accumulation steps taken while the accumulator is large, 3e-5 relative at K = 2304 with the big term first), so the
        // five correction terms are summed while the accumulator is still ~2^-8 of its final size and only the K/16 steps of the
        // (h,h) term run at full magnitude.
        const int plane_of_term[6] = {0, 1, 0, 2, 1, 0};
        int ck = 0;
        for (int t = 0; t < 6; ++t) {
            for (int c = 0; c < g.Cin0 / 64; ++c, ++ck) { p.a_map[ck] = 0; p.a_off[ck] = (int16_t)(plane_of_term[t] * g.Cin0 + c * 64); }
            for (int c = 0; c < g.Cin1 / 64; ++c, ++ck) { p.a_map[ck] = 1; p.a_off[ck] = (int16_t)(plane_of_term[t] * g.Cin1 + c * 64); }
        }
        p.split = 1; p.nchunk0 = ck; p.nchunk1 = 0;
        p.Cin0 = 3 * g.Cin0;                           // channels per pixel of the plane tensor (stride-2 5-D view)
    }
    p.stride2 = (g.stride == 2); p.w_batch_rows = g.w_batch_rows;
    p.num_tiles = g.B * p.nphase * p.tiles_h * p.tiles_w;
    p.a_bytes = p.bh * p.bw * 128;
    memcpy(p.dy, g.dy, sizeof(p.dy)); memcpy(p.dx, g.dx, sizeof(p.dx));
    memcpy(p.wrow, g.wrow, sizeof(p.wrow)); memcpy(p.oy, g.oy, sizeof(p.oy)); memcpy(p.ox, g.ox, sizeof(p.ox));
    p.e = e;
    { const char* dbg = getenv("GTTS_CONV_DBG"); p.dbg = dbg ? atoi(dbg) : 0; }
    pl->N = g.Cout;
    const int apply_extra = e.apply ? conv_tc_halo2_apply_extra_smem(e.apply) : 0;
    const int budget = 227 * 1024 - kMiscBytes - 1024 - apply_extra;
    if (halo_mode) {
        // A ring: halo boxes of 18 x 16 pixels x 64 ch (36 KB); B: resident (all 9*nck tiles) if it fits, else a ring
        const int nck = p.nchunk0 + p.nchunk1, btile = g.Cout * 128, ntiles_b = (convT_halo ? 16 : 9) * nck;
        const int pw = halo_mode == 2 ? 10 : 16;
        const int abytes = (18 * pw * 128 + 1023) / 1024 * 1024;     // stage stride keeps every stage 1024-aligned
        p.a_bytes = abytes;
        p.pass_tiles = 1;
        { const char* pf = getenv("GTTS_HALO_PREFETCH"); p.halo_prefetch = pf ? atoi(pf) : 0; }
        int max_st = 6;
        if (const char* ms = getenv("GTTS_HALO_STAGES")) max_st = atoi(ms);
        // CTA pairs for every halo conv (measured, chunk 16x1720: 64->64 174 -> 136 us, 128->128 134 -> 108 us,
        // 256->64 192 -> 131 us, 512->128 144 -> 125 us, 256->256 134 -> 123 us).  GTTS_CTA2=0: single-CTA kernels.
        const bool cta2 = convT_halo || (halo_mode == 2 && conv_tc_cta2_enabled() && p.num_tiles >= 2 && num_sms >= 2);
        if (g.split && !cta2) { set_error("conv_tc: split convs need the CTA-pair halo kernel or per-tap boxes"); delete pl; return nullptr; }
        if (cta2) {
            // CTA pair: every CTA holds half of each weight tile (Cout/2 rows)
            const int bhalf = g.Cout * 64;
            p.mc = 2;
            if (ntiles_b * bhalf + 3 * abytes <= budget) {
                p.b_resident = 1; p.b_slots = ntiles_b;
                p.stages = (budget - ntiles_b * bhalf) / abytes;
                if (p.stages > max_st) p.stages = max_st;
            } else {
                p.b_resident = 0; p.stages = convT_halo ? 2 * nck : 3;   // ConvT keeps all chunks of a tile resident over its 4 phases
                p.b_slots = (budget - p.stages * abytes) / bhalf;
                if (p.b_slots > 16) p.b_slots = 16;
            }
            if (convT_halo && p.stages < nck) { set_error("conv_tc: ConvT halo variant: not enough A stages"); delete pl; return nullptr; }
            pl->smem = (size_t)p.stages * abytes + (size_t)p.b_slots * bhalf + kMiscBytes + 1024 + apply_extra;
        } else {
        if (ntiles_b <= 16 && max_st >= 6 && ntiles_b * btile + 6 * abytes <= budget) { p.stages = 6; p.b_resident = 1; p.b_slots = ntiles_b; }
        else if (ntiles_b <= 16 && ntiles_b * btile + 4 * abytes <= budget) { p.stages = 4; p.b_resident = 1; p.b_slots = ntiles_b; }
        else if (ntiles_b <= 16 && ntiles_b * btile + 3 * abytes <= budget) { p.stages = 3; p.b_resident = 1; p.b_slots = ntiles_b; }
        else if (ntiles_b <= 16 && ntiles_b * btile + 2 * abytes <= budget) { p.stages = 2; p.b_resident = 1; p.b_slots = ntiles_b; }
        else {
            // streamed weights: two A tiles per pass of the weight ring, A ring double-buffered across chunks
            p.pass_tiles = 2;
            if (const char* pt = getenv("GTTS_PASS_TILES")) p.pass_tiles = atoi(pt) == 1 ? 1 : 2;
            p.stages = p.pass_tiles == 2 ? 4 : 3; p.b_resident = 0;
            p.b_slots = (budget - p.stages * abytes) / btile;
            if (p.b_slots > 16) p.b_slots = 16;
        }
        pl->smem = (size_t)p.stages * abytes + (size_t)p.b_slots * btile + kMiscBytes + 1024;
        }
    } else {
        const int btile = g.Cout * 128, nwt = g.nphase * g.ntaps * (p.nchunk0 + p.nchunk1);
        const bool want_res = getenv("GTTS_TAP_RESIDENT") ? atoi(getenv("GTTS_TAP_RESIDENT")) != 0 : true;
        // 1-D halo mode (the vocoder's dilated Conv1d stacks): H = 1, stride 1, several taps along W, weights resident
        int halo = 0;
        bool h1 = g.Hg == 1 && g.Hin == 1 && g.stride == 1 && g.nphase == 1 && g.ntaps > 1 && g.w_batch_rows == 0 && !g.split &&
                  !p.mc && !(getenv("GTTS_HALO1D") && atoi(getenv("GTTS_HALO1D")) == 0);
        for (int t = 0; t < g.ntaps && h1; ++t) {
            if (g.dy[0][t] != 0) h1 = false;
            const int a = g.dx[0][t] < 0 ? -g.dx[0][t] : g.dx[0][t];
            if (a > halo) halo = a;
        }
        if (h1 && halo >= 1 && p.bw + 2 * halo <= 256) {
            const int rows = p.bw + 2 * halo;
            const int a_stage = (rows * 128 + 1023) / 1024 * 1024;
            if (want_res && nwt * btile + 2 * a_stage <= budget && nwt * btile < (1 << 20)) {
                p.halo1d = halo; p.a_stage = a_stage; p.a_bytes = rows * 128;
                for (int t = 0; t < g.ntaps; ++t) p.tap_row[t] = (int16_t)(halo + g.dx[0][t]);
                p.b_resident = 1; p.b_slots = nwt;
                int stages = (budget - nwt * btile) / a_stage;
                if (stages > 8) stages = 8;
                p.stages = stages;
                pl->smem = (size_t)stages * a_stage + (size_t)nwt * btile + kMiscBytes + 1024;
            } else if (!(getenv("GTTS_HALO1D") && atoi(getenv("GTTS_HALO1D")) == 1)) {
                // weights streamed through their own ring (GTTS_HALO1D=1: resident-weight layers only)
                int slots = (budget - 3 * a_stage) / btile;
                if (slots > 16) slots = 16;
                if (slots >= 4) {
                    p.halo1d = halo; p.a_stage = a_stage; p.a_bytes = rows * 128;
                    for (int t = 0; t < g.ntaps; ++t) p.tap_row[t] = (int16_t)(halo + g.dx[0][t]);
                    p.b_resident = 0; p.b_slots = slots; p.stages = 3;
                    pl->smem = (size_t)3 * a_stage + (size_t)slots * btile + kMiscBytes + 1024;
                }
            }
        }
        if (p.halo1d) {
        } else if (want_res && g.w_batch_rows == 0 && nwt * btile + 4 * kABytes <= budget && nwt * btile < (1 << 20)) {
            p.b_resident = 1; p.b_slots = nwt;
            int stages = (budget - nwt * btile) / kABytes;
            if (stages > 8) stages = 8;
            p.stages = stages;
            pl->smem = (size_t)stages * kABytes + (size_t)nwt * btile + kMiscBytes + 1024;
        } else {
            const int stage_bytes = kABytes + btile;
            int stages = budget / stage_bytes;
            if (stages > 8) stages = 8;
            p.stages = stages;
            pl->smem = (size_t)stages * stage_bytes + kMiscBytes + 1024;
        }
    }
    pl->grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    if (pl->grid > 256) pl->grid = 256;                    // GN partial buffers hold 256 CTA slots per sample
    {
        // 2-CTA clusters with weight-tile multicast: the N=256 layers are bound by L2->SMEM fill, 2/3 of it weights
        const char* mce = getenv("GTTS_MC");
        const int want = mce ? atoi(mce) : 0;      // measured: no gain on B200 (these layers are MMA-, not fill-bound)
        if (p.mc != 2)
            p.mc = (want && !halo_mode && !p.b_resident && g.Cout == 256 && g.nphase == 1 && g.w_batch_rows == 0 && p.num_tiles >= 2 &&
                    num_sms >= 2) ? 1 : 0;
        if (p.mc) {
            int gmax = num_sms & ~1;
            if (gmax > 256) gmax = 256;
            const int need = (p.num_tiles + 1) & ~1;
            pl->grid = need < gmax ? need : gmax;
        }
    }
    if (e.apply) {
        // per-sample grid barrier: every CTA of the grid must be on an SM at the same time
        if (p.mc != 2) { set_error("conv_tc: the GroupNorm-apply epilogue needs the CTA-pair halo kernel (geometry not eligible)"); delete pl; return nullptr; }
        const int cores = conv_tc_halo2_max_grid(g.Cout, e.residual != nullptr, pl->smem, e.apply == 2);
        if (cores >= 2 && cores < pl->grid) pl->grid = cores & ~1;
        if (const char* ge = getenv("GTTS_APPLY_GRID")) { const int gg = atoi(ge) & ~1; if (gg >= 2 && gg < pl->grid) pl->grid = gg; }
        if (e.apply == 1 && !apply_run_fits(p.tiles_h * p.tiles_w, pl->grid, g.Cout)) {
            set_error("conv_tc: GroupNorm-apply epilogue: a sample's tiles do not fit the TMEM accumulator ring at grid " + std::to_string(pl->grid));
            delete pl;
            return nullptr;
        }
    }

    bool ok = true;
    const uint64_t H = g.Hin, W = g.Win;
    auto make_a = [&](CUtensorMap* m, const void* src, int C) {
        if (!p.stride2) {
            uint64_t dims[4] = {(uint64_t)C, W, H, (uint64_t)g.B};
            uint64_t str[3] = {(uint64_t)C * 2, W * C * 2, H * W * C * 2};
            uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
            if (halo_mode) { box[1] = halo_mode == 2 ? 10 : 16; box[2] = 18; }
            if (p.halo1d) box[1] = (uint32_t)(p.bw + 2 * p.halo1d);
            if (p.halo_t) {                                          // dims (C, H, W, B): H is the fast box dimension
                uint64_t dims_t[4] = {(uint64_t)C, H, W, (uint64_t)g.B};
                uint64_t str_t[3] = {W * C * 2, (uint64_t)C * 2, H * W * C * 2};
                return encode_map(m, src, 4, dims_t, str_t, box);
            }
            return encode_map(m, src, 4, dims, str, box);
        } else {
            uint64_t dims[5] = {(uint64_t)2 * C, W / 2, 2, H / 2, (uint64_t)g.B};
            uint64_t str[4] = {(uint64_t)2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, 1};
            return encode_map(m, src, 5, dims, str, box);
        }
    };
    const int cmul = g.split ? 3 : 1;                  // split: [hi | mid | lo] planes per pixel
    ok = ok && make_a(&pl->mapA0, src0, cmul * g.Cin0);
    if (g.Cin1 > 0) ok = ok && make_a(&pl->mapA1, src1, cmul * g.Cin1);
    else pl->mapA1 = pl->mapA0;
    {
        const uint64_t K = (uint64_t)(g.split ? 6 : 1) * (uint64_t)(g.Cin0 + g.Cin1);
        uint64_t dims[2] = {K, (uint64_t)weight_rows};
        uint64_t str[1] = {K * 2};
        uint32_t box[2] = {64, (uint32_t)g.Cout};
        ok = ok && encode_map(&pl->mapW, weight, 2, dims, str, box);
        uint32_t boxh[2] = {64, (uint32_t)g.Cout / 2};
        ok = ok && encode_map(&pl->mapWh, weight, 2, dims, str, boxh);
    }
    if (!ok) { delete pl; return nullptr; }
    if (e.in_stats && p.mc != 2) {
        set_error("conv_tc: the fused input transform needs the CTA-pair halo kernel (geometry not eligible)");
        delete pl;
        return nullptr;
    }
    return pl;
}

void conv_tc_plan_destroy(TcConvPlan* p) { delete p; }
int conv_tc_plan_grid(const TcConvPlan* p) { return p->grid; }
void conv_tc_plan_set_debug(TcConvPlan* p, unsigned long long* dbg_out) { p->p.dbg_out = dbg_out; }

namespace {
template <int N, bool kStats, bool kRes, bool kMask, bool kOutF32 = false, bool kAct = false>
int launch_variant(const TcConvPlan* pl, cudaStream_t stream) {
    static bool attr_set = false;
    auto k = conv_tc_kernel<N, kStats, kRes, kMask, kOutF32, kAct>;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GTTS_CHECK_CUDA(launch_pdl(k, dim3(pl->grid), dim3(kThreads), pl->smem, stream, pl->p.mc ? 2 : 1, pl->mapA0, pl->mapA1,
                               pl->mapW, pl->mapWh, pl->p));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template <int N>
int launch_n(const TcConvPlan* pl, cudaStream_t stream) {
    const ConvEpilogue& e = pl->p.e;
    const bool st = e.gn_partials != nullptr, rs = e.residual != nullptr, mk = e.mask != nullptr;
    if (e.out_f32) {                                   // fp32 activations (fp32 mode on the tensor cores)
        if (st) return launch_variant<N, true, false, false, true>(pl, stream);
        if (rs && mk) return launch_variant<N, false, true, true, true>(pl, stream);
        if (rs) return launch_variant<N, false, true, false, true>(pl, stream);
        if (mk) return launch_variant<N, false, false, true, true>(pl, stream);
        return launch_variant<N, false, false, false, true>(pl, stream);
    }
    if (e.act_out || e.out2) {                         // 1-D conv stacks: leaky-ReLU output(s), optional residual
        if (rs) return launch_variant<N, false, true, false, false, true>(pl, stream);
        return launch_variant<N, false, false, false, false, true>(pl, stream);
    }
    if (st) return launch_variant<N, true, false, false>(pl, stream);
    if (rs && mk) return launch_variant<N, false, true, true>(pl, stream);
    if (rs) return launch_variant<N, false, true, false>(pl, stream);
    if (mk) return launch_variant<N, false, false, true>(pl, stream);
    return launch_variant<N, false, false, false>(pl, stream);
}
}  // namespace

int conv_tc_launch(const TcConvPlan* pl, cudaStream_t stream) {
    if (pl->p.num_tiles == 0) return 0;
    if (pl->p.halo_mode) return pl->p.mc == 2 ? conv_tc_halo2_launch(pl, stream) : conv_tc_halo_launch(pl, stream);
    if (pl->N == 64) return launch_n<64>(pl, stream);
    if (pl->N == 128) return launch_n<128>(pl, stream);
    return launch_n<256>(pl, stream);
}

}  // namespace gtts
