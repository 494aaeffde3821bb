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
#include <cuda.h>

#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

constexpr int kABytes = 128 * 128;                 // 128 pixel rows x 64 bf16
constexpr int kMiscBytes = 4096;                   // barriers + epilogue scratch

struct TcParams {
    int bh, bw, tiles_h, tiles_w, nphase, B;
    int Hg, Wg, Hout, Wout, out_step;
    int ntaps, nchunk0, nchunk1, Cin0;
    int stride2, w_batch_rows, num_tiles, a_bytes, stages;
    int8_t dy[4][9], dx[4][9];
    int wrow[4][9];
    int oy[4], ox[4];
    ConvEpilogue e;
};

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

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

constexpr int kStatSlots = 4;          // ring of per-tile GroupNorm partials between epilogue warps and the stats warp
constexpr int kThreads = 384;          // warps 0-3: TMA, MMA, TMEM alloc, stats; warps 4-11: epilogue

// kStats: GroupNorm partial statistics of (acc + bias); kRes: + residual; kMask: * mask[b][w]
template <int N, bool kStats, bool kRes, bool kMask>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapW, const TcParams p) {
    constexpr int kBBytes = N * 128;
    constexpr int kStage = kABytes + kBBytes;
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t kTmemCols = 2 * N;          // 128 / 256 / 512: power of two
    constexpr int kColsPerWarp = N / 2;            // two epilogue warps share each TMEM lane quarter
    constexpr int kGsz = N / 8;                    // channels per GroupNorm group (4 groups per column half)

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* misc = smem + (size_t)p.stages * kStage;
    uint64_t* full = reinterpret_cast<uint64_t*>(misc);            // [8]
    uint64_t* empty = full + 8;                                    // [8]
    uint64_t* tfull = empty + 8;                                   // [2]
    uint64_t* tempty = tfull + 2;                                  // [2]
    uint64_t* sfull = tempty + 2;                                  // [kStatSlots]
    uint64_t* sempty = sfull + kStatSlots;                         // [kStatSlots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + kStatSlots);
    float* s_bias = reinterpret_cast<float*>(misc + 512);          // [256]
    float* s_ring = reinterpret_cast<float*>(misc + 1536);         // [kStatSlots][8 warps][8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
        for (int i = 0; i < kStatSlots; ++i) { mbar_init(&sfull[i], 8); mbar_init(&sempty[i], 1); }
        mbar_fence_init();
    } else if (warp == 2) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    for (int i = tid; i < N; i += kThreads) s_bias[i] = p.e.bias ? p.e.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nkb = p.ntaps * (p.nchunk0 + p.nchunk1);
    const int tiles_per_phase = p.tiles_h * p.tiles_w;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h;
                const int ph = (tile / tiles_per_phase) % p.nphase, b = tile / (tiles_per_phase * p.nphase);
                const int h0 = th * p.bh, w0 = tw * p.bw;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int dy = p.dy[ph][tap], dx = p.dx[ph][tap];
                    const int wr = p.wrow[ph][tap] + b * p.w_batch_rows;
                    for (int ck = 0; ck < p.nchunk0 + p.nchunk1; ++ck) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * kStage;
                        mbar_expect_tx(&full[stage], (uint32_t)(p.a_bytes + kBBytes));
                        if (p.stride2) {
                            // 5-D view (2C, W/2, 2, H/2, B): input pixel 2*o + d, d in {-1,0,1}
                            const int px = dx & 1, py = dy & 1;
                            tma_load_5d(&mapA0, &full[stage], sa, px * p.Cin0 + ck * 64, w0 + (dx < 0 ? -1 : 0), py,
                                        h0 + (dy < 0 ? -1 : 0), b);
                        } else if (ck < p.nchunk0) {
                            tma_load_4d(&mapA0, &full[stage], sa, ck * 64, w0 + dx, h0 + dy, b);
                        } else {
                            tma_load_4d(&mapA1, &full[stage], sa, (ck - p.nchunk0) * 64, w0 + dx, h0 + dy, b);
                        }
                        tma_load_2d(&mapW, &full[stage], sa + kABytes, ck * 64, wr);
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&tempty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStage);
                    const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
                    const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)                     // 4 x (K = 16 bf16 = 32 bytes)
                        tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                   (uint32_t)((kb | k) != 0));
                    tc_commit(&empty[stage]);                       // smem slot free when these MMAs retire
                    if (kb == nkb - 1) tc_commit(&tfull[buf]);      // accumulator complete
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 3) {
        // ================================================================ GroupNorm statistics warp
        if (kStats) {
            const ConvEpilogue& e = p.e;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int slot = it % kStatSlots;
                const int b = tile / tiles_per_phase, slot_in_sample = tile - b * tiles_per_phase;
                mbar_wait(&sfull[slot], (uint32_t)(it / kStatSlots) & 1u);
                // value k (0..7 sums, 8..15 sums of squares) of group g = k & 7 lives in column half g >> 2
                float v = 0.f;
                if (lane < 16) {
                    const int g = lane & 7, which = lane >> 3, half = g >> 2, idx = which * 4 + (g & 3);
                    const float* r = s_ring + (slot * 8 + half * 4) * 8 + idx;
                    v = (r[0] + r[8]) + (r[16] + r[24]);
                    e.gn_partials[((size_t)b * tiles_per_phase + slot_in_sample) * 16 + lane] = v;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty[slot]);
            }
        }
    } else if (warp >= 4) {
        // ================================================================ epilogue (8 warps, 256 threads)
        const int ew = warp - 4, wq = ew & 3, half = ew >> 2;
        const int row = wq * 32 + lane;                              // TMEM lane = pixel row of the tile
        const ConvEpilogue& e = p.e;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(e.residual);
        const int cbase = half * kColsPerWarp;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h;
            const int ph = (tile / tiles_per_phase) % p.nphase, b = tile / (tiles_per_phase * p.nphase);
            const int hl = row / p.bw, wl = row - hl * p.bw;
            const int j = th * p.bh + hl, i = tw * p.bw + wl;
            const bool valid = (hl < p.bh) && (j < p.Hg) && (i < p.Wg);
            const int oh = j * p.out_step + p.oy[ph], ow = i * p.out_step + p.ox[ph];
            const size_t opix = valid ? ((size_t)b * p.Hout + oh) * p.Wout + ow : 0;
            float m = 1.0f;
            if (kMask) m = valid ? e.mask[(size_t)b * p.Wout + ow] : 0.f;

            mbar_wait(&tfull[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * N + cbase);

            float st[8];                                             // [0..3] sums, [4..7] sums of squares
#pragma unroll
            for (int g = 0; g < 8; ++g) st[g] = 0.f;

#pragma unroll
            for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + (uint32_t)c0, r);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[cbase + c0 + q4 * 4]);
                    f[q4 * 4 + 0] = __uint_as_float(r[q4 * 4 + 0]) + b4.x;
                    f[q4 * 4 + 1] = __uint_as_float(r[q4 * 4 + 1]) + b4.y;
                    f[q4 * 4 + 2] = __uint_as_float(r[q4 * 4 + 2]) + b4.z;
                    f[q4 * 4 + 3] = __uint_as_float(r[q4 * 4 + 3]) + b4.w;
                }
                if (kStats) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int g = (c0 + q) / kGsz;               // local group 0..3 (compile time)
                        const float x = valid ? f[q] : 0.f;
                        st[g] += x;
                        st[4 + g] = fmaf(x, x, st[4 + g]);
                    }
                }
                if (valid) {
                    if (kRes) {
                        const uint4* rp = reinterpret_cast<const uint4*>(res + opix * N + cbase + c0);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            const uint4 u = __ldg(rp + v4);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                f[v4 * 8 + 2 * k] += __uint_as_float(w[k] << 16);
                                f[v4 * 8 + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
                            }
                        }
                    }
                    if (kMask) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) f[q] *= m;
                    }
                    uint4* op = reinterpret_cast<uint4*>(out + opix * N + cbase + c0);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[v4 * 8 + 2 * k], f[v4 * 8 + 2 * k + 1]);
                            w[k] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        op[v4] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            // all TMEM reads of this buffer are complete: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty[buf]);

            if (kStats) {
                const float t = warp_reduce8(st, lane);
                const int slot = it % kStatSlots;
                mbar_wait(&sempty[slot], ((uint32_t)(it / kStatSlots) & 1u) ^ 1u);
                if ((lane & 3) == 0) s_ring[(slot * 8 + ew) * 8 + (lane >> 2)] = t;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sfull[slot]);            // release: orders the ring writes of this warp
            }
        }
    }

    // ---- GroupNorm statistics: ONE fence + ticket per CTA (not per tile).  Every CTA publishes how many tiles of each
    // sample it contributed; whoever completes a sample's count reduces that sample's partials in a fixed order.
    if (kStats) __threadfence();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (kStats) {
        const ConvEpilogue& e = p.e;
        int* s_nfin = reinterpret_cast<int*>(misc + 2560);
        int* s_fin = s_nfin + 1;
        double* s_red = reinterpret_cast<double*>(smem);             // pipeline buffers are idle now: [24][16]
        if (tid == 0) *s_nfin = 0;
        __syncthreads();
        if (warp == 0) {
            const int G = (int)gridDim.x, bx = (int)blockIdx.x, tps = tiles_per_phase;
            for (int b = lane; b < p.B; b += 32) {
                const int lo = b * tps, hi = lo + tps - 1;           // tiles of sample b: lo..hi; mine: bx + i*G
                const int i_min = lo > bx ? (lo - bx + G - 1) / G : 0;
                const int i_max = hi >= bx ? (hi - bx) / G : -1;
                const int cnt = i_max - i_min + 1;
                if (cnt > 0) {
                    const unsigned int old = atomicAdd(&e.gn_counters[b], (unsigned int)cnt);
                    if (old + (unsigned int)cnt == (unsigned int)tps) s_fin[atomicAdd(s_nfin, 1)] = b;
                }
            }
        }
        __syncthreads();
        const int nfin = *s_nfin;
        if (nfin > 0) {
            __threadfence();
            const double inv_count = 1.0 / ((double)kGsz * (double)p.Hout * (double)p.Wout);
            for (int f = 0; f < nfin; ++f) {
                const int b = s_fin[f];
                const int k = tid & 15, slice = tid >> 4;            // 24 slices of 16 components
                const float* pp = e.gn_partials + (size_t)b * tiles_per_phase * 16 + k;
                double acc = 0.0;
                for (int sl = slice; sl < tiles_per_phase; sl += kThreads / 16) acc += (double)__ldcg(pp + (size_t)sl * 16);
                s_red[slice * 16 + k] = acc;
                __syncthreads();
                if (tid < 8) {
                    double sum = 0.0, sq = 0.0;
                    for (int sl = 0; sl < kThreads / 16; ++sl) { sum += s_red[sl * 16 + tid]; sq += s_red[sl * 16 + 8 + tid]; }
                    const double mean = sum * inv_count;
                    double var = sq * inv_count - mean * mean;
                    if (var < 0.0) var = 0.0;
                    e.gn_stats[((size_t)b * 8 + tid) * 2 + 0] = (float)mean;
                    e.gn_stats[((size_t)b * 8 + tid) * 2 + 1] = (float)(1.0 / sqrt(var + (double)e.gn_eps));
                }
                if (tid == 0) e.gn_counters[b] = 0u;
                __syncthreads();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

bool encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
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

struct TcConvPlan {
    CUtensorMap mapA0, mapA1, mapW;
    TcParams p;
    int N, grid;
    size_t smem;
};

size_t conv_tc_partials_slots(const ConvGeom& g) {
    int bh, bw;
    pick_tile(g.Hg, g.Wg, &bh, &bw);
    return (size_t)((g.Hg + bh - 1) / bh) * ((g.Wg + bw - 1) / bw);
}

TcConvPlan* conv_tc_plan_create(const ConvGeom& g, const void* src0, const void* src1, const void* weight,
                                int weight_rows, const ConvEpilogue& e, int num_sms) {
    if (!(g.Cout == 64 || g.Cout == 128 || g.Cout == 256)) { set_error("conv_tc: Cout must be 64/128/256"); return nullptr; }
    if (g.Cin0 % 64 || g.Cin1 % 64 || g.Cin0 <= 0) { set_error("conv_tc: Cin must be a multiple of 64"); return nullptr; }
    if (g.stride == 2 && (g.Cin1 != 0 || (g.Hin & 1) || (g.Win & 1))) { set_error("conv_tc: bad stride-2 geometry"); return nullptr; }
    if (e.gn_partials && (g.nphase != 1 || e.residual || e.mask)) { set_error("conv_tc: GN statistics only on plain convs"); return nullptr; }
    TcConvPlan* pl = new TcConvPlan();
    memset(pl, 0, sizeof(*pl));
    TcParams& p = pl->p;
    pick_tile(g.Hg, g.Wg, &p.bh, &p.bw);
    p.tiles_h = (g.Hg + p.bh - 1) / p.bh;
    p.tiles_w = (g.Wg + p.bw - 1) / p.bw;
    p.nphase = g.nphase; p.B = g.B;
    p.Hg = g.Hg; p.Wg = g.Wg; p.Hout = g.Hout; p.Wout = g.Wout; p.out_step = g.out_step;
    p.ntaps = g.ntaps; p.nchunk0 = g.Cin0 / 64; p.nchunk1 = g.Cin1 / 64; p.Cin0 = g.Cin0;
    p.stride2 = (g.stride == 2); p.w_batch_rows = g.w_batch_rows;
    p.num_tiles = g.B * g.nphase * p.tiles_h * p.tiles_w;
    p.a_bytes = p.bh * p.bw * 128;
    memcpy(p.dy, g.dy, sizeof(p.dy)); memcpy(p.dx, g.dx, sizeof(p.dx));
    memcpy(p.wrow, g.wrow, sizeof(p.wrow)); memcpy(p.oy, g.oy, sizeof(p.oy)); memcpy(p.ox, g.ox, sizeof(p.ox));
    p.e = e;
    pl->N = g.Cout;
    const int stage_bytes = kABytes + g.Cout * 128;
    int stages = (227 * 1024 - kMiscBytes - 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    p.stages = stages;
    pl->smem = (size_t)stages * stage_bytes + kMiscBytes + 1024;
    pl->grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;

    bool ok = true;
    const uint64_t H = g.Hin, W = g.Win;
    auto make_a = [&](CUtensorMap* m, const void* src, int C) {
        if (!p.stride2) {
            uint64_t dims[4] = {(uint64_t)C, W, H, (uint64_t)g.B};
            uint64_t str[3] = {(uint64_t)C * 2, W * C * 2, H * W * C * 2};
            uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
            return encode_map(m, src, 4, dims, str, box);
        } else {
            uint64_t dims[5] = {(uint64_t)2 * C, W / 2, 2, H / 2, (uint64_t)g.B};
            uint64_t str[4] = {(uint64_t)2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, 1};
            return encode_map(m, src, 5, dims, str, box);
        }
    };
    ok = ok && make_a(&pl->mapA0, src0, g.Cin0);
    if (g.Cin1 > 0) ok = ok && make_a(&pl->mapA1, src1, g.Cin1);
    else pl->mapA1 = pl->mapA0;
    {
        const uint64_t K = (uint64_t)(g.Cin0 + g.Cin1);
        uint64_t dims[2] = {K, (uint64_t)weight_rows};
        uint64_t str[1] = {K * 2};
        uint32_t box[2] = {64, (uint32_t)g.Cout};
        ok = ok && encode_map(&pl->mapW, weight, 2, dims, str, box);
    }
    if (!ok) { delete pl; return nullptr; }
    return pl;
}

void conv_tc_plan_destroy(TcConvPlan* p) { delete p; }

namespace {
template <int N, bool kStats, bool kRes, bool kMask>
int launch_variant(const TcConvPlan* pl, cudaStream_t stream) {
    static bool attr_set = false;
    auto k = conv_tc_kernel<N, kStats, kRes, kMask>;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    k<<<pl->grid, kThreads, pl->smem, stream>>>(pl->mapA0, pl->mapA1, pl->mapW, pl->p);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
template <int N>
int launch_n(const TcConvPlan* pl, cudaStream_t stream) {
    const ConvEpilogue& e = pl->p.e;
    const bool st = e.gn_partials != nullptr, rs = e.residual != nullptr, mk = e.mask != nullptr;
    if (st) return launch_variant<N, true, false, false>(pl, stream);
    if (rs && mk) return launch_variant<N, false, true, true>(pl, stream);
    if (rs) return launch_variant<N, false, true, false>(pl, stream);
    if (mk) return launch_variant<N, false, false, true>(pl, stream);
    return launch_variant<N, false, false, false>(pl, stream);
}
}  // namespace

int conv_tc_launch(const TcConvPlan* pl, cudaStream_t stream) {
    if (pl->p.num_tiles == 0) return 0;
    if (pl->N == 64) return launch_n<64>(pl, stream);
    if (pl->N == 128) return launch_n<128>(pl, stream);
    return launch_n<256>(pl, stream);
}

}  // namespace gtts
