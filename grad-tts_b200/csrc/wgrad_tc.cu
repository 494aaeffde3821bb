// Weight gradients of the stride-1 convolutions on the 5th-generation tensor cores (bf16 training plan).
//
//   dW[tap][co][ci] = sum_pixels gout[p][co] * x[p + tap][ci]          (reference: autograd of Conv2d, model/diffusion.py:52,70)
//
// is a GEMM whose K dimension is the PIXEL axis.  Both operands are NHWC tensors, i.e. [pixel][channel] rows: K-outermost, or
// "MN-major" in tcgen05 terms -- which the instruction reads natively (a_major = b_major = 1 in the instruction descriptor), so
// the tiles are fetched by exactly the TMA boxes the forward convolution uses (128 pixels x 64 channels, 128-byte swizzle, the
// tap shift in the box coordinates, zero padding = TMA out-of-bounds fill) and no transposed copy of any activation is made:
//     A = gout tile  [K = 128 pixels][M = 128 output channels]   (two 64-channel TMA blocks, LBO apart)
//     B = x tile     [K = 128 pixels][N = 64..256 input channels] (N / 64 TMA blocks, LBO apart), shifted by the tap
//     D[128 x N] (TMEM, fp32) += A^T B, eight K = 16 MMAs per tile.
// A CTA owns one (128 output channels, N input channels, tap, pixel slice) item: warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> fp32 partials [slice][tap][co][ci], summed in fixed order by
// wgrad_reduce_kernel).  Shared-memory descriptor for an MN-major, 128-byte-swizzled operand (canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): rows of 128 bytes are consecutive K (pixels), 8-row groups SBO = 1024 bytes
// apart, 64-channel blocks LBO = 16 KB apart.
#include <cstdlib>

#include "conv_tc_common.cuh"

namespace gtts {

using namespace tc;

namespace {

constexpr int kWtThreads = 256;
constexpr int kBlockBytes = 128 * 128;            // one TMA block: 128 pixels x 64 bf16 channels

struct WgradTcParams {
    int B, tiles_h, tiles_w, bh, bw;
    int ntaps;
    int8_t dy[16], dx[16];
    int Cout, Cin0, Cin1, cin_tot, Nn;
    int m_tiles, n_blocks, slices, tiles_per_slice, num_tiles, stages, a_blocks;
    float* partial;
};

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)(lbo >> 4) << 16;               // leading byte offset: between 64-element blocks along M / N
    d |= (uint64_t)(sbo >> 4) << 32;               // stride byte offset: between 8-row groups along K
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc_mn() {
    // kind::f16: D = f32 (bit 4), A = bf16 (bit 7), B = bf16 (bit 10), A and B MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

template <int N>
__global__ void __launch_bounds__(kWtThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX0,
                const __grid_constant__ CUtensorMap mapX1, const WgradTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    constexpr int kBBlocks = N / 64;
    const int stage_bytes = (p.a_blocks + kBBlocks) * kBlockBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                       // [8]
    uint64_t* empty = bars + 8;                  // [8]
    uint64_t* tfull = bars + 16;                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // item = (((m * n_blocks + nb) * ntaps + tap) * slices + slice)
    int item = blockIdx.x;
    const int slice = item % p.slices; item /= p.slices;
    const int tap = item % p.ntaps; item /= p.ntaps;
    const int nb = item % p.n_blocks, m = item / p.n_blocks;
    const int t_begin = slice * p.tiles_per_slice, t_end = min(p.num_tiles, t_begin + p.tiles_per_slice);
    const int n_tiles = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapG);
        tma_prefetch_desc(&mapX0);
        tma_prefetch_desc(&mapX1);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, N < 32 ? 32 : N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            const int ci0 = nb * N;
            const bool second = ci0 >= p.Cin0;
            const CUtensorMap* mx = second ? &mapX1 : &mapX0;
            const int xchan = second ? ci0 - p.Cin0 : ci0;
            const int dy = p.dy[tap], dx = p.dx[tap];
            const int tiles_per_sample = p.tiles_h * p.tiles_w;
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int b = t / tiles_per_sample, r = t - b * tiles_per_sample;
                const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
                const int h0 = th * p.bh, w0 = tw * p.bw;
                mbar_wait(&empty[stage], phase ^ 1u);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                for (int ab = 0; ab < p.a_blocks; ++ab)
                    tma_load_4d(&mapG, &full[stage], sa + ab * kBlockBytes, m * 128 + ab * 64, w0, h0, b);
                uint8_t* sb = sa + p.a_blocks * kBlockBytes;
#pragma unroll
                for (int bb = 0; bb < kBBlocks; ++bb)
                    tma_load_4d(mx, &full[stage], sb + bb * kBlockBytes, xchan + bb * 64, w0 + dx, h0 + dy, b);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        constexpr uint32_t kIdesc = make_idesc_mn<N>();
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < n_tiles; ++t) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(smem) + (uint32_t)(stage * stage_bytes);
                const uint32_t b0 = a0 + (uint32_t)(p.a_blocks * kBlockBytes);
                // with one 64-channel block of gout (Cout = 64) rows 64..127 of D read whatever follows the block (the first x
                // block): finite values in rows the epilogue never stores
#pragma unroll
                for (int k = 0; k < 8; ++k) {                          // 8 x (K = 16 pixels = 16 rows of 128 bytes)
                    const uint64_t adesc = make_sw128_mnmajor_desc(a0 + (uint32_t)(k * 2048), kBlockBytes, 1024);
                    const uint64_t bdesc = make_sw128_mnmajor_desc(b0 + (uint32_t)(k * 2048), kBlockBytes, 1024);
                    tc_mma_f16(tmem_base, adesc, bdesc, kIdesc, (uint32_t)((t | k) != 0));
                }
                tc_commit(&empty[stage]);
                if (t == n_tiles - 1) tc_commit(tfull);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp >= 4) {
        // ================================================================ epilogue: D (TMEM) -> fp32 partials
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int co = m * 128 + row;
        float* o = p.partial + (((size_t)slice * p.ntaps + tap) * p.Cout + (co < p.Cout ? co : 0)) * p.cin_tot + nb * N;
        if (n_tiles > 0) {
            mbar_wait(tfull, 0u);
            tc_fence_after();
        }
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            if (n_tiles > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) r[q] = 0u;
            }
            if (co < p.Cout) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<uint4*>(o + c0 + q) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, N < 32 ? 32 : N);
    }
}

void pick_tile128(int Hg, int Wg, int* bh_out, int* bw_out) {
    long best = -1;
    int bbh = 1, bbw = 128;
    for (int bh = 1; bh <= 128; ++bh) {
        if (128 % bh) continue;
        const int bw = 128 / bh;
        if (bh > Hg && bh != 1) continue;
        const long tiles = (long)((Hg + bh - 1) / bh) * ((Wg + bw - 1) / bw);
        const long score = tiles * 1024 - bw;
        if (best < 0 || score < best) { best = score; bbh = bh; bbw = bw; }
    }
    *bh_out = bbh;
    *bw_out = bbw;
}

int pick_n(const ConvGeom& g) {
    for (int n : {256, 128, 64})
        if (g.Cin0 % n == 0 && g.Cin1 % n == 0) return n;
    return 0;
}

}  // namespace

struct WgradTcPlan {
    CUtensorMap mapG, mapX0, mapX1;
    WgradTcParams p;
    int grid;
    size_t smem;
};

bool wgrad_tc_eligible(const ConvGeom& g) {
    if (const char* e = getenv("GTTS_WGRAD_TC")) { if (atoi(e) == 0) return false; }
    return g.nphase == 1 && g.stride == 1 && g.out_step == 1 && g.w_batch_rows == 0 && !g.split && g.ntaps >= 1 && g.ntaps <= 16 &&
           g.Cout % 64 == 0 && g.Cin0 % 64 == 0 && g.Cin1 % 64 == 0 && g.Cin0 > 0 && g.Hin == g.Hout && g.Win == g.Wout && pick_n(g) > 0;
}

// floats of the partial buffer ([slice][tap][Cout][cin_tot]) and the number of pixel slices for this geometry
size_t wgrad_tc_partial_floats(const ConvGeom& g, int num_sms, int* slices_out) {
    int bh, bw;
    pick_tile128(g.Hg, g.Wg, &bh, &bw);
    const int num_tiles = g.B * ((g.Hg + bh - 1) / bh) * ((g.Wg + bw - 1) / bw);
    const int n = pick_n(g), cin = g.Cin0 + g.Cin1;
    const int base = ((g.Cout + 127) / 128) * (cin / (n > 0 ? n : 64)) * g.ntaps;
    int slices = (2 * num_sms + base - 1) / base;
    if (slices > 64) slices = 64;
    if (slices > num_tiles) slices = num_tiles;
    if (slices < 1) slices = 1;
    const int tps = (num_tiles + slices - 1) / slices;
    slices = (num_tiles + tps - 1) / tps;
    if (slices_out) *slices_out = slices;
    return (size_t)slices * g.ntaps * g.Cout * cin;
}

WgradTcPlan* wgrad_tc_plan_create(const ConvGeom& g, const void* gout, const void* x0, const void* x1, float* partial, int num_sms) {
    if (!wgrad_tc_eligible(g)) { set_error("wgrad_tc: geometry not eligible"); return nullptr; }
    WgradTcPlan* pl = new WgradTcPlan();
    memset(pl, 0, sizeof(*pl));
    WgradTcParams& p = pl->p;
    pick_tile128(g.Hg, g.Wg, &p.bh, &p.bw);
    p.B = g.B;
    p.tiles_h = (g.Hg + p.bh - 1) / p.bh;
    p.tiles_w = (g.Wg + p.bw - 1) / p.bw;
    p.num_tiles = g.B * p.tiles_h * p.tiles_w;
    p.ntaps = g.ntaps;
    for (int t = 0; t < g.ntaps; ++t) { p.dy[t] = g.dy[0][t]; p.dx[t] = g.dx[0][t]; }
    p.Cout = g.Cout; p.Cin0 = g.Cin0; p.Cin1 = g.Cin1; p.cin_tot = g.Cin0 + g.Cin1;
    p.Nn = pick_n(g);
    p.m_tiles = (g.Cout + 127) / 128;
    p.n_blocks = p.cin_tot / p.Nn;
    p.a_blocks = g.Cout >= 128 ? 2 : 1;
    int slices;
    wgrad_tc_partial_floats(g, num_sms, &slices);
    p.slices = slices;
    p.tiles_per_slice = (p.num_tiles + slices - 1) / slices;
    p.partial = partial;
    const int stage_bytes = (p.a_blocks + p.Nn / 64) * kBlockBytes;
    int stages = (227 * 1024 - 2048) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) { set_error("wgrad_tc: not enough shared memory for two stages"); delete pl; return nullptr; }
    p.stages = stages;
    pl->smem = (size_t)stages * stage_bytes + 1024 + 256;
    pl->grid = p.m_tiles * p.n_blocks * p.ntaps * p.slices;
    bool ok = true;
    auto make = [&](CUtensorMap* mp, const void* src, int C, int Hh, int Ww) {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)Ww, (uint64_t)Hh, (uint64_t)g.B};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Ww * C * 2, (uint64_t)Hh * Ww * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
        return encode_map(mp, src, 4, dims, str, box);
    };
    ok = ok && make(&pl->mapG, gout, g.Cout, g.Hout, g.Wout);
    ok = ok && make(&pl->mapX0, x0, g.Cin0, g.Hin, g.Win);
    if (g.Cin1 > 0) ok = ok && make(&pl->mapX1, x1, g.Cin1, g.Hin, g.Win);
    else pl->mapX1 = pl->mapX0;
    if (!ok) { delete pl; return nullptr; }
    return pl;
}

void wgrad_tc_plan_destroy(WgradTcPlan* p) { delete p; }
int wgrad_tc_slices(const WgradTcPlan* p) { return p->p.slices; }

namespace {
template <int N>
int launch_wt(const WgradTcPlan* pl, cudaStream_t s) {
    static bool attr = false;
    auto k = wgrad_tc_kernel<N>;
    if (!attr) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    k<<<pl->grid, kWtThreads, pl->smem, s>>>(pl->mapG, pl->mapX0, pl->mapX1, pl->p);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace

int wgrad_tc_launch(const WgradTcPlan* pl, cudaStream_t s) {
    if (pl->p.Nn == 256) return launch_wt<256>(pl, s);
    if (pl->p.Nn == 128) return launch_wt<128>(pl, s);
    return launch_wt<64>(pl, s);
}

}  // namespace gtts
