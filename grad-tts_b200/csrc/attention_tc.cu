// Fused k-projection + softmax-over-positions + context accumulation on tcgen05 (C = 64), replacing the mma.sync kernel
// of attention.cu for that width (ncu: legacy HMMA runs at a quarter of the tcgen05 rate and was 52 % busy).
//
//   per 128-pixel tile:  K^T[128 d, 128 px] = Wk[128, C] . X^T          (GEMM1: A = Wk resident in SMEM, B = x tile, K-major)
//                        P = exp(K^T - m_ref)  (bf16, one TMEM lane = one k-channel d per thread, so max / sum are in-thread)
//                        S[128 d, C]  += P[128, 128 px] . X[128 px, C]  (GEMM2: A = P from SMEM, B = the SAME x tile read as an
//                                                                         MN-major operand: channels contiguous, 8-pixel groups 1 KB apart)
//   end of chunk:        ctx[d, e] = sum_c S[d, c] Wv[32 h(d) + e, c]   (one more MMA: S (bf16) . Wv^T, each head keeps its 32 columns)
//
// The accumulator S lives in TMEM for the whole chunk.  The running maximum is LAZY: exponentials are taken against a
// reference maximum m_ref that is only raised (and S, l rescaled through tcgen05.ld/st) when a tile's maximum exceeds it
// by more than kLazyTau -- p may then be as large as e^tau, harmless in bf16/fp32 -- so the steady state has no
// correction pass.  The (m, l, ctx) partials are consistent with whatever m_ref ended up as; attn_merge handles any m.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = softmax (lane quarter w of TMEM = head w).
#include <cuda.h>

#include "common.cuh"
#include "conv_tc_common.cuh"
#include "ops.h"

namespace gtts {

using namespace tc;

namespace {

constexpr int kAtTile = 128;                 // pixels per tile
constexpr int kAtStages = 3;
constexpr float kLazyTau = 8.0f;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B descriptor with an explicit leading byte offset (MN-major operands wider than one 64-element atom)
__device__ __forceinline__ uint64_t make_sw128_desc_lbo(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Persistent: grid = min(#SMs, B * chunks); a CTA walks work items (sample, chunk) = blockIdx.x + k * gridDim.x, so Wk / Wv,
// the TMEM allocation and the barriers are set up once per CTA and the tail imbalance is one item, not one CTA wave.
// Operands wider than 64 channels are stored as C/64 swizzle atoms of [128 rows][64 ch] (16 KB each).
template <int C>
__global__ void __launch_bounds__(256, 1)
attn_xk_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                  float* __restrict__ partials, int B, int n, int chunks, int chunk_len, float tau) {
    constexpr int kAtoms = C / 64;
    constexpr int kXBytes = kAtTile * 128 * kAtoms;         // 128 px x C ch bf16
    constexpr int kWBytes = 128 * 128 * kAtoms;             // 128 rows x C ch bf16
    constexpr int kPBytes = 128 * kAtTile * 2;              // 128 d x 128 px bf16 = 32 KB (two 16 KB K-atoms)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* s_wk = smem;                                   // kAtoms x [128 d][64 ch]  K-major SW128
    uint8_t* s_wv = s_wk + kWBytes;                         // kAtoms x [128 e'][64 ch]
    uint8_t* s_x = s_wv + kWBytes;                          // kAtStages x kAtoms x [128 px][64 ch]
    uint8_t* s_p = s_x + kAtStages * kXBytes;               // 2 x [128 d][128 px]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 2 * kPBytes);
    uint64_t* wfull = bars;                                 // [1]
    uint64_t* xfull = bars + 1;                             // [3]
    uint64_t* xempty = bars + 4;                            // [3]
    uint64_t* d1full = bars + 7;                            // [2]
    uint64_t* d1empty = bars + 9;                           // [2]
    uint64_t* pfull = bars + 11;                            // [2]
    uint64_t* pempty = bars + 13;                           // [2]
    uint64_t* sfull = bars + 15;                            // [1] S (bf16) staged for the Wv product
    uint64_t* ofull = bars + 16;                            // [1] Wv product done
    uint64_t* odone = bars + 17;                            // [1] its result has been read: D1[0] / P[0] reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_items = B * chunks;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapW);
        mbar_init(wfull, 1);
        for (int s = 0; s < kAtStages; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d1full[i], 1); mbar_init(&d1empty[i], 4); mbar_init(&pfull[i], 4); mbar_init(&pempty[i], 1); }
        mbar_init(sfull, 4);
        mbar_init(ofull, 1);
        mbar_init(odone, 4);
        mbar_fence_init();
    } else if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_d1 = tmem, t_s = tmem + 256;           // D1[0] cols 0..127, D1[1] cols 128..255, S cols 256..256+C

    auto item_range = [&](int item, int& b, int& chunk, int& n0, int& nt, int& n1) {
        b = item / chunks; chunk = item - b * chunks;
        n0 = chunk * chunk_len; n1 = min(n, n0 + chunk_len);
        nt = n1 > n0 ? (n1 - n0 + kAtTile - 1) / kAtTile : 0;
    };

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wfull, (uint32_t)(2 * kWBytes));
#pragma unroll
            for (int a = 0; a < kAtoms; ++a) {
                tma_load_2d(&mapW, wfull, s_wk + a * 16384, a * 64, 0);          // k rows
                tma_load_2d(&mapW, wfull, s_wv + a * 16384, a * 64, 128);        // v rows
            }
            int g = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int b, chunk, n0, nt, n1;
                item_range(item, b, chunk, n0, nt, n1);
                for (int i = 0; i < nt; ++i, ++g) {
                    const int s = g % kAtStages;
                    mbar_wait(&xempty[s], (uint32_t)((g / kAtStages) & 1) ^ 1u);
                    mbar_expect_tx(&xfull[s], (uint32_t)kXBytes);
#pragma unroll
                    for (int a = 0; a < kAtoms; ++a)                             // rows past the tensor: zeros
                        tma_load_2d(&mapX, &xfull[s], s_x + s * kXBytes + a * 16384, a * 64, b * n + n0 + i * kAtTile);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer, per item: G1(0); { G1(i+1); G2(i) }...; then S(bf16) . Wv^T
        constexpr uint32_t kIdG1 = make_idesc<128>();                       // M128 N128, A/B K-major
        constexpr uint32_t kIdG2 = make_idesc<C>() | (1u << 16);            // M128 N=C, B MN-major
        const uint32_t wk_addr = smem_u32(s_wk), wv_addr = smem_u32(s_wv);
        auto g1 = [&](int g) {
            const int s = g % kAtStages, bb = g & 1;
            mbar_wait(&xfull[s], (uint32_t)((g / kAtStages) & 1));
            mbar_wait(&d1empty[bb], (uint32_t)((g >> 1) & 1) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t xa = smem_u32(s_x + s * kXBytes);
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                    tc_mma_f16(t_d1 + (uint32_t)(bb * 128), make_sw128_kmajor_desc(wk_addr + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3)),
                               make_sw128_kmajor_desc(xa + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3)), kIdG1, (uint32_t)(k != 0));
                tc_commit(&d1full[bb]);
            }
            __syncwarp();
        };
        auto g2 = [&](int g, int i) {
            const int s = g % kAtStages, bb = g & 1;
            mbar_wait(&pfull[bb], (uint32_t)((g >> 1) & 1));
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kAtTile / 16; ++kk) {
                    const uint64_t pd = make_sw128_kmajor_desc(smem_u32(s_p + bb * kPBytes + (kk >> 2) * 16384)) + (uint64_t)(2 * (kk & 3));
                    // x tile as an MN-major operand: 8-pixel groups 1 KB apart (SBO), 64-channel atoms 16 KB apart (LBO)
                    const uint64_t xd = make_sw128_desc_lbo(smem_u32(s_x + s * kXBytes + kk * 2048), 16384u, 1024u);
                    tc_mma_f16(t_s, pd, xd, kIdG2, (uint32_t)((i | kk) != 0));
                }
                tc_commit(&pempty[bb]);
                tc_commit(&xempty[s]);
            }
            __syncwarp();
        };
        mbar_wait(wfull, 0u);
        int g = 0, it_n = 0;                                                // it_n: items this CTA has processed
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int b, chunk, n0, nt, n1;
            item_range(item, b, chunk, n0, nt, n1);
            if (nt == 0) continue;
            if (it_n > 0) mbar_wait(odone, (uint32_t)((it_n - 1) & 1));     // previous item's result has left D1[0]
            g1(g);
            for (int i = 0; i < nt; ++i) {
                if (i + 1 < nt) g1(g + i + 1);
                g2(g + i, i);
            }
            g += nt;
            mbar_wait(sfull, (uint32_t)(it_n & 1));
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(s_p);
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                    tc_mma_f16(t_d1, make_sw128_kmajor_desc(sa + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3)),
                               make_sw128_kmajor_desc(wv_addr + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3)), kIdG1, (uint32_t)(k != 0));
                tc_commit(ofull);
            }
            __syncwarp();
            ++it_n;
        }
    } else if (warp >= 4) {
        // ---- softmax: thread = k-channel d = 32 (warp-4) + lane = TMEM lane
        const int wq = warp - 4, d = wq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        int g = 0, it_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int b, chunk, n0, nt, n1;
            item_range(item, b, chunk, n0, nt, n1);
            if (nt == 0) {                                                     // empty chunk: neutral element of the merge
                float* part = partials + (((size_t)b * 4 + wq) * chunks + chunk) * 1088;
                part[lane] = -INFINITY;
                part[32 + lane] = 0.f;
                for (int e = 0; e < 32; ++e) part[64 + lane * 32 + e] = 0.f;
                continue;
            }
            float m_ref = -INFINITY, l = 0.f;
            for (int i = 0; i < nt; ++i, ++g) {
                const int bb = g & 1;
                const int nvalid = n1 - (n0 + i * kAtTile);                   // pixels of this tile inside the chunk
                mbar_wait(&d1full[bb], (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                const uint32_t td = t_d1 + lane_base + (uint32_t)(bb * 128);
                // one TMEM read of my whole row (128 k values stay in registers), then the buffer goes back to the MMA warp
                uint32_t r[kAtTile];
#pragma unroll
                for (int c0 = 0; c0 < kAtTile; c0 += 32) tmem_ld32(td + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&r[c0]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d1empty[bb]);
                if (nvalid < kAtTile) {                                      // ragged last tile: pixels past the chunk never count
#pragma unroll
                    for (int q = 0; q < kAtTile; ++q) if (q >= nvalid) r[q] = 0xff800000u;   // -inf
                }
                float mx = -INFINITY;
#pragma unroll
                for (int q = 0; q < kAtTile; q += 4)
                    mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(r[q]), __uint_as_float(r[q + 1])),
                                         fmaxf(__uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]))));
                // lazy reference maximum: raise it (and rescale S, l) only when the tile exceeds it by more than tau
                const bool raise = mx > m_ref + tau;                          // first tile: m_ref = -inf -> true
                float factor = 1.0f;
                if (raise) { factor = (i == 0) ? 0.f : __expf(m_ref - mx); m_ref = mx; }
                if (i > 0 && __any_sync(0xffffffffu, raise)) {
                    // GEMM2 of the previous tile must have retired before S is touched
                    mbar_wait(&pempty[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = 0; c0 < C; c0 += 32) {
                        uint32_t sr[32];
                        tmem_ld32(t_s + lane_base + (uint32_t)c0, sr);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; ++q) sr[q] = __float_as_uint(__uint_as_float(sr[q]) * factor);
                        tmem_st32(t_s + lane_base + (uint32_t)c0, sr);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                }
                l *= factor;
                // P buffer bb free again? (GEMM2 two tiles back retired)
                mbar_wait(&pempty[bb], (uint32_t)((g >> 1) & 1) ^ 1u);
                // p = 2^(k*log2e - m_ref*log2e) -> bf16 -> SMEM (K-major, 128-byte swizzle); row sum of the ROUNDED values.
                // exp(-inf) = 0 takes care of the masked pixels.
                uint8_t* prow = s_p + bb * kPBytes + d * 128;
                // (k - m_ref) FIRST, then * log2(e): with the product k*log2(e) - m_ref*log2(e) the two roundings differ by up to an
                // ulp of the product, which for |k| ~ 1e17 (diverging random-weight runs) is 1e10 in the exponent -> inf -> NaN.
                const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
                const float2 nm = make_float2(-m_ref, -m_ref);
                float2 ls2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c16 = 0; c16 < kAtTile / 8; ++c16) {                 // 16-byte chunks of 8 pixels
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 tq = fmul2(fadd2(make_float2(__uint_as_float(r[c16 * 8 + 2 * q]), __uint_as_float(r[c16 * 8 + 2 * q + 1])), nm), l2e);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(ex2_approx(tq.x), ex2_approx(tq.y));
                        w[q] = *reinterpret_cast<uint32_t*>(&h2);
                        ls2 = fadd2(ls2, make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u)));
                    }
                    uint8_t* dst = prow + (c16 >> 3) * 16384 + (((c16 & 7) ^ (d & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                l += ls2.x + ls2.y;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&pfull[bb]);
            }
            // ---- end of chunk: S -> bf16 A operand (C/64 atoms in P[0]), one MMA with Wv, then my head's 32 columns
            mbar_wait(&pempty[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
            tc_fence_after();
            {
                uint8_t* srow = s_p + d * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += 32) {
                    uint32_t sr[32];
                    tmem_ld32(t_s + lane_base + (uint32_t)c0, sr);
                    tmem_ld_wait();
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(sr[c16 * 8 + 2 * q]), __uint_as_float(sr[c16 * 8 + 2 * q + 1]));
                            w[q] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        const int ch = (c0 >> 3) + c16;                       // 16-byte chunk along the C channels
                        *reinterpret_cast<uint4*>(srow + (ch >> 3) * 16384 + (((ch & 7) ^ (d & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(sfull);
            }
            mbar_wait(ofull, (uint32_t)(it_n & 1));
            tc_fence_after();
            {
                uint32_t o[32];
                tmem_ld32(t_d1 + lane_base + (uint32_t)(wq * 32), o);         // columns of my head
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(odone);
                float* part = partials + (((size_t)b * 4 + wq) * chunks + chunk) * 1088;
                part[lane] = m_ref;
                part[32 + lane] = l;
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4)
                    *reinterpret_cast<float4*>(&part[64 + lane * 32 + e4 * 4]) =
                        make_float4(__uint_as_float(o[4 * e4]), __uint_as_float(o[4 * e4 + 1]), __uint_as_float(o[4 * e4 + 2]),
                                    __uint_as_float(o[4 * e4 + 3]));
            }
            ++it_n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int C>
int launch_attn_tc(const CUtensorMap& mapX, const CUtensorMap& mapW, float* partials, int B, int n, int chunks, int chunk_len,
                   cudaStream_t s) {
    const size_t smem = (size_t)2 * 128 * 128 * (C / 64) + (size_t)kAtStages * kAtTile * 128 * (C / 64) + 2 * 128 * kAtTile * 2 + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(attn_xk_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    float tau = kLazyTau;
    if (const char* t = getenv("GTTS_ATTN_TAU")) tau = (float)atof(t);        // 0: rescale on every increase (exercises the correction path)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = B * chunks;
    GTTS_CHECK_CUDA(launch_pdl(attn_xk_tc_kernel<C>, dim3(items < sms ? items : sms), dim3(256), smem, s, 1, mapX, mapW, partials, B, n,
                               chunks, chunk_len, tau));
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// x: [B*n][C] bf16 (NHWC activations flattened), wkv: [256][C] bf16 (k rows, then v rows); C = 64 or 128
int attn_xk_tc(const void* x, const void* wkv_bf16, float* partials, int B, int n, int C, int chunks, int chunk_len, cudaStream_t s) {
    GTTS_REQUIRE(C == 64 || C == 128, "attn_xk_tc: C must be 64 or 128");
    CUtensorMap mapX, mapW;
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * (uint64_t)n};
        const uint64_t str[1] = {(uint64_t)C * 2};
        const uint32_t box[2] = {64, (uint32_t)kAtTile};
        if (!encode_map(&mapX, x, 2, dims, str, box)) return 1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, 256};
        const uint64_t str[1] = {(uint64_t)C * 2};
        const uint32_t box[2] = {64, 128};
        if (!encode_map(&mapW, wkv_bf16, 2, dims, str, box)) return 1;
    }
    return C == 64 ? launch_attn_tc<64>(mapX, mapW, partials, B, n, chunks, chunk_len, s)
                   : launch_attn_tc<128>(mapX, mapW, partials, B, n, chunks, chunk_len, s);
}

}  // namespace gtts
