// K10 -- Monotonic Alignment Search on device.
//
// Replaces the host round trip of /root/reference/model/monotonic_align/__init__.py:8-23 and the Cython
// DP of core.pyx:9-45.  One CTA per utterance, one thread per text row, a column sweep over the mel axis
// (every row of column y depends only on column y-1, core.pyx:21-30).  Instead of the 4*t_x*t_y-byte DP
// table the kernel keeps one *direction bit* per cell -- the predicate (v_prev > v_cur) of the forward
// max, which is the same comparison on the same floats that the backtrack evaluates at core.pyx:34 --
// so the result is bit-exact while the table fits in shared memory (200x1000 cells -> 28 KB).
//
// Arithmetic contract: fp32 add / compare only, no FMA, max lowered as (v_prev > v_cur) ? v_prev : v_cur.
#include <cstdlib>

#include "common.cuh"
#include "ops.h"

namespace gtts {

namespace {

constexpr int kGroup = 8;   // columns fetched per thread per prefetch group (two float4)

template <int R, bool kHasMask>
__device__ __forceinline__ void load_group(const float* __restrict__ vrow[R], const float* __restrict__ mrow[R],
                                           const bool (&rowok)[R], int y0, int t_y, bool vec_ok,
                                           float (&dst)[R][kGroup]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!rowok[r] || y0 >= t_y) {
#pragma unroll
            for (int j = 0; j < kGroup; ++j) dst[r][j] = 0.f;
            continue;
        }
        if (vec_ok && y0 + kGroup <= t_y) {
            float4 a = __ldg(reinterpret_cast<const float4*>(vrow[r] + y0));
            float4 b = __ldg(reinterpret_cast<const float4*>(vrow[r] + y0 + 4));
            dst[r][0] = a.x; dst[r][1] = a.y; dst[r][2] = a.z; dst[r][3] = a.w;
            dst[r][4] = b.x; dst[r][5] = b.y; dst[r][6] = b.z; dst[r][7] = b.w;
            if (kHasMask) {
                float4 c = __ldg(reinterpret_cast<const float4*>(mrow[r] + y0));
                float4 d = __ldg(reinterpret_cast<const float4*>(mrow[r] + y0 + 4));
                dst[r][0] = __fmul_rn(dst[r][0], c.x); dst[r][1] = __fmul_rn(dst[r][1], c.y);
                dst[r][2] = __fmul_rn(dst[r][2], c.z); dst[r][3] = __fmul_rn(dst[r][3], c.w);
                dst[r][4] = __fmul_rn(dst[r][4], d.x); dst[r][5] = __fmul_rn(dst[r][5], d.y);
                dst[r][6] = __fmul_rn(dst[r][6], d.z); dst[r][7] = __fmul_rn(dst[r][7], d.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                float v = 0.f;
                if (y0 + j < t_y) {
                    v = __ldg(vrow[r] + y0 + j);
                    if (kHasMask) v = __fmul_rn(v, __ldg(mrow[r] + y0 + j));
                }
                dst[r][j] = v;
            }
        }
    }
}

// bits: one uint32 word per (column y, warp-of-rows w): bit l = direction of row (rr*nthreads + 32*w + l).
template <int R, bool kHasMask, typename PathT>
__global__ void __launch_bounds__(1024, 1)
mas_kernel(const float* __restrict__ value, const float* __restrict__ mask, const int* __restrict__ t_xs,
           const int* __restrict__ t_ys, PathT* __restrict__ path, int tx_max, int ty_max, float max_neg,
           uint32_t* __restrict__ bits_global, size_t bits_words_per_item, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nth >> 5;
    const int rows_cap = nth * R;                 // >= tx_max
    const int words_per_col = rows_cap / 32;

    float* vcol = reinterpret_cast<float*>(smem_raw);                 // [2][rows_cap + 1]
    uint32_t* bits = bits_global ? bits_global + (size_t)b * bits_words_per_item
                                 : reinterpret_cast<uint32_t*>(vcol + 2 * (rows_cap + 1) + 3);
    bits = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(bits) + 3) & ~uintptr_t(3));
    __shared__ int s_len[2];
    __shared__ float s_red[64];

    const float* vb = value + (size_t)b * tx_max * ty_max;
    const float* mb = kHasMask ? mask + (size_t)b * tx_max * ty_max : nullptr;

    // ---- lengths: t_x = sum_x mask[b,x,0], t_y = sum_y mask[b,0,y]   (__init__.py:20-21), or given
    int t_x, t_y;
    if (t_xs != nullptr) {
        t_x = t_xs[b];
        t_y = t_ys[b];
    } else {
        float sx = 0.f, sy = 0.f;
        for (int x = tid; x < tx_max; x += nth) sx += mb[(size_t)x * ty_max];
        for (int y = tid; y < ty_max; y += nth) sy += mb[y];
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        if (lane == 0) { s_red[warp] = sx; s_red[32 + warp] = sy; }
        __syncthreads();
        if (tid == 0) {
            float ax = 0.f, ay = 0.f;
            for (int w = 0; w < nwarp; ++w) { ax += s_red[w]; ay += s_red[32 + w]; }
            s_len[0] = (int)ax;
            s_len[1] = (int)ay;
        }
        __syncthreads();
        t_x = s_len[0];
        t_y = s_len[1];
    }
    // Degenerate inputs: the reference has an empty band for t_x > t_y and indexes out of bounds;
    // here they are reported (status=1) and the path for that item stays all-zero.
    if (t_x > t_y || t_x < 0 || t_y < 0 || t_x > tx_max || t_y > ty_max) {
        if (tid == 0) atomicMax(status, 1);
        return;
    }
    if (t_x == 0 || t_y == 0) return;   // reference loops are empty apart from an OOB write; keep zeros

    const float* vrow[R];
    const float* mrow[R];
    bool rowok[R];
    float myv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int x = r * nth + tid;
        rowok[r] = x < t_x;
        vrow[r] = vb + (size_t)(rowok[r] ? x : 0) * ty_max;
        mrow[r] = kHasMask ? mb + (size_t)(rowok[r] ? x : 0) * ty_max : nullptr;
        myv[r] = 0.f;
    }
    const bool vec_ok = (ty_max % 4 == 0) && ((reinterpret_cast<uintptr_t>(vb) & 15) == 0) &&
                        (!kHasMask || (reinterpret_cast<uintptr_t>(mb) & 15) == 0);

    float cur[R][kGroup], nxt[R][kGroup];
    load_group<R, kHasMask>(vrow, mrow, rowok, 0, t_y, vec_ok, cur);

    int pb = 0;   // vcol buffer holding column y-1
    for (int y0 = 0; y0 < t_y; y0 += kGroup) {
        load_group<R, kHasMask>(vrow, mrow, rowok, y0 + kGroup, t_y, vec_ok, nxt);
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            const int y = y0 + j;
            if (y < t_y) {                                         // uniform across the CTA
                const int lo = max(0, t_x + y - t_y), hi = min(t_x, y + 1);
                const float* vp = vcol + pb * (rows_cap + 1);
                float* vn = vcol + (pb ^ 1) * (rows_cap + 1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int x = r * nth + tid;
                    const bool active = (x >= lo) && (x < hi);
                    float v_cur = (x == y) ? max_neg : myv[r];
                    float v_prev = (x == 0) ? (y == 0 ? 0.f : max_neg) : vp[active ? x - 1 : 0];
                    const bool take = active && (v_prev > v_cur);
                    float nv = __fadd_rn(take ? v_prev : v_cur, cur[r][j]);
                    if (active) { myv[r] = nv; vn[x] = nv; }
                    uint32_t word = __ballot_sync(0xffffffffu, take);
                    if (lane == 0) bits[(size_t)y * words_per_col + r * nwarp + warp] = word;
                }
                pb ^= 1;
                __syncthreads();
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < kGroup; ++j) cur[r][j] = nxt[r][j];
    }

    // ---- backtrack (core.pyx:32-35), one thread; path was zero-filled by the host wrapper
    if (tid == 0) {
        if (bits_global) __threadfence();
        PathT* pout = path + (size_t)b * tx_max * ty_max;
        int index = t_x - 1;
        for (int y = t_y - 1; y >= 0; --y) {
            pout[(size_t)index * ty_max + y] = (PathT)1;
            if (index != 0) {
                bool dec = (index == y);
                if (!dec) {
                    const int rr = index / nth, t = index - rr * nth;
                    uint32_t word = bits[(size_t)y * words_per_col + rr * nwarp + (t >> 5)];
                    dec = (word >> (t & 31)) & 1u;
                }
                if (dec) --index;
            }
        }
    }
}

// ---- warp-serial variant (the default when the problem fits it) ------------------------------------------------------------
// The column sweep above pays one block barrier per mel frame (~0.37 us x 1000 columns at the C2 shape, 0.04 of the HBM roof).
// Here ONE warp carries the whole DP column in registers -- lane l owns the R consecutive text rows l*R .. l*R + R-1, so the
// "row above" is a register of the same lane except for one shuffle per column -- and never meets a barrier inside a tile of 32
// columns; the other seven warps stage value (* mask) tiles of 32 columns into shared memory (coalesced along the mel axis,
// stored [column][row] with an odd lane stride so the DP warp's reads are conflict-free), one tile ahead.  Direction bits: R bits
// per lane per column in shared memory.  Same arithmetic contract as mas_kernel (fp32 add / compare, the same predicate for the
// forward max and the backtrack), hence the same bits.
template <int R, int kCols, typename BitsT, bool kHasMask, typename PathT>
__global__ void __launch_bounds__(256, 1)
mas_warp_kernel(const float* __restrict__ value, const float* __restrict__ mask, const int* __restrict__ t_xs,
                const int* __restrict__ t_ys, PathT* __restrict__ path, int tx_max, int ty_max, float max_neg,
                int* __restrict__ status) {
    constexpr int kRows = 32 * R, kPitch = kRows + 1, kC4 = kCols / 4;   // kCols columns per tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                // [2][kCols][kPitch]
    BitsT* bits = reinterpret_cast<BitsT*>(tiles + 2 * kCols * kPitch);   // [ty_max][32]
    __shared__ int s_len[2];
    __shared__ float s_red[16];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* vb = value + (size_t)b * tx_max * ty_max;
    const float* mb = kHasMask ? mask + (size_t)b * tx_max * ty_max : nullptr;

    int t_x, t_y;
    if (t_xs != nullptr) {
        t_x = t_xs[b];
        t_y = t_ys[b];
    } else {
        float sx = 0.f, sy = 0.f;
        for (int x = tid; x < tx_max; x += 256) sx += mb[(size_t)x * ty_max];
        for (int y = tid; y < ty_max; y += 256) sy += mb[y];
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        if (lane == 0) { s_red[warp] = sx; s_red[8 + warp] = sy; }
        __syncthreads();
        if (tid == 0) {
            float ax = 0.f, ay = 0.f;
            for (int w = 0; w < 8; ++w) { ax += s_red[w]; ay += s_red[8 + w]; }
            s_len[0] = (int)ax;
            s_len[1] = (int)ay;
        }
        __syncthreads();
        t_x = s_len[0];
        t_y = s_len[1];
    }
    if (t_x > t_y || t_x < 0 || t_y < 0 || t_x > tx_max || t_y > ty_max) {
        if (tid == 0) atomicMax(status, 1);
        return;
    }
    if (t_x == 0 || t_y == 0) return;

    const bool vec_ok = (ty_max % 4 == 0) && ((reinterpret_cast<uintptr_t>(vb) & 15) == 0) &&
                        (!kHasMask || (reinterpret_cast<uintptr_t>(mb) & 15) == 0);
    // stage tile k (columns k*kCols .., rows 0 .. t_x-1) into buffer k & 1; threads first .. first+nthr-1 cooperate.  Eight
    // 16-byte loads (and their mask loads) per thread are in flight before the first store: the first version issued them one
    // by one and the DP warp spent its time waiting for the tile (0.27 ms at C2).
    auto stage = [&](int k, int first, int nthr) {
        float* dst = tiles + (size_t)(k & 1) * kCols * kPitch;
        const int y0 = k * kCols;
        const int t = tid - first;
        const int n_items = t_x * kC4;
        for (int i0 = t; i0 < n_items; i0 += 8 * nthr) {
            float4 a[8], m[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * nthr;
                a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (i < n_items) {
                    const int x = i / kC4, c4 = (i - x * kC4) * 4, y = y0 + c4;
                    const float* vp = vb + (size_t)x * ty_max + y;
                    if (vec_ok && y + 4 <= t_y) {
                        a[u] = __ldg(reinterpret_cast<const float4*>(vp));
                        if (kHasMask) m[u] = __ldg(reinterpret_cast<const float4*>(mb + (size_t)x * ty_max + y));
                    } else {
                        float v[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (y + q < t_y) {
                                v[q] = __ldg(vp + q);
                                if (kHasMask) w[q] = __ldg(mb + (size_t)x * ty_max + y + q);
                            }
                        a[u] = make_float4(v[0], v[1], v[2], v[3]);
                        m[u] = make_float4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * nthr;
                if (i < n_items) {
                    const int x = i / kC4, c4 = (i - x * kC4) * 4;
                    float v[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                    if (kHasMask) {
                        v[0] = __fmul_rn(v[0], m[u].x); v[1] = __fmul_rn(v[1], m[u].y);
                        v[2] = __fmul_rn(v[2], m[u].z); v[3] = __fmul_rn(v[3], m[u].w);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[(c4 + q) * kPitch + x] = v[q];
                }
            }
        }
    };
    const int ntiles = (t_y + kCols - 1) / kCols;
    stage(0, 0, 256);
    __syncthreads();
    float myv[R];
    bool rowok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { myv[r] = 0.f; rowok[r] = lane * R + r < t_x; }
    for (int k = 0; k < ntiles; ++k) {
        if (warp != 0) {
            if (k + 1 < ntiles) stage(k + 1, 32, 224);
        } else {
            const float* tile = tiles + (size_t)(k & 1) * kCols * kPitch;
            const int y_end = min(kCols, t_y - k * kCols);
            // row x is active in column y  <=>  0 <= y - x <= t_y - t_x  (and x < t_x): one subtraction and one unsigned compare
            const unsigned span = (unsigned)(t_y - t_x);
            for (int c = 0; c < y_end; ++c) {                          // (unrolling this loop by 4 measured 20 % slower)
                const int y = k * kCols + c;
                const float* col = tile + c * kPitch + lane * R;
                // the row above this lane's first row: the previous lane's last row as it was after column y-1; above row 0 the
                // reference has 0 in the first column and max_neg afterwards (core.pyx:24-27)
                float up = __shfl_up_sync(0xffffffffu, myv[R - 1], 1);
                if (lane == 0) up = (y == 0) ? 0.f : max_neg;
                const int d0 = y - lane * R;
                uint32_t word = 0u;
#pragma unroll
                for (int r = R - 1; r >= 0; --r) {                        // descending: myv[r-1] is still column y-1's value
                    const int d = d0 - r;                                  // y - x
                    const bool active = rowok[r] && ((unsigned)d <= span);
                    const float v_cur = (d == 0) ? max_neg : myv[r];
                    const float v_prev = (r == 0) ? up : myv[r - 1];
                    const bool take = active && (v_prev > v_cur);
                    const float nv = __fadd_rn(take ? v_prev : v_cur, col[r]);
                    myv[r] = active ? nv : myv[r];
                    if (take) word |= 1u << r;
                }
                bits[(size_t)y * 32 + lane] = (BitsT)word;
            }
        }
        __syncthreads();
    }
    // ---- backtrack (core.pyx:32-35), one thread; path was zero-filled by the host wrapper
    if (tid == 0) {
        PathT* pout = path + (size_t)b * tx_max * ty_max;
        int index = t_x - 1;
        for (int y = t_y - 1; y >= 0; --y) {
            pout[(size_t)index * ty_max + y] = (PathT)1;
            if (index != 0) {
                bool dec = (index == y);
                if (!dec) {
                    const int l = index / R, r = index - l * R;
                    dec = ((uint32_t)bits[(size_t)y * 32 + l] >> r) & 1u;
                }
                if (dec) --index;
            }
        }
    }
}

template <int R, int kCols, typename BitsT, typename PathT>
int launch_mas_warp(const float* value, const float* mask, const int* t_xs, const int* t_ys, PathT* path, int B, int tx, int ty,
                    float max_neg, int* status, cudaStream_t stream, bool* done) {
    const size_t smem = (size_t)2 * kCols * (32 * R + 1) * 4 + (size_t)ty * 32 * sizeof(BitsT);
    *done = false;
    if (smem > 200 * 1024) return 0;
    if (mask) {
        auto k = mas_warp_kernel<R, kCols, BitsT, true, PathT>;
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 256, smem, stream>>>(value, mask, t_xs, t_ys, path, tx, ty, max_neg, status);
    } else {
        auto k = mas_warp_kernel<R, kCols, BitsT, false, PathT>;
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 256, smem, stream>>>(value, mask, t_xs, t_ys, path, tx, ty, max_neg, status);
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    *done = true;
    return 0;
}

template <typename PathT>
int launch_mas(const float* value, const float* mask, const int* t_xs, const int* t_ys, PathT* path, int B,
               int tx, int ty, float max_neg, uint32_t* bits_ws, size_t bits_ws_bytes, int* status,
               cudaStream_t stream) {
    GTTS_REQUIRE(B >= 0 && tx >= 1 && ty >= 1, "maximum_path: bad shape");
    GTTS_REQUIRE(tx <= 4096, "maximum_path: t_x > 4096 is not supported");
    GTTS_REQUIRE(mask != nullptr || t_xs != nullptr, "maximum_path: need a mask or explicit lengths");
    if (B == 0) return 0;
    if (!getenv("GTTS_MAS_BLOCK")) {                        // warp-serial variant when rows and direction bits fit (GTTS_MAS_BLOCK=1: off)
        bool done = false;
        int rc = 0;
        if (tx <= 32 * 7) {
            rc = launch_mas_warp<7, 64, uint8_t, PathT>(value, mask, t_xs, t_ys, path, B, tx, ty, max_neg, status, stream, &done);
            if (!rc && !done) rc = launch_mas_warp<7, 32, uint8_t, PathT>(value, mask, t_xs, t_ys, path, B, tx, ty, max_neg, status, stream, &done);
        } else if (tx <= 32 * 13) rc = launch_mas_warp<13, 32, uint16_t, PathT>(value, mask, t_xs, t_ys, path, B, tx, ty, max_neg, status, stream, &done);
        if (rc) return rc;
        if (done) return 0;
    }
    const int R = tx <= 1024 ? 1 : 4;
    int nth = ((tx + R - 1) / R + 31) / 32 * 32;
    const int rows_cap = nth * R;
    size_t vcol_bytes = (size_t)2 * (rows_cap + 1) * 4 + 16;
    size_t bits_bytes = (size_t)ty * (rows_cap / 32) * 4;
    size_t smem = vcol_bytes + bits_bytes;
    uint32_t* bits_global = nullptr;
    size_t words_per_item = 0;
    const size_t kSmemMax = 200 * 1024;
    if (smem > kSmemMax) {
        words_per_item = bits_bytes / 4;
        GTTS_REQUIRE(bits_ws != nullptr && bits_ws_bytes >= (size_t)B * bits_bytes,
                     "maximum_path: direction-bit workspace too small for this t_x*t_y");
        bits_global = bits_ws;
        smem = vcol_bytes;
    }
#define GTTS_MAS_LAUNCH(RR, HM)                                                                          \
    do {                                                                                                 \
        auto k = mas_kernel<RR, HM, PathT>;                                                              \
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<B, nth, smem, stream>>>(value, mask, t_xs, t_ys, path, tx, ty, max_neg, bits_global,         \
                                    words_per_item, status);                                             \
    } while (0)
    if (R == 1) { if (mask) GTTS_MAS_LAUNCH(1, true); else GTTS_MAS_LAUNCH(1, false); }
    else        { if (mask) GTTS_MAS_LAUNCH(4, true); else GTTS_MAS_LAUNCH(4, false); }
#undef GTTS_MAS_LAUNCH
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

size_t mas_bits_workspace_bytes(int B, int tx, int ty) {
    const int R = tx <= 1024 ? 1 : 4;
    int nth = ((tx + R - 1) / R + 31) / 32 * 32;
    size_t bits_bytes = (size_t)ty * (nth * R / 32) * 4;
    size_t smem = (size_t)2 * (nth * R + 1) * 4 + 16 + bits_bytes;
    return smem > 200 * 1024 ? (size_t)B * bits_bytes : 0;
}

int mas_forward_f32(const float* value, const float* mask, const int* t_xs, const int* t_ys, float* path, int B,
                    int tx, int ty, float max_neg, uint32_t* bits_ws, size_t bits_ws_bytes, int* status,
                    cudaStream_t stream) {
    return launch_mas<float>(value, mask, t_xs, t_ys, path, B, tx, ty, max_neg, bits_ws, bits_ws_bytes, status,
                             stream);
}
int mas_forward_i32(const float* value, const float* mask, const int* t_xs, const int* t_ys, int32_t* path, int B,
                    int tx, int ty, float max_neg, uint32_t* bits_ws, size_t bits_ws_bytes, int* status,
                    cudaStream_t stream) {
    return launch_mas<int32_t>(value, mask, t_xs, t_ys, path, B, tx, ty, max_neg, bits_ws, bits_ws_bytes, status,
                               stream);
}

}  // namespace gtts
