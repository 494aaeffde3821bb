// Alignment stage around MAS: the Gaussian log-prior that MAS maximises and the two tensors derived from
// the path.  Reference: model/tts.py:143-149 (log_prior), :155 (logw_), :184-185 (mu_y).  fp32 throughout, like the reference.
//
//   log_prior[b, i, j] = sum_c -0.5 y[b,c,j]^2  -  sum_c 2(-0.5 mu[b,c,i]) y[b,c,j]  +  sum_c -0.5 mu[b,c,i]^2  +  const
//   logw_[b, i]        = log(1e-8 + sum_j attn[b,i,j]) * x_mask[b,i]
//   mu_y[b, c, j]      = sum_i attn[b,i,j] mu[b,c,i]           (attn is a 0/1 path: a gather, exact)
#include "common.cuh"
#include "ops.h"

namespace gtts {
namespace {

constexpr int kLpTile = 64;          // 64 text rows x 64 mel frames per CTA, 4 x 4 outputs per thread

__global__ void __launch_bounds__(256)
log_prior_kernel(const float* __restrict__ mu, const float* __restrict__ y, float* __restrict__ out, int C, int tx, int ty,
                 float cst) {
    __shared__ float s_mu[16][kLpTile + 1];
    __shared__ float s_y[16][kLpTile + 1];
    const int b = blockIdx.z, i0 = blockIdx.y * kLpTile, j0 = blockIdx.x * kLpTile;
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;           // thread -> rows ti + 16a, columns tj + 16c
    const float* mub = mu + (size_t)b * C * tx;
    const float* yb = y + (size_t)b * C * ty;
    float dot[4][4], ysq[4], msq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        ysq[a] = 0.f; msq[a] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) dot[a][c] = 0.f;
    }
    for (int c0 = 0; c0 < C; c0 += 16) {
        for (int e = tid; e < 16 * kLpTile; e += 256) {
            const int cc = e / kLpTile, k = e - cc * kLpTile;
            const bool cv = c0 + cc < C;
            s_mu[cc][k] = (cv && i0 + k < tx) ? mub[(size_t)(c0 + cc) * tx + i0 + k] : 0.f;
            s_y[cc][k] = (cv && j0 + k < ty) ? yb[(size_t)(c0 + cc) * ty + j0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
            float m[4], v[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { m[a] = s_mu[cc][ti + 16 * a]; v[a] = s_y[cc][tj + 16 * a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                msq[a] = fmaf(m[a], m[a], msq[a]);
                ysq[a] = fmaf(v[a], v[a], ysq[a]);
#pragma unroll
                for (int c = 0; c < 4; ++c) dot[a][c] = fmaf(m[a], v[c], dot[a][c]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ti + 16 * a;
        if (i >= tx) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tj + 16 * c;
            // y_square - y_mu_double + mu_square + const with factor = -0.5  (tts.py:146-149)
            if (j < ty) out[((size_t)b * tx + i) * ty + j] = ((-0.5f * ysq[c]) - (-dot[a][c])) + (-0.5f * msq[a]) + cst;
        }
    }
}

// one warp per (b, text row): duration = number of frames on the path
__global__ void __launch_bounds__(256)
logw_kernel(const float* __restrict__ attn, const float* __restrict__ x_mask, float* __restrict__ logw, int rows, int ty) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* a = attn + (size_t)row * ty;
    float s = 0.f;
    for (int j = lane; j < ty; j += 32) s += a[j];
    s = warp_sum(s);
    if (lane == 0) logw[row] = logf(1e-8f + s) * x_mask[row];
}

// mu_y[b,c,j] = sum_i attn[b,i,j] mu[b,c,i]; thread per (b, j): finds the rows with attn != 0 (one for a hard path)
__global__ void __launch_bounds__(128)
mu_y_kernel(const float* __restrict__ attn, const float* __restrict__ mu, float* __restrict__ mu_y, int C, int tx, int ty) {
    const int b = blockIdx.y, j = blockIdx.x * 128 + threadIdx.x;
    if (j >= ty) return;
    const float* ab = attn + (size_t)b * tx * ty + j;
    const float* mub = mu + (size_t)b * C * tx;
    float* o = mu_y + (size_t)b * C * ty + j;
    for (int c = 0; c < C; ++c) o[(size_t)c * ty] = 0.f;
    for (int i = 0; i < tx; ++i) {
        const float w = ab[(size_t)i * ty];
        if (w != 0.f)
            for (int c = 0; c < C; ++c) o[(size_t)c * ty] = fmaf(w, mub[(size_t)c * tx + i], o[(size_t)c * ty]);
    }
}

}  // namespace

int align_log_prior(const float* mu_x, const float* y, float* log_prior, int B, int C, int tx, int ty, cudaStream_t s) {
    GTTS_REQUIRE(B >= 0 && C > 0 && tx >= 0 && ty >= 0, "align_log_prior: bad shape");
    if (B == 0 || tx == 0 || ty == 0) return 0;
    GTTS_REQUIRE(B <= 65535, "align_log_prior: batch too large");
    const float cst = (float)(-0.5 * 1.8378770664093453 * (double)C);        // -0.5 * log(2 pi) * n_feats   (tts.py:144)
    dim3 grid((ty + kLpTile - 1) / kLpTile, (tx + kLpTile - 1) / kLpTile, B);
    log_prior_kernel<<<grid, 256, 0, s>>>(mu_x, y, log_prior, C, tx, ty, cst);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int align_outputs(const float* attn, const float* mu_x, const float* x_mask, float* logw, float* mu_y, int B, int C, int tx,
                  int ty, cudaStream_t s) {
    if (B == 0 || tx == 0 || ty == 0) return 0;
    GTTS_REQUIRE(B <= 65535, "align_outputs: batch too large");
    if (logw) {
        const int rows = B * tx;
        logw_kernel<<<(rows + 7) / 8, 256, 0, s>>>(attn, x_mask, logw, rows, ty);
    }
    if (mu_y) mu_y_kernel<<<dim3((ty + 127) / 128, B), 128, 0, s>>>(attn, mu_x, mu_y, C, tx, ty);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
