// Forward half of the training objective (forward value only; the estimator backward is not built).
// Reference: model/diffusion.py:244-252 (forward_diffusion) and :274-281 (loss_t).  fp32 elementwise, double accumulation.
//
//   c_b   = beta_min t_b + 0.5 (beta_max - beta_min) t_b^2                      get_noise(..., cumulative=True), :219-224
//   xt    = (x0 e^{-c/2} + mu (1 - e^{-c/2}) + z sqrt(1 - e^{-c})) mask          :247-252
//   zm    = z mask
//   loss  = sum_{b,c,j} (est sqrt(1 - e^{-c_b}) + zm)^2 / (sum(mask) n_feats)    :278-280
//
// HBM-bound streams: forward_diffusion reads 3 and writes 2 floats per element (20 B), the loss reads 2 (8 B).  The loss
// reduction has a fixed shape (per-thread double, warp shuffle tree, per-CTA slot, one finishing CTA that adds the slots in
// order), so a given input always gives the same bits.
#include "common.cuh"
#include "ops.h"

namespace gtts {
namespace {

constexpr int kLossCtas = 592;                // 4 per SM: enough loads in flight for an 8 B/element stream

__device__ __forceinline__ float cum_noise(float t, float beta_min, float beta_max) {
    return beta_min * t + 0.5f * (beta_max - beta_min) * (t * t);
}

// one thread per 4 consecutive frames of one (b, c) row; T % 4 == 0 keeps the float4 accesses aligned
__global__ void __launch_bounds__(256)
forward_diffusion_kernel(const float4* __restrict__ x0, const float* __restrict__ mask, const float4* __restrict__ mu,
                         const float* __restrict__ t, const float4* __restrict__ z, float4* __restrict__ xt,
                         float4* __restrict__ zm, int C, int T4, long total4, float beta_min, float beta_max) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total4; i += (long)gridDim.x * 256) {
        const long row = i / T4;
        const int j4 = (int)(i - row * T4), b = (int)(row / C);
        const float c = cum_noise(t[b], beta_min, beta_max);
        const float a = expf(-0.5f * c), s = sqrtf(1.0f - expf(-c));
        const float4 m = reinterpret_cast<const float4*>(mask + (size_t)b * T4 * 4)[j4];
        const float4 x = x0[i], u = mu[i], n = z[i];
        float4 o, q;
        o.x = (x.x * a + u.x * (1.0f - a) + n.x * s) * m.x; q.x = n.x * m.x;
        o.y = (x.y * a + u.y * (1.0f - a) + n.y * s) * m.y; q.y = n.y * m.y;
        o.z = (x.z * a + u.z * (1.0f - a) + n.z * s) * m.z; q.z = n.z * m.z;
        o.w = (x.w * a + u.w * (1.0f - a) + n.w * s) * m.w; q.w = n.w * m.w;
        xt[i] = o;
        zm[i] = q;
    }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; ++w) r += s_red[w];
    }
    __syncthreads();
    return r;                                  // valid in thread 0
}

// ws: [0, kLossCtas) squared-error slots, [kLossCtas, 2 kLossCtas) mask slots, then one uint ticket (zeroed by the launcher)
__global__ void __launch_bounds__(256)
score_loss_kernel(const float4* __restrict__ est, const float4* __restrict__ zm, const float* __restrict__ mask,
                  const float* __restrict__ t, double* __restrict__ ws, float* __restrict__ loss, int B, int C, int T4,
                  long total4, float beta_min, float beta_max) {
    __shared__ double s_red[8];
    __shared__ bool s_last;
    double acc = 0.0, macc = 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total4; i += (long)gridDim.x * 256) {
        const long row = i / T4;
        const int b = (int)(row / C);
        const float s = sqrtf(1.0f - expf(-cum_noise(t[b], beta_min, beta_max)));
        const float4 e = est[i], n = zm[i];
        const float d0 = e.x * s + n.x, d1 = e.y * s + n.y, d2 = e.z * s + n.z, d3 = e.w * s + n.w;
        acc += (double)(d0 * d0) + (double)(d1 * d1) + (double)(d2 * d2) + (double)(d3 * d3);
    }
    const long mtotal = (long)B * T4 * 4;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < mtotal; i += (long)gridDim.x * 256) macc += (double)mask[i];
    const double a = block_sum(acc, s_red), m = block_sum(macc, s_red);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + 2 * kLossCtas);
    if (threadIdx.x == 0) {
        ws[blockIdx.x] = a;
        ws[kLossCtas + blockIdx.x] = m;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double fa = 0.0, fm = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) {
        fa += __ldcg(ws + k);
        fm += __ldcg(ws + kLossCtas + k);
    }
    fa = block_sum(fa, s_red);
    fm = block_sum(fm, s_red);
    if (threadIdx.x == 0) loss[0] = (float)(fa / (fm * (double)C));
}

}  // namespace

size_t score_loss_workspace_bytes() { return (2 * kLossCtas + 1) * sizeof(double); }

int forward_diffusion(const float* x0, const float* mask, const float* mu, const float* t, const float* z, float* xt, float* zm,
                      int B, int C, int T, float beta_min, float beta_max, cudaStream_t s) {
    GTTS_REQUIRE(B > 0 && C > 0 && T > 0 && T % 4 == 0, "forward_diffusion: T must be a positive multiple of 4");
    const long total4 = (long)B * C * (T / 4);
    const int grid = (int)std::min<long>((total4 + 255) / 256, 148L * 16);
    forward_diffusion_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(x0), mask, reinterpret_cast<const float4*>(mu), t,
                                                 reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(xt),
                                                 reinterpret_cast<float4*>(zm), C, T / 4, total4, beta_min, beta_max);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int score_loss(const float* est, const float* zm, const float* mask, const float* t, void* ws, size_t ws_bytes, float* loss, int B,
               int C, int T, float beta_min, float beta_max, cudaStream_t s) {
    GTTS_REQUIRE(B > 0 && C > 0 && T > 0 && T % 4 == 0, "score_loss: T must be a positive multiple of 4");
    GTTS_REQUIRE(ws != nullptr && ws_bytes >= score_loss_workspace_bytes(), "score_loss: workspace too small");
    const long total4 = (long)B * C * (T / 4);
    const int grid = (int)std::min<long>((total4 + 255) / 256, (long)kLossCtas);
    GTTS_CHECK_CUDA(cudaMemsetAsync(static_cast<double*>(ws) + 2 * kLossCtas, 0, sizeof(double), s));
    score_loss_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(est), reinterpret_cast<const float4*>(zm), mask, t,
                                          static_cast<double*>(ws), loss, B, C, T / 4, total4, beta_min, beta_max);
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
