// Micro-benchmarks of the tcgen05 issue path (test hook only; not on the product path).
//
// The halo conv's per-tile cost is (MMA execution) + (a fixed skeleton) instead of max(...).  This kernel isolates what
// the tensor-pipe command stream charges for `tcgen05.commit` relative to `tcgen05.mma`, so the conv pipeline can be
// shaped around measured costs:  one CTA per SM, one issuing thread, `iters` rounds of {n_mma MMAs, n_commit commits},
// one final commit + wait, cycles from clock64.
#include "common.cuh"
#include "conv_tc_common.cuh"
#include "ops.h"

namespace gtts {
namespace {

template <int N>
__global__ void __launch_bounds__(128, 1) mb_issue_kernel(int n_mma, int n_commit, int iters, int wait_each,
                                                          unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + N * 128) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0u;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint64_t adesc = tc::make_sw128_kmajor_desc(smem_u32(smem));
        const uint64_t bdesc = tc::make_sw128_kmajor_desc(smem_u32(smem) + 16384);
        constexpr uint32_t idesc = tc::make_idesc<N>();
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll 4
                for (int m = 0; m < n_mma; ++m)
                    tc_mma_f16(tmem + (uint32_t)((it & 1) * N), adesc + (uint64_t)(2 * (m & 3)), bdesc + (uint64_t)(2 * (m & 3)),
                               idesc, (uint32_t)(m != 0));
                for (int c = 0; c < n_commit; ++c) tc_commit(&bars[c & 3]);
            }
            __syncwarp();
            if (wait_each && n_commit > 0) {            // round trip: commit -> mbarrier -> waiting thread
                mbar_wait(&bars[(n_commit - 1) & 3], ph);
                ph ^= 1u;
                tc_fence_after();
            }
        }
        const long long t1 = clock64();
        if (elect_one()) tc_commit(&bars[7]);
        __syncwarp();
        mbar_wait(&bars[7], 0u);
        const long long t2 = clock64();
        if (tid == 0) {
            out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);      // issue time
            out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);      // until everything retired
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

int microbench_issue(int N, int n_mma, int n_commit, int iters, int wait_each, int grid, unsigned long long* out_dev,
                     cudaStream_t stream) {
    const size_t smem = 16384 + 256 * 128 + 1024;
    if (N == 64) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(mb_issue_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mb_issue_kernel<64><<<grid, 128, smem, stream>>>(n_mma, n_commit, iters, wait_each, out_dev);
    } else if (N == 128) {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(mb_issue_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mb_issue_kernel<128><<<grid, 128, smem, stream>>>(n_mma, n_commit, iters, wait_each, out_dev);
    } else {
        GTTS_CHECK_CUDA(cudaFuncSetAttribute(mb_issue_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mb_issue_kernel<256><<<grid, 128, smem, stream>>>(n_mma, n_commit, iters, wait_each, out_dev);
    }
    GTTS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gtts
