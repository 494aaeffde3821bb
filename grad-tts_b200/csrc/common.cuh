// Shared device/host helpers for the sm_100a Grad-TTS decoder kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>

namespace gtts {

// ---------------------------------------------------------------- host error plumbing
void set_error(const std::string& msg);          // defined in capi.cu
#define GTTS_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::gtts::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +      \
                              " at " + __FILE__ + ":" + std::to_string(__LINE__));             \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define GTTS_REQUIRE(cond, msg)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::gtts::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +         \
                              std::to_string(__LINE__));                                       \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the Euler step is launched with programmaticStreamSerialization: it may start while its
// predecessor drains, runs its prologue, and blocks in pdl_wait() until the predecessor grid has completed and its
// memory is visible.  pdl_trigger() at kernel entry lets the successor be scheduled as soon as resources free up.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();                               // capi.cu (env GTTS_PDL; default off except for small-batch plans)
void pdl_set_override(int on);                    // capi.cu: programmatic dependent launch for the launches that follow (this thread)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    ++n;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- activation element types
template <typename T> struct Act;
template <> struct Act<float> {
    static constexpr int kBytes = 4;
    __device__ __forceinline__ static void load8(const float* p, float (&v)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ static void store8(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __device__ __forceinline__ static float ld(const float* p) { return *p; }
    __device__ __forceinline__ static void st(float* p, float v) { *p = v; }
    struct Packed { float4 a, b; };
    __device__ __forceinline__ static Packed load_packed(const float* p) {
        Packed r;
        r.a = __ldg(reinterpret_cast<const float4*>(p));
        r.b = __ldg(reinterpret_cast<const float4*>(p + 4));
        return r;
    }
    __device__ __forceinline__ static void unpack(const Packed& k, float (&v)[8]) {
        v[0] = k.a.x; v[1] = k.a.y; v[2] = k.a.z; v[3] = k.a.w; v[4] = k.b.x; v[5] = k.b.y; v[6] = k.b.z; v[7] = k.b.w;
    }
};
template <> struct Act<__nv_bfloat16> {
    static constexpr int kBytes = 2;
    __device__ __forceinline__ static void load8(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i]     = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static void store8(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ static float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ __forceinline__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
    typedef uint4 Packed;                                  // 8 bf16 stay packed in 4 registers until used
    __device__ __forceinline__ static Packed load_packed(const __nv_bfloat16* p) {
        return __ldg(reinterpret_cast<const uint4*>(p));
    }
    __device__ __forceinline__ static void unpack(const Packed& u, float (&v)[8]) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i]     = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A thread that owns 32 contiguous bytes of an NHWC pixel row writes
// one full 32-byte sector per instruction instead of two half sectors -- the epilogues' stores are scattered across
// pixels (lane = pixel), so the LSU cost is per sector, not per byte.  Address must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

// ---------------------------------------------------------------- maths
// Mish = x * tanh(softplus(x)), torch softplus(beta=1, threshold=20)  (reference model/diffusion.py:16-18)
template <bool kStrict>
__device__ __forceinline__ float mish(float x) {
    if (kStrict) {
        float sp = (x > 20.0f) ? x : log1pf(expf(x));
        return x * tanhf(sp);
    } else {
        // tanh(log(1+e^x)) = n/(n+2) with n = e^x (e^x + 2); exact algebra, one ex2 + one rcp
        if (x > 20.0f) return x;
        float e = __expf(x);
        float n = e * (e + 2.0f);
        return x * __fdividef(n, n + 2.0f);
    }
}

// Packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100): halves the issue slots of the epilogue arithmetic.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t*>(&r))
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<uint64_t*>(&r))
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return r;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<uint64_t*>(&r))
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Branch-free fast Mish on a pair: y * tanh(softplus(y)) = y * n/(n+2) with n = e(e+2), e = exp(y).  Written as
// y * (1 - 2/d), d = n + 2 = (e+1)^2 + 1: no clamp is needed (e = inf gives d = inf, 1/d = 0, factor 1 -- torch's softplus
// threshold path returns y there as well; y << 0 gives d -> 2, factor -> 0 with absolute error ~1e-7), and d comes from one
// add and one FMA.  ylog2 = y*log2(e) is supplied by the caller.  Two MUFU + four packed instructions per pair.
__device__ __forceinline__ float2 mish2_fast(float2 y, float2 ylog2) {
    const float2 e1 = fadd2(make_float2(ex2_approx(ylog2.x), ex2_approx(ylog2.y)), make_float2(1.0f, 1.0f));
    const float2 d = ffma2(e1, e1, make_float2(1.0f, 1.0f));
    const float2 t = ffma2(make_float2(rcp_approx(d.x), rcp_approx(d.y)), make_float2(-2.0f, -2.0f), make_float2(1.0f, 1.0f));
    return fmul2(y, t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- GroupNorm statistics protocol
// A conv kernel writes, for every (sample b, tile slot s), 16 floats: sum[8 groups], sumsq[8 groups]
// into partials[(b*slots + s)*16 ...]; the last tile of a sample to finish (ticket on counters[b])
// reduces all slots in a fixed order (deterministic, double accumulation) and writes
// stats[b][g] = {mean, rstd}.  `nthreads` threads of the CTA (ids 0..nthreads-1, nthreads>=16,
// multiple of 16, <=256) call this together; `sync` is a barrier over exactly those threads.
struct GnStatsOut {
    float* partials;        // [B][slots][16]
    float* stats;           // [B][8][2]
    unsigned int* counters; // [B], zero on entry, zero again on exit
    int slots;              // tiles contributing per sample
    float inv_count;        // 1 / ((C/8)*H*W)
    float eps;
};

template <typename SyncFn>
__device__ __forceinline__ void gn_stats_publish(const GnStatsOut& o, int b, int slot, int tid, int nthreads,
                                                 const float* tile_vals16 /*smem, valid for tid<16*/,
                                                 double* s_red /*smem [16][16]*/, int* s_flag, SyncFn sync) {
    if (tid < 16) {
        o.partials[((size_t)b * o.slots + slot) * 16 + tid] = tile_vals16[tid];
        __threadfence();
    }
    sync();
    if (tid == 0) {
        __threadfence();
        unsigned int old = atomicAdd(&o.counters[b], 1u);
        *s_flag = (old == (unsigned int)(o.slots - 1));
    }
    sync();
    if (*s_flag) {
        __threadfence();
        const int k = tid & 15, slice = tid >> 4, nslice = nthreads >> 4;
        double acc = 0.0;
        const float* p = o.partials + (size_t)b * o.slots * 16;
        for (int s = slice; s < o.slots; s += nslice) acc += (double)__ldcg(p + (size_t)s * 16 + k);
        s_red[slice * 16 + k] = acc;
        sync();
        if (tid < 8) {
            double sum = 0.0, sq = 0.0;
            for (int s = 0; s < nslice; ++s) { sum += s_red[s * 16 + tid]; sq += s_red[s * 16 + 8 + tid]; }
            double mean = sum * (double)o.inv_count;
            double var = sq * (double)o.inv_count - mean * mean;
            if (var < 0.0) var = 0.0;
            o.stats[((size_t)b * 8 + tid) * 2 + 0] = (float)mean;
            o.stats[((size_t)b * 8 + tid) * 2 + 1] = (float)(1.0 / sqrt(var + (double)o.eps));
        }
        if (tid == 0) o.counters[b] = 0u;
    }
    sync();
}

// ---------------------------------------------------------------- sm_100a PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const void* desc, uint64_t* bar, void* smem, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
// 1-D bulk copy global -> shared (no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const void* desc, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(const void* desc, uint64_t* bar, void* smem, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// 2-D TMA load multicast to the CTAs of `cta_mask` in the cluster (same CTA-relative smem / mbarrier offsets)
__device__ __forceinline__ void tma_load_2d_mc(const void* desc, uint64_t* bar, void* smem, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {      // arrives on `bar` in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one thread of the leader CTA issues M = 256 MMAs that read A (128 rows) and half
// of B (N/2 rows) from EACH CTA's shared memory at the same offsets and write 128 accumulator rows into each CTA's TMEM.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {         // same offset in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed form for TMEM hand-over: the ordering that matters (tcgen05.ld done before the MMA overwrites the buffer) comes
// from tcgen05.wait::ld + tcgen05.fence::before_thread_sync; a cluster-scope RELEASE would additionally wait for the
// thread's outstanding global stores (the previous tile's output) -- measured 181 vs 110 us of epilogue time per launch.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2_mc(uint64_t* bar, uint16_t cta_mask) {     // arrives on `bar` in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, the bytes are counted on `bar_cluster_addr`
// (normally the leader's barrier, mapa_u32(.., 0)).
__device__ __forceinline__ void tma_load_4d_2sm(const void* desc, uint32_t bar_cluster_addr, void* smem, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const void* desc, uint32_t bar_cluster_addr, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 x bf16 -> f32)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace gtts
