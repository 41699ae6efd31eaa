// common.cuh -- shared helpers for the libpcst kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcst.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpcst is written for sm_100a (B200) only"
#endif

namespace pcst {

// SM count of the current device (148 on a B200), queried once per device; grids are sized in multiples of it.
// Host-only plan queries on a machine without a GPU get the B200 value.
int num_sms();

// ---- host-side error plumbing --------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int tuning(const char* key, int dflt);

#define PCST_CHECK_ARG(cond, msg)                          \
    do {                                                   \
        if (!(cond)) {                                     \
            pcst::set_error("%s: %s", __func__, msg);      \
            return PCST_ERR_INVALID;                       \
        }                                                  \
    } while (0)

#define PCST_CUDA(call)                                            \
    do {                                                           \
        int _st = pcst::check_cuda((call), #call);                 \
        if (_st != PCST_OK) return _st;                            \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- packed point layout -------------------------------------------------------------------
// Candidate points are repacked once per call into float4 (x, y, z, |p|^2) rows, padded to a
// whole number of tiles with sentinels whose norm is +inf (never within a radius, never a
// minimum), so that tiles can be moved by 1-D TMA bulk copies with no tail handling.
constexpr int kTilePoints = 1024;                      // candidates per shared-memory tile
constexpr int kTileBytes = kTilePoints * 16;           // 16 KiB

inline int padded_points(int n) { return (int)align_up((size_t)(n > 0 ? n : 1), kTilePoints); }

int launch_pack(const float* xyz, int B, int N, int Npad, float4* out, cudaStream_t stream);

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__

// |p|^2 exactly as torch.sum(x ** 2, -1) forms it on the size-3 axis: (x*x + y*y) + z*z, no FMA.
__device__ __forceinline__ float norm3_sq(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// K=3 dot product exactly as MKL sgemm accumulates it: fma(a2,b2, fma(a1,b1, a0*b0)).
__device__ __forceinline__ float dot3_chain(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// mbarrier (shared::cta) ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst and src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// thread-block cluster ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    cluster_arrive_release();
    cluster_wait_acquire();
}
// address of `local_smem_addr` (a shared::cta u32 address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_v2_u32(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

#endif  // __CUDACC__

}  // namespace pcst
