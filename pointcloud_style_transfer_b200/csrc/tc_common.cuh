// tc_common.cuh -- tcgen05 / TMEM / UMMA-descriptor PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace pcst {

constexpr int kTcM = 128;           // rows per CTA = TMEM lanes
constexpr int kTcEpiThreads = 128;  // warps 0-3
constexpr int kTcThreads = 192;     // + producer warp + MMA warp

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading (K) byte offset between core matrices
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride (M/N) byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
    return d;                                            // base_offset 0, layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4)                      // D format: fp32
           | (1u << 7)                    // A format: bf16
           | (1u << 10)                   // B format: bf16
           | ((uint32_t)(N >> 3) << 17)   // N
           | ((uint32_t)(M >> 4) << 24);  // M;  A and B both K-major (bits 15, 16 = 0)
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// shared::cta -> (peer) shared::cluster bulk copy, completion counted on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                  uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster_addr),
                 "r"(src_cta_addr), "r"(bytes), "r"(bar_cluster_addr)
                 : "memory");
}
// One lane of a converged warp (the same one every time for a given warp): the issuer of single-thread instructions.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// The wait names the registers of the load it completes ("+r"): their uses cannot be scheduled above it, while a load
// issued AFTER the wait into the other buffer stays in flight during the arithmetic on this one.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// Walk an accumulator's n columns 16 at a time with the NEXT 16 already loading: f(r, c0) per block.
template <typename F>
__device__ __forceinline__ void tmem_for_each16(uint32_t taddr, int n, F&& f) {
    uint32_t ra[16], rb[16];
    tmem_ld16_issue(taddr, ra);
    for (int c0 = 0; c0 < n; c0 += 32) {
        tmem_ld_wait(ra);
        const bool second = c0 + 16 < n;
        if (second) tmem_ld16_issue(taddr + (uint32_t)c0 + 16u, rb);
        f(ra, c0);
        if (second) {
            tmem_ld_wait(rb);
            if (c0 + 32 < n) tmem_ld16_issue(taddr + (uint32_t)c0 + 32u, ra);
            f(rb, c0 + 16);
        }
    }
}
// Poll with a pause: the single-lane producer / MMA warps share issue slots with two of the four epilogue warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, bool relaxed) {
    while (!mbar_try_wait(bar, parity)) {
        if (relaxed) __nanosleep(64);
    }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // warps 0-3 only

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 1-D TMA bulk copy global -> the SAME shared-memory offset of every CTA in cta_mask (one L2 read serves the cluster);
// completion is counted on the mbarrier at the same offset in each destination CTA.
__device__ __forceinline__ void tma_load_1d_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                      uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}

// ---- the same primitives on precomputed shared-window addresses.  In a cluster launch a generic pointer to a __shared__
// object is converted with S2UR SR_CgaCtaId at EVERY use; an address formed once and hidden from the optimiser is not.
__device__ __forceinline__ uint32_t tc_opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_addr(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d_addr(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d_multicast_addr(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar,
                                                           uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_dst),
        "l"(gmem_src), "r"(bytes), "r"(bar), "h"(cta_mask)
        : "memory");
}

inline uint32_t tmem_cols_pow2(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }

}  // namespace pcst
