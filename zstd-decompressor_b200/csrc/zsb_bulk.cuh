// zsb_bulk.cuh -- sm_100a bulk asynchronous copies (cp.async.bulk, the one-dimensional form of TMA) and the mbarrier that tracks them.
//
// Used where a CTA moves a whole contiguous tile between HBM and shared memory: the 128 KiB block image of k_exec (shared -> HBM),
// its literal staging (HBM -> shared), the raw / RLE block expansion of k_rawrle.  One elected thread issues the copy, the copy
// engine of the SM moves the bytes, nobody holds registers for them.  Both addresses must be 16-byte aligned and the size a multiple
// of 16: callers copy the unaligned head and tail themselves.  SASS: UBLKCP (the copies), SYNCS (the mbarrier).
#pragma once
#include <stdint.h>

__device__ __forceinline__ void zsb_mbar_init(uint32_t mbar_sa, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_sa), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// the executing thread arrives and announces `bytes` of asynchronous copies that will complete on the barrier
__device__ __forceinline__ void zsb_mbar_expect(uint32_t mbar_sa, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zsb_mbar_wait(uint32_t mbar_sa, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ZSB_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra ZSB_MBAR_WAIT_%=;\n\t}"
        ::"r"(mbar_sa), "r"(parity) : "memory");
}
// HBM -> shared memory, completion counted on the mbarrier (bytes: multiple of 16, < 1 MiB)
__device__ __forceinline__ void zsb_bulk_g2s(uint32_t dst_sa, const void *src, uint32_t bytes, uint32_t mbar_sa) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_sa), "l"(src), "r"(bytes), "r"(mbar_sa) : "memory");
}
// shared memory -> HBM, tracked by the issuing thread's bulk groups
__device__ __forceinline__ void zsb_bulk_s2g(void *dst, uint32_t src_sa, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zsb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread: the shared-memory source has been read (it may be overwritten) / the copy is complete (visible in HBM)
__device__ __forceinline__ void zsb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void zsb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// makes this thread's earlier ordinary shared-memory writes visible to the copy engine (executed by every writer, before the barrier
// behind which the copy is issued)
__device__ __forceinline__ void zsb_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
