// vft_bulk.cuh -- TMA bulk copies (cp.async.bulk, completion on an mbarrier) for rows that a whole CTA shares.
//
// The all-candidate sweeps (k_out_distance_all, k_one_vs_all_warp) evaluate thousands of candidates against ONE profile: the
// out-profile, or the query.  That profile's rows (codeDist table + weights; codes + weights + vectors) are staged once per
// CTA into shared memory by the TMA engine -- one elected thread issues the copies, the bytes land asynchronously while the
// other warps set up, every thread then waits on the mbarrier's phase -- and the per-position gathers of the distance loop
// (ocd[pos][code], the query's weight / vector) become shared-memory reads instead of L1/L2 traffic issued 32 lanes wide with
// their own address arithmetic.  sm_90+ PTX; on sm_100a the copies are UBLKCP instructions.
#pragma once
#include <cstdint>

namespace vft {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dstSmem, const void *srcGlobal, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// up to three global arrays into consecutive shared-memory regions; sizes are multiples of 16 bytes, addresses 16-byte aligned.
// Called by every thread of the CTA; returns once the bytes are visible to all of them.
struct BulkSrc { const void *p; uint32_t bytes; };
__device__ __forceinline__ void cta_bulk_stage(unsigned char *dst, const BulkSrc (&src)[3], uint64_t *bar) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, src[0].bytes + src[1].bytes + src[2].bytes);
        unsigned char *d = dst;
        for (int k = 0; k < 3; k++) {
            const unsigned char *g = static_cast<const unsigned char *>(src[k].p);
            for (uint32_t off = 0; off < src[k].bytes; off += 32768u) {      // pieces of at most 32 KB
                const uint32_t n = src[k].bytes - off < 32768u ? src[k].bytes - off : 32768u;
                bulk_g2s(d + off, g + off, n, bar);
            }
            d += src[k].bytes;
        }
    }
    mbar_wait(bar, 0);
}

}  // namespace vft
