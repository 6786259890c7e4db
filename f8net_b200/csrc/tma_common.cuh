// tma_common.cuh -- TMA (cp.async.bulk.tensor) plumbing shared by the tcgen05 kernels: the host
// side encodes CUtensorMap descriptors through the driver entry point (no link dependency on
// libcuda), the device side issues tiled tensor loads that complete on an mbarrier.
#pragma once
#include <cuda.h>

#include "umma_common.cuh"

namespace f8u {

// 4-D tiled load global -> shared; out-of-bounds elements (negative or past-the-end coordinates)
// are written as zeros, which is how every padding halo of the convolutions is produced.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
// 4-D tiled store shared -> global (bulk async-group completion); elements of the box that fall
// outside the tensor are not written, which clips channel / pixel tails for free.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t src) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
        ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the source shared memory of every committed store may be overwritten again
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed store is complete
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace f8u

namespace f8host {

// Encodes a rank-4 uint8 tensor map.  dims / strides innermost first (strides[0] is implied = 1
// byte and not passed; strides[i] in bytes for i = 1..3).  Returns F8_OK or F8_ERR_CUDA.
int encode_tmap_u8_4d(CUtensorMap *out, const void *base, const uint64_t dims[4], const uint64_t strides[3],
                      const uint32_t box[4], CUtensorMapSwizzle swizzle);
// same for 4-byte elements (the NHWC4 input image seen as one uint32 per pixel)
int encode_tmap_u32_4d(CUtensorMap *out, const void *base, const uint64_t dims[4], const uint64_t strides[3],
                       const uint32_t box[4]);

}  // namespace f8host
