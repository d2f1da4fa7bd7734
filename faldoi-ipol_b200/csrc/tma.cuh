// TMA helpers shared by the tile kernels (tv_tile2_kernel, tv_csad2_kernel, nltv_tile_kernel): a tile of a plane is
// staged into shared memory by one cp.async.bulk.tensor box load per plane (UTMALDG), completion on an mbarrier;
// out-of-frame parts of a box are zero-filled by the TMA unit, which is all the frame borders need (the
// boundary-aware stencils never use those values).
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"
#include "tv_kernels.cuh"

namespace faldoi {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_box(void *dst_smem, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

}  // namespace faldoi
