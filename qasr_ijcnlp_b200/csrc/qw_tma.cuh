// Tensor-map TMA helpers (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) for 3-D fp32 tensors.
// The host side encodes a CUtensorMap through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "qw_async.cuh"
#include "qw_common.cuh"

namespace qw {

// ---- device
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src_smem) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_u32(src_smem))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// float offset of element (row r, 16-byte chunk c) inside a SWIZZLE_128B tile whose rows are 32 floats
// (tile base 1024-byte aligned): the chunk index is XORed with the row index modulo 8.
__device__ __forceinline__ int swz128(int r, int c) { return r * 32 + ((c ^ (r & 7)) << 2); }

// ---- host
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encode_fn();

// fp32 tensor (d2, d1, d0) row-major (d0 contiguous); box (1, box1, box0).  Returns 0 or an error code.
int make_tmap_3d_f32(CUtensorMap* tm, const void* base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                     unsigned box0, unsigned box1, bool swizzle128);

}  // namespace qw
