// Variant probe: each variant runs in its own process (an illegal instruction kills the context).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int V>
__global__ void k(const __grid_constant__ CUtensorMap tm, const float* src, float* out, int n, int c0, int c1, int c2, int pf) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = (unsigned char*)(((uintptr_t)sm + 1023) & ~uintptr_t(1023));
  float* tile = (float*)base;
  uint64_t* bar = (uint64_t*)(base + 32768);
  if (threadIdx.x == 0) {
    if (pf) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (V == 0) {  // 1-D bulk copy
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(n * 4) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(tile)), "l"(src), "r"(n * 4), "r"(s32(bar)) : "memory");
    } else if (V == 1) {  // 2-D tensor, no .tile
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(n * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(tile)), "l"(&tm), "r"(0), "r"(0), "r"(s32(bar)) : "memory");
    } else if (V == 2 || V == 3 || V == 4) {  // 3-D tensor
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(n * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
    }
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  int V = argc > 1 ? atoi(argv[1]) : 0;
  const int D0 = 200, D1 = 80, D2 = 2;
  std::vector<float> h((size_t)D0 * D1 * D2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *dsrc, *dout;
  cudaMalloc(&dsrc, h.size() * 4); cudaMalloc(&dout, 65536 * 4);
  cudaMemcpy(dsrc, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(ce), (int)q, p);
  EncFn enc = (EncFn)p;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int n = 0;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  if (V == 0) n = 512;
  if (V == 1) { cuuint64_t gd[2] = {200, 160}; cuuint64_t gs[1] = {800}; cuuint32_t bx[2] = {32, 16}; n = 512;
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dsrc, gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  if (V == 2) { cuuint64_t gd[3] = {200, 80, 2}; cuuint64_t gs[2] = {800, 64000}; cuuint32_t bx[3] = {32, 16, 1}; n = 512;
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dsrc, gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  if (V == 3) { cuuint64_t gd[3] = {200, 80, 2}; cuuint64_t gs[2] = {800, 64000}; cuuint32_t bx[3] = {36, 16, 1}; n = 576;
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dsrc, gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  if (V == 4) { cuuint64_t gd[3] = {200, 80, 2}; cuuint64_t gs[2] = {800, 64000}; cuuint32_t bx[3] = {32, 32, 1}; n = 1024;
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dsrc, gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  printf("variant %d encode result %d\n", V, (int)r);
  size_t smem = 1024 + 32768 + 64;
  auto launch = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kern<<<1, 128, smem>>>(tm, dsrc, dout, n, argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atoi(argv[3]) : 0, argc > 4 ? atoi(argv[4]) : 0, argc > 5 ? atoi(argv[5]) : 0); };
  switch (V) { case 0: launch(k<0>); break; case 1: launch(k<1>); break; case 2: launch(k<2>); break; case 3: launch(k<3>); break; default: launch(k<4>); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d: %s\n", V, cudaGetErrorString(e));
  if (e == cudaSuccess) { std::vector<float> o(n); cudaMemcpy(o.data(), dout, n * 4, cudaMemcpyDeviceToHost); printf("   out[0..3]=%g %g %g %g  out[32]=%g out[36]=%g\n", o[0], o[1], o[2], o[3], o[32], o[36]); }
  return 0;
}
