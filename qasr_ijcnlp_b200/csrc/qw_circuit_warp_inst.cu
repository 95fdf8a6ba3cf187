// Explicit instantiation unit of the thread-group cooperative circuit kernels: compile with -DQW_T=float|double -DQW_Q=1..12.
#include "qw_circuit_warp.cuh"

#ifndef QW_T
#error "compile with -DQW_T=<float|double> -DQW_Q=<1..12>"
#endif

namespace qw {
namespace wc {

template <typename T, int Q>
int wcirc_forward_tq(const WArgs<T>& a, int grid, cudaStream_t st) {
  const size_t smem = wcirc_smem_bytes<T, Q>(a.Lq, false);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "circuit forward needs %zu bytes of shared memory", smem);
  auto k = wcirc_fwd_kernel<T, Q>;
  if (smem > 48 * 1024) QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    KernelTimer kt(kKCircFwd, st);
    k<<<grid, Cfg<Q>::THREADS, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, int Q>
int wcirc_backward_tq(const WArgs<T>& a, int grid, cudaStream_t st) {
  const size_t smem = wcirc_smem_bytes<T, Q>(a.Lq, true);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "circuit backward needs %zu bytes of shared memory", smem);
  auto k = wcirc_bwd_kernel<T, Q>;
  if (smem > 48 * 1024) QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    KernelTimer kt(kKCircBwd, st);
    k<<<grid, Cfg<Q>::THREADS, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template int wcirc_forward_tq<QW_T, QW_Q>(const WArgs<QW_T>&, int, cudaStream_t);
template int wcirc_backward_tq<QW_T, QW_Q>(const WArgs<QW_T>&, int, cudaStream_t);

#if QW_Q == 1
template <typename T>
int wcirc_finalize_t(const T* part, const T* qw, T* gqw, int G, int PA, int ngates, cudaStream_t st) {
  {
    KernelTimer kt(kKCircFinalize, st);
    wcirc_finalize_kernel<T><<<ngates, 256, 0, st>>>(part, qw, gqw, G, PA, ngates);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template int wcirc_finalize_t<QW_T>(const QW_T*, const QW_T*, QW_T*, int, int, int, cudaStream_t);
#endif

}  // namespace wc
}  // namespace qw
