// Explicit instantiation unit: compile with -DQW_T=float|double -DQW_Q=1..4.
#include "qw_conv1d_plan.cuh"

#ifndef QW_T
#error "compile with -DQW_T=<float|double> -DQW_Q=<1..4>"
#endif

namespace qw {

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
  if (bytes > 48 * 1024) QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <typename T, int Q, int KT, int WPT>
static int launch_fwd(const FwdArgs<T>& a, const Plan& p, cudaStream_t st) {
  const size_t smem = fwd_smem_elems<T, Q>(a.d.C * a.d.K, a.d.O, a.d.Lq, 32 * WPT) * sizeof(T);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "forward needs %zu bytes of shared memory (C*K too large)", smem);
  auto k = qconv_fwd_kernel<T, Q, KT, WPT>;
  if (int e = set_smem(k, smem)) return e;
  {
    KernelTimer kt(kKFwd, st);
    k<<<p.gridF, kThreads, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T, int Q, int KT>
static int launch_fwd_w(const FwdArgs<T>& a, const Plan& p, cudaStream_t st) {
  switch (p.wpt) {
    case 4: return launch_fwd<T, Q, KT, 4>(a, p, st);
    case 2: return launch_fwd<T, Q, KT, 2>(a, p, st);
    default: return launch_fwd<T, Q, KT, 1>(a, p, st);
  }
}
template <typename T, int Q>
int fwd_tq(const FwdArgs<T>& a, const Plan& p, cudaStream_t st) {
  return (a.d.K == 3) ? launch_fwd_w<T, Q, 3>(a, p, st) : launch_fwd_w<T, Q, 0>(a, p, st);
}

template <typename T, int Q, int WPT>
static int launch_bwdA(const BwdAArgs<T>& a, const Plan& p, cudaStream_t st) {
  const size_t smem = bwdA_smem_elems<T, Q>(a.d.O, a.d.Lq, 32 * WPT) * sizeof(T);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "backward(post) needs %zu bytes of shared memory", smem);
  auto k = qconv_bwd_post_kernel<T, Q, WPT>;
  if (int e = set_smem(k, smem)) return e;
  {
    KernelTimer kt(kKBwdPost, st);
    k<<<p.gridA, kThreads, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T, int Q>
static int launch_bwdA_w(const BwdAArgs<T>& a, const Plan& p, cudaStream_t st) {
  switch (p.wpt) {
    case 4: return launch_bwdA<T, Q, 4>(a, p, st);
    case 2: return launch_bwdA<T, Q, 2>(a, p, st);
    default: return launch_bwdA<T, Q, 1>(a, p, st);
  }
}
template <typename T, int Q, int KT>
static int launch_bwdB(const BwdBArgs<T>& a, const Plan& p, cudaStream_t st) {
  const size_t smem = bwdB_smem_elems<T, Q, KT>() * sizeof(T);
  auto k = qconv_bwd_pre_kernel<T, Q, KT>;
  if (int e = set_smem(k, smem)) return e;
  {
    KernelTimer kt(kKBwdPre, st);
    k<<<dim3(p.gridBx, p.nchunks), kThreads, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, int Q>
int bwd_tq(const T* gy, const T* x, const T* pre_save, const T* w_pre, const T* qwts, const T* w_post,
                             T* gx, T* gw_pre, T* gb_pre, T* gqw, T* gw_post, T* gb_post, unsigned char* ws,
                             const ConvDims& d, const Plan& p, cudaStream_t st) {
  const WsLayout<T> wl = ws_layout<T>(d, p);
  T* gpre = reinterpret_cast<T*>(ws + wl.off_gpre);
  T* partA = reinterpret_cast<T*>(ws + wl.off_partA);
  T* partB = reinterpret_cast<T*>(ws + wl.off_partB);
  BwdAArgs<T> aa{gy, pre_save, qwts, w_post, gpre, partA, d, p.tiles_per_utt, p.num_tiles, p.PA};
  if (int e = launch_bwdA_w<T, Q>(aa, p, st)) return e;
  BwdBArgs<T> ab{x, gpre, w_pre, gx, partB, d, p.ptiles_per_utt, p.num_ptiles, p.Cpad};
  if (int e = (p.KT == 3 ? launch_bwdB<T, Q, 3>(ab, p, st) : launch_bwdB<T, Q, 8>(ab, p, st))) return e;
  FinArgs<T> fa{partA, partB, qwts, gw_pre, gb_pre, gqw, gw_post, gb_post, p.gridA, p.PA, p.gridBx, p.PB,
                d.C, d.K, d.O, d.Q, d.Lq, p.KT};
  const int nblk = p.PA / 32 + (p.PB + 31) / 32;
  {
    KernelTimer kt(kKBwdFinalize, st);
    qconv_bwd_finalize_kernel<T><<<nblk, kFinThreads, 0, st>>>(fa);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}


template <typename T, int Q>
int circ_fwd_tq(const CircArgs<T>& a, int grid, cudaStream_t st) {
  {
    KernelTimer kt(kKCircFwd, st);
    circuit_fwd_kernel<T, Q><<<grid, kThreads, 0, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T, int Q>
int circ_bwd_tq(const CircArgs<T>& a, int grid, cudaStream_t st) {
  const size_t smem = ((size_t)a.Lq * Q * kGateStride + (size_t)a.Lq * Q * 8 * (kThreads + 1)) * sizeof(T);
  auto k = circuit_bwd_kernel<T, Q>;
  if (int e = set_smem(k, smem)) return e;
  {
    KernelTimer kt(kKCircBwd, st);
    k<<<grid, kThreads, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}


template int fwd_tq<QW_T, QW_Q>(const FwdArgs<QW_T>&, const Plan&, cudaStream_t);
template int bwd_tq<QW_T, QW_Q>(const QW_T*, const QW_T*, const QW_T*, const QW_T*, const QW_T*, const QW_T*, QW_T*, QW_T*, QW_T*,
                                QW_T*, QW_T*, QW_T*, unsigned char*, const ConvDims&, const Plan&, cudaStream_t);
template int circ_fwd_tq<QW_T, QW_Q>(const CircArgs<QW_T>&, int, cudaStream_t);
template int circ_bwd_tq<QW_T, QW_Q>(const CircArgs<QW_T>&, int, cudaStream_t);

}  // namespace qw
