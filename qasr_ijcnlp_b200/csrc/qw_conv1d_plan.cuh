// Host-side launch plan shared by the dispatcher (qw_conv1d.cu) and the per-(dtype, n_qubits) instantiation
// units (qw_conv1d_inst.cu, compiled once per combination so the build parallelises).
#pragma once
#include "qw_conv1d_kernels.cuh"

namespace qw {

struct Plan {
  int wpt;                 // windows per thread (fwd / bwd A)
  int tw;                  // windows per tile
  int tiles_per_utt, num_tiles;
  int gridF, gridA;        // CTAs
  int ptiles_per_utt, num_ptiles, nchunks, gridBx, Cpad;
  int KT;                  // 3 or 8
  int PA, PB;
};


template <typename T>
struct WsLayout {
  size_t off_gpre, off_partA, off_partB, total;
};
template <typename T>
inline WsLayout<T> ws_layout(const ConvDims& d, const Plan& p) {
  WsLayout<T> w{};
  size_t o = 0;
  w.off_gpre = o;
  o = align_up(o + (size_t)d.B * d.Lout * d.Q * sizeof(T), 256);
  w.off_partA = o;
  o = align_up(o + (size_t)p.gridA * p.PA * sizeof(T), 256);
  w.off_partB = o;
  o = align_up(o + (size_t)p.gridBx * p.PB * sizeof(T), 256);
  w.total = o;
  return w;
}


// per-(T,Q) entry points, explicitly instantiated in qw_conv1d_inst.cu
template <typename T, int Q>
int fwd_tq(const FwdArgs<T>& a, const Plan& p, cudaStream_t st);
template <typename T, int Q>
int bwd_tq(const T* gy, const T* x, const T* pre_save, const T* w_pre, const T* qwts, const T* w_post, T* gx, T* gw_pre,
           T* gb_pre, T* gqw, T* gw_post, T* gb_post, unsigned char* ws, const ConvDims& d, const Plan& p, cudaStream_t st);
template <typename T, int Q>
int circ_fwd_tq(const CircArgs<T>& a, int grid, cudaStream_t st);
template <typename T, int Q>
int circ_bwd_tq(const CircArgs<T>& a, int grid, cudaStream_t st);

}  // namespace qw
