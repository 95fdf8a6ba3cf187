// Host dispatcher of the thread-group cooperative circuit kernels (qw_circuit_warp.cuh): n_qubits 1..12, amplitude / angle
// embedding, any n_layers <= 8.  Used by qw_circuit_* for everything the per-thread q <= 4 amplitude kernels do not cover,
// and by the composed QuantumConv1d path (qw_conv1d_general.cu).
#include "qw_circuit_warp.cuh"

namespace qw {
namespace wc {

int wcirc_grid(long long W, int q) {
  const int gpc = wcirc_gpc(q);
  const long long need = (W + gpc - 1) / gpc;
  const long long cap = (long long)num_sms() * (q == 11 ? 8 : 4);
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

size_t wcirc_workspace_bytes(long long W, int q, int Lq, int elem_size) {
  return align_up((size_t)wcirc_grid(W, q) * wcirc_PA(q, Lq) * (size_t)elem_size, 256);
}

template <typename T>
int wcirc_forward(const T* pre, const T* qwts, T* out, long long W, int q, int Lq, int emb, cudaStream_t st) {
  WArgs<T> a{pre, qwts, nullptr, out, nullptr, nullptr, W, Lq, emb, 0};
  const int grid = wcirc_grid(W, q);
  switch (q) {
#define QW_CASE(QQ) case QQ: return wcirc_forward_tq<T, QQ>(a, grid, st);
    QW_CASE(1) QW_CASE(2) QW_CASE(3) QW_CASE(4) QW_CASE(5) QW_CASE(6) QW_CASE(7) QW_CASE(8) QW_CASE(9) QW_CASE(10) QW_CASE(11) QW_CASE(12)
#undef QW_CASE
  }
  set_error("n_qubits=%d outside [1,12]", q);
  return -2;
}

template <typename T>
int wcirc_backward(const T* pre, const T* qwts, const T* gout, T* gpre, T* gqw, void* ws, long long W, int q, int Lq, int emb,
                   cudaStream_t st) {
  const int grid = wcirc_grid(W, q), PA = wcirc_PA(q, Lq);
  WArgs<T> a{pre, qwts, gout, nullptr, gpre, (T*)ws, W, Lq, emb, PA};
  int e = -2;
  switch (q) {
#define QW_CASE(QQ) case QQ: e = wcirc_backward_tq<T, QQ>(a, grid, st); break;
    QW_CASE(1) QW_CASE(2) QW_CASE(3) QW_CASE(4) QW_CASE(5) QW_CASE(6) QW_CASE(7) QW_CASE(8) QW_CASE(9) QW_CASE(10) QW_CASE(11) QW_CASE(12)
#undef QW_CASE
    default: set_error("n_qubits=%d outside [1,12]", q);
  }
  if (e) return e;
  return wcirc_finalize_t<T>((const T*)ws, qwts, gqw, grid, PA, Lq * q, st);
}

template int wcirc_forward<float>(const float*, const float*, float*, long long, int, int, int, cudaStream_t);
template int wcirc_forward<double>(const double*, const double*, double*, long long, int, int, int, cudaStream_t);
template int wcirc_backward<float>(const float*, const float*, const float*, float*, float*, void*, long long, int, int, int, cudaStream_t);
template int wcirc_backward<double>(const double*, const double*, const double*, double*, double*, void*, long long, int, int, int,
                                    cudaStream_t);

}  // namespace wc
}  // namespace qw
