// Collapsed quadratic-form evaluation of the amplitude-embedded circuit (structure (iii)): OPT-IN, reported beside
// -- never instead of -- the statevector kernels, and the device-side second oracle of the parity suite.
//
// With amplitude embedding the trainable unitary U is data independent and the embedded vector is real with support on the
// first q basis states, so for ANY number of layers
//     <Z_i> = xh^T M_i xh,   xh = pre / ||pre||,   M_i = Re(U[:, :q]^H Z'_i U[:, :q])      (q x q, symmetric)
// where Z'_i is the readout pulled back through the CNOT chain.  The q^3 numbers M are NOT re-derived here: the host evaluates the
// STATEVECTOR kernel (qw_circuit_forward) on q (q + 1) / 2 probe windows -- e_a and (e_a + e_b) / sqrt 2 -- and reads M off them
// (M_aa = f(e_a), M_ab = f((e_a + e_b)/sqrt 2) - (M_aa + M_bb) / 2), so the collapsed path inherits the simulator's semantics
// for every weight and layer count, and its weight gradients flow back through the statevector adjoint kernel on those probes.
// What this file does per window is the data-dependent part only:
//     forward   out_i = xh^T M_i xh                                      (q^3 FMAs instead of ~14 q 2^q n_layers)
//     backward  g_xh = 2 sum_i gout_i M_i xh;  gpre = (g_xh - xh (xh . g_xh)) / ||pre||
//               gM_i[a][b] = sum_windows gout_i xh_a xh_b                (per-CTA partial rows, fixed-order final sum)
// It does not exist for angle embedding (the embedded state then depends on the data non-linearly).
#include "../../include/qw.h"
#include "qw_common.cuh"

namespace qw {
namespace col {

constexpr int kMaxQ = 12;
constexpr int kThreads = 128;

template <typename T>
struct Args {
  const T *pre, *M, *gout;
  T *out, *gpre, *part;  // part: [grid][q^3]
  long long W;
  int q;
};

template <typename T>
__device__ __forceinline__ T rsqrt_t(T v);
template <>
__device__ __forceinline__ float rsqrt_t<float>(float v) { return 1.0f / sqrtf(v); }
template <>
__device__ __forceinline__ double rsqrt_t<double>(double v) { return 1.0 / sqrt(v); }

template <typename T>
__global__ void __launch_bounds__(kThreads) collapsed_fwd_kernel(const Args<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ms = reinterpret_cast<T*>(smem_raw);  // [q][q][q]
  const int q = a.q, q3 = q * q * q;
  for (int e = threadIdx.x; e < q3; e += kThreads) Ms[e] = a.M[e];
  __syncthreads();
  for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < a.W; w += (long long)gridDim.x * kThreads) {
    T x[kMaxQ];
    T ss = T(0);
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j)
      if (j < q) {
        x[j] = a.pre[w * q + j];
        ss = fma(x[j], x[j], ss);
      }
    const T inv = rsqrt_t<T>(ss);  // ||pre|| = 0 -> inf / NaN exactly like the reference (quantum_whisper.py:74)
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j)
      if (j < q) x[j] *= inv;
    for (int i = 0; i < q; ++i) {
      const T* Mi = Ms + i * q * q;
      T o = T(0);
#pragma unroll
      for (int aa = 0; aa < kMaxQ; ++aa)
        if (aa < q) {
          T r = T(0);
#pragma unroll
          for (int bb = 0; bb < kMaxQ; ++bb)
            if (bb < q) r = fma(Mi[aa * q + bb], x[bb], r);
          o = fma(x[aa], r, o);
        }
      a.out[w * q + i] = o;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) collapsed_bwd_kernel(const Args<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int q = a.q, q2 = q * q, q3 = q2 * q;
  T* Ms = reinterpret_cast<T*>(smem_raw);   // [q][q][q]
  T* xs = Ms + q3;                          // [kThreads][q]  xh of the chunk's windows
  T* gs = xs + kThreads * q;                // [kThreads][q]  gout
  for (int e = threadIdx.x; e < q3; e += kThreads) Ms[e] = a.M[e];
  // each thread owns the outputs o = tid, tid + 128, ... of gM (<= 14 at q = 12)
  constexpr int kMaxOwn = (kMaxQ * kMaxQ * kMaxQ + kThreads - 1) / kThreads;
  T acc[kMaxOwn];
#pragma unroll
  for (int k = 0; k < kMaxOwn; ++k) acc[k] = T(0);
  __syncthreads();
  for (long long w0 = (long long)blockIdx.x * kThreads; w0 < a.W; w0 += (long long)gridDim.x * kThreads) {
    const long long w = w0 + threadIdx.x;
    const bool valid = w < a.W;
    T x[kMaxQ], g[kMaxQ];
    T ss = T(0);
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j)
      if (j < q) {
        x[j] = valid ? a.pre[w * q + j] : (j == 0 ? T(1) : T(0));
        g[j] = valid ? a.gout[w * q + j] : T(0);
        ss = fma(x[j], x[j], ss);
      }
    const T inv = rsqrt_t<T>(ss);
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j)
      if (j < q) {
        x[j] *= inv;
        xs[threadIdx.x * q + j] = x[j];
        gs[threadIdx.x * q + j] = g[j];
      }
    // g_xh = 2 sum_i gout_i M_i xh  (M_i symmetric)
    T gx[kMaxQ];
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j) gx[j] = T(0);
    for (int i = 0; i < q; ++i) {
      const T* Mi = Ms + i * q2;
      T gsel = T(0);  // gout_i: a register array read with a run-time index, as a select chain
#pragma unroll
      for (int j = 0; j < kMaxQ; ++j)
        if (j == i) gsel = g[j];
      gsel *= T(2);
#pragma unroll
      for (int aa = 0; aa < kMaxQ; ++aa)
        if (aa < q) {
          T r = T(0);
#pragma unroll
          for (int bb = 0; bb < kMaxQ; ++bb)
            if (bb < q) r = fma(Mi[aa * q + bb], x[bb], r);
          gx[aa] = fma(gsel, r, gx[aa]);
        }
    }
    T dot = T(0);
#pragma unroll
    for (int j = 0; j < kMaxQ; ++j)
      if (j < q) dot = fma(gx[j], x[j], dot);
    if (valid) {
#pragma unroll
      for (int j = 0; j < kMaxQ; ++j)
        if (j < q) a.gpre[w * q + j] = (gx[j] - x[j] * dot) * inv;
    }
    __syncthreads();  // xs / gs of the chunk complete
    // gM_i[a][b] += sum over the chunk's windows of gout_i xh_a xh_b
#pragma unroll
    for (int k = 0; k < kMaxOwn; ++k) {
      const int o = threadIdx.x + k * kThreads;
      if (o < q3) {
        const int i = o / q2, r = o - i * q2, aa = r / q, bb = r - aa * q;
        T s = acc[k];
        for (int t = 0; t < kThreads; ++t) s = fma(gs[t * q + i] * xs[t * q + aa], xs[t * q + bb], s);
        acc[k] = s;
      }
    }
    __syncthreads();  // before the next chunk overwrites xs / gs
  }
  T* prow = a.part + (size_t)blockIdx.x * q3;
#pragma unroll
  for (int k = 0; k < kMaxOwn; ++k) {
    const int o = threadIdx.x + k * kThreads;
    if (o < q3) prow[o] = acc[k];
  }
}

// fixed-order (deterministic) sum of the per-CTA rows, in double
template <typename T>
__global__ void __launch_bounds__(256) collapsed_reduce_kernel(const T* __restrict__ part, T* __restrict__ gM, int rows, int n) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += (double)part[(size_t)r * n + o];
  gM[o] = (T)s;
}

static int grid_for(long long W) {
  const long long need = (W + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <typename T>
static int forward(const T* pre, const T* M, T* out, long long W, int q, cudaStream_t st) {
  Args<T> a{pre, M, nullptr, out, nullptr, nullptr, W, q};
  const size_t smem = (size_t)q * q * q * sizeof(T);
  collapsed_fwd_kernel<T><<<grid_for(W), kThreads, smem, st>>>(a);
  count_launch();
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
static int backward(const T* pre, const T* M, const T* gout, T* gpre, T* gM, void* ws, long long W, int q, cudaStream_t st) {
  const int grid = grid_for(W), q3 = q * q * q;
  Args<T> a{pre, M, gout, nullptr, gpre, (T*)ws, W, q};
  const size_t smem = ((size_t)q3 + 2 * (size_t)kThreads * q) * sizeof(T);
  collapsed_bwd_kernel<T><<<grid, kThreads, smem, st>>>(a);
  collapsed_reduce_kernel<T><<<(q3 + 255) / 256, 256, 0, st>>>((const T*)ws, gM, grid, q3);
  count_launch(2);
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace col
}  // namespace qw

extern "C" {

size_t qw_circuit_collapsed_workspace_bytes(long long W, int q, int elem_size) {
  if (W <= 0 || q < 1 || q > qw::col::kMaxQ) return 0;
  return qw::align_up((size_t)qw::col::grid_for(W) * q * q * q * (size_t)elem_size, 256);
}

#define QW_COLLAPSED_CHECK(name)                                                                                          \
  QW_CHECK_ARG(q >= 1 && q <= qw::col::kMaxQ, -2, name ": n_qubits=%d outside [1, %d]", q, qw::col::kMaxQ);              \
  QW_CHECK_ARG(W > 0, -1, name ": empty batch")

int qw_circuit_forward_collapsed(const float* pre, const float* M, float* out, long long W, int q, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(pre && M && out, -1, "qw_circuit_forward_collapsed: null pointer argument");
  QW_COLLAPSED_CHECK("qw_circuit_forward_collapsed");
  return col::forward<float>(pre, M, out, W, q, (cudaStream_t)stream);
}
int qw_circuit_forward_collapsed_f64(const double* pre, const double* M, double* out, long long W, int q, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(pre && M && out, -1, "qw_circuit_forward_collapsed_f64: null pointer argument");
  QW_COLLAPSED_CHECK("qw_circuit_forward_collapsed_f64");
  return col::forward<double>(pre, M, out, W, q, (cudaStream_t)stream);
}
int qw_circuit_backward_collapsed(const float* pre, const float* M, const float* gout, float* gpre, float* gM, void* workspace,
                                  size_t ws_bytes, long long W, int q, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(pre && M && gout && gpre && gM && workspace, -1, "qw_circuit_backward_collapsed: null pointer argument");
  QW_COLLAPSED_CHECK("qw_circuit_backward_collapsed");
  QW_CHECK_ARG(ws_bytes >= qw_circuit_collapsed_workspace_bytes(W, q, 4), -3, "qw_circuit_backward_collapsed: workspace too small");
  return col::backward<float>(pre, M, gout, gpre, gM, workspace, W, q, (cudaStream_t)stream);
}
int qw_circuit_backward_collapsed_f64(const double* pre, const double* M, const double* gout, double* gpre, double* gM, void* workspace,
                                      size_t ws_bytes, long long W, int q, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(pre && M && gout && gpre && gM && workspace, -1, "qw_circuit_backward_collapsed_f64: null pointer argument");
  QW_COLLAPSED_CHECK("qw_circuit_backward_collapsed_f64");
  QW_CHECK_ARG(ws_bytes >= qw_circuit_collapsed_workspace_bytes(W, q, 8), -3, "qw_circuit_backward_collapsed_f64: workspace too small");
  return col::backward<double>(pre, M, gout, gpre, gM, workspace, W, q, (cudaStream_t)stream);
}

}  // extern "C"
