// Host dispatcher + C ABI for QuantumConv1d (kernels: qw_conv1d_kernels.cuh; instantiations: qw_conv1d_inst.cu).
#include <mutex>

#include "../../include/qw.h"
#include "qw_conv1d_plan.cuh"
#include "qw_circuit_warp_host.h"

namespace qw {

namespace gen {  // qw_conv1d_general.cu: n_qubits > 4 and / or angle embedding
size_t general_workspace_bytes(const ConvDims& d, int elem);
template <typename T>
int general_forward(const T* x, const T* w_pre, const T* b_pre, const T* qwts, const T* w_post, const T* b_post, T* y, T* pre_save,
                    const ConvDims& d, cudaStream_t st);
template <typename T>
int general_backward(const T* gy, const T* x, const T* pre_save, const T* w_pre, const T* qwts, const T* w_post, T* gx, T* gw_pre,
                     T* gb_pre, T* gqw, T* gw_post, T* gb_post, unsigned char* ws, size_t ws_bytes, const ConvDims& d, cudaStream_t st);
}  // namespace gen
static bool is_general(const ConvDims& d) { return d.Q > 4 || d.emb != kEmbAmplitude; }

static std::mutex g_mu;
static int g_num_sms = 0;

int num_sms() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else {
      (void)cudaGetLastError();
      g_num_sms = 148;  // B200
    }
  }
  return g_num_sms;
}

static Plan make_plan(const ConvDims& d) {
  Plan p{};
  const int sms = num_sms();
  const long long W = (long long)d.B * d.Lout;
  p.wpt = (W >= 128LL * sms * 4) ? 4 : (W >= 64LL * sms * 4) ? 2 : 1;
  p.tw = 32 * p.wpt;
  p.tiles_per_utt = (d.Lout + p.tw - 1) / p.tw;
  p.num_tiles = d.B * p.tiles_per_utt;
  p.gridF = p.num_tiles < sms * 16 ? p.num_tiles : sms * 16;
  // bwd A: persistent CTAs with equal tile counts
  {
    const int cap = sms * 6;
    const int tpc = (p.num_tiles + cap - 1) / cap;
    p.gridA = (p.num_tiles + tpc - 1) / tpc;
  }
  p.KT = (d.K == 3) ? 3 : 8;
  p.ptiles_per_utt = (d.L + kTP - 1) / kTP;
  p.num_ptiles = d.B * p.ptiles_per_utt;
  p.nchunks = (d.C + 31) / 32;
  p.Cpad = p.nchunks * 32;
  {
    int cap = sms * 8 / p.nchunks;
    if (cap < 1) cap = 1;
    const int tpc = (p.num_ptiles + cap - 1) / cap;
    p.gridBx = (p.num_ptiles + tpc - 1) / tpc;
  }
  p.PA = partA_len(d.O, d.Q, d.Lq);
  p.PB = p.Cpad * d.Q * p.KT;
  return p;
}

static int check_dims(ConvDims& d) {
  QW_CHECK_ARG(d.B > 0 && d.C > 0 && d.L > 0 && d.K > 0 && d.S > 0 && d.P >= 0 && d.O > 0, -1,
               "bad shape B=%d C=%d L=%d K=%d S=%d P=%d O=%d", d.B, d.C, d.L, d.K, d.S, d.P, d.O);
  QW_CHECK_ARG(d.L + 2 * d.P >= d.K, -1, "kernel_size %d larger than padded length %d", d.K, d.L + 2 * d.P);
  d.Lout = (d.L + 2 * d.P - d.K) / d.S + 1;
  QW_CHECK_ARG(d.Q >= 1 && d.Q <= (long long)d.C * d.K, -1, "n_qubits=%d must be in [1, C*K=%d]", d.Q, d.C * d.K);
  QW_CHECK_ARG(d.Q <= 12, -2, "QuantumConv1d supports n_qubits <= 12 (got %d)", d.Q);
  QW_CHECK_ARG(d.Lq >= 1 && d.Lq <= 8, -2, "n_layers=%d must be in [1,8]", d.Lq);
  QW_CHECK_ARG(d.emb == kEmbAmplitude || d.emb == kEmbAngle, -2, "embedding=%d is neither amplitude (0) nor angle (1)", d.emb);
  QW_CHECK_ARG((long long)d.B * d.C * d.L < (1LL << 40) && (long long)d.B * d.O * d.Lout < (1LL << 40), -1, "tensor too large");
  return 0;
}

template <typename T>
static int conv1d_forward_impl(const T* x, const T* w_pre, const T* b_pre, const T* qwts, const T* w_post, const T* b_post,
                               T* y, T* pre_save, ConvDims d, void* stream, int act = 0) {
  QW_CHECK_ARG(x && w_pre && b_pre && qwts && w_post && b_post && y, -1, "null pointer argument");
  if (int e = check_dims(d)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  QW_CHECK_ARG(act == 0 || act == QW_ACT_GELU, -2, "activation=%d is neither none (0) nor gelu (1)", act);
  if (act) {
    // the fused activation exists in the fast-path kernels only: callers fall back to the plain operator + a separate GELU
    if constexpr (sizeof(T) == 4) {
      if (!is_general(d) && fast_eligible(d, x, y, pre_save, true))
        return fast_forward(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, st, act);
    }
    set_error("the fused activation needs the fast-path regime (fp32, n_qubits=4, amplitude embedding, K=3, stride 1|2, padding 1, aligned shapes)");
    return -2;
  }
  if (is_general(d)) return gen::general_forward<T>(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, st);
  if constexpr (sizeof(T) == 4) {
    if (fast_eligible(d, x, y, pre_save, true)) return fast_forward(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, st);
  }
  const Plan p = make_plan(d);
  FwdArgs<T> a{x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, p.tiles_per_utt, p.num_tiles};
  switch (d.Q) {
    case 1: return fwd_tq<T, 1>(a, p, st);
    case 2: return fwd_tq<T, 2>(a, p, st);
    case 3: return fwd_tq<T, 3>(a, p, st);
    default: return fwd_tq<T, 4>(a, p, st);
  }
}

template <typename T>
static int conv1d_backward_impl(const T* gy, const T* x, const T* pre_save, const T* w_pre, const T* qwts, const T* w_post,
                                T* gx, T* gw_pre, T* gb_pre, T* gqw, T* gw_post, T* gb_post, void* workspace,
                                size_t ws_bytes, ConvDims d, void* stream, const FastDp* dp = nullptr, const T* b_post = nullptr,
                                int act = 0, const FastChain* chain = nullptr) {
  QW_CHECK_ARG((gy || chain) && x && pre_save && w_pre && qwts && w_post && gw_pre && gb_pre && gqw && gw_post && gb_post && workspace,
               -1, "null pointer argument");
  if (int e = check_dims(d)) return e;
  QW_CHECK_ARG(((uintptr_t)workspace & 255) == 0, -1, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  const bool want_dp = dp && dp->world > 1;
  QW_CHECK_ARG(!(want_dp && is_general(d)), -2, "the fused gradient all-reduce needs the fast-path regime (n_qubits=4, amplitude embedding)");
  QW_CHECK_ARG(act == 0 || act == QW_ACT_GELU, -2, "activation=%d is neither none (0) nor gelu (1)", act);
  QW_CHECK_ARG(!act || b_post, -1, "the fused activation needs post_conv.bias");
  if (is_general(d)) {
    QW_CHECK_ARG(!act, -2, "the fused activation needs the fast-path regime (n_qubits=4, amplitude embedding)");
    QW_CHECK_ARG(!chain, -2, "the chained backward needs the fast-path regime (n_qubits=4, amplitude embedding)");
    return gen::general_backward<T>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, ws_bytes, d, st);
  }
  QW_CHECK_ARG(d.K <= 8, -2, "backward supports kernel_size <= 8 (got %d)", d.K);
  if constexpr (sizeof(T) == 4) {
    if (fast_eligible(d, x, gy, gx, false) && (((uintptr_t)pre_save) & 15) == 0) {
      const FastPlan fp = make_fast_plan(d);
      QW_CHECK_ARG(ws_bytes >= fp.ws_bytes, -3, "workspace too small: %zu < %zu", ws_bytes, fp.ws_bytes);
      return fast_backward(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, d, st, dp, b_post, act, chain);
    }
  }
  QW_CHECK_ARG(!chain, -2, "the chained backward needs the fast-path regime (fp32, K=3, stride 1|2, L %% 4 == 0, O %% 4 == 0, aligned tensors)");
  QW_CHECK_ARG(!act, -2, "the fused activation needs the fast-path regime (fp32, K=3, stride 1|2, L %% 4 == 0, O %% 4 == 0, aligned tensors)");
  QW_CHECK_ARG(!want_dp, -2, "the fused gradient all-reduce needs the fast-path regime (fp32, K=3, stride 1|2, L %% 4 == 0, O %% 4 == 0, aligned tensors)");
  const Plan p = make_plan(d);
  const WsLayout<T> wl = ws_layout<T>(d, p);
  QW_CHECK_ARG(ws_bytes >= wl.total, -3, "workspace too small: %zu < %zu", ws_bytes, wl.total);
  switch (d.Q) {
    case 1: return bwd_tq<T, 1>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, d, p, st);
    case 2: return bwd_tq<T, 2>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, d, p, st);
    case 3: return bwd_tq<T, 3>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, d, p, st);
    default: return bwd_tq<T, 4>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, ws, d, p, st);
  }
}

static int circ_grid(long long W) {
  const long long need = (W + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
static int circ_PA(int q, int Lq) { return (int)align_up((size_t)Lq * q * 8, 32); }

template <typename T>
static int circuit_forward_impl(const T* pre, const T* qwts, T* out, long long W, int q, int Lq, int emb, void* stream) {
  QW_CHECK_ARG(pre && qwts && out && W > 0, -1, "null pointer or empty batch");
  QW_CHECK_ARG(q >= 1 && q <= 12, -2, "circuit kernels support n_qubits in [1,12] (got %d)", q);
  QW_CHECK_ARG(Lq >= 1 && Lq <= 8, -2, "n_layers=%d must be in [1,8]", Lq);
  QW_CHECK_ARG(emb == kEmbAmplitude || emb == kEmbAngle, -2, "embedding=%d not supported", emb);
  cudaStream_t st = (cudaStream_t)stream;
  if (q > 4 || emb != kEmbAmplitude) return wc::wcirc_forward<T>(pre, qwts, out, W, q, Lq, emb, st);
  CircArgs<T> a{pre, qwts, nullptr, out, nullptr, nullptr, W, q, Lq, 0};
  const int grid = circ_grid(W);
  switch (q) {
    case 1: return circ_fwd_tq<T, 1>(a, grid, st);
    case 2: return circ_fwd_tq<T, 2>(a, grid, st);
    case 3: return circ_fwd_tq<T, 3>(a, grid, st);
    default: return circ_fwd_tq<T, 4>(a, grid, st);
  }
}

template <typename T>
static int circuit_backward_impl(const T* pre, const T* qwts, const T* gout, T* gpre, T* gqw, void* workspace, size_t ws_bytes,
                                 long long W, int q, int Lq, int emb, void* stream) {
  QW_CHECK_ARG(pre && qwts && gout && gpre && gqw && workspace && W > 0, -1, "null pointer or empty batch");
  QW_CHECK_ARG(q >= 1 && q <= 12, -2, "circuit kernels support n_qubits in [1,12] (got %d)", q);
  QW_CHECK_ARG(Lq >= 1 && Lq <= 8, -2, "n_layers=%d must be in [1,8]", Lq);
  QW_CHECK_ARG(emb == kEmbAmplitude || emb == kEmbAngle, -2, "embedding=%d not supported", emb);
  cudaStream_t st = (cudaStream_t)stream;
  if (q > 4 || emb != kEmbAmplitude) {
    QW_CHECK_ARG(ws_bytes >= wc::wcirc_workspace_bytes(W, q, Lq, (int)sizeof(T)), -3, "workspace too small");
    return wc::wcirc_backward<T>(pre, qwts, gout, gpre, gqw, workspace, W, q, Lq, emb, st);
  }
  const int grid = circ_grid(W), PA = circ_PA(q, Lq);
  QW_CHECK_ARG(ws_bytes >= (size_t)grid * PA * sizeof(T), -3, "workspace too small");
  CircArgs<T> a{pre, qwts, gout, nullptr, gpre, (T*)workspace, W, q, Lq, PA};
  int e = 0;
  switch (q) {
    case 1: e = circ_bwd_tq<T, 1>(a, grid, st); break;
    case 2: e = circ_bwd_tq<T, 2>(a, grid, st); break;
    case 3: e = circ_bwd_tq<T, 3>(a, grid, st); break;
    default: e = circ_bwd_tq<T, 4>(a, grid, st); break;
  }
  if (e) return e;
  {
    KernelTimer kt(kKCircFinalize, st);
    circuit_finalize_kernel<T><<<PA / 32, kFinThreads, 0, st>>>((const T*)workspace, qwts, gqw, grid, PA, q, Lq);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qw

// =============================================================================================== C ABI
using qw::ConvDims;

extern "C" {


int qw_conv1d_forward(const float* x, const float* w_pre, const float* b_pre, const float* qwts, const float* w_post,
                      const float* b_post, float* y, float* pre_save, int B, int C, int L, int K, int S, int P, int O, int q,
                      int n_layers, int embedding, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_forward_impl<float>(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, stream);
}
int qw_conv1d_forward_f64(const double* x, const double* w_pre, const double* b_pre, const double* qwts, const double* w_post,
                          const double* b_post, double* y, double* pre_save, int B, int C, int L, int K, int S, int P, int O,
                          int q, int n_layers, int embedding, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_forward_impl<double>(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, stream);
}

size_t qw_conv1d_workspace_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int elem_size) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, 0, 0};
  if (qw::check_dims(d)) return 0;
  if (d.Q > 4) return qw::gen::general_workspace_bytes(d, elem_size == 8 ? 8 : 4);
  const size_t general = qw::gen::general_workspace_bytes(d, elem_size == 8 ? 8 : 4);  // angle embedding is chosen per call
  const qw::Plan p = qw::make_plan(d);
  if (elem_size == 8) {
    const size_t g64 = qw::ws_layout<double>(d, p).total;
    return g64 > general ? g64 : general;
  }
  const size_t generic = qw::ws_layout<float>(d, p).total;
  const size_t fast = qw::make_fast_plan(d).ws_bytes;
  const size_t m = generic > fast ? generic : fast;
  return m > general ? m : general;
}

int qw_conv1d_backward(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qwts,
                       const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post,
                       void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O, int q, int n_layers,
                       int embedding, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_backward_impl<float>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post,
                                         workspace, ws_bytes, d, stream);
}
int qw_conv1d_backward_f64(const double* gy, const double* x, const double* pre_save, const double* w_pre, const double* qwts,
                           const double* w_post, double* gx, double* gw_pre, double* gb_pre, double* gqw, double* gw_post,
                           double* gb_post, void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O,
                           int q, int n_layers, int embedding, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_backward_impl<double>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post,
                                          workspace, ws_bytes, d, stream);
}

int qw_conv1d_forward_act(const float* x, const float* w_pre, const float* b_pre, const float* qwts, const float* w_post,
                          const float* b_post, float* y, float* pre_save, int B, int C, int L, int K, int S, int P, int O, int q,
                          int n_layers, int embedding, int activation, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_forward_impl<float>(x, w_pre, b_pre, qwts, w_post, b_post, y, pre_save, d, stream, activation);
}
int qw_conv1d_backward_act(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qwts,
                           const float* w_post, const float* b_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post,
                           float* gb_post, void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O, int q,
                           int n_layers, int embedding, int activation, void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  return qw::conv1d_backward_impl<float>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post,
                                         workspace, ws_bytes, d, stream, nullptr, b_post, activation);
}

size_t qw_conv1d_dp_buffer_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int world) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, 0, 0};
  if (qw::check_dims(d) || d.Q != 4 || world < 1 || world > qw::kDpMaxWorld) return 0;
  return qw::fast_dp_buffer_bytes(qw::make_fast_plan(d), world);
}
size_t qw_conv1d_dp_flag_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int world) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, 0, 0};
  if (qw::check_dims(d) || d.Q != 4 || world < 1 || world > qw::kDpMaxWorld) return 0;
  return qw::fast_dp_flag_bytes(qw::make_fast_plan(d));
}
int qw_conv1d_backward_dp(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qwts,
                          const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post,
                          void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O, int q, int n_layers,
                          int embedding, void* const* peer_bufs, void* const* peer_flags, int rank, int world, float scale,
                          void* stream) {
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  QW_CHECK_ARG(world >= 1 && world <= qw::kDpMaxWorld && rank >= 0 && rank < world, -1, "qw_conv1d_backward_dp: bad rank/world %d/%d (world <= %d)",
               rank, world, qw::kDpMaxWorld);
  qw::FastDp dp{};
  if (world > 1) {
    QW_CHECK_ARG(peer_bufs && peer_flags, -1, "qw_conv1d_backward_dp: null peer pointer tables");
    for (int r = 0; r < world; ++r) {
      QW_CHECK_ARG(peer_bufs[r] && peer_flags[r], -1, "qw_conv1d_backward_dp: null peer pointer for rank %d", r);
      dp.bufs[r] = (float*)peer_bufs[r];
      dp.flags[r] = (unsigned*)peer_flags[r];
    }
  }
  dp.rank = rank;
  dp.world = world;
  dp.scale = scale;
  dp.timeout_ns = (unsigned long long)(qw::option(qw::kOptDpTimeoutMs) > 0 ? qw::option(qw::kOptDpTimeoutMs) : 0) * 1000000ull;
  return qw::conv1d_backward_impl<float>(gy, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post,
                                         workspace, ws_bytes, d, stream, &dp);
}

int qw_conv1d_backward_chained(const void* next_workspace, const float* next_w_pre, int next_O, const float* x, const float* pre_save,
                               const float* w_pre, const float* qwts, const float* w_post, const float* b_post, float* gx,
                               float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post, void* workspace,
                               size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int embedding,
                               int activation, void* const* peer_bufs, void* const* peer_flags, int rank, int world, float scale,
                               void* stream) {
  using namespace qw;
  ConvDims d{B, C, L, K, S, P, O, q, n_layers, embedding, 0};
  QW_CHECK_ARG(next_workspace && next_w_pre, -1, "qw_conv1d_backward_chained: null pointer argument");
  if (int e = check_dims(d)) return e;
  // the following layer: in_channels = this O, input length = this L_out, kernel_size 3, stride 2, padding 1, same circuit size
  ConvDims d2{B, O, d.Lout, 3, 2, 1, next_O, q, n_layers, embedding, 0};
  if (int e = check_dims(d2)) return e;
  QW_CHECK_ARG(q == 4 && embedding == kEmbAmplitude && option(kOptGyMma) && option(kOptFastPath) && (O * 3) % 4 == 0 &&
                   (((uintptr_t)next_w_pre) & 15) == 0 && (((uintptr_t)next_workspace) & 255) == 0 && d.Lout % 2 == 0 &&
                   fast_eligible(d2, x, x, x, false),
               -2, "qw_conv1d_backward_chained: outside the chained regime (n_qubits 4, amplitude embedding, tensor-pipe gy kernel, the "
                   "following layer a fast-path kernel_size-3 / stride-2 / padding-1 layer on this layer's output)");
  const FastPlan p2 = make_fast_plan(d2);
  FastChain chain{reinterpret_cast<const float*>(static_cast<const unsigned char*>(next_workspace) + p2.off_gpre), next_w_pre, p2.LP};
  QW_CHECK_ARG(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, -1, "qw_conv1d_backward_chained: bad rank/world %d/%d", rank, world);
  FastDp dp{};
  if (world > 1) {
    QW_CHECK_ARG(peer_bufs && peer_flags, -1, "qw_conv1d_backward_chained: null peer pointer tables");
    for (int r = 0; r < world; ++r) {
      QW_CHECK_ARG(peer_bufs[r] && peer_flags[r], -1, "qw_conv1d_backward_chained: null peer pointer for rank %d", r);
      dp.bufs[r] = (float*)peer_bufs[r];
      dp.flags[r] = (unsigned*)peer_flags[r];
    }
  }
  dp.rank = rank;
  dp.world = world;
  dp.scale = scale;
  dp.timeout_ns = (unsigned long long)(option(kOptDpTimeoutMs) > 0 ? option(kOptDpTimeoutMs) : 0) * 1000000ull;
  return conv1d_backward_impl<float>(nullptr, x, pre_save, w_pre, qwts, w_post, gx, gw_pre, gb_pre, gqw, gw_post, gb_post, workspace, ws_bytes,
                                     d, stream, world > 1 ? &dp : nullptr, b_post, activation, &chain);
}

size_t qw_circuit_workspace_bytes(long long W, int q, int n_layers, int elem_size) {
  if (W <= 0 || q < 1 || q > 12 || n_layers < 1) return 0;
  const int es = elem_size == 8 ? 8 : 4;
  const size_t warp = qw::wc::wcirc_workspace_bytes(W, q, n_layers, es);
  const size_t thread = q <= 4 ? (size_t)qw::circ_grid(W) * qw::circ_PA(q, n_layers) * (size_t)es : 0;
  return warp > thread ? warp : thread;
}
int qw_circuit_forward(const float* pre, const float* qwts, float* out, long long W, int q, int n_layers, int embedding,
                       void* stream) {
  return qw::circuit_forward_impl<float>(pre, qwts, out, W, q, n_layers, embedding, stream);
}
int qw_circuit_forward_f64(const double* pre, const double* qwts, double* out, long long W, int q, int n_layers, int embedding,
                           void* stream) {
  return qw::circuit_forward_impl<double>(pre, qwts, out, W, q, n_layers, embedding, stream);
}
int qw_circuit_backward(const float* pre, const float* qwts, const float* gout, float* gpre, float* gqw, void* workspace,
                        size_t ws_bytes, long long W, int q, int n_layers, int embedding, void* stream) {
  return qw::circuit_backward_impl<float>(pre, qwts, gout, gpre, gqw, workspace, ws_bytes, W, q, n_layers, embedding, stream);
}
int qw_circuit_backward_f64(const double* pre, const double* qwts, const double* gout, double* gpre, double* gqw,
                            void* workspace, size_t ws_bytes, long long W, int q, int n_layers, int embedding, void* stream) {
  return qw::circuit_backward_impl<double>(pre, qwts, gout, gpre, gqw, workspace, ws_bytes, W, q, n_layers, embedding, stream);
}

}  // extern "C"
