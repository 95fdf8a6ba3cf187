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


// ---- fast path (qw_conv1d_fast.cu): fp32, q = 4, K = 3, stride 1|2, aligned shapes
struct FastPlan {
  int tw, tiles_per_utt, num_tiles, rc, chunks_per_tile, gridF;
  int gridGy, PA1, gridAdj, PA2, LP;
  int ptiles_per_utt, num_ptiles, nchunks, Cpad, gridPx, PB;
  size_t off_gout, off_gpre, off_p1, off_p2, off_p3, ws_bytes;
  bool small;  // latency-dominated launch: use programmatic dependent launch
};
constexpr int kHaloL = 8;    // gpre is stored halo-padded per utterance: [B][kHaloL + L_out + kHaloR][4]
constexpr int kHaloR = 144;
// data-parallel context of the fused finalize + gradient all-reduce (fast_finalize_kernel)
constexpr int kDpMaxWorld = 8;
struct FastDp {
  float* bufs[kDpMaxWorld];      // peer r's receive buffer: [2 slots][world][ncol] 64-bit words (epoch << 32 | float bits)
  unsigned* flags[kDpMaxWorld];  // only [rank] is used: epoch[cta], status (local bookkeeping)
  int rank, world, ncol;
  float scale;
  unsigned long long timeout_ns;  // 0 = wait forever
};
inline int fast_dp_columns(const FastPlan& p) { return p.PA1 + p.PA2 + p.PB; }
inline size_t fast_dp_buffer_bytes(const FastPlan& p, int world) { return (size_t)2 * world * fast_dp_columns(p) * 8; }
inline size_t fast_dp_flag_bytes(const FastPlan& p) { return ((size_t)fast_dp_columns(p) / 32 + 1) * sizeof(unsigned); }
struct FastAdjArgs {
  const float *pre_save, *gout, *qw;
  float *gpre_pad, *part;  // part: [grid][PA2] rows: [grad pre_conv.bias 4 + pad 28][Lq*32 gate-gradient matrices M]
  int B, Lout, LP, Lq, PA2;
  long long W;
  int early_trigger;  // experiment switch QW_ADJ_TRIG: trigger the dependent launch right after the wait
  unsigned long long* tl;
};
// chained backward: this layer's incoming gradient is the grad_x of the FOLLOWING layer (kernel_size 3, stride 2, padding 1), which
// its own backward did not write (gx = NULL); the gy kernel rebuilds it from that layer's gpre rows and pre_conv weights
struct FastChain {
  const float* gpre_pad;  // the following layer's halo-padded gpre, [B][LP][4] (inside ITS backward workspace)
  const float* w_pre;     // the following layer's pre_conv.weight (4, O*3)
  int LP;
};
bool fast_eligible(const ConvDims& d, const void* x, const void* y_or_gy, const void* gx, bool fwd);
FastPlan make_fast_plan(const ConvDims& d);
int fast_forward(const float* x, const float* w_pre, const float* b_pre, const float* qwts, const float* w_post,
                 const float* b_post, float* y, float* pre_save, const ConvDims& d, cudaStream_t st, int act = 0);
int fast_backward(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qwts,
                  const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post,
                  unsigned char* ws, const ConvDims& d, cudaStream_t st, const FastDp* dp = nullptr, const float* b_post = nullptr,
                  int act = 0, const FastChain* chain = nullptr);

// fused inference stem (qw_stem.cu)
size_t stem_workspace_bytes(int B, int L);
int stem_forward(const float* x, const float* const* p1, const float* const* p2, const float* pos, float* out, void* ws, size_t ws_bytes,
                 int B, int Cin, int L, int Cmid, int O, int Lq, cudaStream_t st);

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
