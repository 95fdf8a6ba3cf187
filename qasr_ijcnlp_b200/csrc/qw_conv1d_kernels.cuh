// QuantumConv1d forward / backward for sm_100a.
//
// Replaces /root/reference/quantum_whisper.py:95-128 (the Python loop over output columns and batch
// elements around a PennyLane QNode) and the autograd graph behind it.
//
// Kernels (all HBM-bound at q<=4; see DESIGN.md for the per-window byte counts):
//   qconv_fwd_kernel       x -> [window gather + pre_conv] -> [statevector circuit] -> [post_conv] -> y
//                          one CTA per tile of 32*WPT consecutive windows of one utterance; warps split the
//                          channel reduction (phase 1) and the output channels (phase 3); lanes run along
//                          the time axis so every global load/store is a coalesced 128-byte row segment.
//   qconv_bwd_post_kernel  gy -> gout (post_conv^T) -> adjoint circuit -> gpre ; partial sums for
//                          grad post_conv.{weight,bias}, grad pre_conv.bias and the gate-gradient matrices.
//   qconv_bwd_pre_kernel   gpre, x -> gx (pre_conv^T, overlap-add in gather form) and partial sums for
//                          grad pre_conv.weight; lanes run along channels on a shared-memory x tile.
//   qconv_bwd_finalize_kernel  deterministic reduction of the per-CTA partial rows + chain rule from gate
//                          matrices to (phi, theta, omega).
#pragma once
#include "qw_circuit.cuh"
#include "qw_common.cuh"

namespace qw {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

struct ConvDims {
  int B, C, L, K, S, P, O, Q, Lq, emb, Lout;
};

// ------------------------------------------------------------------------------------------ forward
template <typename T>
struct FwdArgs {
  const T *x, *w_pre, *b_pre, *qw, *w_post, *b_post;
  T *y, *pre_save;
  ConvDims d;
  int tiles_per_utt, num_tiles;
};

template <typename T, int Q>
__host__ __device__ inline size_t fwd_smem_elems(int CK, int O, int Lq, int TW) {
  return (size_t)CK * Q + (size_t)O * Q + align_up(O, 4) + 4 + (size_t)Lq * Q * kGateStride + (size_t)kWarps * TW * Q +
         (size_t)TW * Q;
}

template <typename T, int Q, int KT, int WPT>
__global__ void __launch_bounds__(kThreads) qconv_fwd_kernel(const FwdArgs<T> a) {
  constexpr int TW = 32 * WPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int CK = d.C * d.K;
  T* wpre_t = reinterpret_cast<T*>(smem_raw);           // [CK][Q]
  T* wpost = wpre_t + (size_t)CK * Q;                    // [O][Q]
  T* bpost = wpost + (size_t)d.O * Q;                    // [O]
  T* bpre = bpost + align_up(d.O, 4);                    // [4]
  T* gates = bpre + 4;                                   // [Lq][Q][16]
  T* part = gates + (size_t)d.Lq * Q * kGateStride;      // [kWarps][TW][Q]
  T* outs = part + (size_t)kWarps * TW * Q;              // [TW][Q]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- stage parameters once per CTA
  for (int idx = tid; idx < CK * Q; idx += kThreads) {
    const int j = idx / CK, f = idx - j * CK;
    wpre_t[f * Q + j] = a.w_pre[idx];
  }
  for (int idx = tid; idx < d.O * Q; idx += kThreads) wpost[idx] = a.w_post[idx];
  for (int idx = tid; idx < d.O; idx += kThreads) bpost[idx] = a.b_post[idx];
  if (tid < Q) bpre[tid] = a.b_pre[tid];
  if (tid < d.Lq * Q) make_gate<T>(a.qw + tid * 3, gates + tid * kGateStride);
  __syncthreads();

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * TW;
    const T* __restrict__ xb = a.x + (size_t)b * d.C * d.L;

    // ---- phase 1: pre_conv partial sums; warp w owns channels w, w+4, ...
    T acc[WPT][Q];
#pragma unroll
    for (int r = 0; r < WPT; ++r)
#pragma unroll
      for (int j = 0; j < Q; ++j) acc[r][j] = T(0);

    if constexpr (KT > 0) {
#pragma unroll 2
      for (int c = warp; c < d.C; c += kWarps) {
        const T* __restrict__ xr = xb + (size_t)c * d.L;
        T xv[WPT][KT];
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
          const int i = i0 + r * 32 + lane;
          const int lbase = i * d.S - d.P;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const int l = lbase + k;
            xv[r][k] = (i < d.Lout && l >= 0 && l < d.L) ? __ldg(xr + l) : T(0);
          }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          T wv[Q];
          ld_vec<T, Q>(wpre_t + (size_t)(c * KT + k) * Q, wv);
#pragma unroll
          for (int r = 0; r < WPT; ++r)
#pragma unroll
            for (int j = 0; j < Q; ++j) acc[r][j] = fma(wv[j], xv[r][k], acc[r][j]);
        }
      }
    } else {
      for (int f = warp; f < CK; f += kWarps) {
        const int c = f / d.K, k = f - c * d.K;
        const T* __restrict__ xr = xb + (size_t)c * d.L;
        T wv[Q];
        ld_vec<T, Q>(wpre_t + (size_t)f * Q, wv);
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
          const int i = i0 + r * 32 + lane;
          const int l = i * d.S - d.P + k;
          const T xv = (i < d.Lout && l >= 0 && l < d.L) ? __ldg(xr + l) : T(0);
#pragma unroll
          for (int j = 0; j < Q; ++j) acc[r][j] = fma(wv[j], xv, acc[r][j]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < WPT; ++r) st_vec<T, Q>(part + ((size_t)warp * TW + r * 32 + lane) * Q, acc[r]);
    __syncthreads();

    // ---- phase 2: one thread per window: bias, circuit, <Z_i>
    for (int t = tid; t < TW; t += kThreads) {
      const int i = i0 + t;
      T out[Q];
      if (i < d.Lout) {
        T pre[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) pre[j] = bpre[j];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          T pv[Q];
          ld_vec<T, Q>(part + ((size_t)w * TW + t) * Q, pv);
#pragma unroll
          for (int j = 0; j < Q; ++j) pre[j] += pv[j];
        }
        if (a.pre_save) st_vec<T, Q>(a.pre_save + ((size_t)b * d.Lout + i) * Q, pre);
        T re[1 << Q], im[1 << Q];
        circuit_forward_amp<T, Q>(pre, gates, d.Lq, re, im, out);
        if (a.pre_save) st_vec<T, Q>(a.pre_save + ((size_t)(d.B + b) * d.Lout + i) * Q, out);  // plane 1: <Z_i>
      } else {
#pragma unroll
        for (int j = 0; j < Q; ++j) out[j] = T(0);
      }
      st_vec<T, Q>(outs + (size_t)t * Q, out);
    }
    __syncthreads();

    // ---- phase 3: post_conv; warp w owns output channels w, w+4, ...
    T ov[WPT][Q];
#pragma unroll
    for (int r = 0; r < WPT; ++r) ld_vec<T, Q>(outs + (size_t)(r * 32 + lane) * Q, ov[r]);
    T* __restrict__ yb = a.y + (size_t)b * d.O * d.Lout;
#pragma unroll 4
    for (int o = warp; o < d.O; o += kWarps) {
      T wv[Q];
      ld_vec<T, Q>(wpost + (size_t)o * Q, wv);
      const T bv = bpost[o];
#pragma unroll
      for (int r = 0; r < WPT; ++r) {
        const int i = i0 + r * 32 + lane;
        T v = bv;
#pragma unroll
        for (int j = 0; j < Q; ++j) v = fma(wv[j], ov[r][j], v);
        if (i < d.Lout) yb[(size_t)o * d.Lout + i] = v;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ backward A (post + circuit)
template <typename T>
struct BwdAArgs {
  const T *gy, *pre_save, *qw, *w_post;
  T *gpre, *part;  // part: [gridDim.x][PA]
  ConvDims d;
  int tiles_per_utt, num_tiles, PA;
};

// Partial-row layout of buffer A (length PA): [O*Q gw_post][O gb_post][Q gb_pre][pad to 32][Lq*Q*8 gate matrices]
__host__ __device__ inline int partA_moff(int O, int Q) { return (int)align_up((size_t)O * (Q + 1) + Q, 32); }
__host__ __device__ inline int partA_len(int O, int Q, int Lq) { return partA_moff(O, Q) + (int)align_up((size_t)Lq * Q * 8, 32); }

constexpr int kOC = 64;               // output-channel rows staged per chunk
constexpr int kNH = kThreads / kOC;   // time halves in the o-major pass

template <typename T, int Q>
__host__ __device__ inline size_t bwdA_smem_elems(int O, int Lq, int TW) {
  const int NE = Q + Lq * Q * 8;
  return (size_t)kOC * (TW + 4) + (size_t)TW * Q + (size_t)kWarps * TW * Q + align_up((size_t)kNH * O * (Q + 1), 4) +
         align_up((size_t)NE * (TW + 1), 4) + (size_t)O * Q + (size_t)Lq * Q * kGateStride;
}

template <typename T, int Q>
struct SmemGateAcc {
  T* base;  // &macc[Q][t]  (entries for M start after the Q gb_pre rows)
  int stride, layer;
  __device__ __forceinline__ void set_layer(int l) { layer = l; }
  __device__ __forceinline__ void add(int wire, const T (&m)[8]) {
    T* p = base + (size_t)((layer * Q + wire) * 8) * stride;
#pragma unroll
    for (int e = 0; e < 8; ++e) p[e * stride] += m[e];
  }
};

template <typename T, int Q, int WPT>
__global__ void __launch_bounds__(kThreads) qconv_bwd_post_kernel(const BwdAArgs<T> a) {
  constexpr int TW = 32 * WPT;
  constexpr int GS = TW + 4;  // gy_s row stride (== 4 mod 32 words: conflict-free 4-wide column reads)
  constexpr int MS = TW + 1;  // macc row stride
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int NE = Q + d.Lq * Q * 8;
  T* gy_s = reinterpret_cast<T*>(smem_raw);                       // [kOC][GS]
  T* out_s = gy_s + (size_t)kOC * GS;                              // [TW][Q]
  T* gred = out_s + (size_t)TW * Q;                                // [kWarps][TW][Q]
  T* accp = gred + (size_t)kWarps * TW * Q;                        // [kNH][O][Q+1]
  T* macc = accp + align_up((size_t)kNH * d.O * (Q + 1), 4);       // [NE][MS]
  T* wpost = macc + align_up((size_t)NE * MS, 4);                  // [O][Q]
  T* gates = wpost + (size_t)d.O * Q;                              // [Lq][Q][16]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int idx = tid; idx < d.O * Q; idx += kThreads) wpost[idx] = a.w_post[idx];
  for (int idx = tid; idx < kNH * d.O * (Q + 1); idx += kThreads) accp[idx] = T(0);
  for (int idx = tid; idx < NE * MS; idx += kThreads) macc[idx] = T(0);
  if (tid < d.Lq * Q) make_gate<T>(a.qw + tid * 3, gates + tid * kGateStride);
  __syncthreads();

  const int nchunks = (d.O + kOC - 1) / kOC;
  const int o_loc = tid % kOC, th = tid / kOC;  // o-major pass mapping
  constexpr int TH = TW / kNH;                  // time columns per half

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * TW;
    const T* __restrict__ gyb = a.gy + (size_t)b * d.O * d.Lout;

    // ---- phase 0: forward circuit from the saved pre_conv output -> out_s (for grad post_conv.weight)
    for (int t = tid; t < TW; t += kThreads) {
      const int i = i0 + t;
      T out[Q];
      if (i < d.Lout) {
        T pre[Q];
        ld_vec<T, Q>(a.pre_save + ((size_t)b * d.Lout + i) * Q, pre);
        T re[1 << Q], im[1 << Q];
        circuit_forward_amp<T, Q>(pre, gates, d.Lq, re, im, out);
      } else {
#pragma unroll
        for (int j = 0; j < Q; ++j) out[j] = T(0);
      }
      st_vec<T, Q>(out_s + (size_t)t * Q, out);
    }
    __syncthreads();

    T gacc[WPT][Q];
#pragma unroll
    for (int r = 0; r < WPT; ++r)
#pragma unroll
      for (int j = 0; j < Q; ++j) gacc[r][j] = T(0);

    for (int ch = 0; ch < nchunks; ++ch) {
      // ---- phase 1: stream kOC rows of gy (time-major lanes), gout partials, stage into smem
      constexpr int RPW = kOC / kWarps;
#pragma unroll 4
      for (int rr = 0; rr < RPW; ++rr) {
        const int row = warp * RPW + rr;
        const int o = ch * kOC + row;
        T g[WPT];
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
          const int i = i0 + r * 32 + lane;
          g[r] = (o < d.O && i < d.Lout) ? __ldg(gyb + (size_t)o * d.Lout + i) : T(0);
        }
        if (o < d.O) {
          T wv[Q];
          ld_vec<T, Q>(wpost + (size_t)o * Q, wv);
#pragma unroll
          for (int r = 0; r < WPT; ++r)
#pragma unroll
            for (int j = 0; j < Q; ++j) gacc[r][j] = fma(g[r], wv[j], gacc[r][j]);
        }
#pragma unroll
        for (int r = 0; r < WPT; ++r) gy_s[row * GS + r * 32 + lane] = g[r];
      }
      __syncthreads();
      // ---- phase 2: lanes along output channels: grad post_conv.{weight,bias} partial sums
      {
        const int o = ch * kOC + o_loc;
        if (o < d.O) {
          T s[Q + 1];
#pragma unroll
          for (int j = 0; j <= Q; ++j) s[j] = T(0);
          const T* __restrict__ grow = gy_s + o_loc * GS + th * TH;
          const T* __restrict__ orow = out_s + (size_t)th * TH * Q;
#pragma unroll 2
          for (int t = 0; t < TH; t += 4) {
            T gv[4];
            ld_vec<T, 4>(grow + t, gv);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              T ov[Q];
              ld_vec<T, Q>(orow + (size_t)(t + u) * Q, ov);
#pragma unroll
              for (int j = 0; j < Q; ++j) s[j] = fma(gv[u], ov[j], s[j]);
              s[Q] += gv[u];
            }
          }
          T* ap = accp + ((size_t)th * d.O + o) * (Q + 1);
#pragma unroll
          for (int j = 0; j <= Q; ++j) ap[j] += s[j];
        }
      }
      __syncthreads();
    }

    // ---- phase 3: finish gout across warps, adjoint circuit, gpre
#pragma unroll
    for (int r = 0; r < WPT; ++r) st_vec<T, Q>(gred + ((size_t)warp * TW + r * 32 + lane) * Q, gacc[r]);
    __syncthreads();
    for (int t = tid; t < TW; t += kThreads) {
      const int i = i0 + t;
      if (i < d.Lout) {
        T gout[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) gout[j] = T(0);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          T pv[Q];
          ld_vec<T, Q>(gred + ((size_t)w * TW + t) * Q, pv);
#pragma unroll
          for (int j = 0; j < Q; ++j) gout[j] += pv[j];
        }
        T pre[Q], out[Q], gpre[Q];
        ld_vec<T, Q>(a.pre_save + ((size_t)b * d.Lout + i) * Q, pre);
        T re[1 << Q], im[1 << Q];
        const T inv = circuit_forward_amp<T, Q>(pre, gates, d.Lq, re, im, out);
        SmemGateAcc<T, Q> acc{macc + (size_t)Q * MS + t, MS, 0};
        circuit_backward_amp<T, Q>(pre, inv, gates, d.Lq, re, im, gout, gpre, acc);
        st_vec<T, Q>(a.gpre + ((size_t)b * d.Lout + i) * Q, gpre);
#pragma unroll
        for (int j = 0; j < Q; ++j) macc[j * MS + t] += gpre[j];  // grad pre_conv.bias
      }
    }
    __syncthreads();
  }

  // ---- write this CTA's partial row: [O*Q gw_post][O gb_post][Q gb_pre][Lq*Q*8 M]
  T* prow = a.part + (size_t)blockIdx.x * a.PA;
  for (int idx = tid; idx < d.O * (Q + 1); idx += kThreads) {
    const int o = idx / (Q + 1), j = idx - o * (Q + 1);
    T s = T(0);
#pragma unroll
    for (int h = 0; h < kNH; ++h) s += accp[((size_t)h * d.O + o) * (Q + 1) + j];
    if (j < Q) prow[o * Q + j] = s; else prow[d.O * Q + o] = s;
  }
  const int moff = partA_moff(d.O, Q);
  for (int e = warp; e < NE; e += kWarps) {
    T s = T(0);
    for (int t = lane; t < TW; t += 32) s += macc[e * MS + t];
    s = warp_sum(s);
    if (lane == 0) prow[e < Q ? d.O * (Q + 1) + e : moff + (e - Q)] = s;
  }
  for (int e = d.O * (Q + 1) + Q + tid; e < moff; e += kThreads) prow[e] = T(0);
  for (int e = moff + d.Lq * Q * 8 + tid; e < a.PA; e += kThreads) prow[e] = T(0);
}

// ------------------------------------------------------------------------------ backward B (pre_conv^T)
template <typename T>
struct BwdBArgs {
  const T *x, *gpre, *w_pre;
  T *gx, *part;  // part: [gridDim.x][Cpad][Q*KT]
  ConvDims d;
  int tiles_per_utt, num_tiles, Cpad;
};

constexpr int kTP = 128;       // positions per tile
constexpr int kXS = kTP + 4;   // smem row stride

template <typename T, int Q, int KT>
__host__ __device__ inline size_t bwdB_smem_elems() {
  const size_t tile = (size_t)32 * kXS + (size_t)(kTP + KT + 3) * Q;
  const size_t red = (size_t)kWarps * 32 * Q * KT;
  return tile > red ? tile : red;
}

template <typename T, int Q, int KT>
__global__ void __launch_bounds__(kThreads) qconv_bwd_pre_kernel(const BwdBArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  T* x_s = reinterpret_cast<T*>(smem_raw);   // [32][kXS]
  T* gp_s = x_s + (size_t)32 * kXS;           // [<= kTP+KT+3][Q]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * 32;
  const int c = c0 + lane;
  const int CK = d.C * d.K;

  T w[Q][KT];
#pragma unroll
  for (int j = 0; j < Q; ++j)
#pragma unroll
    for (int k = 0; k < KT; ++k) w[j][k] = (c < d.C && k < d.K) ? a.w_pre[(size_t)j * CK + c * d.K + k] : T(0);
  T gw[Q][KT];
#pragma unroll
  for (int j = 0; j < Q; ++j)
#pragma unroll
    for (int k = 0; k < KT; ++k) gw[j][k] = T(0);

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt;
    const int l0 = (tile - b * a.tiles_per_utt) * kTP;
    const int i_lo = floor_div(l0 + d.P - (d.K - 1), d.S);
    const int nwin = floor_div(l0 + kTP - 1 + d.P, d.S) - i_lo + 1;
    // ---- stage the x tile (lanes along time) and the gpre rows of every window touching it
    for (int idx = tid; idx < 32 * kTP; idx += kThreads) {
      const int cl = idx / kTP, l = idx - cl * kTP;
      const int cc = c0 + cl, lg = l0 + l;
      x_s[cl * kXS + l] = (cc < d.C && lg < d.L) ? __ldg(a.x + ((size_t)b * d.C + cc) * d.L + lg) : T(0);
    }
    for (int idx = tid; idx < nwin; idx += kThreads) {
      const int i = i_lo + idx;
      T g[Q];
      if (i >= 0 && i < d.Lout) {
        ld_vec<T, Q>(a.gpre + ((size_t)b * d.Lout + i) * Q, g);
      } else {
#pragma unroll
        for (int j = 0; j < Q; ++j) g[j] = T(0);
      }
      st_vec<T, Q>(gp_s + (size_t)idx * Q, g);
    }
    __syncthreads();
    // ---- lanes along channels; warp w owns positions [32w, 32w+32)
#pragma unroll 2
    for (int u = 0; u < 8; ++u) {
      const int lw = warp * 32 + u * 4;
      T xv[4], gxv[4];
      ld_vec<T, 4>(x_s + lane * kXS + lw, xv);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int lg = l0 + lw + p;
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < d.K) {
            const int num = lg + d.P - k;
            if (num % d.S == 0) {
              const int idx = num / d.S - i_lo;
              T g[Q];
              ld_vec<T, Q>(gp_s + (size_t)idx * Q, g);
#pragma unroll
              for (int j = 0; j < Q; ++j) {
                gw[j][k] = fma(g[j], xv[p], gw[j][k]);
                acc = fma(g[j], w[j][k], acc);
              }
            }
          }
        }
        gxv[p] = acc;
      }
      st_vec<T, 4>(x_s + lane * kXS + lw, gxv);
    }
    __syncthreads();
    if (a.gx) {
      for (int idx = tid; idx < 32 * kTP; idx += kThreads) {
        const int cl = idx / kTP, l = idx - cl * kTP;
        const int cc = c0 + cl, lg = l0 + l;
        if (cc < d.C && lg < d.L) a.gx[((size_t)b * d.C + cc) * d.L + lg] = x_s[cl * kXS + l];
      }
    }
    __syncthreads();
  }
  // ---- cross-warp reduction of the weight-gradient partials, one row per CTA
  T* red = reinterpret_cast<T*>(smem_raw);  // [kWarps][32][Q*KT]
#pragma unroll
  for (int j = 0; j < Q; ++j)
#pragma unroll
    for (int k = 0; k < KT; ++k) red[((size_t)warp * 32 + lane) * (Q * KT) + j * KT + k] = gw[j][k];
  __syncthreads();
  T* prow = a.part + ((size_t)blockIdx.x * a.Cpad + c0) * (Q * KT);
  for (int e = tid; e < 32 * Q * KT; e += kThreads) {
    T s = T(0);
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) s += red[(size_t)wq * 32 * Q * KT + e];
    prow[e] = s;
  }
}

// ------------------------------------------------------------------------------ finalize

template <typename T>
struct FinArgs {
  const T *partA, *partB, *qw;
  T *gw_pre, *gb_pre, *gqw, *gw_post, *gb_post;
  int GA, PA, GB, PB;  // rows / row length of the two partial buffers (PB = 0: no buffer B)
  int C, K, O, Q, Lq, KT;
};

constexpr int kFinThreads = 256;
constexpr int kFinWarps = kFinThreads / 32;

// One block = 32 consecutive entries of a partial row; warp w sums rows w, w+8, ... in double, then the 8
// warp sums are combined in a fixed order (bitwise reproducible for a given grid).
template <typename T>
__global__ void __launch_bounds__(kFinThreads) qconv_bwd_finalize_kernel(const FinArgs<T> a) {
  __shared__ double red[kFinWarps][33];
  __shared__ double tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblkA = a.PA / 32;
  const bool isA = (int)blockIdx.x < nblkA;
  const int blk0 = (isA ? (int)blockIdx.x : (int)blockIdx.x - nblkA) * 32;
  const int p = blk0 + lane;
  const int G = isA ? a.GA : a.GB, P = isA ? a.PA : a.PB;
  const T* __restrict__ part = isA ? a.partA : a.partB;
  double s = 0.0;
  if (p < P) {
    int g = warp;
    for (; g + 3 * kFinWarps < G; g += 4 * kFinWarps) {
      const T v0 = part[(size_t)g * P + p], v1 = part[(size_t)(g + kFinWarps) * P + p];
      const T v2 = part[(size_t)(g + 2 * kFinWarps) * P + p], v3 = part[(size_t)(g + 3 * kFinWarps) * P + p];
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; g < G; g += kFinWarps) s += (double)part[(size_t)g * P + p];
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp != 0) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kFinWarps; ++w) t += red[w][lane];
  tot[lane] = t;
  __syncwarp();
  if (!isA) {
    const int QK = a.Q * a.KT;
    const int c = p / QK, r = p - c * QK, j = r / a.KT, k = r - j * a.KT;
    if (p < P && c < a.C && k < a.K) a.gw_pre[(size_t)j * a.C * a.K + c * a.K + k] = (T)t;
    return;
  }
  const int nW = a.O * a.Q, moff = partA_moff(a.O, a.Q);
  if (p < nW) a.gw_post[p] = (T)t;
  else if (p < nW + a.O) a.gb_post[p - nW] = (T)t;
  else if (p < nW + a.O + a.Q) a.gb_pre[p - nW - a.O] = (T)t;
  if (blk0 >= moff && lane < 4) {
    const int gi = (blk0 - moff) / 8 + lane;  // (layer*Q + wire)
    if (gi < a.Lq * a.Q) {
      double w3[3], g3[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) w3[e] = (double)a.qw[gi * 3 + e];
      gate_grad_to_angles(w3, &tot[lane * 8], g3);
#pragma unroll
      for (int e = 0; e < 3; ++e) a.gqw[gi * 3 + e] = (T)g3[e];
    }
  }
}

// ------------------------------------------------------------------------------ circuit-only kernels (config 4)
template <typename T>
struct CircArgs {
  const T *pre, *qw, *gout;
  T *out, *gpre, *part;
  long long W;
  int Q, Lq, PA;
};

template <typename T, int Q>
__global__ void __launch_bounds__(kThreads) circuit_fwd_kernel(const CircArgs<T> a) {
  __shared__ __align__(16) T gates[8 * 4 * kGateStride];
  if ((int)threadIdx.x < a.Lq * Q) make_gate<T>(a.qw + threadIdx.x * 3, gates + threadIdx.x * kGateStride);
  __syncthreads();
  for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < a.W; w += (long long)gridDim.x * kThreads) {
    T pre[Q], out[Q], re[1 << Q], im[1 << Q];
    ld_vec<T, Q>(a.pre + w * Q, pre);
    circuit_forward_amp<T, Q>(pre, gates, a.Lq, re, im, out);
    st_vec<T, Q>(a.out + w * Q, out);
  }
}

template <typename T, int Q>
struct RegGateAccSmem {
  T* base;
  int stride, layer;
  __device__ __forceinline__ void set_layer(int l) { layer = l; }
  __device__ __forceinline__ void add(int wire, const T (&m)[8]) {
    T* p = base + (size_t)((layer * Q + wire) * 8) * stride;
#pragma unroll
    for (int e = 0; e < 8; ++e) p[e * stride] += m[e];
  }
};

template <typename T, int Q>
__global__ void __launch_bounds__(kThreads) circuit_bwd_kernel(const CircArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MS = kThreads + 1;
  const int NE = a.Lq * Q * 8;
  T* gates = reinterpret_cast<T*>(smem_raw);              // [Lq][Q][16]
  T* macc = gates + (size_t)a.Lq * Q * kGateStride;       // [NE][MS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < a.Lq * Q) make_gate<T>(a.qw + tid * 3, gates + tid * kGateStride);
  for (int idx = tid; idx < NE * MS; idx += kThreads) macc[idx] = T(0);
  __syncthreads();
  for (long long w = (long long)blockIdx.x * kThreads + tid; w < a.W; w += (long long)gridDim.x * kThreads) {
    T pre[Q], out[Q], gout[Q], gpre[Q], re[1 << Q], im[1 << Q];
    ld_vec<T, Q>(a.pre + w * Q, pre);
    ld_vec<T, Q>(a.gout + w * Q, gout);
    const T inv = circuit_forward_amp<T, Q>(pre, gates, a.Lq, re, im, out);
    RegGateAccSmem<T, Q> acc{macc + tid, MS, 0};
    circuit_backward_amp<T, Q>(pre, inv, gates, a.Lq, re, im, gout, gpre, acc);
    st_vec<T, Q>(a.gpre + w * Q, gpre);
  }
  __syncthreads();
  T* prow = a.part + (size_t)blockIdx.x * a.PA;
  for (int e = warp; e < NE; e += kWarps) {
    T s = T(0);
    for (int t = lane; t < kThreads; t += 32) s += macc[e * MS + t];
    s = warp_sum(s);
    if (lane == 0) prow[e] = s;
  }
  for (int e = NE + tid; e < a.PA; e += kThreads) prow[e] = T(0);
}

// circuit-only finalize: PA = align32(Lq*Q*8), gate matrices at offset 0
template <typename T>
__global__ void __launch_bounds__(kFinThreads) circuit_finalize_kernel(const T* __restrict__ part, const T* __restrict__ qw,
                                                                      T* __restrict__ gqw, int G, int PA, int Q, int Lq) {
  __shared__ double red[kFinWarps][33];
  __shared__ double tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk0 = blockIdx.x * 32, p = blk0 + lane;
  double s = 0.0;
  if (p < PA)
    for (int g = warp; g < G; g += kFinWarps) s += (double)part[(size_t)g * PA + p];
  red[warp][lane] = s;
  __syncthreads();
  if (warp != 0) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kFinWarps; ++w) t += red[w][lane];
  tot[lane] = t;
  __syncwarp();
  if (lane < 4) {
    const int gi = blk0 / 8 + lane;
    if (gi < Lq * Q) {
      double w3[3], g3[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) w3[e] = (double)qw[gi * 3 + e];
      gate_grad_to_angles(w3, &tot[lane * 8], g3);
#pragma unroll
      for (int e = 0; e < 3; ++e) gqw[gi * 3 + e] = (T)g3[e];
    }
  }
}


}  // namespace qw
