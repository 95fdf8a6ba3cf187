// General QuantumConv1d path: n_qubits 5..12 and/or angle embedding (any kernel_size, stride, padding; fp32 / fp64).
// The reference accepts any --n_qubits (train_quantum_whisper.py:422, quantum_whisper.py:55); the fused kernels cover
// q <= 4 amplitude (the shipped default), this file composes everything else from streaming kernels around the
// thread-group cooperative circuit simulator (qw_circuit_warp.cuh):
//
//   forward   gen_preconv_fwd (window gather + pre_conv, quantum_whisper.py:107-114) -> wcirc_forward (:64-85)
//             -> gen_postconv_fwd (:125-126).  pre and <Z> land in pre_save (2, W, q), which the backward re-uses.
//   backward  gen_postconv_bwd (gout = W_post^T gy, partials of grad post_conv.{weight,bias}) -> wcirc_backward
//             (gpre, grad quantum_weights) -> gen_preconv_bwd_gx (overlap-add in gather form, no atomics) and
//             gen_preconv_bwd_gw (partials of grad pre_conv.{weight,bias}) -> gen_reduce_rows (deterministic, fp64).
#include "../../include/qw.h"
#include "qw_circuit_warp_host.h"
#include "qw_conv1d_kernels.cuh"

namespace qw {
namespace gen {

constexpr int kT = 128;   // threads per CTA
constexpr int kTW = 32;   // windows per tile (lanes along time)
constexpr int kOC = 128;  // gy rows staged per chunk in the backward

template <typename T>
struct GArgs {
  const T *x, *w_pre, *b_pre, *w_post, *b_post, *gy, *pre, *qout, *gout_c, *gpre_c;
  T *y, *pre_w, *gout, *gx, *part;
  ConvDims d;
  int tiles_per_utt, num_tiles, q, PA, NS;
};

// ---------------------------------------------------------------- forward: window gather + pre_conv -> pre (W, q)
template <typename T, int QP>
__global__ void __launch_bounds__(kT) gen_preconv_fwd_kernel(const GArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int CK = d.C * d.K, q = a.q;
  T* wt = reinterpret_cast<T*>(smem_raw);  // [CK][QP] zero padded
  T* part = wt + (size_t)CK * QP;          // [4][32][QP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < CK * QP; idx += kT) {
    const int f = idx / QP, j = idx - f * QP;
    wt[idx] = j < q ? a.w_pre[(size_t)j * CK + f] : T(0);
  }
  __syncthreads();
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt, i0 = (tile - b * a.tiles_per_utt) * kTW, i = i0 + lane;
    const T* __restrict__ xb = a.x + (size_t)b * d.C * d.L;
    T acc[QP];
#pragma unroll
    for (int j = 0; j < QP; ++j) acc[j] = T(0);
    // channels over the 4 warps, taps inside (no division per feature; the K loads of a channel are issued together)
    const int l0 = i * d.S - d.P;  // quantum_whisper.py:107-110: padded column i*S + k = original column i*S - P + k
    const bool iv = i < d.Lout;
    for (int c = warp; c < d.C; c += 4) {
      const T* __restrict__ xr = xb + (size_t)c * d.L;
      const T* __restrict__ wr = wt + (size_t)c * d.K * QP;
#pragma unroll 3
      for (int k = 0; k < d.K; ++k) {
        const int l = l0 + k;
        const T xv = (iv && l >= 0 && l < d.L) ? __ldg(xr + l) : T(0);
#pragma unroll
        for (int j = 0; j < QP; ++j) acc[j] = fma(wr[k * QP + j], xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < QP; ++j) part[((size_t)warp * kTW + lane) * QP + j] = acc[j];
    __syncthreads();
    for (int idx = tid; idx < kTW * q; idx += kT) {
      const int ii = idx / q, j = idx - ii * q;
      if (i0 + ii < d.Lout) {
        T s = a.b_pre[j];
#pragma unroll
        for (int w = 0; w < 4; ++w) s += part[((size_t)w * kTW + ii) * QP + j];
        a.pre_w[((size_t)b * d.Lout + i0 + ii) * q + j] = s;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- forward: post_conv, column store
template <typename T, int QP>
__global__ void __launch_bounds__(kT) gen_postconv_fwd_kernel(const GArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int q = a.q;
  T* wp = reinterpret_cast<T*>(smem_raw);  // [O][QP]
  T* bp = wp + (size_t)d.O * QP;           // [O]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < d.O * QP; idx += kT) {
    const int o = idx / QP, j = idx - o * QP;
    wp[idx] = j < q ? a.w_post[(size_t)o * q + j] : T(0);
  }
  for (int o = tid; o < d.O; o += kT) bp[o] = a.b_post[o];
  __syncthreads();
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt, i = (tile - b * a.tiles_per_utt) * kTW + lane;
    const bool valid = i < d.Lout;
    T qv[QP];
#pragma unroll
    for (int j = 0; j < QP; ++j) qv[j] = (valid && j < q) ? a.qout[((size_t)b * d.Lout + i) * q + j] : T(0);
    T* __restrict__ yb = a.y + (size_t)b * d.O * d.Lout + i;
    for (int o = warp; o < d.O; o += 4) {
      T s = bp[o];
#pragma unroll
      for (int j = 0; j < QP; ++j) s = fma(wp[(size_t)o * QP + j], qv[j], s);
      if (valid) yb[(size_t)o * d.Lout] = s;
    }
  }
}

// ---------------------------------------------------------------- backward: one pass over gy
template <typename T, int QP>
__global__ void __launch_bounds__(kT) gen_postconv_bwd_kernel(const GArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int q = a.q;
  T* wp = reinterpret_cast<T*>(smem_raw);    // [O][QP]
  T* gwacc = wp + (size_t)d.O * QP;          // [O][QP+1]
  T* gys = gwacc + (size_t)d.O * (QP + 1);   // [kOC][33]
  T* qs = gys + (size_t)kOC * 33;            // [32][QP]
  T* red = qs + (size_t)kTW * QP;            // [4][32][QP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < d.O * QP; idx += kT) {
    const int o = idx / QP, j = idx - o * QP;
    wp[idx] = j < q ? a.w_post[(size_t)o * q + j] : T(0);
  }
  for (int idx = tid; idx < d.O * (QP + 1); idx += kT) gwacc[idx] = T(0);
  __syncthreads();
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt, i0 = (tile - b * a.tiles_per_utt) * kTW, i = i0 + lane;
    const bool valid = i < d.Lout;
    for (int idx = tid; idx < kTW * QP; idx += kT) {
      const int ii = idx / QP, j = idx - ii * QP;
      qs[idx] = (i0 + ii < d.Lout && j < q) ? a.qout[((size_t)b * d.Lout + i0 + ii) * q + j] : T(0);
    }
    T gacc[QP];
#pragma unroll
    for (int j = 0; j < QP; ++j) gacc[j] = T(0);
    const T* __restrict__ gyb = a.gy + (size_t)b * d.O * d.Lout + i;
    // the next chunk's 32 values per thread are fetched into registers while the current chunk is contracted: with the loads issued
    // right before their first use every chunk paid a full memory round trip between two barriers (116 us per launch)
    T nx[kOC / 4];
#pragma unroll
    for (int u = 0; u < kOC / 4; ++u) nx[u] = (valid && warp + 4 * u < d.O) ? __ldg(gyb + (size_t)(warp + 4 * u) * d.Lout) : T(0);
    for (int oc = 0; oc < d.O; oc += kOC) {
#pragma unroll
      for (int u = 0; u < kOC / 4; ++u) gys[(warp + 4 * u) * 33 + lane] = nx[u];
      __syncthreads();
      if (oc + kOC < d.O) {
#pragma unroll
        for (int u = 0; u < kOC / 4; ++u) {
          const int r = oc + kOC + warp + 4 * u;
          nx[u] = (valid && r < d.O) ? __ldg(gyb + (size_t)r * d.Lout) : T(0);
        }
      }
      for (int r = warp; r < kOC && oc + r < d.O; r += 4) {
        const T g = gys[r * 33 + lane];
#pragma unroll
        for (int j = 0; j < QP; ++j) gacc[j] = fma(g, wp[(size_t)(oc + r) * QP + j], gacc[j]);
      }
      if (oc + tid < d.O) {
        T s[QP + 1];
#pragma unroll
        for (int j = 0; j <= QP; ++j) s[j] = T(0);
        for (int ii = 0; ii < kTW; ++ii) {
          const T g = gys[tid * 33 + ii];
#pragma unroll
          for (int j = 0; j < QP; ++j) s[j] = fma(g, qs[ii * QP + j], s[j]);
          s[QP] += g;
        }
#pragma unroll
        for (int j = 0; j <= QP; ++j) gwacc[(size_t)(oc + tid) * (QP + 1) + j] += s[j];
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < QP; ++j) red[((size_t)warp * kTW + lane) * QP + j] = gacc[j];
    __syncthreads();
    for (int idx = tid; idx < kTW * q; idx += kT) {
      const int ii = idx / q, j = idx - ii * q;
      if (i0 + ii < d.Lout) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < 4; ++w) s += red[((size_t)w * kTW + ii) * QP + j];
        a.gout[((size_t)b * d.Lout + i0 + ii) * q + j] = s;
      }
    }
    __syncthreads();
  }
  // partial row: [O*q grad post_conv.weight][O grad post_conv.bias]
  T* prow = a.part + (size_t)blockIdx.x * a.PA;
  for (int idx = tid; idx < d.O * q; idx += kT) {
    const int o = idx / q, j = idx - o * q;
    prow[idx] = gwacc[(size_t)o * (QP + 1) + j];
  }
  for (int o = tid; o < d.O; o += kT) prow[(size_t)d.O * q + o] = gwacc[(size_t)o * (QP + 1) + QP];
}

// ---------------------------------------------------------------- backward: grad_x, gather form
template <typename T, int QP>
__global__ void __launch_bounds__(kT) gen_preconv_bwd_gx_kernel(const GArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ConvDims d = a.d;
  const int CK = d.C * d.K, q = a.q;
  T* wt = reinterpret_cast<T*>(smem_raw);  // [CK][QP]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CK * QP; idx += kT) {
    const int f = idx / QP, j = idx - f * QP;
    wt[idx] = j < q ? a.w_pre[(size_t)j * CK + f] : T(0);
  }
  __syncthreads();
  const int lt = (d.L + kT - 1) / kT;  // position tiles per row
  const long long total = (long long)d.B * d.C * lt;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ltile = (int)(t % lt);
    const long long bc = t / lt;
    const int c = (int)(bc % d.C), b = (int)(bc / d.C);
    const int l = ltile * kT + tid;
    if (l >= d.L) continue;
    T acc = T(0);
    for (int k = 0; k < d.K; ++k) {
      const int num = l + d.P - k;  // window i touches column l with tap k iff i*S - P + k == l
      if (num < 0 || num % d.S) continue;
      const int i = num / d.S;
      if (i >= d.Lout) continue;
      const T* gp = a.gpre_c + ((size_t)b * d.Lout + i) * q;
      const T* wr = wt + (size_t)(c * d.K + k) * QP;
#pragma unroll
      for (int j = 0; j < QP; ++j)
        if (j < q) acc = fma(gp[j], wr[j], acc);
    }
    a.gx[((size_t)b * d.C + c) * d.L + l] = acc;
  }
}

// ---------------------------------------------------------------- backward: partials of grad pre_conv.{weight,bias}
// grid (C, NS): CTA (c, s) reduces window slice s for channel c; partial row s: [q*CK grad weight (j-major)][q grad bias]
template <typename T, int QP>
__global__ void __launch_bounds__(kT) gen_preconv_bwd_gw_kernel(const GArgs<T> a) {
  __shared__ T red[4][QP + 1];
  const ConvDims d = a.d;
  const int CK = d.C * d.K, q = a.q;
  const int c = blockIdx.x, s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long W = (long long)d.B * d.Lout;
  const long long per = (W + a.NS - 1) / a.NS;
  const long long w0 = (long long)s * per, w1 = (w0 + per < W) ? w0 + per : W;
  T* prow = a.part + (size_t)s * a.PA;
  for (int k = 0; k < d.K; ++k) {
    T acc[QP + 1];
#pragma unroll
    for (int j = 0; j <= QP; ++j) acc[j] = T(0);
    for (long long w = w0 + tid; w < w1; w += kT) {
      const int b = (int)(w / d.Lout), i = (int)(w - (long long)b * d.Lout);
      const int l = i * d.S - d.P + k;
      const T xv = (l >= 0 && l < d.L) ? __ldg(a.x + ((size_t)b * d.C + c) * d.L + l) : T(0);
      const T* gp = a.gpre_c + (size_t)w * q;
#pragma unroll
      for (int j = 0; j < QP; ++j) {
        const T g = j < q ? gp[j] : T(0);
        acc[j] = fma(g, xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < QP; ++j) acc[j] = warp_sum<T>(acc[j]);
    if (lane == 0)
#pragma unroll
      for (int j = 0; j < QP; ++j) red[warp][j] = acc[j];
    __syncthreads();
    if (tid < q) prow[(size_t)tid * CK + c * d.K + k] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    __syncthreads();
  }
  if (c == 0) {  // grad pre_conv.bias = sum of gpre over the slice
    T acc[QP];
#pragma unroll
    for (int j = 0; j < QP; ++j) acc[j] = T(0);
    for (long long w = w0 + tid; w < w1; w += kT)
#pragma unroll
      for (int j = 0; j < QP; ++j)
        if (j < q) acc[j] += a.gpre_c[(size_t)w * q + j];
#pragma unroll
    for (int j = 0; j < QP; ++j) acc[j] = warp_sum<T>(acc[j]);
    if (lane == 0)
#pragma unroll
      for (int j = 0; j < QP; ++j) red[warp][j] = acc[j];
    __syncthreads();
    if (tid < q) prow[(size_t)q * CK + tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
  }
}

// ---------------------------------------------------------------- backward: pre_conv^T in ONE pass over x (kernel_size 3)
// grad_x and the partials of grad pre_conv.{weight,bias} from a single read of x -- the general path's counterpart of
// fast_bwd_pre_kernel.  The two kernels above cost 229 us per launch at the stem's conv2 shape (every (b, c, l) thread of the gx
// kernel fetched K * q gpre values, every CTA of the gw kernel walked all 24 000 windows at 2.6 CTAs per SM).  Here:
//   * lanes = 32 consecutive CHANNELS: a lane keeps its channel's pre_conv weights w[q][3] and gradient accumulators gw[q][3] in
//     registers for the CTA lifetime -> one partial row per CTA column, no atomics;
//   * each of the 4 warps owns 32 consecutive positions of the CTA's 128-position tile: the 32 x 32 block of x is read with
//     coalesced 128-byte rows, turned through a pitch-33 shared-memory block (warp-private: __syncwarp only), grad_x is written
//     back over it in place and leaves by coalesced rows again;
//   * the <= 3 windows that touch a position have a warp-uniform index: the <= 36 gpre rows a warp's 32 positions can touch are
//     staged once per tile in a warp-private shared-memory block (coalesced read, rows of windows outside [0, L_out) zeroed, so the
//     inner loop carries no bounds checks) and read back as broadcast vector loads -- uniform GLOBAL loads, 18 per position at
//     q = 6, made the first version of this kernel load/store-unit bound (94 us per launch);
//   * per lane and position 2 q K FMAs against ~12 other instructions.
// grid: (position CTAs, channel chunks); partial row bx: [q * C * 3 grad weight, j-major][q grad bias] (as gen_reduce_rows expects).
// SS: compile-time stride (1, 2) or 0 = any (run-time division per tap).
template <typename T, int QP, int SS>
__global__ void __launch_bounds__(kT) gen_preconv_bwd_k3_kernel(const GArgs<T> a) {
  constexpr int RW = QP * 3 + 1, GR = 36;
  constexpr int kStream = 4 * 32 * 33 + 4 * GR * QP, kRed = 4 * 32 * RW;
  // while streaming: the warps' x blocks [4][32][33] and gpre rows [4][GR][QP]; afterwards the reduction buffer [4][32][RW]
  __shared__ __align__(16) T sbuf[kStream > kRed ? kStream : kRed];
  T(*xt)[32][33] = reinterpret_cast<T(*)[32][33]>(sbuf);
  T(*gsa)[GR][QP] = reinterpret_cast<T(*)[GR][QP]>(sbuf + 4 * 32 * 33);
  T(*red)[32][RW] = reinterpret_cast<T(*)[32][RW]>(sbuf);
  const ConvDims d = a.d;
  const int q = a.q, CK = d.C * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * 32, c = c0 + lane;
  T w[QP][3], gw[QP][3];
#pragma unroll
  for (int j = 0; j < QP; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[j][k] = (j < q && c < d.C) ? a.w_pre[(size_t)j * CK + c * 3 + k] : T(0);
      gw[j][k] = T(0);
    }
  T gb[QP];  // grad pre_conv.bias: lane j of warp 0 in chunk 0 sums gpre[:, j] over the windows this tile owns
#pragma unroll
  for (int j = 0; j < QP; ++j) gb[j] = T(0);
  const int lt = (d.L + 127) / 128;  // position tiles per utterance
  const int total = d.B * lt;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int b = t / lt, l0 = (t - b * lt) * 128, lw = l0 + warp * 32;
    const T* __restrict__ xb = a.x + ((size_t)b * d.C + c0) * d.L;
    T(*blk)[33] = xt[warp];
    T(*gs)[QP] = gsa[warp];
    if (lw < d.L) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) blk[r][lane] = (c0 + r < d.C && lw + lane < d.L) ? __ldg(xb + (size_t)r * d.L + lw + lane) : T(0);
      // gpre rows of the windows i_lo .. i_lo + GR - 1 (floor division: the first tile reaches windows < 0, which are zero rows)
      const int nlo = lw + d.P - 2;
      const int i_lo = SS == 1 ? nlo : SS == 2 ? (nlo >> 1) : (nlo >= 0 ? nlo / d.S : -((-nlo + d.S - 1) / d.S));
      {
        const T* __restrict__ gpb = a.gpre_c + (size_t)b * d.Lout * q;
        for (int e = lane; e < GR * QP; e += 32) {
          const int row = e / QP, j = e - row * QP, i = i_lo + row;
          gs[row][j] = (j < q && i >= 0 && i < d.Lout) ? __ldg(gpb + (size_t)i * q + j) : T(0);
        }
      }
      __syncwarp();
#pragma unroll 2
      for (int p = 0; p < 32; ++p) {
        const T xv = blk[lane][p];
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int num = lw + p + d.P - k;  // window i touches column l with tap k iff i * S - P + k == l
          const int i = SS == 1 ? num : SS == 2 ? (num >> 1) : (num >= 0 ? num / d.S : -((-num + d.S - 1) / d.S));
          const bool hit = SS == 1 ? true : SS == 2 ? (num & 1) == 0 : i * d.S == num;
          if (hit) {  // warp-uniform; i - i_lo lies in [0, GR)
            const T* gp = gs[i - i_lo];
#pragma unroll
            for (int j = 0; j < QP; ++j) {
              const T g = gp[j];
              gw[j][k] = fma(g, xv, gw[j][k]);
              acc = fma(g, w[j][k], acc);
            }
          }
        }
        blk[lane][p] = acc;
      }
      __syncwarp();
      if (a.gx) {
        T* __restrict__ gxb = a.gx + ((size_t)b * d.C + c0) * d.L;
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
          if (c0 + r < d.C && lw + lane < d.L) gxb[(size_t)r * d.L + lw + lane] = blk[r][lane];
      }
      __syncwarp();
    }
    if (blockIdx.y == 0) {
      // windows owned by this tile: tap-0 column i * S - P in [l0, l0 + 128) (the first tile also takes the negative columns)
      const int lo = l0 == 0 ? 0 : (l0 + d.P + d.S - 1) / d.S;
      int hi = (l0 + 128 + d.P + d.S - 1) / d.S;
      hi = hi < d.Lout ? hi : d.Lout;
      if (l0 + 128 >= d.L) hi = d.Lout;
      for (int i = lo + tid; i < hi; i += kT) {
        const T* gp = a.gpre_c + ((size_t)b * d.Lout + i) * q;
#pragma unroll
        for (int j = 0; j < QP; ++j)
          if (j < q) gb[j] += gp[j];
      }
    }
  }
  // ---- cross-warp reduction (fixed order) -> this CTA's columns of partial row blockIdx.x
  __syncthreads();  // every warp is done with its x block: the buffer becomes `red`
#pragma unroll
  for (int j = 0; j < QP; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) red[warp][lane][j * 3 + k] = gw[j][k];
  __syncthreads();
  T* prow = a.part + (size_t)blockIdx.x * a.PA;
  for (int e = tid; e < 32 * q * 3; e += kT) {
    const int ln = e / (q * 3), jk = e - ln * (q * 3);
    const int j = jk / 3, k = jk - j * 3;
    if (c0 + ln < d.C) prow[(size_t)j * CK + (c0 + ln) * 3 + k] = red[0][ln][jk] + red[1][ln][jk] + red[2][ln][jk] + red[3][ln][jk];
  }
  if (blockIdx.y == 0) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < QP; ++j) {
      const T v = warp_sum<T>(gb[j]);
      if (lane == 0) red[warp][0][j] = v;
    }
    __syncthreads();
    if (tid < q) prow[(size_t)q * CK + tid] = red[0][0][tid] + red[1][0][tid] + red[2][0][tid] + red[3][0][tid];
  }
}

// out0[e] (e < n0) / out1[e - n0] = sum_g part[g][e], fixed order, fp64 accumulation.
// A CTA owns 32 columns; its 8 warps take the rows g = warp, warp + 8, ... (128-byte coalesced reads, chains 8 x shorter than one
// thread per column walking all G rows: 592 rows took 33 us per launch that way) and meet in shared memory in warp order.
template <typename T>
__global__ void __launch_bounds__(256) gen_reduce_rows_kernel(const T* __restrict__ part, int G, int P, T* __restrict__ out0, int n0,
                                                              T* __restrict__ out1, int n1) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (e < n0 + n1) {
    // 8 independent loads in flight per thread (the partial rows come from HBM / L2: a chain of dependent loads is pure latency)
    int g = warp;
    for (; g + 56 < G; g += 64) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(size_t)(g + 8 * u) * P + e];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; g < G; g += 8) s += (double)part[(size_t)g * P + e];
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && e < n0 + n1) {
    double t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lane];
    if (e < n0) out0[e] = (T)t;
    else out1[e - n0] = (T)t;
  }
}

// =============================================================================================== host
struct GenPlan {
  int tiles_per_utt, num_tiles, gridT, gridA, PA, NS, PB, gridX, gridP;
  size_t off_gout, off_gpre, off_pa, off_pb, off_circ, total;
};

static GenPlan make_plan(const ConvDims& d, int elem) {
  GenPlan p{};
  const int sms = num_sms();
  p.tiles_per_utt = (d.Lout + kTW - 1) / kTW;
  p.num_tiles = d.B * p.tiles_per_utt;
  p.gridT = p.num_tiles < sms * 8 ? p.num_tiles : sms * 8;
  p.gridA = p.num_tiles < sms * 4 ? p.num_tiles : sms * 4;  // 2 CTAs per SM left the gy pass latency-bound (116 us per launch)
  p.PA = (int)align_up((size_t)d.O * (d.Q + 1), 32);
  p.NS = 16 * sms / d.C;  // C * NS CTAs of 128 threads: ~16 per SM (4 * sms / C left conv2, C = 384, with ONE slice: 2.6 CTAs per SM)
  p.NS = p.NS < 1 ? 1 : p.NS > 64 ? 64 : p.NS;
  const long long W = (long long)d.B * d.Lout;
  if (p.NS > W) p.NS = (int)W;
  {
    // fused pre_conv^T kernel (kernel_size 3): gridP position CTAs per 32-channel chunk, ~8 CTAs per SM in total
    const int chunks = (d.C + 31) / 32, tiles = d.B * ((d.L + 127) / 128);
    int gp = 8 * sms / chunks;
    gp = gp < 1 ? 1 : gp;
    p.gridP = gp < tiles ? gp : tiles;
    if (d.K == 3 && p.gridP > p.NS) p.NS = p.gridP;  // partial rows of either form fit
  }
  p.PB = (int)align_up((size_t)d.Q * d.C * d.K + d.Q, 32);
  {
    const long long t = (long long)d.B * d.C * ((d.L + kT - 1) / kT);
    p.gridX = (int)(t < (long long)sms * 16 ? t : (long long)sms * 16);
  }
  size_t o = 0;
  p.off_gout = o; o = align_up(o + (size_t)W * d.Q * elem, 256);
  p.off_gpre = o; o = align_up(o + (size_t)W * d.Q * elem, 256);
  p.off_pa = o;   o = align_up(o + (size_t)p.gridA * p.PA * elem, 256);
  p.off_pb = o;   o = align_up(o + (size_t)p.NS * p.PB * elem, 256);
  p.off_circ = o; o = align_up(o + wc::wcirc_workspace_bytes(W, d.Q, d.Lq, elem), 256);
  p.total = o;
  return p;
}

size_t general_workspace_bytes(const ConvDims& d, int elem) { return make_plan(d, elem).total; }

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
  QW_CHECK_ARG(bytes <= 227 * 1024, -2, "general QuantumConv1d path needs %zu bytes of shared memory (C*K or O too large)", bytes);
  if (bytes > 48 * 1024) QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <typename T, int QP>
static int forward_qp(GArgs<T> a, const GenPlan& p, const T* qwts, cudaStream_t st) {
  const ConvDims& d = a.d;
  const size_t W = (size_t)d.B * d.Lout;
  {
    const size_t smem = ((size_t)d.C * d.K * QP + 4 * kTW * QP) * sizeof(T);
    auto k = gen_preconv_fwd_kernel<T, QP>;
    if (int e = set_smem(k, smem)) return e;
    KernelTimer kt(kKFwd, st);
    k<<<p.gridT, kT, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  if (int e = wc::wcirc_forward<T>(a.pre_w, qwts, a.pre_w + W * d.Q, (long long)W, d.Q, d.Lq, d.emb, st)) return e;
  {
    const size_t smem = ((size_t)d.O * QP + d.O) * sizeof(T);
    auto k = gen_postconv_fwd_kernel<T, QP>;
    if (int e = set_smem(k, smem)) return e;
    KernelTimer kt(kKFwd, st);
    k<<<p.gridT, kT, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int general_forward(const T* x, const T* w_pre, const T* b_pre, const T* qwts, const T* w_post, const T* b_post, T* y, T* pre_save,
                    const ConvDims& d, cudaStream_t st) {
  QW_CHECK_ARG(pre_save != nullptr, -2,
               "the general QuantumConv1d path (n_qubits > 4 or angle embedding) needs the pre_save buffer (2, B*L_out, q) as scratch");
  const GenPlan p = make_plan(d, (int)sizeof(T));
  const size_t W = (size_t)d.B * d.Lout;
  GArgs<T> a{};
  a.x = x; a.w_pre = w_pre; a.b_pre = b_pre; a.w_post = w_post; a.b_post = b_post;
  a.qout = pre_save + W * d.Q;
  a.y = y; a.pre_w = pre_save;
  a.d = d; a.tiles_per_utt = p.tiles_per_utt; a.num_tiles = p.num_tiles; a.q = d.Q;
  return d.Q <= 4 ? forward_qp<T, 4>(a, p, qwts, st) : d.Q <= 6 ? forward_qp<T, 6>(a, p, qwts, st) : d.Q <= 8 ? forward_qp<T, 8>(a, p, qwts, st) : forward_qp<T, 12>(a, p, qwts, st);
}

template <typename T, int QP>
static int backward_qp(GArgs<T> a, const GenPlan& p, const T* qwts, T* gqw, T* gw_pre, T* gb_pre, T* gw_post, T* gb_post,
                       unsigned char* ws, cudaStream_t st) {
  const ConvDims& d = a.d;
  const size_t W = (size_t)d.B * d.Lout;
  T* partA = reinterpret_cast<T*>(ws + p.off_pa);
  T* partB = reinterpret_cast<T*>(ws + p.off_pb);
  {
    const size_t smem = ((size_t)d.O * QP + (size_t)d.O * (QP + 1) + kOC * 33 + kTW * QP + 4 * kTW * QP) * sizeof(T);
    auto k = gen_postconv_bwd_kernel<T, QP>;
    if (int e = set_smem(k, smem)) return e;
    a.part = partA; a.PA = p.PA;
    KernelTimer kt(kKBwdPost, st);
    k<<<p.gridA, kT, smem, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  if (int e = wc::wcirc_backward<T>(a.pre, qwts, a.gout_c, const_cast<T*>(a.gpre_c), gqw, ws + p.off_circ, (long long)W, d.Q, d.Lq,
                                    d.emb, st))
    return e;
  if (d.K == 3) {
    a.part = partB; a.PA = p.PB;
    KernelTimer kt(kKBwdPre, st);
    const dim3 grid(p.gridP, (d.C + 31) / 32);
    if (d.S == 1) gen_preconv_bwd_k3_kernel<T, QP, 1><<<grid, kT, 0, st>>>(a);
    else if (d.S == 2) gen_preconv_bwd_k3_kernel<T, QP, 2><<<grid, kT, 0, st>>>(a);
    else gen_preconv_bwd_k3_kernel<T, QP, 0><<<grid, kT, 0, st>>>(a);
    QW_CUDA_OK(cudaGetLastError());
  } else if (a.gx) {
    const size_t smem = (size_t)d.C * d.K * QP * sizeof(T);
    auto k = gen_preconv_bwd_gx_kernel<T, QP>;
    if (int e = set_smem(k, smem)) return e;
    KernelTimer kt(kKBwdPre, st);
    k<<<p.gridX, kT, smem, st>>>(a);
    QW_CUDA_OK(cudaGetLastError());
  }
  if (d.K != 3) {
    a.part = partB; a.PA = p.PB; a.NS = p.NS;
    KernelTimer kt(kKBwdPre, st);
    gen_preconv_bwd_gw_kernel<T, QP><<<dim3(d.C, p.NS), kT, 0, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  {
    KernelTimer kt(kKBwdFinalize, st);
    const int nA = d.O * d.Q + d.O;
    gen_reduce_rows_kernel<T><<<(nA + 31) / 32, 256, 0, st>>>(partA, p.gridA, p.PA, gw_post, d.O * d.Q, gb_post, d.O);
  }
  QW_CUDA_OK(cudaGetLastError());
  {
    KernelTimer kt(kKBwdFinalize, st);
    const int nB = d.Q * d.C * d.K + d.Q;
    gen_reduce_rows_kernel<T><<<(nB + 31) / 32, 256, 0, st>>>(partB, d.K == 3 ? p.gridP : p.NS, p.PB, gw_pre, d.Q * d.C * d.K, gb_pre, d.Q);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int general_backward(const T* gy, const T* x, const T* pre_save, const T* w_pre, const T* qwts, const T* w_post, T* gx, T* gw_pre,
                     T* gb_pre, T* gqw, T* gw_post, T* gb_post, unsigned char* ws, size_t ws_bytes, const ConvDims& d, cudaStream_t st) {
  const GenPlan p = make_plan(d, (int)sizeof(T));
  QW_CHECK_ARG(ws_bytes >= p.total, -3, "workspace too small: %zu < %zu", ws_bytes, p.total);
  const size_t W = (size_t)d.B * d.Lout;
  GArgs<T> a{};
  a.x = x; a.w_pre = w_pre; a.w_post = w_post; a.gy = gy;
  a.pre = pre_save; a.qout = pre_save + W * d.Q;
  a.gout = reinterpret_cast<T*>(ws + p.off_gout);
  a.gout_c = a.gout;
  a.gpre_c = reinterpret_cast<T*>(ws + p.off_gpre);
  a.gx = gx;
  a.d = d; a.tiles_per_utt = p.tiles_per_utt; a.num_tiles = p.num_tiles; a.q = d.Q;
  return d.Q <= 4 ? backward_qp<T, 4>(a, p, qwts, gqw, gw_pre, gb_pre, gw_post, gb_post, ws, st)
       : d.Q <= 6 ? backward_qp<T, 6>(a, p, qwts, gqw, gw_pre, gb_pre, gw_post, gb_post, ws, st)
       : d.Q <= 8 ? backward_qp<T, 8>(a, p, qwts, gqw, gw_pre, gb_pre, gw_post, gb_post, ws, st)
                  : backward_qp<T, 12>(a, p, qwts, gqw, gw_pre, gb_pre, gw_post, gb_post, ws, st);
}

template int general_forward<float>(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*,
                                    const ConvDims&, cudaStream_t);
template int general_forward<double>(const double*, const double*, const double*, const double*, const double*, const double*, double*,
                                     double*, const ConvDims&, cudaStream_t);
template int general_backward<float>(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*,
                                     float*, float*, float*, float*, unsigned char*, size_t, const ConvDims&, cudaStream_t);
template int general_backward<double>(const double*, const double*, const double*, const double*, const double*, const double*, double*,
                                      double*, double*, double*, double*, double*, unsigned char*, size_t, const ConvDims&, cudaStream_t);

}  // namespace gen
}  // namespace qw
