// Thread-group cooperative batched statevector simulator + adjoint differentiation, n_qubits 1..12, any number of
// layers, amplitude or angle embedding, fp32 / fp64.  The general form of the QNode of
// /root/reference/quantum_whisper.py:64-85 (BASELINE.json config 4: q in {4,6,8,10,12} x n_layers in {1,2,4}).
//
// Layout: the 2^Q complex amplitudes of ONE window are spread over a group of TG = 2^NT threads, NR = 2^R per
// thread in registers; amplitude index k = reg * TG + tsub (thread bits are the LOW bits, so direct and Gray-permuted
// shared-memory accesses are both bank-conflict free).
//     Q <= 4 : R = Q, one thread per window (no communication at all)
//     Q 5..9 : R = 4, 2..32 lanes per window (several windows per warp)
//     Q = 10 : R = 5, one warp per window
//     Q 11,12: R = 5, 2 / 4 warps per window (the CTA is the group)
// A 1-qubit gate on a register bit is local FMA work; on a lane bit it is a __shfl_xor butterfly; on a warp bit
// (Q >= 11) it goes through the group's shared-memory exchange buffer.  The CNOT chain is the basis permutation
// new[j] = old[j ^ (j >> 1)] (inverse Gray code on the index): registers when NT == 0, else one pass through the
// exchange buffer.
//
// Backward = adjoint method: recompute the final state, lambda = (sum_i gout_i Z_i) psi, then walk the gates in
// reverse: N_ac += conj(lambda_a) psi_c (both AFTER the gate), psi <- G^dag psi, lambda <- G^dag lambda.  N is reduced
// over the warp with a 9-shuffle transposing butterfly and accumulated per (layer, wire) in shared memory; the finalize
// kernel turns it into M = N conj(G) and applies the chain rule to (phi, theta, omega).  Parity checked against
// oracle/qconv_oracle.py (the same math was first prototyped in numpy against it).
#pragma once
#include "qw_circuit.cuh"
#include "qw_common.cuh"

namespace qw {
namespace wc {

template <int Q>
struct Cfg {
  static constexpr int R = Q <= 4 ? Q : (Q <= 9 ? 4 : 5);
  static constexpr int NT = Q - R;
  static constexpr int TG = 1 << NT;
  static constexpr int NR = 1 << R;
  static constexpr int N = 1 << Q;
  static constexpr int THREADS = TG > 32 ? TG : 128;
  static constexpr int GPC = THREADS / TG;          // windows (groups) per CTA iteration
  static constexpr int WARPS = THREADS / 32;
  static constexpr int GS = NT == 0 ? 0 : 2 * N + TG;  // exchange-buffer elements per group (re plane, im plane, skew)
  // (measured: capping the adjoint kernel's registers through __launch_bounds__ min-CTAs -- 255 -> 168 at R = 5, 168 -> 128
  // at R = 4 -- changes nothing, q = 8: -7 %, q = 10: +2 %; the kernels are not occupancy-bound.  Not applied.)
};

template <typename T>
struct WArgs {
  const T *pre, *qw, *gout;
  T *out, *gpre, *part;
  long long W;
  int Lq, emb, PA;
};

__host__ __device__ constexpr int gray(int j) { return j ^ (j >> 1); }
__host__ __device__ constexpr int inv_gray(int j) {
  int x = j;
  x ^= x >> 1;
  x ^= x >> 2;
  x ^= x >> 4;
  x ^= x >> 8;
  return x;
}

template <typename T>
struct Ctx {
  int tsub;   // thread index inside the group
  int lane;
  T* xr;      // group's exchange buffer: re plane [N]
  T* xi;      //                          im plane [N]
};

template <int Q>
__device__ __forceinline__ void group_sync() {
  if constexpr (Cfg<Q>::TG > 32) __syncthreads();
  else if constexpr (Cfg<Q>::TG > 1) __syncwarp();
}

// Access to the partner thread's copy of (vr, vi)[r] across thread bit P, ELEMENT BY ELEMENT (no NR-sized staging buffer in
// registers: at R = 5 that buffer alone was 64 registers and pinned the adjoint kernels at 255 registers / 2 warps per
// scheduler).  Lane bits: a pair of __shfl_xor per element.  Warp bits (Q >= 11): the whole register file of the group goes
// through the exchange buffer once (partner_begin), elements are then read from the partner's slot, partner_end closes it.
template <typename T, int Q, int P>
__device__ __forceinline__ void partner_begin(const T (&vr)[Cfg<Q>::NR], const T (&vi)[Cfg<Q>::NR], const Ctx<T>& c) {
  using C = Cfg<Q>;
  if constexpr (P >= 5) {
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      c.xr[r * C::TG + c.tsub] = vr[r];
      c.xi[r * C::TG + c.tsub] = vi[r];
    }
    __syncthreads();
  }
}
template <typename T, int Q, int P>
__device__ __forceinline__ void partner_get(T mr, T mi, int r, const Ctx<T>& c, T& qr, T& qi) {
  using C = Cfg<Q>;
  if constexpr (P < 5) {
    qr = __shfl_xor_sync(0xffffffffu, mr, 1 << P);
    qi = __shfl_xor_sync(0xffffffffu, mi, 1 << P);
  } else {
    qr = c.xr[r * C::TG + (c.tsub ^ (1 << P))];
    qi = c.xi[r * C::TG + (c.tsub ^ (1 << P))];
  }
}
template <int Q, int P>
__device__ __forceinline__ void partner_end() {
  if constexpr (P >= 5) __syncthreads();
}

// g: 8 reals, row-major complex 2x2 (00r 00i 01r 01i 10r 10i 11r 11i) acting on bit position P
template <typename T, int Q, int P>
__device__ __forceinline__ void apply_gate_w(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const T* __restrict__ g, const Ctx<T>& c) {
  using C = Cfg<Q>;
  if constexpr (P >= C::NT) {
    constexpr int RB = P - C::NT;
    const T g00r = g[0], g00i = g[1], g01r = g[2], g01i = g[3], g10r = g[4], g10i = g[5], g11r = g[6], g11i = g[7];
#pragma unroll
    for (int r0 = 0; r0 < C::NR; ++r0) {
      if ((r0 >> RB) & 1) continue;
      const int r1 = r0 | (1 << RB);
      const T a0r = re[r0], a0i = im[r0], a1r = re[r1], a1i = im[r1];
      re[r0] = g00r * a0r - g00i * a0i + g01r * a1r - g01i * a1i;
      im[r0] = g00r * a0i + g00i * a0r + g01r * a1i + g01i * a1r;
      re[r1] = g10r * a0r - g10i * a0i + g11r * a1r - g11i * a1i;
      im[r1] = g10r * a0i + g10i * a0r + g11r * a1i + g11i * a1r;
    }
  } else {
    const bool hi = (c.tsub >> P) & 1;
    const T ar = hi ? g[6] : g[0], ai = hi ? g[7] : g[1];  // coefficient of my own amplitude
    const T br = hi ? g[4] : g[2], bi = hi ? g[5] : g[3];  // coefficient of the partner's
    partner_begin<T, Q, P>(re, im, c);
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const T mr = re[r], mi = im[r];
      T qr, qi;
      partner_get<T, Q, P>(mr, mi, r, c, qr, qi);
      re[r] = ar * mr - ai * mi + br * qr - bi * qi;
      im[r] = ar * mi + ai * mr + br * qi + bi * qr;
    }
    partner_end<Q, P>();
  }
}

// lane-bit gate with a RUN-TIME bit position (P < min(NT, 5)): the code of the five lane-bit gates differs only in the shuffle
// mask, so one body in a rolled loop replaces five unrolled copies.  ncu showed these kernels starved for instructions
// (stall "no_inst" 42 % forward / 64 % backward at q = 10: ~50 / ~200 KB of straight-line SASS per layer body).
template <typename T, int Q>
__device__ __forceinline__ void apply_gate_lane_rt(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const T* __restrict__ g, int P,
                                                   const Ctx<T>& c) {
  using C = Cfg<Q>;
  const bool hi = (c.tsub >> P) & 1;
  const int mask = 1 << P;
  const T ar = hi ? g[6] : g[0], ai = hi ? g[7] : g[1];
  const T br = hi ? g[4] : g[2], bi = hi ? g[5] : g[3];
#pragma unroll
  for (int r = 0; r < C::NR; ++r) {
    const T mr = re[r], mi = im[r];
    const T qr = __shfl_xor_sync(0xffffffffu, mr, mask), qi = __shfl_xor_sync(0xffffffffu, mi, mask);
    re[r] = ar * mr - ai * mi + br * qr - bi * qi;
    im[r] = ar * mi + ai * mr + br * qi + bi * qr;
  }
}

// new[j] = old[rol(j)]: the amplitude at register index (b_{R-1} ... b_1 b_0) moves to (b_0 b_{R-1} ... b_1), i.e. register bit
// i+1 becomes bit i.  R rotations are the identity, so a rolled loop "gate on register bit 0; rotate" visits every register
// bit with ONE copy of the gate code (2 NR moves per gate buy a 5x smaller loop body that fits the instruction cache).
template <typename T, int Q>
__device__ __forceinline__ void rotate_regs(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR]) {
  using C = Cfg<Q>;
  if constexpr (C::R >= 2) {
    T nr[C::NR], ni[C::NR];
#pragma unroll
    for (int j = 0; j < C::NR; ++j) {
      const int src = ((j << 1) | (j >> (C::R - 1))) & (C::NR - 1);
      nr[j] = re[src];
      ni[j] = im[src];
    }
#pragma unroll
    for (int j = 0; j < C::NR; ++j) {
      re[j] = nr[j];
      im[j] = ni[j];
    }
  }
}

template <typename T, int Q, int P>
struct FwdGatesFrom {  // compile-time bit positions P .. Q-1 (warp bits and register bits)
  static __device__ __forceinline__ void run(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const T* __restrict__ gl, const Ctx<T>& c) {
    if constexpr (P < Q) {
      apply_gate_w<T, Q, P>(re, im, gl + (Q - 1 - P) * kGateStride, c);
      FwdGatesFrom<T, Q, P + 1>::run(re, im, gl, c);
    }
  }
};

template <typename T, int Q, int P>
struct FwdGates {  // one Rot layer; P is kept for source compatibility (always instantiated with 0)
  static __device__ __forceinline__ void run(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const T* __restrict__ gl, const Ctx<T>& c) {
    using C = Cfg<Q>;
    constexpr int NTL = C::NT < 5 ? C::NT : 5;
#pragma unroll 1
    for (int p = 0; p < NTL; ++p) apply_gate_lane_rt<T, Q>(re, im, gl + (Q - 1 - p) * kGateStride, p, c);
    if constexpr (C::NT > NTL) {  // warp bits (Q >= 11): through the exchange buffer, unrolled (1 or 2 gates)
      apply_gate_w<T, Q, 5>(re, im, gl + (Q - 1 - 5) * kGateStride, c);
      if constexpr (C::NT > 6) apply_gate_w<T, Q, 6>(re, im, gl + (Q - 1 - 6) * kGateStride, c);
    }
    if constexpr (C::R >= 3) {    // register bits: one copy of the bit-0 gate, rolled over the R bits by register rotation
#pragma unroll 1
      for (int i = 0; i < C::R; ++i) {
        apply_gate_w<T, Q, C::NT>(re, im, gl + (Q - 1 - C::NT - i) * kGateStride, c);
        rotate_regs<T, Q>(re, im);
      }
    } else {
      FwdGatesFrom<T, Q, C::NT>::run(re, im, gl, c);
    }
  }
};

// CNOT(0,1) ... CNOT(Q-2,Q-1): new[j] = old[gray(j)];  inverse: new[j] = old[inv_gray(j)]
template <typename T, int Q, bool INV>
__device__ __forceinline__ void cnot_chain_w(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const Ctx<T>& c) {
  using C = Cfg<Q>;
  if constexpr (Q == 1) {
    return;
  } else if constexpr (C::NT == 0) {
    T nr[C::NR], ni[C::NR];
#pragma unroll
    for (int j = 0; j < C::NR; ++j) {
      const int s = INV ? inv_gray(j) : gray(j);
      nr[j] = re[s];
      ni[j] = im[s];
    }
#pragma unroll
    for (int j = 0; j < C::NR; ++j) {
      re[j] = nr[j];
      im[j] = ni[j];
    }
  } else {
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      c.xr[r * C::TG + c.tsub] = re[r];
      c.xi[r * C::TG + c.tsub] = im[r];
    }
    group_sync<Q>();
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const int j = r * C::TG + c.tsub;
      const int s = INV ? inv_gray(j) : gray(j);
      re[r] = c.xr[s];
      im[r] = c.xi[s];
    }
    group_sync<Q>();
  }
}

// sum over the group's threads, result in every thread.  V values at once (one smem round for multi-warp groups).
template <typename T, int Q, int V>
__device__ __forceinline__ void group_sum(T (&v)[V], const Ctx<T>& c) {
  using C = Cfg<Q>;
  if constexpr (C::TG > 1) {
#pragma unroll
    for (int m = 1; m < (C::TG < 32 ? C::TG : 32); m <<= 1)
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], m);
    if constexpr (C::TG > 32) {
      const int warp = c.tsub >> 5;
      if (c.lane == 0)
#pragma unroll
        for (int e = 0; e < V; ++e) c.xr[warp * V + e] = v[e];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < V; ++e) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < C::TG / 32; ++w) s += c.xr[w * V + e];
        v[e] = s;
      }
      __syncthreads();
    }
  }
}

// gate matrix of the angle embedding on one wire: RZ(x) RY(x)
template <typename T>
__device__ __forceinline__ void angle_gate(T x, T (&g)[8]) {
  T s, cc;
  sincos(T(0.5) * x, &s, &cc);
  // e^{-ix/2} = cc - i s ; e^{+ix/2} = cc + i s
  g[0] = cc * cc;  g[1] = -s * cc;   // 00: e^{-ix/2} c
  g[2] = -cc * s;  g[3] = s * s;     // 01: -e^{-ix/2} s
  g[4] = cc * s;   g[5] = s * s;     // 10: e^{+ix/2} s
  g[6] = cc * cc;  g[7] = s * cc;    // 11: e^{+ix/2} c
}

template <typename T, int Q, int P>
struct AngleEmbed {
  static __device__ __forceinline__ void run(T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const T (&pre)[Q], const Ctx<T>& c) {
    if constexpr (P < Q) {
      T g[8];
      angle_gate<T>(pre[Q - 1 - P], g);
      apply_gate_w<T, Q, P>(re, im, g, c);
      AngleEmbed<T, Q, P + 1>::run(re, im, pre, c);
    }
  }
};

// initial state; returns 1/||pre|| (amplitude mode) or 0
template <typename T, int Q>
__device__ __forceinline__ T embed_w(const T (&pre)[Q], int emb, T (&re)[Cfg<Q>::NR], T (&im)[Cfg<Q>::NR], const Ctx<T>& c) {
  using C = Cfg<Q>;
  if (emb == kEmbAmplitude) {
    T ss = T(0);
#pragma unroll
    for (int j = 0; j < Q; ++j) ss = fma(pre[j], pre[j], ss);
    const T inv = T(1) / sqrt(ss);  // ||v|| = 0 -> NaN like the reference (quantum_whisper.py:74)
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      T v = T(0);
#pragma unroll
      for (int j = 0; j < Q; ++j)
        if (j / C::TG == r) v = (c.tsub == j % C::TG) ? pre[j] * inv : v;
      re[r] = v;
      im[r] = T(0);
    }
    return inv;
  }
#pragma unroll
  for (int r = 0; r < C::NR; ++r) {
    re[r] = (r == 0 && c.tsub == 0) ? T(1) : T(0);
    im[r] = T(0);
  }
  AngleEmbed<T, Q, 0>::run(re, im, pre, c);
  return T(0);
}

template <typename T, int Q>
__device__ __forceinline__ void forward_w(const T (&pre)[Q], int emb, const T* __restrict__ gates, int Lq, T (&re)[Cfg<Q>::NR],
                                          T (&im)[Cfg<Q>::NR], const Ctx<T>& c, T& inv) {
  inv = embed_w<T, Q>(pre, emb, re, im, c);
  for (int l = 0; l < Lq; ++l) {
    FwdGates<T, Q, 0>::run(re, im, gates + (size_t)l * Q * kGateStride, c);
    cnot_chain_w<T, Q, false>(re, im, c);
  }
}

// per-thread partial of <Z_i>, then group sum
template <typename T, int Q>
__device__ __forceinline__ void readout_w(const T (&re)[Cfg<Q>::NR], const T (&im)[Cfg<Q>::NR], const Ctx<T>& c, T (&out)[Q]) {
  using C = Cfg<Q>;
  T p[C::NR];
  T tot = T(0);
#pragma unroll
  for (int r = 0; r < C::NR; ++r) {
    p[r] = fma(re[r], re[r], im[r] * im[r]);
    tot += p[r];
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int P = Q - 1 - i;
    T s = T(0);
    if (P >= C::NT) {
#pragma unroll
      for (int r = 0; r < C::NR; ++r) s += ((r >> (P >= C::NT ? P - C::NT : 0)) & 1) ? -p[r] : p[r];
    } else {
      s = ((c.tsub >> P) & 1) ? -tot : tot;
    }
    out[i] = s;
  }
  group_sum<T, Q, Q>(out, c);
}

// 8 per-lane values -> warp totals, accumulated into acc[0..7] (shared memory, owned by this warp)
template <typename T>
__device__ __forceinline__ void warp_reduce8_acc(const T (&v)[8], T* acc, int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  T u[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const T keep = b4 ? v[4 + j] : v[j], send = b4 ? v[j] : v[4 + j];
    u[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  T w[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const T keep = b3 ? u[2 + j] : u[j], send = b3 ? u[j] : u[2 + j];
    w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  T z;
  {
    const T keep = b2 ? w[1] : w[0], send = b2 ? w[0] : w[1];
    z = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  z += __shfl_xor_sync(0xffffffffu, z, 2);
  z += __shfl_xor_sync(0xffffffffu, z, 1);
  if ((lane & 3) == 0) acc[lane >> 2] += z;  // element e = b4*4 + b3*2 + b2
  __syncwarp();
}

// one gate of the adjoint sweep on bit position P.  (pr,pi) = psi and (lr,li) = lambda AFTER the gate on entry, BEFORE
// it on exit.  gd = G^dagger (8 reals).  nacc: this warp's 8 accumulators for this (layer, wire).
template <typename T, int Q, int P>
__device__ __forceinline__ void adj_gate_w(T (&pr)[Cfg<Q>::NR], T (&pi)[Cfg<Q>::NR], T (&lr)[Cfg<Q>::NR], T (&li)[Cfg<Q>::NR],
                                           const T* __restrict__ gd, T* nacc, const Ctx<T>& c) {
  using C = Cfg<Q>;
  T n[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) n[e] = T(0);
  if constexpr (P >= C::NT) {
    constexpr int RB = P - C::NT;
#pragma unroll
    for (int r0 = 0; r0 < C::NR; ++r0) {
      if ((r0 >> RB) & 1) continue;
      const int r1 = r0 | (1 << RB);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ka = a ? r1 : r0, kc = cc ? r1 : r0;
          n[(a * 2 + cc) * 2 + 0] += lr[ka] * pr[kc] + li[ka] * pi[kc];
          n[(a * 2 + cc) * 2 + 1] += lr[ka] * pi[kc] - li[ka] * pr[kc];
        }
    }
    apply_gate_w<T, Q, P>(pr, pi, gd, c);
    apply_gate_w<T, Q, P>(lr, li, gd, c);
  } else {
    const bool hi = (c.tsub >> P) & 1;
    const T ar = hi ? gd[6] : gd[0], ai = hi ? gd[7] : gd[1];
    const T br = hi ? gd[4] : gd[2], bi = hi ? gd[5] : gd[3];
    T mmr = T(0), mmi = T(0), mpr = T(0), mpi = T(0);
    partner_begin<T, Q, P>(pr, pi, c);
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const T mr = pr[r], mi = pi[r];
      T qr, qi;
      partner_get<T, Q, P>(mr, mi, r, c, qr, qi);
      mmr += lr[r] * mr + li[r] * mi;
      mmi += lr[r] * mi - li[r] * mr;
      mpr += lr[r] * qr + li[r] * qi;
      mpi += lr[r] * qi - li[r] * qr;
      pr[r] = ar * mr - ai * mi + br * qr - bi * qi;
      pi[r] = ar * mi + ai * mr + br * qi + bi * qr;
    }
    partner_end<Q, P>();
    // row a = my bit; column = my bit (mm) / the other bit (mp)
    n[0] = hi ? T(0) : mmr;  n[1] = hi ? T(0) : mmi;
    n[2] = hi ? T(0) : mpr;  n[3] = hi ? T(0) : mpi;
    n[4] = hi ? mpr : T(0);  n[5] = hi ? mpi : T(0);
    n[6] = hi ? mmr : T(0);  n[7] = hi ? mmi : T(0);
    partner_begin<T, Q, P>(lr, li, c);
#pragma unroll
    for (int r = 0; r < C::NR; ++r) {
      const T mr = lr[r], mi = li[r];
      T qr, qi;
      partner_get<T, Q, P>(mr, mi, r, c, qr, qi);
      lr[r] = ar * mr - ai * mi + br * qr - bi * qi;
      li[r] = ar * mi + ai * mr + br * qi + bi * qr;
    }
    partner_end<Q, P>();
  }
  warp_reduce8_acc<T>(n, nacc, c.lane);
}

// adjoint sweep of one lane-bit gate with a run-time bit position (see apply_gate_lane_rt)
template <typename T, int Q>
__device__ __forceinline__ void adj_gate_lane_rt(T (&pr)[Cfg<Q>::NR], T (&pi)[Cfg<Q>::NR], T (&lr)[Cfg<Q>::NR], T (&li)[Cfg<Q>::NR],
                                                 const T* __restrict__ gd, T* nacc, int P, const Ctx<T>& c) {
  using C = Cfg<Q>;
  const bool hi = (c.tsub >> P) & 1;
  const int mask = 1 << P;
  const T ar = hi ? gd[6] : gd[0], ai = hi ? gd[7] : gd[1];
  const T br = hi ? gd[4] : gd[2], bi = hi ? gd[5] : gd[3];
  T mmr = T(0), mmi = T(0), mpr = T(0), mpi = T(0);
#pragma unroll
  for (int r = 0; r < C::NR; ++r) {
    const T mr = pr[r], mi = pi[r];
    const T qr = __shfl_xor_sync(0xffffffffu, mr, mask), qi = __shfl_xor_sync(0xffffffffu, mi, mask);
    mmr += lr[r] * mr + li[r] * mi;
    mmi += lr[r] * mi - li[r] * mr;
    mpr += lr[r] * qr + li[r] * qi;
    mpi += lr[r] * qi - li[r] * qr;
    pr[r] = ar * mr - ai * mi + br * qr - bi * qi;
    pi[r] = ar * mi + ai * mr + br * qi + bi * qr;
  }
  T n[8];
  n[0] = hi ? T(0) : mmr;  n[1] = hi ? T(0) : mmi;
  n[2] = hi ? T(0) : mpr;  n[3] = hi ? T(0) : mpi;
  n[4] = hi ? mpr : T(0);  n[5] = hi ? mpi : T(0);
  n[6] = hi ? mmr : T(0);  n[7] = hi ? mmi : T(0);
#pragma unroll
  for (int r = 0; r < C::NR; ++r) {
    const T mr = lr[r], mi = li[r];
    const T qr = __shfl_xor_sync(0xffffffffu, mr, mask), qi = __shfl_xor_sync(0xffffffffu, mi, mask);
    lr[r] = ar * mr - ai * mi + br * qr - bi * qi;
    li[r] = ar * mi + ai * mr + br * qi + bi * qr;
  }
  warp_reduce8_acc<T>(n, nacc, c.lane);
}

template <typename T, int Q, int P, int PLO>
struct AdjGatesDown {  // compile-time bit positions P down to PLO (register bits and warp bits)
  static __device__ __forceinline__ void run(T (&pr)[Cfg<Q>::NR], T (&pi)[Cfg<Q>::NR], T (&lr)[Cfg<Q>::NR], T (&li)[Cfg<Q>::NR],
                                             const T* __restrict__ gl, T* nl, const Ctx<T>& c) {
    if constexpr (P >= PLO) {
      adj_gate_w<T, Q, P>(pr, pi, lr, li, gl + (Q - 1 - P) * kGateStride + 8, nl + (Q - 1 - P) * 8, c);
      AdjGatesDown<T, Q, P - 1, PLO>::run(pr, pi, lr, li, gl, nl, c);
    }
  }
};

template <typename T, int Q, int P>
struct AdjGates {  // adjoint sweep of one Rot layer (P kept for source compatibility: always Q - 1)
  static __device__ __forceinline__ void run(T (&pr)[Cfg<Q>::NR], T (&pi)[Cfg<Q>::NR], T (&lr)[Cfg<Q>::NR], T (&li)[Cfg<Q>::NR],
                                             const T* __restrict__ gl, T* nl, const Ctx<T>& c) {
    using C = Cfg<Q>;
    constexpr int NTL = C::NT < 5 ? C::NT : 5;
    if constexpr (C::R >= 3) {  // register bits, rolled (see FwdGates): psi and lambda rotate together
#pragma unroll 1
      for (int i = 0; i < C::R; ++i) {
        adj_gate_w<T, Q, C::NT>(pr, pi, lr, li, gl + (Q - 1 - C::NT - i) * kGateStride + 8, nl + (Q - 1 - C::NT - i) * 8, c);
        rotate_regs<T, Q>(pr, pi);
        rotate_regs<T, Q>(lr, li);
      }
      AdjGatesDown<T, Q, C::NT - 1, NTL>::run(pr, pi, lr, li, gl, nl, c);  // warp bits
    } else {
      AdjGatesDown<T, Q, Q - 1, NTL>::run(pr, pi, lr, li, gl, nl, c);
    }
#pragma unroll 1
    for (int p = NTL - 1; p >= 0; --p)
      adj_gate_lane_rt<T, Q>(pr, pi, lr, li, gl + (Q - 1 - p) * kGateStride + 8, nl + (Q - 1 - p) * 8, p, c);
  }
};

// angle embedding: d L / d pre_i = 2 Re <lambda| A_i |psi>, A = (dE/dx) E^dag = [[-i/2, -e^{-ix}/2], [e^{ix}/2, i/2]]
template <typename T, int Q, int P>
struct AngleGrad {
  static __device__ __forceinline__ void run(const T (&pr)[Cfg<Q>::NR], const T (&pi)[Cfg<Q>::NR], const T (&lr)[Cfg<Q>::NR],
                                             const T (&li)[Cfg<Q>::NR], const T (&pre)[Q], T (&gpre)[Q], const Ctx<T>& c) {
    using C = Cfg<Q>;
    if constexpr (P < Q) {
      const int i = Q - 1 - P;
      T sx, cx;
      sincos(pre[i], &sx, &cx);
      T t = T(0);
      if constexpr (P >= C::NT) {
        constexpr int RB = P - C::NT;
#pragma unroll
        for (int r = 0; r < C::NR; ++r) {
          const bool b = (r >> RB) & 1;
          const int rp = r ^ (1 << RB);
          // (A psi)_r = A_bb psi_r + A_b,1-b psi_rp ;  A_00 = -i/2, A_01 = -(cx - i sx)/2, A_10 = (cx + i sx)/2, A_11 = i/2
          const T dr = b ? T(0.5) * cx : T(-0.5) * cx, di = T(0.5) * sx;
          const T sgn = b ? T(0.5) : T(-0.5);
          const T vr = -sgn * pi[r] + dr * pr[rp] - di * pi[rp];   // (i sgn)(pr + i pi) = -sgn pi + i sgn pr
          const T vi = sgn * pr[r] + dr * pi[rp] + di * pr[rp];
          t += lr[r] * vr + li[r] * vi;
        }
      } else {
        const bool b = (c.tsub >> P) & 1;
        const T dr = b ? T(0.5) * cx : T(-0.5) * cx, di = T(0.5) * sx;
        const T sgn = b ? T(0.5) : T(-0.5);
        partner_begin<T, Q, P>(pr, pi, c);
#pragma unroll
        for (int r = 0; r < C::NR; ++r) {
          T qr, qi;
          partner_get<T, Q, P>(pr[r], pi[r], r, c, qr, qi);
          const T vr = -sgn * pi[r] + dr * qr - di * qi;
          const T vi = sgn * pr[r] + dr * qi + di * qr;
          t += lr[r] * vr + li[r] * vi;
        }
        partner_end<Q, P>();
      }
      gpre[i] = T(2) * t;
      AngleGrad<T, Q, P + 1>::run(pr, pi, lr, li, pre, gpre, c);
    }
  }
};

template <typename T, int Q>
__device__ __forceinline__ Ctx<T> make_ctx(T* xbuf) {
  using C = Cfg<Q>;
  const int tid = threadIdx.x;
  Ctx<T> c;
  c.tsub = tid % C::TG;
  c.lane = tid & 31;
  T* gb = xbuf + (size_t)(tid / C::TG) * C::GS;
  c.xr = gb;
  c.xi = gb + C::N;
  return c;
}

template <typename T, int Q>
__host__ __device__ constexpr size_t wcirc_smem_bytes(int Lq, bool bwd) {
  using C = Cfg<Q>;
  size_t n = (size_t)Lq * Q * kGateStride + (size_t)C::GPC * C::GS + 64;
  if (bwd) n += (size_t)C::WARPS * Lq * Q * 8;
  return n * sizeof(T);
}

template <typename T, int Q>
__global__ void __launch_bounds__(Cfg<Q>::THREADS) wcirc_fwd_kernel(const WArgs<T> a) {
  using C = Cfg<Q>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* gates = reinterpret_cast<T*>(smem_raw);
  T* xbuf = gates + (size_t)a.Lq * Q * kGateStride;
  const int tid = threadIdx.x;
  for (int g = tid; g < a.Lq * Q; g += C::THREADS) make_gate<T>(a.qw + g * 3, gates + g * kGateStride);
  __syncthreads();
  const Ctx<T> c = make_ctx<T, Q>(xbuf);
  const int gid = tid / C::TG;
  for (long long base = (long long)blockIdx.x * C::GPC; base < a.W; base += (long long)gridDim.x * C::GPC) {
    const long long w = base + gid;
    const bool valid = w < a.W;
    T pre[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) pre[j] = valid ? a.pre[w * Q + j] : (j == 0 ? T(1) : T(0));
    T re[C::NR], im[C::NR], out[Q], inv;
    forward_w<T, Q>(pre, a.emb, gates, a.Lq, re, im, c, inv);
    readout_w<T, Q>(re, im, c, out);
    if (valid && c.tsub == 0) {
#pragma unroll
      for (int j = 0; j < Q; ++j) a.out[w * Q + j] = out[j];
    }
    __syncthreads();  // keeps the CTA's warps in phase so they share instruction-cache lines (the loop body is >> L0)
  }
}

template <typename T, int Q>
__global__ void __launch_bounds__(Cfg<Q>::THREADS) wcirc_bwd_kernel(const WArgs<T> a) {
  using C = Cfg<Q>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* gates = reinterpret_cast<T*>(smem_raw);
  T* xbuf = gates + (size_t)a.Lq * Q * kGateStride;
  T* nacc = xbuf + (size_t)C::GPC * C::GS + 32;  // [WARPS][Lq*Q*8]
  const int NE = a.Lq * Q * 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int g = tid; g < a.Lq * Q; g += C::THREADS) make_gate<T>(a.qw + g * 3, gates + g * kGateStride);
  for (int e = tid; e < C::WARPS * NE; e += C::THREADS) nacc[e] = T(0);
  __syncthreads();
  const Ctx<T> c = make_ctx<T, Q>(xbuf);
  T* mynacc = nacc + (size_t)warp * NE;
  const int gid = tid / C::TG;
  for (long long base = (long long)blockIdx.x * C::GPC; base < a.W; base += (long long)gridDim.x * C::GPC) {
    const long long w = base + gid;
    const bool valid = w < a.W;
    T pre[Q], gout[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      pre[j] = valid ? a.pre[w * Q + j] : (j == 0 ? T(1) : T(0));
      gout[j] = valid ? a.gout[w * Q + j] : T(0);
    }
    T pr[C::NR], pi[C::NR], lr[C::NR], li[C::NR], inv;
    forward_w<T, Q>(pre, a.emb, gates, a.Lq, pr, pi, c, inv);
    // lambda = (sum_i gout_i Z_i) psi
    {
      T dthr = T(0);
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        const int P = Q - 1 - i;
        if (P < C::NT) dthr += ((c.tsub >> P) & 1) ? -gout[i] : gout[i];
      }
#pragma unroll
      for (int r = 0; r < C::NR; ++r) {
        T d = dthr;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const int P = Q - 1 - i;
          if (P >= C::NT) d += ((r >> (P >= C::NT ? P - C::NT : 0)) & 1) ? -gout[i] : gout[i];
        }
        lr[r] = d * pr[r];
        li[r] = d * pi[r];
      }
    }
    for (int l = a.Lq - 1; l >= 0; --l) {
      cnot_chain_w<T, Q, true>(pr, pi, c);
      cnot_chain_w<T, Q, true>(lr, li, c);
      AdjGates<T, Q, Q - 1>::run(pr, pi, lr, li, gates + (size_t)l * Q * kGateStride, mynacc + (size_t)l * Q * 8, c);
    }
    T gpre[Q];
    if (a.emb == kEmbAmplitude) {
      // psi0 = v real: dL/dv_k = 2 Re lambda0_k (k < Q); v = pre / ||pre||
      T gv[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) gv[k] = (c.tsub == k % C::TG) ? T(2) * lr[k / C::TG] : T(0);
      group_sum<T, Q, Q>(gv, c);
      T dot = T(0);
#pragma unroll
      for (int k = 0; k < Q; ++k) dot = fma(gv[k], pre[k] * inv, dot);
#pragma unroll
      for (int k = 0; k < Q; ++k) gpre[k] = (gv[k] - pre[k] * inv * dot) * inv;
    } else {
      AngleGrad<T, Q, 0>::run(pr, pi, lr, li, pre, gpre, c);
      group_sum<T, Q, Q>(gpre, c);
    }
    if (valid && c.tsub == 0) {
#pragma unroll
      for (int j = 0; j < Q; ++j) a.gpre[w * Q + j] = gpre[j];
    }
    __syncthreads();  // in-phase warps share instruction-cache lines
  }
  __syncthreads();
  for (int e = tid; e < a.PA; e += C::THREADS) {
    T s = T(0);
    if (e < NE)
      for (int w = 0; w < C::WARPS; ++w) s += nacc[(size_t)w * NE + e];
    a.part[(size_t)blockIdx.x * a.PA + e] = s;
  }
}

// sum the per-CTA rows (fp64), N -> M = N conj(G), chain rule -> d/d(phi, theta, omega).  One 256-thread CTA per gate: thread t
// takes the rows t, t + 256, ... (two 16-byte loads per row, independent across threads -- one warp per gate walked ~40 dependent
// rows per lane and took 17 us for a few hundred KB), then the 8 warps meet in shared memory in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256) wcirc_finalize_kernel(const T* __restrict__ part, const T* __restrict__ qw, T* __restrict__ gqw,
                                                             int G, int PA, int ngates) {
  __shared__ double red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gate = blockIdx.x;
  if (gate >= ngates) return;
  double n[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) n[e] = 0.0;
  for (int g = threadIdx.x; g < G; g += 256)
#pragma unroll
    for (int e = 0; e < 8; ++e) n[e] += (double)part[(size_t)g * PA + gate * 8 + e];
#pragma unroll
  for (int e = 0; e < 8; ++e) n[e] = warp_sum<double>(n[e]);
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][e] = n[e];
  __syncthreads();
  if (threadIdx.x != 0) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    double v = red[0][e];
#pragma unroll
    for (int w = 1; w < 8; ++w) v += red[w][e];
    n[e] = v;
  }
  double w3[3] = {(double)qw[gate * 3], (double)qw[gate * 3 + 1], (double)qw[gate * 3 + 2]};
  double s, c, sp, cp, sm, cm;
  sincos(0.5 * w3[1], &s, &c);
  sincos(0.5 * (w3[0] + w3[2]), &sp, &cp);
  sincos(0.5 * (w3[0] - w3[2]), &sm, &cm);
  const double Gm[8] = {cp * c, -sp * c, -cm * s, -sm * s, cm * s, -sm * s, cp * c, sp * c};
  // M_ab = sum_c N_ac conj(G_cb)
  double m[8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      double mr = 0.0, mi = 0.0;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const double nr = n[(a * 2 + cc) * 2], ni = n[(a * 2 + cc) * 2 + 1];
        const double gr = Gm[(cc * 2 + b) * 2], gi = -Gm[(cc * 2 + b) * 2 + 1];
        mr += nr * gr - ni * gi;
        mi += nr * gi + ni * gr;
      }
      m[(a * 2 + b) * 2] = mr;
      m[(a * 2 + b) * 2 + 1] = mi;
    }
  double g3[3];
  gate_grad_to_angles(w3, m, g3);
#pragma unroll
  for (int e = 0; e < 3; ++e) gqw[gate * 3 + e] = (T)g3[e];
}

// host-visible launchers, instantiated per (T, Q) in qw_circuit_warp_inst.cu
template <typename T, int Q>
int wcirc_forward_tq(const WArgs<T>& a, int grid, cudaStream_t st);
template <typename T, int Q>
int wcirc_backward_tq(const WArgs<T>& a, int grid, cudaStream_t st);
template <typename T>
int wcirc_finalize_t(const T* part, const T* qw, T* gqw, int G, int PA, int ngates, cudaStream_t st);

__host__ inline int wcirc_gpc(int q) { return q <= 4 ? 128 : q <= 9 ? 128 >> (q - 4) : (q == 10 ? 4 : 1); }
__host__ inline int wcirc_PA(int q, int Lq) { return (int)align_up((size_t)Lq * q * 8, 32); }

}  // namespace wc
}  // namespace qw
