// Per-thread batched statevector simulator (q <= 4: the 2^q complex amplitudes of ONE window live in the
// registers of ONE thread) + adjoint differentiation.  Replaces the PennyLane QNode of
// /root/reference/quantum_whisper.py:64-85 (pad -> AmplitudeEmbedding(normalize) -> Rot per wire ->
// CNOT chain -> <Z_i>) and its backprop-through-the-simulator backward.
//
// Conventions (PennyLane default.qubit): wire 0 is the most significant bit of the basis index;
// Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi).
//
// Sparse-support tracking: the amplitude-embedded state has only its first q amplitudes non-zero
// (SURVEY.md 8-a5), so the first Rot layer is applied LSB wire first and every gate only touches the pairs
// whose support mask (a compile-time constant) says can be non-zero.  This is still a state-vector
// simulation of the general circuit (any weights, any number of layers); it just never multiplies by a
// literal zero.  Later layers run with the full mask.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qw {

constexpr int kEmbAmplitude = 0;
constexpr int kEmbAngle = 1;
constexpr int kGateStride = 16;  // per wire: G (8 reals: 00r 00i 01r 01i 10r 10i 11r 11i) then G^dagger (8)

__host__ __device__ constexpr uint32_t full_mask(int Q) { return (Q >= 5) ? 0xffffffffu : ((1u << (1 << Q)) - 1u); }

// support after a 1-qubit gate on bit position p
__host__ __device__ constexpr uint32_t mask_closure(uint32_t m, int p, int N) {
  uint32_t r = m;
  for (int k = 0; k < N; ++k) {
    if (!((k >> p) & 1)) {
      int k1 = k | (1 << p);
      if (((m >> k) & 1u) || ((m >> k1) & 1u)) r |= (1u << k) | (1u << k1);
    }
  }
  return r;
}
// support before stage s of a layer whose gates are applied on bit positions 0,1,...,Q-1 in that order
__host__ __device__ constexpr uint32_t stage_mask(int Q, uint32_t m0, int s) {
  uint32_t m = m0;
  for (int i = 0; i < s; ++i) m = mask_closure(m, i, 1 << Q);
  return m;
}
// CNOT(0,1), CNOT(1,2), ... , CNOT(Q-2,Q-1) applied in that order: amplitude at index k moves to cnot_fwd(k)
__host__ __device__ constexpr int cnot_fwd(int k, int Q) {
  for (int i = 0; i + 1 < Q; ++i)
    if ((k >> (Q - 1 - i)) & 1) k ^= 1 << (Q - 2 - i);
  return k;
}

// out += g * a   (complex), skipping whatever is known to be zero at compile time
template <typename T, bool REAL>
__device__ __forceinline__ void cmac(T& o_r, T& o_i, T gr, T gi, T ar, T ai) {
  o_r = fma(gr, ar, o_r);
  o_i = fma(gi, ar, o_i);
  if (!REAL) {
    o_r = fma(-gi, ai, o_r);
    o_i = fma(gr, ai, o_i);
  }
}

// 1-qubit gate g (8 reals, row major complex 2x2) on bit position P.  INMASK: amplitudes that may be
// non-zero on entry; OUTMASK: amplitudes the caller needs on exit.  REAL: imaginary parts on entry are zero.
template <typename T, int Q, int P, uint32_t INMASK, uint32_t OUTMASK, bool REAL>
__device__ __forceinline__ void apply_gate(T (&re)[1 << Q], T (&im)[1 << Q], const T* __restrict__ g) {
  constexpr int N = 1 << Q;
  const T g00r = g[0], g00i = g[1], g01r = g[2], g01i = g[3], g10r = g[4], g10i = g[5], g11r = g[6], g11i = g[7];
#pragma unroll
  for (int k0 = 0; k0 < N; ++k0) {
    if ((k0 >> P) & 1) continue;
    const int k1 = k0 | (1 << P);
    const bool n0 = (INMASK >> k0) & 1u, n1 = (INMASK >> k1) & 1u;
    const bool w0 = (OUTMASK >> k0) & 1u, w1 = (OUTMASK >> k1) & 1u;
    if (!(n0 || n1) || !(w0 || w1)) continue;
    const T a0r = n0 ? re[k0] : T(0), a0i = (n0 && !REAL) ? im[k0] : T(0);
    const T a1r = n1 ? re[k1] : T(0), a1i = (n1 && !REAL) ? im[k1] : T(0);
    if (w0) {
      T br = T(0), bi = T(0);
      if (n0) cmac<T, REAL>(br, bi, g00r, g00i, a0r, a0i);
      if (n1) cmac<T, REAL>(br, bi, g01r, g01i, a1r, a1i);
      re[k0] = br;
      im[k0] = bi;
    }
    if (w1) {
      T br = T(0), bi = T(0);
      if (n0) cmac<T, REAL>(br, bi, g10r, g10i, a0r, a0i);
      if (n1) cmac<T, REAL>(br, bi, g11r, g11i, a1r, a1i);
      re[k1] = br;
      im[k1] = bi;
    }
  }
}

// M_ab += conj(lam_a) * psi_b summed over the pairs of bit position P.  m[8] = 00r 00i 01r 01i 10r 10i 11r 11i.
// PSIMASK: support of psi (before the gate); REAL: psi is real.
template <typename T, int Q, int P, uint32_t PSIMASK, bool REAL>
__device__ __forceinline__ void accum_gate_grad(const T (&lr)[1 << Q], const T (&li)[1 << Q], const T (&pr)[1 << Q],
                                                const T (&pi)[1 << Q], T (&m)[8]) {
  constexpr int N = 1 << Q;
#pragma unroll
  for (int k0 = 0; k0 < N; ++k0) {
    if ((k0 >> P) & 1) continue;
    const int k1 = k0 | (1 << P);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int kb = b ? k1 : k0;
      if (!((PSIMASK >> kb) & 1u)) continue;
      const T br = pr[kb], bi = REAL ? T(0) : pi[kb];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ka = a ? k1 : k0;
        // conj(l) * p = (lr*pr + li*pi) + i (lr*pi - li*pr)
        m[(a * 2 + b) * 2 + 0] = fma(lr[ka], br, m[(a * 2 + b) * 2 + 0]);
        m[(a * 2 + b) * 2 + 1] = fma(-li[ka], br, m[(a * 2 + b) * 2 + 1]);
        if (!REAL) {
          m[(a * 2 + b) * 2 + 0] = fma(li[ka], bi, m[(a * 2 + b) * 2 + 0]);
          m[(a * 2 + b) * 2 + 1] = fma(lr[ka], bi, m[(a * 2 + b) * 2 + 1]);
        }
      }
    }
  }
}

template <typename T, int Q>
__device__ __forceinline__ void cnot_chain_fwd(T (&re)[1 << Q], T (&im)[1 << Q]) {
  constexpr int N = 1 << Q;
  T nr[N], ni[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    nr[cnot_fwd(k, Q)] = re[k];
    ni[cnot_fwd(k, Q)] = im[k];
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    re[k] = nr[k];
    im[k] = ni[k];
  }
}
template <typename T, int Q>
__device__ __forceinline__ void cnot_chain_bwd(T (&re)[1 << Q], T (&im)[1 << Q]) {
  constexpr int N = 1 << Q;
  T nr[N], ni[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    nr[k] = re[cnot_fwd(k, Q)];
    ni[k] = im[cnot_fwd(k, Q)];
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    re[k] = nr[k];
    im[k] = ni[k];
  }
}

// ---- one Rot layer, forward: stages s = 0..Q-1 act on bit position s (= wire Q-1-s)
template <typename T, int Q, uint32_t M0, bool REAL0, int S>
struct LayerFwd {
  static __device__ __forceinline__ void run(T (&re)[1 << Q], T (&im)[1 << Q], const T* __restrict__ gates) {
    if constexpr (S < Q) {
      constexpr uint32_t MIN = stage_mask(Q, M0, S);
      constexpr uint32_t MOUT = stage_mask(Q, M0, S + 1);
      apply_gate<T, Q, S, MIN, MOUT, (REAL0 && S == 0)>(re, im, gates + (Q - 1 - S) * kGateStride);
      LayerFwd<T, Q, M0, REAL0, S + 1>::run(re, im, gates);
    }
  }
};

// ---- one Rot layer, adjoint sweep: stages s = Q-1..0.  (pr,pi) = psi AFTER the layer on entry, BEFORE it on
// exit (only on the support); (lr,li) = lambda likewise.  macc: [Q][8] per-wire gate-gradient accumulators
// for this layer (added to).  For the sparse first layer the stage-0 "before" state is the known real embedded
// vector, so psi is not un-applied there.
template <typename T, int Q, uint32_t M0, bool REAL0, int S, typename Acc>
struct LayerBwd {
  static __device__ __forceinline__ void run(T (&pr)[1 << Q], T (&pi)[1 << Q], T (&lr)[1 << Q], T (&li)[1 << Q],
                                             const T* __restrict__ gates, Acc& acc) {
    if constexpr (S >= 0) {
      constexpr uint32_t MBEF = stage_mask(Q, M0, S);
      constexpr uint32_t MAFT = stage_mask(Q, M0, S + 1);
      constexpr bool REAL = REAL0 && S == 0;
      const T* gd = gates + (Q - 1 - S) * kGateStride + 8;  // G^dagger
      // psi_before = G^dagger psi_after, restricted to the "before" support
      apply_gate<T, Q, S, MAFT, MBEF, false>(pr, pi, gd);
      T m[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = T(0);
      accum_gate_grad<T, Q, S, MBEF, REAL>(lr, li, pr, pi, m);
      acc.add(Q - 1 - S, m);
      // lambda_before = G^dagger lambda_after, needed only on the "before" support
      apply_gate<T, Q, S, MAFT, MBEF, false>(lr, li, gd);
      LayerBwd<T, Q, M0, REAL0, S - 1, Acc>::run(pr, pi, lr, li, gates, acc);
    }
  }
};

template <typename T, int Q>
__device__ __forceinline__ void readout(const T (&re)[1 << Q], const T (&im)[1 << Q], T (&out)[Q]) {
  constexpr int N = 1 << Q;
  T p[N];
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = fma(re[k], re[k], im[k] * im[k]);
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    T s = T(0);
#pragma unroll
    for (int k = 0; k < N; ++k) s += ((k >> (Q - 1 - i)) & 1) ? -p[k] : p[k];
    out[i] = s;
  }
}

// Forward circuit with amplitude embedding. gates: [n_layers][Q][16] (shared memory). Leaves the final
// state in (re,im) and returns 1/||pre||.
template <typename T, int Q>
__device__ __forceinline__ T circuit_forward_amp(const T (&pre)[Q], const T* __restrict__ gates, int n_layers,
                                                 T (&re)[1 << Q], T (&im)[1 << Q], T (&out)[Q]) {
  constexpr int N = 1 << Q;
  constexpr uint32_t M0 = (1u << Q) - 1u;  // first Q amplitudes
  T ss = T(0);
#pragma unroll
  for (int j = 0; j < Q; ++j) ss = fma(pre[j], pre[j], ss);
  const T inv = T(1) / sqrt(ss);  // ||v|| = 0 -> inf/NaN exactly like the reference (quantum_whisper.py:74)
#pragma unroll
  for (int k = 0; k < N; ++k) {
    re[k] = (k < Q) ? pre[k] * inv : T(0);
    im[k] = T(0);
  }
  LayerFwd<T, Q, M0, true, 0>::run(re, im, gates);
  cnot_chain_fwd<T, Q>(re, im);
  for (int l = 1; l < n_layers; ++l) {
    LayerFwd<T, Q, full_mask(Q), false, 0>::run(re, im, gates + l * Q * kGateStride);
    cnot_chain_fwd<T, Q>(re, im);
  }
  readout<T, Q>(re, im, out);
  return inv;
}

// Adjoint backward given the final state (re,im) from circuit_forward_amp and the cotangent gout.
// Writes gpre (cotangent of pre) and adds per-wire gate-gradient matrices through acc.add(layer, wire, m[8]).
template <typename T, int Q, typename Acc>
__device__ __forceinline__ void circuit_backward_amp(const T (&pre)[Q], T inv, const T* __restrict__ gates, int n_layers,
                                                     T (&re)[1 << Q], T (&im)[1 << Q], const T (&gout)[Q], T (&gpre)[Q],
                                                     Acc& acc) {
  constexpr int N = 1 << Q;
  constexpr uint32_t M0 = (1u << Q) - 1u;
  T lr[N], li[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    T d = T(0);
#pragma unroll
    for (int i = 0; i < Q; ++i) d += ((k >> (Q - 1 - i)) & 1) ? -gout[i] : gout[i];
    lr[k] = d * re[k];
    li[k] = d * im[k];
  }
  for (int l = n_layers - 1; l >= 1; --l) {
    cnot_chain_bwd<T, Q>(re, im);
    cnot_chain_bwd<T, Q>(lr, li);
    acc.set_layer(l);
    LayerBwd<T, Q, full_mask(Q), false, Q - 1, Acc>::run(re, im, lr, li, gates + l * Q * kGateStride, acc);
  }
  cnot_chain_bwd<T, Q>(re, im);
  cnot_chain_bwd<T, Q>(lr, li);
  acc.set_layer(0);
  LayerBwd<T, Q, M0, true, Q - 1, Acc>::run(re, im, lr, li, gates, acc);
  // L = <psi0| U^dag O U |psi0>, psi0 = v real  =>  dL/dv_k = 2 Re(lambda0_k);  v = pre/||pre||
  T gv[Q];
  T dot = T(0);
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    gv[k] = T(2) * lr[k];
    dot = fma(gv[k], pre[k] * inv, dot);
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) gpre[k] = (gv[k] - pre[k] * inv * dot) * inv;
}

// Gate matrices from quantum_weights (phi,theta,omega), computed in the kernel's own precision (sincosf is
// accurate to ~2 ulp; no fast-math).  out: 16 values (G then G^dagger).
template <typename T>
__device__ __forceinline__ void make_gate(const T* __restrict__ w3, T* __restrict__ out) {
  const T phi = w3[0], th = w3[1], om = w3[2];
  T s, c, sp, cp, sm, cm;
  sincos(T(0.5) * th, &s, &c);
  sincos(T(0.5) * (phi + om), &sp, &cp);  // e^{-i(phi+om)/2} = cp - i sp
  sincos(T(0.5) * (phi - om), &sm, &cm);  // e^{+i(phi-om)/2} = cm + i sm
  const T g[8] = {cp * c, -sp * c, -cm * s, -sm * s, cm * s, -sm * s, cp * c, sp * c};
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = g[e];
  // dagger: (G^dag)_ab = conj(G_ba)
  out[8] = g[0];
  out[9] = -g[1];
  out[10] = g[4];
  out[11] = -g[5];
  out[12] = g[2];
  out[13] = -g[3];
  out[14] = g[6];
  out[15] = -g[7];
}

// Chain rule from the summed gate-gradient matrix M_ab = sum conj(lambda_a) psi_b to (phi,theta,omega):
// dL/dp = 2 Re sum_ab (dG_ab/dp) M_ab.
__device__ __forceinline__ void gate_grad_to_angles(const double* w3, const double* m, double* g3) {
  const double phi = w3[0], th = w3[1], om = w3[2];
  double s, c, sp, cp, sm, cm;
  sincos(0.5 * th, &s, &c);
  sincos(0.5 * (phi + om), &sp, &cp);
  sincos(0.5 * (phi - om), &sm, &cm);
  // G entries
  const double G[8] = {cp * c, -sp * c, -cm * s, -sm * s, cm * s, -sm * s, cp * c, sp * c};
  // d/dtheta
  const double Gt[8] = {-0.5 * cp * s, 0.5 * sp * s, -0.5 * cm * c, -0.5 * sm * c,
                        0.5 * cm * c,  -0.5 * sm * c, -0.5 * cp * s, -0.5 * sp * s};
  // multiply-by-(+-i/2) factors: dphi: 00:-i/2 01:+i/2 10:-i/2 11:+i/2 ; domega: 00:-i/2 01:-i/2 10:+i/2 11:+i/2
  const double sphi[4] = {-0.5, 0.5, -0.5, 0.5};
  const double som[4] = {-0.5, -0.5, 0.5, 0.5};
  double gphi = 0, gth = 0, gom = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const double gr = G[2 * e], gi = G[2 * e + 1], mr = m[2 * e], mi = m[2 * e + 1];
    // Re( (i f G) * M ) = f * Re( i (gr + i gi)(mr + i mi) ) = f * ( -(gr*mi + gi*mr) )
    const double re_iGM = -(gr * mi + gi * mr);
    gphi += 2.0 * sphi[e] * re_iGM;
    gom += 2.0 * som[e] * re_iGM;
    gth += 2.0 * (Gt[2 * e] * mr - Gt[2 * e + 1] * mi);
  }
  g3[0] = gphi;
  g3[1] = gth;
  g3[2] = gom;
}

}  // namespace qw
