// Per-frame arithmetic of the fused log-mel kernel: Hann window -> 400-point real FFT -> power, written so the
// same code runs on the device and on the host inside tests/native/logmel_host_check.cpp.
//
// Replaces torch.stft(n_fft=400, hop=160, window=hann(400)) + abs()**2 of
// /root/reference/whisper/whisper/audio.py:147-149.
//
// 400 = 16 x 25 Cooley-Tukey on the REAL frame, n = 25 n1 + n2, k = k1 + 16 k2:
//    pass A (per n2):  Y[k1] = sum_n1 w[n] x[n] W16^(n1 k1)  -- a 16-point real-input FFT, so only k1 = 0..8 exist
//                      (Y[16 - k1] = conj Y[k1]; Y[0], Y[8] real);  A[k1][n2] = Y[k1] W400^(n2 k1)
//    pass B (per k1):  X[k1 + 16 k2] = sum_n2 A[k1][n2] W25^(n2 k2)        (25 = 5 x 5 in registers)
// k1 = 0..8 and k2 = 0..24 give the bins k = k1 + 16 k2; for k2 <= 12 that is k <= 200 directly, for k2 >= 13 the bin
// is > 200 and its power is that of the mirror bin 400 - k (X[400 - k] = conj X[k]), which fills the residues
// 9..15 (mod 16): all 201 bins, no separate real-FFT untangling pass, and only |X|^2 is ever written.
// (The first version packed the frame into a 200-point complex FFT, 8 x 25, plus an untangling pass: 2 000 more
// instructions per frame and a second trip through the work array.)
// Work layout: 17 float planes of 25 (n2) values: plane 0 = Re A[0] (real), planes 2 k1 - 1 / 2 k1 = Re / Im A[k1].
#pragma once
#include "qw_logmel_tables.h"

#if defined(__CUDACC__)
#define QW_HD __host__ __device__ __forceinline__
#else
#define QW_HD inline
#endif

namespace qw {
namespace lm {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kNfreq = 201;
#if defined(__CUDACC__)
// Hann weights and W400 twiddles are read once per CTA with one n2 per LANE (divergent index): global memory, not the
// constant bank.  The 5 x 5 twiddles are indexed uniformly: constant bank.
__device__ const float d_win[400] = QW_TBL_WIN;
__device__ const float d_tw400_re[200] = QW_TBL_TW400_RE;  // [k1 - 1][n2]
__device__ const float d_tw400_im[200] = QW_TBL_TW400_IM;
__constant__ float c_tw25_re[25] = QW_TBL_TW25_RE;
__constant__ float c_tw25_im[25] = QW_TBL_TW25_IM;
#endif
static const float h_win[400] = QW_TBL_WIN;
static const float h_tw400_re[200] = QW_TBL_TW400_RE;
static const float h_tw400_im[200] = QW_TBL_TW400_IM;
static const float h_tw25_re[25] = QW_TBL_TW25_RE;
static const float h_tw25_im[25] = QW_TBL_TW25_IM;

#if defined(__CUDA_ARCH__)
#define QW_TBL(name, i) (c_##name[i])
#else
#define QW_TBL(name, i) (h_##name[i])
#endif

// ---- 8-point DFT of the REAL sequence x_j w_j: Y[0..4] (Y[8 - k] = conj Y[k]); yi[0] = yi[4] = 0 are not written.
// The Hann weights ride in the first butterflies: (x_j w_j) +- (x_{j+4} w_{j+4}) is one FMUL and two FFMAs.
QW_HD void rdft8w(float x0, float x1, float x2, float x3, float x4, float x5, float x6, float x7, float w0, float w1, float w2,
                  float w3, float w4, float w5, float w6, float w7, float (&yr)[5], float (&yi)[5]) {
  const float h = 0.70710678118654752440f;
  const float t0 = x4 * w4, t1 = x5 * w5, t2 = x6 * w6, t3 = x7 * w7;
  const float a0 = x0 * w0 + t0, a1 = x1 * w1 + t1, a2 = x2 * w2 + t2, a3 = x3 * w3 + t3;
  const float b0 = x0 * w0 - t0, b1 = x1 * w1 - t1, b2 = x2 * w2 - t2, b3 = x3 * w3 - t3;
  const float c0 = a0 + a2, c1 = a1 + a3;
  yr[0] = c0 + c1;
  yr[4] = c0 - c1;
  yr[2] = a0 - a2;
  yi[2] = a3 - a1;                 // Y[2] = (a0 - a2) - i (a1 - a3)
  const float s = b1 - b3, t = b1 + b3;
  yr[1] = b0 + h * s;              // Y[1] = b0 + W8 b1 + W8^2 b2 + W8^3 b3
  yi[1] = -(b2 + h * t);
  yr[3] = b0 - h * s;              // Y[3] = b0 + W8^3 b1 + W8^6 b2 + W8^9 b3
  yi[3] = b2 - h * t;
}

// ---- 16-point DFT of the REAL sequence x_j w_j: X[0..8]; xi[0] = xi[8] = 0 are not written.
// X[k] = E[k] + W16^k O[k] with E / O the real 8-point DFTs of the even / odd samples; X[8 - k] = conj(E[k] - W16^k O[k]).
QW_HD void rfft16w(const float (&x)[16], const float (&w)[16], float (&xr)[9], float (&xi)[9]) {
  float er[5], ei[5], orr[5], oi[5];
  rdft8w(x[0], x[2], x[4], x[6], x[8], x[10], x[12], x[14], w[0], w[2], w[4], w[6], w[8], w[10], w[12], w[14], er, ei);
  rdft8w(x[1], x[3], x[5], x[7], x[9], x[11], x[13], x[15], w[1], w[3], w[5], w[7], w[9], w[11], w[13], w[15], orr, oi);
  xr[0] = er[0] + orr[0];
  xr[8] = er[0] - orr[0];
  xr[4] = er[4];                   // W16^4 = -i, E[4] and O[4] real
  xi[4] = -orr[4];
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin (pi / 8)
  const float h = 0.70710678118654752440f;
  // W16^k = (c, -s):  T = W16^k O[k] = (c Or + s Oi, c Oi - s Or)
  {
    const float tr = c1 * orr[1] + s1 * oi[1], ti = c1 * oi[1] - s1 * orr[1];
    xr[1] = er[1] + tr; xi[1] = ei[1] + ti;
    xr[7] = er[1] - tr; xi[7] = ti - ei[1];
  }
  {
    const float tr = h * (orr[2] + oi[2]), ti = h * (oi[2] - orr[2]);
    xr[2] = er[2] + tr; xi[2] = ei[2] + ti;
    xr[6] = er[2] - tr; xi[6] = ti - ei[2];
  }
  {
    const float tr = s1 * orr[3] + c1 * oi[3], ti = s1 * oi[3] - c1 * orr[3];
    xr[3] = er[3] + tr; xi[3] = ei[3] + ti;
    xr[5] = er[3] - tr; xi[5] = ti - ei[3];
  }
}

// ---- forward 5-point DFT in place on (r[o], r[o+s], ..., r[o+4s])
template <int O, int S, int N>
QW_HD void dft5(float (&r)[N], float (&i)[N]) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2pi/5), sin(4pi/5)
  const float x0r = r[O], x0i = i[O];
  const float t1r = r[O + S] + r[O + 4 * S], t1i = i[O + S] + i[O + 4 * S];
  const float t2r = r[O + 2 * S] + r[O + 3 * S], t2i = i[O + 2 * S] + i[O + 3 * S];
  const float t3r = r[O + S] - r[O + 4 * S], t3i = i[O + S] - i[O + 4 * S];
  const float t4r = r[O + 2 * S] - r[O + 3 * S], t4i = i[O + 2 * S] - i[O + 3 * S];
  const float m1r = x0r + c1 * t1r + c2 * t2r, m1i = x0i + c1 * t1i + c2 * t2i;
  const float m2r = x0r + c2 * t1r + c1 * t2r, m2i = x0i + c2 * t1i + c1 * t2i;
  const float n1r = s1 * t3r + s2 * t4r, n1i = s1 * t3i + s2 * t4i;
  const float n2r = s2 * t3r - s1 * t4r, n2i = s2 * t3i - s1 * t4i;
  r[O] = x0r + t1r + t2r;
  i[O] = x0i + t1i + t2i;
  // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2     (-i (a + ib) = b - ia)
  r[O + S] = m1r + n1i;      i[O + S] = m1i - n1r;
  r[O + 4 * S] = m1r - n1i;  i[O + 4 * S] = m1i + n1r;
  r[O + 2 * S] = m2r + n2i;  i[O + 2 * S] = m2i - n2r;
  r[O + 3 * S] = m2r - n2i;  i[O + 3 * S] = m2i + n2r;
}

template <int B_>
struct Dft5Cols {
  template <int N>
  static QW_HD void run(float (&r)[N], float (&i)[N]) {
    if constexpr (B_ < 5) {
      dft5<B_, 5, N>(r, i);  // over a for fixed b: elements 5a + b
      Dft5Cols<B_ + 1>::run(r, i);
    }
  }
};
template <int C_>
struct Dft5Rows {
  template <int N>
  static QW_HD void run(float (&r)[N], float (&i)[N]) {
    if constexpr (C_ < 5) {
      dft5<5 * C_, 1, N>(r, i);  // over b for fixed c: elements 5c + b
      Dft5Rows<C_ + 1>::run(r, i);
    }
  }
};

// ---- forward 25-point DFT.  in: x[n2], n2 = 5a + b.  out: X[k2] returned at index 5c + d with k2 = c + 5d.
QW_HD void dft25(float (&r)[25], float (&i)[25]) {
  Dft5Cols<0>::run(r, i);  // now element 5c + b holds sum_a x[5a+b] W5^(a c)
#pragma unroll
  for (int c = 1; c < 5; ++c) {
#pragma unroll
    for (int b = 1; b < 5; ++b) {
      const float wr = QW_TBL(tw25_re, b * 5 + c), wi = QW_TBL(tw25_im, b * 5 + c);
      const float xr = r[5 * c + b], xi = i[5 * c + b];
      r[5 * c + b] = xr * wr - xi * wi;
      i[5 * c + b] = xr * wi + xi * wr;
    }
  }
  Dft5Rows<0>::run(r, i);  // element 5c + d holds X[c + 5d]
}

constexpr int kPlanes = 17;  // work planes per frame: Re A[0], then Re / Im of A[1..8]

// Pass A for one (frame, n2): 16 taps x Hann -> real FFT16 -> twiddle -> 17 plane values.
//   x[n1]       the frame's sample 25 n1 + n2;  w[n1] its Hann weight (on the device the weights and twiddles of a lane's
//               n2 sit in registers for the whole kernel)
//   twr / twi   W400^(n2 k1) for k1 = 1..8
//   put(pl, v)  stores plane pl of this (frame, n2)
template <typename Put>
QW_HD void pass_a_one(const float (&x)[16], const float (&w)[16], const float (&twr)[8], const float (&twi)[8], Put&& put) {
  float xr[9], xi[9];
  rfft16w(x, w, xr, xi);
  put(0, xr[0]);
#pragma unroll
  for (int k1 = 1; k1 < 8; ++k1) {
    const float wr = twr[k1 - 1], wi = twi[k1 - 1];
    put(2 * k1 - 1, xr[k1] * wr - xi[k1] * wi);
    put(2 * k1, xr[k1] * wi + xi[k1] * wr);
  }
  put(15, xr[8] * twr[7]);
  put(16, xr[8] * twi[7]);
}

// Pass B for one (frame, k1): 25-point DFT over n2 of plane pair k1 -> power of the bins k1 + 16 k2.
//   getr(n2) / geti(n2)   Re / Im A[k1][n2] (geti is not called for k1 == 0: A[0] is real)
//   putp(k, v)            stores P[k], k <= 200
// k1 == 0 and k1 == 8 produce each of their bins twice (k2 and 25 - k2 resp. 24 - k2 mirror onto each other): the
// mirrored copies are skipped.
template <typename GetR, typename GetI, typename PutP>
QW_HD void pass_b_one(int k1, GetR&& getr, GetI&& geti, PutP&& putp) {
  float r[25], i[25];
#pragma unroll
  for (int n2 = 0; n2 < 25; ++n2) {
    r[n2] = getr(n2);
    i[n2] = k1 == 0 ? 0.f : geti(n2);
  }
  dft25(r, i);
  const bool mirror = k1 != 0 && k1 != 8;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const int k2 = c + 5 * d;
      const float p = r[5 * c + d] * r[5 * c + d] + i[5 * c + d] * i[5 * c + d];
      if (k2 <= 12) putp(k1 + 16 * k2, p);
      else if (mirror) putp(400 - 16 * k2 - k1, p);
    }
  }
}

}  // namespace lm
}  // namespace qw
