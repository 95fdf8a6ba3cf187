// Per-frame arithmetic of the fused log-mel kernel: Hann window -> 400-point real FFT -> power, written so the
// same code runs per lane on the device (one STFT frame per lane, the frame's 400 work floats in a
// shared-memory column) and on the host inside tests/native/logmel_host_check.cu.
//
// Replaces torch.stft(n_fft=400, hop=160, window=hann(400)) + abs()**2 of
// /root/reference/whisper/whisper/audio.py:147-149.
//
// 400-point real FFT = 200-point complex FFT of z[n] = x[2n] + i x[2n+1], then untangling.
// 200 = 8 x 25 Cooley-Tukey:  n = 25 n1 + n2,  k = k1 + 8 k2:
//    pass A (per n2):  Y[k1] = sum_n1 z[25 n1 + n2] W8^(n1 k1);  A[k1][n2] = Y[k1] W200^(n2 k1)
//    pass B (per k1):  Z[k1 + 8 k2] = sum_n2 A[k1][n2] W25^(n2 k2)          (25 = 5 x 5 in registers)
// Work-column layout (floats): complex slot s holds re at 2s, im at 2s+1.  Pass A stores A[k1][n2] at slot
// n2*8 + k1; pass B reads slots {n2*8 + k1} and writes Z[k1 + 8 k2] to slot k2*8 + k1 -- the same set, so it is in
// place per k1 and leaves the spectrum in NATURAL order (slot k = Z[k]).  The untangle pass overwrites re-slots with
// the power spectrum: P[k] at float 2k for k < 200, P[200] at float 1 (the im part of Z[0]'s slot).
// The G warps of a CTA split every pass by index (n2, k1, k, mel row = g, g+G, ...); a __syncthreads()
// separates the passes.  Every table index is uniform across a warp -> constant-cache broadcasts.
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
// audio staging skew: sample s of the tile sits at float s + (s >> 5).  Frames are 160 samples apart, so lane f reads
// 165 f + j + (j >> 5): 165 is odd -> the 32 lanes hit 32 distinct banks for every tap j.
constexpr int kLanePitch = 165;

#if defined(__CUDACC__)
__constant__ float c_win[400] = QW_TBL_WIN;
__constant__ float c_tw200_re[200] = QW_TBL_TW200_RE;
__constant__ float c_tw200_im[200] = QW_TBL_TW200_IM;
__constant__ float c_tw25_re[25] = QW_TBL_TW25_RE;
__constant__ float c_tw25_im[25] = QW_TBL_TW25_IM;
__constant__ float c_un_re[101] = QW_TBL_UN_RE;
__constant__ float c_un_im[101] = QW_TBL_UN_IM;
#endif
static const float h_win[400] = QW_TBL_WIN;
static const float h_tw200_re[200] = QW_TBL_TW200_RE;
static const float h_tw200_im[200] = QW_TBL_TW200_IM;
static const float h_tw25_re[25] = QW_TBL_TW25_RE;
static const float h_tw25_im[25] = QW_TBL_TW25_IM;
static const float h_un_re[101] = QW_TBL_UN_RE;
static const float h_un_im[101] = QW_TBL_UN_IM;

#if defined(__CUDA_ARCH__)
#define QW_TBL(name, i) (c_##name[i])
#else
#define QW_TBL(name, i) (h_##name[i])
#endif

QW_HD int pslot(int k) { return k < 200 ? 2 * k : 1; }  // float slot of P[k], k <= 200
QW_HD int skew(int s) { return s + (s >> 5); }           // staging index of tile sample s
// staging index of tap j of the tile's frame f (frame f starts at tile sample 160 f)
QW_HD int tap_index(int f, int j) { return kLanePitch * f + j + (j >> 5); }

// ---- forward 8-point DFT (e^{-2 pi i nk/8}), natural order in and out
QW_HD void dft8(float (&r)[8], float (&i)[8]) {
  const float h = 0.70710678118654752440f;
  float ar[4], ai[4], br[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ar[j] = r[j] + r[j + 4];
    ai[j] = i[j] + i[j + 4];
    br[j] = r[j] - r[j + 4];
    bi[j] = i[j] - i[j + 4];
  }
  // b_j *= W8^j : W8^1 = (1 - i) h, W8^2 = -i, W8^3 = (-1 - i) h
  {
    const float t1r = (br[1] + bi[1]) * h, t1i = (bi[1] - br[1]) * h;
    br[1] = t1r; bi[1] = t1i;
    const float t2r = bi[2], t2i = -br[2];
    br[2] = t2r; bi[2] = t2i;
    const float t3r = (bi[3] - br[3]) * h, t3i = -(br[3] + bi[3]) * h;
    br[3] = t3r; bi[3] = t3i;
  }
  // 4-point DFTs: a -> even outputs, b -> odd outputs
#define QW_DFT4(xr, xi, o0, o1, o2, o3)                                            \
  {                                                                                \
    const float c0r = xr[0] + xr[2], c0i = xi[0] + xi[2];                          \
    const float c1r = xr[1] + xr[3], c1i = xi[1] + xi[3];                          \
    const float d0r = xr[0] - xr[2], d0i = xi[0] - xi[2];                          \
    const float d1r = xi[1] - xi[3], d1i = -(xr[1] - xr[3]); /* (x1 - x3) * (-i) */ \
    r[o0] = c0r + c1r; i[o0] = c0i + c1i;                                          \
    r[o2] = c0r - c1r; i[o2] = c0i - c1i;                                          \
    r[o1] = d0r + d1r; i[o1] = d0i + d1i;                                          \
    r[o3] = d0r - d1r; i[o3] = d0i - d1i;                                          \
  }
  QW_DFT4(ar, ai, 0, 2, 4, 6)
  QW_DFT4(br, bi, 1, 3, 5, 7)
#undef QW_DFT4
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

// Col: float& at(int e) -- the frame's work column.  Aud: float tap(int j) -- windowless sample j of the frame.
template <typename Col, typename Aud>
QW_HD void pass_a(int g, int G, const Aud& aud, Col& col) {
  for (int n2 = g; n2 < 25; n2 += G) {
    float xr[8], xi[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int j = 2 * (25 * n1 + n2);
      xr[n1] = aud.tap(j) * QW_TBL(win, j);
      xi[n1] = aud.tap(j + 1) * QW_TBL(win, j + 1);
    }
    dft8(xr, xi);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
      float yr = xr[k1], yi = xi[k1];
      if (k1 > 0) {
        const float wr = QW_TBL(tw200_re, n2 * 8 + k1), wi = QW_TBL(tw200_im, n2 * 8 + k1);
        const float tr = yr * wr - yi * wi;
        yi = yr * wi + yi * wr;
        yr = tr;
      }
      col.at(2 * (n2 * 8 + k1)) = yr;
      col.at(2 * (n2 * 8 + k1) + 1) = yi;
    }
  }
}

template <typename Col>
QW_HD void pass_b(int g, int G, Col& col) {
  for (int k1 = g; k1 < 8; k1 += G) {
    float r[25], i[25];
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) {
      r[n2] = col.at(2 * (n2 * 8 + k1));
      i[n2] = col.at(2 * (n2 * 8 + k1) + 1);
    }
    dft25(r, i);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const int k2 = c + 5 * d;
        col.at(2 * (k2 * 8 + k1)) = r[5 * c + d];
        col.at(2 * (k2 * 8 + k1) + 1) = i[5 * c + d];
      }
    }
  }
}

// Z (200-point complex spectrum of the packed frame) -> power spectrum of the 400-point real FFT, in place.
template <typename Col>
QW_HD void untangle_power(int g, int G, Col& col) {
  for (int k = g; k <= 100; k += G) {
    if (k == 0) {
      const float zr = col.at(0), zi = col.at(1);
      const float a = zr + zi, b = zr - zi;  // X[0] = Zr + Zi, X[200] = Zr - Zi (both real)
      col.at(0) = a * a;
      col.at(1) = b * b;
    } else {
      const int s0 = 2 * k, s1 = 2 * (200 - k);
      const float ar = col.at(s0), ai = col.at(s0 + 1), br = col.at(s1), bi = col.at(s1 + 1);
      const float er = 0.5f * (ar + br), ei = 0.5f * (ai - bi);   // (Z[k] + conj Z[200-k]) / 2
      const float orr = 0.5f * (ai + bi), oi = -0.5f * (ar - br);  // (Z[k] - conj Z[200-k]) / (2i)
      const float wr = QW_TBL(un_re, k), wi = QW_TBL(un_im, k);
      const float tr = orr * wr - oi * wi, ti = orr * wi + oi * wr;
      const float pr = er + tr, pi = ei + ti;  // X[k]
      const float qr = er - tr, qi = ei - ti;  // conj X[200-k]
      col.at(s0) = pr * pr + pi * pi;
      if (k != 100) col.at(s1) = qr * qr + qi * qi;
    }
  }
}

}  // namespace lm
}  // namespace qw
