// Test-only: runs the log-mel kernel's per-frame arithmetic (qw_logmel_math.cuh, the SAME source the CUDA kernel
// compiles) on the host so the CPU suite can check the FFT factorisation, twiddle tables and plane / bin maps against
// numpy without a GPU.  Built by tests/test_logmel_math_cpu.py with g++; not part of libqw_b200.so.
#include "../../qasr_ijcnlp_b200/csrc/qw_logmel_math.cuh"

extern "C" {
// frames: (n, 400) raw (un-windowed) samples -> power: (n, 201).  `written` (n, 201) counts the stores per bin: the
// bin map of pass B must hit every bin exactly once.
void lm_host_power(const float* frames, float* power, int* written, int n) {
  using namespace qw::lm;
  for (int f = 0; f < n; ++f) {
    float work[kPlanes][25];
    for (int n2 = 0; n2 < 25; ++n2) {  // the device runs one n2 per lane
      float x[16], w[16], twr[8], twi[8];
      for (int n1 = 0; n1 < 16; ++n1) {
        x[n1] = frames[400 * f + 25 * n1 + n2];
        w[n1] = h_win[25 * n1 + n2];
      }
      for (int k1 = 1; k1 <= 8; ++k1) {
        twr[k1 - 1] = h_tw400_re[(k1 - 1) * 25 + n2];
        twi[k1 - 1] = h_tw400_im[(k1 - 1) * 25 + n2];
      }
      pass_a_one(x, w, twr, twi, [&](int pl, float v) { work[pl][n2] = v; });
    }
    for (int k1 = 0; k1 <= 8; ++k1) {  // the device runs one k1 per warp
      const float* re = work[k1 == 0 ? 0 : 2 * k1 - 1];
      const float* im = work[2 * k1];
      pass_b_one(
          k1, [&](int n2) { return re[n2]; }, [&](int n2) { return im[n2]; },
          [&](int k, float v) {
            power[201 * f + k] = v;
            written[201 * f + k] += 1;
          });
    }
  }
}
void lm_dft25(float* r, float* i) {
  float rr[25], ii[25];
  for (int k = 0; k < 25; ++k) { rr[k] = r[k]; ii[k] = i[k]; }
  qw::lm::dft25(rr, ii);
  for (int c = 0; c < 5; ++c)
    for (int d = 0; d < 5; ++d) { r[c + 5 * d] = rr[5 * c + d]; i[c + 5 * d] = ii[5 * c + d]; }
}
// x: 16 real samples -> (re, im)[9]
void lm_rfft16(const float* x, float* re, float* im) {
  float xx[16], ww[16], xr[9], xi[9];
  for (int k = 0; k < 16; ++k) { xx[k] = x[k]; ww[k] = 1.f; }
  xi[0] = xi[8] = 0.f;
  qw::lm::rfft16w(xx, ww, xr, xi);
  for (int k = 0; k < 9; ++k) { re[k] = xr[k]; im[k] = xi[k]; }
}
}
