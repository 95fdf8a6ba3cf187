// Test-only: runs the log-mel kernel's per-frame arithmetic (qw_logmel_math.cuh, the SAME source the CUDA kernel
// compiles) on the host so the CPU suite can check the FFT factorisation, twiddle tables and slot maps against
// numpy without a GPU.  Built by tests/test_logmel_math_cpu.py with g++; not part of libqw_b200.so.
#include "../../qasr_ijcnlp_b200/csrc/qw_logmel_math.cuh"

namespace {
struct Col {
  float v[400];
  float& at(int e) { return v[e]; }
};
struct Aud {
  const float* p;
  float tap(int j) const { return p[j]; }
};
}  // namespace

extern "C" {
// frames: (n, 400) raw (un-windowed) samples -> power: (n, 201); G emulates the CTA's warp split (results must not depend on it)
void lm_host_power(const float* frames, float* power, int n, int G) {
  for (int f = 0; f < n; ++f) {
    Col col;
    Aud aud{frames + 400 * f};
    for (int g = 0; g < G; ++g) qw::lm::pass_a(g, G, aud, col);
    for (int g = 0; g < G; ++g) qw::lm::pass_b(g, G, col);
    for (int g = 0; g < G; ++g) qw::lm::untangle_power(g, G, col);
    for (int k = 0; k <= 200; ++k) power[201 * f + k] = col.v[qw::lm::pslot(k)];
  }
}
int lm_tap_index(int f, int j) { return qw::lm::tap_index(f, j); }
int lm_skew(int s) { return qw::lm::skew(s); }
void lm_dft25(float* r, float* i) {
  float rr[25], ii[25];
  for (int k = 0; k < 25; ++k) { rr[k] = r[k]; ii[k] = i[k]; }
  qw::lm::dft25(rr, ii);
  for (int c = 0; c < 5; ++c)
    for (int d = 0; d < 5; ++d) { r[c + 5 * d] = rr[5 * c + d]; i[c + 5 * d] = ii[5 * c + d]; }
}
void lm_dft8(float* r, float* i) {
  float rr[8], ii[8];
  for (int k = 0; k < 8; ++k) { rr[k] = r[k]; ii[k] = i[k]; }
  qw::lm::dft8(rr, ii);
  for (int k = 0; k < 8; ++k) { r[k] = rr[k]; i[k] = ii[k]; }
}
}
