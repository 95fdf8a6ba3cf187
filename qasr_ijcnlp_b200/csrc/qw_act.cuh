// Exact-erf GELU (torch.nn.functional.gelu, approximate='none') and its derivative for fusing the activation that follows each
// QuantumConv1d in the encoder stem (/root/reference/whisper/whisper/model.py:193-194: x = F.gelu(self.conv1(x)); x = F.gelu(self.conv2(x))).
#pragma once
#include <cuda_runtime.h>

namespace qw {

// gelu(v) = 0.5 v (1 + erf(v / sqrt 2)) = max(v, 0) - 0.5 |v| erfc(|v| / sqrt 2).
// The second form has no 1 - e cancellation, and erfc(u / sqrt 2) = 2^(-u Q(u)) with Q a degree-7 polynomial (weighted
// least-squares fit on [0, 6] of -log2(erfc(u / sqrt 2)) / u, weight u erfc: the error of the RESULT is what is minimised;
// Q stays > 4.8 beyond 6, so the tail underflows to the right limit).  11 instructions instead of erff's 24 (two coefficient
// sets + selects), max |error| 2.7e-7 against the fp64 GELU on [-40, 40] -- ATen's own fp32 gelu is 1.2e-6 off on that range.
// The factor 0.5 rides in the exponent (-u Q - 1) and the final subtraction is one fma.
__device__ __forceinline__ float half_erfc_over_sqrt2(float u) {  // 0.5 erfc(u / sqrt 2), u >= 0
  float q = 2.8103786462452263e-06f;
  q = fmaf(q, u, -3.908000508090481e-05f);
  q = fmaf(q, u, 0.00018477895355317742f);
  q = fmaf(q, u, 0.00014021758397575468f);
  q = fmaf(q, u, -0.007067482452839613f);
  q = fmaf(q, u, 0.05249877646565437f);
  q = fmaf(q, u, 0.4592074155807495f);
  q = fmaf(q, u, 1.151105284690857f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(-u, q, -1.f)));
  return e;
}
__device__ __forceinline__ float gelu_erf(float v) {
  const float u = fabsf(v);
  return fmaf(-u, half_erfc_over_sqrt2(u), fmaxf(v, 0.f));
}
// d/dv gelu(v) = Phi(v) + v phi(v), Phi(v) = v >= 0 ? 1 - e : e with e = 0.5 erfc(|v| / sqrt 2), phi(v) = exp(-v^2 / 2) / sqrt(2 pi).
// Max |error| 1.5e-7 against fp64 on [-12, 12] (checked in float32 arithmetic).
__device__ __forceinline__ float gelu_erf_grad(float v) {
  const float e = half_erfc_over_sqrt2(fabsf(v));
  const float Phi = v >= 0.f ? 1.f - e : e;
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(-0.7213475204444817f * v * v));
  return fmaf(v, 0.3989422804014327f * p, Phi);
}

}  // namespace qw
