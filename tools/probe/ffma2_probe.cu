// Microbenchmark: issue rate of FFMA vs FFMA2 (packed fp32x2, sm_100a) vs FADD2, per SM per clock.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, ACC = 8;

__global__ void k_ffma(float* out, float a, float b) {
  float acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = fmaf(acc[i], a, b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float2 acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, -b);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float* out, float a, float b) {
  float2 acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 b2 = make_float2(b, -b);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = __fadd2_rn(acc[i], b2);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double run(K k, float* out, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<sms * 4, 512>>>(out, 1.0001f, 1e-4f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<sms * 4, 512>>>(out, 1.0001f, 1e-4f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms, khz;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, (size_t)sms * 4 * 512 * 4);
  const double insts = (double)sms * 4 * 512 * ITERS * ACC;  // thread-level instructions
  const double t1 = run(k_ffma, out, sms), t2 = run(k_ffma2, out, sms), t3 = run(k_fadd2, out, sms);
  const double clk = khz * 1e3;
  printf("SMs %d clock %.0f MHz\n", sms, khz / 1e3);
  printf("FFMA : %.3f ms  %.1f thread-instr/clk/SM  %.1f TFLOP/s\n", t1, insts / (t1 * 1e-3) / clk / sms, 2 * insts / (t1 * 1e-3) / 1e12);
  printf("FFMA2: %.3f ms  %.1f thread-instr/clk/SM  %.1f TFLOP/s\n", t2, insts / (t2 * 1e-3) / clk / sms, 4 * insts / (t2 * 1e-3) / 1e12);
  printf("FADD2: %.3f ms  %.1f thread-instr/clk/SM  %.1f TFLOP/s (adds)\n", t3, insts / (t3 * 1e-3) / clk / sms, 2 * insts / (t3 * 1e-3) / 1e12);
  return 0;
}
