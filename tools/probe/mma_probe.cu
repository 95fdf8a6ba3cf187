// Microbenchmark: issue rate of the warp-level mma.sync m16n8k8 TF32 instruction on sm_100a (legacy tensor-core path, SASS HMMA.1688.F32.TF32)
// next to scalar FFMA, for 1..16 warps per SM sub-partition, and the rate of an FFMA stream that shares the sub-partition with an MMA stream.
// Decides whether the rank-4 contractions of the streaming kernels (post_conv, gout, grad post_conv.weight) can move off the FMA pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// NACC independent accumulator tiles per warp
template <int NACC>
__global__ void k_mma(float* out, unsigned seed) {
  float d[NACC][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + (threadIdx.x + i + seed) * 1e-3f);
  b[0] = __float_as_uint(0.5f + seed * 1e-3f);
  b[1] = __float_as_uint(0.25f);
#pragma unroll
  for (int n = 0; n < NACC; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int n = 0; n < NACC; ++n) mma_tf32(d[n], a, b);
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the planned inner loop: per MMA pair 4 LDS.32 + 4 LOP + 4 FADD (hi/lo split of the streamed operand) + 2 MMA
template <int NACC>
__global__ void k_mix(float* out, unsigned seed) {
  __shared__ float tile[32 * 32 * 4];
  for (int i = threadIdx.x; i < 32 * 32 * 4; i += blockDim.x) tile[i] = 1.0f + (i % 97) * 1e-3f;
  __syncthreads();
  float d[NACC][4];
  unsigned bh[2], bl[2];
  bh[0] = __float_as_uint(0.5f + seed * 1e-3f); bh[1] = __float_as_uint(0.25f);
  bl[0] = __float_as_uint(1e-4f); bl[1] = __float_as_uint(2e-4f);
#pragma unroll
  for (int n = 0; n < NACC; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
      for (int n = 0; n < NACC; ++n) {
        const float* p = tile + ((it + n * 8 + kb) & 3) * 1024 + kb * 4;
        const float v0 = p[g * 32 + t], v1 = p[(g + 8) * 32 + t], v2 = p[g * 32 + t + 32 * 16], v3 = p[(g + 8) * 32 + t + 32 * 16];
        unsigned ah[4], al[4];
        ah[0] = __float_as_uint(v0) & 0xffffe000u; ah[1] = __float_as_uint(v1) & 0xffffe000u;
        ah[2] = __float_as_uint(v2) & 0xffffe000u; ah[3] = __float_as_uint(v3) & 0xffffe000u;
        al[0] = __float_as_uint(v0 - __uint_as_float(ah[0])); al[1] = __float_as_uint(v1 - __uint_as_float(ah[1]));
        al[2] = __float_as_uint(v2 - __uint_as_float(ah[2])); al[3] = __float_as_uint(v3 - __uint_as_float(ah[3]));
        mma_tf32(d[n], ah, bh);
        mma_tf32(d[n], al, bl);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, float a, float b) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], a, b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the warps run MMAs, half run FFMAs
__global__ void k_both(float* out, float a, float b) {
  const int warp = threadIdx.x >> 5;
  if ((warp >> 2) & 1) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    float d[4][4];
    unsigned aa[4], bb[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) aa[i] = __float_as_uint(1.0f + (threadIdx.x + i) * 1e-3f);
    bb[0] = __float_as_uint(a); bb[1] = __float_as_uint(b);
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int n = 0; n < 4; ++n) mma_tf32(d[n], aa, bb);
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
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
  const double clk = khz * 1e3;
  float* out;
  cudaMalloc(&out, (size_t)sms * 2048 * 4);
  printf("SMs %d clock %.0f MHz\n", sms, khz / 1e3);
  for (int threads : {128, 256, 512, 1024}) {
    const int wps = threads / 32 / 4;  // warps per sub-partition (1 CTA per SM)
    {
      const double ms = time_ms([&] { k_mma<1><<<sms, threads>>>(out, 1); });
      const double n = (double)sms * (threads / 32) * ITERS * 1;
      printf("mma.m16n8k8.tf32  %2d warps/SMSP, 1 acc tile (dependent chain): %.3f ms  %.2f clk/MMA/SMSP  (latency-bound at 1 warp)\n", wps, ms,
             ms * 1e-3 * clk / (n / sms / 4));
    }
    {
      const double ms = time_ms([&] { k_mma<4><<<sms, threads>>>(out, 1); });
      const double n = (double)sms * (threads / 32) * ITERS * 4;
      printf("mma.m16n8k8.tf32  %2d warps/SMSP, 4 acc tiles: %.3f ms  %.2f clk/MMA/SMSP  %.1f dense TF32 TFLOP/s\n", wps, ms,
             ms * 1e-3 * clk / (n / sms / 4), n * 2048.0 / (ms * 1e-3) / 1e12);
    }
    {
      const double ms = time_ms([&] { k_mma<8><<<sms, threads>>>(out, 1); });
      const double n = (double)sms * (threads / 32) * ITERS * 8;
      printf("mma.m16n8k8.tf32  %2d warps/SMSP, 8 acc tiles: %.3f ms  %.2f clk/MMA/SMSP\n", wps, ms, ms * 1e-3 * clk / (n / sms / 4));
    }
    {
      const double ms = time_ms([&] { k_mix<4><<<sms, threads>>>(out, 1); });
      const double pairs = (double)sms * (threads / 32) * ITERS * 4;  // (4 LDS + 8 ALU + 2 MMA) groups
      printf("split loop (4 LDS + 8 ALU + 2 MMA per group) %2d warps/SMSP: %.3f ms  %.2f clk/group/SMSP  = %.2f clk per 14 instr\n", wps, ms,
             ms * 1e-3 * clk / (pairs / sms / 4), ms * 1e-3 * clk / (pairs / sms / 4));
    }
  }
  {
    const double ms = time_ms([&] { k_ffma<<<sms, 512>>>(out, 1.0001f, 1e-4f); });
    const double n = (double)sms * 16 * ITERS * 8;
    printf("FFMA alone         4 warps/SMSP: %.3f ms  %.2f clk/warp-FFMA/SMSP\n", ms, ms * 1e-3 * clk / (n / sms / 4));
  }
  {
    const double ms = time_ms([&] { k_both<<<sms, 1024>>>(out, 1.0001f, 1e-4f); });
    printf("MMA (4 warps/SMSP, 4 acc) + FFMA (4 warps/SMSP, 8 acc) together: %.3f ms (MMA alone and FFMA alone at the same warp counts above)\n", ms);
  }
  return 0;
}
