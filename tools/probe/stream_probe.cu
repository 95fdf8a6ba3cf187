// Microbenchmark: achievable HBM read bandwidth of the TMA access patterns the streaming kernels use, with NO compute:
// a persistent grid pulls a (B, R, L) fp32 tensor through a shared-memory ring of SWIZZLE_128B boxes and only waits/releases.
// Answers: is the gy pass (192 rows x 32 windows per stage, 128-byte row segments) limited by the memory system or by the kernel?
//   stream_probe B R L  box_rows vert side stages ctas_per_sm [l2promo 0|1|2|3]
//     stage = vert x side boxes of (box_rows x 32 floats); tile = all R rows x (32*side) windows
// nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Args {
  int box_rows, vert, side, stages, tiles_per_utt, num_tiles, stages_per_tile;
  float* sink;
};

__global__ void __launch_bounds__(192) k_stream(const __grid_constant__ CUtensorMap tm, const Args a) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (s32(sm) & 1023u)) & 1023u);
  const int box_elems = a.box_rows * 32;
  const int stage_elems = box_elems * a.vert * a.side;
  float* ring = (float*)base;
  uint64_t* full = (uint64_t*)(ring + (size_t)a.stages * stage_elems);
  uint64_t* empty = full + a.stages;
  const int tid = threadIdx.x, lane = tid & 31, nwarps = blockDim.x >> 5;
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(nwarps) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_tiles * a.stages_per_tile;
  auto issue = [&](int gs) {
    const int n = gs / a.stages_per_tile, h = gs - n * a.stages_per_tile;
    const int tile = blockIdx.x + n * gridDim.x;
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * 32 * a.side;
    const int s = gs % a.stages;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_elems * 4) : "memory");
    for (int v = 0; v < a.vert; ++v)
      for (int sd = 0; sd < a.side; ++sd)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         s32(ring + (size_t)s * stage_elems + (v * a.side + sd) * box_elems)),
                     "l"(&tm), "r"(i0 + 32 * sd), "r"((h * a.vert + v) * a.box_rows), "r"(b), "r"(s32(&full[s]))
                     : "memory");
  };
  auto wait = [&](uint64_t* bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
  };
  if (tid == 0)
    for (int gs = 0; gs < a.stages - 1 && gs < total; ++gs) issue(gs);
  float acc = 0.f;
  for (int gs = 0; gs < total; ++gs) {
    if (tid == 0) {
      const int gn = gs + a.stages - 1;
      if (gn < total) {
        if (gn >= a.stages) wait(&empty[gn % a.stages], ((gn / a.stages) - 1) & 1);
        issue(gn);
      }
    }
    const int s = gs % a.stages;
    wait(&full[s], (gs / a.stages) & 1);
    acc += ring[(size_t)s * stage_elems + tid];  // touch the stage
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
  }
  if (acc == 123.456f) a.sink[tid] = acc;
}

int main(int argc, char** argv) {
  if (argc < 9) {
    printf("usage: stream_probe B R L box_rows vert side stages ctas_per_sm [l2promo]\n");
    return 1;
  }
  const int B = atoi(argv[1]), R = atoi(argv[2]), L = atoi(argv[3]), box_rows = atoi(argv[4]), vert = atoi(argv[5]), side = atoi(argv[6]),
            stages = atoi(argv[7]), cps = atoi(argv[8]), promo = argc > 9 ? atoi(argv[9]) : 2;
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t n = (size_t)B * R * L;
  const int NB = 4;
  float* buf[NB];
  for (int i = 0; i < NB; ++i) {
    cudaMalloc(&buf[i], n * 4);
    cudaMemset(buf[i], 0, n * 4);
  }
  float* sink;
  cudaMalloc(&sink, 4096);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncFn enc = (EncFn)p;
  CUtensorMap tm[NB];
  for (int i = 0; i < NB; ++i) {
    cuuint64_t gd[3] = {(cuuint64_t)L, (cuuint64_t)R, (cuuint64_t)B};
    cuuint64_t gs[2] = {(cuuint64_t)L * 4, (cuuint64_t)L * R * 4};
    cuuint32_t bx[3] = {32, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf[i], gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      return 1;
    }
  }
  Args a{};
  a.box_rows = box_rows; a.vert = vert; a.side = side; a.stages = stages;
  a.tiles_per_utt = (L + 32 * side - 1) / (32 * side);
  a.num_tiles = B * a.tiles_per_utt;
  a.stages_per_tile = (R + box_rows * vert - 1) / (box_rows * vert);
  a.sink = sink;
  const size_t smem = 1024 + (size_t)stages * box_rows * 32 * vert * side * 4 + 2 * stages * 8;
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = a.num_tiles < cps * sms ? a.num_tiles : cps * sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 8; ++i) k_stream<<<grid, 192, smem>>>(tm[i % NB], a);
  cudaDeviceSynchronize();
  float best = 1e9f, sum = 0.f;
  const int reps = 40;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0);
    k_stream<<<grid, 192, smem>>>(tm[i % NB], a);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
    sum += ms;
  }
  // 16 launches back to back (queued: launch overhead hidden, the ramp/tail of each kernel is not)
  cudaEventRecord(e0);
  for (int i = 0; i < 16; ++i) k_stream<<<grid, 192, smem>>>(tm[i % NB], a);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms16;
  cudaEventElapsedTime(&ms16, e0, e1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("B%d R%d L%d box %dx32 stage %dx%d boxes (%.0f KB) x%d stages, %d CTAs/SM (grid %d, %.2f tiles/CTA), promo %d: mean %.2f us (%.0f GB/s)  best %.2f us (%.0f GB/s)  x16 %.2f us each (%.0f GB/s)  smem %zu  %s\n",
         B, R, L, box_rows, vert, side, box_rows * 32 * vert * side * 4 / 1024.0, stages, cps, grid, (double)a.num_tiles / grid, promo, sum / reps * 1e3,
         n * 4 / (sum / reps * 1e-3) / 1e9, best * 1e3, n * 4 / (best * 1e-3) / 1e9, ms16 / 16 * 1e3, n * 4 / (ms16 / 16 * 1e-3) / 1e9, smem, cudaGetErrorString(e));
  return 0;
}
