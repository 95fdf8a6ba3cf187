// Standalone probe: 3-D tensor-map TMA load (no swizzle and SWIZZLE_128B) + TMA store.
#include <cstdio>
#include <vector>
#include "../../qasr_ijcnlp_b200/csrc/qw_tma.cuh"
namespace qw {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
void count_launch(int) {}
bool profiling_enabled() { return false; }
void profile_begin(int, cudaStream_t) {}
void profile_end(int, cudaStream_t) {}
TmapEncodeFn tmap_encode_fn() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (TmapEncodeFn)p;
}
int make_tmap_3d_f32(CUtensorMap* tm, const void* base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                     unsigned box0, unsigned box1, bool swz) {
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstr[2] = {d0 * 4ull, d0 * d1 * 4ull};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = tmap_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); return -1; }
  return 0;
}
}
using namespace qw;
template <int BOX0, int BOX1, bool SWZ, int STEP>
__global__ void probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tms, float* out, int c0, int c1, int c2) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~uintptr_t(1023));
  float* tile = (float*)base;
  uint64_t* bar = (uint64_t*)(base + BOX0 * BOX1 * 4);
  if (threadIdx.x == 0) {
    if (STEP >= 1) { mbar_init(bar, 1); fence_mbar_init(); }
    if (STEP >= 2) tma_prefetch_desc(&tm);
  }
  __syncthreads();
  if (STEP >= 3) {
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, BOX0 * BOX1 * 4);
      tma_load_3d(tile, &tm, c0, c1, c2, bar);
    }
    mbar_wait(bar, 0);
  }
  for (int i = threadIdx.x; i < BOX0 * BOX1; i += blockDim.x) out[i] = (STEP >= 3) ? tile[i] : 1.f;
  if (STEP >= 4) {
    __syncthreads();
    for (int i = threadIdx.x; i < BOX0 * BOX1; i += blockDim.x) tile[i] += 1000.f;
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { tma_store_3d(&tms, c0, c1, c2, tile); bulk_commit(); bulk_wait_all<0>(); }
  }
}
template <int BOX0, int BOX1, bool SWZ, int STEP>
int run(const char* name, float* dsrc, float* ddst, float* dout, int D0, int D1, int D2, int c0, int c1, int c2, const std::vector<float>& h) {
  alignas(64) CUtensorMap tm, tms;
  if (make_tmap_3d_f32(&tm, dsrc, D0, D1, D2, BOX0, BOX1, SWZ)) return 1;
  if (make_tmap_3d_f32(&tms, ddst, D0, D1, D2, BOX0, BOX1, SWZ)) return 1;
  size_t smem = 1024 + BOX0 * BOX1 * 4 + 64;
  cudaFuncSetAttribute(probe<BOX0, BOX1, SWZ, STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(ddst, 0, (size_t)D0 * D1 * D2 * 4);
  probe<BOX0, BOX1, SWZ, STEP><<<1, 128, smem>>>(tm, tms, dout, c0, c1, c2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s step %d: %s\n", name, STEP, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  if (STEP >= 3) {
    std::vector<float> o(BOX0 * BOX1);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < BOX1; ++r) for (int c = 0; c < BOX0; ++c) {
      int gc = c0 + c, gr = c1 + r;
      float want = (gc >= 0 && gc < D0 && gr >= 0 && gr < D1) ? h[((size_t)c2 * D1 + gr) * D0 + gc] : 0.f;
      int idx = SWZ ? (r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3))) : (r * BOX0 + c);
      if (o[idx] != want) { if (bad < 5) printf("   mismatch r=%d c=%d got %f want %f\n", r, c, o[idx], want); ++bad; }
    }
    printf("   load mismatches: %d\n", bad);
  }
  if (STEP >= 4) {
    std::vector<float> d((size_t)D0 * D1 * D2);
    cudaMemcpy(d.data(), ddst, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < D1; ++r) for (int c = 0; c < D0; ++c) {
      bool in = (c >= c0 && c < c0 + BOX0 && r >= c1 && r < c1 + BOX1);
      float want = in ? h[((size_t)c2 * D1 + r) * D0 + c] + 1000.f : 0.f;
      if (d[((size_t)c2 * D1 + r) * D0 + c] != want) ++bad;
    }
    printf("   store mismatches: %d\n", bad);
  }
  return 0;
}
int main() {
  const int D0 = 200, D1 = 80, D2 = 2;
  std::vector<float> h((size_t)D0 * D1 * D2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) * 0.25f;
  float *dsrc, *ddst, *dout;
  cudaMalloc(&dsrc, h.size() * 4); cudaMalloc(&ddst, h.size() * 4); cudaMalloc(&dout, 65536 * 4);
  cudaMemcpy(dsrc, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  if (run<36, 16, false, 1>("noswz36x16", dsrc, ddst, dout, D0, D1, D2, -1, 64, 1, h)) return 1;
  if (run<36, 16, false, 2>("noswz36x16", dsrc, ddst, dout, D0, D1, D2, -1, 64, 1, h)) return 1;
  if (run<36, 16, false, 3>("noswz36x16", dsrc, ddst, dout, D0, D1, D2, -1, 64, 1, h)) return 1;
  if (run<36, 16, false, 3>("noswz36x16 right edge", dsrc, ddst, dout, D0, D1, D2, 191, 72, 0, h)) return 1;
  if (run<32, 32, true, 3>("swz32x32", dsrc, ddst, dout, D0, D1, D2, 32, 64, 1, h)) return 1;
  if (run<32, 32, true, 4>("swz32x32 store", dsrc, ddst, dout, D0, D1, D2, 192, 64, 1, h)) return 1;
  if (run<32, 64, true, 3>("swz32x64", dsrc, ddst, dout, D0, D1, D2, 160, 0, 0, h)) return 1;
  printf("probe done\n");
  return 0;
}
