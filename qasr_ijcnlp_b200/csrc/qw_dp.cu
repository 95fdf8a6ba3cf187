// One-shot all-reduce of the (small) trainable-gradient bucket over NVLink peer memory, fused with the 1/world scaling:
// the data-parallel collective of SURVEY.md 8e for the `freeze_non_quantum_layers` regime (9 440 floats for the two
// quantum layers + a task head).  NCCL stays the path for large buckets (qasr_ijcnlp_b200.dp.GradBucket).
//
// Every rank owns a SYMMETRIC buffer (2 slots x n floats) and a flag array (2 x world uint32 + epoch + status), both
// mapped into every peer's address space (torch symmetric memory does the rendezvous; the library only sees pointers).
// The bucket is cut into <= 64 chunks, one CTA per chunk, each with its own flags (no grid-wide barrier needed):
//   1. copy my chunk into my slot (epoch parity), __threadfence_system, store flag[slot][cta][my rank] = epoch into EVERY
//      peer's flag array (P2P stores);
//   2. spin until my own flag[slot][cta][r] >= epoch for every r (bounded: ~1 s, then status = 1);
//   3. out[i] = scale * sum_r peer_slot_r[i] in fixed rank order, all peer loads of a float4 in flight at once, L1 bypassed
//      -> bitwise identical on all ranks.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is CUDA-graph capturable.
// Double buffering by epoch parity is sufficient: a rank can only be one epoch ahead of the slowest reader.
#include "../../include/qw.h"
#include "qw_common.cuh"

namespace qw {
namespace dp {

constexpr int kMaxWorld = 8;
constexpr int kThreadsAR = 256;
constexpr int kMaxCtas = 64;  // the bucket is cut into <= 64 chunks, one CTA (and one set of flags) per chunk

// flag array (uint32): arrival[slot][cta][rank] at (slot * kMaxCtas + cta) * world + rank; epoch[cta] at 2 * kMaxCtas * world + cta;
// status at 2 * kMaxCtas * world + kMaxCtas (the last word)
__host__ __device__ inline int flag_words(int world) { return 2 * kMaxCtas * world + kMaxCtas + 1; }

struct ARArgs {
  float* grads;                 // (n) in / out, local
  float* bufs[kMaxWorld];       // peer r's symmetric data buffer (2 * n4 * 4 floats)
  unsigned* flags[kMaxWorld];   // peer r's flag array
  long long n, n4;              // elements; float4 groups (ceil)
  int rank, world, per_cta4;    // float4 groups per CTA
  float scale;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cv4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreadsAR) grads_allreduce_p2p_kernel(const ARArgs a) {
  __shared__ unsigned s_epoch;
  __shared__ int s_bad;
  const int tid = threadIdx.x, cta = blockIdx.x;
  unsigned* myflags = a.flags[a.rank];
  const int eidx = 2 * kMaxCtas * a.world + cta;
  if (tid == 0) {
    s_epoch = myflags[eidx] + 1u;
    s_bad = 0;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int slot = epoch & 1u;
  const long long g0 = (long long)cta * a.per_cta4;
  const long long g1 = (g0 + a.per_cta4 < a.n4) ? g0 + a.per_cta4 : a.n4;
  const size_t slot_off = (size_t)slot * a.n4 * 4;
  float* myslot = a.bufs[a.rank] + slot_off;
  // 1. publish my chunk (the slot is padded to whole float4 groups; the ragged tail of `grads` is read element-wise)
  for (long long g = g0 + tid; g < g1; g += kThreadsAR) {
    float4 v;
    if (4 * g + 3 < a.n) v = *reinterpret_cast<const float4*>(a.grads + 4 * g);
    else {
      v.x = 4 * g < a.n ? a.grads[4 * g] : 0.f;
      v.y = 4 * g + 1 < a.n ? a.grads[4 * g + 1] : 0.f;
      v.z = 4 * g + 2 < a.n ? a.grads[4 * g + 2] : 0.f;
      v.w = 0.f;
    }
    *reinterpret_cast<float4*>(myslot + 4 * g) = v;
  }
  __threadfence_system();
  __syncthreads();
  const int fbase = (slot * kMaxCtas + cta) * a.world;
  if (tid < a.world) {
    st_release_sys(a.flags[tid] + fbase + a.rank, epoch);
    // 2. wait for every rank's chunk (bounded: ~1 s, then give up instead of hanging the GPU)
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(myflags + fbase + tid) - epoch) < 0) {
      if (clock64() - t0 > 2000000000LL) {
        s_bad = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) myflags[2 * kMaxCtas * a.world + kMaxCtas] = 1u;
  } else {
    // 3. fixed-order sum over ranks with L1-bypassing peer loads: bitwise identical on every rank
    for (long long g = g0 + tid; g < g1; g += kThreadsAR) {
      float4 v[kMaxWorld];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < a.world) v[r] = ld_cv4(a.bufs[r] + slot_off + 4 * g);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < a.world) {
          s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w;
        }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
      if (4 * g + 3 < a.n) *reinterpret_cast<float4*>(a.grads + 4 * g) = s;
      else {
        if (4 * g < a.n) a.grads[4 * g] = s.x;
        if (4 * g + 1 < a.n) a.grads[4 * g + 1] = s.y;
        if (4 * g + 2 < a.n) a.grads[4 * g + 2] = s.z;
      }
    }
  }
  if (tid == 0) myflags[eidx] = epoch;
}

}  // namespace dp
}  // namespace qw

extern "C" {

size_t qw_grads_allreduce_p2p_buffer_bytes(long long n) { return n > 0 ? (size_t)2 * ((n + 3) / 4) * 4 * sizeof(float) : 0; }
size_t qw_grads_allreduce_p2p_flag_bytes(int world) { return world > 0 ? (size_t)qw::dp::flag_words(world) * sizeof(unsigned) : 0; }

int qw_grads_allreduce_p2p(float* grads, long long n, void* const* peer_bufs, void* const* peer_flags, int rank, int world,
                           float scale, void* stream) {
  using namespace qw;
  using namespace qw::dp;
  QW_CHECK_ARG(grads && peer_bufs && peer_flags && n > 0, -1, "qw_grads_allreduce_p2p: null pointer or empty bucket");
  QW_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, -1, "qw_grads_allreduce_p2p: bad rank/world %d/%d (world <= %d)",
               rank, world, kMaxWorld);
  QW_CHECK_ARG(n <= (1LL << 22), -2, "qw_grads_allreduce_p2p: bucket of %lld floats is too large for the one-shot kernel (use NCCL)", n);
  ARArgs a{};
  a.grads = grads;
  for (int r = 0; r < world; ++r) {
    QW_CHECK_ARG(peer_bufs[r] && peer_flags[r], -1, "qw_grads_allreduce_p2p: null peer pointer for rank %d", r);
    a.bufs[r] = (float*)peer_bufs[r];
    a.flags[r] = (unsigned*)peer_flags[r];
  }
  QW_CHECK_ARG(((uintptr_t)grads & 15) == 0, -1, "qw_grads_allreduce_p2p: the bucket must be 16-byte aligned");
  a.n = n;
  a.n4 = (n + 3) / 4;
  int ctas = (int)((a.n4 + kThreadsAR - 1) / kThreadsAR);
  ctas = ctas < 1 ? 1 : ctas > kMaxCtas ? kMaxCtas : ctas;
  a.per_cta4 = (int)((a.n4 + ctas - 1) / ctas);
  a.rank = rank;
  a.world = world;
  a.scale = scale;
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer kt(kKGradAllReduce, st);
    grads_allreduce_p2p_kernel<<<ctas, kThreadsAR, 0, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
