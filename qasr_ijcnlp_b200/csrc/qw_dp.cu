// One-shot all-reduce of the (small) trainable-gradient bucket over NVLink peer memory, fused with the 1/world scaling:
// the data-parallel collective of SURVEY.md 8e for the `freeze_non_quantum_layers` regime (9 440 floats for the two
// quantum layers + a task head).  NCCL stays the path for large buckets (qasr_ijcnlp_b200.dp.GradBucket).
//
// Every rank owns a SYMMETRIC buffer (2 slots x n floats) and a flag array (2 x world uint32 + epoch + status), both
// mapped into every peer's address space (torch symmetric memory does the rendezvous; the library only sees pointers).
// One CTA per rank:
//   1. copy my gradients into my slot (epoch parity), __threadfence_system, store flag[slot][my rank] = epoch into EVERY
//      peer's flag array (P2P stores);
//   2. spin until my own flag[slot][r] >= epoch for every r (bounded: ~1 s, then status = 1);
//   3. out[i] = scale * sum_r peer_slot_r[i] in fixed rank order with L1-bypassing loads -> bitwise identical on all ranks.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is CUDA-graph capturable.
// Double buffering by epoch parity is sufficient: a rank can only be one epoch ahead of the slowest reader.
#include "../../include/qw.h"
#include "qw_common.cuh"

namespace qw {
namespace dp {

constexpr int kMaxWorld = 8;
constexpr int kThreadsAR = 1024;

struct ARArgs {
  float* grads;                 // (n) in / out, local
  float* bufs[kMaxWorld];       // peer r's symmetric data buffer (2 * n floats)
  unsigned* flags[kMaxWorld];   // peer r's flag array: [2][world] arrival flags, [2*world] = epoch, [2*world + 1] = status
  long long n;
  int rank, world;
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

__global__ void __launch_bounds__(kThreadsAR) grads_allreduce_p2p_kernel(const ARArgs a) {
  __shared__ unsigned s_epoch;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  unsigned* myflags = a.flags[a.rank];
  if (tid == 0) {
    s_epoch = myflags[2 * a.world] + 1u;
    s_bad = 0;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int slot = epoch & 1u;
  float* myslot = a.bufs[a.rank] + (size_t)slot * a.n;
  for (long long i = tid; i < a.n; i += kThreadsAR) myslot[i] = a.grads[i];
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) st_release_sys(a.flags[tid] + slot * a.world + a.rank, epoch);
  if (tid < a.world) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(myflags + slot * a.world + tid) - epoch) < 0) {
      if (clock64() - t0 > 2000000000LL) {  // ~1 s: a peer never arrived; do not hang the GPU
        s_bad = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) myflags[2 * a.world + 1] = 1u;
  } else {
    for (long long i = tid; i < a.n; i += kThreadsAR) {
      float s = 0.f;
      for (int r = 0; r < a.world; ++r) s += __ldcv(a.bufs[r] + (size_t)slot * a.n + i);
      a.grads[i] = s * a.scale;
    }
  }
  if (tid == 0) myflags[2 * a.world] = epoch;
}

}  // namespace dp
}  // namespace qw

extern "C" {

size_t qw_grads_allreduce_p2p_buffer_bytes(long long n) { return n > 0 ? (size_t)2 * n * sizeof(float) : 0; }
size_t qw_grads_allreduce_p2p_flag_bytes(int world) { return world > 0 ? (size_t)(2 * world + 2) * sizeof(unsigned) : 0; }

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
  a.n = n;
  a.rank = rank;
  a.world = world;
  a.scale = scale;
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer kt(kKGradAllReduce, st);
    grads_allreduce_p2p_kernel<<<1, kThreadsAR, 0, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
