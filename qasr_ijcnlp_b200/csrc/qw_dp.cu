// One-shot all-reduce of the (small) trainable-gradient bucket over NVLink peer memory, fused with the 1/world scaling:
// the data-parallel collective of SURVEY.md 8e for the `freeze_non_quantum_layers` regime (9 440 floats for the two
// quantum layers + a task head).  NCCL stays the path for large buckets (qasr_ijcnlp_b200.dp.GradBucket); a QuantumConv1d
// layer can also take the all-reduce into its own backward (qw_conv1d_backward_dp, qw_conv1d_fast.cu).
//
// Every rank owns a RECEIVE buffer of [2 slots][world][n] 64-bit words mapped into every peer's address space (torch symmetric
// memory does the rendezvous; the library only sees pointers) and a local bookkeeping array (epoch per CTA + status).
// The bucket is cut into <= 64 chunks, one CTA per chunk:
//   1. every thread packs (epoch << 32 | float bits) of its elements and stores the word straight into slot
//      [epoch parity][my rank][i] of EVERY peer's buffer (posted P2P stores; data and flag in ONE atomic 8-byte store, the idea
//      of NCCL's LL protocol: no fence, no separate flag, no remote read);
//   2. it polls its own buffer until the word of every rank carries the epoch (NCCL semantics: it WAITS for a late peer; a peer
//      that stays away for the DP_TIMEOUT_MS option -- default 10 min -- makes the kernel trap: a loud launch failure, never a
//      silently un-averaged gradient) and sums the payloads in fixed rank order -> bitwise identical on all ranks;
//      out[i] = scale * sum.
// First version (publish into my own buffer, __threadfence_system, st.release.sys flags into the peers, acquire-spin, remote
// loads): each system-scope fence cost 3-6 us on B200 and the whole exchange ~19 us; this one is one NVLink write latency.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is CUDA-graph capturable.
// Double buffering by epoch parity is sufficient: a rank can only be one epoch ahead of the slowest reader.
#include "../../include/qw.h"
#include "qw_common.cuh"

namespace qw {
namespace dp {

constexpr int kMaxWorld = 8;
constexpr int kThreadsAR = 256;
constexpr int kMaxCtas = 64;  // the bucket is cut into <= 64 chunks, one CTA (and one epoch word) per chunk

// bookkeeping words (uint32, local): epoch[cta] at cta; status at kMaxCtas (the last word)
__host__ __device__ inline int flag_words(int) { return kMaxCtas + 1; }

struct ARArgs {
  float* grads;                        // (n) in / out, local
  unsigned long long* bufs[kMaxWorld]; // peer r's receive buffer: [2][world][n] words
  unsigned* book;                      // my bookkeeping words
  long long n;
  int rank, world, per_cta;            // elements per CTA
  float scale;
  unsigned long long timeout_ns;       // 0 = wait forever
};

__device__ __forceinline__ void st_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreadsAR) grads_allreduce_p2p_kernel(const ARArgs a) {
  __shared__ unsigned s_epoch;
  const int tid = threadIdx.x, cta = blockIdx.x;
  if (tid == 0) s_epoch = a.book[cta] + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  const size_t slot_base = (size_t)(epoch & 1u) * a.world;
  const long long i0 = (long long)cta * a.per_cta;
  const long long i1 = (i0 + a.per_cta < a.n) ? i0 + a.per_cta : a.n;
  // 1. push my elements into every peer's receive buffer
  for (long long i = i0 + tid; i < i1; i += kThreadsAR) {
    const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(a.grads[i]);
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < a.world) st_u64(a.bufs[r] + (slot_base + a.rank) * a.n + i, word);
  }
  // 2. gather: poll my own buffer, fixed rank order.  NCCL semantics: wait for a late peer; only after a.timeout_ns (default
  // 10 min, 0 = forever) trap -- the context dies loudly instead of training on with an un-averaged gradient.
  const unsigned long long* mine = a.bufs[a.rank];
  const unsigned long long t0 = global_ns();
  for (long long i = i0 + tid; i < i1; i += kThreadsAR) {
    float s = 0.f;
    for (int r = 0; r < a.world; ++r) {
      const unsigned long long* src = mine + (slot_base + r) * a.n + i;
      unsigned long long v = ld_u64(src);
      unsigned spins = 0;
      while ((unsigned)(v >> 32) != epoch) {
        if ((++spins & 1023u) == 0 && a.timeout_ns != 0 && global_ns() - t0 > a.timeout_ns) {
          a.book[kMaxCtas] = 1u;
          __threadfence_system();
          __trap();
        }
        v = ld_u64(src);
      }
      s += __uint_as_float((unsigned)v);
    }
    a.grads[i] = s * a.scale;
  }
  __syncthreads();
  if (tid == 0) a.book[cta] = epoch;
}

}  // namespace dp
}  // namespace qw

extern "C" {

size_t qw_grads_allreduce_p2p_buffer_bytes(long long n, int world) {
  return (n > 0 && world > 0) ? (size_t)2 * world * n * sizeof(unsigned long long) : 0;
}
size_t qw_grads_allreduce_p2p_flag_bytes(int world) { return world > 0 ? (size_t)qw::dp::flag_words(world) * sizeof(unsigned) : 0; }

int qw_grads_allreduce_p2p(float* grads, long long n, void* const* peer_bufs, void* const* peer_flags, int rank, int world,
                           float scale, void* stream) {
  using namespace qw;
  using namespace qw::dp;
  QW_CHECK_ARG(grads && peer_bufs && peer_flags && n > 0, -1, "qw_grads_allreduce_p2p: null pointer or empty bucket");
  QW_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, -1, "qw_grads_allreduce_p2p: bad rank/world %d/%d (world <= %d)",
               rank, world, kMaxWorld);
  QW_CHECK_ARG(n <= (1LL << 20), -2, "qw_grads_allreduce_p2p: bucket of %lld floats is too large for the one-shot kernel (use NCCL)", n);
  ARArgs a{};
  a.grads = grads;
  for (int r = 0; r < world; ++r) {
    QW_CHECK_ARG(peer_bufs[r], -1, "qw_grads_allreduce_p2p: null peer pointer for rank %d", r);
    a.bufs[r] = (unsigned long long*)peer_bufs[r];
  }
  QW_CHECK_ARG(peer_flags[rank], -1, "qw_grads_allreduce_p2p: null bookkeeping pointer");
  a.book = (unsigned*)peer_flags[rank];
  a.n = n;
  int ctas = (int)((n + 4 * kThreadsAR - 1) / (4 * kThreadsAR));
  ctas = ctas < 1 ? 1 : ctas > kMaxCtas ? kMaxCtas : ctas;
  a.per_cta = (int)((n + ctas - 1) / ctas);
  a.rank = rank;
  a.world = world;
  a.scale = scale;
  a.timeout_ns = (unsigned long long)(option(kOptDpTimeoutMs) > 0 ? option(kOptDpTimeoutMs) : 0) * 1000000ull;
  cudaStream_t st = (cudaStream_t)stream;
  {
    KernelTimer kt(kKGradAllReduce, st);
    grads_allreduce_p2p_kernel<<<ctas, kThreadsAR, 0, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
