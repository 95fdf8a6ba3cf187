// Shared host/device helpers for libqw_b200.so
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace qw {

// ---- kernel-selection options (A/B switches).  Process-global, read from the environment (QW_<NAME>) on first use and settable through
// qw_set_option(): TEST / EXPERIMENT ONLY -- they pick between equivalent kernels, are not part of the reference-facing contract, and
// changing them while another thread is inside a qw_* call is not supported.
enum Option {
  kOptFastPath = 0,  // FAST_PATH  1: TMA fast path when the shape qualifies; 0: generic kernels
  kOptGyMma,         // GY_MMA     1: gy pass on the tensor pipe (mma.sync 3xTF32); 0: FFMA form
  kOptBwdFused,      // BWD_FUSED  0 (default): gy / adjoint / pre_conv^T kernels; 1: adjoint (+ pre_conv^T of a data layer) inside the gy kernel; 2: adjoint only
  kOptFwdEtma,       // FWD_ETMA   1: forward requests its first x tiles before staging parameters
  kOptFinEarly,      // FIN_EARLY  1: finalize segments 1-2 run before the dependency wait (split backward only)
  kOptAdjTrig,       // ADJ_TRIG   1: adjoint kernel triggers its dependent right after its wait
  kOptPreEx,         // PRE_EX     1: pre_conv^T requests its first x tiles before the dependency wait
  kOptAdjSpec,       // ADJ_SPEC   1: single-layer specialisation of the adjoint kernel
  kOptPreCtas,       // PRE_CTAS   CTAs per SM of the pre_conv^T kernel (default 4)
  kOptGyWarps,       // GY_WARPS   12: wide FFMA gy kernel (only with GY_MMA=0)
  kOptDpTimeoutMs,   // DP_TIMEOUT_MS  how long the NVLink gradient all-reduce kernels wait for a peer before trapping (default 600 000; 0 = forever)
  kOptDbgFwd,        // DBG_FWD    bit mask: forward kernel skips pre_conv FMAs (1) / post_conv FMAs (2) / circuit (4); results are garbage
  kOptDbgGy,         // DBG_GY     1: gy kernel skips its contractions (measures the streaming floor; results are garbage)
  kOptStemChain,     // STEM_CHAIN 1 (default): the stem-level backward (Python _StemTrainFn, bench) chains conv1's gy pass to conv2's gpre rows; 0: grad_x through HBM
  kOptRevTiles,      // REV_TILES  bit 0: forward kernels, bit 1: gy kernel walk their tiles from the last to the first (L2 reuse of the producer's newest lines)
  kOptCount
};
int option(Option o);
int set_option(const char* name, int value);  // 0, or -1 for an unknown name

// ---- error text (thread local) and launch counter
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// Optional per-kernel CUDA-event timing (qw_profile_enable): KernelTimer brackets one launch on `st`.
enum KernelId { kKFwd = 0, kKBwdPost, kKBwdPre, kKBwdFinalize, kKCircFwd, kKCircBwd, kKCircFinalize, kKLogMelStft, kKLogMelFinish, kKBwdAdj, kKLogMelPrep, kKGradAllReduce, kKStem2, kKBwdFused, kKCount };
// the demangled name ncu shows for the kernel last launched under this id (fast path only; "" otherwise)
void note_symbol(int id, const char* fmt, ...);
bool profiling_enabled();
void profile_begin(int id, cudaStream_t st);
void profile_end(int id, cudaStream_t st);
struct KernelTimer {
  int id;
  cudaStream_t st;
  bool on;
  KernelTimer(int id_, cudaStream_t st_) : id(id_), st(st_), on(profiling_enabled()) {
    if (on) profile_begin(id, st);
  }
  ~KernelTimer() {
    count_launch();
    if (on) profile_end(id, st);
  }
};

#define QW_CHECK_ARG(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::qw::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)

#define QW_CUDA_OK(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::qw::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));          \
      return (int)_e;                                                           \
    }                                                                           \
  } while (0)

__host__ __device__ inline int floor_div(int a, int b) {
  int q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- small vector load/store of Q contiguous elements (16-byte aligned when Q*sizeof(T) % 16 == 0)
template <typename T, int Q>
__device__ __forceinline__ void ld_vec(const T* __restrict__ p, T (&v)[Q]) {
  if constexpr (sizeof(T) == 4 && Q == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (sizeof(T) == 4 && Q == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (sizeof(T) == 8 && Q == 4) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else if constexpr (sizeof(T) == 8 && Q == 2) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    v[0] = a.x; v[1] = a.y;
  } else {
#pragma unroll
    for (int j = 0; j < Q; ++j) v[j] = p[j];
  }
}
template <typename T, int Q>
__device__ __forceinline__ void st_vec(T* __restrict__ p, const T (&v)[Q]) {
  if constexpr (sizeof(T) == 4 && Q == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (sizeof(T) == 4 && Q == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else if constexpr (sizeof(T) == 8 && Q == 4) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
  } else if constexpr (sizeof(T) == 8 && Q == 2) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  } else {
#pragma unroll
    for (int j = 0; j < Q; ++j) p[j] = v[j];
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining; everything it does before pdl_wait() (parameter staging into shared memory, barrier init,
// tensor-map prefetch) overlaps the predecessor's tail.  Rules kept by every kernel that uses it: (1) nothing produced by an
// earlier kernel is read, and no global memory is written, before pdl_wait(); (2) pdl_launch() comes right after it, so at
// most two kernels of the chain ever overlap.  QW_PDL=0 turns the launch attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// QW_PDL: 0 = never, 2 = always, unset / 1 = when the caller says the launch is latency-dominated (`small`): measured on B200,
// PDL gains 4 % on the batch-16 stem step (5 tiles per CTA), 2.8 % at batch 64 (20 tiles) and loses 2.3 % at batch 256 (80 tiles).
int pdl_mode();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool small, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_mode() == 2 || (pdl_mode() == 1 && small)) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- debug timeline (qw_timeline_set): when a device buffer is registered, every fast-path launch takes the next slot and its
// CTAs record min(start) / max(end) of %globaltimer there, so the spans and gaps of the kernels INSIDE a replayed CUDA graph
// (which events cannot see) can be read back.  Null pointer (the default): one predictable branch per CTA.
unsigned long long* timeline_next_slot();
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* s) {
  if (s && threadIdx.x == 0) atomicMin(s, global_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* s) {
  if (s && threadIdx.x == 0) atomicMax(s + 1, global_ns());
}

// streaming (evict-first) global store / load for data touched exactly once
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

}  // namespace qw
