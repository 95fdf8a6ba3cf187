// Fused log-mel front end: replaces whisper.log_mel_spectrogram (/root/reference/whisper/whisper/audio.py:110-157),
// batched, with the max of :155 taken per utterance (the reference always passes one utterance).
//
//   logmel_stft_kernel   reflect-padded framing (bit-exact torch.stft(center=True) frame map) -> Hann -> 400-point
//                        real FFT -> power -> mel filterbank -> log10(max(., 1e-10)) -> mel (B, n_mels, T) + per-
//                        utterance running max (one atomic per warp per tile).
//   logmel_finish_kernel max(., max_b - 8) ; (. + 4) / 4, in place (the tensor was just written: L2-resident).
//
// Layout: one STFT frame per LANE, a tile = 32 consecutive frames of one utterance per CTA iteration, the G warps
// of the CTA split the butterflies / mel rows of those 32 frames.  Shared memory per CTA: the tile's audio span
// (34 hops at pitch 161 floats -> conflict-free frame-strided reads) + a 400 x 32 float work array indexed
// [slot][lane] (every access is one 128-byte row) = 73 KB -> 3 CTAs per SM.  HBM traffic is the algorithmic
// minimum: audio read once (hop overlap is served from shared memory), mel written once by 128-byte rows.
#include <cuda_runtime.h>

#include "../../include/qw.h"
#include "qw_common.cuh"
#include "qw_logmel_math.cuh"

namespace qw {
namespace lm {

constexpr int kFrames = 32;                                   // frames per tile == lanes
constexpr int kSpan = (kFrames - 1) * kHop + kNfft;           // 5360 samples feed one tile
constexpr int kHops = (kSpan + kHop - 1) / kHop;              // 34
constexpr int kAudFloats = kHops * kHopPitch;                 // 5474
constexpr int kWorkFloats = kNfft * kFrames;                  // 12800
constexpr int kMaxMels = 256;

struct Args {
  const float* audio;    // (B, n)
  const float* filters;  // (n_mels, 201)
  float* mel;            // (B, n_mels, T)
  float* umax;           // (B) running max, initialised to 0xffffffff ("-inf" for the mixed int/uint atomics)
  int B, n, T, n_mels, tiles_per_utt, num_tiles;
};

struct SmemCol {
  float* base;  // &work[lane]
  __device__ __forceinline__ float& at(int e) { return base[e * kFrames]; }
};
struct SmemAud {
  const float* base;  // &aud[lane * 161]: hop h of the tile starts at aud[h * 161]
  __device__ __forceinline__ float tap(int j) const { return base[(j / kHop) * kHopPitch + (j % kHop)]; }
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <int G>
__global__ void __launch_bounds__(32 * G) logmel_stft_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  float* aud = smem;
  float* work = smem + kAudFloats + 2;  // +2 keeps `work` 16-byte aligned (5476 floats)
  int* mlo = reinterpret_cast<int*>(work + kWorkFloats);
  int* mhi = mlo + kMaxMels;
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;

  // support [lo, hi) of every mel row (first / last non-zero), found once per CTA: warp per row, ballot scan
  for (int m = g; m < a.n_mels; m += G) {
    int lo = kNfreq, hi = 0;
    for (int k0 = 0; k0 < kNfreq; k0 += 32) {
      const int k = k0 + lane;
      const bool nz = k < kNfreq && a.filters[m * kNfreq + k] != 0.f;
      const unsigned bal = __ballot_sync(0xffffffffu, nz);
      if (bal) {
        const int first = k0 + __ffs(bal) - 1, last = k0 + 32 - __clz(bal);
        lo = first < lo ? first : lo;
        hi = last > hi ? last : hi;
      }
    }
    if (lane == 0) {
      mlo[m] = lo < hi ? lo : 0;
      mhi[m] = lo < hi ? hi : 0;
    }
  }

  SmemCol col{work + lane};
  SmemAud au{aud + lane * kHopPitch};
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt;
    const int t0 = (tile - b * a.tiles_per_utt) * kFrames;
    const float* src = a.audio + (size_t)b * a.n;
    const int p0 = t0 * kHop - kNfft / 2;  // original index of the tile's first sample (before reflection)
    __syncthreads();                       // previous tile's readers are done with aud / work
    for (int s = tid; s < kSpan; s += 32 * G) {
      int p = p0 + s;
      p = p < 0 ? -p : p;                          // reflect (no edge repeat), torch.stft center=True
      p = p >= a.n ? 2 * (a.n - 1) - p : p;
      const float v = (p >= 0 && p < a.n) ? __ldg(src + p) : 0.f;  // frames >= T of a ragged last tile read zeros
      aud[s + s / kHop] = v;                       // hop-major, pitch 161
    }
    __syncthreads();
    pass_a(g, G, au, col);
    __syncthreads();
    pass_b(g, G, col);
    __syncthreads();
    untangle_power(g, G, col);
    __syncthreads();
    const int t = t0 + lane;
    float vmax = -INFINITY;
    for (int m = g; m < a.n_mels; m += G) {
      const int lo = mlo[m], hi = mhi[m];
      const float* frow = a.filters + m * kNfreq;
      float acc = 0.f;
      for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(frow + k), work[pslot(k) * kFrames + lane], acc);
      const float v = log10f(fmaxf(acc, 1e-10f));  // audio.py:154
      if (t < a.T) {
        a.mel[((size_t)b * a.n_mels + m) * a.T + t] = v;
        vmax = fmaxf(vmax, v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > -INFINITY) atomic_max_float(a.umax + b, vmax + 0.0f);  // +0.0f: -0 -> +0
  }
}

// audio.py:155-156 on the whole (B, n_mels*T) tensor, in place
__global__ void __launch_bounds__(256) logmel_finish_kernel(float* __restrict__ mel, const float* __restrict__ umax, int B,
                                                            long long per_utt, int vec4) {
  const long long total = (long long)B * per_utt;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec4) {
    const long long n4 = total >> 2, per4 = per_utt >> 2;
    float4* m4 = reinterpret_cast<float4*>(mel);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float fl = umax[i / per4] - 8.0f;
      float4 v = m4[i];
      v.x = (fmaxf(v.x, fl) + 4.0f) / 4.0f;
      v.y = (fmaxf(v.y, fl) + 4.0f) / 4.0f;
      v.z = (fmaxf(v.z, fl) + 4.0f) / 4.0f;
      v.w = (fmaxf(v.w, fl) + 4.0f) / 4.0f;
      m4[i] = v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const float fl = umax[i / per_utt] - 8.0f;
      mel[i] = (fmaxf(mel[i], fl) + 4.0f) / 4.0f;
    }
  }
}

constexpr int kG = 4;
constexpr size_t kSmemBytes = (size_t)(kAudFloats + 2 + kWorkFloats) * sizeof(float) + 2 * kMaxMels * sizeof(int);

}  // namespace lm
}  // namespace qw

extern "C" {

size_t qw_log_mel_workspace_bytes(int B, int n_samples, int n_mels) {
  if (B <= 0 || n_samples <= 0 || n_mels <= 0) return 0;
  return qw::align_up((size_t)B * sizeof(float), 256);
}

int qw_log_mel(const float* audio, const float* filters, float* mel, void* workspace, size_t ws_bytes, int B, int n_samples,
               int n_mels, void* stream) {
  using namespace qw;
  using namespace qw::lm;
  QW_CHECK_ARG(audio && filters && mel && workspace, -1, "qw_log_mel: null pointer argument");
  QW_CHECK_ARG(B > 0 && n_mels > 0 && n_mels <= kMaxMels, -1, "qw_log_mel: bad shape B=%d n_mels=%d (n_mels <= %d)", B, n_mels,
               kMaxMels);
  QW_CHECK_ARG(n_samples > kNfft / 2 && n_samples % kHop == 0, -1,
               "qw_log_mel: n_samples=%d must be a multiple of %d and > %d (reflect padding)", n_samples, kHop, kNfft / 2);
  QW_CHECK_ARG(ws_bytes >= (size_t)B * sizeof(float), -3, "qw_log_mel: workspace too small");
  QW_CHECK_ARG((long long)B * n_samples < (1LL << 40), -1, "qw_log_mel: tensor too large");
  cudaStream_t st = (cudaStream_t)stream;
  Args a{};
  a.audio = audio;
  a.filters = filters;
  a.mel = mel;
  a.umax = (float*)workspace;
  a.B = B;
  a.n = n_samples;
  a.T = n_samples / kHop;  // frame T (the 3001st for 30 s) is dropped, audio.py:149
  a.n_mels = n_mels;
  a.tiles_per_utt = (a.T + kFrames - 1) / kFrames;
  a.num_tiles = B * a.tiles_per_utt;
  QW_CUDA_OK(cudaMemsetAsync(a.umax, 0xff, (size_t)B * sizeof(float), st));
  auto k = logmel_stft_kernel<kG>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const int cap = num_sms() * 3;
  const int grid = a.num_tiles < cap ? a.num_tiles : cap;
  {
    KernelTimer kt(kKLogMelStft, st);
    k<<<grid, 32 * kG, kSmemBytes, st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  {
    const long long per_utt = (long long)n_mels * a.T;
    const long long work = ((long long)B * per_utt + 3) / 4;
    long long blocks = (work + 255) / 256;
    const long long capf = (long long)num_sms() * 8;
    if (blocks > capf) blocks = capf;
    KernelTimer kt(kKLogMelFinish, st);
    const int vec4 = ((per_utt & 3) == 0 && ((uintptr_t)mel & 15) == 0) ? 1 : 0;
    logmel_finish_kernel<<<(int)blocks, 256, 0, st>>>(mel, a.umax, B, per_utt, vec4);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
