// Fused log-mel front end: replaces whisper.log_mel_spectrogram (/root/reference/whisper/whisper/audio.py:110-157),
// batched, with the max of :155 taken per utterance (the reference always passes one utterance).
//
//   logmel_prep_kernel   one CTA: support [lo, hi) of every mel row + the non-zero filter values compacted (whisper's
//                        80 x 201 bank has 391 non-zeros) into the workspace, so the streaming CTAs stage ~3 KB instead
//                        of scanning 64 KB each.
//   logmel_stft_kernel   reflect-padded framing (bit-exact torch.stft(center=True) frame map) -> Hann -> 400-point
//                        real FFT -> power -> mel filterbank -> log10(max(., 1e-10)) -> mel (B, n_mels, T) + per-
//                        utterance running max (one atomic per warp per tile).
//   logmel_finish_kernel max(., max_b - 8) ; (. + 4) / 4, in place (the tensor was just written: L2-resident).
//
// Layout: one STFT frame per LANE, a tile = 32 consecutive frames of one utterance per CTA iteration, the G warps of
// the CTA split the butterflies / mel rows of those 32 frames.  Shared memory per CTA: the tile's audio span (5 360
// samples, skewed by one float per 32 so frame-strided reads are conflict-free) + a 400 x 32 float work array indexed
// [slot][lane] (every access is one 128-byte row) = 76 KB.  The NEXT tile's audio is fetched with cp.async right
// after pass A has consumed the current one, so its HBM latency hides under pass B / untangle / mel.  HBM traffic is
// the algorithmic minimum: audio read once (hop overlap served from shared memory), mel written once by 128-byte rows.
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/qw.h"
#include "qw_common.cuh"
#include "qw_logmel_math.cuh"

namespace qw {
namespace lm {

constexpr int kFrames = 32;                                   // frames per tile == lanes
constexpr int kSpan = (kFrames - 1) * kHop + kNfft;           // 5360 samples feed one tile
constexpr int kAudFloats = ((kSpan + (kSpan >> 5) + 3) / 4) * 4;  // 5528 (skewed)
constexpr int kWorkFloats = kNfft * kFrames;                  // 12800
constexpr int kMaxMels = 256;
constexpr int kMaxNnz = 1024;                                 // compact filter values staged in shared memory

// workspace: [umax: B floats, 256-aligned][meta: lo[256] hi[256] off[256] total pad -> 4 KB][vals: n_mels * 201 floats]
struct Meta {
  int lo[kMaxMels], hi[kMaxMels], off[kMaxMels];
  int total, pad[255];
};

struct Args {
  const float* audio;    // (B, n)
  const float* filters;  // (n_mels, 201)
  float* mel;            // (B, n_mels, T)
  float* umax;           // (B) running max, initialised to 0xffffffff ("-inf" for the mixed int/uint atomics)
  const Meta* meta;
  const float* vals;
  int B, n, T, n_mels, tiles_per_utt, num_tiles;
  int n_in;            // row stride of `audio` = samples stored per utterance (pad_or_trim fused: n_in != n is allowed)
  const int* lengths;  // (B) valid samples per utterance (<= n_in) or NULL: every row holds min(n_in, n) valid samples
};

__global__ void __launch_bounds__(1024) logmel_prep_kernel(const float* __restrict__ filters, int n_mels, Meta* meta, float* vals) {
  __shared__ int slo[kMaxMels], shi[kMaxMels], soff[kMaxMels + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = warp; m < n_mels; m += 32) {
    int lo = kNfreq, hi = 0;
    for (int k0 = 0; k0 < kNfreq; k0 += 32) {
      const int k = k0 + lane;
      const bool nz = k < kNfreq && filters[m * kNfreq + k] != 0.f;
      const unsigned bal = __ballot_sync(0xffffffffu, nz);
      if (bal) {
        const int first = k0 + __ffs(bal) - 1, last = k0 + 32 - __clz(bal);
        lo = first < lo ? first : lo;
        hi = last > hi ? last : hi;
      }
    }
    if (lane == 0) {
      slo[m] = lo < hi ? lo : 0;
      shi[m] = lo < hi ? hi : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    for (int m = 0; m < n_mels; ++m) {
      soff[m] = o;
      o += shi[m] - slo[m];
    }
    soff[n_mels] = o;
    meta->total = o;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    meta->lo[m] = slo[m];
    meta->hi[m] = shi[m];
    meta->off[m] = soff[m];
  }
  for (int m = warp; m < n_mels; m += 32)
    for (int k = slo[m] + lane; k < shi[m]; k += 32) vals[soff[m] + k - slo[m]] = filters[m * kNfreq + k];
}

struct SmemCol {
  float* base;  // &work[lane]
  __device__ __forceinline__ float& at(int e) { return base[e * kFrames]; }
};
struct SmemAud {
  const float* base;  // &aud[165 * lane]
  __device__ __forceinline__ float tap(int j) const { return base[j + (j >> 5)]; }
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int G>
__global__ void __launch_bounds__(32 * G) logmel_stft_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  float* aud = smem;                       // [kAudFloats] skewed tile samples
  float* work = smem + kAudFloats;         // [400][32]
  float* fvals = work + kWorkFloats;       // [kMaxNnz]
  int* mlo = reinterpret_cast<int*>(fvals + kMaxNnz);
  int* mhi = mlo + kMaxMels;
  int* moff = mhi + kMaxMels;
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  constexpr int NT = 32 * G;

  // asynchronous fill of the audio staging buffer for one tile (reflect padding resolved per sample)
  // valid samples of utterance b: whisper.pad_or_trim (audio.py:65-88) fused -- samples past the clip are zeros, samples past n
  // are trimmed -- so a 1 s Speech-Commands clip crosses PCIe and HBM as 16 000 samples, not as 480 000
  auto valid_len = [&](int b) {
    int v = a.lengths ? a.lengths[b] : a.n_in;
    v = v < a.n_in ? v : a.n_in;
    return v < a.n ? v : a.n;
  };
  // a tile all of whose samples (after reflection at the END of the padded signal) lie in the zero padding: every power is 0
  auto tile_is_silent = [&](int tile) {
    const int b = tile / a.tiles_per_utt;
    const int t0 = (tile - b * a.tiles_per_utt) * kFrames;
    const int p0 = t0 * kHop - kNfft / 2, p1 = p0 + kSpan - 1;  // first / last original index (before reflection)
    const int nv = valid_len(b);
    if (p0 < nv || p0 < 0) return false;
    return p1 < a.n || 2 * (a.n - 1) - p1 >= nv;
  };
  auto fill = [&](int tile) {
    const int b = tile / a.tiles_per_utt;
    const int t0 = (tile - b * a.tiles_per_utt) * kFrames;
    const float* src = a.audio + (size_t)b * a.n_in;
    const int nv = valid_len(b);
    const int p0 = t0 * kHop - kNfft / 2;  // original index of the tile's first sample (before reflection)
    if (!tile_is_silent(tile)) {
#pragma unroll 4
      for (int s = tid; s < kSpan; s += NT) {
        int p = p0 + s;
        p = p < 0 ? -p : p;                     // reflect (no edge repeat), torch.stft center=True
        p = p >= a.n ? 2 * (a.n - 1) - p : p;
        float* dst = aud + s + (s >> 5);
        if (p >= 0 && p < nv) cp_async4(dst, src + p);
        else *dst = 0.f;                        // zero padding; frames >= T of a ragged last tile read zeros
      }
    }
    cp_async_commit();
  };

  if ((int)blockIdx.x < a.num_tiles) fill(blockIdx.x);
  const int nnz = a.meta->total;
  const bool CACHED = nnz <= kMaxNnz;  // uniform across the grid
  for (int m = tid; m < a.n_mels; m += NT) {
    mlo[m] = a.meta->lo[m];
    mhi[m] = a.meta->hi[m];
    moff[m] = a.meta->off[m];
  }
  if (CACHED)
    for (int e = tid; e < nnz; e += NT) fvals[e] = a.vals[e];

  SmemCol col{work + lane};
  SmemAud au{aud + kLanePitch * lane};
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_utt;
    const int t0 = (tile - b * a.tiles_per_utt) * kFrames;
    cp_async_wait_all();
    __syncthreads();  // audio of this tile visible; previous tile's mel readers are done with `work`
    const int t = t0 + lane;
    float vmax = -INFINITY;
    if (tile_is_silent(tile)) {
      // zero padding (29/30 of a Speech-Commands batch): power 0 in every bin -> the clamp value, the same expression and bits
      // the arithmetic path produces
      if (tile + (int)gridDim.x < a.num_tiles) fill(tile + gridDim.x);
      const float v = 0.30102999566398120f * __log2f(fmaxf(0.f, 1e-10f));
      if (t < a.T) {
        for (int m = g; m < a.n_mels; m += G) a.mel[((size_t)b * a.n_mels + m) * a.T + t] = v;
        vmax = v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0 && g == 0 && vmax > -INFINITY) atomic_max_float(a.umax + b, vmax + 0.0f);
      continue;
    }
    pass_a(g, G, au, col);
    __syncthreads();
    if (tile + (int)gridDim.x < a.num_tiles) fill(tile + gridDim.x);  // overlaps pass B / untangle / mel
    pass_b(g, G, col);
    __syncthreads();
    untangle_power(g, G, col);
    __syncthreads();
    for (int m = g; m < a.n_mels; m += G) {
      const int lo = mlo[m], hi = mhi[m];
      const bool top = hi == kNfreq;          // P[200] lives at float slot 1, everything else at float 2k
      const int n = (top ? 200 : hi) - lo;
      const float* pw = work + lane + 2 * kFrames * lo;
      float acc0 = 0.f, acc1 = 0.f;
      if (CACHED) {  // filter values from shared memory (shared-space pointer: LDS with immediate offsets)
        const float* fv = fvals + moff[m];
        int e = 0;
#pragma unroll 2
        for (; e + 1 < n; e += 2) {
          acc0 = fmaf(fv[e], pw[e * 2 * kFrames], acc0);
          acc1 = fmaf(fv[e + 1], pw[(e + 1) * 2 * kFrames], acc1);
        }
        if (e < n) acc0 = fmaf(fv[e], pw[e * 2 * kFrames], acc0);
        if (top) acc1 = fmaf(fv[200 - lo], work[kFrames + lane], acc1);
      } else {       // dense / very wide filter banks: values stay in global memory
        const float* fv = a.vals + moff[m];
        for (int e = 0; e < n; ++e) acc0 = fmaf(__ldg(fv + e), pw[e * 2 * kFrames], acc0);
        if (top) acc1 = fmaf(__ldg(fv + 200 - lo), work[kFrames + lane], acc1);
      }
      const float v = 0.30102999566398120f * __log2f(fmaxf(acc0 + acc1, 1e-10f));  // audio.py:154
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

constexpr size_t smem_bytes() { return (size_t)(kAudFloats + kWorkFloats + kMaxNnz) * sizeof(float) + 3 * kMaxMels * sizeof(int); }
static int g_warps = 0;  // 0 = default; tools may override through QW_LOGMEL_WARPS (4 or 8)

template <int G>
static int launch_stft(const Args& a, cudaStream_t st) {
  auto k = logmel_stft_kernel<G>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  const int per_sm = (int)((227 * 1024) / (smem_bytes() + 1024));
  const int cap = num_sms() * per_sm;
  const int grid = a.num_tiles < cap ? a.num_tiles : cap;
  {
    KernelTimer kt(kKLogMelStft, st);
    k<<<grid, 32 * G, smem_bytes(), st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace lm
}  // namespace qw

namespace qw {
namespace lm {
static size_t prep_bytes(int n_mels) { return sizeof(Meta) + align_up((size_t)n_mels * kNfreq * sizeof(float), 256); }

static int run_prepare(const float* filters, int n_mels, void* prep, cudaStream_t st) {
  Meta* meta = (Meta*)prep;
  float* vals = (float*)((unsigned char*)prep + sizeof(Meta));
  {
    KernelTimer kt(kKLogMelPrep, st);
    logmel_prep_kernel<<<1, 1024, 0, st>>>(filters, n_mels, meta, vals);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_prepared(const float* audio, const int* lengths, const void* prep, float* mel, float* umax, int B, int n_in, int n_samples,
                        int n_mels, cudaStream_t st) {
  Args a{};
  a.audio = audio;
  a.n_in = n_in;
  a.lengths = lengths;
  a.filters = nullptr;
  a.mel = mel;
  a.umax = umax;
  a.meta = (const Meta*)prep;
  a.vals = (const float*)((const unsigned char*)prep + sizeof(Meta));
  a.B = B;
  a.n = n_samples;
  a.T = n_samples / kHop;  // frame T (the 3001st for 30 s) is dropped, audio.py:149
  a.n_mels = n_mels;
  a.tiles_per_utt = (a.T + kFrames - 1) / kFrames;
  a.num_tiles = B * a.tiles_per_utt;
  QW_CUDA_OK(cudaMemsetAsync(a.umax, 0xff, (size_t)B * sizeof(float), st));
  if (g_warps == 0) {
    const char* e = getenv("QW_LOGMEL_WARPS");
    g_warps = (e && atoi(e) == 4) ? 4 : 8;
  }
  if (int e = (g_warps == 4 ? launch_stft<4>(a, st) : launch_stft<8>(a, st))) return e;
  {
    const long long per_utt = (long long)n_mels * a.T;
    const long long work = ((long long)B * per_utt + 3) / 4;
    long long blocks = (work + 255) / 256;
    const long long capf = (long long)num_sms() * 8;
    if (blocks > capf) blocks = capf;
    const int vec4 = ((per_utt & 3) == 0 && ((uintptr_t)mel & 15) == 0) ? 1 : 0;
    KernelTimer kt(kKLogMelFinish, st);
    logmel_finish_kernel<<<(int)blocks, 256, 0, st>>>(mel, a.umax, B, per_utt, vec4);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

static int check_shape(int B, int n_samples, int n_mels) {
  QW_CHECK_ARG(B > 0 && n_mels > 0 && n_mels <= kMaxMels, -1, "qw_log_mel: bad shape B=%d n_mels=%d (n_mels <= %d)", B, n_mels, kMaxMels);
  // any length works like the reference: torch.stft(center=True) yields 1 + n / 160 frames and audio.py:149 drops the last one, so
  // T = n / 160 and the samples past the last full hop still feed the last frames (and the reflection at the end)
  QW_CHECK_ARG(n_samples > kNfft / 2 && n_samples >= kHop, -1, "qw_log_mel: n_samples=%d must be > %d (reflect padding) and >= %d (one frame)",
               n_samples, kNfft / 2, kHop);
  QW_CHECK_ARG((long long)B * n_samples < (1LL << 40), -1, "qw_log_mel: tensor too large");
  return 0;
}
}  // namespace lm
}  // namespace qw

extern "C" {

size_t qw_log_mel_prep_bytes(int n_mels) { return n_mels > 0 ? qw::lm::prep_bytes(n_mels) : 0; }

int qw_log_mel_prepare(const float* filters, int n_mels, void* prep, size_t prep_bytes, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(filters && prep, -1, "qw_log_mel_prepare: null pointer argument");
  QW_CHECK_ARG(n_mels > 0 && n_mels <= lm::kMaxMels, -1, "qw_log_mel_prepare: n_mels=%d outside [1, %d]", n_mels, lm::kMaxMels);
  QW_CHECK_ARG(prep_bytes >= lm::prep_bytes(n_mels), -3, "qw_log_mel_prepare: prep buffer too small");
  QW_CHECK_ARG(((uintptr_t)prep & 255) == 0, -1, "qw_log_mel_prepare: prep buffer must be 256-byte aligned");
  return lm::run_prepare(filters, n_mels, prep, (cudaStream_t)stream);
}

int qw_log_mel_prepared(const float* audio, const void* prep, float* mel, void* workspace, size_t ws_bytes, int B, int n_samples,
                        int n_mels, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(audio && prep && mel && workspace, -1, "qw_log_mel_prepared: null pointer argument");
  if (int e = lm::check_shape(B, n_samples, n_mels)) return e;
  QW_CHECK_ARG(ws_bytes >= (size_t)B * sizeof(float), -3, "qw_log_mel_prepared: workspace too small (B floats)");
  QW_CHECK_ARG(((uintptr_t)prep & 255) == 0 && ((uintptr_t)workspace & 3) == 0, -1, "qw_log_mel_prepared: misaligned buffer");
  return lm::run_prepared(audio, nullptr, prep, mel, (float*)workspace, B, n_samples, n_samples, n_mels, (cudaStream_t)stream);
}

int qw_log_mel_padded(const float* audio, const int* lengths, const void* prep, float* mel, void* workspace, size_t ws_bytes, int B,
                      int n_in, int n_samples, int n_mels, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(audio && prep && mel && workspace, -1, "qw_log_mel_padded: null pointer argument");
  if (int e = lm::check_shape(B, n_samples, n_mels)) return e;
  QW_CHECK_ARG(n_in > 0, -1, "qw_log_mel_padded: n_in=%d must be positive", n_in);
  QW_CHECK_ARG(ws_bytes >= (size_t)B * sizeof(float), -3, "qw_log_mel_padded: workspace too small (B floats)");
  QW_CHECK_ARG(((uintptr_t)prep & 255) == 0 && ((uintptr_t)workspace & 3) == 0, -1, "qw_log_mel_padded: misaligned buffer");
  return lm::run_prepared(audio, lengths, prep, mel, (float*)workspace, B, n_in, n_samples, n_mels, (cudaStream_t)stream);
}

size_t qw_log_mel_workspace_bytes(int B, int n_samples, int n_mels) {
  if (B <= 0 || n_samples <= 0 || n_mels <= 0) return 0;
  return qw::align_up((size_t)B * sizeof(float), 256) + qw::lm::prep_bytes(n_mels);
}

int qw_log_mel(const float* audio, const float* filters, float* mel, void* workspace, size_t ws_bytes, int B, int n_samples,
               int n_mels, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(audio && filters && mel && workspace, -1, "qw_log_mel: null pointer argument");
  if (int e = lm::check_shape(B, n_samples, n_mels)) return e;
  QW_CHECK_ARG(ws_bytes >= qw_log_mel_workspace_bytes(B, n_samples, n_mels), -3, "qw_log_mel: workspace too small");
  QW_CHECK_ARG(((uintptr_t)workspace & 255) == 0, -1, "qw_log_mel: workspace must be 256-byte aligned");
  unsigned char* ws = (unsigned char*)workspace;
  void* prep = ws + align_up((size_t)B * sizeof(float), 256);
  if (int e = lm::run_prepare(filters, n_mels, prep, (cudaStream_t)stream)) return e;
  return lm::run_prepared(audio, nullptr, prep, mel, (float*)ws, B, n_samples, n_samples, n_mels, (cudaStream_t)stream);
}

}  // extern "C"
