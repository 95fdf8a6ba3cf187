// Fused log-mel front end: replaces whisper.log_mel_spectrogram (/root/reference/whisper/whisper/audio.py:110-157),
// batched, with the max of :155 taken per utterance (the reference always passes one utterance).
//
//   logmel_prep_kernel   one CTA: support of every mel row + its non-zero filter values compacted and zero-padded to
//                        whole float4 chunks (whisper's 80 x 201 bank has 391 non-zeros) into the workspace, so the
//                        streaming CTAs stage ~3 KB instead of scanning 64 KB each.
//   logmel_stft_kernel   reflect-padded framing (bit-exact torch.stft(center=True) frame map) -> Hann -> 400-point
//                        real FFT -> power -> mel filterbank -> log10(max(., 1e-10)) -> mel (B, n_mels, T) + per-
//                        utterance running max (one atomic per warp per tile).
//   logmel_finish_sel_kernel  audio.py:155-156 -- max(., max_b - 8) ; (. + 4) / 4 -- needs the utterance maximum, i.e. a
//                        second pass.  The stft kernel already stores (v + 4) / 4 and leaves every warp's (min, max) of v per
//                        tile; x -> (x + 4) / 4 is monotone, so the second pass is y = max(y, (max_b - 8 + 4) / 4) and only
//                        has to touch tiles whose minimum lies below max_b - 8 (write-only when the whole tile does): nothing
//                        for noise-like audio, the padded / digitally silent stretches otherwise.  Bit-identical to the
//                        two-step formula.  (logmel_finish_kernel, the unconditional pass, remains for callers that pass
//                        only the B-float workspace.)
//
// A tile = 32 consecutive frames of one utterance per CTA iteration, 9 warps.  The kernel is bound by instruction
// issue (not by HBM: ~1.5 KB of traffic against ~14 k instructions per frame), so the layout is chosen for the fewest
// instructions per frame (qw_logmel_math.cuh: 400 = 16 x 25 on the real frame):
//   audio   the tile's 5 360-sample span, ONE bulk asynchronous copy (TMA engine, no per-sample instructions) issued
//           by one thread as soon as pass A has consumed the previous span, so its HBM latency hides under pass B /
//           mel; edge tiles (reflection, end of the clip, zero padding, misaligned rows) take a per-sample cp.async
//           path, tiles entirely inside the zero padding are not computed at all.
//   pass A  one n2 per LANE (25 of 32), warps over frames: taps 25 n1 + n2 are contiguous across lanes (no staging
//           skew), the 16 Hann weights and 8 twiddles of a lane live in registers for the whole kernel, real FFT16,
//           17 plane values stored to work[(plane * 25 + n2) * 33 + frame] (pitch 33: conflict-free here AND below).
//   pass B  one frame per LANE, warp = k1 (0..8): 25-point DFT in registers, |X|^2 straight into P[bin][frame].
//   mel     one frame per lane, warps over mel rows: float4 filter chunks (broadcast) x 4 power reads.
// HBM traffic is the algorithmic minimum: audio read once (hop overlap served from shared memory), mel written once by
// 128-byte rows.  The first version (one frame per lane everywhere, 200-point complex FFT + untangling, per-sample
// cp.async with a skewed staging index, scalar mel loop) issued 863 warp-instructions per frame
// (profiles/r1_ncu_full_logmel_b64_v2.txt); per source line 26 % of them were the mel loop, 14 % the fill, 11 % tap
// addressing and window-table loads.
#include <cstdlib>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/qw.h"
#include "qw_async.cuh"
#include "qw_common.cuh"
#include "qw_logmel_math.cuh"

namespace qw {
namespace lm {

constexpr int kFrames = 32;                                   // frames per tile
constexpr int kSpan = (kFrames - 1) * kHop + kNfft;           // 5360 samples feed one tile
constexpr int kNW = 9;                                        // warps per CTA: one k1 per warp in pass B
constexpr int kThreads = 32 * kNW;
constexpr int kPitch = 33;                                    // floats between consecutive (plane, n2) rows of `work`
constexpr int kPlaneStride = 25 * kPitch;                     // 825
constexpr int kWorkFloats = ((kPlanes * kPlaneStride + 3) / 4) * 4;  // 14028
constexpr int kPRows = 204;                                   // 201 bins + 3 zero rows (padded filter chunks read past bin 200)
constexpr int kMaxMels = 256;
constexpr int kMaxChunks = 192;                               // filter chunks staged in shared memory (whisper: 131 / 166)

// The filterbank as a chunk stream: a chunk = 4 consecutive bins of one mel row (weights zero-padded).  The chunks of the rows
// warp w owns (m = w, w + 9, ...) are stored consecutively in row order, so a warp walks [cstart[w], cstart[w + 1]) and closes
// a row (log, store, advance the output pointer by 9 rows) whenever a chunk carries the `last` flag.
// workspace: [umax: B floats, 256-aligned][Meta, 256 B][Chunk records]
struct Meta {
  int cstart[kNW + 1];
  int total;             // chunks
  int extra[kNW];        // pass A: the tile's frames 27..31 go to the five warps with the shortest chunk streams (-1: none)
  int pad[64 - 2 * kNW - 2];
};
struct __align__(16) Chunk {
  float4 w;              // weights of bins k .. k + 3
  int off;               // byte offset of P[k][0]
  int last;              // 1: closes its mel row
  int pad[2];
};
static_assert(sizeof(Meta) == 256 && sizeof(Chunk) == 32, "workspace layout");
__host__ __device__ inline size_t max_chunks(int n_mels) { return (size_t)n_mels * ((kNfreq + 3) / 4); }

struct Args {
  const float* audio;    // (B, n)
  float* mel;            // (B, n_mels, T)
  float* umax;           // (B) running max, initialised to 0xffffffff ("-inf" for the mixed int/uint atomics)
  float2* stats;         // (num_tiles, kNW) per-warp (min, max) of the tile's log10 values, or NULL: store v, not (v + 4) / 4
  const Meta* meta;
  const Chunk* chunks;
  int B, n, T, n_mels, tiles_per_utt, num_tiles;
  int n_in;            // row stride of `audio` = samples stored per utterance (pad_or_trim fused: n_in != n is allowed)
  const int* lengths;  // (B) valid samples per utterance (<= n_in) or NULL: every row holds min(n_in, n) valid samples
};

__global__ void __launch_bounds__(1024) logmel_prep_kernel(const float* __restrict__ filters, int n_mels, Meta* meta,
                                                          Chunk* chunks) {
  __shared__ int slo[kMaxMels], scnt[kMaxMels], soff[kMaxMels];
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
      scnt[m] = lo < hi ? (hi - lo + 3) >> 2 : 1;  // an all-zero row keeps one zero chunk: its output is log10(1e-10)
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    for (int w = 0; w < kNW; ++w) {
      meta->cstart[w] = o;
      for (int m = w; m < n_mels; m += kNW) {
        soff[m] = o;
        o += scnt[m];
      }
    }
    meta->cstart[kNW] = o;
    meta->total = o;
    // frames 27..31 of a tile: one each to the five warps with the fewest chunks (mel and the next tile's pass A share a
    // barrier interval, so this evens out the sum)
    bool used[kNW];
    for (int w = 0; w < kNW; ++w) {
      used[w] = false;
      meta->extra[w] = -1;
    }
    for (int j = 0; j < kFrames - 3 * kNW; ++j) {
      int best = -1;
      for (int w = 0; w < kNW; ++w)
        if (!used[w] && (best < 0 || meta->cstart[w + 1] - meta->cstart[w] < meta->cstart[best + 1] - meta->cstart[best])) best = w;
      used[best] = true;
      meta->extra[best] = 3 * kNW + j;
    }
  }
  __syncthreads();
  for (int m = warp; m < n_mels; m += 32)
    for (int c = lane; c < scnt[m]; c += 32) {
      const int k = slo[m] + 4 * c;
      float w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = k + e < kNfreq ? filters[m * kNfreq + k + e] : 0.f;
      Chunk ch;
      ch.w = make_float4(w[0], w[1], w[2], w[3]);
      ch.off = k * kFrames * (int)sizeof(float);
      ch.last = c == scnt[m] - 1 ? 1 : 0;
      ch.pad[0] = ch.pad[1] = 0;
      chunks[soff[m] + c] = ch;
    }
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// log2 of a NORMAL positive number (the argument is clamped to >= 1e-10 first): one MUFU, no denormal pre-scaling
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

constexpr size_t smem_bytes() {
  return (size_t)(kSpan + kWorkFloats + kPRows * kFrames) * sizeof(float) + (size_t)kMaxChunks * sizeof(Chunk) + 16;
}

__global__ void __launch_bounds__(kThreads, 2) logmel_stft_kernel(const Args a) {
  extern __shared__ __align__(128) float smem[];
  float* aud = smem;                        // [kSpan] tile samples (16-byte aligned: bulk-copy destination)
  float* work = aud + kSpan;                // [17 planes][25 n2] rows of pitch 33, one frame per column
  float* P = work + kWorkFloats;            // [204][32] power spectrum, one frame per column
  Chunk* sch = reinterpret_cast<Chunk*>(P + kPRows * kFrames);    // [kMaxChunks] the filterbank's chunk stream
  uint64_t* bar = reinterpret_cast<uint64_t*>(sch + kMaxChunks);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler too (uniform branches, no reconvergence code)

  // valid samples of utterance b: whisper.pad_or_trim (audio.py:65-88) fused -- samples past the clip are zeros, samples past n
  // are trimmed -- so a 1 s Speech-Commands clip crosses PCIe and HBM as 16 000 samples, not as 480 000
  auto valid_len = [&](int b) {
    int v = a.lengths ? a.lengths[b] : a.n_in;
    v = v < a.n_in ? v : a.n_in;
    return v < a.n ? v : a.n;
  };
  // kind 0: every sample of the tile (after reflection at the END of the padded signal) lies in the zero padding -> every power is 0
  //      1: interior tile -- no reflection, all samples valid, 16-byte aligned source -> one bulk copy
  //      2: anything else -> per-sample path
  struct Tile {
    int b, t0, kind;
  };
  auto tile_info = [&](int b, int r) {  // tile r of utterance b
    Tile ti;
    ti.b = b;
    ti.t0 = r * kFrames;
    const int p0 = ti.t0 * kHop - kNfft / 2, p1 = p0 + kSpan - 1;  // first / last original index (before reflection)
    const int nv = valid_len(ti.b);
    if (p0 >= nv && p0 >= 0 && (p1 < a.n || 2 * (a.n - 1) - p1 >= nv)) ti.kind = 0;
    else if (p0 >= 0 && p1 < nv && ((reinterpret_cast<uintptr_t>(a.audio + (size_t)ti.b * a.n_in + p0) & 15) == 0)) ti.kind = 1;
    else ti.kind = 2;
    return ti;
  };
  // asynchronous fill of the audio staging buffer for one tile; completion: the mbarrier (kind 1) or the cp.async group (kind 2)
  auto fill = [&](const Tile& ti) {
    const float* src = a.audio + (size_t)ti.b * a.n_in;
    const int p0 = ti.t0 * kHop - kNfft / 2;  // original index of the tile's first sample (before reflection)
    if (ti.kind == 1) {
      if (tid == 0) {
        fence_proxy_async();  // the span was read / written through the generic proxy until the barrier just passed
        mbar_arrive_expect_tx(bar, kSpan * 4);
        bulk_g2s(aud, src + p0, kSpan * 4, bar);
      }
    } else if (ti.kind == 2) {
      const int nv = valid_len(ti.b);
#pragma unroll 4
      for (int s = tid; s < kSpan; s += kThreads) {
        int p = p0 + s;
        p = p < 0 ? -p : p;                     // reflect (no edge repeat), torch.stft center=True
        p = p >= a.n ? 2 * (a.n - 1) - p : p;
        if (p >= 0 && p < nv) cp_async4(aud + s, src + p);
        else aud[s] = 0.f;                      // zero padding; frames >= T of a ragged last tile read zeros
      }
      cp_async_commit();
    }
  };

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // tile -> (utterance, tile of the utterance): one division here, then stepped by the grid size
  const int step_b = (int)gridDim.x / a.tiles_per_utt, step_r = (int)gridDim.x - step_b * a.tiles_per_utt;
  int nb = (int)blockIdx.x / a.tiles_per_utt, nr = (int)blockIdx.x - nb * a.tiles_per_utt;
  auto advance = [&]() {
    nb += step_b;
    nr += step_r;
    if (nr >= a.tiles_per_utt) {
      nr -= a.tiles_per_utt;
      ++nb;
    }
  };
  Tile cur{0, 0, 0};
  if ((int)blockIdx.x < a.num_tiles) {
    cur = tile_info(nb, nr);
    fill(cur);
  }
  const int nchunks = a.meta->total;
  const bool CACHED = nchunks <= kMaxChunks;  // uniform across the grid
  if (CACHED)
    for (int e = tid; e < nchunks * 2; e += kThreads)
      reinterpret_cast<float4*>(sch)[e] = reinterpret_cast<const float4*>(a.chunks)[e];
  const int c_begin = a.meta->cstart[warp], c_end = a.meta->cstart[warp + 1];
  const int fr_extra = a.meta->extra[warp];
  for (int e = tid; e < (kPRows - kNfreq) * kFrames; e += kThreads) P[kNfreq * kFrames + e] = 0.f;
  // pass A constants of this lane (n2 = lane; lanes 25..31 shadow lane 24: same inputs, same values, same addresses)
  const int n2 = lane < 25 ? lane : 24;
  float win[16], twr[8], twi[8];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) win[n1] = d_win[25 * n1 + n2];
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    twr[k1] = d_tw400_re[25 * k1 + n2];
    twi[k1] = d_tw400_im[25 * k1 + n2];
  }

  uint32_t phase = 0;
  const size_t row_step = (size_t)kNW * a.T;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int next = tile + (int)gridDim.x;
    Tile nxt{0, 0, 0};
    if (next < a.num_tiles) {
      advance();
      nxt = tile_info(nb, nr);
    }
    // The span is visible to a thread once IT has seen the mbarrier phase complete (bulk copy), so an interior tile needs no
    // CTA barrier here: a warp that finishes the previous tile's mel rows early starts pass A at once.  (P is next written in
    // pass B, behind the post-pass-A barrier, which every warp reaches only after its mel rows; `work` was last read before
    // the post-pass-B barrier.)  The per-sample path was filled by all threads: wait for everybody's copies.
    if (cur.kind == 1) {
      mbar_wait(bar, phase);
      phase ^= 1;
    } else if (cur.kind == 2) {
      cp_async_wait_all();
      __syncthreads();
    }
    const bool t_ok = cur.t0 + lane < a.T;
    float vmax = -INFINITY, vmin = INFINITY;
    float* mp = a.mel + (size_t)(cur.b * a.n_mels + warp) * a.T + (cur.t0 + lane);  // this warp's first mel row, this lane's frame
    if (cur.kind == 0) {
      // zero padding (29/30 of a Speech-Commands batch): power 0 in every bin -> the clamp value, the same expression and bits
      // the arithmetic path produces
      if (next < a.num_tiles) fill(nxt);
      const float v = 0.30102999566398120f * lg2_ftz(fmaxf(0.f, 1e-10f));
      const float y = a.stats ? fmaf(v, 0.25f, 1.0f) : v;
      if (t_ok && warp < a.n_mels) {
        for (int m = warp; m < a.n_mels; m += kNW, mp += row_step) *mp = y;
        vmax = v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0) {
        if (warp == 0 && vmax > -INFINITY) atomic_max_float(a.umax + cur.b, vmax + 0.0f);
        if (a.stats) a.stats[(size_t)tile * kNW + warp] = make_float2(vmax > -INFINITY ? vmax : INFINITY, vmax);
      }
      cur = nxt;
      continue;
    }
    // ---- pass A: lane = n2, warps over the tile's frames (w, w + 9, w + 18 and one of 27..31 for five of the warps)
    const int nfr = fr_extra >= 0 ? 4 : 3;
#pragma unroll 1
    for (int it = 0; it < nfr; ++it) {
      const int fr = it < 3 ? warp + kNW * it : fr_extra;
      const float* ap = aud + kHop * fr + n2;
      float x[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) x[n1] = ap[25 * n1];
      float* wp = work + n2 * kPitch + fr;
      pass_a_one(x, win, twr, twi, [&](int pl, float v) { wp[pl * kPlaneStride] = v; });
    }
    __syncthreads();
    if (next < a.num_tiles) fill(nxt);  // overlaps pass B / mel
    // ---- pass B: lane = frame, warp = k1
    {
      const int k1 = warp;
      const float* re = work + (k1 == 0 ? 0 : 2 * k1 - 1) * kPlaneStride + lane;
      const float* im = work + 2 * k1 * kPlaneStride + lane;
      float* pp = P + lane;
      pass_b_one(
          k1, [&](int j) { return re[j * kPitch]; }, [&](int j) { return im[j * kPitch]; },
          [&](int k, float v) { pp[k * kFrames] = v; });
    }
    __syncthreads();
    // ---- mel filterbank + log: lane = frame, each warp walks the chunk stream of its rows
    {
      const float* pl = P + lane;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      auto walk = [&](const Chunk* __restrict__ cp, const Chunk* __restrict__ ce, auto sel) {
        constexpr bool SEL = decltype(sel)::value;
#pragma unroll 1
        for (; cp != ce; ++cp) {
          const float4 f = cp->w;
          const int2 md = *reinterpret_cast<const int2*>(&cp->off);
          const float* pw = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pl) + md.x);
          acc0 = fmaf(f.x, pw[0], acc0);
          acc1 = fmaf(f.y, pw[kFrames], acc1);
          acc2 = fmaf(f.z, pw[2 * kFrames], acc2);
          acc3 = fmaf(f.w, pw[3 * kFrames], acc3);
          if (md.y) {  // row complete (warp-uniform)
            const float v = 0.30102999566398120f * lg2_ftz(fmaxf((acc0 + acc1) + (acc2 + acc3), 1e-10f));  // audio.py:154
            if (t_ok) {
              *mp = SEL ? fmaf(v, 0.25f, 1.0f) : v;  // == (v + 4) / 4 bit for bit (one rounding either way)
              vmax = fmaxf(vmax, v);
              if (SEL) vmin = fminf(vmin, v);
            }
            mp += row_step;
            acc0 = acc1 = acc2 = acc3 = 0.f;
          }
        }
      };
      const Chunk* cb = CACHED ? sch : a.chunks;  // shared memory (broadcast LDS.128 + LDS.64); dense / very wide banks: global
      if (a.stats) walk(cb + c_begin, cb + c_end, std::true_type{});
      else walk(cb + c_begin, cb + c_end, std::false_type{});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if (lane == 0) {
      if (vmax > -INFINITY) atomic_max_float(a.umax + cur.b, vmax + 0.0f);  // +0.0f: -0 -> +0
      if (a.stats) a.stats[(size_t)tile * kNW + warp] = make_float2(vmin, vmax);
    }
    cur = nxt;
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


// Selective second pass (see the header).  Each round a CTA looks at 8 tiles, one per warp (9 (min, max) pairs each); the tiles
// that need work are then rewritten by the whole CTA, rows across warps, so a lone read-modify-write tile in a large batch costs
// one memory round trip rather than a warp's 80 sequential ones (measured: 43 us for ONE such tile among 48 128).
__global__ void __launch_bounds__(256) logmel_finish_sel_kernel(float* __restrict__ mel, const float* __restrict__ umax,
                                                                const float2* __restrict__ stats, int n_mels, int T, int tiles_per_utt,
                                                                int num_tiles) {
  __shared__ int s_mode[8];
  __shared__ float s_yfl[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * 8; base < num_tiles; base += gridDim.x * 8) {
    {
      const int tile = base + warp;
      int mode = 0;  // 0: (v + 4) / 4 is already final; 1: the whole tile sits on the floor (write only); 2: read-modify-write
      float yfl = 0.f;
      if (tile < num_tiles) {
        float2 mm = lane < kNW ? stats[(size_t)tile * kNW + lane] : make_float2(INFINITY, -INFINITY);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          mm.x = fminf(mm.x, __shfl_xor_sync(0xffffffffu, mm.x, o));
          mm.y = fmaxf(mm.y, __shfl_xor_sync(0xffffffffu, mm.y, o));
        }
        const float fl = umax[tile / tiles_per_utt] - 8.0f;
        yfl = fmaf(fl, 0.25f, 1.0f);
        if (mm.x < fl) mode = mm.y <= fl ? 1 : 2;
      }
      if (lane == 0) {
        s_mode[warp] = mode;
        s_yfl[warp] = yfl;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int mode = s_mode[j];
      if (mode == 0) continue;  // CTA-uniform
      const int tile = base + j;
      const int b = tile / tiles_per_utt, t = (tile - b * tiles_per_utt) * kFrames + lane;
      const float yfl = s_yfl[j];
      float* p = mel + ((size_t)b * n_mels + warp) * T + t;
      const size_t step = (size_t)8 * T;
      if (t >= T) {
      } else if (mode == 1) {
#pragma unroll 4
        for (int m = warp; m < n_mels; m += 8, p += step) *p = yfl;
      } else {
#pragma unroll 4
        for (int m = warp; m < n_mels; m += 8, p += step) *p = fmaxf(*p, yfl);
      }
    }
    __syncthreads();
  }
}

static int launch_stft(const Args& a, cudaStream_t st) {
  auto k = logmel_stft_kernel;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  const int per_sm = (int)((227 * 1024) / (smem_bytes() + 1024));
  const int cap = num_sms() * per_sm;
  const int grid = a.num_tiles < cap ? a.num_tiles : cap;
  {
    KernelTimer kt(kKLogMelStft, st);
    k<<<grid, kThreads, smem_bytes(), st>>>(a);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace lm
}  // namespace qw

namespace qw {
namespace lm {
static size_t prep_bytes(int n_mels) { return sizeof(Meta) + align_up(max_chunks(n_mels) * sizeof(Chunk), 256); }

static int run_prepare(const float* filters, int n_mels, void* prep, cudaStream_t st) {
  Meta* meta = (Meta*)prep;
  Chunk* chunks = (Chunk*)((unsigned char*)prep + sizeof(Meta));
  {
    KernelTimer kt(kKLogMelPrep, st);
    logmel_prep_kernel<<<1, 1024, 0, st>>>(filters, n_mels, meta, chunks);
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

// per-call workspace: [umax: B floats][pad to 256][stats: num_tiles x kNW float2]
static size_t call_ws_bytes(int B, int n_samples) {
  const size_t tiles = (size_t)B * ((n_samples / kHop + kFrames - 1) / kFrames);
  return align_up((size_t)B * sizeof(float), 256) + align_up(tiles * kNW * sizeof(float2), 256);
}

static int run_prepared(const float* audio, const int* lengths, const void* prep, float* mel, void* ws, size_t ws_bytes, int B, int n_in,
                        int n_samples, int n_mels, cudaStream_t st) {
  float* umax = (float*)ws;
  // the selective second pass needs the per-tile statistics; a caller that brought only the B floats of the first interface
  // gets the unconditional pass
  const bool sel = ws_bytes >= call_ws_bytes(B, n_samples) && (((uintptr_t)ws & 7) == 0);
  Args a{};
  a.audio = audio;
  a.n_in = n_in;
  a.lengths = lengths;
  a.mel = mel;
  a.umax = umax;
  a.stats = sel ? (float2*)((unsigned char*)ws + align_up((size_t)B * sizeof(float), 256)) : nullptr;
  a.meta = (const Meta*)prep;
  a.chunks = (const Chunk*)((const unsigned char*)prep + sizeof(Meta));
  a.B = B;
  a.n = n_samples;
  a.T = n_samples / kHop;  // frame T (the 3001st for 30 s) is dropped, audio.py:149
  a.n_mels = n_mels;
  a.tiles_per_utt = (a.T + kFrames - 1) / kFrames;
  a.num_tiles = B * a.tiles_per_utt;
  QW_CUDA_OK(cudaMemsetAsync(a.umax, 0xff, (size_t)B * sizeof(float), st));
  if (int e = launch_stft(a, st)) return e;
  if (sel) {
    const int cap = num_sms() * 8, want = (a.num_tiles + 7) / 8;
    KernelTimer kt(kKLogMelFinish, st);
    logmel_finish_sel_kernel<<<want < cap ? want : cap, 256, 0, st>>>(mel, a.umax, a.stats, n_mels, a.T, a.tiles_per_utt,
                                                                                 a.num_tiles);
  } else {
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
  return lm::run_prepared(audio, nullptr, prep, mel, workspace, ws_bytes, B, n_samples, n_samples, n_mels, (cudaStream_t)stream);
}

int qw_log_mel_padded(const float* audio, const int* lengths, const void* prep, float* mel, void* workspace, size_t ws_bytes, int B,
                      int n_in, int n_samples, int n_mels, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(audio && prep && mel && workspace, -1, "qw_log_mel_padded: null pointer argument");
  if (int e = lm::check_shape(B, n_samples, n_mels)) return e;
  QW_CHECK_ARG(n_in > 0, -1, "qw_log_mel_padded: n_in=%d must be positive", n_in);
  QW_CHECK_ARG(ws_bytes >= (size_t)B * sizeof(float), -3, "qw_log_mel_padded: workspace too small (B floats)");
  QW_CHECK_ARG(((uintptr_t)prep & 255) == 0 && ((uintptr_t)workspace & 3) == 0, -1, "qw_log_mel_padded: misaligned buffer");
  return lm::run_prepared(audio, lengths, prep, mel, workspace, ws_bytes, B, n_in, n_samples, n_mels, (cudaStream_t)stream);
}

size_t qw_log_mel_call_workspace_bytes(int B, int n_samples) {
  if (B <= 0 || n_samples <= 0) return 0;
  return qw::lm::call_ws_bytes(B, n_samples);
}

size_t qw_log_mel_workspace_bytes(int B, int n_samples, int n_mels) {
  if (B <= 0 || n_samples <= 0 || n_mels <= 0) return 0;
  return qw::lm::call_ws_bytes(B, n_samples) + qw::lm::prep_bytes(n_mels);
}

int qw_log_mel(const float* audio, const float* filters, float* mel, void* workspace, size_t ws_bytes, int B, int n_samples,
               int n_mels, void* stream) {
  using namespace qw;
  QW_CHECK_ARG(audio && filters && mel && workspace, -1, "qw_log_mel: null pointer argument");
  if (int e = lm::check_shape(B, n_samples, n_mels)) return e;
  QW_CHECK_ARG(ws_bytes >= qw_log_mel_workspace_bytes(B, n_samples, n_mels), -3, "qw_log_mel: workspace too small");
  QW_CHECK_ARG(((uintptr_t)workspace & 255) == 0, -1, "qw_log_mel: workspace must be 256-byte aligned");
  unsigned char* ws = (unsigned char*)workspace;
  const size_t cws = lm::call_ws_bytes(B, n_samples);
  void* prep = ws + cws;
  if (int e = lm::run_prepare(filters, n_mels, prep, (cudaStream_t)stream)) return e;
  return lm::run_prepared(audio, nullptr, prep, mel, ws, cws, B, n_samples, n_samples, n_mels, (cudaStream_t)stream);
}

}  // extern "C"
