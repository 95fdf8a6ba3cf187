// Fast path of QuantumConv1d for the Whisper stem regime: fp32, n_qubits = 4, kernel_size = 3, stride 1 or 2,
// amplitude embedding, L % 4 == 0, L_out % 4 == 0, O % 4 == 0, O <= 576, 16-byte aligned tensors.
// Everything else goes through the generic kernels in qw_conv1d_kernels.cuh.
//
// All four streaming kernels are persistent (grid ~ 2-3 CTAs per SM, static round-robin tile assignment) and
// move their tiles with the TMA engine (cp.async.bulk[.tensor]) through mbarrier-synchronised shared-memory
// rings, so HBM latency is hidden by the ring depth rather than by occupancy:
//
//   fast_fwd_kernel<S,RC>   x tile ring (3-D tensor map, zero padding by OOB fill) -> pre_conv partials with a
//                           "4 rows x 8 lanes x 4 adjacent windows" lane layout -> circuit (thread/window) ->
//                           post_conv with 128-bit coalesced stores.
//   fast_bwd_gy3_kernel     streams gy once (SWIZZLE_128B tiles): gout = post_conv^T gy and grad post_conv.{weight,bias} on the
//                           tensor pipe (mma.sync 3xTF32); its FUSED form carries a small-batch data layer through the adjoint
//                           and pre_conv^T as well.  fast_bwd_gy2_kernel is the FFMA form (A/B switch GY_MMA=0).
//   fast_bwd_adj_kernel     adjoint differentiation of the circuit, one thread per window -> gpre (halo-padded).
//   fast_bwd_pre_kernel<S,PAR,GX>  x tile ring in, grad_x written in place and stored back by TMA; lanes along
//                           channels hold pre_conv weights and their gradient accumulators in registers.
//   fast_finalize_kernel    deterministic reduction of the per-CTA partial rows.
#include <cstdlib>

#include "../../include/qw.h"
#include "qw_act.cuh"
#include "qw_conv1d_plan.cuh"
#include "qw_tma.cuh"

namespace qw {

constexpr int FQ = 4;     // n_qubits on the fast path
constexpr int FTW = 32;   // windows a tile's shared-memory boxes cover (forward, gy streaming)
// (The kernels still carry a run-time tile stride `tw`; it is always FTW -- see make_fast_plan.)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 1024-byte aligned start inside the dynamic shared-memory window.  Computed as an OFFSET from the
// `extern __shared__` symbol so the compiler keeps the pointer in the shared address space (LDS/STS, not LD/ST).
__device__ __forceinline__ unsigned char* align1024(unsigned char* p) {
  return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u);
}

// =============================================================================================== forward
struct FastFwdArgs {
  const float *w_pre, *b_pre, *qw, *w_post, *b_post;
  float *y, *pre_save, *qout_save;
  int B, C, L, P, O, Lq, Lout;
  int tiles_per_utt, num_tiles, chunks_per_tile;
  int tw;  // tile stride in windows
  int early_tma;
  unsigned long long* tl;
  int dbg;  // experiment switch DBG_FWD (results are garbage): 1 = skip the pre_conv FMAs, 2 = skip the post_conv FMAs, 4 = skip the circuit
  int act;  // 1: store gelu(post_conv(<Z>)) -- the activation that follows the layer in the encoder stem, fused into the epilogue
  int rev;  // 1: walk the tiles from the last to the first (the input was just written by the previous kernel in ascending order:
            //    its newest lines are the ones still in the 126 MB L2)
};

constexpr int kFwdStages = 4;
constexpr int kFwdSW = 8;                         // streaming warps
constexpr int kFwdThreads = (kFwdSW + 1) * 32;    // + 1 circuit warp
// x tile columns: the TMA box must start on a 16-byte boundary, so it starts 4 columns before window i0 tap 0 + P
// (P == 1 on the fast path: tap k of local window w sits in column 3 + w*S + k).
template <int S> __host__ __device__ constexpr int fwd_xw() { return S == 1 ? 40 : 72; }

template <int S, int RC>
__host__ __device__ constexpr size_t fast_fwd_smem_bytes(int CK, int O, int Lq) {
  return 1024 + (size_t)kFwdStages * RC * fwd_xw<S>() * 4 +
         ((size_t)CK * FQ + (size_t)O * FQ + align_up(O, 4) + 4 + (size_t)Lq * FQ * kGateStride + 2 * kFwdSW * FTW * FQ +
          2 * FTW * FQ) * 4 +
         (2 * kFwdStages + 8) * 8;
}

// Warp-specialised, mbarrier-coupled pipeline over the tiles of a persistent CTA:
//   warps 0-7 (streaming): pre_conv partial sums of tile n from the TMA ring -> part[n&1]; then post_conv of
//                          tile n-1 (outs[(n-1)&1] -> y, 128-bit coalesced stores).
//   warp  8   (circuit)  : tile n: part[n&1] -> bias, statevector circuit, <Z_i> -> outs[n&1] (+ pre_save/qout_save).
// Hand-offs: pfull/pempty (part), ofull/oempty (outs).  No CTA-wide barrier inside the loop.
// (Specialising this kernel for n_layers == 1 like the adjoint kernel -- 56 -> 48 KB of SASS, 96 -> 72 registers -- changed
// nothing: 131.9 us per step either way.  Its circuit code is one warp's side job, not the whole kernel.)
template <int S, int RC, bool ACT>
__global__ void __launch_bounds__(kFwdThreads) fast_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const FastFwdArgs a) {
  const int Lq = a.Lq;
  constexpr int XW = fwd_xw<S>();
  constexpr int STAGE_ELEMS = RC * XW;
  constexpr int ITS = RC / (4 * kFwdSW);  // kFwdSW warps x 4 rows per iteration
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* base = align1024(smem_dyn);
  float* stage = reinterpret_cast<float*>(base);                    // [kFwdStages][RC][XW]
  float* wpre_t = stage + (size_t)kFwdStages * STAGE_ELEMS;         // [C*3][4]
  const int CK = a.C * 3;
  float* wpost = wpre_t + (size_t)CK * FQ;                          // [O][4]
  float* bpost = wpost + (size_t)a.O * FQ;                          // [O]
  float* bpre = bpost + align_up(a.O, 4);                           // [4]
  float* gates = bpre + 4;                                          // [Lq][4][16]
  float* part = gates + (size_t)Lq * FQ * kGateStride;              // [2][kFwdSW][32][4]
  float* outs = part + 2 * kFwdSW * FTW * FQ;                       // [2][32][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(outs + 2 * FTW * FQ);  // [kFwdStages]
  uint64_t* empty = full + kFwdStages;                              // [kFwdStages]
  uint64_t* pfull = empty + kFwdStages;                             // [2]
  uint64_t* pempty = pfull + 2;                                     // [2]
  uint64_t* ofull = pempty + 2;                                     // [2]
  uint64_t* oempty = ofull + 2;                                     // [2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr = lane >> 3, tl = lane & 7;
  tl_begin(a.tl);

  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_chunks = my_tiles * a.chunks_per_tile;
  const int tile0 = a.rev ? a.num_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x, tstep = a.rev ? -(int)gridDim.x : (int)gridDim.x;
  // producer (thread 0): chunk g of this CTA -> stage g % kFwdStages
  auto issue = [&](int g) {
    const int n = g / a.chunks_per_tile, ch = g - n * a.chunks_per_tile;
    const int tile = tile0 + n * tstep;
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * a.tw;
    const int s = g % kFwdStages;
    mbar_arrive_expect_tx(&full[s], STAGE_ELEMS * 4);
    tma_load_3d(stage + (size_t)s * STAGE_ELEMS, &tm_x, i0 * S - 4, ch * RC, b, &full[s]);
  };
  if (tid == kFwdSW * 32) {  // lane 0 of the circuit warp: it stages no parameters, so it can sit in the dependency wait
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kFwdSW);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], kFwdSW);
      mbar_init(&pempty[s], 1);
      mbar_init(&ofull[s], 1);
      mbar_init(&oempty[s], kFwdSW);
    }
    fence_mbar_init();
    // x may be the previous kernel's output: wait for it here, then get the first tiles moving while the other threads stage
    // the parameters (73 KB of pre_conv weights per CTA at C = 384)
    if (a.early_tma) {
      pdl_wait();
      for (int g = 0; g < kFwdStages - 1 && g < total_chunks; ++g) issue(g);
    }
  }
  // ---- stage parameters (vectorised: 4 features x 4 qubits per step, transposed in registers)
  if (warp < kFwdSW) {
  if ((CK & 3) == 0) {
    for (int u = tid; u < CK / 4; u += kFwdSW * 32) {
      const float4 r0 = ld4(a.w_pre + 0 * (size_t)CK + 4 * u), r1 = ld4(a.w_pre + 1 * (size_t)CK + 4 * u);
      const float4 r2 = ld4(a.w_pre + 2 * (size_t)CK + 4 * u), r3 = ld4(a.w_pre + 3 * (size_t)CK + 4 * u);
      st4(wpre_t + (size_t)(4 * u + 0) * FQ, make_float4(r0.x, r1.x, r2.x, r3.x));
      st4(wpre_t + (size_t)(4 * u + 1) * FQ, make_float4(r0.y, r1.y, r2.y, r3.y));
      st4(wpre_t + (size_t)(4 * u + 2) * FQ, make_float4(r0.z, r1.z, r2.z, r3.z));
      st4(wpre_t + (size_t)(4 * u + 3) * FQ, make_float4(r0.w, r1.w, r2.w, r3.w));
    }
  } else {
    for (int idx = tid; idx < CK * FQ; idx += kFwdSW * 32) {
      const int j = idx / CK, f = idx - j * CK;
      wpre_t[f * FQ + j] = a.w_pre[idx];
    }
  }
  for (int u = tid; u < a.O; u += kFwdSW * 32) st4(wpost + (size_t)u * FQ, ld4(a.w_post + (size_t)u * FQ));
  for (int idx = tid; idx < a.O; idx += kFwdSW * 32) bpost[idx] = a.b_post[idx];
  if (tid < FQ) bpre[tid] = a.b_pre[tid];
  if (tid < Lq * FQ) make_gate<float>(a.qw + tid * 3, gates + tid * kGateStride);
  }
  __syncthreads();
  pdl_wait();    // nothing global is written above
  pdl_launch();
  if (tid == 0 && !a.early_tma)
    for (int g = 0; g < kFwdStages - 1 && g < total_chunks; ++g) issue(g);

  if (warp == kFwdSW) {
    // ======================================================== circuit warp: one lane per window
    for (int n = 0; n < my_tiles; ++n) {
      const int tile = tile0 + n * tstep;
      const int b = tile / a.tiles_per_utt;
      const int i = (tile - b * a.tiles_per_utt) * a.tw + lane;
      const int pb = n & 1, ph = (n >> 1) & 1;
      mbar_wait(&pfull[pb], ph);
      const float* pp = part + (size_t)pb * kFwdSW * FTW * FQ;
      float pre[FQ];
#pragma unroll
      for (int j = 0; j < FQ; ++j) pre[j] = bpre[j];
#pragma unroll
      for (int w = 0; w < kFwdSW; ++w) {
        const float4 pv = ld4(pp + ((size_t)w * FTW + lane) * FQ);
        pre[0] += pv.x; pre[1] += pv.y; pre[2] += pv.z; pre[3] += pv.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pempty[pb]);
      float out[FQ] = {0.f, 0.f, 0.f, 0.f};
      if ((a.dbg & 4) == 0 && lane < a.tw && i < a.Lout) {
        float re[1 << FQ], im[1 << FQ];
        circuit_forward_amp<float, FQ>(pre, gates, Lq, re, im, out);
        const size_t wi = (size_t)b * a.Lout + i;
        if (a.pre_save) st4(a.pre_save + wi * FQ, make_float4(pre[0], pre[1], pre[2], pre[3]));
        if (a.qout_save) st4(a.qout_save + wi * FQ, make_float4(out[0], out[1], out[2], out[3]));
      }
      if (n >= 2) mbar_wait(&oempty[pb], ((n >> 1) - 1) & 1);
      st4(outs + (size_t)pb * FTW * FQ + (size_t)lane * FQ, make_float4(out[0], out[1], out[2], out[3]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&ofull[pb]);
    }
    return;
  }

  // ========================================================== streaming warps
  int g = 0;  // running chunk counter
  for (int n = 0; n <= my_tiles; ++n) {
    // ---- pre_conv partial sums of tile n
    if (n < my_tiles) {
      float acc[4][FQ];
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < FQ; ++j) acc[w][j] = 0.f;
      for (int ch = 0; ch < a.chunks_per_tile; ++ch, ++g) {
        if (tid == 0) {
          const int gn = g + kFwdStages - 1;
          if (gn < total_chunks) {
            if (gn >= kFwdStages) mbar_wait(&empty[gn % kFwdStages], ((gn / kFwdStages) - 1) & 1);
            issue(gn);
          }
        }
        const int s = g % kFwdStages;
        mbar_wait(&full[s], (g / kFwdStages) & 1);
        const float* st = stage + (size_t)s * STAGE_ELEMS;
#pragma unroll
        for (int it = 0; it < ITS; ++it) {
          const int rl = it * (4 * kFwdSW) + warp * 4 + rr;  // row inside the chunk
          const int c = ch * RC + rl;              // channel (rows >= C are zero-filled by TMA; weights masked)
          const float* xr = st + rl * XW;
          float xc[S == 1 ? 6 : 9];
          if constexpr (S == 1) {
            const float4 v0 = ld4(xr + 4 * tl + 4);
            xc[0] = xr[4 * tl + 3];
            xc[1] = v0.x; xc[2] = v0.y; xc[3] = v0.z; xc[4] = v0.w;
            xc[5] = xr[4 * tl + 8];
          } else {
            const float4 v0 = ld4(xr + 8 * tl + 4), v1 = ld4(xr + 8 * tl + 8);
            xc[0] = xr[8 * tl + 3];
            xc[1] = v0.x; xc[2] = v0.y; xc[3] = v0.z; xc[4] = v0.w;
            xc[5] = v1.x; xc[6] = v1.y; xc[7] = v1.z; xc[8] = v1.w;
          }
          if (a.dbg & 1) {
            acc[0][0] += xc[0] + xc[4];
          } else if (c < a.C) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float4 wv = ld4(wpre_t + (size_t)(c * 3 + k) * FQ);
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float xv = xc[w * S + k];
                acc[w][0] = fmaf(wv.x, xv, acc[w][0]);
                acc[w][1] = fmaf(wv.y, xv, acc[w][1]);
                acc[w][2] = fmaf(wv.z, xv, acc[w][2]);
                acc[w][3] = fmaf(wv.w, xv, acc[w][3]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // reduce over the 4 row classes of the warp; the circuit warp sums the kFwdSW warps
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < FQ; ++j) {
          float v = acc[w][j];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          acc[w][j] = v;
        }
      const int pb = n & 1;
      if (n >= 2) mbar_wait(&pempty[pb], ((n >> 1) - 1) & 1);
      if (rr == 0) {
        float* pp = part + (size_t)pb * kFwdSW * FTW * FQ;
#pragma unroll
        for (int w = 0; w < 4; ++w)
          st4(pp + ((size_t)warp * FTW + 4 * tl + w) * FQ, make_float4(acc[w][0], acc[w][1], acc[w][2], acc[w][3]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[pb]);
    }
    // ---- post_conv of tile n-1
    if (n >= 1) {
      const int m = n - 1;
      const int tile = tile0 + m * tstep;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * a.tw;
      const int ob = m & 1;
      mbar_wait(&ofull[ob], (m >> 1) & 1);
      const float* oo = outs + (size_t)ob * FTW * FQ;
      float ov[4][FQ];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float4 v = ld4(oo + (size_t)(4 * tl + w) * FQ);
        ov[w][0] = v.x; ov[w][1] = v.y; ov[w][2] = v.z; ov[w][3] = v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&oempty[ob]);
      if (a.y == nullptr) continue;  // readouts only (fused stem, qw_stem.cu): the circuit warp has stored them
      const bool store_ok = (i0 + 4 * tl) < a.Lout && 4 * tl < a.tw;  // Lout % 4 == 0: a lane's 4 windows are all valid or all invalid
      float* __restrict__ yb = a.y + (size_t)b * a.O * a.Lout + i0 + 4 * tl;
      const int ngroups = a.O >> 2;
#pragma unroll 4
      for (int og = warp; og < ngroups; og += kFwdSW) {
        const int o = og * 4 + rr;
        const float4 wv = ld4(wpost + (size_t)o * FQ);
        const float bv = bpost[o];
        float4 r;
        if (a.dbg & 2) {
          r = make_float4(bv, bv, bv, bv);
        } else {
        r.x = fmaf(wv.w, ov[0][3], fmaf(wv.z, ov[0][2], fmaf(wv.y, ov[0][1], fmaf(wv.x, ov[0][0], bv))));
        r.y = fmaf(wv.w, ov[1][3], fmaf(wv.z, ov[1][2], fmaf(wv.y, ov[1][1], fmaf(wv.x, ov[1][0], bv))));
        r.z = fmaf(wv.w, ov[2][3], fmaf(wv.z, ov[2][2], fmaf(wv.y, ov[2][1], fmaf(wv.x, ov[2][0], bv))));
        r.w = fmaf(wv.w, ov[3][3], fmaf(wv.z, ov[3][2], fmaf(wv.y, ov[3][1], fmaf(wv.x, ov[3][0], bv))));
        }
        if (ACT) {
          r.x = gelu_erf(r.x); r.y = gelu_erf(r.y); r.z = gelu_erf(r.z); r.w = gelu_erf(r.w);
        }
        if (store_ok) st4(yb + (size_t)o * a.Lout, r);
      }
    }
  }
  tl_end(a.tl);
}

// =============================================================================================== backward: gy streaming + adjoint
constexpr int kGyMS = 33;                         // per-lane accumulator column stride
// partial-row layout of the gy kernel: [O*4 grad post_conv.weight][O grad post_conv.bias][pad to 32]
__host__ __device__ inline int gy_plen(int O) { return (int)align_up((size_t)O * 5, 32); }

// =============================================================================================== backward: gy streaming (lean) + adjoint kernel
// Split form of the kernel above (default; QW_GY_FUSED=1 selects the fused one): ncu showed the fused kernel bound by its two
// adjoint warps (a 3 000-instruction dependent chain per tile on one warp) and by instruction-cache misses (98 KB of code
// shared by two roles).  Here the streaming kernel only streams (small code, every warp the same role) and writes
// gout (16 B / window); the adjoint runs as its own kernel with one window per thread across all SMs.
struct FastGy2Args {
  const float* w_post;
  float *gout, *part;  // gout: [B*Lout][4]; part: [gridDim.x][PA1]
  int B, O, Lout, tiles_per_utt, num_tiles, PA1;
  int tw;  // tile stride in windows (the qout box has tw rows)
  unsigned long long* tl;
  const float* b_post = nullptr;  // fused activation (tensor-pipe kernel only)
  int act = 0;
  const float* gen_gpre = nullptr;  // chained backward (tensor-pipe kernel only): see FastGy3Args
  const float* gen_w = nullptr;
  int gen_LP = 0;
};
// NW warps, one thread per stage row (stage = 32*NW output channels x 32 windows), NST-deep ring
__host__ __device__ constexpr size_t fast_gy2_smem_bytes(int O, int NW, int NST) {
  return 1024 + (size_t)NST * (32 * NW * 32) * 4 + (size_t)3 * FTW * FQ * 4 + (size_t)O * FQ * 4 +
         (size_t)NW * FTW * FQ * 4 + (2 * NST) * 8;
}

template <int NHALF, int NW, int NST>
__global__ void __launch_bounds__(32 * NW, 2) fast_bwd_gy2_kernel(const __grid_constant__ CUtensorMap tm_gy,
                                                                    const __grid_constant__ CUtensorMap tm_qout, const FastGy2Args a) {
  constexpr int kGySW = NW, kGyStream = 32 * NW, kGyStages = NST, kGyStageRows = 32 * NW, kGyStageElems = kGyStageRows * 32;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* base = align1024(smem_dyn);
  float* stages = reinterpret_cast<float*>(base);                      // [kGyStages][32*NW][32] swizzled
  float* outs = stages + (size_t)kGyStages * kGyStageElems;            // [3][32][4]
  float* wpost = outs + 3 * FTW * FQ;                                  // [O][4]
  float* gred = wpost + (size_t)a.O * FQ;                              // [kGySW][32][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(gred + kGySW * FTW * FQ);
  uint64_t* empty = full + kGyStages;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr = lane >> 3, tl = lane & 7;
  tl_begin(a.tl);
  if (tid == 0) {
    tma_prefetch_desc(&tm_gy);
    tma_prefetch_desc(&tm_qout);
    for (int s = 0; s < kGyStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGySW);
    }
    fence_mbar_init();
  }
  for (int u = tid; u < a.O; u += kGyStream) st4(wpost + (size_t)u * FQ, ld4(a.w_post + (size_t)u * FQ));
  __syncthreads();
  pdl_wait();
  pdl_launch();

  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_stages = my_tiles * NHALF;
  float wacc[NHALF][FQ + 1];
#pragma unroll
  for (int m = 0; m < NHALF; ++m)
#pragma unroll
    for (int j = 0; j <= FQ; ++j) wacc[m][j] = 0.f;

  auto issue = [&](int gs) {
    const int n = gs / NHALF, h = gs - n * NHALF;
    const int tile = blockIdx.x + n * gridDim.x;
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * a.tw;
    const int s = gs % kGyStages;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(kGyStageElems + (h == 0 ? a.tw * FQ : 0)) * 4);
#pragma unroll
    for (int bx = 0; bx < kGyStageRows / 64; ++bx)
      tma_load_3d(stages + (size_t)s * kGyStageElems + bx * 64 * 32, &tm_gy, i0, h * kGyStageRows + bx * 64, b, &full[s]);
    if (h == 0) tma_load_3d(outs + (n % 3) * FTW * FQ, &tm_qout, 0, i0, b, &full[s]);
  };
  if (tid == 0)
    for (int gs = 0; gs < kGyStages - 1 && gs < total_stages; ++gs) issue(gs);

  int gs = 0;
  const int nchunk = a.tw >> 2;
  for (int n = 0; n < my_tiles; ++n) {
    float gacc[4][FQ];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < FQ; ++j) gacc[u][j] = 0.f;
    const float* os = outs + (n % 3) * FTW * FQ;
#pragma unroll
    for (int h = 0; h < NHALF; ++h, ++gs) {
      if (tid == 0) {
        const int gn = gs + kGyStages - 1;
        if (gn < total_stages) {
          if (gn >= kGyStages) mbar_wait(&empty[gn % kGyStages], ((gn / kGyStages) - 1) & 1);
          issue(gn);
        }
      }
      const int s = gs % kGyStages;
      mbar_wait(&full[s], (gs / kGyStages) & 1);
      const float* gsm = stages + (size_t)s * kGyStageElems;
      const int row0 = h * kGyStageRows;
      // ---- gout partials: lanes (4 rows x 8 chunks of 4 windows); 48 row groups per stage, 8 per warp
#pragma unroll 2
      for (int og = warp; og < kGyStageRows / 4; og += kGySW) {
        const int rl = og * 4 + rr;
        const int r = row0 + rl;
        const float4 gv = ld4(gsm + swz128(rl, tl));
        const float4 wv = (r < a.O) ? ld4(wpost + (size_t)r * FQ) : make_float4(0.f, 0.f, 0.f, 0.f);
        gacc[0][0] = fmaf(gv.x, wv.x, gacc[0][0]); gacc[0][1] = fmaf(gv.x, wv.y, gacc[0][1]);
        gacc[0][2] = fmaf(gv.x, wv.z, gacc[0][2]); gacc[0][3] = fmaf(gv.x, wv.w, gacc[0][3]);
        gacc[1][0] = fmaf(gv.y, wv.x, gacc[1][0]); gacc[1][1] = fmaf(gv.y, wv.y, gacc[1][1]);
        gacc[1][2] = fmaf(gv.y, wv.z, gacc[1][2]); gacc[1][3] = fmaf(gv.y, wv.w, gacc[1][3]);
        gacc[2][0] = fmaf(gv.z, wv.x, gacc[2][0]); gacc[2][1] = fmaf(gv.z, wv.y, gacc[2][1]);
        gacc[2][2] = fmaf(gv.z, wv.z, gacc[2][2]); gacc[2][3] = fmaf(gv.z, wv.w, gacc[2][3]);
        gacc[3][0] = fmaf(gv.w, wv.x, gacc[3][0]); gacc[3][1] = fmaf(gv.w, wv.y, gacc[3][1]);
        gacc[3][2] = fmaf(gv.w, wv.z, gacc[3][2]); gacc[3][3] = fmaf(gv.w, wv.w, gacc[3][3]);
      }
      // ---- grad post_conv.{weight,bias}: thread <-> stage row, accumulators live in registers
      {
        const int rl = tid;  // 0..191
#pragma unroll 2
        for (int c = 0; c < nchunk; ++c) {  // only the tile's own tw windows count towards the weight gradients
          const float4 gv = ld4(gsm + swz128(rl, c));
          const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 ov = ld4(os + (size_t)(4 * c + u) * FQ);
            wacc[h][0] = fmaf(g4[u], ov.x, wacc[h][0]);
            wacc[h][1] = fmaf(g4[u], ov.y, wacc[h][1]);
            wacc[h][2] = fmaf(g4[u], ov.z, wacc[h][2]);
            wacc[h][3] = fmaf(g4[u], ov.w, wacc[h][3]);
            wacc[h][4] += g4[u];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // ---- reduce the 4 row classes, then the 6 warps through shared memory; one warp stores the tile's gout
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < FQ; ++j) {
        float v = gacc[u][j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        gacc[u][j] = v;
      }
    // single gred buffer: the warp that summed tile n-1 out of it ARRIVES on named barrier 2 when it is done reading, every
    // other warp SYNCS on it before overwriting (32 + 32 * (NW - 1) = all threads); the reader knows its own read is over
    float* gr = gred;
    if (n >= 1 && warp != (n - 1) % kGySW) asm volatile("bar.sync 2, %0;" ::"n"(kGyStream) : "memory");
    if (rr == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        st4(gr + ((size_t)warp * FTW + 4 * tl + u) * FQ, make_float4(gacc[u][0], gacc[u][1], gacc[u][2], gacc[u][3]));
    }
    __syncthreads();  // gred complete
    if (warp == n % kGySW) {
      const int tile = blockIdx.x + n * gridDim.x;
      const int b = tile / a.tiles_per_utt;
      const int i = (tile - b * a.tiles_per_utt) * a.tw + lane;
      float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kGySW; ++w) {
        const float4 pv = ld4(gr + ((size_t)w * FTW + lane) * FQ);
        sacc.x += pv.x; sacc.y += pv.y; sacc.z += pv.z; sacc.w += pv.w;
      }
      if (n + 1 < my_tiles) asm volatile("bar.arrive 2, %0;" ::"n"(kGyStream) : "memory");
      if (lane < a.tw && i < a.Lout) st4(a.gout + ((size_t)b * a.Lout + i) * FQ, sacc);
    }
  }
  // ---- partial row of this CTA: [O*4 grad post_conv.weight][O grad post_conv.bias][pad]
  float* prow = a.part + (size_t)blockIdx.x * a.PA1;
#pragma unroll
  for (int m = 0; m < NHALF; ++m) {
    const int r = m * kGyStageRows + tid;
    if (r < a.O) {
      st4(prow + (size_t)r * FQ, make_float4(wacc[m][0], wacc[m][1], wacc[m][2], wacc[m][3]));
      prow[a.O * FQ + r] = wacc[m][4];
    }
  }
  for (int e = a.O * (FQ + 1) + tid; e < a.PA1; e += kGyStream) prow[e] = 0.f;
  tl_end(a.tl);
}

// =============================================================================================== backward: gy streaming on the tensor pipe
// fast_bwd_gy2_kernel is bound by instruction issue, not by HBM: 7 500 warp-instructions per 48 KB tile (ncu, round 1: FFMA 40 %,
// issue slots 41 % busy at 3 warps per scheduler), i.e. 3.2 us per round of 296 tiles against 2.2 us of HBM time.  Both of its
// contractions are rank-4 products of the streamed tile with a small matrix:
//     grad post_conv.weight[o][j] += sum_w gy[o][w] * <Z_j>[w]        (K = windows,  N = 4)
//     gout[w][j]                   = sum_o gy[o][w] * W_post[o][j]    (K = channels, N = 4)
// so they fit the warp-level tensor-core instruction mma.sync.m16n8k8 (TF32 in, fp32 accumulate; SASS HMMA.1688.F32.TF32,
// measured 8.7 clk per sub-partition on B200, tools/probe/mma_probe.cu = 4.6 x the FFMA pipe per MAC and ONE issue slot per
// 1 024 MACs) with the N = 8 columns holding [hi | lo] of the small matrix: a "3 x TF32" product.  The streamed operand is
// split in registers, v = hi + lo with hi = v & 0xffffe000 (exact) and lo = v - hi (exact in fp32, <= 13 significant bits, the
// tensor core keeps 11 of them), and both halves are multiplied with [hi | lo]: the dropped terms are 2^-21 relative, the
// accumulation is fp32 -- inside the 1e-5 budget (tests: same tolerances as the FFMA kernel).  Per 128 elements of the tile
// and contraction: 2 LDS.64 + 8 split ops + 2-4 MMA instead of ~46 LDS/FFMA/FADD: ~3 000 instructions per tile.
//
// Fragment <-> tile mapping (g = lane >> 2, t = lane & 3; the tile is SWIZZLE_128B, rows = channels, 32 windows per row):
//   weight gradient (computed transposed, the streamed tile is the B operand): B[k][n] = gy[8 ob + n][w(k)],
//                    w(t) = 4 kb + 2 (t & 1) + 16 (t >> 1), w(t + 4) = w(t) + 1, so b0/b1 are ONE conflict-free LDS.64 and need no
//                    register shuffling; A[m][k] = rows 0-3 hi(<Z_m>[w(k)]), rows 4-7 lo(<Z_{m-4}>[w(k)]), row 8 = 1 (the bias
//                    gradient rides along for free), rows 9-15 = 0.
//   gout:            A[m][k] = gy[8 kb + r(k)][16 mb + w(m)], r(t) = 2 t, r(t + 4) = 2 t + 1, w(g) = 2 g, w(g + 8) = 2 g + 1
//                    (again two conflict-free LDS.64); B[k][n] = [hi | lo] of W_post[8 kb + r(k)][n & 3], held in registers for the
//                    CTA lifetime (each warp owns 3 of the 24 channel blocks of a stage; partial sums of the 8 warps meet in
//                    shared memory, double-buffered, one __syncthreads per tile).
constexpr int kGy3Warps = 8, kGy3Threads = 32 * kGy3Warps, kGy3Rows = 192, kGy3Stages = 3;
constexpr int kGy3Bpw = kGy3Rows / 8 / kGy3Warps;  // 8-channel blocks of a stage per warp (24 blocks over 8 warps: an even load on the 4 sub-partitions)
static_assert(kGy3Bpw * kGy3Warps * 8 == kGy3Rows, "stage rows must split evenly over the warps");
__host__ __device__ constexpr size_t fast_gy3_smem_bytes(int gen_O = 0) {
  return 1024 + (size_t)kGy3Stages * kGy3Rows * 32 * 4 + (size_t)3 * FTW * FQ * 4 + (size_t)2 * kGy3Warps * FTW * 8 * 4 +
         (size_t)gen_O * 3 * FQ * 4 +            // GEN form: the following layer's pre_conv weights, [O*3][4]
         (3 * kGy3Stages + 2 * 6 + 1) * 8;  // barrier block laid out as in the fused forms
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

// ---- fused small-batch forms (FUSED = 1, 2): the adjoint (and for a data layer the pre_conv^T contraction) ride in the gy kernel.
// At batch 16 a CTA of the persistent grid sees only 3-6 tiles, the backward kernels of a layer run 5-20 us each, and a
// zero-compute TMA streaming kernel of the same footprint needs ~15 us for the 74 MB of gy (tools/probe/stream_probe.cu): launch
// ramp, the ragged last round and the drain of every kernel boundary cost as much as the bytes, while the gy pass itself leaves
// most issue slots idle (it waits for HBM).  So the adjoint differentiation moves INTO those idle slots:
//   warps 0-7  the tensor-pipe gy pass above over a CONTIGUOUS range of <= kFbMaxTiles tiles; gout stays in shared memory
//              (never written to HBM); <Z> and pre_conv outputs of each tile arrive by TMA on a per-tile barrier;
//   warp 8     adjoint warp, one window per lane: the forward recomputation of tile n runs as soon as its pre_conv outputs have
//              landed (i.e. while the gy rows of that tile are still streaming), the backward sweep as soon as the tile's gout
//              is summed.  Gate-gradient matrices accumulate in registers over the CTA's tiles and are reduced once (transposing
//              warp reduction, 31 shuffles for 32 values).  Only the last tile's backward sweep (~1 650 instructions) is exposed.
//              (Round 1 tried two specialised adjoint warps next to an FFMA gy pass and lost: that pass was issue-bound, so the
//              adjoint warps and the streaming warps fought for the same slots.)
//   FUSED = 1  data layer (stride 1, no grad_x: conv1 of the stem): gpre stays in shared memory and the streaming warps finish
//              with grad pre_conv.weight[j][c][k] += gpre[w][j] x[c][w + k - 1] on x tiles requested when the gy ring drains
//              (threads = (c, k) pairs, register accumulators) -- the whole backward of the layer is this kernel + finalize.
//   FUSED = 2  any layer: gpre goes to the halo-padded workspace array the pre_conv^T kernel reads (the adjoint kernel is gone).
// One partial row per CTA and role, reduced by the finalize kernel as before.
constexpr int kFbMaxTiles = 6;   // tiles per CTA the shared-memory slots cover
constexpr int kFbXW = 40;        // x tile columns (as the forward kernel: tap k of local window w = column 3 + w + k)
constexpr int kFbXRows = 96;     // x tile rows (channels; C <= 96)
__host__ __device__ constexpr size_t fast_gy3_fused_smem_bytes() {
  return 1024 + (size_t)kGy3Stages * kGy3Rows * 32 * 4 + (size_t)4 * kFbMaxTiles * FTW * FQ * 4 + (size_t)2 * kGy3Warps * FTW * 8 * 4 +
         (size_t)FQ * kGateStride * 4 + (3 * kGy3Stages + 2 * kFbMaxTiles + 1) * 8;
}
struct FastGy3Args {
  const float *w_post, *qw;
  float *gout, *part;       // plain: gout [B*Lout][4]; part: [grid][PA1]
  float *part2, *part3;     // fused: [grid][PA2], [grid][PB]
  float* gpre_pad;          // FUSED = 2: [B][LP][4]
  int B, C, O, Lout, LP, tiles_per_utt, num_tiles, PA1, PA2, PB;
  unsigned long long* tl;
  int dbg;  // experiment switch DBG_GY: 1 = skip the contractions (streaming floor of the kernel); results are garbage
  const float* b_post;  // act != 0 only
  int act;  // 1: the incoming gradient is that of gelu(post_conv(<Z>)): every stage is multiplied by gelu'(post_conv(<Z>)) in place first
  int rev;  // 1 (plain form): tiles from the last to the first (gy was just written in ascending order by the next layer's pre_conv^T)
  // GEN form: the incoming gradient is NOT read -- it is the grad_x of the FOLLOWING QuantumConv1d (kernel_size 3, stride 2,
  // padding 1, in_channels = this O, input length = this L_out), rebuilt tile by tile from that layer's gpre rows (16 bytes per
  // window, halo-padded, L2-resident) and its pre_conv weights; the following layer's pre_conv^T then never writes grad_x.
  const float* gen_gpre = nullptr;  // [B][gen_LP][4], window i of the following layer at row kHaloL + i
  const float* gen_w = nullptr;     // the following layer's pre_conv.weight (4, O*3)
  int gen_LP = 0;
};
struct RegGateAcc {
  float m[FQ][8];
  __device__ __forceinline__ void set_layer(int) {}
  __device__ __forceinline__ void add(int wire, const float (&mm)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) m[wire][e] += mm[e];
  }
};
// v[32] per lane -> lane L returns sum over the warp of v[L]  (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int ofs = 16; ofs >= 1; ofs >>= 1) {
    const bool up = (lane & ofs) != 0;
#pragma unroll
    for (int i = 0; i < ofs; ++i) {
      const float send = up ? v[i] : v[i + ofs];
      const float keep = up ? v[i + ofs] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
    }
  }
  return v[0];
}
__device__ __forceinline__ void bar_sync_streaming() { asm volatile("bar.sync 1, %0;" ::"n"(kGy3Threads) : "memory"); }

template <int NHALF, int FUSED, bool ACT, bool GEN>
__global__ void __launch_bounds__(FUSED ? kGy3Threads + 32 : kGy3Threads, 2)
    fast_bwd_gy3_kernel(const __grid_constant__ CUtensorMap tm_gy, const __grid_constant__ CUtensorMap tm_qout,
                        const __grid_constant__ CUtensorMap tm_pre, const __grid_constant__ CUtensorMap tm_x, const FastGy3Args a) {
  constexpr int SE = kGy3Rows * 32;  // floats per stage
  constexpr int NQT = FUSED ? kFbMaxTiles : 3;  // <Z> tiles kept (fused: all of the CTA's tiles)
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* base = align1024(smem_dyn);
  float* stages = reinterpret_cast<float*>(base);                      // [kGy3Stages][192][32] swizzled
  float* outs = stages + (size_t)kGy3Stages * SE;                      // [NQT][32][4]   <Z> of the tiles' windows
  float* pre_s = outs + NQT * FTW * FQ;                                // fused: [kFbMaxTiles][32][4] pre_conv outputs
  float* gout_s = pre_s + (FUSED ? kFbMaxTiles * FTW * FQ : 0);        // fused: [kFbMaxTiles][32][4]
  float* gpre_s = gout_s + (FUSED ? kFbMaxTiles * FTW * FQ : 0);       // fused: [kFbMaxTiles][32][4]
  float* gred = gpre_s + (FUSED ? kFbMaxTiles * FTW * FQ : 0);         // [2][6 warps][32 windows][8]
  float* gates = gred + 2 * kGy3Warps * FTW * 8;                       // fused: [4][16]
  float* w2t = gates + (FUSED ? FQ * kGateStride : 0);                 // GEN: [O*3][4] pre_conv weights of the following layer
  uint64_t* full = reinterpret_cast<uint64_t*>(w2t + (GEN ? (size_t)a.O * 3 * FQ : 0));
  uint64_t* empty = full + kGy3Stages;
  uint64_t* xfull = empty + kGy3Stages;                                // FUSED = 1: x tiles of the pre_conv^T phase
  uint64_t* pqfull = xfull + kGy3Stages;                               // fused: [kFbMaxTiles] <Z> + pre_conv outputs of tile n landed
  uint64_t* gfull = pqfull + kFbMaxTiles;                              // fused: [kFbMaxTiles] gout of tile n summed
  uint64_t* gdone = gfull + kFbMaxTiles;                               // fused: the adjoint warp is through

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  tl_begin(a.tl);
  if (tid == 0) {
    tma_prefetch_desc(&tm_gy);
    tma_prefetch_desc(&tm_qout);
    if (FUSED) tma_prefetch_desc(&tm_pre);
    if (FUSED == 1) tma_prefetch_desc(&tm_x);
    for (int s = 0; s < kGy3Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGy3Warps);
      if (FUSED) mbar_init(&xfull[s], 1);
    }
    if (FUSED) {
      for (int n = 0; n < kFbMaxTiles; ++n) {
        mbar_init(&pqfull[n], 1);
        mbar_init(&gfull[n], 1);
      }
      mbar_init(gdone, 1);
    }
    fence_mbar_init();
  }
  if (FUSED && tid < FQ) make_gate<float>(a.qw + tid * 3, gates + tid * kGateStride);

  // plain: round-robin tiles (neighbouring CTAs stream neighbouring row segments); fused: a contiguous range per CTA
  int tile0, tstep, my_tiles;
  if (FUSED) {
    tile0 = (int)(((long long)blockIdx.x * a.num_tiles) / gridDim.x);
    my_tiles = (int)(((long long)(blockIdx.x + 1) * a.num_tiles) / gridDim.x) - tile0;
    tstep = 1;
  } else {
    tile0 = a.rev ? a.num_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x;
    tstep = a.rev ? -(int)gridDim.x : (int)gridDim.x;
    my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  }

  if (FUSED && warp == kGy3Warps) {
    // ========================================================================= adjoint warp: one window per lane
    __syncthreads();
    pdl_wait();
    RegGateAcc acc;
#pragma unroll
    for (int w = 0; w < FQ; ++w)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc.m[w][e] = 0.f;
    float gb[FQ] = {0.f, 0.f, 0.f, 0.f};
    for (int n = 0; n < my_tiles; ++n) {
      const int tile = tile0 + n;
      const int b = tile / a.tiles_per_utt;
      const int i = (tile - b * a.tiles_per_utt) * FTW + lane;
      const bool valid = i < a.Lout;
      mbar_wait(&pqfull[n], 0);
      const float4 pv = valid ? ld4(pre_s + ((size_t)n * FTW + lane) * FQ) : make_float4(1.f, 0.f, 0.f, 0.f);
      const float pre[FQ] = {pv.x, pv.y, pv.z, pv.w};
      float out[FQ], gpre[FQ], re[1 << FQ], im[1 << FQ];
      const float inv = circuit_forward_amp<float, FQ>(pre, gates, 1, re, im, out);
      mbar_wait(&gfull[n], 0);
      const float4 gv = valid ? ld4(gout_s + ((size_t)n * FTW + lane) * FQ) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float gout[FQ] = {gv.x, gv.y, gv.z, gv.w};
      circuit_backward_amp<float, FQ>(pre, inv, gates, 1, re, im, gout, gpre, acc);
#pragma unroll
      for (int j = 0; j < FQ; ++j) gb[j] += gpre[j];
      if (FUSED == 1) st4(gpre_s + ((size_t)n * FTW + lane) * FQ, make_float4(gpre[0], gpre[1], gpre[2], gpre[3]));
      if (FUSED == 2 && valid) st4(a.gpre_pad + ((size_t)b * a.LP + kHaloL + i) * FQ, make_float4(gpre[0], gpre[1], gpre[2], gpre[3]));
    }
    if (FUSED == 1) {
      __syncwarp();
      if (lane == 0) mbar_arrive(gdone);  // gpre_s of every tile is in shared memory
    }
    // partial row of this CTA: [0,4) grad pre_conv.bias, [32,64) gate-gradient matrices, everything else padding
    float mv[32];
#pragma unroll
    for (int w = 0; w < FQ; ++w)
#pragma unroll
      for (int e = 0; e < 8; ++e) mv[w * 8 + e] = acc.m[w][e];
    const float msum = warp_transpose_sum32(mv, lane);
    float* prow2 = a.part2 + (size_t)blockIdx.x * a.PA2;
    float bsum = 0.f;
#pragma unroll
    for (int j = 0; j < FQ; ++j) {
      const float s_ = warp_sum(gb[j]);
      if (lane == j) bsum = s_;
    }
    prow2[lane] = lane < FQ ? bsum : 0.f;
    prow2[32 + lane] = msum;
    for (int e = 64 + lane; e < a.PA2; e += 32) prow2[e] = 0.f;
    if (a.tl && lane == 0) atomicMax(a.tl + 1, global_ns());  // the adjoint warp may be the last one out
    return;
  }

  // B fragments of the gout contraction (parameters: readable before the dependency wait)
  uint32_t bw[NHALF][kGy3Bpw][2];
#pragma unroll
  for (int h = 0; h < NHALF; ++h)
#pragma unroll
    for (int i = 0; i < kGy3Bpw; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int o = h * kGy3Rows + (warp * kGy3Bpw + i) * 8 + 2 * t + e;
        const float v = (o < a.O) ? __ldg(a.w_post + (size_t)o * FQ + (g & 3)) : 0.f;
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        bw[h][i][e] = g < 4 ? hi : lo;
      }
  if (GEN) {  // (4, O*3) -> [O*3][4], 4 features x 4 qubits per step (a parameter: readable before the dependency wait)
    const int CK2 = a.O * 3;
    for (int u = tid; u < CK2 / 4; u += kGy3Threads) {
      const float4 r0 = ld4(a.gen_w + 0 * (size_t)CK2 + 4 * u), r1 = ld4(a.gen_w + 1 * (size_t)CK2 + 4 * u);
      const float4 r2 = ld4(a.gen_w + 2 * (size_t)CK2 + 4 * u), r3 = ld4(a.gen_w + 3 * (size_t)CK2 + 4 * u);
      st4(w2t + (size_t)(4 * u + 0) * FQ, make_float4(r0.x, r1.x, r2.x, r3.x));
      st4(w2t + (size_t)(4 * u + 1) * FQ, make_float4(r0.y, r1.y, r2.y, r3.y));
      st4(w2t + (size_t)(4 * u + 2) * FQ, make_float4(r0.z, r1.z, r2.z, r3.z));
      st4(w2t + (size_t)(4 * u + 3) * FQ, make_float4(r0.w, r1.w, r2.w, r3.w));
    }
  }
  // lane offsets (floats) inside a stage
  int off1[4], off2[2][2];
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) off1[kb] = g * 32 + (((kb + 4 * (t >> 1)) ^ g) << 2) + 2 * (t & 1);
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int e = 0; e < 2; ++e) off2[m][e] = (2 * t + e) * 32 + (((4 * m + (g >> 1)) ^ (2 * t + e)) << 2) + 2 * (g & 1);
  __syncthreads();
  pdl_wait();
  pdl_launch();

  if (FUSED == 2) {
    // zero the halos of gpre_pad (left kHaloL and right kHaloR windows of every utterance): the pre_conv^T kernel reads them
    const int per = (kHaloL + kHaloR) * FQ;
    for (long long idx = (long long)blockIdx.x * kGy3Threads + tid; idx < (long long)a.B * per; idx += (long long)gridDim.x * kGy3Threads) {
      const int b = (int)(idx / per), e = (int)(idx - (long long)b * per);
      const int off = e < kHaloL * FQ ? e : (kHaloL + a.Lout) * FQ + (e - kHaloL * FQ);
      a.gpre_pad[(size_t)b * a.LP * FQ + off] = 0.f;
    }
  }

  const int total_stages = my_tiles * NHALF;
  float d1[NHALF * kGy3Bpw][4];  // [8-channel block][fragment]: rows = hi/lo of <Z_j> (and the ones row), columns = channels
#pragma unroll
  for (int m = 0; m < NHALF * kGy3Bpw; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) d1[m][j] = 0.f;
  const uint32_t one_g0 = (g == 0) ? 0x3f800000u : 0u;  // A rows 8..15: row 8 (lanes g == 0) = 1.0f

  auto issue = [&](int gs) {
    const int n = gs / NHALF, h = gs - n * NHALF;
    const int tile = tile0 + n * tstep;
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * FTW;
    const int s = gs % kGy3Stages;
    mbar_arrive_expect_tx(&full[s], (uint32_t)((GEN ? 0 : SE) + ((h == 0 && !FUSED) ? FTW * FQ : 0)) * 4);
    if (!GEN) {
#pragma unroll
      for (int bx = 0; bx < kGy3Rows / 64; ++bx)
        tma_load_3d(stages + (size_t)s * SE + bx * 64 * 32, &tm_gy, i0, h * kGy3Rows + bx * 64, b, &full[s]);
    }
    if (h == 0) {
      if (FUSED) {
        mbar_arrive_expect_tx(&pqfull[n], (uint32_t)(2 * FTW * FQ) * 4);
        tma_load_3d(outs + n * FTW * FQ, &tm_qout, 0, i0, b, &pqfull[n]);
        tma_load_3d(pre_s + n * FTW * FQ, &tm_pre, 0, i0, b, &pqfull[n]);
      } else {
        tma_load_3d(outs + (n % NQT) * FTW * FQ, &tm_qout, 0, i0, b, &full[s]);
      }
    }
  };
  if (tid == 0)
    for (int gs = 0; gs < kGy3Stages - 1 && gs < total_stages; ++gs) issue(gs);

  int gs = 0;
  for (int n = 0; n < my_tiles; ++n) {
    float d2[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) d2[m][j] = 0.f;
    uint32_t aq[4][4];
    const float* os = outs + (n % NQT) * FTW * FQ;
    float4 gq[3];  // GEN: gpre rows of the following layer's windows m, m + 1, m + 2 that touch this thread's four positions
    if (GEN) {
      const int tile = tile0 + n * tstep;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * FTW;
      const float4* gp = reinterpret_cast<const float4*>(a.gen_gpre) + ((size_t)b * a.gen_LP + kHaloL + (i0 >> 1) + 2 * (tid & 7));
#pragma unroll
      for (int e = 0; e < 3; ++e) gq[e] = __ldg(gp + e);  // rows past the utterance are the zeroed halo: positions >= L_out come out 0
    }
#pragma unroll
    for (int h = 0; h < NHALF; ++h, ++gs) {
      if (tid == 0) {
        const int gn = gs + kGy3Stages - 1;
        if (gn < total_stages) {
          if (gn >= kGy3Stages) mbar_wait(&empty[gn % kGy3Stages], ((gn / kGy3Stages) - 1) & 1);
          // the slot was rewritten in place by generic-proxy stores (fused activation) that the empty barrier has ordered before
          // this point: one proxy fence in the issuing thread orders them before the TMA engine's refill
          if (ACT && !GEN) fence_proxy_async();
          issue(gn);
        }
      }
      const int s = gs % kGy3Stages;
      mbar_wait(&full[s], (gs / kGy3Stages) & 1);
      const float* gsm = stages + (size_t)s * SE;
      if (h == 0) {
        if (FUSED) mbar_wait(&pqfull[n], 0);
        // A fragments of the weight-gradient contraction from this tile's <Z>
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const int w0 = 4 * kb + 2 * (t & 1) + 16 * (t >> 1);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            uint32_t hi, lo;
            split_tf32(os[(w0 + e) * FQ + (g & 3)], hi, lo);
            aq[kb][2 * e] = g < 4 ? hi : lo;
            aq[kb][2 * e + 1] = one_g0;
          }
        }
      }
      if (GEN) {
        // The stage is COMPUTED, not loaded: gy[r][l] = grad_x of the following layer = sum_k sum_j w2[j][r*3+k] gpre2[i(l,k)][j]
        // with l = 2 i - 1 + k (stride 2, padding 1).  A thread owns the positions l0 .. l0 + 3, l0 = i0 + 4 c = 2 m:
        //   l0 (even): tap 1 of window m;  l0+1: tap 0 of m+1 and tap 2 of m;  l0+2: tap 1 of m+1;  l0+3: tap 0 of m+2 and tap 2 of m+1.
        // 24 FMAs + three LDS.128 of weights per 16 bytes of the tile instead of a 73.7 MB tensor written by one kernel and read by
        // the next.  full[s] only guards the <Z> rows here; it cannot complete before every warp has released the slot's previous
        // contents (the issuing thread waits for empty[s] first), so the slot is free to write.  With ACT the gelu' factor of THIS
        // layer is applied before the store (no separate in-place pass).
        float* gmut = stages + (size_t)s * SE;
        constexpr int NIT = kGy3Rows * 8 / kGy3Threads;
        const int c = tid & 7;
        float4 q4[4];
        if (ACT) {
#pragma unroll
          for (int e = 0; e < 4; ++e) q4[e] = ld4(os + (4 * c + e) * FQ);
        }
        auto dot4 = [](const float4& w, const float4& g_) { return fmaf(w.w, g_.w, fmaf(w.z, g_.z, fmaf(w.y, g_.y, w.x * g_.x))); };
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int rl = (tid >> 3) + it * (kGy3Threads / 8);
          const int r = h * kGy3Rows + rl;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < a.O) {
            const float* wr = w2t + (size_t)(r * 3) * FQ;
            const float4 W0 = ld4(wr), W1 = ld4(wr + FQ), W2 = ld4(wr + 2 * FQ);
            v.x = dot4(W1, gq[0]);
            v.y = dot4(W0, gq[1]) + dot4(W2, gq[0]);
            v.z = dot4(W1, gq[1]);
            v.w = dot4(W0, gq[2]) + dot4(W2, gq[1]);
            if (ACT) {
              const float4 wv = __ldg(reinterpret_cast<const float4*>(a.w_post) + r);
              const float bv = __ldg(a.b_post + r);
              float* ve = reinterpret_cast<float*>(&v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float z = fmaf(wv.w, q4[e].w, fmaf(wv.z, q4[e].z, fmaf(wv.y, q4[e].y, fmaf(wv.x, q4[e].x, bv))));
                ve[e] *= gelu_erf_grad(z);
              }
            }
          }
          st4(gmut + swz128(rl, c), v);
        }
        __syncthreads();
      } else if (ACT) {
        // fused activation backward: g <- g * gelu'(z), z = post_conv(<Z>) rebuilt from the 16 bytes per window the forward saved
        // (the pre-activation itself is never stored).  In place on the stage, one pass, 128-byte rows; then everybody syncs.
        // Measured on B200 (batch 16, conv1): this pass takes the kernel from 28 to 70 us -- 27 instructions per element at the
        // ~0.4 IPC these kernels run at (the arithmetic of gelu' is only 9 us of it; the rest is the loads, z and the in-place store)
        // -- against ~45 us for ATen's separate gelu_backward pass (two reads + a write of the tensor) plus the 28 us kernel.
        float* gmut = stages + (size_t)s * SE;
        // a thread keeps its 16-byte chunk column (c = tid & 7: the same four windows in every row it visits), so the <Z> of those
        // windows are loaded once per stage, and the post_conv rows it needs (straight from global memory, 7.7 KB, L1-resident:
        // keeping them in shared memory pushed the CTA past 98 KB) are all requested before the first is used
        constexpr int NIT = kGy3Rows * 8 / kGy3Threads;  // 6 rows per thread and stage
        const int c = tid & 7;
        float4 q4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) q4[e] = ld4(os + (4 * c + e) * FQ);
        float4 wv[NIT];
        float bv[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int r = h * kGy3Rows + (tid >> 3) + it * (kGy3Threads / 8);
          wv[it] = r < a.O ? __ldg(reinterpret_cast<const float4*>(a.w_post) + r) : make_float4(0.f, 0.f, 0.f, 0.f);
          bv[it] = r < a.O ? __ldg(a.b_post + r) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int rl = (tid >> 3) + it * (kGy3Threads / 8);
          if (h * kGy3Rows + rl < a.O) {
            float* p = gmut + swz128(rl, c);
            float4 gv = ld4(p);
            float* ge = reinterpret_cast<float*>(&gv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = fmaf(wv[it].w, q4[e].w, fmaf(wv[it].z, q4[e].z, fmaf(wv[it].y, q4[e].y, fmaf(wv[it].x, q4[e].x, bv[it]))));
              ge[e] *= gelu_erf_grad(z);
            }
            st4(p, gv);
          }
        }
        __syncthreads();
      }
      if (a.dbg) {
        d2[0][0] += gsm[tid];
      } else {
      // ---- grad post_conv.{weight,bias}: this warp's four 8-channel blocks of the stage, K = the tile's 32 windows
#pragma unroll
      for (int i = 0; i < kGy3Bpw; ++i) {
        const float* rp = gsm + (warp * kGy3Bpw + i) * 256;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const float2 u = ld2(rp + off1[kb]);
          uint32_t bh[2], bl[2];
          split_tf32(u.x, bh[0], bl[0]);
          split_tf32(u.y, bh[1], bl[1]);
          mma_tf32(d1[h * kGy3Bpw + i], aq[kb], bh[0], bh[1]);
          mma_tf32(d1[h * kGy3Bpw + i], aq[kb], bl[0], bl[1]);
        }
      }
      // ---- gout partials: this warp's four 8-channel blocks of the stage, both 16-window halves of the tile
#pragma unroll
      for (int i = 0; i < kGy3Bpw; ++i) {
        const float* rp = gsm + (warp * kGy3Bpw + i) * 256;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const float2 u = ld2(rp + off2[m][0]), v = ld2(rp + off2[m][1]);
          uint32_t ah[4], al[4];
          split_tf32(u.x, ah[0], al[0]);
          split_tf32(u.y, ah[1], al[1]);
          split_tf32(v.x, ah[2], al[2]);
          split_tf32(v.y, ah[3], al[3]);
          mma_tf32(d2[m], ah, bw[h][i][0], bw[h][i][1]);
          mma_tf32(d2[m], al, bw[h][i][0], bw[h][i][1]);
        }
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // ---- the 8 warps' partial gout rows meet in shared memory; warp n % 8 sums and stores the tile's gout.  Double-buffered:
    // a warp can only be writing buffer (n + 2) & 1 after the barrier of tile n + 1, which the summing warp of tile n reaches
    // after its reads.
    float* gr = gred + (size_t)(n & 1) * kGy3Warps * FTW * 8 + (size_t)warp * FTW * 8;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      st2(gr + (16 * m + 2 * g) * 8 + 2 * t, make_float2(d2[m][0], d2[m][1]));
      st2(gr + (16 * m + 2 * g + 1) * 8 + 2 * t, make_float2(d2[m][2], d2[m][3]));
    }
    if (FUSED) bar_sync_streaming(); else __syncthreads();
    if (warp == n % kGy3Warps) {
      const float* gq = gred + (size_t)(n & 1) * kGy3Warps * FTW * 8 + (size_t)lane * 8;
      float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kGy3Warps; ++w) {
        const float4 p0 = ld4(gq + (size_t)w * FTW * 8), p1 = ld4(gq + (size_t)w * FTW * 8 + 4);
        sacc.x += p0.x + p1.x; sacc.y += p0.y + p1.y; sacc.z += p0.z + p1.z; sacc.w += p0.w + p1.w;
      }
      if (FUSED) {
        st4(gout_s + ((size_t)n * FTW + lane) * FQ, sacc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&gfull[n]);  // hand the tile to the adjoint warp
      } else {
        const int tile = tile0 + n * tstep;
        const int b = tile / a.tiles_per_utt;
        const int i = (tile - b * a.tiles_per_utt) * FTW + lane;
        if (i < a.Lout) st4(a.gout + ((size_t)b * a.Lout + i) * FQ, sacc);
      }
    }
  }
  // ---- partial row of this CTA: [O*4 grad post_conv.weight][O grad post_conv.bias][pad]
  {
    float* prow = a.part + (size_t)blockIdx.x * a.PA1;
#pragma unroll
    for (int m = 0; m < NHALF * kGy3Bpw; ++m) {
      const int h = m / kGy3Bpw, i = m - h * kGy3Bpw;
      const int o = h * kGy3Rows + (warp * kGy3Bpw + i) * 8 + 2 * t;  // this lane's two channels: o, o + 1
      // rows j (hi half of <Z_j>) and j + 4 (lo half) live in lanes g and g + 4
      const float c0 = d1[m][0] + __shfl_xor_sync(0xffffffffu, d1[m][0], 16);
      const float c1 = d1[m][1] + __shfl_xor_sync(0xffffffffu, d1[m][1], 16);
      if (g < 4) {
        if (o < a.O) prow[(size_t)o * FQ + g] = c0;
        if (o + 1 < a.O) prow[(size_t)(o + 1) * FQ + g] = c1;
      }
      if (g == 0) {  // row 8 = the ones row: grad post_conv.bias
        if (o < a.O) prow[a.O * FQ + o] = d1[m][2];
        if (o + 1 < a.O) prow[a.O * FQ + o + 1] = d1[m][3];
      }
    }
    for (int e = a.O * (FQ + 1) + tid; e < a.PA1; e += kGy3Threads) prow[e] = 0.f;
  }
  if constexpr (FUSED == 1) {
    // =========================================================================== grad pre_conv.weight (streaming warps)
    bar_sync_streaming();  // every streaming warp is done with the gy ring: it now takes the x tiles
    const int xtile_elems = kFbXRows * kFbXW;
    auto issue_x = [&](int n) {
      const int tile = tile0 + n;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * FTW;
      const int s = n % kGy3Stages;
      mbar_arrive_expect_tx(&xfull[s], (uint32_t)xtile_elems * 4);
      tma_load_3d(stages + (size_t)s * SE, &tm_x, i0 - 4, 0, b, &xfull[s]);
    };
    if (tid == 0)
      for (int n = 0; n < kGy3Stages && n < my_tiles; ++n) issue_x(n);
    // thread <-> (channel c, tap k): p = c * 3 + k; slot 1 covers p >= 192 (C <= 96: at most 288 pairs)
    const int nck = a.C * 3;
    const int p0 = tid, p1 = tid + kGy3Threads;
    const int c0 = p0 / 3, k0 = p0 - 3 * c0, c1 = p1 / 3, k1 = p1 - 3 * c1;
    const bool v0 = p0 < nck, v1 = p1 < nck;
    float acc0[FQ] = {0.f, 0.f, 0.f, 0.f}, acc1[FQ] = {0.f, 0.f, 0.f, 0.f};
    mbar_wait(gdone, 0);  // gpre of every tile is in shared memory
    for (int n = 0; n < my_tiles; ++n) {
      const int s = n % kGy3Stages;
      mbar_wait(&xfull[s], (n / kGy3Stages) & 1);
      const float* xs = stages + (size_t)s * SE;
      const float* x0 = xs + (v0 ? c0 * kFbXW + 3 + k0 : 0);
      const float* x1 = xs + (v1 ? c1 * kFbXW + 3 + k1 : 0);
      const float* gp = gpre_s + (size_t)n * FTW * FQ;
#pragma unroll 8
      for (int w = 0; w < FTW; ++w) {
        const float4 gq = ld4(gp + w * FQ);
        const float xa = x0[w], xb = x1[w];
        acc0[0] = fmaf(gq.x, xa, acc0[0]); acc0[1] = fmaf(gq.y, xa, acc0[1]);
        acc0[2] = fmaf(gq.z, xa, acc0[2]); acc0[3] = fmaf(gq.w, xa, acc0[3]);
        acc1[0] = fmaf(gq.x, xb, acc1[0]); acc1[1] = fmaf(gq.y, xb, acc1[1]);
        acc1[2] = fmaf(gq.z, xb, acc1[2]); acc1[3] = fmaf(gq.w, xb, acc1[3]);
      }
      if (n + kGy3Stages < my_tiles) {
        bar_sync_streaming();  // every thread is done with slot s
        if (tid == 0) issue_x(n + kGy3Stages);
      }
    }
    float* prow3 = a.part3 + (size_t)blockIdx.x * a.PB;
    if (v0) {
#pragma unroll
      for (int j = 0; j < FQ; ++j) prow3[c0 * 12 + j * 3 + k0] = acc0[j];
    }
    if (v1) {
#pragma unroll
      for (int j = 0; j < FQ; ++j) prow3[c1 * 12 + j * 3 + k1] = acc1[j];
    }
    for (int e = nck * FQ + tid; e < a.PB; e += kGy3Threads) prow3[e] = 0.f;
  }
  tl_end(a.tl);
}

// adjoint differentiation of the circuit, one window per thread; partial row per CTA: [gb_pre 4 + pad 28][Lq*32 gate matrices]
//
// The forward recomputation of a window needs only pre_save (written by the forward kernel, at least two launches back in the
// stream), so the dependency wait sits INSIDE the loop, between the recomputation and the first read of gout: under programmatic
// dependent launch these small CTAs are resident microseconds before the gy kernel drains, and ~40 % of the dependent chain of
// their first window is done by then.  One copy of the circuit code only -- the kernel is a ~1 650-instruction straight line per
// window and sensitive to instruction-cache misses (a peeled first iteration, i.e. two inlined copies, cost 3.5 us per launch).
//
// Measured and rejected (B200, batch-16 stem step): folding the finalize kernel's work into this kernel and the pre_conv^T
// kernel as leading "reduce" CTAs (grad post_conv.* here, grad pre_conv.bias / quantum_weights there), and reducing the
// pre_conv^T rows by a two-level last-arriver scheme.  The first is +1 us (the reduce CTAs delay the dependent launch trigger and
// grow the code of two cache-sensitive kernels), the second +8 us (fence + atomic + L2 round trips at the tail of every CTA)
// against the programmatically launched 1 024-thread finalize kernel, which costs 3-4 us per layer in the graph.
constexpr int kAdjThreads = 128;

// MULTI = false: the reference circuit (one layer); the n_layers loop of the circuit code folds away: 64 -> 35 KB of SASS,
// 128 -> 96 registers, and the batch-16 step drops from 135.3 to 132.1 us (QW_ADJ_SPEC=0 selects the general kernel for A/B)
template <bool MULTI>
__global__ void __launch_bounds__(kAdjThreads, 4) fast_bwd_adj_kernel(const FastAdjArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  const int Lq = MULTI ? a.Lq : 1;
  const int NE = FQ + Lq * 32;
  float* gates = reinterpret_cast<float*>(smem_dyn);             // [Lq][4][16]
  float* macc = gates + (size_t)Lq * FQ * kGateStride;         // [4 warps][NE][kGyMS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  tl_begin(a.tl);
  if (tid < Lq * FQ) make_gate<float>(a.qw + tid * 3, gates + tid * kGateStride);
  for (int e = tid; e < 4 * NE * kGyMS; e += kAdjThreads) macc[e] = 0.f;
  __syncthreads();
  float* mymacc = macc + (size_t)warp * NE * kGyMS;
  bool waited = false;
  auto wait_once = [&]() {
    if (waited) return;
    waited = true;
    pdl_wait();
    if (a.early_trigger) pdl_launch();
    // zero the halos of gpre_pad (left kHaloL and right kHaloR windows of every utterance)
    const int per = (kHaloL + kHaloR) * FQ;
    for (long long idx = (long long)blockIdx.x * kAdjThreads + tid; idx < (long long)a.B * per; idx += (long long)gridDim.x * kAdjThreads) {
      const int b = (int)(idx / per), e = (int)(idx - (long long)b * per);
      const int off = e < kHaloL * FQ ? e : (kHaloL + a.Lout) * FQ + (e - kHaloL * FQ);
      a.gpre_pad[(size_t)b * a.LP * FQ + off] = 0.f;
    }
  };
  for (long long w0 = (long long)blockIdx.x * kAdjThreads; w0 < a.W; w0 += (long long)gridDim.x * kAdjThreads) {
    const long long w = w0 + tid;
    const bool valid = w < a.W;
    const float4 pv = valid ? ld4(a.pre_save + (size_t)w * FQ) : make_float4(1.f, 0.f, 0.f, 0.f);
    const float pre[FQ] = {pv.x, pv.y, pv.z, pv.w};
    float out[FQ], gpre[FQ], re[1 << FQ], im[1 << FQ];
    const float inv = circuit_forward_amp<float, FQ>(pre, gates, Lq, re, im, out);
    wait_once();
    const float4 gv = valid ? ld4(a.gout + (size_t)w * FQ) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float gout[FQ] = {gv.x, gv.y, gv.z, gv.w};
    SmemGateAcc<float, FQ> acc{mymacc + (size_t)FQ * kGyMS + lane, kGyMS, 0};
    circuit_backward_amp<float, FQ>(pre, inv, gates, Lq, re, im, gout, gpre, acc);
    if (valid) {
      const int b = (int)(w / a.Lout), i = (int)(w - (long long)b * a.Lout);
      st4(a.gpre_pad + ((size_t)b * a.LP + kHaloL + i) * FQ, make_float4(gpre[0], gpre[1], gpre[2], gpre[3]));
    }
#pragma unroll
    for (int j = 0; j < FQ; ++j) mymacc[j * kGyMS + lane] += valid ? gpre[j] : 0.f;
  }
  wait_once();   // a CTA without windows still owes the halo zeroing
  if (!a.early_trigger) pdl_launch();  // late (QW_ADJ_TRIG=0): an early trigger (measured: +0.3 us) lets the next kernel's CTAs crowd the SMs this latency-bound grid needs
  __syncthreads();
  float* prow = a.part + (size_t)blockIdx.x * a.PA2;
  for (int e = warp; e < a.PA2; e += kAdjThreads / 32) {
    // e in [0,4): grad pre_conv.bias; [32, 32+Lq*32): gate matrices; everything else padding
    float v = 0.f;
    const int src = e < FQ ? e : (e >= 32 && e < 32 + Lq * 32) ? FQ + (e - 32) : -1;
    if (src >= 0) {
#pragma unroll
      for (int wq = 0; wq < 4; ++wq) v += macc[((size_t)wq * NE + src) * kGyMS + lane];
      v = warp_sum(v);
    }
    if (lane == 0) prow[e] = v;
  }
  tl_end(a.tl);
}

// =============================================================================================== backward: pre_conv^T
struct FastPreArgs {
  const float *gpre_pad, *w_pre;
  float* part;  // [gridPx][Cpad][12]
  int B, C, L, P, Lout, LP, tiles_per_utt, num_tiles, Cpad;
  int gridPx;  // streaming CTAs per channel chunk; the grid is 1-D: [nchunks x gridPx]
  int early_x;
  unsigned long long* tl;
};
constexpr int kPreSlots = 3;
template <int S> __host__ __device__ constexpr int pre_gpn() { return S == 1 ? 136 : 72; }  // gpre rows staged per tile

template <int S>
__host__ __device__ constexpr size_t fast_pre_smem_bytes() {
  return 1024 + (size_t)kPreSlots * 4096 * 4 + (size_t)kPreSlots * pre_gpn<S>() * FQ * 4 + kPreSlots * 8;
}

template <int S, int PAR, bool GX>
__global__ void __launch_bounds__(kThreads) fast_bwd_pre_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                const __grid_constant__ CUtensorMap tm_gx, const FastPreArgs a) {
  constexpr int GPN = pre_gpn<S>();
  constexpr int NWG = S == 1 ? 6 : 4;  // gpre rows touched by 4 consecutive positions
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* base = align1024(smem_dyn);
  float* xs = reinterpret_cast<float*>(base);                  // [kPreSlots][4 boxes][32 rows][32] swizzled
  float* gps = xs + (size_t)kPreSlots * 4096;                   // [kPreSlots][GPN][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(gps + (size_t)kPreSlots * GPN * FQ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  tl_begin(a.tl);
  const int lin = (int)blockIdx.x;
  const int by = lin / a.gridPx, bx = lin - by * a.gridPx;  // channel chunk, CTA inside the chunk
  const int c0 = by * 32, c = c0 + lane;

  if (tid == 0) {
    tma_prefetch_desc(&tm_x);
    if (GX) tma_prefetch_desc(&tm_gx);
    for (int s = 0; s < kPreSlots; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  float w[FQ][3], gw[FQ][3];
#pragma unroll
  for (int j = 0; j < FQ; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      w[j][k] = (c < a.C) ? a.w_pre[(size_t)j * a.C * 3 + c * 3 + k] : 0.f;
      gw[j][k] = 0.f;
    }
  const int my_tiles = (bx < a.num_tiles) ? (a.num_tiles - 1 - bx) / a.gridPx + 1 : 0;
  // a tile = its x boxes (this layer's forward input: long complete) + its gpre rows (the adjoint kernel's output)
  auto issue_x = [&](int n) {
    const int tile = bx + n * a.gridPx;
    const int b = tile / a.tiles_per_utt;
    const int l0 = (tile - b * a.tiles_per_utt) * 128;
    const int s = n % kPreSlots;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(4096 + GPN * FQ) * 4);
    for (int q = 0; q < 4; ++q) tma_load_3d(xs + (size_t)s * 4096 + q * 1024, &tm_x, l0 + q * 32, c0, b, &full[s]);
  };
  auto issue_g = [&](int n) {
    const int tile = bx + n * a.gridPx;
    const int b = tile / a.tiles_per_utt;
    const int l0 = (tile - b * a.tiles_per_utt) * 128;
    const int i_lo = floor_div(l0 + a.P - 2, S);
    const int s = n % kPreSlots;
    bulk_g2s(gps + (size_t)s * GPN * FQ, a.gpre_pad + ((size_t)b * a.LP + kHaloL + i_lo) * FQ, GPN * FQ * 4, &full[s]);
  };
  auto issue = [&](int n) {
    issue_x(n);
    issue_g(n);
  };
  // the x boxes of the first tiles are requested BEFORE the dependency wait (they overlap the adjoint kernel's tail)
  if (tid == 0 && a.early_x)
    for (int n = 0; n < kPreSlots - 1 && n < my_tiles; ++n) issue_x(n);
  __syncthreads();
  pdl_wait();
  pdl_launch();
  if (tid == 0)
    for (int n = 0; n < kPreSlots - 1 && n < my_tiles; ++n) {
      if (!a.early_x) issue_x(n);
      issue_g(n);
    }

  for (int n = 0; n < my_tiles; ++n) {
    const int s = n % kPreSlots;
    if (tid == 0 && n + kPreSlots - 1 < my_tiles) {
      // slot (n-1) % kPreSlots: its grad_x stores were committed at the end of the previous iteration
      if (GX) bulk_wait_read<0>();
      issue(n + kPreSlots - 1);
    }
    mbar_wait(&full[s], (n / kPreSlots) & 1);
    float* xb = xs + (size_t)s * 4096 + warp * 1024;  // this warp's 32-position box
    const float* gp = gps + (size_t)s * GPN * FQ;
#pragma unroll 2
    for (int cc = 0; cc < 8; ++cc) {
      const int off = swz128(lane, cc);
      const float4 xv4 = ld4(xb + off);
      const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
      const int wb = (S == 1) ? (warp * 32 + cc * 4) : (warp * 16 + cc * 2);
      float4 g[NWG];
#pragma unroll
      for (int q = 0; q < NWG; ++q) g[q] = ld4(gp + (size_t)(wb + q) * FQ);
      float gx[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          // window index relative to wb: stride 1: p - k + 2; stride 2: (p + PAR - k)/2 + 1 when even
          const int num = p + PAR - k;
          const bool valid = (S == 1) || ((num & 1) == 0);
          if (valid) {
            const int q = (S == 1) ? (p - k + 2) : (num / 2 + 1);
            const float4 gq = g[q];
            gw[0][k] = fmaf(gq.x, xv[p], gw[0][k]);
            gw[1][k] = fmaf(gq.y, xv[p], gw[1][k]);
            gw[2][k] = fmaf(gq.z, xv[p], gw[2][k]);
            gw[3][k] = fmaf(gq.w, xv[p], gw[3][k]);
            if (GX) acc = fmaf(gq.w, w[3][k], fmaf(gq.z, w[2][k], fmaf(gq.y, w[1][k], fmaf(gq.x, w[0][k], acc))));
          }
        }
        gx[p] = acc;
      }
      if (GX) st4(xb + off, make_float4(gx[0], gx[1], gx[2], gx[3]));
    }
    if (GX) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        const int tile = bx + n * a.gridPx;
        const int b = tile / a.tiles_per_utt;
        const int l0 = (tile - b * a.tiles_per_utt) * 128;
        for (int q = 0; q < 4; ++q) tma_store_3d(&tm_gx, l0 + q * 32, c0, b, xs + (size_t)s * 4096 + q * 1024);
        bulk_commit();
      }
    } else {
      __syncthreads();  // slot reuse: every warp is done reading before thread 0 re-issues into it
    }
  }
  if (GX && tid == 0) bulk_wait_all<0>();
  __syncthreads();
  // ---- cross-warp reduction of the weight-gradient partials (reuses the tile memory)
  float* red = xs;  // [kWarps][32][12]
#pragma unroll
  for (int j = 0; j < FQ; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) red[((size_t)warp * 32 + lane) * 12 + j * 3 + k] = gw[j][k];
  __syncthreads();
  float* prow = a.part + ((size_t)bx * a.Cpad + c0) * 12;
  for (int e = tid; e < 32 * 12; e += kThreads) {
    float sum = 0.f;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) sum += red[(size_t)wq * 32 * 12 + e];
    prow[e] = sum;
  }
  tl_end(a.tl);
}

// =============================================================================================== finalize
// Deterministic fixed-order fp64 reduction of the per-CTA partial rows of the three kernels above: segment 1 = rows of the gy
// kernel (grad post_conv.*), 2 = rows of the adjoint kernel (grad pre_conv.bias, gate matrices -> grad quantum_weights),
// 3 = rows of the pre_conv^T kernel (grad pre_conv.weight).
// 1 024 threads (32 warps per 32-column block: short dependent chains over the partial rows) for the plain finalize; 512 when
// the gradient all-reduce rides in it, so the whole grid is resident in ONE wave (a CTA holds its SM slot while warp 0 waits
// for the peers: with one 1 024-thread CTA per SM the 208 CTAs of conv2 paid the NVLink round trip twice)
struct FastFinArgs {
  const float *part1, *part2, *part3, *qw;
  float *gw_pre, *gb_pre, *gqw, *gw_post, *gb_post;
  int G1, P1, G2, P2, G3, P3;
  int C, O, Lq;
  unsigned long long* tl;
  FastDp dp;  // world <= 1: plain finalize
  int early12;
};

// Data-parallel training fuses the gradient all-reduce INTO this kernel ("fuse the intra-GPU reduction into that
// kernel's epilogue so the all-reduce input is ready without an extra pass" -- here the all-reduce itself rides in the epilogue).
// Warp 0 of every CTA holds the CTA's 32 reduced columns.  Each lane packs (epoch << 32 | float bits) into ONE 64-bit word and
// stores it straight into slot [epoch parity][my rank][column] of EVERY peer's receive buffer over NVLink (posted P2P stores),
// then polls its own buffer until the words of all ranks carry this epoch, and sums them in fixed rank order -- bitwise
// identical on every rank.  Data and flag travel in the same atomic 8-byte store (the "LL" idea of NCCL's low-latency
// protocol), so there is no fence, no separate flag and no remote read: one NVLink write latency per CTA.  Measured on 2 x
// B200: the first version (data in my own buffer, __threadfence_system, release-store of a flag into the peers, acquire-spin,
// remote loads) spent 3-6 us in EACH system-scope fence and 19 us per exchange in total.
// The raw column sums are exchanged (gate-gradient matrices included: the (phi, theta, omega) chain rule is linear in them and
// is applied after the sum).  The epoch lives in device memory (CUDA-graph capturable); parity double-buffering is enough
// because a rank can only be one epoch ahead of its slowest reader (it cannot finish epoch e + 1 before the slowest rank has
// posted e + 1, i.e. has finished reading e).  A CTA only ever waits for the SAME CTA index of its peers and the whole grid is
// resident, so there is no inter-CTA deadlock.  Like NCCL the kernel WAITS for a late peer; a peer that stays away for
// dp.timeout_ns (default 10 min, option DP_TIMEOUT_MS) trips a trap -- a loud launch failure, never a silently local gradient.
__device__ __forceinline__ void dp_st_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dp_ld_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// called by warp 0 (all 32 lanes); t = this lane's column sum; returns the mean over ranks
__device__ __forceinline__ double dp_allreduce_columns(const FastDp& dp, double t, int lane) {
  const int blk = blockIdx.x, nblk = gridDim.x;
  unsigned* myflags = dp.flags[dp.rank];  // epoch[nblk], status
  unsigned epoch = 0;
  if (lane == 0) epoch = myflags[blk] + 1u;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const size_t slot_base = (size_t)(epoch & 1u) * dp.world;
  const size_t col = (size_t)blk * 32 + lane;
  const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint((float)t);
#pragma unroll
  for (int r = 0; r < kDpMaxWorld; ++r)
    if (r < dp.world) dp_st_u64(reinterpret_cast<unsigned long long*>(dp.bufs[r]) + (slot_base + dp.rank) * dp.ncol + col, word);
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(dp.bufs[dp.rank]);
  double sum = 0.0;
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < dp.world; ++r) {
    const unsigned long long* src = mine + (slot_base + r) * dp.ncol + col;
    unsigned long long v = dp_ld_u64(src);
    unsigned spins = 0;
    while ((unsigned)(v >> 32) != epoch) {
      // NCCL semantics: WAIT for the peer (a late rank -- checkpointing, evaluation, a data-loader stall -- is normal).  Only a
      // peer that stays away for dp.timeout_ns (default 10 min, like the NCCL watchdog; 0 = forever) is an error, and then the
      // kernel traps: the context dies with a launch failure, it never continues with an un-averaged gradient.
      if ((++spins & 1023u) == 0 && dp.timeout_ns != 0 && global_ns() - t0 > dp.timeout_ns) {
        myflags[nblk] = 1u;
        __threadfence_system();
        __trap();
      }
      v = dp_ld_u64(src);
    }
    sum += (double)__uint_as_float((unsigned)v);
  }
  if (lane == 0) myflags[blk] = epoch;
  return sum * (double)dp.scale;
}

template <int kFFThreads>
__global__ void __launch_bounds__(kFFThreads) fast_finalize_kernel(const FastFinArgs a) {
  constexpr int kFFWarps = kFFThreads / 32;
  __shared__ double red[kFFWarps][33];
  __shared__ double tot[32];
  tl_begin(a.tl);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb1 = a.P1 / 32, nb2 = a.P2 / 32;
  const bool seg1 = (int)blockIdx.x < nb1;
  const bool seg2 = !seg1 && (int)blockIdx.x < nb1 + nb2;
  // Segments 1 and 2 reduce the rows of the gy and adjoint kernels.  Those grids completed before this kernel's immediate
  // predecessor (pre_conv^T) passed ITS dependency wait, which is before it triggered this launch -- so their rows are final
  // and visible already, and these CTAs (the lowest block indices: scheduled first, as pre_conv^T CTAs retire) do their whole
  // job, all-reduce included, while pre_conv^T is still streaming.  Only segment 3 waits.
  if (!(a.early12 && (seg1 || seg2))) pdl_wait();
  pdl_launch();
  const int blk = seg1 ? blockIdx.x : seg2 ? blockIdx.x - nb1 : blockIdx.x - nb1 - nb2;
  const int G = seg1 ? a.G1 : seg2 ? a.G2 : a.G3;
  const int P = seg1 ? a.P1 : seg2 ? a.P2 : a.P3;
  const float* __restrict__ part = seg1 ? a.part1 : seg2 ? a.part2 : a.part3;
  const int blk0 = blk * 32, p = blk0 + lane;
  double s = 0.0;
  if (p < P) {
    int g = warp;
    for (; g + 3 * kFFWarps < G; g += 4 * kFFWarps) {
      const float v0 = part[(size_t)g * P + p], v1 = part[(size_t)(g + kFFWarps) * P + p];
      const float v2 = part[(size_t)(g + 2 * kFFWarps) * P + p], v3 = part[(size_t)(g + 3 * kFFWarps) * P + p];
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; g < G; g += kFFWarps) s += (double)part[(size_t)g * P + p];
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp != 0) return;
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kFFWarps; ++w) t += red[w][lane];
  if (a.dp.world > 1) t = dp_allreduce_columns(a.dp, t, lane);
  tl_end(a.tl);
  tot[lane] = t;
  __syncwarp();
  if (seg1) {
    const int nW = a.O * FQ;
    if (p < nW) a.gw_post[p] = (float)t;
    else if (p < nW + a.O) a.gb_post[p - nW] = (float)t;
  } else if (seg2) {
    if (blk0 == 0) {
      if (lane < FQ) a.gb_pre[lane] = (float)t;
    } else if (lane < 4) {
      const int gi = (blk0 - 32) / 8 + lane;
      if (gi < a.Lq * FQ) {
        double w3[3], g3[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) w3[e] = (double)a.qw[gi * 3 + e];
        gate_grad_to_angles(w3, &tot[lane * 8], g3);
#pragma unroll
        for (int e = 0; e < 3; ++e) a.gqw[gi * 3 + e] = (float)g3[e];
      }
    }
  } else {
    const int cch = p / 12, r = p - cch * 12, j = r / 3, k = r - j * 3;
    if (p < P && cch < a.C) a.gw_pre[(size_t)j * a.C * 3 + cch * 3 + k] = (float)t;
  }
}

// =============================================================================================== host
TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return (TmapEncodeFn)p;
  }();
  return fn;
}

int make_tmap_3d_f32(CUtensorMap* tm, const void* base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                     unsigned box0, unsigned box1, bool swizzle128) {
  TmapEncodeFn fn = tmap_encode_fn();
  QW_CHECK_ARG(fn != nullptr, -2, "cuTensorMapEncodeTiled is not available from this driver");
  // the driver API needs the primary context bound to THIS thread (autograd runs backward on its own threads)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    QW_CUDA_OK(cudaFree(nullptr));
    ctx_bound = true;
  }
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstr[2] = {d0 * 4ull, d0 * d1 * 4ull};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QW_CHECK_ARG(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// A/B switches: qw::option() (qw_common.cuh; defaults = the measured winners on B200, batch-16 stem step)
static int flag_fwd_etma() { return option(kOptFwdEtma); }
static int flag_fin_early() { return option(kOptFinEarly); }
static int flag_adj_trig() { return option(kOptAdjTrig); }
static int flag_pre_ex() { return option(kOptPreEx); }
static int flag_gy_mma() { return option(kOptGyMma); }
static int flag_bwd_fused() { return option(kOptBwdFused); }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool fast_eligible(const ConvDims& d, const void* x, const void* y_or_gy, const void* gx, bool fwd) {
  return option(kOptFastPath) && (!fwd || d.P == 1) && d.Q == 4 && d.K == 3 && (d.S == 1 || d.S == 2) && d.emb == kEmbAmplitude && d.L % 4 == 0 &&
         d.Lout % 4 == 0 && d.O % 4 == 0 && d.O <= 576 && d.Lq <= 4 && d.C * 3 * FQ * 4 <= 96 * 1024 && aligned16(x) && aligned16(y_or_gy) &&
         aligned16(gx) && tmap_encode_fn() != nullptr;
}

// NOTE (measured, round 1): sizing the persistent grids for EQUAL tile counts (e.g. 251 CTAs x 3 tiles instead of 296 CTAs of
// which 160 run a third tile) is SLOWER (step 0.1435 -> 0.1525 ms): the kernels are bound by per-CTA pipeline latency, not by
// shared HBM bandwidth, so an SM left with one CTA loses more than the ragged last round costs.  Likewise a 4-lanes-per-window
// adjoint (shorter dependency chain, 4x the warps) lost to one window per thread (17.6 vs 10.9 us): shuffles + the CNOT pass
// through shared memory outweigh the chain shortening at q = 4.  Packed fp32x2 FMAs (FFMA2, __ffma2_rn) in the two contractions
// of fast_bwd_gy2_kernel: no change (26.3 / 19.0 us vs 26.7 / 19.1 us) -- FFMA2 issues at ~0.42x the FFMA rate
// (profiles/r1_fp32_issue_probe.txt) and the kernel is bound by shared-memory / barrier latency at 3 warps per scheduler, not by
// FMA issue slots.
FastPlan make_fast_plan(const ConvDims& d) {
  FastPlan p{};
  const int sms = num_sms();
  // tile stride = tile width.  (Round 1 measured narrower strides with the same 32-window boxes: 32 -> 140.1 us, 28 -> 140.2, 24 -> 154.4
  // per step: the bytes of a round do not shrink with the stride, so the experiment switch is gone.)
  p.tw = FTW;
  p.tiles_per_utt = (d.Lout + p.tw - 1) / p.tw;
  p.num_tiles = d.B * p.tiles_per_utt;
  p.rc = d.C <= 128 ? (int)align_up(d.C, 32) : 64;  // small C: the whole channel range is one TMA box per tile
  p.chunks_per_tile = (d.C + p.rc - 1) / p.rc;
  {
    const size_t smem = 1024 + (size_t)kFwdStages * p.rc * (d.S == 1 ? 40 : 72) * 4 +
                        ((size_t)d.C * 3 * FQ + (size_t)d.O * 5 + 2 * kFwdSW * FTW * FQ + 512) * 4;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : per_sm > 2 ? 2 : per_sm;  // 288 threads x 96 registers: two CTAs per SM
    p.gridF = p.num_tiles < per_sm * sms ? p.num_tiles : per_sm * sms;
  }
  p.gridGy = p.num_tiles < 2 * sms ? p.num_tiles : 2 * sms;
  // programmatic dependent launch while a CTA sees <= 24 tiles: launch / ramp latency still matters (measured with the pre-wait
  // work of this round: batch 16 / 32 / 64 gain 4 / 2.5 / 2.8 %, batch 256 -- 80 tiles per CTA -- loses 2.3 %)
  p.small = p.num_tiles <= 24 * 2 * sms;
  p.PA1 = gy_plen(d.O);
  const long long W = (long long)d.B * d.Lout;
  {
    // one window per thread while the grid fits (measured: 2 windows per thread on half the CTAs is +1 us at batch 16: the
    // second window runs from a warm instruction cache in ~1.5 us, but the chain of the first is what the step waits for)
    const long long need = (W + kAdjThreads - 1) / kAdjThreads;
    const long long cap = (long long)sms * 8;
    p.gridAdj = (int)(need < cap ? need : cap);
  }
  p.PA2 = 32 + (int)align_up((size_t)d.Lq * 32, 32);
  p.LP = kHaloL + d.Lout + kHaloR;
  p.ptiles_per_utt = (d.L + 127) / 128;
  p.num_ptiles = d.B * p.ptiles_per_utt;
  p.nchunks = (d.C + 31) / 32;
  p.Cpad = p.nchunks * 32;
  // CTAs per SM of the pre_conv^T kernel: 4 x 52 KB of shared memory fit; measured at batch 16: 2 -> 133.0, 3 -> 127.9, 4 -> 126.9 us
  const int pre_ctas = option(kOptPreCtas) > 0 ? option(kOptPreCtas) : 4;
  int cap = pre_ctas * sms / p.nchunks;
  if (cap < 1) cap = 1;
  p.gridPx = p.num_ptiles < cap ? p.num_ptiles : cap;
  p.PB = p.Cpad * 12;
  size_t o = 0;
  p.off_gout = o; o = align_up(o + (size_t)W * FQ * 4, 256);
  p.off_gpre = o; o = align_up(o + (size_t)d.B * p.LP * FQ * 4, 256);
  p.off_p1 = o;   o = align_up(o + (size_t)p.gridGy * p.PA1 * 4, 256);
  p.off_p2 = o;   o = align_up(o + (size_t)(p.gridAdj > p.gridGy ? p.gridAdj : p.gridGy) * p.PA2 * 4, 256);  // fused backward: one row per gy CTA
  p.off_p3 = o;   o = align_up(o + (size_t)(p.gridPx > p.gridGy ? p.gridPx : p.gridGy) * p.PB * 4, 256);  // fused backward: one row per gy CTA
  p.ws_bytes = o;
  return p;
}

template <int S, int RC, bool ACT>
static int launch_fast_fwd_t(const CUtensorMap& tm, const FastFwdArgs& a, const FastPlan& p, cudaStream_t st) {
  const size_t smem = fast_fwd_smem_bytes<S, RC>(a.C * 3, a.O, a.Lq);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "fast forward needs %zu bytes of shared memory", smem);
  auto k = fast_fwd_kernel<S, RC, ACT>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_symbol(kKFwd, "fast_fwd_kernel<%d, %d, %d>", S, RC, (int)ACT);
  {
    KernelTimer kt(kKFwd, st);
    QW_CUDA_OK(launch_pdl(p.small, k, dim3(p.gridF), dim3(kFwdThreads), smem, st, tm, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int S, int RC>
static int launch_fast_fwd(const CUtensorMap& tm, const FastFwdArgs& a, const FastPlan& p, cudaStream_t st) {
  return a.act ? launch_fast_fwd_t<S, RC, true>(tm, a, p, st) : launch_fast_fwd_t<S, RC, false>(tm, a, p, st);
}

int fast_forward(const float* x, const float* w_pre, const float* b_pre, const float* qwts, const float* w_post,
                 const float* b_post, float* y, float* pre_save, const ConvDims& d, cudaStream_t st, int act) {
  const FastPlan p = make_fast_plan(d);
  alignas(64) CUtensorMap tm;
  const int xw = d.S == 1 ? fwd_xw<1>() : fwd_xw<2>();
  if (int e = make_tmap_3d_f32(&tm, x, d.L, d.C, d.B, xw, p.rc, false)) return e;
  const size_t W = (size_t)d.B * d.Lout;
  FastFwdArgs a{w_pre, b_pre, qwts, w_post, b_post, y, pre_save, pre_save ? pre_save + W * FQ : nullptr,
                d.B, d.C, d.L, d.P, d.O, d.Lq, d.Lout, p.tiles_per_utt, p.num_tiles, p.chunks_per_tile, p.tw, flag_fwd_etma(), timeline_next_slot(), option(kOptDbgFwd), act, option(kOptRevTiles) & 1};
  if (d.S == 1) {
    switch (p.rc) {
      case 32: return launch_fast_fwd<1, 32>(tm, a, p, st);
      case 96: return launch_fast_fwd<1, 96>(tm, a, p, st);
      case 128: return launch_fast_fwd<1, 128>(tm, a, p, st);
      default: return launch_fast_fwd<1, 64>(tm, a, p, st);
    }
  }
  switch (p.rc) {
    case 32: return launch_fast_fwd<2, 32>(tm, a, p, st);
    case 96: return launch_fast_fwd<2, 96>(tm, a, p, st);
    case 128: return launch_fast_fwd<2, 128>(tm, a, p, st);
    default: return launch_fast_fwd<2, 64>(tm, a, p, st);
  }
}

template <int NHALF, int NW, int NST>
static int launch_fast_gy2(const CUtensorMap& tg, const CUtensorMap& tq, const FastGy2Args& a, const FastPlan& p, cudaStream_t st) {
  const size_t smem = fast_gy2_smem_bytes(a.O, NW, NST);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "fast backward(gy) needs %zu bytes of shared memory", smem);
  auto k = fast_bwd_gy2_kernel<NHALF, NW, NST>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_symbol(kKBwdPost, "fast_bwd_gy2_kernel<%d, %d, %d>", NHALF, NW, NST);
  {
    KernelTimer kt(kKBwdPost, st);
    QW_CUDA_OK(launch_pdl(p.small, k, dim3(p.gridGy), dim3(32 * NW), smem, st, tg, tq, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
// 6 warps x 192-row stages, 3-deep ring (default).  QW_GY_WARPS=12 selects 12 warps x 384-row stages with a 2-deep ring (the
// whole Whisper tile in one stage, twice the warps per scheduler): measured SLOWER on B200 at batch 16 (step 138.0 vs 135.4 us,
// gy 14.7 / 23.2 vs 12.9 / 21.8 us) -- the extra warps do not raise the issue rate of the two contractions, the 2-deep ring
// loses a stage of prefetch, and the fatter CTAs leave less room for the early-resident adjoint CTAs.  Also measured and dropped:
// 3 CTAs per SM of the 6-warp form with a 2-deep ring (444 CTAs): 139.2 us.
template <int NHALF, int FUSED, bool ACT, bool GEN = false>
static int launch_fast_gy3_t(const CUtensorMap& tg, const CUtensorMap& tq, const CUtensorMap& tp, const CUtensorMap& tx, const FastGy3Args& a,
                             const FastPlan& p, cudaStream_t st) {
  const size_t smem = FUSED ? fast_gy3_fused_smem_bytes() : fast_gy3_smem_bytes(GEN ? a.O : 0);
  QW_CHECK_ARG(smem <= 113 * 1024, -2, "chained backward: %zu bytes of shared memory per CTA (out_channels too large)", smem);
  auto k = fast_bwd_gy3_kernel<NHALF, FUSED, ACT, GEN>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_symbol(FUSED ? kKBwdFused : kKBwdPost, "fast_bwd_gy3_kernel<%d, %d, %d, %d>", NHALF, FUSED, (int)ACT, (int)GEN);
  {
    KernelTimer kt(FUSED ? kKBwdFused : kKBwdPost, st);
    QW_CUDA_OK(launch_pdl(p.small, k, dim3(p.gridGy), dim3(FUSED ? kGy3Threads + 32 : kGy3Threads), smem, st, tg, tq, tp, tx, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int NHALF, int FUSED>
static int launch_fast_gy3(const CUtensorMap& tg, const CUtensorMap& tq, const CUtensorMap& tp, const CUtensorMap& tx, const FastGy3Args& a,
                           const FastPlan& p, cudaStream_t st) {
  if constexpr (FUSED == 0) {
    if (a.gen_gpre) return a.act ? launch_fast_gy3_t<NHALF, 0, true, true>(tg, tq, tp, tx, a, p, st) : launch_fast_gy3_t<NHALF, 0, false, true>(tg, tq, tp, tx, a, p, st);
    if (a.act) return launch_fast_gy3_t<NHALF, 0, true>(tg, tq, tp, tx, a, p, st);
  }
  return launch_fast_gy3_t<NHALF, FUSED, false>(tg, tq, tp, tx, a, p, st);
}
template <int FUSED>
static int launch_fast_gy3_any(const CUtensorMap& tg, const CUtensorMap& tq, const CUtensorMap& tp, const CUtensorMap& tx, const FastGy3Args& a,
                               const FastPlan& p, cudaStream_t st) {
  const int nhalf = (a.O + kGy3Rows - 1) / kGy3Rows;
  return nhalf == 1 ? launch_fast_gy3<1, FUSED>(tg, tq, tp, tx, a, p, st)
       : nhalf == 2 ? launch_fast_gy3<2, FUSED>(tg, tq, tp, tx, a, p, st)
                    : launch_fast_gy3<3, FUSED>(tg, tq, tp, tx, a, p, st);
}
static int launch_fast_gy2_any(const CUtensorMap& tg, const CUtensorMap& tq, const FastGy2Args& a, const FastPlan& p, cudaStream_t st) {
  // default: the tensor-pipe form (fast_bwd_gy3_kernel); QW_GY_MMA=0 selects the FFMA form for A/B
  if (flag_gy_mma() && p.tw == FTW) {
    FastGy3Args a3{a.w_post, nullptr, a.gout, a.part, nullptr, nullptr, nullptr, a.B, 0, a.O, a.Lout, 0, a.tiles_per_utt, a.num_tiles, a.PA1, 0, 0, a.tl, option(kOptDbgGy), a.b_post, a.act, (option(kOptRevTiles) >> 1) & 1};
    a3.gen_gpre = a.gen_gpre;
    a3.gen_w = a.gen_w;
    a3.gen_LP = a.gen_LP;
    return launch_fast_gy3_any<0>(tg, tq, tq, tq, a3, p, st);
  }
  QW_CHECK_ARG(a.gen_gpre == nullptr, -2, "the chained backward needs the tensor-pipe gy kernel (option GY_MMA=1)");
  const int forced = option(kOptGyWarps);
  const bool wide = forced == 12;
  if (wide) {
    const int nhalf = (a.O + 383) / 384;
    return nhalf == 1 ? launch_fast_gy2<1, 12, 2>(tg, tq, a, p, st) : launch_fast_gy2<2, 12, 2>(tg, tq, a, p, st);
  }
  const int nhalf = (a.O + 191) / 192;
  return nhalf == 1 ? launch_fast_gy2<1, 6, 3>(tg, tq, a, p, st)
       : nhalf == 2 ? launch_fast_gy2<2, 6, 3>(tg, tq, a, p, st)
                    : launch_fast_gy2<3, 6, 3>(tg, tq, a, p, st);
}

template <int S, int PAR, bool GX>
static int launch_fast_pre(const CUtensorMap& tx, const CUtensorMap& tgx, const FastPreArgs& a, const FastPlan& p, cudaStream_t st) {
  const size_t smem = fast_pre_smem_bytes<S>();
  auto k = fast_bwd_pre_kernel<S, PAR, GX>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_symbol(kKBwdPre, "fast_bwd_pre_kernel<%d, %d, %d>", S, PAR, (int)GX);
  {
    KernelTimer kt(kKBwdPre, st);
    QW_CUDA_OK(launch_pdl(p.small, k, dim3(p.gridPx * p.nchunks), dim3(kThreads), smem, st, tx, tgx, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

int fast_backward(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qwts,
                  const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post,
                  unsigned char* ws, const ConvDims& d, cudaStream_t st, const FastDp* dp, const float* b_post, int act, const FastChain* chain) {
  const FastPlan p = make_fast_plan(d);
  const size_t W = (size_t)d.B * d.Lout;
  float* gpre = reinterpret_cast<float*>(ws + p.off_gpre);
  float* part1 = reinterpret_cast<float*>(ws + p.off_p1);
  float* part3 = reinterpret_cast<float*>(ws + p.off_p3);
  alignas(64) CUtensorMap tm_gy, tm_qout, tm_x, tm_gx;
  if (int e = make_tmap_3d_f32(&tm_qout, pre_save + W * FQ, FQ, d.Lout, d.B, FQ, p.tw, false)) return e;
  if (chain) tm_gy = tm_qout;  // never dereferenced: the chained gy kernel rebuilds its tiles from the following layer's gpre rows
  else if (int e = make_tmap_3d_f32(&tm_gy, gy, d.Lout, d.O, d.B, 32, 64, true)) return e;
  if (int e = make_tmap_3d_f32(&tm_x, x, d.L, d.C, d.B, 32, 32, true)) return e;
  if (int e = make_tmap_3d_f32(&tm_gx, gx ? gx : x, d.L, d.C, d.B, 32, 32, true)) return e;
  float* gout = reinterpret_cast<float*>(ws + p.off_gout);
  float* part2 = reinterpret_cast<float*>(ws + p.off_p2);
  // small batch (<= kFbMaxTiles tiles per CTA), single-layer circuit: the adjoint rides in the gy kernel (fast_bwd_gy3_kernel<.., 1|2>).
  // mode 1 = data layer (no grad_x, stride 1): pre_conv^T as well -- the whole backward is that kernel + finalize;
  // mode 2 = gpre goes to the workspace for the pre_conv^T kernel below.
  int fused = 0;
  QW_CHECK_ARG(!act || (flag_gy_mma() && b_post), -2, "the fused activation needs the tensor-pipe gy kernel (option GY_MMA=1) and post_conv.bias");
  if (!chain && !act && flag_bwd_fused() && flag_gy_mma() && d.Lq == 1 && p.tw == FTW && (long long)p.num_tiles <= (long long)kFbMaxTiles * p.gridGy)
    fused = (gx == nullptr && d.S == 1 && d.P == 1 && d.C <= kFbXRows && flag_bwd_fused() == 1) ? 1 : 2;
  if (fused) {
    alignas(64) CUtensorMap tm_pre, tm_xf;
    if (int e = make_tmap_3d_f32(&tm_pre, pre_save, FQ, d.Lout, d.B, FQ, FTW, false)) return e;
    if (fused == 1)
      if (int e = make_tmap_3d_f32(&tm_xf, x, d.L, d.C, d.B, kFbXW, kFbXRows, false)) return e;
    FastGy3Args a{w_post, qwts, nullptr, part1, part2, part3, gpre, d.B, d.C, d.O, d.Lout, p.LP, p.tiles_per_utt, p.num_tiles, p.PA1, p.PA2,
                  p.PB, timeline_next_slot(), 0, nullptr, 0};
    if (int e = fused == 1 ? launch_fast_gy3_any<1>(tm_gy, tm_qout, tm_pre, tm_xf, a, p, st)
                           : launch_fast_gy3_any<2>(tm_gy, tm_qout, tm_pre, tm_pre, a, p, st))
      return e;
  }
  if (fused == 1) {
    FastFinArgs f{part1, part2, part3, qwts, gw_pre, gb_pre, gqw, gw_post, gb_post, p.gridGy, p.PA1, p.gridGy, p.PA2,
                  p.gridGy, p.PB, d.C, d.O, d.Lq, timeline_next_slot(), FastDp{}, 0};
    const int nblk = p.PA1 / 32 + p.PA2 / 32 + p.PB / 32;
    if (dp && dp->world > 1) {
      f.dp = *dp;
      f.dp.ncol = nblk * 32;
    }
    note_symbol(kKBwdFinalize, "fast_finalize_kernel<%d>", f.dp.world > 1 ? 512 : 1024);
    {
      KernelTimer kt(kKBwdFinalize, st);
      if (f.dp.world > 1) QW_CUDA_OK(launch_pdl(p.small, fast_finalize_kernel<512>, dim3(nblk), dim3(512), 0, st, f));
      else QW_CUDA_OK(launch_pdl(p.small, fast_finalize_kernel<1024>, dim3(nblk), dim3(1024), 0, st, f));
    }
    QW_CUDA_OK(cudaGetLastError());
    return 0;
  }
  // 1) stream gy: gout + partial rows of grad post_conv.{weight,bias}
  if (!fused) {
    FastGy2Args a{w_post, gout, part1, d.B, d.O, d.Lout, p.tiles_per_utt, p.num_tiles, p.PA1, p.tw, timeline_next_slot(), b_post, act};
    if (chain) {
      a.gen_gpre = chain->gpre_pad;
      a.gen_w = chain->w_pre;
      a.gen_LP = chain->LP;
    }
    if (int e = launch_fast_gy2_any(tm_gy, tm_qout, a, p, st)) return e;
  }
  // 2) adjoint differentiation of the circuit, one window per thread
  if (!fused) {
    FastAdjArgs aa{pre_save, gout, qwts, gpre, part2, d.B, d.Lout, p.LP, d.Lq, p.PA2, (long long)W, flag_adj_trig(), timeline_next_slot()};
    const size_t smem = ((size_t)d.Lq * FQ * kGateStride + (size_t)4 * (FQ + d.Lq * 32) * kGyMS) * 4;
    const int adj_spec = option(kOptAdjSpec);
    auto k = (d.Lq == 1 && adj_spec) ? fast_bwd_adj_kernel<false> : fast_bwd_adj_kernel<true>;
    if (smem > 48 * 1024) QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_symbol(kKBwdAdj, "fast_bwd_adj_kernel<%d>", (d.Lq == 1 && adj_spec) ? 0 : 1);
    {
      KernelTimer kt(kKBwdAdj, st);
      QW_CUDA_OK(launch_pdl(p.small, k, dim3(p.gridAdj), dim3(kAdjThreads), smem, st, aa));
    }
    QW_CUDA_OK(cudaGetLastError());
  }
  // 3) pre_conv^T
  {
    FastPreArgs a{gpre, w_pre, part3, d.B, d.C, d.L, d.P, d.Lout, p.LP, p.ptiles_per_utt, p.num_ptiles, p.Cpad, p.gridPx,
                  flag_pre_ex(), timeline_next_slot()};
    const int par = d.P & 1;
    int e;
    if (d.S == 1) e = gx ? launch_fast_pre<1, 0, true>(tm_x, tm_gx, a, p, st) : launch_fast_pre<1, 0, false>(tm_x, tm_gx, a, p, st);
    else if (par) e = gx ? launch_fast_pre<2, 1, true>(tm_x, tm_gx, a, p, st) : launch_fast_pre<2, 1, false>(tm_x, tm_gx, a, p, st);
    else          e = gx ? launch_fast_pre<2, 0, true>(tm_x, tm_gx, a, p, st) : launch_fast_pre<2, 0, false>(tm_x, tm_gx, a, p, st);
    if (e) return e;
  }
  // 4) finalize
  {
    FastFinArgs a{part1, part2, part3, qwts, gw_pre, gb_pre, gqw, gw_post, gb_post, p.gridGy, p.PA1, fused ? p.gridGy : p.gridAdj, p.PA2,
                  p.gridPx, p.PB, d.C, d.O, d.Lq, timeline_next_slot(), FastDp{}, flag_fin_early()};
    const int nblk = p.PA1 / 32 + p.PA2 / 32 + p.PB / 32;
    if (dp && dp->world > 1) {
      a.dp = *dp;
      a.dp.ncol = nblk * 32;
    }
    note_symbol(kKBwdFinalize, "fast_finalize_kernel<%d>", a.dp.world > 1 ? 512 : 1024);
    {
      KernelTimer kt(kKBwdFinalize, st);
      if (a.dp.world > 1) QW_CUDA_OK(launch_pdl(p.small, fast_finalize_kernel<512>, dim3(nblk), dim3(512), 0, st, a));
      else QW_CUDA_OK(launch_pdl(p.small, fast_finalize_kernel<1024>, dim3(nblk), dim3(1024), 0, st, a));
    }
    QW_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace qw
