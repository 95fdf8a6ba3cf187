// Fused TRAINING forward of the encoder stem (/root/reference/whisper/whisper/model.py:193-194,
// /root/reference/quantum_whisper.py:136-137):  y1 = act(conv1(x)),  y2 = act(conv2(y1)),  both QuantumConv1d, n_qubits = 4.
//
// Training needs y1 in HBM (conv2's backward re-reads it for grad pre_conv.weight), so unlike the inference stem (qw_stem.cu) the
// activation IS written -- but it is never read back in the forward: the CTA that produces a 384-channel x 32-column tile of y1
// still holds every value in a register when it stores it, and conv2's pre_conv is a rank-4 contraction of exactly those values.
// One kernel therefore does what fast_fwd_kernel<1> + fast_fwd_kernel<2> did, minus conv2's read of the (B, hidden, L) tensor
// (73.7 MB of the 199.7 MB the two forwards move at batch 16) and minus a kernel boundary.
//
// Per CTA (persistent, a CONTIGUOUS range of tiles so that neighbouring tiles meet in the same CTA):
//   warps 0-7  streaming   A(n): pre_conv1 partial sums of tile n from the TMA ring (as fast_fwd_kernel)
//                          B(n-1): post_conv1 (+ act) of tile n-1 -> y1 (128-bit stores) and, from the same registers, the
//                                  partial sums of conv2's pre_conv: column c of y1 feeds window c/2 with tap 1 (c even) or
//                                  windows (c-1)/2 with tap 2 and (c+1)/2 with tap 0 (c odd)
//                          C(n-2): post_conv2 (+ act) of the 16 conv2 windows of tile n-2 -> y2
//   warp  8    circuit 1   C1(n): conv1's circuit, one window per lane
//   warp  9    circuit 2   first the halo: the tap-0 term of the CTA's first conv2 window comes from the y1 column just LEFT of its
//                          range, which another CTA produces; this warp recomputes that one column from x (240 FMAs, one circuit,
//                          hidden x 4 FMAs).  Then C2(m): conv2's circuit on lanes 0-15 for every tile of the range
// A tile's 16 conv2 windows need 33 columns of y1: its own 32 and the one to their LEFT (tap 0 of its first window).  That term
// flows forward as a "carry": the last (odd) column of tile n contributes tap 0 to the first window of tile n + 1, which the same
// CTA handles next; only the first tile of a CTA's range needs the recomputed halo column.  All hand-offs are mbarriers; there is no CTA-wide
// barrier in the loop.
#include "../../include/qw.h"
#include "qw_act.cuh"
#include "qw_circuit.cuh"
#include "qw_conv1d_plan.cuh"
#include "qw_tma.cuh"

namespace qw {
namespace st {

constexpr int FQ = 4, FTW = 32;
constexpr int kSW = 8;                        // streaming warps
constexpr int kThreadsT = (kSW + 2) * 32;     // + the two circuit warps
constexpr int kStages = 3;
constexpr int XW = 40;                        // x tile columns (stride 1): tap k of local window w = column 3 + w + k
constexpr int kP2 = 17;                       // conv2 windows a tile contributes to: its own 16 + the carry

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ unsigned char* align1024(unsigned char* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
__device__ __forceinline__ float4 fma4(float4 w, float s, float4 a) {
  return make_float4(fmaf(w.x, s, a.x), fmaf(w.y, s, a.y), fmaf(w.z, s, a.z), fmaf(w.w, s, a.w));
}

struct Args {
  const float *w_pre1, *b_pre1, *qw1, *w_post1, *b_post1;
  const float *w_pre2, *b_pre2, *qw2, *w_post2, *b_post2;
  const float* x;
  float *y1, *ps1, *y2, *ps2;   // ps: pre_save (2, W, 4)
  int B, C, L, H, O, Lq;        // x (B,C,L); y1 (B,H,L); y2 (B,O,L/2)
  int tiles_per_utt, num_tiles;
  unsigned long long* tl;
  int dbg;  // experiment switch DBG_FWD (results are garbage): 1 = skip the pre_conv1 FMAs, 2 = skip post_conv / conv2 pre_conv FMAs, 4 = skip the circuits
};

template <int RC>
__host__ __device__ constexpr size_t smem_bytes(int C, int H, int O, int Lq) {
  return 1024 + ((size_t)kStages * RC * XW + (size_t)C * 3 * FQ + (size_t)H * 5 + (size_t)H * 3 * FQ + (size_t)O * 5 + 8 +
                 2 * (size_t)Lq * FQ * kGateStride + 2 * kSW * FTW * FQ + 2 * FTW * FQ + 2 * kSW * kP2 * FQ + 2 * 16 * FQ + 4) * 4 +
         (2 * kStages + 17) * 8;
}

// DBG: the DBG_FWD experiment switches compiled in (a separate instantiation: in the production kernel they cost 2.5 us per step).
template <int RC, bool ACT, bool DBG>
__global__ void __launch_bounds__(kThreadsT, 2) stem_train_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const Args a) {
  constexpr int STAGE = RC * XW;
  constexpr int ITS = RC / (4 * kSW);
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  float* stage = reinterpret_cast<float*>(align1024(smem_dyn));   // [kStages][RC][XW]
  const int CK1 = a.C * 3, CK2 = a.H * 3, Lq = a.Lq;
  float* wpre1 = stage + (size_t)kStages * STAGE;                  // [C*3][4]
  float* wpost1 = wpre1 + (size_t)CK1 * FQ;                        // [H][4]
  float* bpost1 = wpost1 + (size_t)a.H * FQ;                       // [H]
  float* wpre2 = bpost1 + a.H;                                     // [H*3][4]
  float* wpost2 = wpre2 + (size_t)CK2 * FQ;                        // [O][4]
  float* bpost2 = wpost2 + (size_t)a.O * FQ;                       // [O]
  float* bpre = bpost2 + a.O;                                      // [2][4]
  float* gates1 = bpre + 8;                                        // [Lq][4][16]
  float* gates2 = gates1 + (size_t)Lq * FQ * kGateStride;
  float* part = gates2 + (size_t)Lq * FQ * kGateStride;            // [2][kSW][32][4]
  float* outs = part + 2 * kSW * FTW * FQ;                         // [2][32][4]
  float* part2 = outs + 2 * FTW * FQ;                              // [2][kSW][17][4]
  float* outs2 = part2 + 2 * kSW * kP2 * FQ;                       // [2][16][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(outs2 + 2 * 16 * FQ + 4);  // [kStages]
  uint64_t* empty = full + kStages;
  uint64_t* pfull = empty + kStages;   // [2] each below
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;
  uint64_t* oempty = ofull + 2;
  uint64_t* p2full = oempty + 2;
  uint64_t* p2empty = p2full + 2;
  uint64_t* o2full = p2empty + 2;
  uint64_t* o2empty = o2full + 2;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr = lane >> 3, tl = lane & 7;
  const int dbg = DBG ? a.dbg : 0;
  tl_begin(a.tl);
  const int Lout2 = a.L >> 1;
  const int tile0 = (int)(((long long)blockIdx.x * a.num_tiles) / gridDim.x);
  const int my_tiles = (int)(((long long)(blockIdx.x + 1) * a.num_tiles) / gridDim.x) - tile0;

  auto issue = [&](int n) {
    const int tile = tile0 + n;
    const int b = tile / a.tiles_per_utt;
    const int i0 = (tile - b * a.tiles_per_utt) * FTW;
    const int s = n % kStages;
    mbar_arrive_expect_tx(&full[s], STAGE * 4);
    tma_load_3d(stage + (size_t)s * STAGE, &tm_x, i0 - 4, 0, b, &full[s]);
  };
  if (tid == kSW * 32) {  // lane 0 of the circuit warp: stages no parameters, so it can sit in the dependency wait
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kSW);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], kSW);
      mbar_init(&pempty[s], 1);
      mbar_init(&ofull[s], 1);
      mbar_init(&oempty[s], kSW);
      mbar_init(&p2full[s], kSW);
      mbar_init(&p2empty[s], 1);
      mbar_init(&o2full[s], 1);
      mbar_init(&o2empty[s], kSW);
    }
    fence_mbar_init();
    pdl_wait();
    for (int n = 0; n < kStages - 1 && n < my_tiles; ++n) issue(n);
  }
  // ---- stage parameters: every warp but the circuit warp (whose lane 0 sits in the dependency wait above), so the staging
  // overlaps the predecessor's tail
  if (warp != kSW) {
    const int nt = kThreadsT - 32;
    const int tid = (int)threadIdx.x - (warp > kSW ? 32 : 0);
    // (q, C*K) -> [C*K][4]: 4 features x 4 qubits per step, transposed in registers (C*K % 4 == 0: C*3 with C % 4 == 0)
    auto stage_t = [&](const float* __restrict__ src, float* dst, int CK) {
      for (int u = tid; u < CK / 4; u += nt) {
        const float4 r0 = ld4(src + 0 * (size_t)CK + 4 * u), r1 = ld4(src + 1 * (size_t)CK + 4 * u);
        const float4 r2 = ld4(src + 2 * (size_t)CK + 4 * u), r3 = ld4(src + 3 * (size_t)CK + 4 * u);
        st4(dst + (size_t)(4 * u + 0) * FQ, make_float4(r0.x, r1.x, r2.x, r3.x));
        st4(dst + (size_t)(4 * u + 1) * FQ, make_float4(r0.y, r1.y, r2.y, r3.y));
        st4(dst + (size_t)(4 * u + 2) * FQ, make_float4(r0.z, r1.z, r2.z, r3.z));
        st4(dst + (size_t)(4 * u + 3) * FQ, make_float4(r0.w, r1.w, r2.w, r3.w));
      }
    };
    stage_t(a.w_pre1, wpre1, CK1);
    stage_t(a.w_pre2, wpre2, CK2);
    for (int u = tid; u < a.H; u += nt) {
      st4(wpost1 + (size_t)u * FQ, ld4(a.w_post1 + (size_t)u * FQ));
      bpost1[u] = a.b_post1[u];
    }
    for (int u = tid; u < a.O; u += nt) {
      st4(wpost2 + (size_t)u * FQ, ld4(a.w_post2 + (size_t)u * FQ));
      bpost2[u] = a.b_post2[u];
    }
    if (tid < FQ) {
      bpre[tid] = a.b_pre1[tid];
      bpre[4 + tid] = a.b_pre2[tid];
    }
    if (tid < Lq * FQ) make_gate<float>(a.qw1 + tid * 3, gates1 + tid * kGateStride);
    else if (tid >= 64 && tid < 64 + Lq * FQ) make_gate<float>(a.qw2 + (tid - 64) * 3, gates2 + (tid - 64) * kGateStride);
  }
  __syncthreads();
  pdl_wait();  // nothing global is written above
  pdl_launch();

  if (warp == kSW + 1) {
    // ======================================================== circuit warp 2: first the tap-0 term of the range's first conv2 window
    if (my_tiles > 0) {
      const int b = tile0 / a.tiles_per_utt;
      const int i0 = (tile0 - b * a.tiles_per_utt) * FTW;
      float h[FQ] = {0.f, 0.f, 0.f, 0.f};
      if (i0 > 0) {  // (first tile of an utterance: the column left of it is conv2's zero padding)
        const int col = i0 - 1;  // conv1 window: input positions col - 1 .. col + 1
        const float* __restrict__ xb = a.x + (size_t)b * a.C * a.L;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int f = lane; f < CK1; f += 32) {
          const int c = f / 3, k = f - c * 3;
          const int l = col - 1 + k;
          const float xv = (l >= 0 && l < a.L) ? __ldg(xb + (size_t)c * a.L + l) : 0.f;
          acc = fma4(ld4(wpre1 + (size_t)f * FQ), xv, acc);
        }
        float pre[FQ] = {warp_sum<float>(acc.x) + bpre[0], warp_sum<float>(acc.y) + bpre[1], warp_sum<float>(acc.z) + bpre[2],
                         warp_sum<float>(acc.w) + bpre[3]};
        float re[1 << FQ], im[1 << FQ], z[FQ];
        circuit_forward_amp<float, FQ>(pre, gates1, Lq, re, im, z);
        float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = lane; o < a.H; o += 32) {
          const float4 wv = ld4(wpost1 + (size_t)o * FQ);
          float v = fmaf(wv.w, z[3], fmaf(wv.z, z[2], fmaf(wv.y, z[1], fmaf(wv.x, z[0], bpost1[o]))));
          if (ACT) v = gelu_erf(v);
          a2 = fma4(ld4(wpre2 + (size_t)(o * 3) * FQ), v, a2);
        }
        h[0] = warp_sum<float>(a2.x); h[1] = warp_sum<float>(a2.y); h[2] = warp_sum<float>(a2.z); h[3] = warp_sum<float>(a2.w);
      }
      // ---- then this warp is conv2's circuit warp: C2(m) for every tile of the range (conv1's circuit runs on warp 8 at the
      // same time: one warp doing both circuits back to back made its ~2 x 700-instruction dependent chain the tile period)
      float carry[FQ] = {0.f, 0.f, 0.f, 0.f};  // lane 16: tap-0 term of the NEXT tile's first conv2 window
      for (int m = 0; m < my_tiles; ++m) {
        const int tile = tile0 + m;
        const int bb = tile / a.tiles_per_utt;
        const int tu = tile - bb * a.tiles_per_utt;
        const int i2 = tu * 16 + lane;
        const int pb = m & 1;
        mbar_wait(&p2full[pb], (m >> 1) & 1);
        const float* pp = part2 + (size_t)pb * kSW * kP2 * FQ;
        float pre[FQ] = {0.f, 0.f, 0.f, 0.f};
        if (lane < kP2) {
#pragma unroll
          for (int w = 0; w < kSW; ++w) {
            const float4 pv = ld4(pp + ((size_t)w * kP2 + lane) * FQ);
            pre[0] += pv.x; pre[1] += pv.y; pre[2] += pv.z; pre[3] += pv.w;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&p2empty[pb]);
        // carry in: from the previous tile of this CTA (lane 16 of the last round), the halo column (first tile of the range), or
        // nothing (first tile of an utterance: conv2's zero padding)
        float cin[FQ];
#pragma unroll
        for (int j = 0; j < FQ; ++j) cin[j] = m == 0 ? h[j] : __shfl_sync(0xffffffffu, carry[j], 16);
        if (tu == 0) cin[0] = cin[1] = cin[2] = cin[3] = 0.f;
        if (lane == 16) {
#pragma unroll
          for (int j = 0; j < FQ; ++j) carry[j] = pre[j];
        }
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < FQ; ++j) pre[j] += cin[j];
        }
#pragma unroll
        for (int j = 0; j < FQ; ++j) pre[j] += bpre[4 + j];
        float out[FQ] = {0.f, 0.f, 0.f, 0.f};
        if ((dbg & 4) == 0 && lane < 16 && i2 < Lout2) {
          float re[1 << FQ], im[1 << FQ];
          circuit_forward_amp<float, FQ>(pre, gates2, Lq, re, im, out);
          const size_t wi = (size_t)bb * Lout2 + i2;
          st4(a.ps2 + wi * FQ, make_float4(pre[0], pre[1], pre[2], pre[3]));
          st4(a.ps2 + ((size_t)a.B * Lout2 + wi) * FQ, make_float4(out[0], out[1], out[2], out[3]));
        }
        if (m >= 2) mbar_wait(&o2empty[pb], ((m >> 1) - 1) & 1);
        if (lane < 16) st4(outs2 + (size_t)pb * 16 * FQ + (size_t)lane * FQ, make_float4(out[0], out[1], out[2], out[3]));
        __syncwarp();
        if (lane == 0) mbar_arrive(&o2full[pb]);
      }
    }
    return;
  }

  if (warp == kSW) {
    // ======================================================== circuit warp
    for (int n = 0; n < my_tiles; ++n) {
      {  // ---- C1(n): conv1's circuit, one window per lane
        const int tile = tile0 + n;
        const int b = tile / a.tiles_per_utt;
        const int i = (tile - b * a.tiles_per_utt) * FTW + lane;
        const int pb = n & 1;
        mbar_wait(&pfull[pb], (n >> 1) & 1);
        const float* pp = part + (size_t)pb * kSW * FTW * FQ;
        float pre[FQ];
#pragma unroll
        for (int j = 0; j < FQ; ++j) pre[j] = bpre[j];
#pragma unroll
        for (int w = 0; w < kSW; ++w) {
          const float4 pv = ld4(pp + ((size_t)w * FTW + lane) * FQ);
          pre[0] += pv.x; pre[1] += pv.y; pre[2] += pv.z; pre[3] += pv.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pempty[pb]);
        float out[FQ] = {0.f, 0.f, 0.f, 0.f};
        if ((dbg & 4) == 0 && i < a.L) {
          float re[1 << FQ], im[1 << FQ];
          circuit_forward_amp<float, FQ>(pre, gates1, Lq, re, im, out);
          const size_t wi = (size_t)b * a.L + i;
          st4(a.ps1 + wi * FQ, make_float4(pre[0], pre[1], pre[2], pre[3]));
          st4(a.ps1 + ((size_t)a.B * a.L + wi) * FQ, make_float4(out[0], out[1], out[2], out[3]));
        }
        if (n >= 2) mbar_wait(&oempty[pb], ((n >> 1) - 1) & 1);
        st4(outs + (size_t)pb * FTW * FQ + (size_t)lane * FQ, make_float4(out[0], out[1], out[2], out[3]));
        __syncwarp();
        if (lane == 0) mbar_arrive(&ofull[pb]);
      }
    }
    return;
  }

  // ========================================================== streaming warps
  // The outputs leave as plain 128-bit stores from the registers that computed them.  Measured alternatives (B200, batch 16, this
  // kernel alone; tools/dbg_fwd.py): with the y stores removed it runs in 29 us instead of 44 -- the 110 MB it writes are not
  // overlapped with its arithmetic -- but routing them through shared-memory staging + bulk tensor stores did not help: 12 KB
  // chunks with one named barrier per chunk 45.9 us (40.5 with the stores themselves disabled: the barriers), 1.5 KB per-warp
  // chunks with __syncwarp only 45.8 us (36.4 without the stores).  Two small staging buffers cannot decouple a warp from a
  // saturated memory system, and a tile's worth of staging (73 KB) does not fit beside the parameters at two CTAs per SM.
  // De-phasing the two CTAs of every SM (the upper half of the grid starting 1-5 us late, so that store bursts of one CTA meet the
  // compute phases of the other) changes nothing either: 42.5 -> 42.8 ... 44.6 us -- the stall is inside each warp's own issue order.
  for (int n = 0; n <= my_tiles + 1; ++n) {
    // ---- A(n): pre_conv1 partial sums of tile n
    if (n < my_tiles) {
      float acc[4][FQ];
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < FQ; ++j) acc[w][j] = 0.f;
      if (tid == 0) {
        const int gn = n + kStages - 1;
        if (gn < my_tiles) {
          if (gn >= kStages) mbar_wait(&empty[gn % kStages], ((gn / kStages) - 1) & 1);
          issue(gn);
        }
      }
      const int s = n % kStages;
      mbar_wait(&full[s], (n / kStages) & 1);
      const float* st = stage + (size_t)s * STAGE;
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int c = it * (4 * kSW) + warp * 4 + rr;  // channel row (rows >= C are zero-filled by TMA)
        const float* xr = st + c * XW;
        const float4 v0 = ld4(xr + 4 * tl + 4);
        const float xc[6] = {xr[4 * tl + 3], v0.x, v0.y, v0.z, v0.w, xr[4 * tl + 8]};
        if (dbg & 1) {
          acc[0][0] += xc[0] + xc[4];
        } else if (c < a.C) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float4 wv = ld4(wpre1 + (size_t)(c * 3 + k) * FQ);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float xv = xc[w + k];
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
        float* pp = part + (size_t)pb * kSW * FTW * FQ;
#pragma unroll
        for (int w = 0; w < 4; ++w)
          st4(pp + ((size_t)warp * FTW + 4 * tl + w) * FQ, make_float4(acc[w][0], acc[w][1], acc[w][2], acc[w][3]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[pb]);
    }
    // ---- B(n-1): post_conv1 (+ act) -> y1, and conv2's pre_conv partial sums from the same registers
    if (n >= 1 && n <= my_tiles) {
      const int m = n - 1;
      const int tile = tile0 + m;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * FTW;
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
      float4 aA = make_float4(0.f, 0.f, 0.f, 0.f), aB = aA, aC = aA;  // conv2 windows 2 tl, 2 tl + 1, 2 tl + 2 (tile-local)
      const bool ok = (i0 + 4 * tl) < a.L && !(dbg & 8);  // L % 4 == 0: a lane's 4 columns are all valid or all invalid
      float* __restrict__ yb = a.y1 + (size_t)b * a.H * a.L + i0 + 4 * tl;
      const int ngroups = a.H >> 2;
#pragma unroll 2
      for (int og = warp; og < ngroups; og += kSW) {
        const int o = og * 4 + rr;
        const float4 wv = ld4(wpost1 + (size_t)o * FQ);
        const float bv = bpost1[o];
        float4 r;
        if (dbg & 2) {
          if (ok) st4(yb + (size_t)o * a.L, make_float4(bv, bv, bv, bv));
          continue;
        }
        r.x = fmaf(wv.w, ov[0][3], fmaf(wv.z, ov[0][2], fmaf(wv.y, ov[0][1], fmaf(wv.x, ov[0][0], bv))));
        r.y = fmaf(wv.w, ov[1][3], fmaf(wv.z, ov[1][2], fmaf(wv.y, ov[1][1], fmaf(wv.x, ov[1][0], bv))));
        r.z = fmaf(wv.w, ov[2][3], fmaf(wv.z, ov[2][2], fmaf(wv.y, ov[2][1], fmaf(wv.x, ov[2][0], bv))));
        r.w = fmaf(wv.w, ov[3][3], fmaf(wv.z, ov[3][2], fmaf(wv.y, ov[3][1], fmaf(wv.x, ov[3][0], bv))));
        if (ACT) {
          r.x = gelu_erf(r.x); r.y = gelu_erf(r.y); r.z = gelu_erf(r.z); r.w = gelu_erf(r.w);
        }
        if (ok) st4(yb + (size_t)o * a.L, r);
        // columns 4 tl (even: tap 1 of window A), +1 (odd: tap 2 of A, tap 0 of B), +2 (even: tap 1 of B), +3 (odd: tap 2 of B, tap 0 of C).
        // No mask: columns >= L only exist in an utterance's last tile (their r is act(bias), finite) and only feed conv2 windows
        // >= L/2, which are never read out, or the carry, which the next tile -- the first of a new utterance -- drops.
        const float* w2 = wpre2 + (size_t)(o * 3) * FQ;
        const float4 w0 = ld4(w2), w1 = ld4(w2 + FQ), w2v = ld4(w2 + 2 * FQ);
        aA = fma4(w1, r.x, aA);
        aA = fma4(w2v, r.y, aA);
        aB = fma4(w0, r.y, aB);
        aB = fma4(w1, r.z, aB);
        aB = fma4(w2v, r.w, aB);
        aC = fma4(w0, r.w, aC);
      }
      // reduce over the 4 row classes; window C of lane group tl is window A of group tl + 1
      float v2[12] = {aA.x, aA.y, aA.z, aA.w, aB.x, aB.y, aB.z, aB.w, aC.x, aC.y, aC.z, aC.w};
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        v2[e] += __shfl_xor_sync(0xffffffffu, v2[e], 8);
        v2[e] += __shfl_xor_sync(0xffffffffu, v2[e], 16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float up = __shfl_up_sync(0xffffffffu, v2[8 + j], 1, 8);
        if (tl > 0) v2[j] += up;
      }
      const int pb = m & 1;
      if (m >= 2) mbar_wait(&p2empty[pb], ((m >> 1) - 1) & 1);
      if (rr == 0) {
        float* pp = part2 + ((size_t)pb * kSW + warp) * kP2 * FQ;
        st4(pp + (size_t)(2 * tl) * FQ, make_float4(v2[0], v2[1], v2[2], v2[3]));
        st4(pp + (size_t)(2 * tl + 1) * FQ, make_float4(v2[4], v2[5], v2[6], v2[7]));
        if (tl == 7) st4(pp + (size_t)16 * FQ, make_float4(v2[8], v2[9], v2[10], v2[11]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p2full[pb]);
    }
    // ---- C(n-2): post_conv2 (+ act) of the tile's 16 conv2 windows -> y2
    if (n >= 2) {
      const int m = n - 2;
      const int tile = tile0 + m;
      const int b = tile / a.tiles_per_utt;
      const int i2 = (tile - b * a.tiles_per_utt) * 16;
      const int ob = m & 1;
      const int wg = tl & 3, sub = 2 * rr + (tl >> 2);  // 4 windows per lane, 8 channel rows per warp step (a quarter-warp = 2 adjacent rows)
      mbar_wait(&o2full[ob], (m >> 1) & 1);
      const float* oo = outs2 + (size_t)ob * 16 * FQ;
      float ov[4][FQ];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float4 v = ld4(oo + (size_t)(4 * wg + w) * FQ);
        ov[w][0] = v.x; ov[w][1] = v.y; ov[w][2] = v.z; ov[w][3] = v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&o2empty[ob]);
      const bool ok = (i2 + 4 * wg) < Lout2 && !(dbg & 8);  // Lout2 % 4 == 0
      float* __restrict__ yb = a.y2 + (size_t)b * a.O * Lout2 + i2 + 4 * wg;
      const int ngroups = a.O >> 3;
#pragma unroll 2
      for (int og = warp; og < ngroups; og += kSW) {
        const int o = og * 8 + sub;
        const float4 wv = ld4(wpost2 + (size_t)o * FQ);
        const float bv = bpost2[o];
        float4 r;
        r.x = fmaf(wv.w, ov[0][3], fmaf(wv.z, ov[0][2], fmaf(wv.y, ov[0][1], fmaf(wv.x, ov[0][0], bv))));
        r.y = fmaf(wv.w, ov[1][3], fmaf(wv.z, ov[1][2], fmaf(wv.y, ov[1][1], fmaf(wv.x, ov[1][0], bv))));
        r.z = fmaf(wv.w, ov[2][3], fmaf(wv.z, ov[2][2], fmaf(wv.y, ov[2][1], fmaf(wv.x, ov[2][0], bv))));
        r.w = fmaf(wv.w, ov[3][3], fmaf(wv.z, ov[3][2], fmaf(wv.y, ov[3][1], fmaf(wv.x, ov[3][0], bv))));
        if (ACT) {
          r.x = gelu_erf(r.x); r.y = gelu_erf(r.y); r.z = gelu_erf(r.z); r.w = gelu_erf(r.w);
        }
        if (dbg & 2) r = make_float4(bv, bv, bv, bv);
        if (ok) st4(yb + (size_t)o * Lout2, r);
      }
    }
  }
  tl_end(a.tl);
}

template <int RC, bool ACT, bool DBG>
static int launch(const CUtensorMap& tm, const Args& a, int grid, bool small, cudaStream_t st) {
  const size_t smem = smem_bytes<RC>(a.C, a.H, a.O, a.Lq);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "fused training stem needs %zu bytes of shared memory", smem);
  auto k = stem_train_fwd_kernel<RC, ACT, DBG>;
  QW_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_symbol(kKFwd, "stem_train_fwd_kernel<%d, %d>", RC, (int)ACT);
  {
    KernelTimer kt(kKFwd, st);
    QW_CUDA_OK(launch_pdl(small, k, dim3(grid), dim3(kThreadsT), smem, st, tm, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace st
}  // namespace qw

// Measured on B200 (bench.py --stem-forward fused|split, ms per stem step): batch 16 (5.1 tiles per CTA) 0.1163 vs 0.1198, batch 32
// (10.2) 0.2111 vs 0.2091, batch 64 0.3954 vs 0.3737, batch 256 1.482 vs 1.343 -- the fused kernel saves a kernel boundary and
// conv2's read of the activation, which pays while launches are latency-dominated; with many tiles per CTA its longer per-tile
// instruction stream (both layers' epilogues in one warp role) loses to two leaner kernels.
extern "C" int qw_stem_train_forward_preferred(int B, int L) {
  if (B <= 0 || L <= 0) return 0;
  const long long tiles = (long long)B * ((L + qw::st::FTW - 1) / qw::st::FTW);
  return tiles <= 8LL * 2 * qw::num_sms() ? 1 : 0;
}

extern "C" int qw_stem_train_forward(const float* x, const float* w_pre1, const float* b_pre1, const float* qw1, const float* w_post1,
                                     const float* b_post1, const float* w_pre2, const float* b_pre2, const float* qw2,
                                     const float* w_post2, const float* b_post2, float* y1, float* pre_save1, float* y2,
                                     float* pre_save2, int B, int C, int L, int hidden, int O, int n_layers, int activation,
                                     void* stream) {
  using namespace qw;
  QW_CHECK_ARG(x && w_pre1 && b_pre1 && qw1 && w_post1 && b_post1 && w_pre2 && b_pre2 && qw2 && w_post2 && b_post2 && y1 && pre_save1 &&
                   y2 && pre_save2,
               -1, "qw_stem_train_forward: null pointer argument");
  QW_CHECK_ARG(B > 0 && C > 0 && L > 0 && hidden > 0 && O > 0 && n_layers >= 1 && n_layers <= 4, -1, "qw_stem_train_forward: bad shape");
  QW_CHECK_ARG(activation == QW_ACT_NONE || activation == QW_ACT_GELU, -2, "qw_stem_train_forward: activation=%d not supported", activation);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  QW_CHECK_ARG(C <= 96 && C % 4 == 0 && L % 8 == 0 && hidden % 4 == 0 && O % 8 == 0 && hidden <= 384 && O <= 384 && al(x) && al(y1) && al(y2) &&
                   al(pre_save1) && al(pre_save2) && al(w_post1) && al(w_post2) && al(w_pre1) && al(w_pre2) && tmap_encode_fn() != nullptr,
               -2,
               "qw_stem_train_forward: outside the fused regime (C <= 96, C %% 4 == 0, L %% 8 == 0, hidden %% 4 == 0, O %% 8 == 0, hidden, O <= 384, "
               "16-byte aligned tensors): run the two layers through qw_conv1d_forward_act");
  cudaStream_t st = (cudaStream_t)stream;
  st::Args a{};
  a.w_pre1 = w_pre1; a.b_pre1 = b_pre1; a.qw1 = qw1; a.w_post1 = w_post1; a.b_post1 = b_post1;
  a.w_pre2 = w_pre2; a.b_pre2 = b_pre2; a.qw2 = qw2; a.w_post2 = w_post2; a.b_post2 = b_post2;
  a.x = x; a.y1 = y1; a.ps1 = pre_save1; a.y2 = y2; a.ps2 = pre_save2;
  a.B = B; a.C = C; a.L = L; a.H = hidden; a.O = O; a.Lq = n_layers;
  a.tiles_per_utt = (L + st::FTW - 1) / st::FTW;
  a.num_tiles = B * a.tiles_per_utt;
  a.tl = timeline_next_slot();
  a.dbg = option(kOptDbgFwd);
  alignas(64) CUtensorMap tm;
  if (int e = make_tmap_3d_f32(&tm, x, L, C, B, st::XW, 96, false)) return e;
  const int cap = 2 * num_sms();
  const int grid = a.num_tiles < cap ? a.num_tiles : cap;
  const bool small = a.num_tiles <= 24 * cap;
  if (a.dbg)
    return activation == QW_ACT_GELU ? st::launch<96, true, true>(tm, a, grid, small, st) : st::launch<96, false, true>(tm, a, grid, small, st);
  return activation == QW_ACT_GELU ? st::launch<96, true, false>(tm, a, grid, small, st) : st::launch<96, false, false>(tm, a, grid, small, st);
}
