// Fused inference forward of the Whisper stem with two QuantumConv1d layers:
//
//     x = gelu(conv1(mel)); x = gelu(conv2(x)); x = x.permute(0, 2, 1); x = x + positional_embedding
//     (/root/reference/whisper/whisper/model.py:193-198 with conv1/conv2 = QuantumConv1d, /root/reference/quantum_whisper.py:136-137)
//
// The (B, 384, 3000) activation between the two layers never exists.  conv1's post_conv is a rank-4 map of its <Z> readouts,
// h[c, t] = gelu(W_post1[c, :] . q1[t, :] + b_post1[c]), so the first kernel stops at the readouts (16 bytes per time step
// instead of 1 536) and the second kernel rebuilds the taps of h it needs in registers, straight into conv2's pre_conv
// contraction -- then the circuit, post_conv, GELU, the transpose to (B, T, O) and the positional embedding, stored once.
// HBM traffic per utterance: 0.96 MB (mel) + 2.3 MB (out) + 0.1 MB (q1) instead of ~21 MB for the op-by-op sequence; the kernel is
// bound by the FP32 pipe (the erf GELU of 2 x 384 hidden + 384 output values per output frame), not by HBM.
//
//   kernel 1: fast_fwd_kernel<1, RC> of qw_conv1d_fast.cu with y == NULL (TMA ring -> pre_conv -> circuit -> q1).
//   kernel 2: stem2_kernel below (persistent, 8 streaming warps + 1 circuit warp, same mbarrier hand-offs as the forward kernel).
//
// Inference only: nothing is saved for a backward pass (training goes through the two QuantumConv1d operators).
#include "../../include/qw.h"
#include "qw_act.cuh"
#include "qw_async.cuh"
#include "qw_conv1d_plan.cuh"

namespace qw {

namespace {
constexpr int SQ = 4;      // n_qubits
constexpr int STW = 32;    // conv2 windows per tile
constexpr int kSW = 8;     // streaming warps
constexpr int kStemThreads = (kSW + 1) * 32;
constexpr int kQ1Cols = 72;                       // q1 tile: column p <-> time step 2*i0 - 4 + p (tap k of local window w: 3 + 2w + k)
constexpr int kQ1Slots = kQ1Cols + kQ1Cols / 8;   // skewed by one float4 per 8 columns: the 8 lane groups hit 8 distinct bank groups

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ int q1_slot(int p) { return p + (p >> 3); }
__device__ __forceinline__ void bar_stream() { asm volatile("bar.sync 1, %0;" ::"n"(kSW * 32) : "memory"); }

struct Stem2Args {
  const float* q1;                                          // [B][L][4] <Z> readouts of conv1
  const float *w_post1, *b_post1;                           // (C, 4), (C): conv1.post_conv, C = hidden channels
  const float *w_pre2, *b_pre2, *qw2, *w_post2, *b_post2;   // conv2: (4, C*3), (4), (Lq, 4, 3), (O, 4), (O)
  const float* pos;                                         // (Lout, O) or null
  float* out;                                               // (B, Lout, O)
  int B, C, L, O, Lq, Lout, tiles_per_utt, num_tiles;
  unsigned long long* tl;
};

// hidden channels are padded to a multiple of the 32 rows one pass of the 8 streaming warps covers (zero weights: no branch)
__host__ __device__ constexpr int stem2_cpad(int C) { return (int)align_up((size_t)C, 4 * kSW); }
__host__ __device__ constexpr size_t stem2_smem_bytes(int Creal, int O, int Lq) {
  const size_t C = stem2_cpad(Creal);
  return ((size_t)C * 3 * SQ + (size_t)C * SQ + align_up(C, 4) + (size_t)O * SQ + align_up(O, 4) + 4 + (size_t)Lq * SQ * kGateStride +
          (size_t)kQ1Slots * 4 + 2 * kSW * STW * SQ + 2 * STW * SQ) * 4 + 8 * 8;
}

__global__ void __launch_bounds__(kStemThreads, 2) stem2_kernel(const Stem2Args a) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  const int CK = a.C * 3, Cp = stem2_cpad(a.C);
  float* wpre_t = reinterpret_cast<float*>(smem_dyn);            // [Cp*3][4]
  float* wpost1 = wpre_t + (size_t)Cp * 3 * SQ;                  // [Cp][4]
  float* bpost1 = wpost1 + (size_t)Cp * SQ;                      // [Cp]
  float* wpost2_t = bpost1 + Cp;                                 // [4][O]
  float* bpost2 = wpost2_t + (size_t)a.O * SQ;                   // [O]
  float* bpre = bpost2 + align_up(a.O, 4);                       // [4]
  float* gates = bpre + 4;                                       // [Lq][4][16]
  float* q1t = gates + (size_t)a.Lq * SQ * kGateStride;          // [kQ1Slots][4]
  float* part = q1t + kQ1Slots * 4;                              // [2][kSW][32][4]
  float* outs = part + 2 * kSW * STW * SQ;                       // [2][32][4]
  uint64_t* pfull = reinterpret_cast<uint64_t*>(outs + 2 * STW * SQ);  // [2]
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;
  uint64_t* oempty = ofull + 2;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr = lane >> 3, tl = lane & 7;
  tl_begin(a.tl);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], kSW);
      mbar_init(&pempty[s], 1);
      mbar_init(&ofull[s], 1);
      mbar_init(&oempty[s], kSW);
    }
    fence_mbar_init();
  }
  // ---- stage parameters (pre_conv weights transposed to [feature][qubit], post_conv2 weights to [qubit][channel])
  if ((CK & 3) == 0) {
    for (int u = tid; u < CK / 4; u += kStemThreads) {
      const float4 r0 = ld4(a.w_pre2 + 0 * (size_t)CK + 4 * u), r1 = ld4(a.w_pre2 + 1 * (size_t)CK + 4 * u);
      const float4 r2 = ld4(a.w_pre2 + 2 * (size_t)CK + 4 * u), r3 = ld4(a.w_pre2 + 3 * (size_t)CK + 4 * u);
      st4(wpre_t + (size_t)(4 * u + 0) * SQ, make_float4(r0.x, r1.x, r2.x, r3.x));
      st4(wpre_t + (size_t)(4 * u + 1) * SQ, make_float4(r0.y, r1.y, r2.y, r3.y));
      st4(wpre_t + (size_t)(4 * u + 2) * SQ, make_float4(r0.z, r1.z, r2.z, r3.z));
      st4(wpre_t + (size_t)(4 * u + 3) * SQ, make_float4(r0.w, r1.w, r2.w, r3.w));
    }
  } else {
    for (int idx = tid; idx < CK * SQ; idx += kStemThreads) {
      const int j = idx / CK, f = idx - j * CK;
      wpre_t[f * SQ + j] = a.w_pre2[idx];
    }
  }
  for (int u = tid; u < a.C; u += kStemThreads) {
    st4(wpost1 + (size_t)u * SQ, ld4(a.w_post1 + (size_t)u * SQ));
    bpost1[u] = a.b_post1[u];
  }
  for (int u = a.C + tid; u < Cp; u += kStemThreads) {  // padding rows: h = gelu(0) = 0 times zero weights
    st4(wpost1 + (size_t)u * SQ, make_float4(0.f, 0.f, 0.f, 0.f));
    bpost1[u] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) st4(wpre_t + (size_t)(u * 3 + k) * SQ, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  for (int u = tid; u < a.O; u += kStemThreads) {
    const float4 w = ld4(a.w_post2 + (size_t)u * SQ);
    wpost2_t[0 * a.O + u] = w.x;
    wpost2_t[1 * a.O + u] = w.y;
    wpost2_t[2 * a.O + u] = w.z;
    wpost2_t[3 * a.O + u] = w.w;
    bpost2[u] = a.b_post2[u];
  }
  if (tid < SQ) bpre[tid] = a.b_pre2[tid];
  if (tid < a.Lq * SQ) make_gate<float>(a.qw2 + tid * 3, gates + tid * kGateStride);
  __syncthreads();
  pdl_wait();  // q1 is the previous kernel's output
  pdl_launch();

  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kSW) {
    // ======================================================== circuit warp: one lane per window
    for (int n = 0; n < my_tiles; ++n) {
      const int pb = n & 1, ph = (n >> 1) & 1;
      while (!mbar_try_wait(&pfull[pb], ph)) __nanosleep(256);  // a tile of pre_conv takes microseconds: do not burn issue slots
      const float* pp = part + (size_t)pb * kSW * STW * SQ;
      float pre[SQ];
#pragma unroll
      for (int j = 0; j < SQ; ++j) pre[j] = bpre[j];
#pragma unroll
      for (int w = 0; w < kSW; ++w) {
        const float4 pv = ld4(pp + ((size_t)w * STW + lane) * SQ);
        pre[0] += pv.x; pre[1] += pv.y; pre[2] += pv.z; pre[3] += pv.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pempty[pb]);
      float out[SQ], re[1 << SQ], im[1 << SQ];
      circuit_forward_amp<float, SQ>(pre, gates, a.Lq, re, im, out);  // windows past L_out: finite garbage, never stored
      if (n >= 2) mbar_wait(&oempty[pb], ((n >> 1) - 1) & 1);
      st4(outs + (size_t)pb * STW * SQ + (size_t)lane * SQ, make_float4(out[0], out[1], out[2], out[3]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&ofull[pb]);
    }
    return;
  }

  // ========================================================== streaming warps
  const int rows_it = Cp / (4 * kSW);
  // q1 tile of a tile: time steps 2*i0 - 4 ... 2*i0 + 67, one float4 per thread 0..71, fetched one tile ahead into a register
  auto q1_fetch = [&](int n) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < kQ1Cols && n < my_tiles) {
      const int tile = blockIdx.x + n * gridDim.x;
      const int b = tile / a.tiles_per_utt;
      const int t = 2 * (tile - b * a.tiles_per_utt) * STW - 4 + tid;
      if (t >= 0 && t < a.L) v = ld4(a.q1 + ((size_t)b * a.L + t) * SQ);
    }
    return v;
  };
  float4 q1_next = q1_fetch(0);
  for (int n = 0; n <= my_tiles; ++n) {
    if (n < my_tiles) {
      const int tile = blockIdx.x + n * gridDim.x;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * STW;
      bar_stream();  // everyone is done with the previous tile's q1
      if (tid < kQ1Cols) st4(q1t + (size_t)q1_slot(tid) * 4, q1_next);
      bar_stream();
      q1_next = q1_fetch(n + 1);
      // this lane's 4 adjacent windows need the 9 time steps of columns 8*tl + 3 ... 8*tl + 11
      float4 qv[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) qv[m] = ld4(q1t + (size_t)q1_slot(8 * tl + 3 + m) * 4);
      // conv2's zero padding acts on h: the tap at time step -1 (window 0, k = 0) is 0, not gelu(b).  Time steps >= L are only
      // ever touched by windows >= L_out, which are never stored.
      const bool pad_left = (i0 == 0) && (tl == 0);
      float acc[4][SQ];
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < SQ; ++j) acc[w][j] = 0.f;
#pragma unroll 2
      for (int it = 0; it < rows_it; ++it) {
        const int c = it * (4 * kSW) + warp * 4 + rr;  // hidden channel (rows >= C carry zero weights)
        const float4 w1 = ld4(wpost1 + (size_t)c * SQ);
        const float b1 = bpost1[c];
        float xc[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) {
          // same fma order as the forward kernel's post_conv, so h matches the unfused layer output bit for bit
          const float v = fmaf(w1.w, qv[m].w, fmaf(w1.z, qv[m].z, fmaf(w1.y, qv[m].y, fmaf(w1.x, qv[m].x, b1))));
          xc[m] = gelu_erf(v);
        }
        if (pad_left) xc[0] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 wv = ld4(wpre_t + (size_t)(c * 3 + k) * SQ);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float xv = xc[w * 2 + k];
            acc[w][0] = fmaf(wv.x, xv, acc[w][0]);
            acc[w][1] = fmaf(wv.y, xv, acc[w][1]);
            acc[w][2] = fmaf(wv.z, xv, acc[w][2]);
            acc[w][3] = fmaf(wv.w, xv, acc[w][3]);
          }
        }
      }
      // reduce over the 4 row classes of the warp; the circuit warp sums the kSW warps
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int j = 0; j < SQ; ++j) {
          float v = acc[w][j];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          acc[w][j] = v;
        }
      const int pb = n & 1;
      if (n >= 2) mbar_wait(&pempty[pb], ((n >> 1) - 1) & 1);
      if (rr == 0) {
        float* pp = part + (size_t)pb * kSW * STW * SQ;
#pragma unroll
        for (int w = 0; w < 4; ++w)
          st4(pp + ((size_t)warp * STW + 4 * tl + w) * SQ, make_float4(acc[w][0], acc[w][1], acc[w][2], acc[w][3]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[pb]);
    }
    // ---- post_conv + GELU + transpose + positional embedding of tile n-1: warp <-> 4 windows, lane <-> 4 adjacent channels
    if (n >= 1) {
      const int m = n - 1;
      const int tile = blockIdx.x + m * gridDim.x;
      const int b = tile / a.tiles_per_utt;
      const int i0 = (tile - b * a.tiles_per_utt) * STW;
      const int ob = m & 1;
      mbar_wait(&ofull[ob], (m >> 1) & 1);
      const float* oo = outs + (size_t)ob * STW * SQ;
      float4 q2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) q2[e] = ld4(oo + (size_t)(warp * 4 + e) * SQ);
      __syncwarp();
      if (lane == 0) mbar_arrive(&oempty[ob]);
      for (int o = lane * 4; o < a.O; o += 128) {
        const float4 w0 = ld4(wpost2_t + 0 * (size_t)a.O + o), w1 = ld4(wpost2_t + 1 * (size_t)a.O + o);
        const float4 w2 = ld4(wpost2_t + 2 * (size_t)a.O + o), w3 = ld4(wpost2_t + 3 * (size_t)a.O + o);
        const float4 bv = ld4(bpost2 + o);
        float4 pe[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + warp * 4 + e;
          pe[e] = (a.pos && i < a.Lout) ? __ldg(reinterpret_cast<const float4*>(a.pos + (size_t)i * a.O + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + warp * 4 + e;
          if (i < a.Lout) {
            float4 r;
            r.x = fmaf(w3.x, q2[e].w, fmaf(w2.x, q2[e].z, fmaf(w1.x, q2[e].y, fmaf(w0.x, q2[e].x, bv.x))));
            r.y = fmaf(w3.y, q2[e].w, fmaf(w2.y, q2[e].z, fmaf(w1.y, q2[e].y, fmaf(w0.y, q2[e].x, bv.y))));
            r.z = fmaf(w3.z, q2[e].w, fmaf(w2.z, q2[e].z, fmaf(w1.z, q2[e].y, fmaf(w0.z, q2[e].x, bv.z))));
            r.w = fmaf(w3.w, q2[e].w, fmaf(w2.w, q2[e].z, fmaf(w1.w, q2[e].y, fmaf(w0.w, q2[e].x, bv.w))));
            r.x = gelu_erf(r.x); r.y = gelu_erf(r.y); r.z = gelu_erf(r.z); r.w = gelu_erf(r.w);
            r.x += pe[e].x; r.y += pe[e].y; r.z += pe[e].z; r.w += pe[e].w;
            st4(a.out + ((size_t)b * a.Lout + i) * a.O + o, r);
          }
        }
      }
    }
  }
  tl_end(a.tl);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace

size_t stem_workspace_bytes(int B, int L) { return align_up((size_t)2 * B * L * SQ * 4, 256); }

int stem_forward(const float* x, const float* const* p1, const float* const* p2, const float* pos, float* out, void* ws, size_t ws_bytes,
                 int B, int Cin, int L, int Cmid, int O, int Lq, cudaStream_t st) {
  ConvDims d1{B, Cin, L, 3, 1, 1, Cmid, SQ, Lq, kEmbAmplitude, L};
  float* q1 = reinterpret_cast<float*>(ws);
  QW_CHECK_ARG(ws_bytes >= stem_workspace_bytes(B, L), -3, "stem workspace too small: %zu < %zu", ws_bytes, stem_workspace_bytes(B, L));
  QW_CHECK_ARG(fast_eligible(d1, x, x, q1, true) && L % 2 == 0 && O % 4 == 0 && O <= 576 && Cmid <= 576 && aligned16(out) && aligned16(pos), -2,
               "fused stem needs the fast-path regime: fp32, n_qubits=4, L %% 4 == 0, hidden and output channels %% 4 == 0 and <= 576, "
               "16-byte aligned tensors (got B=%d C=%d L=%d hidden=%d O=%d)", B, Cin, L, Cmid, O);
  // 1) conv1 up to the <Z> readouts (y == NULL: no post_conv); plane 1 of the (2, B*L, 4) scratch holds them
  if (int e = fast_forward(x, p1[0], p1[1], p1[2], p1[3], p1[4], nullptr, q1, d1, st)) return e;
  // 2) conv1.post_conv + GELU -> conv2 -> GELU -> (B, T, O) + positional embedding
  const int Lout = (L + 2 - 3) / 2 + 1;
  Stem2Args a{q1 + (size_t)B * L * SQ, p1[3], p1[4], p2[0], p2[1], p2[2], p2[3], p2[4], pos, out,
              B, Cmid, L, O, Lq, Lout, (Lout + STW - 1) / STW, 0, timeline_next_slot()};
  a.num_tiles = B * a.tiles_per_utt;
  const size_t smem = stem2_smem_bytes(Cmid, O, Lq);
  QW_CHECK_ARG(smem <= 227 * 1024, -2, "fused stem needs %zu bytes of shared memory", smem);
  QW_CUDA_OK(cudaFuncSetAttribute(stem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = a.num_tiles < 2 * num_sms() ? a.num_tiles : 2 * num_sms();
  {
    KernelTimer kt(kKStem2, st);
    QW_CUDA_OK(launch_pdl(true, stem2_kernel, dim3(grid), dim3(kStemThreads), smem, st, a));
  }
  QW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qw

extern "C" {

size_t qw_stem_workspace_bytes(int B, int L) { return (B > 0 && L > 0) ? qw::stem_workspace_bytes(B, L) : 0; }

int qw_stem_forward(const float* x, const float* w_pre1, const float* b_pre1, const float* qw1, const float* w_post1, const float* b_post1,
                    const float* w_pre2, const float* b_pre2, const float* qw2, const float* w_post2, const float* b_post2,
                    const float* pos_emb, float* out, void* workspace, size_t ws_bytes, int B, int C, int L, int hidden, int O,
                    int n_layers, void* stream) {
  QW_CHECK_ARG(x && w_pre1 && b_pre1 && qw1 && w_post1 && b_post1 && w_pre2 && b_pre2 && qw2 && w_post2 && b_post2 && out && workspace, -1,
               "null pointer argument");
  QW_CHECK_ARG(B > 0 && C > 0 && L >= 4 && hidden > 0 && O > 0, -1, "bad shape B=%d C=%d L=%d hidden=%d O=%d", B, C, L, hidden, O);
  QW_CHECK_ARG(n_layers >= 1 && n_layers <= 4, -2, "n_layers=%d must be in [1,4] for the fused stem", n_layers);
  QW_CHECK_ARG(((uintptr_t)workspace & 255) == 0, -1, "workspace must be 256-byte aligned");
  const float* p1[5] = {w_pre1, b_pre1, qw1, w_post1, b_post1};
  const float* p2[5] = {w_pre2, b_pre2, qw2, w_post2, b_post2};
  return qw::stem_forward(x, p1, p2, pos_emb, out, workspace, ws_bytes, B, C, L, hidden, O, n_layers, (cudaStream_t)stream);
}

}  // extern "C"
