/* qw.h -- C ABI of the B200-native Quantum-Whisper hot path (libqw_b200.so).
 *
 * Drop-in boundary for the reference's QuantumConv1d operator and the log-mel front end.  The reference
 * is pure Python (PennyLane QNode inside an nn.Module); its FFI for this path is therefore "whatever a
 * torch.autograd.Function can call with raw device pointers".  Each entry point cites the reference code it
 * replaces (paths relative to /root/reference/).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; tensors are dense row-major;
 *   - the library never allocates or frees device memory and keeps no pointer after a call returns
 *     (except *_host entry points, which use an internal per-device staging pool);
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*); no implicit sync;
 *   - return value: 0 ok; <0 bad argument (-1 null/shape, -2 unsupported configuration, -3 workspace too
 *     small); >0 a cudaError_t.  Nothing throws across the ABI.  qw_last_error() gives a thread-local text.
 *   - embedding: 0 = amplitude (the reference circuit), 1 = angle (extension).
 *   - quantum weights qw: (n_layers, q, 3) = [phi, theta, omega] per wire; n_layers = 1 is the reference.
 */
#ifndef QW_B200_H
#define QW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QW_EMB_AMPLITUDE 0
#define QW_EMB_ANGLE 1
#define QW_ABI_VERSION 1
#define QW_ACT_NONE 0
#define QW_ACT_GELU 1

/* ABI version / build info. */
int qw_abi_version(void);
/* Hash of the sources (csrc/, include/, compile flags) this library was built from; the Python loader compares it with the tree
 * next to it so that a stale prebuilt .so is rebuilt (or refused) instead of silently running old kernels. */
const char* qw_build_stamp(void);
const char* qw_last_error(void);
/* Number of kernels this library has launched from the calling process since load (for bench gpu_launches). */
long long qw_launch_count(void);
/* Optional per-kernel CUDA-event timing used by bench.py's roofline leg (not valid during stream capture).
 * qw_profile_read synchronises on the recorded events; kernel ids 0..qw_kernel_name()!="" . */
void qw_profile_enable(int on);
int qw_profile_read(int kernel_id, double* total_ms, long long* count, int reset);
const char* qw_kernel_name(int kernel_id);
/* The kernel symbol (as ncu prints it, e.g. "fast_fwd_kernel<2, 64>") last launched under that id; "" if none yet. */
const char* qw_kernel_symbol(int kernel_id);
/* Debug timeline: register a caller-owned DEVICE buffer of 2*nslots uint64 (pre-filled by the caller with ~0 for even and 0 for
 * odd entries).  Every fast-path QuantumConv1d kernel launched afterwards takes the next slot (round robin from 0) and records
 * min(CTA start) / max(CTA end) of %globaltimer in nanoseconds there -- also inside a replayed CUDA graph, which events cannot
 * see.  dev_buf = NULL turns it off (the default).  Tools only; not part of the reference-facing path. */
int qw_timeline_set(unsigned long long* dev_buf, int nslots);
/* 1 (default): use the TMA fast path when the shape qualifies (fp32, q=4, K=3, stride 1|2, L%4==0, L_out%4==0,
 * O%4==0, O<=576, 16-byte aligned tensors; the forward additionally needs padding==1); 0: always use the generic kernels (used by the parity tests). */
void qw_set_fast_path(int enable);
/* Kernel-selection switches (A/B experiments and parity tests ONLY: they choose between equivalent kernels and are not part of
 * the reference-facing contract).  name = "FAST_PATH", "GY_MMA", "BWD_FUSED", "FWD_ETMA", ... (csrc/qw_common.cuh, enum Option); each
 * is initialised from the environment variable QW_<name> on first use.  Process-global and NOT re-entrant: do not change an
 * option while another thread is inside a qw_* call.  Returns 0, or -1 for an unknown name. */
int qw_set_option(const char* name, int value);
int qw_get_option(const char* name); /* current value, or -1 for an unknown name */

/* ---- QuantumConv1d.forward  (quantum_whisper.py:95-128; circuit :64-85; params :58-59,88)
 * x (B,C,L) -> y (B,O,L_out), L_out = (L+2P-K)/S+1 (:103).  w_pre (q, C*K) with column c*K+k (:58,:111),
 * b_pre (q), qw (n_layers,q,3), w_post (O,q), b_post (O).  pre_save: (2, B*L_out, q) or NULL -- plane 0 the
 * pre_conv outputs, plane 1 the <Z_i> readouts, kept for the backward pass.  q must already be min(n_qubits, C*K) (:55). */
int qw_conv1d_forward(const float* x, const float* w_pre, const float* b_pre, const float* qw, const float* w_post,
                      const float* b_post, float* y, float* pre_save, int B, int C, int L, int K, int S, int P, int O,
                      int q, int n_layers, int embedding, void* stream);
int qw_conv1d_forward_f64(const double* x, const double* w_pre, const double* b_pre, const double* qw,
                          const double* w_post, const double* b_post, double* y, double* pre_save, int B, int C, int L,
                          int K, int S, int P, int O, int q, int n_layers, int embedding, void* stream);

/* ---- backward of the same (autograd through :107-126 and PennyLane backprop) by adjoint
 * differentiation.  gy (B,O,L_out).  gx (B,C,L) or NULL to skip (conv1's input is data).  Parameter
 * gradients are OVERWRITTEN (not accumulated).  workspace: qw_conv1d_workspace_bytes(...) bytes. */
size_t qw_conv1d_workspace_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int elem_size);
int qw_conv1d_backward(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qw,
                       const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post,
                       float* gb_post, void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P,
                       int O, int q, int n_layers, int embedding, void* stream);
int qw_conv1d_backward_f64(const double* gy, const double* x, const double* pre_save, const double* w_pre,
                           const double* qw, const double* w_post, double* gx, double* gw_pre, double* gb_pre,
                           double* gqw, double* gw_post, double* gb_post, void* workspace, size_t ws_bytes, int B,
                           int C, int L, int K, int S, int P, int O, int q, int n_layers, int embedding, void* stream);

/* ---- data-parallel backward: the same as qw_conv1d_backward, with the all-reduce (mean over ranks) of the five parameter
 * gradients FUSED into the backward's last kernel (the reference trains single-process, train_quantum_whisper.py,
 * so this replaces what DistributedDataParallel would add around quantum_whisper.py:95-128).  Every rank owns a receive buffer
 * of qw_conv1d_dp_buffer_bytes(..., world) bytes, zero-initialised once and mapped into every peer's address space (e.g.
 * torch.distributed._symmetric_memory), and a LOCAL zero-initialised bookkeeping buffer of qw_conv1d_dp_flag_bytes(...) bytes;
 * peer_bufs[r] is rank r's receive buffer as seen from the calling process, peer_flags[rank] the caller's own bookkeeping
 * buffer (the other entries are ignored).  Each rank stores (epoch, value) words straight into its peers' buffers over NVLink.
 * After the call (in stream order) gw_pre ... gb_post hold scale * sum over ranks, bitwise identical on every rank; gx stays
 * local.  All ranks must call it the same number of times with the same shapes.  Like an NCCL collective the kernel WAITS for a
 * late peer (checkpointing, evaluation, a data-loader stall); a peer that stays away longer than the DP_TIMEOUT_MS option (default
 * 600 000 ms, 0 = forever; qw_set_option) makes it set the last bookkeeping word to 1 and TRAP -- the CUDA context fails loudly,
 * it never continues with an un-averaged gradient.  Fast-path regime only
 * (-2 otherwise: use qw_conv1d_backward + qw_grads_allreduce_p2p / NCCL).  world == 1: plain backward. */
size_t qw_conv1d_dp_buffer_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int world);
size_t qw_conv1d_dp_flag_bytes(int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int world);
int qw_conv1d_backward_dp(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qw,
                          const float* w_post, float* gx, float* gw_pre, float* gb_pre, float* gqw, float* gw_post,
                          float* gb_post, void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O,
                          int q, int n_layers, int embedding, void* const* peer_bufs, void* const* peer_flags, int rank,
                          int world, float scale, void* stream);

/* ---- CHAINED backward: this layer's incoming gradient is the grad_x of the layer that FOLLOWS it in the stem -- a QuantumConv1d
 * with kernel_size 3, stride 2, padding 1, in_channels = this O, input length = this L_out (conv2 of quantum_whisper.py:137 after
 * conv1 of :136).  Run the following layer's backward first with gx = NULL (it then writes no (B, O, L_out) gradient at all), keep
 * its workspace, and call this instead of qw_conv1d_backward[_act | _dp]: the gy kernel rebuilds every 32-window tile of the
 * gradient from the following layer's gpre rows (16 bytes per window, inside next_workspace) and its pre_conv weights next_w_pre
 * (4, O*3), so the 4*B*O*L_out bytes are neither written nor read.  activation: QW_ACT_NONE | QW_ACT_GELU -- the activation between
 * the two layers (its derivative is applied to the rebuilt tile); b_post may be NULL without it.  world > 1: gradient mean over the
 * ranks as in qw_conv1d_backward_dp (peer tables may be NULL for world == 1).  Same outputs as the unchained calls up to fp32
 * summation order.  -2 outside the regime (fast path, n_qubits 4, option GY_MMA). */
int qw_conv1d_backward_chained(const void* next_workspace, const float* next_w_pre, int next_O, const float* x, const float* pre_save,
                               const float* w_pre, const float* qw, const float* w_post, const float* b_post, float* gx,
                               float* gw_pre, float* gb_pre, float* gqw, float* gw_post, float* gb_post, void* workspace,
                               size_t ws_bytes, int B, int C, int L, int K, int S, int P, int O, int q, int n_layers, int embedding,
                               int activation, void* const* peer_bufs, void* const* peer_flags, int rank, int world, float scale,
                               void* stream);

/* ---- the layer with the activation that follows it in the encoder stem FUSED (whisper/whisper/model.py:193-194:
 * x = F.gelu(self.conv1(x)); x = F.gelu(self.conv2(x)), exact-erf GELU): training-time counterpart of qw_stem_forward.
 * forward:  y = gelu(post_conv(<Z>)) is what gets stored (the pre-activation never exists in HBM); pre_save as qw_conv1d_forward.
 * backward: gy is the gradient wrt that ACTIVATED output; the gy kernel multiplies every staged tile by gelu'(post_conv(<Z>)),
 *           rebuilt from the 16 bytes per window in pre_save (needs b_post), before its contractions -- so neither the GELU forward
 *           (read + write of the (B,O,L_out) tensor) nor the GELU backward (two reads + a write) ever runs as a separate pass.
 * activation: QW_ACT_NONE (then identical to qw_conv1d_forward / qw_conv1d_backward) or QW_ACT_GELU.  Fast-path regime only
 * (fp32, n_qubits = 4, amplitude embedding, K = 3, stride 1|2, padding 1, aligned shapes): -2 otherwise -- apply the plain operator
 * and a separate GELU instead (qasr_ijcnlp_b200.QuantumConv1d.forward_gelu does). */
int qw_conv1d_forward_act(const float* x, const float* w_pre, const float* b_pre, const float* qw, const float* w_post,
                          const float* b_post, float* y, float* pre_save, int B, int C, int L, int K, int S, int P, int O, int q,
                          int n_layers, int embedding, int activation, void* stream);
int qw_conv1d_backward_act(const float* gy, const float* x, const float* pre_save, const float* w_pre, const float* qw,
                           const float* w_post, const float* b_post, float* gx, float* gw_pre, float* gb_pre, float* gqw,
                           float* gw_post, float* gb_post, void* workspace, size_t ws_bytes, int B, int C, int L, int K, int S, int P,
                           int O, int q, int n_layers, int embedding, int activation, void* stream);

/* ---- fused INFERENCE forward of the encoder stem (whisper/whisper/model.py:193-198 with the two QuantumConv1d layers of
 * quantum_whisper.py:136-137):
 *     out[b, t, :] = gelu(conv2(gelu(conv1(x))))[b, :, t] + pos_emb[t, :]
 * x (B,C,L) -> out (B, L/2, O); conv1 = (C -> hidden, K=3, S=1, P=1), conv2 = (hidden -> O, K=3, S=2, P=1), n_qubits = 4,
 * amplitude embedding, gelu = the exact erf form (torch default).  pos_emb (L/2, O) or NULL.  The (B, hidden, L) activation
 * between the layers is never materialised.  Nothing is saved for a backward pass.  Fast-path regime only (-2 otherwise):
 * L % 4 == 0, hidden % 4 == 0, O % 4 == 0, hidden, O <= 576, 16-byte aligned tensors. */
size_t qw_stem_workspace_bytes(int B, int L);
int qw_stem_forward(const float* x, const float* w_pre1, const float* b_pre1, const float* qw1, const float* w_post1,
                    const float* b_post1, const float* w_pre2, const float* b_pre2, const float* qw2, const float* w_post2,
                    const float* b_post2, const float* pos_emb, float* out, void* workspace, size_t ws_bytes, int B, int C,
                    int L, int hidden, int O, int n_layers, void* stream);

/* ---- fused TRAINING forward of the stem: y1 = act(conv1(x)) (B, hidden, L), y2 = act(conv2(y1)) (B, O, L/2) in ONE kernel, with
 * pre_save1 (2, B*L, 4) and pre_save2 (2, B*L/2, 4) exactly as two qw_conv1d_forward_act calls leave them (same values to fp32
 * rounding: conv2's pre_conv sums in a different order), so the layers' qw_conv1d_backward[_act] calls follow unchanged.  y1 is
 * written (the backward needs it) but never read back: conv2's pre_conv is taken from the registers that store it.  conv1 =
 * (C -> hidden, K=3, S=1, P=1), conv2 = (hidden -> O, K=3, S=2, P=1), n_qubits = 4, amplitude embedding; activation QW_ACT_NONE |
 * QW_ACT_GELU (both layers).  Regime: C <= 96, L % 8 == 0, hidden % 4 == 0, O % 8 == 0, hidden, O <= 384, 16-byte aligned tensors;
 * -2 otherwise (run the two layers separately). */
/* 1 when the one-kernel forward is expected to beat two qw_conv1d_forward_act calls for this batch (few tiles per CTA: launches are
 * latency-dominated), 0 when the two leaner kernels win (measured crossover between batch 16 and 32 of 30 s audio on B200). */
int qw_stem_train_forward_preferred(int B, int L);
int qw_stem_train_forward(const float* x, const float* w_pre1, const float* b_pre1, const float* qw1, const float* w_post1,
                          const float* b_post1, const float* w_pre2, const float* b_pre2, const float* qw2, const float* w_post2,
                          const float* b_post2, float* y1, float* pre_save1, float* y2, float* pre_save2, int B, int C, int L,
                          int hidden, int O, int n_layers, int activation, void* stream);

/* ---- the QNode alone (quantum_whisper.py:64-85), batched over W windows: pre (W,q) -> out (W,q);
 * backward: gout (W,q) -> gpre (W,q) and gqw (n_layers,q,3) (overwritten).  BASELINE.json config 4. */
size_t qw_circuit_workspace_bytes(long long W, int q, int n_layers, int elem_size);
int qw_circuit_forward(const float* pre, const float* qw, float* out, long long W, int q, int n_layers, int embedding,
                       void* stream);
int qw_circuit_backward(const float* pre, const float* qw, const float* gout, float* gpre, float* gqw, void* workspace,
                        size_t ws_bytes, long long W, int q, int n_layers, int embedding, void* stream);
int qw_circuit_forward_f64(const double* pre, const double* qw, double* out, long long W, int q, int n_layers,
                           int embedding, void* stream);
int qw_circuit_backward_f64(const double* pre, const double* qw, const double* gout, double* gpre, double* gqw,
                            void* workspace, size_t ws_bytes, long long W, int q, int n_layers, int embedding,
                            void* stream);

/* ---- OPT-IN collapsed evaluation of the amplitude-embedded circuit (SURVEY.md 8a iii): <Z_i> = xh^T M_i xh, xh = pre/||pre||, with the
 * q x q x q matrices M ([i][a][b], symmetric in a, b) supplied by the caller -- qasr_ijcnlp_b200.quantum_circuit(simulator=
 * "collapsed") reads them off the STATEVECTOR kernel evaluated on q(q+1)/2 probe windows, so the semantics (any weights, any
 * n_layers) stay those of qw_circuit_forward.  pre (W,q) -> out (W,q); backward: gout -> gpre (W,q) and gM (q,q,q) = sum over
 * windows of gout_i xh_a xh_b.  Not the headline path and not used by QuantumConv1d: a separately reported mode and the
 * device-side second oracle of the parity tests.  n_qubits 1..12, amplitude embedding only. */
size_t qw_circuit_collapsed_workspace_bytes(long long W, int q, int elem_size);
int qw_circuit_forward_collapsed(const float* pre, const float* M, float* out, long long W, int q, void* stream);
int qw_circuit_backward_collapsed(const float* pre, const float* M, const float* gout, float* gpre, float* gM, void* workspace,
                                  size_t ws_bytes, long long W, int q, void* stream);
int qw_circuit_forward_collapsed_f64(const double* pre, const double* M, double* out, long long W, int q, void* stream);
int qw_circuit_backward_collapsed_f64(const double* pre, const double* M, const double* gout, double* gpre, double* gM,
                                      void* workspace, size_t ws_bytes, long long W, int q, void* stream);

/* ---- whisper.log_mel_spectrogram (whisper/whisper/audio.py:110-157), batched, max taken per utterance.
 * audio (B, n_samples) fp32, any n_samples > 200 (like the reference: T = n_samples / 160 frames, the samples past the last full
 * hop still feed the last frames); filters (n_mels, 201) fp32 (audio.py:91-107); mel (B, n_mels, n_samples/160).
 * workspace: qw_log_mel_workspace_bytes(B, n_samples, n_mels). */
size_t qw_log_mel_workspace_bytes(int B, int n_samples, int n_mels);
int qw_log_mel(const float* audio, const float* filters, float* mel, void* workspace, size_t ws_bytes, int B,
               int n_samples, int n_mels, void* stream);
/* The same in two steps, for callers that keep the filterbank (mel_filters(), audio.py:91-107, is a constant asset):
 * qw_log_mel_prepare analyses `filters` once into `prep` (qw_log_mel_prep_bytes(n_mels) bytes, 256-byte aligned, caller-owned);
 * qw_log_mel_prepared then skips the analysis kernel on every call.  Its workspace is qw_log_mel_call_workspace_bytes(B, n_samples)
 * bytes, 256-byte aligned: B floats for the utterance maxima plus per-tile (min, max) statistics that let the second pass
 * (audio.py:155-156) skip every tile the clamp cannot touch.  A caller that passes only the B floats (>= 4 B bytes, the first
 * version of this interface) still gets the same result through an unconditional second pass. */
size_t qw_log_mel_prep_bytes(int n_mels);
size_t qw_log_mel_call_workspace_bytes(int B, int n_samples);
int qw_log_mel_prepare(const float* filters, int n_mels, void* prep, size_t prep_bytes, void* stream);
int qw_log_mel_prepared(const float* audio, const void* prep, float* mel, void* workspace, size_t ws_bytes, int B,
                        int n_samples, int n_mels, void* stream);
/* whisper.pad_or_trim (audio.py:65-88) fused into the front end -- the batched data path of SURVEY.md 8-f4 (the reference pads every
 * clip to 480 000 samples on the CPU inside Dataset.__getitem__, train_quantum_whisper.py:52-77): audio is (B, n_in) with row
 * stride n_in; utterance b holds lengths[b] valid samples (lengths == NULL: all n_in); the result is exactly
 * log_mel_spectrogram(pad_or_trim(audio[b, :lengths[b]], n_samples)): zero padding up to n_samples, or the first n_samples samples of
 * a longer clip.  Tiles that lie entirely in the padding skip the FFT.  mel (B, n_mels, n_samples/160); workspace as for
 * qw_log_mel_prepared (sized with the OUTPUT length n_samples). */
int qw_log_mel_padded(const float* audio, const int* lengths, const void* prep, float* mel, void* workspace, size_t ws_bytes, int B,
                      int n_in, int n_samples, int n_mels, void* stream);

/* ---- data-parallel training collective (the reference is single-process, train_quantum_whisper.py:195-214):
 * one-shot all-reduce of a small fp32 gradient bucket (n <= 2^20) over NVLink peer memory, fused with the `scale` (1/world)
 * multiply.  grads (n) is reduced in place.  peer_bufs is a HOST array of `world` device pointers: rank r's receive buffer
 * (qw_grads_allreduce_p2p_buffer_bytes(n, world) bytes, zero-initialised once, mapped into this process e.g. by
 * torch.distributed._symmetric_memory); peer_flags[rank] is the caller's own zero-initialised bookkeeping array
 * (qw_grads_allreduce_p2p_flag_bytes(world) bytes; the other entries are ignored).  Each rank stores (epoch, value) words
 * straight into its peers' buffers (no fence, no remote read).  Every rank must enqueue the call the same number of times; the
 * kernel keeps its epoch in the bookkeeping array, so it is graph-capturable.  Like NCCL it waits for a late peer; a peer that
 * stays away longer than the DP_TIMEOUT_MS option (default 600 000 ms, 0 = forever) makes it set the last bookkeeping word to 1 and
 * trap (loud launch failure; the gradient is never left un-averaged). */
size_t qw_grads_allreduce_p2p_buffer_bytes(long long n, int world);
size_t qw_grads_allreduce_p2p_flag_bytes(int world);
int qw_grads_allreduce_p2p(float* grads, long long n, void* const* peer_bufs, void* const* peer_flags, int rank, int world,
                           float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QW_B200_H */
