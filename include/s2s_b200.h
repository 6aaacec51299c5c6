/*
 * s2s_b200.h -- C ABI of libs2s_b200.so: the B200 (sm_100a) implementation of the
 * seq2seq attention-ASR training hot path of Ajay-Wong/seq2seq-attention-asr.
 *
 * The reference has no FFI/plugin interface of its own (it is pure Lua on Torch7).  The
 * boundary this library sits behind is the Torch7 nn.Module protocol of the reference's custom
 * modules; every entry point below names the reference method (file:line under the reference
 * tree) it replaces.  The LuaJIT-FFI binding a maintainer adds is shown in INTEGRATION.md and
 * lives in seq2seq-attention-asr_b200/lua/.
 *
 * Conventions
 *   - plain C, no torch / C++ types.  Every pointer named `const float*` / `float*` /
 *     `const int*` is a DEVICE pointer unless the name ends in `_host`.
 *   - fp32, row-major, batch-first: annotations h [B, Lmax, A], features X [B, Lmax, D],
 *     labels [B, Tmax] (0-based class ids), lengths [B] / tlens [B] (valid frames / labels per
 *     utterance; NULL = all Lmax / Tmax).  Single-utterance ("SGD mode") calls are B = 1.
 *   - every function returns 0 on success, non-zero on error; the message is available from
 *     s2s_last_error() (thread-local).  Nothing throws, nothing frees caller memory, no caller
 *     pointer is retained after the call returns (except the stream given to the context).
 *   - all work is enqueued on the context's stream; calls are asynchronous w.r.t. the host
 *     unless stated otherwise.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef S2S_B200_H
#define S2S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define S2S_VERSION 100

typedef struct s2s_ctx s2s_ctx;

/* Mirrors the `model.*` fields set by timit/model_chorowski_baseline.lua:14-46 (and
 * librispeech/model_chorowski_baseline.lua).  A (annotation depth) = 2*H. */
typedef struct s2s_model_cfg {
    int D;   /* inputFrameSize        (123)  model_chorowski_baseline.lua:15 */
    int H;   /* hiddenFrameSize = outputFrameSize (256)  :16-17              */
    int NL;  /* bidirectional encoder layers (3)  :22-31                     */
    int S;   /* scoreDepth            (512)  :37                             */
    int ST;  /* stateDepth            (256)  :41                             */
    int V;   /* outputDepth           (62)   :43                             */
    int K;   /* hybridAttendFeatureMaps (0 = content-only)  :40              */
    int KF;  /* hybridAttendFilterSize (10)  :39                             */
    int M;   /* mlpDepth              (64)   :44                             */
    int MW;  /* maxout window         (7)    :56                             */
    int MLP; /* decoder MLP stages: 0/1 = Maxout-Linear (:56-57); 2 = Maxout-Linear-Maxout-Linear
                (librispeech/model_vgg.lua:76-80).  NL = 0 gives a decoder-only parameter vector
                (annotations from another encoder, e.g. s2s_vgg_forward).                          */
} s2s_model_cfg;

/* flags for the loss / gradient seed (timit/timit.lua:268-281) */
#define S2S_NORMALIZE_NLL   1   /* nll /= T_b            (timit.lua:270-272) */
#define S2S_NORMALIZE_GRAD  2   /* dlogp = -labelmask/T_b (timit.lua:279-281) */

/* ---- context ------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t; NULL = the legacy default stream (cutorch's default, timit/timit.lua:39). */
int  s2s_ctx_create(int device, void* stream, s2s_ctx** out);
int  s2s_ctx_destroy(s2s_ctx* ctx);
int  s2s_ctx_set_stream(s2s_ctx* ctx, void* stream);
int  s2s_ctx_synchronize(s2s_ctx* ctx);
const char* s2s_last_error(void);
int  s2s_version(void);
/* number of kernels launched by this context since creation (bench.py `gpu_launches`) */
int64_t s2s_ctx_launch_count(s2s_ctx* ctx);
/* launches per kernel class since creation: lets tests and the bench assert WHICH path ran (tcgen05 vs SIMT GEMM,
 * persistent cluster loops vs per-step chains); replayed CUDA graphs add the counts recorded at capture */
#define S2S_KC_GEMM_TC          0   /* gemm_tc_kernel (tcgen05 / TMEM / TMA)                          */
#define S2S_KC_GEMM_SIMT        1   /* gemm_simt_kernel (exact fp32)                                   */
#define S2S_KC_GRU_CLUSTER      2   /* persistent GRU recurrence (one launch per layer and pass)       */
#define S2S_KC_DEC_CLUSTER_FWD  3   /* decoder time loop as one cluster kernel                         */
#define S2S_KC_DEC_CLUSTER_BWD  4   /* decoder backward time loop as one cluster kernel                */
#define S2S_KC_ATTN_STEP        5   /* per-step attention kernels (attn_fwd_kernel / attn_bwd_kernel)  */
#define S2S_KC_LSTM_CLUSTER     6   /* persistent LSTM recurrence                                       */
#define S2S_KC_N                7
int64_t s2s_ctx_kernel_count(s2s_ctx* ctx, int kernel_class);
/* enable (1) / disable (0) CUDA-graph replay of s2s_model_fwdbwd for repeated shapes */
int  s2s_ctx_set_graphs(s2s_ctx* ctx, int enable);
/* Caller-defined graphs: every library call on this context between _begin and _end is captured (not executed) and
 * replayed by _launch in the order of the context's stream.  Same pointers and shapes on every replay; run the
 * sequence eagerly once before capturing; no host read-backs inside.  Workspaces cannot grow while a graph is alive. */
int  s2s_graph_begin(s2s_ctx* ctx);
int  s2s_graph_end(s2s_ctx* ctx, int* graph_id);
int  s2s_graph_launch(s2s_ctx* ctx, int graph_id);
int  s2s_graph_destroy(s2s_ctx* ctx, int graph_id);

/* Per-kernel-class device timing with CUDA events on the context's stream (bench.py roofline).
 * enable != 0 clears the counters and starts recording; read synchronises the stream and returns, per
 * class, the summed kernel time [ms], the launch count and the algorithmic work (bytes for the
 * HBM-bound classes, FLOPs for S2S_PROF_GEMM) the launches were asked to do. */
#define S2S_PROF_ATTN_FWD    0
#define S2S_PROF_ATTN_BWD    1
#define S2S_PROF_ATTN_DVH    2
#define S2S_PROF_GRU_FWD     3
#define S2S_PROF_GRU_BWD     4
#define S2S_PROF_GEMM        5
#define S2S_PROF_DENSE_SMALL 6
#define S2S_PROF_DEC_FWD     7   /* the decoder time loop as one cluster kernel (decoder_cluster.cu) */
#define S2S_PROF_DEC_BWD     8   /* the decoder backward time loop as one cluster kernel */
#define S2S_PROF_GEMM_SIDE   9   /* GEMMs issued on the low-priority side stream (weight gradients on the SMs the cluster kernels leave idle):
                                    overlapped with the recurrences, not on the critical path of the step */
#define S2S_PROF_N           10
int  s2s_ctx_profile(s2s_ctx* ctx, int enable);
int  s2s_ctx_profile_read(s2s_ctx* ctx, double* ms_host, int64_t* count_host, double* work_host);

/* ---- TrainUtils.orthogonalize (TrainUtils.lua:5-26; orthogonalizeGraph applies it to every module with a weight,
 * librispeech/exp0_scriptchecker.lua:49-52).  In place: W [rows, cols] (and bias [rows] when not NULL, treated as one more column)
 * := the orthonormal factor of the QR decomposition taken in the tall orientation (rows >= cols: qr(w).Q, else qr(w^T).Q^T), with
 * LAPACK's Householder sign convention (what torch.qr returns). */
int s2s_orthogonalize(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, float* bias);

/* ---- data-parallel plane: NCCL over NVLink 5 / NVSwitch (no counterpart in the reference: single device, timit/timit.lua:39) ----
 * Only the minibatch shards (timit/timit.lua:240-295 sums per-utterance gradients): every rank holds the full parameters and
 * optimiser state, runs s2s_model_fwdbwd on its shard, the flat gradient is summed over the ranks, and the gradient step
 * (s2s_grad_finalize with the GLOBAL batch size, s2s_adadelta, s2s_model_rownorm_constraint; timit.lua:291-348) is replicated.
 * libnccl.so.2 is opened at run time ($S2S_NCCL_LIB overrides the search); hosts that never call these need no NCCL.
 *   rank 0: s2s_dp_unique_id(id)  ->  the host ships the 128 bytes to every rank  ->  all ranks: s2s_dp_init(ctx, rank, world, id)   */
int  s2s_dp_available(void);                                           /* 1 when libnccl could be opened */
int  s2s_dp_unique_id(void* id_host_128);                              /* ncclGetUniqueId: 128 bytes, host memory */
int  s2s_dp_init(s2s_ctx* ctx, int rank, int world, const void* id_host_128);
int  s2s_dp_rank(s2s_ctx* ctx);
int  s2s_dp_world(s2s_ctx* ctx);
/* G[0, n) := sum over ranks, in place, on the context's stream ("gradients" of timit.lua:229 before :292-295).  No-op for world 1. */
int  s2s_dp_allreduce(s2s_ctx* ctx, float* G, int64_t n);
/* P[0, n) := rank `root`'s copy (initial parameter synchronisation) */
int  s2s_dp_broadcast(s2s_ctx* ctx, float* P, int64_t n, int root);
/* enable != 0: s2s_model_fwdbwd itself sums G over the ranks, bucket by bucket (decoder, then each encoder layer) on a low-priority
 * side stream under the remaining backward pass; on return (in stream order) G is the global gradient sum and s2s_dp_allreduce must
 * not be called again for it. */
int  s2s_dp_set_overlap(s2s_ctx* ctx, int enable);
int  s2s_dp_destroy(s2s_ctx* ctx);                                     /* releases the context's CUDA graphs first (captured collectives reference the communicator) */

/* ---- flat parameter layout (what module:getParameters() flattens to; timit/timit.lua:172) -- */
int64_t s2s_param_count(const s2s_model_cfg* cfg);
/* writes (offset, rows, cols) triples in flat order into out_host[3*max]; returns the count */
int     s2s_param_segments(const s2s_model_cfg* cfg, int64_t* out_host, int max);
/* offset of the first decoder (nn.Attention) parameter = W_V */
int64_t s2s_decoder_param_offset(const s2s_model_cfg* cfg);

/* ---- nn.TemporalConvolutionZeroBias, kW = 1 (TemporalConvolutionZeroBias.lua:37-54) -------- */
/* updateOutput:  y[rows,out] = x[rows,in] . W[out,in]^T   (bias pinned to zero, :38)           */
int s2s_tconv_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y);
/* updateGradInput + accGradParameters: dx[rows,in] = dy . W (overwritten; NULL = skip),
 * dW[out,in] += scale * dy^T . x (NULL = skip); gradBias stays zero (:52-53)                   */
int s2s_tconv_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out,
                          const float* dy, float* dx, float* dW, float scale);

/* ---- nn.LinearZeroBias (LinearZeroBias.lua:31-74) ------------------------------------------ */
int s2s_linear_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y);
int s2s_linear_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out,
                           const float* dy, float* dx, float* dW, float scale);

/* ---- nn.RNN(nn.GRU(in,out), reverse) over whole utterances (RNN.lua:120-201, GRU.lua:8-51) -- */
/* W: the three LinearZeroBias weights (z, r, h~) each [H, H+Din], concat order {prev_h, x}
 * (GRU.lua:22-26), contiguous in that order.  ndir = 1: one direction given by `reverse`;
 * ndir = 2: W holds forward then reverse weights (6 matrices) and the outputs land in the two
 * halves of y as JoinTable(2,2) does (model_chorowski_baseline.lua:24).
 * x [B,Lmax,Din] (row stride ldx), y [B,Lmax,ndir*H].  State needed by backward is kept in
 * `save` (caller-provided, s2s_gru_seq_save_floats(...) floats). */
int64_t s2s_gru_seq_save_floats(int B, int Lmax, int H, int ndir);
int s2s_gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse,
                        const float* x, int ldx, const int* lengths, int B, int Lmax,
                        float* y, float* save);
/* dy [B,Lmax,ndir*H] -> dx [B,Lmax,Din] (overwritten; NULL = skip); dW accumulated (same layout as W) */
int s2s_gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse,
                         const float* x, int ldx, const int* lengths, int B, int Lmax,
                         const float* y, const float* save, const float* dy, float* dx);

/* ---- nn.GRU as a single-step module (GRU.lua:8-51 via nn.Recurrent, Recurrent.lua:104-151) ----------------- */
/* updateOutput({x, prev_h}) -> h ; hprev NULL = zeros (Recurrent.lua:110-112); gates [B,3H] (z | r | h~) kept by the caller */
int s2s_gru_step_forward(s2s_ctx* ctx, const float* W, int Din, int H, const float* x, const float* hprev, int B,
                         float* hnext, float* gates);
/* updateGradInput: dhnext -> dx [B,Din], dhprev [B,H] (overwritten); dW accumulated */
int s2s_gru_step_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, const float* x, const float* hprev, int B,
                          const float* gates, const float* dhnext, float* dx, float* dhprev);

/* ---- nn.LSTM as a single-step module (LSTM.lua:100-136): input {x, prev_h, prev_c} -> {next_h, next_c} ---------- */
/* hprev / cprev NULL = zeros (LSTM.lua:108-109); acts [B,4H] (i | f | g | o) kept by the caller for the backward call */
int s2s_lstm_step_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, const float* x, const float* hprev,
                          const float* cprev, int B, float* hnext, float* cnext, float* acts);
/* dhnext, dcnext (NULL = zeros) -> dx [B,Din], dhprev, dcprev [B,H] (overwritten); dP accumulated */
int s2s_lstm_step_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, const float* x,
                           const float* hprev, const float* cprev, int B, const float* acts, const float* cnext,
                           const float* dhnext, const float* dcnext, float* dx, float* dhprev, float* dcprev);

/* ---- nn.RNN(nn.LSTM(in,out,peepholes), reverse) over whole utterances (LSTM.lua:6-136, RNN.lua:120-201) ---- */
/* P: flat LSTM parameters in the order the module's parameters() yields: for gate in (i, f, g, o):
 * Wx[out,in], bx[out], Wh[out,out], bh[out], and -- with peepholes, except for g -- Wc[out,out], bc[out]
 * (full-matrix peepholes on prev_c for i,f and on next_c for o; LSTM.lua:25-50).  save: s2s_lstm_seq_save_floats floats. */
int64_t s2s_lstm_param_count(int in, int out, int peepholes);
int64_t s2s_lstm_seq_save_floats(int B, int Lmax, int H);
int s2s_lstm_seq_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, int reverse,
                         const float* x, int ldx, const int* lengths, int B, int Lmax, float* y, float* save);
int s2s_lstm_seq_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, int reverse,
                          const float* x, int ldx, const int* lengths, int B, int Lmax,
                          const float* y, const float* save, const float* dy, float* dx);

/* ---- nn.Attention (Attention.lua:305-327) = Vh + nn.RNNAttention(nn.Recurrent(decoder_base_)) */
/* updateOutput: teacher-forced decoder over Tmax steps (RNNAttention.lua:144-185).
 * P = flat parameter vector of the WHOLE model (decoder segments located through cfg).
 * dropmask: NULL or [B,Tmax,ST+A] multiplicative mask on {s,c} before the Maxout
 * (model_chorowski_baseline_dropout.lua:56).  lambda = monotonic-alignment penalty weight
 * (MonotonicAlignment.lua:19-77).  logp [B,Tmax,V].  Forward state (alpha, s, c, Ws, Vh, ...)
 * is kept inside ctx for the matching backward and for s2s_attention_get. */
int s2s_attention_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                          const float* h, const int* lengths, int B, int Lmax,
                          const int* labels, const int* tlens, int Tmax,
                          const float* dropmask, float lambda, float* logp);
/* updateGradInput (+ parameter gradients, which the reference accumulates inside
 * updateGradInput: Attention.lua:325, Recurrent.lua:148).  Must follow the forward with the
 * same arguments.  dlogp [B,Tmax,V]; G (flat, whole model) is ACCUMULATED; dh [B,Lmax,A]
 * overwritten. */
int s2s_attention_backward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G,
                           const float* h, const int* lengths, int B, int Lmax,
                           const int* labels, const int* tlens, int Tmax,
                           const float* dropmask, float lambda, const float* dlogp, float* dh);
/* introspection used by the reference's callers (Attention.lua:214-250, timit/timit.lua:521) */
#define S2S_GET_ALPHA    0   /* decoder:alpha()   [B,Tmax,Lmax] */
#define S2S_GET_WS       1   /* decoder:Ws()      [B,Tmax,S]    */
#define S2S_GET_VH       2   /* decoder.Vh.output [B,Lmax,S]    */
#define S2S_GET_PENALTY  3   /* decoder:penalty() [B,Tmax]      */
#define S2S_GET_STATE    4   /* s_t               [B,Tmax,ST]   */
#define S2S_GET_CONTEXT  5   /* c_t               [B,Tmax,A]    */
int s2s_attention_get(s2s_ctx* ctx, int what, float* dst);

/* One decoder step with explicit hidden state in/out = decoder_base:forward (Attention.lua:366,402),
 * the building block of Attention:BeamSearch.  Vh must have been computed (s2s_tconv_zb_forward
 * with W_V).  yprev[B] = previous label (-1 = zeros_y, RNNAttention.lua:173); alpha_prev / s_prev
 * NULL = zeros (Recurrent.lua:112). */
int s2s_attention_step(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                       const float* h, const float* Vh, const int* lengths, int B, int Lmax,
                       const int* yprev, const float* alpha_prev, const float* s_prev,
                       float* alpha, float* s, float* logp);
/* Attention:BeamSearch (Attention.lua:332-438) for one utterance.  The beams are a batch through the decoder step and the whole
 * search state (scores, top-k selection :406-408, finished list :418-421, label sequences) stays on the device; the host reads a
 * 16-byte status every 8 labels and the winning hypothesis (:435) once.  h [L,A]; writes up to maxlen + 1 labels to out_host
 * (the first label plus maxlen extensions); returns the length through n_out_host and the total log-probability through
 * logp_out_host (may be NULL). */
int s2s_beam_search(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* h, int L,
                    int eos, int beam, int maxlen, int* out_host, int* n_out_host, float* logp_out_host);

/* ---- whole model: encoder -> nn.Attention -> loss (timit/timit.lua:262-282) ------------------ */
/* forward only: nll [B] (device), logp [B,Tmax,V] (NULL = not wanted) */
int s2s_model_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                      const float* X, const int* lengths, int B, int Lmax,
                      const int* labels, const int* tlens, int Tmax,
                      const float* dropmask, float lambda, int flags, float* nll, float* logp);
/* forward + backward of a minibatch; G (flat) is ACCUMULATED (caller zeroes it, as
 * autoencoder:zeroGradParameters() does, timit.lua:233); dX NULL or [B,Lmax,D]. */
int s2s_model_fwdbwd(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G,
                     const float* X, const int* lengths, int B, int Lmax,
                     const int* labels, const int* tlens, int Tmax,
                     const float* dropmask, float lambda, int flags,
                     float* nll, float* logp, float* dX);
/* encoder annotations of the last model call [B,Lmax,A] (encoder.output, timit.lua:397) */
int s2s_model_get_annotations(s2s_ctx* ctx, float* dst);
/* labelmask <-> labels on the device (timit/timit.lua:262 builds labelmask = one-hot(Y) [T,V] / [B,T,V] and feeds it to nn.Attention):
 * labels[r] = index of the positive entry of row r, -1 for an all-zero row (a padded step, or prev_y at t = 1: RNNAttention.lua:172-176) */
int s2s_labels_from_onehot(s2s_ctx* ctx, const float* onehot, int64_t rows, int V, int* labels);
int s2s_onehot(s2s_ctx* ctx, const int* labels, int64_t rows, int V, float* onehot);

/* Loss and gradient seed on their own (timit/timit.lua:262-282), for callers that compose an encoder
 * (e.g. s2s_vgg_forward) with s2s_attention_forward / _backward: nll [B] and/or dlogp [B,T,V] (either may be NULL) */
int s2s_nll_grad_seed(s2s_ctx* ctx, const float* logp, const int* labels, const int* tlens, int B, int T, int V,
                      int flags, float* nll, float* dlogp);

/* ---- weight noise (WeightNoise.lua:17-35, AdaptiveWeightNoise.lua:27-104) ------------------- */
/* eps: injected N(0,1) buffer [n] for parity; NULL = Philox counter RNG seeded by (seed, call counter) */
int s2s_weightnoise_sample(s2s_ctx* ctx, const float* w, const float* eps, uint64_t seed, float sigma, int64_t n, float* sample);
/* weight = [mu ; s = log sigma^2] (2n) */
int s2s_awn_sample(s2s_ctx* ctx, const float* weight, const float* eps, uint64_t seed, int64_t n, float* sample);
/* L = lambda*KL + nll (AdaptiveWeightNoise.lua:63-80); result written to *L_host (synchronises) */
int s2s_awn_forward(s2s_ctx* ctx, const float* weight, int64_t n, double lambda, double nll, double* L_host);
/* gradWeight [2n] overwritten (AdaptiveWeightNoise.lua:82-104); g = dNLL/dw [n] */
int s2s_awn_accgrad(s2s_ctx* ctx, const float* weight, const float* g, int64_t n, double lambda, float* gradWeight);

/* nn.Dropout mask for {s,c} (model_chorowski_baseline_dropout.lua:56): 0 with probability p, else 1/(1-p)
 * (Torch nn.Dropout v2 scaling); Philox counter RNG seeded by (seed, call counter). */
int s2s_dropout_mask(s2s_ctx* ctx, float p, uint64_t seed, int64_t n, float* mask);

/* ---- VGG front-end of librispeech/model_vgg.lua:23-54 (the encoder of BASELINE configs[3]) -----------------------
 * 4 x [SpatialConvolutionMM 3x3 + ReLU] with SpatialMaxPooling(2,1,2,1) / (2,2,2,2), Transpose2 + View,
 * 4 x [TemporalConvolution(k=1) + ReLU].  X [B,3,T,F] (planes, time, frequency) -> h [B, L, OUT], L = (T-8)/2.
 * P: the flat encoder:parameters() order (conv1..4 {weight [nOut, nIn*9], bias}, then the four 1x1 layers {weight, bias}). */
typedef struct s2s_vgg_cfg { int C1, C2, HID, OUT; } s2s_vgg_cfg;   /* 64, 128, 2048, 512 in model_vgg.lua */
int64_t s2s_vgg_param_count(const s2s_vgg_cfg* cfg, int F);
int s2s_vgg_out_len(int T);
int s2s_vgg_forward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, const float* X, int B, int T, int F, float* h);
/* dh [B,L,OUT] -> dP accumulated, dX [B,3,T,F] (nullable); needs the preceding s2s_vgg_forward on the same ctx */
int s2s_vgg_backward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, float* dP, int B, int T, int F,
                     const float* dh, float* dX);

/* ---- gradient step (timit/timit.lua:291-348, TrainUtils.lua:52-104, optim.adadelta) ---------- */
/* g /= batch ; norm ; clip to maxnorm ; g += wd*p ; g += noise_sigma*N(0,1).  The pre-clip norm is
 * written to *gradnorm_host (synchronises) unless NULL. */
int s2s_grad_finalize(s2s_ctx* ctx, float* g, const float* p, int64_t n, int batch, double maxnorm, double wd,
                      const float* noise, uint64_t seed, double noise_sigma, double* gradnorm_host);
int s2s_adadelta(s2s_ctx* ctx, float* x, const float* g, float* v, float* a, int64_t n, double rho, double eps);
/* TrainUtils.columnNormConstraint on one weight matrix (per-ROW L2 norm; TrainUtils.lua:63-85).
 * *nan_host = 1 if a NaN row norm was seen (the reference error()s, :61). */
int s2s_rownorm_constraint(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, double maxval, int* nan_host);
/* columnNormConstraintGraph over every weight matrix of the model (timit.lua:346-348) */
int s2s_model_rownorm_constraint(s2s_ctx* ctx, const s2s_model_cfg* cfg, float* P, double maxval, int* nan_host);

/* WagnerFischer(a, b) (utils.lua:3-27): edit distance between two host label sequences (PER / CER of decodes) */
int s2s_edit_distance(const int* a, int na, const int* b, int nb, int* dist_host);

/* ---- test hooks (not part of the reference surface) ----------------------------------------- */
/* C = alpha * op(A) . op(B) + beta * C (+ bias[n]);  op(A) = A[M,K] (tA=0) or A[K,M]^T (tA=1);
 * op(B) = B[K,N] (tB=0) or B[N,K]^T (tB=1).  impl: 0 = auto, 1 = SIMT fp32, 2 = tcgen05 3xTF32 */
int s2s_gemm_f32(s2s_ctx* ctx, int impl, int tA, int tB, int M, int N, int K, float alpha,
                 const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias);
/* one attention scoring step (score + softmax + context), forward: the K1 microbenchmark kernel */
int s2s_attn_step_forward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                          const int* lengths, int B, int Lmax, int S, int A, float* alpha, float* c);
/* backward of the same step (K2): given dc [B,A], dalpha_in [B,Lmax] (NULL = 0) -> dq [B,S], de [B,Lmax] */
int s2s_attn_step_backward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                           const int* lengths, int B, int Lmax, int S, int A, const float* alpha,
                           const float* dc, const float* dalpha_in, float* dq, float* de);
/* implicit 3x3 convolution on the tcgen05 GEMM, channels-last activations flattened over their grid (rows = pixels, row
 * pitch Ww): out[m,n] = bias[n] + sum_t sum_c in[m + off_t, c] Wp[n, t*C + c], off_t = (t/3)*Ww + t%3, rows past the end
 * read as zeros; dgrad: din[m,c] = sum_t sum_n dout[m - off_t, n] WpT[c, t*N + n]; wgrad: dWp[n, t*C + c] += sum_m
 * dout[m,n] in[m + off_t, c].  C and N multiples of 32.  (The VGG front-end's building blocks, s2s_vgg_forward.) */
int s2s_conv3_forward(s2s_ctx* ctx, const float* in, int64_t Mg, int Ww, int C, const float* Wp, const float* bias,
                      int N, float* out, int relu);
int s2s_conv3_dgrad(s2s_ctx* ctx, const float* dout, int64_t Mg, int Ww, int N, const float* WpT, int C, float* din);
int s2s_conv3_wgrad(s2s_ctx* ctx, const float* dout, const float* in, int64_t Mg, int Ww, int N, int C, float* dWp);
/* location-aware variants (Attention.lua:75-99 with the two convolutions folded into UW [KF,S]):
 * Z[l] = q + Vh[l] + sum_j UW[j] alpha_prev[l + j - pad_left]; the backward also returns d alpha_prev [B,Lmax] */
int s2s_attn_step_forward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                              const int* lengths, int B, int Lmax, int S, int A, int KF, const float* uw,
                              const float* alpha_prev, float* alpha, float* c);
int s2s_attn_step_backward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                               const int* lengths, int B, int Lmax, int S, int A, int KF, const float* uw,
                               const float* alpha_prev, const float* alpha, const float* dc, const float* dalpha_in,
                               float* dq, float* de, float* dalpha_prev);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* S2S_B200_H */
