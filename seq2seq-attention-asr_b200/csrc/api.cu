// api.cu -- extern "C" entry points of libs2s_b200.so (declared in include/s2s_b200.h).
// Each wrapper validates arguments, resets the per-call scratch arena and forwards to the engine.
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include "attention.cuh"
#include "common.cuh"
#include "decoder.cuh"
#include "gru_seq.cuh"
#include "model.cuh"

#include <string.h>

using namespace s2s;

namespace s2s {
int beam_search_impl(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, int L, int eos, int beam, int maxlen,
                     int* out_host, int* n_out_host, float* logp_out_host);
int attention_step_impl(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const float* Vh, const int* lengths, int B, int Lmax,
                        const int* yprev, const float* alpha_prev, const float* s_prev, float* alpha, float* s, float* logp);
}

static void graph_drop(s2s_ctx* ctx) {
    if (ctx->graph.exec) { cudaGraphExecDestroy(ctx->graph.exec); ctx->graph.exec = nullptr; }
    ctx->graph.key.clear(); ctx->graph.seen = 0; ctx->graph.launches = 0; ctx->graph.nocapture = false;
    ctx->wgrad_join_pending = false;      // an event recorded inside a dropped / failed capture must not be waited on by an eager call
    bool user_alive = false;
    for (cudaGraphExec_t g : ctx->user_graphs) user_alive = user_alive || g != nullptr;
    ctx->arena.frozen = user_alive; ctx->persist.frozen = user_alive;      // memory referenced by a live graph must not move
}

namespace s2s {
// every CUDA graph of the context (the cached s2s_model_fwdbwd graph and caller-defined ones): a graph that holds NCCL kernel nodes must be
// gone before its communicator is destroyed (ncclCommDestroy waits for it otherwise)
void graphs_release(s2s_ctx* ctx) {
    graph_drop(ctx);
    for (cudaGraphExec_t& g : ctx->user_graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    ctx->arena.frozen = false; ctx->persist.frozen = false;
}
int labels_from_onehot(s2s_ctx* ctx, const float* onehot, int64_t rows, int V, int* labels);
int onehot_from_labels(s2s_ctx* ctx, const int* labels, int64_t rows, int V, float* onehot);
void vgg_state_free(s2s_ctx* ctx);
int nll_and_seed(s2s_ctx* ctx, const float* logp, const int* labels, const int* tlens, int B, int T, int V, int flags, float* nll, float* dlogp);
}

extern "C" {

int s2s_ctx_destroy(s2s_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->graph.exec) cudaGraphExecDestroy(ctx->graph.exec);
    for (cudaGraphExec_t g : ctx->user_graphs) if (g) cudaGraphExecDestroy(g);
    ctx->arena.release();
    ctx->persist.release();
    if (ctx->counters) cudaFree(ctx->counters);
    for (int i = 0; i < 2; i++) if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    for (int i = 0; i < 6; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    dp_state_free(ctx);
    delete ctx->dec;
    delete ctx->model;
    vgg_state_free(ctx);
    delete ctx;
    return 0;
}

// ---- TemporalConvolutionZeroBias / LinearZeroBias ------------------------------------------------
static int dense_fwd(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y) {
    S2S_REQUIRE(ctx && x && W && y, "null argument");
    S2S_REQUIRE(rows > 0 && rows < (1ll << 31) && in > 0 && out > 0, "bad shape rows=%ld in=%d out=%d", (long)rows, in, out);
    ctx->arena.reset();
    return gemm_f32(ctx, false, true, (int)rows, out, in, 1.f, x, in, W, in, 0.f, y, out);
}
static int dense_bwd(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, const float* dy, float* dx, float* dW, float scale) {
    S2S_REQUIRE(ctx && x && W && dy, "null argument");
    S2S_REQUIRE(rows > 0 && rows < (1ll << 31) && in > 0 && out > 0, "bad shape rows=%ld in=%d out=%d", (long)rows, in, out);
    ctx->arena.reset();
    if (dx) S2S_TRY(gemm_f32(ctx, false, false, (int)rows, in, out, 1.f, dy, out, W, in, 0.f, dx, in));
    if (dW) S2S_TRY(gemm_f32(ctx, true, false, out, in, (int)rows, scale, dy, out, x, in, 1.f, dW, in, nullptr, GemmBatch(), rows >= 4096 ? 8 : 1));
    return 0;
}
int s2s_tconv_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y) { return dense_fwd(ctx, x, rows, in, W, out, y); }
int s2s_tconv_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, const float* dy, float* dx, float* dW, float scale) {
    return dense_bwd(ctx, x, rows, in, W, out, dy, dx, dW, scale);
}
int s2s_linear_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y) { return dense_fwd(ctx, x, rows, in, W, out, y); }
int s2s_linear_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, const float* dy, float* dx, float* dW, float scale) {
    return dense_bwd(ctx, x, rows, in, W, out, dy, dx, dW, scale);
}

// ---- nn.RNN(nn.GRU) ---------------------------------------------------------------------------
int64_t s2s_gru_seq_save_floats(int B, int Lmax, int H, int ndir) { return (int64_t)B * Lmax * ndir * 4 * H; }
int s2s_gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                        const int* lengths, int B, int Lmax, float* y, float* save) {
    S2S_REQUIRE(ctx && W && x && y && save, "gru_seq_forward: null argument");
    S2S_REQUIRE(B > 0 && Lmax > 0 && Din > 0 && ldx >= Din, "gru_seq_forward: bad shape");
    graph_drop(ctx);
    ctx->arena.reset();
    return gru_seq_forward(ctx, W, Din, H, ndir, reverse, x, ldx, lengths, B, Lmax, y, save);
}
int s2s_gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                         const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx) {
    S2S_REQUIRE(ctx && W && dW && x && y && save && dy, "gru_seq_backward: null argument");
    S2S_REQUIRE(B > 0 && Lmax > 0 && Din > 0 && ldx >= Din, "gru_seq_backward: bad shape");
    S2S_REQUIRE(dx == nullptr || ldx == Din, "gru_seq_backward: dx requires a dense x (ldx == Din)");
    graph_drop(ctx);
    ctx->arena.reset();
    return gru_seq_backward(ctx, W, dW, Din, H, ndir, reverse, x, ldx, lengths, B, Lmax, y, save, dy, dx);
}

// ---- nn.Attention -----------------------------------------------------------------------------
int s2s_attention_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* h, const int* lengths, int B, int Lmax,
                          const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, float* logp) {
    S2S_REQUIRE(ctx && P && h && labels && logp, "attention_forward: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    graph_drop(ctx);
    ctx->arena.reset();
    ctx->persist.reset();
    if (ctx->model) ctx->model->valid = false;
    return decoder_forward(ctx, Y, P, h, lengths, B, Lmax, labels, tlens, Tmax, dropmask, lambda, logp);
}
int s2s_attention_backward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G, const float* h, const int* lengths, int B,
                           int Lmax, const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda,
                           const float* dlogp, float* dh) {
    S2S_REQUIRE(ctx && P && G && h && labels && dlogp && dh, "attention_backward: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    ctx->arena.reset();
    return decoder_backward(ctx, Y, P, G, h, lengths, B, Lmax, labels, tlens, Tmax, dropmask, lambda, dlogp, dh);
}
int s2s_attention_get(s2s_ctx* ctx, int what, float* dst) {
    S2S_REQUIRE(ctx && dst, "attention_get: null argument");
    S2S_REQUIRE(ctx->dec && ctx->dec->valid, "attention_get: no forward state on this context");
    DecoderState& d = *ctx->dec;
    const size_t BT = (size_t)d.B * d.T;
    const Layout& Y = d.Y;
    cudaStream_t st = ctx->stream;
    switch (what) {
        case S2S_GET_ALPHA: S2S_CUDA(cudaMemcpyAsync(dst, d.alpha, BT * d.Lmax * 4, cudaMemcpyDeviceToDevice, st)); break;
        case S2S_GET_WS: S2S_CUDA(cudaMemcpyAsync(dst, d.q, BT * Y.S * 4, cudaMemcpyDeviceToDevice, st)); break;
        case S2S_GET_VH: S2S_CUDA(cudaMemcpyAsync(dst, d.Vh, (size_t)d.B * d.Lmax * Y.S * 4, cudaMemcpyDeviceToDevice, st)); break;
        case S2S_GET_PENALTY: S2S_CUDA(cudaMemcpyAsync(dst, d.pen, BT * 4, cudaMemcpyDeviceToDevice, st)); break;
        case S2S_GET_STATE:
            S2S_CUDA(cudaMemcpy2DAsync(dst, (size_t)Y.ST * 4, d.sc, (size_t)(Y.ST + Y.A) * 4, (size_t)Y.ST * 4, BT, cudaMemcpyDeviceToDevice, st));
            break;
        case S2S_GET_CONTEXT:
            S2S_CUDA(cudaMemcpy2DAsync(dst, (size_t)Y.A * 4, d.sc + Y.ST, (size_t)(Y.ST + Y.A) * 4, (size_t)Y.A * 4, BT, cudaMemcpyDeviceToDevice, st));
            break;
        default: return fail("attention_get: unknown selector %d", what);
    }
    return 0;
}
int s2s_attention_step(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* h, const float* Vh, const int* lengths, int B,
                       int Lmax, const int* yprev, const float* alpha_prev, const float* s_prev, float* alpha, float* s, float* logp) {
    S2S_REQUIRE(ctx && P && h && Vh && alpha && s && logp, "attention_step: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    graph_drop(ctx);
    ctx->arena.reset();
    return attention_step_impl(ctx, Y, P, h, Vh, lengths, B, Lmax, yprev, alpha_prev, s_prev, alpha, s, logp);
}
int s2s_beam_search(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* h, int L, int eos, int beam, int maxlen,
                    int* out_host, int* n_out_host, float* logp_out_host) {
    S2S_REQUIRE(ctx && P && h && out_host && n_out_host, "beam_search: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    graph_drop(ctx);
    ctx->arena.reset();
    return beam_search_impl(ctx, Y, P, h, L, eos, beam, maxlen, out_host, n_out_host, logp_out_host);
}

// ---- whole model -------------------------------------------------------------------------------
int s2s_model_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* X, const int* lengths, int B, int Lmax,
                      const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, int flags, float* nll, float* logp) {
    S2S_REQUIRE(ctx && P && X && labels, "model_forward: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    S2S_REQUIRE(Y.NL > 0, "model_forward needs the GRU encoder (NL > 0); decoder-only layouts go through s2s_attention_forward");
    graph_drop(ctx);
    ctx->arena.reset();
    ctx->persist.reset();
    return model_forward(ctx, Y, P, X, lengths, B, Lmax, labels, tlens, Tmax, dropmask, lambda, flags, nll, logp);
}

// One forward+backward is ~800 dependent launches of a few microseconds each: replaying them as a CUDA
// graph removes the host launch cost.  Call 1 with a new signature runs eagerly (sizes the arenas, sets
// kernel attributes); call 2 is captured on an internal stream (the legacy default stream cannot be
// captured) and instantiated; later calls replay.  The arenas are bump allocators rewound per call, so
// every scratch pointer recorded in the graph is the pointer the eager path would use.
int s2s_model_fwdbwd(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G, const float* X, const int* lengths, int B, int Lmax,
                     const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, int flags, float* nll,
                     float* logp, float* dX) {
    S2S_REQUIRE(ctx && P && G && X && labels, "model_fwdbwd: null argument");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    S2S_REQUIRE(Y.NL > 0, "model_fwdbwd needs the GRU encoder (NL > 0); decoder-only layouts go through s2s_attention_forward / _backward");
    auto run = [&]() -> int {
        ctx->arena.reset();
        ctx->persist.reset();
        S2S_TRY(model_forward(ctx, Y, P, X, lengths, B, Lmax, labels, tlens, Tmax, dropmask, lambda, flags, nll, logp, /*backward_follows=*/true));
        return model_backward(ctx, Y, P, G, X, lengths, B, Lmax, labels, tlens, Tmax, dropmask, lambda, flags, dX);
    };
    if (!ctx->graphs || ctx->prof.on || ctx->capturing) {
        if (ctx->graph.exec && !ctx->capturing) graph_drop(ctx);
        return run();
    }
    uint32_t lam_bits; memcpy(&lam_bits, &lambda, 4);
    std::vector<uint64_t> key = {(uint64_t)cfg->D, (uint64_t)cfg->H, (uint64_t)cfg->NL, (uint64_t)cfg->S, (uint64_t)cfg->ST, (uint64_t)cfg->V,
                                 (uint64_t)cfg->K, (uint64_t)cfg->KF, (uint64_t)cfg->M, (uint64_t)cfg->MW + ((uint64_t)cfg->MLP << 32), (uint64_t)B, (uint64_t)Lmax, (uint64_t)Tmax,
                                 (uint64_t)lam_bits, (uint64_t)flags, (uint64_t)P, (uint64_t)G, (uint64_t)X, (uint64_t)lengths, (uint64_t)labels,
                                 (uint64_t)tlens, (uint64_t)dropmask, (uint64_t)nll, (uint64_t)logp, (uint64_t)dX, (uint64_t)ctx->stream};
    if (key != ctx->graph.key) { graph_drop(ctx); ctx->graph.key = key; }
    cudaStream_t user = ctx->stream, side = ctx->side[0];
    if (!ctx->graph.exec) {
        if (ctx->graph.nocapture || ctx->graph.seen++ == 0) return run();   // first sighting: eager warm-up
        // capture on the internal stream
        ctx->arena.frozen = true; ctx->persist.frozen = true;         // a cudaMalloc inside a capture is an error
        const int64_t l0 = ctx->launches;
        int64_t k0[S2S_KC_N];
        for (int i = 0; i < S2S_KC_N; i++) k0[i] = ctx->kcount[i];
        cudaGraph_t graph = nullptr;
        ctx->stream = side;
        cudaError_t e = cudaStreamBeginCapture(side, cudaStreamCaptureModeThreadLocal);
        int rc = 1;
        if (e == cudaSuccess) {
            rc = run();
            e = cudaStreamEndCapture(side, &graph);
        }
        ctx->stream = user;
        ctx->graph.launches = ctx->launches - l0;
        ctx->launches = l0;
        for (int i = 0; i < S2S_KC_N; i++) { ctx->graph.kc[i] = ctx->kcount[i] - k0[i]; ctx->kcount[i] = k0[i]; }
        if (rc != 0 || e != cudaSuccess || !graph) {
            std::string why = rc != 0 ? last_error() : std::string(cudaGetErrorString(e));
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (getenv("S2S_GRAPH_DBG")) fprintf(stderr, "[s2s] CUDA-graph capture of model_fwdbwd failed (%s); staying eager\n", why.c_str());
            graph_drop(ctx);
            ctx->graph.key = key; ctx->graph.nocapture = true;        // do not retry capturing this signature
            S2S_TRY(run());
            return 0;
        }
        e = cudaGraphInstantiate(&ctx->graph.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { ctx->graph.exec = nullptr; graph_drop(ctx); return fail("cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
    }
    // fork from the caller's stream, replay, join
    S2S_CUDA(cudaEventRecord(ctx->ev[0], user));
    S2S_CUDA(cudaStreamWaitEvent(side, ctx->ev[0], 0));
    S2S_CUDA(cudaGraphLaunch(ctx->graph.exec, side));
    S2S_CUDA(cudaEventRecord(ctx->ev[1], side));
    S2S_CUDA(cudaStreamWaitEvent(user, ctx->ev[1], 0));
    ctx->launches += ctx->graph.launches;
    for (int i = 0; i < S2S_KC_N; i++) ctx->kcount[i] += ctx->graph.kc[i];
    // forward state recorded by the capture pass stays valid: same shapes, same arena pointers
    return 0;
}
// ---- caller-defined CUDA graphs ----------------------------------------------------------------------------------
// Everything the library launches for this context between _begin and _end is captured into one graph (on an internal
// stream) instead of being executed; _launch replays it in the order of the context's stream.  The decoder time loop is
// hundreds of microsecond-sized launches, so a replayed step costs about half of the eager one.  Rules for the captured
// region: same pointers and shapes on every replay; run the sequence eagerly once first (workspaces must already be
// large enough -- they cannot grow while a graph references them); no calls that read results back to the host.
int s2s_graph_begin(s2s_ctx* ctx) {
    S2S_REQUIRE(ctx && !ctx->capturing, "graph_begin: bad context or already capturing");
    if (ctx->graph.exec) graph_drop(ctx);
    ctx->capture_user_stream = ctx->stream;
    ctx->capture_l0 = ctx->launches;
    ctx->capture_k0.assign(ctx->kcount, ctx->kcount + S2S_KC_N);
    ctx->arena.frozen = true; ctx->persist.frozen = true;
    cudaError_t e = cudaStreamBeginCapture(ctx->side[0], cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { graph_drop(ctx); return fail("graph_begin: cudaStreamBeginCapture: %s", cudaGetErrorString(e)); }
    ctx->stream = ctx->side[0];
    ctx->capturing = true;
    return 0;
}
int s2s_graph_end(s2s_ctx* ctx, int* graph_id) {
    S2S_REQUIRE(ctx && ctx->capturing && graph_id, "graph_end: not capturing");
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->side[0], &graph);
    ctx->stream = ctx->capture_user_stream;
    ctx->capturing = false;
    const int64_t n = ctx->launches - ctx->capture_l0;
    ctx->launches = ctx->capture_l0;
    std::vector<int64_t> kc(S2S_KC_N, 0);
    for (int i = 0; i < S2S_KC_N && i < (int)ctx->capture_k0.size(); i++) { kc[i] = ctx->kcount[i] - ctx->capture_k0[i]; ctx->kcount[i] = ctx->capture_k0[i]; }
    cudaGraphExec_t exec = nullptr;
    if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || !exec) { cudaGetLastError(); graph_drop(ctx); return fail("graph_end: capture failed: %s", cudaGetErrorString(e)); }
    ctx->user_graphs.push_back(exec);
    ctx->user_graph_launches.push_back(n);
    ctx->user_graph_kc.push_back(kc);
    *graph_id = (int)ctx->user_graphs.size() - 1;
    return 0;
}
int s2s_graph_launch(s2s_ctx* ctx, int graph_id) {
    S2S_REQUIRE(ctx && !ctx->capturing && graph_id >= 0 && graph_id < (int)ctx->user_graphs.size() && ctx->user_graphs[graph_id],
                "graph_launch: invalid graph id %d", graph_id);
    cudaStream_t user = ctx->stream, side = ctx->side[0];
    S2S_CUDA(cudaEventRecord(ctx->ev[0], user));
    S2S_CUDA(cudaStreamWaitEvent(side, ctx->ev[0], 0));
    S2S_CUDA(cudaGraphLaunch(ctx->user_graphs[graph_id], side));
    S2S_CUDA(cudaEventRecord(ctx->ev[1], side));
    S2S_CUDA(cudaStreamWaitEvent(user, ctx->ev[1], 0));
    ctx->launches += ctx->user_graph_launches[graph_id];
    for (int i = 0; i < S2S_KC_N; i++) ctx->kcount[i] += ctx->user_graph_kc[graph_id][i];
    return 0;
}
int s2s_graph_destroy(s2s_ctx* ctx, int graph_id) {
    S2S_REQUIRE(ctx && graph_id >= 0 && graph_id < (int)ctx->user_graphs.size(), "graph_destroy: invalid graph id %d", graph_id);
    if (ctx->user_graphs[graph_id]) {
        cudaStreamSynchronize(ctx->side[0]);
        cudaGraphExecDestroy(ctx->user_graphs[graph_id]);
        ctx->user_graphs[graph_id] = nullptr;
    }
    if (!ctx->graph.exec) graph_drop(ctx);      // unfreezes the workspaces when no graph is left
    return 0;
}

int s2s_labels_from_onehot(s2s_ctx* ctx, const float* onehot, int64_t rows, int V, int* labels) {
    S2S_REQUIRE(ctx && onehot && labels && rows > 0 && V > 0, "labels_from_onehot: bad arguments");
    return labels_from_onehot(ctx, onehot, rows, V, labels);
}
int s2s_onehot(s2s_ctx* ctx, const int* labels, int64_t rows, int V, float* onehot) {
    S2S_REQUIRE(ctx && onehot && labels && rows > 0 && V > 0, "onehot: bad arguments");
    return onehot_from_labels(ctx, labels, rows, V, onehot);
}

// nll[b] = -sum_t logp[b,t,y_t] (/T_b) and dlogp = -labelmask (/T_b), zero beyond T_b   (timit/timit.lua:262-282)
int s2s_nll_grad_seed(s2s_ctx* ctx, const float* logp, const int* labels, const int* tlens, int B, int T, int V, int flags, float* nll,
                      float* dlogp) {
    S2S_REQUIRE(ctx && logp && labels && B > 0 && T > 0 && V > 1 && (nll || dlogp), "nll_grad_seed: bad arguments");
    return nll_and_seed(ctx, logp, labels, tlens, B, T, V, flags, nll, dlogp);
}
int s2s_model_get_annotations(s2s_ctx* ctx, float* dst) {
    S2S_REQUIRE(ctx && dst, "model_get_annotations: null argument");
    S2S_REQUIRE(ctx->model && ctx->model->valid, "model_get_annotations: no forward state on this context");
    ModelState& m = *ctx->model;
    S2S_CUDA(cudaMemcpyAsync(dst, m.acts[m.Y.NL], (size_t)m.B * m.Lmax * m.Y.A * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

// ---- test hooks --------------------------------------------------------------------------------
int s2s_gemm_f32(s2s_ctx* ctx, int impl, int tA, int tB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                 int ldb, float beta, float* C, int ldc, const float* bias) {
    S2S_REQUIRE(ctx && A && B && C, "gemm: null argument");
    ctx->arena.reset();
    return gemm_f32(ctx, tA != 0, tB != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, GemmBatch(), 1, impl);
}
int s2s_attn_step_forward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w, const int* lengths, int B, int Lmax,
                          int S, int A, float* alpha, float* c) {
    S2S_REQUIRE(ctx && Vh && h && q && w && alpha && c, "attn_step_forward: null argument");
    ctx->arena.reset();
    AttnScratch sc;
    S2S_TRY(attn_scratch_alloc(ctx, ctx->arena, B, Lmax, S, A, 0, false, &sc));
    AttnLoc loc;
    return attn_step_fwd(ctx, sc, Vh, h, q, S, w, lengths, B, Lmax, S, A, loc, alpha, Lmax, c, A, nullptr, 0, 0.f, nullptr, 0);
}
int s2s_attn_step_backward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w, const int* lengths, int B,
                           int Lmax, int S, int A, const float* alpha, const float* dc, const float* dalpha_in, float* dq, float* de) {
    S2S_REQUIRE(ctx && Vh && h && q && w && alpha && dc && dq && de, "attn_step_backward: null argument");
    ctx->arena.reset();
    AttnScratch sc;
    S2S_TRY(attn_scratch_alloc(ctx, ctx->arena, B, Lmax, S, A, 0, true, &sc));
    AttnLoc loc;
    return attn_step_bwd(ctx, sc, Vh, h, q, S, w, lengths, B, Lmax, S, A, loc, alpha, Lmax, dc, A, dalpha_in, Lmax, nullptr, 0, 0.f,
                         dq, S, de, Lmax, nullptr, 0);
}

// WagnerFischer(a, b) of utils.lua:3-27: Levenshtein distance between two label sequences (PER / CER scoring of decodes).
// Host integers in, host integer out: the decodes this scores are already on the host (s2s_beam_search).
int s2s_edit_distance(const int* a, int na, const int* b, int nb, int* dist_host) {
    S2S_REQUIRE(dist_host && na >= 0 && nb >= 0 && (a || na == 0) && (b || nb == 0), "edit_distance: bad arguments");
    std::vector<int> prev((size_t)na + 1), cur((size_t)na + 1);
    for (int i = 0; i <= na; i++) prev[i] = i;
    for (int j = 1; j <= nb; j++) {
        cur[0] = j;
        for (int i = 1; i <= na; i++) {
            if (a[i - 1] == b[j - 1]) cur[i] = prev[i - 1];
            else cur[i] = 1 + std::min(prev[i - 1], std::min(prev[i], cur[i - 1]));
        }
        std::swap(prev, cur);
    }
    *dist_host = prev[na];
    return 0;
}

// location-aware variants of the two hooks (cfg5 sweep with K = 16 feature maps folded into UW [KF, S]):
// Z[l] = q + Vh[l] + sum_j UW[j] * alpha_prev[l + j - pad_left]   (Attention.lua:75-99)
int s2s_attn_step_forward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w, const int* lengths, int B, int Lmax,
                              int S, int A, int KF, const float* uw, const float* alpha_prev, float* alpha, float* c) {
    S2S_REQUIRE(ctx && Vh && h && q && w && alpha && c && uw, "attn_step_forward_loc: null argument");
    S2S_REQUIRE(KF >= 1 && KF <= 16, "attn_step_forward_loc: filter size %d out of range (1..16)", KF);
    ctx->arena.reset();
    AttnScratch sc;
    S2S_TRY(attn_scratch_alloc(ctx, ctx->arena, B, Lmax, S, A, KF, false, &sc));
    AttnLoc loc; loc.KF = KF; loc.padl = (KF % 2 == 1) ? (KF - 1) / 2 : KF / 2; loc.uw = uw; loc.alpha_prev = alpha_prev; loc.ld_aprev = Lmax;
    return attn_step_fwd(ctx, sc, Vh, h, q, S, w, lengths, B, Lmax, S, A, loc, alpha, Lmax, c, A, nullptr, 0, 0.f, nullptr, 0);
}
int s2s_attn_step_backward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w, const int* lengths, int B,
                               int Lmax, int S, int A, int KF, const float* uw, const float* alpha_prev, const float* alpha, const float* dc,
                               const float* dalpha_in, float* dq, float* de, float* dalpha_prev) {
    S2S_REQUIRE(ctx && Vh && h && q && w && alpha && dc && dq && de && uw && dalpha_prev, "attn_step_backward_loc: null argument");
    S2S_REQUIRE(KF >= 1 && KF <= 16, "attn_step_backward_loc: filter size %d out of range (1..16)", KF);
    ctx->arena.reset();
    AttnScratch sc;
    S2S_TRY(attn_scratch_alloc(ctx, ctx->arena, B, Lmax, S, A, KF, true, &sc));
    AttnLoc loc; loc.KF = KF; loc.padl = (KF % 2 == 1) ? (KF - 1) / 2 : KF / 2; loc.uw = uw; loc.alpha_prev = alpha_prev; loc.ld_aprev = Lmax;
    return attn_step_bwd(ctx, sc, Vh, h, q, S, w, lengths, B, Lmax, S, A, loc, alpha, Lmax, dc, A, dalpha_in, Lmax, nullptr, 0, 0.f,
                         dq, S, de, Lmax, dalpha_prev, Lmax);
}

}  // extern "C"
