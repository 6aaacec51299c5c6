// model.cu -- timit/model_chorowski_baseline.lua:20-75 end to end: 3 x bidirectional nn.RNN(nn.GRU)
// encoder (JoinTable(2,2){fwd, rev}) -> nn.Attention decoder -> per-utterance NLL and its gradient
// seed (timit/timit.lua:262-282).
#include "model.cuh"

#include "gru_seq.cuh"

namespace s2s {

// nll[b] = -sum_{t<T_b} logp[b,t,y_t]  (/T_b with S2S_NORMALIZE_NLL)     timit.lua:269-272
__global__ void nll_kernel(const float* __restrict__ logp, const int* __restrict__ labels, const int* __restrict__ tlens,
                           int T, int V, int flags, float* __restrict__ nll) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int Tb = tlens ? min(tlens[b], T) : T;
    float acc = 0.f;
    for (int t = lane; t < Tb; t += 32) {
        const int y = labels[(size_t)b * T + t];
        if (y >= 0 && y < V) acc -= logp[((size_t)b * T + t) * V + y];
    }
    acc = warp_sum(acc);
    if (lane == 0) nll[b] = (flags & S2S_NORMALIZE_NLL) ? acc / (float)max(Tb, 1) : acc;
}
// dlogp = -labelmask (/T_b with S2S_NORMALIZE_GRAD), zero for t >= T_b      timit.lua:278-281
__global__ void nll_seed_kernel(const int* __restrict__ labels, const int* __restrict__ tlens, int T, int V, int flags,
                                float* __restrict__ dlogp) {
    const int bt = blockIdx.x, b = bt / T, t = bt % T;
    const int Tb = tlens ? min(tlens[b], T) : T;
    const int y = labels[bt];
    const float g = (t < Tb) ? ((flags & S2S_NORMALIZE_GRAD) ? -1.f / (float)Tb : -1.f) : 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) dlogp[(size_t)bt * V + v] = (v == y) ? g : 0.f;
}

// labelmask <-> labels (timit/timit.lua:262: labelmask = one-hot(Y) [T,V] is what the reference feeds nn.Attention and the loss).
// labels[r] = argmax_v onehot[r, v] if that row has a positive entry, else -1 ("no label": padded step / all-zero prev_y at t = 0)
__global__ void labels_from_onehot_kernel(const float* __restrict__ onehot, int64_t rows, int V, int* __restrict__ labels) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float best = 0.f; int bi = -1;
    for (int v = lane; v < V; v += 32) { const float x = onehot[r * V + v]; if (x > best) { best = x; bi = v; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
    }
    if (lane == 0) labels[r] = bi;
}
__global__ void onehot_kernel(const int* __restrict__ labels, int64_t rows, int V, float* __restrict__ onehot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * V) return;
    const int64_t r = i / V; const int v = (int)(i - r * V);
    onehot[i] = labels[r] == v ? 1.f : 0.f;
}
int labels_from_onehot(s2s_ctx* ctx, const float* onehot, int64_t rows, int V, int* labels) {
    labels_from_onehot_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, ctx->stream>>>(onehot, rows, V, labels);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
int onehot_from_labels(s2s_ctx* ctx, const int* labels, int64_t rows, int V, float* onehot) {
    onehot_kernel<<<(unsigned)ceil_div64(rows * V, 256), 256, 0, ctx->stream>>>(labels, rows, V, onehot);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// standalone loss + gradient seed for callers that compose encoder and decoder themselves (timit/timit.lua:262-282)
int nll_and_seed(s2s_ctx* ctx, const float* logp, const int* labels, const int* tlens, int B, int T, int V, int flags, float* nll, float* dlogp) {
    if (nll) {
        nll_kernel<<<B, 32, 0, ctx->stream>>>(logp, labels, tlens, T, V, flags, nll);
        S2S_LAUNCH_CHECK(ctx);
    }
    if (dlogp) {
        nll_seed_kernel<<<B * T, 64, 0, ctx->stream>>>(labels, tlens, T, V, flags, dlogp);
        S2S_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int model_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* X, const int* lengths, int B, int Lmax,
                  const int* labels, const int* tlens, int T, const float* dropmask, float lambda, int flags, float* nll, float* logp,
                  bool backward_follows) {
    S2S_REQUIRE(B > 0 && Lmax > 0 && T > 0, "model_forward: empty batch");
    if (!ctx->model) ctx->model = new ModelState();
    ModelState& m = *ctx->model;
    m.valid = false; m.B = B; m.Lmax = Lmax; m.T = T; m.Y = Y;
    const int H = Y.H, A = Y.A;
    m.acts[0] = X;
    {   // the decoder's h-independent preparation (folded weights, label inputs of all steps) under the encoder, on the side stream
        static int overlap = -1;
        if (overlap < 0) { const char* e = getenv("S2S_OVERLAP"); overlap = e ? atoi(e) : 1; }
        if (overlap && ctx->side[1] && ctx->stream != ctx->side[1] && !ctx->wgrad_join_pending) {
            cudaStream_t main_stream = ctx->stream;
            S2S_CUDA(cudaEventRecord(ctx->ev[2], main_stream));
            S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev[2], 0));
            ctx->stream = ctx->side[1];
            const int rc = decoder_prepare(ctx, Y, P, labels, B, T, backward_follows);
            ctx->stream = main_stream;
            S2S_TRY(rc);
            S2S_CUDA(cudaEventRecord(ctx->ev[4], ctx->side[1]));
            ctx->dec->prep_pending = true;
        }
    }
    for (int l = 0; l < Y.NL; l++) {
        const int din = l == 0 ? Y.D : A;
        float* out;
        S2S_ALLOC(out, ctx->persist, float, (size_t)B * Lmax * A);
        S2S_ALLOC(m.saves[l], ctx->persist, float, (size_t)B * Lmax * 2 * 4 * H);
        // both directions in one launch; outputs land in the two halves (model_chorowski_baseline.lua:22-24)
        S2S_TRY(gru_seq_forward(ctx, P + Y.enc[l][0][0].off, din, H, 2, 0, m.acts[l], din, lengths, B, Lmax, out, m.saves[l]));
        m.acts[l + 1] = out;
    }
    float* lp = logp;
    if (!lp) { S2S_ALLOC(m.logp, ctx->persist, float, (size_t)B * T * Y.V); lp = m.logp; }
    S2S_TRY(decoder_forward(ctx, Y, P, m.acts[Y.NL], lengths, B, Lmax, labels, tlens, T, dropmask, lambda, lp, backward_follows));
    if (nll) {
        nll_kernel<<<B, 32, 0, ctx->stream>>>(lp, labels, tlens, T, Y.V, flags, nll);
        S2S_LAUNCH_CHECK(ctx);
    }
    m.valid = true;
    return 0;
}

int model_backward(s2s_ctx* ctx, const Layout& Y, const float* P, float* G, const float* X, const int* lengths, int B, int Lmax,
                   const int* labels, const int* tlens, int T, const float* dropmask, float lambda, int flags, float* dX) {
    S2S_REQUIRE(ctx->model && ctx->model->valid, "model backward called without a preceding forward on this context");
    ModelState& m = *ctx->model;
    S2S_REQUIRE(m.B == B && m.Lmax == Lmax && m.T == T && m.Y.n == Y.n && m.acts[0] == X, "model backward: arguments differ from the preceding forward");
    const int H = Y.H, A = Y.A;
    float *dlogp, *dcur;
    S2S_ALLOC(dlogp, ctx->arena, float, (size_t)B * T * Y.V);
    S2S_ALLOC(dcur, ctx->arena, float, (size_t)B * Lmax * A);
    nll_seed_kernel<<<B * T, 64, 0, ctx->stream>>>(labels, tlens, T, Y.V, flags, dlogp);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(decoder_backward(ctx, Y, P, G, m.acts[Y.NL], lengths, B, Lmax, labels, tlens, T, dropmask, lambda, dlogp, dcur,
                             /*defer_wgrad=*/Y.NL > 0));
    // data-parallel overlap (s2s_dp_set_overlap): a bucket of the flat gradient is summed over the ranks as soon as it is complete, on the
    // low-priority side stream, while the remaining recurrences run: the decoder's parameters under the whole encoder backward, encoder
    // layer l under the recurrences of layers l-1 .. 0 (its weight-gradient GEMMs already run there)
    const bool dpo = ctx->dp_overlap && dp_world(ctx) > 1;
    if (dpo) S2S_TRY(dp_allreduce_bucket(ctx, G + Y.WV.off, Y.n - Y.WV.off, true));
    for (int l = Y.NL - 1; l >= 0; l--) {
        const int din = l == 0 ? Y.D : A;
        float* dprev = nullptr;
        if (l > 0) S2S_ALLOC(dprev, ctx->arena, float, (size_t)B * Lmax * A);
        else dprev = dX;
        S2S_TRY(gru_seq_backward(ctx, P + Y.enc[l][0][0].off, G + Y.enc[l][0][0].off, din, H, 2, 0, m.acts[l], din, lengths, B, Lmax,
                                 m.acts[l + 1], m.saves[l], dcur, dprev, /*defer_wgrad=*/l > 0));
        dcur = dprev;
        if (dpo) {
            const int64_t off = Y.enc[l][0][0].off, end = l + 1 < Y.NL ? Y.enc[l + 1][0][0].off : Y.WV.off;
            if (l > 0) {            // behind the layer's weight-gradient GEMMs on the side stream
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = ctx->side[1];
                const int rc = dp_allreduce_bucket(ctx, G + off, end - off, false);
                ctx->stream = main_stream;
                S2S_TRY(rc);
                S2S_CUDA(cudaEventRecord(ctx->ev[3], ctx->side[1]));        // the join event of gru_seq_wgrad_join now also covers the reduce
            } else {
                // layer 0's weight gradients ran on the main stream; its bucket still goes to the side stream so that EVERY collective of
                // the step is issued on one stream, in one order, on every rank (a communicator must not be driven from two streams that
                // are not ordered with respect to each other -- inside a captured graph nothing else would order them)
                S2S_TRY(dp_allreduce_bucket(ctx, G + off, end - off, true));
            }
        }
    }
    S2S_TRY(gru_seq_wgrad_join(ctx));
    if (dpo) S2S_TRY(dp_join(ctx));
    return 0;
}

}  // namespace s2s
