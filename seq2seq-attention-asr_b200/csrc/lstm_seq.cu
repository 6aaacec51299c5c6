// lstm_seq.cu -- nn.RNN(nn.LSTM(diminput, dimoutput, peepholes), reverse) over whole utterances
// (LSTM.lua:6-136, RNN.lua:120-201).  Without peepholes and for H in {128, 256} the recurrence runs in the persistent
// cluster kernels of lstm_cluster.cu; the general path below (any H <= 256, full-matrix peepholes) launches per frame.
//
// Reference step (LSTM.lua:25-58): every gate is Linear(in->out)(x) + Linear(out->out)(h_prev) [+ Linear(out->out)(c)],
// each Linear WITH bias; peepholes are FULL matrices on prev_c (input, forget gates) and next_c (output gate):
//     i = sig(.), f = sig(.), g = tanh(.), c' = f c + i g, o = sig(. [+ W_co c']), h' = o tanh(c')
// Parameter block (flat, the order the module's parameters() yields): for gate in (i, f, g, o):
//     Wx[out,in], bx[out], Wh[out,out], bh[out], [Wc[out,out], bc[out]  if peepholes and gate != g]
//
// Design: the x-products of all four gates and all three biases are ONE time-batched projection per gate
// (tcgen05 GEMM when large); the recurrent and peephole matrices are packed once per call into contiguous
// [4H,H] / [3H,H] blocks so each step is a small-batch product (dense_small) plus a fused elementwise kernel;
// backward emits the gate gradients dA[B,L,4H] and every weight gradient is a time-batched GEMM (K = B*L).
#include "common.cuh"

namespace s2s {

int dense_small_linear(s2s_ctx* ctx, const float* X, int64_t ldx, int B, int K, const float* W, int ldw, int N, const float* bias,
                       const float* add, int64_t ld_add, float* out, int64_t ld_out);
int zero_tail_rows(s2s_ctx* ctx, float* x, const int* lengths, int B, int Lmax, int W);
int lstm_cluster_forward(s2s_ctx* ctx, const float* Whp, const float* xp, const int* lengths, int B, int Lmax, int H, int reverse,
                         float* y, float* cseq, float* acts, bool* handled);
int lstm_cluster_backward(s2s_ctx* ctx, const float* Whp, const int* lengths, int B, int Lmax, int H, int reverse, const float* y,
                          const float* cseq, const float* acts, const float* dy, float* dA, float* hprev, float* cprev, bool* handled);

struct LstmLayout {
    int in, H, peep;
    int64_t Wx[4], bx[4], Wh[4], bh[4], Wc[4], bc[4];   // offsets (Wc/bc = -1 when absent)
    int64_t n;
};
static LstmLayout lstm_layout(int in, int H, int peep) {
    LstmLayout y; y.in = in; y.H = H; y.peep = peep;
    int64_t o = 0;
    for (int g = 0; g < 4; g++) {
        y.Wx[g] = o; o += (int64_t)H * in; y.bx[g] = o; o += H;
        y.Wh[g] = o; o += (int64_t)H * H; y.bh[g] = o; o += H;
        if (peep && g != 2) { y.Wc[g] = o; o += (int64_t)H * H; y.bc[g] = o; o += H; } else { y.Wc[g] = -1; y.bc[g] = -1; }
    }
    y.n = o;
    return y;
}

// pack recurrent [4H,H] and peephole [3H,H] (order i, f, o) blocks and the summed biases [4H]
__global__ void lstm_pack_kernel(const float* __restrict__ P, LstmLayout y, float* __restrict__ Whp, float* __restrict__ Wcp, float* __restrict__ bias) {
    const int H = y.H;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)4 * H * H) { const int g = (int)(i / ((int64_t)H * H)); Whp[i] = P[y.Wh[g] + i - (int64_t)g * H * H]; }
    if (y.peep && i < (int64_t)3 * H * H) {
        const int q = (int)(i / ((int64_t)H * H)); const int g = q == 2 ? 3 : q;
        Wcp[i] = P[y.Wc[g] + i - (int64_t)q * H * H];
    }
    if (i < 4 * H) { const int g = (int)(i / H), j = (int)(i % H); bias[i] = P[y.bx[g] + j] + P[y.bh[g] + j] + (y.Wc[g] >= 0 ? P[y.bc[g] + j] : 0.f); }
}

struct LstmStep {
    const int* lengths; int B, Lmax, H, reverse, s, peep;
    const float* xp;      // [B, Lmax, 4H] input projections + biases
    const float* pre;     // [B, 4H] recurrent (+ input/forget peephole) pre-activations of this step
    const float* pre_o2;  // [B, H]  output-gate peephole product (peep)
    float *hstate, *cstate;       // [B, H]
    float *y, *cseq, *acts;       // [B,Lmax,H], [B,Lmax,H], [B,Lmax,4H]
};
__device__ __forceinline__ bool lstm_time(const LstmStep& p, int b, int* t) {
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    if (p.s >= Lb) return false;
    *t = p.reverse ? Lb - 1 - p.s : p.s;
    return true;
}
// phase 1: i, f, g, c' (and, without peepholes, o and h')
__global__ void lstm_fwd_k1(const LstmStep p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.B * p.H) return;
    const int b = idx / p.H, j = idx - b * p.H, H = p.H;
    int t;
    if (!lstm_time(p, b, &t)) return;
    const float* xr = p.xp + ((size_t)b * p.Lmax + t) * 4 * H;
    const float* pr = p.pre + (size_t)b * 4 * H;
    const float ig = sigmoid_acc(pr[j] + xr[j]), fg = sigmoid_acc(pr[H + j] + xr[H + j]), gg = tanh_acc(pr[2 * H + j] + xr[2 * H + j]);
    const float c = fg * p.cstate[idx] + ig * gg;                                 // LSTM.lua:45-46
    float* ar = p.acts + ((size_t)b * p.Lmax + t) * 4 * H;
    ar[j] = ig; ar[H + j] = fg; ar[2 * H + j] = gg;
    p.cstate[idx] = c;
    p.cseq[((size_t)b * p.Lmax + t) * H + j] = c;
    if (!p.peep) {
        const float og = sigmoid_acc(pr[3 * H + j] + xr[3 * H + j]);
        const float h = og * tanh_acc(c);                                         // LSTM.lua:51
        ar[3 * H + j] = og;
        p.hstate[idx] = h;
        p.y[((size_t)b * p.Lmax + t) * H + j] = h;
    }
}
// phase 2 (peepholes): o = sig(. + W_co c'), h'
__global__ void lstm_fwd_k2(const LstmStep p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.B * p.H) return;
    const int b = idx / p.H, j = idx - b * p.H, H = p.H;
    int t;
    if (!lstm_time(p, b, &t)) return;
    const float og = sigmoid_acc(p.pre[(size_t)b * 4 * H + 3 * H + j] + p.xp[((size_t)b * p.Lmax + t) * 4 * H + 3 * H + j] + p.pre_o2[idx]);
    const float h = og * tanh_acc(p.cstate[idx]);
    p.acts[((size_t)b * p.Lmax + t) * 4 * H + 3 * H + j] = og;
    p.hstate[idx] = h;
    p.y[((size_t)b * p.Lmax + t) * H + j] = h;
}

struct LstmBStep {
    const int* lengths; int B, Lmax, H, reverse, s, peep;
    const float *y, *cseq, *acts, *dy;
    float *dh, *dc;             // [B,H] carries (gradient w.r.t. h_t / c_t from the later step)
    const float* dc_add;        // [B,H] W_co^T da_o (peep)
    float *dA, *hprev, *cprev;  // [B,Lmax,4H], [B,Lmax,H], [B,Lmax,H]
};
__device__ __forceinline__ bool lstm_btime(const LstmBStep& p, int b, int* t, int* tp) {
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    if (p.s >= Lb) return false;
    *t = p.reverse ? Lb - 1 - p.s : p.s;
    *tp = p.s == 0 ? -1 : (p.reverse ? *t + 1 : *t - 1);
    return true;
}
// output gate: da_o = dh tanh(c') o (1-o)            (LSTM.lua:118-136 through the nngraph)
__global__ void lstm_bwd_ko(const LstmBStep p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.B * p.H) return;
    const int b = idx / p.H, j = idx - b * p.H, H = p.H;
    int t, tp;
    float* dar = p.dA + ((size_t)b * p.Lmax) * 4 * H;
    if (!lstm_btime(p, b, &t, &tp)) return;
    const size_t row = (size_t)b * p.Lmax + t;
    const float dh = p.dy[row * H + j] + p.dh[idx];
    const float og = p.acts[row * 4 * H + 3 * H + j];
    const float tc = tanh_acc(p.cseq[row * H + j]);
    dar[(size_t)t * 4 * H + 3 * H + j] = dh * tc * og * (1.f - og);
}
// cell + input/forget/candidate gates
__global__ void lstm_bwd_kc(const LstmBStep p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.B * p.H) return;
    const int b = idx / p.H, j = idx - b * p.H, H = p.H;
    int t, tp;
    if (!lstm_btime(p, b, &t, &tp)) return;
    const size_t row = (size_t)b * p.Lmax + t;
    const float* ar = p.acts + row * 4 * H;
    const float ig = ar[j], fg = ar[H + j], gg = ar[2 * H + j], og = ar[3 * H + j];
    const float dh = p.dy[row * H + j] + p.dh[idx];
    const float tc = tanh_acc(p.cseq[row * H + j]);
    float dc = p.dc[idx] + dh * og * (1.f - tc * tc);
    if (p.peep) dc += p.dc_add[idx];
    const float cp = tp >= 0 ? p.cseq[((size_t)b * p.Lmax + tp) * H + j] : 0.f;
    const float hp = tp >= 0 ? p.y[((size_t)b * p.Lmax + tp) * H + j] : 0.f;
    float* dar = p.dA + row * 4 * H;
    dar[j] = dc * gg * ig * (1.f - ig);
    dar[H + j] = dc * cp * fg * (1.f - fg);
    dar[2 * H + j] = dc * ig * (1.f - gg * gg);
    p.dc[idx] = dc * fg;                       // + W_c{i,f}^T {da_i, da_f} added afterwards (peep)
    p.hprev[row * H + j] = hp;
    p.cprev[row * H + j] = cp;
}
// gather this step's gate gradients of every utterance into a dense [B, 4H] block (rows of inactive utterances = 0)
__global__ void lstm_gather_dA(const LstmBStep p, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.B * 4 * p.H) return;
    const int b = idx / (4 * p.H), j = idx - b * 4 * p.H;
    int t, tp;
    out[idx] = lstm_btime(p, b, &t, &tp) ? p.dA[((size_t)b * p.Lmax + t) * 4 * p.H + j] : 0.f;
}

int lstm_seq_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peep, int reverse, const float* x, int ldx, const int* lengths,
                     int B, int Lmax, float* y, float* save) {
    S2S_REQUIRE(H % 4 == 0 && H <= 256, "lstm_seq: hidden size must be a multiple of 4 and <= 256 (got %d)", H);
    const LstmLayout L = lstm_layout(Din, H, peep);
    Arena& ar = ctx->arena;
    cudaStream_t st = ctx->stream;
    float *Whp, *Wcp = nullptr, *bias, *xp, *pre, *pre_o2 = nullptr, *hstate, *cstate;
    S2S_ALLOC(Whp, ar, float, (size_t)4 * H * H);
    if (peep) S2S_ALLOC(Wcp, ar, float, (size_t)3 * H * H);
    S2S_ALLOC(bias, ar, float, 4 * H);
    S2S_ALLOC(xp, ar, float, (size_t)B * Lmax * 4 * H);
    S2S_ALLOC(pre, ar, float, (size_t)B * 4 * H);
    if (peep) S2S_ALLOC(pre_o2, ar, float, (size_t)B * H);
    S2S_ALLOC(hstate, ar, float, (size_t)B * H);
    S2S_ALLOC(cstate, ar, float, (size_t)B * H);
    float* cseq = save;                                   // [B,Lmax,H]
    float* acts = save + (size_t)B * Lmax * H;            // [B,Lmax,4H]
    lstm_pack_kernel<<<(unsigned)ceil_div64((int64_t)4 * H * H, 256), 256, 0, st>>>(P, L, Whp, Wcp, bias);
    S2S_LAUNCH_CHECK(ctx);
    for (int g = 0; g < 4; g++)
        S2S_TRY(gemm_f32(ctx, false, true, B * Lmax, H, Din, 1.f, x, ldx, P + L.Wx[g], Din, 0.f, xp + g * H, 4 * H, bias + g * H));
    S2S_CUDA(cudaMemsetAsync(hstate, 0, (size_t)B * H * 4, st));      // zeros initial state (LSTM.lua:108-109)
    S2S_CUDA(cudaMemsetAsync(cstate, 0, (size_t)B * H * 4, st));
    if (lengths) {
        S2S_TRY(zero_tail_rows(ctx, y, lengths, B, Lmax, H));
        S2S_TRY(zero_tail_rows(ctx, cseq, lengths, B, Lmax, H));
        S2S_TRY(zero_tail_rows(ctx, acts, lengths, B, Lmax, 4 * H));
    }
    if (!peep) {   // persistent cluster recurrence (lstm_cluster.cu)
        bool handled = false;
        S2S_TRY(lstm_cluster_forward(ctx, Whp, xp, lengths, B, Lmax, H, reverse, y, cseq, acts, &handled));
        if (handled) return 0;
    }
    const int eb = ceil_div(B * H, 256);
    for (int s = 0; s < Lmax; s++) {
        S2S_TRY(dense_small_linear(ctx, hstate, H, B, H, Whp, H, 4 * H, nullptr, nullptr, 0, pre, 4 * H));
        if (peep) S2S_TRY(dense_small_linear(ctx, cstate, H, B, H, Wcp, H, 2 * H, nullptr, pre, 4 * H, pre, 4 * H));   // input, forget peepholes on prev_c
        LstmStep p = {lengths, B, Lmax, H, reverse, s, peep, xp, pre, pre_o2, hstate, cstate, y, cseq, acts};
        lstm_fwd_k1<<<eb, 256, 0, st>>>(p);
        S2S_LAUNCH_CHECK(ctx);
        if (peep) {
            S2S_TRY(dense_small_linear(ctx, cstate, H, B, H, Wcp + (size_t)2 * H * H, H, H, nullptr, nullptr, 0, pre_o2, H));   // output peephole on next_c
            lstm_fwd_k2<<<eb, 256, 0, st>>>(p);
            S2S_LAUNCH_CHECK(ctx);
        }
    }
    return 0;
}

int lstm_seq_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peep, int reverse, const float* x, int ldx,
                      const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx) {
    S2S_REQUIRE(H % 4 == 0 && H <= 256, "lstm_seq: hidden size must be a multiple of 4 and <= 256 (got %d)", H);
    const LstmLayout L = lstm_layout(Din, H, peep);
    Arena& ar = ctx->arena;
    cudaStream_t st = ctx->stream;
    const int BL = B * Lmax;
    float *Whp, *Wcp = nullptr, *bias, *WhpT, *WcoT = nullptr, *WcifT = nullptr, *dA, *hprev, *cprev, *dh, *dc, *dc_add = nullptr, *dAs;
    S2S_ALLOC(Whp, ar, float, (size_t)4 * H * H);
    if (peep) S2S_ALLOC(Wcp, ar, float, (size_t)3 * H * H);
    S2S_ALLOC(bias, ar, float, 4 * H);
    S2S_ALLOC(WhpT, ar, float, (size_t)4 * H * H);
    if (peep) { S2S_ALLOC(WcoT, ar, float, (size_t)H * H); S2S_ALLOC(WcifT, ar, float, (size_t)2 * H * H); S2S_ALLOC(dc_add, ar, float, (size_t)B * H); }
    S2S_ALLOC(dA, ar, float, (size_t)BL * 4 * H);
    S2S_ALLOC(hprev, ar, float, (size_t)BL * H);
    S2S_ALLOC(cprev, ar, float, (size_t)BL * H);
    S2S_ALLOC(dh, ar, float, (size_t)B * H);
    S2S_ALLOC(dc, ar, float, (size_t)B * H);
    S2S_ALLOC(dAs, ar, float, (size_t)B * 4 * H);
    const float* cseq = save;
    const float* acts = save + (size_t)BL * H;
    lstm_pack_kernel<<<(unsigned)ceil_div64((int64_t)4 * H * H, 256), 256, 0, st>>>(P, L, Whp, Wcp, bias);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(transpose_f32(ctx, Whp, 4 * H, H, H, WhpT, 4 * H));                     // [H, 4H]
    if (peep) {
        S2S_TRY(transpose_f32(ctx, Wcp + (size_t)2 * H * H, H, H, H, WcoT, H));     // [H, H]
        S2S_TRY(transpose_f32(ctx, Wcp, 2 * H, H, H, WcifT, 2 * H));                // [H, 2H]
    }
    S2S_CUDA(cudaMemsetAsync(dA, 0, (size_t)BL * 4 * H * 4, st));
    S2S_CUDA(cudaMemsetAsync(hprev, 0, (size_t)BL * H * 4, st));
    S2S_CUDA(cudaMemsetAsync(cprev, 0, (size_t)BL * H * 4, st));
    S2S_CUDA(cudaMemsetAsync(dh, 0, (size_t)B * H * 4, st));
    S2S_CUDA(cudaMemsetAsync(dc, 0, (size_t)B * H * 4, st));
    bool cluster_done = false;
    if (!peep) S2S_TRY(lstm_cluster_backward(ctx, Whp, lengths, B, Lmax, H, reverse, y, cseq, acts, dy, dA, hprev, cprev, &cluster_done));
    const int eb = ceil_div(B * H, 256);
    for (int s = Lmax - 1; s >= 0 && !cluster_done; s--) {                          // RNN.lua:183
        LstmBStep p = {lengths, B, Lmax, H, reverse, s, peep, y, cseq, acts, dy, dh, dc, dc_add, dA, hprev, cprev};
        if (peep) {
            lstm_bwd_ko<<<eb, 256, 0, st>>>(p);
            S2S_LAUNCH_CHECK(ctx);
            lstm_gather_dA<<<ceil_div(B * 4 * H, 256), 256, 0, st>>>(p, dAs);
            S2S_LAUNCH_CHECK(ctx);
            S2S_TRY(dense_small_linear(ctx, dAs + 3 * H, 4 * H, B, H, WcoT, H, H, nullptr, nullptr, 0, dc_add, H));     // d c' += da_o . W_co
        } else {
            lstm_bwd_ko<<<eb, 256, 0, st>>>(p);
            S2S_LAUNCH_CHECK(ctx);
        }
        lstm_bwd_kc<<<eb, 256, 0, st>>>(p);
        S2S_LAUNCH_CHECK(ctx);
        lstm_gather_dA<<<ceil_div(B * 4 * H, 256), 256, 0, st>>>(p, dAs);
        S2S_LAUNCH_CHECK(ctx);
        S2S_TRY(dense_small_linear(ctx, dAs, 4 * H, B, 4 * H, WhpT, 4 * H, H, nullptr, nullptr, 0, dh, H));             // dh_{t-1} = dA . W_h
        if (peep) S2S_TRY(dense_small_linear(ctx, dAs, 4 * H, B, 2 * H, WcifT, 2 * H, H, nullptr, dc, H, dc, H));       // dc_{t-1} += {da_i, da_f} . W_c{i,f}
    }
    // time-batched parameter gradients and dX
    const int sk = BL >= 4096 ? 8 : 1;
    for (int g = 0; g < 4; g++) {
        const float* dAg = dA + g * H;
        S2S_TRY(gemm_f32(ctx, true, false, H, Din, BL, 1.f, dAg, 4 * H, x, ldx, 1.f, dP + L.Wx[g], Din, nullptr, GemmBatch(), sk));
        S2S_TRY(gemm_f32(ctx, true, false, H, H, BL, 1.f, dAg, 4 * H, hprev, H, 1.f, dP + L.Wh[g], H, nullptr, GemmBatch(), sk));
        S2S_TRY(colsum_add(ctx, dAg, BL, H, 4 * H, dP + L.bx[g]));
        S2S_TRY(colsum_add(ctx, dAg, BL, H, 4 * H, dP + L.bh[g]));
        if (L.Wc[g] >= 0) {
            S2S_TRY(gemm_f32(ctx, true, false, H, H, BL, 1.f, dAg, 4 * H, g == 3 ? cseq : cprev, H, 1.f, dP + L.Wc[g], H, nullptr, GemmBatch(), sk));
            S2S_TRY(colsum_add(ctx, dAg, BL, H, 4 * H, dP + L.bc[g]));
        }
        if (dx) S2S_TRY(gemm_f32(ctx, false, false, BL, Din, H, 1.f, dAg, 4 * H, P + L.Wx[g], Din, g == 0 ? 0.f : 1.f, dx, Din));
    }
    return 0;
}

}  // namespace s2s

using namespace s2s;
extern "C" {
int64_t s2s_lstm_param_count(int in, int out, int peepholes) { return lstm_layout(in, out, peepholes).n; }
int64_t s2s_lstm_seq_save_floats(int B, int Lmax, int H) { return (int64_t)B * Lmax * 5 * H; }
int s2s_lstm_seq_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, int reverse, const float* x, int ldx, const int* lengths,
                         int B, int Lmax, float* y, float* save) {
    S2S_REQUIRE(ctx && P && x && y && save, "lstm_seq_forward: null argument");
    S2S_REQUIRE(B > 0 && Lmax > 0 && Din > 0 && ldx >= Din, "lstm_seq_forward: bad shape");
    ctx->arena.reset();
    return lstm_seq_forward(ctx, P, Din, H, peepholes, reverse, x, ldx, lengths, B, Lmax, y, save);
}
int s2s_lstm_seq_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, int reverse, const float* x, int ldx,
                          const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx) {
    S2S_REQUIRE(ctx && P && dP && x && y && save && dy, "lstm_seq_backward: null argument");
    S2S_REQUIRE(B > 0 && Lmax > 0 && Din > 0 && ldx >= Din, "lstm_seq_backward: bad shape");
    S2S_REQUIRE(dx == nullptr || ldx == Din, "lstm_seq_backward: dx requires a dense x (ldx == Din)");
    ctx->arena.reset();
    return lstm_seq_backward(ctx, P, dP, Din, H, peepholes, reverse, x, ldx, lengths, B, Lmax, y, save, dy, dx);
}
}
