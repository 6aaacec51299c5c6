// lstm_step.cu -- nn.LSTM as a single-step module with explicit previous state: updateOutput({x, prev_h, prev_c}) ->
// {next_h, next_c} and updateGradInput (LSTM.lua:6-136; two biases per gate, full-matrix peepholes on prev_c for the
// input/forget gates and on next_c for the output gate).  The sequence kernels (lstm_cluster.cu / lstm_seq.cu) are the hot
// path; this entry completes the module surface (user code stepping an LSTM by hand).
// Parameter block: the flat order of lstm_seq.cu (for gate in i, f, g, o: Wx, bx, Wh, bh [, Wc, bc]).
#include "common.cuh"

namespace s2s {

struct LstmStepLayout { int64_t Wx[4], bx[4], Wh[4], bh[4], Wc[4], bc[4]; };
static LstmStepLayout lstm_step_layout(int in, int H, int peep) {
    LstmStepLayout y;
    int64_t o = 0;
    for (int g = 0; g < 4; g++) {
        y.Wx[g] = o; o += (int64_t)H * in; y.bx[g] = o; o += H;
        y.Wh[g] = o; o += (int64_t)H * H; y.bh[g] = o; o += H;
        if (peep && g != 2) { y.Wc[g] = o; o += (int64_t)H * H; y.bc[g] = o; o += H; } else { y.Wc[g] = -1; y.bc[g] = -1; }
    }
    return y;
}

// pre [B,4H] holds the summed Linear outputs; phase 1: i, f, g, c' (and o, h' without peepholes)
__global__ void lstm_step_f1_kernel(const float* __restrict__ pre, const float* __restrict__ cprev, int B, int H, int peep,
                                    float* __restrict__ acts, float* __restrict__ cnext, float* __restrict__ hnext) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float* pr = pre + (size_t)b * 4 * H;
    const float ig = sigmoid_acc(pr[j]), fg = sigmoid_acc(pr[H + j]), gg = tanh_acc(pr[2 * H + j]);
    const float c = fg * (cprev ? cprev[idx] : 0.f) + ig * gg;                         // LSTM.lua:45-46
    float* ar = acts + (size_t)b * 4 * H;
    ar[j] = ig; ar[H + j] = fg; ar[2 * H + j] = gg;
    cnext[idx] = c;
    if (!peep) {
        const float og = sigmoid_acc(pr[3 * H + j]);
        ar[3 * H + j] = og;
        hnext[idx] = og * tanh_acc(c);                                                // LSTM.lua:51
    }
}
__global__ void lstm_step_f2_kernel(const float* __restrict__ pre, const float* __restrict__ cnext, int B, int H, float* __restrict__ acts,
                                    float* __restrict__ hnext) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float og = sigmoid_acc(pre[(size_t)b * 4 * H + 3 * H + j]);
    acts[(size_t)b * 4 * H + 3 * H + j] = og;
    hnext[idx] = og * tanh_acc(cnext[idx]);
}
// da_o = dh tanh(c') o (1-o)
__global__ void lstm_step_bo_kernel(const float* __restrict__ dh, const float* __restrict__ acts, const float* __restrict__ cnext, int B, int H,
                                    float* __restrict__ dA) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float og = acts[(size_t)b * 4 * H + 3 * H + j];
    dA[(size_t)b * 4 * H + 3 * H + j] = dh[idx] * tanh_acc(cnext[idx]) * og * (1.f - og);
}
// d c' = dcnext + dh o (1 - tanh^2 c') [+ da_o . W_co]; da_i, da_f, da_g; dcprev = d c' f
__global__ void lstm_step_bc_kernel(const float* __restrict__ dh, const float* __restrict__ dcn, const float* __restrict__ dc_add,
                                    const float* __restrict__ acts, const float* __restrict__ cnext, const float* __restrict__ cprev, int B, int H,
                                    float* __restrict__ dA, float* __restrict__ dcprev) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float* ar = acts + (size_t)b * 4 * H;
    const float ig = ar[j], fg = ar[H + j], gg = ar[2 * H + j], og = ar[3 * H + j];
    const float tc = tanh_acc(cnext[idx]);
    float dc = (dcn ? dcn[idx] : 0.f) + dh[idx] * og * (1.f - tc * tc);
    if (dc_add) dc += dc_add[idx];
    const float cp = cprev ? cprev[idx] : 0.f;
    float* dr = dA + (size_t)b * 4 * H;
    dr[j] = dc * gg * ig * (1.f - ig);
    dr[H + j] = dc * cp * fg * (1.f - fg);
    dr[2 * H + j] = dc * ig * (1.f - gg * gg);
    dcprev[idx] = dc * fg;
}

}  // namespace s2s

using namespace s2s;
extern "C" {

// x [B,Din], hprev / cprev [B,H] (NULL = zeros, LSTM.lua:108-109) -> hnext, cnext [B,H]; acts [B,4H] (i | f | g | o) kept for backward
int s2s_lstm_step_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, const float* x, const float* hprev, const float* cprev,
                          int B, float* hnext, float* cnext, float* acts) {
    S2S_REQUIRE(ctx && P && x && hnext && cnext && acts, "lstm_step_forward: null argument");
    S2S_REQUIRE(B > 0 && Din > 0 && H > 0, "lstm_step_forward: bad shape");
    ctx->arena.reset();
    const LstmStepLayout L = lstm_step_layout(Din, H, peepholes);
    float *pre, *zeros;
    S2S_ALLOC(pre, ctx->arena, float, (size_t)B * 4 * H);
    S2S_ALLOC(zeros, ctx->arena, float, (size_t)B * H);
    S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * H * 4, ctx->stream));
    const float* hp = hprev ? hprev : zeros;
    const float* cp = cprev ? cprev : zeros;
    for (int g = 0; g < 4; g++) {
        float* pg = pre + g * H;
        S2S_TRY(gemm_f32(ctx, false, true, B, H, Din, 1.f, x, Din, P + L.Wx[g], Din, 0.f, pg, 4 * H, P + L.bx[g], GemmBatch(), 1, 1));
        S2S_TRY(gemm_f32(ctx, false, true, B, H, H, 1.f, hp, H, P + L.Wh[g], H, 1.f, pg, 4 * H, P + L.bh[g], GemmBatch(), 1, 1));
        if (L.Wc[g] >= 0 && g != 3)
            S2S_TRY(gemm_f32(ctx, false, true, B, H, H, 1.f, cp, H, P + L.Wc[g], H, 1.f, pg, 4 * H, P + L.bc[g], GemmBatch(), 1, 1));
    }
    const int eb = ceil_div(B * H, 256);
    lstm_step_f1_kernel<<<eb, 256, 0, ctx->stream>>>(pre, cprev, B, H, peepholes, acts, cnext, hnext);
    S2S_LAUNCH_CHECK(ctx);
    if (peepholes) {
        S2S_TRY(gemm_f32(ctx, false, true, B, H, H, 1.f, cnext, H, P + L.Wc[3], H, 1.f, pre + 3 * H, 4 * H, P + L.bc[3], GemmBatch(), 1, 1));
        lstm_step_f2_kernel<<<eb, 256, 0, ctx->stream>>>(pre, cnext, B, H, acts, hnext);
        S2S_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// dhnext [B,H], dcnext [B,H] (NULL = zeros) -> dx [B,Din], dhprev, dcprev [B,H] (overwritten); dP accumulated
int s2s_lstm_step_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, const float* x, const float* hprev,
                           const float* cprev, int B, const float* acts, const float* cnext, const float* dhnext, const float* dcnext,
                           float* dx, float* dhprev, float* dcprev) {
    S2S_REQUIRE(ctx && P && dP && x && acts && cnext && dhnext && dx && dhprev && dcprev, "lstm_step_backward: null argument");
    S2S_REQUIRE(B > 0 && Din > 0 && H > 0, "lstm_step_backward: bad shape");
    ctx->arena.reset();
    const LstmStepLayout L = lstm_step_layout(Din, H, peepholes);
    float *dA, *zeros, *dc_add = nullptr;
    S2S_ALLOC(dA, ctx->arena, float, (size_t)B * 4 * H);
    S2S_ALLOC(zeros, ctx->arena, float, (size_t)B * H);
    S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * H * 4, ctx->stream));
    const float* hp = hprev ? hprev : zeros;
    const float* cp = cprev ? cprev : zeros;
    const int eb = ceil_div(B * H, 256);
    lstm_step_bo_kernel<<<eb, 256, 0, ctx->stream>>>(dhnext, acts, cnext, B, H, dA);
    S2S_LAUNCH_CHECK(ctx);
    if (peepholes) {   // the output gate's peephole looks at next_c: d c' += da_o . W_co
        S2S_ALLOC(dc_add, ctx->arena, float, (size_t)B * H);
        S2S_TRY(gemm_f32(ctx, false, false, B, H, H, 1.f, dA + 3 * H, 4 * H, P + L.Wc[3], H, 0.f, dc_add, H, nullptr, GemmBatch(), 1, 1));
    }
    lstm_step_bc_kernel<<<eb, 256, 0, ctx->stream>>>(dhnext, dcnext, dc_add, acts, cnext, cprev, B, H, dA, dcprev);
    S2S_LAUNCH_CHECK(ctx);
    for (int g = 0; g < 4; g++) {
        const float* dAg = dA + g * H;
        // inputs
        S2S_TRY(gemm_f32(ctx, false, false, B, Din, H, 1.f, dAg, 4 * H, P + L.Wx[g], Din, g == 0 ? 0.f : 1.f, dx, Din, nullptr, GemmBatch(), 1, 1));
        S2S_TRY(gemm_f32(ctx, false, false, B, H, H, 1.f, dAg, 4 * H, P + L.Wh[g], H, g == 0 ? 0.f : 1.f, dhprev, H, nullptr, GemmBatch(), 1, 1));
        if (L.Wc[g] >= 0 && g != 3)
            S2S_TRY(gemm_f32(ctx, false, false, B, H, H, 1.f, dAg, 4 * H, P + L.Wc[g], H, 1.f, dcprev, H, nullptr, GemmBatch(), 1, 1));
        // parameters (accGradParameters)
        S2S_TRY(gemm_f32(ctx, true, false, H, Din, B, 1.f, dAg, 4 * H, x, Din, 1.f, dP + L.Wx[g], Din, nullptr, GemmBatch(), 1, 1));
        S2S_TRY(gemm_f32(ctx, true, false, H, H, B, 1.f, dAg, 4 * H, hp, H, 1.f, dP + L.Wh[g], H, nullptr, GemmBatch(), 1, 1));
        S2S_TRY(colsum_add(ctx, dAg, B, H, 4 * H, dP + L.bx[g]));
        S2S_TRY(colsum_add(ctx, dAg, B, H, 4 * H, dP + L.bh[g]));
        if (L.Wc[g] >= 0) {
            S2S_TRY(gemm_f32(ctx, true, false, H, H, B, 1.f, dAg, 4 * H, g == 3 ? cnext : cp, H, 1.f, dP + L.Wc[g], H, nullptr, GemmBatch(), 1, 1));
            S2S_TRY(colsum_add(ctx, dAg, B, H, 4 * H, dP + L.bc[g]));
        }
    }
    return 0;
}

}  // extern "C"
