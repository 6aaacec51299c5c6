// gru_step.cu -- nn.GRU as a single-step module with explicit previous state: updateOutput({x, prev_h}) -> h and
// updateGradInput (GRU.lua:8-51 through nn.Recurrent, Recurrent.lua:104-151).  The sequence kernels (gru_seq.cu) are the
// hot path; this entry exists so the module surface is complete (beam search and user code stepping a GRU by hand).
// z = sig(W_z {h,x}), r = sig(W_r {h,x}), h~ = tanh(W_h {r*h, x}), h' = (1-z) h + z h~        GRU.lua:22-30
#include "common.cuh"

namespace s2s {

__global__ void gru_step_zr_kernel(const float* __restrict__ pre, const float* __restrict__ hp, int B, int H, float* __restrict__ gates, float* __restrict__ rh) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float z = sigmoid_acc(pre[(size_t)b * 2 * H + j]), r = sigmoid_acc(pre[(size_t)b * 2 * H + H + j]);
    gates[(size_t)b * 3 * H + j] = z; gates[(size_t)b * 3 * H + H + j] = r;
    rh[idx] = r * hp[idx];
}
__global__ void gru_step_h_kernel(const float* __restrict__ pre, const float* __restrict__ hp, int B, int H, float* __restrict__ gates, float* __restrict__ hn) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float hc = tanh_acc(pre[idx]), z = gates[(size_t)b * 3 * H + j];
    gates[(size_t)b * 3 * H + 2 * H + j] = hc;
    hn[idx] = (1.f - z) * hp[idx] + z * hc;
}
// dhn -> dA = {daz, -, dah}, dhp partial = dhn (1-z)
__global__ void gru_step_b1_kernel(const float* __restrict__ dhn, const float* __restrict__ hp, const float* __restrict__ gates, int B, int H,
                                   float* __restrict__ dA, float* __restrict__ dhp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float z = gates[(size_t)b * 3 * H + j], hc = gates[(size_t)b * 3 * H + 2 * H + j], d = dhn[idx];
    dA[(size_t)b * 3 * H + 2 * H + j] = d * z * (1.f - hc * hc);
    dA[(size_t)b * 3 * H + j] = d * (hc - hp[idx]) * z * (1.f - z);
    dhp[idx] = d * (1.f - z);
}
// d(r*h) -> dar ; dhp += d(r*h) r
__global__ void gru_step_b2_kernel(const float* __restrict__ drh, const float* __restrict__ hp, const float* __restrict__ gates, int B, int H,
                                   float* __restrict__ dA, float* __restrict__ dhp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const float r = gates[(size_t)b * 3 * H + H + j];
    dA[(size_t)b * 3 * H + H + j] = drh[idx] * hp[idx] * r * (1.f - r);
    dhp[idx] += drh[idx] * r;
}

__global__ void gru_step_rh_kernel(const float* __restrict__ gates, const float* __restrict__ hp, int B, int H, float* __restrict__ rh) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    rh[idx] = gates[(size_t)(idx / H) * 3 * H + H + idx % H] * hp[idx];
}

}  // namespace s2s

using namespace s2s;
extern "C" {

// W: z, r, h~ weights [3][H][H+Din]; x [B,Din], hprev [B,H] (NULL = zeros, Recurrent.lua:110-112) -> hnext [B,H];
// gates [B,3H] (z | r | h~) kept by the caller for the backward call.
int s2s_gru_step_forward(s2s_ctx* ctx, const float* W, int Din, int H, const float* x, const float* hprev, int B, float* hnext, float* gates) {
    S2S_REQUIRE(ctx && W && x && hnext && gates && B > 0 && Din > 0 && H > 0, "gru_step_forward: bad arguments");
    ctx->arena.reset();
    const int ldw = H + Din;
    float *pre, *rh, *zeros = nullptr;
    S2S_ALLOC(pre, ctx->arena, float, (size_t)B * 2 * H);
    S2S_ALLOC(rh, ctx->arena, float, (size_t)B * H);
    if (!hprev) { S2S_ALLOC(zeros, ctx->arena, float, (size_t)B * H); S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * H * 4, ctx->stream)); hprev = zeros; }
    const int eb = ceil_div(B * H, 256);
    // {z,r} pre-activations: h part then x part (LinearZeroBias on the concatenation {prev_h, x}, GRU.lua:22-24)
    S2S_TRY(gemm_f32(ctx, false, true, B, 2 * H, H, 1.f, hprev, H, W, ldw, 0.f, pre, 2 * H, nullptr, GemmBatch(), 1, 1));
    S2S_TRY(gemm_f32(ctx, false, true, B, 2 * H, Din, 1.f, x, Din, W + H, ldw, 1.f, pre, 2 * H, nullptr, GemmBatch(), 1, 1));
    gru_step_zr_kernel<<<eb, 256, 0, ctx->stream>>>(pre, hprev, B, H, gates, rh);
    S2S_LAUNCH_CHECK(ctx);
    const float* Wh = W + (size_t)2 * H * ldw;
    S2S_TRY(gemm_f32(ctx, false, true, B, H, H, 1.f, rh, H, Wh, ldw, 0.f, pre, H, nullptr, GemmBatch(), 1, 1));
    S2S_TRY(gemm_f32(ctx, false, true, B, H, Din, 1.f, x, Din, Wh + H, ldw, 1.f, pre, H, nullptr, GemmBatch(), 1, 1));
    gru_step_h_kernel<<<eb, 256, 0, ctx->stream>>>(pre, hprev, B, H, gates, hnext);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// dhnext [B,H] -> dx [B,Din], dhprev [B,H] (both overwritten); dW accumulated (scale 1)
int s2s_gru_step_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, const float* x, const float* hprev, int B,
                          const float* gates, const float* dhnext, float* dx, float* dhprev) {
    S2S_REQUIRE(ctx && W && dW && x && gates && dhnext && dx && dhprev && B > 0, "gru_step_backward: bad arguments");
    ctx->arena.reset();
    const int ldw = H + Din;
    float *dA, *drh, *rh, *zeros = nullptr;
    S2S_ALLOC(dA, ctx->arena, float, (size_t)B * 3 * H);
    S2S_ALLOC(drh, ctx->arena, float, (size_t)B * H);
    S2S_ALLOC(rh, ctx->arena, float, (size_t)B * H);
    if (!hprev) { S2S_ALLOC(zeros, ctx->arena, float, (size_t)B * H); S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * H * 4, ctx->stream)); hprev = zeros; }
    const int eb = ceil_div(B * H, 256);
    const float* Wh = W + (size_t)2 * H * ldw;
    float* dWh = dW + (size_t)2 * H * ldw;
    gru_step_b1_kernel<<<eb, 256, 0, ctx->stream>>>(dhnext, hprev, gates, B, H, dA, dhprev);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(gemm_f32(ctx, false, false, B, H, H, 1.f, dA + 2 * H, 3 * H, Wh, ldw, 0.f, drh, H, nullptr, GemmBatch(), 1, 1));        // d(r*h) = dah W_h[:, :H]
    gru_step_b2_kernel<<<eb, 256, 0, ctx->stream>>>(drh, hprev, gates, B, H, dA, dhprev);
    S2S_LAUNCH_CHECK(ctx);
    // dhprev += {daz, dar} W_{z,r}[:, :H] ; dx = dA W[:, H:]
    S2S_TRY(gemm_f32(ctx, false, false, B, H, 2 * H, 1.f, dA, 3 * H, W, ldw, 1.f, dhprev, H, nullptr, GemmBatch(), 1, 1));
    S2S_TRY(gemm_f32(ctx, false, false, B, Din, 3 * H, 1.f, dA, 3 * H, W + H, ldw, 0.f, dx, Din, nullptr, GemmBatch(), 1, 1));
    // weight gradients: z, r rows see {h, x}; the candidate rows see {r*h, x}
    S2S_TRY(gemm_f32(ctx, true, false, 2 * H, H, B, 1.f, dA, 3 * H, hprev, H, 1.f, dW, ldw, nullptr, GemmBatch(), 1, 1));
    S2S_TRY(gemm_f32(ctx, true, false, 3 * H, Din, B, 1.f, dA, 3 * H, x, Din, 1.f, dW + H, ldw, nullptr, GemmBatch(), 1, 1));
    // r*h recomputed from the saved gate
    gru_step_rh_kernel<<<eb, 256, 0, ctx->stream>>>(gates, hprev, B, H, rh);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(gemm_f32(ctx, true, false, H, H, B, 1.f, dA + 2 * H, 3 * H, rh, H, 1.f, dWh, ldw, nullptr, GemmBatch(), 1, 1));
    return 0;
}

}  // extern "C"
