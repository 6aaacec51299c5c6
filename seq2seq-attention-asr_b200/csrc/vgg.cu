// vgg.cu -- the convolutional front-end of librispeech/model_vgg.lua:23-54, forward and backward:
//     4 x [SpatialConvolutionMM 3x3 (valid) + ReLU], SpatialMaxPooling(2,1,2,1) after the 2nd and (2,2,2,2) after the 4th,
//     Transpose2 + View to [L, nFeat*H], 4 x [TemporalConvolution(k=1) + ReLU]  ->  annotations h [B, L, OUT].
//
// First CUDA path (round 1): activations live channels-last ([B, time, freq, C]) so that every contraction is a
// row-major GEMM with K contiguous -- the 3x3 convolutions unfold their input into [pixels, 9C] patches (what Torch7's
// SpatialConvolutionMM does too) and run on the tcgen05 GEMM of gemm_tc.cu; bias is fused in the GEMM epilogue, ReLU,
// pooling and the layout changes are streaming kernels.  The module's parameter layouts are kept at the boundary:
// convolution weights [nOut, nIn*3*3] in (plane, kh, kw) order and the first 1x1 layer's [HID, nFeat*H] in (feature, freq)
// order are permuted per call into the channels-last order the GEMMs read, and their gradients are permuted back.
// Utterances are processed in chunks so the patch matrix stays bounded.  Next: implicit GEMM (TMA tap offsets, no patch
// matrix), 64/128-wide GEMM tiles for the 64- and 128-plane layers, ReLU in the GEMM epilogue.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace s2s {

struct VggDims {
    int C1, C2, HID, OUT, B, T, F;
    int H1, W1, H2, W2, Wp1, H3, W3, H4, W4, L, Wq, view;
    int64_t off[16], n;          // parameter offsets: conv1..4 (W, b), t1..4 (W, b)
};
static int vgg_dims(const s2s_vgg_cfg* cfg, int B, int T, int F, VggDims* d) {
    S2S_REQUIRE(cfg && cfg->C1 > 0 && cfg->C2 > 0 && cfg->HID > 0 && cfg->OUT > 0, "vgg: bad configuration");
    S2S_REQUIRE(cfg->C1 % 4 == 0 && cfg->C2 % 4 == 0, "vgg: plane counts must be multiples of 4 (C1=%d C2=%d)", cfg->C1, cfg->C2);
    S2S_REQUIRE(T >= 10 && F >= 12, "vgg: input [T=%d, F=%d] too small for four 3x3 convolutions and two poolings", T, F);
    d->C1 = cfg->C1; d->C2 = cfg->C2; d->HID = cfg->HID; d->OUT = cfg->OUT; d->B = B; d->T = T; d->F = F;
    d->H1 = T - 2; d->W1 = F - 2; d->H2 = T - 4; d->W2 = F - 4; d->Wp1 = d->W2 / 2;
    d->H3 = d->H2 - 2; d->W3 = d->Wp1 - 2; d->H4 = d->H3 - 2; d->W4 = d->W3 - 2;
    d->L = d->H4 / 2; d->Wq = d->W4 / 2; d->view = d->C2 * d->Wq;
    S2S_REQUIRE(d->L >= 1 && d->Wq >= 1, "vgg: input too small");
    const int64_t sz[16] = {(int64_t)d->C1 * 27, d->C1, (int64_t)d->C1 * d->C1 * 9, d->C1, (int64_t)d->C2 * d->C1 * 9, d->C2,
                            (int64_t)d->C2 * d->C2 * 9, d->C2, (int64_t)d->HID * d->view, d->HID, (int64_t)d->HID * d->HID, d->HID,
                            (int64_t)d->HID * d->HID, d->HID, (int64_t)d->OUT * d->HID, d->OUT};
    int64_t o = 0;
    for (int i = 0; i < 16; i++) { d->off[i] = o; o += sz[i]; }
    d->n = o;
    return 0;
}

struct VggState {
    Arena mem;
    bool valid = false, implicit = false;
    VggDims d;
    float *a0 = nullptr, *a1 = nullptr, *a2 = nullptr, *p1 = nullptr, *a3 = nullptr, *a4 = nullptr, *f[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};
void vgg_state_free(s2s_ctx* ctx) {
    if (!ctx->vgg) return;
    ctx->vgg->mem.release();
    delete ctx->vgg;
    ctx->vgg = nullptr;
}

// ---- streaming kernels (channels-last) -----------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int64_t HW, int64_t n, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over [B, HW, C]
    if (i >= n) return;
    const int c = (int)(i % C);
    const int64_t p = (i / C) % HW, b = i / (C * HW);
    y[i] = x[(b * C + c) * HW + p];
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int C, int64_t HW, int64_t n, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over [B, C, HW]
    if (i >= n) return;
    const int64_t p = i % HW;
    const int c = (int)((i / HW) % C);
    const int64_t b = i / (HW * C);
    y[i] = x[(b * HW + p) * C + c];
}
// patches: col[(b, y, x), (kh, kw, c)] = in[b, y + kh, x + kw, c].  VEC = 4: four channels per thread (C % 4 == 0), 16-byte
// loads and stores; one 64-bit division per thread, the rest of the index arithmetic is 32-bit.
template <int VEC>
__global__ void unfold3_kernel(const float* __restrict__ in, int nb, int Hh, int Ww, int C, float* __restrict__ col) {
    const int Ho = Hh - 2, Wo = Ww - 2, CV = C / VEC, KV = 9 * CV;
    const int64_t n = (int64_t)nb * Ho * Wo * KV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / KV;
        const int k = (int)(i - m * KV);
        const int t = k / CV, c = (k - t * CV) * VEC, kh = t / 3, kw = t - 3 * kh;
        const int64_t by = m / Wo;                       // b * Ho + y
        const int x = (int)(m - by * Wo);
        const int b = (int)(by / Ho), y = (int)(by - (int64_t)b * Ho);
        const float* src = in + (((int64_t)b * Hh + y + kh) * Ww + x + kw) * C + c;
        float* dst = col + m * (9 * C) + t * C + c;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
        else *dst = __ldg(src);
    }
}
// din[b, y, x, c] = sum over taps of dcol[(b, y - kh, x - kw), (kh, kw, c)]
template <int VEC>
__global__ void fold3_kernel(const float* __restrict__ dcol, int nb, int Hh, int Ww, int C, float* __restrict__ din) {
    const int Ho = Hh - 2, Wo = Ww - 2, K = 9 * C, CV = C / VEC;
    const int64_t n = (int64_t)nb * Hh * Ww * CV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / CV;                        // (b * Hh + y) * Ww + x
        const int c = (int)(i - p * CV) * VEC;
        const int64_t by = p / Ww;
        const int x = (int)(p - by * Ww);
        const int b = (int)(by / Hh), y = (int)(by - (int64_t)b * Hh);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; kh++) {
            const int yo = y - kh;
            if (yo < 0 || yo >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; kw++) {
                const int xo = x - kw;
                if (xo < 0 || xo >= Wo) continue;
                const float* src = dcol + (((int64_t)b * Ho + yo) * Wo + xo) * K + (kh * 3 + kw) * C + c;
                if (VEC == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(src)); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
                else s.x += __ldg(src);
            }
        }
        if (VEC == 4) *reinterpret_cast<float4*>(din + p * C + c) = s;
        else din[p * C + c] = s.x;
    }
}
__global__ void relu_kernel(float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fmaxf(y[i], 0.f);
}
// dy *= (y > 0)
__global__ void relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dy[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// SpatialMaxPooling(kW, kH, kW, kH), floor mode.  The input lives on a grid of pitch (Hh, Ww) of which (Hv, Wv) is valid
// (the implicit convolutions keep their outputs on the input's grid); the output is compact.
__global__ void pool_fwd_kernel(const float* __restrict__ in, int nb, int Hh, int Ww, int Hv, int Wv, int C, int kH, int kW, float* __restrict__ out) {
    const int Ho = Hv / kH, Wo = Wv / kW;
    const int64_t n = (int64_t)nb * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int xo = (int)((i / C) % Wo), yo = (int)((i / ((int64_t)C * Wo)) % Ho);
        const int64_t b = i / ((int64_t)C * Wo * Ho);
        float m = -INFINITY;
        for (int kh = 0; kh < kH; kh++)
            for (int kw = 0; kw < kW; kw++) m = fmaxf(m, in[((b * Hh + yo * kH + kh) * Ww + xo * kW + kw) * C + c]);
        out[i] = m;
    }
}
// the gradient goes to the first maximum of the window in (kh, kw) scan order; written on the input's grid, zero outside
// the pooled region (so the convolution below sees zeros wherever its output is not valid).  RELU: the pooled tensor is a
// ReLU output, so its backward mask (in > 0) is applied in the same pass.  Four channels per thread (C % 4 == 0).
template <bool RELU>
__global__ void pool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, int nb, int Hh, int Ww, int Hv, int Wv, int C, int kH,
                                int kW, float* __restrict__ din) {
    const int Ho = Hv / kH, Wo = Wv / kW, CV = C >> 2;
    const int64_t n = (int64_t)nb * Hh * Ww * CV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / CV;                        // (b * Hh + y) * Ww + x
        const int c = (int)(i - p * CV) * 4;
        const int64_t by = p / Ww;
        const int x = (int)(p - by * Ww);
        const int b = (int)(by / Hh), y = (int)(by - (int64_t)b * Hh);
        const int yo = y / kH, xo = x / kW;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yo < Ho && xo < Wo) {
            const int me = (y - yo * kH) * kW + (x - xo * kW);
            int4 best = make_int4(0, 0, 0, 0);
            float4 bv = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), mine = bv;
            for (int kh = 0; kh < kH; kh++)
                for (int kw = 0; kw < kW; kw++) {
                    const float4 v = *reinterpret_cast<const float4*>(in + (((int64_t)b * Hh + yo * kH + kh) * Ww + xo * kW + kw) * C + c);
                    const int k = kh * kW + kw;
                    if (v.x > bv.x) { bv.x = v.x; best.x = k; }
                    if (v.y > bv.y) { bv.y = v.y; best.y = k; }
                    if (v.z > bv.z) { bv.z = v.z; best.z = k; }
                    if (v.w > bv.w) { bv.w = v.w; best.w = k; }
                    if (k == me) mine = v;
                }
            const float4 d = *reinterpret_cast<const float4*>(dout + (((int64_t)b * Ho + yo) * Wo + xo) * C + c);
            g.x = (best.x == me && (!RELU || mine.x > 0.f)) ? d.x : 0.f;
            g.y = (best.y == me && (!RELU || mine.y > 0.f)) ? d.y : 0.f;
            g.z = (best.z == me && (!RELU || mine.z > 0.f)) ? d.z : 0.f;
            g.w = (best.w == me && (!RELU || mine.w > 0.f)) ? d.w : 0.f;
        }
        *reinterpret_cast<float4*>(din + p * C + c) = g;
    }
}
// WpT[c, t*N + n] = Wp[n, t*C + c]   (the data-gradient operand of the implicit convolution)
__global__ void wperm_t_kernel(const float* __restrict__ Wp, int N, int C, float* __restrict__ WpT) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over WpT [C, 9, N]
    if (i >= (int64_t)N * C * 9) return;
    const int n = (int)(i % N), t = (int)((i / N) % 9), c = (int)(i / ((int64_t)N * 9));
    WpT[i] = Wp[((int64_t)n * 9 + t) * C + c];
}
// module weight [N, C*S] in (c, s) column order  <->  channels-last [N, S*C] in (s, c) order   (S = 9 taps, or the freq bins)
__global__ void wperm_kernel(const float* __restrict__ W, int N, int C, int S, float* __restrict__ Wp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over Wp
    if (i >= (int64_t)N * C * S) return;
    const int c = (int)(i % C), s = (int)((i / C) % S);
    const int64_t n = i / ((int64_t)C * S);
    Wp[i] = W[(n * C + c) * S + s];
}
__global__ void wperm_back_add_kernel(const float* __restrict__ dWp, int N, int C, int S, float* __restrict__ dW) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over dW
    if (i >= (int64_t)N * C * S) return;
    const int s = (int)(i % S), c = (int)((i / S) % C);
    const int64_t n = i / ((int64_t)C * S);
    dW[i] += dWp[(n * S + s) * C + c];
}

static int grid_for(s2s_ctx* ctx, int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)ctx->sm_count * 32;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
#define VGG_LAUNCH(kernel, n, ...)                                                       \
    do {                                                                                 \
        kernel<<<grid_for(ctx, (n)), 256, 0, ctx->stream>>>(__VA_ARGS__);                \
        S2S_LAUNCH_CHECK(ctx);                                                           \
    } while (0)
#define VGG_LAUNCH_FLAT(kernel, n, ...)                                                  \
    do {                                                                                 \
        kernel<<<(unsigned)ceil_div64((n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__);    \
        S2S_LAUNCH_CHECK(ctx);                                                           \
    } while (0)

// one convolution + ReLU on a chunk of nb utterances: in [nb, Hh, Ww, C] -> out [nb, Hh-2, Ww-2, N]
static int conv_relu_fwd(s2s_ctx* ctx, const float* in, int nb, int Hh, int Ww, int C, const float* Wp, const float* bias, int N, float* col,
                         float* out) {
    const int64_t M = (int64_t)nb * (Hh - 2) * (Ww - 2);
    if (C % 4 == 0) VGG_LAUNCH(unfold3_kernel<4>, M * 9 * C / 4, in, nb, Hh, Ww, C, col);
    else VGG_LAUNCH(unfold3_kernel<1>, M * 9 * C, in, nb, Hh, Ww, C, col);
    S2S_TRY(gemm_f32(ctx, false, true, (int)M, N, 9 * C, 1.f, col, 9 * C, Wp, 9 * C, 0.f, out, N, bias, GemmBatch(), 1, 0, true));   // bias + ReLU
    return 0;
}
// dout [M, N] (gradient w.r.t. the ReLU output; masked here) -> dWp += , db += , din (nullable)
static int conv_relu_bwd(s2s_ctx* ctx, const float* in, const float* out, float* dout, int nb, int Hh, int Ww, int C, const float* Wp, int N,
                         float* col, float* dcol, float* dWp, float* db, float* din) {
    const int64_t M = (int64_t)nb * (Hh - 2) * (Ww - 2);
    VGG_LAUNCH(relu_bwd_kernel, M * N, dout, out, M * N);
    if (C % 4 == 0) VGG_LAUNCH(unfold3_kernel<4>, M * 9 * C / 4, in, nb, Hh, Ww, C, col);
    else VGG_LAUNCH(unfold3_kernel<1>, M * 9 * C, in, nb, Hh, Ww, C, col);
    S2S_TRY(gemm_f32(ctx, true, false, N, 9 * C, (int)M, 1.f, dout, N, col, 9 * C, 1.f, dWp, 9 * C, nullptr, GemmBatch(), 8));
    S2S_TRY(colsum_add(ctx, dout, M, N, N, db));
    if (din) {
        S2S_TRY(gemm_f32(ctx, false, false, (int)M, 9 * C, N, 1.f, dout, N, Wp, 9 * C, 0.f, dcol, 9 * C));
        if (C % 4 == 0) VGG_LAUNCH(fold3_kernel<4>, (int64_t)nb * Hh * Ww * C / 4, dcol, nb, Hh, Ww, C, din);
        else VGG_LAUNCH(fold3_kernel<1>, (int64_t)nb * Hh * Ww * C, dcol, nb, Hh, Ww, C, din);
    }
    return 0;
}

int conv3_tc_forward(s2s_ctx* ctx, const float* in, int64_t Mg, int Ww, int C, const float* Wp, const float* bias, int N, float* out, bool relu);
int conv3_tc_dgrad(s2s_ctx* ctx, const float* dout, int64_t Mg, int Ww, int N, const float* WpT, int C, float* din);
int conv3_tc_wgrad(s2s_ctx* ctx, const float* dout, const float* in, int64_t Mg, int Ww, int N, int C, float* dWp);

// Implicit-GEMM path (default when the plane counts are multiples of 32): conv2-4 read their input through nine TMA row
// offsets instead of a patch matrix, and keep their output on the INPUT's grid -- conv2 on the (H1, W1) grid of conv1's
// output, conv3 and conv4 on the (H2, Wp1) grid of the first pooling -- with a valid region that shrinks by two per layer;
// the poolings compact.  conv1 (3 planes) keeps the explicit patch path.
static bool vgg_implicit(const VggDims& d) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("S2S_VGG_IMPLICIT"); enabled = e ? atoi(e) : 1; }
    return enabled && d.C1 % 32 == 0 && d.C2 % 32 == 0;
}

static int vgg_chunk(const VggDims& d) {
    // patch matrices of the largest layer for one utterance, in floats
    const int64_t per = std::max((int64_t)d.H2 * d.W2 * 9 * d.C1, (int64_t)d.H4 * d.W4 * 9 * d.C2);
    int64_t nb = ((int64_t)96 << 20) / (per > 0 ? per : 1);      // <= 384 MB of patches per chunk
    if (nb < 1) nb = 1;
    if (nb > d.B) nb = d.B;
    return (int)nb;
}

}  // namespace s2s

using namespace s2s;
extern "C" {

int64_t s2s_vgg_param_count(const s2s_vgg_cfg* cfg, int F) {
    VggDims d;
    if (vgg_dims(cfg, 1, 64, F, &d)) return -1;
    return d.n;
}
int s2s_vgg_out_len(int T) { return (T - 8) / 2; }

int s2s_vgg_forward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, const float* X, int B, int T, int F, float* h) {
    S2S_REQUIRE(ctx && P && X && h && B > 0, "vgg_forward: bad arguments");
    VggDims d;
    S2S_TRY(vgg_dims(cfg, B, T, F, &d));
    ctx->arena.reset();
    if (!ctx->vgg) ctx->vgg = new VggState();
    VggState& s = *ctx->vgg;
    s.valid = false;
    s.mem.reset();
    s.d = d;
    const int C1 = d.C1, C2 = d.C2;
    S2S_ALLOC(s.a0, s.mem, float, (size_t)B * T * F * 3);
    S2S_ALLOC(s.a1, s.mem, float, (size_t)B * d.H1 * d.W1 * C1);
    const bool imp = vgg_implicit(d);
    s.implicit = imp;
    // per-utterance sizes: compact (explicit path) or on the producing grid (implicit path)
    const size_t n_a2 = imp ? (size_t)d.H1 * d.W1 * C1 : (size_t)d.H2 * d.W2 * C1;
    const size_t n_a3 = imp ? (size_t)d.H2 * d.Wp1 * C2 : (size_t)d.H3 * d.W3 * C2;
    const size_t n_a4 = imp ? (size_t)d.H2 * d.Wp1 * C2 : (size_t)d.H4 * d.W4 * C2;
    S2S_ALLOC(s.a2, s.mem, float, (size_t)B * n_a2);
    S2S_ALLOC(s.p1, s.mem, float, (size_t)B * d.H2 * d.Wp1 * C1);
    S2S_ALLOC(s.a3, s.mem, float, (size_t)B * n_a3);
    S2S_ALLOC(s.a4, s.mem, float, (size_t)B * n_a4);
    const int64_t rows = (int64_t)B * d.L;
    S2S_ALLOC(s.f[0], s.mem, float, (size_t)rows * d.view);
    for (int k = 1; k <= 3; k++) S2S_ALLOC(s.f[k], s.mem, float, (size_t)rows * d.HID);
    S2S_ALLOC(s.f[4], s.mem, float, (size_t)rows * d.OUT);

    Arena& ar = ctx->arena;
    float *Wp[4], *W1p;
    const int cin[4] = {3, C1, C1, C2}, cout[4] = {C1, C1, C2, C2};
    for (int l = 0; l < 4; l++) {
        S2S_ALLOC(Wp[l], ar, float, (size_t)cout[l] * cin[l] * 9);
        VGG_LAUNCH_FLAT(wperm_kernel, (int64_t)cout[l] * cin[l] * 9, P + d.off[2 * l], cout[l], cin[l], 9, Wp[l]);
    }
    S2S_ALLOC(W1p, ar, float, (size_t)d.HID * d.view);
    VGG_LAUNCH_FLAT(wperm_kernel, (int64_t)d.HID * d.view, P + d.off[8], d.HID, C2, d.Wq, W1p);
    VGG_LAUNCH_FLAT(nchw_to_nhwc_kernel, (int64_t)B * T * F * 3, X, 3, (int64_t)T * F, (int64_t)B * T * F * 3, s.a0);

    const int chunk = imp ? std::min(B, 8) : vgg_chunk(d);
    float* col;
    S2S_ALLOC(col, ar, float, imp ? (size_t)chunk * d.H1 * d.W1 * 27 : (size_t)chunk * std::max((int64_t)d.H2 * d.W2 * 9 * C1, (int64_t)d.H4 * d.W4 * 9 * C2));
    const std::vector<size_t> mk = ar.mark();
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        ar.rewind(mk);                                   // the GEMMs' operand copies of the previous chunk
        float* a1 = s.a1 + (size_t)b0 * d.H1 * d.W1 * C1;
        float* a2 = s.a2 + (size_t)b0 * n_a2;
        float* p1 = s.p1 + (size_t)b0 * d.H2 * d.Wp1 * C1;
        float* a3 = s.a3 + (size_t)b0 * n_a3;
        float* a4 = s.a4 + (size_t)b0 * n_a4;
        float* f0 = s.f[0] + (size_t)b0 * d.L * d.view;
        S2S_TRY(conv_relu_fwd(ctx, s.a0 + (size_t)b0 * T * F * 3, nb, T, F, 3, Wp[0], P + d.off[1], C1, col, a1));
        if (imp) {
            const int64_t M1 = (int64_t)nb * d.H1 * d.W1, M2 = (int64_t)nb * d.H2 * d.Wp1;
            static const int fused = []() { const char* e = getenv("S2S_VGG_FUSED"); return e ? atoi(e) : 7; }();
            S2S_TRY(conv3_tc_forward(ctx, a1, M1, d.W1, C1, Wp[1], P + d.off[3], C1, a2, fused & 1));       // + ReLU; grid (H1, W1), valid (H2, W2)
            if (!(fused & 1)) VGG_LAUNCH(relu_kernel, M1 * C1, a2, M1 * C1);
            VGG_LAUNCH(pool_fwd_kernel, (int64_t)nb * d.H2 * d.Wp1 * C1, a2, nb, d.H1, d.W1, d.H2, d.W2, C1, 1, 2, p1);
            S2S_TRY(conv3_tc_forward(ctx, p1, M2, d.Wp1, C1, Wp[2], P + d.off[5], C2, a3, fused & 2));      // grid (H2, Wp1), valid (H3, W3)
            if (!(fused & 2)) VGG_LAUNCH(relu_kernel, M2 * C2, a3, M2 * C2);
            S2S_TRY(conv3_tc_forward(ctx, a3, M2, d.Wp1, C2, Wp[3], P + d.off[7], C2, a4, fused & 4));      // same grid, valid (H4, W4)
            if (!(fused & 4)) VGG_LAUNCH(relu_kernel, M2 * C2, a4, M2 * C2);
            VGG_LAUNCH(pool_fwd_kernel, (int64_t)nb * d.L * d.view, a4, nb, d.H2, d.Wp1, d.H4, d.W4, C2, 2, 2, f0);
        } else {
            S2S_TRY(conv_relu_fwd(ctx, a1, nb, d.H1, d.W1, C1, Wp[1], P + d.off[3], C1, col, a2));
            VGG_LAUNCH(pool_fwd_kernel, (int64_t)nb * d.H2 * d.Wp1 * C1, a2, nb, d.H2, d.W2, d.H2, d.W2, C1, 1, 2, p1);
            S2S_TRY(conv_relu_fwd(ctx, p1, nb, d.H2, d.Wp1, C1, Wp[2], P + d.off[5], C2, col, a3));
            S2S_TRY(conv_relu_fwd(ctx, a3, nb, d.H3, d.W3, C2, Wp[3], P + d.off[7], C2, col, a4));
            // pooled [nb, L, Wq, C2] is the flattened feature matrix [nb*L, view] in (freq, feature) column order
            VGG_LAUNCH(pool_fwd_kernel, (int64_t)nb * d.L * d.view, a4, nb, d.H4, d.W4, d.H4, d.W4, C2, 2, 2, f0);
        }
    }
    ar.rewind(mk);
    // 1x1 stack over all frames: TemporalConvolution(k = 1) + ReLU   (model_vgg.lua:47-54)
    const float* Wk[4] = {W1p, P + d.off[10], P + d.off[12], P + d.off[14]};
    const int kin[4] = {d.view, d.HID, d.HID, d.HID}, kout[4] = {d.HID, d.HID, d.HID, d.OUT};
    for (int k = 0; k < 4; k++) {
        S2S_TRY(gemm_f32(ctx, false, true, (int)rows, kout[k], kin[k], 1.f, s.f[k], kin[k], Wk[k], kin[k], 0.f, s.f[k + 1], kout[k], P + d.off[9 + 2 * k],
                         GemmBatch(), 1, 0, true));                                   // bias + ReLU in the epilogue
    }
    S2S_CUDA(cudaMemcpyAsync(h, s.f[4], (size_t)rows * d.OUT * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    s.valid = true;
    return 0;
}

int s2s_vgg_backward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, float* dP, int B, int T, int F, const float* dh, float* dX) {
    S2S_REQUIRE(ctx && P && dP && dh && B > 0, "vgg_backward: bad arguments");
    S2S_REQUIRE(ctx->vgg && ctx->vgg->valid, "vgg_backward called without a preceding vgg_forward on this context");
    VggState& s = *ctx->vgg;
    VggDims d;
    S2S_TRY(vgg_dims(cfg, B, T, F, &d));
    S2S_REQUIRE(d.n == s.d.n && s.d.B == B && s.d.T == T && s.d.F == F, "vgg_backward: shapes differ from the preceding forward");
    ctx->arena.reset();
    Arena& ar = ctx->arena;
    const int C1 = d.C1, C2 = d.C2;
    const int cin[4] = {3, C1, C1, C2}, cout[4] = {C1, C1, C2, C2};
    const int64_t rows = (int64_t)B * d.L;
    float *Wp[4], *dWp[4], *W1p, *dW1p;
    for (int l = 0; l < 4; l++) {
        const int64_t n = (int64_t)cout[l] * cin[l] * 9;
        S2S_ALLOC(Wp[l], ar, float, n);
        S2S_ALLOC(dWp[l], ar, float, n);
        VGG_LAUNCH_FLAT(wperm_kernel, n, P + d.off[2 * l], cout[l], cin[l], 9, Wp[l]);
        S2S_CUDA(cudaMemsetAsync(dWp[l], 0, n * 4, ctx->stream));
    }
    S2S_ALLOC(W1p, ar, float, (size_t)d.HID * d.view);
    S2S_ALLOC(dW1p, ar, float, (size_t)d.HID * d.view);
    VGG_LAUNCH_FLAT(wperm_kernel, (int64_t)d.HID * d.view, P + d.off[8], d.HID, C2, d.Wq, W1p);
    S2S_CUDA(cudaMemsetAsync(dW1p, 0, (size_t)d.HID * d.view * 4, ctx->stream));

    // ---- 1x1 stack ------------------------------------------------------------------------------------------------------
    const float* Wk[4] = {W1p, P + d.off[10], P + d.off[12], P + d.off[14]};
    float* dWk[4] = {dW1p, dP + d.off[10], dP + d.off[12], dP + d.off[14]};
    const int kin[4] = {d.view, d.HID, d.HID, d.HID}, kout[4] = {d.HID, d.HID, d.HID, d.OUT};
    // two gradient buffers alternate down the stack; both are sized for the widest layer
    float *ga, *gb;
    S2S_ALLOC(ga, ar, float, (size_t)rows * std::max(std::max(d.HID, d.OUT), d.view));
    S2S_ALLOC(gb, ar, float, (size_t)rows * std::max(std::max(d.HID, d.OUT), d.view));
    S2S_CUDA(cudaMemcpyAsync(ga, dh, (size_t)rows * d.OUT * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int k = 3; k >= 0; k--) {
        VGG_LAUNCH(relu_bwd_kernel, rows * kout[k], ga, s.f[k + 1], rows * kout[k]);
        S2S_TRY(gemm_f32(ctx, true, false, kout[k], kin[k], (int)rows, 1.f, ga, kout[k], s.f[k], kin[k], 1.f, dWk[k], kin[k], nullptr, GemmBatch(), 8));
        S2S_TRY(colsum_add(ctx, ga, rows, kout[k], kout[k], dP + d.off[9 + 2 * k]));
        S2S_TRY(gemm_f32(ctx, false, false, (int)rows, kin[k], kout[k], 1.f, ga, kout[k], Wk[k], kin[k], 0.f, gb, kin[k]));
        std::swap(ga, gb);
    }
    // ga = d f0 [B*L, view] = d pooled [B, L, Wq, C2]

    // ---- convolutional part, chunk by chunk ------------------------------------------------------------------------------
    const bool imp = s.implicit;
    const size_t n_a2 = imp ? (size_t)d.H1 * d.W1 * C1 : (size_t)d.H2 * d.W2 * C1;
    const size_t n_a3 = imp ? (size_t)d.H2 * d.Wp1 * C2 : (size_t)d.H3 * d.W3 * C2;
    const size_t n_a4 = imp ? (size_t)d.H2 * d.Wp1 * C2 : (size_t)d.H4 * d.W4 * C2;
    const int chunk = imp ? std::min(B, 8) : vgg_chunk(d);
    const int64_t colmax = imp ? (int64_t)d.H1 * d.W1 * 27 : std::max((int64_t)d.H2 * d.W2 * 9 * C1, (int64_t)d.H4 * d.W4 * 9 * C2);
    const int64_t actmax = std::max(std::max((int64_t)d.H1 * d.W1 * C1, (int64_t)d.H2 * d.Wp1 * C2), (int64_t)T * F * 3);
    float *col, *dcol, *da, *db_, *WpT[4] = {nullptr, nullptr, nullptr, nullptr};
    S2S_ALLOC(col, ar, float, (size_t)chunk * colmax);
    S2S_ALLOC(dcol, ar, float, (size_t)chunk * colmax);
    S2S_ALLOC(da, ar, float, (size_t)chunk * actmax);
    S2S_ALLOC(db_, ar, float, (size_t)chunk * actmax);
    if (imp) {
        for (int l = 1; l < 4; l++) {
            S2S_ALLOC(WpT[l], ar, float, (size_t)cout[l] * cin[l] * 9);
            VGG_LAUNCH_FLAT(wperm_t_kernel, (int64_t)cout[l] * cin[l] * 9, Wp[l], cout[l], cin[l], WpT[l]);
        }
    }
    float* dXc = nullptr;
    if (dX) S2S_ALLOC(dXc, ar, float, (size_t)B * T * F * 3);
    const std::vector<size_t> mk = ar.mark();
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        ar.rewind(mk);
        const float* a0 = s.a0 + (size_t)b0 * T * F * 3;
        const float* a1 = s.a1 + (size_t)b0 * d.H1 * d.W1 * C1;
        const float* a2 = s.a2 + (size_t)b0 * n_a2;
        const float* p1 = s.p1 + (size_t)b0 * d.H2 * d.Wp1 * C1;
        const float* a3 = s.a3 + (size_t)b0 * n_a3;
        const float* a4 = s.a4 + (size_t)b0 * n_a4;
        const float* gf0 = ga + (size_t)b0 * d.L * d.view;
        if (imp) {
            const int64_t M1 = (int64_t)nb * d.H1 * d.W1, M2 = (int64_t)nb * d.H2 * d.Wp1;
            // pool2 backward onto the (H2, Wp1) grid: zero outside the valid (H4, W4) region, which is what keeps the
            // wrap-around rows of the implicit data / weight gradients out of the sums
            VGG_LAUNCH(pool_bwd_kernel<true>, M2 * C2 / 4, a4, gf0, nb, d.H2, d.Wp1, d.H4, d.W4, C2, 2, 2, da);   // + ReLU mask of a4
            S2S_TRY(conv3_tc_wgrad(ctx, da, a3, M2, d.Wp1, C2, C2, dWp[3]));
            S2S_TRY(colsum_add(ctx, da, M2, C2, C2, dP + d.off[7]));
            S2S_TRY(conv3_tc_dgrad(ctx, da, M2, d.Wp1, C2, WpT[3], C2, db_));                               // d a3 (zero outside (H3, W3))
            VGG_LAUNCH(relu_bwd_kernel, M2 * C2, db_, a3, M2 * C2);
            S2S_TRY(conv3_tc_wgrad(ctx, db_, p1, M2, d.Wp1, C2, C1, dWp[2]));
            S2S_TRY(colsum_add(ctx, db_, M2, C2, C2, dP + d.off[5]));
            S2S_TRY(conv3_tc_dgrad(ctx, db_, M2, d.Wp1, C2, WpT[2], C1, da));                               // d p1
            VGG_LAUNCH(pool_bwd_kernel<true>, M1 * C1 / 4, a2, da, nb, d.H1, d.W1, d.H2, d.W2, C1, 1, 2, db_);   // d a2 on the (H1, W1) grid, + ReLU mask
            S2S_TRY(conv3_tc_wgrad(ctx, db_, a1, M1, d.W1, C1, C1, dWp[1]));
            S2S_TRY(colsum_add(ctx, db_, M1, C1, C1, dP + d.off[3]));
            S2S_TRY(conv3_tc_dgrad(ctx, db_, M1, d.W1, C1, WpT[1], C1, da));                                // d a1
        } else {
            // pool2 backward: d pooled -> d a4
            VGG_LAUNCH(pool_bwd_kernel<false>, (int64_t)nb * d.H4 * d.W4 * C2 / 4, a4, gf0, nb, d.H4, d.W4, d.H4, d.W4, C2, 2, 2, da);
            S2S_TRY(conv_relu_bwd(ctx, a3, a4, da, nb, d.H3, d.W3, C2, Wp[3], C2, col, dcol, dWp[3], dP + d.off[7], db_));      // conv4 -> d a3
            S2S_TRY(conv_relu_bwd(ctx, p1, a3, db_, nb, d.H2, d.Wp1, C1, Wp[2], C2, col, dcol, dWp[2], dP + d.off[5], da));      // conv3 -> d p1
            VGG_LAUNCH(pool_bwd_kernel<false>, (int64_t)nb * d.H2 * d.W2 * C1 / 4, a2, da, nb, d.H2, d.W2, d.H2, d.W2, C1, 1, 2, db_);      // pool1 -> d a2
            S2S_TRY(conv_relu_bwd(ctx, a1, a2, db_, nb, d.H1, d.W1, C1, Wp[1], C1, col, dcol, dWp[1], dP + d.off[3], da));       // conv2 -> d a1
        }
        S2S_TRY(conv_relu_bwd(ctx, a0, a1, da, nb, T, F, 3, Wp[0], C1, col, dcol, dWp[0], dP + d.off[1],
                              dX ? dXc + (size_t)b0 * T * F * 3 : nullptr));                                                  // conv1 -> d x
    }
    ar.rewind(mk);
    for (int l = 0; l < 4; l++)
        VGG_LAUNCH_FLAT(wperm_back_add_kernel, (int64_t)cout[l] * cin[l] * 9, dWp[l], cout[l], cin[l], 9, dP + d.off[2 * l]);
    VGG_LAUNCH_FLAT(wperm_back_add_kernel, (int64_t)d.HID * d.view, dW1p, d.HID, C2, d.Wq, dP + d.off[8]);
    if (dX) VGG_LAUNCH_FLAT(nhwc_to_nchw_kernel, (int64_t)B * T * F * 3, dXc, 3, (int64_t)T * F, (int64_t)B * T * F * 3, dX);
    return 0;
}

}  // extern "C"
