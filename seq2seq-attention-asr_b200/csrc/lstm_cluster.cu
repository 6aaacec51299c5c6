// lstm_cluster.cu -- nn.RNN(nn.LSTM(in, out, false)) as PERSISTENT thread-block-cluster kernels (LSTM.lua:25-58 without
// peepholes -- the configuration the reference instantiates, timit/timit.lua:122-123 -- unrolled by RNN.lua:120-201).
//
//     i = sig(a_i)  f = sig(a_f)  g = tanh(a_g)  o = sig(a_o),   a = W_x x + b_x + W_h h_{t-1} + b_h
//     c_t = f c_{t-1} + i g          h_t = o tanh(c_t)
//
// Same plan as the GRU (gru_seq.cu): the x-products and biases are one time-batched projection; the recurrence is
// ONE launch per direction.  A cluster of H/16 CTAs owns a group of BG utterances; CTA c owns hidden units
// [16c, 16c+16) and keeps the 64 rows (4 gates x 16 units) of the recurrent matrix in REGISTERS for all L steps.
// Without peepholes a step needs the full h_{t-1} only once, so there is a single mat-vec phase and a single DSMEM
// exchange per step (the GRU needs two): every CTA broadcasts its 16-unit slice of h_t with st.async stores that
// signal the receivers' mbarriers.  The state buffer is double-buffered (a fast peer may already broadcast h_t
// while a slow one still reads h_{t-1}); c_t never leaves the registers of the lane that owns the (unit, utterance).
// Backward: the gate gradients of a step (4H values per utterance, exchanged unit-major so that a CTA's slice is
// contiguous) are broadcast once per step and every CTA forms dh_{t-1} for its own units with the transposed weights
// in registers; weight gradients and dX stay time-batched GEMMs (lstm_seq.cu).
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "cluster_rnn.cuh"

namespace cg = cooperative_groups;

namespace s2s {

struct LstmClusterParams {
    const float* Wh;      // [4H, H] packed recurrent weights, rows gate-major (i | f | g | o)
    const float* xp;      // [B, Lmax, 4H] input projections + both biases (forward)
    const int* lengths;
    int B, Lmax, reverse;
    float *y, *cseq, *acts;          // [B,Lmax,H], [B,Lmax,H], [B,Lmax,4H]   (forward: written; backward: read)
    const float* dy;                 // [B,Lmax,H]
    float* dA;                       // [B,Lmax,4H] gate pre-activation gradients (gate-major, like xp)
};

constexpr int LC_UC = 16;            // hidden units per CTA
constexpr int LC_NBP = 4;            // utterances per butterfly pass (8 rows per warp x 4 = 32 lanes)

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(256, 1)
lstm_cluster_fwd_kernel(const LstmClusterParams p) {
    constexpr int CS = H / LC_UC, NI = H / 128, NBP = LC_NBP, R = 8;
    constexpr int NH = (BG + NBP - 1) / NBP;
    constexpr int NB0 = BG < NBP ? BG : NBP, NB1 = BG > NBP ? BG - NBP : 1;
    constexpr unsigned TX = BG * H * 4;
    static_assert(BG <= 2 * NBP, "at most two passes");
    __shared__ __align__(16) float hbuf[2][BG][H];
    __shared__ __align__(16) float stage[BG][LC_UC];
    __shared__ uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int grp = blockIdx.x / CS;
    const bool rev = p.reverse != 0;
    const int b0 = grp * BG;

    // warp w: units 2w, 2w+1 of this CTA's slice; row r of the warp = (unit 2w + r/4, gate r%4)
    float4 w[R][NI];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float* row = p.Wh + ((size_t)(r & 3) * H + crank * LC_UC + 2 * warp + (r >> 2)) * H;
#pragma unroll
        for (int i = 0; i < NI; i++) w[r][i] = __ldg(reinterpret_cast<const float4*>(row + lane * 4 + 128 * i));
    }
    for (int i = tid; i < 2 * BG * H; i += 256) (&hbuf[0][0][0])[i] = 0.f;       // zero initial state (LSTM.lua:108-109)
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }

    // finaliser role of this lane after the butterfly: (row lane/4, utterance lane%4)
    const int bbl = lane % NBP;
    const int gate = (lane / NBP) & 3;
    const int ju = 2 * warp + (lane >> 4);                     // unit within the CTA's slice
    const int j = crank * LC_UC + ju;
    int Lf[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) {
        const int bl = NBP * hf + bbl, b = b0 + bl;
        Lf[hf] = (bl < BG && b < p.B) ? (p.lengths ? p.lengths[b] : p.Lmax) : 0;
    }
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t hbuf_a = smem_u32(&hbuf[0][0][0]), bar_a = smem_u32(&bar);
    cluster_sync_all();

    auto load_xp = [&](int s, int hf) -> float {
        if (s >= Lf[hf]) return 0.f;
        const int t = rev ? Lf[hf] - 1 - s : s;
        return __ldg(p.xp + ((size_t)(b0 + NBP * hf + bbl) * p.Lmax + t) * (4 * H) + gate * H + j);
    };
    float xpn[NH], cst[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) { xpn[hf] = load_xp(0, hf); cst[hf] = 0.f; }
    unsigned parity = 0;
    for (int s = 0; s < Lgrp; s++) {
        const int cur = s & 1, nxt = cur ^ 1;
        float xpv[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) { xpv[hf] = xpn[hf]; xpn[hf] = load_xp(s + 1, hf); }
        if (tid == 0) mbar_expect_tx(&bar, TX);

        float tot[NH];
        if (NH == 1) tot[0] = matvec<H, R, NB0, NBP>(w, hbuf[cur], 0, lane);
        else matvec_pair<H, R, NB0, NB1, NBP>(w, hbuf[cur], lane, tot[0], tot[NH - 1]);
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const float pre = tot[hf] + xpv[hf];
            const float act = gate == 2 ? tanh_acc(pre) : sigmoid_acc(pre);             // LSTM.lua:41-43,49
            // the three other gates of this (unit, utterance) sit 4, 8 and 12 lanes up
            const float fg = __shfl_down_sync(0xffffffffu, act, 4), gg = __shfl_down_sync(0xffffffffu, act, 8),
                        og = __shfl_down_sync(0xffffffffu, act, 12);
            const int bl = NBP * hf + bbl;
            if (bl < BG) {
                const bool on = s < Lf[hf];
                const int t = rev ? Lf[hf] - 1 - s : s;
                const size_t row = (size_t)(b0 + bl) * p.Lmax + t;
                if (on) p.acts[row * 4 * H + gate * H + j] = act;
                if (gate == 0) {
                    float hn = hbuf[cur][bl][j];                                        // inactive: state frozen
                    if (on) {
                        const float c = fg * cst[hf] + act * gg;                        // LSTM.lua:45-46
                        hn = og * tanh_acc(c);                                          // LSTM.lua:51
                        cst[hf] = c;
                        p.cseq[row * H + j] = c;
                        p.y[row * H + j] = hn;
                    }
                    stage[bl][ju] = hn;
                }
            }
        }
        __syncthreads();
        bcast_slice<H, LC_UC, BG>(stage, hbuf_a + (uint32_t)(nxt * BG * H) * 4u, bar_a, crank, warp, lane);
        mbar_wait(&bar, parity);
        parity ^= 1;
    }
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(256, 1)
lstm_cluster_bwd_kernel(const LstmClusterParams p) {
    constexpr int CS = H / LC_UC, K4 = 4 * H, NI4 = K4 / 128, NBP = LC_NBP, R = 2, SL = 4 * LC_UC;
    constexpr int NH = (BG + NBP - 1) / NBP;
    constexpr int NB0 = BG < NBP ? BG : NBP, NB1 = BG > NBP ? BG - NBP : 1;
    constexpr unsigned TX = BG * K4 * 4;
    static_assert(BG <= 2 * NBP, "at most two passes");
    __shared__ __align__(16) float dabuf[2][BG][K4];           // gate gradients of the later step, unit-major: [unit][gate]
    __shared__ __align__(16) float stage[BG][SL];
    __shared__ uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int grp = blockIdx.x / CS;
    const bool rev = p.reverse != 0;
    const int b0 = grp * BG;

    // transposed weights: output unit jo = 16 crank + 2 warp + r; input k = 4 u + gate  <->  Wh[gate H + u][jo]
    float4 w[R][NI4];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int jo = crank * LC_UC + 2 * warp + r;
#pragma unroll
        for (int i = 0; i < NI4; i++) {
            const int u = lane + 32 * i;
            w[r][i] = make_float4(__ldg(p.Wh + ((size_t)0 * H + u) * H + jo), __ldg(p.Wh + ((size_t)1 * H + u) * H + jo),
                                  __ldg(p.Wh + ((size_t)2 * H + u) * H + jo), __ldg(p.Wh + ((size_t)3 * H + u) * H + jo));
        }
    }
    for (int i = tid; i < 2 * BG * K4; i += 256) (&dabuf[0][0][0])[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }

    // finaliser role: lanes 0..7 of a warp hold (unit 2 warp + lane/4, utterance lane%4)
    const int bbl = lane % NBP;
    const int ju = 2 * warp + ((lane >> 2) & 1);
    const int j = crank * LC_UC + ju;
    const bool fin = lane < R * NBP;
    int Lf[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) {
        const int bl = NBP * hf + bbl, b = b0 + bl;
        Lf[hf] = (bl < BG && b < p.B) ? (p.lengths ? p.lengths[b] : p.Lmax) : 0;
    }
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t dabuf_a = smem_u32(&dabuf[0][0][0]), bar_a = smem_u32(&bar);
    cluster_sync_all();

    // saved activations of a step: prefetched one step ahead of their use
    struct Sv { float ig, fg, gg, og, c, cp, dy; };
    auto load_sv = [&](int s, int hf) -> Sv {
        Sv v = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!fin || s < 0 || s >= Lf[hf]) return v;
        const int t = rev ? Lf[hf] - 1 - s : s;
        const size_t row = (size_t)(b0 + NBP * hf + bbl) * p.Lmax + t;
        const float* ar = p.acts + row * 4 * H;
        v.ig = ldg_pinned(ar + j); v.fg = ldg_pinned(ar + H + j); v.gg = ldg_pinned(ar + 2 * H + j); v.og = ldg_pinned(ar + 3 * H + j);
        v.c = ldg_pinned(p.cseq + row * H + j);
        v.dy = ldg_pinned(p.dy + row * H + j);
        if (s > 0) v.cp = ldg_pinned(p.cseq + ((size_t)(b0 + NBP * hf + bbl) * p.Lmax + (rev ? t + 1 : t - 1)) * H + j);
        return v;
    };
    Sv svn[NH];
    float dcc[NH];                                             // d c_t carried from the later step
#pragma unroll
    for (int hf = 0; hf < NH; hf++) { svn[hf] = load_sv(Lgrp - 1, hf); dcc[hf] = 0.f; }
    unsigned parity = 0;
    for (int s = Lgrp - 1, it = 0; s >= 0; s--, it++) {                                 // RNN.lua:183
        const int cur = it & 1, nxt = cur ^ 1;
        Sv sv[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) sv[hf] = svn[hf];
        if (tid == 0) mbar_expect_tx(&bar, TX);

        // dh_t (recurrent part) = W_h^T dA_{t+1} for this CTA's units
        float tot[NH];
        if (NH == 1) tot[0] = matvec<K4, R, NB0, NBP>(w, dabuf[cur], 0, lane);
        else matvec_pair<K4, R, NB0, NB1, NBP>(w, dabuf[cur], lane, tot[0], tot[NH - 1]);
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = NBP * hf + bbl;
            if (fin && bl < BG) {
                float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s < Lf[hf]) {
                    const Sv& v = sv[hf];
                    const int t = rev ? Lf[hf] - 1 - s : s;
                    const float dh = v.dy + tot[hf];
                    const float tc = tanh_acc(v.c);
                    const float dc = dcc[hf] + dh * v.og * (1.f - tc * tc);
                    da.x = dc * v.gg * v.ig * (1.f - v.ig);                             // da_i
                    da.y = dc * v.cp * v.fg * (1.f - v.fg);                             // da_f
                    da.z = dc * v.ig * (1.f - v.gg * v.gg);                             // da_g
                    da.w = dh * tc * v.og * (1.f - v.og);                               // da_o
                    dcc[hf] = dc * v.fg;
                    float* dr = p.dA + ((size_t)(b0 + bl) * p.Lmax + t) * 4 * H;
                    dr[j] = da.x; dr[H + j] = da.y; dr[2 * H + j] = da.z; dr[3 * H + j] = da.w;
                }
                *reinterpret_cast<float4*>(&stage[bl][4 * ju]) = da;
            }
        }
        // the next step's operands, issued only now that this step's have been consumed: issued at the top of the loop, these loads -- the
        // same static instructions, hence the same scoreboard -- made the gate math above wait for their full round trip (gru_seq3.cu)
#pragma unroll
        for (int hf = 0; hf < NH; hf++) svn[hf] = load_sv(s - 1, hf);
        __syncthreads();
        bcast_slice<K4, SL, BG>(stage, dabuf_a + (uint32_t)(nxt * BG * K4) * 4u, bar_a, crank, warp, lane);
        mbar_wait(&bar, parity);
        parity ^= 1;
    }
    cluster_sync_all();
}

// previous state in processing order (the operands of the time-batched weight gradients): hprev[b,t] = y[b,t'] etc.
__global__ void lstm_shift_kernel(const float* __restrict__ y, const float* __restrict__ cseq, const int* __restrict__ lengths, int B, int Lmax,
                                  int H, int reverse, float* __restrict__ hprev, float* __restrict__ cprev) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * Lmax * H) return;
    const int jj = (int)(idx % H);
    const int64_t row = idx / H;
    const int t = (int)(row % Lmax), b = (int)(row / Lmax);
    const int Lb = lengths ? lengths[b] : Lmax;
    float hp = 0.f, cp = 0.f;
    if (t < Lb) {
        const int tp = reverse ? t + 1 : t - 1;
        if (tp >= 0 && tp < Lb) { hp = y[((size_t)b * Lmax + tp) * H + jj]; cp = cseq[((size_t)b * Lmax + tp) * H + jj]; }
    }
    hprev[idx] = hp; cprev[idx] = cp;
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
template <int H, int BG>
static int lc_launch(s2s_ctx* ctx, bool backward, const LstmClusterParams& p, int* max_clusters) {
    constexpr int CS = H / LC_UC;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ceil_div(p.B, BG));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (CS > 8) {   // clusters of 16 are "non-portable": opt in once per kernel
        static bool set[2] = {false, false};
        if (!set[backward]) {
            if (backward) S2S_CUDA(cudaFuncSetAttribute(lstm_cluster_bwd_kernel<H, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            else S2S_CUDA(cudaFuncSetAttribute(lstm_cluster_fwd_kernel<H, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            set[backward] = true;
        }
    }
    if (max_clusters) {
        cudaError_t e = backward ? cudaOccupancyMaxActiveClusters(max_clusters, lstm_cluster_bwd_kernel<H, BG>, &cfg)
                                 : cudaOccupancyMaxActiveClusters(max_clusters, lstm_cluster_fwd_kernel<H, BG>, &cfg);
        if (e != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    prof_begin(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD);      // reported with the recurrence classes
    if (backward) S2S_CUDA(cudaLaunchKernelEx(&cfg, lstm_cluster_bwd_kernel<H, BG>, p));
    else S2S_CUDA(cudaLaunchKernelEx(&cfg, lstm_cluster_fwd_kernel<H, BG>, p));
    ctx->kcount[S2S_KC_LSTM_CLUSTER]++;
    // algorithmic bytes: fwd reads xp (4H), writes y, c (2H) and the gates (4H); bwd reads gates, c, c_prev, dy (7H), writes dA (4H)
    prof_end(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD, 4.0 * p.B * p.Lmax * (backward ? 11.0 : 10.0) * H);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

template <int H>
static int lc_dispatch(s2s_ctx* ctx, bool backward, const LstmClusterParams& p, bool* handled) {
    static int cap[2] = {0, 0};
    if (cap[backward] == 0) {
        int n = 0;
        S2S_TRY((lc_launch<H, 4>(ctx, backward, p, &n)));
        cap[backward] = n > 0 ? n : -1;
        if (getenv("S2S_LSTM_DBG")) fprintf(stderr, "[lstm_cluster] H=%d %s: %d co-resident clusters of %d CTAs\n", H, backward ? "bwd" : "fwd", n, H / LC_UC);
    }
    *handled = false;
    if (cap[backward] < 0) return 0;
    // one wave: the smallest group size whose cluster count is co-resident
    const int sizes[7] = {1, 2, 3, 4, 5, 6, 8};
    const int nsizes = H == 128 ? 7 : 5;      // H = 256: larger groups would exceed the static shared-memory limit of the backward kernel
    int bg = 0;
    for (int i = 0; i < nsizes; i++)
        if (ceil_div(p.B, sizes[i]) <= cap[backward]) { bg = sizes[i]; break; }
    if (!bg) return 0;
    *handled = true;
    switch (bg) {
        case 1: return lc_launch<H, 1>(ctx, backward, p, nullptr);
        case 2: return lc_launch<H, 2>(ctx, backward, p, nullptr);
        case 3: return lc_launch<H, 3>(ctx, backward, p, nullptr);
        case 4: return lc_launch<H, 4>(ctx, backward, p, nullptr);
        case 5: return lc_launch<H, 5>(ctx, backward, p, nullptr);
        default:
            if constexpr (H == 128) {
                if (bg == 6) return lc_launch<H, 6>(ctx, backward, p, nullptr);
                return lc_launch<H, 8>(ctx, backward, p, nullptr);
            } else {
                return fail("lstm_cluster: unsupported group size");
            }
    }
}

// Runs the recurrence of lstm_seq_forward / _backward when the shape is supported (no peepholes, H in {128, 256});
// *handled = false otherwise (the caller falls back to its per-frame path).
int lstm_cluster_forward(s2s_ctx* ctx, const float* Whp, const float* xp, const int* lengths, int B, int Lmax, int H, int reverse,
                         float* y, float* cseq, float* acts, bool* handled) {
    *handled = false;
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("S2S_LSTM_CLUSTER"); enabled = e ? atoi(e) : 1; }
    if (!enabled) return 0;
    LstmClusterParams p = {};
    p.Wh = Whp; p.xp = xp; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.reverse = reverse; p.y = y; p.cseq = cseq; p.acts = acts;
    if (H == 128) return lc_dispatch<128>(ctx, false, p, handled);
    if (H == 256) return lc_dispatch<256>(ctx, false, p, handled);
    return 0;
}
int lstm_cluster_backward(s2s_ctx* ctx, const float* Whp, const int* lengths, int B, int Lmax, int H, int reverse, const float* y,
                          const float* cseq, const float* acts, const float* dy, float* dA, float* hprev, float* cprev, bool* handled) {
    *handled = false;
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("S2S_LSTM_CLUSTER"); enabled = e ? atoi(e) : 1; }
    if (!enabled) return 0;
    LstmClusterParams p = {};
    p.Wh = Whp; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.reverse = reverse;
    p.y = const_cast<float*>(y); p.cseq = const_cast<float*>(cseq); p.acts = const_cast<float*>(acts); p.dy = dy; p.dA = dA;
    if (H == 128) S2S_TRY(lc_dispatch<128>(ctx, true, p, handled));
    else if (H == 256) S2S_TRY(lc_dispatch<256>(ctx, true, p, handled));
    if (!*handled) return 0;
    const int64_t n = (int64_t)B * Lmax * H;
    lstm_shift_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(y, cseq, lengths, B, Lmax, H, reverse, hprev, cprev);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s
