// decoder_cluster.cu -- the teacher-forced decoder TIME LOOP as one persistent thread-block-cluster kernel.
//
// Per decoder step the reference runs (Attention.lua:95-151, RNNAttention.lua:144-190, GRU.lua:22-30)
//     e_l = w . tanh(q_t + Vh_l)      alpha = softmax_l(e)      c_t = sum_l alpha_l h_l
//     u_t = (W_j W_c) c_t + uy_t      {z, r} = sigmoid(G_zr {s_{t-1}, u_t})      h~ = tanh(G_h {r s_{t-1}, u_t})
//     s_t = (1-z) s_{t-1} + z h~      q_{t+1} = W_s s_t + b_s
// which decoder.cu issues as one attention launch + four dense launches per step: five dependent kernels whose
// fixed latencies (~40 us per step) are the largest single item of the training step.  Here a cluster of 16 CTAs owns
// BG utterances for ALL T steps:
//   * the step's weights (W_jc, G_zr, G_h: 512 KB + 1 MB + 512 KB, W_s: 512 KB) stay on chip for the whole call -- every CTA
//     keeps the rows of its 16 decoder units (32 rows of q) in shared memory (128 KB, k-major so lanes = rows read
//     conflict-free) and registers (W_s);
//   * the encoder frames of an utterance are split over the 16 CTAs (flash-decoding style): each CTA streams its rows of
//     Vh and h from L2 once per step, forms local softmax statistics and a partial context, and the partials are
//     reduce-scattered over distributed shared memory (st.async + mbarrier complete_tx) so that CTA j normalises the
//     context slice j and all-gathers it;
//   * the mat-vec phases run with lanes = rows, warps = K-slices (x read as warp-uniform broadcasts), partial sums are
//     combined through shared memory, the gate math is fused, and each phase's slice is all-gathered with one DSMEM
//     exchange (6 exchanges per step, ~0.5 us each, instead of 5 kernel boundaries).
// Everything the backward pass reads (alpha, {s,c}, q, {s,u}, {r s,u}, gates, penalty) is written exactly as the
// per-step path writes it, so decoder_backward is unchanged.
// Supported: ST = 256, S = A = 512 (the Chorowski / VGG model sizes), Lmax <= 1024; content attention or the location-aware term with a
// filter of <= 10 taps (LOC template flag), with or without the monotonicity penalty -- forward and backward.
#include <cooperative_groups.h>

#include "cluster_rnn.cuh"
#include "decoder.cuh"

namespace cg = cooperative_groups;

namespace s2s {

constexpr int DC_CS = 16;          // CTAs per cluster
constexpr int DC_THREADS = 512;
constexpr int DC_ST = 256, DC_A = 512, DC_S = 512;
constexpr int DC_RMAX = 64;        // encoder frames of one utterance per CTA: Lmax <= 16 * 64

struct DecClusterParams {
    const float *Vh, *h, *w, *qbias, *Ws, *Wjc, *Gz, *Gh, *uy;
    const int *lengths, *tlens;
    int B, Lmax, T;
    float lambda;
    float *alpha, *sc, *q, *pen, *su, *rhu, *gates;
    long long* clk;                // optional per-phase clock accumulators (S2S_DEC_PROF)
    const float* uw;               // location term (LOC): folded weights U W_F [KF][S], filter size, left padding
    int KF, padl;
};
constexpr int DC_KFMAX = 10;       // filter taps of the location term the cluster kernel holds in shared memory

template <int BG, int LOC = 0>
struct DcSmem {
    // weights, k-major: Wt[(k4 * ROWS + r) * 4 + kk] = W[row0 + r][4 k4 + kk]
    float wjc[128 * 16 * 4];
    float gz[128 * 32 * 4];
    float gh[128 * 16 * 4];
    // all-gathered vectors
    alignas(16) float q_full[BG][DC_S];        // holds 2 log2(e) q (see dc_tanh4_dot)
    float c_full[BG][DC_A];
    float s_full[BG][DC_ST];
    float u_full[BG][DC_ST];
    float rs_full[BG][DC_ST];
    // reduce-scatter receive: slice of every CTA's partial context + its softmax statistics {m, s, w1, w2}
    float recv_c[DC_CS][BG][32];
    float4 recv_st[DC_CS][BG];
    float part[16][BG][32];        // per-warp (K-slice) partial sums of a mat-vec phase
    alignas(16) float e_s[BG][DC_RMAX], p_s[BG][DC_RMAX + 16], ap_s[BG][DC_RMAX];     // p_s: zero past the slice (the context loop reads blocks of 20)
    alignas(16) float w_s[DC_S];
    alignas(16) float stage[BG][32];
    alignas(16) float zbuf[BG][16];
    float scl_own[BG];
    int l0_s[BG], nr_s[BG], len_s[BG];
    int frow[BG * DC_RMAX];        // this CTA's frames, flattened over its utterances: row of Vh / h ((b0+b) Lmax + l)
    short fb[BG * DC_RMAX], fr[BG * DC_RMAX];      // utterance and frame-within-slice of a flattened frame
    uint64_t bar[6];               // q, cp, c, u, rs, s
    // location-aware term (Attention.lua:75-99, folded): 2 log2(e) U W_F, and alpha_{t-1} over this CTA's frames plus the filter halo
    alignas(16) float uw_s[LOC ? DC_KFMAX * DC_S : 4];
    alignas(16) unsigned long long aph_s[LOC ? BG : 1][DC_RMAX + 16];    // {a, a} pairs (packed FFMA2 operands)
};
enum { BAR_Q = 0, BAR_CP, BAR_C, BAR_U, BAR_RS, BAR_S };

// all-gather: this CTA's [BG][UC] slice (staged in shared memory, row pitch 32) -> columns [UC*crank, +UC) of `buf` ([BG][H]) of
// every CTA of the cluster; warp w serves rank w
template <int H, int UC, int BG>
__device__ __forceinline__ void dc_bcast(const float (*stage)[32], uint32_t buf_a, uint32_t bar_a, unsigned crank, int warp, int lane) {
    constexpr int CPB = UC / 4;
    const uint32_t rbar = mapa_rank(bar_a, warp);
    for (int ch = lane; ch < BG * CPB; ch += 32) {
        const int b = ch / CPB, off = (ch % CPB) * 4;
        const float4 v = *reinterpret_cast<const float4*>(&stage[b][off]);
        st_async_v4(mapa_rank(buf_a + (uint32_t)(b * H + crank * UC + off) * 4u, warp), v, rbar);
    }
}

// one K-slice (32 k per warp) of a mat-vec phase.  ROWS = 32: lane = row, 8 k4 per lane; ROWS = 16: lane = (row, half), 4 k4 per lane.
// Wt is k-major ([k4][ROWS][4]); x rows are read as warp-uniform broadcasts; the warp's partial sums go to part[warp][b][row].
template <int BG, int ROWS>
__device__ __forceinline__ void dc_mv(const float* __restrict__ Wt, const float* __restrict__ x, int xpitch, int k4w, int k4x, int warp, int lane,
                                      float (*part)[BG][32]) {
    constexpr int NK = ROWS == 32 ? 8 : 4;
    const int r = ROWS == 32 ? lane : (lane & 15), sub = ROWS == 32 ? 0 : (lane >> 4) * 4;
    float acc[BG];
#pragma unroll
    for (int b = 0; b < BG; b++) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < NK; i++) {
        const float4 wv = *reinterpret_cast<const float4*>(Wt + ((size_t)(k4w + sub + i) * ROWS + r) * 4);
#pragma unroll
        for (int b = 0; b < BG; b++) acc[b] = dot4(wv, *reinterpret_cast<const float4*>(x + (size_t)b * xpitch + (k4x + sub + i) * 4), acc[b]);
    }
    if (ROWS == 16) {
#pragma unroll
        for (int b = 0; b < BG; b++) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 16);
    }
    if (ROWS == 32 || lane < 16) {
#pragma unroll
        for (int b = 0; b < BG; b++) part[warp][b][r] = acc[b];
    }
}

// acc + w . tanh(x) for four elements with ONE reciprocal, x given as y = 2 log2(e) x (the factor is folded into the staged q and
// one FFMA per element): tanh(x) = 1 - 2 / (2^y + 1); the four denominators are inverted together
// (1 / d_i = (product of the others) / (d_0 d_1 d_2 d_3)), 5 MUFU operations per 4 elements instead of 8, raw ex2.approx / rcp.approx
// (no range fix-up code: y is clamped to 10 k, tanh(10) rounds to 1.0f, so the product stays below 6e34 and nothing is denormal).
constexpr float DC_K = 2.885390081777927f;                   // 2 log2(e)
typedef unsigned long long dc_f2;                            // packed fp32 pair (fma.rn.f32x2 operands)
__device__ __forceinline__ dc_f2 dc_pack2(float a, float b) { dc_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void dc_unpack2(dc_f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ dc_f2 dc_ffma2(dc_f2 a, dc_f2 b, dc_f2 c) { dc_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float dc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float dc_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float dc_tanh4_dot(const float4 w, const float4 v, const float4 qk, float acc) {
    const float d0 = dc_ex2(fminf(fmaf(DC_K, v.x, qk.x), 10.f * DC_K)) + 1.f, d1 = dc_ex2(fminf(fmaf(DC_K, v.y, qk.y), 10.f * DC_K)) + 1.f;
    const float d2 = dc_ex2(fminf(fmaf(DC_K, v.z, qk.z), 10.f * DC_K)) + 1.f, d3 = dc_ex2(fminf(fmaf(DC_K, v.w, qk.w), 10.f * DC_K)) + 1.f;
    const float p01 = d0 * d1, p23 = d2 * d3;
    const float r = -2.0f * dc_rcp(p01 * p23);
    const float r01 = r * p23, r23 = r * p01;                // -2 / (d0 d1), -2 / (d2 d3)
    acc = fmaf(w.x, fmaf(r01, d1, 1.f), acc);
    acc = fmaf(w.y, fmaf(r01, d0, 1.f), acc);
    acc = fmaf(w.z, fmaf(r23, d3, 1.f), acc);
    return fmaf(w.w, fmaf(r23, d2, 1.f), acc);
}

template <int BG, int LOC>
__global__ void __launch_bounds__(DC_THREADS, 1)
dec_cluster_fwd_kernel(const DecClusterParams p) {
    extern __shared__ __align__(128) unsigned char dc_smem_raw[];
    DcSmem<BG, LOC>& sm = *reinterpret_cast<DcSmem<BG, LOC>*>(dc_smem_raw);
    constexpr int ST = DC_ST, A = DC_A, S = DC_S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int b0 = (blockIdx.x / DC_CS) * BG;
    const int T = p.T, Lmax = p.Lmax;

    // ---- one-time staging -------------------------------------------------------------------------------------------
    {   // weights: global rows are k-contiguous (coalesced scalar reads: segments of the flat parameter vector are not
        // 16-byte aligned in general), shared layout is k-major
        for (int i = tid; i < 16 * 512; i += DC_THREADS) {
            const int r = i >> 9, k = i & 511;
            sm.wjc[((size_t)(k >> 2) * 16 + r) * 4 + (k & 3)] = p.Wjc[(size_t)(16 * crank + r) * A + k];
            sm.gh[((size_t)(k >> 2) * 16 + r) * 4 + (k & 3)] = p.Gh[(size_t)(16 * crank + r) * 2 * ST + k];
        }
        for (int i = tid; i < 32 * 512; i += DC_THREADS) {
            const int r = i >> 9, k = i & 511;
            const int n = r < 16 ? 16 * crank + r : ST + 16 * crank + r - 16;        // z rows, then r rows of this CTA's units
            sm.gz[((size_t)(k >> 2) * 32 + r) * 4 + (k & 3)] = p.Gz[(size_t)n * 2 * ST + k];
        }
        for (int i = tid; i < S; i += DC_THREADS) sm.w_s[i] = p.w[i];
        if constexpr (LOC != 0) {
            for (int i = tid; i < DC_KFMAX * S; i += DC_THREADS) sm.uw_s[i] = i < p.KF * S ? DC_K * p.uw[i] : 0.f;     // taps past KF: zero weights
            for (int i = tid; i < BG * (DC_RMAX + 16); i += DC_THREADS) (&sm.aph_s[0][0])[i] = 0ull;                     // alpha_{-1} = 0
        }
        for (int i = tid; i < BG * S; i += DC_THREADS) (&sm.q_full[0][0])[i] = DC_K * p.qbias[i % S];      // q_0 = W_s 0 + b_s
        for (int i = tid; i < BG * ST; i += DC_THREADS) (&sm.s_full[0][0])[i] = 0.f;               // s_0 = 0 (Recurrent.lua:112)
        for (int i = tid; i < BG * DC_RMAX; i += DC_THREADS) (&sm.ap_s[0][0])[i] = 0.f;            // alpha_{-1} = 0
        for (int i = tid; i < BG * (DC_RMAX + 16); i += DC_THREADS) (&sm.p_s[0][0])[i] = 0.f;
        if (tid < BG) {
            const int b = b0 + tid;
            const int Lb = b < p.B ? (p.lengths ? p.lengths[b] : Lmax) : 0;
            const int Rb = (Lb + DC_CS - 1) / DC_CS;
            const int l0 = (int)crank * Rb;
            sm.len_s[tid] = Lb; sm.l0_s[tid] = l0; sm.nr_s[tid] = max(0, min(Rb, Lb - l0));
        }
        if (tid == 0) {
            for (int i = 0; i < 6; i++) mbar_init(&sm.bar[i], 1);
            fence_mbar_init();
        }
    }
    // W_s rows of this CTA's q slice: lane = row, warp = K-slice of 16
    float4 wq[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float* wr = p.Ws + (size_t)(32 * crank + lane) * ST + 16 * warp + 4 * i;
        wq[i] = make_float4(__ldg(wr), __ldg(wr + 1), __ldg(wr + 2), __ldg(wr + 3));
    }
    __syncthreads();
    if (tid < BG * 32) {   // q_0 slice of the saved state
        const int b = tid >> 5, k = tid & 31;
        if (b0 + b < p.B) p.q[((size_t)(b0 + b) * T) * S + 32 * crank + k] = p.qbias[32 * crank + k];
    }
    int NR = 0;
#pragma unroll
    for (int b = 0; b < BG; b++) NR += sm.nr_s[b];
    for (int f = tid; f < NR; f += DC_THREADS) {
        int r = f, b = 0;
#pragma unroll
        for (int bb = 0; bb < BG - 1; bb++)
            if (b == bb && r >= sm.nr_s[bb]) { r -= sm.nr_s[bb]; b = bb + 1; }
        sm.frow[f] = (b0 + b) * Lmax + sm.l0_s[b] + r; sm.fb[f] = (short)b; sm.fr[f] = (short)r;
    }

    const uint32_t q_a = smem_u32(&sm.q_full[0][0]), c_a = smem_u32(&sm.c_full[0][0]), s_a = smem_u32(&sm.s_full[0][0]);
    const uint32_t u_a = smem_u32(&sm.u_full[0][0]), rs_a = smem_u32(&sm.rs_full[0][0]);
    const uint32_t rc_a = smem_u32(&sm.recv_c[0][0][0]), rst_a = smem_u32(&sm.recv_st[0][0]);
    uint32_t bar_a[6];
#pragma unroll
    for (int i = 0; i < 6; i++) bar_a[i] = smem_u32(&sm.bar[i]);
    cluster_sync_all();   // every CTA of the cluster is resident and has initialised its barriers / buffers

    constexpr unsigned TX_CP = DC_CS * BG * (128 + 16), TX_512 = BG * 512 * 4, TX_256 = BG * 256 * 4;
    long long tck = 0;
    const bool prof = p.clk != nullptr && blockIdx.x == 0 && tid == 0;
#define DC_TICK(i) do { if (prof) { const long long n_ = clock64(); p.clk[i] += n_ - tck; tck = n_; } } while (0)
    if (prof) tck = clock64();

    float4 va[2][4], vb[2][4];
    auto load_pair = [&](int f0, float4 (&v)[2][4]) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int f = f0 + 16 * j < NR ? f0 + 16 * j : f0;
            const float* vp = p.Vh + (size_t)sm.frow[f] * S + lane * 4;
#pragma unroll
            for (int i = 0; i < 4; i++) v[j][i] = ldg_stream(vp + 128 * i);
        }
    };
    auto score_pair = [&](int f0, const float4 (&v)[2][4]) {
        if constexpr (LOC != 0) {
            // both frames of the pass together: every U W_F load is shared by the two frames, alpha_{t-1} comes as broadcast {a, a}
            // pairs, and the taps are packed FFMA2 (Attention.lua:86-99: + U W_F alpha_{t-1}[l + j - pad_left])
            const int fA = f0, fB = f0 + 16 < NR ? f0 + 16 : f0;
            const int bA = sm.fb[fA], bB = sm.fb[fB];
            const dc_f2* apA = &sm.aph_s[bA][sm.fr[fA]];
            const dc_f2* apB = &sm.aph_s[bB][sm.fr[fB]];
            float accA = 0.f, accB = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 qA = *reinterpret_cast<const float4*>(&sm.q_full[bA][lane * 4 + 128 * i]);
                const float4 qB = *reinterpret_cast<const float4*>(&sm.q_full[bB][lane * 4 + 128 * i]);
                dc_f2 zA0 = dc_pack2(qA.x, qA.y), zA1 = dc_pack2(qA.z, qA.w), zB0 = dc_pack2(qB.x, qB.y), zB1 = dc_pack2(qB.z, qB.w);
#pragma unroll 5
                for (int jj = 0; jj < DC_KFMAX; jj++) {
                    const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(&sm.uw_s[jj * S + lane * 4 + 128 * i]);
                    const dc_f2 aA = apA[jj], aB = apB[jj];
                    zA0 = dc_ffma2(aA, u.x, zA0); zA1 = dc_ffma2(aA, u.y, zA1);
                    zB0 = dc_ffma2(aB, u.x, zB0); zB1 = dc_ffma2(aB, u.y, zB1);
                }
                float4 zA, zB;
                dc_unpack2(zA0, zA.x, zA.y); dc_unpack2(zA1, zA.z, zA.w); dc_unpack2(zB0, zB.x, zB.y); dc_unpack2(zB1, zB.z, zB.w);
                const float4 wv = *reinterpret_cast<const float4*>(&sm.w_s[lane * 4 + 128 * i]);
                accA = dc_tanh4_dot(wv, v[0][i], zA, accA);
                accB = dc_tanh4_dot(wv, v[1][i], zB, accB);
            }
            accA = warp_sum(accA); accB = warp_sum(accB);
            if (lane == 0) { sm.e_s[bA][sm.fr[fA]] = accA; sm.e_s[bB][sm.fr[fB]] = accB; }   // (a duplicated tail frame rewrites the same value)
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int f = f0 + 16 * j < NR ? f0 + 16 * j : f0;
                const int b = sm.fb[f];
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 qv = *reinterpret_cast<const float4*>(&sm.q_full[b][lane * 4 + 128 * i]);
                    const float4 wv = *reinterpret_cast<const float4*>(&sm.w_s[lane * 4 + 128 * i]);
                    acc = dc_tanh4_dot(wv, v[j][i], qv, acc);
                }
                acc = warp_sum(acc);
                if (lane == 0) sm.e_s[b][sm.fr[f]] = acc;        // (a duplicated tail frame rewrites the same value)
            }
        }
    };
    __syncthreads();                                         // frame table
    if (warp < NR) load_pair(warp, va);

    unsigned parity = 0;
    for (int t = 0; t < T; t++) {
        if (tid == 0) {
            mbar_expect_tx(&sm.bar[BAR_CP], TX_CP); mbar_expect_tx(&sm.bar[BAR_C], TX_512);
            mbar_expect_tx(&sm.bar[BAR_U], TX_256); mbar_expect_tx(&sm.bar[BAR_RS], TX_256); mbar_expect_tx(&sm.bar[BAR_S], TX_256);
            if (t + 1 < T) mbar_expect_tx(&sm.bar[BAR_Q], TX_512);
        }
        if constexpr (LOC != 0) __syncthreads();                 // alpha_{t-1} window staged at the end of the previous step
        // uy_t of the u rows this thread finalises (fetched early)
        float uyv = 0.f;
        if (tid < BG * 16 && b0 + (tid >> 4) < p.B) uyv = __ldg(p.uy + ((size_t)(b0 + (tid >> 4)) * T + t) * ST + 16 * crank + (tid & 15));

        // ---- scoring: warp per encoder frame; two frames per pass, the next pass's rows already in flight (the first pass
        // of a step was issued before the wait for q_t) (Attention.lua:95-121) --------------------------------------------
        {
            int f0 = warp;
            while (f0 < NR) {
                if (f0 + 32 < NR) load_pair(f0 + 32, vb);
                score_pair(f0, va);
                f0 += 32;
                if (f0 >= NR) break;
                if (f0 + 32 < NR) load_pair(f0 + 32, va);
                score_pair(f0, vb);
                f0 += 32;
            }
        }
        __syncthreads();
        DC_TICK(0);

        // ---- partial context: thread = (float4 column c4, row group g), warp w holds the 8 columns of CTA w's slice x 4 row groups;
        // blocks of 20 frames of two utterances (2 x 5 loads per thread) are in flight at once, the first block is issued before the
        // statistics.  Row indices are clamped instead of predicated: p_s is zero past the slice. ---------------------------------
        const int c4 = 8 * warp + (lane & 7), g = lane >> 3;     // a quarter-warp reads 128 contiguous bytes of one frame
        float4 hx[2][5];
        auto ctx_load = [&](int bp, int base) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (bp + j >= BG) continue;
                const int b = bp + j, nr = sm.nr_s[b];
                if (nr <= base) {                                    // (uniform) nothing left of this utterance here
#pragma unroll
                    for (int u = 0; u < 5; u++) hx[j][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    continue;
                }
                const float* hp = p.h + ((size_t)(b0 + b) * Lmax + sm.l0_s[b]) * A + c4 * 4;
#pragma unroll
                for (int u = 0; u < 5; u++) hx[j][u] = ldg_stream(hp + (size_t)min(base + g + 4 * u, nr - 1) * A);
            }
        };
        ctx_load(0, 0);

        // ---- local softmax statistics: warp b -> utterance b --------------------------------------------------------------
        if (warp < BG) {
            const int b = warp, nr = sm.nr_s[b], Lb = sm.len_s[b], l0 = sm.l0_s[b];
            const float e0 = lane < nr ? sm.e_s[b][lane] : -INFINITY, e1 = lane + 32 < nr ? sm.e_s[b][lane + 32] : -INFINITY;
            const float m = warp_max(fmaxf(e0, e1));
            const float p0 = lane < nr ? expf(e0 - m) : 0.f, p1 = lane + 32 < nr ? expf(e1 - m) : 0.f;
            sm.p_s[b][lane] = p0; sm.p_s[b][lane + 32] = p1;
            const float ssum = warp_sum(p0 + p1);
            // penalty terms (Attention.lua:123-135): sum_l (L_b - l) alpha_l and sum_l (L_b - l) alpha_{t-1,l}
            const float k0 = (float)(Lb - l0 - lane), k1 = (float)(Lb - l0 - lane - 32);
            const float w1 = warp_sum(k0 * p0 + k1 * p1);
            const float w2 = warp_sum((lane < nr ? k0 * sm.ap_s[b][lane] : 0.f) + (lane + 32 < nr ? k1 * sm.ap_s[b][lane + 32] : 0.f));
            if (lane < DC_CS)   // statistics to every CTA of the cluster
                st_async_v4(mapa_rank(rst_a + (uint32_t)(crank * BG + b) * 16u, lane), make_float4(m, ssum, w1, w2), mapa_rank(bar_a[BAR_CP], lane));
        }
        __syncthreads();
        DC_TICK(8);
        {
            const int dst = warp;                                    // owner of this warp's columns
            const uint32_t dbar = mapa_rank(bar_a[BAR_CP], dst);
#pragma unroll
            for (int bp = 0; bp < BG; bp += 2) {
                float4 acc[2];
                acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int nrm = max(sm.nr_s[bp], bp + 1 < BG ? sm.nr_s[bp + 1] : 0);
                for (int base = 0; base < nrm || base == 0; base += 20) {
                    if (base > 0) ctx_load(bp, base);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        if (bp + j < BG) {
#pragma unroll
                            for (int u = 0; u < 5; u++) {
                                const float pv = sm.p_s[bp + j][base + g + 4 * u];
                                acc[j].x = fmaf(pv, hx[j][u].x, acc[j].x); acc[j].y = fmaf(pv, hx[j][u].y, acc[j].y);
                                acc[j].z = fmaf(pv, hx[j][u].z, acc[j].z); acc[j].w = fmaf(pv, hx[j][u].w, acc[j].w);
                            }
                        }
                    }
                }
                if (bp + 2 < BG) ctx_load(bp + 2, 0);                // next pair in flight during the reduction of this one
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (bp + j < BG) {
                        const int b = bp + j;
#pragma unroll
                        for (int o = 8; o <= 16; o <<= 1) {
                            acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, o); acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, o);
                            acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, o); acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, o);
                        }
                        if (g == 0) st_async_v4(mapa_rank(rc_a + (uint32_t)((crank * BG + b) * 32 + (c4 & 7) * 4) * 4u, dst), acc[j], dbar);
                    }
                }
            }
        }
        DC_TICK(1);
        mbar_wait(&sm.bar[BAR_CP], parity);
        DC_TICK(2);

        // ---- combine: this CTA normalises context columns [32 crank, +32) of every utterance ----------------------------------
        if (tid < BG * 32) {
            const int b = tid >> 5, k = tid & 31;
            float M = -INFINITY;
#pragma unroll
            for (int i = 0; i < DC_CS; i++) M = fmaxf(M, sm.recv_st[i][b].x);
            float den = 0.f, cv = 0.f, w1 = 0.f, w2 = 0.f, own = 0.f;
#pragma unroll
            for (int i = 0; i < DC_CS; i++) {
                const float4 st = sm.recv_st[i][b];
                const float sc = M > -INFINITY ? expf(st.x - M) : 0.f;
                den = fmaf(st.y, sc, den);
                cv = fmaf(sc, sm.recv_c[i][b][k], cv);
                w1 = fmaf(sc, st.z, w1); w2 += st.w;
                if (i == (int)crank) own = sc;
            }
            const float inv = den > 0.f ? 1.0f / den : 0.f;
            cv *= inv;
            sm.stage[b][k] = cv;
            if (k == 0) sm.scl_own[b] = own * inv;
            if (b0 + b < p.B) {
                p.sc[((size_t)(b0 + b) * T + t) * (ST + A) + ST + 32 * crank + k] = cv;
                if (k == 0 && crank == 0) {
                    const bool padded = p.tlens && t >= p.tlens[b0 + b];
                    p.pen[(size_t)(b0 + b) * T + t] = padded ? 0.f : p.lambda * fmaxf(w1 * inv - w2, 0.f);
                }
            }
        }
        __syncthreads();
        dc_bcast<A, 32, BG>(sm.stage, c_a, bar_a[BAR_C], crank, warp, lane);
        // alpha_t of this CTA's frames (the broadcast is in flight meanwhile)
        for (int i = tid; i < BG * DC_RMAX; i += DC_THREADS) {
            const int b = i / DC_RMAX, r = i % DC_RMAX;
            if (r < sm.nr_s[b]) {
                const float a = sm.p_s[b][r] * sm.scl_own[b];
                sm.ap_s[b][r] = a;
                p.alpha[((size_t)(b0 + b) * T + t) * Lmax + sm.l0_s[b] + r] = a;
            }
        }
        if constexpr (LOC != 0) __threadfence();                 // the neighbours read this slice's alpha_t (filter halo) through L2
        mbar_wait(&sm.bar[BAR_C], parity);
        DC_TICK(3);

        // ---- u_t = W_jc c_t + uy_t   (Attention.lua:150-151, folded) --------------------------------------------------------
        dc_mv<BG, 16>(sm.wjc, &sm.c_full[0][0], A, 8 * warp, 8 * warp, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 16) {
            const int b = tid >> 4, r = tid & 15;
            float v = uyv;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][r];
            sm.stage[b][r] = v;
            if (b0 + b < p.B) {
                const size_t row = (size_t)(b0 + b) * T + t;
                p.su[row * 2 * ST + ST + 16 * crank + r] = v;
                p.rhu[row * 2 * ST + ST + 16 * crank + r] = v;
            }
        }
        __syncthreads();
        dc_bcast<ST, 16, BG>(sm.stage, u_a, bar_a[BAR_U], crank, warp, lane);
        mbar_wait(&sm.bar[BAR_U], parity);
        DC_TICK(4);

        // ---- z, r = sigmoid(G_zr {s_{t-1}, u_t}) ; r * s_{t-1}   (GRU.lua:22-25) -----------------------------------------------
        if (warp < 8) dc_mv<BG, 32>(sm.gz, &sm.s_full[0][0], ST, 8 * warp, 8 * warp, warp, lane, sm.part);
        else dc_mv<BG, 32>(sm.gz, &sm.u_full[0][0], ST, 8 * warp, 8 * warp - 64, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 32) {
            const int b = tid >> 5, rr = tid & 31;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][rr];
            const float g = sigmoid_acc(v);
            const size_t row = (size_t)(b0 + b) * T + t;
            const bool live = b0 + b < p.B;
            if (rr < 16) {
                sm.zbuf[b][rr] = g;
                if (live) p.gates[row * 3 * ST + 16 * crank + rr] = g;
            } else {
                const int j = 16 * crank + rr - 16;
                const float rs = g * sm.s_full[b][j];
                sm.stage[b][rr - 16] = rs;
                if (live) { p.gates[row * 3 * ST + ST + j] = g; p.rhu[row * 2 * ST + j] = rs; }
            }
        }
        __syncthreads();
        dc_bcast<ST, 16, BG>(sm.stage, rs_a, bar_a[BAR_RS], crank, warp, lane);
        mbar_wait(&sm.bar[BAR_RS], parity);
        DC_TICK(5);

        // ---- h~ = tanh(G_h {r s_{t-1}, u_t}) ; s_t = (1-z) s_{t-1} + z h~   (GRU.lua:26-30) ------------------------------------
        if (warp < 8) dc_mv<BG, 16>(sm.gh, &sm.rs_full[0][0], ST, 8 * warp, 8 * warp, warp, lane, sm.part);
        else dc_mv<BG, 16>(sm.gh, &sm.u_full[0][0], ST, 8 * warp, 8 * warp - 64, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 16) {
            const int b = tid >> 4, r = tid & 15, j = 16 * crank + r;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][r];
            const float hc = tanh_acc(v), z = sm.zbuf[b][r], sp = sm.s_full[b][j];
            const float sn = (1.f - z) * sp + z * hc;
            sm.stage[b][r] = sn;
            if (b0 + b < p.B) {
                const size_t row = (size_t)(b0 + b) * T + t;
                p.gates[row * 3 * ST + 2 * ST + j] = hc;
                p.sc[row * (ST + A) + j] = sn;
                if (t + 1 < T) p.su[(row + 1) * 2 * ST + j] = sn;
            }
        }
        __syncthreads();
        dc_bcast<ST, 16, BG>(sm.stage, s_a, bar_a[BAR_S], crank, warp, lane);
        mbar_wait(&sm.bar[BAR_S], parity);
        DC_TICK(6);

        // ---- q_{t+1} = W_s s_t + b_s   (Attention.lua:65-67) -------------------------------------------------------------------
        if (t + 1 < T) {
            float acc[BG];
#pragma unroll
            for (int b = 0; b < BG; b++) {
                acc[b] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) acc[b] = dot4(wq[i], *reinterpret_cast<const float4*>(&sm.s_full[b][16 * warp + 4 * i]), acc[b]);
                sm.part[warp][b][lane] = acc[b];
            }
            __syncthreads();
            if (tid < BG * 32) {
                const int b = tid >> 5, k = tid & 31;
                float v = __ldg(p.qbias + 32 * crank + k);
#pragma unroll
                for (int w = 0; w < 16; w++) v += sm.part[w][b][k];
                sm.stage[b][k] = DC_K * v;
                if (b0 + b < p.B) p.q[((size_t)(b0 + b) * T + t + 1) * S + 32 * crank + k] = v;
            }
            __syncthreads();
            dc_bcast<S, 32, BG>(sm.stage, q_a, bar_a[BAR_Q], crank, warp, lane);
            if (warp < NR) load_pair(warp, va);              // Vh does not depend on q: the next step's first frames are fetched under the exchange
            if constexpr (LOC != 0) {   // alpha_t over this CTA's frames and the filter halo (written by the neighbours three exchanges ago)
                for (int i = tid; i < BG * (DC_RMAX + 16); i += DC_THREADS) {
                    const int b = i / (DC_RMAX + 16), x = i % (DC_RMAX + 16);
                    const int l = sm.l0_s[b] + x - p.padl;
                    float a = 0.f;
                    if (x < sm.nr_s[b] + p.KF - 1 && l >= 0 && l < sm.len_s[b]) a = __ldcg(p.alpha + ((size_t)(b0 + b) * T + t) * Lmax + l);
                    sm.aph_s[b][x] = dc_pack2(a, a);
                }
            }
            mbar_wait(&sm.bar[BAR_Q], parity);
        }
        DC_TICK(7);
        parity ^= 1;
    }
#undef DC_TICK
    cluster_sync_all();
}



// =================================================================================================================================
// backward time loop (RNNAttention.lua:233-253 with the deferred accumulations of decoder.cu), t = T-1 .. 0, same cluster geometry:
//   A  ds_t[own 16]  = carry + ds_mlp + dq_{t+1} . W_s[:, own]          dah, daz, carry = ds (1-z)           -> all-gather {dah, daz}
//   B  d(r s)[own], du_h[own] = dah . G_h[:, own | ST+own]                dar, carry += d(r s) r               -> all-gather dar
//   C  carry += {daz,dar} . G_zr[:, own] ; du[own] = du_h + {daz,dar} . G_zr[:, ST+own]                        -> all-gather du
//   D  dc_t[own 32]  = dc_mlp + du . W_jc[:, own]                                                              -> all-gather dc
//   E  own frames: dalpha_l = dc . h_l ; dot = sum alpha dalpha (cluster all-reduce) ; de_l = alpha_l (dalpha_l - dot)
//   F  own frames: dq += de_l w (1 - tanh^2(Vh_l + q_t))  -> reduce-scatter, CTA j sums dq[32j, +32)              -> all-gather dq_t
// The transposed weights (K-contiguous rows of the per-call copies W_s^T, G_h^T, G_zr^T, W_jc^T) stay on chip: 128 KB of shared memory
// (k-major) + W_s^T in registers.  Outputs are the arrays the deferred GEMMs / attn_dvh of decoder_backward read: dA = daz|dar|dah,
// du, dc, dq, de.  The alignment carry d alpha_{t-1} (location term through the hoisted Jacobian V1, monotonicity penalty) runs inside
// the kernel when the model has one (LOC / p.carry).
struct DecClusterBwdParams {
    const float *Vh, *h, *w, *q, *alpha, *gates, *su, *dsc;
    const float *WsT, *GhT, *GzrT, *WjcT;
    const int* lengths;
    int B, Lmax, T;
    float *dA, *du_all, *dc_all, *dq_all, *de_all;
    long long* clk;
    // alignment carry d alpha_{t-1} (RNNAttention.lua:246 through Attention.lua:86-99 and MonotonicAlignment.lua:49-75): location term
    // (folded weights U W_F [KF][S], left padding, the hoisted Jacobian V1 [B,T,Lmax,KF] of attn_v1) and / or the monotonicity penalty
    const float *uw, *V1, *pen;
    int KF, padl, carry;
    float lambda;
};

template <int BG, int LOC = 0>
struct DcBwdSmem {
    float wb[64 * 32 * 4];         // phase B rows: G_h^T   [own 16 | ST + own 16] x K = ST,   k-major
    float wc[128 * 32 * 4];        // phase C rows: G_zr^T  [own 16 | ST + own 16] x K = 2 ST
    float wd[64 * 32 * 4];         // phase D rows: W_jc^T  [own 32]               x K = ST
    alignas(16) float dq_full[BG][DC_S];
    float dahz_full[BG][2 * DC_ST];        // dah | daz
    float dar_full[BG][DC_ST];
    float du_full[BG][DC_ST];
    float dc_full[BG][DC_A];
    float q_full[BG][DC_S];        // 2 log2(e) q_t
    union {
        float part[16][BG][32];    // mat-vec partial sums (phases A-D)
        float recv_dq[DC_CS][BG][32];      // reduce-scatter receive (phase F): never live at the same time (see the exchange order)
    };
    float4 recv_dot[DC_CS][BG];
    alignas(16) float al_s[BG][DC_RMAX], dal_s[BG][DC_RMAX], de_s[BG][DC_RMAX + 16];    // de_s: zero past the slice (phase F reads blocks of 20)
    alignas(16) float w_s[DC_S];
    alignas(16) float stage[BG][32], stage2[BG][32];
    alignas(16) float carry_s[BG][16], duh_s[BG][16];
    int l0_s[BG], nr_s[BG], len_s[BG];
    int frow[BG * DC_RMAX];
    short fb[BG * DC_RMAX], fr[BG * DC_RMAX];
    uint64_t bar[7];
    alignas(16) float cin_s[BG][DC_RMAX];      // carry into d alpha_t of this CTA's frames: penalty of step t, penalty and location term of step t+1
    float pg_s[2][BG];             // lambda where the penalty of step t (slot t & 1) is active
    alignas(16) float uw_s[LOC ? DC_KFMAX * DC_S : 4];             // 2 log2(e) U W_F  (read as float4: the members above it do not add up to a multiple of 16 bytes for every BG)
    alignas(16) float apw_s[LOC ? BG : 1][DC_RMAX + 32];           // alpha_{t-1} over this CTA's frames plus the filter halo
};
enum { BB_X1 = 0, BB_X2, BB_X3, BB_X4, BB_X5, BB_X6A, BB_X6B };

// generic K-slice of a mat-vec phase: the warp covers NK * (ROWS == 32 ? 1 : 2) k4 starting at k4w (weights) / k4x (x)
template <int BG, int ROWS, int NK>
__device__ __forceinline__ void dc_mvg(const float* __restrict__ Wt, const float* __restrict__ x, int xpitch, int k4w, int k4x, int warp, int lane,
                                       float (*part)[BG][32]) {
    const int r = ROWS == 32 ? lane : (lane & 15), sub = ROWS == 32 ? 0 : (lane >> 4) * NK;
    float acc[BG];
#pragma unroll
    for (int b = 0; b < BG; b++) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < NK; i++) {
        const float4 wv = *reinterpret_cast<const float4*>(Wt + ((size_t)(k4w + sub + i) * ROWS + r) * 4);
#pragma unroll
        for (int b = 0; b < BG; b++) acc[b] = dot4(wv, *reinterpret_cast<const float4*>(x + (size_t)b * xpitch + (k4x + sub + i) * 4), acc[b]);
    }
    if (ROWS == 16) {
#pragma unroll
        for (int b = 0; b < BG; b++) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 16);
    }
    if (ROWS == 32 || lane < 16) {
#pragma unroll
        for (int b = 0; b < BG; b++) part[warp][b][r] = acc[b];
    }
}

// (1 - tanh^2(x)) / 4 for four elements, x given through y = 2 log2(e) x as in dc_tanh4_dot: with d = 2^y + 1 and r = 1 / d,
// 1 - tanh^2 = 4 r (1 - r); one reciprocal per four elements.  The factor 4 w is applied once per column after the frame sum.
__device__ __forceinline__ float4 dc_dtanh4(const float4 v, const float4 qk) {
    const float d0 = dc_ex2(fminf(fmaf(DC_K, v.x, qk.x), 10.f * DC_K)) + 1.f, d1 = dc_ex2(fminf(fmaf(DC_K, v.y, qk.y), 10.f * DC_K)) + 1.f;
    const float d2 = dc_ex2(fminf(fmaf(DC_K, v.z, qk.z), 10.f * DC_K)) + 1.f, d3 = dc_ex2(fminf(fmaf(DC_K, v.w, qk.w), 10.f * DC_K)) + 1.f;
    const float p01 = d0 * d1, p23 = d2 * d3;
    const float R = dc_rcp(p01 * p23);
    const float r01 = R * p23, r23 = R * p01;
    const float r0 = r01 * d1, r1 = r01 * d0, r2 = r23 * d3, r3 = r23 * d2;
    return make_float4(fmaf(-r0, r0, r0), fmaf(-r1, r1, r1), fmaf(-r2, r2, r2), fmaf(-r3, r3, r3));
}

template <int BG, int LOC>
__global__ void __launch_bounds__(DC_THREADS, 1)
dec_cluster_bwd_kernel(const DecClusterBwdParams p) {
    extern __shared__ __align__(128) unsigned char dc_smem_raw[];
    DcBwdSmem<BG, LOC>& sm = *reinterpret_cast<DcBwdSmem<BG, LOC>*>(dc_smem_raw);
    constexpr int ST = DC_ST, A = DC_A, S = DC_S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int b0 = (blockIdx.x / DC_CS) * BG;
    const int T = p.T, Lmax = p.Lmax;

    // ---- one-time staging (the per-call transposed copies are K-contiguous and 16-byte aligned) ----------------------------------
    for (int i = tid; i < 32 * 64; i += DC_THREADS) {          // K = ST phases: 64 k4
        const int r = i >> 6, k4 = i & 63;
        const int nb = r < 16 ? 16 * crank + r : ST + 16 * crank + r - 16;
        *reinterpret_cast<float4*>(sm.wb + ((size_t)k4 * 32 + r) * 4) = *reinterpret_cast<const float4*>(p.GhT + (size_t)nb * ST + 4 * k4);
        *reinterpret_cast<float4*>(sm.wd + ((size_t)k4 * 32 + r) * 4) = *reinterpret_cast<const float4*>(p.WjcT + (size_t)(32 * crank + r) * ST + 4 * k4);
    }
    for (int i = tid; i < 32 * 128; i += DC_THREADS) {         // K = 2 ST: 128 k4
        const int r = i >> 7, k4 = i & 127;
        const int nb = r < 16 ? 16 * crank + r : ST + 16 * crank + r - 16;
        *reinterpret_cast<float4*>(sm.wc + ((size_t)k4 * 32 + r) * 4) = *reinterpret_cast<const float4*>(p.GzrT + (size_t)nb * 2 * ST + 4 * k4);
    }
    for (int i = tid; i < S; i += DC_THREADS) sm.w_s[i] = p.w[i];
    for (int i = tid; i < BG * S; i += DC_THREADS) (&sm.dq_full[0][0])[i] = 0.f;          // dq_T = 0
    for (int i = tid; i < BG * 16; i += DC_THREADS) (&sm.carry_s[0][0])[i] = 0.f;         // Recurrent.lua:134
    for (int i = tid; i < BG * (DC_RMAX + 16); i += DC_THREADS) (&sm.de_s[0][0])[i] = 0.f;
    for (int i = tid; i < BG * DC_RMAX; i += DC_THREADS) (&sm.cin_s[0][0])[i] = 0.f;
    if constexpr (LOC != 0) {
        for (int i = tid; i < DC_KFMAX * S; i += DC_THREADS) sm.uw_s[i] = i < p.KF * S ? DC_K * p.uw[i] : 0.f;     // taps past KF: zero weights
    }
    if (tid < BG) {
        const int b = b0 + tid;
        const int Lb = b < p.B ? (p.lengths ? p.lengths[b] : Lmax) : 0;
        const int Rb = (Lb + DC_CS - 1) / DC_CS;
        const int l0 = (int)crank * Rb;
        sm.l0_s[tid] = l0; sm.nr_s[tid] = max(0, min(Rb, Lb - l0)); sm.len_s[tid] = Lb;
        sm.pg_s[0][tid] = 0.f; sm.pg_s[1][tid] = 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < 7; i++) mbar_init(&sm.bar[i], 1);
        fence_mbar_init();
    }
    // W_s^T rows of this CTA's 16 units, phase-A mapping: lane = (row, half), warp = 32 k
    float4 wa[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        wa[i] = *reinterpret_cast<const float4*>(p.WsT + (size_t)(16 * crank + (lane & 15)) * S + (8 * warp + 4 * (lane >> 4) + i) * 4);
    __syncthreads();
    int NR = 0;
#pragma unroll
    for (int b = 0; b < BG; b++) NR += sm.nr_s[b];
    for (int f = tid; f < NR; f += DC_THREADS) {
        int r = f, b = 0;
#pragma unroll
        for (int bb = 0; bb < BG - 1; bb++)
            if (b == bb && r >= sm.nr_s[bb]) { r -= sm.nr_s[bb]; b = bb + 1; }
        sm.frow[f] = (b0 + b) * Lmax + sm.l0_s[b] + r; sm.fb[f] = (short)b; sm.fr[f] = (short)r;
    }
    uint32_t bar_a[7];
#pragma unroll
    for (int i = 0; i < 7; i++) bar_a[i] = smem_u32(&sm.bar[i]);
    const uint32_t dahz_a = smem_u32(&sm.dahz_full[0][0]), dar_a = smem_u32(&sm.dar_full[0][0]), du_a = smem_u32(&sm.du_full[0][0]);
    const uint32_t dc_a = smem_u32(&sm.dc_full[0][0]), dq_a = smem_u32(&sm.dq_full[0][0]);
    const uint32_t rdot_a = smem_u32(&sm.recv_dot[0][0]), rdq_a = smem_u32(&sm.recv_dq[0][0][0]);
    cluster_sync_all();

    constexpr unsigned TX_256 = BG * 256 * 4, TX_512 = BG * 512 * 4, TX_DOT = DC_CS * BG * 16, TX_DQ = DC_CS * BG * 128;
    long long tck = 0;
    const bool prof = p.clk != nullptr && blockIdx.x == 0 && tid == 0;
#define DC_TICK(i) do { if (prof) { const long long n_ = clock64(); p.clk[i] += n_ - tck; tck = n_; } } while (0)
    if (prof) tck = clock64();

    // operands saved by the forward pass, per epilogue thread; each set is re-fetched for step t-1 as soon as its epilogue of step t
    // has consumed it, a whole step ahead of its use
    float pa_z = 0.f, pa_hc = 0.f, pa_sp = 0.f, pa_ds = 0.f, pb_r = 0.f, pb_sp = 0.f, pd_dc = 0.f;
    auto load_pa = [&](int t) {
        if (t >= 0 && tid < BG * 16 && b0 + (tid >> 4) < p.B) {
            const size_t row = (size_t)(b0 + (tid >> 4)) * T + t;
            const int j = 16 * crank + (tid & 15);
            pa_z = __ldg(p.gates + row * 3 * ST + j); pa_hc = __ldg(p.gates + row * 3 * ST + 2 * ST + j);
            pa_sp = __ldg(p.su + row * 2 * ST + j); pa_ds = __ldg(p.dsc + row * (ST + A) + j);
        }
    };
    auto load_pb = [&](int t) {
        if (t >= 0 && tid < BG * 32 && (tid & 31) < 16 && b0 + (tid >> 5) < p.B) {
            const size_t row = (size_t)(b0 + (tid >> 5)) * T + t;
            pb_r = __ldg(p.gates + row * 3 * ST + ST + 16 * crank + (tid & 31)); pb_sp = __ldg(p.su + row * 2 * ST + 16 * crank + (tid & 31));
        }
    };
    auto load_pd = [&](int t) {
        if (t >= 0 && tid < BG * 32 && b0 + (tid >> 5) < p.B)
            pd_dc = __ldg(p.dsc + ((size_t)(b0 + (tid >> 5)) * T + t) * (ST + A) + ST + 32 * crank + (tid & 31));
    };
    load_pa(T - 1); load_pb(T - 1); load_pd(T - 1);

    unsigned parity = 0;
    for (int t = T - 1; t >= 0; t--) {
        if (tid == 0) {
            mbar_expect_tx(&sm.bar[BB_X1], TX_512); mbar_expect_tx(&sm.bar[BB_X2], TX_256); mbar_expect_tx(&sm.bar[BB_X3], TX_256);
            mbar_expect_tx(&sm.bar[BB_X4], TX_512); mbar_expect_tx(&sm.bar[BB_X5], TX_DOT); mbar_expect_tx(&sm.bar[BB_X6A], TX_DQ);
            mbar_expect_tx(&sm.bar[BB_X6B], TX_512);
        }
        if (p.carry && tid < BG)     // MonotonicAlignment.lua:49-75: the penalty of step t feeds d alpha_t (+) and d alpha_{t-1} (-) where it is active
            sm.pg_s[t & 1][tid] = (p.lambda != 0.f && b0 + tid < p.B && __ldg(p.pen + (size_t)(b0 + tid) * T + t) > 0.f) ? p.lambda : 0.f;

        // ---- A: ds_t, elementwise GRU backward ----------------------------------------------------------------------------------
        {
            const int r = lane & 15, half = lane >> 4;
            float acc[BG];
#pragma unroll
            for (int b = 0; b < BG; b++) {
                acc[b] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) acc[b] = dot4(wa[i], *reinterpret_cast<const float4*>(&sm.dq_full[b][(8 * warp + 4 * half + i) * 4]), acc[b]);
                acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 16);
                if (lane < 16) sm.part[warp][b][r] = acc[b];
            }
        }
        __syncthreads();
        if (tid < BG * 16) {
            const int b = tid >> 4, r = tid & 15, j = 16 * crank + r;
            float v = sm.carry_s[b][r] + pa_ds;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][r];
            const float z = pa_z, hc = pa_hc;
            const float dah = v * z * (1.f - hc * hc), daz = v * (hc - pa_sp) * z * (1.f - z);      // GRU.lua:26-30 reversed
            sm.carry_s[b][r] = v * (1.f - z);
            sm.stage[b][r] = dah; sm.stage2[b][r] = daz;
            if (b0 + b < p.B) {
                const size_t row = (size_t)(b0 + b) * T + t;
                p.dA[row * 3 * ST + 2 * ST + j] = dah; p.dA[row * 3 * ST + j] = daz;
            }
        }
        load_pa(t - 1);
        __syncthreads();
        dc_bcast<2 * ST, 16, BG>(sm.stage, dahz_a, bar_a[BB_X1], crank, warp, lane);
        dc_bcast<2 * ST, 16, BG>(sm.stage2, dahz_a + ST * 4u, bar_a[BB_X1], crank, warp, lane);
        // q_t and alpha_t of this CTA's frames (saved by the forward pass), staged under the exchange
        for (int i = tid; i < BG * DC_RMAX; i += DC_THREADS) {
            const int b = i / DC_RMAX, r = i % DC_RMAX;
            sm.al_s[b][r] = r < sm.nr_s[b] ? __ldg(p.alpha + ((size_t)(b0 + b) * T + t) * Lmax + sm.l0_s[b] + r) : 0.f;
        }
        if (p.carry) {
            // d alpha_t of this CTA's frames from outside the context path: + g_t (penalty of step t), and from step t+1: - g_{t+1} and
            // the location term sum_j de_{t+1}[x] V1_{t+1}[x][j], x = l - j + pad_left (de_{t+1} of the neighbours' frames through L2:
            // written, fenced and followed by two cluster exchanges in step t+1)
            for (int i = tid; i < BG * DC_RMAX; i += DC_THREADS) {
                const int b = i / DC_RMAX, r = i % DC_RMAX;
                if (r < sm.nr_s[b]) {
                    const int Lb = sm.len_s[b], l = sm.l0_s[b] + r;
                    float c = sm.pg_s[t & 1][b] * (float)(Lb - l);
                    if (t + 1 < T) {
                        c -= sm.pg_s[(t + 1) & 1][b] * (float)(Lb - l);
                        if constexpr (LOC != 0) {
                            const size_t row = ((size_t)(b0 + b) * T + t + 1) * Lmax;
                            for (int jj = 0; jj < p.KF; jj++) {
                                const int x = l - jj + p.padl;
                                if (x >= 0 && x < Lb) c = fmaf(__ldcg(p.de_all + row + x), __ldg(p.V1 + (row + x) * p.KF + jj), c);
                            }
                        }
                    }
                    sm.cin_s[b][r] = c;
                }
            }
        }
        if constexpr (LOC != 0) {
            for (int i = tid; i < BG * (DC_RMAX + 32); i += DC_THREADS) {
                const int b = i / (DC_RMAX + 32), x = i % (DC_RMAX + 32);
                const int l = sm.l0_s[b] + x - p.padl;
                sm.apw_s[b][x] = (t > 0 && b0 + b < p.B && l >= 0 && l < sm.len_s[b]) ? __ldg(p.alpha + ((size_t)(b0 + b) * T + t - 1) * Lmax + l) : 0.f;
            }
        }
        {
            float qv[BG];
#pragma unroll
            for (int b = 0; b < BG; b++) qv[b] = b0 + b < p.B ? __ldg(p.q + ((size_t)(b0 + b) * T + t) * S + tid) : 0.f;      // S == DC_THREADS
#pragma unroll
            for (int b = 0; b < BG; b++) sm.q_full[b][tid] = DC_K * qv[b];
        }
        mbar_wait(&sm.bar[BB_X1], parity);
        DC_TICK(0);

        // ---- B: d(r s), du_h = dah . G_h ; dar -----------------------------------------------------------------------------------
        dc_mvg<BG, 32, 4>(sm.wb, &sm.dahz_full[0][0], 2 * ST, 4 * warp, 4 * warp, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 32) {
            const int b = tid >> 5, rr = tid & 31;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][rr];
            if (rr < 16) {
                const float dar = v * pb_sp * pb_r * (1.f - pb_r);                                 // GRU.lua:24-25 reversed
                sm.carry_s[b][rr] += v * pb_r;
                sm.stage[b][rr] = dar;
                if (b0 + b < p.B) p.dA[((size_t)(b0 + b) * T + t) * 3 * ST + ST + 16 * crank + rr] = dar;
            } else {
                sm.duh_s[b][rr - 16] = v;
            }
        }
        load_pb(t - 1);
        __syncthreads();
        dc_bcast<ST, 16, BG>(sm.stage, dar_a, bar_a[BB_X2], crank, warp, lane);
        mbar_wait(&sm.bar[BB_X2], parity);
        DC_TICK(1);

        // ---- C: d{s_{t-1}, u} += {daz, dar} . G_zr ----------------------------------------------------------------------------------
        if (warp < 8) dc_mvg<BG, 32, 8>(sm.wc, &sm.dahz_full[0][0] + ST, 2 * ST, 8 * warp, 8 * warp, warp, lane, sm.part);
        else dc_mvg<BG, 32, 8>(sm.wc, &sm.dar_full[0][0], ST, 8 * warp, 8 * warp - 64, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 32) {
            const int b = tid >> 5, rr = tid & 31;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][rr];
            if (rr < 16) {
                sm.carry_s[b][rr] += v;
            } else {
                const float du = v + sm.duh_s[b][rr - 16];
                sm.stage[b][rr - 16] = du;
                if (b0 + b < p.B) p.du_all[((size_t)(b0 + b) * T + t) * ST + 16 * crank + rr - 16] = du;
            }
        }
        __syncthreads();
        dc_bcast<ST, 16, BG>(sm.stage, du_a, bar_a[BB_X3], crank, warp, lane);
        mbar_wait(&sm.bar[BB_X3], parity);
        DC_TICK(2);

        // ---- D: dc_t = dc_mlp + du . W_jc ---------------------------------------------------------------------------------------------
        dc_mvg<BG, 32, 4>(sm.wd, &sm.du_full[0][0], ST, 4 * warp, 4 * warp, warp, lane, sm.part);
        __syncthreads();
        if (tid < BG * 32) {
            const int b = tid >> 5, k = tid & 31;
            float v = pd_dc;
#pragma unroll
            for (int w = 0; w < 16; w++) v += sm.part[w][b][k];
            sm.stage[b][k] = v;
            if (b0 + b < p.B) p.dc_all[((size_t)(b0 + b) * T + t) * A + 32 * crank + k] = v;
        }
        load_pd(t - 1);
        __syncthreads();
        dc_bcast<A, 32, BG>(sm.stage, dc_a, bar_a[BB_X4], crank, warp, lane);
        mbar_wait(&sm.bar[BB_X4], parity);
        DC_TICK(3);

        // ---- E: dalpha_l = dc . h_l over this CTA's frames (Attention.lua:132-134 reversed) ----------------------------------------
        for (int f0 = warp; f0 < NR; f0 += 32) {
            float4 hv[2][4];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int f = f0 + 16 * j < NR ? f0 + 16 * j : f0;
                const float* hp = p.h + (size_t)sm.frow[f] * A + lane * 4;
#pragma unroll
                for (int i = 0; i < 4; i++) hv[j][i] = ldg_stream(hp + 128 * i);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int f = f0 + 16 * j < NR ? f0 + 16 * j : f0;
                const int b = sm.fb[f];
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) acc = dot4(hv[j][i], *reinterpret_cast<const float4*>(&sm.dc_full[b][lane * 4 + 128 * i]), acc);
                acc = warp_sum(acc);
                if (lane == 0) sm.dal_s[b][sm.fr[f]] = acc + sm.cin_s[b][sm.fr[f]];
            }
        }
        __syncthreads();
        if (warp < BG) {   // sum_l alpha dalpha over this CTA's frames -> every CTA of the cluster
            const int b = warp, nr = sm.nr_s[b];
            const float d0 = lane < nr ? sm.al_s[b][lane] * sm.dal_s[b][lane] : 0.f;
            const float d1 = lane + 32 < nr ? sm.al_s[b][lane + 32] * sm.dal_s[b][lane + 32] : 0.f;
            const float dotp = warp_sum(d0 + d1);
            if (lane < DC_CS)
                st_async_v4(mapa_rank(rdot_a + (uint32_t)(crank * BG + b) * 16u, lane), make_float4(dotp, 0.f, 0.f, 0.f), mapa_rank(bar_a[BB_X5], lane));
        }
        mbar_wait(&sm.bar[BB_X5], parity);
        if (warp < BG) {   // de_l = alpha_l (dalpha_l - <alpha, dalpha>)   (SoftMax backward)
            const int b = warp, nr = sm.nr_s[b];
            const float dot = warp_sum(lane < DC_CS ? sm.recv_dot[lane][b].x : 0.f);
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const int r = lane + 32 * hh;
                const float de = r < nr ? sm.al_s[b][r] * (sm.dal_s[b][r] - dot) : 0.f;
                sm.de_s[b][r] = de;
                if (r < nr) p.de_all[((size_t)(b0 + b) * T + t) * Lmax + sm.l0_s[b] + r] = de;
            }
            if constexpr (LOC != 0) __threadfence();             // the neighbours read these de_t (filter halo) through L2 in step t-1
        }
        __syncthreads();
        DC_TICK(4);

        // ---- F: dq_t = sum_l de_l w (1 - tanh^2(Vh_l + q_t))   (Attention.lua:95-121 reversed) -------------------------------------------
        // thread = (float4 column c4, row group g) like the context phase of the forward kernel: warp w holds the 8 columns of CTA w's
        // slice x 4 row groups, so the CTA's partial needs no shared-memory accumulation: two shuffles, then straight to the owner.
        if constexpr (LOC != 0) {
            // location-aware energies: Z_l = q_t + Vh_l + U W_F alpha_{t-1}[l + j - pad_left].  One utterance at a time; a thread's five
            // frames are CONSECUTIVE so their filter windows share a 14-value run of alpha_{t-1}; the thread's columns of U W_F stay in
            // registers for the whole phase.
            const int c4 = 8 * warp + (lane & 7), g = lane >> 3, dst = warp;
            const uint32_t dbar = mapa_rank(bar_a[BB_X6A], dst);
            float4 w4 = *reinterpret_cast<const float4*>(&sm.w_s[c4 * 4]);
            w4.x *= 4.f; w4.y *= 4.f; w4.z *= 4.f; w4.w *= 4.f;
            float4 uwr[DC_KFMAX];
#pragma unroll
            for (int jj = 0; jj < DC_KFMAX; jj++) uwr[jj] = *reinterpret_cast<const float4*>(&sm.uw_s[jj * S + c4 * 4]);
#pragma unroll 1
            for (int b = 0; b < BG; b++) {
                const int nr = sm.nr_s[b];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 qk = *reinterpret_cast<const float4*>(&sm.q_full[b][c4 * 4]);
                const float* vp = p.Vh + ((size_t)(b0 + b) * Lmax + sm.l0_s[b]) * S + c4 * 4;
                for (int base = 0; base < nr; base += 20) {
                    const int r0 = base + 5 * g;
                    float4 vx[5];
#pragma unroll
                    for (int u = 0; u < 5; u++) vx[u] = ldg_stream(vp + (size_t)min(r0 + u, nr - 1) * S);
                    float a[DC_KFMAX + 4];
#pragma unroll
                    for (int x = 0; x < DC_KFMAX + 4; x++) a[x] = sm.apw_s[b][r0 + x];
#pragma unroll
                    for (int u = 0; u < 5; u++) {
                        float4 z = qk;
#pragma unroll
                        for (int jj = 0; jj < DC_KFMAX; jj++) {
                            z.x = fmaf(a[u + jj], uwr[jj].x, z.x); z.y = fmaf(a[u + jj], uwr[jj].y, z.y);
                            z.z = fmaf(a[u + jj], uwr[jj].z, z.z); z.w = fmaf(a[u + jj], uwr[jj].w, z.w);
                        }
                        const float de = sm.de_s[b][r0 + u];                           // zero past the slice
                        const float4 gv = dc_dtanh4(vx[u], z);
                        acc.x = fmaf(de, gv.x, acc.x); acc.y = fmaf(de, gv.y, acc.y);
                        acc.z = fmaf(de, gv.z, acc.z); acc.w = fmaf(de, gv.w, acc.w);
                    }
                }
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                }
                if (g == 0)
                    st_async_v4(mapa_rank(rdq_a + (uint32_t)((crank * BG + b) * 32 + (c4 & 7) * 4) * 4u, dst),
                                make_float4(w4.x * acc.x, w4.y * acc.y, w4.z * acc.z, w4.w * acc.w), dbar);
            }
        } else {
            const int c4 = 8 * warp + (lane & 7), g = lane >> 3, dst = warp;
            const uint32_t dbar = mapa_rank(bar_a[BB_X6A], dst);
            float4 w4 = *reinterpret_cast<const float4*>(&sm.w_s[c4 * 4]);
            w4.x *= 4.f; w4.y *= 4.f; w4.z *= 4.f; w4.w *= 4.f;
            float4 vx[2][5];
            auto f_load = [&](int bp, int base) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (bp + j >= BG) continue;
                    const int b = bp + j, nr = sm.nr_s[b];
                    if (nr <= base) {
#pragma unroll
                        for (int u = 0; u < 5; u++) vx[j][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        continue;
                    }
                    const float* vp = p.Vh + ((size_t)(b0 + b) * Lmax + sm.l0_s[b]) * S + c4 * 4;
#pragma unroll
                    for (int u = 0; u < 5; u++) vx[j][u] = ldg_stream(vp + (size_t)min(base + g + 4 * u, nr - 1) * S);
                }
            };
            f_load(0, 0);
#pragma unroll
            for (int bp = 0; bp < BG; bp += 2) {
                float4 acc[2];
                acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int nrm = max(sm.nr_s[bp], bp + 1 < BG ? sm.nr_s[bp + 1] : 0);
                for (int base = 0; base < nrm || base == 0; base += 20) {
                    if (base > 0) f_load(bp, base);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        if (bp + j < BG) {
                            const float4 qk = *reinterpret_cast<const float4*>(&sm.q_full[bp + j][c4 * 4]);
#pragma unroll
                            for (int u = 0; u < 5; u++) {
                                const float de = sm.de_s[bp + j][base + g + 4 * u];           // zero past the slice
                                const float4 gv = dc_dtanh4(vx[j][u], qk);
                                acc[j].x = fmaf(de, gv.x, acc[j].x); acc[j].y = fmaf(de, gv.y, acc[j].y);
                                acc[j].z = fmaf(de, gv.z, acc[j].z); acc[j].w = fmaf(de, gv.w, acc[j].w);
                            }
                        }
                    }
                }
                if (bp + 2 < BG) f_load(bp + 2, 0);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (bp + j < BG) {
#pragma unroll
                        for (int o = 8; o <= 16; o <<= 1) {
                            acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, o); acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, o);
                            acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, o); acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, o);
                        }
                        if (g == 0)
                            st_async_v4(mapa_rank(rdq_a + (uint32_t)((crank * BG + bp + j) * 32 + (c4 & 7) * 4) * 4u, dst),
                                        make_float4(w4.x * acc[j].x, w4.y * acc[j].y, w4.z * acc[j].z, w4.w * acc[j].w), dbar);
                    }
                }
            }
        }
        mbar_wait(&sm.bar[BB_X6A], parity);
        DC_TICK(5);
        if (tid < BG * 32) {
            const int b = tid >> 5, k = tid & 31;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < DC_CS; i++) v += sm.recv_dq[i][b][k];
            sm.stage[b][k] = v;
            if (b0 + b < p.B) p.dq_all[((size_t)(b0 + b) * T + t) * S + 32 * crank + k] = v;
        }
        __syncthreads();
        dc_bcast<S, 32, BG>(sm.stage, dq_a, bar_a[BB_X6B], crank, warp, lane);
        mbar_wait(&sm.bar[BB_X6B], parity);
        DC_TICK(6);
        parity ^= 1;
    }
#undef DC_TICK
    cluster_sync_all();
}

template <int BG, int LOC>
static int dc_launch(s2s_ctx* ctx, const DecClusterParams& p, int* max_clusters) {
    static bool attr = false;
    const size_t smem = sizeof(DcSmem<BG, LOC>);
    if (!attr) {
        S2S_CUDA(cudaFuncSetAttribute(dec_cluster_fwd_kernel<BG, LOC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        S2S_CUDA(cudaFuncSetAttribute(dec_cluster_fwd_kernel<BG, LOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(DC_CS * ceil_div(p.B, BG));
    cfg.blockDim = dim3(DC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = DC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (max_clusters) {
        if (cudaOccupancyMaxActiveClusters(max_clusters, dec_cluster_fwd_kernel<BG, LOC>, &cfg) != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    S2S_CUDA(cudaLaunchKernelEx(&cfg, dec_cluster_fwd_kernel<BG, LOC>, p));
    ctx->kcount[S2S_KC_DEC_CLUSTER_FWD]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

static int dc_enabled() {
    const char* e = getenv("S2S_DEC_CLUSTER");      // read per call: tests and benchmarks compare both paths in one process
    return e ? atoi(e) : 1;
}

// The decoder time loop of decoder_forward on the cluster kernel.  *handled = false (nothing launched) when the shapes are
// not the ones the kernel is built for or clusters of 16 cannot be scheduled; the caller then runs the per-step path.
int decoder_cluster_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax, const int* tlens,
                            int T, float lambda, const float* uy, DecoderState& d, bool* handled) {
    *handled = false;
    const int KF = Y.K > 0 ? Y.KF : 0;
    if (!dc_enabled() || Y.ST != DC_ST || Y.A != DC_A || Y.S != DC_S || KF > DC_KFMAX || Lmax > DC_CS * DC_RMAX) return 0;
    if (KF > 0) { const char* e = getenv("S2S_DEC_CLUSTER_LOC"); if (e && !atoi(e)) return 0; }
    static int cap = -1;
    DecClusterParams p = {};
    p.Vh = d.Vh; p.h = h; p.w = P + Y.we.off; p.qbias = d.qbias; p.Ws = P + Y.Ws.off; p.Wjc = d.Wjc; p.Gz = P + Y.Gz.off; p.Gh = P + Y.Gh.off;
    p.uy = uy; p.lengths = lengths; p.tlens = tlens; p.B = B; p.Lmax = Lmax; p.T = T; p.lambda = lambda;
    p.alpha = d.alpha; p.sc = d.sc; p.q = d.q; p.pen = d.pen; p.su = d.su; p.rhu = d.rhu; p.gates = d.gates;
    p.uw = d.uw; p.KF = KF; p.padl = KF > 0 ? ((KF % 2 == 1) ? (KF - 1) / 2 : KF / 2) : 0;          // Attention.lua:77-85
    if (cap < 0) { int n = 0; S2S_TRY((dc_launch<5, 1>(ctx, p, &n))); cap = n; }
    if (cap < 1) return 0;
    static long long* clk = nullptr;
    static int prof = -1;
    if (prof < 0) { const char* e = getenv("S2S_DEC_PROF"); prof = e ? atoi(e) : 0; }
    if (prof && !ctx->capturing) {
        if (!clk) { S2S_CUDA(cudaMalloc(&clk, 16 * sizeof(long long))); }
        S2S_CUDA(cudaMemsetAsync(clk, 0, 16 * sizeof(long long), ctx->stream));
        p.clk = clk;
    }
    // frames past an utterance's length keep alpha = 0 (the per-step kernel writes those zeros every step)
    S2S_CUDA(cudaMemsetAsync(d.alpha, 0, (size_t)B * T * Lmax * sizeof(float), ctx->stream));
    int bg = 1;
    while (bg < 5 && ceil_div(B, bg) > cap) bg++;          // one wave of clusters when possible
    { const char* e = getenv("S2S_DEC_BG"); if (e && atoi(e) >= 1 && atoi(e) <= 5) bg = atoi(e); }
    prof_begin(ctx, S2S_PROF_DEC_FWD);
    if (KF > 0) {
        switch (bg) {
            case 1: S2S_TRY((dc_launch<1, 1>(ctx, p, nullptr))); break;
            case 2: S2S_TRY((dc_launch<2, 1>(ctx, p, nullptr))); break;
            case 3: S2S_TRY((dc_launch<3, 1>(ctx, p, nullptr))); break;
            case 4: S2S_TRY((dc_launch<4, 1>(ctx, p, nullptr))); break;
            default: S2S_TRY((dc_launch<5, 1>(ctx, p, nullptr))); break;
        }
    } else {
        switch (bg) {
            case 1: S2S_TRY((dc_launch<1, 0>(ctx, p, nullptr))); break;
            case 2: S2S_TRY((dc_launch<2, 0>(ctx, p, nullptr))); break;
            case 3: S2S_TRY((dc_launch<3, 0>(ctx, p, nullptr))); break;
            case 4: S2S_TRY((dc_launch<4, 0>(ctx, p, nullptr))); break;
            default: S2S_TRY((dc_launch<5, 0>(ctx, p, nullptr))); break;
        }
    }
    prof_end(ctx, S2S_PROF_DEC_FWD, 4.0 * B * T * ((double)Lmax * (DC_S + DC_A)));
    if (p.clk) {
        long long hclk[16];
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
        S2S_CUDA(cudaMemcpy(hclk, clk, sizeof(hclk), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[dec_cluster] B=%d L=%d T=%d BG=%d clocks/step: score %lld | stats %lld | ctx %lld | wait cp %lld | combine+c %lld | u %lld | zr %lld | h %lld | q %lld\n",
                B, Lmax, T, bg, hclk[0] / T, hclk[8] / T, hclk[1] / T, hclk[2] / T, hclk[3] / T, hclk[4] / T, hclk[5] / T, hclk[6] / T, hclk[7] / T);
    }
    *handled = true;
    return 0;
}


template <int BG, int LOC>
static int dcb_launch(s2s_ctx* ctx, const DecClusterBwdParams& p, int* max_clusters) {
    static bool attr = false;
    const size_t smem = sizeof(DcBwdSmem<BG, LOC>);
    static_assert(sizeof(DcBwdSmem<BG, LOC>) <= 227 * 1024, "decoder cluster backward: shared memory");
    if (!attr) {
        S2S_CUDA(cudaFuncSetAttribute(dec_cluster_bwd_kernel<BG, LOC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        S2S_CUDA(cudaFuncSetAttribute(dec_cluster_bwd_kernel<BG, LOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(DC_CS * ceil_div(p.B, BG));
    cfg.blockDim = dim3(DC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = DC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (max_clusters) {
        if (cudaOccupancyMaxActiveClusters(max_clusters, dec_cluster_bwd_kernel<BG, LOC>, &cfg) != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    S2S_CUDA(cudaLaunchKernelEx(&cfg, dec_cluster_bwd_kernel<BG, LOC>, p));
    ctx->kcount[S2S_KC_DEC_CLUSTER_BWD]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

bool decoder_cluster_backward_eligible(const Layout& Y, int Lmax, float lambda) {
    const int KF = Y.K > 0 ? Y.KF : 0;
    { const char* e = getenv("S2S_DEC_CLUSTER_BWD"); if (e && !atoi(e)) return false; }
    (void)lambda;                // the alignment carry (location term, monotonicity penalty) runs inside the cluster kernel
    return dc_enabled() && Y.ST == DC_ST && Y.A == DC_A && Y.S == DC_S && KF <= DC_KFMAX && Lmax <= DC_CS * DC_RMAX;
}

// The time loop of decoder_backward on the cluster kernel (same conditions as the forward one).  V1 (location path): the hoisted
// Jacobian of attn_v1, [B, T, Lmax, KF].
int decoder_cluster_backward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax, int T, float lambda,
                             const DecoderState& d, const float* WsT, const float* GhT, const float* GzrT, const float* WjcT, const float* dsc,
                             const float* V1, float* dA, float* du_all, float* dc_all, float* dq_all, float* de_all, bool* handled) {
    *handled = false;
    if (!decoder_cluster_backward_eligible(Y, Lmax, lambda)) return 0;
    static int cap = -1;
    const int KF = Y.K > 0 ? Y.KF : 0;
    S2S_REQUIRE(KF == 0 || V1 != nullptr, "decoder_cluster_backward: the location path needs V1");
    DecClusterBwdParams p = {};
    p.uw = d.uw; p.V1 = V1; p.pen = d.pen; p.KF = KF; p.padl = KF > 0 ? ((KF % 2 == 1) ? (KF - 1) / 2 : KF / 2) : 0;      // Attention.lua:77-85
    p.lambda = lambda; p.carry = (KF > 0 || lambda != 0.f) ? 1 : 0;
    p.Vh = d.Vh; p.h = h; p.w = P + Y.we.off; p.q = d.q; p.alpha = d.alpha; p.gates = d.gates; p.su = d.su; p.dsc = dsc;
    p.WsT = WsT; p.GhT = GhT; p.GzrT = GzrT; p.WjcT = WjcT; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.T = T;
    p.dA = dA; p.du_all = du_all; p.dc_all = dc_all; p.dq_all = dq_all; p.de_all = de_all;
    if (cap < 0) { int n = 0; S2S_TRY((dcb_launch<5, 1>(ctx, p, &n))); cap = n; }
    if (cap < 1) return 0;
    static long long* clk = nullptr;
    static int prof = -1;
    if (prof < 0) { const char* e = getenv("S2S_DEC_PROF"); prof = e ? atoi(e) : 0; }
    if (prof && !ctx->capturing) {
        if (!clk) { S2S_CUDA(cudaMalloc(&clk, 16 * sizeof(long long))); }
        S2S_CUDA(cudaMemsetAsync(clk, 0, 16 * sizeof(long long), ctx->stream));
        p.clk = clk;
    }
    S2S_CUDA(cudaMemsetAsync(de_all, 0, (size_t)B * T * Lmax * sizeof(float), ctx->stream));       // frames past an utterance's length
    int bg = 1;
    while (bg < 5 && ceil_div(B, bg) > cap) bg++;
    { const char* e = getenv("S2S_DEC_BG"); if (e && atoi(e) >= 1 && atoi(e) <= 5) bg = atoi(e); }
    prof_begin(ctx, S2S_PROF_DEC_BWD);
    if (KF > 0) {
        switch (bg) {
            case 1: S2S_TRY((dcb_launch<1, 1>(ctx, p, nullptr))); break;
            case 2: S2S_TRY((dcb_launch<2, 1>(ctx, p, nullptr))); break;
            case 3: S2S_TRY((dcb_launch<3, 1>(ctx, p, nullptr))); break;
            case 4: S2S_TRY((dcb_launch<4, 1>(ctx, p, nullptr))); break;
            default: S2S_TRY((dcb_launch<5, 1>(ctx, p, nullptr))); break;
        }
    } else {
        switch (bg) {
            case 1: S2S_TRY((dcb_launch<1, 0>(ctx, p, nullptr))); break;
            case 2: S2S_TRY((dcb_launch<2, 0>(ctx, p, nullptr))); break;
            case 3: S2S_TRY((dcb_launch<3, 0>(ctx, p, nullptr))); break;
            case 4: S2S_TRY((dcb_launch<4, 0>(ctx, p, nullptr))); break;
            default: S2S_TRY((dcb_launch<5, 0>(ctx, p, nullptr))); break;
        }
    }
    prof_end(ctx, S2S_PROF_DEC_BWD, 4.0 * B * T * ((double)Lmax * (DC_S + DC_A)));
    if (p.clk) {
        long long hclk[16];
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
        S2S_CUDA(cudaMemcpy(hclk, clk, sizeof(hclk), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[dec_cluster bwd] B=%d L=%d T=%d BG=%d clocks/step: A %lld | B %lld | C %lld | D %lld | E (dalpha, dot, de) %lld | F (dq) %lld | dq gather %lld\n",
                B, Lmax, T, bg, hclk[0] / T, hclk[1] / T, hclk[2] / T, hclk[3] / T, hclk[4] / T, hclk[5] / T, hclk[6] / T);
    }
    *handled = true;
    return 0;
}

}  // namespace s2s
