// gru_seq5.cu -- fifth generation of the persistent cluster GRU recurrence (nn.RNN(nn.GRU), RNN.lua:120-201, GRU.lua:22-30):
// generation 3 (gru_seq3.cu: warp-specialised, two pipelined sub-batches, dedicated owner warps) with the mat-vec re-mapped so that it
// reads HALF as much shared memory.
//
// The timeline of generation 3 (S2S_GRU_TRACE, profiles/r02_gru_trace.txt) shows the four mat-vec phases of a step taking 2190 of its
// 3140 cycles against an FMA floor of 960, whatever the accumulator layout -- and 640 warp-wide LDS.128 per step per CTA: a 16-byte
// shared-memory read is served in four passes even when all 32 lanes read the same address, so the state broadcasts alone keep the
// shared-memory pipe busy for ~2200 cycles of every step.  The FMA : LDS ratio was 4 packed FMAs per load in phase 1 and 2 in phase 2
// because a lane owned ONE unit (row) of each gate over the warp's 32-wide K-slice.
// Here a lane owns TWO units over HALF the K-slice (lane = (k-half, unit pair)): the same 96 weights per lane, the same FMAs, but every
// state load now feeds twice as many of them and a phase needs 4 instead of 8 loads per utterance; the two k-halves are added with one
// shuffle per output and the lower half-warp writes the partial sums.  Everything else is generation 3's.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "cluster_rnn.cuh"
#include "common.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {

typedef unsigned long long g5_f2;
__device__ __forceinline__ g5_f2 g5_pack(float a, float b) { g5_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float g5_hsum(g5_f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
__device__ __forceinline__ g5_f2 g5_fma2(g5_f2 a, g5_f2 b, g5_f2 c) { g5_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void g5_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void g5_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// lane = (hf = lane >> 4: k-half, up = lane & 15: unit pair).  w2[8 j + kk] = weights of unit 2 up + j for k = k0 + 2 kk, 2 kk + 1
// (k0 = start of this lane's 16-wide k-half).  out[(LO + b) * ostride + 2 up + j] = sum over the warp's 32-wide K-slice.
template <int H, int LO, int N>
__device__ __forceinline__ void g5_mv(const g5_f2 (&w2)[16], const float (*x)[H], int k0, float* out, int ostride, int lane) {
    g5_f2 a0[2][N > 0 ? N : 1], a1[2][N > 0 ? N : 1];
#pragma unroll
    for (int b = 0; b < N; b++) { a0[0][b] = 0ull; a0[1][b] = 0ull; a1[0][b] = 0ull; a1[1][b] = 0ull; }
#pragma unroll
    for (int k4 = 0; k4 < 4; k4++) {
        ulonglong2 xv[N > 0 ? N : 1];
#pragma unroll
        for (int b = 0; b < N; b++) xv[b] = *reinterpret_cast<const ulonglong2*>(&x[LO + b][k0 + 4 * k4]);
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int b = 0; b < N; b++) a0[j][b] = g5_fma2(w2[8 * j + 2 * k4], xv[b].x, a0[j][b]);
#pragma unroll
            for (int b = 0; b < N; b++) a1[j][b] = g5_fma2(w2[8 * j + 2 * k4 + 1], xv[b].y, a1[j][b]);
        }
    }
#pragma unroll
    for (int b = 0; b < N; b++) {
        float v0 = g5_hsum(a0[0][b]) + g5_hsum(a1[0][b]), v1 = g5_hsum(a0[1][b]) + g5_hsum(a1[1][b]);
        v0 += __shfl_xor_sync(0xffffffffu, v0, 16); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
        if (lane < 16) *reinterpret_cast<float2*>(&out[(size_t)(LO + b) * ostride + 2 * lane]) = make_float2(v0, v1);
    }
}
// two gates that read the same state slice: the broadcasts are shared
template <int H, int LO, int N, bool SAME>
__device__ __forceinline__ void g5_mv2(const g5_f2 (&wa)[16], const g5_f2 (&wb)[16], const float (*xa)[H], const float (*xb)[H], int k0,
                                       float* outa, float* outb, int ostride, int lane) {
    g5_f2 a[2][N > 0 ? N : 1], c[2][N > 0 ? N : 1];
#pragma unroll
    for (int b = 0; b < N; b++) { a[0][b] = 0ull; a[1][b] = 0ull; c[0][b] = 0ull; c[1][b] = 0ull; }
#pragma unroll
    for (int k4 = 0; k4 < 4; k4++) {
        ulonglong2 xv[N > 0 ? N : 1], yv[N > 0 ? N : 1];
#pragma unroll
        for (int b = 0; b < N; b++) {
            xv[b] = *reinterpret_cast<const ulonglong2*>(&xa[LO + b][k0 + 4 * k4]);
            yv[b] = SAME ? xv[b] : *reinterpret_cast<const ulonglong2*>(&xb[LO + b][k0 + 4 * k4]);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int b = 0; b < N; b++) { a[j][b] = g5_fma2(wa[8 * j + 2 * k4], xv[b].x, a[j][b]); c[j][b] = g5_fma2(wb[8 * j + 2 * k4], yv[b].x, c[j][b]); }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int b = 0; b < N; b++) { a[j][b] = g5_fma2(wa[8 * j + 2 * k4 + 1], xv[b].y, a[j][b]); c[j][b] = g5_fma2(wb[8 * j + 2 * k4 + 1], yv[b].y, c[j][b]); }
        }
    }
#pragma unroll
    for (int b = 0; b < N; b++) {
        float a0 = g5_hsum(a[0][b]), a1 = g5_hsum(a[1][b]), c0 = g5_hsum(c[0][b]), c1 = g5_hsum(c[1][b]);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16); a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
        c0 += __shfl_xor_sync(0xffffffffu, c0, 16); c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
        if (lane < 16) {
            *reinterpret_cast<float2*>(&outa[(size_t)(LO + b) * ostride + 2 * lane]) = make_float2(a0, a1);
            *reinterpret_cast<float2*>(&outb[(size_t)(LO + b) * ostride + 2 * lane]) = make_float2(c0, c1);
        }
    }
}

template <int CS>
__device__ __forceinline__ void g5_send(const uint32_t (&delta)[CS], uint32_t buf_a, uint32_t bar_a, float4 v) {
#pragma unroll
    for (int d = 0; d < CS; d++) st_async_v4(buf_a + delta[d], v, bar_a + delta[d]);
}
__device__ __forceinline__ g5_f2 g5_add2(g5_f2 a, g5_f2 b) { g5_f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// sum of the CS K-slice partials of one (gate, utterance, unit quad): packed adds in a tree -- 2 (CS - 1) instructions, log2 CS deep.
// The owner warps' FP instructions queue on the FMA pipe that the other sub-batch's mat-vec saturates (profiles/r02_gru_trace.txt), so
// their COUNT and dependent depth, not their flops, set the length of a finalisation (28 scalar adds, 7 deep, before).
template <int CS>
__device__ __forceinline__ float4 g5_sum4(const float* part, int stride) {
    static_assert((CS & (CS - 1)) == 0, "power of two");
    g5_f2 lo[CS], hi[CS];
#pragma unroll
    for (int w = 0; w < CS; w++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(part + (size_t)w * stride);
        lo[w] = v.x; hi[w] = v.y;
    }
#pragma unroll
    for (int n = CS; n > 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n / 2; i++) { lo[i] = g5_add2(lo[2 * i], lo[2 * i + 1]); hi[i] = g5_add2(hi[2 * i], hi[2 * i + 1]); }
    }
    float4 s;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(s.x), "=f"(s.y) : "l"(lo[0]));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(s.z), "=f"(s.w) : "l"(hi[0]));
    return s;
}

enum { G5_BAR_P1A = 1, G5_BAR_P1B, G5_BAR_P2A, G5_BAR_P2B };

// ---------------------------------------------------------------------------------------------------------------------------------
// forward.  Sub-batch A = utterances [0, NA), B = [NA, NA + NB) of the cluster's group (NB may be 0).
// ---------------------------------------------------------------------------------------------------------------------------------
// TRACE: timeline experiment (S2S_GRU_TRACE=1, benchmarks/gru_micro.py): CTA 0 of cluster 0 writes clock64() at the events of steps
// 100..103 into p.clk -- mat-vec warps 0 and 7: slots 0..7 / 8..15, owner warps of sub-batches A / B: slots 16..19 / 20..23
template <int H, int NA, int NB, bool TRACE = false>
__global__ void __launch_bounds__(H + 128, 1)
gru5_fwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32, BG = NA + NB, NTB = H + 64, NT = H + 128;      // NTB: participants of one named barrier (mat-vec + one owner pair)
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ __align__(16) float zbuf[BG][32];
    __shared__ uint64_t bar_h[2][CS], bar_rh[2][CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    constexpr unsigned TXA = NA * 32 * 4, TXB = NB * 32 * 4;      // bytes one source CTA sends per exchange and sub-batch
    const bool owner = warp >= CS;                                 // warps CS, CS+1: finalisation; warps < CS: mat-vec

    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);
    const bool tr_on = TRACE && p.clk != nullptr && blockIdx.x == 0 && lane == (warp >= H / 32 ? 8 : 0);      // owner warps: lane 8 sends r*h
#define G5_TR(slot) do { if (TRACE) { if (tr_on && s >= 100 && s < 104) p.clk[(s - 100) * 32 + (slot)] = clock64(); } } while (0)

    for (int i = tid; i < BG * H; i += NT) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }      // Recurrent.lua:13,112
    if (!owner && lane == 0) {
        mbar_init(&bar_h[0][warp], 1); mbar_init(&bar_rh[0][warp], 1); mbar_init(&bar_h[1][warp], 1); mbar_init(&bar_rh[1][warp], 1);
        fence_mbar_init();
        mbar_expect_tx(&bar_h[0][warp], TXA); mbar_expect_tx(&bar_rh[0][warp], TXA);
        if (NB > 0) { mbar_expect_tx(&bar_h[1][warp], TXB); mbar_expect_tx(&bar_rh[1][warp], TXB); }
    }
    __syncthreads();
    cluster_sync_all();   // every CTA of the cluster is resident and has initialised its barriers / buffers

    if (!owner) {
        // =========================== mat-vec warps ===========================
        g5_f2 wz2[16], wr2[16], wh2[16];      // rows 32 crank + 2 (lane & 15) + {0, 1} of each gate, columns [32 warp + 16 (lane >> 4), +16) of the h block
        {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * crank + 2 * (lane & 15) + j) * p.ldw + 32 * warp + 16 * (lane >> 4);
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    wz2[8 * j + k / 2] = g5_pack(Wd[k], Wd[k + 1]);
                    wr2[8 * j + k / 2] = g5_pack(Wd[(size_t)H * p.ldw + k], Wd[(size_t)H * p.ldw + k + 1]);
                    wh2[8 * j + k / 2] = g5_pack(Wd[(size_t)2 * H * p.ldw + k], Wd[(size_t)2 * H * p.ldw + k + 1]);
                }
            }
        }
        const int k0 = 32 * warp + 16 * (lane >> 4);
        for (int s = 0; s < Lgrp; s++) {
            const unsigned ph = (unsigned)(s - 1) & 1u, pr = (unsigned)s & 1u;
            // phase 1, sub-batch A then B: each as soon as its source CTA's slice of h_{s-1} has landed
            const int trb = warp == 0 ? 0 : (warp == CS - 1 ? 8 : 24);      // (slots 24.. : scratch of the other warps, never read)
            if (s > 0) { mbar_wait(&bar_h[0][warp], ph); if (lane == 0) mbar_expect_tx(&bar_h[0][warp], TXA); }
            G5_TR(trb + 0);
            g5_mv2<H, 0, NA, true>(wz2, wr2, hbuf, hbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);
            if (!(p.dbg & 4)) __threadfence_block();
            g5_bar_arrive(G5_BAR_P1A, NTB);
            G5_TR(trb + 1);
            if (NB > 0) {
                if (s > 0) { mbar_wait(&bar_h[1][warp], ph); if (lane == 0) mbar_expect_tx(&bar_h[1][warp], TXB); }
                G5_TR(trb + 2);
                g5_mv2<H, NA, NB, true>(wz2, wr2, hbuf, hbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);
                if (!(p.dbg & 4)) __threadfence_block();
                g5_bar_arrive(G5_BAR_P1B, NTB);
                G5_TR(trb + 3);
            }
            // phase 2
            mbar_wait(&bar_rh[0][warp], pr); if (lane == 0) mbar_expect_tx(&bar_rh[0][warp], TXA);
            G5_TR(trb + 4);
            g5_mv<H, 0, NA>(wh2, rhbuf, k0, &part2[warp][0][0], 32, lane);
            if (!(p.dbg & 4)) __threadfence_block();
            g5_bar_arrive(G5_BAR_P2A, NTB);
            G5_TR(trb + 5);
            if (NB > 0) {
                mbar_wait(&bar_rh[1][warp], pr); if (lane == 0) mbar_expect_tx(&bar_rh[1][warp], TXB);
                G5_TR(trb + 6);
                g5_mv<H, NA, NB>(wh2, rhbuf, k0, &part2[warp][0][0], 32, lane);
                if (!(p.dbg & 4)) __threadfence_block();
                g5_bar_arrive(G5_BAR_P2B, NTB);
                G5_TR(trb + 7);
            }
        }
        if (Lgrp > 0) {      // the last h' slices have landed: nothing is in flight towards this CTA
            mbar_wait(&bar_h[0][warp], (unsigned)(Lgrp - 1) & 1u);
            if (NB > 0) mbar_wait(&bar_h[1][warp], (unsigned)(Lgrp - 1) & 1u);
        }
    } else {
        // =========================== owner warps: 64 threads per sub-batch ===========================
        const int sb = (warp - CS) >> 1, ft = tid - H - 64 * sb;
        const int N = sb ? NB : NA, LO = sb ? NA : 0;
        if (N > 0) {
            // roles: phase 1 -> (utterance, gate, quad) for ft < 16 N ; phase 2 -> (utterance, quad) for ft < 8 N
            const bool fin1 = ft < 16 * N, fin2 = ft < 8 * N;
            const int f1b = LO + (ft >> 4), f1g = (ft >> 3) & 1, f1q = ft & 7, f2b = LO + (ft >> 3), f2q = ft & 7;
            const int u1 = 32 * crank + 4 * f1q, u2 = 32 * crank + 4 * f2q;
            const int L1 = (fin1 && b0 + f1b < p.B) ? (p.lengths ? p.lengths[b0 + f1b] : p.Lmax) : 0;
            const int L2 = (fin2 && b0 + f2b < p.B) ? (p.lengths ? p.lengths[b0 + f2b] : p.Lmax) : 0;
            uint32_t delta[CS];
#pragma unroll
            for (int d = 0; d < CS; d++) delta[d] = mapa_rank(smem_u32(&hbuf[0][0]), d) - smem_u32(&hbuf[0][0]);
            const uint32_t rh_dst = smem_u32(&rhbuf[f1b][u1]), h_dst = smem_u32(&hbuf[f2b][u2]);
            const uint32_t barrh_a = smem_u32(&bar_rh[sb][crank]), barh_a = smem_u32(&bar_h[sb][crank]);    // "from CTA crank" slots
            const int bar1 = sb ? G5_BAR_P1B : G5_BAR_P1A, bar2 = sb ? G5_BAR_P2B : G5_BAR_P2A;
            // input projections do not depend on the recurrence: step s+1's values are fetched while step s runs
            auto load_xp = [&](int s, int b, int Lb, int gate, int u) -> float4 {
                if (s >= Lb) return make_float4(0.f, 0.f, 0.f, 0.f);
                const int t = rev ? Lb - 1 - s : s;
                return __ldg(reinterpret_cast<const float4*>(p.xp + ((size_t)(b0 + b) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + gate * H + u));
            };
            float4 xp1n = load_xp(0, f1b, L1, f1g, u1), xp2n = load_xp(0, f2b, L2, 2, u2);
            for (int s = 0; s < Lgrp; s++) {
                const float4 xp1 = xp1n, xp2 = xp2n;
                xp1n = load_xp(s + 1, f1b, L1, f1g, u1);
                xp2n = load_xp(s + 1, f2b, L2, 2, u2);
                const int tro = ((warp - CS) & 1) ? 28 : 16 + 4 * sb;        // first warp of each owner pair
                g5_bar_sync(bar1, NTB);
                G5_TR(tro + 0);
                if (fin1) {
                    float4 v = g5_sum4<CS>(&part1[0][f1g][f1b][4 * f1q], 2 * BG * 32);
                    v.x = sigmoid_acc(v.x + xp1.x); v.y = sigmoid_acc(v.y + xp1.y); v.z = sigmoid_acc(v.z + xp1.z); v.w = sigmoid_acc(v.w + xp1.w);   // GRU.lua:23-24
                    const bool act = s < L1;
                    const int t = rev ? L1 - 1 - s : s;
                    float* sv = p.save + (((size_t)(b0 + f1b) * p.Lmax + t) * p.ndir + dir) * 4 * H;
                    if (f1g == 0) {
                        *reinterpret_cast<float4*>(&zbuf[f1b][4 * f1q]) = v;
                        if (act) *reinterpret_cast<float4*>(sv + u1) = v;
                    } else {
                        const float4 hp = *reinterpret_cast<const float4*>(&hbuf[f1b][u1]);
                        const float4 rh = make_float4(v.x * hp.x, v.y * hp.y, v.z * hp.z, v.w * hp.w);   // GRU.lua:25
                        g5_send<CS>(delta, rh_dst, barrh_a, rh);
                        if (act) { *reinterpret_cast<float4*>(sv + H + u1) = v; *reinterpret_cast<float4*>(sv + 3 * H + u1) = rh; }
                    }
                }
                G5_TR(tro + 1);
                g5_bar_sync(bar2, NTB);                      // (also orders the z quads written above before their readers below)
                G5_TR(tro + 2);
                if (fin2) {
                    const float4 v = g5_sum4<CS>(&part2[0][f2b][4 * f2q], BG * 32);
                    const float4 hp = *reinterpret_cast<const float4*>(&hbuf[f2b][u2]);
                    float4 hn = hp;                                                                    // inactive: state frozen
                    if (s < L2) {
                        const float4 hc = make_float4(tanh_acc(v.x + xp2.x), tanh_acc(v.y + xp2.y), tanh_acc(v.z + xp2.z), tanh_acc(v.w + xp2.w));   // GRU.lua:26
                        const float4 z = *reinterpret_cast<const float4*>(&zbuf[f2b][4 * f2q]);
                        hn = make_float4((1.f - z.x) * hp.x + z.x * hc.x, (1.f - z.y) * hp.y + z.y * hc.y,
                                         (1.f - z.z) * hp.z + z.z * hc.z, (1.f - z.w) * hp.w + z.w * hc.w);     // GRU.lua:27-30
                        const int t = rev ? L2 - 1 - s : s;
                        const size_t row = (size_t)(b0 + f2b) * p.Lmax + t;
                        *reinterpret_cast<float4*>(p.save + (row * p.ndir + dir) * 4 * H + 2 * H + u2) = hc;
                        *reinterpret_cast<float4*>(p.y + row * (p.ndir * H) + dir * H + u2) = hn;
                    }
                    g5_send<CS>(delta, h_dst, barh_a, hn);
                }
                G5_TR(tro + 3);
            }
        }
    }
#undef G5_TR
    __syncthreads();
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward: per step (reverse recurrence order) and sub-batch
//   E   owner (utterance, unit quad): dh = dy + carry ; dah = dh z (1 - h~^2) ; daz = dh (h~ - h_prev) z (1-z)       -> all-gather dah, daz
//   P1  warp w: W_h[:, own]^T dah[slice w], W_z[:, own]^T daz[slice w]   ->  d(r h) ; dar = d(r h) h_prev r (1-r)    -> all-gather dar
//   P2  warp w: W_r[:, own]^T dar[slice w]                               ->  carry = dh (1-z) + d(r h) r + W_z^T daz + W_r^T dar
// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int NA, int NB>
__global__ void __launch_bounds__(H + 128, 1)
gru5_bwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32, BG = NA + NB, NTB = H + 64, NT = H + 128;      // NTB: participants of one named barrier (mat-vec + one owner pair)
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ uint64_t bar_a[2][CS], bar_r[2][CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    constexpr unsigned TXA = NA * 32 * 4, TXB = NB * 32 * 4;
    const bool owner = warp >= CS;

    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    if (!owner && lane == 0) {
        mbar_init(&bar_a[0][warp], 1); mbar_init(&bar_r[0][warp], 1); mbar_init(&bar_a[1][warp], 1); mbar_init(&bar_r[1][warp], 1);
        fence_mbar_init();
        mbar_expect_tx(&bar_a[0][warp], 2 * TXA); mbar_expect_tx(&bar_r[0][warp], TXA);
        if (NB > 0) { mbar_expect_tx(&bar_a[1][warp], 2 * TXB); mbar_expect_tx(&bar_r[1][warp], TXB); }
    }
    __syncthreads();
    cluster_sync_all();

    if (!owner) {
        // transposed recurrent weights: input unit (32 crank + lane), output units [32 warp, +32) of each gate
        g5_f2 wz2[16], wr2[16], wh2[16];
        {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * warp + 16 * (lane >> 4)) * p.ldw + 32 * crank + 2 * (lane & 15) + j;
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    wz2[8 * j + k / 2] = g5_pack(Wd[(size_t)k * p.ldw], Wd[(size_t)(k + 1) * p.ldw]);
                    wr2[8 * j + k / 2] = g5_pack(Wd[(size_t)(H + k) * p.ldw], Wd[(size_t)(H + k + 1) * p.ldw]);
                    wh2[8 * j + k / 2] = g5_pack(Wd[(size_t)(2 * H + k) * p.ldw], Wd[(size_t)(2 * H + k + 1) * p.ldw]);
                }
            }
        }
        const int k0 = 32 * warp + 16 * (lane >> 4);
        unsigned par = 0;
        for (int s = Lgrp - 1; s >= 0; s--, par ^= 1u) {                                      // RNN.lua:183
            mbar_wait(&bar_a[0][warp], par); if (lane == 0) mbar_expect_tx(&bar_a[0][warp], 2 * TXA);
            g5_mv2<H, 0, NA, false>(wh2, wz2, ahbuf, azbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);
            if (!(p.dbg & 4)) __threadfence_block();
            g5_bar_arrive(G5_BAR_P1A, NTB);
            if (NB > 0) {
                mbar_wait(&bar_a[1][warp], par); if (lane == 0) mbar_expect_tx(&bar_a[1][warp], 2 * TXB);
                g5_mv2<H, NA, NB, false>(wh2, wz2, ahbuf, azbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);
                if (!(p.dbg & 4)) __threadfence_block();
                g5_bar_arrive(G5_BAR_P1B, NTB);
            }
            mbar_wait(&bar_r[0][warp], par); if (lane == 0) mbar_expect_tx(&bar_r[0][warp], TXA);
            g5_mv<H, 0, NA>(wr2, arbuf, k0, &part2[warp][0][0], 32, lane);
            if (!(p.dbg & 4)) __threadfence_block();
            g5_bar_arrive(G5_BAR_P2A, NTB);
            if (NB > 0) {
                mbar_wait(&bar_r[1][warp], par); if (lane == 0) mbar_expect_tx(&bar_r[1][warp], TXB);
                g5_mv<H, NA, NB>(wr2, arbuf, k0, &part2[warp][0][0], 32, lane);
                if (!(p.dbg & 4)) __threadfence_block();
                g5_bar_arrive(G5_BAR_P2B, NTB);
            }
        }
    } else {
        const int sb = (warp - CS) >> 1, ft = tid - H - 64 * sb;
        const int N = sb ? NB : NA, LO = sb ? NA : 0;
        const bool own = ft < 8 * N;                             // owner of (utterance, unit quad) for the whole sequence
        if (N > 0) {
            const int ob = LO + (ft >> 3), oq = ft & 7, uo = 32 * crank + 4 * oq;
            const int Lo = (own && b0 + ob < p.B) ? (p.lengths ? p.lengths[b0 + ob] : p.Lmax) : 0;
            uint32_t delta[CS];
#pragma unroll
            for (int d = 0; d < CS; d++) delta[d] = mapa_rank(smem_u32(&ahbuf[0][0]), d) - smem_u32(&ahbuf[0][0]);
            const uint32_t ah_dst = smem_u32(&ahbuf[ob][uo]), az_dst = smem_u32(&azbuf[ob][uo]), ar_dst = smem_u32(&arbuf[ob][uo]);
            const uint32_t bara_a = smem_u32(&bar_a[sb][crank]), barr_a = smem_u32(&bar_r[sb][crank]);
            const int bar1 = sb ? G5_BAR_P1B : G5_BAR_P1A, bar2 = sb ? G5_BAR_P2B : G5_BAR_P2A;
            // saved activations / incoming gradients do not depend on the recurrence: prefetched one step ahead
            struct Pre { float4 z, r, hc, hp, dy; };
            auto load_pre = [&](int s) -> Pre {
                Pre q;
                q.z = q.r = q.hc = q.hp = q.dy = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s < 0 || s >= Lo) return q;
                const int t = rev ? Lo - 1 - s : s;
                const size_t row = (size_t)(b0 + ob) * p.Lmax + t;
                const float* sv = p.save + (row * p.ndir + dir) * 4 * H;
                q.z = ldg_stream(sv + uo); q.r = ldg_stream(sv + H + uo);
                q.hc = ldg_stream(sv + 2 * H + uo);
                if (s > 0) {                                                                      // RNN.lua:186-192
                    const int tp = rev ? t + 1 : t - 1;
                    q.hp = ldg_stream(p.y + ((size_t)(b0 + ob) * p.Lmax + tp) * (p.ndir * H) + dir * H + uo);
                }
                q.dy = ldg_stream(p.dy + row * (p.ndir * H) + dir * H + uo);
                return q;
            };
            Pre nxt = load_pre(Lgrp - 1);
            float4 carry = make_float4(0.f, 0.f, 0.f, 0.f), dhp = carry, rr = carry, hpv = carry;
            auto phase_e = [&](int s) {          // elementwise part of step s; sends dah, daz
                if (!own) return;
                const Pre cur = nxt;                                     // (prefetch discipline: see gru_seq3.cu -- the next loads go behind the sends)
                float4 dah = make_float4(0.f, 0.f, 0.f, 0.f), daz = dah;
                dhp = dah;
                if (s < Lo) {
                    const int t = rev ? Lo - 1 - s : s;
                    const size_t row = (size_t)(b0 + ob) * p.Lmax + t;
                    asm volatile("mov.b32 %0, %4; mov.b32 %1, %5; mov.b32 %2, %6; mov.b32 %3, %7;"
                                 : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w) : "f"(cur.r.x), "f"(cur.r.y), "f"(cur.r.z), "f"(cur.r.w));
                    asm volatile("mov.b32 %0, %4; mov.b32 %1, %5; mov.b32 %2, %6; mov.b32 %3, %7;"
                                 : "=f"(hpv.x), "=f"(hpv.y), "=f"(hpv.z), "=f"(hpv.w) : "f"(cur.hp.x), "f"(cur.hp.y), "f"(cur.hp.z), "f"(cur.hp.w));
#define G5_E(c)                                                                               \
                    {                                                                         \
                        const float dh = cur.dy.c + carry.c;              /* RNN.lua:193-194 */ \
                        dah.c = dh * cur.z.c * (1.f - cur.hc.c * cur.hc.c);                   \
                        daz.c = dh * (cur.hc.c - cur.hp.c) * cur.z.c * (1.f - cur.z.c);       \
                        dhp.c = dh * (1.f - cur.z.c);                                         \
                    }
                    G5_E(x) G5_E(y) G5_E(z) G5_E(w)
#undef G5_E
                    float* da = p.dA + row * (p.ndir * H3) + dir * H3;
                    *reinterpret_cast<float4*>(da + uo) = daz; *reinterpret_cast<float4*>(da + 2 * H + uo) = dah;
                    *reinterpret_cast<float4*>(p.hp_all + (row * p.ndir + dir) * H + uo) = cur.hp;
                }
                g5_send<CS>(delta, ah_dst, bara_a, dah);
                g5_send<CS>(delta, az_dst, bara_a, daz);
                nxt = load_pre(s - 1);
            };
            if (Lgrp > 0) phase_e(Lgrp - 1);
            for (int s = Lgrp - 1; s >= 0; s--) {
                g5_bar_sync(bar1, NTB);
                float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), tz = pr;
                if (own) {
                    const float4 th = g5_sum4<CS>(&part1[0][0][ob][4 * oq], 2 * BG * 32);
                    float4 dar = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s < Lo) {
                        tz = g5_sum4<CS>(&part1[0][1][ob][4 * oq], 2 * BG * 32);
                        dar = make_float4(th.x * hpv.x * rr.x * (1.f - rr.x), th.y * hpv.y * rr.y * (1.f - rr.y),
                                          th.z * hpv.z * rr.z * (1.f - rr.z), th.w * hpv.w * rr.w * (1.f - rr.w));
                        pr = make_float4(th.x * rr.x, th.y * rr.y, th.z * rr.z, th.w * rr.w);
                        const int t = rev ? Lo - 1 - s : s;
                        *reinterpret_cast<float4*>(p.dA + ((size_t)(b0 + ob) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + H + uo) = dar;
                    }
                    g5_send<CS>(delta, ar_dst, barr_a, dar);
                }
                g5_bar_sync(bar2, NTB);
                if (own && s < Lo) {
                    const float4 tr = g5_sum4<CS>(&part2[0][ob][4 * oq], BG * 32);
                    carry = make_float4(dhp.x + pr.x + tz.x + tr.x, dhp.y + pr.y + tz.y + tr.y, dhp.z + pr.z + tz.z + tr.z, dhp.w + pr.w + tz.w + tr.w);
                }
                if (s > 0) phase_e(s - 1);                   // the next step's E right behind the carry
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG, bool BWD>
static int g5_launch_geo(s2s_ctx* ctx, const GruSeqParams& p, int* max_clusters) {
    constexpr int CS = H / 32, NA = (BG + 1) / 2, NB = BG - NA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ceil_div(p.B, BG) * p.ndir);
    cfg.blockDim = dim3(H + 128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void (*kern)(const GruSeqParams);
    if constexpr (BWD) kern = gru5_bwd_kernel<H, NA, NB>; else kern = gru5_fwd_kernel<H, NA, NB>;
    if (max_clusters) {
        if (cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg) != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    if constexpr (!BWD && H == 256 && BG == 5) {
        static int trace = -1;
        if (trace < 0) { const char* e = getenv("S2S_GRU_TRACE"); trace = e ? atoi(e) : 0; }
        if (trace && !ctx->capturing) {      // timeline experiment: one traced launch, printed as cycles relative to the first event of step 100
            static long long* buf = nullptr;
            if (!buf) S2S_CUDA(cudaMalloc(&buf, 4 * 32 * sizeof(long long)));
            S2S_CUDA(cudaMemsetAsync(buf, 0, 4 * 32 * sizeof(long long), ctx->stream));
            GruSeqParams q = p; q.clk = buf;
            S2S_CUDA(cudaLaunchKernelEx(&cfg, gru5_fwd_kernel<H, NA, NB, true>, q));
            long long hb[4 * 32];
            S2S_CUDA(cudaMemcpyAsync(hb, buf, sizeof(hb), cudaMemcpyDeviceToHost, ctx->stream));
            S2S_CUDA(cudaStreamSynchronize(ctx->stream));
            const long long t0 = hb[0];
            static const char* names[24] = {"mv0 hA landed", "mv0 P1A done", "mv0 hB landed", "mv0 P1B done", "mv0 rhA landed", "mv0 P2A done", "mv0 rhB landed", "mv0 P2B done",
                                            "mv7 hA landed", "mv7 P1A done", "mv7 hB landed", "mv7 P1B done", "mv7 rhA landed", "mv7 P2A done", "mv7 rhB landed", "mv7 P2B done",
                                            "ownA bar1", "ownA fin1 sent", "ownA bar2", "ownA fin2 sent", "ownB bar1", "ownB fin1 sent", "ownB bar2", "ownB fin2 sent"};
            for (int st = 0; st < 4; st++) {
                fprintf(stderr, "[gru5 trace] step %d:", 100 + st);
                for (int e = 0; e < 24; e++) fprintf(stderr, " %s=%lld", names[e], hb[st * 32 + e] - t0);
                fprintf(stderr, "\n");
            }
            return 0;
        }
    }
    S2S_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return 0;
}

template <int H, bool BWD>
static int g5_launch_hb(s2s_ctx* ctx, const GruSeqParams& p) {
    static int cap = 0;            // co-resident clusters, queried once per process (one device per process: s2s_ctx_create)
    if (cap == 0) {
        int n = 0;
        S2S_TRY((g5_launch_geo<H, 4, BWD>(ctx, p, &n)));
        cap = n > 0 ? n : 1;
    }
    // one wave of clusters: the smallest group size for which every cluster is co-resident (a second wave would double the time).
    // The backward kernel's static shared memory limits H = 256 to groups of 7: larger batches take more than one wave.
    constexpr int BGMAX = (H == 256 && BWD) ? 7 : 8;
    int bg = 1;
    while (bg < BGMAX && p.ndir * ceil_div(p.B, bg) > cap) bg++;
    { const char* e = getenv("S2S_GRU_BG"); if (e && atoi(e) >= 1 && atoi(e) <= BGMAX) bg = atoi(e); }
    switch (bg) {
        case 1: return g5_launch_geo<H, 1, BWD>(ctx, p, nullptr);
        case 2: return g5_launch_geo<H, 2, BWD>(ctx, p, nullptr);
        case 3: return g5_launch_geo<H, 3, BWD>(ctx, p, nullptr);
        case 4: return g5_launch_geo<H, 4, BWD>(ctx, p, nullptr);
        case 5: return g5_launch_geo<H, 5, BWD>(ctx, p, nullptr);
        case 6: return g5_launch_geo<H, 6, BWD>(ctx, p, nullptr);
        case 7: return g5_launch_geo<H, 7, BWD>(ctx, p, nullptr);
        default: return g5_launch_geo<H, BGMAX, BWD>(ctx, p, nullptr);
    }
}

// launches the recurrence of one layer (all directions and utterances); H in {128, 256}
int gru_cluster5_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H) {
    if (H == 256) return backward ? g5_launch_hb<256, true>(ctx, p) : g5_launch_hb<256, false>(ctx, p);
    return backward ? g5_launch_hb<128, true>(ctx, p) : g5_launch_hb<128, false>(ctx, p);
}

}  // namespace s2s
