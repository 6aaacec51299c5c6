// gru_seq4.cu -- fourth generation of the persistent cluster GRU recurrence (nn.RNN(nn.GRU), RNN.lua:120-201, GRU.lua:22-30):
// gru_seq3.cu with the utterances of a cluster cut into sub-batches of AT MOST TWO.
//
// The step period of generation 3 is the dependent chain of ONE sub-batch (mat-vec -> named barrier -> 8-partial sum + gate math ->
// 8 st.async -> DSMEM flight, twice per step); the other sub-batch rides in its gaps.  That chain grows with the sub-batch: measured
// 1.39 us per frame-step with sub-batches of (2, 2) against 1.58 with (3, 2) and 1.72 with (3, 3), while the eight mat-vec warps are
// busy for only about a third of the period.  So a group of 5 utterances runs as THREE independent chains (2, 2, 1) and a group of 7 or
// 8 as four: every chain is the short one, and the mat-vec warps simply visit one more sub-batch per phase.
//   * sub-batch i = utterances [2i, 2i + 2) of the group, with its own per-source mbarriers and its own pair of named barriers;
//   * ONE dedicated owner warp per sub-batch (16 (utterance, gate, quad) roles per utterance in phase 1, 8 in phase 2: a sub-batch of two
//     needs exactly 32 lanes), so the CTA keeps the size of generation 3 (8 mat-vec warps + 4 owner warps);
//   * everything else (weights in registers, lane = row, warp = K-slice = source CTA, warp-uniform 16-byte state broadcasts, packed
//     fma.rn.f32x2, per-destination address deltas, backward owners keeping r / h_prev / the dh carry in registers) is generation 3's.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cluster_rnn.cuh"
#include "common.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {

typedef unsigned long long g4_f2;
__device__ __forceinline__ g4_f2 g4_pack(float a, float b) { g4_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float g4_hsum(g4_f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
__device__ __forceinline__ g4_f2 g4_fma2(g4_f2 a, g4_f2 b, g4_f2 c) { g4_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void g4_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void g4_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// out[b][lane] (b in [LO, LO+N)) = sum_{k<32} w[k] x[b][k0+k]: this lane's row over the warp's K-slice
template <int H, int LO, int N>
__device__ __forceinline__ void g4_mv(const g4_f2 (&w2)[16], const float (*x)[H], int k0, float* out, int ostride, int lane) {
    g4_f2 a[N > 0 ? N : 1];
#pragma unroll
    for (int b = 0; b < N; b++) a[b] = 0ull;
#pragma unroll
    for (int k4 = 0; k4 < 8; k4++) {
#pragma unroll
        for (int b = 0; b < N; b++) {
            const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&x[LO + b][k0 + 4 * k4]);
            a[b] = g4_fma2(w2[2 * k4], xv.x, a[b]);
            a[b] = g4_fma2(w2[2 * k4 + 1], xv.y, a[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < N; b++) out[(size_t)(LO + b) * ostride + lane] = g4_hsum(a[b]);
}
// two gates that read the same state slice: the broadcasts are shared
template <int H, int LO, int N, bool SAME>
__device__ __forceinline__ void g4_mv2(const g4_f2 (&wa)[16], const g4_f2 (&wb)[16], const float (*xa)[H], const float (*xb)[H], int k0,
                                       float* outa, float* outb, int ostride, int lane) {
    g4_f2 a[N > 0 ? N : 1], c[N > 0 ? N : 1];
#pragma unroll
    for (int b = 0; b < N; b++) { a[b] = 0ull; c[b] = 0ull; }
#pragma unroll
    for (int k4 = 0; k4 < 8; k4++) {
#pragma unroll
        for (int b = 0; b < N; b++) {
            const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&xa[LO + b][k0 + 4 * k4]);
            const ulonglong2 yv = SAME ? xv : *reinterpret_cast<const ulonglong2*>(&xb[LO + b][k0 + 4 * k4]);
            a[b] = g4_fma2(wa[2 * k4], xv.x, a[b]); a[b] = g4_fma2(wa[2 * k4 + 1], xv.y, a[b]);
            c[b] = g4_fma2(wb[2 * k4], yv.x, c[b]); c[b] = g4_fma2(wb[2 * k4 + 1], yv.y, c[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < N; b++) { outa[(size_t)(LO + b) * ostride + lane] = g4_hsum(a[b]); outb[(size_t)(LO + b) * ostride + lane] = g4_hsum(c[b]); }
}

template <int CS>
__device__ __forceinline__ void g4_send(const uint32_t (&delta)[CS], uint32_t buf_a, uint32_t bar_a, float4 v) {
#pragma unroll
    for (int d = 0; d < CS; d++) st_async_v4(buf_a + delta[d], v, bar_a + delta[d]);
}
template <int CS>
__device__ __forceinline__ float4 g4_sum4(const float* part, int stride) {
    float4 s = *reinterpret_cast<const float4*>(part);
#pragma unroll
    for (int w = 1; w < CS; w++) {
        const float4 v = *reinterpret_cast<const float4*>(part + (size_t)w * stride);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    return s;
}


constexpr int G4_MAXSB = 4;
// named barriers: 1 + sb (phase 1), 1 + G4_MAXSB + sb (phase 2)
__device__ __forceinline__ int g4_bar1(int sb) { return 1 + sb; }
__device__ __forceinline__ int g4_bar2(int sb) { return 1 + G4_MAXSB + sb; }

// ---------------------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(H + 128, 1)
gru4_fwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32, NSB = (BG + 1) / 2, NTB = H + 32, NT = H + 128;      // NTB: participants of one named barrier (mat-vec warps + one owner warp)
    static_assert(NSB <= G4_MAXSB, "at most 8 utterances per cluster");
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ __align__(16) float zbuf[BG][32];
    __shared__ uint64_t bar_h[NSB][CS], bar_rh[NSB][CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    const bool owner = warp >= CS;

    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    for (int i = tid; i < BG * H; i += NT) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }      // Recurrent.lua:13,112
    if (!owner && lane == 0) {
#pragma unroll
        for (int sb = 0; sb < NSB; sb++) {
            const unsigned tx = (unsigned)((BG - 2 * sb) < 2 ? (BG - 2 * sb) : 2) * 32 * 4;      // bytes one source CTA sends per exchange
            mbar_init(&bar_h[sb][warp], 1); mbar_init(&bar_rh[sb][warp], 1);
            fence_mbar_init();
            mbar_expect_tx(&bar_h[sb][warp], tx); mbar_expect_tx(&bar_rh[sb][warp], tx);
        }
    }
    __syncthreads();
    cluster_sync_all();   // every CTA of the cluster is resident and has initialised its barriers / buffers

    if (!owner) {
        // =========================== mat-vec warps ===========================
        g4_f2 wz2[16], wr2[16], wh2[16];      // row (32 crank + lane) of each gate, columns [32 warp, +32) of the h block
        {
            const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * crank + lane) * p.ldw + 32 * warp;
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                wz2[k / 2] = g4_pack(Wd[k], Wd[k + 1]);
                wr2[k / 2] = g4_pack(Wd[(size_t)H * p.ldw + k], Wd[(size_t)H * p.ldw + k + 1]);
                wh2[k / 2] = g4_pack(Wd[(size_t)2 * H * p.ldw + k], Wd[(size_t)2 * H * p.ldw + k + 1]);
            }
        }
        const int k0 = 32 * warp;
#define G4_P1(SB)                                                                                                                    \
        if constexpr (SB < NSB) {                                                                                                    \
            constexpr int N_ = (BG - 2 * SB) < 2 ? (BG - 2 * SB) : 2;                                                                \
            if (s > 0) { mbar_wait(&bar_h[SB][warp], ph); if (lane == 0) mbar_expect_tx(&bar_h[SB][warp], N_ * 128u); }              \
            g4_mv2<H, 2 * SB, N_, true>(wz2, wr2, hbuf, hbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);          \
            __threadfence_block();                                                                                                   \
            g4_bar_arrive(g4_bar1(SB), NTB);                                                                                         \
        }
#define G4_P2(SB)                                                                                                                    \
        if constexpr (SB < NSB) {                                                                                                    \
            constexpr int N_ = (BG - 2 * SB) < 2 ? (BG - 2 * SB) : 2;                                                                \
            mbar_wait(&bar_rh[SB][warp], pr); if (lane == 0) mbar_expect_tx(&bar_rh[SB][warp], N_ * 128u);                           \
            g4_mv<H, 2 * SB, N_>(wh2, rhbuf, k0, &part2[warp][0][0], 32, lane);                                                      \
            __threadfence_block();                                                                                                   \
            g4_bar_arrive(g4_bar2(SB), NTB);                                                                                         \
        }
        for (int s = 0; s < Lgrp; s++) {
            const unsigned ph = (unsigned)(s - 1) & 1u, pr = (unsigned)s & 1u;
            // phase 1 of every sub-batch as soon as its source CTA's slice of h_{s-1} has landed, then phase 2 likewise
            G4_P1(0) G4_P1(1) G4_P1(2) G4_P1(3)
            G4_P2(0) G4_P2(1) G4_P2(2) G4_P2(3)
        }
#undef G4_P1
#undef G4_P2
        if (Lgrp > 0) {      // the last h' slices have landed: nothing is in flight towards this CTA
#pragma unroll
            for (int sb = 0; sb < NSB; sb++) mbar_wait(&bar_h[sb][warp], (unsigned)(Lgrp - 1) & 1u);
        }
    } else if (warp - CS < NSB) {
        // =========================== owner warp of sub-batch sb: 32 lanes ===========================
        const int sb = warp - CS, ft = lane;
        const int N = (BG - 2 * sb) < 2 ? (BG - 2 * sb) : 2, LO = 2 * sb;
        // roles: phase 1 -> (utterance, gate, quad) for ft < 16 N ; phase 2 -> (utterance, quad) for ft < 8 N
        const bool fin1 = ft < 16 * N, fin2 = ft < 8 * N;
        const int f1b = LO + (ft >> 4), f1g = (ft >> 3) & 1, f1q = ft & 7, f2b = LO + (ft >> 3), f2q = ft & 7;
        const int u1 = 32 * crank + 4 * f1q, u2 = 32 * crank + 4 * f2q;
        const int L1 = (fin1 && b0 + f1b < p.B) ? (p.lengths ? p.lengths[b0 + f1b] : p.Lmax) : 0;
        const int L2 = (fin2 && b0 + f2b < p.B) ? (p.lengths ? p.lengths[b0 + f2b] : p.Lmax) : 0;
        uint32_t delta[CS];
#pragma unroll
        for (int d = 0; d < CS; d++) delta[d] = mapa_rank(smem_u32(&hbuf[0][0]), d) - smem_u32(&hbuf[0][0]);
        const uint32_t rh_dst = smem_u32(&rhbuf[fin1 ? f1b : 0][u1]), h_dst = smem_u32(&hbuf[fin2 ? f2b : 0][u2]);
        const uint32_t barrh_a = smem_u32(&bar_rh[sb][crank]), barh_a = smem_u32(&bar_h[sb][crank]);    // "from CTA crank" slots
        const int bar1 = g4_bar1(sb), bar2 = g4_bar2(sb);
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
            g4_bar_sync(bar1, NTB);
            if (fin1) {
                float4 v = g4_sum4<CS>(&part1[0][f1g][f1b][4 * f1q], 2 * BG * 32);
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
                    g4_send<CS>(delta, rh_dst, barrh_a, rh);
                    if (act) { *reinterpret_cast<float4*>(sv + H + u1) = v; *reinterpret_cast<float4*>(sv + 3 * H + u1) = rh; }
                }
            }
            g4_bar_sync(bar2, NTB);                      // (also orders the z quads written above before their readers below)
            if (fin2) {
                const float4 v = g4_sum4<CS>(&part2[0][f2b][4 * f2q], BG * 32);
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
                g4_send<CS>(delta, h_dst, barh_a, hn);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward: per step (reverse recurrence order) and sub-batch
//   E   owner (utterance, unit quad): dh = dy + carry ; dah = dh z (1 - h~^2) ; daz = dh (h~ - h_prev) z (1-z)       -> all-gather dah, daz
//   P1  warp w: W_h[:, own]^T dah[slice w], W_z[:, own]^T daz[slice w]   ->  d(r h) ; dar = d(r h) h_prev r (1-r)    -> all-gather dar
//   P2  warp w: W_r[:, own]^T dar[slice w]                               ->  carry = dh (1-z) + d(r h) r + W_z^T daz + W_r^T dar
// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(H + 128, 1)
gru4_bwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32, NSB = (BG + 1) / 2, NTB = H + 32;
    static_assert(NSB <= G4_MAXSB, "at most 8 utterances per cluster");
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ uint64_t bar_a[NSB][CS], bar_r[NSB][CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    const bool owner = warp >= CS;

    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    if (!owner && lane == 0) {
#pragma unroll
        for (int sb = 0; sb < NSB; sb++) {
            const unsigned tx = (unsigned)((BG - 2 * sb) < 2 ? (BG - 2 * sb) : 2) * 32 * 4;
            mbar_init(&bar_a[sb][warp], 1); mbar_init(&bar_r[sb][warp], 1);
            fence_mbar_init();
            mbar_expect_tx(&bar_a[sb][warp], 2 * tx); mbar_expect_tx(&bar_r[sb][warp], tx);
        }
    }
    __syncthreads();
    cluster_sync_all();

    if (!owner) {
        // transposed recurrent weights: input unit (32 crank + lane), output units [32 warp, +32) of each gate
        g4_f2 wz2[16], wr2[16], wh2[16];
        {
            const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * warp) * p.ldw + 32 * crank + lane;
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                wz2[k / 2] = g4_pack(Wd[(size_t)k * p.ldw], Wd[(size_t)(k + 1) * p.ldw]);
                wr2[k / 2] = g4_pack(Wd[(size_t)(H + k) * p.ldw], Wd[(size_t)(H + k + 1) * p.ldw]);
                wh2[k / 2] = g4_pack(Wd[(size_t)(2 * H + k) * p.ldw], Wd[(size_t)(2 * H + k + 1) * p.ldw]);
            }
        }
        const int k0 = 32 * warp;
        unsigned par = 0;
#define G4_B1(SB)                                                                                                                    \
        if constexpr (SB < NSB) {                                                                                                    \
            constexpr int N_ = (BG - 2 * SB) < 2 ? (BG - 2 * SB) : 2;                                                                \
            mbar_wait(&bar_a[SB][warp], par); if (lane == 0) mbar_expect_tx(&bar_a[SB][warp], 2 * N_ * 128u);                        \
            g4_mv2<H, 2 * SB, N_, false>(wh2, wz2, ahbuf, azbuf, k0, &part1[warp][0][0][0], &part1[warp][1][0][0], 32, lane);       \
            __threadfence_block();                                                                                                   \
            g4_bar_arrive(g4_bar1(SB), NTB);                                                                                         \
        }
#define G4_B2(SB)                                                                                                                    \
        if constexpr (SB < NSB) {                                                                                                    \
            constexpr int N_ = (BG - 2 * SB) < 2 ? (BG - 2 * SB) : 2;                                                                \
            mbar_wait(&bar_r[SB][warp], par); if (lane == 0) mbar_expect_tx(&bar_r[SB][warp], N_ * 128u);                            \
            g4_mv<H, 2 * SB, N_>(wr2, arbuf, k0, &part2[warp][0][0], 32, lane);                                                      \
            __threadfence_block();                                                                                                   \
            g4_bar_arrive(g4_bar2(SB), NTB);                                                                                         \
        }
        for (int s = Lgrp - 1; s >= 0; s--, par ^= 1u) {                                      // RNN.lua:183
            G4_B1(0) G4_B1(1) G4_B1(2) G4_B1(3)
            G4_B2(0) G4_B2(1) G4_B2(2) G4_B2(3)
        }
#undef G4_B1
#undef G4_B2
    } else if (warp - CS < NSB) {
        const int sb = warp - CS, ft = lane;
        const int N = (BG - 2 * sb) < 2 ? (BG - 2 * sb) : 2, LO = 2 * sb;
        const bool own = ft < 8 * N;                             // owner of (utterance, unit quad) for the whole sequence
        const int ob = own ? LO + (ft >> 3) : LO, oq = ft & 7, uo = 32 * crank + 4 * oq;
        const int Lo = (own && b0 + ob < p.B) ? (p.lengths ? p.lengths[b0 + ob] : p.Lmax) : 0;
        uint32_t delta[CS];
#pragma unroll
        for (int d = 0; d < CS; d++) delta[d] = mapa_rank(smem_u32(&ahbuf[0][0]), d) - smem_u32(&ahbuf[0][0]);
        const uint32_t ah_dst = smem_u32(&ahbuf[ob][uo]), az_dst = smem_u32(&azbuf[ob][uo]), ar_dst = smem_u32(&arbuf[ob][uo]);
        const uint32_t bara_a = smem_u32(&bar_a[sb][crank]), barr_a = smem_u32(&bar_r[sb][crank]);
        const int bar1 = g4_bar1(sb), bar2 = g4_bar2(sb);
        // saved activations / incoming gradients do not depend on the recurrence: prefetched one step ahead
        struct Pre { float4 z, r, hc, hp, dy; };
        auto load_pre = [&](int s) -> Pre {
            Pre q;
            q.z = q.r = q.hc = q.hp = q.dy = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < 0 || s >= Lo) return q;
            const int t = rev ? Lo - 1 - s : s;
            const size_t row = (size_t)(b0 + ob) * p.Lmax + t;
            const float* sv = p.save + (row * p.ndir + dir) * 4 * H;
            q.z = __ldg(reinterpret_cast<const float4*>(sv + uo)); q.r = __ldg(reinterpret_cast<const float4*>(sv + H + uo));
            q.hc = __ldg(reinterpret_cast<const float4*>(sv + 2 * H + uo));
            if (s > 0) {                                                                      // RNN.lua:186-192
                const int tp = rev ? t + 1 : t - 1;
                q.hp = __ldg(reinterpret_cast<const float4*>(p.y + ((size_t)(b0 + ob) * p.Lmax + tp) * (p.ndir * H) + dir * H + uo));
            }
            q.dy = __ldg(reinterpret_cast<const float4*>(p.dy + row * (p.ndir * H) + dir * H + uo));
            return q;
        };
        Pre nxt = load_pre(Lgrp - 1);
        float4 carry = make_float4(0.f, 0.f, 0.f, 0.f), dhp = carry, rr = carry, hpv = carry;
        auto phase_e = [&](int s) {          // elementwise part of step s; sends dah, daz
            if (!own) return;
            const Pre cur = nxt;
            nxt = load_pre(s - 1);
            float4 dah = make_float4(0.f, 0.f, 0.f, 0.f), daz = dah;
            dhp = dah;
            if (s < Lo) {
                const int t = rev ? Lo - 1 - s : s;
                const size_t row = (size_t)(b0 + ob) * p.Lmax + t;
                rr = cur.r; hpv = cur.hp;
#define G4_E(c)                                                                               \
                {                                                                             \
                    const float dh = cur.dy.c + carry.c;              /* RNN.lua:193-194 */   \
                    dah.c = dh * cur.z.c * (1.f - cur.hc.c * cur.hc.c);                       \
                    daz.c = dh * (cur.hc.c - cur.hp.c) * cur.z.c * (1.f - cur.z.c);           \
                    dhp.c = dh * (1.f - cur.z.c);                                             \
                }
                G4_E(x) G4_E(y) G4_E(z) G4_E(w)
#undef G4_E
                float* da = p.dA + row * (p.ndir * H3) + dir * H3;
                *reinterpret_cast<float4*>(da + uo) = daz; *reinterpret_cast<float4*>(da + 2 * H + uo) = dah;
                *reinterpret_cast<float4*>(p.hp_all + (row * p.ndir + dir) * H + uo) = cur.hp;
            }
            g4_send<CS>(delta, ah_dst, bara_a, dah);
            g4_send<CS>(delta, az_dst, bara_a, daz);
        };
        if (Lgrp > 0) phase_e(Lgrp - 1);
        for (int s = Lgrp - 1; s >= 0; s--) {
            g4_bar_sync(bar1, NTB);
            float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), tz = pr;
            if (own) {
                const float4 th = g4_sum4<CS>(&part1[0][0][ob][4 * oq], 2 * BG * 32);
                float4 dar = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s < Lo) {
                    tz = g4_sum4<CS>(&part1[0][1][ob][4 * oq], 2 * BG * 32);
                    dar = make_float4(th.x * hpv.x * rr.x * (1.f - rr.x), th.y * hpv.y * rr.y * (1.f - rr.y),
                                      th.z * hpv.z * rr.z * (1.f - rr.z), th.w * hpv.w * rr.w * (1.f - rr.w));
                    pr = make_float4(th.x * rr.x, th.y * rr.y, th.z * rr.z, th.w * rr.w);
                    const int t = rev ? Lo - 1 - s : s;
                    *reinterpret_cast<float4*>(p.dA + ((size_t)(b0 + ob) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + H + uo) = dar;
                }
                g4_send<CS>(delta, ar_dst, barr_a, dar);
            }
            g4_bar_sync(bar2, NTB);
            if (own && s < Lo) {
                const float4 tr = g4_sum4<CS>(&part2[0][ob][4 * oq], BG * 32);
                carry = make_float4(dhp.x + pr.x + tz.x + tr.x, dhp.y + pr.y + tz.y + tr.y, dhp.z + pr.z + tz.z + tr.z, dhp.w + pr.w + tz.w + tr.w);
            }
            if (s > 0) phase_e(s - 1);                   // the next step's E right behind the carry
        }
    }
    __syncthreads();
    cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG, bool BWD>
static int g4_launch_geo(s2s_ctx* ctx, const GruSeqParams& p, int* max_clusters) {
    constexpr int CS = H / 32;
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
    if constexpr (BWD) kern = gru4_bwd_kernel<H, BG>; else kern = gru4_fwd_kernel<H, BG>;
    if (max_clusters) {
        if (cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg) != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    S2S_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return 0;
}

template <int H, bool BWD>
static int g4_launch_hb(s2s_ctx* ctx, const GruSeqParams& p) {
    static int cap = 0;            // co-resident clusters, queried once per process (one device per process: s2s_ctx_create)
    if (cap == 0) {
        int n = 0;
        S2S_TRY((g4_launch_geo<H, 4, BWD>(ctx, p, &n)));
        cap = n > 0 ? n : 1;
    }
    // one wave of clusters: the smallest group size for which every cluster is co-resident (a second wave would double the time).
    // The backward kernel's static shared memory limits H = 256 to groups of 7: larger batches take more than one wave.
    constexpr int BGMAX = (H == 256 && BWD) ? 7 : 8;
    int bg = 1;
    while (bg < BGMAX && p.ndir * ceil_div(p.B, bg) > cap) bg++;
    { const char* e = getenv("S2S_GRU_BG"); if (e && atoi(e) >= 1 && atoi(e) <= BGMAX) bg = atoi(e); }
    switch (bg) {
        case 1: return g4_launch_geo<H, 1, BWD>(ctx, p, nullptr);
        case 2: return g4_launch_geo<H, 2, BWD>(ctx, p, nullptr);
        case 3: return g4_launch_geo<H, 3, BWD>(ctx, p, nullptr);
        case 4: return g4_launch_geo<H, 4, BWD>(ctx, p, nullptr);
        case 5: return g4_launch_geo<H, 5, BWD>(ctx, p, nullptr);
        case 6: return g4_launch_geo<H, 6, BWD>(ctx, p, nullptr);
        case 7: return g4_launch_geo<H, 7, BWD>(ctx, p, nullptr);
        default: return g4_launch_geo<H, BGMAX, BWD>(ctx, p, nullptr);
    }
}

// launches the recurrence of one layer (all directions and utterances); H in {128, 256}
int gru_cluster4_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H) {
    if (H == 256) return backward ? g4_launch_hb<256, true>(ctx, p) : g4_launch_hb<256, false>(ctx, p);
    return backward ? g4_launch_hb<128, true>(ctx, p) : g4_launch_hb<128, false>(ctx, p);
}

}  // namespace s2s
