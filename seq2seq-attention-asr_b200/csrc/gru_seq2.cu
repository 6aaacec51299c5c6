// gru_seq2.cu -- second generation of the persistent cluster GRU recurrence (nn.RNN(nn.GRU), RNN.lua:120-201, GRU.lua:22-30).
//
// Same cluster geometry as gru_seq.cu (a cluster of CS = H/32 CTAs owns BG utterances of one direction for all L steps, CTA c owns
// hidden units [32c, 32c+32), recurrent weights live in registers, the state is exchanged through distributed shared memory with
// st.async + mbarrier complete_tx), but the step is re-cut around the exchange instead of around the reduction:
//   * warp w of every CTA owns the K-slice [32w, 32w+32) of the state -- exactly the slice CTA w produces -- and lane = output row.
//     A warp waits only for ITS source CTA's slice (one mbarrier per source) and starts its FMAs the moment that slice lands, so the
//     skew between the 8 producers is absorbed instead of serialised behind one all-slices barrier; the own slice needs no hop.
//   * the state is read as warp-uniform 16-byte broadcasts (one wavefront each), every lane accumulates its row for all BG utterances:
//     no shuffle butterfly, no power-of-two padding of the group size (the first generation ran a second pass for the 5th utterance
//     of a group: +0.54 us per step), the cost is linear in BG.
//   * the K-slices are summed through shared memory by quad-owner threads (unit quad, utterance) that apply the gate math and send
//     their float4 straight to all peers: ONE block barrier per phase (two per step) instead of two, no staging pass.
//   * optional packed fma.rn.f32x2 over even / odd k (the mat-vec is now FMA-issue bound: 2 warps per scheduler, no shuffles).
// Backward mirrors it with the transposed weights in registers; the elementwise part, both finalisations and the dh carry of a
// (unit quad, utterance) belong to one thread for the whole sequence, so r, h_prev and the partial carries never leave its registers.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cluster_rnn.cuh"
#include "common.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {

typedef unsigned long long g2_f2;                            // packed fp32 pair
__device__ __forceinline__ g2_f2 g2_pack(float a, float b) { g2_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float g2_hsum(g2_f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
__device__ __forceinline__ g2_f2 g2_fma2(g2_f2 a, g2_f2 b, g2_f2 c) { g2_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// acc[b] += sum_{k < 32} w[k] x[b][k] for this lane's row: x = the warp's K-slice of a [BG][H] state buffer, read as broadcasts
template <int H, int BG, int PK>
struct G2Acc {
    float a[BG];
    g2_f2 a2[BG];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int b = 0; b < BG; b++) { a[b] = 0.f; a2[b] = 0ull; }
    }
    __device__ __forceinline__ void mac(const float (&w)[32], const g2_f2 (&w2)[16], const float (*x)[H], int k0) {
#pragma unroll
        for (int k4 = 0; k4 < 8; k4++) {
#pragma unroll
            for (int b = 0; b < BG; b++) {
                if (PK) {
                    const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&x[b][k0 + 4 * k4]);
                    a2[b] = g2_fma2(w2[2 * k4], xv.x, a2[b]);
                    a2[b] = g2_fma2(w2[2 * k4 + 1], xv.y, a2[b]);
                } else {
                    const float4 xv = *reinterpret_cast<const float4*>(&x[b][k0 + 4 * k4]);
                    a[b] = fmaf(w[4 * k4], xv.x, a[b]); a[b] = fmaf(w[4 * k4 + 1], xv.y, a[b]);
                    a[b] = fmaf(w[4 * k4 + 2], xv.z, a[b]); a[b] = fmaf(w[4 * k4 + 3], xv.w, a[b]);
                }
            }
        }
    }
    __device__ __forceinline__ float get(int b) const { return PK ? g2_hsum(a2[b]) : a[b]; }
};

// float4 of this thread's unit quad to the same place in every CTA of the cluster, signalling the receivers' per-source barrier
template <int CS>
__device__ __forceinline__ void g2_send(uint32_t buf_a, uint32_t bar_a, float4 v) {
#pragma unroll
    for (int d = 0; d < CS; d++) st_async_v4(mapa_rank(buf_a, d), v, mapa_rank(bar_a, d));
}
template <int CS>
__device__ __forceinline__ float4 g2_sum4(const float* part, int stride) {       // sum of the CS K-slice partials of one quad
    float4 s = *reinterpret_cast<const float4*>(part);
#pragma unroll
    for (int w = 1; w < CS; w++) {
        const float4 v = *reinterpret_cast<const float4*>(part + (size_t)w * stride);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG, int PK>
__global__ void __launch_bounds__(H, 1)
gru2_fwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32;
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ __align__(16) float zbuf[BG][32];
    __shared__ uint64_t bar_h[CS], bar_rh[CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    constexpr unsigned TXS = BG * 32 * 4;                    // bytes one source CTA sends per exchange

    // recurrent weights -> registers: row (32 crank + lane) of each gate, columns [32 warp, +32) of the h block
    float wz[32], wr[32], wh[32];           // (only one of the two forms is live per instantiation)
    g2_f2 wz2[16], wr2[16], wh2[16];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * crank + lane) * p.ldw + 32 * warp;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const float z0 = Wd[k], z1 = Wd[k + 1], r0 = Wd[(size_t)H * p.ldw + k], r1 = Wd[(size_t)H * p.ldw + k + 1];
            const float h0 = Wd[(size_t)2 * H * p.ldw + k], h1 = Wd[(size_t)2 * H * p.ldw + k + 1];
            if (PK) { wz2[k / 2] = g2_pack(z0, z1); wr2[k / 2] = g2_pack(r0, r1); wh2[k / 2] = g2_pack(h0, h1); }
            else { wz[k] = z0; wz[k + 1] = z1; wr[k] = r0; wr[k + 1] = r1; wh[k] = h0; wh[k + 1] = h1; }
        }
    }
    for (int i = tid; i < BG * H; i += H) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }      // Recurrent.lua:13,112
    if (lane == 0) {
        mbar_init(&bar_h[warp], 1); mbar_init(&bar_rh[warp], 1);
        fence_mbar_init();
        mbar_expect_tx(&bar_h[warp], TXS); mbar_expect_tx(&bar_rh[warp], TXS);
    }

    // quad-owner roles.  phase 1: tid < 16 BG -> (utterance, gate, unit quad); phase 2: tid < 8 BG -> (utterance, unit quad)
    const bool fin1 = tid < 16 * BG, fin2 = tid < 8 * BG;
    const int f1b = tid >> 4, f1g = (tid >> 3) & 1, f1q = tid & 7;
    const int f2b = tid >> 3, f2q = tid & 7;
    const int L1 = (fin1 && b0 + f1b < p.B) ? (p.lengths ? p.lengths[b0 + f1b] : p.Lmax) : 0;
    const int L2 = (fin2 && b0 + f2b < p.B) ? (p.lengths ? p.lengths[b0 + f2b] : p.Lmax) : 0;
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t h_a = smem_u32(&hbuf[0][0]), rh_a = smem_u32(&rhbuf[0][0]);
    const uint32_t barh_a = smem_u32(&bar_h[crank]), barrh_a = smem_u32(&bar_rh[crank]);       // "from CTA crank" slot, same offset in every CTA
    const int u1 = 32 * crank + 4 * f1q, u2 = 32 * crank + 4 * f2q;                           // first unit of the quad
    __syncthreads();
    cluster_sync_all();   // every CTA of the cluster is resident and has initialised its barriers / buffers

    // input projections do not depend on the recurrence: step s+1's values are fetched while step s runs
    auto load_xp = [&](int s, int b, int Lb, int gate, int u) -> float4 {
        if (s >= Lb || (p.dbg & 2)) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int t = rev ? Lb - 1 - s : s;
        return __ldg(reinterpret_cast<const float4*>(p.xp + ((size_t)(b0 + b) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + gate * H + u));
    };
    float4 xp1n = load_xp(0, f1b, L1, f1g, u1), xp2n = load_xp(0, f2b, L2, 2, u2);
    long long tck = 0;
    const bool prof = p.clk != nullptr && blockIdx.x == 0 && (tid == 0 || tid == H - 32);
    long long* clk = p.clk + (tid == 0 ? 0 : 8);
#define G2_TICK(i) do { if (prof) { const long long n_ = clock64(); clk[i] += n_ - tck; tck = n_; } } while (0)
    if (prof) tck = clock64();

    for (int s = 0; s < Lgrp; s++) {
        const float4 xp1 = xp1n, xp2 = xp2n;
        xp1n = load_xp(s + 1, f1b, L1, f1g, u1);
        xp2n = load_xp(s + 1, f2b, L2, 2, u2);

        // ---- phase 1: z, r over this warp's K-slice, as soon as its source CTA's slice of h_{s-1} has landed ---------------------
        if (s > 0) {
            mbar_wait(&bar_h[warp], (unsigned)(s - 1) & 1u);
            if (lane == 0) mbar_expect_tx(&bar_h[warp], TXS);
        }
        G2_TICK(0);
        {
            G2Acc<H, BG, PK> az, ar;
            az.zero(); ar.zero();
            if (!(p.dbg & 4)) {
                az.mac(wz, wz2, hbuf, 32 * warp);
                ar.mac(wr, wr2, hbuf, 32 * warp);
            }
#pragma unroll
            for (int b = 0; b < BG; b++) { part1[warp][0][b][lane] = az.get(b); part1[warp][1][b][lane] = ar.get(b); }
        }
        G2_TICK(1);
        __syncthreads();
        G2_TICK(2);
        if (fin1) {
            float4 v = g2_sum4<CS>(&part1[0][f1g][f1b][4 * f1q], 2 * BG * 32);
            v.x = sigmoid_acc(v.x + xp1.x); v.y = sigmoid_acc(v.y + xp1.y);                    // GRU.lua:23-24
            v.z = sigmoid_acc(v.z + xp1.z); v.w = sigmoid_acc(v.w + xp1.w);
            const bool act = s < L1 && !(p.dbg & 1);
            const int t = rev ? L1 - 1 - s : s;
            float* sv = p.save + (((size_t)(b0 + f1b) * p.Lmax + t) * p.ndir + dir) * 4 * H;
            if (f1g == 0) {
                *reinterpret_cast<float4*>(&zbuf[f1b][4 * f1q]) = v;
                if (act) *reinterpret_cast<float4*>(sv + u1) = v;
            } else {
                const float4 hp = *reinterpret_cast<const float4*>(&hbuf[f1b][u1]);
                const float4 rh = make_float4(v.x * hp.x, v.y * hp.y, v.z * hp.z, v.w * hp.w);   // GRU.lua:25
                g2_send<CS>(rh_a + (uint32_t)(f1b * H + u1) * 4u, barrh_a, rh);
                if (act) { *reinterpret_cast<float4*>(sv + H + u1) = v; *reinterpret_cast<float4*>(sv + 3 * H + u1) = rh; }
            }
        }

        // ---- phase 2: h~ over this warp's K-slice of r*h ; h' -----------------------------------------------------------------------
        G2_TICK(3);
        mbar_wait(&bar_rh[warp], (unsigned)s & 1u);
        if (lane == 0) mbar_expect_tx(&bar_rh[warp], TXS);
        G2_TICK(4);
        {
            G2Acc<H, BG, PK> ah;
            ah.zero();
            if (!(p.dbg & 4)) ah.mac(wh, wh2, rhbuf, 32 * warp);
#pragma unroll
            for (int b = 0; b < BG; b++) part2[warp][b][lane] = ah.get(b);
        }
        G2_TICK(5);
        __syncthreads();
        G2_TICK(6);
        if (fin2) {
            const float4 v = g2_sum4<CS>(&part2[0][f2b][4 * f2q], BG * 32);
            const float4 hp = *reinterpret_cast<const float4*>(&hbuf[f2b][u2]);
            float4 hn = hp;                                                                    // inactive: state frozen
            if (s < L2) {
                const float4 hc = make_float4(tanh_acc(v.x + xp2.x), tanh_acc(v.y + xp2.y), tanh_acc(v.z + xp2.z), tanh_acc(v.w + xp2.w));   // GRU.lua:26
                const float4 z = *reinterpret_cast<const float4*>(&zbuf[f2b][4 * f2q]);
                hn = make_float4((1.f - z.x) * hp.x + z.x * hc.x, (1.f - z.y) * hp.y + z.y * hc.y,
                                 (1.f - z.z) * hp.z + z.z * hc.z, (1.f - z.w) * hp.w + z.w * hc.w);     // GRU.lua:27-30
                const int t = rev ? L2 - 1 - s : s;
                const size_t row = (size_t)(b0 + f2b) * p.Lmax + t;
                if (!(p.dbg & 1)) {
                    *reinterpret_cast<float4*>(p.save + (row * p.ndir + dir) * 4 * H + 2 * H + u2) = hc;
                    *reinterpret_cast<float4*>(p.y + row * (p.ndir * H) + dir * H + u2) = hn;
                }
            }
            g2_send<CS>(h_a + (uint32_t)(f2b * H + u2) * 4u, barh_a, hn);
        }
        G2_TICK(7);
    }
#undef G2_TICK
    if (Lgrp > 0) mbar_wait(&bar_h[warp], (unsigned)(Lgrp - 1) & 1u);      // the last h' slices have landed: nothing is in flight
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward: per step (reverse recurrence order)
//   E   owner (utterance, unit quad): dh = dy + carry ; dah = dh z (1 - h~^2) ; daz = dh (h~ - h_prev) z (1-z)       -> all-gather dah, daz
//   P1  warp w: W_h[:, own]^T dah[slice w], W_z[:, own]^T daz[slice w]   ->  d(r h) ; dar = d(r h) h_prev r (1-r)    -> all-gather dar
//   P2  warp w: W_r[:, own]^T dar[slice w]                               ->  carry = dh (1-z) + d(r h) r + W_z^T daz + W_r^T dar
// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG, int PK>
__global__ void __launch_bounds__(H, 1)
gru2_bwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32;
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float part1[CS][2][BG][32];
    __shared__ __align__(16) float part2[CS][BG][32];
    __shared__ uint64_t bar_a[CS], bar_r[CS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;
    constexpr unsigned TXS = BG * 32 * 4;

    // transposed recurrent weights -> registers: input unit (32 crank + lane), output units [32 warp, +32) of each gate
    float wz[32], wr[32], wh[32];           // (only one of the two forms is live per instantiation)
    g2_f2 wz2[16], wr2[16], wh2[16];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw + (size_t)(32 * warp) * p.ldw + 32 * crank + lane;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const float z0 = Wd[(size_t)k * p.ldw], z1 = Wd[(size_t)(k + 1) * p.ldw];
            const float r0 = Wd[(size_t)(H + k) * p.ldw], r1 = Wd[(size_t)(H + k + 1) * p.ldw];
            const float h0 = Wd[(size_t)(2 * H + k) * p.ldw], h1 = Wd[(size_t)(2 * H + k + 1) * p.ldw];
            if (PK) { wz2[k / 2] = g2_pack(z0, z1); wr2[k / 2] = g2_pack(r0, r1); wh2[k / 2] = g2_pack(h0, h1); }
            else { wz[k] = z0; wz[k + 1] = z1; wr[k] = r0; wr[k + 1] = r1; wh[k] = h0; wh[k + 1] = h1; }
        }
    }
    if (lane == 0) {
        mbar_init(&bar_a[warp], 1); mbar_init(&bar_r[warp], 1);
        fence_mbar_init();
        mbar_expect_tx(&bar_a[warp], 2 * TXS); mbar_expect_tx(&bar_r[warp], TXS);
    }

    const bool own = tid < 8 * BG;                             // owner of (utterance ob, unit quad oq) for the whole sequence
    const int ob = tid >> 3, oq = tid & 7, uo = 32 * crank + 4 * oq;
    const int Lo = (own && b0 + ob < p.B) ? (p.lengths ? p.lengths[b0 + ob] : p.Lmax) : 0;
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t ah_a = smem_u32(&ahbuf[0][0]), az_a = smem_u32(&azbuf[0][0]), ar_a = smem_u32(&arbuf[0][0]);
    const uint32_t bara_a = smem_u32(&bar_a[crank]), barr_a = smem_u32(&bar_r[crank]);
    __syncthreads();
    cluster_sync_all();

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
        if (s > 0) {                                                                          // RNN.lua:186-192
            const int tp = rev ? t + 1 : t - 1;
            q.hp = __ldg(reinterpret_cast<const float4*>(p.y + ((size_t)(b0 + ob) * p.Lmax + tp) * (p.ndir * H) + dir * H + uo));
        }
        q.dy = __ldg(reinterpret_cast<const float4*>(p.dy + row * (p.ndir * H) + dir * H + uo));
        return q;
    };
    Pre nxt = load_pre(Lgrp - 1);
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);           // dE/dh flowing to the previous recurrence step
    int it = 0;
    for (int s = Lgrp - 1; s >= 0; s--, it++) {                                               // RNN.lua:183
        const unsigned par = (unsigned)it & 1u;
        // ---- E: elementwise part (owners) -------------------------------------------------------------------------------------
        float4 dhp = make_float4(0.f, 0.f, 0.f, 0.f), rr = dhp, hpv = dhp;
        const bool act = own && s < Lo;
        if (own) {
            const Pre cur = nxt;
            nxt = load_pre(s - 1);
            float4 dah = make_float4(0.f, 0.f, 0.f, 0.f), daz = dah;
            if (act) {
                const int t = rev ? Lo - 1 - s : s;
                const size_t row = (size_t)(b0 + ob) * p.Lmax + t;
                rr = cur.r; hpv = cur.hp;
#define G2_E(c)                                                                               \
                {                                                                             \
                    const float dh = cur.dy.c + carry.c;              /* RNN.lua:193-194 */   \
                    dah.c = dh * cur.z.c * (1.f - cur.hc.c * cur.hc.c);                       \
                    daz.c = dh * (cur.hc.c - cur.hp.c) * cur.z.c * (1.f - cur.z.c);           \
                    dhp.c = dh * (1.f - cur.z.c);                                             \
                }
                G2_E(x) G2_E(y) G2_E(z) G2_E(w)
#undef G2_E
                float* da = p.dA + row * (p.ndir * H3) + dir * H3;
                *reinterpret_cast<float4*>(da + uo) = daz; *reinterpret_cast<float4*>(da + 2 * H + uo) = dah;
                *reinterpret_cast<float4*>(p.hp_all + (row * p.ndir + dir) * H + uo) = cur.hp;
            }
            g2_send<CS>(ah_a + (uint32_t)(ob * H + uo) * 4u, bara_a, dah);
            g2_send<CS>(az_a + (uint32_t)(ob * H + uo) * 4u, bara_a, daz);
        }

        // ---- P1: d(r h) = W_h[:, :H]^T dah ; W_z[:, :H]^T daz -------------------------------------------------------------------
        mbar_wait(&bar_a[warp], par);
        if (lane == 0) mbar_expect_tx(&bar_a[warp], 2 * TXS);
        {
            G2Acc<H, BG, PK> ah, az;
            ah.zero(); az.zero();
            ah.mac(wh, wh2, ahbuf, 32 * warp);
            az.mac(wz, wz2, azbuf, 32 * warp);
#pragma unroll
            for (int b = 0; b < BG; b++) { part1[warp][0][b][lane] = ah.get(b); part1[warp][1][b][lane] = az.get(b); }
        }
        __syncthreads();
        float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), tz = pr;
        if (own) {
            const float4 th = g2_sum4<CS>(&part1[0][0][ob][4 * oq], 2 * BG * 32);
            float4 dar = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act) {
                tz = g2_sum4<CS>(&part1[0][1][ob][4 * oq], 2 * BG * 32);
                dar = make_float4(th.x * hpv.x * rr.x * (1.f - rr.x), th.y * hpv.y * rr.y * (1.f - rr.y),
                                  th.z * hpv.z * rr.z * (1.f - rr.z), th.w * hpv.w * rr.w * (1.f - rr.w));
                pr = make_float4(th.x * rr.x, th.y * rr.y, th.z * rr.z, th.w * rr.w);
                const int t = rev ? Lo - 1 - s : s;
                *reinterpret_cast<float4*>(p.dA + ((size_t)(b0 + ob) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + H + uo) = dar;
            }
            g2_send<CS>(ar_a + (uint32_t)(ob * H + uo) * 4u, barr_a, dar);
        }

        // ---- P2: W_r[:, :H]^T dar ; carry -----------------------------------------------------------------------------------------
        mbar_wait(&bar_r[warp], par);
        if (lane == 0) mbar_expect_tx(&bar_r[warp], TXS);
        {
            G2Acc<H, BG, PK> ar;
            ar.zero();
            ar.mac(wr, wr2, arbuf, 32 * warp);
#pragma unroll
            for (int b = 0; b < BG; b++) part2[warp][b][lane] = ar.get(b);
        }
        __syncthreads();
        if (act) {
            const float4 tr = g2_sum4<CS>(&part2[0][ob][4 * oq], BG * 32);
            carry = make_float4(dhp.x + pr.x + tz.x + tr.x, dhp.y + pr.y + tz.y + tr.y, dhp.z + pr.z + tz.z + tr.z, dhp.w + pr.w + tz.w + tr.w);
        }
        // (the owners' next E may overwrite nothing a peer still reads: ahbuf/azbuf slices are re-sent only after every CTA's P2
        //  finalisation, i.e. after every CTA has passed the block barrier behind its P1 reads)
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------------------------------
template <int H, int BG, int PK, bool BWD>
static int g2_launch_geo(s2s_ctx* ctx, const GruSeqParams& p, int* max_clusters) {
    constexpr int CS = H / 32;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ceil_div(p.B, BG) * p.ndir);
    cfg.blockDim = dim3(H);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void (*kern)(const GruSeqParams);
    if constexpr (BWD) kern = gru2_bwd_kernel<H, BG, PK>; else kern = gru2_fwd_kernel<H, BG, PK>;
    if (max_clusters) {
        if (cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg) != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    static int prof = -1;
    static long long* clk = nullptr;
    if (prof < 0) { const char* e = getenv("S2S_GRU_PROF"); prof = e ? atoi(e) : 0; }
    GruSeqParams q = p;
    if (prof && !BWD && !ctx->capturing) {
        if (!clk) S2S_CUDA(cudaMalloc(&clk, 16 * sizeof(long long)));
        S2S_CUDA(cudaMemsetAsync(clk, 0, 16 * sizeof(long long), ctx->stream));
        q.clk = clk;
    }
    S2S_CUDA(cudaLaunchKernelEx(&cfg, kern, q));
    if (q.clk) {
        long long hc[16];
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
        S2S_CUDA(cudaMemcpy(hc, clk, sizeof(hc), cudaMemcpyDeviceToHost));
        for (int w = 0; w < 2; w++)
            fprintf(stderr, "[gru2 fwd H=%d BG=%d warp %s] clocks/step: wait h %lld | mv1 %lld | bar1 %lld | fin1 %lld | wait rh %lld | mv2 %lld | bar2 %lld | fin2 %lld\n", H, BG,
                    w ? "last" : "0", hc[8 * w + 0] / p.Lmax, hc[8 * w + 1] / p.Lmax, hc[8 * w + 2] / p.Lmax, hc[8 * w + 3] / p.Lmax, hc[8 * w + 4] / p.Lmax,
                    hc[8 * w + 5] / p.Lmax, hc[8 * w + 6] / p.Lmax, hc[8 * w + 7] / p.Lmax);
    }
    return 0;
}

template <int H, int PK, bool BWD>
static int g2_launch_h(s2s_ctx* ctx, const GruSeqParams& p, int cap, int* cap_out) {
    if (cap_out) return g2_launch_geo<H, 4, PK, BWD>(ctx, p, cap_out);
    // one wave of clusters: the smallest group size for which every cluster is co-resident (a second wave would double the time).
    // The backward kernel's static shared memory limits H = 256 to groups of 7: larger batches take more than one wave.
    constexpr int BGMAX = (H == 256 && BWD) ? 7 : 8;
    int bg = 1;
    { const char* e = getenv("S2S_GRU_BGMIN"); if (e && atoi(e) >= 1 && atoi(e) <= BGMAX) bg = atoi(e); }
    while (bg < BGMAX && p.ndir * ceil_div(p.B, bg) > cap) bg++;
    { const char* e = getenv("S2S_GRU_BG"); if (e && atoi(e) >= 1 && atoi(e) <= BGMAX) bg = atoi(e); }
    switch (bg) {
        case 1: return g2_launch_geo<H, 1, PK, BWD>(ctx, p, nullptr);
        case 2: return g2_launch_geo<H, 2, PK, BWD>(ctx, p, nullptr);
        case 3: return g2_launch_geo<H, 3, PK, BWD>(ctx, p, nullptr);
        case 4: return g2_launch_geo<H, 4, PK, BWD>(ctx, p, nullptr);
        case 5: return g2_launch_geo<H, 5, PK, BWD>(ctx, p, nullptr);
        case 6: return g2_launch_geo<H, 6, PK, BWD>(ctx, p, nullptr);
        case 7: return g2_launch_geo<H, 7, PK, BWD>(ctx, p, nullptr);
        default: return g2_launch_geo<H, BGMAX, PK, BWD>(ctx, p, nullptr);
    }
}

template <int H, bool BWD>
static int g2_launch_hb(s2s_ctx* ctx, const GruSeqParams& p) {
    static int cap = 0;            // co-resident clusters, queried once per process (one context per process and device: see s2s_ctx_create)
    static int pk = -1;
    if (pk < 0) { const char* e = getenv("S2S_GRU_PK"); pk = e ? atoi(e) : 1; }
    if (cap == 0) {
        int n = 0;
        S2S_TRY((g2_launch_h<H, 1, BWD>(ctx, p, 0, &n)));
        cap = n > 0 ? n : 1;
    }
    return pk ? g2_launch_h<H, 1, BWD>(ctx, p, cap, nullptr) : g2_launch_h<H, 0, BWD>(ctx, p, cap, nullptr);
}

// launches the recurrence of one layer (all directions and utterances); H in {128, 256}
int gru_cluster2_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H) {
    if (H == 256) return backward ? g2_launch_hb<256, true>(ctx, p) : g2_launch_hb<256, false>(ctx, p);
    return backward ? g2_launch_hb<128, true>(ctx, p) : g2_launch_hb<128, false>(ctx, p);
}

}  // namespace s2s
