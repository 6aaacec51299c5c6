// gru_seq.cu -- nn.RNN(nn.GRU) over whole utterances as PERSISTENT thread-block-cluster kernels.
//
// Reference: RNN.lua:120-201 unrolls one nn.GRU clone per frame (GRU.lua:22-30):
//     z = sigmoid(W_z {h, x})   r = sigmoid(W_r {h, x})   h~ = tanh(W_h {r*h, x})   h' = (1-z) h + z h~
// with LinearZeroBias weights [H, H+Din] and concat order {prev_h, x} (no biases).
//
// B200 design:
//   * the x-columns of the three weights are time-batched into ONE GEMM per layer
//     (xp = X . W[:, H:]^T for both directions, N = ndir*3H) -- the reference never batches over time;
//   * the recurrence runs in one launch per layer: a cluster of CS = H/32 CTAs owns one group of
//     BG (4..8) utterances of one direction; CTA c owns hidden units [32c, 32c+32) and keeps its
//     3 x 32 rows of the recurrent weights in REGISTERS for all L steps (96 floats per thread);
//     h_{t-1} and r*h_{t-1} live in shared memory and are exchanged every step through distributed
//     shared memory: each CTA broadcasts its 32-unit slice with 16-byte st.async stores that signal
//     the receiver's mbarrier (complete_tx), so the exchange costs one DSMEM hop and never waits for
//     the global stores of the saved activations -- no HBM round trip for the state, no cluster-wide
//     fence, no kernel launch per step;
//   * forward/reverse directions and batch groups are independent clusters of the same launch
//     (2 dirs x 8 groups x 8 CTAs = 128 SMs at the Chorowski TIMIT batch of 32);
//   * backward mirrors it with the transposed weights in registers and emits the gate
//     pre-activation gradients dA[B,L,ndir,3H]; every weight gradient and dX are time-batched GEMMs
//     afterwards (K = B*L), instead of one rank-1 update per frame.
//   * per-utterance lengths: a step is inactive for utterance b once s >= L_b; the reverse direction
//     starts at L_b - 1, so padded batches reproduce the reference's per-utterance results.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "cluster_rnn.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {


// Geometry of one cluster.  UC hidden units per CTA (cluster of H/UC CTAs), BG utterances per cluster.
//   UC = 32: clusters of 8 (portable), butterflies of 8 rows x 4 utterances -> groups of 4 utterances are optimal
//   UC = 16: clusters of 16 (non-portable), butterflies of 4 rows x 8 utterances -> 8 utterances in one pass;
//            used when the UC = 32 layout would need more clusters than fit in one wave (B = 32: 16 > 15).
template <int H, int UC, int BG>
struct Geo {
    static constexpr int CS = H / UC;            // CTAs per cluster
    static constexpr int R1 = UC / 4;            // phase-1 rows per warp (2 UC rows over 8 warps)
    static constexpr int R2 = UC / 8;            // phase-2 rows per warp
    static constexpr int NBP = 32 / R1;          // utterances per butterfly pass (4 or 8)
    static constexpr int NH = (BG + NBP - 1) / NBP;          // passes
    static constexpr int NB0 = BG < NBP ? BG : NBP;          // utterances in pass 0
    static constexpr int NB1 = BG > NBP ? BG - NBP : 1;      // utterances in pass 1 (if any)
    static constexpr int NI = H / 128;
    static constexpr unsigned TX = BG * H * 4;   // bytes every CTA receives per exchange
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H, int UC, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_fwd_kernel(const GruSeqParams p) {
    using G = Geo<H, UC, BG>;
    constexpr int CS = G::CS, R1 = G::R1, R2 = G::R2, NBP = G::NBP, NH = G::NH, NI = G::NI;
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float stage[BG][UC];
    __shared__ float zbuf[BG][UC];
    __shared__ uint64_t barA, barB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;

    // ---- recurrent weights -> registers (coalesced: lane owns k = 4 lane + 128 i) -------------------
    const int g1 = warp >> 2;                      // phase-1 gate of this warp's rows: 0 = z, 1 = r
    float4 w1[R1][NI], w2[R2][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
#pragma unroll
        for (int r = 0; r < R1; r++) {
            const float* row = Wd + ((size_t)g1 * H + crank * UC + R1 * (warp & 3) + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w1[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
#pragma unroll
        for (int r = 0; r < R2; r++) {
            const float* row = Wd + ((size_t)2 * H + crank * UC + R2 * warp + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w2[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
    }
    for (int i = tid; i < BG * H; i += 256) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }   // Recurrent.lua:13,112
    if (tid == 0) { mbar_init(&barA, 1); mbar_init(&barB, 1); fence_mbar_init(); }

    // finalizer roles (see matvec)
    const int bbl = lane % NBP;                                 // utterance within a pass
    const int ju1 = R1 * (warp & 3) + lane / NBP;               // phase-1 unit within this CTA's slice
    const int ju2 = R2 * warp + (lane & 15) / NBP;              // phase-2 unit (lanes < 16)
    const int j1 = crank * UC + ju1, j2 = crank * UC + ju2;
    int Lf[NH];                                                 // length of the utterance this lane finalises, per pass
#pragma unroll
    for (int hf = 0; hf < NH; hf++) {
        const int bl = NBP * hf + bbl, b = b0 + bl;
        Lf[hf] = (bl < BG && b < p.B) ? (p.lengths ? p.lengths[b] : p.Lmax) : 0;
    }
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t hbuf_a = smem_u32(&hbuf[0][0]), rhbuf_a = smem_u32(&rhbuf[0][0]);
    const uint32_t barA_a = smem_u32(&barA), barB_a = smem_u32(&barB);
    cluster_sync_all();   // every CTA of the cluster is resident and has initialised its barriers / buffers

    // input projections are independent of the recurrence: step s+1's values are fetched while step s runs
    auto load_xp = [&](int s, int hf, int gate, int j) -> float {
        if (s >= Lf[hf]) return 0.f;
        const int t = rev ? Lf[hf] - 1 - s : s;
        return __ldg(p.xp + ((size_t)(b0 + NBP * hf + bbl) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + gate * H + j);
    };
    float xp1n[NH], xp2n[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) { xp1n[hf] = load_xp(0, hf, g1, j1); xp2n[hf] = lane < 16 ? load_xp(0, hf, 2, j2) : 0.f; }
    unsigned parity = 0;
    for (int s = 0; s < Lgrp; s++) {
        float xp1[NH], xp2[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            xp1[hf] = xp1n[hf]; xp2[hf] = xp2n[hf];
            xp1n[hf] = load_xp(s + 1, hf, g1, j1);
            xp2n[hf] = lane < 16 ? load_xp(s + 1, hf, 2, j2) : 0.f;
        }
        if (tid == 0 && !(p.dbg & 1)) { mbar_expect_tx(&barA, G::TX); mbar_expect_tx(&barB, G::TX); }

        // ---- phase 1: z, r ------------------------------------------------------------------
        float tot1[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) tot1[hf] = 0.f;
        if (!(p.dbg & 2)) {
            if (NH == 1) tot1[0] = matvec<H, R1, G::NB0, NBP>(w1, hbuf, 0, lane);
            else matvec_pair<H, R1, G::NB0, G::NB1, NBP>(w1, hbuf, lane, tot1[0], tot1[NH - 1]);
        }
        float g1v[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) g1v[hf] = sigmoid_acc(tot1[hf] + xp1[hf]);    // GRU.lua:23-24 (both passes in flight)
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = NBP * hf + bbl;
            if (bl < BG) {
                const bool act = s < Lf[hf];
                const int t = rev ? Lf[hf] - 1 - s : s;
                const float g = g1v[hf];
                float* sv = p.save + (((size_t)(b0 + bl) * p.Lmax + t) * p.ndir + dir) * 4 * H;
                if (g1 == 0) {
                    zbuf[bl][ju1] = g;
                } else {
                    const float rh = g * hbuf[bl][j1];                                 // GRU.lua:25
                    stage[bl][ju1] = rh;
                    if (act) sv[3 * H + j1] = rh;
                }
                if (act) sv[g1 * H + j1] = g;
            }
        }
        __syncthreads();
        if (!(p.dbg & 1)) {
            bcast_slice<H, UC, BG>(stage, rhbuf_a, barA_a, crank, warp, lane);
            mbar_wait(&barA, parity);
        }

        // ---- phase 2: h~, h' ------------------------------------------------------------------
        float tot2[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) tot2[hf] = 0.f;
        if (!(p.dbg & 2)) {
            if (NH == 1) tot2[0] = matvec<H, R2, G::NB0, NBP>(w2, rhbuf, 0, lane);
            else matvec_pair<H, R2, G::NB0, G::NB1, NBP>(w2, rhbuf, lane, tot2[0], tot2[NH - 1]);
        }
        float hcv[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) hcv[hf] = tanh_acc(tot2[hf] + xp2[hf]);       // GRU.lua:26 (both passes in flight)
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = NBP * hf + bbl;
            if (lane < 16 && bl < BG) {
                const bool act = s < Lf[hf];
                const int t = rev ? Lf[hf] - 1 - s : s;
                const float hp = hbuf[bl][j2];
                float hn = hp;                                                         // inactive: state frozen
                if (act) {
                    const float hc = hcv[hf];
                    const float z = zbuf[bl][ju2];
                    hn = (1.f - z) * hp + z * hc;                                      // GRU.lua:27-30
                    const size_t row = (size_t)(b0 + bl) * p.Lmax + t;
                    p.save[(row * p.ndir + dir) * 4 * H + 2 * H + j2] = hc;
                    p.y[row * (p.ndir * H) + dir * H + j2] = hn;
                }
                stage[bl][ju2] = hn;
            }
        }
        __syncthreads();
        if (!(p.dbg & 1)) {
            bcast_slice<H, UC, BG>(stage, hbuf_a, barB_a, crank, warp, lane);
            mbar_wait(&barB, parity);
        }
        parity ^= 1;
    }
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// forward, two software-pipelined half-batches per cluster: the utterances of a cluster are split into sub-batches
// 0 = [0, NA) and 1 = [NA, BG) with their own mbarriers; while the r*h (or h') slices of one half are in flight over
// DSMEM the CTA runs the other half's mat-vec, so most of the ~0.55 us exchange latency of each of the two dependent
// phases of a step is covered by useful work instead of a wait:
//     P1(0) send | P1(1) send | wait(0) P2(0) send | wait(1) P2(1) send | [next step] wait(0) P1(0) ...
// ---------------------------------------------------------------------------------------------
template <int UC, int NB>
__device__ __forceinline__ void bcast_rows(const float (*stage)[UC], int lo, uint32_t buf_a, uint32_t bar_a, unsigned crank, int warp, int lane, int Hh) {
    constexpr int CPB = UC / 4;
    const int CS = Hh / UC;
    for (int d = warp; d < CS; d += 8) {
        const uint32_t rbar = mapa_rank(bar_a, d);
        for (int ch = lane; ch < NB * CPB; ch += 32) {
            const int b = lo + ch / CPB, off = (ch % CPB) * 4;
            const float4 v = *reinterpret_cast<const float4*>(&stage[b][off]);
            st_async_v4(mapa_rank(buf_a + (uint32_t)(b * Hh + crank * UC + off) * 4u, d), v, rbar);
        }
    }
}

template <int H, int UC, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_fwd_pipe_kernel(const GruSeqParams p) {
    using G = Geo<H, UC, BG>;
    constexpr int CS = G::CS, R1 = G::R1, R2 = G::R2, NBP = G::NBP, NI = G::NI;
    constexpr int NA = (BG + 1) / 2, NBb = BG - NA;          // sub-batch sizes (NA <= NBP)
    static_assert(NA <= NBP && NBb >= 1, "pipelined GRU: 2 <= BG <= 2 NBP");
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float stage[BG][UC];
    __shared__ float zbuf[BG][UC];
    __shared__ uint64_t barA[2], barB[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;

    const int g1 = warp >> 2;
    float4 w1[R1][NI], w2[R2][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
#pragma unroll
        for (int r = 0; r < R1; r++) {
            const float* row = Wd + ((size_t)g1 * H + crank * UC + R1 * (warp & 3) + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w1[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
#pragma unroll
        for (int r = 0; r < R2; r++) {
            const float* row = Wd + ((size_t)2 * H + crank * UC + R2 * warp + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w2[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
    }
    for (int i = tid; i < BG * H; i += 256) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }
    if (tid == 0) { mbar_init(&barA[0], 1); mbar_init(&barA[1], 1); mbar_init(&barB[0], 1); mbar_init(&barB[1], 1); fence_mbar_init(); }

    const int bbl = lane % NBP;                                 // utterance within a sub-batch
    const int ju1 = R1 * (warp & 3) + lane / NBP;
    const int ju2 = R2 * warp + (lane & 15) / NBP;
    const int j1 = crank * UC + ju1, j2 = crank * UC + ju2;
    int Lf[2];
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
        const int bl = (hf ? NA : 0) + bbl, b = b0 + bl;
        Lf[hf] = (bbl < (hf ? NBb : NA) && b < p.B) ? (p.lengths ? p.lengths[b] : p.Lmax) : 0;
    }
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t hbuf_a = smem_u32(&hbuf[0][0]), rhbuf_a = smem_u32(&rhbuf[0][0]);
    cluster_sync_all();

    auto load_xp = [&](int s, int hf, int gate, int j) -> float {
        if (s >= Lf[hf]) return 0.f;
        const int t = rev ? Lf[hf] - 1 - s : s;
        return __ldg(p.xp + ((size_t)(b0 + (hf ? NA : 0) + bbl) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + gate * H + j);
    };
    float xp1n[2], xp2n[2];
#pragma unroll
    for (int hf = 0; hf < 2; hf++) { xp1n[hf] = load_xp(0, hf, g1, j1); xp2n[hf] = lane < 16 ? load_xp(0, hf, 2, j2) : 0.f; }

    // phase 1 / phase 2 of one sub-batch
    auto phase1 = [&](auto HF, int s, float xp1) {
        constexpr int hf = decltype(HF)::value;
        constexpr int NBH = hf ? NBb : NA, LO = hf ? NA : 0;
        const float tot = matvec<H, R1, NBH, NBP>(w1, hbuf, LO, lane);
        const float g = sigmoid_acc(tot + xp1);                                    // GRU.lua:23-24
        if (bbl < NBH) {
            const int bl = LO + bbl;
            const bool act = s < Lf[hf];
            const int t = rev ? Lf[hf] - 1 - s : s;
            float* sv = p.save + (((size_t)(b0 + bl) * p.Lmax + t) * p.ndir + dir) * 4 * H;
            if (g1 == 0) {
                zbuf[bl][ju1] = g;
            } else {
                const float rh = g * hbuf[bl][j1];                                 // GRU.lua:25
                stage[bl][ju1] = rh;
                if (act) sv[3 * H + j1] = rh;
            }
            if (act) sv[g1 * H + j1] = g;
        }
        __syncthreads();
        bcast_rows<UC, NBH>(stage, LO, rhbuf_a, smem_u32(&barA[hf]), crank, warp, lane, H);
    };
    auto phase2 = [&](auto HF, int s, float xp2) {
        constexpr int hf = decltype(HF)::value;
        constexpr int NBH = hf ? NBb : NA, LO = hf ? NA : 0;
        const float tot = matvec<H, R2, NBH, NBP>(w2, rhbuf, LO, lane);
        const float hc = tanh_acc(tot + xp2);                                      // GRU.lua:26
        if (lane < 16 && bbl < NBH) {
            const int bl = LO + bbl;
            const bool act = s < Lf[hf];
            const int t = rev ? Lf[hf] - 1 - s : s;
            const float hp = hbuf[bl][j2];
            float hn = hp;                                                         // inactive: state frozen
            if (act) {
                const float z = zbuf[bl][ju2];
                hn = (1.f - z) * hp + z * hc;                                      // GRU.lua:27-30
                const size_t row = (size_t)(b0 + bl) * p.Lmax + t;
                p.save[(row * p.ndir + dir) * 4 * H + 2 * H + j2] = hc;
                p.y[row * (p.ndir * H) + dir * H + j2] = hn;
            }
            stage[bl][ju2] = hn;
        }
        __syncthreads();
        bcast_rows<UC, NBH>(stage, LO, hbuf_a, smem_u32(&barB[hf]), crank, warp, lane, H);
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;

    unsigned parity = 0;
    for (int s = 0; s < Lgrp; s++) {
        float xp1[2], xp2[2];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            xp1[hf] = xp1n[hf]; xp2[hf] = xp2n[hf];
            xp1n[hf] = load_xp(s + 1, hf, g1, j1);
            xp2n[hf] = lane < 16 ? load_xp(s + 1, hf, 2, j2) : 0.f;
        }
        // a barrier is re-armed only after its previous phase has been observed complete by the arming thread
        if (tid == 0) { mbar_expect_tx(&barA[0], NA * H * 4); mbar_expect_tx(&barA[1], NBb * H * 4); }
        if (s > 0) mbar_wait(&barB[0], parity ^ 1);             // h' of sub-batch 0 from the previous step
        if (tid == 0) mbar_expect_tx(&barB[0], NA * H * 4);
        phase1(I0(), s, xp1[0]);
        if (s > 0) mbar_wait(&barB[1], parity ^ 1);
        if (tid == 0) mbar_expect_tx(&barB[1], NBb * H * 4);
        phase1(I1(), s, xp1[1]);
        mbar_wait(&barA[0], parity);
        phase2(I0(), s, xp2[0]);
        mbar_wait(&barA[1], parity);
        phase2(I1(), s, xp2[1]);
        parity ^= 1;
    }
    if (Lgrp > 0) { mbar_wait(&barB[0], parity ^ 1); mbar_wait(&barB[1], parity ^ 1); }
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int H, int UC, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_bwd_kernel(const GruSeqParams p) {
    using G = Geo<H, UC, BG>;
    constexpr int CS = G::CS, R1 = G::R1, R2 = G::R2, NBP = G::NBP, NH = G::NH, NI = G::NI;
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float stage_h[BG][UC], stage_z[BG][UC], stage_r[BG][UC];
    __shared__ float stash_r[BG][UC], stash_hp[BG][UC];
    __shared__ float part1[BG][UC], part2[BG][UC];
    __shared__ uint64_t barA, barB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;

    // transposed recurrent weights -> registers.  phase 1 rows (R1 per warp): warps 0-3 -> W_h^T (applied to
    // dah), warps 4-7 -> W_z^T (applied to daz); phase 2 rows (R2 per warp): W_r^T (applied to dar).  Row =
    // input unit owned by this CTA, reduction over the output units j = 4 lane + 128 i + e.
    const int g1 = warp >> 2;
    float4 w1[R1][NI], w2[R2][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
        const float* Wg = Wd + (size_t)(g1 == 0 ? 2 : 0) * H * p.ldw + crank * UC + R1 * (warp & 3);
#pragma unroll
        for (int r = 0; r < R1; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w1[r][i] = make_float4(Wg[j * p.ldw + r], Wg[(j + 1) * p.ldw + r], Wg[(j + 2) * p.ldw + r], Wg[(j + 3) * p.ldw + r]);
            }
        const float* Wr = Wd + (size_t)H * p.ldw + crank * UC + R2 * warp;
#pragma unroll
        for (int r = 0; r < R2; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w2[r][i] = make_float4(Wr[j * p.ldw + r], Wr[(j + 1) * p.ldw + r], Wr[(j + 2) * p.ldw + r], Wr[(j + 3) * p.ldw + r]);
            }
    }
    for (int i = tid; i < BG * H; i += 256) { (&ahbuf[0][0])[i] = 0.f; (&azbuf[0][0])[i] = 0.f; (&arbuf[0][0])[i] = 0.f; }
    if (tid == 0) { mbar_init(&barA, 1); mbar_init(&barB, 1); fence_mbar_init(); }

    // roles: owner (elementwise part + carry) = phase-2 finaliser: lanes < 16 -> (unit ju2, utterance lane % NBP) per pass
    const int bbl = lane % NBP;
    const int ju1 = R1 * (warp & 3) + lane / NBP;
    const int ju2 = R2 * warp + (lane & 15) / NBP;
    const int j1 = crank * UC + ju1, j2 = crank * UC + ju2;
    const bool owner_lane = lane < 16;
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

    const uint32_t ah_a = smem_u32(&ahbuf[0][0]), az_a = smem_u32(&azbuf[0][0]), ar_a = smem_u32(&arbuf[0][0]);
    const uint32_t barA_a = smem_u32(&barA), barB_a = smem_u32(&barB);
    float carry[NH];            // dE/dh flowing to the previous recurrence step (owner lanes)
#pragma unroll
    for (int hf = 0; hf < NH; hf++) carry[hf] = 0.f;
    cluster_sync_all();

    // saved activations / incoming gradients do not depend on the recurrence: prefetch one step ahead
    struct Pre { float z, r, hc, hp, dy; };
    auto load_pre = [&](int s, int hf) -> Pre {
        Pre q = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (!owner_lane || s < 0 || s >= Lf[hf]) return q;
        const int t = rev ? Lf[hf] - 1 - s : s;
        const int b = b0 + NBP * hf + bbl;
        const size_t row = (size_t)b * p.Lmax + t;
        const float* sv = p.save + (row * p.ndir + dir) * 4 * H;
        q.z = __ldg(sv + j2); q.r = __ldg(sv + H + j2); q.hc = __ldg(sv + 2 * H + j2);
        if (s > 0) {                                                                  // RNN.lua:186-192
            const int tp = rev ? t + 1 : t - 1;
            q.hp = __ldg(p.y + ((size_t)b * p.Lmax + tp) * (p.ndir * H) + dir * H + j2);
        }
        q.dy = __ldg(p.dy + row * (p.ndir * H) + dir * H + j2);
        return q;
    };
    Pre nxt[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) nxt[hf] = load_pre(Lgrp - 1, hf);
    unsigned parity = 0;
    for (int s = Lgrp - 1; s >= 0; s--) {                                           // RNN.lua:183
        float dhp_part[NH];
        if (tid == 0 && !(p.dbg & 1)) { mbar_expect_tx(&barA, 2 * G::TX); mbar_expect_tx(&barB, G::TX); }
        // ---- elementwise part (owners) ---------------------------------------------------------
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const Pre cur = nxt[hf];
            nxt[hf] = load_pre(s - 1, hf);
            dhp_part[hf] = 0.f;
            const int bl = NBP * hf + bbl;
            if (owner_lane && bl < BG) {
                float dah = 0.f, daz = 0.f;
                if (s < Lf[hf]) {
                    const int t = rev ? Lf[hf] - 1 - s : s;
                    const size_t row = (size_t)(b0 + bl) * p.Lmax + t;
                    const float z = cur.z, r = cur.r, hc = cur.hc, hp = cur.hp;
                    const float dh = cur.dy + carry[hf];                              // RNN.lua:193-194
                    dah = dh * z * (1.f - hc * hc);
                    daz = dh * (hc - hp) * z * (1.f - z);
                    dhp_part[hf] = dh * (1.f - z);
                    float* da = p.dA + row * (p.ndir * H3) + dir * H3;
                    da[j2] = daz; da[2 * H + j2] = dah;
                    p.hp_all[(row * p.ndir + dir) * H + j2] = hp;
                    stash_r[bl][ju2] = r; stash_hp[bl][ju2] = hp;
                }
                stage_h[bl][ju2] = dah; stage_z[bl][ju2] = daz;
            }
        }
        __syncthreads();
        if (!(p.dbg & 1)) {
            bcast_slice<H, UC, BG>(stage_h, ah_a, barA_a, crank, warp, lane);
            bcast_slice<H, UC, BG>(stage_z, az_a, barA_a, crank, warp, lane);
            mbar_wait(&barA, parity);
        }

        // ---- phase 1: d(r*h) = W_h[:, :H]^T dah ; W_z[:, :H]^T daz -------------------------------
        const float (*src1)[H] = g1 == 0 ? ahbuf : azbuf;
        float tot1[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) tot1[hf] = 0.f;
        if (!(p.dbg & 2)) {   // (the paired variant measured slower here: register pressure)
            tot1[0] = matvec<H, R1, G::NB0, NBP>(w1, src1, 0, lane);
            if (NH > 1) tot1[NH - 1] = matvec<H, R1, G::NB1, NBP>(w1, src1, NBP, lane);
        }
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const float tot = tot1[hf];
            const int bl = NBP * hf + bbl;
            if (bl < BG) {
                const bool act = s < Lf[hf];
                const int t = rev ? Lf[hf] - 1 - s : s;
                if (g1 == 0) {
                    float dar = 0.f, pr = 0.f;
                    if (act) {
                        const float r = stash_r[bl][ju1], hp = stash_hp[bl][ju1];
                        dar = tot * hp * r * (1.f - r);
                        pr = tot * r;
                        p.dA[((size_t)(b0 + bl) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + H + j1] = dar;
                    }
                    part1[bl][ju1] = pr;
                    stage_r[bl][ju1] = dar;
                } else {
                    part2[bl][ju1] = act ? tot : 0.f;
                }
            }
        }
        __syncthreads();
        if (!(p.dbg & 1)) {
            bcast_slice<H, UC, BG>(stage_r, ar_a, barB_a, crank, warp, lane);
            mbar_wait(&barB, parity);
        }

        // ---- phase 2: W_r[:, :H]^T dar ; carry ---------------------------------------------------
        float tot2[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) tot2[hf] = 0.f;
        if (!(p.dbg & 2)) {   // (the paired variant measured slower here: register pressure)
            tot2[0] = matvec<H, R2, G::NB0, NBP>(w2, arbuf, 0, lane);
            if (NH > 1) tot2[NH - 1] = matvec<H, R2, G::NB1, NBP>(w2, arbuf, NBP, lane);
        }
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = NBP * hf + bbl;
            if (owner_lane && bl < BG && s < Lf[hf]) carry[hf] = dhp_part[hf] + part1[bl][ju2] + part2[bl][ju2] + tot2[hf];
        }
        parity ^= 1;
    }
    cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------
// backward, two software-pipelined half-batches per cluster (see gru_seq_fwd_pipe_kernel):
//     E(0) send | E(1) send | wait P1(0) send | wait P1(1) send | wait P2(0) | wait P2(1)      (P2 leaves the carry in registers)
// ---------------------------------------------------------------------------------------------
template <int H, int UC, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_bwd_pipe_kernel(const GruSeqParams p) {
    using G = Geo<H, UC, BG>;
    constexpr int CS = G::CS, R1 = G::R1, R2 = G::R2, NBP = G::NBP, NI = G::NI;
    constexpr int NA = (BG + 1) / 2, NBb = BG - NA;
    static_assert(NA <= NBP && NBb >= 1, "pipelined GRU: 2 <= BG <= 2 NBP");
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float stage_h[BG][UC], stage_z[BG][UC], stage_r[BG][UC];
    __shared__ float stash_r[BG][UC], stash_hp[BG][UC];
    __shared__ float part1[BG][UC], part2[BG][UC];
    __shared__ uint64_t barA[2], barB[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;

    const int g1 = warp >> 2;
    float4 w1[R1][NI], w2[R2][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
        const float* Wg = Wd + (size_t)(g1 == 0 ? 2 : 0) * H * p.ldw + crank * UC + R1 * (warp & 3);
#pragma unroll
        for (int r = 0; r < R1; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w1[r][i] = make_float4(Wg[j * p.ldw + r], Wg[(j + 1) * p.ldw + r], Wg[(j + 2) * p.ldw + r], Wg[(j + 3) * p.ldw + r]);
            }
        const float* Wr = Wd + (size_t)H * p.ldw + crank * UC + R2 * warp;
#pragma unroll
        for (int r = 0; r < R2; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w2[r][i] = make_float4(Wr[j * p.ldw + r], Wr[(j + 1) * p.ldw + r], Wr[(j + 2) * p.ldw + r], Wr[(j + 3) * p.ldw + r]);
            }
    }
    for (int i = tid; i < BG * H; i += 256) { (&ahbuf[0][0])[i] = 0.f; (&azbuf[0][0])[i] = 0.f; (&arbuf[0][0])[i] = 0.f; }
    if (tid == 0) { mbar_init(&barA[0], 1); mbar_init(&barA[1], 1); mbar_init(&barB[0], 1); mbar_init(&barB[1], 1); fence_mbar_init(); }

    const int bbl = lane % NBP;
    const int ju1 = R1 * (warp & 3) + lane / NBP;
    const int ju2 = R2 * warp + (lane & 15) / NBP;
    const int j1 = crank * UC + ju1, j2 = crank * UC + ju2;
    const bool owner_lane = lane < 16;
    int Lf[2];
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
        const int bl = (hf ? NA : 0) + bbl, b = b0 + bl;
        Lf[hf] = (bbl < (hf ? NBb : NA) && b < p.B) ? (p.lengths ? p.lengths[b] : p.Lmax) : 0;
    }
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t ah_a = smem_u32(&ahbuf[0][0]), az_a = smem_u32(&azbuf[0][0]), ar_a = smem_u32(&arbuf[0][0]);
    float carry[2] = {0.f, 0.f};            // dE/dh flowing to the previous recurrence step (owner lanes)
    float dhp_part[2] = {0.f, 0.f};
    cluster_sync_all();

    struct Pre { float z, r, hc, hp, dy; };
    auto load_pre = [&](int s, int hf) -> Pre {
        Pre q = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (!owner_lane || s < 0 || s >= Lf[hf]) return q;
        const int t = rev ? Lf[hf] - 1 - s : s;
        const int b = b0 + (hf ? NA : 0) + bbl;
        const size_t row = (size_t)b * p.Lmax + t;
        const float* sv = p.save + (row * p.ndir + dir) * 4 * H;
        q.z = __ldg(sv + j2); q.r = __ldg(sv + H + j2); q.hc = __ldg(sv + 2 * H + j2);
        if (s > 0) {                                                                  // RNN.lua:186-192
            const int tp = rev ? t + 1 : t - 1;
            q.hp = __ldg(p.y + ((size_t)b * p.Lmax + tp) * (p.ndir * H) + dir * H + j2);
        }
        q.dy = __ldg(p.dy + row * (p.ndir * H) + dir * H + j2);
        return q;
    };
    Pre nxt[2];
#pragma unroll
    for (int hf = 0; hf < 2; hf++) nxt[hf] = load_pre(Lgrp - 1, hf);

    auto elementwise = [&](auto HF, int s) {
        constexpr int hf = decltype(HF)::value;
        constexpr int NBH = hf ? NBb : NA, LO = hf ? NA : 0;
        const Pre cur = nxt[hf];
        nxt[hf] = load_pre(s - 1, hf);
        dhp_part[hf] = 0.f;
        if (owner_lane && bbl < NBH) {
            const int bl = LO + bbl;
            float dah = 0.f, daz = 0.f;
            if (s < Lf[hf]) {
                const int t = rev ? Lf[hf] - 1 - s : s;
                const size_t row = (size_t)(b0 + bl) * p.Lmax + t;
                const float z = cur.z, r = cur.r, hc = cur.hc, hp = cur.hp;
                const float dh = cur.dy + carry[hf];                              // RNN.lua:193-194
                dah = dh * z * (1.f - hc * hc);
                daz = dh * (hc - hp) * z * (1.f - z);
                dhp_part[hf] = dh * (1.f - z);
                float* da = p.dA + row * (p.ndir * H3) + dir * H3;
                da[j2] = daz; da[2 * H + j2] = dah;
                p.hp_all[(row * p.ndir + dir) * H + j2] = hp;
                stash_r[bl][ju2] = r; stash_hp[bl][ju2] = hp;
            }
            stage_h[bl][ju2] = dah; stage_z[bl][ju2] = daz;
        }
        __syncthreads();
        bcast_rows<UC, NBH>(stage_h, LO, ah_a, smem_u32(&barA[hf]), crank, warp, lane, H);
        bcast_rows<UC, NBH>(stage_z, LO, az_a, smem_u32(&barA[hf]), crank, warp, lane, H);
    };
    auto phase1 = [&](auto HF, int s) {
        constexpr int hf = decltype(HF)::value;
        constexpr int NBH = hf ? NBb : NA, LO = hf ? NA : 0;
        const float (*src1)[H] = g1 == 0 ? ahbuf : azbuf;
        const float tot = matvec<H, R1, NBH, NBP>(w1, src1, LO, lane);
        if (bbl < NBH) {
            const int bl = LO + bbl;
            const bool act = s < Lf[hf];
            const int t = rev ? Lf[hf] - 1 - s : s;
            if (g1 == 0) {
                float dar = 0.f, pr = 0.f;
                if (act) {
                    const float r = stash_r[bl][ju1], hp = stash_hp[bl][ju1];
                    dar = tot * hp * r * (1.f - r);
                    pr = tot * r;
                    p.dA[((size_t)(b0 + bl) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + H + j1] = dar;
                }
                part1[bl][ju1] = pr;
                stage_r[bl][ju1] = dar;
            } else {
                part2[bl][ju1] = act ? tot : 0.f;
            }
        }
        __syncthreads();
        bcast_rows<UC, NBH>(stage_r, LO, ar_a, smem_u32(&barB[hf]), crank, warp, lane, H);
    };
    auto phase2 = [&](auto HF, int s) {
        constexpr int hf = decltype(HF)::value;
        constexpr int NBH = hf ? NBb : NA, LO = hf ? NA : 0;
        const float tot = matvec<H, R2, NBH, NBP>(w2, arbuf, LO, lane);
        const int bl = LO + bbl;
        if (owner_lane && bbl < NBH && s < Lf[hf]) carry[hf] = dhp_part[hf] + part1[bl][ju2] + part2[bl][ju2] + tot;
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;

    unsigned parity = 0;
    for (int s = Lgrp - 1; s >= 0; s--) {                                           // RNN.lua:183
        if (tid == 0) {
            mbar_expect_tx(&barA[0], 2 * NA * H * 4); mbar_expect_tx(&barA[1], 2 * NBb * H * 4);
            mbar_expect_tx(&barB[0], NA * H * 4); mbar_expect_tx(&barB[1], NBb * H * 4);
        }
        elementwise(I0(), s);
        elementwise(I1(), s);
        mbar_wait(&barA[0], parity);
        phase1(I0(), s);
        mbar_wait(&barA[1], parity);
        phase1(I1(), s);
        mbar_wait(&barB[0], parity);
        phase2(I0(), s);
        mbar_wait(&barB[1], parity);
        phase2(I1(), s);
        parity ^= 1;
    }
    cluster_sync_all();
}

// rows t >= L_b of a [B, Lmax, W] tensor := 0 (padding must not leak NaNs into the time-batched GEMMs).  The tail of an
// utterance is one contiguous span; a few CTAs per utterance stream zeros over it (and exit at once when there is none).
// blockIdx.z selects one of up to two buffers (the two padded outputs of a recurrence call are cleared by ONE launch)
__global__ void zero_tail_rows_kernel(float* __restrict__ x0, int W0, float* __restrict__ x1, int W1, const int* __restrict__ lengths, int Lmax) {
    const int b = blockIdx.y;
    float* x = blockIdx.z ? x1 : x0;
    const int W = blockIdx.z ? W1 : W0;
    const int len = min(max(lengths[b], 0), Lmax);
    const size_t n = (size_t)(Lmax - len) * W;
    float* r = x + ((size_t)b * Lmax + len) * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) r[i] = 0.f;
}
int zero_tail_rows2(s2s_ctx* ctx, float* x0, int W0, float* x1, int W1, const int* lengths, int B, int Lmax) {
    if (!lengths) return 0;
    zero_tail_rows_kernel<<<dim3(8, B, 2), 256, 0, ctx->stream>>>(x0, W0, x1, W1, lengths, Lmax);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
int zero_tail_rows(s2s_ctx* ctx, float* x, const int* lengths, int B, int Lmax, int W) {
    if (!lengths) return 0;
    zero_tail_rows_kernel<<<dim3(8, B, 1), 256, 0, ctx->stream>>>(x, W, nullptr, 0, lengths, Lmax);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

template <int H, int UC, int BG>
static int launch_cluster_geo(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int* max_clusters) {
    constexpr int CS = H / UC;
    const int ngroups = ceil_div(p.B, BG);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ngroups * p.ndir);
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
            if (backward) S2S_CUDA(cudaFuncSetAttribute(gru_seq_bwd_kernel<H, UC, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            else S2S_CUDA(cudaFuncSetAttribute(gru_seq_fwd_kernel<H, UC, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            set[backward] = true;
        }
    }
    if (max_clusters) {   // occupancy query only
        cudaError_t e = backward ? cudaOccupancyMaxActiveClusters(max_clusters, gru_seq_bwd_kernel<H, UC, BG>, &cfg)
                                 : cudaOccupancyMaxActiveClusters(max_clusters, gru_seq_fwd_kernel<H, UC, BG>, &cfg);
        if (e != cudaSuccess) { *max_clusters = 0; cudaGetLastError(); }
        return 0;
    }
    static int pipe = -1;
    if (pipe < 0) { const char* e = getenv("S2S_GRU_PIPE"); pipe = e ? atoi(e) : 0; }
    prof_begin(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD);
    static int pipeb = -1;
    // measured (B=32, L=300, H=256, both directions): forward 2.30 vs 2.25 us/step (the extra butterflies and barriers of two
    // sequential half-batches cost what the hidden exchange latency saves: opt-in), backward 2.83 vs 3.04 us/step (default)
    if (pipeb < 0) { const char* e = getenv("S2S_GRU_PIPE_BWD"); pipeb = e ? atoi(e) : 1; }
    if (backward && pipeb) {
        if (CS > 8) {
            static bool setb = false;
            if (!setb) { S2S_CUDA(cudaFuncSetAttribute(gru_seq_bwd_pipe_kernel<H, UC, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); setb = true; }
        }
        S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_bwd_pipe_kernel<H, UC, BG>, p));
    } else if (backward) S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_bwd_kernel<H, UC, BG>, p));
    else if (pipe) {
        if (CS > 8) {
            static bool setp = false;
            if (!setp) { S2S_CUDA(cudaFuncSetAttribute(gru_seq_fwd_pipe_kernel<H, UC, BG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); setp = true; }
        }
        S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_fwd_pipe_kernel<H, UC, BG>, p));
    } else S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_fwd_kernel<H, UC, BG>, p));
    {   // algorithmic bytes per launch: fwd reads xp (3H) and writes y (H) + save (4H) per direction;
        // bwd reads save z,r,h~ (3H) + h_prev (H) + dy (H) and writes dA (3H) + h_prev (H)
        const double per = backward ? 9.0 * H : 8.0 * H;
        prof_end(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD, 4.0 * p.B * p.Lmax * p.ndir * per);
    }
    ctx->kcount[S2S_KC_GRU_CLUSTER]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// Cluster geometry: a cluster that does not fit in the first wave runs after the others and doubles the time of this
// latency-bound kernel, so the launch is sized to ONE wave.  Co-resident cluster counts are queried once (15 clusters
// of 8 CTAs / 8 clusters of 16 CTAs on a 148-SM B200).  Preference: groups of 4 utterances on clusters of 8; else
// 8 utterances on clusters of 16 (one butterfly pass); else larger groups on clusters of 8 (two passes).
template <int H>
static int launch_cluster(s2s_ctx* ctx, bool backward, const GruSeqParams& p) {
    // S2S_GRU_GEN: 5 = generation 3 with two units per lane over half the K-slice (gru_seq5.cu); 4 = sub-batches of at most two utterances
    // (gru_seq4.cu, a negative result); 3 = warp-specialised, two pipelined sub-batches (gru_seq3.cu); 2 = gru_seq2.cu; 1 = first generation
    static int gen = -1;
    if (gen < 0) { const char* e = getenv("S2S_GRU_GEN"); gen = e ? atoi(e) : 3; }
    if (gen >= 2) {
        prof_begin(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD);
        static int gen_bwd = -1;       // S2S_GRU_GEN_BWD: generation of the backward kernel alone (default: same as S2S_GRU_GEN)
        if (gen_bwd < 0) { const char* e = getenv("S2S_GRU_GEN_BWD"); gen_bwd = e ? atoi(e) : 0; }
        // default: generation 3, except the backward kernel at H = 256, where generation 5's mapping (half the shared-memory reads of the
        // two-buffer phase-1 product) is 6% faster once the prefetch discipline removed the larger stall (1.68 vs 1.78 us per frame-step;
        // at H = 128 generation 3 is the faster one: 0.85 vs 0.88)
        const int g = backward ? (gen_bwd >= 2 ? gen_bwd : (gen == 3 && H == 256 ? 5 : gen)) : gen;
        if (g >= 5) S2S_TRY(gru_cluster5_launch(ctx, backward, p, H));
        else if (g == 4) S2S_TRY(gru_cluster4_launch(ctx, backward, p, H));
        else if (g == 3) S2S_TRY(gru_cluster3_launch(ctx, backward, p, H));
        else S2S_TRY(gru_cluster2_launch(ctx, backward, p, H));
        prof_end(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD, 4.0 * p.B * p.Lmax * p.ndir * (backward ? 9.0 : 8.0) * H);
        ctx->kcount[S2S_KC_GRU_CLUSTER]++;
        S2S_LAUNCH_CHECK(ctx);
        return 0;
    }
    static int cap8[2] = {0, 0}, cap16[2] = {-1, -1};
    if (cap8[backward] == 0) {
        int n = 0;
        S2S_TRY((launch_cluster_geo<H, 32, 4>(ctx, backward, p, &n)));
        cap8[backward] = n > 0 ? n : 1;
    }
    int bg = 4, uc = 32;
    { const char* e = getenv("S2S_GRU_UC"); if (e) uc = atoi(e) == 16 ? -16 : -32; }   // experiments: force a geometry
    if (p.ndir * ceil_div(p.B, 4) > cap8[backward] && uc == -16) {   // measured: no faster than two-pass groups on clusters of 8 -> opt-in only
        if (H == 256) {
            if (cap16[backward] < 0) {
                int n = 0;
                S2S_TRY((launch_cluster_geo<256, 16, 8>(ctx, backward, p, &n)));
                cap16[backward] = n;
            }
            if (p.ndir * ceil_div(p.B, 8) <= cap16[backward] || uc == -16) uc = 16;
        }
    }
    if (uc == 16 && H == 256) {
        bg = 5;
        while (bg < 8 && p.ndir * ceil_div(p.B, bg) > cap16[backward]) bg++;
        switch (bg) {
            case 5: return launch_cluster_geo<256, 16, 5>(ctx, backward, p, nullptr);
            case 6: return launch_cluster_geo<256, 16, 6>(ctx, backward, p, nullptr);
            case 7: return launch_cluster_geo<256, 16, 7>(ctx, backward, p, nullptr);
            default: return launch_cluster_geo<256, 16, 8>(ctx, backward, p, nullptr);
        }
    }
    while (bg < 8 && p.ndir * ceil_div(p.B, bg) > cap8[backward]) bg++;
    { const char* e = getenv("S2S_GRU_BG"); if (e && atoi(e) >= 4 && atoi(e) <= 8) bg = atoi(e); }
    switch (bg) {
        case 4: return launch_cluster_geo<H, 32, 4>(ctx, backward, p, nullptr);
        case 5: return launch_cluster_geo<H, 32, 5>(ctx, backward, p, nullptr);
        case 6: return launch_cluster_geo<H, 32, 6>(ctx, backward, p, nullptr);
        case 7: return launch_cluster_geo<H, 32, 7>(ctx, backward, p, nullptr);
        default: return launch_cluster_geo<H, 32, 8>(ctx, backward, p, nullptr);
    }
}

int gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                    const int* lengths, int B, int Lmax, float* y, float* save) {
    S2S_REQUIRE(H == 128 || H == 256, "gru_seq: hidden size %d not supported by the cluster kernel (128 or 256)", H);
    S2S_REQUIRE(ndir == 1 || ndir == 2, "gru_seq: ndir must be 1 or 2");
    const int ldw = H + Din, N = ndir * 3 * H;
    float* xp;
    S2S_ALLOC(xp, ctx->arena, float, (size_t)B * Lmax * N);
    // time-batched input projections for all gates and both directions (LinearZeroBias.lua:42, x columns)
    S2S_TRY(gemm_f32(ctx, false, true, B * Lmax, N, Din, 1.f, x, ldx, W + H, ldw, 0.f, xp, N));
    if (lengths) {
        S2S_TRY(zero_tail_rows2(ctx, y, ndir * H, save, ndir * 4 * H, lengths, B, Lmax));
    }
    GruSeqParams p = {};
    p.W = W; p.ldw = ldw; p.xp = xp; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.ndir = ndir; p.reverse0 = reverse;
    p.y = y; p.save = save;
    { const char* e = getenv("S2S_GRU_DBG"); p.dbg = e ? atoi(e) : 0; }
    if (H == 128) return launch_cluster<128>(ctx, false, p);
    return launch_cluster<256>(ctx, false, p);
}

int gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                     const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx, bool defer_wgrad) {
    S2S_REQUIRE(H == 128 || H == 256, "gru_seq: hidden size %d not supported by the cluster kernel (128 or 256)", H);
    S2S_REQUIRE(ndir == 1 || ndir == 2, "gru_seq: ndir must be 1 or 2");
    const int ldw = H + Din, N = ndir * 3 * H, BL = B * Lmax;
    float *dA, *hp_all;
    S2S_ALLOC(dA, ctx->arena, float, (size_t)BL * N);
    S2S_ALLOC(hp_all, ctx->arena, float, (size_t)BL * ndir * H);
    if (lengths) {
        S2S_TRY(zero_tail_rows2(ctx, dA, N, hp_all, ndir * H, lengths, B, Lmax));
    }
    GruSeqParams p = {};
    p.W = W; p.ldw = ldw; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.ndir = ndir; p.reverse0 = reverse;
    p.y = const_cast<float*>(y); p.save = const_cast<float*>(save); p.dy = dy; p.dA = dA; p.hp_all = hp_all;
    { const char* e = getenv("S2S_GRU_DBG"); p.dbg = e ? atoi(e) : 0; }
    if (H == 128) S2S_TRY(launch_cluster<128>(ctx, true, p));
    else S2S_TRY(launch_cluster<256>(ctx, true, p));
    // dX = dA . W[:, H:]   (LinearZeroBias.lua:50-65, x columns), both directions summed as nngraph does: the next layer's recurrence
    // waits for it, so it goes first
    if (dx) S2S_TRY(gemm_f32(ctx, false, false, BL, Din, N, 1.f, dA, N, W + H, ldw, 0.f, dx, Din));
    // time-batched weight gradients (K = B*L) instead of one rank-1 update per frame (LinearZeroBias.lua:67-74).  Nothing downstream
    // reads them before the gradient step: with defer_wgrad they run on a low-priority side stream with a persistent grid limited
    // to the SMs the cluster kernels leave idle, under the NEXT layer's recurrence (S2S_OVERLAP=1).
    static int overlap = -1;
    if (overlap < 0) { const char* e = getenv("S2S_OVERLAP"); overlap = e ? atoi(e) : 1; }
    const bool fork = defer_wgrad && overlap && ctx->side[1] && ctx->stream != ctx->side[1];
    cudaStream_t main_stream = ctx->stream;
    struct SwapGuard {      // whatever path leaves this function, the context's stream and GEMM grid limit are restored
        s2s_ctx* c; cudaStream_t s; bool on;
        ~SwapGuard() { if (on) { c->stream = s; c->gemm_sm_limit = 0; } }
    } guard{ctx, main_stream, false};
    if (fork) {
        S2S_CUDA(cudaEventRecord(ctx->ev[2], main_stream));
        S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev[2], 0));
        guard.on = true;
        ctx->stream = ctx->side[1];
        ctx->gemm_sm_limit = ctx->sm_count - 112 > 16 ? ctx->sm_count - 112 : 0;
    }
    int rc = 0;
    {
        const int sk = 8;
        TcCacheScope tc_scope(ctx);        // dA^T is prepared once by the x-column product and reused by the h-column blocks
        // x columns of all gates, both directions:  dW[:, H:] += dA^T X
        rc = gemm_f32(ctx, true, false, N, Din, BL, 1.f, dA, N, x, ldx, 1.f, dW + H, ldw, nullptr, GemmBatch(), sk);
        for (int d = 0; d < ndir && !rc; d++) {
            float* dWd = dW + (size_t)d * 3 * H * ldw;
            // z, r gates see h_prev; the candidate sees r*h_prev (GRU.lua:23-26)
            rc = gemm_f32(ctx, true, false, 2 * H, H, BL, 1.f, dA + (size_t)d * 3 * H, N, hp_all + (size_t)d * H, ndir * H, 1.f, dWd, ldw,
                          nullptr, GemmBatch(), sk);
            if (!rc)
                rc = gemm_f32(ctx, true, false, H, H, BL, 1.f, dA + (size_t)d * 3 * H + 2 * H, N, save + (size_t)d * 4 * H + 3 * H, ndir * 4 * H, 1.f,
                              dWd + (size_t)2 * H * ldw, ldw, nullptr, GemmBatch(), sk);
        }
    }
    if (fork && !rc) {
        S2S_CUDA(cudaEventRecord(ctx->ev[3], ctx->side[1]));
        ctx->wgrad_join_pending = true;
    }
    return rc;
}

int gru_seq_wgrad_join(s2s_ctx* ctx) {
    if (ctx->wgrad_join_pending) {
        S2S_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev[3], 0));
        ctx->wgrad_join_pending = false;
    }
    return 0;
}

}  // namespace s2s
