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

#include "common.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {

__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16-byte store into another CTA's shared memory that also signals that CTA's mbarrier
// (complete_tx of 16 bytes): data + arrival in one message, no cluster-wide fence.
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                   "r"(__float_as_uint(v.w)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 4 values over the 4 lanes {l, l^1, l^2, l^3}: lane with (lane & 3) == i ends with the sum of value i
__device__ __forceinline__ float reduce4_transpose(float (&v)[4], int lane) {
    {
        const float s0 = (lane & 2) ? v[0] : v[2], k0 = (lane & 2) ? v[2] : v[0];
        const float s1 = (lane & 2) ? v[1] : v[3], k1 = (lane & 2) ? v[3] : v[1];
        v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
        v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
    }
    const float s = (lane & 1) ? v[0] : v[1], k = (lane & 1) ? v[1] : v[0];
    return k + __shfl_xor_sync(0xffffffffu, s, 1);
}

struct GruSeqParams {
    const float* W;        // [ndir][3][H][ldw]  (forward: rows used as-is; backward: read transposed)
    int ldw;               // H + Din
    const float* xp;       // [B, Lmax, ndir*3H] time-batched input projections (forward only)
    const int* lengths;
    int B, Lmax, ndir, reverse0;   // reverse0: direction of dir index 0 (ndir == 1 case)
    float* y;              // [B, Lmax, ndir*H]
    float* save;           // [B, Lmax, ndir, 4H]: z | r | h~ | r*h_prev
    // backward
    const float* dy;       // [B, Lmax, ndir*H]
    float* dA;             // [B, Lmax, ndir*3H]: daz | dar | dah   (same column order as xp)
    float* hp_all;         // [B, Lmax, ndir, H]
    int dbg;               // timing experiments only (S2S_GRU_DBG): 1 = skip the DSMEM exchange, 2 = skip the mat-vec loops
};

// Broadcast this CTA's [BG][32] slice (staged in local shared memory) into columns
// [32*crank, 32*crank+32) of buffer `buf_a` ([BG][H]) of EVERY CTA of the cluster: warp w sends to rank w,
// lane -> 16-byte chunks (utterance, 4 units).  At most two 16-byte st.async per thread.
template <int H, int BG>
__device__ __forceinline__ void bcast_slice(const float (*stage)[32], uint32_t buf_a, uint32_t bar_a, unsigned crank, int warp, int lane) {
    constexpr int CS = H / 32;
    if (warp < CS) {
        const uint32_t rbar = mapa_rank(bar_a, warp);
#pragma unroll
        for (int ch = lane; ch < BG * 8; ch += 32) {
            const int b = ch >> 3, off = (ch & 7) * 4;
            const float4 v = *reinterpret_cast<const float4*>(&stage[b][off]);
            st_async_v4(mapa_rank(buf_a + (uint32_t)(b * H + crank * 32 + off) * 4u, warp), v, rbar);
        }
    }
}
// Transposed butterfly reduction: every lane holds NV partial sums (NV a power of two <= 32); afterwards
// lane l holds the complete sum number (l mod NV).  NV - 1 + log2(32/NV) shuffles instead of 5 NV.
template <int NV>
__device__ __forceinline__ float bfly(float (&v)[NV], int lane) {
#pragma unroll
    for (int s = NV / 2; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float send = (lane & s) ? v[i] : v[i + s];
            const float keep = (lane & s) ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    float r = v[0];
#pragma unroll
    for (int o = NV; o < 32; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float c) {
    c = fmaf(a.x, b.x, c); c = fmaf(a.y, b.y, c); c = fmaf(a.z, b.z, c); return fmaf(a.w, b.w, c);
}

// Mat-vec mapping shared by forward and backward.  Lanes split K (lane owns k = 4 lane + 128 i), warps own
// rows: phase 1 = 64 rows (ROWS = 8 per warp), phase 2 = 32 rows (ROWS = 4 per warp).  The K-slice of the
// state vector is loaded ONCE per thread and phase (NI LDS.128 per utterance) and reused for all of the
// warp's rows, so the per-step shared-memory traffic is BG*H*4 bytes per WARP instead of per row; the partial
// sums of a warp are reduced with the transposed butterfly, which also hands each (row, utterance) result to
// its own lane: the value for (row l/4 [ROWS = 8] or (l%16)/4 [ROWS = 4], utterance b_lo + l%4) ends in lane l.
// Groups of more than 4 utterances are processed as two halves; NB = utterances in this half (1..4).
template <int H, int ROWS, int NB>
__device__ __forceinline__ float matvec(const float4 (&w)[ROWS][H / 128], const float (*src)[H], int b_lo, int lane) {
    constexpr int NI = H / 128;
    constexpr int NBP = NB == 1 ? 1 : (NB == 2 ? 2 : 4);
    constexpr int NV = ROWS * NBP;
    float4 x[NB][NI];
#pragma unroll
    for (int bb = 0; bb < NB; bb++)
#pragma unroll
        for (int i = 0; i < NI; i++) x[bb][i] = *reinterpret_cast<const float4*>(&src[b_lo + bb][lane * 4 + 128 * i]);
    float acc[NV];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
#pragma unroll
        for (int bb = 0; bb < NBP; bb++) {
            float a = 0.f;
            if (bb < NB) {
#pragma unroll
                for (int i = 0; i < NI; i++) a = dot4(w[r][i], x[bb][i], a);
            }
            acc[r * NBP + bb] = a;
        }
    float tot = bfly<NV>(acc, lane);
    if (NBP < 4) {   // hand (row, utterance) to the lane layout used by the callers
        const int row = ROWS == 8 ? (lane >> 2) : ((lane & 15) >> 2);
        tot = __shfl_sync(0xffffffffu, tot, (row * NBP + (lane & 3)) & 31);
    }
    return tot;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_fwd_kernel(const GruSeqParams p) {
    constexpr int NI = H / 128;
    constexpr int NH = (BG + 3) / 4;          // halves of up to 4 utterances
    constexpr int NB2 = BG - 4 > 0 ? BG - 4 : 1;   // utterances in the second half
    constexpr unsigned TX = BG * H * 4;       // bytes every CTA receives per exchange
    constexpr int CS = H / 32;
    __shared__ __align__(16) float hbuf[BG][H];
    __shared__ __align__(16) float rhbuf[BG][H];
    __shared__ __align__(16) float stage[BG][32];
    __shared__ float zbuf[BG][32];
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
    float4 w1[8][NI], w2[4][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float* row = Wd + ((size_t)g1 * H + crank * 32 + 8 * (warp & 3) + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w1[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float* row = Wd + ((size_t)2 * H + crank * 32 + 4 * warp + r) * p.ldw;
#pragma unroll
            for (int i = 0; i < NI; i++) { const float* s = row + lane * 4 + 128 * i; w2[r][i] = make_float4(s[0], s[1], s[2], s[3]); }
        }
    }
    for (int i = tid; i < BG * H; i += 256) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }   // Recurrent.lua:13,112
    if (tid == 0) { mbar_init(&barA, 1); mbar_init(&barB, 1); fence_mbar_init(); }

    // finalizer roles (see matvec8 / matvec4)
    const int ju1 = 8 * (warp & 3) + (lane >> 2);          // phase-1 unit within this CTA's slice
    const int ju2 = 4 * warp + ((lane & 15) >> 2);         // phase-2 unit (lanes < 16)
    const int j1 = crank * 32 + ju1, j2 = crank * 32 + ju2;
    int Lf[NH];                                            // length of the utterance this lane finalises, per half
#pragma unroll
    for (int hf = 0; hf < NH; hf++) {
        const int bl = 4 * hf + (lane & 3), b = b0 + bl;
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
        return __ldg(p.xp + ((size_t)(b0 + 4 * hf + (lane & 3)) * p.Lmax + t) * (p.ndir * H3) + dir * H3 + gate * H + j);
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
        if (tid == 0 && !(p.dbg & 1)) { mbar_expect_tx(&barA, TX); mbar_expect_tx(&barB, TX); }

        // ---- phase 1: z, r ------------------------------------------------------------------
        float tot1[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            tot1[hf] = 0.f;
            if (!(p.dbg & 2)) tot1[hf] = hf == 0 ? matvec<H, 8, (BG < 4 ? BG : 4)>(w1, hbuf, 0, lane) : matvec<H, 8, NB2>(w1, hbuf, 4, lane);
        }
        float g1v[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) g1v[hf] = sigmoid_acc(tot1[hf] + xp1[hf]);    // GRU.lua:23-24 (both halves in flight)
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = 4 * hf + (lane & 3);
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
            bcast_slice<H, BG>(stage, rhbuf_a, barA_a, crank, warp, lane);
            mbar_wait(&barA, parity);
        }

        // ---- phase 2: h~, h' ------------------------------------------------------------------
        float tot2[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            tot2[hf] = 0.f;
            if (!(p.dbg & 2)) tot2[hf] = hf == 0 ? matvec<H, 4, (BG < 4 ? BG : 4)>(w2, rhbuf, 0, lane) : matvec<H, 4, NB2>(w2, rhbuf, 4, lane);
        }
        float hcv[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) hcv[hf] = tanh_acc(tot2[hf] + xp2[hf]);       // GRU.lua:26 (both halves in flight)
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const int bl = 4 * hf + (lane & 3);
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
            bcast_slice<H, BG>(stage, hbuf_a, barB_a, crank, warp, lane);
            mbar_wait(&barB, parity);
        }
        parity ^= 1;
    }
    cluster_sync_all();   // no CTA exits while a peer may still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int H, int BG>
__global__ void __launch_bounds__(256, 1)
gru_seq_bwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32;
    constexpr int NI = H / 128;
    constexpr int NH = (BG + 3) / 4;
    constexpr int NB2 = BG - 4 > 0 ? BG - 4 : 1;
    constexpr unsigned TX = BG * H * 4;
    __shared__ __align__(16) float ahbuf[BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[BG][H];   // daz
    __shared__ __align__(16) float arbuf[BG][H];   // dar
    __shared__ __align__(16) float stage_h[BG][32], stage_z[BG][32], stage_r[BG][32];
    __shared__ float stash_r[BG][32], stash_hp[BG][32];
    __shared__ float part1[BG][32], part2[BG][32];
    __shared__ uint64_t barA, barB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + BG - 1) / BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * BG;
    const int H3 = 3 * H;

    // transposed recurrent weights -> registers.  phase 1 rows (8 per warp): warps 0-3 -> W_h^T (applied to
    // dah), warps 4-7 -> W_z^T (applied to daz); phase 2 rows (4 per warp): W_r^T (applied to dar).  Row =
    // input unit owned by this CTA, reduction over the output units j = 4 lane + 128 i + e.
    const int g1 = warp >> 2;
    float4 w1[8][NI], w2[4][NI];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
        const float* Wg = Wd + (size_t)(g1 == 0 ? 2 : 0) * H * p.ldw + crank * 32 + 8 * (warp & 3);
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w1[r][i] = make_float4(Wg[j * p.ldw + r], Wg[(j + 1) * p.ldw + r], Wg[(j + 2) * p.ldw + r], Wg[(j + 3) * p.ldw + r]);
            }
        const float* Wr = Wd + (size_t)H * p.ldw + crank * 32 + 4 * warp;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t j = lane * 4 + 128 * i;
                w2[r][i] = make_float4(Wr[j * p.ldw + r], Wr[(j + 1) * p.ldw + r], Wr[(j + 2) * p.ldw + r], Wr[(j + 3) * p.ldw + r]);
            }
    }
    for (int i = tid; i < BG * H; i += 256) { (&ahbuf[0][0])[i] = 0.f; (&azbuf[0][0])[i] = 0.f; (&arbuf[0][0])[i] = 0.f; }
    if (tid == 0) { mbar_init(&barA, 1); mbar_init(&barB, 1); fence_mbar_init(); }

    // roles: owner (elementwise part + carry) = phase-2 finaliser: lanes < 16 -> (unit ju2, utterance lane%4) per half
    const int ju1 = 8 * (warp & 3) + (lane >> 2);
    const int ju2 = 4 * warp + ((lane & 15) >> 2);
    const int j1 = crank * 32 + ju1, j2 = crank * 32 + ju2;
    const bool owner_lane = lane < 16;
    int Lf[NH];
#pragma unroll
    for (int hf = 0; hf < NH; hf++) {
        const int bl = 4 * hf + (lane & 3), b = b0 + bl;
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
        const int b = b0 + 4 * hf + (lane & 3);
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
        if (tid == 0 && !(p.dbg & 1)) { mbar_expect_tx(&barA, 2 * TX); mbar_expect_tx(&barB, TX); }
        // ---- elementwise part (owners) ---------------------------------------------------------
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const Pre cur = nxt[hf];
            nxt[hf] = load_pre(s - 1, hf);
            dhp_part[hf] = 0.f;
            const int bl = 4 * hf + (lane & 3);
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
            bcast_slice<H, BG>(stage_h, ah_a, barA_a, crank, warp, lane);
            bcast_slice<H, BG>(stage_z, az_a, barA_a, crank, warp, lane);
            mbar_wait(&barA, parity);
        }

        // ---- phase 1: d(r*h) = W_h[:, :H]^T dah ; W_z[:, :H]^T daz -------------------------------
        const float (*src1)[H] = g1 == 0 ? ahbuf : azbuf;
        float tot1[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            tot1[hf] = 0.f;
            if (!(p.dbg & 2)) tot1[hf] = hf == 0 ? matvec<H, 8, (BG < 4 ? BG : 4)>(w1, src1, 0, lane) : matvec<H, 8, NB2>(w1, src1, 4, lane);
        }
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const float tot = tot1[hf];
            const int bl = 4 * hf + (lane & 3);
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
            bcast_slice<H, BG>(stage_r, ar_a, barB_a, crank, warp, lane);
            mbar_wait(&barB, parity);
        }

        // ---- phase 2: W_r[:, :H]^T dar ; carry ---------------------------------------------------
        float tot2[NH];
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            tot2[hf] = 0.f;
            if (!(p.dbg & 2)) tot2[hf] = hf == 0 ? matvec<H, 4, (BG < 4 ? BG : 4)>(w2, arbuf, 0, lane) : matvec<H, 4, NB2>(w2, arbuf, 4, lane);
        }
#pragma unroll
        for (int hf = 0; hf < NH; hf++) {
            const float tot = tot2[hf];
            const int bl = 4 * hf + (lane & 3);
            if (owner_lane && bl < BG && s < Lf[hf]) carry[hf] = dhp_part[hf] + part1[bl][ju2] + part2[bl][ju2] + tot;
        }
        parity ^= 1;
    }
    cluster_sync_all();
}

// rows t >= L_b of a [B, Lmax, W] tensor := 0 (padding must not leak NaNs into the time-batched GEMMs)
__global__ void zero_tail_rows_kernel(float* __restrict__ x, const int* __restrict__ lengths, int Lmax, int W) {
    const int b = blockIdx.y, t = blockIdx.x;
    if (t < lengths[b]) return;
    float* r = x + ((size_t)b * Lmax + t) * W;
    for (int i = threadIdx.x; i < W; i += blockDim.x) r[i] = 0.f;
}
int zero_tail_rows(s2s_ctx* ctx, float* x, const int* lengths, int B, int Lmax, int W) {
    if (!lengths) return 0;
    zero_tail_rows_kernel<<<dim3(Lmax, B), 128, 0, ctx->stream>>>(x, lengths, Lmax, W);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

template <int H, int BG>
static int launch_cluster_bg(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int* max_clusters) {
    constexpr int CS = H / 32;
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
    if (max_clusters) {   // occupancy query only
        if (backward) S2S_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, gru_seq_bwd_kernel<H, BG>, &cfg));
        else S2S_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, gru_seq_fwd_kernel<H, BG>, &cfg));
        return 0;
    }
    prof_begin(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD);
    if (backward) S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_bwd_kernel<H, BG>, p));
    else S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_fwd_kernel<H, BG>, p));
    {   // algorithmic bytes per launch: fwd reads xp (3H) and writes y (H) + save (4H) per direction;
        // bwd reads save z,r,h~ (3H) + h_prev (H) + dy (H) and writes dA (3H) + h_prev (H)
        const double per = backward ? 9.0 * H : 8.0 * H;
        prof_end(ctx, backward ? S2S_PROF_GRU_BWD : S2S_PROF_GRU_FWD, 4.0 * p.B * p.Lmax * p.ndir * per);
    }
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// Utterances per cluster: the smallest group size whose cluster count fits in ONE wave (a cluster that
// does not fit runs after the others and doubles the time of this latency-bound kernel).  The number of
// co-resident clusters is queried once per (H, direction) -- 15 clusters of 8 CTAs on a 148-SM B200.
template <int H>
static int launch_cluster(s2s_ctx* ctx, bool backward, const GruSeqParams& p) {
    static int max_active[2] = {0, 0};
    if (max_active[backward] == 0) {
        int n = 0;
        S2S_TRY((launch_cluster_bg<H, 4>(ctx, backward, p, &n)));
        max_active[backward] = n > 0 ? n : 1;
    }
    const int cap = max_active[backward];
    int bg = 4;
    while (bg < 8 && p.ndir * ceil_div(p.B, bg) > cap) bg++;
    { const char* e = getenv("S2S_GRU_BG"); if (e && atoi(e) >= 4 && atoi(e) <= 8) bg = atoi(e); }
    switch (bg) {
        case 4: return launch_cluster_bg<H, 4>(ctx, backward, p, nullptr);
        case 5: return launch_cluster_bg<H, 5>(ctx, backward, p, nullptr);
        case 6: return launch_cluster_bg<H, 6>(ctx, backward, p, nullptr);
        case 7: return launch_cluster_bg<H, 7>(ctx, backward, p, nullptr);
        default: return launch_cluster_bg<H, 8>(ctx, backward, p, nullptr);
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
        S2S_TRY(zero_tail_rows(ctx, y, lengths, B, Lmax, ndir * H));
        S2S_TRY(zero_tail_rows(ctx, save, lengths, B, Lmax, ndir * 4 * H));
    }
    GruSeqParams p = {};
    p.W = W; p.ldw = ldw; p.xp = xp; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.ndir = ndir; p.reverse0 = reverse;
    p.y = y; p.save = save;
    { const char* e = getenv("S2S_GRU_DBG"); p.dbg = e ? atoi(e) : 0; }
    if (H == 128) return launch_cluster<128>(ctx, false, p);
    return launch_cluster<256>(ctx, false, p);
}

int gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                     const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx) {
    S2S_REQUIRE(H == 128 || H == 256, "gru_seq: hidden size %d not supported by the cluster kernel (128 or 256)", H);
    S2S_REQUIRE(ndir == 1 || ndir == 2, "gru_seq: ndir must be 1 or 2");
    const int ldw = H + Din, N = ndir * 3 * H, BL = B * Lmax;
    float *dA, *hp_all;
    S2S_ALLOC(dA, ctx->arena, float, (size_t)BL * N);
    S2S_ALLOC(hp_all, ctx->arena, float, (size_t)BL * ndir * H);
    if (lengths) {
        S2S_TRY(zero_tail_rows(ctx, dA, lengths, B, Lmax, N));
        S2S_TRY(zero_tail_rows(ctx, hp_all, lengths, B, Lmax, ndir * H));
    }
    GruSeqParams p = {};
    p.W = W; p.ldw = ldw; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.ndir = ndir; p.reverse0 = reverse;
    p.y = const_cast<float*>(y); p.save = const_cast<float*>(save); p.dy = dy; p.dA = dA; p.hp_all = hp_all;
    { const char* e = getenv("S2S_GRU_DBG"); p.dbg = e ? atoi(e) : 0; }
    if (H == 128) S2S_TRY(launch_cluster<128>(ctx, true, p));
    else S2S_TRY(launch_cluster<256>(ctx, true, p));
    // time-batched gradients (K = B*L) instead of one rank-1 update per frame (LinearZeroBias.lua:67-74)
    const int sk = 8;
    // x columns of all gates, both directions:  dW[:, H:] += dA^T X
    S2S_TRY(gemm_f32(ctx, true, false, N, Din, BL, 1.f, dA, N, x, ldx, 1.f, dW + H, ldw, nullptr, GemmBatch(), sk));
    for (int d = 0; d < ndir; d++) {
        float* dWd = dW + (size_t)d * 3 * H * ldw;
        // z, r gates see h_prev; the candidate sees r*h_prev (GRU.lua:23-26)
        S2S_TRY(gemm_f32(ctx, true, false, 2 * H, H, BL, 1.f, dA + (size_t)d * 3 * H, N, hp_all + (size_t)d * H, ndir * H, 1.f, dWd, ldw,
                         nullptr, GemmBatch(), sk));
        S2S_TRY(gemm_f32(ctx, true, false, H, H, BL, 1.f, dA + (size_t)d * 3 * H + 2 * H, N, save + (size_t)d * 4 * H + 3 * H, ndir * 4 * H, 1.f,
                         dWd + (size_t)2 * H * ldw, ldw, nullptr, GemmBatch(), sk));
    }
    // dX = dA . W[:, H:]   (LinearZeroBias.lua:50-65, x columns), both directions summed as nngraph does
    if (dx) S2S_TRY(gemm_f32(ctx, false, false, BL, Din, N, 1.f, dA, N, W + H, ldw, 0.f, dx, Din));
    return 0;
}

}  // namespace s2s
