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
//     GRU_BG utterances of one direction; CTA c owns hidden units [32c, 32c+32) and keeps its
//     3 x 32 rows of the recurrent weights in REGISTERS for all L steps (96 floats per thread);
//     h_{t-1} and r*h_{t-1} live in shared memory and are exchanged every step through distributed
//     shared memory (st.shared::cluster) + two cluster barriers -- no HBM round trip for the state,
//     no kernel launch per step;
//   * forward/reverse directions and batch groups are independent clusters of the same launch
//     (2 dirs x 8 groups x 8 CTAs = 128 SMs at the Chorowski TIMIT batch of 32);
//   * backward mirrors it with the transposed weights in registers and emits the gate
//     pre-activation gradients dA[B,L,ndir,3H]; every weight gradient and dX are time-batched GEMMs
//     afterwards (K = B*L), instead of one rank-1 update per frame.
//   * per-utterance lengths: a step is inactive for utterance b once s >= L_b; the reverse direction
//     starts at L_b - 1, so padded batches reproduce the reference's per-utterance results.
#include <cooperative_groups.h>

#include "common.cuh"
#include "gru_seq.cuh"

namespace cg = cooperative_groups;

namespace s2s {

__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
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
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256, 1)
gru_seq_fwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32;         // cluster size
    constexpr int N1 = H / 16;         // float4 chunks per thread, phase 1 (4 k-quarters)
    constexpr int N2 = H / 32;         // float4 chunks per thread, phase 2 (8 k-eighths)
    __shared__ __align__(16) float hbuf[GRU_BG][H];
    __shared__ __align__(16) float rhbuf[GRU_BG][H];
    __shared__ float zbuf[GRU_BG][32];

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + GRU_BG - 1) / GRU_BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * GRU_BG;
    const int H3 = 3 * H;

    // ---- recurrent weights -> registers -----------------------------------------------------
    const int r1 = tid >> 2, kq = tid & 3;          // phase 1: 64 rows (z: 0-31, r: 32-63) x 4 k-quarters
    const int g1 = r1 >> 5, ju1 = r1 & 31;
    const int r2 = tid >> 3, k8 = tid & 7;          // phase 2: 32 rows (h~) x 8 k-eighths
    float4 w1[N1], w2[N2];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
        const float* row1 = Wd + ((size_t)g1 * H + crank * 32 + ju1) * p.ldw;
#pragma unroll
        for (int i = 0; i < N1; i++) {
            const float* s = row1 + kq * 4 + 16 * i;
            w1[i] = make_float4(s[0], s[1], s[2], s[3]);
        }
        const float* row2 = Wd + ((size_t)2 * H + crank * 32 + r2) * p.ldw;
#pragma unroll
        for (int i = 0; i < N2; i++) {
            const float* s = row2 + k8 * 4 + 32 * i;
            w2[i] = make_float4(s[0], s[1], s[2], s[3]);
        }
    }
    for (int i = tid; i < GRU_BG * H; i += 256) { (&hbuf[0][0])[i] = 0.f; (&rhbuf[0][0])[i] = 0.f; }   // Recurrent.lua:13,112

    // finalizer roles: phase 1 -> (row r1, utterance kq); phase 2 -> lanes k8 < 4: (unit r2, utterance k8)
    const int bf1 = b0 + kq;
    const int Lf1 = bf1 < p.B ? (p.lengths ? p.lengths[bf1] : p.Lmax) : 0;
    const int bf2 = b0 + (k8 & 3);
    const int Lf2 = (k8 < 4 && bf2 < p.B) ? (p.lengths ? p.lengths[bf2] : p.Lmax) : 0;
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < GRU_BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t hbuf_a = smem_u32(&hbuf[0][0]), rhbuf_a = smem_u32(&rhbuf[0][0]);
    const int j1 = crank * 32 + ju1, j2 = crank * 32 + r2;
    cluster_sync_all();

    // input projections are independent of the recurrence: step s+1's values are fetched while step s runs
    auto load_xp1 = [&](int s) -> float {
        if (s >= Lf1) return 0.f;
        const int t = rev ? Lf1 - 1 - s : s;
        return __ldg(p.xp + ((size_t)bf1 * p.Lmax + t) * (p.ndir * H3) + dir * H3 + g1 * H + j1);
    };
    auto load_xp2 = [&](int s) -> float {
        if (s >= Lf2) return 0.f;
        const int t = rev ? Lf2 - 1 - s : s;
        return __ldg(p.xp + ((size_t)bf2 * p.Lmax + t) * (p.ndir * H3) + dir * H3 + 2 * H + j2);
    };
    float xp1n = load_xp1(0), xp2n = load_xp2(0);
    for (int s = 0; s < Lgrp; s++) {
        const bool act1 = s < Lf1, act2 = s < Lf2;
        const int t1 = rev ? Lf1 - 1 - s : s, t2 = rev ? Lf2 - 1 - s : s;
        const float xp1 = xp1n, xp2 = xp2n;
        xp1n = load_xp1(s + 1); xp2n = load_xp2(s + 1);

        // ---- phase 1: z, r ------------------------------------------------------------------
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < N1; i++) {
#pragma unroll
            for (int b = 0; b < GRU_BG; b++) {
                const float4 x = *reinterpret_cast<const float4*>(&hbuf[b][kq * 4 + 16 * i]);
                acc[b] = fmaf(w1[i].x, x.x, acc[b]); acc[b] = fmaf(w1[i].y, x.y, acc[b]);
                acc[b] = fmaf(w1[i].z, x.z, acc[b]); acc[b] = fmaf(w1[i].w, x.w, acc[b]);
            }
        }
        {
            const float tot = reduce4_transpose(acc, lane);
            if (act1) {
                const float g = sigmoid_acc(tot + xp1);                               // GRU.lua:23-24
                float* sv = p.save + (((size_t)bf1 * p.Lmax + t1) * p.ndir + dir) * 4 * H;
                sv[g1 * H + j1] = g;
                if (g1 == 0) {
                    zbuf[kq][ju1] = g;
                } else {
                    const float rh = g * hbuf[kq][j1];                                // GRU.lua:25
                    sv[3 * H + j1] = rh;
                    const uint32_t off = rhbuf_a + (uint32_t)(kq * H + j1) * 4u;
#pragma unroll
                    for (int c = 0; c < CS; c++) st_cluster_f32(mapa_rank(off, c), rh);
                }
            }
        }
        cluster_sync_all();

        // ---- phase 2: h~, h' ------------------------------------------------------------------
        float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < N2; i++) {
#pragma unroll
            for (int b = 0; b < GRU_BG; b++) {
                const float4 x = *reinterpret_cast<const float4*>(&rhbuf[b][k8 * 4 + 32 * i]);
                acc2[b] = fmaf(w2[i].x, x.x, acc2[b]); acc2[b] = fmaf(w2[i].y, x.y, acc2[b]);
                acc2[b] = fmaf(w2[i].z, x.z, acc2[b]); acc2[b] = fmaf(w2[i].w, x.w, acc2[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < GRU_BG; b++) acc2[b] += __shfl_xor_sync(0xffffffffu, acc2[b], 4);
        {
            const float tot = reduce4_transpose(acc2, lane);
            if (act2) {
                const int bb = k8 & 3;
                const float hc = tanh_acc(tot + xp2);                                 // GRU.lua:26
                const float z = zbuf[bb][r2], hp = hbuf[bb][j2];
                const float hn = (1.f - z) * hp + z * hc;                             // GRU.lua:27-30
                p.save[(((size_t)bf2 * p.Lmax + t2) * p.ndir + dir) * 4 * H + 2 * H + j2] = hc;
                p.y[((size_t)bf2 * p.Lmax + t2) * (p.ndir * H) + dir * H + j2] = hn;
                const uint32_t off = hbuf_a + (uint32_t)(bb * H + j2) * 4u;
#pragma unroll
                for (int c = 0; c < CS; c++) st_cluster_f32(mapa_rank(off, c), hn);
            }
        }
        cluster_sync_all();
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256, 1)
gru_seq_bwd_kernel(const GruSeqParams p) {
    constexpr int CS = H / 32;
    constexpr int N1 = H / 16;
    constexpr int N2 = H / 32;
    __shared__ __align__(16) float ahbuf[GRU_BG][H];   // dah (all units)
    __shared__ __align__(16) float azbuf[GRU_BG][H];   // daz
    __shared__ __align__(16) float arbuf[GRU_BG][H];   // dar
    __shared__ float stash_r[GRU_BG][32], stash_hp[GRU_BG][32];
    __shared__ float part1[GRU_BG][32], part2[GRU_BG][32];

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned crank = cg::this_cluster().block_rank();
    const int cluster_id = blockIdx.x / CS;
    const int ngroups = (p.B + GRU_BG - 1) / GRU_BG;
    const int dir = cluster_id / ngroups, grp = cluster_id % ngroups;
    const bool rev = p.ndir == 2 ? dir == 1 : p.reverse0 != 0;
    const int b0 = grp * GRU_BG;
    const int H3 = 3 * H;

    // transposed recurrent weights -> registers.  phase 1 rows: g1 = 0 -> W_h^T (on dah), 1 -> W_z^T (on daz);
    // phase 2 rows: W_r^T (on dar).  Row i = input unit owned by this CTA, reduction over output units j.
    const int r1 = tid >> 2, kq = tid & 3;
    const int g1 = r1 >> 5, ju1 = r1 & 31;
    const int r2 = tid >> 3, k8 = tid & 7;
    float4 w1[N1], w2[N2];
    {
        const float* Wd = p.W + (size_t)dir * 3 * H * p.ldw;
        const float* Wg = Wd + (size_t)(g1 == 0 ? 2 : 0) * H * p.ldw + crank * 32 + ju1;
#pragma unroll
        for (int i = 0; i < N1; i++) {
            const int j = kq * 4 + 16 * i;
            w1[i] = make_float4(Wg[(size_t)j * p.ldw], Wg[(size_t)(j + 1) * p.ldw], Wg[(size_t)(j + 2) * p.ldw], Wg[(size_t)(j + 3) * p.ldw]);
        }
        const float* Wr = Wd + (size_t)H * p.ldw + crank * 32 + r2;
#pragma unroll
        for (int i = 0; i < N2; i++) {
            const int j = k8 * 4 + 32 * i;
            w2[i] = make_float4(Wr[(size_t)j * p.ldw], Wr[(size_t)(j + 1) * p.ldw], Wr[(size_t)(j + 2) * p.ldw], Wr[(size_t)(j + 3) * p.ldw]);
        }
    }
    for (int i = tid; i < GRU_BG * H; i += 256) { (&ahbuf[0][0])[i] = 0.f; (&azbuf[0][0])[i] = 0.f; (&arbuf[0][0])[i] = 0.f; }

    // elementwise owner role (= phase-2 finalizer): lanes k8 < 4 own (unit r2, utterance k8)
    const bool owner = k8 < 4;
    const int bo = b0 + (k8 & 3);
    const int Lo = (owner && bo < p.B) ? (p.lengths ? p.lengths[bo] : p.Lmax) : 0;
    const int bf1 = b0 + kq;
    const int Lf1 = bf1 < p.B ? (p.lengths ? p.lengths[bf1] : p.Lmax) : 0;
    int Lgrp = 0;
#pragma unroll
    for (int b = 0; b < GRU_BG; b++)
        if (b0 + b < p.B) Lgrp = max(Lgrp, p.lengths ? p.lengths[b0 + b] : p.Lmax);

    const uint32_t ah_a = smem_u32(&ahbuf[0][0]), az_a = smem_u32(&azbuf[0][0]), ar_a = smem_u32(&arbuf[0][0]);
    const int j1 = crank * 32 + ju1, j2 = crank * 32 + r2;
    float carry = 0.f;          // dE/dh flowing to the previous recurrence step (owner threads)
    cluster_sync_all();

    // saved activations / incoming gradients do not depend on the recurrence: prefetch one step ahead
    struct Pre { float z, r, hc, hp, dy; };
    auto load_pre = [&](int s) -> Pre {
        Pre q = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (s < 0 || s >= Lo) return q;
        const int t = rev ? Lo - 1 - s : s;
        const size_t row = (size_t)bo * p.Lmax + t;
        const float* sv = p.save + (row * p.ndir + dir) * 4 * H;
        q.z = __ldg(sv + j2); q.r = __ldg(sv + H + j2); q.hc = __ldg(sv + 2 * H + j2);
        if (s > 0) {                                                                  // RNN.lua:186-192
            const int tp = rev ? t + 1 : t - 1;
            q.hp = __ldg(p.y + ((size_t)bo * p.Lmax + tp) * (p.ndir * H) + dir * H + j2);
        }
        q.dy = __ldg(p.dy + row * (p.ndir * H) + dir * H + j2);
        return q;
    };
    Pre nxt = load_pre(Lgrp - 1);
    for (int s = Lgrp - 1; s >= 0; s--) {                                           // RNN.lua:183
        const bool acto = s < Lo, act1 = s < Lf1;
        const int to = rev ? Lo - 1 - s : s, t1 = rev ? Lf1 - 1 - s : s;
        float dhp_part = 0.f;
        const Pre cur = nxt;
        nxt = load_pre(s - 1);
        // ---- elementwise part (owners) ---------------------------------------------------------
        if (acto) {
            const int bb = k8 & 3;
            const size_t row = (size_t)bo * p.Lmax + to;
            const float z = cur.z, r = cur.r, hc = cur.hc, hp = cur.hp;
            const float dh = cur.dy + carry;                                          // RNN.lua:193-194
            const float dah = dh * z * (1.f - hc * hc);
            const float daz = dh * (hc - hp) * z * (1.f - z);
            dhp_part = dh * (1.f - z);
            float* da = p.dA + row * (p.ndir * H3) + dir * H3;
            da[j2] = daz; da[2 * H + j2] = dah;
            p.hp_all[(row * p.ndir + dir) * H + j2] = hp;
            stash_r[bb][r2] = r; stash_hp[bb][r2] = hp;
            const uint32_t o = (uint32_t)(bb * H + j2) * 4u;
#pragma unroll
            for (int c = 0; c < CS; c++) { st_cluster_f32(mapa_rank(ah_a + o, c), dah); st_cluster_f32(mapa_rank(az_a + o, c), daz); }
        }
        cluster_sync_all();

        // ---- phase 1: d(r*h) = W_h[:, :H]^T dah ; W_z[:, :H]^T daz -------------------------------
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const float (*src)[H] = g1 == 0 ? ahbuf : azbuf;
#pragma unroll
            for (int i = 0; i < N1; i++) {
#pragma unroll
                for (int b = 0; b < GRU_BG; b++) {
                    const float4 x = *reinterpret_cast<const float4*>(&src[b][kq * 4 + 16 * i]);
                    acc[b] = fmaf(w1[i].x, x.x, acc[b]); acc[b] = fmaf(w1[i].y, x.y, acc[b]);
                    acc[b] = fmaf(w1[i].z, x.z, acc[b]); acc[b] = fmaf(w1[i].w, x.w, acc[b]);
                }
            }
        }
        {
            const float tot = reduce4_transpose(acc, lane);
            if (g1 == 0) {
                float dar = 0.f, pr = 0.f;
                if (act1) {
                    const float r = stash_r[kq][ju1], hp = stash_hp[kq][ju1];
                    dar = tot * hp * r * (1.f - r);
                    pr = tot * r;
                    p.dA[((size_t)bf1 * p.Lmax + t1) * (p.ndir * H3) + dir * H3 + H + j1] = dar;
                }
                part1[kq][ju1] = pr;
                const uint32_t o = ar_a + (uint32_t)(kq * H + j1) * 4u;
#pragma unroll
                for (int c = 0; c < CS; c++) st_cluster_f32(mapa_rank(o, c), dar);
            } else {
                part2[kq][ju1] = act1 ? tot : 0.f;
            }
        }
        cluster_sync_all();

        // ---- phase 2: W_r[:, :H]^T dar ; carry ---------------------------------------------------
        float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < N2; i++) {
#pragma unroll
            for (int b = 0; b < GRU_BG; b++) {
                const float4 x = *reinterpret_cast<const float4*>(&arbuf[b][k8 * 4 + 32 * i]);
                acc2[b] = fmaf(w2[i].x, x.x, acc2[b]); acc2[b] = fmaf(w2[i].y, x.y, acc2[b]);
                acc2[b] = fmaf(w2[i].z, x.z, acc2[b]); acc2[b] = fmaf(w2[i].w, x.w, acc2[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < GRU_BG; b++) acc2[b] += __shfl_xor_sync(0xffffffffu, acc2[b], 4);
        {
            const float tot = reduce4_transpose(acc2, lane);
            if (acto) carry = dhp_part + part1[k8 & 3][r2] + part2[k8 & 3][r2] + tot;
        }
        // (the next iteration's owner writes to ah/az happen after every CTA passed the barrier above;
        //  part1/part2/stash are re-written only after the next barrier A)
    }
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

template <int H>
static int launch_cluster(s2s_ctx* ctx, bool backward, const GruSeqParams& p) {
    constexpr int CS = H / 32;
    const int ngroups = ceil_div(p.B, GRU_BG);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ngroups * p.ndir);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (backward) S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_bwd_kernel<H>, p));
    else S2S_CUDA(cudaLaunchKernelEx(&cfg, gru_seq_fwd_kernel<H>, p));
    S2S_LAUNCH_CHECK(ctx);
    return 0;
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
