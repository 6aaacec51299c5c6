// cluster_rnn.cuh -- device helpers shared by the persistent thread-block-cluster recurrences (gru_seq.cu, lstm_cluster.cu):
// DSMEM slice broadcast with st.async + mbarrier complete_tx, the transposed butterfly reduction and the
// lanes-split-K / warps-own-rows mat-vec.
#pragma once
#include "common.cuh"

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

// Broadcast this CTA's [BG][UC] slice (staged in local shared memory) into columns [UC*crank, UC*crank+UC) of
// buffer `buf_a` ([BG][H]) of EVERY CTA of the cluster with 16-byte st.async stores that signal the receiver's
// mbarrier: warp w serves ranks w, w+8, ...; lane -> 16-byte chunks (utterance, 4 units).
template <int H, int UC, int BG>
__device__ __forceinline__ void bcast_slice(const float (*stage)[UC], uint32_t buf_a, uint32_t bar_a, unsigned crank, int warp, int lane) {
    constexpr int CS = H / UC, CPB = UC / 4;
#pragma unroll
    for (int d = warp; d < CS; d += 8) {
        const uint32_t rbar = mapa_rank(bar_a, d);
#pragma unroll
        for (int ch = lane; ch < BG * CPB; ch += 32) {
            const int b = ch / CPB, off = (ch % CPB) * 4;
            const float4 v = *reinterpret_cast<const float4*>(&stage[b][off]);
            st_async_v4(mapa_rank(buf_a + (uint32_t)(b * H + crank * UC + off) * 4u, d), v, rbar);
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
// rows: phase 1 = 2 UC rows (ROWS = R1 per warp), phase 2 = UC rows (ROWS = R2 per warp).  The K-slice of the
// state vector is loaded ONCE per thread and phase (NI LDS.128 per utterance) and reused for all of the warp's
// rows, so the per-step shared-memory traffic is BG*H*4 bytes per WARP instead of per row; the partial sums of a
// warp are reduced with the transposed butterfly, which also hands each (row, utterance) result to its own lane:
// with G = ROWS * NBP (32 in phase 1, 16 in phase 2), lane l ends with (row (l % G) / NBP, utterance b_lo + l % NBP).
// NB = utterances present in this pass (<= NBP).
template <int H, int ROWS, int NB, int NBP>
__device__ __forceinline__ float matvec(const float4 (&w)[ROWS][H / 128], const float (*src)[H], int b_lo, int lane) {
    constexpr int NI = H / 128;
    constexpr int NBX = NB == 1 ? 1 : (NB == 2 ? 2 : (NB <= 4 ? 4 : 8));    // padded to a power of two
    constexpr int NV = ROWS * NBX;
    float4 x[NB][NI];
#pragma unroll
    for (int bb = 0; bb < NB; bb++)
#pragma unroll
        for (int i = 0; i < NI; i++) x[bb][i] = *reinterpret_cast<const float4*>(&src[b_lo + bb][lane * 4 + 128 * i]);
    float acc[NV];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
#pragma unroll
        for (int bb = 0; bb < NBX; bb++) {
            float a = 0.f;
            if (bb < NB) {
#pragma unroll
                for (int i = 0; i < NI; i++) a = dot4(w[r][i], x[bb][i], a);
            }
            acc[r * NBX + bb] = a;
        }
    float tot = bfly<NV>(acc, lane);
    if (NBX < NBP) {   // hand (row, utterance) to the canonical lane layout
        const int idx = lane % (ROWS * NBP);
        tot = __shfl_sync(0xffffffffu, tot, ((idx / NBP) * NBX + (idx % NBP)) & 31);
    }
    return tot;
}

// Two passes (groups of more than NBP utterances) as ONE instruction stream: both K-slices are loaded, both sets of
// partial sums are formed, and the two butterflies advance stage by stage together, so their shuffle latencies overlap
// instead of adding up (with two warps per scheduler the passes are latency-bound, not throughput-bound).
template <int H, int ROWS, int NB0, int NB1, int NBP>
__device__ __forceinline__ void matvec_pair(const float4 (&w)[ROWS][H / 128], const float (*src)[H], int lane, float& tot0, float& tot1) {
    constexpr int NI = H / 128;
    constexpr int NX0 = NB0 == 1 ? 1 : (NB0 == 2 ? 2 : (NB0 <= 4 ? 4 : 8));
    constexpr int NX1 = NB1 == 1 ? 1 : (NB1 == 2 ? 2 : (NB1 <= 4 ? 4 : 8));
    constexpr int NV0 = ROWS * NX0, NV1 = ROWS * NX1;
    float4 x0[NB0][NI], x1[NB1][NI];
#pragma unroll
    for (int bb = 0; bb < NB0; bb++)
#pragma unroll
        for (int i = 0; i < NI; i++) x0[bb][i] = *reinterpret_cast<const float4*>(&src[bb][lane * 4 + 128 * i]);
#pragma unroll
    for (int bb = 0; bb < NB1; bb++)
#pragma unroll
        for (int i = 0; i < NI; i++) x1[bb][i] = *reinterpret_cast<const float4*>(&src[NBP + bb][lane * 4 + 128 * i]);
    float a0[NV0], a1[NV1];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
#pragma unroll
        for (int bb = 0; bb < NX0; bb++) {
            float a = 0.f;
            if (bb < NB0) {
#pragma unroll
                for (int i = 0; i < NI; i++) a = dot4(w[r][i], x0[bb][i], a);
            }
            a0[r * NX0 + bb] = a;
        }
#pragma unroll
        for (int bb = 0; bb < NX1; bb++) {
            float a = 0.f;
            if (bb < NB1) {
#pragma unroll
                for (int i = 0; i < NI; i++) a = dot4(w[r][i], x1[bb][i], a);
            }
            a1[r * NX1 + bb] = a;
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        if (s < NV0) {
#pragma unroll
            for (int i = 0; i < s; i++) {
                if (i + s < NV0) {
                    const float send = (lane & s) ? a0[i] : a0[i + s];
                    const float keep = (lane & s) ? a0[i + s] : a0[i];
                    a0[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                }
            }
        }
        if (s < NV1) {
#pragma unroll
            for (int i = 0; i < s; i++) {
                if (i + s < NV1) {
                    const float send = (lane & s) ? a1[i] : a1[i + s];
                    const float keep = (lane & s) ? a1[i + s] : a1[i];
                    a1[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                }
            }
        }
    }
    float r0 = a0[0], r1 = a1[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        if (o >= NV0) r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        if (o >= NV1) r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    }
    if (NX0 < NBP) { const int idx = lane % (ROWS * NBP); r0 = __shfl_sync(0xffffffffu, r0, ((idx / NBP) * NX0 + (idx % NBP)) & 31); }
    if (NX1 < NBP) { const int idx = lane % (ROWS * NBP); r1 = __shfl_sync(0xffffffffu, r1, ((idx / NBP) * NX1 + (idx % NBP)) & 31); }
    tot0 = r0; tot1 = r1;
}

}  // namespace s2s
