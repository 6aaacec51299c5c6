// attention.cu -- the location-aware / content attention step of nn.Attention as three fused,
// HBM-bound sm_100a kernels.
//
// Reference semantics (per utterance, per decoder step; Attention.lua:65-135):
//     Z[l,:] = q + Vh[l,:] (+ U F[l,:])        e[l] = w . tanh(Z[l,:])
//     alpha  = softmax_l(e)                     c    = sum_l alpha[l] h[l,:]
// The reference materialises ~20 [L,S] temporaries per step; here one launch reads Vh and h exactly
// once per step (the algorithmic minimum, A_f = 4 B (L S + L A + 2L + S + A) bytes):
//
//   attn_fwd_kernel   grid (ceil(Lmax/32), B): a CTA owns 32 encoder frames of one utterance.
//                     The h tile is fetched by ONE 1-D bulk-TMA copy (cp.async.bulk -> smem,
//                     mbarrier complete_tx) issued before the scoring phase so it lands while the
//                     warps stream Vh rows through registers (128-bit ld.global.nc, L1 no-allocate)
//                     and reduce e[l] with warp shuffles.  Chunk-local softmax statistics and the
//                     partial context are combined flash-decoding style by the LAST CTA of the
//                     utterance to arrive (atomic ticket), so there is no second launch.
//   attn_bwd_kernel   same tiling, single pass: because de = alpha (dalpha - <alpha,dalpha>) is
//                     affine in the not-yet-known dot product, each chunk accumulates the two
//                     vectors P1 = sum_l alpha dalpha g_l and P2 = sum_l alpha g_l
//                     (g_l = w (1 - tanh^2 Z_l)); the last CTA forms dq = P1 - dot P2 and de.
//   attn_dvh_kernel   deferred accumulation: dVh = sum_t de_t g_t in one pass after the time loop
//                     (replaces the reference's per-step [L,S]+[L,A] read-modify-write,
//                     RNNAttention.lua:247), tanh recomputed instead of stored.
//
// The location term is folded: U (W_F * alpha_pad + b_F) = UW * alpha_pad + Ub with UW = U W_F
// ([KF,S], built once per call), i.e. KF instead of K FMAs per element and no F temporary.
#include "attention.cuh"

namespace s2s {

constexpr int ATT_MAXCH = 128;   // Lmax <= 4096
constexpr int LOC_MAXKF = 16;

struct AttnFwdParams {
    const float *Vh, *h, *q, *w;
    int64_t ldq;
    const int* lengths;
    int B, Lmax, nch;
    // location
    int KF, padl;
    const float *uw, *alpha_prev;
    int64_t ld_aprev;
    // scratch
    float *E, *part_ms, *part_c;
    unsigned* counters;
    // outputs
    float *alpha, *c, *pen;
    int64_t ld_alpha, ld_c, ld_pen;
    float lambda;
    const float* app;   // alpha_{t-1} for the penalty
    int64_t ld_app;
    const int* tlens;   // padded decoder steps (tstep >= T_b) record no penalty
    int tstep;
    int dbg;            // timing experiments (S2S_ATT_DBG): 1 = no combine, 2 = no scoring, 4 = no context
};

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 b) { return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w)); }

// packed fp32 pairs: FFMA2 (fma.rn.f32x2) does two FMAs per issue slot -- the location term is issue-bound
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int NS, int NA, int LOC, int RIF>
__global__ void __launch_bounds__(ATT_THREADS, (RIF == 2 && !LOC) ? 3 : 2)
attn_fwd_kernel(const AttnFwdParams p) {
    constexpr int S = NS * 128, A = NA * 128;
    constexpr int GROUPS = 8 / NA;              // row groups in the context phase
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* hs = reinterpret_cast<float*>(smem_raw);                 // [ATT_R][A]
    float* w_s = hs + ATT_R * A;                                     // [S]
    float* uw_s = w_s + S;                                           // [KF][S]        (LOC)
    float2* ap_s = reinterpret_cast<float2*>(uw_s + (LOC ? p.KF * S : 0));   // [ATT_R+KF-1] {a, a} pairs   (LOC)
    constexpr int KFT = LOC > 1 ? LOC : 0;                           // exact filter size known at compile time
    __shared__ float e_s[ATT_R], p_s[ATT_R];
    __shared__ __align__(16) float red[1024];
    __shared__ float scl_s[ATT_MAXCH];
    __shared__ float wred[8];
    __shared__ uint64_t bar;
    __shared__ int is_last;

    const int b = blockIdx.y, ch = blockIdx.x;
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    const int l0 = ch * ATT_R;
    if (l0 >= Lb) return;
    const int nrows = min(ATT_R, Lb - l0);
    const int nchb = (Lb + ATT_R - 1) / ATT_R;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    pdl_trigger();
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0 && !(p.dbg & 4)) {
        const unsigned bytes = (unsigned)nrows * A * 4u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(hs, p.h + ((size_t)b * p.Lmax + l0) * A, bytes, &bar);
    }
    if (LOC) {
        for (int i = tid; i < p.KF * S; i += ATT_THREADS) uw_s[i] = p.uw[i];
    }

    // ---- scoring: warp per row, 4 rows per warp processed as two pairs (keeps the kernel under 85 registers so
    // three CTAs share an SM: 24 warps of loads in flight, and the 320-CTA grid of the B=32, L=300 step is one wave).
    // h, Vh, w and the folded location weights are inputs of the whole decoder call, so their loads are issued
    // BEFORE the programmatic-dependency wait and may overlap the tail of the previous kernel of the chain.
    for (int i = tid; i < S; i += ATT_THREADS) w_s[i] = p.w[i];
    const float* vbase = p.Vh + ((size_t)b * p.Lmax + l0) * S + lane * 4;
    float4 v[RIF][NS];
#pragma unroll
    for (int j = 0; j < RIF; j++) {
        const int r = warp + 8 * j;
        if (r < nrows && !(p.dbg & 2)) {
#pragma unroll
            for (int i = 0; i < NS; i++) v[j][i] = ldg_stream(vbase + (size_t)r * S + i * 128);
        }
    }
    pdl_wait();                                      // q_t (and alpha_{t-1}) come from the kernels just before this one
    if (LOC) {
        for (int x = tid; x < ATT_R + p.KF - 1; x += ATT_THREADS) {
            int l = l0 + x - p.padl;
            const float a = (p.alpha_prev && l >= 0 && l < Lb) ? p.alpha_prev[(size_t)b * p.ld_aprev + l] : 0.f;
            ap_s[x] = make_float2(a, a);
        }
    }
    float4 qv[NS];
    {
        const float* qb = p.q + (size_t)b * p.ldq + lane * 4;
#pragma unroll
        for (int i = 0; i < NS; i++) qv[i] = __ldcg(reinterpret_cast<const float4*>(qb + i * 128));
    }
    __syncthreads();                                 // w_s (uw_s, ap_s) staged
#pragma unroll
    for (int pr = 0; pr < 4 / RIF; pr++) {
        if (LOC) {
            // location term: the RIF rows of this pass share every UW load (one LDS.128 per (jj, column group) instead of
            // one per row); alpha_{t-1} comes as broadcast {a, a} pairs so each row costs two FFMA2 per (jj, column group)
            float acc[RIF];
#pragma unroll
            for (int j = 0; j < RIF; j++) acc[j] = 0.f;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                f2_t z[RIF][2];
#pragma unroll
                for (int j = 0; j < RIF; j++) {
                    const float4 t = f4add(v[j][i], qv[i]);
                    z[j][0] = pack2(t.x, t.y); z[j][1] = pack2(t.z, t.w);
                }
                if constexpr (KFT > 0) {
#pragma unroll
                    for (int jj = 0; jj < KFT; jj++) {
                        const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(uw_s + jj * S + lane * 4 + i * 128);
#pragma unroll
                        for (int j = 0; j < RIF; j++) {
                            const f2_t a = *reinterpret_cast<const f2_t*>(ap_s + warp + 8 * (RIF * pr + j) + jj);
                            z[j][0] = ffma2(a, u.x, z[j][0]); z[j][1] = ffma2(a, u.y, z[j][1]);
                        }
                    }
                } else {
                    for (int jj = 0; jj < p.KF; jj++) {
                        const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(uw_s + jj * S + lane * 4 + i * 128);
#pragma unroll
                        for (int j = 0; j < RIF; j++) {
                            const f2_t a = *reinterpret_cast<const f2_t*>(ap_s + warp + 8 * (RIF * pr + j) + jj);
                            z[j][0] = ffma2(a, u.x, z[j][0]); z[j][1] = ffma2(a, u.y, z[j][1]);
                        }
                    }
                }
                const float4 wv = *reinterpret_cast<const float4*>(w_s + lane * 4 + i * 128);
#pragma unroll
                for (int j = 0; j < RIF; j++) {
                    float4 t;
                    unpack2(z[j][0], t.x, t.y); unpack2(z[j][1], t.z, t.w);
                    acc[j] = fmaf(wv.x, tanh_acc(t.x), acc[j]);
                    acc[j] = fmaf(wv.y, tanh_acc(t.y), acc[j]);
                    acc[j] = fmaf(wv.z, tanh_acc(t.z), acc[j]);
                    acc[j] = fmaf(wv.w, tanh_acc(t.w), acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < RIF; j++) {
                const int r = warp + 8 * (RIF * pr + j);
                const float a = warp_sum(acc[j]);
                if (r < nrows && lane == 0) e_s[r] = a;
            }
        } else {
#pragma unroll
            for (int j = 0; j < RIF; j++) {
                const int r = warp + 8 * (RIF * pr + j);
                if (r < nrows) {
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < NS; i++) {
                        const float4 z = f4add(v[j][i], qv[i]);
                        const float4 wv = *reinterpret_cast<const float4*>(w_s + lane * 4 + i * 128);
                        acc = fmaf(wv.x, tanh_acc(z.x), acc);
                        acc = fmaf(wv.y, tanh_acc(z.y), acc);
                        acc = fmaf(wv.z, tanh_acc(z.z), acc);
                        acc = fmaf(wv.w, tanh_acc(z.w), acc);
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) e_s[r] = acc;
                }
            }
        }
        if (RIF == 2 && pr == 0) {   // second pair of rows
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int r = warp + 8 * (2 + j);
                if (r < nrows && !(p.dbg & 2)) {
#pragma unroll
                    for (int i = 0; i < NS; i++) v[j][i] = ldg_stream(vbase + (size_t)r * S + i * 128);
                }
            }
        }
    }
    __syncthreads();

    // ---- chunk-local softmax statistics ---------------------------------------------------------
    {
        const float ev = lane < nrows ? e_s[lane] : -INFINITY;
        const float m = warp_max(ev);
        const float pv = lane < nrows ? expf(ev - m) : 0.f;
        const float ssum = warp_sum(pv);
        if (warp == 0) {
            p_s[lane] = pv;
            if (lane < nrows) p.E[(size_t)b * p.Lmax + l0 + lane] = ev;
            if (lane == 0) {
                p.part_ms[((size_t)b * p.nch + ch) * 2 + 0] = m;
                p.part_ms[((size_t)b * p.nch + ch) * 2 + 1] = ssum;
            }
        }
    }
    __syncthreads();

    // ---- partial context from the TMA-staged h tile ---------------------------------------------
    if (!(p.dbg & 4)) mbar_wait(&bar, 0);
    {
        const int g = tid / (A / 4), c4 = tid % (A / 4);
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* hs4 = reinterpret_cast<const float4*>(hs);
        for (int r = g; r < nrows; r += GROUPS) acc4 = f4fma(p_s[r], hs4[r * (A / 4) + c4], acc4);
        if (GROUPS > 1) {
            float4* red4 = reinterpret_cast<float4*>(red);
            red4[g * (A / 4) + c4] = acc4;
            __syncthreads();
            if (tid < A / 4) {
                float4 s4 = red4[tid];
#pragma unroll
                for (int gg = 1; gg < GROUPS; gg++) s4 = f4add(s4, red4[gg * (A / 4) + tid]);
                *reinterpret_cast<float4*>(p.part_c + ((size_t)b * p.nch + ch) * A + tid * 4) = s4;
            }
        } else {
            *reinterpret_cast<float4*>(p.part_c + ((size_t)b * p.nch + ch) * A + tid * 4) = acc4;
        }
    }

    // ---- ticket: the last CTA of this utterance combines ---------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&p.counters[b], 1u);
        is_last = (prev == (unsigned)(nchb - 1));
    }
    __syncthreads();
    if (!is_last || (p.dbg & 1)) { if (is_last && tid == 0) p.counters[b] = 0u; return; }
    __threadfence();

    // every warp: chunk statistics -> registers (<= 4 chunks per lane), one L2 round trip
    float mc[ATT_MAXCH / 32], sc_[ATT_MAXCH / 32];
#pragma unroll
    for (int u = 0; u < ATT_MAXCH / 32; u++) {
        const int c = lane + 32 * u;
        mc[u] = c < nchb ? ldcg1(p.part_ms + ((size_t)b * p.nch + c) * 2) : -INFINITY;
        sc_[u] = c < nchb ? ldcg1(p.part_ms + ((size_t)b * p.nch + c) * 2 + 1) : 0.f;
    }
    float M = -INFINITY;
#pragma unroll
    for (int u = 0; u < ATT_MAXCH / 32; u++) M = fmaxf(M, mc[u]);
    M = warp_max(M);
    float den = 0.f;
#pragma unroll
    for (int u = 0; u < ATT_MAXCH / 32; u++) den += sc_[u] * expf(mc[u] - M);
    den = warp_sum(den);
    const float inv = 1.0f / den;
    if (warp == 0) {
#pragma unroll
        for (int u = 0; u < ATT_MAXCH / 32; u++)
            if (lane + 32 * u < nchb) scl_s[lane + 32 * u] = expf(mc[u] - M) * inv;
    }
    __syncthreads();

    float penacc = 0.f;
    for (int l0a = 0; l0a < p.Lmax; l0a += 4 * ATT_THREADS) {
        float ev[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int l = l0a + u * ATT_THREADS + tid;
            ev[u] = l < Lb ? ldcg1(p.E + (size_t)b * p.Lmax + l) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int l = l0a + u * ATT_THREADS + tid;
            if (l < p.Lmax) {
                float a = 0.f;
                if (l < Lb) {
                    a = expf(ev[u] - M) * inv;
                    if (p.pen) {
                        const float ap = p.app ? p.app[(size_t)b * p.ld_app + l] : 0.f;
                        penacc += (float)(Lb - l) * (a - ap);
                    }
                }
                p.alpha[(size_t)b * p.ld_alpha + l] = a;
            }
        }
    }
    {   // context: thread -> (chunk group cg, float4 column); 8 independent loads in flight per thread
        constexpr int CG = ATT_THREADS / (A / 4);            // chunk groups (2 for A = 512)
        const int a4 = tid % (A / 4), cg = tid / (A / 4);
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = cg; c0 < nchb; c0 += 8 * CG) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int c = c0 + u * CG;
                t[u] = c < nchb ? ldcg4(p.part_c + ((size_t)b * p.nch + c) * A + a4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int c = c0 + u * CG;
                if (c < nchb) acc4 = f4fma(scl_s[c], t[u], acc4);
            }
        }
        if (CG > 1) {
            float4* red4 = reinterpret_cast<float4*>(red);
            __syncthreads();
            red4[cg * (A / 4) + a4] = acc4;
            __syncthreads();
            if (tid < A / 4) {
                float4 s4 = red4[tid];
#pragma unroll
                for (int gg = 1; gg < CG; gg++) s4 = f4add(s4, red4[gg * (A / 4) + tid]);
                *reinterpret_cast<float4*>(p.c + (size_t)b * p.ld_c + tid * 4) = s4;
            }
        } else {
            *reinterpret_cast<float4*>(p.c + (size_t)b * p.ld_c + a4 * 4) = acc4;
        }
    }
    if (p.pen) {
        penacc = warp_sum(penacc);
        if (lane == 0) wred[warp] = penacc;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; i++) t += wred[i];
            const bool padded = p.tlens && p.tstep >= p.tlens[b];
            p.pen[(size_t)b * p.ld_pen] = padded ? 0.f : p.lambda * fmaxf(t, 0.f);
        }
    }
    if (tid == 0) p.counters[b] = 0u;
}

// =================================================================================================
struct AttnBwdParams {
    const float *Vh, *h, *q, *w;
    int64_t ldq;
    const int* lengths;
    int B, Lmax, nch;
    int KF, padl;
    const float *uw, *alpha_prev;
    int64_t ld_aprev;
    const float *alpha, *dc, *dalpha_in, *pen;
    int64_t ld_alpha, ld_dc, ld_dain, ld_pen;
    float lambda;
    float *dalpha_s, *part_P, *part_dot, *V1;
    unsigned* counters;
    float *dq, *de, *dalpha_prev;
    int64_t ld_dq, ld_de, ld_dap;
};

template <int NS, int NA, int LOC>
__global__ void __launch_bounds__(ATT_THREADS, LOC ? 2 : 3)
attn_bwd_kernel(const AttnBwdParams p) {
    constexpr int S = NS * 128, A = NA * 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int HS = (ATT_R * A > 8 * 2 * S) ? ATT_R * A : 8 * 2 * S;
    float* hs = reinterpret_cast<float*>(smem_raw);              // [ATT_R][A]   h tile (bulk TMA)
    float* red = hs;                                             // [8][2][S]    cross-warp reduction, aliases the (then dead) h tile
    float* w_s = hs + HS;                                        // [S]
    float* dc_s = w_s + S;                                       // [A]
    float* q_s = dc_s + A;                                       // [S]
    float* uw_s = q_s + S;                                       // [KF][S]      (LOC)
    float2* ap_s = reinterpret_cast<float2*>(uw_s + (LOC ? p.KF * S : 0));   // [ATT_R+KF-1] {a, a} pairs (LOC)
    constexpr int KFT = LOC > 1 ? LOC : 0;
    __shared__ float dot_s[8];
    __shared__ uint64_t bar;
    __shared__ int is_last;

    const int b = blockIdx.y, ch = blockIdx.x;
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    const int l0 = ch * ATT_R;
    if (l0 >= Lb) return;
    const int nrows = min(ATT_R, Lb - l0);
    const int nchb = (Lb + ATT_R - 1) / ATT_R;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    pdl_trigger();
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        const unsigned bytes = (unsigned)nrows * A * 4u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(hs, p.h + ((size_t)b * p.Lmax + l0) * A, bytes, &bar);
    }
    // call-level inputs first (overlaps the previous kernel of the chain), then wait for dc_t / d alpha_t
    for (int i = tid; i < S; i += ATT_THREADS) w_s[i] = p.w[i];
    if (LOC) {
        for (int i = tid; i < p.KF * S; i += ATT_THREADS) uw_s[i] = p.uw[i];
        for (int x = tid; x < ATT_R + p.KF - 1; x += ATT_THREADS) {
            int l = l0 + x - p.padl;
            const float a = (p.alpha_prev && l >= 0 && l < Lb) ? p.alpha_prev[(size_t)b * p.ld_aprev + l] : 0.f;
            ap_s[x] = make_float2(a, a);
        }
    }
    for (int i = tid; i < S; i += ATT_THREADS) q_s[i] = p.q[(size_t)b * p.ldq + i];
    pdl_wait();
    for (int i = tid; i < A; i += ATT_THREADS) dc_s[i] = __ldcg(p.dc + (size_t)b * p.ld_dc + i);
    const float pen_g = (p.pen && p.lambda != 0.f && p.pen[(size_t)b * p.ld_pen] > 0.f) ? p.lambda : 0.f;

    float4 P1[NS], P2[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) { P1[i] = make_float4(0.f, 0.f, 0.f, 0.f); P2[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    float dotp = 0.f;

    const float* vbase = p.Vh + ((size_t)b * p.Lmax + l0) * S + lane * 4;
    // Vh rows stream through registers one at a time (three CTAs per SM hide the latency); h rows come from the
    // TMA-staged tile
    __syncthreads();            // w_s / dc_s / q_s / uw_s / ap_s staged
    mbar_wait(&bar, 0);         // h tile landed
    if constexpr (!LOC) {
        float4 vv[NS];
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const int r = warp + 8 * j;
            if (r >= nrows) break;
#pragma unroll
            for (int i = 0; i < NS; i++) vv[i] = ldg_stream(vbase + (size_t)r * S + i * 128);
            const int l = l0 + r;
            float da = 0.f;
#pragma unroll
            for (int i = 0; i < NA; i++) {
                const float4 hv = *reinterpret_cast<const float4*>(hs + (size_t)r * A + lane * 4 + i * 128);
                const float4 dv = *reinterpret_cast<const float4*>(dc_s + lane * 4 + i * 128);
                da = fmaf(hv.x, dv.x, da); da = fmaf(hv.y, dv.y, da); da = fmaf(hv.z, dv.z, da); da = fmaf(hv.w, dv.w, da);
            }
            da = warp_sum(da);
            if (p.dalpha_in) da += p.dalpha_in[(size_t)b * p.ld_dain + l];
            da += pen_g * (float)(Lb - l);
            const float a = p.alpha[(size_t)b * p.ld_alpha + l];
            const float x = a * da;
            dotp += x;
            if (lane == 0) p.dalpha_s[(size_t)b * p.Lmax + l] = da;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                const float4 z = f4add(vv[i], *reinterpret_cast<const float4*>(q_s + lane * 4 + i * 128));
                const float4 wv = *reinterpret_cast<const float4*>(w_s + lane * 4 + i * 128);
                float4 g;
                { float t = tanh_acc(z.x); g.x = wv.x * (1.f - t * t); }
                { float t = tanh_acc(z.y); g.y = wv.y * (1.f - t * t); }
                { float t = tanh_acc(z.z); g.z = wv.z * (1.f - t * t); }
                { float t = tanh_acc(z.w); g.w = wv.w * (1.f - t * t); }
                P1[i] = f4fma(x, g, P1[i]);
                P2[i] = f4fma(a, g, P2[i]);
            }
        }
    } else {
        // location path: two rows per pass share every UW load (the kernel is shared-memory-bandwidth bound otherwise:
        // 2 KF LDS.128 per row and column group); alpha_{t-1} windows and the V1 partial sums live in registers
        float4 vv[2][NS];
#pragma unroll 1
        for (int j = 0; j < 4; j += 2) {
            const int r0 = warp + 8 * j;
            if (r0 >= nrows) break;
            const int r1 = r0 + 8;
            const bool ok1 = r1 < nrows;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                vv[0][i] = ldg_stream(vbase + (size_t)r0 * S + i * 128);
                vv[1][i] = ok1 ? ldg_stream(vbase + (size_t)r1 * S + i * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float a2[2], x2[2];
            constexpr int NV = KFT > 0 ? KFT : LOC_MAXKF;
            float v1a[2][NV];                                  // V1 partial sums of the two rows
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int r = k ? r1 : r0;
                const bool ok = k ? ok1 : true;
                const int l = l0 + r;
                float da = 0.f;
                if (ok) {
#pragma unroll
                    for (int i = 0; i < NA; i++) {
                        const float4 hv = *reinterpret_cast<const float4*>(hs + (size_t)r * A + lane * 4 + i * 128);
                        const float4 dv = *reinterpret_cast<const float4*>(dc_s + lane * 4 + i * 128);
                        da = fmaf(hv.x, dv.x, da); da = fmaf(hv.y, dv.y, da); da = fmaf(hv.z, dv.z, da); da = fmaf(hv.w, dv.w, da);
                    }
                }
                da = warp_sum(da);
                float a = 0.f;
                if (ok) {
                    if (p.dalpha_in) da += p.dalpha_in[(size_t)b * p.ld_dain + l];
                    da += pen_g * (float)(Lb - l);
                    a = p.alpha[(size_t)b * p.ld_alpha + l];
                    if (lane == 0) p.dalpha_s[(size_t)b * p.Lmax + l] = da;
                } else {
                    da = 0.f;
                }
                a2[k] = a; x2[k] = a * da;
                dotp += x2[k];
#pragma unroll
                for (int jj = 0; jj < NV; jj++) v1a[k][jj] = 0.f;
            }
            const int r1c = ok1 ? r1 : r0;                     // a missing second row recomputes the first (its results are dropped)
#pragma unroll
            for (int i = 0; i < NS; i++) {
                const float4 qv = *reinterpret_cast<const float4*>(q_s + lane * 4 + i * 128);
                float4 z0 = f4add(vv[0][i], qv), z1 = f4add(vv[1][i], qv);
#pragma unroll
                for (int jj = 0; jj < NV; jj++) {
                    if (KFT > 0 || jj < p.KF) {
                        const float4 u = *reinterpret_cast<const float4*>(uw_s + jj * S + lane * 4 + i * 128);
                        z0 = f4fma(ap_s[r0 + jj].x, u, z0);             // broadcast loads: one wavefront each
                        z1 = f4fma(ap_s[r1c + jj].x, u, z1);
                    }
                }
                const float4 wv = *reinterpret_cast<const float4*>(w_s + lane * 4 + i * 128);
                float4 g0, g1;
                { float t = tanh_acc(z0.x); g0.x = wv.x * (1.f - t * t); }
                { float t = tanh_acc(z0.y); g0.y = wv.y * (1.f - t * t); }
                { float t = tanh_acc(z0.z); g0.z = wv.z * (1.f - t * t); }
                { float t = tanh_acc(z0.w); g0.w = wv.w * (1.f - t * t); }
                { float t = tanh_acc(z1.x); g1.x = wv.x * (1.f - t * t); }
                { float t = tanh_acc(z1.y); g1.y = wv.y * (1.f - t * t); }
                { float t = tanh_acc(z1.z); g1.z = wv.z * (1.f - t * t); }
                { float t = tanh_acc(z1.w); g1.w = wv.w * (1.f - t * t); }
                P1[i] = f4fma(x2[0], g0, P1[i]); P1[i] = f4fma(x2[1], g1, P1[i]);
                P2[i] = f4fma(a2[0], g0, P2[i]); P2[i] = f4fma(a2[1], g1, P2[i]);
#pragma unroll
                for (int jj = 0; jj < NV; jj++) {
                    if (KFT > 0 || jj < p.KF) {
                        const float4 u = *reinterpret_cast<const float4*>(uw_s + jj * S + lane * 4 + i * 128);
                        v1a[0][jj] = fmaf(g0.x, u.x, fmaf(g0.y, u.y, fmaf(g0.z, u.z, fmaf(g0.w, u.w, v1a[0][jj]))));
                        v1a[1][jj] = fmaf(g1.x, u.x, fmaf(g1.y, u.y, fmaf(g1.z, u.z, fmaf(g1.w, u.w, v1a[1][jj]))));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float v1[LOC_MAXKF];
#pragma unroll
                for (int jj = 0; jj < LOC_MAXKF; jj++) v1[jj] = jj < NV ? v1a[k][jj < NV ? jj : 0] : 0.f;
                // 16 values over 32 lanes: transpose-reduce; lane (x & 15) ends with value index (x & 15)
#pragma unroll
                for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
                    for (int i = 0; i < s; i++) {
                        const float send = (lane & s) ? v1[i] : v1[i + s];
                        const float keep = (lane & s) ? v1[i + s] : v1[i];
                        v1[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                    }
                }
                v1[0] += __shfl_xor_sync(0xffffffffu, v1[0], 16);
                const int r = k ? r1 : r0;
                if ((k == 0 || ok1) && lane < p.KF) p.V1[((size_t)b * p.Lmax + l0 + r) * p.KF + lane] = v1[0];
            }
        }
    }

    // ---- cross-warp reduction of P1 / P2 / dot (red aliases the h tile: wait until every warp is done with it) ----
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NS; i++) {
        *reinterpret_cast<float4*>(red + (warp * 2 + 0) * S + lane * 4 + i * 128) = P1[i];
        *reinterpret_cast<float4*>(red + (warp * 2 + 1) * S + lane * 4 + i * 128) = P2[i];
    }
    if (lane == 0) dot_s[warp] = dotp;
    __syncthreads();
    for (int idx = tid; idx < 2 * S; idx += ATT_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int wg = 0; wg < 8; wg++) s += red[wg * 2 * S + idx];
        p.part_P[((size_t)b * p.nch + ch) * 2 * S + idx] = s;
    }
    if (tid == 0) {
        float s = 0.f;
        for (int wg = 0; wg < 8; wg++) s += dot_s[wg];
        p.part_dot[(size_t)b * p.nch + ch] = s;
    }

    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&p.counters[b], 1u);
        is_last = (prev == (unsigned)(nchb - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last CTA of the utterance: dq, de, d alpha_{t-1}; independent loads batched 8 deep -------------
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < ATT_MAXCH / 32; u++) {
        const int c = lane + 32 * u;
        dot += c < nchb ? ldcg1(p.part_dot + (size_t)b * p.nch + c) : 0.f;
    }
    dot = warp_sum(dot);
    for (int i = tid; i < S; i += ATT_THREADS) {
        float s1 = 0.f, s2 = 0.f;
        for (int c0 = 0; c0 < nchb; c0 += 8) {
            float t1[8], t2[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool ok = c0 + u < nchb;
                t1[u] = ok ? ldcg1(p.part_P + ((size_t)b * p.nch + c0 + u) * 2 * S + i) : 0.f;
                t2[u] = ok ? ldcg1(p.part_P + ((size_t)b * p.nch + c0 + u) * 2 * S + S + i) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) { s1 += t1[u]; s2 += t2[u]; }
        }
        p.dq[(size_t)b * p.ld_dq + i] = s1 - dot * s2;
    }
    for (int l0a = 0; l0a < p.Lmax; l0a += 4 * ATT_THREADS) {
        float av[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int l = l0a + u * ATT_THREADS + tid;
            av[u] = l < Lb ? p.alpha[(size_t)b * p.ld_alpha + l] : 0.f;
            dv[u] = l < Lb ? ldcg1(p.dalpha_s + (size_t)b * p.Lmax + l) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int l = l0a + u * ATT_THREADS + tid;
            if (l < p.Lmax) p.de[(size_t)b * p.ld_de + l] = l < Lb ? av[u] * (dv[u] - dot) : 0.f;
        }
    }
    if (p.dalpha_prev) {
        for (int l = tid; l < p.Lmax; l += ATT_THREADS) {
            float d = 0.f;
            if (l < Lb) {
                d = -pen_g * (float)(Lb - l);
                if (LOC) {
                    for (int jj = 0; jj < p.KF; jj++) {
                        const int x = l - jj + p.padl;      // frame whose window position jj looks at l
                        if (x >= 0 && x < Lb) {
                            const float dex = p.alpha[(size_t)b * p.ld_alpha + x] * (ldcg1(p.dalpha_s + (size_t)b * p.Lmax + x) - dot);
                            d = fmaf(dex, ldcg1(p.V1 + ((size_t)b * p.Lmax + x) * p.KF + jj), d);
                        }
                    }
                }
            }
            p.dalpha_prev[(size_t)b * p.ld_dap + l] = d;
        }
    }
    if (tid == 0) p.counters[b] = 0u;
}

// =================================================================================================
// deferred dVh / dw_e / dUW: thread per score column, DVH_R frames per CTA held in registers,
// time loop outermost so q_t is read once per step.
struct DvhParams {
    const float *Vh, *q_all, *de_all, *w;
    const int *lengths, *tlens;
    int B, Lmax, T, S;
    int KF, padl;
    const float *uw, *alpha_all;    // alpha_all [B,T,Lmax]: alpha_t (alpha_{t-1} is row t-1; zeros for t = 0)
    float *dVh, *dwe, *duw;
};

// LOC: 0 = content attention, 1 = location term with a run-time filter size (<= LOC_MAXKF), > 1 = exact filter size
template <int LOC>
__global__ void __launch_bounds__(128)
attn_dvh_kernel(const DvhParams p) {
    constexpr int KFC = LOC > 1 ? LOC : LOC_MAXKF;            // filter taps held in registers
    extern __shared__ float sm[];
    float* de_s = sm;                                         // [T][DVH_R]
    float* ap_s = de_s + p.T * DVH_R;                         // [T][DVH_R+KF-1]  (LOC)
    const int b = blockIdx.y, l0 = blockIdx.x * DVH_R, i = blockIdx.z * 128 + threadIdx.x;
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    const int Tb = p.tlens ? min(p.tlens[b], p.T) : p.T;
    const int S = p.S;
    float* out = p.dVh + ((size_t)b * p.Lmax + l0) * S + i;
    if (l0 >= Lb) {
        for (int r = 0; r < DVH_R && l0 + r < p.Lmax; r++) out[(size_t)r * S] = 0.f;
        return;
    }
    const int nrows = min(DVH_R, Lb - l0);
    for (int e = threadIdx.x; e < Tb * DVH_R; e += 128) {
        const int t = e / DVH_R, r = e % DVH_R;
        de_s[e] = r < nrows ? p.de_all[((size_t)b * p.T + t) * p.Lmax + l0 + r] : 0.f;
    }
    const int W = DVH_R + p.KF - 1;
    if (LOC) {
        for (int e = threadIdx.x; e < Tb * W; e += 128) {
            const int t = e / W, x = e % W, l = l0 + x - p.padl;
            ap_s[e] = (t > 0 && l >= 0 && l < Lb) ? p.alpha_all[((size_t)b * p.T + t - 1) * p.Lmax + l] : 0.f;
        }
    }
    __syncthreads();

    float v[DVH_R], acc[DVH_R];
    const float* vb = p.Vh + ((size_t)b * p.Lmax + l0) * S + i;
#pragma unroll
    for (int r = 0; r < DVH_R; r++) { v[r] = r < nrows ? vb[(size_t)r * S] : 0.f; acc[r] = 0.f; }
    const float wi = p.w[i];
    float dw = 0.f;
    float uwr[LOC ? KFC : 1], duwr[LOC ? KFC : 1];
    if (LOC) {
#pragma unroll
        for (int jj = 0; jj < KFC; jj++) { uwr[jj] = jj < p.KF ? p.uw[jj * S + i] : 0.f; duwr[jj] = 0.f; }
    }
    const float* qb = p.q_all + (size_t)b * p.T * S + i;
    for (int t = 0; t < Tb; t++) {
        const float q = qb[(size_t)t * S];
        // the alpha_{t-1} window of this CTA's frames: read once per step (warp-uniform broadcasts) and reused by every frame and tap
        // (one LDS per FMA before: the kernel was bound by shared-memory instruction issue); taps past KF have zero weights
        float a[LOC ? DVH_R + KFC - 1 : 1];
        if (LOC) {
#pragma unroll
            for (int x = 0; x < DVH_R + KFC - 1; x++) a[x] = (LOC > 1 || x < W) ? ap_s[t * W + x] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < DVH_R; r++) {
            const float d = de_s[t * DVH_R + r];
            float z = q + v[r];
            if (LOC) {
#pragma unroll
                for (int jj = 0; jj < KFC; jj++) z = fmaf(uwr[jj], a[r + jj], z);
            }
            const float th = tanh_acc(z);
            const float dz = d * wi * (1.f - th * th);
            acc[r] += dz;
            dw = fmaf(d, th, dw);
            if (LOC) {
#pragma unroll
                for (int jj = 0; jj < KFC; jj++) duwr[jj] = fmaf(dz, a[r + jj], duwr[jj]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < DVH_R; r++)
        if (l0 + r < p.Lmax) out[(size_t)r * S] = r < nrows ? acc[r] : 0.f;
    atomicAdd(p.dwe + i, dw);
    if (LOC) {
#pragma unroll
        for (int jj = 0; jj < KFC; jj++)
            if (jj < p.KF) atomicAdd(p.duw + jj * S + i, duwr[jj]);
    }
}

// =================================================================================================
// V1[b,t,l,j] = sum_s UW[j,s] w_s (1 - tanh^2 Z_t[l,s]) for ALL decoder steps at once (location path).  It is the Jacobian of the
// energy e_t[l] with respect to alpha_{t-1}[l + j - pad_left] (Attention.lua:86-99 reversed): the alignment carry of the backward
// time loop is  dalpha_{t-1}[i] = sum_j de_t[i - j + pad_left] V1[t, i - j + pad_left, j].  V1 does not depend on any gradient, so it
// is hoisted out of the sequential loop: one throughput-bound launch over all (b, t, l) instead of a shuffle reduction inside every
// step of the latency-bound loop.  Warp = two consecutive frames (they share every UW load and all but one value of the alpha_{t-1}
// window), Vh rows stay in registers across the T steps, S = 512.
struct V1Params {
    const float *Vh, *q_all, *w, *uw, *alpha_all;
    const int *lengths, *tlens;
    int B, Lmax, T, KF, padl;
    float* V1;                   // [B, T, Lmax, KF]
};
constexpr int V1_F = 4;                  // frames per warp: the 12 shared-memory vectors of one s-quad (q, w, 10 taps of U W_F) serve 4 frames
constexpr int V1_R = 8 * V1_F;           // frames per CTA
constexpr int V1_S = 512;

template <int KFT>               // exact filter size (10) or 0 = run-time size <= LOC_MAXKF
__global__ void __launch_bounds__(256, 1)
attn_v1_kernel(const V1Params p) {
    constexpr int NT = KFT > 0 ? KFT : LOC_MAXKF;
    constexpr int S = V1_S, F = V1_F;
    extern __shared__ __align__(16) float v1_sm[];
    float* uw_s = v1_sm;                                   // [NT][S]   (taps past KF: zero weights)
    float* w_s = uw_s + NT * S;                            // [S]
    float* q_s = w_s + S;                                  // [2][S]
    float* ap_s = q_s + 2 * S;                             // [T][V1_R + NT]   alpha_{t-1}[l0 - padl + x]
    const int b = blockIdx.y, l0 = blockIdx.x * V1_R;
    const int Lb = p.lengths ? p.lengths[b] : p.Lmax;
    const int Tb = p.tlens ? min(p.tlens[b], p.T) : p.T;
    if (l0 >= Lb || Tb < 2) return;
    // the steps are independent: blockIdx.z takes an equal share of [1, Tb) so that the grid is several waves deep (B * L / 32 frame
    // blocks alone are ~2 waves of long CTAs: a third of the machine would idle through the last one)
    const int tA = 1 + (int)(((long long)(Tb - 1) * blockIdx.z) / gridDim.z), tB = 1 + (int)(((long long)(Tb - 1) * (blockIdx.z + 1)) / gridDim.z);
    if (tA >= tB) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int W = V1_R + NT;
    for (int i = tid; i < NT * S; i += 256) uw_s[i] = i < p.KF * S ? p.uw[i] : 0.f;
    for (int i = tid; i < S; i += 256) w_s[i] = p.w[i];
    for (int e = tid; e < (tB - tA) * W; e += 256) {
        const int t = tA + e / W, x = e % W, l = l0 + x - p.padl;
        ap_s[e] = (l >= 0 && l < Lb) ? p.alpha_all[((size_t)b * p.T + t - 1) * p.Lmax + l] : 0.f;
    }
    const int r0 = F * warp, la = l0 + r0;
    float4 vv[F][4];
#pragma unroll
    for (int f = 0; f < F; f++)
#pragma unroll
        for (int i = 0; i < 4; i++)
            vv[f][i] = la + f < Lb ? ldg_stream(p.Vh + ((size_t)b * p.Lmax + la + f) * S + lane * 4 + 128 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* qb = p.q_all + (size_t)b * p.T * S;
    // q_t is staged through registers one step ahead: the global load is issued before the step's arithmetic and stored behind it
    float qn0 = qb[(size_t)tA * S + tid], qn1 = qb[(size_t)tA * S + 256 + tid];
    q_s[(tA & 1) * S + tid] = qn0; q_s[(tA & 1) * S + 256 + tid] = qn1;
    __syncthreads();
    if (la >= Lb) {                                        // warp without frames: only helps staging q
        for (int t = tA; t < tB; t++) {
            if (t + 1 < tB) { q_s[((t + 1) & 1) * S + tid] = qb[(size_t)(t + 1) * S + tid]; q_s[((t + 1) & 1) * S + 256 + tid] = qb[(size_t)(t + 1) * S + 256 + tid]; }
            __syncthreads();
        }
        return;
    }
    for (int t = tA; t < tB; t++) {
        const float* qs = q_s + (t & 1) * S;
        if (t + 1 < tB) { qn0 = qb[(size_t)(t + 1) * S + tid]; qn1 = qb[(size_t)(t + 1) * S + 256 + tid]; }
        float a[NT + F - 1];
#pragma unroll
        for (int x = 0; x < NT + F - 1; x++) a[x] = ap_s[(t - tA) * W + r0 + x];
        float v1a[F][NT];
#pragma unroll
        for (int f = 0; f < F; f++)
#pragma unroll
            for (int jj = 0; jj < NT; jj++) v1a[f][jj] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float4 qv = *reinterpret_cast<const float4*>(qs + lane * 4 + 128 * i);
            float4 z[F];
#pragma unroll
            for (int f = 0; f < F; f++) z[f] = f4add(vv[f][i], qv);
            float4 u[NT];
#pragma unroll
            for (int jj = 0; jj < NT; jj++) {
                u[jj] = *reinterpret_cast<const float4*>(uw_s + jj * S + lane * 4 + 128 * i);
#pragma unroll
                for (int f = 0; f < F; f++) z[f] = f4fma(a[f + jj], u[jj], z[f]);
            }
            const float4 wv = *reinterpret_cast<const float4*>(w_s + lane * 4 + 128 * i);
#pragma unroll
            for (int f = 0; f < F; f++) {
                float4 g;
                { float th = tanh_acc(z[f].x); g.x = wv.x * (1.f - th * th); }
                { float th = tanh_acc(z[f].y); g.y = wv.y * (1.f - th * th); }
                { float th = tanh_acc(z[f].z); g.z = wv.z * (1.f - th * th); }
                { float th = tanh_acc(z[f].w); g.w = wv.w * (1.f - th * th); }
#pragma unroll
                for (int jj = 0; jj < NT; jj++)
                    v1a[f][jj] = fmaf(g.x, u[jj].x, fmaf(g.y, u[jj].y, fmaf(g.z, u[jj].z, fmaf(g.w, u[jj].w, v1a[f][jj]))));
            }
        }
        // F x 16 partial sums over 32 lanes, two frames at a time: transposed butterfly, lane x ends with value (x & 15) of frame (x >> 4)
#pragma unroll
        for (int fp = 0; fp < F; fp += 2) {
            float v[16];
#pragma unroll
            for (int jj = 0; jj < 16; jj++) {
                const float s0 = jj < NT ? v1a[fp][jj < NT ? jj : 0] : 0.f, s1 = jj < NT ? v1a[fp + 1][jj < NT ? jj : 0] : 0.f;
                const float send = (lane & 16) ? s0 : s1, keep = (lane & 16) ? s1 : s0;
                v[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int sft = 8; sft >= 1; sft >>= 1) {
#pragma unroll
                for (int i = 0; i < sft; i++) {
                    const float send = (lane & sft) ? v[i] : v[i + sft];
                    const float keep = (lane & sft) ? v[i + sft] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                }
            }
            const int fr = fp + (lane >> 4), jj = lane & 15;
            if (la + fr < Lb && jj < p.KF) p.V1[(((size_t)b * p.T + t) * p.Lmax + la + fr) * p.KF + jj] = v[0];
        }
        if (t + 1 < tB) { q_s[((t + 1) & 1) * S + tid] = qn0; q_s[((t + 1) & 1) * S + 256 + tid] = qn1; }
        __syncthreads();                                    // q_{t+1} staged; q_t buffer free for t+2
    }
}

int attn_v1(s2s_ctx* ctx, const float* Vh, const float* q_all, const float* w, const float* uw, const float* alpha_all, const int* lengths,
            const int* tlens, int B, int Lmax, int T, int S, int KF, int padl, float* V1) {
    S2S_REQUIRE(S == V1_S && KF >= 1 && KF <= LOC_MAXKF, "attn_v1: needs S = 512 and a filter of 1..16 taps");
    V1Params p;
    p.Vh = Vh; p.q_all = q_all; p.w = w; p.uw = uw; p.alpha_all = alpha_all; p.lengths = lengths; p.tlens = tlens;
    p.B = B; p.Lmax = Lmax; p.T = T; p.KF = KF; p.padl = padl; p.V1 = V1;
    const int NT = KF == 10 ? 10 : LOC_MAXKF;
    const size_t smem = ((size_t)NT * V1_S + 3 * V1_S + (size_t)T * (V1_R + NT)) * sizeof(float);
    S2S_REQUIRE(smem <= 200 * 1024, "attn_v1: T=%d too large for the shared-memory staging", T);
    dim3 grid(ceil_div(Lmax, V1_R), B, T >= 12 ? 3 : 1);
    prof_begin(ctx, S2S_PROF_ATTN_DVH);
    if (KF == 10) {
        S2S_CUDA(cudaFuncSetAttribute(attn_v1_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_v1_kernel<10><<<grid, 256, smem, ctx->stream>>>(p);
    } else {
        S2S_CUDA(cudaFuncSetAttribute(attn_v1_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_v1_kernel<0><<<grid, 256, smem, ctx->stream>>>(p);
    }
    prof_end(ctx, S2S_PROF_ATTN_DVH, 4.0 * B * ((double)Lmax * S + (double)T * S + (double)T * Lmax * (KF + 1)));   // one read of Vh, one write of V1
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// =================================================================================================
// host wrappers
// =================================================================================================
int attn_scratch_alloc(s2s_ctx* ctx, Arena& arena, int B, int Lmax, int S, int A, int KF, bool backward, AttnScratch* sc) {
    S2S_REQUIRE(Lmax <= ATT_MAXCH * ATT_R, "Lmax=%d exceeds the attention kernel limit %d", Lmax, ATT_MAXCH * ATT_R);
    S2S_REQUIRE(B <= 4000, "B=%d exceeds the ticket-counter capacity 4000 (counters[4000..] are the dense-chain grid barrier, decoder.cu)", B);
    sc->nch = ceil_div(Lmax, ATT_R);
    S2S_ALLOC(sc->E, arena, float, (size_t)B * Lmax);
    S2S_ALLOC(sc->part_ms, arena, float, (size_t)B * sc->nch * 2);
    S2S_ALLOC(sc->part_c, arena, float, (size_t)B * sc->nch * A);
    if (backward) {
        S2S_ALLOC(sc->dalpha, arena, float, (size_t)B * Lmax);
        S2S_ALLOC(sc->part_P, arena, float, (size_t)B * sc->nch * 2 * S);
        S2S_ALLOC(sc->part_dot, arena, float, (size_t)B * sc->nch);
        if (KF > 0) S2S_ALLOC(sc->V1, arena, float, (size_t)B * Lmax * KF);
    }
    sc->counters = ctx->counters;
    return 0;
}

template <int NS, int NA, int LOC, int RIF>
static int launch_fwd_rif(s2s_ctx* ctx, const AttnFwdParams& p, int KF) {
    size_t smem = (size_t)ATT_R * NA * 128 * 4 + (size_t)NS * 128 * 4 + (LOC ? ((size_t)KF * NS * 128 + 2 * (ATT_R + KF)) * 4 : 0);
    static bool attr_set = false;   // per instantiation
    static size_t attr_smem = 0;
    if (!attr_set || smem > attr_smem) {
        S2S_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NS, NA, LOC, RIF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true; attr_smem = smem;
    }
    prof_begin(ctx, S2S_PROF_ATTN_FWD);
    S2S_CUDA(launch_kernel(attn_fwd_kernel<NS, NA, LOC, RIF>, dim3(p.nch, p.B), dim3(ATT_THREADS), smem, ctx->stream, ctx->pdl, p));
    {   // algorithmic bytes A_f = 4 B (L S + L A + 2L + S + A)   
        const double S = NS * 128.0, A = NA * 128.0, L = p.Lmax;
        prof_end(ctx, S2S_PROF_ATTN_FWD, 4.0 * p.B * (L * S + L * A + 2 * L + S + A));
    }
    ctx->kcount[S2S_KC_ATTN_STEP]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
// Small grids (the L2-resident decoder step) are latency-bound: three light CTAs per SM.  Large grids are
// bandwidth-bound: two CTAs per SM with all four rows of every warp in flight (measured 84-91% of HBM peak).
template <int NS, int NA, int LOC>
static int launch_fwd(s2s_ctx* ctx, const AttnFwdParams& p, int KF) {
    if constexpr (!LOC) {   // the location path keeps alpha_{t-1} windows in registers: two rows in flight only
        if ((long)p.nch * p.B >= 6L * ctx->sm_count) return launch_fwd_rif<NS, NA, LOC, 4>(ctx, p, KF);
    }
    return launch_fwd_rif<NS, NA, LOC, 2>(ctx, p, KF);
}
template <int NS, int NA, int LOC>
static int launch_bwd(s2s_ctx* ctx, const AttnBwdParams& p, int KF) {
    size_t hsz = (size_t)ATT_R * NA * 128 > (size_t)16 * NS * 128 ? (size_t)ATT_R * NA * 128 : (size_t)16 * NS * 128;
    size_t smem = (hsz + (size_t)2 * NS * 128 + (size_t)NA * 128) * 4 + (LOC ? ((size_t)KF * NS * 128 + 2 * (ATT_R + KF)) * 4 : 0);
    static bool attr_set = false;
    static size_t attr_smem = 0;
    if (!attr_set || smem > attr_smem) {
        S2S_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<NS, NA, LOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true; attr_smem = smem;
    }
    prof_begin(ctx, S2S_PROF_ATTN_BWD);
    S2S_CUDA(launch_kernel(attn_bwd_kernel<NS, NA, LOC>, dim3(p.nch, p.B), dim3(ATT_THREADS), smem, ctx->stream, ctx->pdl, p));
    {   // A_b,min = 4 B (L S + L A + 6L + 2S + 2A)   (deferred accumulation)
        const double S = NS * 128.0, A = NA * 128.0, L = p.Lmax;
        prof_end(ctx, S2S_PROF_ATTN_BWD, 4.0 * p.B * (L * S + L * A + 6 * L + 2 * S + 2 * A));
    }
    ctx->kcount[S2S_KC_ATTN_STEP]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

#define ATT_DISPATCH(FN, LOCV, ...)                                                               \
    do {                                                                                          \
        const int ns__ = S / 128, na__ = A / 128;                                                 \
        if (ns__ == 1 && na__ == 1) return FN<1, 1, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 1 && na__ == 2) return FN<1, 2, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 2 && na__ == 1) return FN<2, 1, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 2 && na__ == 2) return FN<2, 2, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 2 && na__ == 4) return FN<2, 4, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 4 && na__ == 2) return FN<4, 2, LOCV>(__VA_ARGS__);                           \
        if (ns__ == 4 && na__ == 4) return FN<4, 4, LOCV>(__VA_ARGS__);                           \
        return fail("attention kernels: unsupported (S=%d, A=%d); supported S,A in {128,256,512} with |log2(S/A)|<=1", S, A); \
    } while (0)

static int check_dims(int S, int A, int KF) {
    S2S_REQUIRE(S % 128 == 0 && A % 128 == 0 && S >= 128 && A >= 128, "attention kernels need S and A to be multiples of 128 (S=%d A=%d)", S, A);
    S2S_REQUIRE(KF >= 0 && KF <= LOC_MAXKF, "location filter size %d not supported (max %d)", KF, LOC_MAXKF);
    return 0;
}

int attn_step_fwd(s2s_ctx* ctx, const AttnScratch& sc, const float* Vh, const float* h, const float* q, int64_t ldq, const float* w,
                  const int* lengths, int B, int Lmax, int S, int A, const AttnLoc& loc, float* alpha, int64_t ld_alpha, float* c,
                  int64_t ld_c, float* pen, int64_t ld_pen, float lambda, const float* app, int64_t ld_app, const int* tlens, int tstep) {
    S2S_TRY(check_dims(S, A, loc.KF));
    AttnFwdParams p;
    p.Vh = Vh; p.h = h; p.q = q; p.w = w; p.ldq = ldq; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.nch = sc.nch;
    p.KF = loc.KF; p.padl = loc.padl; p.uw = loc.uw; p.alpha_prev = loc.alpha_prev; p.ld_aprev = loc.ld_aprev;
    p.E = sc.E; p.part_ms = sc.part_ms; p.part_c = sc.part_c; p.counters = sc.counters;
    p.alpha = alpha; p.c = c; p.pen = pen; p.ld_alpha = ld_alpha; p.ld_c = ld_c; p.ld_pen = ld_pen; p.lambda = lambda;
    p.app = app; p.ld_app = ld_app; p.tlens = tlens; p.tstep = tstep;
    { static const int dbg = []() { const char* e = getenv("S2S_ATT_DBG"); return e ? atoi(e) : 0; }(); p.dbg = dbg; }
    if (loc.KF == 10) ATT_DISPATCH(launch_fwd, 10, ctx, p, loc.KF);      // the reference's default filter size (Attention.lua:17)
    else if (loc.KF > 0) ATT_DISPATCH(launch_fwd, 1, ctx, p, loc.KF);
    else ATT_DISPATCH(launch_fwd, 0, ctx, p, 0);
}

int attn_step_bwd(s2s_ctx* ctx, const AttnScratch& sc, const float* Vh, const float* h, const float* q, int64_t ldq, const float* w,
                  const int* lengths, int B, int Lmax, int S, int A, const AttnLoc& loc, const float* alpha, int64_t ld_alpha,
                  const float* dc, int64_t ld_dc, const float* dalpha_in, int64_t ld_dain, const float* pen, int64_t ld_pen,
                  float lambda, float* dq, int64_t ld_dq, float* de, int64_t ld_de, float* dalpha_prev, int64_t ld_dap) {
    S2S_TRY(check_dims(S, A, loc.KF));
    S2S_REQUIRE(sc.dalpha && sc.part_P, "attention scratch was not allocated for backward");
    AttnBwdParams p;
    p.Vh = Vh; p.h = h; p.q = q; p.w = w; p.ldq = ldq; p.lengths = lengths; p.B = B; p.Lmax = Lmax; p.nch = sc.nch;
    p.KF = loc.KF; p.padl = loc.padl; p.uw = loc.uw; p.alpha_prev = loc.alpha_prev; p.ld_aprev = loc.ld_aprev;
    p.alpha = alpha; p.dc = dc; p.dalpha_in = dalpha_in; p.pen = pen;
    p.ld_alpha = ld_alpha; p.ld_dc = ld_dc; p.ld_dain = ld_dain; p.ld_pen = ld_pen; p.lambda = lambda;
    p.dalpha_s = sc.dalpha; p.part_P = sc.part_P; p.part_dot = sc.part_dot; p.V1 = sc.V1; p.counters = sc.counters;
    p.dq = dq; p.de = de; p.dalpha_prev = dalpha_prev; p.ld_dq = ld_dq; p.ld_de = ld_de; p.ld_dap = ld_dap;
    if (loc.KF == 10) ATT_DISPATCH(launch_bwd, 10, ctx, p, loc.KF);
    else if (loc.KF > 0) ATT_DISPATCH(launch_bwd, 1, ctx, p, loc.KF);
    else ATT_DISPATCH(launch_bwd, 0, ctx, p, 0);
}

int attn_dvh(s2s_ctx* ctx, const float* Vh, const float* q_all, const float* de_all, const float* w, const int* lengths,
             const int* tlens, int B, int Lmax, int T, int S, const AttnLoc& loc, float* dVh, float* dwe, float* duw) {
    S2S_REQUIRE(S % 128 == 0, "attn_dvh: S must be a multiple of 128");
    DvhParams p;
    p.Vh = Vh; p.q_all = q_all; p.de_all = de_all; p.w = w; p.lengths = lengths; p.tlens = tlens;
    p.B = B; p.Lmax = Lmax; p.T = T; p.S = S; p.KF = loc.KF; p.padl = loc.padl; p.uw = loc.uw; p.alpha_all = loc.alpha_prev;
    p.dVh = dVh; p.dwe = dwe; p.duw = duw;
    size_t smem = ((size_t)T * DVH_R + (loc.KF > 0 ? (size_t)T * (DVH_R + loc.KF - 1) : 0)) * 4;
    S2S_REQUIRE(smem <= 200 * 1024, "attn_dvh: T=%d too large for the shared-memory staging", T);
    dim3 grid(ceil_div(Lmax, DVH_R), B, S / 128);
    prof_begin(ctx, S2S_PROF_ATTN_DVH);
    if (loc.KF == 10) {
        S2S_CUDA(cudaFuncSetAttribute(attn_dvh_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_dvh_kernel<10><<<grid, 128, smem, ctx->stream>>>(p);
    } else if (loc.KF > 0) {
        S2S_CUDA(cudaFuncSetAttribute(attn_dvh_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_dvh_kernel<1><<<grid, 128, smem, ctx->stream>>>(p);
    } else {
        if (smem > 48 * 1024) S2S_CUDA(cudaFuncSetAttribute(attn_dvh_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_dvh_kernel<0><<<grid, 128, smem, ctx->stream>>>(p);
    }
    prof_end(ctx, S2S_PROF_ATTN_DVH, 4.0 * B * ((double)2 * Lmax * S + (double)T * S + (double)T * Lmax));   // one read of Vh, one write of dVh
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s
