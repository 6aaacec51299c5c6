// decoder.cu -- nn.Attention = Vh precompute + nn.RNNAttention(nn.Recurrent(decoder_base_)), teacher
// forced, forward and backward (Attention.lua:39-211,305-327; RNNAttention.lua:144-253;
// Recurrent.lua:104-151; GRU.lua:22-30; Maxout.lua:14-19; model_chorowski_baseline.lua:48-59).
//
// What is sequential in t stays in the time loop (Ws s_{t-1}, the attention step, c->u, the decoder
// GRU); everything teacher forcing makes independent of the recurrence is hoisted into time-batched
// GEMMs over M = B*T rows: the y_{t-1} input path, the Maxout MLP + LogSoftMax, and every weight
// gradient.  The per-step [L,S]/[L,A] gradient read-modify-write of RNNAttention.lua:247 is replaced
// by the deferred accumulation kernels of attention.cu.
#include "decoder.cuh"
#include "model.cuh"

namespace s2s {

// =================================================================================================
// dense_small: Y[b, n] = epi( sum_k X[b,k] W[n,k] )  for a handful of rows b (the minibatch) --
// the per-step matrix-vector products of the decoder (Linear / LinearZeroBias of Attention.lua:65-67,
// 149-151 and the decoder GRU, GRU.lua:22-30, applied to B rows at once).
// Mapping: lane = minibatch row, warp = one eighth of K.  Each lane keeps ITS row's K-slice in registers
// (loaded once, straight from global), the CTA's NT weight rows are staged in shared memory and read as
// warp-uniform broadcasts (one pass per LDS.128), so there is no shuffle reduction and no re-read of X;
// the eight K-slices are summed through shared memory and the GRU gate math is fused into the epilogue.
// =================================================================================================
enum { EPI_LINEAR = 0, EPI_GRU_ZR = 1, EPI_GRU_H = 2, EPI_BWD_DRHU = 3, EPI_BWD_DS = 4, EPI_BWD_DSU_DC = 5 };
struct DenseEpi {
    int mode = EPI_LINEAR;
    const float* bias = nullptr;
    const float* add = nullptr; int64_t ld_add = 0;
    float* out = nullptr; int64_t ld_out = 0;
    float* out2 = nullptr; int64_t ld_out2 = 0; int n2_start = 0;   // second copy of columns n >= n2_start
    // GRU epilogues
    int ST = 0;
    float* gates = nullptr; int64_t ld_gates = 0;          // z | r | h~
    const float* sprev = nullptr; int64_t ld_sprev = 0;
    float* rh_out = nullptr; int64_t ld_rh = 0;
    float* s_out = nullptr; int64_t ld_s = 0;
    float* s_out2 = nullptr; int64_t ld_s2 = 0;
    // backward epilogues (decoder GRU, GRU.lua:22-30 reversed)
    float* dA = nullptr; int64_t ld_dA = 0;                // daz | dar | dah of the step being written
    float* dsu = nullptr;                                  // [B, 2ST] d{s_{t-1}, u}
    const float* dsc = nullptr; int64_t ld_dsc = 0;        // d{s,c} from the MLP for the step being prepared
    const float* gates_n = nullptr; const float* su_n = nullptr;   // gates / {s,u} of the step being prepared (EPI_BWD_DS)
};

// Loads of activations go through L2 (ld.global.cg): in the fused chain kernel below they may have been written
// earlier in the same launch by another CTA, and L1 is not coherent across SMs.
__device__ __forceinline__ void dense_epilogue(const DenseEpi& e, int b, int n, float v) {
    if (e.mode == EPI_LINEAR) {
        if (e.bias) v += e.bias[n];
        if (e.add) v += __ldcg(e.add + (size_t)b * e.ld_add + n);
        e.out[(size_t)b * e.ld_out + n] = v;
        if (e.out2 && n >= e.n2_start) e.out2[(size_t)b * e.ld_out2 + n - e.n2_start] = v;
    } else if (e.mode == EPI_GRU_ZR) {
        const float g = sigmoid_acc(v);                                // GRU.lua:23-24
        e.gates[(size_t)b * e.ld_gates + n] = g;
        if (n >= e.ST) e.rh_out[(size_t)b * e.ld_rh + n - e.ST] = g * __ldcg(e.sprev + (size_t)b * e.ld_sprev + n - e.ST);   // GRU.lua:25
    } else if (e.mode == EPI_GRU_H) {
        const float hc = tanh_acc(v);                                  // GRU.lua:26
        const float z = __ldcg(e.gates + (size_t)b * e.ld_gates + n);
        const float sp = __ldcg(e.sprev + (size_t)b * e.ld_sprev + n);
        const float s = (1.f - z) * sp + z * hc;                       // GRU.lua:27-30
        e.gates[(size_t)b * e.ld_gates + 2 * e.ST + n] = hc;
        e.s_out[(size_t)b * e.ld_s + n] = s;
        if (e.s_out2) e.s_out2[(size_t)b * e.ld_s2 + n] = s;
    } else if (e.mode == EPI_BWD_DRHU) {
        // v = (dah . G_h)[n]: n < ST -> d(r*s_{t-1}); n >= ST -> the candidate gate's share of du
        if (n < e.ST) {
            const float r = __ldcg(e.gates + (size_t)b * e.ld_gates + e.ST + n), sp = __ldcg(e.sprev + (size_t)b * e.ld_sprev + n);
            e.dA[(size_t)b * e.ld_dA + e.ST + n] = v * sp * r * (1.f - r);     // dar
            e.dsu[(size_t)b * 2 * e.ST + n] = __ldcg(e.dsu + (size_t)b * 2 * e.ST + n) + v * r;
        } else {
            e.dsu[(size_t)b * 2 * e.ST + n] = v;
        }
    } else if (e.mode == EPI_BWD_DSU_DC) {
        // rows n < 2ST: d{s_{t-1}, u} += {daz, dar} . G_zr ; rows n >= 2ST: dc_t = dc_mlp + du . W_jc with du substituted (folded weights)
        if (n < 2 * e.ST) {
            v += __ldcg(e.dsu + (size_t)b * 2 * e.ST + n);
            e.dsu[(size_t)b * 2 * e.ST + n] = v;
            if (n >= e.ST) e.out2[(size_t)b * e.ld_out2 + n - e.ST] = v;
        } else {
            e.out[(size_t)b * e.ld_out + n - 2 * e.ST] = v + __ldcg(e.dsc + (size_t)b * e.ld_dsc + n - 2 * e.ST);
        }
    } else {   // EPI_BWD_DS: v = (dq . W_s)[n]; ds_{t-1} complete -> elementwise GRU backward of step t-1
        const float ds = v + __ldcg(e.dsu + (size_t)b * 2 * e.ST + n) + __ldcg(e.dsc + (size_t)b * e.ld_dsc + n);
        const float z = __ldcg(e.gates_n + (size_t)b * e.ld_gates + n), hc = __ldcg(e.gates_n + (size_t)b * e.ld_gates + 2 * e.ST + n);
        const float sp = __ldcg(e.su_n + (size_t)b * e.ld_sprev + n);
        e.dA[(size_t)b * e.ld_dA + 2 * e.ST + n] = ds * z * (1.f - hc * hc);        // dah
        e.dA[(size_t)b * e.ld_dA + n] = ds * (hc - sp) * z * (1.f - z);            // daz
        e.dsu[(size_t)b * 2 * e.ST + n] = ds * (1.f - z);
    }
}

// The epilogue's activation operands, fetched BEFORE the dot products so their L2 round trip overlaps the main loop
// (each launch of the decoder chain is a handful of dependent L2 latencies long; this removes one of them).
struct EpiOps { float a, b, c, d, e; };
__device__ __forceinline__ EpiOps dense_prefetch(const DenseEpi& e, int b, int n) {
    EpiOps o = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (e.mode == EPI_LINEAR) {
        if (e.add) o.a = __ldcg(e.add + (size_t)b * e.ld_add + n);
    } else if (e.mode == EPI_GRU_ZR) {
        if (n >= e.ST) o.a = __ldcg(e.sprev + (size_t)b * e.ld_sprev + n - e.ST);
    } else if (e.mode == EPI_GRU_H) {
        o.a = __ldcg(e.gates + (size_t)b * e.ld_gates + n);
        o.b = __ldcg(e.sprev + (size_t)b * e.ld_sprev + n);
    } else if (e.mode == EPI_BWD_DRHU) {
        if (n < e.ST) {
            o.a = __ldcg(e.gates + (size_t)b * e.ld_gates + e.ST + n);
            o.b = __ldcg(e.sprev + (size_t)b * e.ld_sprev + n);
            o.c = __ldcg(e.dsu + (size_t)b * 2 * e.ST + n);
        }
    } else if (e.mode == EPI_BWD_DSU_DC) {
        o.a = n < 2 * e.ST ? __ldcg(e.dsu + (size_t)b * 2 * e.ST + n) : __ldcg(e.dsc + (size_t)b * e.ld_dsc + n - 2 * e.ST);
    } else {
        o.a = __ldcg(e.dsu + (size_t)b * 2 * e.ST + n);
        o.b = __ldcg(e.dsc + (size_t)b * e.ld_dsc + n);
        o.c = __ldcg(e.gates_n + (size_t)b * e.ld_gates + n);
        o.d = __ldcg(e.gates_n + (size_t)b * e.ld_gates + 2 * e.ST + n);
        o.e = __ldcg(e.su_n + (size_t)b * e.ld_sprev + n);
    }
    return o;
}
__device__ __forceinline__ void dense_epilogue_pf(const DenseEpi& e, int b, int n, float v, const EpiOps& o) {
    if (e.mode == EPI_LINEAR) {
        if (e.bias) v += e.bias[n];
        v += o.a;
        e.out[(size_t)b * e.ld_out + n] = v;
        if (e.out2 && n >= e.n2_start) e.out2[(size_t)b * e.ld_out2 + n - e.n2_start] = v;
    } else if (e.mode == EPI_GRU_ZR) {
        const float g = sigmoid_acc(v);                                // GRU.lua:23-24
        e.gates[(size_t)b * e.ld_gates + n] = g;
        if (n >= e.ST) e.rh_out[(size_t)b * e.ld_rh + n - e.ST] = g * o.a;   // GRU.lua:25
    } else if (e.mode == EPI_GRU_H) {
        const float hc = tanh_acc(v);                                  // GRU.lua:26
        const float s = (1.f - o.a) * o.b + o.a * hc;                  // GRU.lua:27-30
        e.gates[(size_t)b * e.ld_gates + 2 * e.ST + n] = hc;
        e.s_out[(size_t)b * e.ld_s + n] = s;
        if (e.s_out2) e.s_out2[(size_t)b * e.ld_s2 + n] = s;
    } else if (e.mode == EPI_BWD_DRHU) {
        if (n < e.ST) {
            e.dA[(size_t)b * e.ld_dA + e.ST + n] = v * o.b * o.a * (1.f - o.a);     // dar
            e.dsu[(size_t)b * 2 * e.ST + n] = o.c + v * o.a;
        } else {
            e.dsu[(size_t)b * 2 * e.ST + n] = v;
        }
    } else if (e.mode == EPI_BWD_DSU_DC) {
        v += o.a;
        if (n < 2 * e.ST) {
            e.dsu[(size_t)b * 2 * e.ST + n] = v;
            if (n >= e.ST) e.out2[(size_t)b * e.ld_out2 + n - e.ST] = v;
        } else {
            e.out[(size_t)b * e.ld_out + n - 2 * e.ST] = v;
        }
    } else {
        const float ds = v + o.a + o.b;
        const float z = o.c, hc = o.d, sp = o.e;
        e.dA[(size_t)b * e.ld_dA + 2 * e.ST + n] = ds * z * (1.f - hc * hc);        // dah
        e.dA[(size_t)b * e.ld_dA + n] = ds * (hc - sp) * z * (1.f - z);            // daz
        e.dsu[(size_t)b * 2 * e.ST + n] = ds * (1.f - z);
    }
}

template <int KS4>
__global__ void __launch_bounds__(256)
dense_small_kernel(const float* __restrict__ X, int64_t ldx, int B, int K, const float* __restrict__ W, int ldw, int N, int NT, const DenseEpi e) {
    constexpr int KP4 = 8 * KS4;                     // padded K in float4 units
    extern __shared__ __align__(16) float sm[];
    float4* ws4 = reinterpret_cast<float4*>(sm);     // [NT][KP4]
    float* part = sm + (size_t)NT * KP4 * 4;         // [8][NT][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * NT;
    pdl_trigger();                                   // let the next kernel of the chain start its own prologue
    // prologue independent of the previous kernel: stage this CTA's weight rows (parameters / per-call copies)
    for (int idx = tid; idx < NT * KP4; idx += 256) {
        const int c = idx / KP4, k4 = idx - c * KP4, n = n0 + c;
        ws4[idx] = (n < N && k4 * 4 < K) ? ldg4_any(W + (size_t)n * ldw + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();                                      // X (and everything the epilogue touches) is produced upstream
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        float4 x[KS4];
#pragma unroll
        for (int i = 0; i < KS4; i++) {
            const int k = (warp * KS4 + i) * 4;
            x[i] = (b < B && k < K) ? __ldcg(reinterpret_cast<const float4*>(X + (size_t)b * ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // this thread's epilogue output (NT * 32 <= 256: at most one per thread) and its operands
        const int oc = tid >> 5, on = n0 + oc, obr = b0 + lane;
        const bool ovalid = oc < NT && on < N && obr < B;
        EpiOps ops = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (ovalid) ops = dense_prefetch(e, obr, on);
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < NT; c++) {
            const float4* wr = ws4 + c * KP4 + warp * KS4;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < KS4; i++) {
                const float4 w4 = wr[i];
                if (i & 1) { a1 = fmaf(w4.x, x[i].x, a1); a1 = fmaf(w4.y, x[i].y, a1); a1 = fmaf(w4.z, x[i].z, a1); a1 = fmaf(w4.w, x[i].w, a1); }
                else { a0 = fmaf(w4.x, x[i].x, a0); a0 = fmaf(w4.y, x[i].y, a0); a0 = fmaf(w4.z, x[i].z, a0); a0 = fmaf(w4.w, x[i].w, a0); }
            }
            part[(warp * NT + c) * 32 + lane] = a0 + a1;
        }
        __syncthreads();
        if (ovalid) {
            float v = 0.f;
#pragma unroll
            for (int wg = 0; wg < 8; wg++) v += part[(wg * NT + oc) * 32 + lane];
            dense_epilogue_pf(e, obr, on, v, ops);
        }
    }
}

static int dense_small(s2s_ctx* ctx, const float* X, int64_t ldx, int B, int K, const float* W, int ldw, int N, const DenseEpi& e) {
    S2S_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "dense_small: K (%d) and ldx (%ld) must be multiples of 4 and X 16-byte aligned", K, (long)ldx);
    S2S_REQUIRE(K <= 1024, "dense_small: K=%d > 1024 not supported", K);
    const int ks4 = ceil_div(K, 32);
    const int NT = N >= 512 ? 8 : 4;
    dim3 grid(ceil_div(N, NT));
    prof_begin(ctx, S2S_PROF_DENSE_SMALL);
#define DS_LAUNCH(KS)                                                                                           \
    do {                                                                                                        \
        const size_t smem = ((size_t)NT * 8 * KS * 4 + (size_t)8 * NT * 32) * 4;                                \
        static size_t attr = 0;                                                                                 \
        if (smem > 48 * 1024 && smem > attr) {                                                                  \
            S2S_CUDA(cudaFuncSetAttribute(dense_small_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr = smem;                                                                                        \
        }                                                                                                       \
        S2S_CUDA(launch_kernel(dense_small_kernel<KS>, grid, dim3(256), smem, ctx->stream, ctx->pdl, X, ldx, B, K, W, ldw, N, NT, e)); \
    } while (0)
    if (ks4 <= 1) DS_LAUNCH(1);
    else if (ks4 <= 2) DS_LAUNCH(2);
    else if (ks4 <= 4) DS_LAUNCH(4);
    else if (ks4 <= 8) DS_LAUNCH(8);
    else if (ks4 <= 12) DS_LAUNCH(12);
    else if (ks4 <= 16) DS_LAUNCH(16);
    else if (ks4 <= 24) DS_LAUNCH(24);
    else DS_LAUNCH(32);
#undef DS_LAUNCH
    prof_end(ctx, S2S_PROF_DENSE_SMALL, 4.0 * ((double)N * K + (double)B * K + (double)B * N));
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// =================================================================================================
// dense_chain: up to four DEPENDENT dense_small products in ONE launch.
// Between two attention steps the decoder runs a chain of tiny products on the B minibatch rows
//   forward :  u_t -> {z, r} -> h~ / s_t -> q_{t+1}
//   backward:  [ds_{t} / GRU elementwise] -> d{r*s, u} -> d{s, u} -> dc_t
// each of which needs the complete result of the previous one.  As separate launches each link costs a
// launch + drain (~5.5 us for ~0.5 us of work).  Here a grid of CH_G co-resident CTAs (cooperative launch)
// runs all links back to back: every CTA owns N/CH_G output columns of every link, stages those weight rows
// for ALL links in shared memory up front, and the links are separated by a grid-wide barrier (one atomic
// arrive + a generation word to spin on; ~1 us) instead of a kernel boundary.
// =================================================================================================
constexpr int CH_G = 128;          // CTAs (<= SM count: all co-resident)
constexpr int CH_KS4 = 16;         // K <= 512
constexpr int CH_NTMAX = 8;
constexpr int CH_MAXPH = 4;
struct ChainPhase {
    const float* X; int64_t ldx; int K; const float* W; int ldw; int N; int NT; int ws_off4;   // ws_off4: float4 offset of this phase's rows
    DenseEpi e;
};
struct ChainParams {
    int nph, B;
    unsigned* bar;                 // {arrival count, generation}
    ChainPhase ph[CH_MAXPH];
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the (cooperative) grid; gen is the caller's copy of the generation word, read before its first arrival
__device__ __forceinline__ void chain_grid_barrier(unsigned* bar, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned old = atomicAdd(&bar[0], 1u);
        if (old == gridDim.x - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(&bar[1], 1u);
        } else {
            const long long t0 = clock64();
            while (ld_acquire_u32(&bar[1]) == gen) {
                if (clock64() - t0 > 2000000000ll) __trap();     // a lost CTA must not hang the GPU
            }
        }
        gen++;
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
dense_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ __align__(16) float sm[];
    float4* ws4 = reinterpret_cast<float4*>(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ws_total4 = 0;
#pragma unroll
    for (int ip = 0; ip < CH_MAXPH; ip++) {
        if (ip < p.nph) {
            const ChainPhase& ph = p.ph[ip];
            const int KP4 = 8 * ((ph.K + 31) / 32);
            const int n0 = blockIdx.x * ph.NT;
            for (int idx = tid; idx < ph.NT * KP4; idx += 256) {
                const int c = idx / KP4, k4 = idx - c * KP4, n = n0 + c;
                ws4[ph.ws_off4 + idx] = (n < ph.N && k4 * 4 < ph.K) ? ldg4_any(ph.W + (size_t)n * ph.ldw + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ws_total4 = ph.ws_off4 + ph.NT * KP4;
        }
    }
    float* part = sm + (size_t)ws_total4 * 4;      // [8][CH_NTMAX][32]
    unsigned gen = 0;
    if (tid == 0) gen = ld_acquire_u32(&p.bar[1]);
    __syncthreads();
#pragma unroll
    for (int ip = 0; ip < CH_MAXPH; ip++) {
        if (ip >= p.nph) break;
        if (ip > 0) chain_grid_barrier(p.bar, gen);
        const ChainPhase& ph = p.ph[ip];
        const int NT = ph.NT, K = ph.K, ks4 = (K + 31) / 32, KP4 = 8 * ks4;
        const int n0 = blockIdx.x * NT;
        if (n0 >= ph.N) continue;                 // no columns of this link for this CTA (it still joins the barriers)
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int b = b0 + lane;
            float4 x[CH_KS4];
#pragma unroll
            for (int i = 0; i < CH_KS4; i++) {
                const int k = (warp * ks4 + i) * 4;
                x[i] = (i < ks4 && b < p.B && k < K) ? __ldcg(reinterpret_cast<const float4*>(ph.X + (size_t)b * ph.ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            for (int c = 0; c < NT; c++) {
                const float4* wr = ws4 + ph.ws_off4 + c * KP4 + warp * ks4;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int i = 0; i < CH_KS4; i++) {
                    if (i < ks4) {
                        const float4 w4 = wr[i];
                        if (i & 1) { a1 = fmaf(w4.x, x[i].x, a1); a1 = fmaf(w4.y, x[i].y, a1); a1 = fmaf(w4.z, x[i].z, a1); a1 = fmaf(w4.w, x[i].w, a1); }
                        else { a0 = fmaf(w4.x, x[i].x, a0); a0 = fmaf(w4.y, x[i].y, a0); a0 = fmaf(w4.z, x[i].z, a0); a0 = fmaf(w4.w, x[i].w, a0); }
                    }
                }
                part[(warp * CH_NTMAX + c) * 32 + lane] = a0 + a1;
            }
            __syncthreads();
            for (int o = tid; o < NT * 32; o += 256) {
                const int c = o >> 5, bb = o & 31, n = n0 + c, br = b0 + bb;
                float v = 0.f;
#pragma unroll
                for (int wg = 0; wg < 8; wg++) v += part[(wg * CH_NTMAX + c) * 32 + bb];
                if (n < ph.N && br < p.B) dense_epilogue(ph.e, br, n, v);
            }
        }
    }
}

struct ChainLink { const float* X; int64_t ldx; int K; const float* W; int ldw; int N; DenseEpi e; };

static bool dense_chain_on() {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("S2S_CHAIN"); enabled = e ? atoi(e) : 0; }
    return enabled != 0;
}
static bool dense_chain_supported(const ChainLink* l, int n) {
    static int enabled = -1;
    // Measured on B200 (cfg2, CUDA-graph replay): one chain launch = 29.5 us against 4 x 5.5 us for the separate
    // launches -- a kernel boundary inside a graph costs about as much as the grid barrier (~2 us), and with 128 CTAs
    // each link re-reads the whole X from L2.  Kept as an opt-in (S2S_CHAIN=1) for experiments; default off.
    if (enabled < 0) { const char* e = getenv("S2S_CHAIN"); enabled = e ? atoi(e) : 0; }
    if (!enabled || n < 1 || n > CH_MAXPH) return false;
    for (int i = 0; i < n; i++) {
        if (l[i].K % 4 || l[i].K > 32 * CH_KS4 || l[i].ldx % 4 || (reinterpret_cast<uintptr_t>(l[i].X) & 15)) return false;
        if (ceil_div(l[i].N, CH_G) > CH_NTMAX) return false;
    }
    return true;
}
// the links must be supported (dense_chain_supported); same per-link summation order as dense_small
static int dense_chain(s2s_ctx* ctx, const ChainLink* l, int n, int B) {
    ChainParams p = {};
    p.nph = n; p.B = B; p.bar = ctx->counters + 4000;
    int off4 = 0;
    double work = 0;
    for (int i = 0; i < n; i++) {
        ChainPhase& ph = p.ph[i];
        ph.X = l[i].X; ph.ldx = l[i].ldx; ph.K = l[i].K; ph.W = l[i].W; ph.ldw = l[i].ldw; ph.N = l[i].N; ph.e = l[i].e;
        ph.NT = ceil_div(l[i].N, CH_G); ph.ws_off4 = off4;
        off4 += ph.NT * 8 * ceil_div(l[i].K, 32);
        work += 4.0 * ((double)l[i].N * l[i].K + (double)B * l[i].K + (double)B * l[i].N);
    }
    const size_t smem = (size_t)off4 * 16 + (size_t)8 * CH_NTMAX * 32 * 4;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        S2S_CUDA(cudaFuncSetAttribute(dense_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CH_G); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;     // co-residency of all CTAs is required by the barrier
    static int coop = -1;
    if (coop < 0) { const char* e = getenv("S2S_CHAIN_COOP"); coop = e ? atoi(e) : 1; }
    cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
    prof_begin(ctx, S2S_PROF_DENSE_SMALL);
    S2S_CUDA(cudaLaunchKernelEx(&cfg, dense_chain_kernel, p));
    prof_end(ctx, S2S_PROF_DENSE_SMALL, work);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// =================================================================================================
// elementwise / gather kernels
// =================================================================================================
// y_in[b,t,:] = b_y + W_y[:, y_{t-1}]   (Linear(V,ST) on the one-hot previous label, Attention.lua:149;
// zeros_y at t = 0, RNNAttention.lua:172-176)
__global__ void yin_gather_kernel(const float* __restrict__ Wy, const float* __restrict__ by, const int* __restrict__ labels,
                                  int B, int T, int ST, int V, float* __restrict__ yin) {
    const int bt = blockIdx.x, t = bt % T, b = bt / T;
    int y = t > 0 ? labels[(size_t)b * T + t - 1] : -1;
    if (y >= V) y = -1;
    for (int i = threadIdx.x; i < ST; i += blockDim.x)
        yin[(size_t)bt * ST + i] = by[i] + (y >= 0 ? Wy[(size_t)i * V + y] : 0.f);
}
// dWy[:, y_{t-1}] += dyin[b,t,:]
__global__ void wy_scatter_kernel(const float* __restrict__ dyin, const int* __restrict__ labels, int B, int T, int ST, int V, float* __restrict__ dWy) {
    const int bt = blockIdx.x, t = bt % T, b = bt / T;
    if (t == 0) return;
    const int y = labels[(size_t)b * T + t - 1];
    if (y < 0 || y >= V) return;
    for (int i = threadIdx.x; i < ST; i += blockDim.x) atomicAdd(dWy + (size_t)i * V + y, dyin[(size_t)bt * ST + i]);
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ m, float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * m[i];
}
// Maxout: groups of MW consecutive units (Maxout.lua:15-19); first maximum wins on ties
__global__ void maxout_fwd_kernel(const float* __restrict__ pre, int64_t rows, int M, int MW, float* __restrict__ mo, int* __restrict__ midx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * M) return;
    const float* p = pre + i * MW;
    int best = 0; float bv = p[0];
    for (int j = 1; j < MW; j++) if (p[j] > bv) { bv = p[j]; best = j; }
    mo[i] = bv; midx[i] = best;
}
__global__ void maxout_bwd_kernel(const float* __restrict__ dmo, const int* __restrict__ midx, int64_t rows, int M, int MW, float* __restrict__ dm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * M) return;
    const int best = midx[i]; const float g = dmo[i];
    for (int j = 0; j < MW; j++) dm[i * MW + j] = j == best ? g : 0.f;
}
// LogSoftMax over V (warp per row), in place + copy
__global__ void logsoftmax_kernel(float* __restrict__ x, int64_t rows, int V, float* __restrict__ copy) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = x + row * V;
    float m = -INFINITY;
    for (int i = lane; i < V; i += 32) m = fmaxf(m, r[i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < V; i += 32) s += expf(r[i] - m);
    s = warp_sum(s);
    const float lz = m + logf(s);
    for (int i = lane; i < V; i += 32) { const float v = r[i] - lz; r[i] = v; if (copy) copy[row * V + i] = v; }
}
// dlogits = dlogp - exp(logp) * sum(dlogp)
// (rows t >= T_b are padding: their gradient is forced to zero)
__global__ void logsoftmax_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ dlogp, int64_t rows, int V,
                                      const int* __restrict__ tlens, int T, float* __restrict__ dlogits) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (tlens && (int)(row % T) >= tlens[row / T]) {
        for (int i = lane; i < V; i += 32) dlogits[row * V + i] = 0.f;
        return;
    }
    float s = 0.f;
    for (int i = lane; i < V; i += 32) s += dlogp[row * V + i];
    s = warp_sum(s);
    for (int i = lane; i < V; i += 32) dlogits[row * V + i] = dlogp[row * V + i] - expf(logp[row * V + i]) * s;
}
// decoder GRU backward, elementwise part 1 (GRU.lua:27-30 reversed)
//   ds = ds_mlp + ds_carry ; dah = ds z (1-hc^2) ; daz = ds (hc - sp) z (1-z) ; dsu[:, :ST] = ds (1-z)
__global__ void gru_bwd_e1_kernel(const float* __restrict__ dsc, int64_t ld_dsc, const float* __restrict__ ds_carry,
                                  const float* __restrict__ gates, int64_t ld_g, const float* __restrict__ su, int64_t ld_su,
                                  int B, int ST, float* __restrict__ dA, int64_t ld_dA, float* __restrict__ dsu) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * ST) return;
    const int b = idx / ST, j = idx - b * ST;
    const float ds = dsc[(size_t)b * ld_dsc + j] + ds_carry[idx];
    const float z = gates[(size_t)b * ld_g + j], hc = gates[(size_t)b * ld_g + 2 * ST + j], sp = su[(size_t)b * ld_su + j];
    dA[(size_t)b * ld_dA + 2 * ST + j] = ds * z * (1.f - hc * hc);
    dA[(size_t)b * ld_dA + j] = ds * (hc - sp) * z * (1.f - z);
    dsu[(size_t)b * 2 * ST + j] = ds * (1.f - z);
}
// part 2: drh = drhu[:, :ST] ; dar = drh sp r (1-r) ; dsu[:, :ST] += drh r ; dsu[:, ST:] = drhu[:, ST:]
__global__ void gru_bwd_e2_kernel(const float* __restrict__ drhu, const float* __restrict__ gates, int64_t ld_g,
                                  const float* __restrict__ su, int64_t ld_su, int B, int ST,
                                  float* __restrict__ dA, int64_t ld_dA, float* __restrict__ dsu) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * ST) return;
    const int b = idx / ST, j = idx - b * ST;
    const float drh = drhu[(size_t)b * 2 * ST + j];
    const float r = gates[(size_t)b * ld_g + ST + j], sp = su[(size_t)b * ld_su + j];
    dA[(size_t)b * ld_dA + ST + j] = drh * sp * r * (1.f - r);
    dsu[(size_t)b * 2 * ST + j] += drh * r;
    dsu[(size_t)b * 2 * ST + ST + j] = drhu[(size_t)b * 2 * ST + ST + j];
}
// UW[j][i] = sum_m U[i,m] WF[m,j] ; qbias[i] = bs[i] + sum_m U[i,m] bF[m]
__global__ void loc_fold_kernel(const float* __restrict__ U, const float* __restrict__ WF, const float* __restrict__ bF,
                                const float* __restrict__ bs, int S, int K, int KF, float* __restrict__ uw, float* __restrict__ qbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    float ub = 0.f;
    for (int m = 0; m < K; m++) ub = fmaf(U[(size_t)i * K + m], bF[m], ub);
    qbias[i] = bs[i] + ub;
    for (int j = 0; j < KF; j++) {
        float a = 0.f;
        for (int m = 0; m < K; m++) a = fmaf(U[(size_t)i * K + m], WF[(size_t)m * KF + j], a);
        uw[(size_t)j * S + i] = a;
    }
}
// gradients of the folded location parameters back to U, WF, bF:
//   dU[i,m] += sum_j dUW[j][i] WF[m,j] + dUb[i] bF[m] ; dWF[m,j] += sum_i U[i,m] dUW[j][i] ; dbF[m] += sum_i U[i,m] dUb[i]
// one block per feature map m: the sums over the S score units are block reductions (no atomics: 17 values per block)
constexpr int LOC_MAXKF = 16;        // filter taps (attention.cu)
__global__ void __launch_bounds__(256)
loc_unfold_kernel(const float* __restrict__ U, const float* __restrict__ WF, const float* __restrict__ bF,
                  const float* __restrict__ duw, const float* __restrict__ dub, int S, int K, int KF,
                  float* __restrict__ dU, float* __restrict__ dWF, float* __restrict__ dbF) {
    __shared__ float red[8][LOC_MAXKF + 1];
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc[LOC_MAXKF + 1];
#pragma unroll
    for (int j = 0; j <= LOC_MAXKF; j++) acc[j] = 0.f;
    const float bm = bF[m];
    for (int i = tid; i < S; i += 256) {
        const float db = dub[i], u = U[(size_t)i * K + m];
        float a = db * bm;
#pragma unroll
        for (int j = 0; j < LOC_MAXKF; j++) {
            if (j < KF) {
                const float g = duw[(size_t)j * S + i];
                a = fmaf(g, WF[(size_t)m * KF + j], a);
                acc[j] = fmaf(u, g, acc[j]);
            }
        }
        acc[LOC_MAXKF] = fmaf(u, db, acc[LOC_MAXKF]);
        dU[(size_t)i * K + m] += a;
    }
#pragma unroll
    for (int j = 0; j <= LOC_MAXKF; j++) {
        float v = acc[j];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (tid <= LOC_MAXKF) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) v += red[w][tid];
        if (tid < KF) dWF[(size_t)m * KF + tid] += v;
        else if (tid == LOC_MAXKF) dbF[m] += v;
    }
}

static void pad_lr(int kf, int* pl) { *pl = (kf % 2 == 1) ? (kf - 1) / 2 : kf / 2; }   // Attention.lua:77-85

// =================================================================================================
// forward
// =================================================================================================
// The part of the decoder forward that does not depend on the annotations h: the folded location weights U W_F and q bias, the folded
// context Linears W_j[:, :ST] W_c, and the teacher-forced label path y_in / its share of u for all steps.  model_forward issues it on the
// side stream at the START of the step, so it runs under the encoder instead of between the encoder and the decoder time loop;
// decoder_forward runs it inline when nobody did.  Buffers: persistent arena (read again by the backward pass) + per-call scratch (uy).
int decoder_prepare(s2s_ctx* ctx, const Layout& Y, const float* P, const int* labels, int B, int T, bool backward_follows) {
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V, KF = Y.K > 0 ? Y.KF : 0;
    if (!ctx->dec) ctx->dec = new DecoderState();
    DecoderState& d = *ctx->dec;
    d.valid = false; d.prepared = false; d.prep_pending = false; d.tw_valid = false;
    Arena& pa = ctx->persist;
    const size_t BT = (size_t)B * T;
    cudaStream_t st = ctx->stream;
    float* bjc;
    S2S_ALLOC(d.qbias, pa, float, S);
    S2S_ALLOC(d.Wjc, pa, float, (size_t)ST * A);
    if (KF > 0) S2S_ALLOC(d.uw, pa, float, (size_t)KF * S); else d.uw = nullptr;
    S2S_ALLOC(d.yin, pa, float, BT * ST);
    S2S_ALLOC(d.uy, ctx->arena, float, BT * ST);
    S2S_ALLOC(bjc, ctx->arena, float, ST);
    // location fold / q bias
    if (KF > 0) {
        loc_fold_kernel<<<ceil_div(S, 128), 128, 0, st>>>(P + Y.U.off, P + Y.WF.off, P + Y.bF.off, P + Y.bs.off, S, Y.K, KF, d.uw, d.qbias);
        S2S_LAUNCH_CHECK(ctx);
    } else {
        S2S_CUDA(cudaMemcpyAsync(d.qbias, P + Y.bs.off, S * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // The two input Linears on the context (Attention.lua:150-151) have no nonlinearity between them:
    //   u = W_j {W_c c + b_c, y_in} + b_j = (W_j[:, :ST] W_c) c + W_j[:, ST:] y_in + (W_j[:, :ST] b_c + b_j)
    // so the time loop needs ONE product on c_t; c_in itself (needed for dW_j) is recomputed time-batched.
    S2S_TRY(gemm_f32(ctx, false, false, ST, A, ST, 1.f, P + Y.Wj.off, 2 * ST, P + Y.Wc.off, A, 0.f, d.Wjc, A, nullptr, GemmBatch(), 1, 1));
    S2S_TRY(gemm_f32(ctx, false, true, 1, ST, ST, 1.f, P + Y.bc.off, ST, P + Y.Wj.off, 2 * ST, 0.f, bjc, ST, P + Y.bj.off, GemmBatch(), 1, 1));
    // teacher-forced input path, hoisted out of the loop: y_in (Attention.lua:149) and its share of u
    yin_gather_kernel<<<(unsigned)BT, 128, 0, st>>>(P + Y.Wy.off, P + Y.by.off, labels, B, T, ST, V, d.yin);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(gemm_f32(ctx, false, true, (int)BT, ST, ST, 1.f, d.yin, ST, P + Y.Wj.off + ST, 2 * ST, 0.f, d.uy, ST, bjc));
    if (backward_follows) {      // the backward loop's transposed weights (K-contiguous rows): off the critical path of the backward pass
        S2S_ALLOC(d.GhT, pa, float, (size_t)2 * ST * ST);
        S2S_ALLOC(d.GzrT, pa, float, (size_t)2 * ST * 2 * ST);
        S2S_ALLOC(d.WjcT, pa, float, (size_t)A * ST);
        S2S_ALLOC(d.WsT, pa, float, (size_t)ST * S);
        S2S_TRY(transpose_f32(ctx, P + Y.Gh.off, ST, 2 * ST, 2 * ST, d.GhT, ST));
        S2S_TRY(transpose_f32(ctx, P + Y.Gz.off, 2 * ST, 2 * ST, 2 * ST, d.GzrT, 2 * ST));
        S2S_TRY(transpose_f32(ctx, d.Wjc, ST, A, A, d.WjcT, ST));
        S2S_TRY(transpose_f32(ctx, P + Y.Ws.off, S, ST, ST, d.WsT, S));
        d.tw_valid = true;
    }
    d.prepared = true; d.prep_B = B; d.prep_T = T; d.prep_n = Y.n; d.prep_epoch = ctx->persist.epoch; d.prep_epoch_scratch = ctx->arena.epoch;
    return 0;
}

int decoder_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax,
                    const int* labels, const int* tlens, int T, const float* dropmask, float lambda, float* logp_out, bool prefetch_v1) {
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V, M = Y.M, MW = Y.MW, KF = Y.K > 0 ? Y.KF : 0;
    S2S_REQUIRE(B > 0 && Lmax > 0 && T > 0, "decoder_forward: empty batch (B=%d Lmax=%d T=%d)", B, Lmax, T);
    S2S_REQUIRE(ST % 4 == 0 && A % 4 == 0, "decoder: ST and A must be multiples of 4");
    if (!(ctx->dec && ctx->dec->prepared && ctx->dec->prep_B == B && ctx->dec->prep_T == T && ctx->dec->prep_n == Y.n &&
          ctx->dec->prep_epoch == ctx->persist.epoch && ctx->dec->prep_epoch_scratch == ctx->arena.epoch))
        S2S_TRY(decoder_prepare(ctx, Y, P, labels, B, T));           // nobody ran the h-independent part ahead of time
    DecoderState& d = *ctx->dec;
    d.prepared = false;
    d.valid = false; d.B = B; d.Lmax = Lmax; d.T = T; d.Y = Y; d.lambda = lambda; d.has_drop = dropmask != nullptr;
    d.V1 = nullptr; d.v1_pending = false;
    Arena& pa = ctx->persist;
    const size_t BT = (size_t)B * T;
    S2S_ALLOC(d.Vh, pa, float, (size_t)B * Lmax * S);
    S2S_ALLOC(d.alpha, pa, float, BT * Lmax);
    S2S_ALLOC(d.sc, pa, float, BT * (ST + A));
    S2S_ALLOC(d.q, pa, float, BT * S);
    S2S_ALLOC(d.pen, pa, float, BT);
    S2S_ALLOC(d.cin, pa, float, BT * ST);
    S2S_ALLOC(d.su, pa, float, BT * 2 * ST);
    S2S_ALLOC(d.rhu, pa, float, BT * 2 * ST);
    S2S_ALLOC(d.gates, pa, float, BT * 3 * ST);
    S2S_ALLOC(d.mo, pa, float, BT * M);
    S2S_ALLOC(d.midx, pa, int, BT * M);
    if (Y.MLP == 2) { S2S_ALLOC(d.l1, pa, float, BT * M); S2S_ALLOC(d.mo2, pa, float, BT * M); S2S_ALLOC(d.midx2, pa, int, BT * M); }
    S2S_ALLOC(d.logp, pa, float, BT * V);
    if (dropmask) S2S_ALLOC(d.scm, pa, float, BT * (ST + A)); else d.scm = d.sc;
    S2S_TRY(attn_scratch_alloc(ctx, pa, B, Lmax, S, A, KF, false, &d.att));
    Arena& ar = ctx->arena;
    float *mpre, *zeros;
    float* const uy = d.uy;
    S2S_ALLOC(mpre, ar, float, BT * M * MW);
    S2S_ALLOC(zeros, ar, float, (size_t)B * ST);
    cudaStream_t st = ctx->stream;

    // Vh = TemporalConvolutionZeroBias(A,S,1)(h)   (Attention.lua:44; bias pinned to zero)
    S2S_TRY(gemm_f32(ctx, false, true, B * Lmax, S, A, 1.f, h, A, P + Y.WV.off, A, 0.f, d.Vh, S));
    int padl = 0;
    if (KF > 0) pad_lr(KF, &padl);
    if (d.prep_pending) {        // the h-independent part ran on the side stream (model_forward): join
        S2S_CUDA(cudaStreamWaitEvent(st, ctx->ev[4], 0));
        d.prep_pending = false;
    }
    S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * ST * sizeof(float), st));
    S2S_CUDA(cudaMemsetAsync(d.su, 0, BT * 2 * ST * sizeof(float), st));     // s_0 = 0 (Recurrent.lua:112)

    const int64_t ldsc = (int64_t)T * (ST + A), ldsu = (int64_t)T * 2 * ST, ldg = (int64_t)T * 3 * ST;
    bool clustered = false;     // the whole time loop in one persistent cluster kernel (decoder_cluster.cu) when the shapes allow
    S2S_TRY(decoder_cluster_forward(ctx, Y, P, h, lengths, B, Lmax, tlens, T, lambda, uy, d, &clustered));
    if (clustered && prefetch_v1 && KF > 0 && decoder_cluster_backward_eligible(Y, Lmax, lambda) && ctx->side[1] && st != ctx->side[1] &&
        !ctx->wgrad_join_pending) {
        // A backward pass follows (s2s_model_fwdbwd): its location path needs V1 = d e_t / d alpha_{t-1}, which depends on forward results
        // only.  Formed here on the low-priority side stream it runs beside the MLP, the loss and the MLP backward instead of in front of
        // the backward time loop.
        static int overlap = -1;
        if (overlap < 0) { const char* e = getenv("S2S_OVERLAP"); overlap = e ? atoi(e) : 1; }
        if (overlap) {
            S2S_ALLOC(d.V1, pa, float, BT * Lmax * KF);
            S2S_CUDA(cudaEventRecord(ctx->ev[2], st));
            S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev[2], 0));
            struct StreamGuard { s2s_ctx* c; cudaStream_t s; ~StreamGuard() { c->stream = s; } } g{ctx, st};
            ctx->stream = ctx->side[1];
            S2S_CUDA(cudaMemsetAsync(d.V1, 0, BT * Lmax * KF * sizeof(float), ctx->stream));
            S2S_TRY(attn_v1(ctx, d.Vh, d.q, P + Y.we.off, d.uw, d.alpha, lengths, tlens, B, Lmax, T, S, KF, padl, d.V1));
            S2S_CUDA(cudaEventRecord(ctx->ev[5], ctx->side[1]));
            d.v1_pending = true;
        }
    }
    if (!clustered) {   // q_0 = W_s s_0 + b_s with s_0 = 0   (Attention.lua:65-67, Recurrent.lua:112)
        DenseEpi e; e.bias = d.qbias; e.out = d.q; e.ld_out = (int64_t)T * S;
        S2S_TRY(dense_small(ctx, zeros, ST, B, ST, P + Y.Ws.off, ST, S, e));
    }
    for (int t = 0; t < T && !clustered; t++) {
        {   // attention step (Attention.lua:95-135)
            AttnLoc loc; loc.KF = KF; loc.padl = padl; loc.uw = d.uw;
            loc.alpha_prev = t ? d.alpha + (size_t)(t - 1) * Lmax : nullptr; loc.ld_aprev = (int64_t)T * Lmax;
            S2S_TRY(attn_step_fwd(ctx, d.att, d.Vh, h, d.q + (size_t)t * S, (int64_t)T * S, P + Y.we.off, lengths, B, Lmax, S, A, loc,
                                  d.alpha + (size_t)t * Lmax, (int64_t)T * Lmax, d.sc + (size_t)t * (ST + A) + ST, ldsc,
                                  d.pen + t, T, lambda, t ? d.alpha + (size_t)(t - 1) * Lmax : nullptr, (int64_t)T * Lmax, tlens, t));
        }
        // the dependent chain up to the next attention step: u_t -> {z, r} -> s_t -> q_{t+1}
        ChainLink L[4];
        int nl = 0;
        {   // u_t = (W_j[:, :ST] W_c) c_t + uy_t   (Attention.lua:150-151) -> second halves of {s,u} and {r*s,u}
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.add = uy + (size_t)t * ST; e.ld_add = (int64_t)T * ST;
            e.out = d.su + (size_t)t * 2 * ST + ST; e.ld_out = ldsu;
            e.out2 = d.rhu + (size_t)t * 2 * ST + ST; e.ld_out2 = ldsu; e.n2_start = 0;
            k.X = d.sc + (size_t)t * (ST + A) + ST; k.ldx = ldsc; k.K = A; k.W = d.Wjc; k.ldw = A; k.N = ST;
        }
        {   // z, r = sigmoid(G_{z,r} {s_{t-1}, u})   (GRU.lua:22-24) ; r * s_{t-1}  (:25)
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.mode = EPI_GRU_ZR; e.ST = ST; e.gates = d.gates + (size_t)t * 3 * ST; e.ld_gates = ldg;
            e.sprev = d.su + (size_t)t * 2 * ST; e.ld_sprev = ldsu; e.rh_out = d.rhu + (size_t)t * 2 * ST; e.ld_rh = ldsu;
            k.X = d.su + (size_t)t * 2 * ST; k.ldx = ldsu; k.K = 2 * ST; k.W = P + Y.Gz.off; k.ldw = 2 * ST; k.N = 2 * ST;
        }
        {   // h~ = tanh(G_h {r*s_{t-1}, u}) ; s_t = (1-z) s_{t-1} + z h~   (GRU.lua:26-30)
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.mode = EPI_GRU_H; e.ST = ST; e.gates = d.gates + (size_t)t * 3 * ST; e.ld_gates = ldg;
            e.sprev = d.su + (size_t)t * 2 * ST; e.ld_sprev = ldsu;
            e.s_out = d.sc + (size_t)t * (ST + A); e.ld_s = ldsc;
            e.s_out2 = t + 1 < T ? d.su + (size_t)(t + 1) * 2 * ST : nullptr; e.ld_s2 = ldsu;
            k.X = d.rhu + (size_t)t * 2 * ST; k.ldx = ldsu; k.K = 2 * ST; k.W = P + Y.Gh.off; k.ldw = 2 * ST; k.N = ST;
        }
        if (t + 1 < T) {   // q_{t+1} = W_s s_t + b_s   (Attention.lua:65-67)
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.bias = d.qbias; e.out = d.q + (size_t)(t + 1) * S; e.ld_out = (int64_t)T * S;
            k.X = d.sc + (size_t)t * (ST + A); k.ldx = ldsc; k.K = ST; k.W = P + Y.Ws.off; k.ldw = ST; k.N = S;
        }
        if (dense_chain_supported(L, nl)) {
            S2S_TRY(dense_chain(ctx, L, nl, B));
        } else {
            for (int i = 0; i < nl; i++) S2S_TRY(dense_small(ctx, L[i].X, L[i].ldx, B, L[i].K, L[i].W, L[i].ldw, L[i].N, L[i].e));
        }
    }

    // c_in = W_c c + b_c for all steps (Attention.lua:150): only the backward needs it
    S2S_TRY(gemm_f32(ctx, false, true, (int)BT, ST, A, 1.f, d.sc + ST, ST + A, P + Y.Wc.off, A, 0.f, d.cin, ST, P + Y.bc.off));
    // decoder MLP, time-batched: JoinTable{s,c} -> [Dropout] -> Maxout -> Linear -> LogSoftMax
    // (model_chorowski_baseline.lua:53-59, model_chorowski_baseline_dropout.lua:56)
    if (dropmask) {
        const int64_t n = (int64_t)BT * (ST + A);
        mul_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(d.sc, dropmask, d.scm, n);
        S2S_LAUNCH_CHECK(ctx);
    }
    S2S_TRY(gemm_f32(ctx, false, true, (int)BT, M * MW, ST + A, 1.f, d.scm, ST + A, P + Y.Wm.off, ST + A, 0.f, mpre, M * MW, P + Y.bm.off));
    maxout_fwd_kernel<<<(unsigned)ceil_div64((int64_t)BT * M, 256), 256, 0, st>>>(mpre, (int64_t)BT, M, MW, d.mo, d.midx);
    S2S_LAUNCH_CHECK(ctx);
    const float* top = d.mo;
    if (Y.MLP == 2) {   // Linear(M,M) -> Maxout(M,M,MW)   (librispeech/model_vgg.lua:78-79)
        float* mpre2;
        S2S_ALLOC(mpre2, ar, float, BT * M * MW);
        S2S_TRY(gemm_f32(ctx, false, true, (int)BT, M, M, 1.f, d.mo, M, P + Y.Wl.off, M, 0.f, d.l1, M, P + Y.bl.off));
        S2S_TRY(gemm_f32(ctx, false, true, (int)BT, M * MW, M, 1.f, d.l1, M, P + Y.Wm2.off, M, 0.f, mpre2, M * MW, P + Y.bm2.off));
        maxout_fwd_kernel<<<(unsigned)ceil_div64((int64_t)BT * M, 256), 256, 0, st>>>(mpre2, (int64_t)BT, M, MW, d.mo2, d.midx2);
        S2S_LAUNCH_CHECK(ctx);
        top = d.mo2;
    }
    S2S_TRY(gemm_f32(ctx, false, true, (int)BT, V, M, 1.f, top, M, P + Y.Wo.off, M, 0.f, d.logp, V, P + Y.bo.off));
    logsoftmax_kernel<<<(unsigned)ceil_div64((int64_t)BT, 8), 256, 0, st>>>(d.logp, (int64_t)BT, V, logp_out);
    S2S_LAUNCH_CHECK(ctx);
    d.valid = true;
    return 0;
}

// =================================================================================================
// backward
// =================================================================================================
int decoder_backward(s2s_ctx* ctx, const Layout& Y, const float* P, float* G, const float* h, const int* lengths, int B, int Lmax,
                     const int* labels, const int* tlens, int T, const float* dropmask, float lambda, const float* dlogp, float* dh,
                     bool defer_wgrad) {
    S2S_REQUIRE(ctx->dec && ctx->dec->valid, "attention backward called without a preceding forward on this context");
    DecoderState& d = *ctx->dec;
    S2S_REQUIRE(d.B == B && d.Lmax == Lmax && d.T == T && d.Y.n == Y.n, "attention backward: shapes differ from the preceding forward");
    S2S_REQUIRE((dropmask != nullptr) == d.has_drop, "attention backward: dropmask presence differs from the forward");
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V, M = Y.M, MW = Y.MW, KF = Y.K > 0 ? Y.KF : 0;
    const size_t BT = (size_t)B * T;
    const int iBT = (int)BT;
    cudaStream_t st = ctx->stream;
    Arena& ar = ctx->arena;
    const bool carry_alpha = KF > 0 || lambda != 0.f;

    float *dlogits, *dmo, *dm, *dsc, *dA, *du_all, *dcin_all, *dc_all, *dq_all, *de_all, *dyin, *dVh;
    float *ds_carry, *dsu, *dac[2] = {nullptr, nullptr}, *GhT, *GzrT, *WjcT, *WsT, *duw = nullptr;
    S2S_ALLOC(dlogits, ar, float, BT * V);
    S2S_ALLOC(dmo, ar, float, BT * M);
    S2S_ALLOC(dm, ar, float, BT * M * MW);
    S2S_ALLOC(dsc, ar, float, BT * (ST + A));
    S2S_ALLOC(dA, ar, float, BT * 3 * ST);
    S2S_ALLOC(du_all, ar, float, BT * ST);
    S2S_ALLOC(dcin_all, ar, float, BT * ST);
    S2S_ALLOC(dc_all, ar, float, BT * A);
    S2S_ALLOC(dq_all, ar, float, BT * S);
    S2S_ALLOC(de_all, ar, float, BT * Lmax);
    S2S_ALLOC(dyin, ar, float, BT * ST);
    S2S_ALLOC(dVh, ar, float, (size_t)B * Lmax * S);
    S2S_ALLOC(ds_carry, ar, float, (size_t)B * ST);
    S2S_ALLOC(dsu, ar, float, (size_t)B * 2 * ST);
    if (carry_alpha) { S2S_ALLOC(dac[0], ar, float, (size_t)B * Lmax); S2S_ALLOC(dac[1], ar, float, (size_t)B * Lmax); }
    if (d.tw_valid) { GhT = d.GhT; GzrT = d.GzrT; WjcT = d.WjcT; WsT = d.WsT; }      // made by decoder_prepare, under the encoder
    else {
        S2S_ALLOC(GhT, ar, float, (size_t)2 * ST * ST);
        S2S_ALLOC(GzrT, ar, float, (size_t)2 * ST * 2 * ST);
        S2S_ALLOC(WjcT, ar, float, (size_t)A * ST);
        S2S_ALLOC(WsT, ar, float, (size_t)ST * S);
    }
    if (KF > 0) { S2S_ALLOC(duw, ar, float, (size_t)KF * S); S2S_CUDA(cudaMemsetAsync(duw, 0, (size_t)KF * S * sizeof(float), st)); }
    AttnScratch att;
    S2S_TRY(attn_scratch_alloc(ctx, ar, B, Lmax, S, A, KF, true, &att));

    // transposed copies of the recurrent-chain weights so every in-loop product is K-contiguous
    if (!d.tw_valid) {
        S2S_TRY(transpose_f32(ctx, P + Y.Gh.off, ST, 2 * ST, 2 * ST, GhT, ST));          // GhT [2ST, ST]
        S2S_TRY(transpose_f32(ctx, P + Y.Gz.off, 2 * ST, 2 * ST, 2 * ST, GzrT, 2 * ST)); // rows: z then r (contiguous segments)
        S2S_TRY(transpose_f32(ctx, d.Wjc, ST, A, A, WjcT, ST));                          // (W_j[:, :ST] W_c)^T  [A, ST]
        S2S_TRY(transpose_f32(ctx, P + Y.Ws.off, S, ST, ST, WsT, S));                    // WsT [ST, S]
    }
    // The last two links of the chain, d{s_{t-1},u} += {daz,dar} G_zr and dc_t = dc_mlp + du W_jc, as ONE product: du is linear in
    // {daz, dar, dah} (du = dah G_h[:, ST:] + {daz,dar} G_zr[:, ST:]), so dc_t = dc_mlp + {daz,dar,dah} . W3[2ST:] with
    //   W3 [(2ST + A), 3ST] = [ G_zr^T | 0 ;  W_jc^T G_zr^T[ST:] | W_jc^T G_h^T[ST:] ]
    // -- one dependent launch less per decoder step.
    float* W3 = nullptr;
    const bool fuse3 = 3 * ST <= 1024 && !dense_chain_on() && !decoder_cluster_backward_eligible(Y, Lmax, lambda);   // (the cluster kernel keeps its own weights)
    if (fuse3) {
        S2S_ALLOC(W3, ar, float, (size_t)(2 * ST + A) * 3 * ST);
        S2S_CUDA(cudaMemsetAsync(W3, 0, (size_t)(2 * ST + A) * 3 * ST * sizeof(float), st));
        S2S_CUDA(cudaMemcpy2DAsync(W3, (size_t)3 * ST * sizeof(float), GzrT, (size_t)2 * ST * sizeof(float), (size_t)2 * ST * sizeof(float), 2 * ST,
                                   cudaMemcpyDeviceToDevice, st));
        float* W3c = W3 + (size_t)2 * ST * 3 * ST;
        S2S_TRY(gemm_f32(ctx, false, false, A, 2 * ST, ST, 1.f, WjcT, ST, GzrT + (size_t)ST * 2 * ST, 2 * ST, 0.f, W3c, 3 * ST, nullptr, GemmBatch(), 1, 1));
        S2S_TRY(gemm_f32(ctx, false, false, A, ST, ST, 1.f, WjcT, ST, GhT + (size_t)ST * ST, ST, 0.f, W3c + 2 * ST, 3 * ST, nullptr, GemmBatch(), 1, 1));
    }

    // location path on the cluster kernel: the Jacobian of the energies w.r.t. alpha_{t-1} does not depend on any gradient -> one
    // throughput-bound launch over all steps (rows t = 0 and padded steps stay zero: they only ever multiply de = 0).  It needs nothing from
    // the MLP backward below, so with S2S_OVERLAP (default) it runs beside it on the side stream and joins before the time loop.
    int padl = 0;
    if (KF > 0) pad_lr(KF, &padl);
    const bool cluster_ok = decoder_cluster_backward_eligible(Y, Lmax, lambda);
    float* V1 = nullptr;
    struct StreamGuard {      // whatever path leaves this function, the context's stream is restored
        s2s_ctx* c; cudaStream_t s; bool on;
        ~StreamGuard() { if (on) c->stream = s; }
    } v1_guard{ctx, st, false};
    if (cluster_ok && KF > 0 && d.V1) {
        V1 = d.V1;                                   // formed behind the forward loop (decoder_forward(prefetch_v1))
        if (d.v1_pending) { S2S_CUDA(cudaStreamWaitEvent(st, ctx->ev[5], 0)); d.v1_pending = false; }
    } else if (cluster_ok && KF > 0) {
        static int overlap = -1;
        if (overlap < 0) { const char* e = getenv("S2S_OVERLAP"); overlap = e ? atoi(e) : 1; }
        S2S_ALLOC(V1, ar, float, BT * Lmax * KF);
        if (overlap && ctx->side[1] && st != ctx->side[1] && !ctx->wgrad_join_pending) {
            S2S_CUDA(cudaEventRecord(ctx->ev[2], st));
            S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev[2], 0));
            v1_guard.on = true;
            ctx->stream = ctx->side[1];
        }
        S2S_CUDA(cudaMemsetAsync(V1, 0, BT * Lmax * KF * sizeof(float), ctx->stream));
        S2S_TRY(attn_v1(ctx, d.Vh, d.q, P + Y.we.off, d.uw, d.alpha, lengths, tlens, B, Lmax, T, S, KF, padl, V1));
        if (v1_guard.on) {
            S2S_CUDA(cudaEventRecord(ctx->ev[3], ctx->side[1]));
            ctx->stream = st;
        }
    }

    // ---- time-batched MLP backward (model_chorowski_baseline.lua:53-59 reversed) ----------------
    logsoftmax_bwd_kernel<<<(unsigned)ceil_div64((int64_t)BT, 8), 256, 0, st>>>(d.logp, dlogp, (int64_t)BT, V, tlens, T, dlogits);
    S2S_LAUNCH_CHECK(ctx);
    // (only the data path is on the critical chain to the time loop; the weight gradients of these layers are formed with the other
    // deferred products after it)
    S2S_TRY(gemm_f32(ctx, false, false, iBT, M, V, 1.f, dlogits, V, P + Y.Wo.off, M, 0.f, dmo, M));
    float *dl1 = nullptr, *dm2 = nullptr;
    if (Y.MLP == 2) {   // second Maxout and Linear(M,M) backward (librispeech/model_vgg.lua:78-79)
        S2S_ALLOC(dl1, ar, float, BT * M);
        S2S_ALLOC(dm2, ar, float, BT * M * MW);
        maxout_bwd_kernel<<<(unsigned)ceil_div64((int64_t)BT * M, 256), 256, 0, st>>>(dmo, d.midx2, (int64_t)BT, M, MW, dm2);
        S2S_LAUNCH_CHECK(ctx);
        S2S_TRY(gemm_f32(ctx, false, false, iBT, M, M * MW, 1.f, dm2, M * MW, P + Y.Wm2.off, M, 0.f, dl1, M));
        S2S_TRY(gemm_f32(ctx, false, false, iBT, M, M, 1.f, dl1, M, P + Y.Wl.off, M, 0.f, dmo, M));
    }
    maxout_bwd_kernel<<<(unsigned)ceil_div64((int64_t)BT * M, 256), 256, 0, st>>>(dmo, d.midx, (int64_t)BT, M, MW, dm);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(gemm_f32(ctx, false, false, iBT, ST + A, M * MW, 1.f, dm, M * MW, P + Y.Wm.off, ST + A, 0.f, dsc, ST + A));
    if (dropmask) {
        const int64_t n = (int64_t)BT * (ST + A);
        mul_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(dsc, dropmask, dsc, n);
        S2S_LAUNCH_CHECK(ctx);
    }

    // ---- time loop, t = T-1 .. 0 (RNNAttention.lua:233) -------------------------------------------
    S2S_CUDA(cudaMemsetAsync(ds_carry, 0, (size_t)B * ST * sizeof(float), st));      // Recurrent.lua:134
    if (carry_alpha) S2S_CUDA(cudaMemsetAsync(dac[0], 0, (size_t)B * Lmax * sizeof(float), st));
    const int64_t ldsc = (int64_t)T * (ST + A), ldsu = (int64_t)T * 2 * ST, ldg = (int64_t)T * 3 * ST, lddA = (int64_t)T * 3 * ST;
    const int eb = ceil_div(B * ST, 256);
    bool clustered = false;      // the whole loop in one persistent cluster kernel (decoder_cluster.cu) when the shapes allow
    if (cluster_ok) {
        if (v1_guard.on) { S2S_CUDA(cudaStreamWaitEvent(st, ctx->ev[3], 0)); v1_guard.on = false; }
        S2S_TRY(decoder_cluster_backward(ctx, Y, P, h, lengths, B, Lmax, T, lambda, d, WsT, GhT, GzrT, WjcT, dsc, V1, dA, du_all, dc_all, dq_all, de_all, &clustered));
    }
    // elementwise GRU backward of the LAST step (ds_carry = 0); later steps get it fused into the W_s product
    if (!clustered) gru_bwd_e1_kernel<<<eb, 256, 0, st>>>(dsc + (size_t)(T - 1) * (ST + A), ldsc, ds_carry, d.gates + (size_t)(T - 1) * 3 * ST, ldg,
                                           d.su + (size_t)(T - 1) * 2 * ST, ldsu, B, ST, dA + (size_t)(T - 1) * 3 * ST, lddA, dsu);
    S2S_LAUNCH_CHECK(ctx);
    for (int t = T - 1; t >= 0 && !clustered; t--) {
        const int cur = (T - 1 - t) & 1;
        // the dependent chain between two attention backward steps:
        //   [ds_t and the elementwise GRU backward of step t] -> d{r*s, u} -> d{s_{t-1}, u} -> dc_t
        ChainLink L[4];
        int nl = 0;
        if (t < T - 1) {   // ds_t = dsu[:, :ST] + dq_{t+1} . W_s + ds_mlp[t], then the elementwise GRU backward of step t
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.mode = EPI_BWD_DS; e.ST = ST; e.dsu = dsu; e.dsc = dsc + (size_t)t * (ST + A); e.ld_dsc = ldsc;
            e.gates_n = d.gates + (size_t)t * 3 * ST; e.ld_gates = ldg; e.su_n = d.su + (size_t)t * 2 * ST; e.ld_sprev = ldsu;
            e.dA = dA + (size_t)t * 3 * ST; e.ld_dA = lddA;
            k.X = dq_all + (size_t)(t + 1) * S; k.ldx = (int64_t)T * S; k.K = S; k.W = WsT; k.ldw = S; k.N = ST;
        }
        {   // d{r*s, u} = dah . G_h ; dar ; dsu[:, :ST] += d(r*s) r
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.mode = EPI_BWD_DRHU; e.ST = ST; e.gates = d.gates + (size_t)t * 3 * ST; e.ld_gates = ldg;
            e.sprev = d.su + (size_t)t * 2 * ST; e.ld_sprev = ldsu; e.dA = dA + (size_t)t * 3 * ST; e.ld_dA = lddA; e.dsu = dsu;
            k.X = dA + (size_t)t * 3 * ST + 2 * ST; k.ldx = lddA; k.K = ST; k.W = GhT; k.ldw = ST; k.N = 2 * ST;
        }
        if (fuse3) {   // d{s_{t-1}, u} += {daz, dar} . G_{z,r}  and  dc_t, one launch (see W3 above)
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.mode = EPI_BWD_DSU_DC; e.ST = ST; e.dsu = dsu; e.out2 = du_all + (size_t)t * ST; e.ld_out2 = (int64_t)T * ST;
            e.dsc = dsc + (size_t)t * (ST + A) + ST; e.ld_dsc = ldsc; e.out = dc_all + (size_t)t * A; e.ld_out = (int64_t)T * A;
            k.X = dA + (size_t)t * 3 * ST; k.ldx = lddA; k.K = 3 * ST; k.W = W3; k.ldw = 3 * ST; k.N = 2 * ST + A;
        } else {   // d{s_{t-1}, u} += {daz, dar} . G_{z,r}
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.add = dsu; e.ld_add = 2 * ST; e.out = dsu; e.ld_out = 2 * ST;
            e.out2 = du_all + (size_t)t * ST; e.ld_out2 = (int64_t)T * ST; e.n2_start = ST;
            k.X = dA + (size_t)t * 3 * ST; k.ldx = lddA; k.K = 2 * ST; k.W = GzrT; k.ldw = 2 * ST; k.N = 2 * ST;
        }
        if (!fuse3) {   // dc_t = dc_mlp + du . (W_j[:, :ST] W_c)
            ChainLink& k = L[nl++]; DenseEpi& e = k.e;
            e.add = dsc + (size_t)t * (ST + A) + ST; e.ld_add = ldsc; e.out = dc_all + (size_t)t * A; e.ld_out = (int64_t)T * A;
            k.X = dsu + ST; k.ldx = 2 * ST; k.K = ST; k.W = WjcT; k.ldw = ST; k.N = A;
        }
        if (dense_chain_supported(L, nl)) {
            S2S_TRY(dense_chain(ctx, L, nl, B));
        } else {
            for (int i = 0; i < nl; i++) S2S_TRY(dense_small(ctx, L[i].X, L[i].ldx, B, L[i].K, L[i].W, L[i].ldw, L[i].N, L[i].e));
        }
        {   // attention step backward
            AttnLoc loc; loc.KF = KF; loc.padl = padl; loc.uw = d.uw;
            loc.alpha_prev = t ? d.alpha + (size_t)(t - 1) * Lmax : nullptr; loc.ld_aprev = (int64_t)T * Lmax;
            S2S_TRY(attn_step_bwd(ctx, att, d.Vh, h, d.q + (size_t)t * S, (int64_t)T * S, P + Y.we.off, lengths, B, Lmax, S, A, loc,
                                  d.alpha + (size_t)t * Lmax, (int64_t)T * Lmax, dc_all + (size_t)t * A, (int64_t)T * A,
                                  carry_alpha ? dac[cur] : nullptr, Lmax, d.pen + t, T, lambda,
                                  dq_all + (size_t)t * S, (int64_t)T * S, de_all + (size_t)t * Lmax, (int64_t)T * Lmax,
                                  carry_alpha ? dac[cur ^ 1] : nullptr, Lmax));
        }
    }

    // Nothing downstream reads the decoder's weight gradients before the gradient step, only dh gates the encoder backward.  With
    // defer_wgrad (model_backward, S2S_OVERLAP=1) the critical path below is attn_dvh -> dh and every weight-gradient product moves to the
    // low-priority side stream with a grid limited to the SMs the cluster kernels leave idle, under the encoder's recurrences
    // (same mechanism as gru_seq_backward; the caller joins through gru_seq_wgrad_join).
    static int overlap = -1;
    if (overlap < 0) { const char* e = getenv("S2S_OVERLAP"); overlap = e ? atoi(e) : 1; }
    const bool fork = defer_wgrad && overlap && ctx->side[1] && st != ctx->side[1];
    struct SwapGuard {
        s2s_ctx* c; cudaStream_t s; bool on;
        ~SwapGuard() { if (on) { c->stream = s; c->gemm_sm_limit = 0; } }
    } guard{ctx, st, false};
    const int side_limit = ctx->sm_count - 112 > 16 ? ctx->sm_count - 112 : 0;
    auto to_side = [&]() { if (fork) { guard.on = true; ctx->stream = ctx->side[1]; ctx->gemm_sm_limit = side_limit; } };
    auto to_main = [&]() { if (fork) { ctx->stream = st; ctx->gemm_sm_limit = 0; guard.on = false; } };
    if (fork) {
        S2S_CUDA(cudaEventRecord(ctx->ev[2], st));
        S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev[2], 0));
    }

    // ---- critical path: deferred dVh / dw_e (/ dUW), then dh ---------------------------------------
    // dh = alpha^T dc (context path, Attention.lua:132-134: dh[b,l,:] = sum_t alpha_t[b,l] dc_t[b,:]) + dVh W_V.  The first term needs
    // the time loop only, so with the side stream it is formed beside attn_dvh and the data-gradient product accumulates onto it.
    const int BL = B * Lmax;
    {
        to_side();
        ctx->gemm_sm_limit = 0;
        GemmBatch gb; gb.count = B; gb.sA = (int64_t)T * Lmax; gb.sB = (int64_t)T * A; gb.sC = (int64_t)Lmax * A;
        S2S_TRY(gemm_f32(ctx, true, false, Lmax, A, T, 1.f, d.alpha, Lmax, dc_all, A, 0.f, dh, A, nullptr, gb, 1, 1));
        if (fork) S2S_CUDA(cudaEventRecord(ctx->ev[5], ctx->side[1]));
        to_main();
    }
    {
        AttnLoc loc; loc.KF = KF; loc.padl = padl; loc.uw = d.uw; loc.alpha_prev = d.alpha;
        S2S_TRY(attn_dvh(ctx, d.Vh, d.q, de_all, P + Y.we.off, lengths, tlens, B, Lmax, T, S, loc, dVh, G + Y.we.off, duw));
    }
    if (fork) {
        S2S_CUDA(cudaEventRecord(ctx->ev[4], st));
        S2S_CUDA(cudaStreamWaitEvent(st, ctx->ev[5], 0));
    }
    // TemporalConvolutionZeroBias backward, data part (TemporalConvolutionZeroBias.lua:42-48)
    S2S_TRY(gemm_f32(ctx, false, false, BL, A, S, 1.f, dVh, S, P + Y.WV.off, A, 1.f, dh, A));

    // ---- deferred weight gradients over M = B*T rows ----------------------------------------------
    to_side();
    const cudaStream_t ws = ctx->stream;
    // decoder MLP (model_chorowski_baseline.lua:53-59): Linear(M,V), [Maxout, Linear(M,M)], Maxout's Linear(ST+A, M*MW)
    S2S_TRY(gemm_f32(ctx, true, false, V, M, iBT, 1.f, dlogits, V, Y.MLP == 2 ? d.mo2 : d.mo, M, 1.f, G + Y.Wo.off, M, nullptr, GemmBatch(), 8));
    S2S_TRY(colsum_add(ctx, dlogits, BT, V, V, G + Y.bo.off));
    if (Y.MLP == 2) {
        S2S_TRY(gemm_f32(ctx, true, false, M * MW, M, iBT, 1.f, dm2, M * MW, d.l1, M, 1.f, G + Y.Wm2.off, M, nullptr, GemmBatch(), 4));
        S2S_TRY(colsum_add(ctx, dm2, BT, M * MW, M * MW, G + Y.bm2.off));
        S2S_TRY(gemm_f32(ctx, true, false, M, M, iBT, 1.f, dl1, M, d.mo, M, 1.f, G + Y.Wl.off, M, nullptr, GemmBatch(), 8));
        S2S_TRY(colsum_add(ctx, dl1, BT, M, M, G + Y.bl.off));
    }
    S2S_TRY(gemm_f32(ctx, true, false, M * MW, ST + A, iBT, 1.f, dm, M * MW, d.scm, ST + A, 1.f, G + Y.Wm.off, ST + A, nullptr, GemmBatch(), 4));
    S2S_TRY(colsum_add(ctx, dm, BT, M * MW, M * MW, G + Y.bm.off));
    // decoder GRU (GRU.lua:23-26): dG_{z,r} += {daz,dar}^T {s,u} ; dG_h += dah^T {r*s,u}
    S2S_TRY(gemm_f32(ctx, true, false, 2 * ST, 2 * ST, iBT, 1.f, dA, 3 * ST, d.su, 2 * ST, 1.f, G + Y.Gz.off, 2 * ST, nullptr, GemmBatch(), 4));
    S2S_TRY(gemm_f32(ctx, true, false, ST, 2 * ST, iBT, 1.f, dA + 2 * ST, 3 * ST, d.rhu, 2 * ST, 1.f, G + Y.Gh.off, 2 * ST, nullptr, GemmBatch(), 4));
    // Linear(2ST,ST) on {c_in, y_in} (Attention.lua:151)
    S2S_TRY(gemm_f32(ctx, true, false, ST, ST, iBT, 1.f, du_all, ST, d.cin, ST, 1.f, G + Y.Wj.off, 2 * ST, nullptr, GemmBatch(), 4));
    S2S_TRY(gemm_f32(ctx, true, false, ST, ST, iBT, 1.f, du_all, ST, d.yin, ST, 1.f, G + Y.Wj.off + ST, 2 * ST, nullptr, GemmBatch(), 4));
    S2S_TRY(colsum_add(ctx, du_all, BT, ST, ST, G + Y.bj.off));
    // Linear(A,ST) on c (Attention.lua:150): dc_in = du . W_j[:, :ST], time-batched
    S2S_TRY(gemm_f32(ctx, false, false, iBT, ST, ST, 1.f, du_all, ST, P + Y.Wj.off, 2 * ST, 0.f, dcin_all, ST));
    S2S_TRY(gemm_f32(ctx, true, false, ST, A, iBT, 1.f, dcin_all, ST, d.sc + ST, ST + A, 1.f, G + Y.Wc.off, A, nullptr, GemmBatch(), 4));
    S2S_TRY(colsum_add(ctx, dcin_all, BT, ST, ST, G + Y.bc.off));
    // Linear(V,ST) on the one-hot label (Attention.lua:149)
    S2S_TRY(gemm_f32(ctx, false, false, iBT, ST, ST, 1.f, du_all, ST, P + Y.Wj.off + ST, 2 * ST, 0.f, dyin, ST));
    S2S_TRY(colsum_add(ctx, dyin, BT, ST, ST, G + Y.by.off));
    wy_scatter_kernel<<<(unsigned)BT, 128, 0, ws>>>(dyin, labels, B, T, ST, V, G + Y.Wy.off);
    S2S_LAUNCH_CHECK(ctx);
    // Ws (Attention.lua:66): dW_s += dq^T s_{t-1} ; db_s += sum dq
    S2S_TRY(gemm_f32(ctx, true, false, S, ST, iBT, 1.f, dq_all, S, d.su, 2 * ST, 1.f, G + Y.Ws.off, ST, nullptr, GemmBatch(), 4));
    S2S_TRY(colsum_add(ctx, dq_all, BT, S, S, G + Y.bs.off));

    // ---- products that need attn_dvh's outputs ------------------------------------------------------
    if (fork) S2S_CUDA(cudaStreamWaitEvent(ws, ctx->ev[4], 0));
    if (KF > 0) {
        float* dub;
        S2S_ALLOC(dub, ar, float, S);
        S2S_CUDA(cudaMemsetAsync(dub, 0, S * sizeof(float), ws));
        S2S_TRY(colsum_add(ctx, dq_all, BT, S, S, dub));
        loc_unfold_kernel<<<Y.K, 256, 0, ws>>>(P + Y.U.off, P + Y.WF.off, P + Y.bF.off, duw, dub, S, Y.K, KF,
                                                            G + Y.U.off, G + Y.WF.off, G + Y.bF.off);
        S2S_LAUNCH_CHECK(ctx);
    }
    // TemporalConvolutionZeroBias backward, weight part (TemporalConvolutionZeroBias.lua:50-54): gradBias stays zero
    S2S_TRY(gemm_f32(ctx, true, false, S, A, BL, 1.f, dVh, S, h, A, 1.f, G + Y.WV.off, A, nullptr, GemmBatch(), 8));
    if (fork) {
        S2S_CUDA(cudaEventRecord(ctx->ev[3], ctx->side[1]));
        ctx->wgrad_join_pending = true;
    }
    to_main();
    return 0;
}

}  // namespace s2s

// =================================================================================================
// single decoder step with explicit hidden state (decoder_base:forward, Attention.lua:366,402) and
// Attention:BeamSearch (Attention.lua:332-438) with the beams batched on the device
// =================================================================================================
namespace s2s {

__global__ void yin_step_kernel(const float* __restrict__ Wy, const float* __restrict__ by, const int* __restrict__ yprev,
                                int ST, int V, float* __restrict__ yin) {
    const int b = blockIdx.x;
    int y = yprev ? yprev[b] : -1;
    if (y >= V) y = -1;
    for (int i = threadIdx.x; i < ST; i += blockDim.x) yin[(size_t)b * ST + i] = by[i] + (y >= 0 ? Wy[(size_t)i * V + y] : 0.f);
}

int attention_step_impl(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const float* Vh, const int* lengths, int B, int Lmax,
                        const int* yprev, const float* alpha_prev, const float* s_prev, float* alpha, float* s, float* logp) {
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V, M = Y.M, MW = Y.MW, KF = Y.K > 0 ? Y.KF : 0;
    S2S_REQUIRE(M % 4 == 0, "attention_step: mlpDepth must be a multiple of 4");
    Arena& ar = ctx->arena;
    cudaStream_t st = ctx->stream;
    float *zeros, *q, *qbias, *uw = nullptr, *sc, *yin, *uy, *cin, *su, *rhu, *gates, *mpre, *mo;
    int* midx;
    S2S_ALLOC(zeros, ar, float, (size_t)B * ST);
    S2S_ALLOC(q, ar, float, (size_t)B * S);
    S2S_ALLOC(qbias, ar, float, S);
    if (KF > 0) S2S_ALLOC(uw, ar, float, (size_t)KF * S);
    S2S_ALLOC(sc, ar, float, (size_t)B * (ST + A));
    S2S_ALLOC(yin, ar, float, (size_t)B * ST);
    S2S_ALLOC(uy, ar, float, (size_t)B * ST);
    S2S_ALLOC(cin, ar, float, (size_t)B * ST);
    S2S_ALLOC(su, ar, float, (size_t)B * 2 * ST);
    S2S_ALLOC(rhu, ar, float, (size_t)B * 2 * ST);
    S2S_ALLOC(gates, ar, float, (size_t)B * 3 * ST);
    S2S_ALLOC(mpre, ar, float, (size_t)B * M * MW);
    S2S_ALLOC(mo, ar, float, (size_t)B * M);
    S2S_ALLOC(midx, ar, int, (size_t)B * M);
    AttnScratch att;
    S2S_TRY(attn_scratch_alloc(ctx, ar, B, Lmax, S, A, KF, false, &att));
    int padl = 0;
    if (KF > 0) {
        pad_lr(KF, &padl);
        loc_fold_kernel<<<ceil_div(S, 128), 128, 0, st>>>(P + Y.U.off, P + Y.WF.off, P + Y.bF.off, P + Y.bs.off, S, Y.K, KF, uw, qbias);
        S2S_LAUNCH_CHECK(ctx);
    } else {
        S2S_CUDA(cudaMemcpyAsync(qbias, P + Y.bs.off, S * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (!s_prev) { S2S_CUDA(cudaMemsetAsync(zeros, 0, (size_t)B * ST * sizeof(float), st)); s_prev = zeros; }
    S2S_CUDA(cudaMemcpy2DAsync(su, (size_t)2 * ST * 4, s_prev, (size_t)ST * 4, (size_t)ST * 4, B, cudaMemcpyDeviceToDevice, st));
    yin_step_kernel<<<B, 128, 0, st>>>(P + Y.Wy.off, P + Y.by.off, yprev, ST, V, yin);
    S2S_LAUNCH_CHECK(ctx);
    { DenseEpi e; e.bias = P + Y.bj.off; e.out = uy; e.ld_out = ST; S2S_TRY(dense_small(ctx, yin, ST, B, ST, P + Y.Wj.off + ST, 2 * ST, ST, e)); }
    { DenseEpi e; e.bias = qbias; e.out = q; e.ld_out = S; S2S_TRY(dense_small(ctx, s_prev, ST, B, ST, P + Y.Ws.off, ST, S, e)); }
    {
        AttnLoc loc; loc.KF = KF; loc.padl = padl; loc.uw = uw; loc.alpha_prev = alpha_prev; loc.ld_aprev = Lmax;
        S2S_TRY(attn_step_fwd(ctx, att, Vh, h, q, S, P + Y.we.off, lengths, B, Lmax, S, A, loc, alpha, Lmax, sc + ST, ST + A, nullptr, 0, 0.f, nullptr, 0));
    }
    { DenseEpi e; e.bias = P + Y.bc.off; e.out = cin; e.ld_out = ST; S2S_TRY(dense_small(ctx, sc + ST, ST + A, B, A, P + Y.Wc.off, A, ST, e)); }
    { DenseEpi e; e.add = uy; e.ld_add = ST; e.out = su + ST; e.ld_out = 2 * ST; e.out2 = rhu + ST; e.ld_out2 = 2 * ST; e.n2_start = 0;
      S2S_TRY(dense_small(ctx, cin, ST, B, ST, P + Y.Wj.off, 2 * ST, ST, e)); }
    { DenseEpi e; e.mode = EPI_GRU_ZR; e.ST = ST; e.gates = gates; e.ld_gates = 3 * ST; e.sprev = su; e.ld_sprev = 2 * ST; e.rh_out = rhu; e.ld_rh = 2 * ST;
      S2S_TRY(dense_small(ctx, su, 2 * ST, B, 2 * ST, P + Y.Gz.off, 2 * ST, 2 * ST, e)); }
    { DenseEpi e; e.mode = EPI_GRU_H; e.ST = ST; e.gates = gates; e.ld_gates = 3 * ST; e.sprev = su; e.ld_sprev = 2 * ST;
      e.s_out = sc; e.ld_s = ST + A; e.s_out2 = s; e.ld_s2 = ST;
      S2S_TRY(dense_small(ctx, rhu, 2 * ST, B, 2 * ST, P + Y.Gh.off, 2 * ST, ST, e)); }
    { DenseEpi e; e.bias = P + Y.bm.off; e.out = mpre; e.ld_out = M * MW; S2S_TRY(dense_small(ctx, sc, ST + A, B, ST + A, P + Y.Wm.off, ST + A, M * MW, e)); }
    maxout_fwd_kernel<<<(unsigned)ceil_div64((int64_t)B * M, 256), 256, 0, st>>>(mpre, B, M, MW, mo, midx);
    S2S_LAUNCH_CHECK(ctx);
    if (Y.MLP == 2) {   // Linear(M,M) -> Maxout(M,M,MW)   (librispeech/model_vgg.lua:78-79); mpre is reused for the second stage
        float* l1;
        S2S_ALLOC(l1, ar, float, (size_t)B * M);
        { DenseEpi e; e.bias = P + Y.bl.off; e.out = l1; e.ld_out = M; S2S_TRY(dense_small(ctx, mo, M, B, M, P + Y.Wl.off, M, M, e)); }
        { DenseEpi e; e.bias = P + Y.bm2.off; e.out = mpre; e.ld_out = M * MW; S2S_TRY(dense_small(ctx, l1, M, B, M, P + Y.Wm2.off, M, M * MW, e)); }
        maxout_fwd_kernel<<<(unsigned)ceil_div64((int64_t)B * M, 256), 256, 0, st>>>(mpre, B, M, MW, mo, midx);
        S2S_LAUNCH_CHECK(ctx);
    }
    { DenseEpi e; e.bias = P + Y.bo.off; e.out = logp; e.ld_out = V; S2S_TRY(dense_small(ctx, mo, M, B, M, P + Y.Wo.off, M, V, e)); }
    logsoftmax_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, st>>>(logp, B, V, nullptr);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

__global__ void replicate_rows_kernel(const float* __restrict__ src, int64_t n, int copies, float* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = src[i];
    for (int c = 0; c < copies; c++) dst[(size_t)c * n + i] = v;
}
// gather beam states: dst[k] = src[sel[k]]
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ sel, int width, float* __restrict__ dst) {
    const int k = blockIdx.x, s = sel[k];
    for (int i = threadIdx.x; i < width; i += blockDim.x) dst[(size_t)k * width + i] = src[(size_t)s * width + i];
}

// ---- beam bookkeeping on the device (Attention.lua:390-432) --------------------------------------------------------------------
// One block per label: adds the beams' scores to the step's log-probabilities (:404), takes the top (K0 - finished) candidates
// (:406-408, highest first, lowest flat index on ties -- torch.topk's order on distinct values), moves candidates that end in <eos> or
// hit the length limit to the finished list (:418-421) and makes the others the next beams (source row + label for the state gather).
struct BeamState { int nb, nfin, done, count; };
__global__ void __launch_bounds__(256)
beam_select_kernel(const float* __restrict__ lp, int V, int K0, int eos, int maxlen, int first, int ML, BeamState* __restrict__ bs,
                   const float* __restrict__ bp_in, float* __restrict__ bp_out, const int* __restrict__ seq_in, int* __restrict__ seq_out,
                   float* __restrict__ fin_p, int* __restrict__ fin_len, int* __restrict__ fin_seq, float* __restrict__ work,
                   unsigned char* __restrict__ used, int* __restrict__ ysel, int* __restrict__ ssel) {
    __shared__ float rv[8];
    __shared__ int ri[8];
    __shared__ int sel_i[64];
    __shared__ float sel_v[64];
    __shared__ int dst_slot[64];         // >= 0: next beam slot, < 0: -(finished slot) - 1
    __shared__ int s_nn, s_nfin;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!first && bs->done) return;
    const int nb = first ? 1 : bs->nb, nfin0 = first ? 0 : bs->nfin, count = first ? 0 : bs->count + 1;
    const int n = nb * V;
    for (int i = tid; i < n; i += 256) { work[i] = first ? lp[i] : lp[i] + bp_in[i / V]; used[i] = 0; }
    __syncthreads();
    const int want = K0 - nfin0, kk = want < n ? want : n;
    for (int a = 0; a < kk; a++) {
        float bv = 0.f; int bi = -1;
        for (int i = tid; i < n; i += 256)
            if (!used[i]) { const float v = work[i]; if (bi < 0 || v > bv) { bv = v; bi = i; } }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { rv[warp] = bv; ri[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; w++)
                if (ri[w] >= 0 && (bi < 0 || rv[w] > bv || (rv[w] == bv && ri[w] < bi))) { bv = rv[w]; bi = ri[w]; }
            sel_i[a] = bi; sel_v[a] = bv; used[bi] = 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int nn = 0, nf = nfin0;
        for (int k = 0; k < kk; k++) {
            const int j = sel_i[k] % V;
            if (j == eos || (!first && count == maxlen)) { dst_slot[k] = -nf - 1; fin_p[nf] = sel_v[k]; fin_len[nf] = count + 1; nf++; }
            else { dst_slot[k] = nn; ssel[nn] = sel_i[k] / V; ysel[nn] = j; bp_out[nn] = sel_v[k]; nn++; }
        }
        for (int m = nn; m < K0; m++) { ssel[m] = nn ? ssel[0] : 0; ysel[m] = nn ? ysel[0] : 0; }      // dead rows repeat a live one
        s_nn = nn; s_nfin = nf;
        bs->nb = nn; bs->nfin = nf; bs->count = count; bs->done = (nf >= K0 || nn == 0 || count >= maxlen) ? 1 : 0;
    }
    __syncthreads();
    // label sequences: the parent's `count` labels followed by the new one
    for (int e = tid; e < kk * (count + 1); e += 256) {
        const int k = e / (count + 1), t = e % (count + 1);
        const int src = sel_i[k] / V, j = sel_i[k] % V;
        const int v = t < count ? seq_in[(size_t)src * ML + t] : j;
        if (dst_slot[k] >= 0) seq_out[(size_t)dst_slot[k] * ML + t] = v;
        else fin_seq[(size_t)(-dst_slot[k] - 1) * ML + t] = v;
    }
}

int beam_search_impl(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, int L, int eos, int beam, int maxlen,
                     int* out_host, int* n_out_host, float* logp_out_host) {
    S2S_REQUIRE(L > 0 && beam > 0 && beam <= 64 && maxlen > 0, "beam_search: bad arguments (L=%d beam=%d maxlen=%d)", L, beam, maxlen);
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V;
    const int K0 = beam < V ? beam : V;
    cudaStream_t st = ctx->stream;
    // long-lived buffers of this call come from the persistent arena (attention_step_impl resets nothing
    // but bumps ctx->arena on every step, so the per-step scratch is rewound manually)
    ctx->persist.reset();
    if (ctx->dec) ctx->dec->valid = false;
    if (ctx->model) ctx->model->valid = false;
    Arena& pa = ctx->persist;
    float *hK, *VhK, *alpha[2], *sst[2], *lp;
    int *ysel, *ssel;
    S2S_ALLOC(hK, pa, float, (size_t)K0 * L * A);
    S2S_ALLOC(VhK, pa, float, (size_t)K0 * L * S);
    for (int i = 0; i < 2; i++) { S2S_ALLOC(alpha[i], pa, float, (size_t)K0 * L); S2S_ALLOC(sst[i], pa, float, (size_t)K0 * ST); }
    S2S_ALLOC(lp, pa, float, (size_t)K0 * V);
    S2S_ALLOC(ysel, pa, int, K0);
    S2S_ALLOC(ssel, pa, int, K0);
    int host_topk = 0;              // S2S_BEAM_HOST=1: the first version (scores to the host every label), kept for A/B tests
    { const char* e = getenv("S2S_BEAM_HOST"); if (e) host_topk = atoi(e); }
    replicate_rows_kernel<<<(unsigned)ceil_div64((int64_t)L * A, 256), 256, 0, st>>>(h, (int64_t)L * A, K0, hK);
    S2S_LAUNCH_CHECK(ctx);
    S2S_TRY(gemm_f32(ctx, false, true, K0 * L, S, A, 1.f, hK, A, P + Y.WV.off, A, 0.f, VhK, S));       // Attention.lua:355

    if (!host_topk) {
        // ---- device-resident search: no score ever leaves the GPU; the host reads 16 bytes of status every 8 labels ----------------
        const int ML = maxlen + 2;
        BeamState* bs; float *bp[2], *fin_p, *work; int *seq[2], *fin_len, *fin_seq; unsigned char* used;
        S2S_ALLOC(bs, pa, BeamState, 1);
        for (int i = 0; i < 2; i++) { S2S_ALLOC(bp[i], pa, float, K0); S2S_ALLOC(seq[i], pa, int, (size_t)K0 * ML); }
        S2S_ALLOC(fin_p, pa, float, K0); S2S_ALLOC(fin_len, pa, int, K0); S2S_ALLOC(fin_seq, pa, int, (size_t)K0 * ML);
        S2S_ALLOC(work, pa, float, (size_t)K0 * V); S2S_ALLOC(used, pa, unsigned char, (size_t)K0 * V);
        S2S_CUDA(cudaMemsetAsync(bs, 0, sizeof(BeamState), st));
        BeamState hs = {0, 0, 0, 0};
        auto status = [&]() -> int {
            S2S_CUDA(cudaMemcpyAsync(&hs, bs, sizeof(BeamState), cudaMemcpyDeviceToHost, st));
            S2S_CUDA(cudaStreamSynchronize(st));
            return 0;
        };
        // first step from the zero state (Attention.lua:366-387): one row, then its state replicated to every beam
        ctx->arena.reset();
        S2S_TRY(attention_step_impl(ctx, Y, P, hK, VhK, nullptr, 1, L, nullptr, nullptr, nullptr, alpha[0], sst[0], lp));
        int sp = 0;
        beam_select_kernel<<<1, 256, 0, st>>>(lp, V, K0, eos, maxlen, 1, ML, bs, bp[sp], bp[sp ^ 1], seq[sp], seq[sp ^ 1], fin_p, fin_len, fin_seq,
                                              work, used, ysel, ssel);
        S2S_LAUNCH_CHECK(ctx);
        sp ^= 1;
        gather_rows_kernel<<<K0, 128, 0, st>>>(alpha[0], ssel, L, alpha[1]);
        S2S_LAUNCH_CHECK(ctx);
        gather_rows_kernel<<<K0, 128, 0, st>>>(sst[0], ssel, ST, sst[1]);
        S2S_LAUNCH_CHECK(ctx);
        S2S_TRY(status());
        const int cur = 1;
        for (int count = 1; count <= maxlen && !hs.done; count++) {                        // Attention.lua:390
            ctx->arena.reset();
            // all K0 rows every label (finished beams leave dead rows that repeat a live one): the batch shape never depends on device state
            S2S_TRY(attention_step_impl(ctx, Y, P, hK, VhK, nullptr, K0, L, ysel, alpha[cur], sst[cur], alpha[cur ^ 1], sst[cur ^ 1], lp));
            beam_select_kernel<<<1, 256, 0, st>>>(lp, V, K0, eos, maxlen, 0, ML, bs, bp[sp], bp[sp ^ 1], seq[sp], seq[sp ^ 1], fin_p, fin_len,
                                                  fin_seq, work, used, ysel, ssel);
            S2S_LAUNCH_CHECK(ctx);
            sp ^= 1;
            gather_rows_kernel<<<K0, 128, 0, st>>>(alpha[cur ^ 1], ssel, L, alpha[cur]);
            S2S_LAUNCH_CHECK(ctx);
            gather_rows_kernel<<<K0, 128, 0, st>>>(sst[cur ^ 1], ssel, ST, sst[cur]);
            S2S_LAUNCH_CHECK(ctx);
            if ((count & 7) == 0 || count == maxlen) S2S_TRY(status());                     // (steps issued after `done` change nothing)
        }
        if (!hs.done) S2S_TRY(status());
        *n_out_host = 0;
        if (hs.nfin > 0) {
            std::vector<float> fp(hs.nfin); std::vector<int> fl(hs.nfin), fs((size_t)hs.nfin * ML);
            S2S_CUDA(cudaMemcpyAsync(fp.data(), fin_p, hs.nfin * sizeof(float), cudaMemcpyDeviceToHost, st));
            S2S_CUDA(cudaMemcpyAsync(fl.data(), fin_len, hs.nfin * sizeof(int), cudaMemcpyDeviceToHost, st));
            S2S_CUDA(cudaMemcpyAsync(fs.data(), fin_seq, (size_t)hs.nfin * ML * sizeof(int), cudaMemcpyDeviceToHost, st));
            S2S_CUDA(cudaStreamSynchronize(st));
            int best = 0;
            for (int k = 1; k < hs.nfin; k++) if (fp[k] > fp[best]) best = k;               // Attention.lua:435
            *n_out_host = fl[best];
            for (int i = 0; i < fl[best]; i++) out_host[i] = fs[(size_t)best * ML + i];
            if (logp_out_host) *logp_out_host = fp[best];
        }
        return 0;
    }

    struct Hyp { std::vector<int> y; float p; };
    std::vector<Hyp> beams, fin;
    std::vector<float> lph((size_t)K0 * V);
    std::vector<int> hy(K0), hs(K0);
    auto topk = [&](const float* v, int n, int k, std::vector<int>& idx) {   // descending, lowest index first on ties
        idx.clear();
        std::vector<char> used(n, 0);
        for (int a = 0; a < k; a++) {
            int best = -1;
            for (int i = 0; i < n; i++) if (!used[i] && (best < 0 || v[i] > v[best])) best = i;
            used[best] = 1; idx.push_back(best);
        }
    };
    // first step from the zero state (Attention.lua:366-387)
    ctx->arena.reset();
    S2S_TRY(attention_step_impl(ctx, Y, P, hK, VhK, nullptr, 1, L, nullptr, nullptr, nullptr, alpha[0], sst[0], lp));
    S2S_CUDA(cudaMemcpyAsync(lph.data(), lp, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
    S2S_CUDA(cudaStreamSynchronize(st));
    std::vector<int> idx;
    topk(lph.data(), V, K0, idx);
    int cur = 0;
    {
        int nb = 0;
        for (int k = 0; k < K0; k++) {
            Hyp hyp; hyp.y.push_back(idx[k]); hyp.p = lph[idx[k]];
            if (idx[k] == eos) fin.push_back(hyp);
            else { beams.push_back(hyp); hs[nb] = 0; hy[nb] = idx[k]; nb++; }
        }
        if (nb > 0) {
            S2S_CUDA(cudaMemcpyAsync(ssel, hs.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
            S2S_CUDA(cudaMemcpyAsync(ysel, hy.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
            gather_rows_kernel<<<nb, 128, 0, st>>>(alpha[0], ssel, L, alpha[1]);
            S2S_LAUNCH_CHECK(ctx);
            gather_rows_kernel<<<nb, 128, 0, st>>>(sst[0], ssel, ST, sst[1]);
            S2S_LAUNCH_CHECK(ctx);
            cur = 1;
        }
    }
    int count = 0;
    while ((int)fin.size() < K0 && count < maxlen && !beams.empty()) {                 // Attention.lua:390
        count++;
        const int nb = (int)beams.size();
        ctx->arena.reset();
        S2S_TRY(attention_step_impl(ctx, Y, P, hK, VhK, nullptr, nb, L, ysel, alpha[cur], sst[cur], alpha[cur ^ 1], sst[cur ^ 1], lp));
        S2S_CUDA(cudaMemcpyAsync(lph.data(), lp, (size_t)nb * V * 4, cudaMemcpyDeviceToHost, st));
        S2S_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < nb; k++) for (int j = 0; j < V; j++) lph[(size_t)k * V + j] += beams[k].p;     // :404
        const int want = K0 - (int)fin.size();
        const int kk = want < nb * V ? want : nb * V;
        topk(lph.data(), nb * V, kk, idx);                                               // :406-408
        std::vector<Hyp> nxt;
        int nn = 0;
        for (int k = 0; k < kk; k++) {
            const int i = idx[k] / V, j = idx[k] % V;
            Hyp hyp = beams[i]; hyp.y.push_back(j); hyp.p = lph[(size_t)i * V + j];
            if (j == eos || count == maxlen) fin.push_back(hyp);                          // :418-421
            else { nxt.push_back(hyp); hs[nn] = i; hy[nn] = j; nn++; }
        }
        beams.swap(nxt);
        if (nn > 0) {
            S2S_CUDA(cudaMemcpyAsync(ssel, hs.data(), nn * sizeof(int), cudaMemcpyHostToDevice, st));
            S2S_CUDA(cudaMemcpyAsync(ysel, hy.data(), nn * sizeof(int), cudaMemcpyHostToDevice, st));
            gather_rows_kernel<<<nn, 128, 0, st>>>(alpha[cur ^ 1], ssel, L, alpha[cur]);
            S2S_LAUNCH_CHECK(ctx);
            gather_rows_kernel<<<nn, 128, 0, st>>>(sst[cur ^ 1], ssel, ST, sst[cur]);
            S2S_LAUNCH_CHECK(ctx);
            S2S_CUDA(cudaStreamSynchronize(st));   // hs/hy are reused by the next iteration
        }
    }
    *n_out_host = 0;
    if (!fin.empty()) {
        size_t best = 0;
        for (size_t k = 1; k < fin.size(); k++) if (fin[k].p > fin[best].p) best = k;       // :435
        *n_out_host = (int)fin[best].y.size();
        for (size_t i = 0; i < fin[best].y.size(); i++) out_host[i] = fin[best].y[i];
        if (logp_out_host) *logp_out_host = fin[best].p;
    }
    return 0;
}

}  // namespace s2s

namespace s2s {
// plain linear use of the small-batch product kernel by other translation units (lstm_seq.cu)
int dense_small_linear(s2s_ctx* ctx, const float* X, int64_t ldx, int B, int K, const float* W, int ldw, int N, const float* bias,
                       const float* add, int64_t ld_add, float* out, int64_t ld_out) {
    DenseEpi e; e.bias = bias; e.add = add; e.ld_add = ld_add; e.out = out; e.ld_out = ld_out;
    return dense_small(ctx, X, ldx, B, K, W, ldw, N, e);
}
}  // namespace s2s
