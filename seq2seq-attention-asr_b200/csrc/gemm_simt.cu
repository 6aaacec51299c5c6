// gemm_simt.cu -- exact-fp32 SIMT GEMM used for every dense contraction whose result must match the
// reference's fp32 BLAS to <= 1e-4 (small/odd shapes, bias gradients, weight gradients).  The large
// time-batched projections are routed to the tcgen05 kernel in gemm_tc.cu when it is enabled.
//
// Reference call sites this replaces: TemporalConvolutionZeroBias.lua:39,45,51 (Vh and its gradients),
// LinearZeroBias.lua:42,58,70 (GRU gate products, time-batched here), the stock nn.Linear products of
// Attention.lua:149-151 and model_chorowski_baseline.lua:56-57.
#include <algorithm>

#include "common.cuh"

namespace s2s {

constexpr int GEMM_BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(bool tA, bool tB, int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                 const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc,
                 const float* __restrict__ bias, int64_t sA, int64_t sB, int64_t sC, int splitk_relu) {
    constexpr int NT = (BM / TM) * (BN / TN);
    const bool relu = splitk_relu < 0;              // splitk_relu = -1: no split-K, ReLU epilogue
    const int splitk = relu ? 1 : splitk_relu;
    constexpr int LA = BM * GEMM_BK / NT;   // elements of A per thread per tile
    constexpr int LB = BN * GEMM_BK / NT;
    __shared__ __align__(16) float As[2][GEMM_BK][BM + 4];
    __shared__ __align__(16) float Bs[2][GEMM_BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int kbeg = 0, kend = K;
    if (splitk > 1) {
        int per = ((K + splitk - 1) / splitk + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
        kbeg = blockIdx.z * per;
        kend = min(K, kbeg + per);
        if (kbeg >= kend) return;
    } else {
        A += (int64_t)blockIdx.z * sA; B += (int64_t)blockIdx.z * sB; C += (int64_t)blockIdx.z * sC;
    }

    float ra[LA], rb[LB];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; i++) {
            int e = tid + i * NT, m, k;
            if (tA) { m = e % BM; k = e / BM; } else { k = e % GEMM_BK; m = e / GEMM_BK; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = tA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LB; i++) {
            int e = tid + i * NT, n, k;
            if (!tB) { n = e % BN; k = e / BN; } else { k = e % GEMM_BK; n = e / GEMM_BK; }
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < kend) v = tB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; i++) {
            int e = tid + i * NT, m, k;
            if (tA) { m = e % BM; k = e / BM; } else { k = e % GEMM_BK; m = e / GEMM_BK; }
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; i++) {
            int e = tid + i * NT, n, k;
            if (!tB) { n = e % BN; k = e / BN; } else { k = e % GEMM_BK; n = e / GEMM_BK; }
            Bs[buf][k][n] = rb[i];
        }
    };

    // thread tile: TM rows split in groups of 4 spread BM/(TM/4) apart (conflict-free float4 LDS)
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    constexpr int GM = TM / 4, GN = TN / 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    gload(kbeg);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += GEMM_BK) {
        const bool more = k0 + GEMM_BK < kend;
        if (more) gload(k0 + GEMM_BK);
#pragma unroll
        for (int k = 0; k < GEMM_BK; k++) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; g++) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / GM) + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; g++) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / GN) + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        int gm = m0 + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int gn = n0 + (j / 4) * (BN / GN) + tx * 4 + (j % 4);
            if (gn >= N) continue;
            float v = alpha * acc[i][j];
            float* c = C + (int64_t)gm * ldc + gn;
            if (splitk > 1) {
                if (bias && blockIdx.z == 0) v += bias[gn];
                atomicAdd(c, v);
            } else {
                if (bias) v += bias[gn];
                if (beta != 0.f) v += beta * (*c);
                *c = relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
}

int gemm_tc_f32(s2s_ctx* ctx, bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda,
                const float* B, int ldb, float beta, float* C, int ldc, const float* bias, bool* handled, bool force, bool relu);

int gemm_f32(s2s_ctx* ctx, bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
             int ldb, float beta, float* C, int ldc, const float* bias, GemmBatch batch, int splitk, int impl, bool relu) {
    if (M <= 0 || N <= 0) return 0;
    S2S_REQUIRE(K >= 0, "gemm: K<0");
    S2S_REQUIRE(!(splitk > 1 && batch.count > 1), "gemm: split-K and batching are exclusive");
    S2S_REQUIRE(!(splitk > 1 && beta != 1.f), "gemm: split-K requires beta == 1 (accumulate)");
    S2S_REQUIRE(!(relu && splitk > 1), "gemm: the ReLU epilogue cannot be combined with split-K");
    if (impl != 1 && batch.count == 1) {
        bool handled = false;
        S2S_TRY(gemm_tc_f32(ctx, tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, &handled, impl == 2, relu));
        if (handled) return 0;
        S2S_REQUIRE(impl != 2, "gemm: tcgen05 path requested but shape/alignment not supported (M=%d N=%d K=%d)", M, N, K);
    }
    if (K == 0) return 0;
    if (splitk > 1) {   // few output tiles and a long K (weight gradients over all frames): split K until the grid fills the SMs
        const long tiles64 = (long)ceil_div(M, 64) * ceil_div(N, 64);
        const int want = (int)std::min<long>((long)ctx->sm_count * 2 / std::max(tiles64, 1L), (long)K / 512);
        if (want > splitk) splitk = want;
    }
    int z = splitk > 1 ? splitk : batch.count;
    prof_begin(ctx, S2S_PROF_GEMM);
    // pick the tile so the grid covers the SMs: big tiles only when they still give >= 1 wave
    long tiles128 = (long)ceil_div(M, 128) * ceil_div(N, 128) * z;
    if (tiles128 >= ctx->sm_count) {
        dim3 grid(ceil_div(N, 128), ceil_div(M, 128), z);
        gemm_simt_kernel<128, 128, 8, 8><<<grid, 256, 0, ctx->stream>>>(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias,
                                                                       batch.sA, batch.sB, batch.sC, relu ? -1 : splitk);
    } else {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 64), z);
        gemm_simt_kernel<64, 64, 4, 4><<<grid, 256, 0, ctx->stream>>>(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias,
                                                                     batch.sA, batch.sB, batch.sC, relu ? -1 : splitk);
    }
    prof_end(ctx, S2S_PROF_GEMM, 2.0 * M * N * (double)K * batch.count);
    ctx->kcount[S2S_KC_GEMM_SIMT]++;
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s
