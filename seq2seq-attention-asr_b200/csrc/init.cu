// init.cu -- TrainUtils.orthogonalize (TrainUtils.lua:5-26; applied to every module with a weight by orthogonalizeGraph,
// librispeech/exp0_scriptchecker.lua:49-52): the weight (with its bias appended as one more column when the module has one) is
// replaced by the orthonormal factor of its QR decomposition, taken in the TALL orientation:
//     w [r, c]:   r >= c :  q = qr(w).Q            (orthonormal columns)
//                 r <  c :  q = qr(w^T).Q^T        (orthonormal rows)
// torch.qr is LAPACK geqrf + orgqr; the same Householder convention is used here (beta = -sign(alpha) ||x||, tau = (beta - alpha) / beta,
// v scaled to v_k = 1, Q = H_0 H_1 ... H_{n-1} applied to the leading columns of the identity), so the signs of the result match.
// Initialisation-time code: one CTA per matrix, the matrix stays in L2; O(m n^2) flops (a 769 x 448 Maxout layer: ~0.3 GFLOP).
#include "common.cuh"

namespace s2s {

constexpr int QR_THREADS = 1024;

__device__ __forceinline__ float qr_block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float s = lane < QR_THREADS / 32 ? red[lane] : 0.f;
        s = warp_sum(s);
        if (lane == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// A, Q: column-major [n][m] (column j at A + j*m), m >= n.  On exit Q holds the m x n orthonormal factor.
__global__ void __launch_bounds__(QR_THREADS, 1)
householder_qr_kernel(float* __restrict__ A, float* __restrict__ Q, float* __restrict__ tau, int m, int n) {
    __shared__ float red[33];
    __shared__ float s_tau, s_scale;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = QR_THREADS / 32;
    for (int k = 0; k < n; k++) {                                 // geqrf
        float* ck = A + (size_t)k * m;
        float ss = 0.f;
        for (int i = k + 1 + tid; i < m; i += QR_THREADS) ss = fmaf(ck[i], ck[i], ss);
        ss = qr_block_sum(ss, red);
        if (tid == 0) {
            const float alpha = ck[k];
            if (ss == 0.f) { s_tau = 0.f; s_scale = 0.f; }       // H = I
            else {
                const float beta = -copysignf(sqrtf(alpha * alpha + ss), alpha);
                s_tau = (beta - alpha) / beta;
                s_scale = 1.f / (alpha - beta);
                ck[k] = beta;
            }
            tau[k] = s_tau;
        }
        __syncthreads();
        const float t = s_tau, sc = s_scale;
        for (int i = k + 1 + tid; i < m; i += QR_THREADS) ck[i] *= sc;          // v (v_k = 1 implicit)
        __syncthreads();
        if (t != 0.f) {
            for (int j = k + 1 + warp; j < n; j += nw) {          // trailing columns: A_j -= tau (v . A_j) v
                float* cj = A + (size_t)j * m;
                float w = lane == 0 ? cj[k] : 0.f;
                for (int i = k + 1 + lane; i < m; i += 32) w = fmaf(ck[i], cj[i], w);
                w = warp_sum(w) * t;
                if (lane == 0) cj[k] -= w;
                for (int i = k + 1 + lane; i < m; i += 32) cj[i] = fmaf(-w, ck[i], cj[i]);
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < m * n; idx += QR_THREADS) Q[idx] = (idx % m) == (idx / m) ? 1.f : 0.f;   // leading columns of I
    __syncthreads();
    for (int k = n - 1; k >= 0; k--) {                            // orgqr: Q = H_k Q
        const float* ck = A + (size_t)k * m;
        const float t = tau[k];
        if (t != 0.f) {
            for (int j = k + warp; j < n; j += nw) {
                float* qj = Q + (size_t)j * m;
                float w = lane == 0 ? qj[k] : 0.f;
                for (int i = k + 1 + lane; i < m; i += 32) w = fmaf(ck[i], qj[i], w);
                w = warp_sum(w) * t;
                if (lane == 0) qj[k] -= w;
                for (int i = k + 1 + lane; i < m; i += 32) qj[i] = fmaf(-w, ck[i], qj[i]);
            }
        }
        __syncthreads();
    }
}

// gather the logical matrix w [r, C] (weight [r, c] row-major plus the optional bias column) into the tall column-major A,
// and scatter the factor back
__global__ void qr_gather_kernel(const float* __restrict__ W, const float* __restrict__ bias, int r, int c, int tall_is_transposed, int m,
                                 float* __restrict__ A) {
    const int C = c + (bias ? 1 : 0);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)r * C) return;
    const int i = (int)(idx / C), j = (int)(idx % C);                       // w[i][j]
    const float v = j < c ? W[(size_t)i * c + j] : bias[i];
    if (tall_is_transposed) A[(size_t)i * m + j] = v;                       // A = w^T (m = C, n = r): column i of A = row i of w
    else A[(size_t)j * m + i] = v;                                          // A = w   (m = r, n = C)
}
__global__ void qr_scatter_kernel(const float* __restrict__ Q, int r, int c, int tall_is_transposed, int m, float* __restrict__ W,
                                  float* __restrict__ bias) {
    const int C = c + (bias ? 1 : 0);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)r * C) return;
    const int i = (int)(idx / C), j = (int)(idx % C);
    const float v = tall_is_transposed ? Q[(size_t)i * m + j] : Q[(size_t)j * m + i];
    if (j < c) W[(size_t)i * c + j] = v; else bias[i] = v;
}

int orthogonalize(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, float* bias) {
    const int r = (int)rows, c = (int)cols, C = c + (bias ? 1 : 0);
    const int transposed = r < C;
    const int m = transposed ? C : r, n = transposed ? r : C;
    float *A, *Q, *tau;
    S2S_ALLOC(A, ctx->arena, float, (size_t)m * n);
    S2S_ALLOC(Q, ctx->arena, float, (size_t)m * n);
    S2S_ALLOC(tau, ctx->arena, float, n);
    const unsigned blocks = (unsigned)ceil_div64((int64_t)r * C, 256);
    qr_gather_kernel<<<blocks, 256, 0, ctx->stream>>>(W, bias, r, c, transposed, m, A);
    S2S_LAUNCH_CHECK(ctx);
    householder_qr_kernel<<<1, QR_THREADS, 0, ctx->stream>>>(A, Q, tau, m, n);
    S2S_LAUNCH_CHECK(ctx);
    qr_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(Q, r, c, transposed, m, W, bias);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s

using namespace s2s;
extern "C" __attribute__((visibility("default"))) int s2s_orthogonalize(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, float* bias) {
    S2S_REQUIRE(ctx && W && rows > 0 && cols > 0 && rows < (1 << 20) && cols < (1 << 20), "orthogonalize: bad arguments");
    ctx->arena.reset();
    return orthogonalize(ctx, W, rows, cols, bias);
}
