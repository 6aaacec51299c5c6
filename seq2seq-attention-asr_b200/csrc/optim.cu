// optim.cu -- weight noise, adaptive weight noise, the gradient step and the row-norm constraint as
// fused elementwise / reduction kernels over the flat parameter vector.
//
// Reference: WeightNoise.lua:17-35, AdaptiveWeightNoise.lua:27-104, timit/timit.lua:291-348,
// TrainUtils.lua:52-104, and `optim.adadelta` (un-vendored Torch7 optim; published algorithm).
// torch.randn cannot be reproduced on a GPU, so every sampler takes an optional injected N(0,1)
// buffer (parity tests) and otherwise draws from a Philox-4x32-10 counter stream.
#include "common.cuh"

namespace s2s {

// ---- Philox4x32-10 + Box-Muller ---------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t stream, uint64_t idx, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// four N(0,1) draws for element group idx (elements 4*idx .. 4*idx+3)
__device__ __forceinline__ void randn4(uint64_t seed, uint64_t stream, uint64_t idx, float (&z)[4]) {
    uint32_t u[4];
    philox4(seed, stream, idx, u);
    const float k = 2.3283064365386963e-10f;   // 2^-32
    const float u0 = ((float)u[0] + 0.5f) * k, u1 = ((float)u[1] + 0.5f) * k, u2 = ((float)u[2] + 0.5f) * k, u3 = ((float)u[3] + 0.5f) * k;
    const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c); z[0] = r0 * c; z[1] = r0 * s;
    __sincosf(6.283185307179586f * u3, &s, &c); z[2] = r1 * c; z[3] = r1 * s;
}

// sample = w + sigma * eps   (WeightNoise.lua:17-22)   |  mu + exp(s/2) * eps  (AdaptiveWeightNoise.lua:27-38)
__global__ void noise_sample_kernel(const float* __restrict__ w, const float* __restrict__ logvar, const float* __restrict__ eps,
                                    uint64_t seed, uint64_t stream, float sigma, int64_t n, float* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 elements
    const int64_t i0 = g * 4;
    if (i0 >= n) return;
    float z[4];
    if (eps) {
#pragma unroll
        for (int j = 0; j < 4; j++) z[j] = i0 + j < n ? eps[i0 + j] : 0.f;
    } else {
        randn4(seed, stream, (uint64_t)g, z);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j < n) {
            const float sd = logvar ? sqrtf(expf(logvar[i0 + j])) : sigma;
            out[i0 + j] = fmaf(z[j], sd, w[i0 + j]);
        }
    }
}

// generic double-precision block reduction helper
__device__ __forceinline__ void block_add_double(double v, double* dst) {
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(dst, v);
    }
    __syncthreads();
}

// stats[0] = sum mu ; (second pass) stats[1] = sum (mu - mean)^2, stats[2] = sum exp(s), stats[3] = sum s
__global__ void awn_stats1_kernel(const float* __restrict__ weight, int64_t n, double* __restrict__ stats) {
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a += (double)weight[i];
    block_add_double(a, stats + 0);
}
__global__ void awn_stats2_kernel(const float* __restrict__ weight, int64_t n, double* __restrict__ stats) {
    const double mean = stats[0] / (double)n;
    double q = 0.0, s2 = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = (double)weight[i] - mean;
        q += d * d;
        const float s = weight[n + i];
        s2 += (double)expf(s);
        ss += (double)s;
    }
    block_add_double(q, stats + 1);
    block_add_double(s2, stats + 2);
    block_add_double(ss, stats + 3);
}
// dmu = lambda (mu - a_mu)/a_s2 + g ; ds = lambda (sig2/(2 a_s2) - 1/2) + g^2 sig2 / 2     (AdaptiveWeightNoise.lua:82-104)
__global__ void awn_accgrad_kernel(const float* __restrict__ weight, const float* __restrict__ g, int64_t n, double lambda,
                                   const double* __restrict__ stats, float* __restrict__ gw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sig2 = exp((double)weight[n + i]);
    const double gi = (double)g[i];
    const double dLNds = 0.5 * gi * gi * sig2;
    if (lambda > 0) {
        const double amu = stats[0] / (double)n;
        double as2 = stats[2] / (double)n + stats[1] / (double)n;
        if (as2 < 1e-12) as2 = 1e-12;                                           // AdaptiveWeightNoise.lua:3,71
        gw[i] = (float)(lambda * ((double)weight[i] - amu) / as2 + gi);
        gw[n + i] = (float)(lambda * 0.5 / as2 * sig2 - lambda * 0.5 + dLNds);
    } else {
        gw[i] = g[i];
        gw[n + i] = (float)dLNds;
    }
}

// ---- gradient step -------------------------------------------------------------------------------
__global__ void sumsq_scaled_kernel(const float* __restrict__ g, int64_t n, float inv_batch, double* __restrict__ out) {
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = g[i] * inv_batch;
        a += (double)v * (double)v;
    }
    block_add_double(a, out);
}
// g = clip(g / batch) + wd p + noise_sigma N(0,1)      (timit.lua:292-315), one pass
__global__ void grad_finalize_kernel(float* __restrict__ g, const float* __restrict__ p, int64_t n, int batch, double maxnorm, float wd,
                                     const float* __restrict__ noise, uint64_t seed, uint64_t stream, float noise_sigma,
                                     const double* __restrict__ sumsq) {
    const int64_t grp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = grp * 4;
    if (i0 >= n) return;
    const double nrm = sqrt(*sumsq);
    const float clip = nrm > maxnorm ? (float)(maxnorm / nrm) : 1.f;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise_sigma != 0.f) {
        if (noise) {
#pragma unroll
            for (int j = 0; j < 4; j++) z[j] = i0 + j < n ? noise[i0 + j] : 0.f;
        } else {
            randn4(seed, stream, (uint64_t)grp, z);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j < n) {
            float v = g[i0 + j];
            if (batch > 1) v = v / (float)batch;
            v *= clip;
            if (wd > 0.f) v = fmaf(wd, p[i0 + j], v);
            if (noise_sigma != 0.f) v = fmaf(z[j], noise_sigma, v);
            g[i0 + j] = v;
        }
    }
}
// optim.adadelta: v = rho v + (1-rho) g^2 ; d = sqrt(a+eps)/sqrt(v+eps) g ; x -= d ; a = rho a + (1-rho) d^2
__global__ void adadelta_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ v, float* __restrict__ a,
                                int64_t n, float rho, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float vi = rho * v[i] + (1.f - rho) * gi * gi;
    const float d = sqrtf(a[i] + eps) / sqrtf(vi + eps) * gi;
    v[i] = vi;
    x[i] -= d;
    a[i] = rho * a[i] + (1.f - rho) * d * d;
}
// TrainUtils.columnNormConstraint: rows whose L2 norm (+1e-8) >= maxval are divided by (norm+1e-8)/maxval.
// One warp per row.
__global__ void rownorm_kernel(float* __restrict__ W, int64_t rows, int64_t cols, float maxval, int* __restrict__ nan_flag) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = W + row * cols;
    float ss = 0.f;
    for (int64_t j = lane; j < cols; j += 32) ss = fmaf(r[j], r[j], ss);
    ss = warp_sum(ss);
    if (ss != ss) { if (lane == 0) atomicExch(nan_flag, 1); return; }           // TrainUtils.lua:55-62
    const float norm = sqrtf(ss) + 1e-8f;
    if (norm >= maxval) {
        const float div = norm / maxval;
        for (int64_t j = lane; j < cols; j += 32) r[j] = r[j] / div;
    }
}

// every weight matrix of the model in ONE launch: a table of (offset, rows, cols, first global row) segments; warp per row
constexpr int RN_MAXSEG = 40;
struct RownormTable { int nseg; int64_t off[RN_MAXSEG]; int rows[RN_MAXSEG]; int cols[RN_MAXSEG]; int first[RN_MAXSEG + 1]; };
__global__ void rownorm_multi_kernel(float* __restrict__ P, const __grid_constant__ RownormTable t, float maxval, int* __restrict__ nan_flag) {
    const int grow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (grow >= t.first[t.nseg]) return;
    int sg = 0;
    while (grow >= t.first[sg + 1]) sg++;
    const int cols = t.cols[sg];
    float* r = P + t.off[sg] + (int64_t)(grow - t.first[sg]) * cols;
    float ss = 0.f;
    for (int j = lane; j < cols; j += 32) ss = fmaf(r[j], r[j], ss);
    ss = warp_sum(ss);
    if (ss != ss) { if (lane == 0) atomicExch(nan_flag, 1); return; }           // TrainUtils.lua:55-62
    const float norm = sqrtf(ss) + 1e-8f;
    if (norm >= maxval) {
        const float div = norm / maxval;
        for (int j = lane; j < cols; j += 32) r[j] = r[j] / div;
    }
}

static int grid1d(s2s_ctx* ctx, int64_t n, int per) {
    int64_t b = (n + per - 1) / per;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int rownorm_launch(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, double maxval, int* nan_flag_dev) {
    rownorm_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, ctx->stream>>>(W, rows, cols, (float)maxval, nan_flag_dev);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s

using namespace s2s;

extern "C" {

int s2s_weightnoise_sample(s2s_ctx* ctx, const float* w, const float* eps, uint64_t seed, float sigma, int64_t n, float* sample) {
    S2S_REQUIRE(ctx && w && sample && n > 0, "weightnoise_sample: bad arguments");
    noise_sample_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, ctx->stream>>>(w, nullptr, eps, seed, ctx->rng_calls++, sigma, n, sample);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
int s2s_awn_sample(s2s_ctx* ctx, const float* weight, const float* eps, uint64_t seed, int64_t n, float* sample) {
    S2S_REQUIRE(ctx && weight && sample && n > 0, "awn_sample: bad arguments");
    noise_sample_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, ctx->stream>>>(weight, weight + n, eps, seed, ctx->rng_calls++, 0.f, n, sample);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
static int awn_stats(s2s_ctx* ctx, const float* weight, int64_t n, double** stats_out) {
    ctx->arena.reset();
    double* stats;
    S2S_ALLOC(stats, ctx->arena, double, 4);
    S2S_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(double), ctx->stream));
    const int g = grid1d(ctx, n, 1024);
    awn_stats1_kernel<<<g, 256, 0, ctx->stream>>>(weight, n, stats);
    S2S_LAUNCH_CHECK(ctx);
    awn_stats2_kernel<<<g, 256, 0, ctx->stream>>>(weight, n, stats);
    S2S_LAUNCH_CHECK(ctx);
    *stats_out = stats;
    return 0;
}
int s2s_awn_forward(s2s_ctx* ctx, const float* weight, int64_t n, double lambda, double nll, double* L_host) {
    S2S_REQUIRE(ctx && weight && L_host && n > 0, "awn_forward: bad arguments");
    if (!(lambda > 0)) { *L_host = nll; return 0; }
    double* stats;
    S2S_TRY(awn_stats(ctx, weight, n, &stats));
    double h[4];
    S2S_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    double as2 = h[2] / (double)n + h[1] / (double)n;
    if (as2 < 1e-12) as2 = 1e-12;
    double KL = 0.5 * ((double)n * log(as2) - h[3]);            // AdaptiveWeightNoise.lua:73-77
    KL += 0.5 / as2 * h[1];
    KL += 0.5 / as2 * h[2] - (double)n / 2;
    *L_host = lambda * KL + nll;
    return 0;
}
int s2s_awn_accgrad(s2s_ctx* ctx, const float* weight, const float* g, int64_t n, double lambda, float* gradWeight) {
    S2S_REQUIRE(ctx && weight && g && gradWeight && n > 0, "awn_accgrad: bad arguments");
    double* stats = nullptr;
    if (lambda > 0) S2S_TRY(awn_stats(ctx, weight, n, &stats));
    awn_accgrad_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(weight, g, n, lambda, stats, gradWeight);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

int s2s_grad_finalize(s2s_ctx* ctx, float* g, const float* p, int64_t n, int batch, double maxnorm, double wd,
                      const float* noise, uint64_t seed, double noise_sigma, double* gradnorm_host) {
    S2S_REQUIRE(ctx && g && n > 0 && batch >= 1, "grad_finalize: bad arguments");
    S2S_REQUIRE(wd <= 0 || p, "grad_finalize: weight decay needs the parameter vector");
    ctx->arena.reset();
    double* ss;
    S2S_ALLOC(ss, ctx->arena, double, 1);
    S2S_CUDA(cudaMemsetAsync(ss, 0, sizeof(double), ctx->stream));
    sumsq_scaled_kernel<<<grid1d(ctx, n, 1024), 256, 0, ctx->stream>>>(g, n, 1.f / (float)batch, ss);
    S2S_LAUNCH_CHECK(ctx);
    grad_finalize_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, ctx->stream>>>(g, p, n, batch, maxnorm, (float)wd, noise, seed,
                                                                                          ctx->rng_calls++, (float)noise_sigma, ss);
    S2S_LAUNCH_CHECK(ctx);
    if (gradnorm_host) {
        double h;
        S2S_CUDA(cudaMemcpyAsync(&h, ss, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
        *gradnorm_host = sqrt(h);
    }
    return 0;
}
int s2s_adadelta(s2s_ctx* ctx, float* x, const float* g, float* v, float* a, int64_t n, double rho, double eps) {
    S2S_REQUIRE(ctx && x && g && v && a && n > 0, "adadelta: bad arguments");
    adadelta_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(x, g, v, a, n, (float)rho, (float)eps);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
int s2s_rownorm_constraint(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, double maxval, int* nan_host) {
    S2S_REQUIRE(ctx && W && rows > 0 && cols > 0, "rownorm_constraint: bad arguments");
    ctx->arena.reset();
    int* flag;
    S2S_ALLOC(flag, ctx->arena, int, 1);
    S2S_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    S2S_TRY(rownorm_launch(ctx, W, rows, cols, maxval, flag));
    if (nan_host) {
        S2S_CUDA(cudaMemcpyAsync(nan_host, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}
int s2s_model_rownorm_constraint(s2s_ctx* ctx, const s2s_model_cfg* cfg, float* P, double maxval, int* nan_host) {
    S2S_REQUIRE(ctx && P, "model_rownorm_constraint: bad arguments");
    Layout Y;
    S2S_TRY(make_layout(cfg, &Y));
    ctx->arena.reset();
    int* flag;
    S2S_ALLOC(flag, ctx->arena, int, 1);
    S2S_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    // every module with a .weight reachable through apply2graph (TrainUtils.lua:137-184): all weight matrices
    RownormTable t = {};
    t.first[0] = 0;
    auto go = [&](const Seg& s) -> int {
        if (!s.rows) return 0;
        S2S_REQUIRE(t.nseg < RN_MAXSEG, "model_rownorm_constraint: too many weight matrices");
        t.off[t.nseg] = s.off; t.rows[t.nseg] = s.rows; t.cols[t.nseg] = s.cols; t.first[t.nseg + 1] = t.first[t.nseg] + s.rows;
        t.nseg++;
        return 0;
    };
    for (int l = 0; l < Y.NL; l++) for (int d = 0; d < 2; d++) for (int g = 0; g < 3; g++) S2S_TRY(go(Y.enc[l][d][g]));
    S2S_TRY(go(Y.WV)); S2S_TRY(go(Y.Ws)); S2S_TRY(go(Y.WF)); S2S_TRY(go(Y.U)); S2S_TRY(go(Y.we));
    S2S_TRY(go(Y.Wy)); S2S_TRY(go(Y.Wc)); S2S_TRY(go(Y.Wj)); S2S_TRY(go(Y.Gz)); S2S_TRY(go(Y.Gr)); S2S_TRY(go(Y.Gh));
    S2S_TRY(go(Y.Wm)); S2S_TRY(go(Y.Wl)); S2S_TRY(go(Y.Wm2)); S2S_TRY(go(Y.Wo));
    rownorm_multi_kernel<<<(unsigned)ceil_div(t.first[t.nseg], 8), 256, 0, ctx->stream>>>(P, t, (float)maxval, flag);
    S2S_LAUNCH_CHECK(ctx);
    if (nan_host) {
        S2S_CUDA(cudaMemcpyAsync(nan_host, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

}  // extern "C"

// ---- nn.Dropout mask (model_chorowski_baseline_dropout.lua:56; Torch nn.Dropout v2: keep with prob 1-p, scale 1/(1-p)) ----
namespace s2s {
__global__ void dropout_mask_kernel(float p, uint64_t seed, uint64_t stream, int64_t n, float* __restrict__ mask) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = g * 4;
    if (i0 >= n) return;
    uint32_t u[4];
    philox4(seed, stream, (uint64_t)g, u);
    const float scale = 1.f / (1.f - p);
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (i0 + j < n) mask[i0 + j] = ((float)u[j] * 2.3283064365386963e-10f >= p) ? scale : 0.f;
}
}  // namespace s2s

extern "C" int s2s_dropout_mask(s2s_ctx* ctx, float p, uint64_t seed, int64_t n, float* mask) {
    S2S_REQUIRE(ctx && mask && n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments (p=%f)", (double)p);
    s2s::dropout_mask_kernel<<<(unsigned)s2s::ceil_div64(s2s::ceil_div64(n, 4), 256), 256, 0, ctx->stream>>>(p, seed, ctx->rng_calls++, n, mask);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}
