// core.cu -- context lifetime, workspace arena, flat parameter layout and small utility kernels.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace s2s {

std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return 1;
}

void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    for (auto& c : chunks) {
        if (c.size - c.used >= bytes) {
            void* p = c.base + c.used;
            c.used += bytes;
            return p;
        }
    }
    if (frozen) {
        fail("workspace arena is frozen by a captured graph but needs %zu more bytes", bytes);
        return nullptr;
    }
    size_t sz = bytes > (size_t)(64u << 20) ? bytes : (size_t)(64u << 20);
    if (sz < total) sz = total;   // geometric growth
    if (sz < bytes) sz = bytes;
    char* base = nullptr;
    cudaError_t e = cudaMalloc((void**)&base, sz);
    if (e != cudaSuccess) {
        fail("cudaMalloc(%zu) for workspace failed: %s", sz, cudaGetErrorString(e));
        return nullptr;
    }
    chunks.push_back({base, sz, bytes});
    total += sz;
    return base;
}
void Arena::release() {
    for (auto& c : chunks) cudaFree(c.base);
    chunks.clear();
    total = 0;
    frozen = false;
}

static cudaEvent_t prof_event(s2s_ctx* ctx) {
    if (!ctx->prof.pool.empty()) { cudaEvent_t e = ctx->prof.pool.back(); ctx->prof.pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
static int prof_class(const s2s_ctx* ctx, int cls) {
    return (cls == S2S_PROF_GEMM && ctx->side[1] && ctx->stream == ctx->side[1]) ? S2S_PROF_GEMM_SIDE : cls;
}
void prof_begin(s2s_ctx* ctx, int cls) {
    if (!ctx->prof.on) return;
    cls = prof_class(ctx, cls);
    ProfRec r; r.cls = cls; r.a = prof_event(ctx); r.b = prof_event(ctx);
    cudaEventRecord(r.a, ctx->stream);
    ctx->prof.recs.push_back(r);
}
void prof_end(s2s_ctx* ctx, int cls, double work) {
    if (!ctx->prof.on || ctx->prof.recs.empty()) return;
    cls = prof_class(ctx, cls);
    cudaEventRecord(ctx->prof.recs.back().b, ctx->stream);
    ctx->prof.work[cls] += work;
    ctx->prof.count[cls] += 1;
}

// Flat layout.  Reference sources for every segment:
//   encoder GRUs        timit/model_chorowski_baseline.lua:22-31, GRU.lua:23-26
//   Vh                  Attention.lua:44      Ws  Attention.lua:66
//   location conv / U   Attention.lua:90-91   e   Attention.lua:110
//   y/c/joint linears   Attention.lua:149-151
//   decoder GRU         model_chorowski_baseline.lua:50
//   Maxout + output     model_chorowski_baseline.lua:56-57, Maxout.lua:15
// ZeroBias convolutions keep their (dead) bias slot in the flat vector
// (TemporalConvolutionZeroBias.lua:14-16) so the parameter count matches the reference's.
static void seg(Seg* s, int64_t* off, int rows, int cols) {
    s->off = *off; s->rows = rows; s->cols = cols;
    *off += (int64_t)rows * cols;
}
int make_layout(const s2s_model_cfg* c, Layout* Y) {
    S2S_REQUIRE(c != nullptr, "cfg is NULL");
    S2S_REQUIRE(c->D > 0 && c->H > 0 && c->NL >= 0 && c->NL <= 8 && c->S > 0 && c->ST > 0 && c->V > 1 && c->K >= 0 && c->M > 0 && c->MW > 0,
                "invalid model cfg (D=%d H=%d NL=%d S=%d ST=%d V=%d K=%d KF=%d M=%d MW=%d)", c->D, c->H, c->NL, c->S, c->ST, c->V, c->K, c->KF, c->M, c->MW);
    S2S_REQUIRE(c->K == 0 || c->KF > 0, "K>0 needs KF>0");
    S2S_REQUIRE(c->MLP >= 0 && c->MLP <= 2, "invalid model cfg: MLP=%d (0/1 = Maxout-Linear, 2 = Maxout-Linear-Maxout-Linear)", c->MLP);
    memset(Y, 0, sizeof(*Y));
    Y->D = c->D; Y->H = c->H; Y->NL = c->NL; Y->S = c->S; Y->A = 2 * c->H; Y->ST = c->ST; Y->V = c->V;
    Y->K = c->K; Y->KF = c->KF; Y->M = c->M; Y->MW = c->MW; Y->MLP = c->MLP == 2 ? 2 : 1;
    int64_t o = 0;
    for (int l = 0; l < Y->NL; l++) {
        int din = l == 0 ? Y->D : 2 * Y->H;
        for (int d = 0; d < 2; d++)
            for (int g = 0; g < 3; g++) seg(&Y->enc[l][d][g], &o, Y->H, Y->H + din);
    }
    seg(&Y->WV, &o, Y->S, Y->A); seg(&Y->bV, &o, Y->S, 1);
    seg(&Y->Ws, &o, Y->S, Y->ST); seg(&Y->bs, &o, Y->S, 1);
    if (Y->K > 0) {
        seg(&Y->WF, &o, Y->K, Y->KF); seg(&Y->bF, &o, Y->K, 1);
        seg(&Y->U, &o, Y->S, Y->K); seg(&Y->bU, &o, Y->S, 1);
    }
    seg(&Y->we, &o, 1, Y->S); seg(&Y->be, &o, 1, 1);
    seg(&Y->Wy, &o, Y->ST, Y->V); seg(&Y->by, &o, Y->ST, 1);
    seg(&Y->Wc, &o, Y->ST, Y->A); seg(&Y->bc, &o, Y->ST, 1);
    seg(&Y->Wj, &o, Y->ST, 2 * Y->ST); seg(&Y->bj, &o, Y->ST, 1);
    seg(&Y->Gz, &o, Y->ST, 2 * Y->ST); seg(&Y->Gr, &o, Y->ST, 2 * Y->ST); seg(&Y->Gh, &o, Y->ST, 2 * Y->ST);
    seg(&Y->Wm, &o, Y->M * Y->MW, Y->ST + Y->A); seg(&Y->bm, &o, Y->M * Y->MW, 1);
    if (Y->MLP == 2) {   // librispeech/model_vgg.lua:78-79
        seg(&Y->Wl, &o, Y->M, Y->M); seg(&Y->bl, &o, Y->M, 1);
        seg(&Y->Wm2, &o, Y->M * Y->MW, Y->M); seg(&Y->bm2, &o, Y->M * Y->MW, 1);
    }
    seg(&Y->Wo, &o, Y->V, Y->M); seg(&Y->bo, &o, Y->V, 1);
    Y->n = o;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
__global__ void fill_kernel(float* p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
int fill_f32(s2s_ctx* ctx, float* p, int64_t n, float v) {
    if (n <= 0) return 0;
    if (v == 0.0f) { S2S_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), ctx->stream)); return 0; }
    int blocks = (int)((n + 255) / 256); if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    fill_kernel<<<blocks, 256, 0, ctx->stream>>>(p, n, v);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

// out[n] += sum_m X[m,n].  Block = 32 columns x 8 row-lanes, grid = (ceil(N/32), row splits).
__global__ void colsum_kernel(const float* __restrict__ X, int64_t M, int N, int ldx, float* __restrict__ out) {
    __shared__ float red[8][33];
    int col = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (col < N)
        for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8) acc += X[m * ldx + col];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += red[i][threadIdx.x];
        atomicAdd(out + col, s);
    }
}
int colsum_add(s2s_ctx* ctx, const float* X, int64_t M, int N, int ldx, float* out) {
    if (M <= 0 || N <= 0) return 0;
    int gy = (int)((M + 255) / 256); if (gy > 64) gy = 64; if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(ceil_div(N, 32), gy), dim3(32, 8), 0, ctx->stream>>>(X, M, N, ldx, out);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out, int ld_out) {
    __shared__ float t[32][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int r = blockIdx.y * 32 + i;
        t[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
    }
    __syncthreads();
    int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int c2 = blockIdx.x * 32 + i;
        if (c2 < cols && r2 < rows) out[(size_t)c2 * ld_out + r2] = t[threadIdx.x][i];
    }
}
int transpose_f32(s2s_ctx* ctx, const float* in, int rows, int cols, int ld_in, float* out, int ld_out) {
    transpose_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32)), dim3(32, 8), 0, ctx->stream>>>(in, rows, cols, ld_in, out, ld_out);
    S2S_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace s2s

using namespace s2s;

extern "C" {

int s2s_version(void) { return S2S_VERSION; }
const char* s2s_last_error(void) { return last_error().c_str(); }

int s2s_ctx_create(int device, void* stream, s2s_ctx** out) {
    S2S_REQUIRE(out != nullptr, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device available (%s); libs2s_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    S2S_REQUIRE(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
    // the per-kernel cudaFuncSetAttribute flags and cluster-occupancy figures are cached once per process, and they are per-device
    // state: every context of a process must live on the same device (one process per GPU, as bench.py and the Lua host run)
    static int process_device = -1;
    if (process_device < 0) process_device = device;
    S2S_REQUIRE(device == process_device, "this process already uses device %d: libs2s_b200 supports one device per process (one process per GPU)", process_device);
    S2S_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    S2S_CUDA(cudaGetDeviceProperties(&prop, device));
    S2S_REQUIRE(prop.major >= 10, "device %d is sm_%d%d; libs2s_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
    s2s_ctx* c = new s2s_ctx();
    c->device = device;
    { const char* e = getenv("S2S_PDL"); if (e) c->pdl = atoi(e) != 0; }
    { const char* e = getenv("S2S_GRAPHS"); if (e) c->graphs = atoi(e) != 0; }      // debugging: S2S_GRAPHS=0 keeps s2s_model_fwdbwd eager
    c->sm_count = prop.multiProcessorCount;
    c->stream = (cudaStream_t)stream;   // NULL = the legacy default stream (what cutorch uses, timit/timit.lua:39)
    c->own_stream = false;
    {   // side[0] (graph capture origin) at the highest priority, side[1] (overlapped branches) at the lowest
        int lo = 0, hi = 0;
        S2S_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        S2S_CUDA(cudaStreamCreateWithPriority(&c->side[0], cudaStreamNonBlocking, hi));
        S2S_CUDA(cudaStreamCreateWithPriority(&c->side[1], cudaStreamNonBlocking, lo));
    }
    for (int i = 0; i < 6; i++) S2S_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
    S2S_CUDA(cudaMalloc((void**)&c->counters, 4096 * sizeof(unsigned)));
    S2S_CUDA(cudaMemset(c->counters, 0, 4096 * sizeof(unsigned)));
    *out = c;
    return 0;
}

int s2s_ctx_set_stream(s2s_ctx* ctx, void* stream) {
    S2S_REQUIRE(ctx, "ctx is NULL");
    ctx->stream = (cudaStream_t)stream;
    return 0;
}
int s2s_ctx_synchronize(s2s_ctx* ctx) {
    S2S_REQUIRE(ctx, "ctx is NULL");
    S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int64_t s2s_ctx_launch_count(s2s_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t s2s_ctx_kernel_count(s2s_ctx* ctx, int k) { return (ctx && k >= 0 && k < S2S_KC_N) ? ctx->kcount[k] : -1; }
int s2s_ctx_set_graphs(s2s_ctx* ctx, int enable) {
    S2S_REQUIRE(ctx, "ctx is NULL");
    ctx->graphs = enable != 0;
    return 0;
}

int s2s_ctx_profile(s2s_ctx* ctx, int enable) {
    S2S_REQUIRE(ctx, "ctx is NULL");
    S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof.recs) { ctx->prof.pool.push_back(r.a); ctx->prof.pool.push_back(r.b); }
    ctx->prof.recs.clear();
    for (int i = 0; i < S2S_PROF_N; i++) { ctx->prof.work[i] = 0; ctx->prof.count[i] = 0; }
    ctx->prof.on = enable != 0;
    return 0;
}
int s2s_ctx_profile_read(s2s_ctx* ctx, double* ms_host, int64_t* count_host, double* work_host) {
    S2S_REQUIRE(ctx && ms_host && count_host && work_host, "profile_read: null argument");
    S2S_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < S2S_PROF_N; i++) { ms_host[i] = 0; count_host[i] = ctx->prof.count[i]; work_host[i] = ctx->prof.work[i]; }
    for (auto& r : ctx->prof.recs) {
        float ms = 0.f;
        S2S_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_host[r.cls] += ms;
    }
    return 0;
}

int64_t s2s_param_count(const s2s_model_cfg* cfg) {
    Layout Y;
    if (make_layout(cfg, &Y)) return -1;
    return Y.n;
}
int64_t s2s_decoder_param_offset(const s2s_model_cfg* cfg) {
    Layout Y;
    if (make_layout(cfg, &Y)) return -1;
    return Y.WV.off;
}
int s2s_param_segments(const s2s_model_cfg* cfg, int64_t* out, int max) {
    Layout Y;
    if (make_layout(cfg, &Y)) return -1;
    int n = 0;
    auto put = [&](const Seg& s) {
        if (s.rows == 0) return;
        if (n < max) { out[3 * n] = s.off; out[3 * n + 1] = s.rows; out[3 * n + 2] = s.cols; }
        n++;
    };
    for (int l = 0; l < Y.NL; l++) for (int d = 0; d < 2; d++) for (int g = 0; g < 3; g++) put(Y.enc[l][d][g]);
    put(Y.WV); put(Y.bV); put(Y.Ws); put(Y.bs); put(Y.WF); put(Y.bF); put(Y.U); put(Y.bU); put(Y.we); put(Y.be);
    put(Y.Wy); put(Y.by); put(Y.Wc); put(Y.bc); put(Y.Wj); put(Y.bj); put(Y.Gz); put(Y.Gr); put(Y.Gh);
    put(Y.Wm); put(Y.bm); put(Y.Wl); put(Y.bl); put(Y.Wm2); put(Y.bm2); put(Y.Wo); put(Y.bo);
    return n;
}

}  // extern "C"
