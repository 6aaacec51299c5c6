// common.cuh -- context, workspace arena, error plumbing and device helpers shared by all
// translation units of libs2s_b200.so (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/s2s_b200.h"

namespace s2s {

// ---------------------------------------------------------------------------------------------
// error plumbing: every API returns int; message kept thread-local (s2s_last_error)
// ---------------------------------------------------------------------------------------------
std::string& last_error();
int fail(const char* fmt, ...);

#define S2S_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return ::s2s::fail("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define S2S_TRY(call)                                                                           \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != 0) return r__;                                                               \
    } while (0)
#define S2S_REQUIRE(cond, ...)                                                                  \
    do {                                                                                        \
        if (!(cond)) return ::s2s::fail(__VA_ARGS__);                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------
// workspace arena: bump allocator over a list of cudaMalloc'd chunks.  reset() rewinds; the
// allocation sequence of a call is a pure function of its shapes, so pointers repeat across
// calls with equal shapes (required for CUDA-graph replay).
// ---------------------------------------------------------------------------------------------
struct Arena {
    struct Chunk { char* base; size_t size; size_t used; };
    std::vector<Chunk> chunks;
    size_t total = 0;
    bool frozen = false;   // set while a CUDA graph that references arena memory exists
    void* alloc(size_t bytes);            // returns nullptr on failure (message in last_error)
    uint64_t epoch = 0;    // number of resets: pointers handed out in an earlier epoch are stale
    void reset() { for (auto& c : chunks) c.used = 0; epoch++; }
    // bump-pointer snapshot / restore: scratch of a loop iteration is handed back for the next one (same stream => ordered)
    std::vector<size_t> mark() const { std::vector<size_t> m; for (auto& c : chunks) m.push_back(c.used); return m; }
    void rewind(const std::vector<size_t>& m) { for (size_t i = 0; i < chunks.size(); i++) chunks[i].used = i < m.size() ? m[i] : 0; }
    void release();
    template <typename T> T* get(size_t n) { return (T*)alloc(n * sizeof(T)); }
};

struct ProfRec { int cls; cudaEvent_t a, b; };
struct Prof {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double work[S2S_PROF_N] = {0};
    int64_t count[S2S_PROF_N] = {0};
};

// cached CUDA graph of one s2s_model_fwdbwd call signature (shapes + every pointer argument)
struct GraphCache {
    std::vector<uint64_t> key;
    int seen = 0;                  // eager calls with this key so far
    bool nocapture = false;        // capture failed once for this key: stay eager
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;          // kernel launches recorded in the graph
    int64_t kc[S2S_KC_N] = {0};    // per-class launch counts recorded in the graph
};

// prepared (hi/lo-split, K-contiguous) GEMM operands that stay valid for the duration of a TcCacheScope (gemm_tc.cu)
struct TcCacheEntry { const float* src; int K, cols, ld; bool transposed; const float* hi; const float* lo; int ld_hi, ld_lo; bool mn; };

struct DecoderState;   // decoder.cu
struct VggState;       // vgg.cu
struct ModelState;     // model.cu
struct DpState;        // dp_nccl.cu

}  // namespace s2s

struct s2s_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side[2] = {nullptr, nullptr};     // internal streams for independent branches
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int sm_count = 148;
    int64_t launches = 0;
    int64_t kcount[S2S_KC_N] = {0};   // launches per kernel class (s2s_ctx_kernel_count): which path ran is testable
    bool graphs = true;
    bool pdl = false;          // programmatic dependent launch for the decoder's per-step kernel chain (S2S_PDL=1 enables;
                               // measured neutral-to-negative under CUDA-graph replay, so off by default)
    s2s::Arena arena;          // per-call scratch (reset at the start of each top-level call)
    s2s::Arena persist;        // state that survives between forward and backward
    s2s::DecoderState* dec = nullptr;
    s2s::ModelState* model = nullptr;
    s2s::VggState* vgg = nullptr;
    s2s::DpState* dp = nullptr;     // NCCL communicator of the data-parallel plane (s2s_dp_init)
    bool dp_overlap = false;        // s2s_model_fwdbwd reduces the gradient buckets itself, under the remaining backward pass
    unsigned* counters = nullptr;   // zero-initialised device counters for last-block-done patterns
    uint64_t rng_calls = 0;
    s2s::Prof prof;
    s2s::GraphCache graph;
    // caller-defined graphs (s2s_graph_begin / _end / _launch): sequences of library calls captured once and replayed
    std::vector<cudaGraphExec_t> user_graphs;
    std::vector<int64_t> user_graph_launches;
    std::vector<std::vector<int64_t>> user_graph_kc;
    bool capturing = false;
    cudaStream_t capture_user_stream = nullptr;
    int64_t capture_l0 = 0;
    std::vector<int64_t> capture_k0;
    bool tc_cache_on = false;
    std::vector<s2s::TcCacheEntry> tc_cache;
    // weight-gradient GEMMs of one encoder layer overlapped with the next layer's recurrence (S2S_OVERLAP=1): they run on side[1]
    // with a persistent grid limited to the SMs the cluster kernels leave idle
    int gemm_sm_limit = 0;
    bool wgrad_join_pending = false;
};

namespace s2s {

#define S2S_LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return ::s2s::fail("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

// event-timed scope around one kernel launch (no-ops unless s2s_ctx_profile is on)
void prof_begin(s2s_ctx* ctx, int cls);
void prof_end(s2s_ctx* ctx, int cls, double work);

#define S2S_ALLOC(ptr, arena, T, n)                                                             \
    do {                                                                                        \
        (ptr) = (arena).get<T>((size_t)(n));                                                    \
        if (!(ptr)) return 1;                                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// flat parameter layout (builder-defined; identical to the order the Lua shim's parameters()
// returns, so getParameters() flattens to it -- timit/timit.lua:172)
// ---------------------------------------------------------------------------------------------
struct Seg { int64_t off = 0; int rows = 0, cols = 0; };
struct Layout {
    int D, H, NL, S, A, ST, V, K, KF, M, MW, MLP;
    Seg enc[8][2][3];   // layer, direction (0 fwd, 1 reverse), gate (z, r, h~): [H, H+Din]  GRU.lua:23-26
    Seg WV, bV, Ws, bs, WF, bF, U, bU, we, be, Wy, by, Wc, bc, Wj, bj, Gz, Gr, Gh, Wm, bm, Wl, bl, Wm2, bm2, Wo, bo;
    int64_t n;
};
int make_layout(const s2s_model_cfg* cfg, Layout* Y);   // validates cfg; 0 on success

// ---------------------------------------------------------------------------------------------
// GEMM entry (gemm_simt.cu / gemm_tc.cu)
// ---------------------------------------------------------------------------------------------
struct GemmBatch { int count = 1; int64_t sA = 0, sB = 0, sC = 0; };
// C = alpha*op(A)*op(B) + beta*C (+bias[n]).  splitk > 1 => partial products are atomically added
// (requires beta == 1, i.e. accumulate-into-C semantics).
int gemm_f32(s2s_ctx* ctx, bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, const float* bias = nullptr,
             GemmBatch batch = GemmBatch(), int splitk = 1, int impl = 0, bool relu = false);   // relu: C = max(., 0) (needs splitk == 1)

// While a scope is alive, operands prepared for the tcgen05 GEMM are remembered and reused by later gemm_f32 calls that
// read the same matrix (or a column block of it, in the transposed forms).  The caller guarantees that the source
// matrices are not modified inside the scope.
struct TcCacheScope {
    s2s_ctx* ctx;
    explicit TcCacheScope(s2s_ctx* c) : ctx(c) { c->tc_cache.clear(); c->tc_cache_on = true; }
    ~TcCacheScope() { ctx->tc_cache.clear(); ctx->tc_cache_on = false; }
};

// column sums: out[n] += sum_m X[m, n]   (bias gradients)
int colsum_add(s2s_ctx* ctx, const float* X, int64_t M, int N, int ldx, float* out);
int transpose_f32(s2s_ctx* ctx, const float* in, int rows, int cols, int ld_in, float* out, int ld_out);   // out[c][r] = in[r][c]
int fill_f32(s2s_ctx* ctx, float* p, int64_t n, float v);

// data-parallel plane (dp_nccl.cu)
void dp_state_free(s2s_ctx* ctx);
int dp_allreduce_bucket(s2s_ctx* ctx, float* G, int64_t n, bool on_side_stream);   // sum over ranks, in place; no-op without a communicator
int dp_join(s2s_ctx* ctx);                                                          // context stream waits for the side-stream buckets
int dp_world(const s2s_ctx* ctx);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// tanh accurate to ~1e-7 absolute: 1 - 2/(exp(2x)+1) with ex2.approx + rcp.approx.
// (tanh.approx.f32 is only good to ~5e-4 and would break the 1e-4 parity bound.)
__device__ __forceinline__ float tanh_acc(float x) {
    float t = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, t + 1.0f);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// streaming 128-bit global load (read-once data: bypass L1 allocation)
__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// read-only scalar load the compiler may neither move nor re-execute (register prefetch of a later loop iteration: see the prefetch
// discipline in gru_seq3.cu's backward kernel)
__device__ __forceinline__ float ldg_pinned(const float* p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// 4 consecutive floats from a pointer that may not be 16-byte aligned (segments of the flat parameter
// vector start at arbitrary float offsets, e.g. after the 1-element bias of the `e` convolution)
__device__ __forceinline__ float4 ldg4_any(const float* p) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
// L2-coherent loads for data produced by other CTAs of the same launch
__device__ __forceinline__ float ldcg1(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may
// start while its predecessor in the stream is still running; everything it does before pdl_wait() must not
// depend on (or overwrite inputs of) recent kernels.  pdl_wait() returns once the predecessor grid has completed
// and its writes are visible; both calls are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait suspends the thread for a hardware-defined interval per poll; a phase that never completes (a programming
// error in a producer) traps after ~2^26 polls instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls > (1u << 26)) __trap();
    }
}
// global -> shared bulk copy (TMA engine, 1-D): bytes multiple of 16, both addresses 16B-aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif  // __CUDACC__

// launch with (pdl = true) or without the programmatic-stream-serialization attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace s2s
