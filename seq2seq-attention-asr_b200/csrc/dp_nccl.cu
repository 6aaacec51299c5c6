// dp_nccl.cu -- the data-parallel plane behind the C ABI (SURVEY 8b group 6 / 8e): one NCCL communicator per context,
// all-reduce (sum) of the flat fp32 gradient over NVLink 5 / NVSwitch, bucketed so that a layer's gradients can be reduced on a
// side stream while the remaining recurrences of the backward pass still run.
//
// The reference has no collective at all (single device, timit/timit.lua:39); the minibatch is the only axis of the path that
// shards (timit/timit.lua:240-295 sums per-utterance gradients), so the whole distributed surface is:
//     s2s_dp_unique_id  (rank 0)  ->  [host side ships the 128 bytes to the other ranks]  ->  s2s_dp_init (every rank)
//     s2s_dp_allreduce / s2s_dp_broadcast on caller buffers;  s2s_dp_destroy.
// libnccl is opened at run time (dlopen): libs2s_b200.so itself has no link-time dependency on it, a host that never calls
// s2s_dp_* (single GPU, the Lua default) does not need NCCL installed.  Search order: $S2S_NCCL_LIB, a libnccl.so.2 already
// mapped into the process (e.g. the one a Python host's torch brought), the system libnccl.so.2.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace s2s {

void graphs_release(s2s_ctx* ctx);      // api.cu

typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };          // ncclFloat32, ncclSum (nccl.h)

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* env = getenv("S2S_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);      // reuse a copy the host process already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
    api.handle = h;
#define S2S_SYM(field, name) *(void**)(&api.field) = dlsym(h, name)
    S2S_SYM(GetUniqueId, "ncclGetUniqueId"); S2S_SYM(CommInitRank, "ncclCommInitRank"); S2S_SYM(AllReduce, "ncclAllReduce");
    S2S_SYM(Broadcast, "ncclBroadcast"); S2S_SYM(CommDestroy, "ncclCommDestroy"); S2S_SYM(GroupStart, "ncclGroupStart");
    S2S_SYM(GroupEnd, "ncclGroupEnd"); S2S_SYM(GetErrorString, "ncclGetErrorString"); S2S_SYM(GetVersion, "ncclGetVersion");
#undef S2S_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast || !api.CommDestroy || !api.GetErrorString) {
        dlclose(h); api.handle = nullptr;
        return nullptr;
    }
    return &api;
}

struct DpState {
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool join_pending = false;
};

#define S2S_NCCL(api, call)                                                                                          \
    do {                                                                                                             \
        int r__ = (call);                                                                                            \
        if (r__ != 0) return ::s2s::fail("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, r__, (api)->GetErrorString(r__)); \
    } while (0)

void dp_state_free(s2s_ctx* ctx) {
    DpState* d = ctx->dp;
    if (!d) return;
    NcclApi* api = nccl_api();
    if (d->comm && api) api->CommDestroy(d->comm);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->ev_join) cudaEventDestroy(d->ev_join);
    delete d;
    ctx->dp = nullptr;
}

// Bucketed overlap: reduce G[off, off+n) on the low-priority side stream once everything enqueued so far on the context's stream
// (the producer of that bucket) has finished; dp_join() makes the context's stream wait for every bucket issued since.  The calls are
// made at the same program points on every rank, so the NCCL operations are enqueued in the same order everywhere.  Inside a CUDA-graph
// capture the fork / join events become graph edges and the collective a graph node (NCCL supports capture).
int dp_allreduce_bucket(s2s_ctx* ctx, float* G, int64_t n, bool on_side_stream) {
    DpState* d = ctx->dp;
    if (!d || d->world <= 1 || n <= 0) return 0;
    NcclApi* api = nccl_api();
    S2S_REQUIRE(api, "dp: libnccl is not available");
    cudaStream_t st = ctx->stream;
    if (on_side_stream && ctx->side[1] && ctx->stream != ctx->side[1]) {
        S2S_CUDA(cudaEventRecord(d->ev_fork, ctx->stream));
        S2S_CUDA(cudaStreamWaitEvent(ctx->side[1], d->ev_fork, 0));
        st = ctx->side[1];
    }
    S2S_NCCL(api, api->AllReduce(G, G, (size_t)n, NCCL_FLOAT32, NCCL_SUM, d->comm, st));
    if (st != ctx->stream) {
        S2S_CUDA(cudaEventRecord(d->ev_join, st));
        d->join_pending = true;
    }
    return 0;
}
int dp_join(s2s_ctx* ctx) {
    DpState* d = ctx->dp;
    if (d && d->join_pending) {
        S2S_CUDA(cudaStreamWaitEvent(ctx->stream, d->ev_join, 0));
        d->join_pending = false;
    }
    return 0;
}
int dp_world(const s2s_ctx* ctx) { return ctx->dp ? ctx->dp->world : 1; }

}  // namespace s2s

using namespace s2s;

extern "C" {

int s2s_dp_available(void) { return nccl_api() != nullptr; }

int s2s_dp_unique_id(void* id_host_128) {
    S2S_REQUIRE(id_host_128, "dp_unique_id: null argument");
    NcclApi* api = nccl_api();
    S2S_REQUIRE(api, "dp_unique_id: libnccl.so.2 could not be opened (set S2S_NCCL_LIB)");
    nccl_uid id;
    S2S_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_host_128, &id, sizeof(id));
    return 0;
}

int s2s_dp_init(s2s_ctx* ctx, int rank, int world, const void* id_host_128) {
    S2S_REQUIRE(ctx && id_host_128, "dp_init: null argument");
    S2S_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dp_init: rank %d of %d", rank, world);
    S2S_REQUIRE(!ctx->dp, "dp_init: this context already has a communicator");
    NcclApi* api = nccl_api();
    S2S_REQUIRE(api, "dp_init: libnccl.so.2 could not be opened (set S2S_NCCL_LIB)");
    S2S_CUDA(cudaSetDevice(ctx->device));
    DpState* d = new DpState();
    d->rank = rank; d->world = world;
    nccl_uid id;
    memcpy(&id, id_host_128, sizeof(id));
    int r = api->CommInitRank(&d->comm, world, id, rank);
    if (r != 0) { delete d; return fail("dp_init: ncclCommInitRank -> %d (%s)", r, api->GetErrorString(r)); }
    cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming);
    ctx->dp = d;
    return 0;
}

int s2s_dp_rank(s2s_ctx* ctx) { return ctx && ctx->dp ? ctx->dp->rank : 0; }
int s2s_dp_world(s2s_ctx* ctx) { return ctx && ctx->dp ? ctx->dp->world : 1; }

int s2s_dp_allreduce(s2s_ctx* ctx, float* G, int64_t n) {
    S2S_REQUIRE(ctx && G && n >= 0, "dp_allreduce: bad argument");
    if (!ctx->dp || ctx->dp->world <= 1) return 0;       // single rank: the sum over ranks is the buffer itself
    S2S_TRY(dp_join(ctx));
    return dp_allreduce_bucket(ctx, G, n, false);
}

int s2s_dp_broadcast(s2s_ctx* ctx, float* P, int64_t n, int root) {
    S2S_REQUIRE(ctx && P && n >= 0, "dp_broadcast: bad argument");
    if (!ctx->dp || ctx->dp->world <= 1) return 0;
    NcclApi* api = nccl_api();
    S2S_REQUIRE(api, "dp: libnccl is not available");
    S2S_REQUIRE(root >= 0 && root < ctx->dp->world, "dp_broadcast: root %d of %d", root, ctx->dp->world);
    S2S_NCCL(api, api->Broadcast(P, P, (size_t)n, NCCL_FLOAT32, root, ctx->dp->comm, ctx->stream));
    return 0;
}

int s2s_dp_set_overlap(s2s_ctx* ctx, int enable) {
    S2S_REQUIRE(ctx, "dp_set_overlap: null context");
    if (ctx->graph.exec) return fail("dp_set_overlap: call it before the first s2s_model_fwdbwd (a captured step would keep the old setting)");
    ctx->dp_overlap = enable != 0;
    return 0;
}

int s2s_dp_destroy(s2s_ctx* ctx) {
    S2S_REQUIRE(ctx, "dp_destroy: null context");
    if (ctx->dp) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->side[1]) cudaStreamSynchronize(ctx->side[1]);
        graphs_release(ctx);          // captured collectives reference the communicator
        dp_state_free(ctx);
    }
    return 0;
}

}  // extern "C"
