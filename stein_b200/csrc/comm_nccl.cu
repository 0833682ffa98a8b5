// comm_nccl.cu -- built-in collective hooks on NCCL (NVLink 5 / NVSwitch on a B200 node).
//
// The reference has no distributed code; the particle-sharded
// iteration of SURVEY.md section 8(e) needs an all-gather of the particle / score shards, u64
// all-reduces of the median counters and one f64 all-reduce of sum(phi^2).  The hooks of
// `stein_comm` can be supplied by the host program (the Python package can back them with
// torch.distributed), but a callback into Python costs ~50-80 us per collective, i.e. more
// than the collective itself at these sizes.  Here the library drives NCCL directly on the
// context's stream: one ncclAllGather / ncclAllReduce enqueue per hook, no host round trip.
//
// NCCL is resolved at run time from the libnccl.so.2 that is already loaded in the process
// (PyTorch's bundled copy), so libstein_b200.so has no link-time dependency on it; only the
// public types and enum values of nccl.h are used.
#include <dlfcn.h>

#include "common.cuh"

namespace stein {

// the subset of nccl.h this file needs (values fixed by the NCCL ABI)
typedef struct ncclComm *nccl_comm_t;
typedef struct {
    char internal[128];
} nccl_unique_id;
enum { NCCL_SUCCESS = 0, NCCL_SUM = 0, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host program already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return api;
    api.GetUniqueId = (int (*)(nccl_unique_id *))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_unique_id, int))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    api.AllReduce =
        (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather;
    return api;
}

struct NcclState {
    stein_ctx *ctx;
    nccl_comm_t comm;       // collectives on the ctx stream
    nccl_comm_t comm_side;  // all-gathers on a caller-given stream, concurrent with the first one
};

static int hook_allreduce_u64(void *user, void *buf, int64_t count) {
    NcclState *st = (NcclState *)user;
    return nccl_api().AllReduce(buf, buf, (size_t)count, NCCL_UINT64, NCCL_SUM, st->comm, st->ctx->stream);
}
static int hook_allreduce_f64(void *user, void *buf, int64_t count) {
    NcclState *st = (NcclState *)user;
    return nccl_api().AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, st->comm, st->ctx->stream);
}
static int hook_allgather_f32(void *user, const void *send, void *recv, int64_t count) {
    NcclState *st = (NcclState *)user;
    return nccl_api().AllGather(send, recv, (size_t)count, NCCL_FLOAT32, st->comm, st->ctx->stream);
}

static int hook_allgather_f32_on(void *user, const void *send, void *recv, int64_t count, void *stream) {
    NcclState *st = (NcclState *)user;
    return nccl_api().AllGather(send, recv, (size_t)count, NCCL_FLOAT32, st->comm_side, (cudaStream_t)stream);
}

void nccl_release(stein_ctx *ctx) {
    NcclState *st = (NcclState *)ctx->nccl_state;
    if (!st) return;
    if (st->comm_side) nccl_api().CommDestroy(st->comm_side);
    if (st->comm) nccl_api().CommDestroy(st->comm);
    delete st;
    ctx->nccl_state = nullptr;
}

}  // namespace stein

using namespace stein;

extern "C" {

int stein_nccl_unique_id(void *id_out) {
    if (!id_out) return fail(nullptr, STEIN_ERR_INVALID, "null id buffer");
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(nullptr, STEIN_ERR_COMM, "libnccl.so.2 not available: %s", dlerror());
    nccl_unique_id id;
    const int rc = api.GetUniqueId(&id);
    if (rc != NCCL_SUCCESS)
        return fail(nullptr, STEIN_ERR_COMM, "ncclGetUniqueId: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
    memcpy(id_out, &id, sizeof(id));
    return STEIN_OK;
}

int stein_ctx_init_nccl(stein_ctx *ctx, int rank, int world, const void *id, const void *id_side) {
    STEIN_REQUIRE(ctx, ctx != nullptr && id != nullptr, "null ctx / id");
    STEIN_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(ctx, STEIN_ERR_COMM, "libnccl.so.2 not available");
    STEIN_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_release(ctx);
    if (world == 1) return stein_ctx_set_comm(ctx, nullptr);
    NcclState *st = new NcclState{ctx, nullptr, nullptr};
    nccl_unique_id uid;
    memcpy(&uid, id, sizeof(uid));
    int rc = api.CommInitRank(&st->comm, world, uid, rank);
    if (rc == NCCL_SUCCESS && id_side) {
        memcpy(&uid, id_side, sizeof(uid));
        rc = api.CommInitRank(&st->comm_side, world, uid, rank);
    }
    if (rc != NCCL_SUCCESS) {
        if (st->comm) api.CommDestroy(st->comm);
        delete st;
        return fail(ctx, STEIN_ERR_COMM, "ncclCommInitRank: %s", api.GetErrorString ? api.GetErrorString(rc) : "?");
    }
    ctx->nccl_state = st;
    stein_comm c{};
    c.rank = rank;
    c.world = world;
    c.user = st;
    c.allreduce_sum_u64 = hook_allreduce_u64;
    c.allreduce_sum_f64 = hook_allreduce_f64;
    c.allgather_f32 = hook_allgather_f32;
    c.allgather_f32_on = st->comm_side ? hook_allgather_f32_on : nullptr;
    return stein_ctx_set_comm(ctx, &c);
}

}  // extern "C"
