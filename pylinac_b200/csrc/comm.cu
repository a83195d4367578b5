// Multi-GPU plumbing: the PF / WL / Starshot / FieldAnalysis batches shard by frame index with no data-path
// collective; the only exchange is the final gather of the fixed-size per-frame result structs,
// done with ONE ncclAllGather over NVLink.  NCCL is resolved at run time (dlopen) so that libepid.so loads on a
// box without it and shares whatever libnccl.so.2 the process already has.
#include <dlfcn.h>

#include "common.cuh"

namespace epid {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int /*ncclDataType_t*/, nccl_comm_t, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);

static struct {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
} N;

static int load_nccl() {
    if (N.lib) return EPID_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !N.lib; i++) N.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!N.lib) { set_error("NCCL not found: %s", dlerror()); return EPID_ERR_NCCL; }
    N.get_uid = (fn_get_uid)dlsym(N.lib, "ncclGetUniqueId");
    N.init_rank = (fn_init_rank)dlsym(N.lib, "ncclCommInitRank");
    N.all_gather = (fn_all_gather)dlsym(N.lib, "ncclAllGather");
    N.all_reduce = (fn_all_reduce)dlsym(N.lib, "ncclAllReduce");
    N.destroy = (fn_destroy)dlsym(N.lib, "ncclCommDestroy");
    N.errstr = (fn_errstr)dlsym(N.lib, "ncclGetErrorString");
    if (!N.get_uid || !N.init_rank || !N.all_gather || !N.destroy) { set_error("NCCL symbols missing"); return EPID_ERR_NCCL; }
    return EPID_OK;
}

#define EPID_NCCL(call)                                                                          \
    do {                                                                                         \
        int _r = (call);                                                                         \
        if (_r != 0) { set_error("%s failed: %s", #call, N.errstr ? N.errstr(_r) : "?"); return EPID_ERR_NCCL; } \
    } while (0)

}  // namespace epid

using namespace epid;

extern "C" {

int32_t epid_comm_unique_id(void* id128) {
    EPID_REQUIRE(id128, EPID_ERR_INVALID, "id buffer is NULL");
    int rc = load_nccl();
    if (rc != EPID_OK) return rc;
    nccl_uid id;
    EPID_NCCL(N.get_uid(&id));
    memcpy(id128, &id, sizeof(id));
    return EPID_OK;
}

int32_t epid_comm_init(epid_ctx* ctx, int32_t nranks, int32_t rank, const void* id128) {
    EPID_REQUIRE(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, EPID_ERR_INVALID, "bad argument");
    int rc = load_nccl();
    if (rc != EPID_OK) return rc;
    EPID_CUDA(cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm_t comm = nullptr;
    EPID_NCCL(N.init_rank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return EPID_OK;
}

int32_t epid_comm_destroy(epid_ctx* ctx) {
    if (ctx && ctx->nccl_comm && N.destroy) {
        cudaSetDevice(ctx->device);
        N.destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return EPID_OK;
}

int32_t epid_comm_info(const epid_ctx* ctx, int32_t* nranks, int32_t* rank) {
    EPID_REQUIRE(ctx && nranks && rank, EPID_ERR_INVALID, "NULL argument");
    *nranks = ctx->nccl_comm ? ctx->nranks : 1;
    *rank = ctx->nccl_comm ? ctx->rank : 0;
    return EPID_OK;
}

int32_t epid_gather_results(epid_ctx* ctx, const void* local, size_t bytes_per_rank, void* all) {
    EPID_REQUIRE(ctx && local && all, EPID_ERR_INVALID, "NULL argument");
    if (ctx->nranks == 1 || !ctx->nccl_comm) {
        EPID_REQUIRE(ctx->nranks == 1, EPID_ERR_NCCL, "communicator not initialised");
        memcpy(all, local, bytes_per_rank);
        return EPID_OK;
    }
    EPID_CUDA(cudaSetDevice(ctx->device));
    const size_t total = bytes_per_rank * (size_t)(ctx->nranks + 1);
    int rc = ensure_scratch(ctx, total + 256);
    if (rc != EPID_OK) return rc;
    char* d_local = (char*)ctx->scratch;
    char* d_all = d_local + (bytes_per_rank + 255) / 256 * 256;
    EPID_CUDA(cudaMemcpyAsync(d_local, local, bytes_per_rank, cudaMemcpyHostToDevice, ctx->stream));
    EPID_NCCL(N.all_gather(d_local, d_all, bytes_per_rank, 0 /* ncclInt8 / ncclChar */, ctx->nccl_comm, ctx->stream));
    EPID_CUDA(cudaMemcpyAsync(all, d_all, bytes_per_rank * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    EPID_CUDA(cudaStreamSynchronize(ctx->stream));
    return EPID_OK;
}

int32_t epid_barrier(epid_ctx* ctx) {
    EPID_REQUIRE(ctx, EPID_ERR_INVALID, "ctx is NULL");
    if (ctx->nranks == 1 || !ctx->nccl_comm) return epid_sync(ctx);
    EPID_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, 256);
    if (rc != EPID_OK) return rc;
    EPID_CUDA(cudaMemsetAsync(ctx->scratch, 0, 4, ctx->stream));
    EPID_NCCL(N.all_reduce(ctx->scratch, ctx->scratch, 1, 2 /* ncclInt32 */, 0 /* ncclSum */, ctx->nccl_comm, ctx->stream));
    EPID_CUDA(cudaStreamSynchronize(ctx->stream));
    return EPID_OK;
}

}  // extern "C"
