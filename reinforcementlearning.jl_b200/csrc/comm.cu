// comm.cu — multi-GPU exchange: env-index data parallelism needs exactly one sum all-reduce of
// the flat gradient (np fp32, ~36 KB) per optimiser step plus two doubles for the global
// advantage normalisation.  One process per GPU; the communicator is NCCL over
// NVLink 5 / NVSwitch, resolved at run time with dlopen("libnccl.so.2") so the library has no
// link-time dependency and shares the NCCL already loaded by the host process (e.g. torch's).
#include <dlfcn.h>

#include "common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
typedef ncclResult_t (*fn_GetUniqueId)(ncclUniqueId*);
typedef ncclResult_t (*fn_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_CommDestroy)(ncclComm_t);
typedef const char* (*fn_GetErrorString)(ncclResult_t);

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
};
NcclApi g_api;

int load_nccl() {
    if (g_api.lib) return B200RL_OK;
    const char* override_path = getenv("B200RL_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        b200rl_set_error("b200rl_comm: cannot dlopen libnccl.so.2 (%s); set B200RL_NCCL_LIB", dlerror());
        return B200RL_ERR_NCCL;
    }
    g_api.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_api.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_api.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_api.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_api.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.CommDestroy) {
        b200rl_set_error("b200rl_comm: libnccl is missing required symbols");
        return B200RL_ERR_NCCL;
    }
    g_api.lib = h;
    return B200RL_OK;
}
#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != 0) {                                                                                   \
            b200rl_set_error("%s -> NCCL error %d (%s)", #expr, _r, g_api.GetErrorString ? g_api.GetErrorString(_r) : "?"); \
            return B200RL_ERR_NCCL;                                                                      \
        }                                                                                                \
    } while (0)
}  // namespace

struct b200rl_comm_state {
    ncclComm_t comm;
    int nranks, rank;
};

void b200rl_comm_destroy_internal(b200rl_ctx* ctx) {
    if (ctx->comm) {
        if (g_api.CommDestroy) g_api.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}
int b200rl_comm_world(b200rl_ctx* ctx) { return ctx->comm ? ctx->comm->nranks : 1; }
int b200rl_comm_allreduce_internal(b200rl_ctx* ctx, void* buf, int64_t n, int is_double) {
    REQUIRE(ctx->comm, B200RL_ERR_INVALID, "no communicator");
    NCCL_TRY(g_api.AllReduce(buf, buf, (size_t)n, is_double ? ncclFloat64 : ncclFloat32, ncclSum, ctx->comm->comm, ctx->stream));
    return B200RL_OK;
}

extern "C" {
/* rank 0 creates the 128-byte NCCL unique id and ships it to the other ranks out of band */
int b200rl_comm_unique_id(void* id128_out) {
    REQUIRE(id128_out, B200RL_ERR_INVALID, "null out");
    TRY(load_nccl());
    ncclUniqueId id;
    NCCL_TRY(g_api.GetUniqueId(&id));
    memcpy(id128_out, &id, sizeof id);
    return B200RL_OK;
}
/* one process per GPU: attach ctx to rank `rank` of an `nranks` communicator */
int b200rl_comm_init(b200rl_ctx* ctx, int nranks, int rank, const void* id128) {
    TRY(ctx_bind(ctx));
    REQUIRE(id128 && nranks >= 1 && rank >= 0 && rank < nranks, B200RL_ERR_INVALID, "bad argument");
    REQUIRE(!ctx->comm, B200RL_ERR_INVALID, "communicator already initialised");
    TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c;
    NCCL_TRY(g_api.CommInitRank(&c, nranks, id, rank));
    ctx->comm = new b200rl_comm_state{c, nranks, rank};
    return B200RL_OK;
}
/* in-place sum all-reduce of a DEVICE fp32 buffer on the ctx stream */
int b200rl_comm_allreduce_f32(b200rl_ctx* ctx, float* dev_buf, int64_t n) {
    TRY(ctx_bind(ctx));
    REQUIRE(dev_buf && n > 0, B200RL_ERR_INVALID, "bad argument");
    return b200rl_comm_allreduce_internal(ctx, dev_buf, n, 0);
}
}
