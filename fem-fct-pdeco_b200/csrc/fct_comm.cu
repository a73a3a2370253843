// Multi-GPU plumbing: row-block partition with one-ring halo exchange over NCCL (ncclSend/ncclRecv between
// chain neighbours, rank r <-> r+-1) and all-reduced scalars.  NCCL is loaded at run time (dlopen of
// libnccl.so.2 -- the copy torch already mapped into the process when there is one) so that single-GPU use
// has no NCCL dependency at all.
//
// Local numbering: a rank holds the contiguous global row range [G0,G1) (owned rows [R0,R1) plus the rows its
// owned rows reference); local index = global - G0, so halo ranges are contiguous: [0,row_begin) comes from
// rank-1 and [row_end,n) from rank+1.  Nothing here has a counterpart in the reference (single process).
#include "fct_common.cuh"
#include "../../include/fctpdeco.h"

#include <dlfcn.h>

// The handful of NCCL 2.x declarations this file needs, stated locally (ABI-stable since NCCL 2.0) so that the library
// builds on machines without the NCCL development headers; the functions themselves are resolved with dlsym below.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6,
               ncclFloat32 = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    FCT_CHECK(h, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define LOAD(sym, name)                                                                  \
    *(void**)(&g_nccl.sym) = dlsym(h, name);                                             \
    FCT_CHECK(g_nccl.sym, "NCCL symbol %s not found", name)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.handle = h;
    return 0;
}

#define FCT_NCCL(call)                                                                                  \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            fct_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

struct fct_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

extern "C" int fct_nccl_unique_id(void* id_out) {
    FCT_CHECK(id_out, "fct_nccl_unique_id: null argument");
    if (nccl_load()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    FCT_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" int fct_ctx_init_comm(fct_ctx* ctx, const void* id_bytes, int32_t rank, int32_t world, int32_t slb,
                                 int32_t sle, int32_t shb, int32_t she) {
    FCT_CHECK(ctx && id_bytes, "fct_ctx_init_comm: null argument");
    FCT_CHECK(world >= 1 && rank >= 0 && rank < world, "fct_ctx_init_comm: bad rank/world");
    FCT_CHECK(!ctx->comm, "fct_ctx_init_comm: communicator already initialised");
    FCT_CHECK(0 <= slb && slb <= sle && sle <= ctx->n && 0 <= shb && shb <= she && she <= ctx->n,
              "fct_ctx_init_comm: bad send ranges");
    if (nccl_load()) return 1;
    FCT_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    fct_comm* c = new fct_comm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        fct_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return 1;
    }
    ctx->comm = c;
    ctx->send_lo[0] = slb; ctx->send_lo[1] = sle;
    ctx->send_hi[0] = shb; ctx->send_hi[1] = she;
    return 0;
}

void fct_comm_destroy(fct_ctx* ctx) {
    if (ctx->comm) {
        if (ctx->comm->comm) g_nccl.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}

// exchange element ranges of a device array with the chain neighbours:
//   to rank-1: [slo0,slo1)   from rank-1: [rlo0,rlo1)   to rank+1: [shi0,shi1)   from rank+1: [rhi0,rhi1)
static int exchange_ranges(fct_ctx* ctx, double* a, int64_t slo0, int64_t slo1, int64_t rlo0, int64_t rlo1,
                           int64_t shi0, int64_t shi1, int64_t rhi0, int64_t rhi1) {
    fct_comm* c = ctx->comm;
    FCT_NCCL(g_nccl.GroupStart());
    if (c->rank > 0) {
        if (slo1 > slo0) FCT_NCCL(g_nccl.Send(a + slo0, (size_t)(slo1 - slo0), ncclDouble, c->rank - 1, c->comm, ctx->stream));
        if (rlo1 > rlo0) FCT_NCCL(g_nccl.Recv(a + rlo0, (size_t)(rlo1 - rlo0), ncclDouble, c->rank - 1, c->comm, ctx->stream));
    }
    if (c->rank + 1 < c->world) {
        if (shi1 > shi0) FCT_NCCL(g_nccl.Send(a + shi0, (size_t)(shi1 - shi0), ncclDouble, c->rank + 1, c->comm, ctx->stream));
        if (rhi1 > rhi0) FCT_NCCL(g_nccl.Recv(a + rhi0, (size_t)(rhi1 - rhi0), ncclDouble, c->rank + 1, c->comm, ctx->stream));
    }
    FCT_NCCL(g_nccl.GroupEnd());
    return 0;
}

bool fct_p2p_ready(const fct_ctx* ctx);
int fct_p2p_exchange(fct_ctx* ctx, double* v0, double* v1, const unsigned long long* cond);

int fct_halo_exchange_if(fct_ctx* ctx, double* vec) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    if (!ctx->capturing) ctx->exchanges++;
    if (fct_p2p_ready(ctx)) return fct_p2p_exchange(ctx, vec, nullptr, nullptr);
    return exchange_ranges(ctx, vec, ctx->send_lo[0], ctx->send_lo[1], 0, ctx->row_begin, ctx->send_hi[0],
                           ctx->send_hi[1], ctx->row_end, ctx->n);
}

// exchange only if the device word *cond is nonzero (peer mailboxes only; *cond must be the same on every rank)
int fct_halo_exchange_cond(fct_ctx* ctx, double* vec, const unsigned long long* cond) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    FCT_CHECK(fct_p2p_ready(ctx), "fct_halo_exchange_cond: needs the peer mailboxes");
    return fct_p2p_exchange(ctx, vec, nullptr, cond);
}

extern "C" int fct_halo_exchange(fct_ctx* ctx, double* vec) {
    FCT_CHECK(ctx && vec, "fct_halo_exchange: null argument");
    return fct_halo_exchange_if(ctx, vec);
}

// max-allreduce of two nonnegative doubles stored as bit patterns (ordering of nonnegative IEEE doubles equals
// the ordering of their bit patterns, so an integer max is exact)
int fct_halo_allreduce_max2(fct_ctx* ctx, unsigned long long* two_words) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    FCT_NCCL(g_nccl.AllReduce(two_words, two_words, 2, ncclUint64, ncclMax, ctx->comm->comm, ctx->stream));
    return 0;
}

int fct_allreduce_sum_host(fct_ctx* ctx, double* v) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    FCT_CUDA(cudaMemcpyAsync(ctx->red + 8, v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FCT_NCCL(g_nccl.AllReduce(ctx->red + 8, ctx->red + 8, 1, ncclDouble, ncclSum, ctx->comm->comm, ctx->stream));
    FCT_CUDA(cudaMemcpyAsync(v, ctx->red + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fct_allreduce_sum_dev(fct_ctx* ctx, double* dev, int count) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    FCT_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, ctx->comm->comm, ctx->stream));
    return 0;
}

// two vectors in one message (R+ and R-)
int fct_halo_exchange2_if(fct_ctx* ctx, double* v0, double* v1) {
    if (!ctx->comm || ctx->comm->world == 1) return 0;
    if (fct_p2p_ready(ctx)) return fct_p2p_exchange(ctx, v0, v1, nullptr);
    if (fct_halo_exchange_if(ctx, v0)) return 1;
    return v1 ? fct_halo_exchange_if(ctx, v1) : 0;
}
