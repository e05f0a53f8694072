// Cross-rank exchange for the batch-sharded int8 path: when a batch is split over GPUs, every rank must
// quantise with the min / max of the WHOLE tensor to stay bit-identical to the unsharded reference
// (src/ops/quantize.rs:352-434 computes one range per tensor).  That is an all-reduce of two numbers per
// DynamicQuantizeLinear.  NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy already loaded in the process if
// there is one), so the library itself keeps no link-time dependency on it.
#include <dlfcn.h>

#include "common.h"

struct Id128 {  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
    char bytes[128];
};

namespace {

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return false;
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) return false;
    g_nccl = a;
    return true;
}

constexpr int kNcclInt32 = 2;  // ncclInt32
constexpr int kNcclMax = 2;    // ncclMax
constexpr int kNcclMin = 3;    // ncclMin

}  // namespace

struct rten_comm {
    void* nccl = nullptr;
    int rank = 0, world = 1;
};

namespace rtb {

// min / max all-reduce of the two ORDERED-INT encoded floats produced by the local min/max kernel (rowops.cu): integer
// min / max on that encoding is the float min / max, and is exact and order independent.
rten_status comm_allreduce_minmax(rten_ctx* ctx, rten_comm* comm, int* mm) {
    if (!comm || comm->world <= 1) return RTEN_OK;
    cudaStream_t s = launch_stream(ctx);
    int r = g_nccl.AllReduce(mm, mm, 1, kNcclInt32, kNcclMin, comm->nccl, s);
    if (r == 0) r = g_nccl.AllReduce(mm + 1, mm + 1, 1, kNcclInt32, kNcclMax, comm->nccl, s);
    if (r != 0) {
        ctx->err = std::string("ncclAllReduce failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        return RTEN_ERR_NCCL;
    }
    count_launch(ctx, 2);
    return RTEN_OK;
}

}  // namespace rtb

extern "C" {

rten_status rten_b200_comm_unique_id(void* id128) {
    if (!id128) return RTEN_ERR_INVALID_VALUE;
    if (!load_nccl()) return RTEN_ERR_NCCL;
    return g_nccl.GetUniqueId(id128) == 0 ? RTEN_OK : RTEN_ERR_NCCL;
}

rten_status rten_b200_comm_create(rten_ctx* ctx, const void* id128, int rank, int world, rten_comm** out) {
    if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return RTEN_ERR_INVALID_VALUE;
    *out = nullptr;
    if (!load_nccl()) return rtb::fail(ctx, RTEN_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(ctx->device);
    Id128 id;
    memcpy(id.bytes, id128, sizeof(id.bytes));
    rten_comm* c = new rten_comm();
    c->rank = rank;
    c->world = world;
    const int r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
    if (r != 0) {
        ctx->err = std::string("ncclCommInitRank failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        delete c;
        return RTEN_ERR_NCCL;
    }
    *out = c;
    return RTEN_OK;
}

void rten_b200_comm_destroy(rten_comm* comm) {
    if (!comm) return;
    if (comm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->nccl);
    delete comm;
}

}  // extern "C"
