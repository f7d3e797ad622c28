// Multi-GPU exchange step: an NCCL communicator per context, loaded at run time.
//
// The library itself has no link-time dependency on NCCL (a single-GPU host needs none): libnccl.so.2
// is dlopen()ed on the first dfb_comm_* call.  A multi-process host (one process per GPU, e.g. Rscript
// workers or torchrun ranks) creates the id on rank 0 with dfb_comm_unique_id, ships its 128 bytes to
// the other ranks by whatever means it has (a file, MPI, a socket, torch.distributed) and every rank
// calls dfb_comm_init_rank; a single process driving all GPUs of the box calls dfb_comm_init_all.
// The only collectives of the pipeline live in dfb_merge_results (pvalues.cu): the genome-wide pooling
// of mergeResults() (R/mergeResults.R:216-235, 300-307, 380-386).
#include <dlfcn.h>
#include <nccl.h>

#include "data.cuh"

namespace dfb {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // an already loaded copy (e.g. the one PyTorch brought) is found first by its soname
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            auto sym = [&](const char* s) { return dlsym(api.handle, s); };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather ||
                !api.AllReduce || !api.GroupStart || !api.GroupEnd || !api.GetErrorString) {
                dlclose(api.handle);
                api.handle = nullptr;
            }
        }
    }
    return api.handle ? &api : nullptr;
}

#define DFB_NCCL(api, expr)                                                                              \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            dfb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r));     \
            return DFB_ECUDA;                                                                            \
        }                                                                                                \
    } while (0)

int comm_all_gather(dfb_ctx* ctx, const void* send, void* recv, size_t count, int dtype_bytes) {
    if (ctx->world <= 1) {
        if (send != recv) DFB_CUDA(cudaMemcpyAsync(recv, send, count * dtype_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return DFB_OK;
    }
    NcclApi* api = nccl_api();
    DFB_REQUIRE(api && ctx->nccl_comm, "no communicator: call dfb_comm_init_rank / dfb_comm_init_all first");
    DFB_NCCL(api, api->AllGather(send, recv, count * dtype_bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DFB_OK;
}

int comm_all_reduce_sum_u64(dfb_ctx* ctx, void* buf, size_t count) {
    if (ctx->world <= 1) return DFB_OK;
    NcclApi* api = nccl_api();
    DFB_REQUIRE(api && ctx->nccl_comm, "no communicator: call dfb_comm_init_rank / dfb_comm_init_all first");
    DFB_NCCL(api, api->AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_comm_unique_id(void* id_out) {
    DFB_REQUIRE(id_out, "dfb_comm_unique_id: NULL output");
    NcclApi* api = nccl_api();
    DFB_REQUIRE(api, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    static_assert(sizeof(ncclUniqueId) == DFB_COMM_ID_BYTES, "NCCL unique id size");
    ncclUniqueId id;
    DFB_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return DFB_OK;
}

int dfb_comm_init_rank(dfb_ctx* ctx, int world, int rank, const void* id) {
    DFB_REQUIRE(ctx && id && world >= 1 && rank >= 0 && rank < world, "dfb_comm_init_rank: bad argument");
    NcclApi* api = nccl_api();
    DFB_REQUIRE(api, "libnccl.so.2 could not be loaded");
    dfb_comm_destroy(ctx);
    DFB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId nid;
    memcpy(&nid, id, sizeof nid);
    ncclComm_t comm = nullptr;
    DFB_NCCL(api, api->CommInitRank(&comm, world, nid, rank));
    ctx->nccl_comm = comm;
    ctx->world = world;
    ctx->rank = rank;
    return DFB_OK;
}

int dfb_comm_init_all(dfb_ctx* const* ctxs, int ngpu) {
    DFB_REQUIRE(ctxs && ngpu >= 1 && ngpu <= 64, "dfb_comm_init_all: bad argument");
    NcclApi* api = nccl_api();
    DFB_REQUIRE(api, "libnccl.so.2 could not be loaded");
    std::vector<int> devs((size_t)ngpu);
    std::vector<ncclComm_t> comms((size_t)ngpu, nullptr);
    for (int i = 0; i < ngpu; ++i) {
        DFB_REQUIRE(ctxs[i] != nullptr, "dfb_comm_init_all: NULL context %d", i);
        dfb_comm_destroy(ctxs[i]);
        devs[(size_t)i] = ctxs[i]->device;
    }
    DFB_NCCL(api, api->CommInitAll(comms.data(), ngpu, devs.data()));
    for (int i = 0; i < ngpu; ++i) {
        ctxs[i]->nccl_comm = comms[(size_t)i];
        ctxs[i]->world = ngpu;
        ctxs[i]->rank = i;
    }
    return DFB_OK;
}

void dfb_comm_destroy(dfb_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    NcclApi* api = nccl_api();
    if (api) {
        cudaStreamSynchronize(ctx->stream);
        api->CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
}

int dfb_comm_info(const dfb_ctx* ctx, int* world, int* rank) {
    DFB_REQUIRE(ctx, "dfb_comm_info: NULL context");
    if (world) *world = ctx->world;
    if (rank) *rank = ctx->rank;
    return DFB_OK;
}

}  // extern "C"
