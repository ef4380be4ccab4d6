// comm.cu — the one collective of the path: the NCCL sum of the ranks' accumulation buffers.
//
// The reference is single-GPU (libs/DXRFramework/RtContext.cpp:23, NodeMask always 0); this is new work of the B200 design:
// one process per GPU, every GPU holds the whole BVH, the frame is split by sample index and / or by screen strips
// (rt_dispatch_rays_interleaved), and ONE reduce per output frame sums the ranks' running means onto the root — each rank's
// weight (its share of the samples) is applied INSIDE the reduction (ncclRedOpCreatePreMulSum), so there is no separate
// scaling pass over the 33-133 MB buffer.
//
// libnccl is bound at run time (dlopen), not at link time: single-GPU users of librt_core need no NCCL, and inside a
// process that already carries one (PyTorch bundles its own libnccl.so.2) the SAME library instance is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*RedOpCreatePreMulSum)(ncclRedOp_t *, void *, ncclDataType_t, ncclScalarResidence_t, ncclComm_t) = nullptr;
    ncclResult_t (*RedOpDestroy)(ncclRedOp_t, ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return RT_OK;
    const char *names[] = {getenv("RT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the instance the process already carries, if any
        if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        rt_set_error("libnccl.so.2 not found (%s): multi-GPU accumulation needs NCCL; set RT_NCCL_LIB to its path", dlerror());
        return RT_ERR_UNSUPPORTED;
    }
#define RT_SYM(field, name)                                                    \
    do {                                                                       \
        *reinterpret_cast<void **>(&g_nccl.field) = dlsym(h, name);            \
        if (!g_nccl.field) {                                                   \
            rt_set_error("libnccl lacks %s (need NCCL >= 2.11)", name);       \
            return RT_ERR_UNSUPPORTED;                                         \
        }                                                                      \
    } while (0)
    RT_SYM(GetUniqueId, "ncclGetUniqueId");
    RT_SYM(CommInitRank, "ncclCommInitRank");
    RT_SYM(CommDestroy, "ncclCommDestroy");
    RT_SYM(Reduce, "ncclReduce");
    RT_SYM(AllReduce, "ncclAllReduce");
    RT_SYM(RedOpCreatePreMulSum, "ncclRedOpCreatePreMulSum");
    RT_SYM(RedOpDestroy, "ncclRedOpDestroy");
    RT_SYM(GetErrorString, "ncclGetErrorString");
    RT_SYM(GetVersion, "ncclGetVersion");
#undef RT_SYM
    g_nccl.handle = h;
    return RT_OK;
}

#define RT_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) {                                                                        \
            rt_set_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
            return RT_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

}  // namespace

struct rt_comm {
    rt_context *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "rt_core.h RT_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");

extern "C" {

int rt_comm_get_unique_id(uint8_t *id) {
    RT_REQUIRE(id != nullptr, "id");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId u;
    RT_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return RT_OK;
}

int rt_comm_create(rt_context *ctx, const uint8_t *id, int world_size, int rank, rt_comm **out) {
    RT_REQUIRE(ctx && id && out, "null argument");
    RT_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "rank / world size");
    *out = nullptr;
    int rc = load_nccl();
    if (rc) return rc;
    RT_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    rt_comm *c = new rt_comm();
    c->ctx = ctx, c->world = world_size, c->rank = rank;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world_size, u, rank);
    if (r != ncclSuccess) {
        rt_set_error("ncclCommInitRank(world %d, rank %d) failed: %s", world_size, rank, g_nccl.GetErrorString(r));
        delete c;
        return RT_ERR_CUDA;
    }
    *out = c;
    return RT_OK;
}

int rt_comm_destroy(rt_comm *comm) {
    if (!comm) return RT_OK;
    if (comm->comm && g_nccl.CommDestroy) {
        cudaSetDevice(comm->ctx->device);
        cudaStreamSynchronize(comm->ctx->stream);
        g_nccl.CommDestroy(comm->comm);
    }
    delete comm;
    return RT_OK;
}

int rt_comm_info(const rt_comm *comm, int *world_size, int *rank, int *nccl_version) {
    RT_REQUIRE(comm != nullptr, "comm");
    if (world_size) *world_size = comm->world;
    if (rank) *rank = comm->rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
    }
    return RT_OK;
}

int rt_accum_reduce(rt_context *ctx, rt_comm *comm, const float *send, float *recv, uint64_t count, float weight, int root) {
    RT_REQUIRE(ctx && comm && comm->ctx == ctx, "context / communicator");
    RT_REQUIRE(root >= -1 && root < comm->world, "root rank (-1 = all ranks receive)");
    RT_REQUIRE(count == 0 || send != nullptr, "send buffer");
    RT_REQUIRE(count == 0 || recv != nullptr || (root >= 0 && comm->rank != root), "receive buffer on the root");
    RT_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return RT_OK;
    // sum_r weight_r * send_r: the weight is multiplied in as the data enters the reduction (no extra pass, no extra launch)
    ncclRedOp_t op;
    RT_NCCL(g_nccl.RedOpCreatePreMulSum(&op, &weight, ncclFloat32, ncclScalarHostImmediate, comm->comm));
    if (recv == nullptr) recv = const_cast<float *>(send);  // non-root ranks of a rooted reduce: NCCL never writes it there
    ncclResult_t r = root < 0 ? g_nccl.AllReduce(send, recv, count, ncclFloat32, op, comm->comm, ctx->stream)
                              : g_nccl.Reduce(send, recv, count, ncclFloat32, op, root, comm->comm, ctx->stream);
    g_nccl.RedOpDestroy(op, comm->comm);
    if (r != ncclSuccess) {
        rt_set_error("ncclReduce of %llu floats failed: %s", (unsigned long long)count, g_nccl.GetErrorString(r));
        return RT_ERR_CUDA;
    }
    return RT_OK;
}

}  // extern "C"
