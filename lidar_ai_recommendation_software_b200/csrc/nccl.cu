// lidar_nccl_* — NCCL plumbing behind the C ABI (include/lidar_b200.h).
//
// The two collectives of the point-sharded path: MAX over the packed bounding box and SUM over
// the int32 density grid.  The reference has no communication code at all; these entry points exist so that a
// caller without torch.distributed can run the three-enqueue form of the sharded density
// (lidar_scan_bbox_packed -> MAX -> lidar_scan_hist -> SUM -> lidar_scan_finish), and so that the fused NVLink
// kernel (scan.cu) has a baseline to be checked against.
//
// libnccl.so.2 is resolved at run time: in a PyTorch process that is the NCCL torch already loaded (same
// soname), otherwise the system library.  The few declarations needed are restated here so that the core builds
// without nccl.h.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace lidar {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
// values of ncclDataType_t / ncclRedOp_t (nccl.h, stable since NCCL 2.0)
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    });
    return api;
}

#define LIDAR_NCCL_TRY(expr)                                                                       \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != 0) {                                                                             \
            ::lidar::set_error("NCCL error %d (%s) in `%s`", _r, nccl_api().GetErrorString(_r), #expr); \
            return LIDAR_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

}  // namespace lidar

using namespace lidar;

extern "C" {

int lidar_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int lidar_nccl_unique_id(void* h_id128) {
    LIDAR_REQUIRE(h_id128 != nullptr, LIDAR_ERR_INVALID, "lidar_nccl_unique_id: NULL id");
    LIDAR_REQUIRE(nccl_api().ok, LIDAR_ERR_INVALID, "lidar_nccl_unique_id: libnccl.so.2 not found");
    ncclUniqueId id;
    LIDAR_NCCL_TRY(nccl_api().GetUniqueId(&id));
    memcpy(h_id128, id.internal, sizeof(id.internal));
    return LIDAR_OK;
}

int lidar_nccl_comm_init(const void* h_id128, int rank, int world, void** comm_out) {
    LIDAR_REQUIRE(h_id128 && comm_out && world >= 1 && rank >= 0 && rank < world, LIDAR_ERR_INVALID,
                  "lidar_nccl_comm_init: bad argument (rank %d of %d)", rank, world);
    LIDAR_REQUIRE(nccl_api().ok, LIDAR_ERR_INVALID, "lidar_nccl_comm_init: libnccl.so.2 not found");
    ncclUniqueId id;
    memcpy(id.internal, h_id128, sizeof(id.internal));
    ncclComm_t comm = nullptr;
    LIDAR_NCCL_TRY(nccl_api().CommInitRank(&comm, world, id, rank));
    *comm_out = comm;
    return LIDAR_OK;
}

int lidar_nccl_comm_destroy(void* comm) {
    if (!comm) return LIDAR_OK;
    LIDAR_REQUIRE(nccl_api().ok, LIDAR_ERR_INVALID, "lidar_nccl_comm_destroy: libnccl.so.2 not found");
    LIDAR_NCCL_TRY(nccl_api().CommDestroy(static_cast<ncclComm_t>(comm)));
    return LIDAR_OK;
}

int lidar_nccl_allreduce(void* comm, void* d_buf, int64_t count, int op, void* stream) {
    LIDAR_REQUIRE(comm && d_buf && count >= 0, LIDAR_ERR_INVALID, "lidar_nccl_allreduce: bad argument");
    LIDAR_REQUIRE(op == LIDAR_NCCL_SUM_I32 || op == LIDAR_NCCL_MAX_F64, LIDAR_ERR_INVALID, "lidar_nccl_allreduce: unknown op %d", op);
    LIDAR_REQUIRE(nccl_api().ok, LIDAR_ERR_INVALID, "lidar_nccl_allreduce: libnccl.so.2 not found");
    if (count == 0) return LIDAR_OK;
    const int dt = op == LIDAR_NCCL_SUM_I32 ? kNcclInt32 : kNcclFloat64;
    const int ro = op == LIDAR_NCCL_SUM_I32 ? kNcclSum : kNcclMax;
    LIDAR_NCCL_TRY(nccl_api().AllReduce(d_buf, d_buf, (size_t)count, dt, ro, static_cast<ncclComm_t>(comm), as_stream(stream)));
    return LIDAR_OK;
}

}  // extern "C"
