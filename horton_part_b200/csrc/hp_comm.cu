// Multi-GPU plumbing behind the C ABI: thin wrappers over NCCL so that a consumer of
// libhp_b200.so alone -- the ctypes route of INTEGRATION.md section 2, without the Python classes and without
// torch.distributed -- can run the sharded stockholder iteration: each rank evaluates the promolecule of ALL atoms
// on its own atom blocks, solves its own atoms, and ONE sum all-reduce of the zero-filled state vector
// [entropy | msd | charges | propars] per iteration acts as the all-gather (x + 0 = x exactly).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch has already loaded when the classes are
// in use, the system one otherwise), so the library itself has no link-time dependency on it and single-GPU
// users never touch it.  Only the handful of types below are needed from nccl.h; they are restated to keep
// the build header-free (NCCL >= 2.x ABI: ncclUniqueId = 128 opaque bytes, ncclFloat64 = 8, ncclSum = 0).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "hp_common.cuh"

namespace {

struct NcclId { char internal[128]; };
using NcclComm = void*;
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_once;

int load_nccl() {
    std::call_once(g_once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            g_nccl.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
        }
        if (!g_nccl.handle) return;
        auto sym = [&](const char* n) { return dlsym(g_nccl.handle, n); };
        g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
        g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
        g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
        g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
        g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
        g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
        g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!g_nccl.handle || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce ||
        !g_nccl.AllGather) {
        hp::set_error("hp_comm: libnccl.so.2 could not be loaded (%s)", g_nccl.handle ? "missing symbols" : dlerror());
        return HP_ERR_CUDA;
    }
    return HP_OK;
}

int check_nccl(int rc, const char* what) {
    if (rc == 0) return HP_OK;
    hp::set_error("%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return HP_ERR_CUDA;
}

}  // namespace

using namespace hp;

extern "C" int32_t hp_comm_nccl_version(void) {
    int v = 0;
    if (load_nccl() != HP_OK || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return 0;
    return v;
}

extern "C" int hp_comm_unique_id(void* id128_host) {
    HP_REQUIRE(id128_host, "null output");
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    rc = check_nccl(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    if (rc) return rc;
    memcpy(id128_host, &id, sizeof(id));
    return HP_OK;
}

extern "C" int hp_comm_init(int32_t world, int32_t rank, const void* id128_host, void** comm_out) {
    HP_REQUIRE(world > 0 && rank >= 0 && rank < world && id128_host && comm_out, "bad arguments");
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    memcpy(&id, id128_host, sizeof(id));
    NcclComm comm = nullptr;
    rc = check_nccl(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    if (rc) return rc;
    *comm_out = comm;
    return HP_OK;
}

extern "C" int hp_comm_allreduce(void* comm, double* buf, int64_t count, int32_t op_max, void* stream) {
    HP_REQUIRE(comm && buf && count >= 0, "bad arguments");
    if (count == 0) return HP_OK;
    return check_nccl(g_nccl.AllReduce(buf, buf, size_t(count), kNcclFloat64, op_max ? kNcclMax : kNcclSum, comm,
                                       as_stream(stream)), "ncclAllReduce");
}

extern "C" int hp_comm_allgather(void* comm, const double* send, double* recv, int64_t count_per_rank, void* stream) {
    HP_REQUIRE(comm && send && recv && count_per_rank >= 0, "bad arguments");
    if (count_per_rank == 0) return HP_OK;
    return check_nccl(g_nccl.AllGather(send, recv, size_t(count_per_rank), kNcclFloat64, comm, as_stream(stream)),
                      "ncclAllGather");
}

extern "C" int hp_comm_destroy(void* comm) {
    if (!comm) return HP_OK;
    return check_nccl(g_nccl.CommDestroy(comm), "ncclCommDestroy");
}
