// comm.cu — NCCL binding for the accumulator gather (see comm.h).
#include "comm.h"

#include <dlfcn.h>
#include <string.h>

#include <mutex>

namespace rfw {

// the part of nccl.h this file needs (NCCL's C ABI: opaque communicator, 128-byte unique id passed by value, int enums)
struct NcclUniqueId { char internal[COMM_UNIQUE_ID_BYTES]; };
typedef int nccl_result_t;  // ncclSuccess = 0
enum { NCCL_FLOAT32 = 7 };  // ncclFloat32 / ncclFloat
struct NcclApi {
    void* lib = nullptr;
    nccl_result_t (*GetVersion)(int*) = nullptr;
    nccl_result_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
    nccl_result_t (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    nccl_result_t (*CommDestroy)(void*) = nullptr;
    nccl_result_t (*CommAbort)(void*) = nullptr;
    nccl_result_t (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(nccl_result_t) = nullptr;
    std::string error;
};

static NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // by soname first: if the process has an NCCL loaded already (PyTorch), this returns that very copy
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) { a.error = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
        auto sym = [&](const char* n) -> void* {
            void* p = dlsym(a.lib, n);
            if (!p && a.error.empty()) a.error = std::string("NCCL symbol missing: ") + n;
            return p;
        };
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.CommAbort = reinterpret_cast<decltype(a.CommAbort)>(sym("ncclCommAbort"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return a;
}

static std::string nccl_err(const char* what, nccl_result_t r) {
    NcclApi& a = api();
    return std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error") + " (" + std::to_string(r) + ")";
}
#define NCCL_CK(call, what)                     \
    do {                                        \
        const nccl_result_t r_ = (call);        \
        if (r_ != 0) return nccl_err(what, r_); \
    } while (0)

std::string comm_version(int* out_version) {
    NcclApi& a = api();
    if (!a.error.empty()) return a.error;
    NCCL_CK(a.GetVersion(out_version), "ncclGetVersion");
    return "";
}

std::string comm_unique_id(uint8_t out[COMM_UNIQUE_ID_BYTES]) {
    NcclApi& a = api();
    if (!a.error.empty()) return a.error;
    NcclUniqueId id;
    NCCL_CK(a.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, COMM_UNIQUE_ID_BYTES);
    return "";
}

std::string comm_init(Comm& c, const uint8_t idb[COMM_UNIQUE_ID_BYTES], uint32_t rank, uint32_t world) {
    NcclApi& a = api();
    if (!a.error.empty()) return a.error;
    if (world == 0 || rank >= world) return "comm_init: rank out of range";
    comm_destroy(c);
    NcclUniqueId id;
    memcpy(id.internal, idb, COMM_UNIQUE_ID_BYTES);
    void* comm = nullptr;
    NCCL_CK(a.CommInitRank(&comm, (int)world, id, (int)rank), "ncclCommInitRank");
    c.nccl_comm = comm; c.rank = rank; c.world = world; c.warmed = false;
    return "";
}

void comm_destroy(Comm& c) {
    if (c.nccl_comm && api().CommDestroy) api().CommDestroy(c.nccl_comm);
    c.nccl_comm = nullptr; c.rank = 0; c.world = 1; c.warmed = false;
}

void comm_abort(Comm& c) {
    if (c.nccl_comm && api().CommAbort) api().CommAbort(c.nccl_comm);
    c.nccl_comm = nullptr; c.rank = 0; c.world = 1; c.warmed = false;
}

std::string comm_gather(Comm& c, const float* d_send, float* d_recv, size_t count, uint32_t root, cudaStream_t stream) {
    NcclApi& a = api();
    if (!c.active()) return "comm_gather: no communicator (rfwb200_comm_init)";
    if (root >= c.world) {
        NCCL_CK(a.AllGather(d_send, d_recv, count, NCCL_FLOAT32, c.nccl_comm, stream), "ncclAllGather");
        return "";
    }
    // gather to one rank: every other rank sends its block, the root posts one receive per peer (one fused group: the
    // transfers run concurrently over NVLink / NVSwitch) and copies its own block locally
    NCCL_CK(a.GroupStart(), "ncclGroupStart");
    if (c.rank == root) {
        for (uint32_t r = 0; r < c.world; r++) {
            if (r == root) continue;
            const nccl_result_t rr = a.Recv(d_recv + (size_t)r * count, count, NCCL_FLOAT32, (int)r, c.nccl_comm, stream);
            if (rr != 0) { a.GroupEnd(); return nccl_err("ncclRecv", rr); }
        }
    } else {
        const nccl_result_t rs = a.Send(d_send, count, NCCL_FLOAT32, (int)root, c.nccl_comm, stream);
        if (rs != 0) { a.GroupEnd(); return nccl_err("ncclSend", rs); }
    }
    NCCL_CK(a.GroupEnd(), "ncclGroupEnd");
    if (c.rank == root && cudaMemcpyAsync(d_recv + (size_t)root * count, d_send, count * sizeof(float), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return std::string("comm_gather: local copy: ") + cudaGetErrorString(cudaGetLastError());
    return "";
}

}  // namespace rfw
