// comm.h — the one collective of the multi-GPU path: the accumulator gather (BASELINE.json north_star: "tiles
// merged over NCCL on NVLink only for the final accumulation gather").  NCCL's C API, bound at run time: librfwb200.so has no
// link-time dependency on libnccl (a single-GPU host needs none), and a host process that already carries an NCCL — PyTorch
// bundles its own libnccl.so.2 — shares that copy instead of loading a second one next to it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace rfw {

static constexpr int COMM_UNIQUE_ID_BYTES = 128;  // sizeof(ncclUniqueId)

struct Comm {
    void* nccl_comm = nullptr;  // ncclComm_t
    uint32_t rank = 0, world = 1;
    bool warmed = false;        // a gather has completed: the peer-to-peer connections exist (NCCL sets them up inside the first call, blocking)
    bool active() const { return nccl_comm != nullptr; }
};

// every function returns an empty string on success, else a message
std::string comm_unique_id(uint8_t out[COMM_UNIQUE_ID_BYTES]);
std::string comm_init(Comm& c, const uint8_t id[COMM_UNIQUE_ID_BYTES], uint32_t rank, uint32_t world);
void comm_destroy(Comm& c);
// every rank sends `count` floats from d_send; with root < world only the root receives (d_recv: world * count floats, rank
// r's block at r * count; the root's own block is copied device-to-device), with root >= world every rank receives (all-gather)
std::string comm_gather(Comm& c, const float* d_send, float* d_recv, size_t count, uint32_t root, cudaStream_t stream);
std::string comm_version(int* out_version);
// a peer died or never called: tears the communicator down without waiting for the outstanding collective (ncclCommAbort)
void comm_abort(Comm& c);

}  // namespace rfw
