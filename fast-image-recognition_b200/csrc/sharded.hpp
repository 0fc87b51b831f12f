// sharded.hpp — what the multi-GPU translation units share: the run-time NCCL binding and the communicator handle.
#pragma once
#include "fir_common.cuh"
#include <nccl.h>      // types and enums only; every call goes through the table below (dlopen of libnccl.so.2)

namespace fir {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    std::string error;
};

NcclApi* nccl_api();      // nullptr when no NCCL library can be loaded

#define FIR_NCCL_TRY(expr)                                                                                       \
    do {                                                                                                         \
        ncclResult_t _r = (expr);                                                                                \
        if (_r != ncclSuccess) return ::fir::fail(FIR_ERR_NCCL, std::string(#expr) + ": " + nccl_api()->GetErrorString(_r)); \
    } while (0)

}  // namespace fir

// one rank of the communicator + its grow-only staging buffers (kept apart from the gallery's bump workspace, which every
// search call resets)
struct fir_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    bool owns_comm = true;
    enum { SLOTS = 24 };
    void* buf[SLOTS] = {};
    size_t cap[SLOTS] = {};
};


namespace fir {
int comm_take(fir_comm* c, int slot, size_t bytes, cudaStream_t s, void** out);   // grow-only staging buffer `slot` of the rank
}
