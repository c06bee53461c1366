// NCCL, loaded at run time.  libbess_b200.so has no link-time dependency on NCCL: single-GPU users never touch it, and
// inside a torch process dlopen("libnccl.so.2") resolves to the copy torch already mapped, so there is exactly one
// NCCL in the process.  Declarations come from <nccl.h> (compile time only).
#pragma once
#include <nccl.h>

#include <string>

namespace bess {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

// throws EngineError when libnccl.so.2 cannot be loaded
const NcclApi &nccl_api();

}  // namespace bess
