// NCCL access without a link-time dependency: libt4b.so resolves the handful of NCCL entry points it needs with
// dlopen at first use (preferring the copy already loaded into the process, e.g. the one bundled with the host's
// torch), so single-GPU users never need NCCL at all.  Types come from nccl.h (header only).
#pragma once
#include <nccl.h>

#include "../common.h"

namespace t4b {
namespace nccl {

struct Api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int*);
};
// Throws Error(ST_UNSUPPORTED) when no libnccl.so.2 can be loaded.
const Api& api();
void check(ncclResult_t r, const char* what);

}  // namespace nccl
}  // namespace t4b
