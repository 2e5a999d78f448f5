// pn2_nccl.cuh -- NCCL bound at first use with dlopen, not at link time: a process that also imports torch must end
// up with ONE libnccl.so.2 (torch's bundled 2.28 needs symbols the system 2.27 lacks).  Order: the copy already
// mapped into the process, $PN2_NCCL_LIB (pn2gpu.py points it at torch's bundled library), the system library.
#pragma once
#include <nccl.h>
#include "pn2_common.cuh"

struct Pn2NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok = false;
};
extern Pn2NcclApi g_nccl;
bool nccl_load();          // defined in pn2_let.cu

#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define NCCL_TRY(expr)                                                                          \
    do {                                                                                        \
        ncclResult_t r_ = (expr);                                                               \
        if (r_ != ncclSuccess) {                                                                \
            pn2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(r_)); \
            return PN2_ERR_NCCL;                                                                \
        }                                                                                       \
    } while (0)
