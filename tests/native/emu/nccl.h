// TEST INFRASTRUCTURE: just enough of <nccl.h> for comm.cuh to compile in the -DGM_CPU_EMU build (no collective is ever called
// there: the emulated multi-rank step uses the peer-memory gather, whose "peers" are blocks in host memory).
#pragma once
typedef enum { ncclSuccess = 0, ncclInternalError = 3 } ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
#define NCCL_UNIQUE_ID_BYTES 128
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclUint32 = 3, ncclUint64 = 5 } ncclDataType_t;
