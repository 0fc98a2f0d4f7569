// tests/cuda_emu/nccl.h -- TEST INFRASTRUCTURE ONLY: the NCCL types pft_b200.cu names, for the
// emulated build.  The functions themselves are looked up with dlopen (PFT_NCCL_LIB), so the
// multi-rank CPU tests point that at libnccl_emu.so (emu_nccl.cpp), which routes the calls to
// hooks that the test registers (torch.distributed gloo).
#pragma once
#include <cuda_runtime.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclInt32 = 2, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
