// tests/cuda_emu/emu_nccl.cpp -- TEST INFRASTRUCTURE ONLY: a libnccl stand-in for the CPU
// multi-rank tests.  Point PFT_NCCL_LIB at the built libnccl_emu.so; the test registers hooks
// (ctypes callbacks around torch.distributed/gloo) that move the bytes.
#include "nccl.h"
#include <stdio.h>

extern "C" {
typedef void (*emu_p2p_fn)(void *buf, size_t bytes, int peer, int is_send);
typedef void (*emu_group_fn)(int end);
typedef void (*emu_allreduce_max_fn)(double *buf, size_t count);
static emu_p2p_fn h_p2p; static emu_group_fn h_group; static emu_allreduce_max_fn h_allmax;
void emu_nccl_set_hooks(emu_p2p_fn p2p, emu_group_fn group, emu_allreduce_max_fn allmax) { h_p2p = p2p; h_group = group; h_allmax = allmax; }

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) { for (int i = 0; i < 128; i++) id->internal[i] = (char)i; return ncclSuccess; }
ncclResult_t ncclCommInitRank(ncclComm_t *c, int, ncclUniqueId, int) { *c = (ncclComm_t)1; return h_p2p ? ncclSuccess : ncclInternalError; }
ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
ncclResult_t ncclSend(const void *b, size_t n, ncclDataType_t, int peer, ncclComm_t, cudaStream_t) { h_p2p((void *)b, n * 8, peer, 1); return ncclSuccess; }
ncclResult_t ncclRecv(void *b, size_t n, ncclDataType_t, int peer, ncclComm_t, cudaStream_t) { h_p2p(b, n * 8, peer, 0); return ncclSuccess; }
ncclResult_t ncclAllReduce(const void *s, void *r, size_t n, ncclDataType_t, ncclRedOp_t op, ncclComm_t, cudaStream_t)
{ if (op != ncclMax || s != r) return ncclInternalError; h_allmax((double *)r, n); return ncclSuccess; }
ncclResult_t ncclGroupStart(void) { h_group(0); return ncclSuccess; }
ncclResult_t ncclGroupEnd(void) { h_group(1); return ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t) { return "emulated nccl error"; }
}
