// tests/cuda_emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A single-OS-thread, fiber-based stand-in for the slice of the CUDA runtime and device
// language that porousfreezethaw_b200/csrc uses.  Compiling pft_b200.cu with
//     g++ -x c++ -DPFT_CUDA_EMU -Itests/cuda_emu ...
// yields libpft_b200_emu.so: the SAME host control flow and the SAME kernel source, executed
// one CTA at a time with one fiber per CUDA thread (__syncthreads = yield to the scheduler).
// It lets the -m "not gpu" tests check kernel indexing, the fused stage algebra and the
// integrator's control flow against the oracle on a box without a GPU.
//
// Several host threads may launch at the same time (the in-process multi-GPU group drives one
// "device" per worker thread): the scheduler state is thread-local, `__shared__` variables are
// per-thread statics, and the system-scope loads/stores of the peer protocol are real atomics, so
// the halo-flag and error-maximum exchange between slabs is stepped for real on the CPU.
//
// It is never loaded by the product: porousfreezethaw_b200/_lib.py only loads libpft_b200.so
// (nvcc, sm_100a).  It finds no races inside a CTA (fibers run in a fixed order) and says nothing
// about performance.
#pragma once
#ifndef PFT_CUDA_EMU
#error "cuda_emu/cuda_runtime.h is only for -DPFT_CUDA_EMU builds"
#endif

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <functional>

struct uint3 { unsigned int x, y, z; };
struct dim3 {
	unsigned int x, y, z;
	dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace pft_emu {
extern thread_local uint3 g_threadIdx, g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;
extern thread_local unsigned char *g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
void syncthreads();
void yield();                        // let the other fibers of the CTA run (spin-wait loops)
void syncwarp();                     // barrier among the fibers of one warp (32 consecutive linear thread ids)
struct CUtensorMap_fwd;
uint32_t tensor_load_3d(void *dst, const void *map, int x, int y, int z);   // returns the bytes written
}

#define threadIdx (pft_emu::g_threadIdx)
#define blockIdx  (pft_emu::g_blockIdx)
#define blockDim  (pft_emu::g_blockDim)
#define gridDim   (pft_emu::g_gridDim)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define PFT_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(pft_emu::g_dyn_smem)
#define PFT_DYN_SMEM_ALIGNED(type, name) type *name = reinterpret_cast<type *>(pft_emu::g_dyn_smem)

static inline void __syncthreads() { pft_emu::syncthreads(); }
static inline void __syncwarp() { pft_emu::syncwarp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }

static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }

// ---- runtime API slice --------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600, cudaErrorMemoryAllocation = 2 };
typedef struct pft_emu_stream *cudaStream_t;
typedef struct pft_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); p->major = 10; p->minor = 0; p->multiProcessorCount = 4; strcpy(p->name, "emulated sm_100"); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T **d, void *h, unsigned) { *d = (T *)h; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return 1; }       // no peers on a CPU
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return 1; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
enum { cudaEnableDefault = 0 };
typedef int cudaDriverEntryPointQueryResult;
enum { cudaDriverEntryPointSuccess = 0 };
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long flags, cudaDriverEntryPointQueryResult *st);
