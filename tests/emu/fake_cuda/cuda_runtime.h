// TEST INFRASTRUCTURE -- a host stand-in for the few CUDA runtime calls gs_horus.cu makes, so that the library's
// HOST side (handle management, slab layout in prepare(), stream / word-table uploads, run / stats / fetch, error
// codes) can be exercised through the real C ABI on a box without a GPU (tests/test_horus_abi_emu.py).
// "Device" memory is host memory; the kernels are replaced by gs_horus.cu's own host loop (#ifndef __CUDACC__).
#pragma once
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum cudaLimit { cudaLimitStackSize = 0 };
enum { cudaStreamNonBlocking = 1 };

#define __host__
#define __device__
#define __forceinline__ inline

static inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetLimit(size_t *v, cudaLimit) { *v = 1024; return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
