// TEST HARNESS (not product code): the handful of CUDA runtime names divshot_b200/csrc/densify.cu uses, restated for
// a plain g++ build so the CPU suite can execute that file's kernel bodies and host orchestration serially
// (tests/native/densify_emul.cpp).  "Device" memory is malloc'd and filled with a 0xCD pattern so that a read of
// memory the code never wrote shows up as garbage instead of as a lucky zero.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline

enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorAssert = 710 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
typedef void* cudaStream_t;

template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
    if (!*p) return cudaErrorMemoryAllocation;
    std::memset(*p, 0xCD, bytes);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline int atomicAdd(int32_t* p, int v) { const int old = *p; *p += v; return old; }
