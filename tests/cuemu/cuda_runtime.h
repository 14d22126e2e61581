// tests/cuemu/cuda_runtime.h -- TEST INFRASTRUCTURE: the handful of CUDA runtime names that
// include/gunrock/b200/runtime.cuh mentions, so that the kernel headers compile under the CPU emulator.
// Nothing here does anything: the emulator never calls the host launchers; a call aborts.
#pragma once

#include <cstddef>
#include <cstdio>
#include <cstdlib>

typedef int cudaError_t;
typedef struct cuemu_stream* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };

inline void cuemu_no_runtime(const char* what) {
  std::fprintf(stderr, "cuemu: CUDA runtime call %s reached under the emulator\n", what);
  std::abort();
}
inline const char* cudaGetErrorName(cudaError_t) { return "cuemu"; }
inline const char* cudaGetErrorString(cudaError_t) { return "cuemu"; }
inline cudaError_t cudaGetDevice(int*) { cuemu_no_runtime("cudaGetDevice"); return 1; }
inline cudaError_t cudaDeviceGetAttribute(int*, cudaDeviceAttr, int) { cuemu_no_runtime("cudaDeviceGetAttribute"); return 1; }
template <typename T>
inline cudaError_t cudaMalloc(T**, std::size_t) { cuemu_no_runtime("cudaMalloc"); return 1; }
inline cudaError_t cudaFree(void*) { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void*, int, std::size_t, cudaStream_t) { cuemu_no_runtime("cudaMemsetAsync"); return 1; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { cuemu_no_runtime("cudaStreamQuery"); return 1; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { cuemu_no_runtime("cudaStreamSynchronize"); return 1; }
template <typename T>
inline cudaError_t cudaMallocHost(T**, std::size_t) { cuemu_no_runtime("cudaMallocHost"); return 1; }
inline cudaError_t cudaFreeHost(void*) { return cudaSuccess; }
