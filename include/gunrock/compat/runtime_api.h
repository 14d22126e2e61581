/**
 * @file runtime_api.h
 * @brief HIP spellings of the CUDA runtime, for code written against the reference's HIP-first tree
 * (include/gunrock/compat/runtime_api.h:33-157 maps the same names with macros).  The reference's unit tests and
 * user code call `hipDeviceSynchronize()`, compare against `hipSuccess`, print `hipGetErrorString(status)` ...;
 * with this header those sources compile unchanged with plain nvcc.
 *
 * No HIP backend stands behind it: every name is a typed alias, a constant or an inline forwarder to the CUDA
 * runtime call of the same meaning (no macros, so nothing leaks into code that does not ask for it by name).
 * Covered: errors, device / stream / event management, memory, attributes and occupancy -- what the reference's
 * headers, examples and unit tests use.
 */
#ifndef GUNROCK_B200_COMPAT_RUNTIME_API_H
#define GUNROCK_B200_COMPAT_RUNTIME_API_H

#include <cstddef>
#include <utility>

#include <cuda_runtime.h>

// ---- types ------------------------------------------------------------------------------------------------
using hipError_t = cudaError_t;
using hipStream_t = cudaStream_t;
using hipEvent_t = cudaEvent_t;
using hipDeviceProp_t = cudaDeviceProp;
using hipFuncAttributes = cudaFuncAttributes;
using hipFuncCache_t = cudaFuncCache;
using hipSharedMemConfig = cudaSharedMemConfig;
using hipMemcpyKind = cudaMemcpyKind;
using hipDeviceAttribute_t = cudaDeviceAttr;

// ---- constants --------------------------------------------------------------------------------------------
constexpr hipError_t hipSuccess = cudaSuccess;
constexpr hipError_t hipErrorUnknown = cudaErrorUnknown;
constexpr hipError_t hipErrorInvalidValue = cudaErrorInvalidValue;
constexpr hipError_t hipErrorNotSupported = cudaErrorNotSupported;
constexpr unsigned int hipStreamNonBlocking = cudaStreamNonBlocking;
constexpr unsigned int hipEventDisableTiming = cudaEventDisableTiming;
constexpr hipMemcpyKind hipMemcpyHostToDevice = cudaMemcpyHostToDevice;
constexpr hipMemcpyKind hipMemcpyDeviceToHost = cudaMemcpyDeviceToHost;
constexpr hipMemcpyKind hipMemcpyDeviceToDevice = cudaMemcpyDeviceToDevice;
constexpr hipMemcpyKind hipMemcpyDefault = cudaMemcpyDefault;
constexpr hipFuncCache_t hipFuncCachePreferNone = cudaFuncCachePreferNone;
constexpr hipFuncCache_t hipFuncCachePreferShared = cudaFuncCachePreferShared;
constexpr hipFuncCache_t hipFuncCachePreferL1 = cudaFuncCachePreferL1;
constexpr hipFuncCache_t hipFuncCachePreferEqual = cudaFuncCachePreferEqual;
constexpr hipSharedMemConfig hipSharedMemBankSizeDefault = cudaSharedMemBankSizeDefault;
constexpr hipSharedMemConfig hipSharedMemBankSizeFourByte = cudaSharedMemBankSizeFourByte;
constexpr hipSharedMemConfig hipSharedMemBankSizeEightByte = cudaSharedMemBankSizeEightByte;
constexpr hipDeviceAttribute_t hipDeviceAttributeMaxGridDimX = cudaDevAttrMaxGridDimX;
constexpr hipDeviceAttribute_t hipDeviceAttributeClockRate = cudaDevAttrClockRate;
constexpr hipDeviceAttribute_t hipDeviceAttributeMemoryClockRate = cudaDevAttrMemoryClockRate;

// ---- calls: `hipX(args...)` is `cudaX(args...)` ----------------------------------------------------------------
#define GUNROCK_B200_HIP_FORWARD(hip_name, cuda_name)                 \
  template <typename... args_t>                                        \
  inline auto hip_name(args_t&&... args)                               \
      ->decltype(cuda_name(std::forward<args_t>(args)...)) {           \
    return cuda_name(std::forward<args_t>(args)...);                   \
  }
GUNROCK_B200_HIP_FORWARD(hipGetErrorString, cudaGetErrorString)
GUNROCK_B200_HIP_FORWARD(hipGetErrorName, cudaGetErrorName)
GUNROCK_B200_HIP_FORWARD(hipGetLastError, cudaGetLastError)
GUNROCK_B200_HIP_FORWARD(hipPeekAtLastError, cudaPeekAtLastError)
GUNROCK_B200_HIP_FORWARD(hipSetDevice, cudaSetDevice)
GUNROCK_B200_HIP_FORWARD(hipGetDevice, cudaGetDevice)
GUNROCK_B200_HIP_FORWARD(hipGetDeviceCount, cudaGetDeviceCount)
GUNROCK_B200_HIP_FORWARD(hipGetDeviceProperties, cudaGetDeviceProperties)
GUNROCK_B200_HIP_FORWARD(hipDeviceSynchronize, cudaDeviceSynchronize)
GUNROCK_B200_HIP_FORWARD(hipDeviceGetAttribute, cudaDeviceGetAttribute)
GUNROCK_B200_HIP_FORWARD(hipDeviceEnablePeerAccess, cudaDeviceEnablePeerAccess)
GUNROCK_B200_HIP_FORWARD(hipDeviceCanAccessPeer, cudaDeviceCanAccessPeer)
GUNROCK_B200_HIP_FORWARD(hipDriverGetVersion, cudaDriverGetVersion)
GUNROCK_B200_HIP_FORWARD(hipRuntimeGetVersion, cudaRuntimeGetVersion)
GUNROCK_B200_HIP_FORWARD(hipStreamCreate, cudaStreamCreate)
GUNROCK_B200_HIP_FORWARD(hipStreamCreateWithFlags, cudaStreamCreateWithFlags)
GUNROCK_B200_HIP_FORWARD(hipStreamSynchronize, cudaStreamSynchronize)
GUNROCK_B200_HIP_FORWARD(hipStreamDestroy, cudaStreamDestroy)
GUNROCK_B200_HIP_FORWARD(hipEventCreate, cudaEventCreate)
GUNROCK_B200_HIP_FORWARD(hipEventCreateWithFlags, cudaEventCreateWithFlags)
GUNROCK_B200_HIP_FORWARD(hipEventRecord, cudaEventRecord)
GUNROCK_B200_HIP_FORWARD(hipEventSynchronize, cudaEventSynchronize)
GUNROCK_B200_HIP_FORWARD(hipEventElapsedTime, cudaEventElapsedTime)
GUNROCK_B200_HIP_FORWARD(hipEventDestroy, cudaEventDestroy)
GUNROCK_B200_HIP_FORWARD(hipMalloc, cudaMalloc)
GUNROCK_B200_HIP_FORWARD(hipFree, cudaFree)
GUNROCK_B200_HIP_FORWARD(hipHostMalloc, cudaMallocHost)
GUNROCK_B200_HIP_FORWARD(hipHostFree, cudaFreeHost)
GUNROCK_B200_HIP_FORWARD(hipMemcpy, cudaMemcpy)
GUNROCK_B200_HIP_FORWARD(hipMemcpyAsync, cudaMemcpyAsync)
GUNROCK_B200_HIP_FORWARD(hipMemset, cudaMemset)
GUNROCK_B200_HIP_FORWARD(hipMemsetAsync, cudaMemsetAsync)
GUNROCK_B200_HIP_FORWARD(hipMemGetInfo, cudaMemGetInfo)
GUNROCK_B200_HIP_FORWARD(hipFuncGetAttributes, cudaFuncGetAttributes)
GUNROCK_B200_HIP_FORWARD(hipOccupancyMaxActiveBlocksPerMultiprocessor, cudaOccupancyMaxActiveBlocksPerMultiprocessor)
GUNROCK_B200_HIP_FORWARD(hipLaunchCooperativeKernel, cudaLaunchCooperativeKernel)
#undef GUNROCK_B200_HIP_FORWARD

#endif  // GUNROCK_B200_COMPAT_RUNTIME_API_H
