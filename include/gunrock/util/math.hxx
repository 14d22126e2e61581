/**
 * @file math.hxx
 * @brief `math::atomic::{add,min,max,cas,exch}` (include/gunrock/util/math.hxx:75-136).
 *
 * The reference keys the device branch on `__HIP_DEVICE_COMPILE__`, which nvcc never defines, so
 * its CUDA build silently takes the non-atomic host fallback.  Here the device
 * branch is keyed on `__CUDA_ARCH__`; float/double min/max use sign-aware integer atomics on the
 * IEEE bit pattern (one instruction, same returned old value as the reference's CAS loop in
 * include/gunrock/cuda/atomic_functions.hxx:35-121).
 */
#pragma once

#include <algorithm>

#include <cuda_runtime.h>

namespace gunrock {
namespace gcuda {

__device__ __forceinline__ float atomicMin(float* addr, float value) {
  return (value >= 0.0f)
             ? __int_as_float(::atomicMin(reinterpret_cast<int*>(addr), __float_as_int(value)))
             : __uint_as_float(
                   ::atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(value)));
}
__device__ __forceinline__ float atomicMax(float* addr, float value) {
  return (value >= 0.0f)
             ? __int_as_float(::atomicMax(reinterpret_cast<int*>(addr), __float_as_int(value)))
             : __uint_as_float(
                   ::atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(value)));
}
__device__ __forceinline__ double atomicMin(double* addr, double value) {
  using ll = long long;
  using ull = unsigned long long;
  return (value >= 0.0)
             ? __longlong_as_double(
                   ::atomicMin(reinterpret_cast<ll*>(addr), __double_as_longlong(value)))
             : __longlong_as_double(static_cast<ll>(::atomicMax(
                   reinterpret_cast<ull*>(addr), static_cast<ull>(__double_as_longlong(value)))));
}
__device__ __forceinline__ double atomicMax(double* addr, double value) {
  using ll = long long;
  using ull = unsigned long long;
  return (value >= 0.0)
             ? __longlong_as_double(
                   ::atomicMax(reinterpret_cast<ll*>(addr), __double_as_longlong(value)))
             : __longlong_as_double(static_cast<ll>(::atomicMin(
                   reinterpret_cast<ull*>(addr), static_cast<ull>(__double_as_longlong(value)))));
}

}  // namespace gcuda

namespace math {

/// Generic host/device min / max (used by geo.hxx and friends).
template <typename type_t>
__host__ __device__ __forceinline__ constexpr type_t min(const type_t& a, const type_t& b) {
  return b < a ? b : a;
}
template <typename type_t>
__host__ __device__ __forceinline__ constexpr type_t max(const type_t& a, const type_t& b) {
  return a < b ? b : a;
}

namespace atomic {

template <typename type_t>
__host__ __device__ __forceinline__ type_t add(type_t* address, const type_t& value) {
#ifdef __CUDA_ARCH__
  return ::atomicAdd(address, value);
#else
  type_t old = *address;
  *address += value;
  return old;
#endif
}

template <typename type_t>
__host__ __device__ __forceinline__ type_t min(type_t* address, const type_t& value) {
#ifdef __CUDA_ARCH__
  if constexpr (std::is_floating_point<type_t>::value)
    return gcuda::atomicMin(address, value);
  else
    return ::atomicMin(address, value);
#else
  type_t old = *address;
  *address = std::min(old, value);
  return old;
#endif
}

template <typename type_t>
__host__ __device__ __forceinline__ type_t max(type_t* address, const type_t& value) {
#ifdef __CUDA_ARCH__
  if constexpr (std::is_floating_point<type_t>::value)
    return gcuda::atomicMax(address, value);
  else
    return ::atomicMax(address, value);
#else
  type_t old = *address;
  *address = std::max(old, value);
  return old;
#endif
}

template <typename type_t>
__host__ __device__ __forceinline__ type_t cas(type_t* address,
                                               const type_t& compare,
                                               const type_t& value) {
#ifdef __CUDA_ARCH__
  return ::atomicCAS(address, compare, value);
#else
  type_t old = *address;
  *address = (old == compare) ? value : old;
  return old;
#endif
}

template <typename type_t>
__host__ __device__ __forceinline__ type_t exch(type_t* address, const type_t& value) {
#ifdef __CUDA_ARCH__
  return ::atomicExch(address, value);
#else
  type_t old = *address;
  *address = value;
  return old;
#endif
}

}  // namespace atomic
}  // namespace math
}  // namespace gunrock
