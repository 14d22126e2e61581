/**
 * @file memory.hxx
 * @brief Memory spaces and raw allocation helpers (drop-in for the reference's
 * include/gunrock/memory.hxx:33-122: `memory_space_t`, `allocate`, `free`, `raw_pointer_cast`).
 */
#pragma once

#include <cstddef>
#include <cstdlib>

#include <cuda_runtime.h>
#include <thrust/device_ptr.h>

#include <gunrock/error.hxx>

namespace gunrock {
namespace memory {

enum memory_space_t { device, host };

template <typename type_t>
inline type_t* allocate(std::size_t bytes, memory_space_t space = memory_space_t::device) {
  void* p = nullptr;
  if (bytes == 0)
    return nullptr;
  if (space == memory_space_t::device)
    error::throw_if_exception(cudaMalloc(&p, bytes), "cudaMalloc failed");
  else
    error::throw_if_exception(cudaMallocHost(&p, bytes), "cudaMallocHost failed");
  return static_cast<type_t*>(p);
}

template <typename type_t>
inline void free(type_t* p, memory_space_t space = memory_space_t::device) {
  if (!p)
    return;
  if (space == memory_space_t::device)
    error::throw_if_exception(cudaFree(p), "cudaFree failed");
  else
    error::throw_if_exception(cudaFreeHost(p), "cudaFreeHost failed");
}

template <typename type_t>
inline type_t* raw_pointer_cast(type_t* p) {
  return p;
}
template <typename type_t>
inline type_t* raw_pointer_cast(thrust::device_ptr<type_t> p) {
  return thrust::raw_pointer_cast(p);
}

}  // namespace memory
}  // namespace gunrock
