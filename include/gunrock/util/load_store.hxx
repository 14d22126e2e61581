/**
 * @file load_store.hxx
 * @brief `thread::load` / `thread::store` (include/gunrock/util/load_store.hxx:57-85, which wraps
 * cub::ThreadLoad/ThreadStore with the default modifier).  Plain generic loads/stores here; the
 * streaming column-index reads inside the advance kernels use explicit cache hints instead
 * (include/gunrock/b200/ptx.cuh).
 */
#pragma once

namespace gunrock {
namespace thread {

template <typename type_t>
__host__ __device__ __forceinline__ type_t load(type_t* ptr) {
  return *ptr;
}
template <typename type_t>
__host__ __device__ __forceinline__ type_t load(const type_t* ptr) {
  return *ptr;
}
template <typename type_t>
__host__ __device__ __forceinline__ void store(type_t* ptr, const type_t& value) {
  *ptr = value;
}

}  // namespace thread
}  // namespace gunrock
