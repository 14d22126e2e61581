/**
 * @file vector.hxx
 * @brief `vector_t<T, space>`: thrust device/host vector alias (include/gunrock/container/vector.hxx:26-31).
 * Thrust containers stay at the API boundary (the examples use them directly,
 * examples/algorithms/bfs/bfs.cu:52-53); nothing on the hot path touches Thrust algorithms.
 */
#pragma once

#include <type_traits>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <gunrock/memory.hxx>

namespace gunrock {

using namespace memory;

template <typename type_t, memory_space_t space>
using vector_t = std::conditional_t<space == memory_space_t::host,
                                    thrust::host_vector<type_t>,
                                    thrust::device_vector<type_t>>;

}  // namespace gunrock
