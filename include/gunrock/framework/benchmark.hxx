/**
 * @file benchmark.hxx
 * @brief Run counters the examples collect around every run (include/gunrock/framework/benchmark.hxx:43-98):
 * `benchmark::host_benchmark_t`, `INIT_BENCH / EXTRACT / DESTROY_BENCH`, and the device-side
 * `LOG_EDGE_VISITED / LOG_VERTEX_VISITED` hooks.
 *
 * The reference keeps a global `__managed__` struct bumped with atomicAdd from inside kernels and
 * only when built with ESSENTIALS_COLLECT_METRICS.  Here the advance kernels always account the
 * edges they inspect in their control block (gunrock/b200/runtime.cuh ctrl_t::edges); the enactor
 * adds them up into this host-side record at the end of enact(), so metrics cost nothing extra
 * and need no special build.
 */
#pragma once

#include <cstddef>

namespace gunrock {
namespace benchmark {

struct host_benchmark_t {
  unsigned long long edges_visited = 0;
  unsigned long long vertices_visited = 0;
  int search_depth = 0;
  float total_runtime = 0.0f;
};

namespace detail {
inline host_benchmark_t& current() {
  static host_benchmark_t instance;
  return instance;
}
}  // namespace detail

inline void INIT_BENCH() {
  detail::current() = host_benchmark_t();
}
inline host_benchmark_t EXTRACT() {
  return detail::current();
}
inline void DESTROY_BENCH() {
  detail::current() = host_benchmark_t();
}

/// Source-compatible no-ops for user kernels written against the reference's hooks.
__host__ __device__ inline void LOG_EDGE_VISITED(std::size_t = 1) {}
__host__ __device__ inline void LOG_VERTEX_VISITED(std::size_t = 1) {}

}  // namespace benchmark
}  // namespace gunrock
