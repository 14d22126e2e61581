/**
 * @file for.hxx
 * @brief `operators::parallel_for::execute` (include/gunrock/framework/operators/for/for.hxx:25-107):
 * apply `op` to every vertex / edge / weight of the graph, or to every element of a frontier.
 * Grid-stride kernels on the context stream replace thrust::for_each; no host synchronisation.
 */
#pragma once

#include <gunrock/b200/ptx.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/frontier/frontier.hxx>  // parallel_for over a frontier; callers include only this header
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace parallel_for {

namespace detail {
template <typename frontier_t, typename operator_t>
__global__ void frontier_for_kernel(frontier_t f, operator_t op) {
  const std::size_t n = f.get_number_of_elements();
  for (std::size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
    auto x = f.get_element_at(i);
    op(x);
  }
}
}  // namespace detail

/// vertex / edge / weight over the whole graph.
template <parallel_for_each_t type, typename graph_t, typename operator_t>
std::enable_if_t<type != parallel_for_each_t::element> execute(graph_t& G,
                                                               operator_t op,
                                                               gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  const int grid = b200::device_info_t::get().sm_count * 8;
  if constexpr (type == parallel_for_each_t::vertex) {
    int n = static_cast<int>(G.get_number_of_vertices());
    auto f = [op] __device__(int i) {
      typename graph_t::vertex_type v = i;
      op(v);
    };
    b200::for_each_index<<<grid, 256, 0, ws.stream>>>(n, f);
  } else if constexpr (type == parallel_for_each_t::edge) {
    int n = static_cast<int>(G.get_number_of_edges());
    auto f = [op] __device__(int i) {
      typename graph_t::edge_type e = i;
      op(e);
    };
    b200::for_each_index<<<grid, 256, 0, ws.stream>>>(n, f);
  } else {
    int n = static_cast<int>(G.get_number_of_edges());
    auto values = G.get_nonzero_values();
    auto f = [op, values] __device__(int i) {
      typename graph_t::weight_type w = values[i];
      op(w);
    };
    b200::for_each_index<<<grid, 256, 0, ws.stream>>>(n, f);
  }
  ws.launches += 1;
  error::throw_if_exception(cudaGetLastError(), "parallel_for launch");
}

/// element: every entry of a frontier (for.hxx:84-107).
template <parallel_for_each_t type, typename frontier_t, typename operator_t>
std::enable_if_t<type == parallel_for_each_t::element> execute(frontier_t& f,
                                                               operator_t op,
                                                               gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  const int grid = b200::device_info_t::get().sm_count * 8;
  detail::frontier_for_kernel<<<grid, 256, 0, ws.stream>>>(f, op);
  ws.launches += 1;
  error::throw_if_exception(cudaGetLastError(), "parallel_for launch");
}

}  // namespace parallel_for
}  // namespace operators
}  // namespace gunrock
