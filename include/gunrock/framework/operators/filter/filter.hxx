/**
 * @file filter.hxx
 * @brief `operators::filter::execute` (include/gunrock/framework/operators/filter/filter.hxx:72-100
 * explicit form, :144-168 enactor form, :187-211 `execute_runtime`).  An element x survives iff
 * `util::limits::is_valid(x) && op(x)`; `op` is never called on invalid elements.
 * predicated / remove / compact -> stable look-back select (gunrock/b200/filter.cuh);
 * bypass -> same size, rejected entries become invalid, output may alias input.
 * `compact` throws in the reference (filter/compact.hxx:21-24); it is implemented here.
 */
#pragma once

#include <gunrock/b200/filter.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace filter {

namespace detail {
template <typename type_t, typename operator_t>
struct pred_adapter_t {
  operator_t op;
  __device__ __forceinline__ bool operator()(int x) const {
    type_t v = x;
    return op(v);
  }
};
}  // namespace detail

template <filter_algorithm_t alg_type, typename graph_t, typename operator_t, typename frontier_t>
void execute(graph_t& G,
             operator_t op,
             frontier_t* input,
             frontier_t* output,
             gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  using type_t = typename frontier_t::type_t;
  detail::pred_adapter_t<type_t, operator_t> f{op};
  std::size_t bound = input->size_upper_bound();
  const bool unique_in = input->is_known_unique();
  if (output->get_capacity() < bound || output->get_capacity() == 0)
    output->reserve(bound ? bound : 1);
  output->bind_stream(ws.stream);
  const int* in = reinterpret_cast<const int*>(input->get());
  int* out = reinterpret_cast<int*>(output->get());
  if (alg_type == filter_algorithm_t::bypass) {
    b200::launch_filter_bypass(ws, in, input->count_ptr(), out, output->count_ptr(), f);
  } else {
    error::throw_if_exception(in == out, "predicated/remove/compact filter cannot run in place");
    b200::launch_filter_select(ws, in, input->count_ptr(), static_cast<int>(bound), out,
                               output->count_ptr(), f);
  }
  output->mark_produced(ws.stream, nullptr, /*unique=*/unique_in);  // a filter never adds an occurrence
}

template <filter_algorithm_t alg_type, typename graph_t, typename enactor_type, typename operator_t>
void execute(graph_t& G,
             enactor_type* E,
             operator_t op,
             gcuda::multi_context_t& context,
             bool swap_buffers = true) {
  execute<alg_type>(G, op, E->get_input_frontier(), E->get_output_frontier(), context);
  if (swap_buffers)
    E->swap_frontier_buffers();
}

template <typename graph_t, typename enactor_type, typename operator_t>
void execute_runtime(graph_t& G,
                     enactor_type* E,
                     operator_t op,
                     filter_algorithm_t alg,
                     gcuda::multi_context_t& context,
                     bool swap_buffers = true) {
  switch (alg) {
    case filter_algorithm_t::remove:
      execute<filter_algorithm_t::remove>(G, E, op, context, swap_buffers);
      break;
    case filter_algorithm_t::predicated:
      execute<filter_algorithm_t::predicated>(G, E, op, context, swap_buffers);
      break;
    case filter_algorithm_t::compact:
      execute<filter_algorithm_t::compact>(G, E, op, context, swap_buffers);
      break;
    case filter_algorithm_t::bypass:
      execute<filter_algorithm_t::bypass>(G, E, op, context, swap_buffers);
      break;
    default:
      error::throw_if_exception(cudaErrorUnknown, "Filter algorithm not supported.");
  }
}

}  // namespace filter
}  // namespace operators
}  // namespace gunrock
