/**
 * @file advance.hxx
 * @brief `operators::advance::execute` -- same three entry points as the reference
 * (include/gunrock/framework/operators/advance/advance.hxx:94-133 eight-argument form, :196-225
 * enactor form, :247-275 `execute_runtime`), dispatching to the sm_100a kernels in
 * include/gunrock/b200/advance.cuh instead of block_mapped.hxx / merge_path.hxx / thread_mapped.hxx.
 *
 * Operator contract unchanged (advance.hxx:35-49): `op(source, neighbor, edge, weight) -> bool`,
 * called once per (input entry, out-edge); `true` adds the neighbour (or the edge id for an edge
 * output) to the output frontier.  Behavioural differences, all documented in DESIGN.md:
 *   - the output frontier is compact (no invalid slots); a following filter sees only valid ids.  The
 *     reference's layout (one slot per edge rank, -1 for rejected edges, merge_path.hxx:218-279) is an
 *     opt-in: `context.get_context(0)->reference_advance_output(true)` or -DGUNROCK_B200_REFERENCE_ADVANCE_OUTPUT;
 *   - nothing synchronises with the host; the output size stays on the device until asked for;
 *   - `direction` is accepted and ignored exactly as in the reference -- pull and
 *     direction-optimised traversal are provided by the fused enactors (gunrock/b200/bfs.cuh);
 *   - multi-context execution (`context.size() != 1`) is rejected here as in the reference
 *     (advance.hxx:129-132); the partitioned multi-GPU path lives above the C ABI.
 */
#pragma once

#include <type_traits>

#include <gunrock/b200/advance.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace advance {

namespace detail {

/// Adapts the user's `op(vertex_t const&, vertex_t const&, edge_t const&, weight_t const&)`.
template <typename graph_t,
          typename operator_t,
          bool float_weights = std::is_same<typename graph_t::weight_type, float>::value>
struct op_adapter_t {
  operator_t op;
  __device__ __forceinline__ bool operator()(int src, int dst, int edge, float w) const {
    typename graph_t::vertex_type s = src, d = dst;
    typename graph_t::edge_type e = edge;
    typename graph_t::weight_type wt = w;
    return op(s, d, e, wt);
  }
};
/// weight_t other than float (int, double ...): the kernels walk the structure only (graph_t::structure_view) and
/// the edge's weight is read here, in its own type, by edge id.
template <typename graph_t, typename operator_t>
struct op_adapter_t<graph_t, operator_t, false> {
  operator_t op;
  const typename graph_t::weight_type* typed_values;
  __device__ __forceinline__ bool operator()(int src, int dst, int edge, float) const {
    typename graph_t::vertex_type s = src, d = dst;
    typename graph_t::edge_type e = edge;
    typename graph_t::weight_type wt =
        typed_values ? typed_values[edge] : static_cast<typename graph_t::weight_type>(1);
    return op(s, d, e, wt);
  }
};
template <typename graph_t, typename operator_t>
op_adapter_t<graph_t, operator_t> make_op_adapter(graph_t& G, operator_t op) {
  if constexpr (std::is_same<typename graph_t::weight_type, float>::value)
    return op_adapter_t<graph_t, operator_t>{op};
  else
    return op_adapter_t<graph_t, operator_t>{op, G.get_nonzero_values()};
}
template <typename graph_t>
b200::csr_view_t kernel_view(graph_t& G) {
  if constexpr (std::is_same<typename graph_t::weight_type, float>::value)
    return G.csr_view();
  else
    return G.structure_view();
}

inline b200::lb_t to_lb(load_balance_t lb) {
  switch (lb) {
    case load_balance_t::thread_mapped:
      return b200::lb_t::thread_mapped;
    case load_balance_t::merge_path:
    case load_balance_t::merge_path_v2:
      return b200::lb_t::merge_path;
    default:
      return b200::lb_t::block_mapped;
  }
}

}  // namespace detail

/**
 * @brief Eight-argument form: explicit input / output frontiers.
 * `segments` (the reference's scanned work domain) is accepted for source compatibility; the
 * degree scan lives in the context's workspace.
 */
template <load_balance_t lb,
          advance_direction_t direction,
          advance_io_type_t input_type,
          advance_io_type_t output_type,
          typename graph_t,
          typename operator_t,
          typename frontier_t,
          typename work_tiles_t>
void execute(graph_t& G,
             operator_t op,
             frontier_t* input,
             frontier_t* output,
             work_tiles_t& segments,
             gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  static_assert(input_type != advance_io_type_t::edges,
                "edge input frontiers are not supported by the B200 advance");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  auto view = detail::kernel_view(G);
  auto f = detail::make_op_adapter(G, op);
  b200::advance_launch_t cfg;
  cfg.lb = detail::to_lb(lb);

  const int* in = nullptr;
  const int* in_count = nullptr;
  int in_bound = view.n_vertices;
  if (input_type != advance_io_type_t::graph) {
    in = reinterpret_cast<const int*>(input->get());
    in_count = input->count_ptr();
    in_bound = static_cast<int>(input->size_upper_bound());
  }
  b200::ctrl_t* ctrl = nullptr;
  if constexpr (output_type == advance_io_type_t::none) {
    b200::launch_advance<b200::advance_output_t::none, false, true>(
        ws, view, in, in_count, in_bound, nullptr, nullptr, 0, f, cfg, &ctrl);
  } else {
    // A duplicate-free frontier expands into at most E slots (the reference's enactor reserves
    // max(E,V)*1.5 up front, enactor.hxx:161-193), so nothing has to visit the host.  A frontier that may
    // hold the same vertex several times (the output of an earlier advance, a user-filled list) is not
    // bounded by E: its output is sized from the degree sum, as the reference does before every advance
    // (block_mapped.hxx:205-217) -- one pinned-memory poll, no stream synchronise.  The overflow flag the
    // kernels raise stays as the backstop (it reaches the frontier's own storage, mark_produced).
    std::size_t want = static_cast<std::size_t>(
        view.n_edges > view.n_vertices ? view.n_edges : view.n_vertices);
    const bool ranked = context0->reference_advance_output();
    if (input_type != advance_io_type_t::graph && (ranked || !input->is_known_unique())) {
      const unsigned long long total = b200::frontier_degree_total(ws, view, in, in_count, in_bound);
      error::throw_if_exception(total > 0x7fffffffull, "advance: the output frontier would exceed 2^31 entries");
      if (static_cast<std::size_t>(total) > want)
        want = static_cast<std::size_t>(total);
      cfg.edges_upper_bound = static_cast<long long>(total);
    }
    if (output->get_capacity() < want)
      output->reserve(want);
    output->bind_stream(ws.stream);
    output->set_number_of_elements(0);
    constexpr b200::advance_output_t kOut = output_type == advance_io_type_t::edges
                                                ? b200::advance_output_t::edges
                                                : b200::advance_output_t::vertices;
    if (ranked) {
      // the reference's shape: slot = edge rank, -1 where `op` said false; every load balancer shares the
      // rank-ordered tile kernel (the layout is the merge-path rank space)
      const std::size_t slots = input_type == advance_io_type_t::graph
                                    ? static_cast<std::size_t>(view.n_edges)
                                    : static_cast<std::size_t>(cfg.edges_upper_bound);
      b200::launch_advance_ranked<kOut, true>(
          ws, view, in, in_count, in_bound, reinterpret_cast<int*>(output->get()), output->count_ptr(),
          static_cast<int>(slots < output->get_capacity() ? slots : output->get_capacity()), f, cfg, &ctrl);
    } else {
      b200::launch_advance<kOut, false, true>(
          ws, view, in, in_count, in_bound, reinterpret_cast<int*>(output->get()), output->count_ptr(),
          static_cast<int>(output->get_capacity()), f, cfg, &ctrl);
    }
    output->mark_produced(ws.stream, ctrl);
  }
}

/// Enactor form: uses (and by default swaps) the enactor's ping-pong buffers (advance.hxx:196-225).
template <load_balance_t lb = load_balance_t::merge_path,
          advance_direction_t direction = advance_direction_t::forward,
          advance_io_type_t input_type = advance_io_type_t::vertices,
          advance_io_type_t output_type = advance_io_type_t::vertices,
          typename graph_t,
          typename enactor_type,
          typename operator_type>
void execute(graph_t& G,
             enactor_type* E,
             operator_type op,
             gcuda::multi_context_t& context,
             bool swap_buffers = true) {
  execute<lb, direction, input_type, output_type>(G, op, E->get_input_frontier(),
                                                  E->get_output_frontier(),
                                                  E->scanned_work_domain, context);
  if (swap_buffers && (output_type != advance_io_type_t::none))
    E->swap_frontier_buffers();
}

/// Runtime load-balance selection (advance.hxx:247-275).
template <typename graph_t, typename enactor_type, typename operator_type>
void execute_runtime(graph_t& G,
                     enactor_type* E,
                     operator_type op,
                     load_balance_t lb,
                     gcuda::multi_context_t& context,
                     bool swap_buffers = true) {
  constexpr auto fwd = advance_direction_t::forward;
  constexpr auto v = advance_io_type_t::vertices;
  switch (lb) {
    case load_balance_t::thread_mapped:
      execute<load_balance_t::thread_mapped, fwd, v, v>(G, E, op, context, swap_buffers);
      break;
    case load_balance_t::merge_path:
    case load_balance_t::merge_path_v2:
      execute<load_balance_t::merge_path, fwd, v, v>(G, E, op, context, swap_buffers);
      break;
    case load_balance_t::block_mapped:
    case load_balance_t::warp_mapped:
    case load_balance_t::bucketing:
    case load_balance_t::work_stealing:
      execute<load_balance_t::block_mapped, fwd, v, v>(G, E, op, context, swap_buffers);
      break;
    default:
      error::throw_if_exception(cudaErrorUnknown, "Load balance type not supported.");
  }
}

}  // namespace advance
}  // namespace operators
}  // namespace gunrock
