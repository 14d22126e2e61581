/**
 * @file problem.hxx
 * @brief `gunrock::problem_t<graph_t>`: the per-algorithm data slice an enactor works on.
 *
 * Interface contract kept from the reference (include/gunrock/framework/problem.hxx:28-58) because user
 * algorithms derive from it: the graph view is held BY VALUE (`graph_slice`), the multi-context is
 * shared (`context`), `init()` / `reset()` are the two customisation points, and a problem can be
 * neither copied nor assigned (a copy would alias device arrays owned by the original).
 */
#pragma once

#include <memory>
#include <utility>

#include <gunrock/cuda/context.hxx>

namespace gunrock {

namespace detail {
/// Deleting copy operations once, here, keeps every problem type non-copyable by inheritance.
struct pinned_object_t {
  pinned_object_t() = default;
  pinned_object_t(const pinned_object_t&) = delete;
  pinned_object_t& operator=(const pinned_object_t&) = delete;
};
}  // namespace detail

template <typename graph_t>
struct problem_t : private detail::pinned_object_t {
  typedef typename graph_t::vertex_type vertex_t;
  typedef typename graph_t::edge_type edge_t;
  typedef typename graph_t::weight_type weight_t;
  typedef std::shared_ptr<gcuda::multi_context_t> context_ptr_t;

  graph_t graph_slice;    ///< non-owning device view, copied from the caller's graph
  context_ptr_t context;  ///< streams / events / workspace of every device in use

  problem_t() : graph_slice(nullptr), context() {}
  problem_t(graph_t& G, context_ptr_t ctx) : graph_slice(G), context(std::move(ctx)) {}
  virtual ~problem_t() = default;

  /// allocate the algorithm's device state (called once by the algorithm's run())
  virtual void init() = 0;
  /// bring that state back to "before the first iteration" (called before every enact())
  virtual void reset() = 0;

  graph_t get_graph() { return graph_slice; }
  context_ptr_t get_multi_context() { return context; }
  auto get_single_context(gcuda::device_id_t ordinal = 0) { return context->get_context(ordinal); }
};

}  // namespace gunrock
