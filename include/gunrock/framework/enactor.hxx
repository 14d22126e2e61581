/**
 * @file enactor.hxx
 * @brief `gunrock::enactor_t<problem_t>` -- the bulk-synchronous driver
 * (include/gunrock/framework/enactor.hxx:78-344): ping-pong frontier buffers, `enact()` =
 * prepare_frontier + `while (!is_converged) { loop(); ++iteration; }` + finalize, timed with CUDA
 * events on the context stream (:266-288).  Public members keep the reference's names.
 *
 * Difference: nothing inside the loop forces a host sync except the convergence test itself
 * (the default `is_converged` needs the active frontier's size, which is read back lazily from
 * the device -- one small D2H per iteration instead of the reference's 3-5 blocking calls).
 */
#pragma once

#include <memory>
#include <vector>

#include <thrust/device_vector.h>

#include <gunrock/cuda/context.hxx>
#include <gunrock/framework/benchmark.hxx>
#include <gunrock/framework/frontier/frontier.hxx>

namespace gunrock {

struct enactor_properties_t {
  /// Output frontiers are reserved at max(E, V) * this factor (enactor.hxx:31-54).
  float frontier_sizing_factor{1.5f};
  std::size_t number_of_frontier_buffers{2};
  /// true: the algorithm does not use the enactor's frontiers (PageRank), none are allocated.
  bool self_manage_frontiers{false};
  enactor_properties_t() = default;
};

template <typename algorithm_problem_t,
          frontier::frontier_kind_t frontier_kind = frontier::frontier_kind_t::vertex_frontier,
          frontier::frontier_view_t frontier_view = frontier::frontier_view_t::vector>
struct enactor_t {
  using vertex_t = typename algorithm_problem_t::vertex_t;
  using edge_t = typename algorithm_problem_t::edge_t;
  using frontier_t = frontier::frontier_t<vertex_t, edge_t, frontier_kind>;

  enactor_properties_t properties;
  std::shared_ptr<gcuda::multi_context_t> context;
  algorithm_problem_t* problem;
  std::vector<frontier_t> frontiers;
  /// Kept for source compatibility (the reference's degree-scan scratch); the B200 advance keeps
  /// its scan in the context workspace, so this stays empty.
  thrust::device_vector<edge_t> scanned_work_domain;
  frontier_t* active_frontier;
  frontier_t* inactive_frontier;
  int buffer_selector;
  int iteration;

  enactor_t(const enactor_t& rhs) = delete;
  enactor_t& operator=(const enactor_t& rhs) = delete;

  enactor_t(algorithm_problem_t* _problem,
            std::shared_ptr<gcuda::multi_context_t> _context,
            enactor_properties_t _properties = enactor_properties_t())
      : properties(_properties),
        context(_context),
        problem(_problem),
        frontiers(_properties.number_of_frontier_buffers),
        active_frontier(&frontiers[0]),
        inactive_frontier(&frontiers[1]),
        buffer_selector(0),
        iteration(0) {
    auto stream = context->get_context(0)->stream();
    for (auto& buffer : frontiers)
      buffer.bind_stream(stream);
    if (!properties.self_manage_frontiers) {
      auto g = problem->get_graph();
      std::size_t initial_size = (g.get_number_of_edges() > g.get_number_of_vertices())
                                     ? g.get_number_of_edges()
                                     : g.get_number_of_vertices();
      for (auto& buffer : frontiers) {
        buffer.set_resizing_factor(properties.frontier_sizing_factor);
        buffer.reserve(initial_size);
      }
    }
  }
  virtual ~enactor_t() = default;

  algorithm_problem_t* get_problem() { return problem; }
  frontier_t* get_input_frontier() { return active_frontier; }
  frontier_t* get_output_frontier() { return inactive_frontier; }
  enactor_t* get_enactor() { return this; }

  void swap_frontier_buffers() {
    buffer_selector ^= 1;
    active_frontier = &frontiers[buffer_selector];
    inactive_frontier = &frontiers[buffer_selector ^ 1];
  }

  /// Run the algorithm to convergence; returns elapsed milliseconds (enactor.hxx:243-288).
  float enact() {
    iteration = 0;
    buffer_selector = 0;
    active_frontier = &frontiers[0];
    inactive_frontier = &frontiers[1];
    auto single_context = context->get_context(0);
    auto stream = single_context->stream();
    for (auto& frontier : frontiers) {
      frontier.bind_stream(stream);
      frontier.set_number_of_elements(0);
    }
    auto& timer = single_context->timer();
    auto& ws = single_context->workspace();
    const unsigned long long edges0 = ws.edges_accounted;
    timer.reset();
    timer.begin(stream);
    prepare_frontier(get_input_frontier(), *context);
    while (!is_converged(*context)) {
      loop(*context);
      ++iteration;
    }
    finalize(*context);
    auto runtime = timer.end(stream);
    auto& bench = benchmark::detail::current();
    bench.search_depth = iteration;
    bench.total_runtime = runtime;
    bench.edges_visited += ws.edges_accounted - edges0;
    return runtime;
  }

  virtual void loop(gcuda::multi_context_t& context) = 0;
  virtual void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) {}
  virtual bool is_converged(gcuda::multi_context_t& context) { return active_frontier->is_empty(); }
  virtual void finalize(gcuda::multi_context_t& context) {}
};

}  // namespace gunrock
