/**
 * @file bfs.hxx
 * @brief Breadth-first search -- same surface as include/gunrock/algorithms/bfs.hxx:
 * `bfs::param_t`, `bfs::result_t`, `bfs::problem_t`, `bfs::enactor_t`, `bfs::run` (param/result
 * form :162-182 and the legacy pointer form :201-215).
 *
 * Two execution paths behind that surface:
 *   - `run()` drives the fused B200 enactor (gunrock/b200/bfs.cuh): bitmap visited set, compacting
 *     top-down advance with the load balancer named in `options.advance_load_balance`, and -- when
 *     `options.advance_direction` asks for it -- the bottom-up sweep / Beamer switch.
 *   - `problem_t` + `enactor_t` remain a plain operator-level formulation (advance with a user
 *     lambda, optional filter) for callers that instantiate them directly, exactly like the
 *     reference's loop (:93-147); define GUNROCK_B200_OPERATOR_PATH to make `run()` use it too.
 * Result: distances[v] = depth, INT_MAX if unreachable; predecessors untouched (:29).
 */
#pragma once

#include <limits>

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/b200/bfs.cuh>
#include <gunrock/b200/bfs_multi.cuh>

namespace gunrock {
namespace bfs {

template <typename vertex_t>
struct param_t {
  vertex_t single_source;
  options_t options;
  param_t(vertex_t _single_source, options_t _options = options_t())
      : single_source(_single_source), options(_options) {}
};

template <typename vertex_t>
struct result_t {
  vertex_t* distances;
  vertex_t* predecessors;
  result_t(vertex_t* _distances, vertex_t* _predecessors)
      : distances(_distances), predecessors(_predecessors) {}
};

template <typename graph_t, typename param_type, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  param_type param;
  result_type result;

  problem_t(graph_t& G,
            param_type& _param,
            result_type& _result,
            std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), param(_param), result(_result) {}

  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  void init() override {}

  void reset() override {
    auto ctx = this->get_single_context();
    auto n = this->get_graph().get_number_of_vertices();
    auto distances = this->result.distances;
    auto source = this->param.single_source;
    auto fill = [distances, source] __device__(int i) {
      distances[i] = (i == source) ? 0 : std::numeric_limits<vertex_t>::max();
    };
    b200::for_each_index<<<b200::device_info_t::get().sm_count * 8, 256, 0, ctx->stream()>>>(
        static_cast<int>(n), fill);
  }
};

template <typename problem_t>
struct enactor_t : gunrock::enactor_t<problem_t> {
  enactor_t(problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::enactor_t<problem_t>(_problem, _context) {}

  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;
  using frontier_t = typename enactor_t<problem_t>::frontier_t;

  void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) override {
    f->push_back(this->get_problem()->param.single_source);
  }

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    auto distances = P->result.distances;
    vertex_t next_level = this->iteration + 1;

    // Level-synchronous claim: the first thread to lower the label owns the vertex.
    auto claim = [distances, next_level] __host__ __device__(
                     vertex_t const& source, vertex_t const& neighbor, edge_t const& edge,
                     weight_t const& weight) -> bool {
      vertex_t before = math::atomic::min(&distances[neighbor], next_level);
      return next_level < before;
    };
    operators::advance::execute_runtime(G, E, claim, P->param.options.advance_load_balance,
                                        context);
    if (P->param.options.enable_filter) {
      auto keep_all = [] __host__ __device__(vertex_t const& vertex) -> bool { return true; };
      operators::filter::execute_runtime(G, E, keep_all, P->param.options.filter_algorithm,
                                         context);
    }
  }
};

namespace detail {
/**
 * @brief `bfs::run` over a `gcuda::multi_context_t` of several devices (the reference throws here,
 * include/gunrock/framework/operators/advance/advance.hxx:129-132).  The graph and `result.distances` live on
 * the FIRST context's device, as for a single-device run; the graph is cut 1-D across the devices once (cached
 * on the context), every device runs its rank of the level loop on its own context's stream with the frontier
 * exchange done by the kernels over peer memory, and the depths are gathered back into `result.distances`.
 * Same result contract as the single-device run (bit-identical depths).
 */
template <typename graph_t>
float run_multi(graph_t& G,
                param_t<typename graph_t::vertex_type>& param,
                result_t<typename graph_t::vertex_type>& result,
                gcuda::multi_context_t& context) {
  auto ctx = context.get_context(0);
  b200::part_bfs_config_t cfg;
  cfg.advance.lb = operators::advance::detail::to_lb(param.options.advance_load_balance);
  cfg.direction = static_cast<int>(param.options.advance_direction);
  b200::csr_view_t out_view = G.csr_view();
  b200::csr_view_t in_view;  // row_offsets == nullptr: pull disabled
  if (cfg.direction != 0) {
    if (G.has_csc())
      in_view = G.csc_view();
    else if (G.properties.symmetric)
      in_view = out_view;
    else
      cfg.direction = 0;  // no transpose available: stay top-down
  }
  auto& cache = ctx->template scratch<b200::multi_bfs_cache_t>();
  b200::part_bfs_report_t report;
  b200::bfs_prepare_multi(context, cache, out_view, in_view);  // ingest: outside the timed region
  auto& timer = ctx->timer();
  timer.reset();
  timer.begin(ctx->stream());
  int depth = b200::bfs_run_multi(context, cache, out_view, in_view, static_cast<int>(param.single_source),
                                  reinterpret_cast<int*>(result.distances), cfg, &report);
  float ms = timer.end(ctx->stream());
  auto& bench = benchmark::detail::current();
  bench.search_depth = depth;
  bench.total_runtime = ms;
  bench.edges_visited += report.edges_total;
  bench.vertices_visited += report.verts_total;
  return ms;
}
}  // namespace detail

template <typename graph_t>
float run(graph_t& G,
          param_t<typename graph_t::vertex_type>& param,
          result_t<typename graph_t::vertex_type>& result,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using vertex_t = typename graph_t::vertex_type;
  using param_type = param_t<vertex_t>;
  using result_type = result_t<vertex_t>;
#ifdef GUNROCK_B200_OPERATOR_PATH
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type>;
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  return enactor.enact();
#else
  error::throw_if_exception(context->size() < 1, "empty multi_context_t");
  auto ctx = context->get_context(0);
  auto& ws = ctx->workspace();
  if (context->size() > 1)  // several devices: the 1-D partitioned traversal (gunrock/b200/bfs_multi.cuh)
    return detail::run_multi(G, param, result, *context);
  b200::bfs_config_t cfg;
  cfg.advance.lb = operators::advance::detail::to_lb(param.options.advance_load_balance);
  cfg.direction = static_cast<int>(param.options.advance_direction);
  b200::csr_view_t out_view = G.csr_view();
  b200::csr_view_t in_view;  // row_offsets == nullptr: pull disabled
  if (cfg.direction != 0) {
    if (G.has_csc())
      in_view = G.csc_view();
    else if (G.properties.symmetric)
      in_view = out_view;
    else
      cfg.direction = 0;  // no transpose available: stay top-down
  }
  std::vector<b200::bfs_level_stat_t> levels;
  auto& timer = ctx->timer();
  timer.reset();
  timer.begin(ctx->stream());
  int depth = b200::bfs_run(ws, ctx->template scratch<b200::bfs_scratch_t>(), out_view, in_view,
                            static_cast<int>(param.single_source), result.distances, cfg, &levels);
  float ms = timer.end(ctx->stream());
  auto& bench = benchmark::detail::current();
  bench.search_depth = depth;
  bench.total_runtime = ms;
  for (auto& l : levels) {
    bench.edges_visited += l.edges_inspected;
    bench.vertices_visited += static_cast<unsigned long long>(l.frontier);
  }
  return ms;
#endif
}

template <typename graph_t>
float run(graph_t& G,
          typename graph_t::vertex_type& single_source,
          typename graph_t::vertex_type* distances,
          typename graph_t::vertex_type* predecessors,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using vertex_t = typename graph_t::vertex_type;
  param_t<vertex_t> param(single_source);
  result_t<vertex_t> result(distances, predecessors);
  return run(G, param, result, context);
}

}  // namespace bfs
}  // namespace gunrock
