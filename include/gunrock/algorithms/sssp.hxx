/**
 * @file sssp.hxx
 * @brief Single-source shortest paths -- same surface as include/gunrock/algorithms/sssp.hxx
 * (`param_t`, `result_t`, `problem_t`, `enactor_t`, both `run` overloads :176-230).
 * `run()` drives the fused B200 enactor (gunrock/b200/sssp.cuh); `problem_t`/`enactor_t` are the
 * operator-level formulation (advance with the relax lambda, bypass filter on iteration stamps,
 * optional uniquify) for direct instantiation or GUNROCK_B200_OPERATOR_PATH builds.
 * Result: fp32 distances, FLT_MAX if unreachable, bit-identical to the reference's CPU validator
 * (argument in gunrock/b200/sssp.cuh).
 */
#pragma once

#include <limits>

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/b200/sssp.cuh>
#include <gunrock/b200/part_multi.cuh>

namespace gunrock {
namespace sssp {

template <typename vertex_t>
struct param_t {
  vertex_t single_source;
  options_t options;
  param_t(vertex_t _single_source, options_t _options = options_t())
      : single_source(_single_source), options(_options) {}
};

template <typename vertex_t, typename weight_t>
struct result_t {
  weight_t* distances;
  vertex_t* predecessors;
  result_t(weight_t* _distances, vertex_t* _predecessors, vertex_t n_vertices)
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

  thrust::device_vector<vertex_t> visited;  // iteration stamps

  void init() override {
    visited.resize(this->get_graph().get_number_of_vertices());
  }

  void reset() override {
    auto ctx = this->get_single_context();
    auto n = this->get_graph().get_number_of_vertices();
    auto distances = this->result.distances;
    auto stamps = visited.data().get();
    auto source = this->param.single_source;
    auto fill = [distances, stamps, source] __device__(int i) {
      distances[i] = (i == source) ? weight_t(0) : std::numeric_limits<weight_t>::max();
      stamps[i] = -1;
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
    auto stamps = P->visited.data().get();
    auto iteration = this->iteration;

    auto relax = [distances] __host__ __device__(vertex_t const& source, vertex_t const& neighbor,
                                                 edge_t const& edge, weight_t const& weight) -> bool {
      weight_t candidate = thread::load(&distances[source]) + weight;
      weight_t before = math::atomic::min(&distances[neighbor], candidate);
      return candidate < before;
    };
    // One occurrence per vertex and iteration survives (exchange makes the stamp test race-free).
    auto first_this_iteration = [stamps, iteration] __host__ __device__(
                                    vertex_t const& vertex) -> bool {
      return math::atomic::exch(&stamps[vertex], static_cast<vertex_t>(iteration)) != iteration;
    };
    operators::advance::execute_runtime(G, E, relax, P->param.options.advance_load_balance,
                                        context);
    operators::filter::execute<operators::filter_algorithm_t::predicated>(
        G, E, first_this_iteration, context);
    if (P->param.options.enable_uniquify)
      operators::uniquify::execute<operators::uniquify_algorithm_t::unique>(
          E, context, P->param.options.best_effort_uniquify, P->param.options.uniquify_percent);
  }
};

template <typename graph_t>
float run(graph_t& G,
          param_t<typename graph_t::vertex_type>& param,
          result_t<typename graph_t::vertex_type, typename graph_t::weight_type>& result,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using vertex_t = typename graph_t::vertex_type;
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<vertex_t>;
  using result_type = result_t<vertex_t, weight_t>;
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
  b200::advance_launch_t cfg;
  cfg.lb = operators::advance::detail::to_lb(param.options.advance_load_balance);
  if (context->size() > 1) {
    // several devices: the 1-D partitioned SSSP (gunrock/b200/part_multi.cuh) -- the reference declares
    // multi_context_t (cuda/context.hxx:146-216) and throws here; same distances, bit for bit
    auto& cache = ctx->template scratch<b200::multi_sssp_cache_t>();
    b200::csr_view_t view = G.csr_view();
    b200::multi_partition(*context, cache, view);  // ingest: outside the timed region
    b200::part_sssp_report_t report;
    auto& timer = ctx->timer();
    timer.reset();
    timer.begin(ctx->stream());
    int iters = b200::sssp_run_multi(*context, cache, view, static_cast<int>(param.single_source),
                                     result.distances, cfg, &report);
    float ms = timer.end(ctx->stream());
    auto& bench = benchmark::detail::current();
    bench.search_depth = iters;
    bench.total_runtime = ms;
    bench.edges_visited += report.edges_relaxed;
    bench.vertices_visited += report.verts_total;
    return ms;
  }
  std::vector<b200::sssp_level_stat_t> levels;
  auto& timer = ctx->timer();
  timer.reset();
  timer.begin(ctx->stream());
  int iters = b200::sssp_run(ws, ctx->template scratch<b200::sssp_scratch_t>(), G.csr_view(),
                             static_cast<int>(param.single_source), result.distances, cfg, &levels);
  float ms = timer.end(ctx->stream());
  auto& bench = benchmark::detail::current();
  bench.search_depth = iters;
  bench.total_runtime = ms;
  for (auto& l : levels) {
    bench.edges_visited += l.edges_relaxed;
    bench.vertices_visited += static_cast<unsigned long long>(l.frontier);
  }
  return ms;
#endif
}

template <typename graph_t>
float run(graph_t& G,
          typename graph_t::vertex_type& single_source,
          typename graph_t::weight_type* distances,
          typename graph_t::vertex_type* predecessors,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using vertex_t = typename graph_t::vertex_type;
  using weight_t = typename graph_t::weight_type;
  param_t<vertex_t> param(single_source);
  result_t<vertex_t, weight_t> result(distances, predecessors, G.get_number_of_vertices());
  return run(G, param, result, context);
}

}  // namespace sssp
}  // namespace gunrock
