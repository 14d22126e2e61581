/**
 * @file pr.hxx
 * @brief PageRank -- same surface as include/gunrock/algorithms/pr.hxx (`param_t`, `result_t`,
 * `problem_t`, `enactor_t`, both `run` overloads :211-265); alpha / tol semantics, convergence rule
 * (`max|p - plast| < tol`, tested from the second iteration on, no cap) unchanged (:172-195).
 *
 * `run()` drives the deterministic pull enactor (gunrock/b200/pr.cuh) over the transpose: if the
 * graph view carries CSC arrays they are used, a symmetric graph is its own transpose, otherwise
 * the transpose is built once per (graph, context) on the device and cached.
 * `problem_t` / `enactor_t` keep the reference's push formulation on the generic operators
 * (parallel_for over edges + atomic add) for direct instantiation / GUNROCK_B200_OPERATOR_PATH.
 */
#pragma once

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/b200/pr.cuh>
#include <gunrock/b200/transpose.cuh>
#include <gunrock/b200/part_multi.cuh>

namespace gunrock {
namespace pr {

template <typename weight_t>
struct param_t {
  weight_t alpha;
  weight_t tol;
  options_t options;
  param_t(weight_t _alpha, weight_t _tol, options_t _options = options_t())
      : alpha(_alpha), tol(_tol), options(_options) {}
};

template <typename weight_t>
struct result_t {
  weight_t* p;
  result_t(weight_t* _p) : p(_p) {}
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

  thrust::device_vector<weight_t> plast;     // ranks of the previous iteration
  thrust::device_vector<weight_t> iweights;  // alpha / (sum of out-weights), 0 for dangling
  thrust::device_vector<weight_t> scratch;   // dangling mass / error accumulators

  void init() override {
    auto n = this->get_graph().get_number_of_vertices();
    plast.resize(n);
    iweights.resize(n);
    scratch.resize(2);
  }

  void reset() override {
    auto ctx = this->get_single_context();
    auto g = this->get_graph();
    auto n = g.get_number_of_vertices();
    auto alpha = this->param.alpha;
    auto p = this->result.p;
    auto pl = plast.data().get();
    auto iw = iweights.data().get();
    auto fill = [=] __device__(int v) {
      p[v] = static_cast<weight_t>(1.0 / n);
      pl[v] = 0;
      weight_t val = 0;
      edge_t start = g.get_starting_edge(v);
      edge_t end = start + g.get_number_of_neighbors(v);
      for (edge_t e = start; e < end; ++e)
        val += g.get_edge_weight(e);
      iw[v] = val != 0 ? alpha / val : 0;
    };
    b200::for_each_index<<<b200::device_info_t::get().sm_count * 8, 256, 0, ctx->stream()>>>(
        static_cast<int>(n), fill);
  }
};

template <typename problem_t>
struct enactor_t : gunrock::enactor_t<problem_t> {
  enactor_t(problem_t* _problem,
            std::shared_ptr<gcuda::multi_context_t> _context,
            enactor_properties_t _properties)
      : gunrock::enactor_t<problem_t>(_problem, _context, _properties) {}

  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;

  void loop(gcuda::multi_context_t& context) override {
    auto P = this->get_problem();
    auto G = P->get_graph();
    auto n = G.get_number_of_vertices();
    auto p = P->result.p;
    auto plast = P->plast.data().get();
    auto iweights = P->iweights.data().get();
    auto acc = P->scratch.data().get();
    auto alpha = P->param.alpha;
    auto ctx = context.get_context(0);
    cudaMemsetAsync(acc, 0, 2 * sizeof(weight_t), ctx->stream());
    // plast <- p, dangling mass
    auto stash = [=] __host__ __device__(vertex_t const& v) {
      plast[v] = p[v];
      if (iweights[v] == 0)
        math::atomic::add(acc, alpha * p[v]);
    };
    operators::parallel_for::execute<operators::parallel_for_each_t::vertex>(G, stash, context);
    auto base = [=] __host__ __device__(vertex_t const& v) { p[v] = (1 - alpha + acc[0]) / n; };
    operators::parallel_for::execute<operators::parallel_for_each_t::vertex>(G, base, context);
    auto spread = [=] __host__ __device__(edge_t const& e) {
      auto src = G.get_source_vertex(e);
      auto dst = G.get_destination_vertex(e);
      math::atomic::add(p + dst, plast[src] * iweights[src] * G.get_edge_weight(e));
    };
    operators::parallel_for::execute<operators::parallel_for_each_t::edge>(G, spread, context);
  }

  bool is_converged(gcuda::multi_context_t& context) override {
    if (this->iteration == 0)
      return false;
    auto P = this->get_problem();
    auto G = P->get_graph();
    auto p = P->result.p;
    auto plast = P->plast.data().get();
    auto acc = P->scratch.data().get() + 1;
    auto ctx = context.get_context(0);
    cudaMemsetAsync(acc, 0, sizeof(weight_t), ctx->stream());
    auto diff = [=] __host__ __device__(vertex_t const& v) {
      weight_t d = p[v] - plast[v];
      math::atomic::max(acc, d < 0 ? -d : d);
    };
    operators::parallel_for::execute<operators::parallel_for_each_t::vertex>(G, diff, context);
    weight_t err = 0;
    cudaMemcpyAsync(&err, acc, sizeof(weight_t), cudaMemcpyDeviceToHost, ctx->stream());
    ctx->synchronize();
    return err < P->param.tol;
  }
};

namespace detail {
/// Per-context cache of the fused PageRank (owned by the context: gcuda::standard_context_t::scratch).
struct pr_cache_t {
  b200::pr_scratch_t scratch;
  b200::transpose_t transpose;
  b200::graph_key_t transposed_for;  // the CSR the cached transpose was built from (identity, not address)
};
}  // namespace detail

template <typename graph_t>
float run(graph_t& G,
          param_t<typename graph_t::weight_type>& param,
          result_t<typename graph_t::weight_type>& result,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<weight_t>;
  using result_type = result_t<weight_t>;
#ifdef GUNROCK_B200_OPERATOR_PATH
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type>;
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_properties_t props;
  props.self_manage_frontiers = true;
  enactor_type enactor(&problem, context, props);
  return enactor.enact();
#else
  error::throw_if_exception(context->size() < 1, "empty multi_context_t");
  auto ctx = context->get_context(0);
  auto& ws = ctx->workspace();
  auto& cache = ctx->template scratch<detail::pr_cache_t>();
  b200::csr_view_t out_view = G.csr_view();
  b200::csr_view_t in_view;
  if (G.has_csc()) {
    in_view = G.csc_view();
  } else if (G.properties.symmetric) {
    in_view = out_view;
  } else {
    if (!cache.transposed_for.matches(out_view)) {  // ingest step, outside the timed loop
      cache.transpose.build(ws, out_view);            // (gives the transpose a fresh identity, so the
      cache.transposed_for.set(out_view);             //  tile table keyed on it is rebuilt as well)
    }
    in_view = cache.transpose.view;
  }
  auto& timer = ctx->timer();
  if (context->size() > 1) {
    // several devices: the destination-partitioned pull (gunrock/b200/part_multi.cuh); the reference declares
    // multi_context_t (cuda/context.hxx:146-216) and throws here
    auto& mcache = ctx->template scratch<b200::multi_pr_cache_t>();
    b200::multi_partition(*context, mcache, in_view);  // ingest: outside the timed region
    timer.reset();
    timer.begin(ctx->stream());
    int iters = b200::pr_run_multi(*context, mcache, out_view, in_view, param.alpha, param.tol, 0, result.p);
    float ms = timer.end(ctx->stream());
    auto& bench = benchmark::detail::current();
    bench.search_depth = iters;
    bench.total_runtime = ms;
    bench.edges_visited += static_cast<unsigned long long>(out_view.n_edges) * iters;
    bench.vertices_visited += static_cast<unsigned long long>(out_view.n_vertices) * iters;
    return ms;
  }
  timer.reset();
  timer.begin(ctx->stream());
  int iters = b200::pr_run(ws, cache.scratch, out_view, in_view, param.alpha, param.tol, 0, result.p);
  float ms = timer.end(ctx->stream());
  auto& bench = benchmark::detail::current();
  bench.search_depth = iters;
  bench.total_runtime = ms;
  bench.edges_visited += static_cast<unsigned long long>(out_view.n_edges) * iters;
  bench.vertices_visited += static_cast<unsigned long long>(out_view.n_vertices) * iters;
  return ms;
#endif
}

template <typename graph_t>
float run(graph_t& G,
          typename graph_t::weight_type alpha,
          typename graph_t::weight_type tol,
          typename graph_t::weight_type* p,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using weight_t = typename graph_t::weight_type;
  param_t<weight_t> param(alpha, tol);
  result_t<weight_t> result(p);
  return run(G, param, result, context);
}

}  // namespace pr
}  // namespace gunrock
