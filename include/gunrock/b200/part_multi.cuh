/**
 * @file part_multi.cuh
 * @brief SSSP and PageRank over SEVERAL devices of one process: what `sssp::run` / `pr::run` do when they are
 * handed a `gcuda::multi_context_t` with more than one context (the reference declares that surface,
 * include/gunrock/cuda/context.hxx:146-216, and throws for `size() != 1` in every algorithm and operator,
 * SURVEY.md F6).  Companion of bfs_multi.cuh:
 *   1. the caller's device graph (CSR with values for SSSP; the CSC -- in-edge rows -- for PageRank) is cut 1-D
 *      (vertex v -> rank v % P, row v / P) by kernels on the first context's device that write each rank's rows
 *      straight into that rank's memory; cached per graph identity and device list;
 *   2. one host thread per rank runs the loop of part_loops.cuh on its context's stream; the collectives are
 *      `thread_exchange_t` (peer loads + a host barrier between the threads);
 *   3. a gather kernel interleaves the ranks' slices into the caller's result array.
 * Results: SSSP distances bit-exact with the single-device run; PageRank uses the same iweights as the
 * single-device run (row sums taken from the whole CSR on the first device), ranks within the 1e-6 tolerance.
 */
#pragma once

#include <exception>
#include <memory>
#include <thread>
#include <vector>

#include <gunrock/b200/bfs_multi.cuh>
#include <gunrock/b200/part_loops.cuh>

namespace gunrock {
namespace b200 {

/// out[v] = slice[v % P][v / P] (peer loads from the launching device)
template <typename T>
static __global__ void part_interleave_any_kernel(peer_table_t t, int nparts, int n_global, T* __restrict__ out) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_global; v += gridDim.x * blockDim.x)
    out[v] = static_cast<const T*>(t.p[v % nparts])[v / nparts];
}

/// Row sums of the whole CSR for PageRank's iweights: out-degrees, or -- for a graph with values -- the row's
/// weights added sequentially in fp32 exactly as the single-device reset does (pr.cuh pr_reset_kernel,
/// pr.hxx:65-93), widened to fp64 for part_pr_begin.
static __global__ void multi_pr_rowsum_kernel(csr_view_t g, int* __restrict__ outdeg, double* __restrict__ outweight) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < g.n_vertices; v += gridDim.x * blockDim.x) {
    const int s = g.row_offsets[v], e = g.row_offsets[v + 1];
    if (outweight) {
      float val = 0.0f;
      for (int k = s; k < e; ++k)
        val = __fadd_rn(val, g.values[k]);
      outweight[v] = static_cast<double>(val);
    } else {
      outdeg[v] = e - s;
    }
  }
}

/// One rank of a multi-device SSSP / PageRank: its rows, its state, its end of the exchange.
struct multi_loop_rank_t {
  int device = 0;
  workspace_t* ws = nullptr;
  dbuf_t<int> ro, ci;
  dbuf_t<float> vals;
  csr_view_t view;
  partition_t pt;
  part_sssp_state_t sssp;
  part_pr_state_t pr;
  dbuf_t<unsigned long long> part_deg;
  dbuf_t<int> msg_out, msg_in;
  dbuf_t<long long> stats;
  thread_exchange_t x;
};

struct multi_loop_cache_t {
  graph_key_t key;
  std::vector<int> devices;
  std::vector<std::unique_ptr<multi_loop_rank_t>> ranks;
  thread_hub_t hub;
  dbuf_t<int> outdeg;        // PageRank: row sums of the whole graph, on the first device
  dbuf_t<double> outweight;
  graph_key_t rowsum_for;
  bool matches(const csr_view_t& g, const std::vector<int>& devs) const {
    return !ranks.empty() && key.matches(g) && devices == devs;
  }
};
struct multi_sssp_cache_t : multi_loop_cache_t {};  // distinct types: one scratch slot each on the context
struct multi_pr_cache_t : multi_loop_cache_t {};

/// Cut `g` (with its values) across the devices of `contexts`; cached per graph identity + device list.
template <typename multi_context_type>
inline void multi_partition(multi_context_type& contexts, multi_loop_cache_t& cache, const csr_view_t& g) {
  const int P = static_cast<int>(contexts.size());
  if (P > kMaxPeers)
    throw std::runtime_error("a multi_context_t run takes at most 16 devices");
  std::vector<int> devs(P);
  for (int r = 0; r < P; ++r)
    devs[r] = contexts.contexts[r]->ordinal();
  if (cache.matches(g, devs))
    return;
  const int home = devs[0];
  auto* ctx0 = contexts.get_context(0);  // also makes the home device current
  workspace_t& ws0 = ctx0->workspace();
  cache.ranks.clear();
  bool distinct = false;
  for (int r = 1; r < P; ++r)
    distinct = distinct || devs[r] != home;
  if (distinct)
    contexts.enable_peer_access();
  for (int r = 0; r < P; ++r) {
    std::unique_ptr<multi_loop_rank_t> R(new multi_loop_rank_t());
    R->device = devs[r];
    R->ws = &contexts.contexts[r]->workspace();
    R->pt = partition_t::make(g.n_vertices, P, r);
    R->view = partition_rows_to(ws0, g, R->pt, devs[r], home, R->ro, R->ci, &R->vals);
    cache.ranks.push_back(std::move(R));
  }
  B2G_CHECK(cudaSetDevice(home));
  cache.key.set(g);
  cache.devices = devs;
}

/// Run `body(rank)` on one host thread per rank (the rank's device current), rethrow the first failure; a failing
/// rank aborts the hub so that its peers leave their barriers.
template <typename F>
inline void multi_run_ranks(multi_loop_cache_t& cache, int home_device, F body) {
  const int P = static_cast<int>(cache.ranks.size());
  cache.hub.reset(P);
  std::vector<std::thread> threads;
  std::vector<std::exception_ptr> errors(P);
  for (int r = 0; r < P; ++r) {
    threads.emplace_back([&, r]() {
      try {
        auto& R = *cache.ranks[r];
        B2G_CHECK(cudaSetDevice(R.device));
        R.x.bind(&cache.hub, r, P);
        body(R);
      } catch (...) {
        errors[r] = std::current_exception();
        cache.hub.abort();
      }
    });
  }
  for (auto& t : threads)
    t.join();
  B2G_CHECK(cudaSetDevice(home_device));
  // the rank that failed first carries the real message; the others only report the abort
  std::exception_ptr first;
  for (auto& e : errors) {
    if (!e)
      continue;
    try {
      std::rethrow_exception(e);
    } catch (const std::exception& ex) {
      if (!first || std::string(ex.what()).find("a peer rank") == std::string::npos)
        first = e;
    } catch (...) {
      first = e;
    }
  }
  if (first)
    std::rethrow_exception(first);
}

/**
 * @brief SSSP from `source` over the devices of `contexts`.  `g` (with values) and `distances` (V floats) are
 * resident on the first context's device.  Returns the iteration count.
 */
template <typename multi_context_type>
inline int sssp_run_multi(multi_context_type& contexts, multi_sssp_cache_t& cache, const csr_view_t& g, int source,
                          float* distances, const advance_launch_t& lcfg, part_sssp_report_t* report) {
  if (source < 0 || source >= g.n_vertices)
    throw std::runtime_error("sssp: source out of range");
  if (!g.values)
    throw std::runtime_error("sssp: the graph has no edge values");
  multi_partition(contexts, cache, g);
  const int P = static_cast<int>(contexts.size());
  const int home = contexts.contexts[0]->ordinal();
  workspace_t& ws0 = contexts.get_context(0)->workspace();
  std::vector<part_sssp_report_t> reports(P);
  multi_run_ranks(cache, home, [&](multi_loop_rank_t& R) {
    part_sssp_run(*R.ws, R.view, R.pt, R.sssp, R.part_deg, R.msg_out, R.msg_in, R.stats, R.x, source, 0, lcfg,
                  &reports[R.pt.part]);
  });
  peer_table_t table{};
  for (int r = 0; r < P; ++r)
    table.p[r] = cache.ranks[r]->sssp.dist.ptr;
  part_interleave_any_kernel<float><<<device_info_t::get().sm_count * 8, 256, 0, ws0.stream>>>(table, P, g.n_vertices,
                                                                                              distances);
  ws0.launches += 1;
  B2G_CHECK(cudaGetLastError());
  if (report)
    *report = reports[0];
  return reports[0].iterations;
}

/**
 * @brief PageRank over the devices of `contexts`.  `g` = the CSR (row sums -> iweights), `in_g` = the in-edge rows
 * (CSC, or the CSR itself for a symmetric graph) that get partitioned, `p` (V floats): all on the first context's
 * device.  Returns the iteration count.
 */
template <typename multi_context_type>
inline int pr_run_multi(multi_context_type& contexts, multi_pr_cache_t& cache, const csr_view_t& g,
                        const csr_view_t& in_g, float alpha, float tol, int max_iter, float* p) {
  multi_partition(contexts, cache, in_g);
  const int P = static_cast<int>(contexts.size());
  const int home = contexts.contexts[0]->ordinal();
  workspace_t& ws0 = contexts.get_context(0)->workspace();
  const int sms = device_info_t::get().sm_count;
  const bool weighted = in_g.values != nullptr && g.values != nullptr;
  if (!cache.rowsum_for.matches(g)) {  // ingest: the whole graph's row sums, read by every rank through peer loads
    if (weighted)
      cache.outweight.ensure(static_cast<size_t>(g.n_vertices) + 16);
    else
      cache.outdeg.ensure(static_cast<size_t>(g.n_vertices) + 16);
    multi_pr_rowsum_kernel<<<sms * 8, 256, 0, ws0.stream>>>(g, weighted ? nullptr : cache.outdeg.ptr,
                                                            weighted ? cache.outweight.ptr : nullptr);
    ws0.launches += 1;
    B2G_CHECK(cudaStreamSynchronize(ws0.stream));
    cache.rowsum_for.set(g);
  }
  std::vector<int> iters(P, 0);
  multi_run_ranks(cache, home, [&](multi_loop_rank_t& R) {
    part_pr_begin(*R.ws, R.view, R.pt, R.pr, alpha, weighted ? nullptr : cache.outdeg.ptr,
                  weighted ? cache.outweight.ptr : nullptr);
    iters[R.pt.part] = part_pr_run(*R.ws, R.pr, R.stats, R.x, alpha, tol, max_iter);
  });
  peer_table_t table{};
  for (int r = 0; r < P; ++r)
    table.p[r] = cache.ranks[r]->pr.p.ptr;
  part_interleave_any_kernel<float><<<sms * 8, 256, 0, ws0.stream>>>(table, P, g.n_vertices, p);
  ws0.launches += 1;
  B2G_CHECK(cudaGetLastError());
  return iters[0];
}

}  // namespace b200
}  // namespace gunrock
