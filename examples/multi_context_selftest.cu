// `bfs::run`, `sssp::run` and `pr::run` over a multi-device `gcuda::multi_context_t` (the surface the reference
// declares in include/gunrock/cuda/context.hxx:146-216 and never runs: its operators throw for size() != 1).
//
//   multi_context_selftest <scale> <device> <device> ...      e.g.  multi_context_selftest 20 0 1 2 3
//
// Builds a symmetric RMAT-like graph of 2^scale vertices and a directed one, runs BFS (push only and
// direction-optimised) through the UNCHANGED call `gunrock::bfs::run(G, param, result, context)` with a
// context of the listed devices, and checks the depths against a host BFS and against the single-device run;
// then SSSP (fp32 distances bit-equal to the single-device run) and PageRank (ranks within 1e-6 relative of it)
// through `sssp::run` / `pr::run` with the same context.
// The same device may be listed several times (ranks then share it: how the single-GPU test box runs this;
// set CUDA_MODULE_LOADING=EAGER for that case -- lazy module loading synchronises the device at a kernel's
// first launch, which dead-locks against another rank's spinning barrier kernel on the SAME device).
// Prints "ALL OK" and returns 0 on success.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <vector>

#include <gunrock/algorithms/bfs.hxx>
#include <gunrock/algorithms/pr.hxx>
#include <gunrock/algorithms/sssp.hxx>

using namespace gunrock;
using namespace memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;
using coo_host_t = format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t>;
using csr_dev_t = format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t>;
using csc_dev_t = format::csc_t<memory_space_t::device, vertex_t, edge_t, weight_t>;

static int failures = 0;
#define CHECK(cond)                                                 \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                   \
    }                                                               \
  } while (0)

static unsigned long long mix(unsigned long long x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// RMAT-like pairs (a=.57 b=.19 c=.19 d=.05), self loops dropped, duplicates kept (the kernels do not care)
static coo_host_t make_graph(int scale, int pairs_per_vertex, bool mirror, unsigned long long seed) {
  const int n = 1 << scale;
  std::vector<int> I, J;
  const long long pairs = 1ll * n * pairs_per_vertex;
  for (long long k = 0; k < pairs; ++k) {
    int u = 0, v = 0;
    unsigned long long h = 0;
    for (int l = 0; l < scale; ++l) {
      if ((l & 3) == 0)
        h = mix(seed + 0x9E3779B97F4A7C15ull * (k * 16 + (l >> 2) + 1));
      unsigned r = (h >> (16 * (l & 3))) & 0xffffu;
      u = (u << 1) | (r >= 49807u);
      v = (v << 1) | ((r >= 37356u && r < 49807u) || r >= 62259u);
    }
    if (u == v)
      continue;
    I.push_back(u);
    J.push_back(v);
    if (mirror) {
      I.push_back(v);
      J.push_back(u);
    }
  }
  coo_host_t coo(n, n, (int)I.size());
  for (size_t k = 0; k < I.size(); ++k) {
    coo.row_indices[k] = I[k];
    coo.column_indices[k] = J[k];
    // a weight per UNORDERED vertex pair: w(u -> v) == w(v -> u), as a symmetric graph's PageRank assumes when it
    // pulls over the CSR itself
    const unsigned long long lo = I[k] < J[k] ? I[k] : J[k], hi = I[k] < J[k] ? J[k] : I[k];
    coo.nonzero_values[k] = 1.0f + static_cast<float>(mix(seed ^ (lo * 0x100000001B3ull + hi)) % 63);
  }
  return coo;
}

static std::vector<int> host_bfs(const thrust::host_vector<int>& ro, const thrust::host_vector<int>& ci, int n,
                                 int src) {
  std::vector<int> d(n, std::numeric_limits<int>::max());
  std::queue<int> q;
  d[src] = 0;
  q.push(src);
  while (!q.empty()) {
    int u = q.front();
    q.pop();
    for (int e = ro[u]; e < ro[u + 1]; ++e)
      if (d[ci[e]] == std::numeric_limits<int>::max()) {
        d[ci[e]] = d[u] + 1;
        q.push(ci[e]);
      }
  }
  return d;
}

template <typename graph_type>
static void check_runs(graph_type& G, const thrust::host_vector<int>& ro, const thrust::host_vector<int>& ci,
                       std::shared_ptr<gcuda::multi_context_t> multi, std::shared_ptr<gcuda::multi_context_t> single,
                       const char* what) {
  const int n = G.get_number_of_vertices();
  int hub = 0;
  for (int v = 1; v < n; ++v)
    if (ro[v + 1] - ro[v] > ro[hub + 1] - ro[hub])
      hub = v;
  for (int src : {hub, 1, n - 1}) {
    auto want = host_bfs(ro, ci, n, src);
    for (auto dir : {operators::advance_direction_t::forward, operators::advance_direction_t::optimized})
      for (auto lb : {operators::load_balance_t::block_mapped, operators::load_balance_t::merge_path}) {
        options_t opt;
        opt.advance_direction = dir;
        opt.advance_load_balance = lb;
        thrust::device_vector<int> dist(n, -7), pred(n), dist1(n, -7);
        bfs::param_t<int> param(src, opt);
        bfs::result_t<int> result(dist.data().get(), pred.data().get());
        float ms = bfs::run(G, param, result, multi);
        bfs::result_t<int> result1(dist1.data().get(), pred.data().get());
        bfs::run(G, param, result1, single);
        thrust::host_vector<int> h(dist), h1(dist1);
        long long bad = 0, bad1 = 0;
        for (int v = 0; v < n; ++v) {
          bad += h[v] != want[v];
          bad1 += h1[v] != want[v];
        }
        std::printf("%s src %d %s %s: %zu devices %.3f ms, wrong depths: %lld (single device: %lld)\n", what, src,
                    dir == operators::advance_direction_t::forward ? "push" : "optimised",
                    lb == operators::load_balance_t::merge_path ? "merge_path" : "block_mapped", multi->size(), ms, bad,
                    bad1);
        CHECK(bad == 0 && bad1 == 0);
      }
  }
}

/// SSSP and PageRank: the multi-device run against the single-device run of the same call.
template <typename graph_type>
static void check_sssp_pr(graph_type& G, const thrust::host_vector<int>& ro,
                          std::shared_ptr<gcuda::multi_context_t> multi, std::shared_ptr<gcuda::multi_context_t> single,
                          const char* what) {
  const int n = G.get_number_of_vertices();
  int hub = 0;
  for (int v = 1; v < n; ++v)
    if (ro[v + 1] - ro[v] > ro[hub + 1] - ro[hub])
      hub = v;
  for (int src : {hub, n - 1})
    for (auto lb : {operators::load_balance_t::block_mapped, operators::load_balance_t::merge_path}) {
      options_t opt;
      opt.advance_load_balance = lb;
      thrust::device_vector<float> dist(n, -7.0f), dist1(n, -7.0f);
      thrust::device_vector<int> pred(n);
      sssp::param_t<int> param(src, opt);
      sssp::result_t<int, float> result(dist.data().get(), pred.data().get(), n);
      float ms = sssp::run(G, param, result, multi);
      sssp::result_t<int, float> result1(dist1.data().get(), pred.data().get(), n);
      sssp::run(G, param, result1, single);
      thrust::host_vector<float> h(dist), h1(dist1);
      long long bad = 0, reached = 0;
      for (int v = 0; v < n; ++v) {
        unsigned a, b;
        std::memcpy(&a, &h[v], 4);
        std::memcpy(&b, &h1[v], 4);
        bad += a != b;
        reached += h1[v] < std::numeric_limits<float>::max();
      }
      std::printf("%s sssp src %d %s: %zu devices %.3f ms, %lld reached, distances differing from the single-device run: %lld\n",
                  what, src, lb == operators::load_balance_t::merge_path ? "merge_path" : "block_mapped", multi->size(), ms,
                  reached, bad);
      CHECK(bad == 0 && reached >= 1);
    }
  {
    thrust::device_vector<float> p(n, -1.0f), p1(n, -1.0f);
    float ms = pr::run(G, 0.85f, 1e-6f, p.data().get(), multi);
    pr::run(G, 0.85f, 1e-6f, p1.data().get(), single);
    thrust::host_vector<float> h(p), h1(p1);
    double worst = 0.0, sum = 0.0;
    for (int v = 0; v < n; ++v) {
      const double ref = h1[v] > 0 ? h1[v] : 1e-30;
      const double rel = std::fabs(static_cast<double>(h[v]) - h1[v]) / ref;
      worst = rel > worst ? rel : worst;
      sum += h[v];
    }
    std::printf("%s pagerank: %zu devices %.3f ms, sum of ranks %.6f, worst relative difference to the single-device run %.3g\n",
                what, multi->size(), ms, sum, worst);
    CHECK(worst <= 1e-6 && std::fabs(sum - 1.0) < 1e-3);
  }
}

int main(int argc, char** argv) {
  const int scale = argc > 1 ? std::atoi(argv[1]) : 14;
  thrust::host_vector<gcuda::device_id_t> devices;
  for (int i = 2; i < argc; ++i)
    devices.push_back(std::atoi(argv[i]));
  if (devices.empty()) {
    devices.push_back(0);
    devices.push_back(0);
  }
  auto multi = std::make_shared<gcuda::multi_context_t>(devices);
  auto single = std::make_shared<gcuda::multi_context_t>(devices[0]);
  CHECK(multi->size() == devices.size());
  multi->enable_peer_access();
  cudaSetDevice(devices[0]);
  {  // symmetric graph: the ranks' CSR rows double as their in-edge lists
    auto coo = make_graph(scale, 8, true, 0x5EED);
    csr_dev_t csr;
    csr.from_coo(coo);
    thrust::host_vector<int> ro(csr.row_offsets), ci(csr.column_indices);
    graph::graph_properties_t props;
    props.symmetric = true;
    auto G = graph::build<memory_space_t::device>(props, csr);
    check_runs(G, ro, ci, multi, single, "symmetric");
    check_sssp_pr(G, ro, multi, single, "symmetric");
  }
  {  // directed graph with a CSC: both are cut across the devices
    auto coo = make_graph(scale > 12 ? scale - 2 : scale, 12, false, 0xD1CE);
    csr_dev_t csr;
    csr.from_coo(coo);
    csc_dev_t csc;
    csc.from_csr(csr);
    thrust::host_vector<int> ro(csr.row_offsets), ci(csr.column_indices);
    graph::graph_properties_t props;
    props.directed = true;
    auto G = graph::build<memory_space_t::device>(props, csr, csc);
    check_runs(G, ro, ci, multi, single, "directed+csc");
    check_sssp_pr(G, ro, multi, single, "directed+csc");
  }
  if (failures == 0)
    std::printf("ALL OK\n");
  return failures ? 1 : 0;
}
