// Self-test of the per-graph caches behind the fused enactors ("stale caches keyed by raw
// device pointers"): ONE context runs direction-optimised BFS and PageRank on two DIFFERENT directed graphs
// of identical size back to back, the second graph refilled into the SAME csr_t / csc_t objects (so every
// device address is reused), then a third time after the context was destroyed and recreated.  Every
// result is compared with a host computation in this file.  Prints "ALL OK" and returns 0 on success.
#include <cmath>
#include <cstdio>
#include <queue>
#include <vector>

#include <gunrock/algorithms/bfs.hxx>
#include <gunrock/algorithms/pr.hxx>

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

// Directed graph on n vertices with exactly 3 out-edges per vertex < n - 40 (so that two seeds give the
// same V and E): v -> (a v + 1) % m, (b v + 7) % m, hub-ish v -> v % 17.  The last 40 vertices have no
// out-edges (dangling for PageRank); vertices >= m have no in-edges (pre-marked by the pull sweep).
static coo_host_t make_graph(int n, int a, int b) {
  const int m = n - 25, rows = n - 40;
  coo_host_t coo(n, n, rows * 3);
  int k = 0;
  for (int v = 0; v < rows; ++v)
    for (int t : {static_cast<int>((1ll * a * v + 1) % m), static_cast<int>((1ll * b * v + 7) % m), v % 17}) {
      coo.row_indices[k] = v;
      coo.column_indices[k] = t;
      coo.nonzero_values[k] = 1.0f;
      ++k;
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

// include/gunrock/algorithms/pr.hxx:65-195 restated in double precision (checked at 1e-4 relative here:
// this test is about WHICH graph was pulled, the 1e-6 parity tests live in tests/).
static std::vector<double> host_pr(const thrust::host_vector<int>& ro, const thrust::host_vector<int>& ci, int n) {
  const double alpha = 0.85, tol = 1e-6;
  std::vector<double> p(n, 1.0 / n), last(n);
  for (int it = 0; it < 1000; ++it) {
    last = p;
    double dangling = 0;
    for (int v = 0; v < n; ++v)
      if (ro[v + 1] == ro[v])
        dangling += alpha * last[v];
    for (int v = 0; v < n; ++v)
      p[v] = (1 - alpha + dangling) / n;
    for (int u = 0; u < n; ++u)
      for (int e = ro[u]; e < ro[u + 1]; ++e)
        p[ci[e]] += alpha * last[u] / (ro[u + 1] - ro[u]);
    double err = 0;
    for (int v = 0; v < n; ++v)
      err = std::max(err, std::fabs(p[v] - last[v]));
    if (err < tol)
      break;
  }
  return p;
}

static void run_and_check(csr_dev_t& csr, csc_dev_t& csc, std::shared_ptr<gcuda::multi_context_t> context,
                          const char* what) {
  const int n = csr.number_of_rows;
  thrust::host_vector<int> ro(csr.row_offsets), ci(csr.column_indices);
  graph::graph_properties_t props;
  props.directed = true;
  // BFS, direction-optimised, on a csr + csc graph (the pull sweep keeps a per-graph map of vertices without
  // in-edges); PageRank on the csr-only graph (the enactor builds and caches the transpose + tile table)
  auto G2 = graph::build<memory_space_t::device>(props, csr, csc);
  auto G1 = graph::build<memory_space_t::device>(props, csr);
  for (int src : {0, 5}) {
    thrust::device_vector<int> dist(n), pred(n);
    options_t opt;
    opt.advance_direction = operators::advance_direction_t::optimized;
    bfs::param_t<int> param(src, opt);
    bfs::result_t<int> result(dist.data().get(), pred.data().get());
    bfs::run(G2, param, result, context);
    thrust::host_vector<int> h(dist);
    auto want = host_bfs(ro, ci, n, src);
    int bad = 0;
    for (int v = 0; v < n; ++v)
      bad += h[v] != want[v];
    if (bad)
      std::printf("%s: bfs from %d: %d wrong depths\n", what, src, bad);
    CHECK(bad == 0);
  }
  {
    thrust::device_vector<float> p(n);
    pr::param_t<float> param(0.85f, 1e-6f);
    pr::result_t<float> result(p.data().get());
    pr::run(G1, param, result, context);
    thrust::host_vector<float> h(p);
    auto want = host_pr(ro, ci, n);
    int bad = 0;
    for (int v = 0; v < n; ++v)
      bad += std::fabs(h[v] - want[v]) > 1e-4 * want[v];
    if (bad)
      std::printf("%s: pagerank: %d ranks off\n", what, bad);
    CHECK(bad == 0);
  }
}

int main() {
  const int n = 200000;
  csr_dev_t csr;
  csc_dev_t csc;
  auto context = std::make_shared<gcuda::multi_context_t>(0);
  const void* first_offsets = nullptr;
  int reused = 0;
  const int seeds[3][2] = {{3, 11}, {5, 13}, {7, 19}};
  for (int round = 0; round < 3; ++round) {
    auto coo = make_graph(n, seeds[round][0], seeds[round][1]);
    csr.from_coo(coo);  // refills the same object: thrust keeps the allocation when the size is unchanged
    csc.from_csr(csr);
    if (round == 0)
      first_offsets = csr.row_offsets.data().get();
    else
      reused += first_offsets == csr.row_offsets.data().get();
    if (round == 2)  // a NEW context at (possibly) the address of the old one must start from empty scratch
      context = std::make_shared<gcuda::multi_context_t>(0);
    run_and_check(csr, csc, context, round == 0 ? "graph A" : round == 1 ? "graph B (same addresses)" : "graph C (new context)");
  }
  std::printf("device addresses reused by %d of 2 refills\n", reused);
  if (failures == 0)
    std::printf("ALL OK\n");
  return failures ? 1 : 0;
}
