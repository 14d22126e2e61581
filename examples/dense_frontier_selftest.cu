// Self-test of the dense frontier views: frontier_t<..., bitmap> and <..., boolmap>
// -- host operations, device accessors, conversion from / to the vector view, and one BFS level written
// with a vector input frontier and a bitmap output frontier on the raw advance operator.
// Prints "ALL OK" and returns 0 when every check passes.
#include <algorithm>
#include <cstdio>
#include <set>
#include <vector>

#include <gunrock/algorithms/algorithms.hxx>

using namespace gunrock;
using namespace memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;
using vector_frontier_t = frontier::frontier_t<vertex_t, edge_t>;
using bitmap_frontier_t =
    frontier::frontier_t<vertex_t, edge_t, frontier::frontier_kind_t::vertex_frontier, frontier::frontier_view_t::bitmap>;
using boolmap_frontier_t =
    frontier::frontier_t<vertex_t, edge_t, frontier::frontier_kind_t::vertex_frontier, frontier::frontier_view_t::boolmap>;

static int failures = 0;
#define CHECK(cond)                                                 \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                   \
    }                                                               \
  } while (0)

template <typename dense_t>
__global__ void probe_kernel(dense_t f, int universe, int* present, int* via_get) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < universe; v += gridDim.x * blockDim.x) {
    present[v] = f.contains(v) ? 1 : 0;
    via_get[v] = f.get_element_at(v);
  }
}
template <typename dense_t>
__global__ void insert_kernel(dense_t f, const int* ids, int n, int* fresh) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (f.set_element_at(ids[i]))
      atomicAdd(fresh, 1);
}
template <typename dense_t>
__global__ void remove_kernel(dense_t f, const int* ids, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    f.remove_element(ids[i]);
}

static std::vector<int> sorted_host(vector_frontier_t& f) {
  std::vector<int> h(f.get_number_of_elements());
  if (!h.empty())
    cudaMemcpy(h.data(), f.data(), h.size() * sizeof(int), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  return h;
}

template <typename dense_t>
void exercise(const char* name, cudaStream_t stream) {
  for (int universe : {1, 31, 32, 33, 1000, 4099}) {
    dense_t f(universe);
    f.bind_stream(stream);
    CHECK(f.get_universe() == (std::size_t)universe && f.is_empty());
    // host inserts
    std::set<int> expect;
    for (int v : {0, universe / 2, universe - 1}) {
      f.push_back(v);
      expect.insert(v);
    }
    CHECK(f.get_number_of_elements() == expect.size());
    // device inserts with duplicates: `fresh` counts first insertions only (bitmap: atomic, exact)
    std::vector<int> ids;
    for (int k = 0; k < 300; ++k)
      ids.push_back((k * 37) % universe);
    int *d_ids, *d_fresh;
    cudaMalloc(&d_ids, ids.size() * sizeof(int));
    cudaMalloc(&d_fresh, sizeof(int));
    cudaMemcpyAsync(d_ids, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(d_fresh, 0, sizeof(int), stream);
    std::size_t before = expect.size();
    for (int v : ids)
      expect.insert(v);
    insert_kernel<<<4, 128, 0, stream>>>(f, d_ids, (int)ids.size(), d_fresh);
    int fresh = -1;
    cudaMemcpyAsync(&fresh, d_fresh, sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    if (f.get_view() == frontier::frontier_view_t::bitmap)
      CHECK(fresh == (int)(expect.size() - before));
    CHECK(f.get_number_of_elements() == expect.size());
    // device probes
    int *d_present, *d_get;
    cudaMalloc(&d_present, universe * sizeof(int));
    cudaMalloc(&d_get, universe * sizeof(int));
    probe_kernel<<<8, 128, 0, stream>>>(f, universe, d_present, d_get);
    std::vector<int> present(universe), got(universe);
    cudaMemcpyAsync(present.data(), d_present, universe * sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(got.data(), d_get, universe * sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    bool ok = true;
    for (int v = 0; v < universe; ++v) {
      bool in = expect.count(v) != 0;
      ok = ok && present[v] == (in ? 1 : 0) && got[v] == (in ? v : -1);
    }
    CHECK(ok);
    // dense -> vector -> dense round trip
    vector_frontier_t vec;
    vec.bind_stream(stream);
    f.to_vector(vec);
    CHECK(sorted_host(vec) == std::vector<int>(expect.begin(), expect.end()));
    dense_t g(universe);
    g.bind_stream(stream);
    g.from_vector(vec);
    CHECK(g.get_number_of_elements() == expect.size());
    // removal, fill, clear
    remove_kernel<<<4, 128, 0, stream>>>(g, d_ids, (int)ids.size());
    std::set<int> left = expect;
    for (int v : ids)
      left.erase(v);
    CHECK(g.get_number_of_elements() == left.size());
    g.fill(1);
    CHECK(g.get_number_of_elements() == (std::size_t)universe);
    g.fill(0);
    CHECK(g.is_empty());
    bool threw = false;
    try {
      g.fill(2);
    } catch (std::exception&) {
      threw = true;
    }
    CHECK(threw);
    // copies alias the storage (kernel arguments are copies)
    dense_t alias = f;
    alias.clear();
    CHECK(f.is_empty());
    cudaFree(d_ids);
    cudaFree(d_fresh);
    cudaFree(d_present);
    cudaFree(d_get);
  }
  std::printf("%s view ok\n", name);
}

int main() {
  auto context = std::make_shared<gcuda::multi_context_t>(0);
  cudaStream_t stream = context->get_context(0)->stream();
  exercise<bitmap_frontier_t>("bitmap", stream);
  exercise<boolmap_frontier_t>("boolmap", stream);

  // ---- one BFS level: vector input frontier, bitmap "visited" + bitmap output frontier ----------
  const int n = 2000;
  std::vector<int> I, J;
  for (int v = 0; v < n; ++v)
    for (int t : {(v * 3 + 1) % n, (v * 7 + 5) % n, (v + 1) % n}) {
      I.push_back(v);
      J.push_back(t);
    }
  format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t> coo(n, n, (int)I.size());
  for (size_t k = 0; k < I.size(); ++k) {
    coo.row_indices[k] = I[k];
    coo.column_indices[k] = J[k];
    coo.nonzero_values[k] = 1.0f;
  }
  format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t> csr;
  csr.from_coo(coo);
  auto G = graph::build<memory_space_t::device>(graph::graph_properties_t(), csr);

  vector_frontier_t in, scratch_out;
  in.bind_stream(stream);
  std::set<int> frontier_set = {0, 17, 999};
  for (int v : frontier_set)
    in.push_back(v);
  bitmap_frontier_t visited(n), next(n);
  visited.bind_stream(stream);
  next.bind_stream(stream);
  visited.from_vector(in);
  auto claim = [visited, next] __device__(vertex_t const& src, vertex_t const& dst, edge_t const& e,
                                         weight_t const& w) -> bool {
    if (visited.set_element_at(dst))  // first visit: goes to the next (dense) frontier
      next.set_element_at(dst);
    return false;                     // nothing for the vector output
  };
  std::vector<int> unused_segments;
  operators::advance::execute<operators::load_balance_t::block_mapped, operators::advance_direction_t::forward,
                              operators::advance_io_type_t::vertices, operators::advance_io_type_t::none>(
      G, claim, &in, &scratch_out, unused_segments, *context);
  std::set<int> expect_next;
  for (int v : frontier_set)
    for (int t : {(v * 3 + 1) % n, (v * 7 + 5) % n, (v + 1) % n})
      if (!frontier_set.count(t))
        expect_next.insert(t);
  vector_frontier_t out_vec;
  out_vec.bind_stream(stream);
  next.to_vector(out_vec);
  CHECK(sorted_host(out_vec) == std::vector<int>(expect_next.begin(), expect_next.end()));
  CHECK(visited.get_number_of_elements() == frontier_set.size() + expect_next.size());

  if (failures == 0)
    std::printf("ALL OK\n");
  return failures == 0 ? 0 : 1;
}
