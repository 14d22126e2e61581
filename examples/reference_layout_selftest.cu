// Self-test of the opt-in REFERENCE OUTPUT LAYOUT of operators::advance::execute
// (standard_context_t::reference_advance_output / -DGUNROCK_B200_REFERENCE_ADVANCE_OUTPUT): the output frontier
// has one slot per (input entry, out-edge) in edge-rank order, -1 where the functor returned false, and as many
// elements as the input frontier's out-degree sum -- what the reference's advance writes
// (/root/reference/include/gunrock/framework/operators/advance/merge_path.hxx:218-279) -- instead of the compact
// frontier that is this library's default.  Checked for every load balancer, vertex and edge outputs, an input
// with an invalid slot and a duplicate, the whole graph as the input, and a filter consuming the result.
// Prints "ALL OK" and returns 0 when every check passes.
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <set>
#include <vector>

#include <gunrock/algorithms/algorithms.hxx>

using namespace gunrock;
using namespace memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                       \
    }                                                                   \
  } while (0)

template <typename frontier_t>
std::vector<int> to_host(frontier_t& f) {
  std::vector<int> h(f.get_number_of_elements());
  if (!h.empty())
    cudaMemcpy(h.data(), f.data(), h.size() * sizeof(int), cudaMemcpyDeviceToHost);
  return h;
}

int main() {
  // ---- a small directed graph with a hub, duplicates of nothing, one isolated vertex ----------
  const int n = 600;
  std::vector<int> I, J;
  std::vector<float> W;
  for (int v = 1; v < 500; ++v) {  // hub 0 -> 1..499
    I.push_back(0);
    J.push_back(v);
    W.push_back(1.0f + (v % 7));
  }
  for (int v = 1; v < 599; ++v)  // chain v -> v+1, and v -> (7v mod 599)
    for (int t : {v + 1, (7 * v) % 599}) {
      I.push_back(v);
      J.push_back(t);
      W.push_back(0.5f + (v % 5));
    }
  format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t> coo(n, n, (int)I.size());
  for (size_t k = 0; k < I.size(); ++k) {
    coo.row_indices[k] = I[k];
    coo.column_indices[k] = J[k];
    coo.nonzero_values[k] = W[k];
  }
  format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t> csr;
  csr.from_coo(coo);
  graph::graph_properties_t props;
  props.directed = true;
  auto G = graph::build<memory_space_t::device>(props, csr);
  thrust::host_vector<int> ro(csr.row_offsets), ci(csr.column_indices);
  CHECK(G.get_number_of_vertices() == n && G.get_number_of_edges() == (int)I.size());

  auto context = std::make_shared<gcuda::multi_context_t>(0);
  auto& ctx = *context;
  using frontier_t = frontier::frontier_t<vertex_t, edge_t>;

  thrust::device_vector<int> marks(n);
  int* m = marks.data().get();
  auto keep_not_mult_of_3 = [m] __host__ __device__(vertex_t const& s, vertex_t const& d,
                                                    edge_t const& e, weight_t const& w) -> bool {
    math::atomic::add(m + d, 1);
    return d % 3 != 0;
  };
  thrust::device_vector<edge_t> segments;
  CHECK(!ctx.get_context(0)->reference_advance_output());  // off unless asked for

  // ---- advance with the reference's output layout (opt-in): one slot per edge rank, -1 = rejected ---
  {
    ctx.get_context(0)->reference_advance_output(true);
    std::vector<int> in_r = {0, 5, -1, 17, 5, 598, 599};  // an invalid slot and a duplicate in the input
    std::vector<int> exp_v, exp_e;
    for (int v : in_r) {
      if (v < 0)
        continue;
      for (int e = ro[v]; e < ro[v + 1]; ++e) {
        exp_v.push_back(ci[e] % 3 != 0 ? ci[e] : -1);
        exp_e.push_back(ci[e] % 3 != 0 ? e : -1);
      }
    }
    auto all = [] __host__ __device__(vertex_t const& v) -> bool { return true; };
    auto run_ranked = [&](auto lb_tag) {
      constexpr operators::load_balance_t lb = decltype(lb_tag)::value;
      thrust::fill(marks.begin(), marks.end(), 0);
      frontier_t in, out, out_e;
      for (int v : in_r)
        in.push_back(v);
      operators::advance::execute<lb, operators::advance_direction_t::forward,
                                  operators::advance_io_type_t::vertices,
                                  operators::advance_io_type_t::vertices>(G, keep_not_mult_of_3, &in, &out,
                                                                          segments, ctx);
      CHECK(to_host(out) == exp_v);  // position = edge rank (merge_path.hxx:218-279)
      operators::advance::execute<lb, operators::advance_direction_t::forward,
                                  operators::advance_io_type_t::vertices,
                                  operators::advance_io_type_t::edges>(G, keep_not_mult_of_3, &in, &out_e,
                                                                       segments, ctx);
      CHECK(to_host(out_e) == exp_e);
      // the usual follow-up: a filter drops the invalid slots
      frontier_t kept;
      operators::filter::execute<operators::filter_algorithm_t::predicated>(G, all, &out, &kept, ctx);
      std::vector<int> valid;
      for (int x : exp_v)
        if (x >= 0)
          valid.push_back(x);
      CHECK(to_host(kept) == valid);
    };
    run_ranked(std::integral_constant<operators::load_balance_t, operators::load_balance_t::thread_mapped>{});
    run_ranked(std::integral_constant<operators::load_balance_t, operators::load_balance_t::block_mapped>{});
    run_ranked(std::integral_constant<operators::load_balance_t, operators::load_balance_t::merge_path>{});
    {  // whole graph as the input: E slots, slot e belongs to edge e
      frontier_t in, out;
      operators::advance::execute<operators::load_balance_t::merge_path,
                                  operators::advance_direction_t::forward,
                                  operators::advance_io_type_t::graph,
                                  operators::advance_io_type_t::vertices>(G, keep_not_mult_of_3, &in, &out,
                                                                          segments, ctx);
      auto h = to_host(out);
      bool ok = h.size() == ci.size();
      for (size_t e = 0; ok && e < h.size(); ++e)
        ok = h[e] == (ci[e] % 3 != 0 ? ci[e] : -1);
      CHECK(ok);
    }
    ctx.get_context(0)->reference_advance_output(false);
  }

  if (failures == 0)
    std::printf("ALL OK\n");
  return failures == 0 ? 0 : 1;
}
