// tests/cuemu/emu_more.cpp -- TEST INFRASTRUCTURE: more product kernels under the CPU emulator (see cuemu.h):
//   * lookback_scan_kernel (scan.cuh): single-pass exclusive scan / stable select with decoupled look-back;
//   * one PageRank iteration (pr.cuh): reset, prepare (deterministic dangling sum, last-CTA fold), the TMA-tiled
//     pull kernel and the fix-up of rows that cross tiles -- against a float64 evaluation of the same formula.
// Usage: emu_more <seed>; prints "EMU OK <checks>".
#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include <cuemu.h>

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>

#include "scan_kernels.gen.cuh"
#include "pr_kernels.gen.cuh"

using namespace gunrock::b200;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                 \
  do {                                                              \
    ++checks;                                                       \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                   \
    }                                                               \
  } while (0)

static void check_scan(std::mt19937& rng) {
  for (int n : {0, 1, 31, 2048, 2049, 9000}) {
    std::vector<int> in(n), excl(n + 1, -1), kept(n + 1, -1);
    for (auto& x : in)
      x = static_cast<int>(rng() % 5);
    const int* src = in.data();
    int* out_excl = excl.data();
    int* out_kept = kept.data();
    int total = -1;
    std::vector<unsigned long long> state(n / (64 * 8) + 8, 0ull);
    ctrl_t ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    const int n_copy = n;
    auto value = [=](int i) -> int { return src[i]; };
    auto emit = [=](int i, int e, int) { out_excl[i] = e; };
    cuemu::launch(3, 64, 0, 1, [&] {
      lookback_scan_kernel<64, 8>(&n_copy, 0, value, emit, &total, out_excl, state.data(), &ctrl, 7u); });
    std::vector<int> ref(n + 1, 0);
    for (int i = 0; i < n; ++i)
      ref[i + 1] = ref[i] + in[i];
    bool ok = total == ref[n];
    for (int i = 0; i <= n && ok; ++i)
      ok = excl[i] == ref[i];  // excl[n] = grand total (total_at_n)
    CHECK(ok);
    // stable select of the odd values, second launch on the SAME status words with the next epoch
    std::memset(&ctrl, 0, sizeof ctrl);
    int n_kept = -1;
    auto keep = [=](int i) -> int { return src[i] & 1; };
    auto put = [=](int i, int e, int k) {
      if (k)
        out_kept[e] = i;
    };
    cuemu::launch(2, 64, 0, 1, [&] {
      lookback_scan_kernel<64, 8>(&n_copy, 0, keep, put, &n_kept, nullptr, state.data(), &ctrl, 8u); });
    std::vector<int> want;
    for (int i = 0; i < n; ++i)
      if (in[i] & 1)
        want.push_back(i);
    CHECK(n_kept == static_cast<int>(want.size()) &&
          std::equal(want.begin(), want.end(), kept.begin()));
  }
  std::printf("look-back scan ok\n");
}

static void check_pagerank_iteration(std::mt19937& rng) {
  // CSC of a random directed graph with a few very long in-rows (they cross 2048-edge tiles) and dangling vertices
  const int V = 3000;
  std::vector<std::vector<int>> in_nbrs(V);
  std::vector<int> outdeg(V, 0);
  auto add = [&](int u, int v) {
    in_nbrs[v].push_back(u);
    ++outdeg[u];
  };
  for (int e = 0; e < 20000; ++e) {
    int u = static_cast<int>(rng() % V), v = static_cast<int>(rng() % V);
    if (u % 10 == 9)
      continue;  // dangling sources: no out-edges
    add(u, v);
  }
  for (int v : {5, 1700})  // hub destinations: 5000 in-edges each
    for (int k = 0; k < 5000; ++k) {
      int u = static_cast<int>(rng() % V);
      if (u % 10 != 9)
        add(u, v);
    }
  std::vector<int> t_ro(V + 1, 0), t_ci;
  for (int v = 0; v < V; ++v) {
    t_ro[v + 1] = t_ro[v] + static_cast<int>(in_nbrs[v].size());
    t_ci.insert(t_ci.end(), in_nbrs[v].begin(), in_nbrs[v].end());
  }
  const int E = static_cast<int>(t_ci.size());
  for (int i = 0; i < 16; ++i)
    t_ci.push_back(0);
  std::vector<int> g_ro(V + 1, 0);  // only the out-degrees matter for the reset kernel (unweighted)
  for (int v = 0; v < V; ++v)
    g_ro[v + 1] = g_ro[v] + outdeg[v];
  csr_view_t g, t;
  g.n_vertices = V;
  g.n_edges = E;
  g.row_offsets = g_ro.data();
  t.n_vertices = V;
  t.n_edges = E;
  t.row_offsets = t_ro.data();
  t.column_indices = t_ci.data();
  const float alpha = 0.85f;
  std::vector<float> p(V), plast(V), iw(V), c(V);
  cuemu::launch(2, 64, 0, 1, [&] { pr_reset_kernel(g, alpha, p.data(), plast.data(), iw.data()); });
  const int ntiles = std::max(1, (E + kPrTile - 1) / kPrTile);
  std::vector<int> first_owned(ntiles + 4), tail_row(ntiles + 4, -9);
  std::vector<double> head(ntiles + 4, 0.0), tail(ntiles + 4, 0.0), partials(kPrPartials, 0.0);
  cuemu::launch(1, 64, 0, 1, [&] { pr_tile_table_kernel(t_ro.data(), V, ntiles, first_owned.data()); });
  for (int iter = 0; iter < 2; ++iter) {
    unsigned err_bits[2] = {0u, 0u};
    float base = -1.0f;
    cuemu::launch(4, 128, 0, 1, [&] {
      pr_prepare_kernel<128>(V, alpha, p.data(), iw.data(), plast.data(), c.data(), partials.data(), err_bits + 1, &base); });
    // float64 evaluation of the same formula on the same inputs
    double dsum = 0.0;
    for (int v = 0; v < V; ++v)
      if (iw[v] == 0.0f)
        dsum += static_cast<double>(alpha * plast[v]);
    const float base_ref = ((1.0f - alpha) + static_cast<float>(dsum)) / static_cast<float>(V);
    CHECK(std::fabs(base - base_ref) <= 1e-6f * base_ref);
    std::vector<float> want(V);
    float err_ref = 0.0f;
    for (int v = 0; v < V; ++v) {
      double acc = 0.0;
      for (int e = t_ro[v]; e < t_ro[v + 1]; ++e)
        acc += static_cast<double>(c[t_ci[e]]);
      want[v] = static_cast<float>(static_cast<double>(base) + acc);
      err_ref = std::max(err_ref, std::fabs(want[v] - plast[v]));
    }
    ctrl_t ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    cuemu::launch(3, 256, 0, 1, [&] {
      pr_pull_tile_kernel<256, false>(t, ntiles, first_owned.data(), c.data(), plast.data(), &base, p.data(),
                                      head.data(), tail.data(), tail_row.data(), err_bits, &ctrl); });
    cuemu::launch(2, 64, 0, 1, [&] {
      pr_fixup_kernel(t, ntiles, tail_row.data(), head.data(), tail.data(), &base, plast.data(), p.data(), err_bits); });
    bool ok = true;
    double sum = 0.0;
    for (int v = 0; v < V; ++v) {
      ok = ok && std::fabs(p[v] - want[v]) <= 1e-6f * std::fabs(want[v]);
      sum += p[v];
    }
    CHECK(ok);
    float err;
    std::memcpy(&err, err_bits, 4);
    CHECK(std::fabs(err - err_ref) <= 1e-6f * err_ref + 1e-12f);
    CHECK(std::fabs(sum - 1.0) < 1e-3);
    std::printf("pagerank iteration %d: sum %.6f, err %.3e, rows crossing tiles folded by the fix-up\n", iter, sum, err);
  }
}

int main(int argc, char** argv) {
  std::mt19937 rng(argc > 1 ? std::atoi(argv[1]) : 1);
  check_scan(rng);
  check_pagerank_iteration(rng);
  if (failures == 0)
    std::printf("EMU OK %d\n", checks);
  return failures == 0 ? 0 : 1;
}
