// tests/cuemu/emu_more.cpp -- TEST INFRASTRUCTURE: more product kernels under the CPU emulator (see cuemu.h):
//   * lookback_scan_kernel (scan.cuh): single-pass exclusive scan / stable select with decoupled look-back;
//   * one PageRank iteration (pr.cuh): reset, prepare (deterministic dangling sum, last-CTA fold), the TMA-tiled
//     pull kernel and the fix-up of rows that cross tiles -- against a float64 evaluation of the same formula;
//   * the pull side of the direction-optimised BFS (bfs.cuh): reset, map of vertices without in-edges, queue <->
//     bitmap, one bottom-up sweep over the visited words and one over the list of still-unvisited vertices.
// Usage: emu_more <seed>; prints "EMU OK <checks>".
#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include <cuemu.h>

#include <climits>
#include <set>

#include "advance_kernels.gen.cuh"
#include "scan_kernels.gen.cuh"
#include "pr_kernels.gen.cuh"
#include "bfs_kernels.gen.cuh"

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

/// The pull side of the direction-optimised BFS (bfs.cuh): queue <-> bitmap conversion, the map of vertices
/// without in-edges, one bottom-up sweep over every visited word and one over the list of still-unvisited vertices.
static void check_bottom_up(std::mt19937& rng) {
  const int V = 5000;
  std::vector<std::vector<int>> adj(V);
  for (int e = 0; e < 7000; ++e) {  // a sparse, skewed core on the ids below 3000 ...
    int a = static_cast<int>(3000 * std::pow((rng() % 100000) / 100000.0, 3.0)), b = static_cast<int>(rng() % 3000);
    if (a == b || a % 9 == 4 || b % 9 == 4)
      continue;  // vertices = 4 mod 9 stay isolated
    adj[a].push_back(b);
    adj[b].push_back(a);
  }
  int prev = 0;  // ... and a comb hanging off the source: discovered a few vertices per level, for many levels
  for (int v = 3000; v < V; ++v) {
    if (v % 9 == 4)
      continue;
    int parent = (v % 3 == 0) ? prev : std::max(0, v - 40);
    if (parent % 9 == 4)
      parent = prev;
    adj[v].push_back(parent);
    adj[parent].push_back(v);
    prev = v;
  }
  std::vector<int> ro(V + 1, 0), ci;
  for (int v = 0; v < V; ++v) {
    std::sort(adj[v].begin(), adj[v].end());
    adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
    ro[v + 1] = ro[v] + static_cast<int>(adj[v].size());
    ci.insert(ci.end(), adj[v].begin(), adj[v].end());
  }
  ci.resize(ci.size() + 16, 0);
  csr_view_t g;
  g.n_vertices = V;
  g.n_edges = ro[V];
  g.row_offsets = ro.data();
  g.column_indices = ci.data();
  const int words = (V + 31) / 32;
  // map of vertices without in-edges (symmetric graph: the isolated ones, plus the tail bits past V)
  std::vector<unsigned> dead(words + 4, 0u);
  cuemu::launch(2, 64, 0, 1, [&] { bfs_unreachable_map_kernel(ro.data(), V, dead.data()); });
  bool ok = true;
  for (int v = 0; v < words * 32; ++v)
    ok = ok && (((dead[v >> 5] >> (v & 31)) & 1u) == (v >= V || adj[v].empty()));
  CHECK(ok);
  // reset from a source, then two levels top-down on the host to get a frontier worth pulling from
  const int source = 0;
  std::vector<int> dist(V), q0(V + 64), counts(8, 0);
  std::vector<unsigned> visited(words + 4), fbm(words + 4, 0xdeadbeefu), nbm(words + 4, 0u);
  cuemu::launch(2, 64, 0, 1, [&] {
    bfs_reset_kernel(dist.data(), visited.data(), fbm.data(), V, source, q0.data(), counts.data(), dead.data()); });
  CHECK(dist[source] == 0 && dist[1] == INT_MAX && counts[0] == 1 && q0[0] == source && fbm[3] == 0u);
  std::vector<int> ref(V, INT_MAX), frontier{source};
  ref[source] = 0;
  for (int level = 0; level < 2; ++level) {
    std::vector<int> next;
    for (int v : frontier)
      for (int u : adj[v])
        if (ref[u] == INT_MAX) {
          ref[u] = level + 1;
          next.push_back(u);
        }
    frontier.swap(next);
  }
  for (int v = 0; v < V; ++v)
    if (ref[v] != INT_MAX) {
      dist[v] = ref[v];
      visited[v >> 5] |= 1u << (v & 31);
    }
  // frontier queue -> bitmap
  std::vector<int> q(frontier);
  q.resize(V + 64);
  int q_count = static_cast<int>(frontier.size());
  std::fill(fbm.begin(), fbm.end(), 0u);
  cuemu::launch(2, 64, 0, 1, [&] { queue_to_bitmap_kernel(q.data(), &q_count, fbm.data()); });
  // level 3, bottom-up over every word; the still-unvisited vertices are collected into a list
  ctrl_t ctrl;
  std::memset(&ctrl, 0, sizeof ctrl);
  std::vector<int> unv0(V + 64, -1), unv1(V + 64, -1);
  int next_count = 0, unv_count[2] = {0, 0};
  cuemu::launch(3, 256, 0, 1, [&] {
    bfs_bottom_up_kernel<256, 8>(g, visited.data(), fbm.data(), nbm.data(), dist.data(), 3, &ctrl, &next_count,
                                 unv0.data(), &unv_count[0]); });
  std::set<int> in_frontier(frontier.begin(), frontier.end()), found, still;
  for (int v = 0; v < V; ++v)
    if (ref[v] == INT_MAX && !adj[v].empty()) {
      bool hit = false;
      for (int u : adj[v])
        hit = hit || in_frontier.count(u);
      (hit ? found : still).insert(v);
    }
  ok = next_count == static_cast<int>(found.size());
  for (int v = 0; v < V; ++v) {
    const bool f = found.count(v) != 0;
    ok = ok && (((nbm[v >> 5] >> (v & 31)) & 1u) == f) && (dist[v] == (f ? 3 : ref[v]));
    ok = ok && (((visited[v >> 5] >> (v & 31)) & 1u) == (f || ref[v] != INT_MAX || adj[v].empty()));
  }
  CHECK(ok);
  {  // the same two levels through the second-generation pull kernels: K1 (one probe from first_nb per
     // unvisited vertex, words out) + K2 (full search from the second in-neighbour for K1's retry map)
    std::vector<int2> first_nb(V + 64, int2{12345, 12345});
    cuemu::launch(3, 64, 0, 1, [&] { bfs_first_neighbor_kernel(g, first_nb.data()); });
    bool fn_ok = true;
    for (int v = 0; v < V; ++v) {
      const int n = static_cast<int>(adj[v].size());
      const int want_x = n ? adj[v][0] : -1;
      const int want_y = n >= 2 ? (adj[v][1] | (n == 2 ? kNoMoreNeighbors : 0)) : -1;
      fn_ok = fn_ok && first_nb[v].x == want_x && first_nb[v].y == want_y;
    }
    CHECK(fn_ok);
    std::vector<unsigned> visited_c(words + 4, 0u), nbm_c(words + 4, 0xdeadbeefu), retry_map(words + 4, 0xdeadbeefu);
    std::vector<int> dist_c(V);
    for (int v = 0; v < words * 32; ++v) {
      if (v < V)
        dist_c[v] = ref[v] != INT_MAX ? ref[v] : -5;  // unlabelled so far: the reset left them unwritten (fill mode)
      if (v >= V || ref[v] != INT_MAX || adj[v].empty())
        visited_c[v >> 5] |= 1u << (v & 31);
    }
    int found_c = 0;
    ctrl_t ca, cb;
    std::memset(&ca, 0, sizeof ca);
    std::memset(&cb, 0, sizeof cb);
    const bitmap_frontier_t in_f{fbm.data()};
    cuemu::launch(3, 256, 0, 1, [&] {
      bfs_pull_first_kernel<256>(V, first_nb.data(), visited_c.data(), in_f, nbm_c.data(), retry_map.data(),
                                 dist_c.data(), 3, &ca, &found_c, dead.data(), 1, source); });
    // K1 alone: exactly the vertices whose first or second in-neighbour is in the frontier; misses with a third
    // in-neighbour to look at -> retry map; the probe count is 1 per vertex + 1 per first-probe miss with a second
    std::set<int> k1_found, k1_retry, k1_single;
    unsigned long long want_probes = 0;
    for (int v = 0; v < V; ++v)
      if (ref[v] == INT_MAX && !adj[v].empty()) {
        ++want_probes;
        bool hit = in_frontier.count(adj[v][0]) != 0;
        if (!hit && adj[v].size() >= 2) {
          ++want_probes;
          hit = in_frontier.count(adj[v][1]) != 0;
        }
        if (hit)
          k1_found.insert(v);
        else
          (adj[v].size() <= 2 ? k1_single : k1_retry).insert(v);
      }
    CHECK(found_c == static_cast<int>(k1_found.size()));
    bool maps_ok = true;
    for (int v = 0; v < V; ++v)
      maps_ok = maps_ok && (((retry_map[v >> 5] >> (v & 31)) & 1u) == (k1_retry.count(v) != 0)) &&
                (((nbm_c[v >> 5] >> (v & 31)) & 1u) == (k1_found.count(v) != 0));
    CHECK(maps_ok);
    CHECK(ca.edges == want_probes && ca.hub_count == static_cast<int>(k1_retry.size()));
    cuemu::launch(3, 256, 0, 1, [&] {
      bfs_pull_rest_kernel<256, 32, 8>(g, retry_map.data(), visited_c.data(), in_f, nbm_c.data(), dist_c.data(), 3, &cb,
                                       &found_c); });
    CHECK(found_c == next_count && dist_c == dist);  // incl. INT_MAX written by K1's fill mode for the unreached
    {  // the deferred fill when no pull level happens: every unlabelled vertex except the source
      std::vector<int> d2(V);
      std::vector<unsigned> vis2(words + 4, 0u);
      for (int v = 0; v < words * 32; ++v) {
        if (v < V)
          d2[v] = ref[v] != INT_MAX ? ref[v] : -9;
        if (v >= V || ref[v] != INT_MAX || adj[v].empty())
          vis2[v >> 5] |= 1u << (v & 31);
      }
      cuemu::launch(3, 64, 0, 1, [&] { bfs_fill_unreached_kernel(d2.data(), vis2.data(), dead.data(), V, source); });
      bool fill_ok = true;
      for (int v = 0; v < V; ++v)
        fill_ok = fill_ok && d2[v] == ref[v];
      CHECK(fill_ok);
    }
    bool same_maps = true;
    for (int v = 0; v < V; ++v)
      same_maps = same_maps && (((nbm_c[v >> 5] ^ nbm[v >> 5]) >> (v & 31)) & 1u) == 0 &&
                  (((visited_c[v >> 5] ^ visited[v >> 5]) >> (v & 31)) & 1u) == 0;
    CHECK(same_maps);
    CHECK(ca.edges + cb.edges <= ctrl.edges + static_cast<unsigned long long>(V));  // never more probes than the sweep (+ slack)
    // next level: frontier = the level-3 vertices (nbm_c)
    std::set<int> f2, s2;
    for (int v : still) {
      bool hit = false;
      for (int u : adj[v])
        hit = hit || found.count(u);
      (hit ? f2 : s2).insert(v);
    }
    std::vector<unsigned> nbm_d(words + 4, 0xdeadbeefu);
    int found_d = 0;
    std::memset(&ca, 0, sizeof ca);
    std::memset(&cb, 0, sizeof cb);
    const bitmap_frontier_t in_f2{nbm_c.data()};
    cuemu::launch(3, 256, 0, 1, [&] {
      bfs_pull_first_kernel<256>(V, first_nb.data(), visited_c.data(), in_f2, nbm_d.data(), retry_map.data(),
                                 dist_c.data(), 4, &ca, &found_d, nullptr, 0, -1, /*batch_words=*/8); });
    cuemu::launch(3, 256, 0, 1, [&] {
      bfs_pull_rest_kernel<256, 32, 8>(g, retry_map.data(), visited_c.data(), in_f2, nbm_d.data(), dist_c.data(), 4, &cb,
                                       &found_d, /*batch_words=*/16); });
    bool ok2 = found_d == static_cast<int>(f2.size());
    for (int v = 0; v < V; ++v) {
      const bool f = f2.count(v) != 0;
      ok2 = ok2 && (((nbm_d[v >> 5] >> (v & 31)) & 1u) == f) && (dist_c[v] == (f ? 4 : (found.count(v) ? 3 : ref[v])));
      ok2 = ok2 && (((visited_c[v >> 5] >> (v & 31)) & 1u) == (f || found.count(v) || ref[v] != INT_MAX || adj[v].empty()));
    }
    CHECK(ok2);
    CHECK(!f2.empty() && !s2.empty() && !k1_retry.empty() && !k1_single.empty());
  }
  std::set<int> listed(unv0.begin(), unv0.begin() + unv_count[0]);
  CHECK(listed == still && static_cast<int>(listed.size()) == unv_count[0]);
  CHECK(ctrl.edges > 0 && ctrl.edges <= static_cast<unsigned long long>(ro[V]));
  // bitmap -> queue of the new frontier
  std::vector<int> q2(V + 64, -1);
  int q2_count = 0;
  cuemu::launch(2, 64, 0, 1, [&] { bitmap_to_queue_kernel(nbm.data(), words, q2.data(), &q2_count); });
  CHECK(std::set<int>(q2.begin(), q2.begin() + q2_count) == found && q2_count == static_cast<int>(found.size()));
  // level 4, bottom-up over the LIST of unvisited vertices with the level-3 vertices as the frontier
  std::vector<unsigned> nbm2(words + 4, 0u);
  std::memset(&ctrl, 0, sizeof ctrl);
  int next_count2 = 0;
  cuemu::launch(3, 256, 0, 1, [&] {
    bfs_bottom_up_list_kernel<256, 8>(g, unv0.data(), &unv_count[0], visited.data(), nbm.data(), nbm2.data(),
                                      dist.data(), 4, &ctrl, &next_count2, unv1.data(), &unv_count[1]); });
  std::set<int> found2, still2;
  for (int v : still) {
    bool hit = false;
    for (int u : adj[v])
      hit = hit || found.count(u);
    (hit ? found2 : still2).insert(v);
  }
  ok = next_count2 == static_cast<int>(found2.size());
  for (int v : found2)
    ok = ok && dist[v] == 4 && ((nbm2[v >> 5] >> (v & 31)) & 1u) && ((visited[v >> 5] >> (v & 31)) & 1u);
  CHECK(ok);
  CHECK(std::set<int>(unv1.begin(), unv1.begin() + unv_count[1]) == still2);
  CHECK(!found2.empty() && !still2.empty());  // the list kernel had something to find and something to pass on
  std::printf("bottom-up: level 3 found %zu (sweep), level 4 found %zu (list), %zu still unvisited\n", found.size(),
              found2.size(), still2.size());
}

int main(int argc, char** argv) {
  std::mt19937 rng(argc > 1 ? std::atoi(argv[1]) : 1);
  check_scan(rng);
  check_pagerank_iteration(rng);
  check_bottom_up(rng);
  if (failures == 0)
    std::printf("EMU OK %d\n", checks);
  return failures == 0 ? 0 : 1;
}
