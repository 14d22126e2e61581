// tests/cuemu/emu_advance.cpp -- TEST INFRASTRUCTURE: runs the advance kernels of include/gunrock/b200/advance.cuh
// (kernel section only; generated header advance_kernels.gen.cuh, see tests/test_cuemu_kernels.py) under the CPU
// emulator and checks every launch against a straightforward expansion of the same frontier:
//   * merge_path_partition_kernel + advance_merge_path_kernel (CTA tiles of 2048 edges; the kernel user lambdas get):
//     also validates the emulator itself;
//   * advance_warp_path_kernel: warp-private spans, 4 / 8 chunks in flight (the fused BFS / SSSP functors' default);
//   * advance_binned_kernel + advance_hub_kernel ("block_mapped"; the hub kernel's cp.async.bulk + mbarrier staging is
//     emulated as an immediate copy), advance_thread_mapped_kernel, advance_tail_kernel (several levels per launch);
//   * the reference's output layout of advance_merge_path_kernel (one slot per edge rank, -1 for rejected edges);
//   * both with the BFS claim functor (bitmap test-and-set) and the SSSP relax functor (needs the source id,
//     reads weights), degree-sum accounting on.
// Usage: emu_advance <seed>;  prints "EMU OK <checks>" and returns 0 when every check passes.
#include <cstdio>
#include <map>
#include <numeric>
#include <random>
#include <set>
#include <vector>

#include <cuemu.h>

#include "advance_kernels.gen.cuh"
#include "functors.gen.cuh"
#include "dense_kernels.gen.cuh"

using namespace gunrock::b200;

static int failures = 0, checks = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    ++checks;                                                           \
    if (!(cond)) {                                                      \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                       \
    }                                                                   \
  } while (0)

struct graph_t {
  int V = 0;
  std::vector<int> ro, ci;
  std::vector<float> w;
  csr_view_t view() const {
    csr_view_t g;
    g.n_vertices = V;
    g.n_edges = static_cast<int>(ci.size());
    g.row_offsets = ro.data();
    g.column_indices = ci.data();
    g.values = w.data();
    return g;
  }
};

/// Power-law-ish random graph: low ids are hubs (as in the bench graph), some vertices have no edges.
static graph_t make_graph(int V, int avg_deg, std::mt19937& rng) {
  graph_t g;
  g.V = V;
  std::vector<std::vector<int>> adj(V);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  const long long E = static_cast<long long>(V) * avg_deg;
  for (long long e = 0; e < E; ++e) {
    int a = static_cast<int>(V * std::pow(u(rng), 3.0));  // skewed towards 0
    int b = static_cast<int>(V * std::pow(u(rng), 1.5));
    if (a == b)
      continue;
    adj[a].push_back(b);
    adj[b].push_back(a);
  }
  g.ro.assign(V + 1, 0);
  for (int v = 0; v < V; ++v) {
    if (v % 7 == 3)
      adj[v].clear();  // isolated rows inside the id range
    std::sort(adj[v].begin(), adj[v].end());
    adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
    g.ro[v + 1] = g.ro[v] + static_cast<int>(adj[v].size());
  }
  for (int v = 0; v < V; ++v)
    for (int x : adj[v])
      if (x % 7 != 3)
        g.ci.push_back(x);
      else
        g.ci.push_back((x + 1) % V == v ? x : (x + 1) % V);  // keep the row length, avoid isolated targets
  g.w.resize(g.ci.size());
  for (auto& x : g.w)
    x = 1.0f + static_cast<float>(rng() % 63);
  for (int i = 0; i < 16; ++i)  // padding the kernels may over-read (16-byte slabs)
    g.ci.push_back(0), g.w.push_back(1.0f);
  return g;
}

struct frontier_case_t {
  std::vector<int> in;        // frontier ids
  std::vector<int> scanned;   // exclusive degree scan (n + 1 entries)
  std::vector<int> row_base;  // CSR offset of every frontier row
};
static frontier_case_t make_frontier(const graph_t& g, std::vector<int> ids) {
  frontier_case_t f;
  f.in = std::move(ids);
  const int n = static_cast<int>(f.in.size());
  f.scanned.assign(n + 1, 0);
  f.row_base.assign(n + 2, 0);
  for (int i = 0; i < n; ++i) {
    int v = f.in[i];  // -1 = invalid slot of a bypass-filtered frontier: no edges (advance.cuh frontier_degree_scan)
    f.scanned[i + 1] = f.scanned[i] + (v >= 0 ? g.ro[v + 1] - g.ro[v] : 0);
    f.row_base[i] = v >= 0 ? g.ro[v] : 0;
  }
  return f;
}

enum class kind_t { cta2048, warp4, warp8, binned, binned_plain_loads, thread };

struct run_out_t {
  std::vector<int> out;
  std::vector<unsigned> visited;
  std::vector<int> dist;
  ctrl_t ctrl;
};

template <int kTile>
static std::vector<int> partition(const frontier_case_t& f) {
  const int n = static_cast<int>(f.in.size());
  const int total = f.scanned[n];
  std::vector<int> rows(static_cast<size_t>(total) / kTile + 8, -12345);
  const int n_copy = n;
  cuemu::launch(2, 64, 0, 1, [&] { merge_path_partition_kernel<kTile>(f.scanned.data(), &n_copy, 0, rows.data()); });
  return rows;
}

/// One advance level with the BFS claim functor.
static run_out_t run_bfs(const graph_t& g, const frontier_case_t& f, const std::vector<unsigned>& visited0,
                         kind_t kind, int grid_ctas) {
  run_out_t r;
  const int n = static_cast<int>(f.in.size());
  r.visited = visited0;
  r.dist.assign(g.V, 0x7fffffff);
  r.out.assign(static_cast<size_t>(g.V) + 64, -7);
  int out_count = 0;
  std::memset(&r.ctrl, 0, sizeof r.ctrl);
  advance_params_t p;
  p.g = g.view();
  p.in = f.in.data();
  p.in_count = &n;
  p.out = r.out.data();
  p.out_count = &out_count;
  p.out_capacity = static_cast<int>(r.out.size());
  p.ctrl = &r.ctrl;
  p.row_base = f.row_base.data();
  bfs_claim_op op{r.visited.data(), r.dist.data(), 5};
  constexpr auto kV = advance_input_t::vertices;
  constexpr auto kO = advance_output_t::vertices;
  std::vector<int> rows, hubs(static_cast<size_t>(g.V) + 16, -1);
  if (kind == kind_t::thread) {
    cuemu::launch(grid_ctas, 256, 0, 1, [&] { advance_thread_mapped_kernel<256, kV, kO, true, false>(p, op); });
  } else if (kind == kind_t::binned || kind == kind_t::binned_plain_loads) {
    // "block_mapped": CTA walk + the slab kernel for the deferred hub rows (cp.async.bulk staging, or -- arrays not
    // 16-byte aligned -- plain coalesced loads)
    p.hub_threshold = 64;
    p.hubs = hubs.data();
    p.hub_capacity = static_cast<int>(hubs.size());
    p.tma_ok = kind == kind_t::binned && (reinterpret_cast<uintptr_t>(g.ci.data()) & 15u) == 0;
    p.entries_per_ticket = 64;
    cuemu::launch(grid_ctas, 256, 0, 1, [&] { advance_binned_kernel<256, kV, kO, true, false>(p, op); });
    std::vector<hub_slab_t> slabs(static_cast<size_t>(g.ro[g.V]) / 2048 + hubs.size() + 64);
    p.hub_slabs = slabs.data();
    p.hub_slab_capacity = static_cast<int>(slabs.size());
    cuemu::launch(2, 64, 0, 1, [&] { advance_hub_table_kernel<2048>(p); });
    cuemu::launch(2, 256, 0, 1, [&] { advance_hub_kernel<256, 2048, kO, true, false>(p, op); });
  } else if (kind == kind_t::cta2048) {
    rows = partition<2048>(f);
    p.tile_rows = rows.data();
    cuemu::launch(grid_ctas, 256, 0, 1, [&] { advance_merge_path_kernel<256, 2048, kV, kO, true, false, false>(p, f.scanned.data(), op); });
  } else {
    rows = partition<256>(f);
    p.tile_rows = rows.data();
    constexpr int kThreads = 64;  // two warps per CTA
    constexpr int kWarpBytes = warp_path_ints<256, false>() * 4;
    const size_t stage = (kThreads / 32) * kWarpBytes;
    if (kind == kind_t::warp4)
      cuemu::launch(grid_ctas, kThreads, stage, 1, [&] {
        advance_warp_path_kernel<kThreads, 1, 256, 4, kV, kO, true, false>(p, f.scanned.data(), op); });
    else
      cuemu::launch(grid_ctas, kThreads, stage, 1, [&] {
        advance_warp_path_kernel<kThreads, 1, 256, 8, kV, kO, true, false>(p, f.scanned.data(), op); });
  }
  r.out.resize(out_count);
  return r;
}

static void check_bfs(const graph_t& g, const frontier_case_t& f, const std::vector<unsigned>& visited0,
                      const run_out_t& r, const char* what) {
  std::set<int> expect;
  unsigned long long total = 0, deg_sum = 0;
  for (int v : f.in) {
    if (v < 0)
      continue;
    total += static_cast<unsigned>(g.ro[v + 1] - g.ro[v]);
    for (int e = g.ro[v]; e < g.ro[v + 1]; ++e) {
      int d = g.ci[e];
      if (!((visited0[d >> 5] >> (d & 31)) & 1u))
        expect.insert(d);
    }
  }
  for (int d : expect)
    deg_sum += static_cast<unsigned>(g.ro[d + 1] - g.ro[d]);
  std::vector<int> got = r.out;
  std::sort(got.begin(), got.end());
  const bool same = got == std::vector<int>(expect.begin(), expect.end());  // each unvisited neighbour exactly once
  if (!same)
    std::printf("  [%s] emitted %zu, expected %zu\n", what, got.size(), expect.size());
  CHECK(same);
  CHECK(r.ctrl.edges == total);
  CHECK(r.ctrl.deg_sum == deg_sum);
  CHECK(r.ctrl.overflow == 0);
  bool labels = true, bits = true;
  for (int v = 0; v < g.V; ++v) {
    const bool was = (visited0[v >> 5] >> (v & 31)) & 1u, now = (r.visited[v >> 5] >> (v & 31)) & 1u;
    labels = labels && (r.dist[v] == (expect.count(v) ? 5 : 0x7fffffff));
    bits = bits && (now == (was || expect.count(v) != 0));
  }
  CHECK(labels);
  CHECK(bits);
}

/// Robustness of the default kernels: an output frontier that is too small must raise ctrl.overflow and never
/// write past its capacity; a full hub list keeps the rows in the CTA walk.
static void run_and_check_limits(const graph_t& g, const frontier_case_t& f, const std::vector<unsigned>& visited0) {
  const int n = static_cast<int>(f.in.size());
  constexpr auto kV = advance_input_t::vertices;
  constexpr auto kO = advance_output_t::vertices;
  for (int which = 0; which < 3; ++which) {
    std::vector<unsigned> visited = visited0;
    std::vector<int> dist(g.V, 0x7fffffff), out(static_cast<size_t>(g.V) + 64, -7), hubs(8, -1);
    const int cap = which == 2 ? static_cast<int>(out.size()) : 100;  // 0, 1: too small on purpose
    int out_count = 0;
    ctrl_t ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    advance_params_t p;
    p.g = g.view();
    p.in = f.in.data();
    p.in_count = &n;
    p.out = out.data();
    p.out_count = &out_count;
    p.out_capacity = cap;
    p.ctrl = &ctrl;
    p.row_base = f.row_base.data();
    bfs_claim_op op{visited.data(), dist.data(), 5};
    if (which == 0) {
      std::vector<int> rows = partition<2048>(f);
      p.tile_rows = rows.data();
      cuemu::launch(2, 256, 0, 1, [&] { advance_merge_path_kernel<256, 2048, kV, kO, true, false, false>(p, f.scanned.data(), op); });
    } else {
      p.hub_threshold = 64;
      p.hubs = hubs.data();
      p.hub_capacity = which == 2 ? 2 : static_cast<int>(hubs.size());  // 2: the hub list overflows
      p.tma_ok = 1;
      cuemu::launch(2, 256, 0, 1, [&] { advance_binned_kernel<256, kV, kO, true, false>(p, op); });
      std::vector<hub_slab_t> slabs(static_cast<size_t>(g.ro[g.V]) / 2048 + hubs.size() + 64);
    p.hub_slabs = slabs.data();
    p.hub_slab_capacity = static_cast<int>(slabs.size());
    cuemu::launch(2, 64, 0, 1, [&] { advance_hub_table_kernel<2048>(p); });
    cuemu::launch(2, 256, 0, 1, [&] { advance_hub_kernel<256, 2048, kO, true, false>(p, op); });
    }
    if (which == 2) {  // nothing may be lost when the hub list is full
      run_out_t r;
      r.out.assign(out.begin(), out.begin() + out_count);
      r.visited = visited;
      r.dist = dist;
      r.ctrl = ctrl;
      r.ctrl.hub_count = 0;
      check_bfs(g, f, visited0, r, "hub list full");
    } else {
      CHECK(ctrl.overflow == 1);
      bool untouched = true;
      for (size_t i = cap; i < out.size(); ++i)
        untouched = untouched && out[i] == -7;
      CHECK(untouched);
    }
  }
}

/// The reference's output layout (advance_merge_path_kernel<..., kRanked = true>, launch_advance_ranked): slot r belongs
/// to edge rank r of the frontier's expansion -- the neighbour / edge id where the functor said true, -1 elsewhere --
/// and the count is the frontier's out-degree sum (merge_path.hxx:218-279).  A capacity below the sum raises overflow
/// and writes nothing past it.
struct keep_even_op {
  int* calls;
  __device__ bool operator()(int, int dst, int, float) const {
    atomicAdd(calls, 1);
    return (dst & 1) == 0;
  }
};
static void run_and_check_ranked(const graph_t& g, const frontier_case_t& f, int grid_ctas) {
  const int n = static_cast<int>(f.in.size());
  const int total = f.scanned[n];
  constexpr auto kV = advance_input_t::vertices;
  std::vector<int> rows = partition<2048>(f);
  for (int which = 0; which < 3; ++which) {  // 0: vertex output, 1: edge output, 2: capacity too small
    const int cap = which == 2 ? total / 2 : total + 7;
    std::vector<int> out(static_cast<size_t>(total) + 64, -7);
    int out_count = -3, calls = 0;
    ctrl_t ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    advance_params_t p;
    p.g = g.view();
    p.in = f.in.data();
    p.in_count = &n;
    p.out = out.data();
    p.out_count = &out_count;
    p.out_capacity = cap;
    p.ctrl = &ctrl;
    p.row_base = f.row_base.data();
    p.tile_rows = rows.data();
    keep_even_op op{&calls};
    if (which == 1)
      cuemu::launch(grid_ctas, 256, 0, 1, [&] {
        advance_merge_path_kernel<256, 2048, kV, advance_output_t::edges, false, false, true>(p, f.scanned.data(), op); });
    else
      cuemu::launch(grid_ctas, 256, 0, 1, [&] {
        advance_merge_path_kernel<256, 2048, kV, advance_output_t::vertices, false, false, true>(p, f.scanned.data(), op); });
    std::vector<int> expect;
    for (int v : f.in) {
      if (v < 0)
        continue;
      for (int e = g.ro[v]; e < g.ro[v + 1]; ++e)
        expect.push_back((g.ci[e] & 1) == 0 ? (which == 1 ? e : g.ci[e]) : -1);
    }
    CHECK(static_cast<int>(expect.size()) == total);
    CHECK(calls == total);                         // the functor runs once per edge whatever the capacity
    CHECK(out_count == std::min(total, cap));
    CHECK(ctrl.overflow == (which == 2 && total > cap ? 1 : 0));
    CHECK(ctrl.edges == static_cast<unsigned long long>(total));
    bool same = true, untouched = true;
    for (int r = 0; r < std::min(total, cap); ++r)
      same = same && out[r] == expect[r];
    for (size_t r = std::min(total, cap); r < out.size(); ++r)
      untouched = untouched && out[r] == -7;
    CHECK(same);
    CHECK(untouched);
  }
  std::printf("ranked layout: frontier %4d rows %7d slots ok\n", n, total);
}

/// One relaxation sweep with the SSSP functor (reads the source id and the weights).
static void run_and_check_sssp(const graph_t& g, const frontier_case_t& f, int mode, int grid_ctas) {
  const bool warp_path = mode == 1;
  const int n = static_cast<int>(f.in.size());
  std::vector<float> dist(g.V, 3.0e38f), dist0;
  for (int i = 0; i < n; ++i)
    dist[f.in[i]] = static_cast<float>(i % 11);
  dist0 = dist;
  std::vector<int> stamp(g.V, -1), out(static_cast<size_t>(g.V) + 64, -7);
  int out_count = 0;
  ctrl_t ctrl;
  std::memset(&ctrl, 0, sizeof ctrl);
  advance_params_t p;
  p.g = g.view();
  p.in = f.in.data();
  p.in_count = &n;
  p.out = out.data();
  p.out_count = &out_count;
  p.out_capacity = static_cast<int>(out.size());
  p.ctrl = &ctrl;
  p.row_base = f.row_base.data();
  sssp_relax_op op{dist.data(), stamp.data(), 3};
  constexpr auto kV = advance_input_t::vertices;
  constexpr auto kO = advance_output_t::vertices;
  std::vector<int> rows;
  if (warp_path) {
    rows = partition<256>(f);
    p.tile_rows = rows.data();
    constexpr int kWarpBytes = warp_path_ints<256, true>() * 4;
    if (grid_ctas % 2)
      cuemu::launch(grid_ctas, 64, 2 * kWarpBytes, 1, [&] {
        advance_warp_path_kernel<64, 1, 256, 4, kV, kO, true, true>(p, f.scanned.data(), op); });
    else  // 8 chunks in flight: what the SSSP functor runs by default
      cuemu::launch(grid_ctas, 64, 2 * kWarpBytes, 1, [&] {
        advance_warp_path_kernel<64, 1, 256, 8, kV, kO, true, true>(p, f.scanned.data(), op); });
  } else if (mode == 2) {  // block_mapped: hub rows staged with TWO bulk copies per slab (indices + weights)
    std::vector<int> hubs(static_cast<size_t>(g.V) + 16, -1);
    p.hub_threshold = 64;
    p.hubs = hubs.data();
    p.hub_capacity = static_cast<int>(hubs.size());
    p.tma_ok = (reinterpret_cast<uintptr_t>(g.ci.data()) & 15u) == 0 && (reinterpret_cast<uintptr_t>(g.w.data()) & 15u) == 0;
    p.entries_per_ticket = 128;
    cuemu::launch(grid_ctas, 256, 0, 1, [&] { advance_binned_kernel<256, kV, kO, true, true>(p, op); });
    std::vector<hub_slab_t> slabs(static_cast<size_t>(g.ro[g.V]) / 2048 + hubs.size() + 64);
    p.hub_slabs = slabs.data();
    p.hub_slab_capacity = static_cast<int>(slabs.size());
    cuemu::launch(2, 64, 0, 1, [&] { advance_hub_table_kernel<2048>(p); });
    cuemu::launch(2, 256, 0, 1, [&] { advance_hub_kernel<256, 2048, kO, true, true>(p, op); });
  } else {
    rows = partition<2048>(f);
    p.tile_rows = rows.data();
    cuemu::launch(grid_ctas, 256, 0, 1, [&] { advance_merge_path_kernel<256, 2048, kV, kO, true, true, false>(p, f.scanned.data(), op); });
  }
  // expected: dist = min over candidates of (dist0[src] + w) -- sources keep their start value unless relaxed too.
  // (a source may itself be lowered during the sweep; candidates then use either value: accept both bounds)
  std::vector<float> lo = dist0, hi = dist0;  // lo: best case (sources already lowered), hi: sources at start value
  for (int v : f.in)
    for (int e = g.ro[v]; e < g.ro[v + 1]; ++e)
      hi[g.ci[e]] = std::min(hi[g.ci[e]], dist0[v] + g.w[e]);
  bool within = true, queued_ok = true;
  std::set<int> q(out.begin(), out.begin() + out_count);
  for (int v = 0; v < g.V; ++v) {
    within = within && dist[v] <= hi[v] && dist[v] >= 0.0f;
    if (dist[v] < dist0[v])
      queued_ok = queued_ok && q.count(v) == 1;  // every lowered vertex is in the next frontier ...
  }
  CHECK(within);
  CHECK(queued_ok);
  CHECK(static_cast<int>(q.size()) == out_count);  // ... exactly once (stamp dedup)
  for (int v : q)
    queued_ok = queued_ok && dist[v] < dist0[v];
  CHECK(queued_ok);
}

/// advance_tail_kernel: level after level in ONE single-CTA launch, against a level-synchronous BFS on the host.
static void run_and_check_tail(const graph_t& g, int start, int max_levels) {
  std::vector<unsigned> visited((g.V + 31) / 32 + 4, 0u);
  std::vector<int> dist(g.V, 0x7fffffff), q0(static_cast<size_t>(g.V) + 64, -1), q1(q0);
  visited[start >> 5] |= 1u << (start & 31);
  dist[start] = 0;
  q0[0] = start;
  int counts[2] = {1, 0};
  tail_report_t rep;
  std::memset(&rep, 0, sizeof rep);
  cuemu::launch(1, 256, 0, 1, [&] {
    advance_tail_kernel<256, false>(g.view(), q0.data(), q1.data(), counts, 0, 0, max_levels, ~0ull,
                                    bfs_claim_maker{visited.data(), dist.data()}, &rep, 42);
  });
  std::vector<int> ref(g.V, 0x7fffffff), cur{start}, nxt;
  ref[start] = 0;
  int level = 0;
  while (!cur.empty() && level < max_levels) {
    nxt.clear();
    for (int v : cur)
      for (int e = g.ro[v]; e < g.ro[v + 1]; ++e)
        if (ref[g.ci[e]] == 0x7fffffff) {
          ref[g.ci[e]] = level + 1;
          nxt.push_back(g.ci[e]);
        }
    cur.swap(nxt);
    ++level;
  }
  CHECK(rep.seq == 42 && rep.levels == level);
  CHECK(rep.count == static_cast<int>(cur.size()) && counts[rep.cur] == rep.count);
  CHECK(dist == ref);
  std::printf("tail: %d levels from %d, frontier left %d\n", rep.levels, start, rep.count);
}

/// A whole traversal: the level loop of the enactor on the host (degree scan + partition + advance per level),
/// the kernels under emulation, the visited map evolving
/// from level to level.  Depths must equal a plain BFS.
static void run_and_check_whole_bfs(const graph_t& g, int source, kind_t kind, const char* name) {
  std::vector<unsigned> visited((g.V + 31) / 32 + 4, 0u);
  std::vector<int> dist(g.V, 0x7fffffff);
  visited[source >> 5] |= 1u << (source & 31);
  dist[source] = 0;
  std::vector<int> frontier{source};
  int level = 0;
  while (!frontier.empty()) {
    const frontier_case_t f = make_frontier(g, frontier);
    run_out_t r;
    {  // run_bfs labels with 5: relabel through a scratch copy, keep the visited map
      r = run_bfs(g, f, visited, kind, 3);
      visited = r.visited;
      for (int v : r.out)
        dist[v] = level + 1;
    }
    frontier = r.out;
    ++level;
  }
  std::vector<int> ref(g.V, 0x7fffffff), cur{source}, nxt;
  ref[source] = 0;
  for (int l = 0; !cur.empty(); ++l) {
    nxt.clear();
    for (int v : cur)
      for (int e = g.ro[v]; e < g.ro[v + 1]; ++e)
        if (ref[g.ci[e]] == 0x7fffffff) {
          ref[g.ci[e]] = l + 1;
          nxt.push_back(g.ci[e]);
        }
    cur.swap(nxt);
  }
  CHECK(dist == ref);
  std::printf("whole bfs %-6s from %d: %d levels\n", name, source, level);
}

/// Kernels of the dense frontier views (framework/frontier/dense_frontier.hxx), bitmap (1 bit) and boolmap (8).
static void run_and_check_dense_frontier(std::mt19937& rng) {
  using namespace gunrock::frontier::detail;
  for (int bits : {1, 8})
    for (int universe : {1, 31, 32, 33, 1000, 4099}) {
      const size_t n_words = bits == 1 ? (universe + 31) / 32 : (universe + 3) / 4;
      std::vector<unsigned> words(n_words + 8, 0u);
      std::set<int> expect;
      std::vector<int> list;
      for (int k = 0; k < 300; ++k) {
        int v = static_cast<int>(rng() % universe);
        list.push_back(k % 50 == 7 ? -1 : v);  // invalid slots are skipped
        if (k % 50 != 7)
          expect.insert(v);
      }
      const int n_list = static_cast<int>(list.size());
      cuemu::launch(2, 64, 0, 1, [&] { dense_from_list_kernel(words.data(), bits, list.data(), &n_list); });
      if (universe > 40) {
        cuemu::launch(2, 64, 0, 1, [&] { dense_set_range_kernel(words.data(), bits, size_t(universe - 9), size_t(9)); });
        for (int v = universe - 9; v < universe; ++v)
          expect.insert(v);
      }
      int count = 0;
      cuemu::launch(2, 64, 0, 1, [&] { dense_count_kernel(words.data(), n_words, bits, &count); });
      CHECK(count == static_cast<int>(expect.size()));
      std::vector<int> out(static_cast<size_t>(universe) + 8, -5);
      int out_count = 0, overflow = 0;
      cuemu::launch(2, 64, 0, 1, [&] {
        dense_to_list_kernel(words.data(), bits, size_t(universe), out.data(), &out_count, universe, &overflow); });
      out.resize(out_count);
      std::sort(out.begin(), out.end());
      CHECK(out == std::vector<int>(expect.begin(), expect.end()) && overflow == 0);
    }
  std::printf("dense frontier kernels ok\n");
}


int main(int argc, char** argv) {
  std::mt19937 rng(argc > 1 ? std::atoi(argv[1]) : 1);
  const int seed = argc > 1 ? std::atoi(argv[1]) : 1;
  const graph_t g = make_graph(seed % 2 ? 6000 : 20000, seed % 2 ? 6 : 9, rng);
  std::vector<unsigned> visited0((g.V + 31) / 32 + 4, 0u);
  for (int i = 0; i < g.V / 5; ++i) {
    int v = static_cast<int>(rng() % g.V);
    visited0[v >> 5] |= 1u << (v & 31);
  }
  // frontiers: hubs (long rows spanning many spans), a mixed bag with duplicates and degree-0 rows, one short row
  std::vector<std::vector<int>> frontiers;
  frontiers.push_back({0, 1, 2, 4, 5});
  std::vector<int> mixed;
  for (int i = 0; i < 700; ++i)
    mixed.push_back(static_cast<int>(rng() % g.V));
  mixed.push_back(3);
  mixed.push_back(3);
  frontiers.push_back(mixed);
  frontiers.push_back({g.V - 2});
  std::vector<int> holes = mixed;  // a bypass-filtered frontier: invalid slots between the ids
  for (size_t i = 0; i < holes.size(); i += 3)
    holes[i] = -1;
  frontiers.push_back(holes);
  std::vector<int> everyone(g.V);
  std::iota(everyone.begin(), everyone.end(), 0);
  frontiers.push_back(everyone);  // every row, the isolated ones included: many rows per span, many spans
  const struct { kind_t k; const char* name; int grid; } kinds[] = {
      {kind_t::cta2048, "cta2048", 3}, {kind_t::warp4, "warp4", 3}, {kind_t::warp8, "warp8", 2},
      {kind_t::binned, "binned", 3},   {kind_t::binned_plain_loads, "binned-ld", 2},
      {kind_t::thread, "thread", 2}};
  for (auto& ids : frontiers) {
    const frontier_case_t f = make_frontier(g, ids);
    for (auto& k : kinds) {
      run_out_t r = run_bfs(g, f, visited0, k.k, k.grid);
      check_bfs(g, f, visited0, r, k.name);
      std::printf("bfs %-8s frontier %4zu rows %7d edges -> %5zu claimed\n", k.name, f.in.size(), f.scanned.back(),
                  r.out.size());
    }
    if (f.scanned.back() > 2000)
      run_and_check_limits(g, f, visited0);
    run_and_check_ranked(g, f, 3);
    if (std::find(f.in.begin(), f.in.end(), -1) != f.in.end())
      continue;  // the SSSP sweep below labels its sources
    run_and_check_sssp(g, f, 0, 2);
    run_and_check_sssp(g, f, 1, 3);
    run_and_check_sssp(g, f, 1, 2);
    run_and_check_sssp(g, f, 2, 2);
  }
  for (auto& k : kinds)
    if (k.k == kind_t::cta2048 || k.k == kind_t::warp4 || k.k == kind_t::warp8)
      run_and_check_whole_bfs(g, 0, k.k, k.name);
  run_and_check_whole_bfs(g, g.V - 2, kind_t::warp4, "warp4");
  run_and_check_dense_frontier(rng);
  run_and_check_tail(g, 0, 2);
  run_and_check_tail(g, g.V - 2, 16);
  if (failures == 0)
    std::printf("EMU OK %d\n", checks);
  return failures == 0 ? 0 : 1;
}
