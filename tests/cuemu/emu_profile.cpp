// tests/cuemu/emu_profile.cpp -- TEST / DESIGN INFRASTRUCTURE: which memory path do the probes of one push level
// take?  Runs level 1 (frontier = the neighbours of the bench source) of a graph written by
// profiles/micro/emulated_probe_paths.py through the default merge_path kernel and through the on-chip-copy
// variants under the CPU emulator and prints, per kernel, how many visited-bit probes went to the global map
// (L1 / L2), to the copy in the CTA's own shared memory and to a cluster peer's (DSMEM) -- with the copy sized in
// proportion to the graph as the launcher would size it on a scale-22 graph (1.27 M of 4.19 M vertices per CTA).
// Usage: emu_profile <graph.csr> <source>
#include <cstdio>
#include <vector>

#include <cuemu.h>

#include "advance_kernels.gen.cuh"
#include "functors.gen.cuh"

using namespace gunrock::b200;

int main(int argc, char** argv) {
  if (argc < 3)
    return 2;
  FILE* f = std::fopen(argv[1], "rb");
  int hdr[3];
  if (!f || std::fread(hdr, 4, 3, f) != 3)
    return 2;
  const int V = hdr[0], E = hdr[2];
  std::vector<int> ro(V + 1), ci(static_cast<size_t>(E) + 16, 0);
  if (std::fread(ro.data(), 4, V + 1, f) != static_cast<size_t>(V + 1) || std::fread(ci.data(), 4, E, f) != static_cast<size_t>(E))
    return 2;
  std::fclose(f);
  const int source = std::atoi(argv[2]);
  csr_view_t g;
  g.n_vertices = V;
  g.n_edges = E;
  g.row_offsets = ro.data();
  g.column_indices = ci.data();
  // level-1 frontier and the visited map at its start
  std::vector<int> in(ci.begin() + ro[source], ci.begin() + ro[source + 1]);
  const int n = static_cast<int>(in.size());
  std::vector<unsigned> visited0((V + 31) / 32 + 4, 0u);
  visited0[source >> 5] |= 1u << (source & 31);
  for (int v : in)
    visited0[v >> 5] |= 1u << (v & 31);
  std::vector<int> scanned(n + 1, 0), row_base(n + 2, 0);
  for (int i = 0; i < n; ++i) {
    scanned[i + 1] = scanned[i] + (ro[in[i] + 1] - ro[in[i]]);
    row_base[i] = ro[in[i]];
  }
  const int total = scanned[n];
  std::printf("graph %d vertices %d edges; level 1 from %d: %d rows, %d edges\n", V, E, source, n, total);
  constexpr auto kV = advance_input_t::vertices;
  constexpr auto kO = advance_output_t::vertices;
  const int map_words = (V + 31) / 32;
  const double per_cta_fraction = 1270000.0 / 4194304.0;  // what one CTA's copy covers of a scale-22 graph
  for (int variant : {0, 2, 5, 6}) {
    std::vector<unsigned> visited = visited0;
    std::vector<int> dist(V, 0x7fffffff), out(static_cast<size_t>(V) + 64), rows;
    int out_count = 0;
    ctrl_t ctrl;
    std::memset(&ctrl, 0, sizeof ctrl);
    advance_params_t p;
    p.g = g;
    p.in = in.data();
    p.in_count = &n;
    p.out = out.data();
    p.out_count = &out_count;
    p.out_capacity = static_cast<int>(out.size());
    p.ctrl = &ctrl;
    p.row_base = row_base.data();
    bfs_claim_op op{visited.data(), dist.data(), 2};
    cuemu_counters = cuemu_counters_t{};
    const int k = variant == 0 ? 0 : variant == 2 ? 1 : variant == 5 ? 2 : 4;
    if (variant == 0) {
      rows.assign(static_cast<size_t>(total) / 2048 + 8, 0);
      cuemu::launch(8, 64, 0, 1, [&] { merge_path_partition_kernel<2048>(scanned.data(), &n, 0, rows.data()); });
      p.tile_rows = rows.data();
      cuemu::launch(16, 256, 0, 1, [&] { advance_merge_path_kernel<256, 2048, kV, kO, true, false>(p, scanned.data(), op); });
    } else {
      rows.assign(static_cast<size_t>(total) / 256 + 8, 0);
      cuemu::launch(8, 64, 0, 1, [&] { merge_path_partition_kernel<256>(scanned.data(), &n, 0, rows.data()); });
      p.tile_rows = rows.data();
      constexpr int kThreads = 256;
      const int lines_cta = std::max(1, static_cast<int>(per_cta_fraction * V / 1024.0));
      const int lines_map = (map_words + 31) / 32;
      long long lines = static_cast<long long>(lines_cta) * k;
      if (lines > lines_map)
        lines = (lines_map + k - 1) / k * k;
      const int bits = static_cast<int>(lines * 1024);
      const size_t smem = (kThreads / 32) * warp_path_ints<256, false>() * 4 + static_cast<size_t>(lines / k) * 128;
      const int grid = 16;  // CTAs sharing the level (each cluster has its own copy, as on the GPU)
      if (k == 1)
        cuemu::launch(grid, kThreads, smem, 1, [&] {
          advance_warp_path_kernel<kThreads, 1, 256, 8, 1, kV, kO, true, false>(p, scanned.data(), bits, map_words, op); });
      else if (k == 2)
        cuemu::launch(grid, kThreads, smem, 2, [&] {
          advance_warp_path_kernel<kThreads, 1, 256, 8, 2, kV, kO, true, false>(p, scanned.data(), bits, map_words, op); });
      else
        cuemu::launch(grid, kThreads, smem, 4, [&] {
          advance_warp_path_kernel<kThreads, 1, 256, 8, 4, kV, kO, true, false>(p, scanned.data(), bits, map_words, op); });
    }
    const auto& c = cuemu_counters;
    const double e = static_cast<double>(total);
    std::printf("variant %d: claimed %d | per edge: global probes %.3f, own-CTA copy %.3f, cluster peer %.3f, copy updates %.4f\n",
                variant, out_count, c.cached_probes / e, c.shared_probes / e, c.dsmem_probes / e, c.shared_merges / e);
  }
  return 0;
}
