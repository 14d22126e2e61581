/**
 * @file bfs.cuh
 * @brief Fused level-synchronous BFS enactor for B200: bitmap visited set, compacting top-down
 * (push) advance, bottom-up (pull) sweep and the Beamer direction switch.
 *
 * Path replaced: gunrock::bfs::enactor_t::loop + enactor_t::enact
 *   include/gunrock/algorithms/bfs.hxx:93-147, include/gunrock/framework/enactor.hxx:243-288.
 * Result contract kept (bfs.hxx:59-69, examples/algorithms/bfs/bfs_cpu.hxx:32-63):
 *   distances[v] = BFS depth, INT_MAX when unreachable; predecessors untouched.
 * The reference's per-edge `atomicMin(&distances[n], iter+1)` (bfs.hxx:125-127) is replaced by a
 * test-and-set on a V-bit visited map that stays L2 resident (8 MiB at scale 26): edges into
 * already-visited vertices cost one cached 4-byte read and no atomic.  Depths are identical
 * because both are level-synchronous: a vertex is claimed in the first level that reaches it.
 * The direction-optimised variant (advance_direction_t::optimized is a dead parameter in the
 * reference) is new design: Beamer et al.'s alpha/beta switch.
 */
#pragma once

#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <gunrock/b200/advance.cuh>

namespace gunrock {
namespace b200 {

/// Top-down edge functor: claim `dst` on the visited bitmap, label it, keep it.
struct bfs_claim_op {
  static constexpr bool kNeedsSource = false;
  unsigned* visited;
  int* dist;
  int next_level;
  __device__ __forceinline__ bool operator()(int, int dst, int, float) const {
    if (!bitmap_test_and_set(visited, dst))
      return false;
    dist[dst] = next_level;
    return true;
  }
  /// two-phase protocol (advance.cuh op_traits): the probe is a pure cached load ...
  __device__ __forceinline__ unsigned prefetch(int dst) const {
    return ld_cached(visited + (dst >> 5));
  }
  /// ... and only edges whose probe saw a clear bit pay for the atomic.
  __device__ __forceinline__ bool commit(int, int dst, int, float, unsigned word) const {
    const unsigned bit = 1u << (dst & 31);
    if (word & bit)
      return false;
    if (atomicOr(visited + (dst >> 5), bit) & bit)
      return false;
    dist[dst] = next_level;
    return true;
  }
  /// merge_path runs on the warp-private-span kernel, 4 chunks in flight (advance.cuh op_merge_path_kernel)
  static constexpr int kMergePathKernel = 1;
};

/// Builds the claim functor of a given level (advance_tail_kernel runs several levels per launch).
struct bfs_claim_maker {
  unsigned* visited;
  int* dist;
  __device__ __forceinline__ bfs_claim_op operator()(int level) const {
    return bfs_claim_op{visited, dist, level + 1};
  }
};

/// The reference's own functor (bfs.hxx:105-128), kept selectable for like-for-like runs.
struct bfs_atomic_min_op {
  static constexpr bool kNeedsSource = false;
  int* dist;
  int next_level;
  __device__ __forceinline__ bool operator()(int, int dst, int, float) const {
    int old = atomicMin(dist + dst, next_level);
    return next_level < old;
  }
};

/// One bit per vertex with NO in-edges: such a vertex can never be discovered, so the bottom-up
/// sweep need not look at it again every level (60 % of an RMAT-26 graph is isolated vertices).
/// The bits are OR-ed into the visited map at reset; top-down never targets them, and the source
/// gets its own bit and label regardless.
static __global__ void bfs_unreachable_map_kernel(const int* __restrict__ in_offsets, int n_vertices,
                                                  unsigned* __restrict__ map) {
  const int words = (n_vertices + 31) / 32;
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < words; wi += warps) {
    int v = (wi << 5) + lane;
    bool dead = v >= n_vertices || in_offsets[v + 1] == in_offsets[v];
    unsigned m = __ballot_sync(kFull, dead);
    if (lane == 0)
      map[wi] = m;
  }
}

/// `fill_dist` == 0: only the source is labelled here; the INT_MAX of everything that stays unreached is written
/// later, by the first pull level (K1 writes a label for every vertex it looks at anyway) or by
/// bfs_fill_unreached_kernel when the traversal ends without one -- the 4 x V bytes of labels are then written once
/// per run instead of twice (268 MB at scale 26: ~45 us of a 0.85 ms traversal).
static __global__ void bfs_reset_kernel(int* dist, unsigned* visited, unsigned* fbm, int n_vertices,
                                 int source, int* q0, int* counts,
                                 const unsigned* __restrict__ premark = nullptr, int fill_dist = 1) {
  const int words = (n_vertices + 31) / 32;
  const int n = fill_dist ? n_vertices : words;
  if (!fill_dist && blockIdx.x == 0 && threadIdx.x == 0)
    dist[source] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += gridDim.x * blockDim.x) {
    if (fill_dist)
      dist[i] = (i == source) ? 0 : INT_MAX;
    if (i < words) {
      unsigned bit = (i == (source >> 5)) ? (1u << (source & 31)) : 0u;
      visited[i] = bit | (premark ? premark[i] : 0u);
      fbm[i] = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    q0[0] = source;
    counts[0] = 1;
    counts[1] = 0;
  }
}

/// The deferred half of the reset (see bfs_reset_kernel): INT_MAX for every vertex that was never labelled --
/// not visited, or pre-marked as having no in-edges -- except the source.
static __global__ void bfs_fill_unreached_kernel(int* dist, const unsigned* __restrict__ visited,
                                                 const unsigned* __restrict__ premark, int n_vertices, int source) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vertices; v += gridDim.x * blockDim.x) {
    const unsigned bit = 1u << (v & 31);
    const bool labelled = (visited[v >> 5] & bit) && !(premark && (premark[v >> 5] & bit));
    if (!labelled && v != source)
      dist[v] = INT_MAX;
  }
}

static __global__ void queue_to_bitmap_kernel(const int* __restrict__ q, const int* __restrict__ count,
                                       unsigned* __restrict__ bm) {
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int v = q[i];
    atomicOr(bm + (v >> 5), 1u << (v & 31));
  }
}

/// Enumerate set bits into a queue (order is irrelevant to BFS; one atomic per warp pass).
/// `ro` / `deg_sum` (optional): also add up the out-degrees of the enumerated vertices (the first push level after
/// pull levels needs the frontier's out-degree sum, which the pull kernels do not compute).
static __global__ void bitmap_to_queue_kernel(const unsigned* __restrict__ bm, int words, int* q,
                                       int* count, const int* __restrict__ ro = nullptr,
                                       unsigned long long* deg_sum = nullptr) {
  unsigned long long ds = 0;
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int w0 = gw * 32; w0 < words; w0 += warps * 32) {
    int wi = w0 + lane;
    unsigned w = wi < words ? bm[wi] : 0u;
    int c = __popc(w);
    int incl = warp_inclusive_sum(c);
    int total = __shfl_sync(kFull, incl, 31);
    if (!total)
      continue;
    int base = 0;
    if (lane == 0)
      base = atomicAdd(count, total);
    base = __shfl_sync(kFull, base, 0) + incl - c;
    int vb = wi << 5;
    while (w) {
      int b = __ffs(w) - 1;
      w &= w - 1;
      q[base++] = vb + b;
      if (ro)
        ds += static_cast<unsigned>(ro[vb + b + 1] - ro[vb + b]);
    }
  }
  if (deg_sum) {
    ds = warp_sum(ds);
    if (lane == 0 && ds)
      atomicAdd(deg_sum, ds);
  }
}

/**
 * @brief Bottom-up sweep.  One warp owns one 32-vertex word of the visited map, so the map and
 * the next-frontier map are updated with plain stores (no atomics).  Each lane walks its
 * vertex's in-neighbours until one is in the current frontier; lanes still searching after
 * kSerial steps are finished by the whole warp, 32 neighbours per step.
 * `in` is the CSC (or the CSR itself for a symmetric graph).
 */
template <int kThreads, int kSerial>
__global__ void __launch_bounds__(kThreads)
bfs_bottom_up_kernel(csr_view_t in, unsigned* __restrict__ visited,
                     const unsigned* __restrict__ frontier, unsigned* __restrict__ next, int* dist,
                     int next_level, ctrl_t* ctrl, int* next_count, int* unv_out, int* unv_count) {
  // vertices still unvisited after this sweep are collected (staged per warp, one atomic per ~100)
  // so that the following bottom-up levels walk a dense list instead of sweeping every word
  __shared__ int s_emit[kThreads / 32][kEmitCap];
  warp_emitter_t<kEmitCap, false> em;
  em.init(s_emit[threadIdx.x >> 5], unv_out, unv_count, in.n_vertices, nullptr, ctrl);
  const int lane = lane_id();
  const int words = (in.n_vertices + 31) / 32;
  const int warps = (gridDim.x * kThreads) >> 5;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int* __restrict__ ro = in.row_offsets;
  const int* __restrict__ ci = in.column_indices;
  unsigned long long scanned = 0, found_deg = 0;
  int found_cnt = 0;
  // 32 words per warp pass: one coalesced load decides which of them still hold unvisited
  // vertices (late levels: almost none), so the sweep is bandwidth- not latency-bound.
  for (int w0 = gw * 32; w0 < words; w0 += warps * 32) {
    const int my_wi = w0 + lane;
    const unsigned my_vis = my_wi < words ? visited[my_wi] : 0xffffffffu;
    if (my_wi < words && my_vis == 0xffffffffu)
      next[my_wi] = 0;
    unsigned todo = __ballot_sync(kFull, my_vis != 0xffffffffu);
    while (todo) {
    const int src_lane = __ffs(todo) - 1;
    todo &= todo - 1;
    const int wi = w0 + src_lane;
    const unsigned vis = __shfl_sync(kFull, my_vis, src_lane);
    const int v = (wi << 5) + lane;
    bool searching = v < in.n_vertices && !((vis >> lane) & 1u);
    bool found = false;
    int start = 0, end = 0;
    if (searching) {
      start = ro[v];
      end = ro[v + 1];
    }
    const int deg = end - start;
    int e = start;
    for (int k = 0; k < kSerial; ++k) {
      if (searching && e < end) {
        int u = ci[e++];
        ++scanned;
        if (bitmap_test(frontier, u)) {
          found = true;
          searching = false;
        }
      }
    }
    if (e >= end)
      searching = false;
    unsigned rest = __ballot_sync(kFull, searching);
    while (rest) {
      int leader = __ffs(rest) - 1;
      rest &= rest - 1;
      int s = __shfl_sync(kFull, e, leader);
      int t = __shfl_sync(kFull, end, leader);
      bool hit = false;
      for (int off = s; off < t && !hit; off += 32) {
        int idx = off + lane;
        bool mine = false;
        if (idx < t) {
          ++scanned;
          mine = bitmap_test(frontier, ci[idx]);
        }
        hit = __any_sync(kFull, mine);
      }
      if (lane == leader)
        found = hit;
    }
    const unsigned fm = __ballot_sync(kFull, found);
    if (found) {
      dist[v] = next_level;
      found_deg += static_cast<unsigned>(deg);
    }
    if (lane == 0) {
      next[wi] = fm;
      if (fm)
        visited[wi] = vis | fm;
    }
    found_cnt += found ? 1 : 0;
    if (unv_out)
      em.push(v < in.n_vertices && !((vis >> lane) & 1u) && !found, v);
    }
  }
  if (unv_out)
    em.flush();
  scanned = warp_sum(scanned);
  found_deg = warp_sum(found_deg);
  found_cnt = warp_sum(found_cnt);
  if (lane == 0) {
    if (scanned)
      atomicAdd(&ctrl->edges, scanned);
    if (found_cnt) {
      atomicAdd(&ctrl->deg_sum, found_deg);
      atomicAdd(next_count, found_cnt);
    }
  }
}

/**
 * @brief Bottom-up over a LIST of still-unvisited vertices (second and later consecutive pull
 * levels).  After the first sweep only a few percent of the vertices are still unvisited but they
 * are spread over almost every 32-vertex word, so sweeping words again would run with 2-3 active
 * lanes per warp; here every lane has a vertex.  Found vertices set their bit in `next` / `visited`
 * with atomicOr; the rest go to the next list.
 */
template <int kThreads, int kSerial>
__global__ void __launch_bounds__(kThreads)
bfs_bottom_up_list_kernel(csr_view_t in, const int* __restrict__ unv_in,
                          const int* __restrict__ unv_in_count, unsigned* visited,
                          const unsigned* __restrict__ frontier, unsigned* next, int* dist,
                          int next_level, ctrl_t* ctrl, int* next_count, int* unv_out, int* unv_count) {
  __shared__ int s_emit[kThreads / 32][kEmitCap];
  warp_emitter_t<kEmitCap, false> em;
  em.init(s_emit[threadIdx.x >> 5], unv_out, unv_count, in.n_vertices, nullptr, ctrl);
  const int lane = lane_id();
  const int n = *unv_in_count;
  const int* __restrict__ ro = in.row_offsets;
  const int* __restrict__ ci = in.column_indices;
  unsigned long long scanned = 0, found_deg = 0;
  int found_cnt = 0;
  for (;;) {
    int base = 0;
    if (lane == 0)
      base = atomicAdd(&ctrl->work, 32);
    base = __shfl_sync(kFull, base, 0);
    if (base >= n)
      break;
    const int i = base + lane;
    bool searching = i < n;
    const int v = searching ? unv_in[i] : 0;
    bool found = false;
    int start = 0, end = 0;
    if (searching) {
      start = ro[v];
      end = ro[v + 1];
    }
    const int deg = end - start;
    int e = start;
    for (int k = 0; k < kSerial; ++k) {
      if (searching && e < end) {
        int u = ci[e++];
        ++scanned;
        if (bitmap_test(frontier, u)) {
          found = true;
          searching = false;
        }
      }
    }
    if (e >= end)
      searching = false;
    unsigned rest = __ballot_sync(kFull, searching);
    while (rest) {
      int leader = __ffs(rest) - 1;
      rest &= rest - 1;
      int s = __shfl_sync(kFull, e, leader);
      int t = __shfl_sync(kFull, end, leader);
      bool hit = false;
      for (int off = s; off < t && !hit; off += 32) {
        int idx = off + lane;
        bool mine = false;
        if (idx < t) {
          ++scanned;
          mine = bitmap_test(frontier, ci[idx]);
        }
        hit = __any_sync(kFull, mine);
      }
      if (lane == leader)
        found = hit;
    }
    if (found) {
      dist[v] = next_level;
      found_deg += static_cast<unsigned>(deg);
      const unsigned bit = 1u << (v & 31);
      atomicOr(next + (v >> 5), bit);
      atomicOr(visited + (v >> 5), bit);
      ++found_cnt;
    }
    em.push(i < n && !found, v);
  }
  em.flush();
  scanned = warp_sum(scanned);
  found_deg = warp_sum(found_deg);
  found_cnt = warp_sum(found_cnt);
  if (lane == 0) {
    if (scanned)
      atomicAdd(&ctrl->edges, scanned);
    if (found_cnt) {
      atomicAdd(&ctrl->deg_sum, found_deg);
      atomicAdd(next_count, found_cnt);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pull levels, second generation: the FIRST-IN-NEIGHBOURS shortcut.
//
// Measured on the bench graph (RMAT-26, profiles/r2_*): the first bottom-up level inspects 33 M column indices
// for 27 M unvisited vertices -- almost every vertex finds its parent with its first or second probe, because
// in-neighbour lists are sorted and the low ids are the hubs.  Yet the sweep above pays, per vertex, a pair of row
// offsets and a whole 32-byte sector of column indices fetched from HBM to use 4 bytes of it (1.37 GB of DRAM
// traffic for 132 MB of useful indices) at the DRAM's random-sector rate (~72 G sectors/s, 29 % of the copy
// bandwidth), one dependent chain per lane.  So the graph carries one more per-vertex array, built once like
// the transpose:
//     head[v] = { first in-neighbour | -1,  second in-neighbour | -1, bit 31 set when there is no third }
// and a pull level becomes two kernels over 32-vertex words of the visited map:
//   K1 (bfs_pull_first_kernel): a warp expands the unvisited bits of 32 words into a queue in shared memory and
//       walks it 32 vertices at a time: one 8-byte load of head[v] (ascending v: nearly sequential), one probe,
//       a second probe for those that missed -- which settles ~95 % of the vertices that can be found at this
//       level without touching row offsets or column indices.  Results are WORDS (next frontier, visited, and
//       the vertices that need K2) assembled in shared memory and written with coalesced plain stores.
//   K2 (bfs_pull_rest_kernel): the full search from the THIRD in-neighbour on, for the few vertices K1 marked.
// Depths are those of the sweep above: a vertex is labelled at the first level at which ANY in-neighbour is
// in the frontier; which neighbour is probed first does not matter.
// ---------------------------------------------------------------------------------------------------------
constexpr int kNoMoreNeighbors = static_cast<int>(0x80000000u);

static __global__ void bfs_first_neighbor_kernel(csr_view_t in, int2* __restrict__ head) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < in.n_vertices; v += gridDim.x * blockDim.x) {
    const int s = in.row_offsets[v], e = in.row_offsets[v + 1];
    int2 h = make_int2(-1, -1);
    if (e > s)
      h.x = in.column_indices[s];
    if (e > s + 1)
      h.y = in.column_indices[s + 1] | (e - s == 2 ? kNoMoreNeighbors : 0);
    head[v] = h;
  }
}

/// "is u in the current frontier" for the single-GPU enactor: one bit per vertex.
struct bitmap_frontier_t {
  const unsigned* bm;
  __device__ __forceinline__ bool operator()(int u) const { return bitmap_test(bm, u); }
};

/**
 * @brief K1 (see above).  `next` receives every word of the next frontier this kernel can decide (K2 ORs its
 * finds in afterwards), `retry_map` the vertices K2 has to search.
 */
template <int kThreads, typename FrontierTest>
__global__ void __launch_bounds__(kThreads, 2048 / kThreads)  // 32 registers: a full SM of warps, latency is the enemy
bfs_pull_first_kernel(int n_vertices, const int2* __restrict__ head, unsigned* __restrict__ visited,
                      FrontierTest in_frontier, unsigned* __restrict__ next, unsigned* __restrict__ retry_map,
                      int* dist, int next_level, ctrl_t* ctrl, int* next_count,
                      const unsigned* __restrict__ fill_premark = nullptr, int fill = 0, int source = -1,
                      int batch_words = 32) {
  // batch_words (8 / 16 / 32): words a warp takes per pass.  32 amortises the bookkeeping best; a graph (or a
  // rank's share of one) with fewer than a few passes per resident warp takes smaller batches so that every warp
  // has work -- a pass is a chain of dependent memory round trips, its length is the kernel's floor.
  // fill != 0 (first pull level of a run whose reset left the labels unwritten): also write INT_MAX for every
  // vertex that is not found at this level and was not labelled before -- the unvisited ones this kernel looks
  // at anyway, and the ones pre-marked as having no in-edges.
  constexpr int kWarps = kThreads / 32;
  constexpr int kU = 4;  // vertices per lane per step: kU independent head[] loads, then kU probes in flight
  __shared__ unsigned short s_q[kWarps][1024];  // vertex = (w0 << 5) + entry
  __shared__ unsigned s_found[kWarps][32], s_retry[kWarps][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned short* q = s_q[warp];
  const int words = (n_vertices + 31) / 32;
  const int warps = (gridDim.x * kThreads) >> 5;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  unsigned probes = 0, found_cnt = 0, retry_cnt = 0;
  for (int w0 = gw * batch_words; w0 < words; w0 += warps * batch_words) {
    const int my_wi = w0 + lane;
    const bool mine = lane < batch_words && my_wi < words;  // this lane holds a word of the batch
    const unsigned my_vis = mine ? visited[my_wi] : 0xffffffffu;
    unsigned m = ~my_vis;
    if (my_wi == words - 1 && (n_vertices & 31))
      m &= (1u << (n_vertices & 31)) - 1u;  // bits past the last vertex are not vertices
    const int c = __popc(m);
    const int incl = warp_inclusive_sum(c);
    const int total = __shfl_sync(kFull, incl, 31);
    s_found[warp][lane] = 0;
    s_retry[warp][lane] = 0;
    int at = incl - c;
    while (m) {  // ascending vertex ids: the head[] loads below walk memory forwards
      const int b = __ffs(m) - 1;
      m &= m - 1;
      q[at++] = static_cast<unsigned short>((lane << 5) + b);
    }
    __syncwarp();
    const int vbase = w0 << 5;
    if (fill && fill_premark) {  // vertices without in-edges: never looked at below, never labelled by anybody
      const unsigned my_dead = mine ? fill_premark[my_wi] : 0u;
      unsigned any = __ballot_sync(kFull, my_dead != 0u);
      while (any) {
        const int wl = __ffs(any) - 1;
        any &= any - 1;
        const unsigned bits = __shfl_sync(kFull, my_dead, wl);
        const int v = ((w0 + wl) << 5) + lane;
        if (((bits >> lane) & 1u) && v < n_vertices && v != source)
          dist[v] = INT_MAX;
      }
    }
    for (int i0 = 0; i0 < total; i0 += 32 * kU) {
      int v[kU];
      int2 h[kU];
      bool hit[kU], again[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int i = i0 + 32 * k + lane;
        v[k] = i < total ? vbase + q[i] : -1;
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        h[k] = make_int2(-1, -1);
        if (v[k] >= 0)
          h[k] = __ldg(head + v[k]);
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        hit[k] = h[k].x != -1 && in_frontier(h[k].x);
        probes += h[k].x != -1 ? 1u : 0u;
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        again[k] = h[k].x != -1 && !hit[k] && h[k].y != -1;
        if (again[k]) {
          ++probes;
          hit[k] = in_frontier(h[k].y & ~kNoMoreNeighbors);
        }
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        if (v[k] < 0)
          continue;
        const unsigned bit = 1u << (v[k] & 31);
        const int wl = (v[k] >> 5) - w0;
        if (hit[k]) {
          dist[v[k]] = next_level;
          atomicOr(&s_found[warp][wl], bit);
        } else {
          if (fill)
            dist[v[k]] = INT_MAX;  // K2 overwrites it if it finds the vertex after all
          if (again[k] && !(h[k].y & kNoMoreNeighbors))
            atomicOr(&s_retry[warp][wl], bit);
        }
      }
    }
    __syncwarp();
    if (mine) {
      const unsigned fm = s_found[warp][lane], rm = s_retry[warp][lane];
      next[my_wi] = fm;
      retry_map[my_wi] = rm;
      if (fm)
        visited[my_wi] = my_vis | fm;
      found_cnt += __popc(fm);
      retry_cnt += __popc(rm);
    }
    __syncwarp();  // the queue and the word arrays are rewritten by the next pass
  }
  probes = warp_sum(probes);
  found_cnt = warp_sum(found_cnt);
  retry_cnt = warp_sum(retry_cnt);
  if (lane == 0) {
    if (probes)
      atomicAdd(&ctrl->edges, static_cast<unsigned long long>(probes));
    if (found_cnt)
      atomicAdd(next_count, static_cast<int>(found_cnt));
    if (retry_cnt)
      atomicAdd(&ctrl->hub_count, static_cast<int>(retry_cnt));  // K2's population (statistics only)
  }
}

/// Words per pass of the pull kernels (K1 / K2).  32 everywhere: measured on B200 with maps of 131 K words (RMAT-22)
/// and 262 K words (one of eight ranks of RMAT-26), smaller batches that would keep more warps busy lost -- the first
/// pull level took 0.058 / 0.055 / 0.049 ms with 8 / 16 / 32 words on RMAT-22 (profiles/r2_l_pull_batch_ab.txt): the
/// queue walk amortises better over long batches than idle warps cost.  B2G_PULL_BATCH=8|16|32 keeps the A/B.
inline int pull_batch_words(int /*words*/, int /*warps*/) {
  static const char* force = std::getenv("B2G_PULL_BATCH");
  if (force) {
    const int f = std::atoi(force);
    if (f == 8 || f == 16 || f == 32)
      return f;
  }
  return 32;
}

/// Set v's bit in `map` (RED.OR, no return value).
__device__ __forceinline__ void bitmap_set(unsigned* map, bool on, int v) {
  if (on)
    atomicOr(map + (v >> 5), 1u << (v & 31));
}

/**
 * @brief K2: the full search from the THIRD in-neighbour on, for the vertices K1 marked in `retry_map`
 * (~4 % of the unvisited vertices of the first pull level, fewer than one set bit per word).  A warp takes kWords words
 * per pass, expands their set bits into its own queue in shared memory (warp scan of the popcounts), and
 * walks the queue 32 vertices at a time: kSerial interleaved probes per lane, rows still open after that are
 * finished by the whole warp.  Found vertices set their bits with RED.OR.
 */
template <int kThreads, int kWords, int kSerial, typename FrontierTest>
__global__ void __launch_bounds__(kThreads)
bfs_pull_rest_kernel(csr_view_t in, const unsigned* __restrict__ retry_map, unsigned* visited,
                     FrontierTest in_frontier, unsigned* next, int* dist, int next_level, ctrl_t* ctrl,
                     int* next_count, int batch_words = kWords) {
  static_assert(kWords <= 32, "one lane per word");
  __shared__ int s_q[kThreads / 32][kWords * 32];
  int* q = s_q[threadIdx.x >> 5];
  const int lane = lane_id();
  const int words = (in.n_vertices + 31) / 32;
  const int warps = (gridDim.x * kThreads) >> 5;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int* __restrict__ ro = in.row_offsets;
  const int* __restrict__ ci = in.column_indices;
  unsigned long long scanned = 0;
  int found_cnt = 0;
  for (int w0 = gw * batch_words; w0 < words; w0 += warps * batch_words) {
    const int my_wi = w0 + lane;
    unsigned m = (lane < batch_words && my_wi < words) ? retry_map[my_wi] : 0u;
    const int c = __popc(m);
    const int incl = warp_inclusive_sum(c);
    const int total = __shfl_sync(kFull, incl, 31);
    if (!total)
      continue;
    int at = incl - c;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      q[at++] = (my_wi << 5) + b;
    }
    __syncwarp();
    for (int i0 = 0; i0 < total; i0 += 32) {
      const int i = i0 + lane;
      const int v = i < total ? q[i] : -1;
      int e = 0, end = 0;
      if (v >= 0) {
        e = ro[v] + 2;  // the first two in-neighbours were K1's probes
        end = ro[v + 1];
      }
      bool searching = e < end, found = false;
      for (int s = 0; s < kSerial; ++s) {
        if (searching) {
          const int u = ci[e++];
          ++scanned;
          if (in_frontier(u)) {
            found = true;
            searching = false;
          } else if (e >= end) {
            searching = false;
          }
        }
      }
      unsigned rest = __ballot_sync(kFull, searching);
      while (rest) {
        const int leader = __ffs(rest) - 1;
        rest &= rest - 1;
        const int s0 = __shfl_sync(kFull, e, leader);
        const int t = __shfl_sync(kFull, end, leader);
        bool any = false;
        for (int off = s0; off < t && !any; off += 32) {
          const int idx = off + lane;
          bool mine = false;
          if (idx < t) {
            ++scanned;
            mine = in_frontier(ci[idx]);
          }
          any = __any_sync(kFull, mine);
        }
        if (lane == leader)
          found = any;
      }
      if (found) {
        dist[v] = next_level;
        ++found_cnt;
      }
      bitmap_set(next, found, v);
      bitmap_set(visited, found, v);
    }
    __syncwarp();  // the queue is rewritten by the next pass
  }
  scanned = warp_sum(scanned);
  found_cnt = warp_sum(found_cnt);
  if (lane == 0) {
    if (scanned)
      atomicAdd(&ctrl->edges, scanned);
    if (found_cnt)
      atomicAdd(next_count, found_cnt);
  }
}

struct bfs_level_stat_t {
  int direction;  // 0 = top-down (push), 1 = bottom-up (pull)
  int frontier;   // vertices in the input frontier
  unsigned long long frontier_edges;  // sum of their degrees
  unsigned long long edges_inspected; // column indices actually read
  float kernel_ms = 0.0f;             // device time of this level's advance / sweep kernels
};

struct bfs_config_t {
  advance_launch_t advance;
  int direction = 0;    // 0 push only, 1 pull only (after the first level), 2 optimised
  int use_atomic_min_op = 0;  // 1: the reference's per-edge atomicMin functor
  double alpha = 14.0;  // top-down -> bottom-up when m_f > m_u / alpha
  double beta = 24.0;   // bottom-up -> top-down when n_f < V / beta
};

/// Persistent per-graph BFS scratch (allocated once; nothing is allocated inside run()).
struct bfs_scratch_t {
  dbuf_t<unsigned> visited, fbm, nbm, unreachable;
  dbuf_t<int> unv[2];                    // still-unvisited vertices (consecutive bottom-up levels)
  dbuf_t<unsigned> retry_map;            // pull levels: vertices whose first in-neighbour missed (K2's input)
  dbuf_t<int2> first_nb;                 // per vertex: its first two in-neighbours (bfs_first_neighbor_kernel)
  graph_key_t unreachable_for;           // the (in-edge) graph the unreachable map was built from
  graph_key_t first_nb_for;              // ... and the one first_nb was built from
  dbuf_t<int> q[2];
  dbuf_t<int> counts;  // [0],[1] queue sizes
  struct host_fb_t {
    int count;
    int overflow;
    unsigned long long deg_sum;
    unsigned long long edges;
    volatile int seq;  // written last (after a system fence): the host polls it
  };
  host_fb_t* h_fb = nullptr;  // pinned
  int seq = 0;
  tail_report_t* h_tail = nullptr;  // pinned
  cudaEvent_t ev[128] = {};   // per-level event pairs (first 64 levels are timed)
  ~bfs_scratch_t() {
    if (h_fb)
      cudaFreeHost(h_fb);
    if (h_tail)
      cudaFreeHost(h_tail);
    for (auto e : ev)
      if (e)
        cudaEventDestroy(e);
  }
  void ensure(int V) {
    size_t words = (static_cast<size_t>(V) + 31) / 32 + 4;
    visited.ensure(words);
    fbm.ensure(words);
    nbm.ensure(words);
    q[0].ensure(static_cast<size_t>(V) + 64);
    q[1].ensure(static_cast<size_t>(V) + 64);
    counts.ensure(8);
    if (!h_fb) {
      B2G_CHECK(cudaMallocHost(&h_fb, sizeof(host_fb_t)));
      h_fb->seq = 0;
    }
    if (!h_tail) {
      B2G_CHECK(cudaMallocHost(&h_tail, sizeof(tail_report_t)));
      h_tail->seq = 0;
    }
    if (!ev[0])
      for (auto& e : ev)
        B2G_CHECK(cudaEventCreate(&e));
  }
};

static __global__ void bfs_feedback_kernel(const int* count, const ctrl_t* a, const ctrl_t* b,
                                    bfs_scratch_t::host_fb_t* fb, int seq) {
  fb->count = *count;
  unsigned long long ds = a ? a->deg_sum : 0, ed = a ? a->edges : 0;
  int ov = a ? a->overflow : 0;
  if (b) {
    ds += b->deg_sum;
    ed += b->edges;
    ov |= b->overflow;
  }
  fb->deg_sum = ds;
  fb->edges = ed;
  fb->overflow = ov;
  __threadfence_system();
  fb->seq = seq;
}

/**
 * @brief Run BFS from `source` into `dist` (device, V ints).  `out_g` is the CSR, `in_g` the
 * transpose used by the bottom-up sweep (pass the CSR again for symmetric graphs; pass a view with
 * row_offsets == nullptr to disable pull).  Returns the number of levels executed.
 */
inline int bfs_run(workspace_t& ws, bfs_scratch_t& sc, const csr_view_t& out_g,
                   const csr_view_t& in_g, int source, int* dist, const bfs_config_t& cfg,
                   std::vector<bfs_level_stat_t>* levels = nullptr) {
  const int V = out_g.n_vertices;
  const int sms = device_info_t::get().sm_count;
  sc.ensure(V);
  cudaStream_t st = ws.stream;
  const int words = (V + 31) / 32;
  const bool can_pull =
      in_g.row_offsets != nullptr && cfg.direction != 0 && !cfg.use_atomic_min_op;
  const unsigned* premark = nullptr;
  // B2G_BFS_PULL_LEGACY=1: the first-generation pull kernels (one sweep / list kernel per level), kept for A/B runs
  static const bool legacy_pull = std::getenv("B2G_BFS_PULL_LEGACY") != nullptr;
  if (can_pull) {  // per-graph map of vertices without in-edges (built once, like the transpose)
    if (!sc.unreachable_for.matches(in_g)) {  // keyed on graph identity + addresses + sizes, not an address
      sc.unreachable.ensure(static_cast<size_t>(words) + 4);
      bfs_unreachable_map_kernel<<<sms * 8, 256, 0, st>>>(in_g.row_offsets, V, sc.unreachable.ptr);
      sc.unreachable_for.set(in_g);
      ws.launches += 1;
    }
    premark = sc.unreachable.ptr;
    if (!legacy_pull && !sc.first_nb_for.matches(in_g)) {  // per-graph shortcut array of the pull levels
      sc.first_nb.ensure(static_cast<size_t>(V) + 64);
      bfs_first_neighbor_kernel<<<sms * 8, 256, 0, st>>>(in_g, sc.first_nb.ptr);
      sc.first_nb_for.set(in_g);
      ws.launches += 1;
    }
  }
  // Direction-optimised runs leave the labels of unreached vertices to the first pull level (bfs_reset_kernel)
  static const bool eager_fill = std::getenv("B2G_BFS_EAGER_FILL") != nullptr;
  bool dist_filled = !(can_pull && !legacy_pull && !eager_fill);
  bfs_reset_kernel<<<sms * 8, 256, 0, st>>>(dist, sc.visited.ptr, sc.fbm.ptr, V, source,
                                             sc.q[0].ptr, sc.counts.ptr, premark, dist_filled ? 1 : 0);
  ws.launches += 1;
  int cur = 0;
  int level = 0;
  bool bottom_up = false;      // representation of the current frontier: queue (false) / bitmap
  long long n_f = 1;           // frontier vertices
  // frontier out-degree sum.  Not known for the source (probing it would cost a host round trip);
  // level 0 always runs top-down on the small path and reports it.
  unsigned long long m_f = 0;
  bool explored_counted = false;
  bool m_known = false;  // the pull kernels do not report the new frontier's out-degree sum (it would cost a pair
                         // of row offsets per found vertex); the first push level after them runs like level 0
  unsigned long long explored = 0;
  unsigned* fbm = sc.fbm.ptr;
  unsigned* nbm = sc.nbm.ptr;
  bool unv_valid = false;  // sc.unv[unv_cur] lists exactly the vertices still unvisited
  int unv_cur = 0;
  // B2G_TRACE=1: host-side timeline of every level (microseconds since the run started)
  static const bool trace = std::getenv("B2G_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto now_us = [&]() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  };
  while (n_f > 0) {
    const double t_begin = trace ? now_us() : 0.0;
    // ---- tiny queue frontier: run the tail of the traversal in one single-CTA launch ------------
    if (level > 0 && !bottom_up && m_known && !cfg.use_atomic_min_op &&
        static_cast<long long>(m_f) < cfg.advance.small_frontier_edges) {
      if (level < 64)
        B2G_CHECK(cudaEventRecord(sc.ev[2 * level], st));
      advance_tail_kernel<1024, false><<<1, 1024, 0, st>>>(
          out_g, sc.q[0].ptr, sc.q[1].ptr, sc.counts.ptr, cur, level, 16,
          static_cast<unsigned long long>(cfg.advance.small_frontier_edges),
          bfs_claim_maker{sc.visited.ptr, dist}, sc.h_tail, ++sc.seq);
      if (level < 64)
        B2G_CHECK(cudaEventRecord(sc.ev[2 * level + 1], st));
      ws.launches += 1;
      wait_for_sequence(&sc.h_tail->seq, sc.seq, st);
      const tail_report_t& t = *sc.h_tail;
      for (int k = 0; k < t.levels; ++k) {
        explored += t.edges[k];
        if (levels)
          levels->push_back({0, t.frontier[k], t.edges[k], t.edges[k]});
      }
      // events were recorded once for the whole launch: later levels of it report 0 ms
      for (int k = 1; k < t.levels && level + k < 64; ++k) {
        B2G_CHECK(cudaEventRecord(sc.ev[2 * (level + k)], st));
        B2G_CHECK(cudaEventRecord(sc.ev[2 * (level + k) + 1], st));
      }
      level += t.levels;
      cur = t.cur;
      n_f = t.count;
      m_f = t.deg_sum;
      m_known = true;
      unv_valid = false;
      continue;
    }
    // ---- choose direction for this level (Beamer et al.) --------------------------------
    bool want_bottom_up = bottom_up;
    if (can_pull && level > 0) {
      unsigned long long m_u = static_cast<unsigned long long>(out_g.n_edges) - explored;
      if (cfg.direction == 1)
        want_bottom_up = level > 0;
      else if (!bottom_up)
        want_bottom_up = m_known && static_cast<double>(m_f) > static_cast<double>(m_u) / cfg.alpha;
      else
        want_bottom_up = !(static_cast<double>(n_f) < static_cast<double>(V) / cfg.beta);
    }
    if (level > 0 && !explored_counted)
      explored += m_f;
    explored_counted = false;
    if (level < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * level], st));
    ctrl_t* ca = nullptr;
    ctrl_t* cb = nullptr;
    const int* count_ptr;
    if (want_bottom_up) {
      if (!bottom_up) {  // queue -> bitmap
        B2G_CHECK(cudaMemsetAsync(fbm, 0, sizeof(unsigned) * words, st));
        queue_to_bitmap_kernel<<<sms * 4, 256, 0, st>>>(sc.q[cur].ptr, sc.counts.ptr + cur, fbm);
        ws.launches += 1;
      }
      ca = ws.next_ctrl();
      if (legacy_pull) {
        sc.unv[0].ensure(static_cast<size_t>(V) + 64);
        sc.unv[1].ensure(static_cast<size_t>(V) + 64);
      }
      B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + 2, 0, sizeof(int), st));
      if (!legacy_pull) {
        // K1 (one probe per unvisited vertex, from first_nb) + K2 (full search for K1's misses): see above
        cb = ws.next_ctrl();
        sc.retry_map.ensure(static_cast<size_t>(words) + 4);
        const bitmap_frontier_t in_frontier{fbm};
        bfs_pull_first_kernel<256><<<sms * 8, 256, 0, st>>>(V, sc.first_nb.ptr, sc.visited.ptr, in_frontier, nbm,
                                                            sc.retry_map.ptr, dist, level + 1, ca,
                                                            sc.counts.ptr + 2, premark, dist_filled ? 0 : 1, source,
                                                            pull_batch_words(words, sms * 64));
        dist_filled = true;
        bfs_pull_rest_kernel<256, 32, 8><<<sms * 6, 256, 0, st>>>(in_g, sc.retry_map.ptr, sc.visited.ptr, in_frontier,
                                                                  nbm, dist, level + 1, cb, sc.counts.ptr + 2,
                                                                  pull_batch_words(words, sms * 48));
        ws.launches += 1;
      } else if (!unv_valid) {  // first pull level of a run of pull levels: sweep every visited word
        B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + 4, 0, sizeof(int), st));
        bfs_bottom_up_kernel<256, 8><<<sms * 8, 256, 0, st>>>(
            in_g, sc.visited.ptr, fbm, nbm, dist, level + 1, ca, sc.counts.ptr + 2,
            sc.unv[0].ptr, sc.counts.ptr + 4);
        unv_cur = 0;
        unv_valid = true;
      } else {  // later pull levels: dense list of the vertices still unvisited
        const int o = unv_cur ^ 1;
        B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + 4 + o, 0, sizeof(int), st));
        B2G_CHECK(cudaMemsetAsync(nbm, 0, sizeof(unsigned) * words, st));
        bfs_bottom_up_list_kernel<256, 8><<<sms * 8, 256, 0, st>>>(
            in_g, sc.unv[unv_cur].ptr, sc.counts.ptr + 4 + unv_cur, sc.visited.ptr, fbm,
            nbm, dist, level + 1, ca, sc.counts.ptr + 2, sc.unv[o].ptr, sc.counts.ptr + 4 + o);
        unv_cur = o;
      }
      ws.launches += 1;
      unsigned* t = fbm;
      fbm = nbm;
      nbm = t;
      bottom_up = true;
      count_ptr = sc.counts.ptr + 2;
    } else {
      unv_valid = false;  // a push level claims vertices the list does not know about
      if (bottom_up) {  // bitmap -> queue
        B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + cur, 0, sizeof(int), st));
        ctrl_t* cq = ws.next_ctrl();
        bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(fbm, words, sc.q[cur].ptr, sc.counts.ptr + cur,
                                                        m_known ? nullptr : out_g.row_offsets,
                                                        m_known ? nullptr : &cq->deg_sum);
        ws.launches += 1;
        bottom_up = false;
        if (!m_known) {
          // one more pinned-memory poll, once per run: with the queue's out-degree sum known the tail of the
          // traversal (usually everything that is left) runs in the single-launch multi-level kernel
          bfs_feedback_kernel<<<1, 1, 0, st>>>(sc.counts.ptr + cur, cq, nullptr, sc.h_fb, ++sc.seq);
          ws.launches += 1;
          wait_for_sequence(&sc.h_fb->seq, sc.seq, st);
          m_f = sc.h_fb->deg_sum;
          m_known = true;
          explored_counted = true;
          if (level < 64)  // the level's event pair is re-recorded when the loop comes back to this level
            B2G_CHECK(cudaEventRecord(sc.ev[2 * level + 1], st));
          continue;
        }
      }
      int nxt = cur ^ 1;
      B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + nxt, 0, sizeof(int), st));
      int ub = static_cast<int>(n_f < V ? n_f : V);
      advance_launch_t lcfg = cfg.advance;
      lcfg.avg_degree = (level > 0 && m_known && n_f > 0) ? static_cast<double>(m_f) / static_cast<double>(n_f) : 0.0;
      if (level == 0 || !m_known) {
        lcfg.lb = lb_t::block_mapped;  // rows of unknown total length: binned kernel + hub slabs
      } else if (lcfg.lb == lb_t::merge_path &&
                 static_cast<long long>(m_f) < cfg.advance.mid_frontier_edges) {
        lcfg.lb = lb_t::block_mapped;  // mid-size frontier: skip the scan + partition launches
      } else if (static_cast<long long>(m_f) < cfg.advance.small_frontier_edges) {
        lcfg.lb = lb_t::block_mapped;  // one kernel: warp/thread bins only
        lcfg.hub_threshold = 1 << 30;
      }
      if (cfg.use_atomic_min_op) {
        bfs_atomic_min_op op{dist, level + 1};
        launch_advance<advance_output_t::vertices, true, false>(
            ws, out_g, sc.q[cur].ptr, sc.counts.ptr + cur, ub, sc.q[nxt].ptr,
            sc.counts.ptr + nxt, V, op, lcfg, &ca);
      } else {
        bfs_claim_op op{sc.visited.ptr, dist, level + 1};
        launch_advance<advance_output_t::vertices, true, false>(
            ws, out_g, sc.q[cur].ptr, sc.counts.ptr + cur, ub, sc.q[nxt].ptr,
            sc.counts.ptr + nxt, V, op, lcfg, &ca);
      }
      cur = nxt;
      count_ptr = sc.counts.ptr + cur;
    }
    if (level < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * level + 1], st));
    bfs_feedback_kernel<<<1, 1, 0, st>>>(count_ptr, ca, cb, sc.h_fb, ++sc.seq);
    ws.launches += 1;
    const double t_enq = trace ? now_us() : 0.0;
    wait_for_sequence(&sc.h_fb->seq, sc.seq, st);
    if (trace)
      std::fprintf(stderr, "[b2g] level %d: begin %.1f enqueued %.1f feedback %.1f (n_f=%lld m_f=%llu)\n",
                   level, t_begin, t_enq, now_us(), n_f, m_f);
    if (sc.h_fb->overflow)
      throw std::runtime_error("bfs: output frontier overflow");
    if (levels)
      levels->push_back({want_bottom_up ? 1 : 0, static_cast<int>(n_f), m_f, sc.h_fb->edges});
    if (level == 0)
      explored += sc.h_fb->edges;  // the source's degree, learnt from the level it just ran
    n_f = sc.h_fb->count;
    m_f = sc.h_fb->deg_sum;
    m_known = !want_bottom_up || legacy_pull;  // pull levels (second generation) report no out-degree sum
    if (!m_known)
      m_f = sc.h_fb->edges;  // a stand-in for the `explored` estimate of the direction heuristic only
    ++level;
  }
  if (!dist_filled) {  // no pull level happened: write the unreached vertices' labels now
    bfs_fill_unreached_kernel<<<sms * 8, 256, 0, st>>>(dist, sc.visited.ptr, premark, V, source);
    ws.launches += 1;
  }
  if (levels)
    for (int l = 0; l < level && l < 64; ++l)
      cudaEventElapsedTime(&(*levels)[l].kernel_ms, sc.ev[2 * l], sc.ev[2 * l + 1]);
  return level;
}

}  // namespace b200
}  // namespace gunrock
