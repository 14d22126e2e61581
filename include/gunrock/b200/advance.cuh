/**
 * @file advance.cuh
 * @brief Blackwell-native neighbour-expansion (advance) kernels with in-kernel output compaction.
 *
 * These replace the reference load balancers
 *   block_mapped_kernel   include/gunrock/framework/operators/advance/block_mapped.hxx:67-191
 *   merge_path_kernel     include/gunrock/framework/operators/advance/merge_path.hxx:112-280
 *   thread_mapped lambda  include/gunrock/framework/operators/advance/thread_mapped.hxx:58-81
 * with the same operator contract (advance.hxx:35-49): `op(src, dst, edge, weight) -> bool` is
 * called exactly once per (input-frontier occurrence, out-edge); a `true` puts `dst` in the output
 * frontier.  Differences by design:
 *   - the output frontier is COMPACT: rejected edges occupy no slot (the reference writes -1 and
 *     leaves the culling to a second full pass, filter/predicated.hxx:30).  Compaction is done
 *     in-kernel with warp ballots into a per-warp shared-memory staging buffer that is flushed with
 *     one global atomicAdd per ~100 vertices;
 *   - frontier sizes live in device memory, no host round trip between operators;
 *   - power-law rows are degree-binned: rows >= hub_threshold are deferred to a grid-wide bin whose
 *     column-index slabs are staged into shared memory with cp.async.bulk (TMA engine) and walked
 *     by the CTA; rows >= 32 are walked by a warp; short rows are packed by a warp scan;
 *   - grids are persistent (multiples of the SM count) and fetch work through an atomic cursor.
 */
#pragma once

#include <type_traits>
#include <utility>

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>
#include <gunrock/b200/scan.cuh>

namespace gunrock {
namespace b200 {

enum class advance_input_t { vertices, graph };
enum class advance_output_t { vertices, edges, none };

/**
 * Optional two-phase protocol for edge functors.  A functor that also provides
 *     token_t prefetch(int dst) const;                          // loads only, no side effects
 *     bool    commit(int src, int dst, int edge, float w, token_t) const;
 * lets the kernels issue the prefetches of several edges back to back (memory-level parallelism)
 * before any of the dependent atomics.  Plain `bool operator()(src, dst, edge, w)` functors (user
 * lambdas) are called as they are.
 */
template <typename Op, typename = void>
struct op_traits {
  static constexpr bool two_phase = false;
  using token_t = int;
};
template <typename Op>
struct op_traits<Op, std::void_t<decltype(&Op::commit)>> {
  static constexpr bool two_phase = true;
  using token_t = decltype(std::declval<const Op&>().prefetch(0));
};
template <typename Op>
__device__ __forceinline__ typename op_traits<Op>::token_t op_prefetch(const Op& op, int dst) {
  if constexpr (op_traits<Op>::two_phase)
    return op.prefetch(dst);
  else
    return 0;
}
template <typename Op>
__device__ __forceinline__ bool op_commit(const Op& op, int src, int dst, int e, float w,
                                          typename op_traits<Op>::token_t tok) {
  if constexpr (op_traits<Op>::two_phase)
    return op.commit(src, dst, e, w, tok);
  else
    return op(src, dst, e, w);
}
/// Optional `int emit_as(int dst) const`: the value stored in the output frontier for a kept
/// vertex (the partitioned BFS stores local row ids while column indices are global ids).
template <typename Op, typename = void>
struct op_has_emit_as : std::false_type {};
template <typename Op>
struct op_has_emit_as<Op, std::void_t<decltype(&Op::emit_as)>> : std::true_type {};
template <typename Op>
__device__ __forceinline__ int op_emit(const Op& op, int dst) {
  if constexpr (op_has_emit_as<Op>::value)
    return op.emit_as(dst);
  else
    return dst;
}
/// Optional `static constexpr bool kNeedsSource = false`: the functor ignores its `src` argument,
/// so kernels need not stage source ids (saves 8 KiB of shared memory per merge_path CTA, which
/// goes to L1 and raises the hit rate of the bitmap probes).
template <typename Op, typename = void>
struct op_needs_source : std::true_type {};
template <typename Op>
struct op_needs_source<Op, std::void_t<decltype(Op::kNeedsSource)>>
    : std::integral_constant<bool, Op::kNeedsSource> {};
constexpr int kBatch = 4;  // 32-edge chunks whose loads are issued back to back per warp

/// Optional `static constexpr int kMergePathKernel = 1 | 4`: which merge_path kernel serves this functor when the
/// caller does not say (advance_launch_t::variant < 0): 0 = the CTA-tile kernel (advance_merge_path_kernel),
/// 1 / 4 = warp-private spans (advance_warp_path_kernel) with 4 / 8 chunks in flight.  Functors that do not
/// declare it (user lambdas, the partitioned paths' functors) get the CTA-tile kernel and do not instantiate the
/// others.  Measured on the bench graphs (profiles/r2_bench_matrix.md): BFS push RMAT-22 0.776 -> 0.718 ms with 1,
/// SSSP RMAT-24 7.49 -> 7.06 ms with 4.
template <typename Op, typename = void>
struct op_merge_path_kernel : std::integral_constant<int, 0> {};
template <typename Op>
struct op_merge_path_kernel<Op, std::void_t<decltype(Op::kMergePathKernel)>>
    : std::integral_constant<int, Op::kMergePathKernel> {};

/// Per-warp staging buffer: ballot-compacted appends, flushed with one global atomic.
template <int kCap, bool kDegSum>
struct warp_emitter_t {
  int* s_buf;  // this warp's kCap ints of shared memory
  int cnt;     // warp-uniform fill level
  int* out;
  int* out_count;
  int out_capacity;
  const int* row_offsets;
  ctrl_t* ctrl;

  __device__ __forceinline__ void init(int* smem, int* out_, int* out_count_, int cap,
                                       const int* ro, ctrl_t* c) {
    s_buf = smem;
    cnt = 0;
    out = out_;
    out_count = out_count_;
    out_capacity = cap;
    row_offsets = ro;
    ctrl = c;
  }
  /// Must be called by all 32 lanes (converged).
  __device__ __forceinline__ void push(bool keep, int item) {
    unsigned m = __ballot_sync(kFull, keep);
    if (m) {
      if (keep)
        s_buf[cnt + __popc(m & lanemask_lt())] = item;
      cnt += __popc(m);
      if (cnt > kCap - 32)
        flush();
    }
  }
  __device__ __forceinline__ void flush() {
    __syncwarp();
    if (cnt) {
      int base = 0;
      if (lane_id() == 0)
        base = atomicAdd(out_count, cnt);
      base = __shfl_sync(kFull, base, 0);
      if (base + cnt > out_capacity) {  // never write past the frontier buffer; host raises
        if (lane_id() == 0)
          ctrl->overflow = 1;
        cnt = max(0, out_capacity - base);
      }
      unsigned long long ds = 0;
      for (int i = lane_id(); i < cnt; i += 32) {
        int v = s_buf[i];
        out[base + i] = v;
        if (kDegSum)
          ds += static_cast<unsigned>(row_offsets[v + 1] - row_offsets[v]);
      }
      if (kDegSum) {
        ds = warp_sum(ds);
        if (lane_id() == 0 && ds)
          atomicAdd(&ctrl->deg_sum, ds);
      }
    }
    __syncwarp();
    cnt = 0;
  }
};

/// One slab of a deferred hub row: kChunk (or fewer) consecutive edges of one row.
struct __align__(16) hub_slab_t {
  int e0;   // CSR position of the slab's first edge
  int cnt;  // edges in the slab (<= kChunk); <= 0 marks "no more work" in the kernels' shared-memory ring
  int src;  // the row's vertex
  int pad;
};

struct advance_params_t {
  csr_view_t g;
  const int* in = nullptr;        // input frontier (null when input is the whole graph)
  const int* in_count = nullptr;  // device count (null when input is the whole graph)
  int* out = nullptr;
  int* out_count = nullptr;
  int out_capacity = 0;
  int* hubs = nullptr;  // deferred rows
  int hub_capacity = 0;
  ctrl_t* ctrl = nullptr;
  int hub_threshold = 1 << 30;
  int tma_ok = 0;                  // column_indices / values are 16-byte aligned
  int entries_per_ticket = 256;    // block_mapped: frontier entries a CTA draws at a time (<= 256)
  int hub_slab_capacity = 0;
  hub_slab_t* hub_slabs = nullptr;  // block_mapped: slab table of the deferred rows (advance_hub_table_kernel)
  const int* tile_rows = nullptr;  // merge_path: first row of every tile
  const int* row_base = nullptr;   // merge_path: CSR offset of every frontier row (next to the scan)
};

constexpr int kEmitCap = 128;  // ints per warp in the staging buffer

/// Shared-memory ints one warp of advance_warp_path_kernel owns (must match the kernel's layout).
template <int kSpan, bool kSrc>
constexpr int warp_path_ints() {
  return kEmitCap + (kSpan + 36) + (kSpan + 36) / 2 + (kSrc ? (kSpan + 36) : 0);
}

/**
 * @brief "block_mapped": equal number of frontier entries per CTA, as in the reference
 * (block_mapped.hxx:67-191: CTA loads 256 entries, block-scans their degrees, threads stride over
 * the CTA's edge range and binary-search the owner per edge), rebuilt around the merge_path walk:
 *   - a persistent CTA draws 256 entries per ticket; rows >= hub_threshold (or >= 256 edges when the
 *     CTA's share would exceed kCtaBudget) are deferred to the grid bin (TMA slab kernel);
 *   - the remaining degrees are block-scanned and the non-empty rows compacted into shared memory;
 *   - every warp then walks 256-rank spans of the CTA's rank space: one binary search per span, row
 *     starts inside each 32 ranks turned into a bit mask with one REDUX, kBatch chunks of loads in
 *     flight -- no per-edge search, no separate warp / thread bins.
 * Four block barriers per 256 entries; shared memory 7 KiB, so L1 keeps ~200 KiB for the probes.
 */
template <int kThreads, advance_input_t kIn, advance_output_t kOut, bool kDegSum, bool kWeights,
          typename Op>
// Resident CTAs per SM the kernel is compiled for.  Measured (profiles/r2_r_occupancy_ab.txt): the unweighted
// functors (BFS claims) gain 7 % at 6 CTAs / 40 registers (BFS push RMAT-22 block_mapped 0.788 -> 0.733 ms, no
// spills); the weighted ones (SSSP relax) LOSE at 40 or 32 registers (8.06 -> 9.57 / 8.18 ms: fewer loads in flight
// per thread cost more than the extra warps hide) and keep their natural 48.
#ifdef B2G_BINNED_CTAS  // A/B builds only (-DB2G_BINNED_CTAS=n)
__global__ void __launch_bounds__(kThreads, B2G_BINNED_CTAS)
#else
__global__ void __launch_bounds__(kThreads, kWeights ? 5 : 6)
#endif
advance_binned_kernel(advance_params_t p, Op op) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kSpan = 256;            // ranks per warp span
  constexpr int kCtaBudget = 1 << 16;   // edges a CTA keeps for itself per ticket
  constexpr int kSpill = 256;           // rows at least this long are deferred when over budget
  constexpr bool kSrc = op_needs_source<Op>::value;
  __shared__ int s_emit[kWarps][kEmitCap];
  __shared__ int s_rank[kThreads + 36];  // first rank of each live row of this ticket
  __shared__ int s_base[kThreads + 36];  // (CSR offset of the row's first edge) - (its first rank)
  __shared__ int s_vert[kSrc ? kThreads + 36 : 1];
  __shared__ int s_wsum[kWarps], s_wlive[kWarps], s_wlong[kWarps];
  __shared__ int s_ticket;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int* __restrict__ ro = p.g.row_offsets;
  const int* __restrict__ ci = p.g.column_indices;
  const float* __restrict__ vals = p.g.values;
  const int n = (kIn == advance_input_t::graph) ? p.g.n_vertices : *p.in_count;

  warp_emitter_t<kEmitCap, kDegSum> em;
  em.init(s_emit[warp], p.out, p.out_count, p.out_capacity, ro, p.ctrl);
  unsigned long long edges_seen = 0;

  // Entries per ticket are chosen by the host from the frontier's average degree so that a ticket
  // is worth ~16K edges: small enough to balance hub-heavy frontiers over ~10^3 resident CTAs,
  // large enough to amortise the four barriers.
  const int ept = min(kThreads, max(1, p.entries_per_ticket));
  if (threadIdx.x == 0)
    s_ticket = atomicAdd(&p.ctrl->work, ept);
  for (;;) {
    __syncthreads();  // [A] previous ticket retired, new ticket visible
    const int base = s_ticket;
    if (base >= n)
      break;
    const int idx = base + threadIdx.x;
    int v = -1;
    if (idx < n && threadIdx.x < ept)
      v = (kIn == advance_input_t::graph) ? idx : p.in[idx];
    int start = 0, deg = 0;
    if (v >= 0) {
      start = ro[v];
      deg = ro[v + 1] - start;
    }
    edges_seen += static_cast<unsigned>(deg);
    // ---- grid bin --------------------------------------------------------------------------
    const int wl = warp_sum(deg >= kSpill ? deg : 0);
    if (lane == 0)
      s_wlong[warp] = wl;
    __syncthreads();  // [B1]
    int cta_long = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
      cta_long += s_wlong[w];
    const int defer_at = (cta_long > kCtaBudget && p.hub_threshold < (1 << 30))
                             ? min(p.hub_threshold, kSpill)
                             : p.hub_threshold;
    if (deg >= defer_at) {
      int slot = atomicAdd(&p.ctrl->hub_count, 1);
      if (slot < p.hub_capacity) {  // list full (duplicate-heavy frontier): keep the row here
        p.hubs[slot] = v;
        deg = 0;
      }
    }
    // ---- block scan of the degrees + compaction of the live rows ---------------------------
    const int incl = warp_inclusive_sum(deg);
    const unsigned live_m = __ballot_sync(kFull, deg > 0);
    if (lane == 31)
      s_wsum[warp] = incl;
    if (lane == 0)
      s_wlive[warp] = __popc(live_m);
    __syncthreads();  // [B2]
    int rank0 = incl - deg, slot = __popc(live_m & lanemask_lt()), total = 0, nrows = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      int ws = s_wsum[w], wc = s_wlive[w];
      if (w < warp) {
        rank0 += ws;
        slot += wc;
      }
      total += ws;
      nrows += wc;
    }
    if (deg > 0) {
      s_rank[slot] = rank0;
      s_base[slot] = start - rank0;
      if (kSrc)
        s_vert[slot] = v;
    }
    if (threadIdx.x < 33)
      s_rank[nrows + threadIdx.x] = total;  // sentinels
    __syncthreads();  // [C]
    if (threadIdx.x == 0)
      s_ticket = atomicAdd(&p.ctrl->work, ept);  // next ticket, read after barrier [A]
    // ---- walk: warp w takes spans w, w + kWarps, ... of the CTA's rank space ---------------
    for (int w_begin = warp * kSpan; w_begin < total; w_begin += kWarps * kSpan) {
      const int w_end = min(total, w_begin + kSpan);
      int a = 0, b = nrows;  // last live row with s_rank <= w_begin
      while (b - a > 1) {
        int mid = (a + b) >> 1;
        if (s_rank[mid] <= w_begin)
          a = mid;
        else
          b = mid;
      }
      for (int r0 = w_begin; r0 < w_end; r0 += 32 * kBatch) {
        int row[kBatch], e[kBatch], nb[kBatch], u[kBatch];
        float w[kBatch];
        bool valid[kBatch], keep[kBatch];
        typename op_traits<Op>::token_t tok[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          const int rk = r0 + 32 * k;
          int nxt = s_rank[min(a + 1 + lane, nrows + 32)] - rk;
          unsigned bit = (nxt >= 0 && nxt < 32) ? (1u << nxt) : 0u;
          unsigned starts = __reduce_or_sync(kFull, bit);
          row[k] = min(a + __popc(starts & (0xffffffffu >> (31 - lane))), nrows - 1);
          valid[k] = rk + lane < w_end;
          a += __popc(starts);
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          u[k] = kSrc ? s_vert[row[k]] : -1;
          e[k] = s_base[row[k]] + r0 + 32 * k + lane;
          nb[k] = valid[k] ? ld_stream(ci + e[k]) : -1;
          w[k] = (kWeights && vals && valid[k]) ? ld_stream(vals + e[k]) : 1.0f;
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
          if (valid[k])
            tok[k] = op_prefetch(op, nb[k]);
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
          keep[k] = valid[k] && op_commit(op, u[k], nb[k], e[k], w[k], tok[k]);
        if (kOut != advance_output_t::none) {
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
            em.push(keep[k], kOut == advance_output_t::edges ? e[k] : op_emit(op, nb[k]));
        }
      }
    }
  }
  if (kOut != advance_output_t::none)
    em.flush();
  edges_seen = warp_sum(edges_seen);
  if (lane == 0 && edges_seen)
    atomicAdd(&p.ctrl->edges, edges_seen);
}

/**
 * @brief Grid bin, step 1: cut every deferred hub row into kChunk-edge slabs.  One warp per hub row reserves
 * a range of the slab table with ONE atomic and writes the row's descriptors; the running total stays in
 * ctrl->pad[0].  (The first version dealt batches of 512 hub rows to fixed groups of CTAs: on a power-law
 * frontier the first batches hold the rows with hundreds of slabs and the last ones rows with two, so most
 * CTAs idled -- SSSP RMAT-24 block_mapped ran 11.1 ms against merge_path's 7.5.  A table of slabs can be dealt
 * evenly whatever the degrees are.)
 */
template <int kChunk>
__global__ void advance_hub_table_kernel(advance_params_t p) {
  const int lane = lane_id();
  const int n_hubs = min(p.ctrl->hub_count, p.hub_capacity);
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int* __restrict__ ro = p.g.row_offsets;
  for (int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < n_hubs; h += warps) {
    const int v = p.hubs[h];
    const int s = ro[v], e = ro[v + 1];
    const int c = (e - s + kChunk - 1) / kChunk;
    int base = 0;
    if (lane == 0) {
      base = atomicAdd(&p.ctrl->pad[0], c);
      if (base + c > p.hub_slab_capacity) {  // a frontier with the same hub many times over: the host sized the
        atomicAdd(&p.ctrl->pad[0], -c);     // table from the graph's edge count (edges_upper_bound says better)
        p.ctrl->overflow = 1;
        base = -1;
      }
    }
    base = __shfl_sync(kFull, base, 0);
    if (base < 0)
      continue;
    for (int j = lane; j < c; j += 32) {
      hub_slab_t d;
      d.e0 = s + j * kChunk;
      d.cnt = min(kChunk, e - d.e0);
      d.src = v;
      d.pad = 0;
      p.hub_slabs[base + j] = d;
    }
  }
}

/**
 * @brief Grid bin, step 2: a persistent grid walks the slab table.  A CTA draws kRange consecutive slabs per
 * atomic (warp 0 fetches the NEXT range's descriptors into a two-deep ring in shared memory while the CTA
 * walks the current one), and every slab is staged HBM -> shared memory by ONE thread issuing a cp.async.bulk
 * (TMA engine, double buffered on two mbarriers: slab i+1 is in flight while slab i is walked) and then walked
 * by the CTA, kHB edges per thread with all their loads in flight at once (two-phase functors).
 * When the CSR arrays are not 16-byte aligned (p.tma_ok == 0) the slab is read with plain coalesced loads.
 */
template <int kThreads, int kChunk, advance_output_t kOut, bool kDegSum, bool kWeights, typename Op>
#ifndef B2G_HUB_CTAS_W
#define B2G_HUB_CTAS_W 4  // resident CTAs per SM the weighted slab kernel is compiled for (A/B builds: -DB2G_HUB_CTAS_W=n)
#endif
__global__ void __launch_bounds__(kThreads, kWeights ? B2G_HUB_CTAS_W : 5)
advance_hub_kernel(advance_params_t p, Op op) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kHB = kChunk / kThreads;  // slab edges per thread: all of them in flight at once
  constexpr int kRange = 8;               // most slabs per ticket (ring slots)
  constexpr int kSlab = kChunk + 4;       // +4: slabs start at a 16-byte aligned column index
  static_assert(kChunk % 4 == 0 && kRange <= 32, "slab size must keep 16-byte granularity");
  __shared__ int s_emit[kWarps][kEmitCap];
  __shared__ hub_slab_t s_desc[2][kRange];
  __shared__ __align__(16) int s_idx[2][kSlab];
  __shared__ __align__(16) float s_val[kWeights ? 2 : 1][kWeights ? kSlab : 4];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int* __restrict__ ro = p.g.row_offsets;
  const int* __restrict__ ci = p.g.column_indices;
  const float* __restrict__ vals = p.g.values;
  const bool use_vals = kWeights && vals != nullptr;
  const bool tma = p.tma_ok != 0;
  const int total = p.ctrl->pad[0];  // slabs in the table (advance_hub_table_kernel)
  if (total == 0)
    return;
  // slabs per ticket: 8 when there is plenty of work, fewer when the table is short (one hub row of a BFS
  // source is ~300 slabs: ranges of 8 would leave all but 36 CTAs idle); never below 2 (ring visibility below)
  const int range = min(kRange, max(2, total / static_cast<int>(gridDim.x)));

  warp_emitter_t<kEmitCap, kDegSum> em;
  em.init(s_emit[warp], p.out, p.out_count, p.out_capacity, ro, p.ctrl);

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  // warp 0: draw a range of slabs and put its descriptors (or end markers) into ring slot `slot`
  auto fetch_range = [&](int slot) {
    int base = 0;
    if (lane == 0)
      base = atomicAdd(&p.ctrl->tile, range);
    base = __shfl_sync(kFull, base, 0);
    if (lane < kRange) {  // slots past `range` hold end markers and are never reached
      hub_slab_t d;
      d.e0 = 0, d.cnt = -1, d.src = -1, d.pad = 0;
      if (lane < range && base + lane < total)
        d = p.hub_slabs[base + lane];
      s_desc[slot][lane] = d;
    }
  };
  // The bulk copy moves whole 16-byte groups.  The group that holds the slab's last edge may reach past the END OF
  // THE ARRAY (by at most 3 elements, only for the last slab of the last row when E % 4 != 0): owned graphs are
  // padded, but a view over the caller's arrays (b2g_graph_view_csr, graph::build over a user's csr_t) is not --
  // so the copy stops at the last whole group inside the array and `tail_fix` loads the rest with plain loads.
  const int n_edges = p.g.n_edges;
  auto issue = [&](const hub_slab_t& d, int buf) {  // one thread
    const int a0 = d.e0 & ~3;
    const int a1 = min((d.e0 + d.cnt + 3) & ~3, n_edges & ~3);
    const uint32_t bytes = a1 > a0 ? static_cast<uint32_t>(a1 - a0) * 4u : 0u;
    mbar_expect_tx(&s_bar[buf], use_vals ? 2 * bytes : bytes);  // 0 bytes: a plain arrival, the phase completes
    if (bytes) {
      bulk_g2s(&s_idx[buf][0], ci + a0, bytes, &s_bar[buf]);
      if (use_vals)
        bulk_g2s(&s_val[kWeights ? buf : 0][0], vals + a0, bytes, &s_bar[buf]);
    }
  };
  auto tail_fix = [&](const hub_slab_t& d, int buf) {  // all threads; uniform condition
    const int a0 = d.e0 & ~3;
    const int covered = n_edges & ~3;  // first element the bulk copy could not bring
    if (d.e0 + d.cnt > covered) {
      const int i = covered + static_cast<int>(threadIdx.x);
      if (threadIdx.x < 3 && i < d.e0 + d.cnt && i >= a0) {
        s_idx[buf][i - a0] = ci[i];
        if (use_vals)
          s_val[kWeights ? buf : 0][i - a0] = vals[i];
      }
      __syncthreads();
    }
  };
  if (warp == 0)
    fetch_range(0);
  __syncthreads();  // mbarriers initialised, first range visible
  if (tma && threadIdx.x == 0 && s_desc[0][0].cnt > 0)
    issue(s_desc[0][0], 0);
  unsigned phase_bits = 0;
  int ring = 0, i = 0, buf = 0;
  for (;;) {
    if (i == 0 && warp == 0)
      fetch_range(ring ^ 1);  // visible to everybody after this slab's closing barrier (range >= 2 of them follow)
    const hub_slab_t d = s_desc[ring][i];
    if (d.cnt <= 0)
      break;  // uniform: every thread reads the same descriptor
    if (tma && threadIdx.x == 0) {
      // the next slab's descriptor: in this range, or the first of the next one (fetched range - 1 barriers ago;
      // with a range of 1 it would not be visible yet)
      const hub_slab_t nd = (i + 1 < range) ? s_desc[ring][i + 1] : s_desc[ring ^ 1][0];
      if (nd.cnt > 0)
        issue(nd, buf ^ 1);
    }
    const int u = d.src;
    const int e0 = d.e0, cnt = d.cnt;
    const int a0 = e0 & ~3;
    if (tma) {
      mbar_wait(&s_bar[buf], (phase_bits >> buf) & 1u);
      phase_bits ^= 1u << buf;
      tail_fix(d, buf);
    }
    for (int i0 = 0; i0 < cnt; i0 += kThreads * kHB) {
      int e[kHB], nbv[kHB];
      float w[kHB];
      bool valid[kHB], keep[kHB];
      typename op_traits<Op>::token_t tok[kHB];
#pragma unroll
      for (int k = 0; k < kHB; ++k) {
        int j = i0 + k * kThreads + threadIdx.x;
        e[k] = e0 + j;
        valid[k] = j < cnt;
        nbv[k] = -1;
        w[k] = 1.0f;
        if (valid[k]) {
          if (tma) {
            nbv[k] = s_idx[buf][e[k] - a0];
            if (use_vals)
              w[k] = s_val[kWeights ? buf : 0][e[k] - a0];
          } else {
            nbv[k] = ld_stream(ci + e[k]);
            if (use_vals)
              w[k] = ld_stream(vals + e[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kHB; ++k)
        if (valid[k])
          tok[k] = op_prefetch(op, nbv[k]);
#pragma unroll
      for (int k = 0; k < kHB; ++k)
        keep[k] = valid[k] && op_commit(op, u, nbv[k], e[k], w[k], tok[k]);
      if (kOut != advance_output_t::none) {
#pragma unroll
        for (int k = 0; k < kHB; ++k)
          em.push(keep[k], kOut == advance_output_t::edges ? e[k] : op_emit(op, nbv[k]));
      }
    }
    __syncthreads();  // all reads of s_idx[buf] / s_desc[ring][i] retire before they are refilled
    buf ^= 1;
    if (++i == range) {
      i = 0;
      ring ^= 1;
    }
  }
  if (kOut != advance_output_t::none)
    em.flush();
}

/// One thread per frontier entry, serial walk ("thread_mapped", thread_mapped.hxx:58-81).
template <int kThreads, advance_input_t kIn, advance_output_t kOut, bool kDegSum, bool kWeights,
          typename Op>
__global__ void __launch_bounds__(kThreads)
advance_thread_mapped_kernel(advance_params_t p, Op op) {
  constexpr int kWarps = kThreads / 32;
  __shared__ int s_emit[kWarps][kEmitCap];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int* __restrict__ ro = p.g.row_offsets;
  const int* __restrict__ ci = p.g.column_indices;
  const float* __restrict__ vals = p.g.values;
  const int n = (kIn == advance_input_t::graph) ? p.g.n_vertices : *p.in_count;
  warp_emitter_t<kEmitCap, kDegSum> em;
  em.init(s_emit[warp], p.out, p.out_count, p.out_capacity, ro, p.ctrl);
  unsigned long long edges_seen = 0;
  for (;;) {
    int base = 0;
    if (lane == 0)
      base = atomicAdd(&p.ctrl->work, 32);
    base = __shfl_sync(kFull, base, 0);
    if (base >= n)
      break;
    int idx = base + lane;
    int v = -1;
    if (idx < n)
      v = (kIn == advance_input_t::graph) ? idx : p.in[idx];
    int start = 0, deg = 0;
    if (v >= 0) {
      start = ro[v];
      deg = ro[v + 1] - start;
    }
    edges_seen += static_cast<unsigned>(deg);
    int maxdeg = deg;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      maxdeg = max(maxdeg, __shfl_xor_sync(kFull, maxdeg, d));
    for (int k = 0; k < maxdeg; ++k) {
      bool keep = false;
      int nb = -1;
      int e = start + k;
      if (k < deg) {
        nb = ci[e];
        float w = (kWeights && vals) ? vals[e] : 1.0f;
        keep = op(v, nb, e, w);
      }
      if (kOut != advance_output_t::none)
        em.push(keep, kOut == advance_output_t::edges ? e : op_emit(op, nb));
    }
  }
  if (kOut != advance_output_t::none)
    em.flush();
  edges_seen = warp_sum(edges_seen);
  if (lane == 0 && edges_seen)
    atomicAdd(&p.ctrl->edges, edges_seen);
}

/// merge_path partition: tile t starts in row tile_rows[t] = max{ i : scanned[i] <= t*kTile }.
/// One thread per tile, so the log2(n) dependent loads of each search overlap across tiles.
template <int kTile>
__global__ void merge_path_partition_kernel(const int* __restrict__ scanned,
                                            const int* __restrict__ n_ptr,
                                            int n_fixed,
                                            int* __restrict__ tile_rows) {
  const int n = n_ptr ? *n_ptr : n_fixed;
  if (n == 0)
    return;
  const int total = scanned[n];
  const int ntiles = (total + kTile - 1) / kTile;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
    long long r = static_cast<long long>(t) * kTile;
    int row;
    if (r >= total) {
      row = n;
    } else {
      int lo = 0, hi = n;  // scanned[lo] <= r < scanned[hi]
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (scanned[mid] <= r)
          lo = mid;
        else
          hi = mid;
      }
      row = lo;
    }
    tile_rows[t] = row;
  }
}

/**
 * @brief Edge-balanced tiles over the degree scan of the frontier ("merge_path").
 * scanned[i] = exclusive prefix of deg(frontier[i]), scanned[n] = total (device resident).
 * A tile is kTile consecutive edge ranks.  The CTA stages the non-empty rows that overlap the
 * tile (first rank, CSR base, vertex) in shared memory; each warp then walks 32 ranks at a time:
 * one warp-uniform search gives the row window, lanes finish with a <= 5-step local search.
 * Column loads are coalesced within rows; every tile costs the same number of edges whatever
 * the degree skew.
 *
 * kRanked = the REFERENCE's output layout instead of the compact one (opt-in, launch_advance_ranked): the
 * output has one slot per edge rank -- out[rank] = the neighbour (or edge id) where the functor returned true,
 * -1 where it did not, *out_count = scanned[n] -- which is what merge_path.hxx:218-279 writes and what
 * callers that index the output by edge rank rely on.
 */
template <int kThreads, int kTile, advance_input_t kIn, advance_output_t kOut, bool kDegSum,
          bool kWeights, bool kRanked, typename Op>
__global__ void __launch_bounds__(kThreads)
advance_merge_path_kernel(advance_params_t p, const int* __restrict__ scanned, Op op) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kRows = kTile + 36;  // kTile edges overlap at most kTile non-empty rows (+ 33 sentinels)
  __shared__ int s_emit[kWarps][kEmitCap];
  constexpr bool kSrc = op_needs_source<Op>::value;
  static_assert(kTile < 65536 - 64, "row starts are kept as 16-bit offsets from the tile start");
  __shared__ unsigned short s_rank[kRows];  // first rank of each staged row, relative to the tile
  __shared__ int s_base[kRows];             // (CSR offset of the row's first edge) - (its first rank)
  __shared__ int s_vert[kSrc ? kRows : 1];
  __shared__ int s_wcount[kWarps];
  __shared__ int s_tile, s_row0, s_row1;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int* __restrict__ ro = p.g.row_offsets;
  const int* __restrict__ ci = p.g.column_indices;
  const float* __restrict__ vals = p.g.values;
  const int n = (kIn == advance_input_t::graph) ? p.g.n_vertices : *p.in_count;
  if (n == 0)
    return;
  const int total = scanned[n];
  const int ntiles = (total + kTile - 1) / kTile;
  warp_emitter_t<kEmitCap, kDegSum> em;
  em.init(s_emit[warp], p.out, p.out_count, p.out_capacity, ro, p.ctrl);

  // Dynamic tile tickets (CTAs that become resident late simply take fewer tiles).  The ticket of
  // the NEXT tile is drawn by thread 0 while the current tile is being walked, so a tile costs
  // three block barriers in total: [A] previous tile retired + ticket visible, [B] per-warp live
  // row counts visible, [C] staged rows + sentinels visible.
  // Thread 0 also fetches the row window of the ticketed tile, so by the time the CTA reaches
  // barrier [A] the only memory latency left in staging is one parallel round of loads
  // (scan values + row bases).
  auto draw_ticket = [&]() {
    int t = atomicAdd(&p.ctrl->work, 1);
    s_tile = t;
    if (t < ntiles) {
      s_row0 = p.tile_rows[t];
      s_row1 = p.tile_rows[t + 1];
    }
  };
  if (threadIdx.x == 0)
    draw_ticket();
  for (;;) {
    __syncthreads();  // [A]
    const int tile = s_tile;
    if (tile >= ntiles)
      break;
    const int r_begin = tile * kTile;
    const int r_end = min(total, r_begin + kTile);
    const int row0 = s_row0;
    const int row1 = min(n - 1, s_row1);  // last row that can overlap the tile
    int nrows = 0;  // identical in every thread
    for (int i0 = row0; i0 <= row1; i0 += kThreads) {
      int i = i0 + threadIdx.x;
      int sc = 0, rb = 0, vv = 0;
      bool live = false;
      if (i <= row1) {
        sc = scanned[i];
        int sc_next = scanned[i + 1];
        rb = p.row_base[i];  // independent of the scan loads: one round of memory latency
        if (kSrc)
          vv = (kIn == advance_input_t::graph) ? i : p.in[i];
        live = sc_next > sc && sc < r_end && sc_next > r_begin;
      }
      unsigned m = __ballot_sync(kFull, live);
      if (i0 != row0)
        __syncthreads();  // s_wcount of the previous batch has been consumed
      if (lane == 0)
        s_wcount[warp] = __popc(m);
      __syncthreads();  // [B]
      int off = nrows, batch = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        int c = s_wcount[w];
        if (w < warp)
          off += c;
        batch += c;
      }
      if (live) {
        int slot = off + __popc(m & lanemask_lt());
        s_rank[slot] = static_cast<unsigned short>(max(sc, r_begin) - r_begin);
        s_base[slot] = rb - sc;
        if (kSrc)
          s_vert[slot] = vv;
      }
      nrows += batch;
    }
    if (threadIdx.x < 33)
      s_rank[nrows + threadIdx.x] =
          static_cast<unsigned short>(r_end - r_begin);  // sentinels: no row starts past the tile
    __syncthreads();  // [C]
    if (threadIdx.x == 0)
      draw_ticket();  // next ticket + its row window, consumed after barrier [A]
    // Each warp owns a contiguous span of the tile and walks it 32 ranks at a time, so its row
    // cursor only moves forward: one binary search per span, then per chunk the row starts that
    // fall inside the 32 ranks are turned into a bit mask (one REDUX) and every lane derives its
    // row as cursor + popc(mask & lanes_le) -- no per-edge search.
    constexpr int kSpan = kTile / kWarps;
    const int w_begin = r_begin + warp * kSpan;
    const int w_end = min(r_end, w_begin + kSpan);
    if (w_begin < w_end) {
      int a = 0, b = nrows;  // last staged row with s_rank <= w_begin
      while (b - a > 1) {
        int mid = (a + b) >> 1;
        if (static_cast<int>(s_rank[mid]) + r_begin <= w_begin)
          a = mid;
        else
          b = mid;
      }
      for (int r0 = w_begin; r0 < w_end; r0 += 32 * kBatch) {
        int row[kBatch], e[kBatch], nb[kBatch], u[kBatch];
        float w[kBatch];
        bool valid[kBatch], keep[kBatch];
        typename op_traits<Op>::token_t tok[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          // invariant: s_rank[a] <= rk <= s_rank[a+1].  Row starts of rows a+1.. that fall inside
          // [rk, rk+32) become bits (strictly increasing starts => at most 32 of them).
          const int rk = r0 + 32 * k;
          int nxt = static_cast<int>(s_rank[min(a + 1 + lane, nrows + 32)]) + r_begin - rk;
          unsigned bit = (nxt >= 0 && nxt < 32) ? (1u << nxt) : 0u;
          unsigned starts = __reduce_or_sync(kFull, bit);
          row[k] = min(a + __popc(starts & (0xffffffffu >> (31 - lane))), nrows - 1);
          valid[k] = rk + lane < w_end;
          a += __popc(starts);
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          u[k] = kSrc ? s_vert[row[k]] : -1;
          e[k] = s_base[row[k]] + r0 + 32 * k + lane;  // s_base = CSR offset - first rank of the row
          nb[k] = valid[k] ? ld_stream(ci + e[k]) : -1;
          w[k] = (kWeights && vals && valid[k]) ? ld_stream(vals + e[k]) : 1.0f;
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
          if (valid[k])
            tok[k] = op_prefetch(op, nb[k]);
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
          keep[k] = valid[k] && op_commit(op, u[k], nb[k], e[k], w[k], tok[k]);
        if (kOut != advance_output_t::none) {
#pragma unroll
          for (int k = 0; k < kBatch; ++k) {
            const int item = kOut == advance_output_t::edges ? e[k] : op_emit(op, nb[k]);
            if constexpr (kRanked) {
              const int rank = r0 + 32 * k + lane;  // coalesced: consecutive lanes, consecutive slots
              if (valid[k] && rank < p.out_capacity)
                p.out[rank] = keep[k] ? item : -1;
            } else {
              em.push(keep[k], item);
            }
          }
        }
      }
    }
  }
  if (kOut != advance_output_t::none && !kRanked)
    em.flush();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&p.ctrl->edges, static_cast<unsigned long long>(total));
    if (kRanked && kOut != advance_output_t::none) {
      *p.out_count = min(total, p.out_capacity);
      if (total > p.out_capacity)
        p.ctrl->overflow = 1;
    }
  }
}

/**
 * @brief merge_path with WARP-PRIVATE spans -- no block barrier anywhere (the default for the fused BFS / SSSP
 * functors, see op_merge_path_kernel).
 *
 * Why: the ncu capture of advance_merge_path_kernel on the bench graph (profiles/r1_b_merge_path_v5_ncu.md)
 * shows a latency-bound kernel -- issue slots 43 % busy, L1 40 %, DRAM 12 % -- whose two largest stall
 * reasons are the scoreboard (10.9 warps per issue) and the three block barriers of every 2048-edge
 * tile (8.9 warps per issue): whenever one warp of a CTA waits for its probes, the other seven end up
 * waiting for it at the next barrier.  Here the unit of work is a SPAN of kSpan consecutive edge ranks
 * owned by one warp: the partition kernel gives the first frontier row of every span, the warp stages
 * the (at most kSpan) rows that overlap its span in its own slice of shared memory, then walks the span
 * with the same REDUX row-mask walk as the CTA kernel.  Warps never wait for each other; a warp draws
 * kTicket spans per atomic.
 * (Round 2 also measured copies of the visited map in shared memory / distributed shared memory in front of
 * the probes: 1.3x - 2x SLOWER -- a probe through DSMEM runs at 52 G/s against 362 G/s through L1 -- removed.)
 */
template <int kThreads, int kMinCtas, int kSpan, int kB, advance_input_t kIn, advance_output_t kOut, bool kDegSum,
          bool kWeights, typename Op>
__global__ void __launch_bounds__(kThreads, kMinCtas)
advance_warp_path_kernel(advance_params_t p, const int* __restrict__ scanned, Op op) {
  constexpr int kRows = kSpan + 36;  // kSpan ranks overlap at most kSpan non-empty rows (+ 33 sentinels)
  constexpr int kTicket = 8;         // spans per work-cursor atomic (lanes 0..kTicket hold their first rows)
  constexpr bool kSrc = op_needs_source<Op>::value;
  constexpr int kWarpInts = warp_path_ints<kSpan, kSrc>();
  static_assert(kSpan % 32 == 0 && kSpan < 65536 - 64 && kRows % 2 == 0, "span layout");
  unsigned char* smem_raw = dynamic_smem();
  // layout per warp: [emit | base | (vert) | rank (16 bit)]
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int* mine = reinterpret_cast<int*>(smem_raw) + warp * kWarpInts;
  int* s_base = mine + kEmitCap;  // (CSR offset of the row's first edge) - (its first rank)
  int* s_vert = s_base + kRows;   // only when the functor reads its source
  unsigned short* s_rank =        // first rank of each staged row, relative to the span start
      reinterpret_cast<unsigned short*>(s_base + kRows + (kSrc ? kRows : 0));
  const int* __restrict__ ro = p.g.row_offsets;
  const int* __restrict__ ci = p.g.column_indices;
  const float* __restrict__ vals = p.g.values;
  const int n = (kIn == advance_input_t::graph) ? p.g.n_vertices : *p.in_count;
  if (n == 0)
    return;
  const int total = scanned[n];
  const int nspans = (total + kSpan - 1) / kSpan;
  warp_emitter_t<kEmitCap, kDegSum> em;
  em.init(mine, p.out, p.out_count, p.out_capacity, ro, p.ctrl);

  for (;;) {
    int first = 0;
    if (lane == 0)
      first = atomicAdd(&p.ctrl->work, kTicket);
    first = __shfl_sync(kFull, first, 0);
    if (first >= nspans)
      break;
    const int last = min(nspans, first + kTicket);
    // p.tile_rows[k] = row holding rank k * kSpan (n past the end): entries 0 .. nspans exist
    const int my_row = (first + lane <= last) ? p.tile_rows[first + lane] : 0;
    for (int sp = first; sp < last; ++sp) {
      const int row0 = __shfl_sync(kFull, my_row, sp - first);
      const int row1 = min(n - 1, __shfl_sync(kFull, my_row, sp - first + 1));
      const int r_begin = sp * kSpan;
      const int r_end = min(total, r_begin + kSpan);
      // ---- stage the rows that overlap [r_begin, r_end) ------------------------------------
      int nrows = 0;  // warp-uniform
      for (int i0 = row0; i0 <= row1; i0 += 32) {
        const int i = i0 + lane;
        int sc = 0, sc_next = 0, rb = 0, vv = 0;
        bool live = false;
        if (i <= row1) {
          sc = scanned[i];
          sc_next = scanned[i + 1];
          rb = p.row_base[i];
          if (kSrc)
            vv = (kIn == advance_input_t::graph) ? i : p.in[i];
          live = sc_next > sc && sc < r_end && sc_next > r_begin;
        }
        const unsigned m = __ballot_sync(kFull, live);
        if (live) {
          const int slot = nrows + __popc(m & lanemask_lt());
          s_rank[slot] = static_cast<unsigned short>(max(sc, r_begin) - r_begin);
          s_base[slot] = rb - sc;
          if (kSrc)
            s_vert[slot] = vv;
        }
        nrows += __popc(m);
      }
      s_rank[nrows + lane] = static_cast<unsigned short>(r_end - r_begin);  // sentinels
      if (lane == 0)
        s_rank[nrows + 32] = static_cast<unsigned short>(r_end - r_begin);
      __syncwarp();
      // ---- walk: slot 0 is the row that holds rank r_begin, the cursor only moves forward ----
      int a = 0;
      for (int r0 = r_begin; r0 < r_end; r0 += 32 * kB) {
        int row[kB], e[kB], nb[kB], u[kB];
        float w[kB];
        bool valid[kB], keep[kB];
        typename op_traits<Op>::token_t tok[kB];
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          const int rk = r0 + 32 * k;
          int nxt = static_cast<int>(s_rank[min(a + 1 + lane, nrows + 32)]) + r_begin - rk;
          unsigned bit = (nxt >= 0 && nxt < 32) ? (1u << nxt) : 0u;
          unsigned starts = __reduce_or_sync(kFull, bit);
          row[k] = min(a + __popc(starts & (0xffffffffu >> (31 - lane))), nrows - 1);
          valid[k] = rk + lane < r_end;
          a += __popc(starts);
        }
#pragma unroll
        for (int k = 0; k < kB; ++k) {
          u[k] = kSrc ? s_vert[row[k]] : -1;
          e[k] = s_base[row[k]] + r0 + 32 * k + lane;
          nb[k] = valid[k] ? ld_stream(ci + e[k]) : -1;
          w[k] = (kWeights && vals && valid[k]) ? ld_stream(vals + e[k]) : 1.0f;
        }
#pragma unroll
        for (int k = 0; k < kB; ++k)
          if (valid[k])
            tok[k] = op_prefetch(op, nb[k]);
#pragma unroll
        for (int k = 0; k < kB; ++k)
          keep[k] = valid[k] && op_commit(op, u[k], nb[k], e[k], w[k], tok[k]);
        if (kOut != advance_output_t::none) {
#pragma unroll
          for (int k = 0; k < kB; ++k)
            em.push(keep[k], kOut == advance_output_t::edges ? e[k] : op_emit(op, nb[k]));
        }
      }
      __syncwarp();  // every lane is done with this span's rows before the next span restages them
    }
  }
  if (kOut != advance_output_t::none)
    em.flush();
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&p.ctrl->edges, static_cast<unsigned long long>(total));
}

/// Report of advance_tail_kernel (written to pinned host memory by the kernel).
struct tail_report_t {
  int levels;    // levels executed by this launch
  int count;     // size of the frontier it stopped at
  int cur;       // which of the two queues holds that frontier
  int pad;
  unsigned long long deg_sum;  // out-degree sum of that frontier
  unsigned long long edges[16];
  int frontier[16];
  volatile int seq;  // written last, after a system fence (the host polls it)
};

/**
 * @brief Tail of a traversal in ONE launch: while the frontier stays tiny (its out-degree sum below
 * `edge_budget`), a single CTA runs level after level -- warp per frontier vertex, ballot-compacted
 * appends through a shared counter into the other global queue, one block barrier per level --
 * instead of paying a launch + host round trip per level (RMAT traversals end with 2-4 levels of a
 * few hundred edges each).  `make_op(level)` builds the level's edge functor.
 */
template <int kThreads, bool kWeights, typename OpMaker>
__global__ void __launch_bounds__(kThreads, 1)  // one CTA per launch: take the registers
advance_tail_kernel(csr_view_t g, int* q0, int* q1, int* counts, int cur, int first_level,
                    int max_levels, unsigned long long edge_budget, OpMaker make_op,
                    tail_report_t* rep, int seq) {
  __shared__ int s_cnt;
  __shared__ unsigned long long s_deg, s_edges;
  const int lane = lane_id();
  const int* __restrict__ ro = g.row_offsets;
  const int* __restrict__ ci = g.column_indices;
  const float* __restrict__ vals = g.values;
  int* q[2] = {q0, q1};
  int n = counts[cur];
  int level = first_level, done = 0;
  unsigned long long deg_sum = 0;
  for (;;) {
    if (threadIdx.x == 0) {
      s_cnt = 0;
      s_deg = 0;
      s_edges = 0;
    }
    __syncthreads();
    auto op = make_op(level);
    const int* in = q[cur];
    int* out = q[cur ^ 1];
    unsigned long long my_deg = 0, my_edges = 0;
    // one lane per frontier row (late levels: thousands of rows of a few edges each); rows of 32+
    // edges are walked by the whole warp, one coalesced chunk at a time
    for (int base = 0; base < n; base += kThreads) {
      const int i = base + static_cast<int>(threadIdx.x);
      int v = -1, s = 0, d = 0;
      if (i < n) {
        v = in[i];
        if (v >= 0) {
          s = ro[v];
          d = ro[v + 1] - s;
        }
      }
      my_edges += static_cast<unsigned>(d);
      unsigned big = __ballot_sync(kFull, d >= 32);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bv = __shfl_sync(kFull, v, src), bs = __shfl_sync(kFull, s, src),
                  bd = __shfl_sync(kFull, d, src);
        for (int off = 0; off < bd; off += 32) {
          const int e = bs + off + lane;
          bool keep = false;
          int nb = -1;
          if (off + lane < bd) {
            nb = ci[e];
            float w = (kWeights && vals) ? vals[e] : 1.0f;
            keep = op(bv, nb, e, w);
          }
          const unsigned m = __ballot_sync(kFull, keep);
          if (m) {
            int at = 0;
            if (lane == 0)
              at = atomicAdd(&s_cnt, __popc(m));
            at = __shfl_sync(kFull, at, 0);
            if (keep) {
              int x = op_emit(op, nb);
              out[at + __popc(m & lanemask_lt())] = x;
              my_deg += static_cast<unsigned>(ro[x + 1] - ro[x]);
            }
          }
        }
      }
      if (d < 32) {
        for (int k = 0; k < d; ++k) {
          const int e = s + k;
          const int nb = ci[e];
          float w = (kWeights && vals) ? vals[e] : 1.0f;
          if (op(v, nb, e, w)) {
            int x = op_emit(op, nb);
            out[atomicAdd(&s_cnt, 1)] = x;
            my_deg += static_cast<unsigned>(ro[x + 1] - ro[x]);
          }
        }
      }
    }
    my_deg = warp_sum(my_deg);
    my_edges = warp_sum(my_edges);
    if (lane == 0) {
      if (my_deg)
        atomicAdd(&s_deg, my_deg);
      if (my_edges)
        atomicAdd(&s_edges, my_edges);
    }
    __syncthreads();
    if (threadIdx.x == 0 && done < 16) {
      rep->edges[done] = s_edges;
      rep->frontier[done] = n;
    }
    n = s_cnt;
    deg_sum = s_deg;
    cur ^= 1;
    ++level;
    ++done;
    if (n == 0 || deg_sum >= edge_budget || done >= max_levels)
      break;
    __syncthreads();  // everyone has read s_cnt / s_deg before they are cleared
  }
  if (threadIdx.x == 0) {
    counts[cur] = n;
    rep->levels = done;
    rep->count = n;
    rep->cur = cur;
    rep->deg_sum = deg_sum;
    __threadfence_system();
    rep->seq = seq;
  }
}

// ---------------------------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------------------------
enum class lb_t { thread_mapped, block_mapped, merge_path };

struct advance_launch_t {
  lb_t lb = lb_t::block_mapped;
  int hub_threshold = 4096;  // rows >= this go to the grid (TMA slab) bin; block_mapped only
  int ctas_per_sm = 8;
  /// frontiers whose out-degree sum is below this take the single-kernel path (no scan, no hub
  /// pass): fixed per-level cost matters more than balance there.
  long long small_frontier_edges = 1 << 12;
  /// fused enactors only: merge_path requests below this many frontier edges use the scan-free
  /// CTA walk (block_mapped kernels) -- the scan + partition launches cost more than they save.
  long long mid_frontier_edges = 1 << 20;
  /// block_mapped: average out-degree of the frontier if the caller knows it (0 = unknown).
  double avg_degree = 0.0;
  /// merge_path kernel: < 0 = the functor's own choice (op_merge_path_kernel), 0 = CTA tiles of 2048 edges
  /// (advance_merge_path_kernel), 1 = warp-private spans, 256-thread CTAs, 4 chunks in flight, registers capped for
  /// 6 CTAs per SM, 4 = the same with 8 chunks in flight, capped for 4 CTAs per SM (advance_warp_path_kernel).
  /// Functors without a choice of their own always run 0.
  int variant = -1;
  /// Upper bound on the out-degree sum of the input frontier when the caller knows one above the graph's edge
  /// count (a frontier that holds vertices several times); sizes the hub slab table.  0 = the edge count.
  long long edges_upper_bound = 0;
};


/// Allocate, once, everything launch_advance may need for frontiers of up to `n_upper_bound` rows of
/// `g`, so that a later run performs no cudaMalloc / cudaFree (both synchronise the device, which a
/// run that overlaps with other streams' spinning barrier kernels must not do -- bfs_p2p.cuh).
inline void reserve_advance_workspace(workspace_t& ws, const csr_view_t& g, int n_upper_bound) {
  ws.scanned.ensure(2 * static_cast<size_t>(n_upper_bound) + 4);
  ws.tile_rows.ensure((static_cast<size_t>(1) << 31) / 2048 + 4);
  ws.hubs.ensure(static_cast<size_t>(g.n_edges / 256 + 1024));
  ws.hub_slabs.ensure(2 * (static_cast<size_t>(g.n_edges) / 2048 + static_cast<size_t>(g.n_edges / 256 + 1024) + 64));
  const int max_tiles = (n_upper_bound + 256 * 8 - 1) / (256 * 8) + 1;
  const size_t had = ws.tile_state.cap;
  ws.tile_state.ensure(max_tiles);
  if (ws.tile_state.cap != had)
    B2G_CHECK(cudaMemsetAsync(ws.tile_state.ptr, 0, ws.tile_state.cap * 8, ws.stream));
}

/// Degree scan of the frontier for merge_path (replaces helpers.hxx:41-111): scanned[0..n],
/// scanned[n] = total, all written by the one look-back scan kernel.
inline const int* frontier_degree_scan(workspace_t& ws,
                                       const csr_view_t& g,
                                       const int* in,
                                       const int* in_count,
                                       int n_upper_bound,
                                       const int** row_base_out) {
  int* scanned = ws.scanned.ensure(2 * static_cast<size_t>(n_upper_bound) + 4);
  int* row_base = scanned + n_upper_bound + 2;  // CSR offset of every frontier row
  const int* ro = g.row_offsets;
  auto value = [=] __device__(int i) -> int {
    int v = in[i];
    return v >= 0 ? ro[v + 1] - ro[v] : 0;
  };
  auto emit = [=] __device__(int i, int excl, int) {
    scanned[i] = excl;
    int v = in[i];
    row_base[i] = v >= 0 ? ro[v] : 0;
  };
  lookback_scan(ws, in_count, 0, n_upper_bound, value, emit, nullptr, scanned);
  *row_base_out = row_base;
  return scanned;
}

/// Sum of the out-degrees of a vertex frontier (invalid entries count 0), for exact output sizing.
static __global__ void frontier_degree_total_kernel(csr_view_t g, const int* __restrict__ in,
                                                    const int* __restrict__ in_count, ctrl_t* ctrl) {
  const int n = *in_count;
  unsigned long long sum = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = in[i];
    if (v >= 0 && v < g.n_vertices)
      sum += static_cast<unsigned>(g.row_offsets[v + 1] - g.row_offsets[v]);
  }
  sum = warp_sum(sum);
  if (lane_id() == 0 && sum)
    atomicAdd(&ctrl->deg_sum, sum);
}
static __global__ void publish_degree_total_kernel(const ctrl_t* ctrl, workspace_t::host_value_t* h, int seq) {
  h->value = ctrl->deg_sum;
  __threadfence_system();
  h->seq = seq;
}

/**
 * @brief The exact number of output slots an advance over `in` can produce: the out-degree sum of the
 * frontier, read back through pinned memory (one cheap poll, no stream synchronise).  This is what the
 * reference computes -- and blocks on -- before EVERY advance (advance/helpers.hxx:127-161,
 * block_mapped.hxx:205-217); here it is only needed for frontiers that may hold the same vertex more than
 * once, whose expansion is not bounded by the number of edges of the graph.
 */
inline unsigned long long frontier_degree_total(workspace_t& ws, const csr_view_t& g, const int* in,
                                                const int* in_count, int in_upper_bound) {
  const int sms = device_info_t::get().sm_count;
  ctrl_t* c = ws.next_ctrl();
  auto* h = ws.host_value();
  int blocks = (in_upper_bound + 255) / 256;
  blocks = blocks < 1 ? 1 : (blocks > sms * 8 ? sms * 8 : blocks);
  frontier_degree_total_kernel<<<blocks, 256, 0, ws.stream>>>(g, in, in_count, c);
  publish_degree_total_kernel<<<1, 1, 0, ws.stream>>>(c, h, ++ws.value_seq);
  ws.launches += 2;
  B2G_CHECK(cudaGetLastError());
  wait_for_sequence(&h->seq, ws.value_seq, ws.stream);
  return h->value;
}

inline bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

/// merge_path proper: partition into kTile-edge tiles, then the CTA kernel.
/// Ranks are int32 (as in the reference, merge_path.hxx:325-327), so 2^31/kTile tiles bound every
/// possible frontier, duplicates included.
template <int kTile, advance_output_t kOut, bool kDegSum, bool kWeights, bool kRanked = false, typename Op>
inline void launch_merge_path_tiles(workspace_t& ws, advance_params_t& p, const int* scanned,
                                    bool graph_in, int grid, Op op) {
  constexpr int kThreads = 256;
  const int sms = device_info_t::get().sm_count;
  int* tile_rows = ws.tile_rows.ensure((static_cast<size_t>(1) << 31) / kTile + 4);
  merge_path_partition_kernel<kTile><<<sms, 256, 0, ws.stream>>>(scanned, p.in_count, p.g.n_vertices,
                                                                 tile_rows);
  p.tile_rows = tile_rows;
  p.ctrl = ws.next_ctrl();
  if (graph_in)
    advance_merge_path_kernel<kThreads, kTile, advance_input_t::graph, kOut, kDegSum, kWeights, kRanked>
        <<<grid, kThreads, 0, ws.stream>>>(p, scanned, op);
  else
    advance_merge_path_kernel<kThreads, kTile, advance_input_t::vertices, kOut, kDegSum, kWeights, kRanked>
        <<<grid, kThreads, 0, ws.stream>>>(p, scanned, op);
}

/// merge_path with warp-private spans: span partition + advance_warp_path_kernel (variant 1 or 4).
template <advance_output_t kOut, bool kDegSum, bool kWeights, typename Op>
inline void launch_warp_path(workspace_t& ws, advance_params_t& p, const int* scanned, bool graph_in, int variant,
                             int ctas_per_sm, Op op) {
  constexpr int kSpan = 256;
  constexpr bool kSrc = op_needs_source<Op>::value;
  constexpr int kWarpBytes = warp_path_ints<kSpan, kSrc>() * 4;
  const int sms = device_info_t::get().sm_count;
  int* span_rows = ws.tile_rows.ensure((static_cast<size_t>(1) << 31) / kSpan + 4);
  // 8x the tiles of the CTA kernel: a full wave of searchers instead of one CTA per SM
  merge_path_partition_kernel<kSpan><<<sms * 8, 256, 0, ws.stream>>>(scanned, p.in_count,
                                                                     p.g.n_vertices, span_rows);
  p.tile_rows = span_rows;
  p.ctrl = ws.next_ctrl();
  constexpr auto kGraph = advance_input_t::graph;
  constexpr auto kVerts = advance_input_t::vertices;
  constexpr int kThreads = 256;
  const int smem = (kThreads / 32) * kWarpBytes;
  const int grid = sms * ctas_per_sm;
  if (variant == 4) {
    if (graph_in)
      advance_warp_path_kernel<kThreads, 4, kSpan, 8, kGraph, kOut, kDegSum, kWeights>
          <<<grid, kThreads, smem, ws.stream>>>(p, scanned, op);
    else
      advance_warp_path_kernel<kThreads, 4, kSpan, 8, kVerts, kOut, kDegSum, kWeights>
          <<<grid, kThreads, smem, ws.stream>>>(p, scanned, op);
  } else {
    if (graph_in)
      advance_warp_path_kernel<kThreads, 6, kSpan, kBatch, kGraph, kOut, kDegSum, kWeights>
          <<<grid, kThreads, smem, ws.stream>>>(p, scanned, op);
    else
      advance_warp_path_kernel<kThreads, 6, kSpan, kBatch, kVerts, kOut, kDegSum, kWeights>
          <<<grid, kThreads, smem, ws.stream>>>(p, scanned, op);
  }
}

/**
 * @brief Launch one advance.  `in == nullptr` means the whole graph is the input frontier.
 * `in_upper_bound` bounds *in_count (buffer sizing only).  `ctrl_out` (optional) returns the
 * control block holding deg_sum / edges / overflow for this launch.
 */
template <advance_output_t kOut, bool kDegSum, bool kWeights, typename Op>
inline void launch_advance(workspace_t& ws,
                           const csr_view_t& g,
                           const int* in,
                           const int* in_count,
                           int in_upper_bound,
                           int* out,
                           int* out_count,
                           int out_capacity,
                           Op op,
                           const advance_launch_t& cfg,
                           ctrl_t** ctrl_out = nullptr) {
  constexpr int kThreads = 256;
  constexpr int kTile = 2048;
  const int sms = device_info_t::get().sm_count;
  advance_params_t p;
  p.g = g;
  p.in = in;
  p.in_count = in_count;
  p.out = out;
  p.out_count = out_count;
  p.out_capacity = out_capacity;
  const int grid = sms * cfg.ctas_per_sm;
  const bool graph_in = (in == nullptr);
  if (cfg.lb == lb_t::thread_mapped) {
    p.ctrl = ws.next_ctrl();
    if (graph_in)
      advance_thread_mapped_kernel<kThreads, advance_input_t::graph, kOut, kDegSum, kWeights>
          <<<grid, kThreads, 0, ws.stream>>>(p, op);
    else
      advance_thread_mapped_kernel<kThreads, advance_input_t::vertices, kOut, kDegSum, kWeights>
          <<<grid, kThreads, 0, ws.stream>>>(p, op);
  } else if (cfg.lb == lb_t::merge_path) {
    // the CSR offsets ARE the degree scan when the whole graph is the frontier
    const int* row_base = g.row_offsets;  // whole graph: row i starts at row_offsets[i]
    const int* scanned = graph_in ? g.row_offsets
                                  : frontier_degree_scan(ws, g, in, in_count, in_upper_bound, &row_base);
    p.row_base = row_base;
    bool launched = false;
    if constexpr (op_merge_path_kernel<Op>::value != 0) {
      const int variant = cfg.variant < 0 ? op_merge_path_kernel<Op>::value : cfg.variant;
      if (variant == 1 || variant == 4) {
        launch_warp_path<kOut, kDegSum, kWeights>(ws, p, scanned, graph_in, variant, cfg.ctas_per_sm, op);
        launched = true;
      }
    }
    if (!launched)
      launch_merge_path_tiles<kTile, kOut, kDegSum, kWeights>(ws, p, scanned, graph_in, grid, op);
  } else {
    p.ctrl = ws.next_ctrl();
    p.hub_threshold = cfg.hub_threshold < 32 ? 32 : cfg.hub_threshold;
    p.hub_capacity = cfg.hub_threshold < (1 << 30) ? g.n_edges / 256 + 1024 : 16;
    p.hubs = ws.hubs.ensure(static_cast<size_t>(p.hub_capacity));
    // every deferred row has >= 1 slab and at most one partial one: E / 2048 + rows bounds the table
    const size_t edge_bound = cfg.edges_upper_bound > g.n_edges ? static_cast<size_t>(cfg.edges_upper_bound)
                                                                : static_cast<size_t>(g.n_edges);
    p.hub_slab_capacity = static_cast<int>(edge_bound / 2048 + static_cast<size_t>(p.hub_capacity) + 64);
    p.hub_slabs = reinterpret_cast<hub_slab_t*>(ws.hub_slabs.ensure(2 * static_cast<size_t>(p.hub_slab_capacity)));
    p.tma_ok = aligned16(g.column_indices) && (!kWeights || !g.values || aligned16(g.values));
    if (cfg.avg_degree > 0.0) {
      int want = static_cast<int>(16384.0 / cfg.avg_degree);
      int ept = 8;
      while (ept * 2 <= want && ept < 256)
        ept *= 2;
      p.entries_per_ticket = ept;
    }
    if (graph_in)
      advance_binned_kernel<kThreads, advance_input_t::graph, kOut, kDegSum, kWeights>
          <<<grid, kThreads, 0, ws.stream>>>(p, op);
    else
      advance_binned_kernel<kThreads, advance_input_t::vertices, kOut, kDegSum, kWeights>
          <<<grid, kThreads, 0, ws.stream>>>(p, op);
    if (cfg.hub_threshold < (1 << 30)) {
      advance_hub_table_kernel<2048><<<sms * 2, 256, 0, ws.stream>>>(p);
      advance_hub_kernel<kThreads, 2048, kOut, kDegSum, kWeights>
          <<<sms * (kWeights ? 4 : 5), kThreads, 0, ws.stream>>>(p, op);
      ws.launches += 1;
    } else {
      ws.launches -= 1;
    }
  }
  ws.launches += (cfg.lb == lb_t::thread_mapped) ? 1 : 2;
  if (ctrl_out)
    *ctrl_out = p.ctrl;
  B2G_CHECK(cudaGetLastError());
}

/**
 * @brief One advance with the REFERENCE's output layout (opt-in; `standard_context_t::reference_advance_output`):
 * slot r of the output belongs to edge rank r of the input frontier's expansion and holds the neighbour (edge id
 * for an edge output) or -1, `*out_count` = the frontier's out-degree sum -- merge_path.hxx:218-279 /
 * block_mapped.hxx:150-176 / thread_mapped.hxx:58-81 all produce a frontier of that shape (block_mapped with a
 * CTA-order permutation of the rows).  Every load balancer takes the tile kernel here: the layout IS the
 * merge-path rank space.  `out_capacity` must hold the degree sum (the caller sizes it, advance.hxx).
 */
template <advance_output_t kOut, bool kWeights, typename Op>
inline void launch_advance_ranked(workspace_t& ws, const csr_view_t& g, const int* in, const int* in_count,
                                  int in_upper_bound, int* out, int* out_count, int out_capacity, Op op,
                                  const advance_launch_t& cfg, ctrl_t** ctrl_out = nullptr) {
  static_assert(kOut != advance_output_t::none, "a ranked advance has an output frontier");
  constexpr int kTile = 2048;
  const int sms = device_info_t::get().sm_count;
  advance_params_t p;
  p.g = g;
  p.in = in;
  p.in_count = in_count;
  p.out = out;
  p.out_count = out_count;
  p.out_capacity = out_capacity;
  const bool graph_in = (in == nullptr);
  const int* row_base = g.row_offsets;
  const int* scanned = graph_in ? g.row_offsets
                                : frontier_degree_scan(ws, g, in, in_count, in_upper_bound, &row_base);
  p.row_base = row_base;
  launch_merge_path_tiles<kTile, kOut, false, kWeights, true>(ws, p, scanned, graph_in, sms * cfg.ctas_per_sm, op);
  ws.launches += 2;
  if (ctrl_out)
    *ctrl_out = p.ctrl;
  B2G_CHECK(cudaGetLastError());
}

}  // namespace b200
}  // namespace gunrock
