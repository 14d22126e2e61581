/**
 * @file bfs_partitioned.cuh
 * @brief Per-rank kernels of the multi-GPU BFS: 1-D vertex partition, one process per GPU, the
 * per-level remote-frontier exchange done by the host side with NCCL (torch.distributed
 * all_to_all_single / all_gather / all_reduce over NVLink) -- gunrock_b200/multi_gpu.py.
 *
 * The reference has no multi-GPU execution at all (advance/filter throw for context.size() != 1,
 * SURVEY.md F6); this is new design following SURVEY.md section 8e.
 *
 * Partition: CYCLIC.  Global vertex v lives on rank v % P at local row v / P.  That is the
 * contiguous-block partition of the fixed relabelling pi(v) = (v % P) * ceil(V/P) + v / P, and it
 * spreads RMAT's low-id hubs evenly without a permutation array.  A rank stores the CSR rows of its
 * vertices with GLOBAL column ids (and, for pull, the CSC rows -- the same arrays for a symmetric
 * graph), its slice of `distances` and of the visited bitmap.
 *
 * Top-down level = the ordinary advance (any load balancer) with `part_claim_op`:
 *   local neighbour  -> test-and-set in the local visited map, label, emit its LOCAL row id;
 *   remote neighbour -> first time this rank sees it (a V-bit "sent" map, 8 MiB at scale 26):
 *                       append it to the owner's send buffer (warp-aggregated per owner).
 * After the exchange, `part_claim_received_kernel` claims the received ids on their owner.
 * Bottom-up level = all-gather of the frontier bitmap (V/8 bytes total), then a purely local
 * sweep (`part_bottom_up_kernel`) -- no all-to-all.
 */
#pragma once

#include <gunrock/b200/advance.cuh>
#include <gunrock/b200/bfs.cuh>
#include <gunrock/b200/sssp.cuh>

namespace gunrock {
namespace b200 {

struct partition_t {
  int nparts = 1;
  int part = 0;
  int shift = 0;        // log2(nparts) when nparts is a power of two, else -1
  int n_global = 0;     // global vertex count
  int n_local = 0;      // rows owned by this rank
  __host__ __device__ __forceinline__ int owner(int v) const {
    return shift >= 0 ? (v & (nparts - 1)) : (v % nparts);
  }
  __host__ __device__ __forceinline__ int local(int v) const {
    return shift >= 0 ? (v >> shift) : (v / nparts);
  }
  __host__ __device__ __forceinline__ int global(int l) const { return l * nparts + part; }
  /// rows owned by rank r
  __host__ __device__ __forceinline__ int rows_of(int r) const {
    return (n_global - r + nparts - 1) / nparts;
  }
  static partition_t make(int n_global, int nparts, int part) {
    partition_t p;
    p.nparts = nparts;
    p.part = part;
    p.n_global = n_global;
    p.shift = -1;
    for (int s = 0; s < 31; ++s)
      if ((1 << s) == nparts)
        p.shift = s;
    p.n_local = p.rows_of(part);
    return p;
  }
};

/// Top-down edge functor of the partitioned BFS (two-phase protocol, see advance.cuh).
struct part_claim_op {
  static constexpr bool kNeedsSource = false;
  partition_t pt;
  unsigned* visited;   // local rows
  unsigned* sent;      // global ids already forwarded by this rank
  int* dist;           // local rows
  int next_level;
  int* send_buf;       // nparts x send_cap
  int* send_count;     // nparts
  int send_cap;
  int* overflow;

  __device__ __forceinline__ unsigned prefetch(int dst) const {
    return pt.owner(dst) == pt.part ? ld_cached(visited + (pt.local(dst) >> 5))
                                    : ld_cached(sent + (dst >> 5));
  }
  __device__ __forceinline__ bool commit(int, int dst, int, float, unsigned word) const {
    const int own = pt.owner(dst);
    if (own == pt.part) {
      const int l = pt.local(dst);
      const unsigned bit = 1u << (l & 31);
      if (word & bit)
        return false;
      if (atomicOr(visited + (l >> 5), bit) & bit)
        return false;
      dist[l] = next_level;
      return true;
    }
    const unsigned bit = 1u << (dst & 31);
    if (word & bit)
      return false;
    if (atomicOr(sent + (dst >> 5), bit) & bit)
      return false;
    // warp-aggregated append to the owner's send buffer
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, own);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (lane_id() == leader)
      base = atomicAdd(send_count + own, __popc(grp));
    base = __shfl_sync(grp, base, leader);
    const int slot = base + __popc(grp & lanemask_lt());
    if (slot < send_cap)
      send_buf[static_cast<size_t>(own) * send_cap + slot] = dst;
    else
      *overflow = 1;
    return false;
  }
  __device__ __forceinline__ bool operator()(int s, int d, int e, float w) const {
    return commit(s, d, e, w, prefetch(d));
  }
  /// the local frontier holds LOCAL row ids
  __device__ __forceinline__ int emit_as(int dst) const { return pt.local(dst); }
};

/// Claim ids received from peers (global ids owned by this rank) into the next local frontier.
static __global__ void part_claim_received_kernel(partition_t pt, const int* __restrict__ recv,
                                                  int n_recv, unsigned* visited, int* dist,
                                                  int next_level, const int* __restrict__ ro,
                                                  int* out, int* out_count,
                                                  unsigned long long* deg_sum) {
  const int lane = lane_id();
  unsigned long long ds = 0;
  for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n_recv;
       i0 += gridDim.x * blockDim.x) {
    int i = i0 + lane;
    bool won = false;
    int l = 0;
    if (i < n_recv) {
      l = pt.local(recv[i]);
      won = bitmap_test_and_set(visited, l);
      if (won) {
        dist[l] = next_level;
        ds += static_cast<unsigned>(ro[l + 1] - ro[l]);
      }
    }
    unsigned m = __ballot_sync(kFull, won);
    if (m) {
      int base = 0;
      if (lane == 0)
        base = atomicAdd(out_count, __popc(m));
      base = __shfl_sync(kFull, base, 0);
      if (won)
        out[base + __popc(m & lanemask_lt())] = l;
    }
  }
  ds = warp_sum(ds);
  if (lane == 0 && ds)
    atomicAdd(deg_sum, ds);
}

/// local queue (local row ids) -> local frontier bitmap
static __global__ void part_queue_to_bitmap_kernel(const int* __restrict__ q,
                                                   const int* __restrict__ count, unsigned* bm) {
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int l = q[i];
    atomicOr(bm + (l >> 5), 1u << (l & 31));
  }
}

/// "is global vertex u in the current frontier" against the all-gathered frontier bitmap (rank-major: words
/// [r * words_per_rank, (r + 1) * words_per_rank) hold rank r's local rows) -- the FrontierTest of the pull
/// kernels K1 / K2 (bfs.cuh) on a partitioned graph, whose column indices are global ids.
struct part_frontier_t {
  partition_t pt;
  const unsigned* all;
  int words_per_rank;
  __device__ __forceinline__ bool operator()(int u) const {
    const int l = pt.local(u);
    return (__ldg(all + pt.owner(u) * words_per_rank + (l >> 5)) >> (l & 31)) & 1u;
  }
};

/// Where the bottom-up sweep puts the words of the next frontier: the local bitmap (NCCL variant; the
/// peer-memory variant in bfs_p2p.cuh stores them into every rank's copy instead).
struct local_word_sink_t {
  unsigned* next;
  __device__ __forceinline__ void zero(int wi) const { next[wi] = 0; }  // lanes hold different words
  __device__ __forceinline__ void word(int wi, unsigned v) const {      // warp-uniform word
    if (lane_id() == 0)
      next[wi] = v;
  }
};

/**
 * @brief Bottom-up sweep over the rows this rank owns.  `frontier_all` is the all-gathered frontier
 * bitmap, laid out rank-major: words [r * words_per_rank, (r+1) * words_per_rank) hold rank r's
 * local rows.  Same structure as bfs_bottom_up_kernel (one warp per visited word).
 */
template <int kThreads, int kSerial, typename Sink>
__global__ void __launch_bounds__(kThreads)
part_bottom_up_kernel(partition_t pt, csr_view_t in, int words_per_rank,
                      unsigned* __restrict__ visited, const unsigned* __restrict__ frontier_all,
                      Sink sink, int* dist, int next_level, ctrl_t* ctrl, int* next_count) {
  const int lane = lane_id();
  const int words = (pt.n_local + 31) / 32;
  const int warps = (gridDim.x * kThreads) >> 5;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int* __restrict__ ro = in.row_offsets;
  const int* __restrict__ ci = in.column_indices;
  auto in_frontier = [&](int u) -> bool {
    int l = pt.local(u);
    return (__ldg(frontier_all + pt.owner(u) * words_per_rank + (l >> 5)) >> (l & 31)) & 1u;
  };
  unsigned long long scanned = 0, found_deg = 0;
  int found_cnt = 0;
  // 32 words per warp pass: one coalesced load decides which of them still hold unvisited
  // vertices (late levels: almost none), so the sweep is bandwidth- not latency-bound.
  for (int w0 = gw * 32; w0 < words; w0 += warps * 32) {
    const int my_wi = w0 + lane;
    const unsigned my_vis = my_wi < words ? visited[my_wi] : 0xffffffffu;
    if (my_wi < words && my_vis == 0xffffffffu)
      sink.zero(my_wi);
    unsigned todo = __ballot_sync(kFull, my_vis != 0xffffffffu);
    while (todo) {
    const int src_lane = __ffs(todo) - 1;
    todo &= todo - 1;
    const int wi = w0 + src_lane;
    const unsigned vis = __shfl_sync(kFull, my_vis, src_lane);
    const int v = (wi << 5) + lane;
    bool searching = v < pt.n_local && !((vis >> lane) & 1u);
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
        if (in_frontier(u)) {
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
          mine = in_frontier(ci[idx]);
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
    sink.word(wi, fm);  // warp-uniform arguments, every lane calls
    if (lane == 0 && fm)
      visited[wi] = vis | fm;
    found_cnt += found ? 1 : 0;
    }
  }
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

static __global__ void part_reset_kernel(partition_t pt, int source, int* dist, unsigned* visited,
                                         unsigned* sent, int sent_words, int* q0, int* counts,
                                         const unsigned* __restrict__ premark = nullptr) {
  const int lwords = (pt.n_local + 31) / 32;
  const int n = max(pt.n_local, sent_words);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < pt.n_local)
      dist[i] = 0x7fffffff;
    if (i < lwords)
      visited[i] = premark ? premark[i] : 0u;
    if (i < sent_words)
      sent[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[0] = counts[1] = counts[2] = 0;
    if (pt.owner(source) == pt.part)
      counts[0] = 1, q0[0] = pt.local(source);
  }
}
static __global__ void part_seed_kernel(partition_t pt, int source, int* dist, unsigned* visited) {
  if (pt.owner(source) == pt.part) {
    int l = pt.local(source);
    dist[l] = 0;
    visited[l >> 5] |= 1u << (l & 31);
  }
}

/// Device state of one rank's share of a partitioned BFS (allocated once per graph).
struct part_bfs_state_t {
  partition_t pt;
  dbuf_t<unsigned> visited, sent, fbm, nbm, unreachable, retry_map;
  dbuf_t<int2> first_nb;                 // per local row: first two in-neighbours (global ids; bfs_first_neighbor_kernel)
  graph_key_t unreachable_for, first_nb_for;
  dbuf_t<int> q[2], counts, send_count, overflow, dist;
  dbuf_t<int> send_buf;
  int send_cap = 0;
  int cur = 0;
  bool frontier_is_bitmap = false;
  struct host_fb_t {
    int count;
    int overflow;
    unsigned long long deg_sum;
    unsigned long long edges;
    int send_count[64];
  };
  host_fb_t* h_fb = nullptr;
  ctrl_t* last_ctrl = nullptr;
  ~part_bfs_state_t() {
    if (h_fb)
      cudaFreeHost(h_fb);
  }
  int local_words() const { return (pt.n_local + 31) / 32; }
  /// words per rank in the all-gathered frontier bitmap (identical on every rank)
  int words_per_rank() const { return (pt.rows_of(0) + 31) / 32; }
  void ensure(const partition_t& p, int send_capacity) {
    pt = p;
    size_t lw = static_cast<size_t>(words_per_rank()) + 4;
    visited.ensure(lw);
    fbm.ensure(lw);
    nbm.ensure(lw);
    retry_map.ensure(lw);
    first_nb.ensure(static_cast<size_t>(p.n_local) + 64);
    sent.ensure((static_cast<size_t>(p.n_global) + 31) / 32 + 4);
    q[0].ensure(static_cast<size_t>(p.n_local) + 64);
    q[1].ensure(static_cast<size_t>(p.n_local) + 64);
    dist.ensure(static_cast<size_t>(p.n_local) + 64);
    counts.ensure(8);
    send_count.ensure(64);
    overflow.ensure(4);
    send_cap = send_capacity;
    send_buf.ensure(static_cast<size_t>(p.nparts) * send_capacity + 64);
    if (!h_fb)
      B2G_CHECK(cudaMallocHost(&h_fb, sizeof(host_fb_t)));
  }
};

/// Fixed-layout message for the sync-free exchange: row o = [count_o, ids ... (<= cap_s)].
static __global__ void part_pack_kernel(const int* __restrict__ send_buf, const int* __restrict__ send_count,
                                        int send_cap, int nparts, int cap_s, int* __restrict__ msg) {
  for (int o = blockIdx.y; o < nparts; o += gridDim.y) {
    const int n = send_count[o];
    int* row = msg + static_cast<size_t>(o) * (cap_s + 1);
    if (blockIdx.x == 0 && threadIdx.x == 0)
      row[0] = n;
    const int m = min(n, cap_s);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      row[1 + i] = send_buf[static_cast<size_t>(o) * send_cap + i];
  }
}

/// Claim the ids of every peer's packed message (rows of cap_s+1 ints); a message whose count
/// exceeds cap_s raises *overflow (the host then falls back to the two-phase exchange).
static __global__ void part_claim_packed_kernel(partition_t pt, const int* __restrict__ msgs, int cap_s,
                                                unsigned* visited, int* dist, int next_level,
                                                const int* __restrict__ ro, int* out, int* out_count,
                                                unsigned long long* deg_sum, int* overflow) {
  const int lane = lane_id();
  unsigned long long ds = 0;
  for (int src = blockIdx.y; src < pt.nparts; src += gridDim.y) {
    if (src == pt.part)
      continue;
    const int* row = msgs + static_cast<size_t>(src) * (cap_s + 1);
    int n = row[0];
    if (n > cap_s) {
      if (blockIdx.x == 0 && threadIdx.x == 0)
        *overflow = 1;
      n = cap_s;
    }
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n; i0 += gridDim.x * blockDim.x) {
      int i = i0 + lane;
      bool won = false;
      int l = 0;
      if (i < n) {
        l = pt.local(row[1 + i]);
        won = bitmap_test_and_set(visited, l);
        if (won) {
          dist[l] = next_level;
          ds += static_cast<unsigned>(ro[l + 1] - ro[l]);
        }
      }
      unsigned m = __ballot_sync(kFull, won);
      if (m) {
        int base = 0;
        if (lane == 0)
          base = atomicAdd(out_count, __popc(m));
        base = __shfl_sync(kFull, base, 0);
        if (won)
          out[base + __popc(m & lanemask_lt())] = l;
      }
    }
  }
  ds = warp_sum(ds);
  if (lane == 0 && ds)
    atomicAdd(deg_sum, ds);
}

/// Level statistics into a device tensor (int64[4]: frontier size, its out-degree sum, edges
/// inspected, overflow) -- all-reduced by the host side without reading them first.
static __global__ void part_stats_kernel(const int* count, const ctrl_t* c, unsigned long long* extra_deg,
                                         const int* overflow, long long* stats) {
  stats[0] = *count;
  stats[1] = static_cast<long long>((c ? c->deg_sum : 0) + *extra_deg);
  stats[2] = static_cast<long long>(c ? c->edges : 0);
  stats[3] = (c ? c->overflow : 0) | *overflow;
  *extra_deg = 0;
}

static __global__ void part_feedback_kernel(const int* count, const ctrl_t* c, const int* send_count,
                                            const int* overflow, int nparts,
                                            part_bfs_state_t::host_fb_t* fb) {
  fb->count = *count;
  fb->overflow = (c ? c->overflow : 0) | *overflow;
  fb->deg_sum = c ? c->deg_sum : 0;
  fb->edges = c ? c->edges : 0;
  for (int i = 0; i < nparts && i < 64; ++i)
    fb->send_count[i] = send_count[i];
}

// ---------------------------------------------------------------------------------------------
// Partitioned SSSP: the same push exchange carrying (vertex, fp32 distance)
// pairs; the receiver applies atomicMin.  Distances only ever decrease and every improvement is
// re-expanded by its owner, so the distributed run reaches the same least fixed point as the
// single-GPU run and the reference's CPU validator (bit-exact; see sssp.cuh).
// ---------------------------------------------------------------------------------------------
struct part_relax_op {
  partition_t pt;
  float* dist;        // local rows
  int* stamp;         // local rows
  float* best_sent;   // global ids: smallest candidate this rank already forwarded
  int iteration;
  int* send_ids;      // nparts x send_cap
  float* send_vals;   // nparts x send_cap
  int* send_count;
  int send_cap;
  int* overflow;

  /// `src` is a LOCAL row id (the frontier holds local ids), `dst` a global id.
  __device__ __forceinline__ float prefetch(int dst) const {
    return pt.owner(dst) == pt.part ? ld_relaxed(dist + pt.local(dst)) : ld_relaxed(best_sent + dst);
  }
  __device__ __forceinline__ bool commit(int src, int dst, int, float w, float current) const {
    const float nd = __fadd_rn(ld_relaxed(dist + src), w);
    if (!(nd < current))
      return false;
    const int own = pt.owner(dst);
    if (own == pt.part) {
      const int l = pt.local(dst);
      float old = atomic_min_float(dist + l, nd);
      if (!(nd < old))
        return false;
      return atomicExch(stamp + l, iteration) != iteration;
    }
    float old = atomic_min_float(best_sent + dst, nd);
    if (!(nd < old))
      return false;
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, own);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (lane_id() == leader)
      base = atomicAdd(send_count + own, __popc(grp));
    base = __shfl_sync(grp, base, leader);
    const int slot = base + __popc(grp & lanemask_lt());
    if (slot < send_cap) {
      send_ids[static_cast<size_t>(own) * send_cap + slot] = dst;
      send_vals[static_cast<size_t>(own) * send_cap + slot] = nd;
    } else {
      *overflow = 1;
    }
    return false;
  }
  __device__ __forceinline__ bool operator()(int s, int d, int e, float w) const {
    return commit(s, d, e, w, prefetch(d));
  }
  __device__ __forceinline__ int emit_as(int dst) const { return pt.local(dst); }
};

/// msg row o = [count, ids[cap_s], value bits[cap_s]]
static __global__ void part_pack_pairs_kernel(const int* __restrict__ send_ids,
                                              const float* __restrict__ send_vals,
                                              const int* __restrict__ send_count, int send_cap,
                                              int nparts, int cap_s, int* __restrict__ msg) {
  for (int o = blockIdx.y; o < nparts; o += gridDim.y) {
    const int n = send_count[o];
    int* row = msg + static_cast<size_t>(o) * (2 * cap_s + 1);
    if (blockIdx.x == 0 && threadIdx.x == 0)
      row[0] = n;
    const int m = min(n, cap_s);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
      row[1 + i] = send_ids[static_cast<size_t>(o) * send_cap + i];
      row[1 + cap_s + i] = __float_as_int(send_vals[static_cast<size_t>(o) * send_cap + i]);
    }
  }
}

static __global__ void part_relax_packed_kernel(partition_t pt, const int* __restrict__ msgs, int cap_s,
                                                float* dist, int* stamp, int iteration,
                                                const int* __restrict__ ro, int* out, int* out_count,
                                                unsigned long long* deg_sum, int* overflow) {
  const int lane = lane_id();
  unsigned long long ds = 0;
  for (int src = blockIdx.y; src < pt.nparts; src += gridDim.y) {
    if (src == pt.part)
      continue;
    const int* row = msgs + static_cast<size_t>(src) * (2 * cap_s + 1);
    int n = row[0];
    if (n > cap_s) {
      if (blockIdx.x == 0 && threadIdx.x == 0)
        *overflow = 1;
      n = cap_s;
    }
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < n; i0 += gridDim.x * blockDim.x) {
      int i = i0 + lane;
      bool won = false;
      int l = 0;
      if (i < n) {
        l = pt.local(row[1 + i]);
        float nd = __int_as_float(row[1 + cap_s + i]);
        float old = atomic_min_float(dist + l, nd);
        if (nd < old)
          won = atomicExch(stamp + l, iteration) != iteration;
        if (won)
          ds += static_cast<unsigned>(ro[l + 1] - ro[l]);
      }
      unsigned m = __ballot_sync(kFull, won);
      if (m) {
        int base = 0;
        if (lane == 0)
          base = atomicAdd(out_count, __popc(m));
        base = __shfl_sync(kFull, base, 0);
        if (won)
          out[base + __popc(m & lanemask_lt())] = l;
      }
    }
  }
  ds = warp_sum(ds);
  if (lane == 0 && ds)
    atomicAdd(deg_sum, ds);
}

static __global__ void part_sssp_reset_kernel(partition_t pt, int source, float* dist, int* stamp,
                                              float* best_sent, int* q0, int* counts) {
  const int n = max(pt.n_local, pt.n_global);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < pt.n_local) {
      dist[i] = (pt.owner(source) == pt.part && pt.local(source) == i) ? 0.0f : FLT_MAX;
      stamp[i] = -1;
    }
    if (i < pt.n_global)
      best_sent[i] = FLT_MAX;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[0] = counts[1] = 0;
    if (pt.owner(source) == pt.part)
      counts[0] = 1, q0[0] = pt.local(source);
  }
}

struct part_sssp_state_t {
  partition_t pt;
  dbuf_t<float> dist, best_sent, send_vals;
  dbuf_t<int> stamp, q[2], counts, send_ids, send_count, overflow;
  int send_cap = 0, cur = 0;
  void ensure(const partition_t& p, int send_capacity) {
    pt = p;
    dist.ensure(static_cast<size_t>(p.n_local) + 64);
    stamp.ensure(static_cast<size_t>(p.n_local) + 64);
    best_sent.ensure(static_cast<size_t>(p.n_global) + 64);
    q[0].ensure(static_cast<size_t>(p.n_local) + 64);
    q[1].ensure(static_cast<size_t>(p.n_local) + 64);
    counts.ensure(8);
    send_count.ensure(64);
    overflow.ensure(4);
    send_cap = send_capacity;
    send_ids.ensure(static_cast<size_t>(p.nparts) * send_capacity + 64);
    send_vals.ensure(static_cast<size_t>(p.nparts) * send_capacity + 64);
  }
};

}  // namespace b200
}  // namespace gunrock
