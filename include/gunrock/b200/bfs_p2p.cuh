/**
 * @file bfs_p2p.cuh
 * @brief Multi-GPU BFS whose frontier exchange is done BY THE KERNELS over NVLink peer memory:
 * no NCCL call, no host code between the phases of a level.
 *
 * Same algorithm and partition as bfs_partitioned.cuh (cyclic 1-D vertex cut, direction-optimised,
 * one process per GPU).  Every rank owns one "window" of device memory that all peers map (CUDA IPC
 * between processes, plain pointers between simulated ranks of one process):
 *
 *   flags  [64]                 arrival epochs, slot r written by rank r          (barrier)
 *   stats  [2][16][4] int64     per-level statistics, slot r written by rank r    (all-reduce)
 *   front  [2][P * words]       frontier bitmap of ALL vertices, rank-major       (all-gather)
 *   inbox  [P][cap + 1]         row r = [count, ids...] forwarded by rank r       (all-to-all)
 *
 * top-down level : the ordinary advance with `p2p_claim_op` -- a remote neighbour seen for the first
 *                  time is stored straight into its owner's inbox row (slot reserved in a local,
 *                  warp-aggregated counter; the store itself crosses NVLink) -> `p2p_sync_kernel`
 *                  (publishes the counts, barrier) -> owner claims its inbox -> `p2p_sync_kernel`
 *                  (statistics: every rank writes its 4 numbers into every peer, barrier, local sum).
 * bottom-up level: `part_bottom_up_kernel` with a sink that stores each next-frontier word into
 *                  every peer's `front[next]` -- the sweep IS the all-gather -> `p2p_sync_kernel`.
 * The barrier is an epoch flag per peer (release store after __threadfence_system, acquire spin,
 * bounded by a time-out that raises an error instead of hanging the box).  `front` and `stats` are
 * double-buffered by level / epoch parity, so a rank that runs ahead never overwrites what a slower
 * peer is still reading.  The host (C++) reads ONE pinned feedback record per level.
 *
 * The reference has no multi-GPU execution; SURVEY.md 8e asks for the per-iteration
 * remote-frontier exchange over NVLink -- the NCCL variant of it lives in multi_gpu.py.
 */
#pragma once

#include <chrono>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <gunrock/b200/bfs_partitioned.cuh>

namespace gunrock {
namespace b200 {

constexpr int kMaxPeers = 16;

struct p2p_window_t {
  char* base[kMaxPeers];  // every rank's window as mapped in THIS process
  int nparts = 1;
  int me = 0;
  int words = 0;  // frontier words per rank (identical on every rank)
  int cap = 0;    // ids per inbox row

  static constexpr size_t kFlagsOff = 0;
  static constexpr size_t kStatsOff = 256;
  static constexpr size_t kFrontOff = 4096;
  __host__ __device__ static size_t align256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }
  __host__ __device__ size_t front_bytes() const {
    return align256(sizeof(unsigned) * static_cast<size_t>(nparts) * words);
  }
  __host__ __device__ size_t inbox_row_ints() const { return (static_cast<size_t>(cap) + 1 + 63) & ~static_cast<size_t>(63); }
  __host__ __device__ size_t inbox_off() const { return kFrontOff + 2 * front_bytes(); }
  __host__ __device__ size_t bytes() const {
    return inbox_off() + sizeof(int) * inbox_row_ints() * nparts;
  }
  __host__ __device__ unsigned* flags(int r) const { return reinterpret_cast<unsigned*>(base[r] + kFlagsOff); }
  __host__ __device__ long long* stats(int r, int parity) const {
    return reinterpret_cast<long long*>(base[r] + kStatsOff) + parity * (kMaxPeers * 4);
  }
  /// rank r's copy of the whole frontier bitmap (buffer `parity`)
  __host__ __device__ unsigned* front(int r, int parity) const {
    return reinterpret_cast<unsigned*>(base[r] + kFrontOff + parity * front_bytes());
  }
  /// row `from` of rank r's inbox
  __host__ __device__ int* inbox(int r, int from) const {
    return reinterpret_cast<int*>(base[r] + inbox_off()) + inbox_row_ints() * from;
  }
};

/// Top-down edge functor: local neighbours are claimed in place, remote ones are stored into the
/// owner's inbox over NVLink (each global id forwarded at most once per rank: the `sent` map).
struct p2p_claim_op {
  static constexpr bool kNeedsSource = false;
  partition_t pt;
  p2p_window_t w;
  unsigned* visited;
  unsigned* sent;
  int* dist;
  int next_level;
  int* send_count;  // local, nparts counters
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
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, own);
    const int leader = __ffs(grp) - 1;
    int slot = 0;
    if (lane_id() == leader)
      slot = atomicAdd(send_count + own, __popc(grp));
    slot = __shfl_sync(grp, slot, leader) + __popc(grp & lanemask_lt());
    if (slot < w.cap)
      w.inbox(own, pt.part)[1 + slot] = dst;  // peer store
    else
      *overflow = 1;
    return false;
  }
  __device__ __forceinline__ bool operator()(int s, int d, int e, float wgt) const {
    return commit(s, d, e, wgt, prefetch(d));
  }
  __device__ __forceinline__ int emit_as(int dst) const { return pt.local(dst); }
};

/// Sink of the bottom-up sweep that stores every next-frontier word into ALL ranks' copies of the
/// frontier bitmap: the sweep performs the all-gather itself.
struct peer_word_sink_t {
  p2p_window_t w;
  int parity;
  __device__ __forceinline__ void zero(int wi) const {  // lanes hold consecutive words: coalesced
    const size_t at = static_cast<size_t>(w.me) * w.words + wi;
    for (int r = 0; r < w.nparts; ++r)
      w.front(r, parity)[at] = 0;
  }
  __device__ __forceinline__ void word(int wi, unsigned v) const {  // warp-uniform: lane r -> peer r
    const int lane = lane_id();
    if (lane < w.nparts)
      w.front(lane, parity)[static_cast<size_t>(w.me) * w.words + wi] = v;
  }
};

/// Feedback record read by the host once per level (pinned memory).
struct p2p_feedback_t {
  long long count;     // global next-frontier size
  long long deg_sum;   // its out-degree sum
  long long edges;     // edges inspected by the level (all ranks)
  int overflow;
  int timed_out;
  int late_peer;            // diagnostics of a time-out: which peer, the epoch it had published,
  unsigned late_seen;       // and the epoch that was awaited
  unsigned late_epoch;
  volatile int seq;
};

/**
 * @brief Barrier across the ranks (+ optional payloads).  One warp; lane r talks to peer r.
 *   send_count != nullptr : publish inbox counts (top-down, before the owners claim)
 *   kStats                : all-reduce (sum) of this level's statistics and host feedback
 */
template <bool kStats>
__global__ void p2p_sync_kernel(p2p_window_t w, unsigned epoch, const int* send_count, const int* count,
                                const ctrl_t* c, unsigned long long* extra_deg, const int* overflow,
                                p2p_feedback_t* fb, int seq, unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  const int parity = epoch & 1;
  bool late = false;
  if (lane < w.nparts) {
    if (send_count)
      w.inbox(lane, w.me)[0] = lane == w.me ? 0 : send_count[lane];
    if (kStats) {
      long long* s = w.stats(lane, parity) + 4 * w.me;
      s[0] = *count;
      s[1] = static_cast<long long>((c ? c->deg_sum : 0) + *extra_deg);
      s[2] = static_cast<long long>(c ? c->edges : 0);
      s[3] = (c ? c->overflow : 0) | *overflow;
    }
    __threadfence_system();
    st_release_sys(w.flags(lane) + w.me, epoch);
    const unsigned* mine = w.flags(w.me) + lane;
    const unsigned long long t0 = global_timer_ns();
    // epochs only grow; signed distance tolerates wrap-around
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
      if (global_timer_ns() - t0 > timeout_ns) {
        late = true;
        fb->late_peer = lane;
        fb->late_seen = ld_acquire_sys(mine);
        fb->late_epoch = epoch;
        break;
      }
    }
  }
  late = __any_sync(kFull, late);
  if (kStats) {
    __syncwarp();
    if (lane == 0) {
      *extra_deg = 0;
      long long t[4] = {0, 0, 0, 0};
      const long long* s = w.stats(w.me, parity);
      for (int r = 0; r < w.nparts; ++r)
        for (int k = 0; k < 4; ++k)
          t[k] += s[4 * r + k];
      fb->count = t[0];
      fb->deg_sum = t[1];
      fb->edges = t[2];
      fb->overflow = t[3] != 0;
      if (late)
        fb->timed_out = 1;  // sticky: the host clears it when a run starts
      __threadfence_system();
      fb->seq = seq;
    }
  } else if (late && lane == 0) {
    fb->timed_out = 1;  // reported with the level's statistics record
  }
}

/// Report of p2p_tail_kernel (pinned host memory).
struct p2p_tail_report_t {
  int levels;     // levels executed by this launch (each consumed two barrier epochs)
  int cur;        // which local queue holds the frontier it stopped at
  int timed_out;
  int pad;
  long long count;    // GLOBAL size of that frontier
  long long deg_sum;  // its GLOBAL out-degree sum
  long long edges[16];     // global edges inspected per level
  long long frontier[16];  // global frontier size per level
  volatile int seq;
};

/**
 * @brief Distributed tail of the traversal in ONE launch per rank: while the global frontier stays
 * tiny, a single CTA per GPU runs level after level -- expand (remote neighbours stored into the
 * owners' inboxes), barrier, claim the inbox, statistics barrier -- without a host round trip or a
 * kernel boundary between levels (a top-down level otherwise costs ~5 launches and ~60 us of fixed
 * latency per rank).  All ranks enter and leave it on the same global statistics.
 */
template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1)  // one CTA per launch: take the registers
p2p_tail_kernel(csr_view_t g, partition_t pt, p2p_window_t w, unsigned epoch0, int* q0, int* q1, int* counts,
                int cur, int first_level, long long first_frontier, int max_levels, long long edge_budget,
                unsigned* visited, unsigned* sent, int* dist, int* overflow, p2p_tail_report_t* rep, int seq,
                unsigned long long timeout_ns) {
  __shared__ int s_cnt, s_late;
  __shared__ int s_send[kMaxPeers];
  __shared__ unsigned long long s_deg, s_edges;
  __shared__ long long s_glob[4];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int* __restrict__ ro = g.row_offsets;
  const int* __restrict__ ci = g.column_indices;
  int* q[2] = {q0, q1};
  int n = counts[cur];
  int level = first_level, done = 0;
  unsigned epoch = epoch0;
  long long glob_n = first_frontier, glob_m = 0;
  if (threadIdx.x == 0)
    s_late = 0;
  auto barrier = [&](unsigned e) {  // threads r < P: publish epoch e to peer r, wait for peer r
    if (threadIdx.x < w.nparts) {
      const int r = threadIdx.x;
      __threadfence_system();
      st_release_sys(w.flags(r) + w.me, e);
      const unsigned* mine = w.flags(w.me) + r;
      const unsigned long long t0 = global_timer_ns();
      while (static_cast<int>(ld_acquire_sys(mine) - e) < 0)
        if (global_timer_ns() - t0 > timeout_ns) {
          s_late = 1;
          break;
        }
    }
  };
  for (;;) {
    if (threadIdx.x == 0) {
      s_cnt = 0;
      s_deg = 0;
      s_edges = 0;
    }
    if (threadIdx.x < kMaxPeers)
      s_send[threadIdx.x] = 0;
    __syncthreads();
    p2p_claim_op op{pt, w, visited, sent, dist, level + 1, s_send, overflow};
    const int* in = q[cur];
    int* out = q[cur ^ 1];
    unsigned long long my_deg = 0, my_edges = 0;
    // ---- expand: one lane per frontier row, rows of 32+ edges walked by the whole warp -----------
    for (int base = 0; base < n; base += kThreads) {
      const int i = base + static_cast<int>(threadIdx.x);
      int v = -1, s = 0, d = 0;
      if (i < n) {
        v = in[i];
        s = ro[v];
        d = ro[v + 1] - s;
      }
      my_edges += static_cast<unsigned>(d);
      unsigned big = __ballot_sync(kFull, d >= 32);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bv = __shfl_sync(kFull, v, src), bs = __shfl_sync(kFull, s, src),
                  bd = __shfl_sync(kFull, d, src);
        for (int off = 0; off < bd; off += 32) {
          bool keep = false;
          int nb = -1;
          if (off + lane < bd) {
            nb = ci[bs + off + lane];
            keep = op(bv, nb, bs + off + lane, 1.0f);
          }
          const unsigned m = __ballot_sync(kFull, keep);
          if (m) {
            int at = 0;
            if (lane == 0)
              at = atomicAdd(&s_cnt, __popc(m));
            at = __shfl_sync(kFull, at, 0);
            if (keep) {
              const int x = pt.local(nb);
              out[at + __popc(m & lanemask_lt())] = x;
              my_deg += static_cast<unsigned>(ro[x + 1] - ro[x]);
            }
          }
        }
      }
      if (d < 32) {
        for (int k = 0; k < d; ++k) {
          const int nb = ci[s + k];
          if (op(v, nb, s + k, 1.0f)) {
            const int x = pt.local(nb);
            out[atomicAdd(&s_cnt, 1)] = x;
            my_deg += static_cast<unsigned>(ro[x + 1] - ro[x]);
          }
        }
      }
    }
    __syncthreads();  // every peer store of this CTA is ordered before the flags below
    // ---- barrier A: publish the inbox counts ------------------------------------------------------
    if (threadIdx.x < w.nparts)
      w.inbox(threadIdx.x, w.me)[0] = threadIdx.x == w.me ? 0 : min(s_send[threadIdx.x], w.cap);
    barrier(epoch);
    __syncthreads();
    // ---- claim what the peers forwarded (their stores bypass this SM's L1: read through L2) -------
    for (int src = 0; src < w.nparts; ++src) {
      if (src == w.me)
        continue;
      const int* row = w.inbox(w.me, src);
      const int n_in = __ldcg(row);
      for (int i0 = warp * 32; i0 < n_in; i0 += kThreads) {
        const int i = i0 + lane;
        bool won = false;
        int l = 0;
        if (i < n_in) {
          l = pt.local(__ldcg(row + 1 + i));
          won = bitmap_test_and_set(visited, l);
          if (won) {
            dist[l] = level + 1;
            my_deg += static_cast<unsigned>(ro[l + 1] - ro[l]);
          }
        }
        const unsigned m = __ballot_sync(kFull, won);
        if (m) {
          int base = 0;
          if (lane == 0)
            base = atomicAdd(&s_cnt, __popc(m));
          base = __shfl_sync(kFull, base, 0);
          if (won)
            out[base + __popc(m & lanemask_lt())] = l;
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
    // ---- statistics barrier: every rank writes its numbers into every peer, then sums ------------
    const int parity = (epoch + 1) & 1;
    if (threadIdx.x < w.nparts) {
      long long* st = w.stats(threadIdx.x, parity) + 4 * w.me;
      st[0] = s_cnt;
      st[1] = static_cast<long long>(s_deg);
      st[2] = static_cast<long long>(s_edges);
      st[3] = 0;
    }
    barrier(epoch + 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t[3] = {0, 0, 0};
      const long long* st = w.stats(w.me, parity);
      for (int r = 0; r < w.nparts; ++r)
        for (int k = 0; k < 3; ++k)
          t[k] += __ldcg(st + 4 * r + k);
      s_glob[0] = t[0];
      s_glob[1] = t[1];
      s_glob[2] = t[2];
      if (done < 16) {
        rep->edges[done] = t[2];
        rep->frontier[done] = glob_n;
      }
    }
    __syncthreads();
    glob_n = s_glob[0];
    glob_m = s_glob[1];
    n = s_cnt;
    cur ^= 1;
    ++level;
    ++done;
    epoch += 2;
    if (glob_n == 0 || glob_m >= edge_budget || done >= max_levels || s_late)
      break;
    __syncthreads();  // s_cnt / s_glob were read by everyone before they are cleared
  }
  if (threadIdx.x == 0) {
    counts[cur] = n;
    rep->levels = done;
    rep->cur = cur;
    rep->count = glob_n;
    rep->deg_sum = glob_m;
    rep->timed_out = s_late;
    __threadfence_system();
    rep->seq = seq;
  }
}

/// K2's edge count joins K1's control block (the statistics barrier reads ONE block)
static __global__ void part_fold_edges_kernel(ctrl_t* into, const ctrl_t* from) {
  into->edges += from->edges;
}

/// push this rank's segment of front[parity] (already complete in its own window) to every peer:
/// 16-byte stores, enough CTAs to keep NVLink busy (segments are 16-byte aligned: words % 4 == 0)
static __global__ void p2p_push_segment_kernel(p2p_window_t w, int parity) {
  const size_t at = static_cast<size_t>(w.me) * w.words;
  const uint4* src = reinterpret_cast<const uint4*>(w.front(w.me, parity) + at);
  const int n4 = w.words >> 2;
  for (int r = blockIdx.y; r < w.nparts; r += gridDim.y) {
    if (r == w.me)
      continue;
    uint4* dst = reinterpret_cast<uint4*>(w.front(r, parity) + at);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
}


/// How one rank runs its share of a partitioned BFS (same meaning as bfs_config_t of the single-GPU enactor).
struct part_bfs_config_t {
  advance_launch_t advance;
  int direction = 2;    // 0 push only, 1 pull after the first level, 2 optimised (Beamer on GLOBAL counts)
  double alpha = 14.0;
  double beta = 24.0;
};

/// What a rank reports about a run (global per-level statistics: every rank sees the same numbers).
struct part_bfs_report_t {
  int levels = 0;
  unsigned long long edges_total = 0, verts_total = 0;
  int level_direction[64] = {};
  int level_frontier[64] = {};
  unsigned long long level_edges[64] = {};
};

/// Peer-memory exchange state of one rank: its own window, the peers' mappings (CUDA IPC between processes,
/// plain peer pointers between the devices of one process).
struct p2p_state_t {
  p2p_window_t w;
  void* own = nullptr;
  size_t own_bytes = 0;
  void* opened[kMaxPeers] = {};  // cudaIpcOpenMemHandle mappings to close
  bool attached = false;
  unsigned epoch = 0;
  int seq = 0;
  p2p_feedback_t* h_fb = nullptr;
  p2p_tail_report_t* h_tail = nullptr;
  void release() {
    if (h_tail)
      cudaFreeHost(h_tail);
    h_tail = nullptr;
    for (auto& o : opened)
      if (o) {
        cudaIpcCloseMemHandle(o);
        o = nullptr;
      }
    if (own)
      cudaFree(own);
    own = nullptr;
    if (h_fb)
      cudaFreeHost(h_fb);
    h_fb = nullptr;
    attached = false;
  }
};


/**
 * @brief Allocate (once) this rank's window and everything a traversal needs, so that the run itself never
 * calls cudaMalloc / cudaFree (a device-wide synchronisation while a peer's barrier kernel is spinning).
 */
inline void part_p2p_prepare(workspace_t& ws, const csr_view_t& view, const partition_t& pt, part_bfs_state_t& S,
                             dbuf_t<unsigned long long>& part_deg, p2p_state_t& P, bool may_pull) {
  if (P.own)
    return;
  if (pt.nparts > kMaxPeers)
    throw std::runtime_error("peer-memory exchange: at most 16 ranks");
  P.w = p2p_window_t{};
  P.w.nparts = pt.nparts;
  P.w.me = pt.part;
  P.w.words = (((pt.rows_of(0) + 31) / 32) + 3) & ~3;  // 16-byte aligned segments
  // a global id is forwarded at most once per rank, so a row never holds more than the owner's rows
  P.w.cap = pt.rows_of(0) + 64;
  P.own_bytes = P.w.bytes();
  B2G_CHECK(cudaMalloc(&P.own, P.own_bytes));
  B2G_CHECK(cudaMemset(P.own, 0, P.own_bytes));
  B2G_CHECK(cudaMallocHost(&P.h_fb, sizeof(p2p_feedback_t)));
  memset(P.h_fb, 0, sizeof(p2p_feedback_t));
  B2G_CHECK(cudaMallocHost(&P.h_tail, sizeof(p2p_tail_report_t)));
  memset(P.h_tail, 0, sizeof(p2p_tail_report_t));
  S.ensure(pt, 1);
  part_deg.ensure(2);
  reserve_advance_workspace(ws, view, pt.n_local);
  if (may_pull)
    S.unreachable.ensure(static_cast<size_t>(S.words_per_rank()) + 4);
  B2G_CHECK(cudaDeviceSynchronize());
}

/// Map the peers' windows: plain device pointers (ranks of one process with peer access enabled) ...
inline void part_p2p_attach_pointers(p2p_state_t& P, void* const* windows) {
  for (int r = 0; r < P.w.nparts; ++r)
    P.w.base[r] = r == P.w.me ? static_cast<char*>(P.own) : static_cast<char*>(windows[r]);
  P.attached = true;
}

/**
 * @brief One rank's level loop of the direction-optimised BFS over a 1-D (cyclic) partitioned graph, the
 * frontier exchange done by the kernels over peer memory (file header).  COLLECTIVE: every rank calls it
 * with the same source / total_edges / cfg; the host reads one pinned feedback record per level.  Used by
 * the C ABI (`b2g_part_bfs_p2p`, one process per GPU) and by the header API (`bfs::run` with a multi-device
 * `gcuda::multi_context_t`, one host thread per device).  `in_view` = this rank's in-edge lists (the CSR
 * itself for a symmetric graph; row_offsets == nullptr disables pull).  Everything is enqueued on ws.stream.
 */
inline void part_bfs_p2p_run(workspace_t& ws, const csr_view_t& view, const csr_view_t& in_view,
                             const partition_t& pt, part_bfs_state_t& S, dbuf_t<unsigned long long>& part_deg,
                             p2p_state_t& P, int source, long long total_edges, const part_bfs_config_t& cfg,
                             part_bfs_report_t* out) {
  {
    cudaStream_t st = ws.stream;
    const p2p_window_t w = P.w;
    const int np = w.nparts;
    const int sms = device_info_t::get().sm_count;
    const bool can_pull = in_view.row_offsets != nullptr && cfg.direction != 0;
    const double alpha = cfg.alpha;
    const double beta = cfg.beta;
    static const bool trace = std::getenv("B2G_TRACE") != nullptr;
    static const char* timeout_env = std::getenv("B2G_P2P_TIMEOUT_MS");
    const unsigned long long timeout_ns =
        (timeout_env ? std::strtoull(timeout_env, nullptr, 10) : 20000ull) * 1000ull * 1000ull;

    // ---- reset (the part of b2g_part_bfs_begin that matters here; no send buffer) ----------------
    S.ensure(pt, 1);
    part_deg.ensure(2);
    const int sent_words = (pt.n_global + 31) / 32;
    const unsigned* premark = nullptr;
    if (can_pull) {
      if (!S.unreachable_for.matches(in_view)) {
        S.unreachable.ensure(static_cast<size_t>(S.words_per_rank()) + 4);
        bfs_unreachable_map_kernel<<<sms * 8, 256, 0, st>>>(in_view.row_offsets, pt.n_local,
                                                            S.unreachable.ptr);
        S.unreachable_for.set(in_view);
      }
      premark = S.unreachable.ptr;
      if (!S.first_nb_for.matches(in_view)) {  // the pull kernels' per-row shortcut (bfs.cuh), built once per graph
        bfs_first_neighbor_kernel<<<sms * 8, 256, 0, st>>>(in_view, S.first_nb.ptr);
        S.first_nb_for.set(in_view);
      }
    }
    part_reset_kernel<<<sms * 8, 256, 0, st>>>(pt, source, S.dist.ptr, S.visited.ptr, S.sent.ptr,
                                                sent_words, S.q[0].ptr, S.counts.ptr, premark);
    part_seed_kernel<<<1, 1, 0, st>>>(pt, source, S.dist.ptr, S.visited.ptr);
    B2G_CHECK(cudaMemsetAsync(S.overflow.ptr, 0, sizeof(int), st));
    B2G_CHECK(cudaMemsetAsync(part_deg.ptr, 0, 16, st));
    ws.launches += 2;
    P.h_fb->timed_out = 0;

    static const bool use_tail = std::getenv("B2G_P2P_NO_TAIL") == nullptr;
    const long long tail_budget = 1 << 16;  // global frontier out-degree below which the tail kernel runs
    const int push_ctas = std::max(8, std::min(sms * 4 / np, (w.words / 4 + 255) / 256));
    // B2G_TRACE: CUDA-event stamps between the phases of every level (device time of this rank)
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    auto mark = [&](const std::string& name) {
      if (!trace)
        return;
      cudaEvent_t e;
      B2G_CHECK(cudaEventCreate(&e));
      B2G_CHECK(cudaEventRecord(e, st));
      marks.emplace_back(name, e);
    };
    mark("begin");
    int cur = 0, level = 0, parity = 0;  // parity: which `front` buffer holds the current frontier
    bool is_bitmap = false, bottom_up = false;
    long long n_f = 1, m_f = 0, explored = 0;
    bool m_known = true;  // false after a pull level: K1 / K2 do not report the new frontier's out-degree sum
    unsigned long long edges_total = 0, verts_total = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto sync = [&](bool with_stats, const int* send_count, const int* count_ptr, const ctrl_t* c) {
      ++P.epoch;
      if (with_stats)
        p2p_sync_kernel<true><<<1, 32, 0, st>>>(w, P.epoch, send_count, count_ptr, c, part_deg.ptr,
                                                 S.overflow.ptr, P.h_fb, ++P.seq, timeout_ns);
      else
        p2p_sync_kernel<false><<<1, 32, 0, st>>>(w, P.epoch, send_count, nullptr, nullptr, nullptr,
                                                  nullptr, P.h_fb, 0, timeout_ns);
      ws.launches += 1;
    };
    while (n_f > 0) {
      mark("L" + std::to_string(level) + ":");
      bool go_up = false;
      if (can_pull && level > 0) {
        if (cfg.direction == 1)
          go_up = true;
        else if (!bottom_up)
          go_up = m_known && static_cast<double>(m_f) > static_cast<double>(total_edges - explored) / alpha;
        else
          go_up = !(static_cast<double>(n_f) < static_cast<double>(pt.n_global) / beta);
      }
      unsigned* my_front = w.front(w.me, parity) + static_cast<size_t>(w.me) * w.words;
      // ---- first push level after pull levels: learn the frontier's GLOBAL out-degree sum (the pull kernels do
      //      not compute it) with one more statistics barrier, once per run; the tail kernel below usually takes
      //      everything that is left
      if (!go_up && !m_known) {
        ctrl_t* cq = ws.next_ctrl();
        if (is_bitmap) {
          B2G_CHECK(cudaMemsetAsync(S.counts.ptr + cur, 0, sizeof(int), st));
          bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(my_front, S.local_words(), S.q[cur].ptr, S.counts.ptr + cur,
                                                          view.row_offsets, &cq->deg_sum);
          ws.launches += 1;
          is_bitmap = false;
        }
        sync(true, nullptr, S.counts.ptr + cur, cq);
        wait_for_sequence(&P.h_fb->seq, P.seq, st);
        if (P.h_fb->timed_out)
          throw std::runtime_error("b2g_part_bfs_p2p: a peer did not reach the statistics barrier of the pull -> push switch");
        n_f = P.h_fb->count;
        m_f = P.h_fb->deg_sum;
        m_known = true;
        bottom_up = false;
        mark("switch");
        continue;
      }
      // ---- tiny global frontier: the distributed tail kernel runs level after level on its own ------
      if (!go_up && level > 0 && use_tail && m_known && m_f < tail_budget) {
        if (is_bitmap) {
          B2G_CHECK(cudaMemsetAsync(S.counts.ptr + cur, 0, sizeof(int), st));
          bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(my_front, S.local_words(), S.q[cur].ptr,
                                                          S.counts.ptr + cur);
          ws.launches += 1;
          is_bitmap = false;
        }
        P.h_tail->timed_out = 0;
        p2p_tail_kernel<1024><<<1, 1024, 0, st>>>(
            view, pt, w, P.epoch + 1, S.q[0].ptr, S.q[1].ptr, S.counts.ptr, cur, level, n_f, 16,
            tail_budget, S.visited.ptr, S.sent.ptr, S.dist.ptr, S.overflow.ptr, P.h_tail, ++P.seq, timeout_ns);
        ws.launches += 1;
        mark("tail");
        wait_for_sequence(&P.h_tail->seq, P.seq, st);
        const p2p_tail_report_t& t = *P.h_tail;
        P.epoch += 2u * static_cast<unsigned>(t.levels);
        if (t.timed_out)
          throw std::runtime_error("b2g_part_bfs_p2p: a peer did not reach a barrier of the tail kernel (time-out)");
        for (int k = 0; k < t.levels; ++k) {
          if (out && level + k < 64) {
            out->level_direction[level + k] = 0;
            out->level_frontier[level + k] = static_cast<int>(t.frontier[k]);
            out->level_edges[level + k] = static_cast<unsigned long long>(t.edges[k]);
          }
          edges_total += static_cast<unsigned long long>(t.edges[k]);
          verts_total += static_cast<unsigned long long>(t.frontier[k]);
          explored += t.edges[k];
        }
        if (trace)
          std::fprintf(stderr, "[b2g-p2p] rank %d epoch %u levels %d..%d in the tail kernel, n_f=%lld m_f=%lld\n",
                       w.me, P.epoch, level, level + t.levels - 1, t.count, t.deg_sum);
        level += t.levels;
        cur = t.cur;
        n_f = t.count;
        m_f = t.deg_sum;
        m_known = true;
        bottom_up = false;
        continue;
      }
      if (level > 0)
        explored += m_f;
      ctrl_t* c = nullptr;
      const int* count_ptr = nullptr;
      if (go_up) {
        if (!is_bitmap) {  // queue -> bitmap in my segment, pushed to every peer, barrier
          B2G_CHECK(cudaMemsetAsync(my_front, 0, sizeof(unsigned) * w.words, st));
          part_queue_to_bitmap_kernel<<<sms * 4, 256, 0, st>>>(S.q[cur].ptr, S.counts.ptr + cur, my_front);
          if (np > 1) {
            p2p_push_segment_kernel<<<dim3(push_ctas, np), 256, 0, st>>>(w, parity);
            sync(false, nullptr, nullptr, nullptr);
          }
          ws.launches += 2;
          is_bitmap = true;
        }
        c = ws.next_ctrl();
        B2G_CHECK(cudaMemsetAsync(S.counts.ptr + 2, 0, sizeof(int), st));
        const unsigned* all = w.front(w.me, parity);
        // K1 + K2 over this rank's rows (bfs.cuh): next-frontier words land in my segment of front[parity ^ 1] in my
        // own window (K1 plain stores, K2 RED.OR), then one kernel pushes the segment to every peer
        unsigned* nxt_seg = w.front(w.me, parity ^ 1) + static_cast<size_t>(w.me) * w.words;
        const part_frontier_t in_frontier{pt, all, w.words};
        ctrl_t* c2 = ws.next_ctrl();
        bfs_pull_first_kernel<256><<<sms * 8, 256, 0, st>>>(pt.n_local, S.first_nb.ptr, S.visited.ptr, in_frontier,
                                                            nxt_seg, S.retry_map.ptr, S.dist.ptr, level + 1, c,
                                                            S.counts.ptr + 2, nullptr, 0, -1,
                                                            pull_batch_words(S.local_words(), sms * 64));
        bfs_pull_rest_kernel<256, 32, 8><<<sms * 6, 256, 0, st>>>(in_view, S.retry_map.ptr, S.visited.ptr, in_frontier,
                                                                  nxt_seg, S.dist.ptr, level + 1, c2, S.counts.ptr + 2,
                                                                  pull_batch_words(S.local_words(), sms * 48));
        part_fold_edges_kernel<<<1, 1, 0, st>>>(c, c2);
        mark("sweep");
        if (np > 1) {
          p2p_push_segment_kernel<<<dim3(push_ctas, np), 256, 0, st>>>(w, parity ^ 1);
          ws.launches += 1;
        }
        ws.launches += 2;
        mark("push");
        ws.launches += 1;
        parity ^= 1;
        count_ptr = S.counts.ptr + 2;
      } else {
        if (is_bitmap) {  // bitmap -> queue (my segment of the current frontier map)
          B2G_CHECK(cudaMemsetAsync(S.counts.ptr + cur, 0, sizeof(int), st));
          bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(my_front, S.local_words(), S.q[cur].ptr,
                                                          S.counts.ptr + cur);
          ws.launches += 1;
          is_bitmap = false;
        }
        const int nxt = cur ^ 1;
        B2G_CHECK(cudaMemsetAsync(S.counts.ptr + nxt, 0, sizeof(int), st));
        B2G_CHECK(cudaMemsetAsync(S.send_count.ptr, 0, 64 * sizeof(int), st));
        p2p_claim_op op{pt, w, S.visited.ptr, S.sent.ptr, S.dist.ptr, level + 1, S.send_count.ptr,
                        S.overflow.ptr};
        // same path selection as the single-GPU enactor (bfs.cuh), on this rank's share of the frontier
        advance_launch_t lcfg = cfg.advance;
        const long long m_rank = m_f / np;
        lcfg.avg_degree = (level > 0 && m_known && n_f > 0) ? static_cast<double>(m_f) / static_cast<double>(n_f) : 0.0;
        if (level == 0 || !m_known) {
          lcfg.lb = lb_t::block_mapped;  // rows of unknown total length
        } else if (m_rank < lcfg.small_frontier_edges) {
          lcfg.lb = lb_t::block_mapped;  // one kernel: warp / thread bins only
          lcfg.hub_threshold = 1 << 30;
        } else if (lcfg.lb == lb_t::merge_path && m_rank < lcfg.mid_frontier_edges) {
          lcfg.lb = lb_t::block_mapped;  // skip the scan + partition launches
        }
        launch_advance<advance_output_t::vertices, true, false>(
            ws, view, S.q[cur].ptr, S.counts.ptr + cur, pt.n_local, S.q[nxt].ptr,
            S.counts.ptr + nxt, pt.n_local, op, lcfg, &c);
        mark("advance");
        if (np > 1) {
          sync(false, S.send_count.ptr, nullptr, nullptr);
          mark("barrier");
          part_claim_packed_kernel<<<dim3(std::max(16, sms * 2 / np), np), 256, 0, st>>>(
              pt, w.inbox(w.me, 0), static_cast<int>(w.inbox_row_ints()) - 1, S.visited.ptr, S.dist.ptr,
              level + 1, view.row_offsets, S.q[nxt].ptr, S.counts.ptr + nxt, part_deg.ptr,
              S.overflow.ptr);
          ws.launches += 1;
          mark("claim");
        }
        cur = nxt;
        count_ptr = S.counts.ptr + cur;
      }
      sync(true, nullptr, count_ptr, c);
      mark("stats");
      wait_for_sequence(&P.h_fb->seq, P.seq, st);
      if (P.h_fb->timed_out)
        throw std::runtime_error("b2g_part_bfs_p2p: rank " + std::to_string(w.me) + " level " +
                                 std::to_string(level) + ": peer " + std::to_string(P.h_fb->late_peer) +
                                 " did not reach barrier epoch " + std::to_string(P.h_fb->late_epoch) +
                                 " (published " + std::to_string(P.h_fb->late_seen) + ", time-out)");
      if (P.h_fb->overflow)
        throw std::runtime_error("b2g_part_bfs_p2p: frontier / inbox overflow");
      if (trace)
        std::fprintf(stderr, "[b2g-p2p] rank %d epoch %u level %d %s n_f=%lld m_f=%lld edges=%lld t=%.1f us\n",
                     w.me, P.epoch, level, go_up ? "up" : "down", n_f, m_f, P.h_fb->edges,
                     std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
      if (out && level < 64) {
        out->level_direction[level] = go_up ? 1 : 0;
        out->level_frontier[level] = static_cast<int>(n_f);
        out->level_edges[level] = static_cast<unsigned long long>(P.h_fb->edges);
      }
      edges_total += static_cast<unsigned long long>(P.h_fb->edges);
      verts_total += static_cast<unsigned long long>(n_f);
      if (level == 0)
        explored += P.h_fb->edges;
      n_f = P.h_fb->count;
      m_f = P.h_fb->deg_sum;
      m_known = !go_up;
      if (!m_known)
        m_f = P.h_fb->edges;  // stand-in for the `explored` estimate only
      bottom_up = go_up;
      ++level;
    }
    B2G_CHECK(cudaStreamSynchronize(st));
    if (trace) {
      std::string line = "[b2g-p2p] rank " + std::to_string(w.me) + " phases (us):";
      for (size_t i = 1; i < marks.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
        char buf[64];
        std::snprintf(buf, sizeof buf, " %s=%.1f", marks[i].first.c_str(), ms * 1e3f);
        line += buf;
      }
      std::fprintf(stderr, "%s\n", line.c_str());
      for (auto& m : marks)
        cudaEventDestroy(m.second);
    }
    S.cur = cur;
    S.frontier_is_bitmap = false;
    if (out) {
      out->levels = level;
      out->edges_total = edges_total;
      out->verts_total = verts_total;
    }
  }
}

}  // namespace b200
}  // namespace gunrock
