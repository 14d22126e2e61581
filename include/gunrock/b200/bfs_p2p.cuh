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

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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

/// push this rank's segment of front[parity] (already complete in its own window) to every peer
static __global__ void p2p_push_segment_kernel(p2p_window_t w, int parity) {
  const size_t at = static_cast<size_t>(w.me) * w.words;
  const unsigned* src = w.front(w.me, parity) + at;
  for (int r = blockIdx.y; r < w.nparts; r += gridDim.y) {
    if (r == w.me)
      continue;
    unsigned* dst = w.front(r, parity) + at;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.words; i += gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
}

}  // namespace b200
}  // namespace gunrock
