/**
 * @file bfs_nccl.cuh
 * @brief Multi-GPU BFS whose per-level frontier exchange is NCCL over NVLink / NVSwitch, driven from C++:
 * the level loop of SURVEY.md 8e ("NCCL all-to-all = ncclGroupStart; ncclSend / ncclRecv per peer;
 * ncclGroupEnd", "ncclAllGather of the frontier bitmap", "ncclAllReduce of the level statistics") with no
 * Python and no stream synchronisation inside a level -- everything of a level is enqueued on the rank's one
 * stream and the host polls ONE pinned record per level (runtime.cuh wait_for_sequence).
 *
 * Same partition, kernels and per-rank state as the peer-memory variant (bfs_partitioned.cuh, bfs_p2p.cuh):
 *   top-down level : advance with `part_claim_op`, whose send buffer IS the outgoing message -- row o =
 *                    [count, ids ... (<= cap)] -- then one grouped Send/Recv of cap + 1 ints per peer (cap is
 *                    derived on every rank from the GLOBAL frontier out-degree the previous level reported,
 *                    so the sizes agree without a count round trip), then `part_claim_packed_kernel`;
 *   bottom-up level: ncclAllGather of the ranks' frontier words (V / 8 bytes in total), then the purely local
 *                    sweep (`part_bottom_up_kernel`);
 *   every level    : one ncclAllReduce(sum) of 4 x int64 (frontier size, its out-degree sum, edges inspected,
 *                    overflow), copied to pinned memory by a 1-thread kernel.
 * The reference has no multi-GPU execution; NCCL is bound at run time (dlopen of libnccl.so.2:
 * the copy torch already loaded when one process per GPU runs under torch.distributed, the system one for a
 * plain C++ program), so nothing here adds a link-time dependency to the library or to users of the headers.
 */
#pragma once

#include <dlfcn.h>

#include <nccl.h>  // types and enums only; every entry point is resolved with dlsym

#include <gunrock/b200/bfs_p2p.cuh>
#include <gunrock/b200/part_loops.cuh>

namespace gunrock {
namespace b200 {

/// The NCCL entry points this file uses, bound once per process.
struct nccl_api_t {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  static nccl_api_t& get() {
    static nccl_api_t api = load();
    return api;
  }
  void check(ncclResult_t r, const char* what) const {
    if (r != ncclSuccess)
      throw std::runtime_error(std::string("NCCL: ") + what + " -> " + (GetErrorString ? GetErrorString(r) : "error"));
  }

 private:
  static nccl_api_t load() {
    nccl_api_t a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle)
        break;
    }
    if (!a.handle)
      throw std::runtime_error("NCCL exchange requested but libnccl.so.2 cannot be loaded (dlopen)");
    auto sym = [&](const char* n) {
      void* p = dlsym(a.handle, n);
      if (!p)
        throw std::runtime_error(std::string("libnccl: missing symbol ") + n);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.CommCount = reinterpret_cast<decltype(a.CommCount)>(sym("ncclCommCount"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    return a;
  }
};

/// Pinned record the host polls once per level.
struct nccl_feedback_t {
  long long v[4];  // global: next frontier size, its out-degree sum, edges inspected, overflow
  volatile int seq;
};

/// NCCL-exchange state of one rank.
struct nccl_state_t {
  ncclComm_t comm = nullptr;
  bool owns_comm = false;
  dbuf_t<int> msg_out, msg_in;   // nparts rows of cap + 1 ints
  dbuf_t<unsigned> all;          // the all-gathered frontier words, rank-major
  dbuf_t<long long> stats;       // 4 x int64, all-reduced in place
  nccl_feedback_t* h_fb = nullptr;
  int seq = 0;
  size_t cap_full = 0;
  void release() {
    if (h_fb)
      cudaFreeHost(h_fb);
    h_fb = nullptr;
    if (comm && owns_comm)
      nccl_api_t::get().CommDestroy(comm);
    comm = nullptr;
  }
  ~nccl_state_t() { release(); }
  /// Everything a traversal needs, allocated once (the run itself allocates nothing).
  void prepare(workspace_t& ws, const csr_view_t& view, const partition_t& pt, part_bfs_state_t& S,
               dbuf_t<unsigned long long>& part_deg, bool may_pull) {
    cap_full = static_cast<size_t>(pt.rows_of(0)) + 64;  // a peer is sent at most the rows it owns, once each
    msg_out.ensure(static_cast<size_t>(pt.nparts) * (cap_full + 1) + 64);
    msg_in.ensure(static_cast<size_t>(pt.nparts) * (cap_full + 1) + 64);
    S.ensure(pt, 1);
    all.ensure(static_cast<size_t>(pt.nparts) * S.words_per_rank() + 64);
    stats.ensure(8);
    part_deg.ensure(2);
    reserve_advance_workspace(ws, view, pt.n_local);
    if (may_pull)
      S.unreachable.ensure(static_cast<size_t>(S.words_per_rank()) + 4);
    if (!h_fb) {
      B2G_CHECK(cudaMallocHost(&h_fb, sizeof(nccl_feedback_t)));
      memset(h_fb, 0, sizeof(nccl_feedback_t));
    }
  }
};

/// count header of every outgoing row: msg[o * (cap + 1)] = min(send_count[o], cap)
static __global__ void nccl_row_headers_kernel(const int* __restrict__ send_count, int nparts, int cap,
                                               int* __restrict__ msg, int* overflow) {
  const int o = threadIdx.x;
  if (o < nparts) {
    const int n = send_count[o];
    if (n > cap)
      *overflow = 1;
    msg[static_cast<size_t>(o) * (cap + 1)] = min(n, cap);
  }
}

static __global__ void nccl_feedback_kernel(const long long* __restrict__ stats, nccl_feedback_t* fb, int seq) {
  fb->v[0] = stats[0];
  fb->v[1] = stats[1];
  fb->v[2] = stats[2];
  fb->v[3] = stats[3];
  __threadfence_system();
  fb->seq = seq;
}

/// The exchange policy of part_loops.cuh over one rank's NCCL communicator (SSSP / PageRank loops).
struct nccl_exchange_t {
  nccl_state_t* N = nullptr;
  int rank_ = 0, nparts_ = 1;
  int rank() const { return rank_; }
  int nparts() const { return nparts_; }
  void publish_and_wait(workspace_t& ws, cudaStream_t st) {
    nccl_feedback_kernel<<<1, 1, 0, st>>>(N->stats.ptr, N->h_fb, ++N->seq);
    ws.launches += 1;
    wait_for_sequence(&N->h_fb->seq, N->seq, st);
  }
  void all_to_all_rows(const int* out, int* in, size_t row_len, cudaStream_t st) {
    const nccl_api_t& nccl = nccl_api_t::get();
    nccl.check(nccl.GroupStart(), "ncclGroupStart");
    for (int p = 0; p < nparts_; ++p) {
      if (p == rank_)
        continue;
      nccl.check(nccl.Send(out + p * row_len, row_len, ncclInt32, p, N->comm, st), "ncclSend");
      nccl.check(nccl.Recv(in + p * row_len, row_len, ncclInt32, p, N->comm, st), "ncclRecv");
    }
    nccl.check(nccl.GroupEnd(), "ncclGroupEnd");
  }
  /// `d_stats` must be N->stats.ptr (the record the feedback kernel publishes).
  void reduce_stats(workspace_t& ws, long long* d_stats, long long h_out[4], cudaStream_t st) {
    const nccl_api_t& nccl = nccl_api_t::get();
    if (nparts_ > 1)
      nccl.check(nccl.AllReduce(d_stats, d_stats, 4, ncclInt64, ncclSum, N->comm, st), "ncclAllReduce(stats)");
    publish_and_wait(ws, st);
    for (int k = 0; k < 4; ++k)
      h_out[k] = N->h_fb->v[k];
  }
  void all_reduce_sum(double* v, size_t n, cudaStream_t st) {
    if (nparts_ > 1)
      nccl_api_t::get().check(nccl_api_t::get().AllReduce(v, v, n, ncclFloat64, ncclSum, N->comm, st),
                              "ncclAllReduce(fp64 sum)");
  }
  void all_reduce_sum(int* v, size_t n, cudaStream_t st) {
    if (nparts_ > 1)
      nccl_api_t::get().check(nccl_api_t::get().AllReduce(v, v, n, ncclInt32, ncclSum, N->comm, st),
                              "ncclAllReduce(int32 sum)");
  }
  void all_gather(const float* local, float* all, size_t n, cudaStream_t st) {
    nccl_api_t::get().check(nccl_api_t::get().AllGather(local, all, n, ncclFloat32, N->comm, st), "ncclAllGather");
  }
  /// max of one non-negative float: its bit pattern reduced as an integer (`d_scratch` must be N->stats.ptr).
  float reduce_max(workspace_t& ws, const float* d_v, long long* d_scratch, cudaStream_t st) {
    part_float_bits_kernel<<<1, 1, 0, st>>>(d_v, d_scratch);
    ws.launches += 1;
    if (nparts_ > 1)
      nccl_api_t::get().check(nccl_api_t::get().AllReduce(d_scratch, d_scratch, 1, ncclInt64, ncclMax, N->comm, st),
                              "ncclAllReduce(max)");
    publish_and_wait(ws, st);
    const int bits = static_cast<int>(N->h_fb->v[0]);
    float f;
    std::memcpy(&f, &bits, sizeof f);
    return f;
  }
};

/**
 * @brief One rank's level loop with the NCCL exchange (file header).  COLLECTIVE over the communicator in `N`.
 * Same arguments and result as `part_bfs_p2p_run`.
 */
inline void part_bfs_nccl_run(workspace_t& ws, const csr_view_t& view, const csr_view_t& in_view,
                              const partition_t& pt, part_bfs_state_t& S, dbuf_t<unsigned long long>& part_deg,
                              nccl_state_t& N, int source, long long total_edges, const part_bfs_config_t& cfg,
                              part_bfs_report_t* out) {
  const nccl_api_t& nccl = nccl_api_t::get();
  if (!N.comm)
    throw std::runtime_error("part_bfs_nccl_run: no communicator");
  cudaStream_t st = ws.stream;
  const int np = pt.nparts;
  const int sms = device_info_t::get().sm_count;
  const bool can_pull = in_view.row_offsets != nullptr && cfg.direction != 0;
  static const bool trace = std::getenv("B2G_TRACE") != nullptr;
  const int wpr = S.words_per_rank();

  // ---- reset ---------------------------------------------------------------------------------------------
  const int sent_words = (pt.n_global + 31) / 32;
  const unsigned* premark = nullptr;
  if (can_pull) {
    if (!S.unreachable_for.matches(in_view)) {
      S.unreachable.ensure(static_cast<size_t>(wpr) + 4);
      bfs_unreachable_map_kernel<<<sms * 8, 256, 0, st>>>(in_view.row_offsets, pt.n_local, S.unreachable.ptr);
      S.unreachable_for.set(in_view);
    }
    premark = S.unreachable.ptr;
    if (!S.first_nb_for.matches(in_view)) {  // the pull kernels' per-row shortcut (bfs.cuh), built once per graph
      bfs_first_neighbor_kernel<<<sms * 8, 256, 0, st>>>(in_view, S.first_nb.ptr);
      S.first_nb_for.set(in_view);
    }
  }
  part_reset_kernel<<<sms * 8, 256, 0, st>>>(pt, source, S.dist.ptr, S.visited.ptr, S.sent.ptr, sent_words,
                                              S.q[0].ptr, S.counts.ptr, premark);
  part_seed_kernel<<<1, 1, 0, st>>>(pt, source, S.dist.ptr, S.visited.ptr);
  B2G_CHECK(cudaMemsetAsync(S.overflow.ptr, 0, sizeof(int), st));
  B2G_CHECK(cudaMemsetAsync(part_deg.ptr, 0, 16, st));
  part_seed_stats_kernel<<<1, 1, 0, st>>>(pt, source, view.row_offsets, N.stats.ptr);
  ws.launches += 3;
  auto reduce_and_publish = [&]() {
    if (np > 1)
      nccl.check(nccl.AllReduce(N.stats.ptr, N.stats.ptr, 4, ncclInt64, ncclSum, N.comm, st), "ncclAllReduce(stats)");
    nccl_feedback_kernel<<<1, 1, 0, st>>>(N.stats.ptr, N.h_fb, ++N.seq);
    ws.launches += 1;
    wait_for_sequence(&N.h_fb->seq, N.seq, st);
  };
  reduce_and_publish();  // every rank learns the source's degree (the size of level 0's exchange)
  long long n_f = N.h_fb->v[0], m_f = N.h_fb->v[1], explored = 0;
  bool m_known = true;  // false after a pull level: K1 / K2 do not report the new frontier's out-degree sum
  if (n_f != 1)
    throw std::runtime_error("part_bfs_nccl_run: the source is owned by no rank");

  int cur = 0, level = 0;
  bool is_bitmap = false, bottom_up = false;
  unsigned long long edges_total = 0, verts_total = 0;
  unsigned* fbm = S.fbm.ptr;  // local words of the current / next frontier (bottom-up representation)
  unsigned* nbm = S.nbm.ptr;
  const auto t0 = std::chrono::steady_clock::now();
  while (n_f > 0) {
    bool go_up = false;
    if (can_pull && level > 0) {
      if (cfg.direction == 1)
        go_up = true;
      else if (!bottom_up)
        go_up = m_known && static_cast<double>(m_f) > static_cast<double>(total_edges - explored) / cfg.alpha;
      else
        go_up = !(static_cast<double>(n_f) < static_cast<double>(pt.n_global) / cfg.beta);
    }
    if (!go_up && !m_known) {
      // first push level after pull levels: learn the frontier's GLOBAL out-degree sum (the pull kernels do not
      // compute it) with one more all-reduce, once per run -- it sizes the exchange of the coming push levels
      ctrl_t* cq = ws.next_ctrl();
      if (is_bitmap) {
        B2G_CHECK(cudaMemsetAsync(S.counts.ptr + cur, 0, sizeof(int), st));
        bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(fbm, S.local_words(), S.q[cur].ptr, S.counts.ptr + cur,
                                                        view.row_offsets, &cq->deg_sum);
        ws.launches += 1;
        is_bitmap = false;
      }
      part_stats_kernel<<<1, 1, 0, st>>>(S.counts.ptr + cur, cq, part_deg.ptr, S.overflow.ptr, N.stats.ptr);
      ws.launches += 1;
      reduce_and_publish();
      n_f = N.h_fb->v[0];
      m_f = N.h_fb->v[1];
      m_known = true;
      bottom_up = false;
      continue;
    }
    explored += m_f;
    ctrl_t* c = nullptr;
    const int* count_ptr = nullptr;
    if (go_up) {
      if (!is_bitmap) {  // queue -> my words of the frontier bitmap
        B2G_CHECK(cudaMemsetAsync(fbm, 0, sizeof(unsigned) * wpr, st));
        part_queue_to_bitmap_kernel<<<sms * 4, 256, 0, st>>>(S.q[cur].ptr, S.counts.ptr + cur, fbm);
        ws.launches += 1;
        is_bitmap = true;
      }
      const unsigned* all = fbm;
      if (np > 1) {
        nccl.check(nccl.AllGather(fbm, N.all.ptr, static_cast<size_t>(wpr), ncclUint32, N.comm, st),
                   "ncclAllGather(frontier)");
        all = N.all.ptr;
      }
      c = ws.next_ctrl();
      ctrl_t* c2 = ws.next_ctrl();
      B2G_CHECK(cudaMemsetAsync(S.counts.ptr + 2, 0, sizeof(int), st));
      // K1 + K2 over this rank's rows (bfs.cuh): K1 writes every next-frontier word, K2 ORs its finds in
      const part_frontier_t in_frontier{pt, all, wpr};
      bfs_pull_first_kernel<256><<<sms * 8, 256, 0, st>>>(pt.n_local, S.first_nb.ptr, S.visited.ptr, in_frontier, nbm,
                                                          S.retry_map.ptr, S.dist.ptr, level + 1, c, S.counts.ptr + 2, nullptr, 0,
                                                          -1, pull_batch_words(S.local_words(), sms * 64));
      bfs_pull_rest_kernel<256, 32, 8><<<sms * 6, 256, 0, st>>>(in_view, S.retry_map.ptr, S.visited.ptr, in_frontier, nbm,
                                                                S.dist.ptr, level + 1, c2, S.counts.ptr + 2,
                                                                pull_batch_words(S.local_words(), sms * 48));
      part_fold_edges_kernel<<<1, 1, 0, st>>>(c, c2);
      ws.launches += 3;
      std::swap(fbm, nbm);
      count_ptr = S.counts.ptr + 2;
    } else {
      if (is_bitmap) {  // my words -> queue
        B2G_CHECK(cudaMemsetAsync(S.counts.ptr + cur, 0, sizeof(int), st));
        bitmap_to_queue_kernel<<<sms * 4, 256, 0, st>>>(fbm, S.local_words(), S.q[cur].ptr, S.counts.ptr + cur);
        ws.launches += 1;
        is_bitmap = false;
      }
      const int nxt = cur ^ 1;
      // no rank can forward more ids to a peer than the frontier has out-edges, nor than the peer owns rows
      long long cap_ll = !m_known ? static_cast<long long>(N.cap_full) : (m_f < 256 ? 256 : m_f);
      if (cap_ll > static_cast<long long>(N.cap_full))
        cap_ll = static_cast<long long>(N.cap_full);
      const int cap = static_cast<int>(cap_ll);
      const size_t row = static_cast<size_t>(cap) + 1;
      B2G_CHECK(cudaMemsetAsync(S.counts.ptr + nxt, 0, sizeof(int), st));
      B2G_CHECK(cudaMemsetAsync(S.send_count.ptr, 0, 64 * sizeof(int), st));
      // the functor's send buffer IS the outgoing message: row o starts at msg_out + o * (cap + 1), ids from + 1
      part_claim_op op{pt, S.visited.ptr, S.sent.ptr, S.dist.ptr, level + 1, N.msg_out.ptr + 1, S.send_count.ptr,
                       static_cast<int>(row), S.overflow.ptr};
      advance_launch_t lcfg = cfg.advance;
      const long long m_rank = m_f / np;
      lcfg.avg_degree = (level > 0 && m_known && n_f > 0) ? static_cast<double>(m_f) / static_cast<double>(n_f) : 0.0;
      if (level == 0 || !m_known) {
        lcfg.lb = lb_t::block_mapped;  // rows of unknown total length: binned kernel + hub slabs
      } else if (m_rank < lcfg.small_frontier_edges) {
        lcfg.lb = lb_t::block_mapped;
        lcfg.hub_threshold = 1 << 30;
      } else if (lcfg.lb == lb_t::merge_path && m_rank < lcfg.mid_frontier_edges) {
        lcfg.lb = lb_t::block_mapped;
      }
      launch_advance<advance_output_t::vertices, true, false>(ws, view, S.q[cur].ptr, S.counts.ptr + cur, pt.n_local,
                                                               S.q[nxt].ptr, S.counts.ptr + nxt, pt.n_local, op, lcfg,
                                                               &c);
      if (np > 1) {
        nccl_row_headers_kernel<<<1, 64, 0, st>>>(S.send_count.ptr, np, cap, N.msg_out.ptr, S.overflow.ptr);
        nccl.check(nccl.GroupStart(), "ncclGroupStart");
        for (int p = 0; p < np; ++p) {
          if (p == pt.part)
            continue;
          nccl.check(nccl.Send(N.msg_out.ptr + p * row, row, ncclInt32, p, N.comm, st), "ncclSend");
          nccl.check(nccl.Recv(N.msg_in.ptr + p * row, row, ncclInt32, p, N.comm, st), "ncclRecv");
        }
        nccl.check(nccl.GroupEnd(), "ncclGroupEnd");
        part_claim_packed_kernel<<<dim3(std::max(16, sms * 2 / np), np), 256, 0, st>>>(
            pt, N.msg_in.ptr, cap, S.visited.ptr, S.dist.ptr, level + 1, view.row_offsets, S.q[nxt].ptr,
            S.counts.ptr + nxt, part_deg.ptr, S.overflow.ptr);
        ws.launches += 2;
      }
      cur = nxt;
      count_ptr = S.counts.ptr + cur;
    }
    part_stats_kernel<<<1, 1, 0, st>>>(count_ptr, c, part_deg.ptr, S.overflow.ptr, N.stats.ptr);
    ws.launches += 1;
    reduce_and_publish();
    if (N.h_fb->v[3])
      throw std::runtime_error("part_bfs_nccl_run: frontier / message overflow");
    if (trace)
      std::fprintf(stderr, "[b2g-nccl] rank %d level %d %s n_f=%lld m_f=%lld edges=%lld t=%.1f us\n", pt.part, level,
                   go_up ? "up" : "down", n_f, m_f, N.h_fb->v[2],
                   std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    if (out && level < 64) {
      out->level_direction[level] = go_up ? 1 : 0;
      out->level_frontier[level] = static_cast<int>(n_f);
      out->level_edges[level] = static_cast<unsigned long long>(N.h_fb->v[2]);
    }
    edges_total += static_cast<unsigned long long>(N.h_fb->v[2]);
    verts_total += static_cast<unsigned long long>(n_f);
    n_f = N.h_fb->v[0];
    m_f = N.h_fb->v[1];
    m_known = !go_up;
    if (!m_known)
      m_f = N.h_fb->v[2];  // stand-in for the `explored` estimate only
    bottom_up = go_up;
    ++level;
  }
  B2G_CHECK(cudaStreamSynchronize(st));
  S.cur = cur;
  S.frontier_is_bitmap = false;
  if (out) {
    out->levels = level;
    out->edges_total = edges_total;
    out->verts_total = verts_total;
  }
}

}  // namespace b200
}  // namespace gunrock
