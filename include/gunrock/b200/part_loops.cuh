/**
 * @file part_loops.cuh
 * @brief The iteration loops of the partitioned SSSP and PageRank, written once over an
 * EXCHANGE policy `X` -- the handful of collectives the loops need -- so that the same loop runs
 *   - one process per GPU over NCCL (`nccl_exchange_t`, bfs_nccl.cuh; C ABI b2g_part_sssp_nccl / b2g_part_pr_nccl),
 *   - several devices of ONE process behind `gcuda::multi_context_t` (`thread_exchange_t` below: one host thread
 *     per rank, peers' device memory read directly -- peer access or the same device -- and a host barrier between
 *     the threads; part_multi.cuh, `sssp::run` / `pr::run` with a context of several devices).
 * The per-rank kernels are those of bfs_partitioned.cuh (SSSP: `part_relax_op`, packed (vertex, distance) rows)
 * and pr.cuh (`part_pr_begin / prepare / pull`).  The reference has no multi-GPU execution.
 *
 * What `X` provides (all COLLECTIVE over the ranks, stream-ordered on `st` where they touch device memory):
 *   int  rank(), nparts();
 *   void all_to_all_rows(const int* out, int* in, size_t row_len, st)   row p of `out` -> rank p's row [rank()] of `in`
 *   void reduce_stats(ws, long long* d_stats, long long h_out[4], st)   sum of 4 x int64, result on the HOST
 *   void all_reduce_sum(double* v, size_t n, st)                        in place
 *   void all_gather(const float* local, float* all, size_t n, st)      rank-major
 *   float reduce_max(ws, const float* d_v, long long* d_scratch, st)    max of one non-negative float, on the HOST
 */
#pragma once

#include <cstring>
#include <stdexcept>

#include <gunrock/b200/bfs_p2p.cuh>  // kMaxPeers
#include <gunrock/b200/bfs_partitioned.cuh>
#include <gunrock/b200/pr.cuh>
#include <gunrock/b200/thread_hub.hxx>

namespace gunrock {
namespace b200 {

/// The source's degree on its owner, 0 elsewhere: the run's first reduction tells every rank the size of level 0.
static __global__ void part_seed_stats_kernel(partition_t pt, int source, const int* __restrict__ ro, long long* stats) {
  stats[0] = stats[1] = stats[2] = stats[3] = 0;
  if (pt.owner(source) == pt.part) {
    const int l = pt.local(source);
    stats[0] = 1;
    stats[1] = ro[l + 1] - ro[l];
  }
}

/// stats[0] = bit pattern of a non-negative float (orders like the float), for an integer max-reduction.
static __global__ void part_float_bits_kernel(const float* v, long long* stats) {
  stats[0] = __float_as_int(*v);
  stats[1] = stats[2] = stats[3] = 0;
}

static_assert(kHubMaxRanks == kMaxPeers, "thread_hub_t and the peer-memory windows agree on the rank limit");

struct peer_table_t {
  const void* p[kMaxPeers];
};
/// out[i] = sum over the ranks, in rank order, of their arrays (peer loads): the same bits on every rank.
template <typename T>
static __global__ void part_sum_peers_kernel(peer_table_t t, int nparts, size_t n, T* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    T acc = static_cast<const T*>(t.p[0])[i];
    for (int r = 1; r < nparts; ++r)
      acc += static_cast<const T*>(t.p[r])[i];
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Exchange between the threads of one process, one thread per rank.
// ---------------------------------------------------------------------------------------------------------------
/// One rank's end of the hub.  Needs the peers' device memory to be addressable from this rank's device (peer
/// access enabled, or ranks sharing a device).
struct thread_exchange_t {
  thread_hub_t* hub = nullptr;
  int rank_ = 0, nparts_ = 1;
  int stats_phase = 0, max_phase = 0;
  dbuf_t<double> tmp;
  long long* h_stats = nullptr;  // pinned: 4 x int64 + one float
  ~thread_exchange_t() {
    if (h_stats)
      cudaFreeHost(h_stats);
  }
  void bind(thread_hub_t* h, int rank, int nparts) {
    hub = h;
    rank_ = rank;
    nparts_ = nparts;
    stats_phase = max_phase = 0;
    if (!h_stats)
      B2G_CHECK(cudaMallocHost(&h_stats, 64));
  }
  int rank() const { return rank_; }
  int nparts() const { return nparts_; }

  void all_to_all_rows(const int* out, int* in, size_t row_len, cudaStream_t st) {
    hub->ptr[rank_] = out;
    B2G_CHECK(cudaStreamSynchronize(st));  // my rows are complete
    hub->barrier();
    for (int p = 0; p < nparts_; ++p)
      if (p != rank_)
        B2G_CHECK(cudaMemcpyAsync(in + p * row_len, static_cast<const int*>(hub->ptr[p]) + rank_ * row_len,
                                  row_len * sizeof(int), cudaMemcpyDefault, st));
    B2G_CHECK(cudaStreamSynchronize(st));
    hub->barrier();  // nobody refills its rows before every peer has read them
  }
  void reduce_stats(workspace_t&, long long* d_stats, long long h_out[4], cudaStream_t st) {
    B2G_CHECK(cudaMemcpyAsync(h_stats, d_stats, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    B2G_CHECK(cudaStreamSynchronize(st));
    const int ph = stats_phase++ & 1;  // two sets of slots: a rank can be at most one reduction ahead of a peer
    std::memcpy(hub->vals[ph][rank_], h_stats, 4 * sizeof(long long));
    hub->barrier();
    for (int k = 0; k < 4; ++k) {
      long long s = 0;
      for (int p = 0; p < nparts_; ++p)
        s += hub->vals[ph][p][k];
      h_out[k] = s;
    }
  }
  void all_reduce_sum(double* v, size_t n, cudaStream_t st) {
    if (nparts_ == 1)
      return;
    tmp.ensure(n + 16);
    hub->ptr[rank_] = v;
    B2G_CHECK(cudaStreamSynchronize(st));
    hub->barrier();
    peer_table_t t{};
    for (int p = 0; p < nparts_; ++p)
      t.p[p] = hub->ptr[p];
    const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
    part_sum_peers_kernel<double><<<grid, 256, 0, st>>>(t, nparts_, n, tmp.ptr);
    B2G_CHECK(cudaStreamSynchronize(st));
    hub->barrier();  // every peer has read my array: it may be overwritten now
    B2G_CHECK(cudaMemcpyAsync(v, tmp.ptr, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  void all_gather(const float* local, float* all, size_t n, cudaStream_t st) {
    hub->ptr[rank_] = local;
    B2G_CHECK(cudaStreamSynchronize(st));
    hub->barrier();
    for (int p = 0; p < nparts_; ++p)
      B2G_CHECK(cudaMemcpyAsync(all + p * n, hub->ptr[p], n * sizeof(float), cudaMemcpyDefault, st));
    B2G_CHECK(cudaStreamSynchronize(st));
    hub->barrier();
  }
  float reduce_max(workspace_t&, const float* d_v, long long*, cudaStream_t st) {
    float* h = reinterpret_cast<float*>(h_stats + 4);
    B2G_CHECK(cudaMemcpyAsync(h, d_v, sizeof(float), cudaMemcpyDeviceToHost, st));
    B2G_CHECK(cudaStreamSynchronize(st));
    const int ph = max_phase++ & 1;
    hub->fvals[ph][rank_] = *h;
    hub->barrier();
    float m = 0.0f;
    for (int p = 0; p < nparts_; ++p)
      m = hub->fvals[ph][p] > m ? hub->fvals[ph][p] : m;
    return m;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// SSSP: frontier Bellman-Ford, push exchange of (vertex, fp32 distance) pairs, the receiver applies atomicMin.
// ---------------------------------------------------------------------------------------------------------------
struct part_sssp_report_t {
  int iterations = 0;
  unsigned long long edges_relaxed = 0, verts_total = 0;
};

/**
 * @brief One rank's SSSP loop.  COLLECTIVE over `x`.  Per iteration: relax (advance with `part_relax_op`) ->
 * pack rows [count, ids[cap], distance bits[cap]] -> all-to-all of the rows, cap derived on every rank from the
 * frontier's GLOBAL out-degree sum (no rank forwards more pairs to a peer than the frontier has out-edges), so the
 * sizes agree without a count round trip -> apply -> reduction of 4 x int64 (next frontier size, its out-degree
 * sum, edges relaxed, overflow), the iteration's one host synchronisation.  A row that overflows restarts the run
 * with rows four times as long.  Result: `S.dist` (owned fp32 distances, bit-exact with the single-GPU run).
 */
template <typename X>
inline void part_sssp_run(workspace_t& ws, const csr_view_t& view, const partition_t& pt, part_sssp_state_t& S,
                          dbuf_t<unsigned long long>& part_deg, dbuf_t<int>& msg_out, dbuf_t<int>& msg_in,
                          dbuf_t<long long>& stats, X& x, int source, int send_capacity, const advance_launch_t& lcfg,
                          part_sssp_report_t* out) {
  if (!view.values)
    throw std::runtime_error("partitioned sssp: the graph has no edge values");
  cudaStream_t st = ws.stream;
  const int np = pt.nparts;
  const int rows = pt.rows_of(0) + 64;
  int cap_s = send_capacity > 0 ? send_capacity : std::min(rows, 1 << 20);
  stats.ensure(8);
  part_deg.ensure(2);
  int it = 0;
  unsigned long long relaxed = 0, verts = 0;
  long long h[4];
  for (;;) {
    S.ensure(pt, std::max(cap_s, rows));
    const size_t row_full = 2 * static_cast<size_t>(cap_s) + 1;
    msg_out.ensure(static_cast<size_t>(np) * row_full + 64);
    msg_in.ensure(static_cast<size_t>(np) * row_full + 64);
    part_sssp_reset_kernel<<<device_info_t::get().sm_count * 8, 256, 0, st>>>(pt, source, S.dist.ptr, S.stamp.ptr,
                                                                              S.best_sent.ptr, S.q[0].ptr, S.counts.ptr);
    B2G_CHECK(cudaMemsetAsync(S.overflow.ptr, 0, sizeof(int), st));
    B2G_CHECK(cudaMemsetAsync(part_deg.ptr, 0, 16, st));
    part_seed_stats_kernel<<<1, 1, 0, st>>>(pt, source, view.row_offsets, stats.ptr);
    ws.launches += 2;
    S.cur = 0;
    x.reduce_stats(ws, stats.ptr, h, st);  // every rank learns the source's degree: the bound of iteration 0's rows
    long long n_f = h[0], m_f = h[1];
    if (n_f != 1)
      throw std::runtime_error("partitioned sssp: the source is owned by no rank");
    bool overflowed = false;
    it = 0;
    relaxed = verts = 0;
    while (n_f > 0) {
      const int nxt = S.cur ^ 1;
      const int cap = static_cast<int>(std::min<long long>(cap_s, std::max<long long>(m_f, 256)));
      const size_t row = 2 * static_cast<size_t>(cap) + 1;
      B2G_CHECK(cudaMemsetAsync(S.counts.ptr + nxt, 0, sizeof(int), st));
      B2G_CHECK(cudaMemsetAsync(S.send_count.ptr, 0, 64 * sizeof(int), st));
      part_relax_op op{pt,          S.dist.ptr,     S.stamp.ptr,      S.best_sent.ptr, it,
                       S.send_ids.ptr, S.send_vals.ptr, S.send_count.ptr, S.send_cap,      S.overflow.ptr};
      ctrl_t* c = nullptr;
      launch_advance<advance_output_t::vertices, true, true>(ws, view, S.q[S.cur].ptr, S.counts.ptr + S.cur, pt.n_local,
                                                            S.q[nxt].ptr, S.counts.ptr + nxt, pt.n_local, op, lcfg, &c);
      if (np > 1) {
        part_pack_pairs_kernel<<<dim3(32, np), 256, 0, st>>>(S.send_ids.ptr, S.send_vals.ptr, S.send_count.ptr,
                                                             S.send_cap, np, cap, msg_out.ptr);
        x.all_to_all_rows(msg_out.ptr, msg_in.ptr, row, st);
        part_relax_packed_kernel<<<dim3(64, np), 256, 0, st>>>(pt, msg_in.ptr, cap, S.dist.ptr, S.stamp.ptr, it,
                                                               view.row_offsets, S.q[nxt].ptr, S.counts.ptr + nxt,
                                                               part_deg.ptr, S.overflow.ptr);
        ws.launches += 2;
      }
      S.cur = nxt;
      part_stats_kernel<<<1, 1, 0, st>>>(S.counts.ptr + S.cur, c, part_deg.ptr, S.overflow.ptr, stats.ptr);
      ws.launches += 1;
      x.reduce_stats(ws, stats.ptr, h, st);
      if (h[3]) {
        overflowed = true;
        break;
      }
      verts += static_cast<unsigned long long>(n_f);
      relaxed += static_cast<unsigned long long>(h[2]);
      n_f = h[0];
      m_f = h[1];
      ++it;
    }
    if (!overflowed)
      break;
    if (cap_s > (1 << 28))
      throw std::runtime_error("partitioned sssp: message rows overflow at the largest capacity");
    cap_s *= 4;
  }
  B2G_CHECK(cudaStreamSynchronize(st));
  if (out) {
    out->iterations = it;
    out->edges_relaxed = relaxed;
    out->verts_total = verts;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PageRank (pull): the rank owns the destination vertices v % P == rank with their in-edges.
// ---------------------------------------------------------------------------------------------------------------
/**
 * @brief One rank's PageRank loop, after `part_pr_begin`.  COLLECTIVE over `x`.  Per iteration: prepare
 * (c = plast * iweights, dangling partial) -> all-gather(c) + all-reduce(dangling, fp64 sum) -> pull ->
 * max-reduction of the error, the iteration's one host synchronisation.  Same recurrence and stopping rule as
 * include/gunrock/algorithms/pr.hxx:107-195 (converged = max |p - plast| < tol, checked once iteration >= 1).
 * Returns the iteration count; the owned ranks are in `S.p`.
 */
template <typename X>
inline int part_pr_run(workspace_t& ws, part_pr_state_t& S, dbuf_t<long long>& stats, X& x, float alpha, float tol,
                       int max_iter) {
  cudaStream_t st = ws.stream;
  const int np = x.nparts();
  const size_t R = static_cast<size_t>(S.rows_per_rank);
  S.c_local.ensure(R + 16);
  S.c_all.ensure(static_cast<size_t>(np) * R + 16);
  stats.ensure(8);
  B2G_CHECK(cudaMemsetAsync(S.c_local.ptr, 0, sizeof(float) * R, st));
  int it = 0;
  for (;;) {
    if (it > 0 && x.reduce_max(ws, S.err.ptr, stats.ptr, st) < tol)
      break;
    if (max_iter > 0 && it >= max_iter)
      break;
    part_pr_prepare(ws, S, alpha, S.c_local.ptr, S.dsum.ptr);
    const float* c_all = S.c_local.ptr;
    if (np > 1) {
      x.all_gather(S.c_local.ptr, S.c_all.ptr, R, st);
      x.all_reduce_sum(S.dsum.ptr, 1, st);
      c_all = S.c_all.ptr;
    }
    part_pr_pull(ws, S, alpha, c_all, S.dsum.ptr, S.err.ptr);
    ++it;
  }
  B2G_CHECK(cudaStreamSynchronize(st));
  return it;
}

}  // namespace b200
}  // namespace gunrock
