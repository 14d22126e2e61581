/**
 * @file sssp.cuh
 * @brief Frontier Bellman-Ford SSSP enactor on the compacting advance kernels.
 *
 * Path replaced: gunrock::sssp::enactor_t::loop, include/gunrock/algorithms/sssp.hxx:104-159
 * (advance with the `shortest_path` lambda :116-130, bypass filter with `remove_completed_paths`
 * :132-151).  Result contract (sssp.hxx:66-79, examples/algorithms/sssp/sssp_cpu.hxx:36-67):
 * fp32 distances, FLT_MAX when unreachable, source = 0.
 *
 * Bit-exactness: every stored value is fl32(dist[u] + w) for some in-edge, dist only decreases,
 * and a vertex whose distance dropped is always re-expanded, so the run ends at the least fixed
 * point of d[v] = min_u fl32(d[u] + w_uv) -- the same array the reference's lazy Dijkstra reaches
 * (SURVEY.md 8a R16).  The add is a plain fp32 add (no FMA contraction is possible: one operand
 * pair), the min is an integer atomicMin/atomicMax on the IEEE bit pattern, which is order
 * preserving for same-sign floats -- cheaper than the reference's CAS loop
 * (include/gunrock/cuda/atomic_functions.hxx:35-45) and returns the same old value.
 *
 * The reference's racy `visited[v] == iteration` dedup (sssp.hxx:134-139) becomes an atomicExch
 * on the stamp inside the edge functor, so a vertex enters the next frontier exactly once per
 * iteration and the separate filter pass disappears.
 */
#pragma once

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include <gunrock/b200/advance.cuh>

namespace gunrock {
namespace b200 {

/// fp32 atomic min valid for any mix of signs (no NaNs): returns the previous value.
__device__ __forceinline__ float atomic_min_float(float* addr, float value) {
  return (value >= 0.0f)
             ? __int_as_float(atomicMin(reinterpret_cast<int*>(addr), __float_as_int(value)))
             : __uint_as_float(
                   atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(value)));
}

struct sssp_relax_op {
  static constexpr int kMergePathKernel = 4;  // warp-private spans, 8 chunks in flight (advance.cuh)
  float* dist;
  int* stamp;
  int iteration;
  __device__ __forceinline__ bool operator()(int src, int dst, int, float w) const {
    float nd = __fadd_rn(ld_relaxed(dist + src), w);
    float old = atomic_min_float(dist + dst, nd);
    if (!(nd < old))
      return false;
    return atomicExch(stamp + dst, iteration) != iteration;
  }
  /// two-phase protocol: read the target's current distance first; a candidate that cannot
  /// lower it needs no atomic at all (distances only decrease, so a stale read is conservative:
  /// it can only let through a candidate the atomicMin then rejects).
  __device__ __forceinline__ float prefetch(int dst) const { return ld_relaxed(dist + dst); }
  __device__ __forceinline__ bool commit(int src, int dst, int, float w, float current) const {
    float nd = __fadd_rn(ld_relaxed(dist + src), w);
    if (!(nd < current))
      return false;
    float old = atomic_min_float(dist + dst, nd);
    if (!(nd < old))
      return false;
    return atomicExch(stamp + dst, iteration) != iteration;
  }
};

struct sssp_relax_maker {
  float* dist;
  int* stamp;
  __device__ __forceinline__ sssp_relax_op operator()(int iteration) const {
    return sssp_relax_op{dist, stamp, iteration};
  }
};

static __global__ void sssp_reset_kernel(float* dist, int* stamp, int n_vertices, int source, int* q0,
                                  int* counts) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vertices;
       i += gridDim.x * blockDim.x) {
    dist[i] = (i == source) ? 0.0f : FLT_MAX;
    stamp[i] = -1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    q0[0] = source;
    counts[0] = 1;
    counts[1] = 0;
  }
}

/// Report of one bucket selection of the experimental near/far schedule (pinned host memory).
struct sssp_bucket_report_t {
  int count;                   // vertices selected into the near queue
  int pad;
  unsigned long long deg_sum;  // their out-degree sum
  unsigned min_far_bits;       // smallest finite distance >= hi (fp32 bits; 0x7f7fffff = none)
  volatile int seq;
};

struct sssp_scratch_t {
  dbuf_t<int> stamp;
  dbuf_t<unsigned long long> bucket;  // near/far: [0] out-degree sum, [1] low word = min far distance bits
  sssp_bucket_report_t* h_bucket = nullptr;
  dbuf_t<int> q[2];
  dbuf_t<int> counts;
  struct host_fb_t {
    int count;
    int overflow;
    unsigned long long edges;
    unsigned long long deg_sum;
    volatile int seq;
  };
  int seq = 0;
  host_fb_t* h_fb = nullptr;
  tail_report_t* h_tail = nullptr;
  cudaEvent_t ev[128] = {};
  ~sssp_scratch_t() {
    if (h_fb)
      cudaFreeHost(h_fb);
    if (h_tail)
      cudaFreeHost(h_tail);
    if (h_bucket)
      cudaFreeHost(h_bucket);
    for (auto e : ev)
      if (e)
        cudaEventDestroy(e);
  }
  void ensure(int V) {
    stamp.ensure(static_cast<size_t>(V) + 64);
    q[0].ensure(static_cast<size_t>(V) + 64);
    q[1].ensure(static_cast<size_t>(V) + 64);
    counts.ensure(4);
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

static __global__ void sssp_feedback_kernel(const int* count, const ctrl_t* c,
                                     sssp_scratch_t::host_fb_t* fb, int seq) {
  fb->count = *count;
  fb->overflow = c->overflow;
  fb->edges = c->edges;
  fb->deg_sum = c->deg_sum;
  __threadfence_system();
  fb->seq = seq;
}

struct sssp_level_stat_t {
  int frontier;
  unsigned long long edges_relaxed;
  float kernel_ms = 0.0f;
};

/// Returns the number of iterations; `dist` is a device array of V floats.
inline int sssp_run(workspace_t& ws, sssp_scratch_t& sc, const csr_view_t& g, int source,
                    float* dist, const advance_launch_t& cfg,
                    std::vector<sssp_level_stat_t>* levels = nullptr) {
  const int V = g.n_vertices;
  const int sms = device_info_t::get().sm_count;
  sc.ensure(V);
  cudaStream_t st = ws.stream;
  sssp_reset_kernel<<<sms * 8, 256, 0, st>>>(dist, sc.stamp.ptr, V, source, sc.q[0].ptr,
                                              sc.counts.ptr);
  ws.launches += 1;
  int cur = 0, iteration = 0;
  long long n_f = 1;
  unsigned long long m_f = 0;  // out-degree sum of the frontier (0 = unknown, first iteration)
  while (n_f > 0) {
    // tiny frontier: finish (or continue) the run inside one single-CTA launch
    if (iteration > 0 && static_cast<long long>(m_f) < cfg.small_frontier_edges) {
      if (iteration < 64)
        B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration], st));
      advance_tail_kernel<1024, true><<<1, 1024, 0, st>>>(
          g, sc.q[0].ptr, sc.q[1].ptr, sc.counts.ptr, cur, iteration, 16,
          static_cast<unsigned long long>(cfg.small_frontier_edges),
          sssp_relax_maker{dist, sc.stamp.ptr}, sc.h_tail, ++sc.seq);
      if (iteration < 64)
        B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration + 1], st));
      ws.launches += 1;
      wait_for_sequence(&sc.h_tail->seq, sc.seq, st);
      const tail_report_t& t = *sc.h_tail;
      for (int k = 0; k < t.levels; ++k)
        if (levels)
          levels->push_back({t.frontier[k], t.edges[k]});
      for (int k = 1; k < t.levels && iteration + k < 64; ++k) {
        B2G_CHECK(cudaEventRecord(sc.ev[2 * (iteration + k)], st));
        B2G_CHECK(cudaEventRecord(sc.ev[2 * (iteration + k) + 1], st));
      }
      iteration += t.levels;
      cur = t.cur;
      n_f = t.count;
      m_f = t.deg_sum;
      continue;
    }
    int nxt = cur ^ 1;
    B2G_CHECK(cudaMemsetAsync(sc.counts.ptr + nxt, 0, sizeof(int), st));
    sssp_relax_op op{dist, sc.stamp.ptr, iteration};
    ctrl_t* c = nullptr;
    if (iteration < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration], st));
    advance_launch_t lcfg = cfg;
    lcfg.avg_degree = (iteration > 0 && n_f > 0) ? static_cast<double>(m_f) / static_cast<double>(n_f) : 0.0;
    if (iteration == 0) {
      lcfg.lb = lb_t::block_mapped;  // one row of unknown length: binned kernel + hub slabs
    } else if (lcfg.lb == lb_t::merge_path && static_cast<long long>(m_f) < cfg.mid_frontier_edges) {
      lcfg.lb = lb_t::block_mapped;  // mid-size frontier: skip the scan + partition launches
    } else if (static_cast<long long>(m_f) < cfg.small_frontier_edges) {
      lcfg.lb = lb_t::block_mapped;  // single-kernel path for tiny frontiers
      lcfg.hub_threshold = 1 << 30;
    }
    launch_advance<advance_output_t::vertices, true, true>(
        ws, g, sc.q[cur].ptr, sc.counts.ptr + cur, static_cast<int>(n_f < V ? n_f : V),
        sc.q[nxt].ptr, sc.counts.ptr + nxt, V, op, lcfg, &c);
    if (iteration < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration + 1], st));
    sssp_feedback_kernel<<<1, 1, 0, st>>>(sc.counts.ptr + nxt, c, sc.h_fb, ++sc.seq);
    ws.launches += 1;
    wait_for_sequence(&sc.h_fb->seq, sc.seq, st);
    if (sc.h_fb->overflow)
      throw std::runtime_error("sssp: output frontier overflow");
    if (levels)
      levels->push_back({static_cast<int>(n_f), sc.h_fb->edges});
    n_f = sc.h_fb->count;
    m_f = sc.h_fb->deg_sum;
    cur = nxt;
    ++iteration;
  }
  if (levels)
    for (int l = 0; l < iteration && l < 64; ++l)
      cudaEventElapsedTime(&(*levels)[l].kernel_ms, sc.ev[2 * l], sc.ev[2 * l + 1]);
  return iteration;
}

}  // namespace b200
}  // namespace gunrock
