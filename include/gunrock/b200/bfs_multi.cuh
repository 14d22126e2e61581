/**
 * @file bfs_multi.cuh
 * @brief Direction-optimised BFS over SEVERAL devices of one process: what `bfs::run` does when it is handed
 * a `gcuda::multi_context_t` with more than one context.
 *
 * The reference declares the surface -- `multi_context_t(thrust::host_vector<device_id_t>)`, `size()`,
 * `get_context(i)`, `enable_peer_access()` (include/gunrock/cuda/context.hxx:146-216) -- and then throws in every
 * operator for `size() != 1` (advance.hxx:129-132).  Here the same object runs the partitioned
 * traversal of SURVEY.md 8e:
 *   1. the caller's device CSR is cut 1-D (cyclic vertex cut: vertex v -> rank v % P, row v / P) by two kernels
 *      on the first context's device that write each rank's rows straight into that rank's memory (peer
 *      stores); cached per graph identity, like the transpose;
 *   2. every rank owns a peer-memory window (bfs_p2p.cuh); peers are plain pointers inside one process;
 *   3. one host thread per device runs `part_bfs_p2p_run` on its context's stream -- the frontier exchange is
 *      done by the kernels over NVLink, the host reads one pinned record per level;
 *   4. a gather kernel interleaves the ranks' slices into the caller's `distances` (peer loads).
 * The one-process-per-GPU form of the same loop (C ABI `b2g_part_bfs_p2p`, CUDA IPC windows, torchrun) is
 * what `bench.py --gpus N` times.
 */
#pragma once

#include <cstdlib>
#include <exception>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <gunrock/b200/bfs_p2p.cuh>
#if defined(__has_include)
#if __has_include(<nccl.h>)
#include <gunrock/b200/bfs_nccl.cuh>  // the NCCL exchange (B2G_EXCHANGE=nccl); bound with dlopen, no link dependency
#define GUNROCK_B200_HAS_NCCL 1
#endif
#endif

namespace gunrock {
namespace b200 {

/// One warp per local row: copy the row's column indices (and values) into the rank's arrays.
static __global__ void part_gather_rows_kernel(const int* __restrict__ ro, const int* __restrict__ ci,
                                               const float* __restrict__ vals, int nparts, int part, int n_local,
                                               const int* __restrict__ lro, int* __restrict__ lci,
                                               float* __restrict__ lvals) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_local; i += warps) {
    const int g = i * nparts + part;
    const int s = ro[g], e = ro[g + 1], d = lro[i];
    for (int k = lane; k < e - s; k += 32) {
      lci[d + k] = ci[s + k];
      if (lvals)
        lvals[d + k] = vals[s + k];
    }
  }
}

struct gather_table_t {
  const int* part[kMaxPeers];
};
/// distances[v] = slice[v % P][v / P]
static __global__ void part_interleave_kernel(gather_table_t t, int nparts, int n_global, int* __restrict__ out) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_global; v += gridDim.x * blockDim.x)
    out[v] = t.part[v % nparts][v / nparts];
}

/// One rank's share: its rows, its traversal state, its window.
struct multi_rank_t {
  int device = 0;
  workspace_t* ws = nullptr;  // the rank's context's workspace (bound to that context's stream)
  dbuf_t<int> ro, ci, in_ro, in_ci;
  csr_view_t view, in_view;
  partition_t pt;
  part_bfs_state_t S;
  dbuf_t<unsigned long long> part_deg;
  p2p_state_t P;
#ifdef GUNROCK_B200_HAS_NCCL
  nccl_state_t N;
#endif
  ~multi_rank_t() { P.release(); }
};

struct multi_bfs_cache_t {
  graph_key_t key, in_key;
  std::vector<int> devices;
  std::vector<std::unique_ptr<multi_rank_t>> ranks;
  long long total_edges = 0;
  bool use_nccl = false;  // B2G_EXCHANGE=nccl: NCCL collectives instead of the kernels' peer-memory exchange
  bool matches(const csr_view_t& g, const csr_view_t& in_g, const std::vector<int>& devs) const {
    return !ranks.empty() && key.matches(g) && devices == devs &&
           (in_g.row_offsets == nullptr ? in_key.uid == 0 : (in_g.row_offsets == g.row_offsets || in_key.matches(in_g)));
  }
};

/// Cut `g` (resident on the current device) into rank `r`'s rows, written into `ro_buf` / `ci_buf` that live on
/// the rank's device (allocated here after the row scan told the edge count).
inline csr_view_t partition_rows_to(workspace_t& ws0, const csr_view_t& g, const partition_t& pt, int rank_device,
                                    int home_device, dbuf_t<int>& ro_buf, dbuf_t<int>& ci_buf,
                                    dbuf_t<float>* vals_buf = nullptr) {
  const int n_local = pt.n_local, P = pt.nparts, part = pt.part;
  B2G_CHECK(cudaSetDevice(rank_device));
  int* lro = ro_buf.ensure(static_cast<size_t>(n_local) + 1 + 16);
  B2G_CHECK(cudaSetDevice(home_device));
  const int* ro = g.row_offsets;
  auto value = [=] __device__(int i) -> int {
    const int v = i * P + part;
    return ro[v + 1] - ro[v];
  };
  auto emit = [=] __device__(int i, int excl, int) { lro[i] = excl; };
  if (n_local > 0)
    lookback_scan(ws0, nullptr, n_local, n_local, value, emit, nullptr, lro);
  else
    B2G_CHECK(cudaMemsetAsync(lro, 0, sizeof(int), ws0.stream));
  int n_edges = 0;
  B2G_CHECK(cudaMemcpyAsync(&n_edges, lro + n_local, sizeof(int), cudaMemcpyDeviceToHost, ws0.stream));
  B2G_CHECK(cudaStreamSynchronize(ws0.stream));
  B2G_CHECK(cudaSetDevice(rank_device));
  int* lci = ci_buf.ensure(static_cast<size_t>(n_edges) + 16 + 64);  // + 64 B: TMA slabs may over-read the tail
  float* lvals = (vals_buf && g.values) ? vals_buf->ensure(static_cast<size_t>(n_edges) + 16 + 64) : nullptr;
  B2G_CHECK(cudaSetDevice(home_device));
  if (n_local > 0 && n_edges > 0) {
    const int sms = device_info_t::get().sm_count;
    part_gather_rows_kernel<<<sms * 8, 256, 0, ws0.stream>>>(g.row_offsets, g.column_indices, lvals ? g.values : nullptr,
                                                             P, part, n_local, lro, lci, lvals);
    ws0.launches += 1;
  }
  B2G_CHECK(cudaStreamSynchronize(ws0.stream));
  csr_view_t v;
  v.n_vertices = n_local;
  v.n_edges = n_edges;
  v.row_offsets = lro;
  v.column_indices = lci;
  v.values = lvals;
  v.uid = next_graph_uid();
  return v;
}

/**
 * @brief Ingest step of the multi-device BFS (untimed, like the transpose): cut the graph across the devices
 * of `contexts` and create + connect the ranks' peer-memory windows.  Cached per graph identity and device
 * list; a second call with the same graph returns immediately.
 */
template <typename multi_context_type>
inline void bfs_prepare_multi(multi_context_type& contexts, multi_bfs_cache_t& cache, const csr_view_t& g,
                              const csr_view_t& in_g) {
  const int P = static_cast<int>(contexts.size());
  if (P > kMaxPeers)
    throw std::runtime_error("bfs over a multi_context_t: at most 16 devices");
  std::vector<int> devs(P);
  for (int r = 0; r < P; ++r)
    devs[r] = contexts.contexts[r]->ordinal();
  const int home = devs[0];
  auto* ctx0 = contexts.get_context(0);  // also makes the home device current
  workspace_t& ws0 = ctx0->workspace();
  // ---- 1. partition (ingest, cached per graph identity + device list) ------------------------------------
  if (cache.matches(g, in_g, devs))
    return;
  {
    cache.ranks.clear();
    bool distinct = false;
    for (int r = 1; r < P; ++r)
      distinct = distinct || devs[r] != home;
    if (distinct)
      contexts.enable_peer_access();  // the partition kernels, the windows and the gather use peer pointers
    const bool own_in = in_g.row_offsets != nullptr && in_g.row_offsets != g.row_offsets;
    for (int r = 0; r < P; ++r) {
      std::unique_ptr<multi_rank_t> R(new multi_rank_t());
      R->device = devs[r];
      R->ws = &contexts.contexts[r]->workspace();
      R->pt = partition_t::make(g.n_vertices, P, r);
      R->view = partition_rows_to(ws0, g, R->pt, devs[r], home, R->ro, R->ci);
      if (own_in)
        R->in_view = partition_rows_to(ws0, in_g, R->pt, devs[r], home, R->in_ro, R->in_ci);
      else if (in_g.row_offsets)
        R->in_view = R->view;  // symmetric graph: the local CSR doubles as the local CSC
      cache.ranks.push_back(std::move(R));
    }
    // ---- 2. windows ------------------------------------------------------------------------------------
    void* windows[kMaxPeers] = {};
    for (int r = 0; r < P; ++r) {
      auto& R = *cache.ranks[r];
      B2G_CHECK(cudaSetDevice(R.device));
      part_p2p_prepare(*R.ws, R.view, R.pt, R.S, R.part_deg, R.P, R.in_view.row_offsets != nullptr);
      windows[r] = R.P.own;
    }
    for (int r = 0; r < P; ++r)
      part_p2p_attach_pointers(cache.ranks[r]->P, windows);
    const char* ex = std::getenv("B2G_EXCHANGE");
    cache.use_nccl = ex != nullptr && std::string(ex) == "nccl";
    if (cache.use_nccl) {
#ifdef GUNROCK_B200_HAS_NCCL
      if (!distinct && P > 1)
        throw std::runtime_error("B2G_EXCHANGE=nccl needs one distinct device per context (ncclCommInitAll)");
      const nccl_api_t& nccl = nccl_api_t::get();
      std::vector<ncclComm_t> comms(P);
      nccl.check(nccl.CommInitAll(comms.data(), P, devs.data()), "ncclCommInitAll");
      for (int r = 0; r < P; ++r) {
        auto& R = *cache.ranks[r];
        B2G_CHECK(cudaSetDevice(R.device));
        R.N.comm = comms[r];
        R.N.owns_comm = true;
        R.N.prepare(*R.ws, R.view, R.pt, R.S, R.part_deg, R.in_view.row_offsets != nullptr);
        B2G_CHECK(cudaDeviceSynchronize());
      }
#else
      throw std::runtime_error("B2G_EXCHANGE=nccl: this translation unit was compiled without <nccl.h>");
#endif
    }
    B2G_CHECK(cudaSetDevice(home));
    cache.key.set(g);
    if (own_in)
      cache.in_key.set(in_g);
    else
      cache.in_key.clear();
    cache.devices = devs;
    cache.total_edges = g.n_edges;
  }

}

/**
 * @brief BFS from `source` over the devices of `contexts` (anything with `size()`, `contexts[i]` and
 * `get_context(i)` returning an object with `ordinal()`, `stream()`, `workspace()`, i.e.
 * gcuda::multi_context_t).  `g` / `in_g` are views of the caller's graph resident on the FIRST context's
 * device; `distances` (V ints) lives there too.  Returns the number of levels; `report` receives the global
 * per-level statistics.
 */
template <typename multi_context_type>
inline int bfs_run_multi(multi_context_type& contexts, multi_bfs_cache_t& cache, const csr_view_t& g,
                         const csr_view_t& in_g, int source, int* distances, const part_bfs_config_t& cfg,
                         part_bfs_report_t* report) {
  if (source < 0 || source >= g.n_vertices)
    throw std::runtime_error("bfs: source out of range");
  bfs_prepare_multi(contexts, cache, g, in_g);
  const int P = static_cast<int>(contexts.size());
  const int home = contexts.contexts[0]->ordinal();
  auto* ctx0 = contexts.get_context(0);
  workspace_t& ws0 = ctx0->workspace();

  // ---- 3. one host thread per device runs its rank's level loop ------------------------------------------
  std::vector<std::thread> threads;
  std::vector<std::exception_ptr> errors(P);
  std::vector<part_bfs_report_t> reports(P);
  for (int r = 0; r < P; ++r) {
    threads.emplace_back([&, r]() {
      try {
        auto& R = *cache.ranks[r];
        B2G_CHECK(cudaSetDevice(R.device));
        csr_view_t in_view = cfg.direction != 0 ? R.in_view : csr_view_t();
#ifdef GUNROCK_B200_HAS_NCCL
        if (cache.use_nccl) {
          part_bfs_nccl_run(*R.ws, R.view, in_view, R.pt, R.S, R.part_deg, R.N, source, cache.total_edges, cfg,
                            &reports[r]);
          return;
        }
#endif
        part_bfs_p2p_run(*R.ws, R.view, in_view, R.pt, R.S, R.part_deg, R.P, source, cache.total_edges, cfg,
                         &reports[r]);
      } catch (...) {
        errors[r] = std::current_exception();
      }
    });
  }
  for (auto& t : threads)
    t.join();
  B2G_CHECK(cudaSetDevice(home));
  for (auto& e : errors)
    if (e)
      std::rethrow_exception(e);

  // ---- 4. gather the ranks' slices into the caller's array (peer loads from the home device) -------------
  gather_table_t table{};
  for (int r = 0; r < P; ++r)
    table.part[r] = cache.ranks[r]->S.dist.ptr;
  const int sms = device_info_t::get().sm_count;
  part_interleave_kernel<<<sms * 8, 256, 0, ws0.stream>>>(table, P, g.n_vertices, distances);
  ws0.launches += 1;
  B2G_CHECK(cudaGetLastError());
  if (report)
    *report = reports[0];
  return reports[0].levels;
}

}  // namespace b200
}  // namespace gunrock
