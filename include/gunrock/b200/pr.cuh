/**
 * @file pr.cuh
 * @brief PageRank power iteration as a deterministic PULL over the transpose (CSC).
 *
 * Path replaced: gunrock::pr::problem_t::reset + enactor_t::loop + is_converged,
 *   include/gunrock/algorithms/pr.hxx:65-93, :107-152, :172-195.
 * The reference spreads along out-edges with one thread per EDGE, a log2(V)-step binary search
 * for the source (graph/csr.hxx:66-81) and a contended fp32 atomicAdd per edge (pr.hxx:140-152),
 * plus four V-sized Thrust passes per iteration (copy_n, transform_reduce, fill_n and the
 * max-abs-diff reduce).  Here one V-sized prepare pass and one pull pass do all of it:
 *   prepare : plast = p; c[u] = plast[u] * iweights[u]; dangling partial sums
 *   pull    : p[v] = base + sum_{u -> v} c[u] * w(u,v); err = max |p - plast|  (no atomics on p)
 * Arithmetic contract (the test checker restates the same one): each product is formed in fp32
 * exactly as the reference lambda does ((plast*iw)*w), a destination's products are accumulated
 * in fp64 together with the base term and rounded to fp32 once; the dangling sum is accumulated
 * in fp64 and rounded once.  Convergence rule unchanged: checked before every iteration from the
 * second on, strict `max|p - plast| < tol`, no iteration cap unless max_iter > 0.
 */
#pragma once

#include <cstdlib>

#include <stdexcept>
#include <vector>

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>

namespace gunrock {
namespace b200 {

constexpr int kPrPartials = 1024;  // fixed number of dangling-sum partials (deterministic tree)
constexpr int kPrTile = 2048;      // in-edges per pull tile (8 KiB of source ids per TMA slab)

struct pr_scratch_t {
  dbuf_t<float> plast, iw, c;
  dbuf_t<double> partials;           // kPrPartials dangling partial sums
  dbuf_t<double> head, tail;         // per-tile partial sums of rows crossing a tile boundary
  dbuf_t<int> first_owned;           // per tile: first row whose first in-edge rank >= tile start
  dbuf_t<int> tail_row;              // per tile: row left incomplete at the tile end, or -1
  dbuf_t<unsigned> err;              // [0] fp32 bits of max|p - plast|, [1] prepare ticket
  dbuf_t<float> base;                // (1 - alpha + dangling) / V of the current iteration
  graph_key_t tiled_for;             // the in-edge graph the tile table was built for
  float* h_err = nullptr;
  cudaEvent_t ev[128] = {};
  ~pr_scratch_t() {
    if (h_err)
      cudaFreeHost(h_err);
    for (auto e : ev)
      if (e)
        cudaEventDestroy(e);
  }
  void ensure(int V, int E) {
    plast.ensure(static_cast<size_t>(V) + 16);
    iw.ensure(static_cast<size_t>(V) + 16);
    c.ensure(static_cast<size_t>(V) + 16);
    partials.ensure(kPrPartials);
    size_t tiles = static_cast<size_t>(E) / kPrTile + 4;
    head.ensure(tiles);
    tail.ensure(tiles);
    first_owned.ensure(tiles);
    tail_row.ensure(tiles);
    err.ensure(4);
    base.ensure(4);
    if (!h_err)
      B2G_CHECK(cudaMallocHost(&h_err, sizeof(float)));
    if (!ev[0])
      for (auto& e : ev)
        B2G_CHECK(cudaEventCreate(&e));
  }
};

/// reset (pr.hxx:65-93): p = (float)(1.0/V), plast = 0, iweights = alpha / rowsum (sequential
/// fp32 sum of the row's weights, as get_weight does) or 0.
static __global__ void pr_reset_kernel(csr_view_t g, float alpha, float* p, float* plast, float* iw) {
  const int V = g.n_vertices;
  const float p0 = static_cast<float>(1.0 / static_cast<double>(V));
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    p[v] = p0;
    plast[v] = 0.0f;
    int s = g.row_offsets[v], e = g.row_offsets[v + 1];
    float val = 0.0f;
    if (g.values) {
      for (int k = s; k < e; ++k)
        val = __fadd_rn(val, g.values[k]);
    } else {
      // a row of 1.0f weights sums exactly to deg while deg <= 2^24; beyond that the sequential
      // fp32 sum saturates at 2^24 exactly like the reference's loop.
      int deg = e - s;
      val = deg <= (1 << 24) ? static_cast<float>(deg) : 16777216.0f;
    }
    iw[v] = val != 0.0f ? __fdiv_rn(alpha, val) : 0.0f;
  }
}

/// Tile table of the CSC (built once per graph): first_owned[t] = smallest row r with
/// offsets[r] >= t * kPrTile, first_owned[ntiles] = V.
static __global__ void pr_tile_table_kernel(const int* __restrict__ offsets, int V, int ntiles,
                                            int* __restrict__ first_owned) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
    if (t == ntiles) {
      first_owned[t] = V;
      continue;
    }
    long long t0 = static_cast<long long>(t) * kPrTile;
    int lo = -1, hi = V;  // offsets[lo] < t0 <= offsets[hi]   (offsets[V] = E >= t0)
    while (hi - lo > 1) {
      int mid = lo + ((hi - lo) >> 1);
      if (offsets[mid] < t0)
        lo = mid;
      else
        hi = mid;
    }
    first_owned[t] = hi;
  }
}

/// prepare: plast = p, c = plast*iw, kPrPartials deterministic fp64 partial dangling sums; the
/// last CTA to finish folds them (fixed order) into base = (1 - alpha + dsum) / V.
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
pr_prepare_kernel(int V, float alpha, const float* __restrict__ p, const float* __restrict__ iw,
                  float* __restrict__ plast, float* __restrict__ c, double* __restrict__ partials,
                  unsigned* ticket, float* base_out) {
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  // contiguous slice per CTA so the partial is independent of the grid's scheduling
  const int per = (V + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(V, lo + per);
  double acc = 0.0;
  for (int v = lo + threadIdx.x; v < hi; v += kThreads) {
    float pv = p[v], w = iw[v];
    plast[v] = pv;
    c[v] = __fmul_rn(pv, w);
    if (w == 0.0f)
      acc += static_cast<double>(__fmul_rn(alpha, pv));
  }
  acc = warp_sum(acc);
  if (lane_id() == 0)
    s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w)
      t += s_red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double t = 0.0;
    const int chunk = (gridDim.x + 31) / 32;
    for (int i = 0; i < chunk; ++i) {
      int k = threadIdx.x * chunk + i;
      if (k < static_cast<int>(gridDim.x))
        t += __ldcg(partials + k);
    }
    t = warp_sum(t);  // xor tree: fixed order
    if (threadIdx.x == 0) {
      const float dsum = static_cast<float>(t);
      *base_out = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, alpha), dsum), static_cast<float>(V));
      *ticket = 0;
    }
  }
}

/**
 * @brief pull, tile version.  The CSC edge array is cut into kPrTile-edge tiles handed out by an
 * atomic ticket; a tile's source ids are staged HBM -> shared memory by ONE thread with
 * cp.async.bulk (TMA engine, SASS UBLKCP; the next tile's slab is in flight while this one is
 * reduced).  Threads gather c[src] (x weight) for the slab into shared memory (8 independent
 * gathers in flight per thread), then the rows that START inside the tile are reduced: short
 * segments by one thread, segments >= 64 by a warp, all in fp64.  Rows that cross a tile end leave
 * a partial (tail of this tile / head of the following ones) that pr_fixup_kernel folds in a
 * fixed order, so the result does not depend on scheduling.
 */
template <int kThreads, bool kWeights, int kMinCtas = 6>
__global__ void __launch_bounds__(kThreads, kMinCtas)
pr_pull_tile_kernel(csr_view_t t, int ntiles, const int* __restrict__ first_owned,
                    const float* __restrict__ c, const float* __restrict__ plast,
                    const float* __restrict__ base_ptr, float* __restrict__ p,
                    double* __restrict__ head, double* __restrict__ tail,
                    int* __restrict__ tail_row, unsigned* err_bits, ctrl_t* ctrl) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kLong = 64;
  constexpr int kMaxLong = kPrTile / kLong + 2;
  __shared__ __align__(16) int s_src[2][kPrTile];
  __shared__ __align__(16) float s_w[kWeights ? 2 : 1][kWeights ? kPrTile : 4];
  __shared__ float s_x[kPrTile];
  __shared__ int s_long_row[kMaxLong];  // row id, or -1 for the head segment
  __shared__ int s_long_lo[kMaxLong], s_long_hi[kMaxLong];
  __shared__ int s_nlong;
  __shared__ int s_ticket[2];
  __shared__ __align__(8) uint64_t s_bar[2];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int V = t.n_vertices, E = t.n_edges;
  const int* __restrict__ ro = t.row_offsets;
  const double base = static_cast<double>(*base_ptr);
  float err = 0.0f;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // The bulk copy moves whole 16-byte groups and never reaches past the end of the arrays (a view over the caller's
  // arrays is not padded): the last tile's trailing E % 4 elements are loaded with plain loads after the wait.
  auto issue = [&](int tile, int buf) {  // thread 0 only
    int t0 = tile * kPrTile;
    int cnt = min(kPrTile, E - t0);
    uint32_t bytes = static_cast<uint32_t>(max(cnt, 0) & ~3) * 4u;
    mbar_expect_tx(&s_bar[buf], kWeights ? 2 * bytes : bytes);  // 0 bytes: a plain arrival
    if (bytes) {
      bulk_g2s(&s_src[buf][0], t.column_indices + t0, bytes, &s_bar[buf]);
      if (kWeights)
        bulk_g2s(&s_w[kWeights ? buf : 0][0], t.values + t0, bytes, &s_bar[buf]);
    }
  };
  if (threadIdx.x == 0) {
    s_ticket[0] = atomicAdd(&ctrl->work, 1);
    if (s_ticket[0] < ntiles)
      issue(s_ticket[0], 0);
  }
  __syncthreads();
  int buf = 0;
  unsigned phase_bits = 0;
  for (;;) {
    const int tile = s_ticket[buf];
    if (tile >= ntiles)
      break;
    if (threadIdx.x == 0) {  // prefetch the next tile's slab
      int nt = atomicAdd(&ctrl->work, 1);
      s_ticket[buf ^ 1] = nt;
      if (nt < ntiles)
        issue(nt, buf ^ 1);
      s_nlong = 0;
    }
    const int t0 = tile * kPrTile;
    const int t1 = min(E, t0 + kPrTile);
    const int cnt = t1 - t0;
    mbar_wait(&s_bar[buf], (phase_bits >> buf) & 1u);
    phase_bits ^= 1u << buf;
    if (cnt & 3) {  // uniform: only the last tile of a graph with E % 4 != 0
      const int k = (cnt & ~3) + static_cast<int>(threadIdx.x);
      if (threadIdx.x < 3 && k < cnt) {
        s_src[buf][k] = t.column_indices[t0 + k];
        if (kWeights)
          s_w[kWeights ? buf : 0][k] = t.values[t0 + k];
      }
      __syncthreads();
    }
    // ---- gather contributions of the slab ------------------------------------------------
#pragma unroll
    for (int j = 0; j < kPrTile / kThreads; ++j) {
      int k = threadIdx.x + j * kThreads;
      if (k < cnt) {
        float x = c[s_src[buf][k]];
        if (kWeights)
          x = __fmul_rn(x, s_w[kWeights ? buf : 0][k]);
        s_x[k] = x;
      }
    }
    __syncthreads();
    // ---- rows owned by this tile ------------------------------------------------------------
    const int lo = first_owned[tile];
    const int hi = (tile == ntiles - 1) ? V : first_owned[tile + 1];
    const int first_start = lo < V ? ro[lo] : E;
    if (threadIdx.x == 0) {
      tail_row[tile] = -1;
      if (first_start > t0) {  // row lo-1 started earlier and reaches into this tile: head segment
        int slot = atomicAdd(&s_nlong, 1);
        s_long_row[slot] = -1;
        s_long_lo[slot] = 0;
        s_long_hi[slot] = min(first_start, t1) - t0;
      }
    }
    for (int r = lo + threadIdx.x; r < hi; r += kThreads) {
      const int s = ro[r], e = ro[r + 1];
      const int seg_hi = min(e, t1) - t0, seg_lo = s - t0;
      if (seg_hi - seg_lo >= kLong || e > t1) {  // long, or crosses the tile end (tail)
        int slot = atomicAdd(&s_nlong, 1);
        s_long_row[slot] = r;
        s_long_lo[slot] = seg_lo;
        s_long_hi[slot] = seg_hi;
      } else {
        double acc = 0.0;
        for (int k = seg_lo; k < seg_hi; ++k)
          acc += static_cast<double>(s_x[k]);
        float pv = static_cast<float>(base + acc);
        p[r] = pv;
        err = fmaxf(err, fabsf(pv - plast[r]));
      }
    }
    __syncthreads();
    const int nlong = s_nlong;
    for (int i = warp; i < nlong; i += kWarps) {
      const int r = s_long_row[i], a = s_long_lo[i], b = s_long_hi[i];
      double acc = 0.0;
      for (int k = a + lane; k < b; k += 32)
        acc += static_cast<double>(s_x[k]);
      acc = warp_sum(acc);
      if (lane == 0) {
        if (r < 0) {
          head[tile] = acc;
        } else if (ro[r + 1] > t1) {
          tail[tile] = acc;
          tail_row[tile] = r;
        } else {
          float pv = static_cast<float>(base + acc);
          p[r] = pv;
          err = fmaxf(err, fabsf(pv - plast[r]));
        }
      }
    }
    __syncthreads();  // s_x, s_long_*, s_src[buf] are free again
    buf ^= 1;
  }
  err = warp_max(err);
  if (lane == 0 && err > 0.0f)
    atomicMax(err_bits, __float_as_uint(err));
}

/// Fold the partials of rows that span several tiles: total = tail(t) + sum of head(t+1 .. tB),
/// lanes strided over the following tiles, fixed xor-tree order.  One warp per tile.
static __global__ void pr_fixup_kernel(csr_view_t t, int ntiles, const int* __restrict__ tail_row,
                                       const double* __restrict__ head,
                                       const double* __restrict__ tail,
                                       const float* __restrict__ base_ptr,
                                       const float* __restrict__ plast, float* __restrict__ p,
                                       unsigned* err_bits) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const double base = static_cast<double>(*base_ptr);
  float err = 0.0f;
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += warps) {
    const int r = tail_row[tile];
    if (r < 0)
      continue;
    const int last_tile = (t.row_offsets[r + 1] - 1) / kPrTile;
    double acc = 0.0;
    for (int k = tile + 1 + lane; k <= last_tile; k += 32)
      acc += head[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      float pv = static_cast<float>(base + (tail[tile] + acc));
      p[r] = pv;
      err = fmaxf(err, fabsf(pv - plast[r]));
    }
  }
  if (lane == 0 && err > 0.0f)
    atomicMax(err_bits, __float_as_uint(err));
}

// ---------------------------------------------------------------------------------------------
// Partitioned (multi-GPU) PageRank: the rank owns the destination vertices v % P == part with their
// in-edges; per iteration the host side all-gathers c = plast * iweights (4 V bytes in total),
// all-reduces the dangling sum (fp64) and the error (max), SURVEY.md section 8e.  The pull itself is
// the same tile kernel: the column ids are remapped once to positions in the gathered array.
// ---------------------------------------------------------------------------------------------
/// outdeg[src] += 1 for every local in-edge (global array, all-reduced by the host afterwards).
static __global__ void part_pr_outdeg_kernel(const int* __restrict__ src_ids, int n_edges, int* outdeg) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x)
    atomicAdd(outdeg + src_ids[e], 1);
}
/// outweight[src] += w(e) for every local in-edge, in fp64 (global array, all-reduced afterwards): the weighted
/// graph's row sums.  The single-GPU reset sums a row's weights sequentially in fp32 as the reference's loop does
/// (pr.hxx:65-93); a vertex's out-edges are spread over the ranks here, so the sum is taken in fp64 and rounded
/// once -- it can differ from the sequential fp32 sum in the last place, well inside PageRank's 1e-6 tolerance.
static __global__ void part_pr_outweight_kernel(const int* __restrict__ src_ids, const float* __restrict__ w,
                                                int n_edges, double* outweight) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x)
    atomicAdd(outweight + src_ids[e], static_cast<double>(w[e]));
}
/// position of global vertex u in the rank-major gathered array: (u % P) * R + u / P
static __global__ void part_pr_remap_kernel(const int* __restrict__ src_ids, int n_edges, int nparts,
                                            int rows_per_rank, int* __restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    int u = src_ids[e];
    out[e] = (u % nparts) * rows_per_rank + u / nparts;
  }
}
/// p = 1/V, plast = 0, iweights of the owned vertices from the global out-degrees.
static __global__ void part_pr_reset_kernel(int n_local, int nparts, int part, int n_global, float alpha,
                                            const int* __restrict__ outdeg, float* p, float* plast,
                                            float* iw) {
  const float p0 = static_cast<float>(1.0 / static_cast<double>(n_global));
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_local; l += gridDim.x * blockDim.x) {
    p[l] = p0;
    plast[l] = 0.0f;
    int d = outdeg[l * nparts + part];
    float val = d <= (1 << 24) ? static_cast<float>(d) : 16777216.0f;
    iw[l] = val != 0.0f ? __fdiv_rn(alpha, val) : 0.0f;
  }
}
/// The weighted form: iweights from the global fp64 row sums of the weights.
static __global__ void part_pr_reset_weighted_kernel(int n_local, int nparts, int part, int n_global, float alpha,
                                                     const double* __restrict__ outweight, float* p, float* plast,
                                                     float* iw) {
  const float p0 = static_cast<float>(1.0 / static_cast<double>(n_global));
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_local; l += gridDim.x * blockDim.x) {
    p[l] = p0;
    plast[l] = 0.0f;
    const float val = static_cast<float>(outweight[l * nparts + part]);
    iw[l] = val != 0.0f ? __fdiv_rn(alpha, val) : 0.0f;
  }
}
/// plast = p, c = plast*iw, this rank's fp64 dangling partial (last CTA folds the per-CTA partials).
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
part_pr_prepare_kernel(int n_local, float alpha, const float* __restrict__ p,
                       const float* __restrict__ iw, float* __restrict__ plast,
                       float* __restrict__ c, double* __restrict__ partials, unsigned* ticket,
                       double* dsum_out) {
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  const int per = (n_local + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n_local, lo + per);
  double acc = 0.0;
  for (int v = lo + threadIdx.x; v < hi; v += kThreads) {
    float pv = p[v], w = iw[v];
    plast[v] = pv;
    c[v] = __fmul_rn(pv, w);
    if (w == 0.0f)
      acc += static_cast<double>(__fmul_rn(alpha, pv));
  }
  acc = warp_sum(acc);
  if (lane_id() == 0)
    s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w)
      t += s_red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double t = 0.0;
    const int chunk = (gridDim.x + 31) / 32;
    for (int i = 0; i < chunk; ++i) {
      int k = threadIdx.x * chunk + i;
      if (k < static_cast<int>(gridDim.x))
        t += __ldcg(partials + k);
    }
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      *dsum_out = t;
      *ticket = 0;
    }
  }
}
/// base = (1 - alpha + (float)dsum_global) / V ; also publishes the previous error bits and clears them.
static __global__ void part_pr_base_kernel(const double* dsum_global, float alpha, int n_global,
                                           float* base_out) {
  const float dsum = static_cast<float>(*dsum_global);
  *base_out = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, alpha), dsum), static_cast<float>(n_global));
}
static __global__ void part_pr_err_kernel(unsigned* err_bits, float* err_out) {
  *err_out = __uint_as_float(*err_bits);
  *err_bits = 0;
}

/// Per-rank state of a partitioned PageRank.
struct part_pr_state_t {
  int nparts = 1, part = 0, n_global = 0, n_local = 0, rows_per_rank = 0;
  dbuf_t<int> outdeg;       // global out-degrees (after the all-reduce); the C++ NCCL loop's copy
  dbuf_t<double> outweight; // weighted graphs: global row sums of the weights (same)
  dbuf_t<float> c_local, c_all;  // the C++ NCCL loop's gather buffers (b2g_part_pr_nccl)
  dbuf_t<int> remapped;     // column ids -> positions in the gathered c array
  dbuf_t<float> p;          // owned ranks
  dbuf_t<double> dsum;      // [0] local dangling partial (in/out of the all-reduce)
  dbuf_t<float> err;        // [0] local max |p - plast| (in/out of the all-reduce)
  pr_scratch_t sc;
  csr_view_t t;             // local in-edge rows with remapped columns
  ctrl_t* ctrl = nullptr;
};

// ---- the per-rank steps of one partitioned PageRank (enqueued on ws.stream; the collectives between them belong
// to the caller: C ABI b2g_part_pr_* + torch.distributed, or part_pr_run in part_loops.cuh) -------------------------
/// Per-graph setup (column remap + tile table, cached on the row-offsets pointer) and reset of p / plast / iweights
/// from the GLOBAL out-degrees or, for a graph with values, the global fp64 row sums of the weights (exactly one
/// of the two is non-null).  `view` = the rank's in-edge rows with global source ids.
template <typename partition_type>  // b200::partition_t (bfs_partitioned.cuh); a template keeps this header light
inline void part_pr_begin(workspace_t& ws, const csr_view_t& view, const partition_type& pt, part_pr_state_t& S,
                          float alpha, const int* outdeg_global, const double* outweight_global) {
  cudaStream_t st = ws.stream;
  const int sms = device_info_t::get().sm_count;
  S.nparts = pt.nparts;
  S.part = pt.part;
  S.n_global = pt.n_global;
  S.n_local = pt.n_local;
  S.rows_per_rank = pt.rows_of(0);
  S.sc.ensure(S.n_local, view.n_edges);
  S.p.ensure(static_cast<size_t>(S.n_local) + 16);
  S.dsum.ensure(2);
  S.err.ensure(2);
  if (S.remapped.cap < static_cast<size_t>(view.n_edges) + 16 || S.t.row_offsets != view.row_offsets) {
    S.remapped.ensure(static_cast<size_t>(view.n_edges) + 16);
    if (view.n_edges)
      part_pr_remap_kernel<<<sms * 8, 256, 0, st>>>(view.column_indices, view.n_edges, S.nparts, S.rows_per_rank,
                                                    S.remapped.ptr);
    S.t.n_vertices = S.n_local;
    S.t.n_edges = view.n_edges;
    S.t.row_offsets = view.row_offsets;
    S.t.column_indices = S.remapped.ptr;
    S.t.uid = next_graph_uid();
    const int ntiles = view.n_edges > 0 ? (view.n_edges + kPrTile - 1) / kPrTile : 1;
    pr_tile_table_kernel<<<sms * 2, 256, 0, st>>>(S.t.row_offsets, S.n_local, ntiles, S.sc.first_owned.ptr);
    S.sc.tiled_for.set(S.t);
    ws.launches += 2;
  }
  S.t.values = outweight_global ? view.values : nullptr;  // the remap keeps the edge order
  if (outweight_global)
    part_pr_reset_weighted_kernel<<<sms * 8, 256, 0, st>>>(S.n_local, S.nparts, S.part, S.n_global, alpha,
                                                           outweight_global, S.p.ptr, S.sc.plast.ptr, S.sc.iw.ptr);
  else
    part_pr_reset_kernel<<<sms * 8, 256, 0, st>>>(S.n_local, S.nparts, S.part, S.n_global, alpha, outdeg_global,
                                                  S.p.ptr, S.sc.plast.ptr, S.sc.iw.ptr);
  B2G_CHECK(cudaMemsetAsync(S.sc.err.ptr, 0, 2 * sizeof(unsigned), st));
  ws.launches += 1;
  B2G_CHECK(cudaGetLastError());
}

/// plast = p, c_local = plast * iweights (rows_per_rank floats, zero padded), dsum_local = the dangling partial.
inline void part_pr_prepare(workspace_t& ws, part_pr_state_t& S, float alpha, float* c_local, double* dsum_local) {
  cudaStream_t st = ws.stream;
  if (S.rows_per_rank > S.n_local)  // zero the padding slot(s) of the gathered layout
    B2G_CHECK(cudaMemsetAsync(c_local + S.n_local, 0, sizeof(float) * (S.rows_per_rank - S.n_local), st));
  part_pr_prepare_kernel<256><<<kPrPartials, 256, 0, st>>>(S.n_local, alpha, S.p.ptr, S.sc.iw.ptr, S.sc.plast.ptr,
                                                           c_local, S.sc.partials.ptr, S.sc.err.ptr + 1, dsum_local);
  ws.launches += 1;
  B2G_CHECK(cudaGetLastError());
}

/// One pull over the owned rows from the gathered c and the global dangling sum; err_local = max |p - plast|.
inline void part_pr_pull(workspace_t& ws, part_pr_state_t& S, float alpha, const float* c_all,
                         const double* dsum_global, float* err_local) {
  cudaStream_t st = ws.stream;
  const int sms = device_info_t::get().sm_count;
  const int ntiles = S.t.n_edges > 0 ? (S.t.n_edges + kPrTile - 1) / kPrTile : 1;
  part_pr_base_kernel<<<1, 1, 0, st>>>(dsum_global, alpha, S.n_global, S.sc.base.ptr);
  ctrl_t* ctrl = ws.next_ctrl();
  if (S.t.values)
    pr_pull_tile_kernel<256, true><<<sms * 8, 256, 0, st>>>(S.t, ntiles, S.sc.first_owned.ptr, c_all, S.sc.plast.ptr,
                                                            S.sc.base.ptr, S.p.ptr, S.sc.head.ptr, S.sc.tail.ptr,
                                                            S.sc.tail_row.ptr, S.sc.err.ptr, ctrl);
  else
    pr_pull_tile_kernel<256, false><<<sms * 8, 256, 0, st>>>(S.t, ntiles, S.sc.first_owned.ptr, c_all, S.sc.plast.ptr,
                                                             S.sc.base.ptr, S.p.ptr, S.sc.head.ptr, S.sc.tail.ptr,
                                                             S.sc.tail_row.ptr, S.sc.err.ptr, ctrl);
  pr_fixup_kernel<<<sms, 256, 0, st>>>(S.t, ntiles, S.sc.tail_row.ptr, S.sc.head.ptr, S.sc.tail.ptr, S.sc.base.ptr,
                                       S.sc.plast.ptr, S.p.ptr, S.sc.err.ptr);
  part_pr_err_kernel<<<1, 1, 0, st>>>(S.sc.err.ptr, err_local);
  ws.launches += 4;
  B2G_CHECK(cudaGetLastError());
}

static __global__ void pr_err_feedback_kernel(unsigned* err_bits, float* h_err) {
  *h_err = __uint_as_float(*err_bits);
  *err_bits = 0;
}

struct pr_iter_stat_t {
  float kernel_ms = 0.0f;  // prepare + pull + fixup of one iteration
};

/// Returns the iteration count.  g = CSR (out-edges, for iweights), t = CSC (in-edges, pulled).
inline int pr_run(workspace_t& ws, pr_scratch_t& sc, const csr_view_t& g, const csr_view_t& t,
                  float alpha, float tol, int max_iter, float* p,
                  std::vector<pr_iter_stat_t>* iters_out = nullptr) {
  const int V = g.n_vertices, E = t.n_edges;
  const int sms = device_info_t::get().sm_count;
  sc.ensure(V, E);
  cudaStream_t st = ws.stream;
  const int ntiles = E > 0 ? (E + kPrTile - 1) / kPrTile : 1;
  if (!sc.tiled_for.matches(t)) {  // per-graph table (ingest-like, not per iteration); keyed on identity
    pr_tile_table_kernel<<<sms * 2, 256, 0, st>>>(t.row_offsets, V, ntiles, sc.first_owned.ptr);
    sc.tiled_for.set(t);
    ws.launches += 1;
  }
  pr_reset_kernel<<<sms * 8, 256, 0, st>>>(g, alpha, p, sc.plast.ptr, sc.iw.ptr);
  B2G_CHECK(cudaMemsetAsync(sc.err.ptr, 0, 2 * sizeof(unsigned), st));
  ws.launches += 1;
  const bool tma_ok = (reinterpret_cast<uintptr_t>(t.column_indices) & 15u) == 0 &&
                      (!t.values || (reinterpret_cast<uintptr_t>(t.values) & 15u) == 0);
  if (!tma_ok)
    throw std::runtime_error("pagerank pull needs 16-byte aligned CSC arrays (TMA slabs)");
  int iteration = 0;
  for (;;) {
    if (iteration > 0) {
      pr_err_feedback_kernel<<<1, 1, 0, st>>>(sc.err.ptr, sc.h_err);
      ws.launches += 1;
      B2G_CHECK(cudaStreamSynchronize(st));
      if (*sc.h_err < tol)
        break;
    }
    if (max_iter > 0 && iteration >= max_iter)
      break;
    if (iteration < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration], st));
    pr_prepare_kernel<256><<<kPrPartials, 256, 0, st>>>(V, alpha, p, sc.iw.ptr, sc.plast.ptr,
                                                        sc.c.ptr, sc.partials.ptr, sc.err.ptr + 1,
                                                        sc.base.ptr);
    ctrl_t* ctrl = ws.next_ctrl();
    const int grid = sms * 8;  // tickets are dynamic; residency is capped by the hardware
    static const int full_occupancy = [] {  // A/B: B2G_PR_CTAS=8 caps the unweighted kernel at 32 registers
      const char* e = std::getenv("B2G_PR_CTAS");
      return e ? std::atoi(e) : 0;
    }();
    if (t.values)
      pr_pull_tile_kernel<256, true><<<grid, 256, 0, st>>>(
          t, ntiles, sc.first_owned.ptr, sc.c.ptr, sc.plast.ptr, sc.base.ptr, p, sc.head.ptr,
          sc.tail.ptr, sc.tail_row.ptr, sc.err.ptr, ctrl);
    else if (full_occupancy == 8)
      pr_pull_tile_kernel<256, false, 8><<<grid, 256, 0, st>>>(
          t, ntiles, sc.first_owned.ptr, sc.c.ptr, sc.plast.ptr, sc.base.ptr, p, sc.head.ptr,
          sc.tail.ptr, sc.tail_row.ptr, sc.err.ptr, ctrl);
    else
      pr_pull_tile_kernel<256, false><<<grid, 256, 0, st>>>(
          t, ntiles, sc.first_owned.ptr, sc.c.ptr, sc.plast.ptr, sc.base.ptr, p, sc.head.ptr,
          sc.tail.ptr, sc.tail_row.ptr, sc.err.ptr, ctrl);
    pr_fixup_kernel<<<sms, 256, 0, st>>>(t, ntiles, sc.tail_row.ptr, sc.head.ptr, sc.tail.ptr,
                                         sc.base.ptr, sc.plast.ptr, p, sc.err.ptr);
    if (iteration < 64)
      B2G_CHECK(cudaEventRecord(sc.ev[2 * iteration + 1], st));
    ws.launches += 3;
    B2G_CHECK(cudaGetLastError());
    ++iteration;
  }
  if (iters_out) {
    iters_out->resize(iteration);
    for (int l = 0; l < iteration && l < 64; ++l)
      cudaEventElapsedTime(&(*iters_out)[l].kernel_ms, sc.ev[2 * l], sc.ev[2 * l + 1]);
  }
  return iteration;
}

}  // namespace b200
}  // namespace gunrock
