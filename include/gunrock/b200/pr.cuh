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

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>

namespace gunrock {
namespace b200 {

constexpr int kPrPartials = 1024;  // fixed number of dangling-sum partials (deterministic tree)

struct pr_scratch_t {
  dbuf_t<float> plast, iw, c;
  dbuf_t<double> partials;
  dbuf_t<unsigned> err;  // fp32 bit pattern of max |p - plast|
  float* h_err = nullptr;
  ~pr_scratch_t() {
    if (h_err)
      cudaFreeHost(h_err);
  }
  void ensure(int V) {
    plast.ensure(static_cast<size_t>(V) + 16);
    iw.ensure(static_cast<size_t>(V) + 16);
    c.ensure(static_cast<size_t>(V) + 16);
    partials.ensure(kPrPartials);
    err.ensure(4);
    if (!h_err)
      B2G_CHECK(cudaMallocHost(&h_err, sizeof(float)));
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

/// prepare: plast = p, c = plast*iw, and kPrPartials deterministic fp64 partial dangling sums.
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
pr_prepare_kernel(int V, float alpha, const float* __restrict__ p, const float* __restrict__ iw,
                  float* __restrict__ plast, float* __restrict__ c, double* __restrict__ partials) {
  __shared__ double s_red[kThreads / 32];
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
  }
}

/**
 * @brief pull: one warp fetches 32 destination rows; rows >= 32 in-edges are reduced by the whole
 * warp (coalesced index stream, fp64 shuffle reduction), shorter rows by their lane.
 */
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
pr_pull_kernel(csr_view_t t, float alpha, const float* __restrict__ c,
               const float* __restrict__ plast, const double* __restrict__ partials,
               float* __restrict__ p, unsigned* err_bits, ctrl_t* ctrl) {
  const int lane = lane_id();
  const int V = t.n_vertices;
  const int* __restrict__ ro = t.row_offsets;
  const int* __restrict__ ci = t.column_indices;
  const float* __restrict__ vals = t.values;
  // base term, identical in every thread: fixed-order fp64 sum of the partials
  double ds = 0.0;
  for (int i = 0; i < kPrPartials; ++i)
    ds += partials[i];
  const float dsum = static_cast<float>(ds);
  const float base_f = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, alpha), dsum), static_cast<float>(V));
  const double base = static_cast<double>(base_f);
  float err = 0.0f;
  for (;;) {
    int b = 0;
    if (lane == 0)
      b = atomicAdd(&ctrl->work, 32);
    b = __shfl_sync(kFull, b, 0);
    if (b >= V)
      break;
    const int v = b + lane;
    int start = 0, deg = 0;
    if (v < V) {
      start = ro[v];
      deg = ro[v + 1] - start;
    }
    double acc = 0.0;
    bool mine_done = false;
    unsigned big = __ballot_sync(kFull, deg >= 32);
    while (big) {
      int leader = __ffs(big) - 1;
      big &= big - 1;
      int s = __shfl_sync(kFull, start, leader);
      int d = __shfl_sync(kFull, deg, leader);
      double part = 0.0;
      for (int off = lane; off < d; off += 32) {
        int u = ld_stream(ci + s + off);
        float x = c[u];
        if (vals)
          x = __fmul_rn(x, ld_stream(vals + s + off));
        part += static_cast<double>(x);
      }
      part = warp_sum(part);
      if (lane == leader) {
        acc = part;
        mine_done = true;
      }
    }
    if (!mine_done) {
      for (int k = 0; k < deg; ++k) {
        int u = ci[start + k];
        float x = c[u];
        if (vals)
          x = __fmul_rn(x, vals[start + k]);
        acc += static_cast<double>(x);
      }
    }
    if (v < V) {
      float pv = static_cast<float>(base + acc);
      p[v] = pv;
      err = fmaxf(err, fabsf(pv - plast[v]));
    }
  }
  err = warp_max(err);
  if (lane == 0 && err > 0.0f)
    atomicMax(err_bits, __float_as_uint(err));
}

static __global__ void pr_err_feedback_kernel(unsigned* err_bits, float* h_err) {
  *h_err = __uint_as_float(*err_bits);
  *err_bits = 0;
}

/// Returns the iteration count.  g = CSR (out-edges, for iweights), t = CSC (in-edges, pulled).
inline int pr_run(workspace_t& ws, pr_scratch_t& sc, const csr_view_t& g, const csr_view_t& t,
                  float alpha, float tol, int max_iter, float* p) {
  const int V = g.n_vertices;
  const int sms = device_info_t::get().sm_count;
  sc.ensure(V);
  cudaStream_t st = ws.stream;
  pr_reset_kernel<<<sms * 8, 256, 0, st>>>(g, alpha, p, sc.plast.ptr, sc.iw.ptr);
  B2G_CHECK(cudaMemsetAsync(sc.err.ptr, 0, sizeof(unsigned), st));
  ws.launches += 1;
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
    pr_prepare_kernel<256><<<kPrPartials, 256, 0, st>>>(V, alpha, p, sc.iw.ptr, sc.plast.ptr,
                                                        sc.c.ptr, sc.partials.ptr);
    ctrl_t* ctrl = ws.next_ctrl();
    pr_pull_kernel<256><<<sms * 8, 256, 0, st>>>(t, alpha, sc.c.ptr, sc.plast.ptr, sc.partials.ptr,
                                                 p, sc.err.ptr, ctrl);
    ws.launches += 2;
    B2G_CHECK(cudaGetLastError());
    ++iteration;
  }
  return iteration;
}

}  // namespace b200
}  // namespace gunrock
