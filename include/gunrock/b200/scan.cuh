/**
 * @file scan.cuh
 * @brief Single-pass device-wide exclusive scan / stable select with decoupled look-back.
 *
 * Replaces the reference's Thrust calls on the hot path:
 *   thrust::transform_exclusive_scan  include/gunrock/framework/operators/advance/helpers.hxx:70-79
 *   thrust::transform_reduce          include/gunrock/framework/operators/advance/helpers.hxx:150-158
 *   thrust::copy_if / remove_copy_if  include/gunrock/framework/operators/filter/predicated.hxx:30,
 *                                     filter/remove.hxx:28
 * Differences by design: the element count is read from device memory (no host round trip, the
 * reference blocks on a D2H copy of the scan total, helpers.hxx:106-110), tiles are handed out
 * by an atomic ticket so a persistent grid of (#SM x k) CTAs can run any size, and the tile
 * status words carry a launch epoch so the status array never needs clearing.
 */
#pragma once

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>

namespace gunrock {
namespace b200 {

namespace scan_detail {
constexpr unsigned kInvalid = 0, kAggregate = 1, kPrefix = 2;
__host__ __device__ __forceinline__ unsigned long long pack(unsigned epoch, unsigned flag, int v) {
  return (static_cast<unsigned long long>((epoch << 2) | flag) << 32) |
         static_cast<unsigned>(v);
}
}  // namespace scan_detail

/**
 * @brief value(i) -> int is evaluated exactly once per i in [0, n); emit(i, exclusive_prefix,
 * value) is then called for every i.  total_out (optional) receives the grand total.
 * n is *n_ptr if n_ptr != nullptr, else n_fixed.
 */
template <int kThreads, int kItems, typename ValueF, typename EmitF>
__global__ void __launch_bounds__(kThreads)
lookback_scan_kernel(const int* __restrict__ n_ptr,
                     int n_fixed,
                     ValueF value,
                     EmitF emit,
                     int* total_out,
                     int* total_at_n,  // optional array: total_at_n[n] = grand total
                     unsigned long long* state,
                     ctrl_t* ctrl,
                     unsigned epoch) {
  using namespace scan_detail;
  constexpr int kTile = kThreads * kItems;
  constexpr int kWarps = kThreads / 32;
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int ntiles = (n + kTile - 1) / kTile;
  __shared__ int s_tile;
  __shared__ int s_warp[kWarps];
  __shared__ int s_prefix;
  const int lane = lane_id(), warp = threadIdx.x >> 5;

  if (n == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      if (total_out)
        *total_out = 0;
      if (total_at_n)
        total_at_n[0] = 0;
    }
    return;
  }
  for (;;) {
    if (threadIdx.x == 0)
      s_tile = atomicAdd(&ctrl->tile, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= ntiles)
      break;
    const int base = tile * kTile + threadIdx.x * kItems;
    int vals[kItems];
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      int idx = base + k;
      vals[k] = (idx < n) ? value(idx) : 0;
      tsum += vals[k];
    }
    int incl = warp_inclusive_sum(tsum);
    if (lane == 31)
      s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      int x = s_warp[w];
      if (w < warp)
        warp_off += x;
      block_total += x;
    }
    if (warp == 0) {
      if (lane == 0)
        st_release(&state[tile], pack(epoch, tile == 0 ? kPrefix : kAggregate, block_total));
      int exclusive = 0;
      if (tile > 0) {
        int look = tile - 1;
        for (;;) {
          int idx = look - lane;
          unsigned long long st = idx >= 0 ? ld_acquire(&state[idx]) : pack(epoch, kPrefix, 0);
          unsigned hi = static_cast<unsigned>(st >> 32);
          unsigned flag = ((hi >> 2) == epoch) ? (hi & 3u) : kInvalid;
          if (__any_sync(kFull, flag == kInvalid))
            continue;  // a predecessor has not published yet
          int v = static_cast<int>(static_cast<unsigned>(st));
          unsigned pm = __ballot_sync(kFull, flag == kPrefix);
          if (pm) {
            int first = __ffs(pm) - 1;
            exclusive += warp_sum(lane <= first ? v : 0);
            break;
          }
          exclusive += warp_sum(v);
          look -= 32;
        }
        if (lane == 0)
          st_release(&state[tile], pack(epoch, kPrefix, exclusive + block_total));
      }
      if (lane == 0)
        s_prefix = exclusive;
    }
    __syncthreads();
    int run = s_prefix + warp_off + (incl - tsum);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      int idx = base + k;
      if (idx < n)
        emit(idx, run, vals[k]);
      run += vals[k];
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) {
      if (total_out)
        *total_out = s_prefix + block_total;
      if (total_at_n)
        total_at_n[n] = s_prefix + block_total;
    }
    __syncthreads();  // s_tile / s_warp reuse
  }
}

/// Host launcher: persistent grid, ticket + epoch bookkeeping from the workspace.
template <typename ValueF, typename EmitF>
inline void lookback_scan(workspace_t& ws,
                          const int* n_ptr,
                          int n_fixed,
                          int n_upper_bound,
                          ValueF value,
                          EmitF emit,
                          int* total_out,
                          int* total_at_n = nullptr) {
  constexpr int kThreads = 256, kItems = 8;
  unsigned& epoch = ws.scan_epoch;
  int max_tiles = (n_upper_bound + kThreads * kItems - 1) / (kThreads * kItems) + 1;
  size_t had = ws.tile_state.cap;
  ws.tile_state.ensure(max_tiles);
  if (ws.tile_state.cap != had)  // fresh allocation: clear once, epochs do the rest
    B2G_CHECK(cudaMemsetAsync(ws.tile_state.ptr, 0, ws.tile_state.cap * 8, ws.stream));
  epoch = (epoch + 1) & 0x3fffffffu;
  if (epoch == 0) {
    B2G_CHECK(cudaMemsetAsync(ws.tile_state.ptr, 0, ws.tile_state.cap * 8, ws.stream));
    epoch = 1;
  }
  ctrl_t* ctrl = ws.next_ctrl();
  int grid = device_info_t::get().sm_count * 4;
  if (grid > max_tiles)
    grid = max_tiles;
  lookback_scan_kernel<kThreads, kItems><<<grid, kThreads, 0, ws.stream>>>(
      n_ptr, n_fixed, value, emit, total_out, total_at_n, ws.tile_state.ptr, ctrl, epoch);
  ws.launches += 1;
  B2G_CHECK(cudaGetLastError());
}

}  // namespace b200
}  // namespace gunrock
