/**
 * @file ptx.cuh
 * @brief sm_100a device primitives used by the frontier kernels: warp scans/ballots, mbarrier +
 * 1-D bulk async copy (TMA engine, SASS `UBLKCP`), cache-hinted loads, bitmap helpers.
 *
 * Nothing here comes from Thrust/CUB/ModernGPU; it replaces the reference's use of
 * cub::BlockScan / cub::ThreadLoad (include/gunrock/framework/operators/advance/block_mapped.hxx:89,123,
 * include/gunrock/util/load_store.hxx:57-85).
 */
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace gunrock {
namespace b200 {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() {
  return threadIdx.x & 31;
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

/// Inclusive warp scan (Kogge-Stone over shuffles).
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T x) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T y = __shfl_up_sync(kFull, x, d);
    if (lane_id() >= d)
      x += y;
  }
  return x;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    x += __shfl_xor_sync(kFull, x, d);
  return x;
}

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    x = fmaxf(x, __shfl_xor_sync(kFull, x, d));
  return x;
}

/// Streaming (read-once) 4-byte load: do not pollute L1 with the column-index stream.
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

/// Relaxed (non-coherent-free) load used for racy-but-monotone reads of bitmaps / labels.
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_relaxed(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/// System-scope release / acquire on a 32-bit flag (peer-memory barriers between GPUs, bfs_p2p.cuh).
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
/// Nanosecond wall clock of the GPU (time-outs of spinning barriers).
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (cp.async.bulk, executed by the TMA unit; SASS UBLKCP).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
}
/// Make mbarrier.init visible to the async proxy (the TMA unit) before the first copy.
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
/// Global -> shared bulk copy. dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem,
                                         const void* src_gmem,
                                         uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
/// Order prior generic-proxy accesses to shared memory before later async-proxy (TMA) writes.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Shared-window accesses by address, thread-block clusters and distributed shared memory (DSMEM).
// `addr` is a shared-window address (smem_u32) valid in the calling CTA; `rank` names the CTA of the cluster
// whose copy of that address is meant (mapa), read / updated over the SM-to-SM network.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_shared_u32(uint32_t addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_shared_or(uint32_t addr, unsigned value) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(value) : "memory");
}
__device__ __forceinline__ unsigned ld_dsmem_u32(uint32_t addr, unsigned rank) {
  uint32_t remote;
  unsigned v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote));
  return v;
}
__device__ __forceinline__ void red_dsmem_or(uint32_t addr, unsigned rank, unsigned value) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
  asm volatile("red.shared::cluster.or.b32 [%0], %1;" ::"r"(remote), "r"(value) : "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
/// Full cluster barrier (every thread of every CTA of the cluster).
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/// Base of the dynamic shared memory of this CTA (`extern __shared__`), 16-byte aligned.
__device__ __forceinline__ unsigned char* dynamic_smem() {
  extern __shared__ __align__(16) unsigned char b2g_dynamic_smem[];
  return b2g_dynamic_smem;
}

// ---------------------------------------------------------------------------------------------
// Bitmap helpers (32 vertices per word).
// ---------------------------------------------------------------------------------------------
/// L1-cacheable weak load.  Used for the visited / frontier bitmaps: bits only ever go 0 -> 1, so
/// a stale line can only answer "not set", and every such answer is re-checked by an atomic.
__device__ __forceinline__ unsigned ld_cached(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
/// Read-only (for the whole kernel) bitmap probe through the non-coherent path.
__device__ __forceinline__ bool bitmap_test(const unsigned* bm, int v) {
  return (__ldg(bm + (v >> 5)) >> (v & 31)) & 1u;
}
/// Returns true iff this call set the bit (i.e. the caller "won" the vertex).
__device__ __forceinline__ bool bitmap_test_and_set(unsigned* bm, int v) {
  unsigned bit = 1u << (v & 31);
  unsigned* w = bm + (v >> 5);
  if (ld_cached(w) & bit)
    return false;
  return !(atomicOr(w, bit) & bit);
}

/// Generic grid-stride kernel: f(i) for i in [0, n).
template <typename F>
__global__ void for_each_index(int n, F f) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    f(i);
}

}  // namespace b200
}  // namespace gunrock
