// tests/cuemu/gunrock/b200/ptx.cuh -- TEST INFRASTRUCTURE: plain C++ stand-ins, under tests/cuemu/cuemu.h, for
// every helper of include/gunrock/b200/ptx.cuh (the one header of the product that holds inline PTX).  Same
// names, same signatures, functional behaviour only; this directory precedes include/ on the emulator's include
// path, so the kernel headers pick this file up unchanged.
#pragma once

#include <chrono>

#include <cuemu.h>

namespace gunrock {
namespace b200 {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

/// Event counters (relaxed, approximate under concurrency): how often each memory path was taken.
struct cuemu_counters_t {
  unsigned long long cached_probes = 0;   // ld_cached: L1-cached global probes (visited / frontier bitmaps)
  unsigned long long shared_probes = 0;   // ld_shared_u32: on-chip copy, own CTA
  unsigned long long dsmem_probes = 0;    // ld_dsmem_u32: on-chip copy, another CTA of the cluster
  unsigned long long shared_merges = 0;   // red_shared_or + red_dsmem_or
  unsigned long long stream_loads = 0;    // ld_stream (column indices / weights)
};
inline cuemu_counters_t cuemu_counters;
inline void cuemu_count(unsigned long long& c) { __atomic_fetch_add(&c, 1ull, __ATOMIC_RELAXED); }

inline int lane_id() { return static_cast<int>(threadIdx.x & 31); }
inline unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

template <typename T>
inline T warp_inclusive_sum(T x) {
  for (int d = 1; d < 32; d <<= 1) {
    T y = __shfl_up_sync(kFull, x, static_cast<unsigned>(d));
    if (lane_id() >= d)
      x += y;
  }
  return x;
}
template <typename T>
inline T warp_sum(T x) {
  for (int d = 16; d > 0; d >>= 1)
    x += __shfl_xor_sync(kFull, x, d);
  return x;
}
inline float warp_max(float x) {
  for (int d = 16; d > 0; d >>= 1)
    x = fmaxf(x, __shfl_xor_sync(kFull, x, d));
  return x;
}

inline int ld_stream(const int* p) {
  cuemu_count(cuemu_counters.stream_loads);
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
inline float ld_stream(const float* p) { return *p; }
inline unsigned ld_relaxed(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline int ld_relaxed(const int* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline float ld_relaxed(const float* p) {
  unsigned u = __atomic_load_n(reinterpret_cast<const unsigned*>(p), __ATOMIC_RELAXED);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline unsigned long long ld_acquire(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

inline void st_release_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned ld_acquire_sys(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long global_timer_ns() {
  return static_cast<unsigned long long>(std::chrono::duration_cast<std::chrono::nanoseconds>(
                                             std::chrono::steady_clock::now().time_since_epoch()).count());
}

// ---- shared-window addresses: offsets into the CTA's dynamic shared memory --------------------------------
inline unsigned char* dynamic_smem() { return cuemu::t_cta->smem.data(); }
inline uint32_t smem_u32(const void* p) {
  const unsigned char* base = cuemu::t_cta->smem.data();
  const unsigned char* q = static_cast<const unsigned char*>(p);
  if (q < base || q >= base + cuemu::t_cta->smem.size()) {
    std::fprintf(stderr, "cuemu: smem_u32 of a pointer outside the dynamic shared memory\n");
    std::abort();
  }
  return static_cast<uint32_t>(q - base);
}
inline unsigned* cuemu_word(cuemu::cta_ctx* c, uint32_t addr) {
  if (addr % 4 != 0 || addr + 4 > c->smem.size()) {
    std::fprintf(stderr, "cuemu: shared-window access out of range (%u of %zu)\n", addr, c->smem.size());
    std::abort();
  }
  return reinterpret_cast<unsigned*>(c->smem.data() + addr);
}
inline unsigned ld_shared_u32(uint32_t addr) {
  cuemu_count(cuemu_counters.shared_probes);
  return __atomic_load_n(cuemu_word(cuemu::t_cta, addr), __ATOMIC_RELAXED);
}
inline void red_shared_or(uint32_t addr, unsigned value) {
  cuemu_count(cuemu_counters.shared_merges);
  __atomic_fetch_or(cuemu_word(cuemu::t_cta, addr), value, __ATOMIC_SEQ_CST);
}
inline cuemu::cta_ctx* cuemu_peer(unsigned rank) {
  auto* cl = cuemu::t_cta->cluster;
  if (rank >= cl->ctas.size()) {
    std::fprintf(stderr, "cuemu: cluster rank %u out of range\n", rank);
    std::abort();
  }
  return cl->ctas[rank];
}
inline unsigned ld_dsmem_u32(uint32_t addr, unsigned rank) {
  cuemu_count(cuemu_counters.dsmem_probes);
  return __atomic_load_n(cuemu_word(cuemu_peer(rank), addr), __ATOMIC_RELAXED);
}
inline void red_dsmem_or(uint32_t addr, unsigned rank, unsigned value) {
  cuemu_count(cuemu_counters.shared_merges);
  __atomic_fetch_or(cuemu_word(cuemu_peer(rank), addr), value, __ATOMIC_SEQ_CST);
}
inline unsigned cluster_cta_rank() { return cuemu::t_cta->rank; }
inline void cluster_barrier() { cuemu::wide_barrier(true); }

// ---- mbarrier + bulk copy: the copy happens at issue, the barrier word keeps (phase, pending bytes) --------
inline void mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  __atomic_fetch_add(bar, static_cast<uint64_t>(bytes) << 1, __ATOMIC_SEQ_CST);
}
inline void cuemu_complete_tx(uint64_t* bar, uint32_t bytes) {
  uint64_t old = __atomic_load_n(bar, __ATOMIC_SEQ_CST), want;
  do {
    uint64_t pending = (old >> 1) - bytes;
    want = pending ? ((pending << 1) | (old & 1)) : ((old & 1) ^ 1);  // last byte: flip the phase
  } while (!__atomic_compare_exchange_n(bar, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  return (__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1) != parity;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity))
    cuemu::yield();  // the copy may be issued by a warp mate that has not run yet
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  std::memcpy(dst_smem, src_gmem, bytes);
  cuemu_complete_tx(bar, bytes);
}
inline void fence_proxy_async() {}

// ---- bitmap helpers ----------------------------------------------------------------------------------------
inline unsigned ld_cached(const unsigned* p) {
  cuemu_count(cuemu_counters.cached_probes);
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
inline bool bitmap_test(const unsigned* bm, int v) { return (ld_cached(bm + (v >> 5)) >> (v & 31)) & 1u; }
inline bool bitmap_test_and_set(unsigned* bm, int v) {
  unsigned bit = 1u << (v & 31);
  unsigned* w = bm + (v >> 5);
  if (ld_cached(w) & bit)
    return false;
  return !(atomicOr(w, bit) & bit);
}

template <typename F>
void for_each_index(int n, F f) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    f(i);
}

}  // namespace b200
}  // namespace gunrock
