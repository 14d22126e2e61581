// tests/cuemu/cuemu.h -- TEST INFRASTRUCTURE: a small CPU emulator for the CUDA device code of this repository.
//
// It exists so that kernel LOGIC (index arithmetic, work distribution, compaction, functor protocols) can be
// exercised by the CPU test suite: a kernel template from include/gunrock/b200/*.cuh is compiled by g++ with the
// CUDA keywords defined away and run with one OS thread per CUDA thread.  Warp collectives (__shfl*, __ballot,
// __reduce_or, __syncwarp) exchange through a per-warp barrier, __syncthreads is a per-CTA barrier, atomics are
// the compiler's, `__shared__` arrays are function statics (CTAs run one after the other) and dynamic shared
// memory is a per-CTA buffer -- the CTAs of a thread-block cluster run concurrently and reach each other's
// buffer for the distributed-shared-memory helpers.  Inline PTX is confined to gunrock/b200/ptx.cuh, which this
// directory shadows with a plain C++ version (tests/cuemu/gunrock/b200/ptx.cuh).
// It models functional behaviour only: no timing, no memory model weaker than sequential consistency, and every
// full-mask collective must be reached by all 32 lanes (true of every kernel here by construction).
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace cuemu {

struct uint3_t {
  unsigned x = 0, y = 0, z = 0;
};
struct dim3_t {
  unsigned x = 1, y = 1, z = 1;
};

struct warp_ctx {
  std::barrier<> bar{32};
  unsigned long long slot[32];
};
struct cluster_ctx;
struct cta_ctx {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<warp_ctx>> warps;
  std::vector<unsigned char> smem;  // dynamic shared memory
  unsigned rank = 0;                // rank in the cluster
  cluster_ctx* cluster = nullptr;
};
struct cluster_ctx {
  std::vector<cta_ctx*> ctas;
  std::unique_ptr<std::barrier<>> bar;  // every thread of every CTA
};

inline thread_local cta_ctx* t_cta = nullptr;
inline thread_local warp_ctx* t_warp = nullptr;
inline thread_local unsigned t_lane = 0;

}  // namespace cuemu

inline thread_local cuemu::uint3_t threadIdx, blockIdx;
inline cuemu::dim3_t blockDim, gridDim;

namespace cuemu {

/// Run `body` as a kernel: grid x block threads, `cluster` consecutive CTAs at a time (1 = plain launch).
/// blockDim.x must be a multiple of 32.
inline void launch(unsigned grid, unsigned block, std::size_t dynamic_smem_bytes, unsigned cluster,
                   const std::function<void()>& body) {
  if (block % 32 != 0 || grid % cluster != 0) {
    std::fprintf(stderr, "cuemu::launch: bad shape\n");
    std::abort();
  }
  ::blockDim.x = block;
  ::gridDim.x = grid;
  for (unsigned first = 0; first < grid; first += cluster) {
    cluster_ctx cl;
    cl.bar = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(cluster) * block);
    std::vector<std::unique_ptr<cta_ctx>> ctas;
    for (unsigned r = 0; r < cluster; ++r) {
      auto c = std::make_unique<cta_ctx>();
      c->nthreads = block;
      c->bar = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(block));
      for (unsigned w = 0; w < block / 32; ++w)
        c->warps.push_back(std::make_unique<warp_ctx>());
      c->smem.assign(dynamic_smem_bytes + 64, 0xCD);  // poison: nothing may rely on zeroed shared memory
      c->rank = r;
      c->cluster = &cl;
      cl.ctas.push_back(c.get());
      ctas.push_back(std::move(c));
    }
    std::vector<std::thread> threads;
    for (unsigned r = 0; r < cluster; ++r)
      for (unsigned t = 0; t < block; ++t)
        threads.emplace_back([&, r, t] {
          cta_ctx* c = ctas[r].get();
          t_cta = c;
          t_warp = c->warps[t / 32].get();
          t_lane = t % 32;
          ::threadIdx.x = t;
          ::blockIdx.x = first + r;
          body();
          // a thread that has left the kernel no longer takes part in any barrier (as on the GPU)
          t_warp->bar.arrive_and_drop();
          c->bar->arrive_and_drop();
          cl.bar->arrive_and_drop();
        });
    for (auto& th : threads)
      th.join();
  }
}

template <typename T>
inline T exchange(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffles move at most 8 bytes");
  warp_ctx* w = t_warp;
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof v);
  w->slot[t_lane] = raw;
  w->bar.arrive_and_wait();
  raw = w->slot[src_lane & 31];
  w->bar.arrive_and_wait();  // nobody overwrites a slot somebody still reads
  T out;
  std::memcpy(&out, &raw, sizeof out);
  return out;
}
inline unsigned gather_or(unsigned mine) {
  warp_ctx* w = t_warp;
  w->slot[t_lane] = mine;
  w->bar.arrive_and_wait();
  unsigned all = 0;
  for (int l = 0; l < 32; ++l)
    all |= static_cast<unsigned>(w->slot[l]);
  w->bar.arrive_and_wait();
  return all;
}

}  // namespace cuemu

// ---- warp collectives (full masks only) -------------------------------------------------------------------
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  return cuemu::exchange(v, static_cast<unsigned>(src));
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  T got = cuemu::exchange(v, cuemu::t_lane >= d ? cuemu::t_lane - d : cuemu::t_lane);
  return cuemu::t_lane >= d ? got : v;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int d) {
  return cuemu::exchange(v, cuemu::t_lane ^ static_cast<unsigned>(d));
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  return cuemu::gather_or(pred ? (1u << cuemu::t_lane) : 0u);
}
inline bool __any_sync(unsigned m, bool pred) {
  return __ballot_sync(m, pred) != 0;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  return cuemu::gather_or(v);
}
inline void __syncwarp(unsigned = 0xffffffffu) {
  cuemu::t_warp->bar.arrive_and_wait();
}
inline void __syncthreads() {
  cuemu::t_cta->bar->arrive_and_wait();
}

// ---- scalar intrinsics ------------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---- atomics (sequentially consistent) ---------------------------------------------------------------------
template <typename T>
inline T cuemu_fetch_op(T* p, T v, T (*op)(T, T)) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (!__atomic_compare_exchange_n(p, &old, op(old, v), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int expected, int desired) {
  __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;
}
inline int atomicMin(int* p, int v) { return cuemu_fetch_op<int>(p, v, [](int a, int b) { return a < b ? a : b; }); }
inline int atomicMax(int* p, int v) { return cuemu_fetch_op<int>(p, v, [](int a, int b) { return a > b ? a : b; }); }
inline unsigned atomicMin(unsigned* p, unsigned v) {
  return cuemu_fetch_op<unsigned>(p, v, [](unsigned a, unsigned b) { return a < b ? a : b; });
}
inline unsigned atomicMax(unsigned* p, unsigned v) {
  return cuemu_fetch_op<unsigned>(p, v, [](unsigned a, unsigned b) { return a > b ? a : b; });
}
