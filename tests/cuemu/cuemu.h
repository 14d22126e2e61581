// tests/cuemu/cuemu.h -- TEST INFRASTRUCTURE: a small CPU emulator for the CUDA device code of this repository.
//
// It exists so that kernel LOGIC (index arithmetic, work distribution, compaction, functor protocols) can be
// exercised by the CPU test suite: a kernel template from include/gunrock/b200/*.cuh is compiled by g++ with the
// CUDA keywords defined away and run with one OS thread per WARP, the 32 lanes being cooperatively scheduled
// fibers on it: a lane runs until it reaches a warp collective (__shfl*, __ballot, __reduce_or, __syncwarp) or a
// barrier, then its warp mates get the thread -- lanes are never in lockstep, so a missing __syncwarp shows.
// __syncthreads / barrier.cluster are OS barriers between the warps' threads, atomics are the compiler's, `__shared__` arrays are function statics (CTAs run one after the other) and dynamic shared
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

#include <ucontext.h>

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

struct cluster_ctx;
struct cta_ctx;

/// One warp = one OS thread that runs its 32 lanes as cooperatively scheduled fibers (ucontext): a lane runs
/// until it reaches a warp collective or a barrier, then the next runnable lane gets the thread.  Warp
/// collectives therefore cost a user-space context switch, not a futex.
struct warp_ctx {
  static constexpr std::size_t kStack = 256 * 1024;
  ucontext_t scheduler;
  ucontext_t lane_ctx[32];
  std::vector<unsigned char> stacks;
  bool done[32];
  int alive = 32;    // lanes that have not left the kernel
  int arrived = 0;   // lanes waiting at the current warp-level rendezvous
  unsigned gen = 0;  // rendezvous generation
  unsigned long long slot[32];
  unsigned index = 0;  // warp index in the CTA
  cta_ctx* cta = nullptr;
  const std::function<void()>* body = nullptr;
};
struct cta_ctx {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> bar;  // one participant per warp (its OS thread)
  std::vector<std::unique_ptr<warp_ctx>> warps;
  std::vector<unsigned char> smem;  // dynamic shared memory
  unsigned rank = 0;                // rank in the cluster
  unsigned block = 0;               // blockIdx.x
  cluster_ctx* cluster = nullptr;
};
struct cluster_ctx {
  std::vector<cta_ctx*> ctas;
  std::unique_ptr<std::barrier<>> bar;  // one participant per warp of every CTA
};

inline thread_local cta_ctx* t_cta = nullptr;
inline thread_local warp_ctx* t_warp = nullptr;
inline thread_local unsigned t_lane = 0;

}  // namespace cuemu

inline thread_local cuemu::uint3_t threadIdx, blockIdx;
inline cuemu::dim3_t blockDim, gridDim;

namespace cuemu {

/// Give the OS thread to the next lane of this warp (called by a lane that has to wait for its warp mates).
inline void yield() {
  warp_ctx* w = t_warp;
  swapcontext(&w->lane_ctx[t_lane], &w->scheduler);
}
/// All lanes of the warp that are still inside the kernel meet here.
inline void warp_rendezvous() {
  warp_ctx* w = t_warp;
  if (++w->arrived == w->alive) {
    w->arrived = 0;
    ++w->gen;
    return;
  }
  const unsigned my_gen = w->gen;
  while (w->gen == my_gen)
    yield();
}
/// The first lane still alive does the OS-level part of a block / cluster barrier for its warp.
inline bool is_warp_leader() {
  warp_ctx* w = t_warp;
  for (unsigned l = 0; l < 32; ++l)
    if (!w->done[l])
      return l == t_lane;
  return false;
}
inline void lane_trampoline(unsigned lane) {
  warp_ctx* w = t_warp;
  (*w->body)();
  w->done[lane] = true;
  --w->alive;
  if (w->alive > 0 && w->arrived == w->alive) {  // the others were only waiting for this lane
    w->arrived = 0;
    ++w->gen;
  }
  swapcontext(&w->lane_ctx[lane], &w->scheduler);  // never resumed
}
inline void run_warp(warp_ctx* w) {
  t_warp = w;
  t_cta = w->cta;
  ::blockIdx.x = w->cta->block;
  w->stacks.resize(32 * warp_ctx::kStack);
  for (unsigned l = 0; l < 32; ++l) {
    w->done[l] = false;
    getcontext(&w->lane_ctx[l]);
    w->lane_ctx[l].uc_stack.ss_sp = w->stacks.data() + l * warp_ctx::kStack;
    w->lane_ctx[l].uc_stack.ss_size = warp_ctx::kStack;
    w->lane_ctx[l].uc_link = &w->scheduler;
    makecontext(&w->lane_ctx[l], reinterpret_cast<void (*)()>(lane_trampoline), 1, l);
  }
  while (w->alive > 0)
    for (unsigned l = 0; l < 32; ++l)
      if (!w->done[l]) {
        t_lane = l;
        ::threadIdx.x = w->index * 32 + l;
        swapcontext(&w->scheduler, &w->lane_ctx[l]);
      }
  // a warp that has left the kernel no longer takes part in block / cluster barriers (as on the GPU)
  w->cta->bar->arrive_and_drop();
  w->cta->cluster->bar->arrive_and_drop();
}

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
  const unsigned warps = block / 32;
  for (unsigned first = 0; first < grid; first += cluster) {
    cluster_ctx cl;
    cl.bar = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(cluster) * warps);
    std::vector<std::unique_ptr<cta_ctx>> ctas;
    for (unsigned r = 0; r < cluster; ++r) {
      auto c = std::make_unique<cta_ctx>();
      c->nthreads = block;
      c->bar = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(warps));
      for (unsigned w = 0; w < warps; ++w) {
        c->warps.push_back(std::make_unique<warp_ctx>());
        c->warps.back()->index = w;
        c->warps.back()->cta = c.get();
        c->warps.back()->body = &body;
      }
      c->smem.assign(dynamic_smem_bytes + 64, 0xCD);  // poison: nothing may rely on zeroed shared memory
      c->rank = r;
      c->block = first + r;
      c->cluster = &cl;
      cl.ctas.push_back(c.get());
      ctas.push_back(std::move(c));
    }
    std::vector<std::thread> threads;
    for (unsigned r = 0; r < cluster; ++r)
      for (unsigned w = 0; w < warps; ++w)
        threads.emplace_back(run_warp, ctas[r]->warps[w].get());
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
  warp_rendezvous();
  raw = w->slot[src_lane & 31];
  warp_rendezvous();  // nobody overwrites a slot somebody still reads
  T out;
  std::memcpy(&out, &raw, sizeof out);
  return out;
}
inline unsigned gather_or(unsigned mine) {
  warp_ctx* w = t_warp;
  w->slot[t_lane] = mine;
  warp_rendezvous();
  unsigned all = 0;
  for (int l = 0; l < 32; ++l)
    if (!w->done[l])
      all |= static_cast<unsigned>(w->slot[l]);
  warp_rendezvous();
  return all;
}
/// Block-level (cluster == false) or cluster-level barrier: lanes meet in their warp, the warp's leader meets
/// the other warps at the OS barrier, then the lanes are released.
inline void wide_barrier(bool cluster) {
  warp_rendezvous();
  if (is_warp_leader()) {
    if (cluster)
      t_cta->cluster->bar->arrive_and_wait();
    else
      t_cta->bar->arrive_and_wait();
  }
  warp_rendezvous();
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
  cuemu::warp_rendezvous();
}
inline void __syncthreads() {
  cuemu::wide_barrier(false);
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
struct int2 {
  int x, y;
};
inline int2 make_int2(int x, int y) { return int2{x, y}; }
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
