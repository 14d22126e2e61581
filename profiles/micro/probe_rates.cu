// Microbenchmark for the next push-BFS experiment: how fast can one SM
// answer random 4-byte "is this bit set" probes from (a) L1/L2 (`ld.global.ca`), (b) L2 only
// (`ld.global.cg`), (c) its own shared memory, (d) a 50/50 mix, (e) distributed shared memory of an
// 8-CTA cluster.  Prints giga-probes per second for the whole GPU.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_rates probe_rates.cu && ./probe_rates
// Not part of the product or the tests; numbers printed here are design inputs, not bench values.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CHECK(x)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(1);                                                                   \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ unsigned ld_ca(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_cg(const unsigned* p) {
  unsigned v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

constexpr int kIlp = 4;

// mode 0: ld.global.ca, 1: ld.global.cg, 2: shared, 3: half shared / half global
template <int kMode>
__global__ void probe_kernel(const unsigned* __restrict__ table, unsigned table_words, unsigned smem_words,
                             int iters, unsigned* sink) {
  extern __shared__ unsigned s_table[];
  if (kMode >= 2) {
    for (unsigned i = threadIdx.x; i < smem_words; i += blockDim.x)
      s_table[i] = table[i % table_words];
    __syncthreads();
  }
  unsigned h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    unsigned idx[kIlp], val[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      h = mix(h + 0x9e3779b9U);
      idx[k] = h;
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      if (kMode == 0)
        val[k] = ld_ca(table + idx[k] % table_words);
      else if (kMode == 1)
        val[k] = ld_cg(table + idx[k] % table_words);
      else if (kMode == 2)
        val[k] = s_table[idx[k] % smem_words];
      else
        val[k] = (idx[k] & 0x10000u) ? s_table[idx[k] % smem_words] : ld_ca(table + idx[k] % table_words);
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k)
      acc += (val[k] >> (idx[k] & 31)) & 1u;
  }
  if (acc == 0xffffffffu)
    *sink = acc;
}

// distributed shared memory: every CTA of an 8-CTA cluster holds a slice, probes go to a random slice
__global__ void __cluster_dims__(8, 1, 1)
dsmem_probe_kernel(const unsigned* __restrict__ table, unsigned table_words, unsigned slice_words, int iters,
                   unsigned* sink) {
  extern __shared__ unsigned s_table[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  for (unsigned i = threadIdx.x; i < slice_words; i += blockDim.x)
    s_table[i] = table[(rank * slice_words + i) % table_words];
  cluster.sync();
  unsigned h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    unsigned idx[kIlp], val[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      h = mix(h + 0x9e3779b9U);
      idx[k] = h;
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
      const unsigned* remote = cluster.map_shared_rank(s_table, (idx[k] >> 20) & 7u);
      val[k] = remote[idx[k] % slice_words];
    }
#pragma unroll
    for (int k = 0; k < kIlp; ++k)
      acc += (val[k] >> (idx[k] & 31)) & 1u;
  }
  cluster.sync();
  if (acc == 0xffffffffu)
    *sink = acc;
}

template <typename Launch>
double time_ms(Launch&& launch) {
  cudaEvent_t a, b;
  CHECK(cudaEventCreate(&a));
  CHECK(cudaEventCreate(&b));
  launch();  // warm-up
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(a));
  for (int r = 0; r < 5; ++r)
    launch();
  CHECK(cudaEventRecord(b));
  CHECK(cudaEventSynchronize(b));
  float ms = 0;
  CHECK(cudaEventElapsedTime(&ms, a, b));
  return ms / 5.0;
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int iters = 2048;
  unsigned* sink;
  CHECK(cudaMalloc(&sink, 4));
  std::printf("%s, %d SMs\n", prop.name, sms);
  for (unsigned table_kib : {512u, 8192u}) {  // visited map of RMAT-22 / RMAT-26
    const unsigned words = table_kib * 256;
    unsigned* table;
    CHECK(cudaMalloc(&table, words * 4ull));
    CHECK(cudaMemset(table, 0x5a, words * 4ull));
    for (int ctas_per_sm : {8, 4, 2}) {
      const int threads = 2048 / ctas_per_sm;  // 64 warps per SM in every configuration
      const unsigned smem_bytes = (200u * 1024u) / ctas_per_sm;
      const double probes = double(sms) * ctas_per_sm * threads * double(iters) * kIlp;
      auto report = [&](const char* what, double ms) {
        std::printf("table %5u KiB  %d CTAs/SM x %4d thr  %-28s %8.1f Gprobe/s\n", table_kib, ctas_per_sm, threads,
                    what, probes / ms / 1e6);
      };
      CHECK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      CHECK(cudaFuncSetAttribute(probe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      report("ld.global.ca", time_ms([&] {
               probe_kernel<0><<<sms * ctas_per_sm, threads>>>(table, words, 0, iters, sink);
             }));
      report("ld.global.cg", time_ms([&] {
               probe_kernel<1><<<sms * ctas_per_sm, threads>>>(table, words, 0, iters, sink);
             }));
      report("shared (private copy)", time_ms([&] {
               probe_kernel<2><<<sms * ctas_per_sm, threads, smem_bytes>>>(table, words, smem_bytes / 4, iters, sink);
             }));
      report("half shared, half ld.ca", time_ms([&] {
               probe_kernel<3><<<sms * ctas_per_sm, threads, smem_bytes>>>(table, words, smem_bytes / 4, iters, sink);
             }));
    }
    {  // 8-CTA clusters, 1 CTA per SM, 200 KiB slice each => 1.6 MiB map per cluster
      const unsigned slice_bytes = 200u * 1024u;
      CHECK(cudaFuncSetAttribute(dsmem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, slice_bytes));
      const int grid = (sms / 8) * 8;
      const double probes = double(grid) * 1024 * double(iters) * kIlp;
      double ms = time_ms([&] {
        dsmem_probe_kernel<<<grid, 1024, slice_bytes>>>(table, words, slice_bytes / 4, iters, sink);
      });
      std::printf("table %5u KiB  cluster of 8 x 1024 thr      %-28s %8.1f Gprobe/s\n", table_kib,
                  "distributed shared memory", probes / ms / 1e6);
    }
    CHECK(cudaFree(table));
  }
  CHECK(cudaGetLastError());
  return 0;
}
