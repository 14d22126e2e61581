/* oracle/gunrock_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's algorithms for the advance->filter->compute hot
 * path (BFS / SSSP / PageRank over CSR) and of the ingest steps either side of it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker.  The product path (gunrock_b200/csrc, include/gunrock)
 * never links, imports or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Pinning status (see tests/test_oracle.py, tests/golden/):
 *   - orc_load_mtx, orc_csr_from_coo, orc_bfs, orc_sssp : PINNED against the reference's own
 *     code compiled from /root/reference (oracle/_ref, oracle/ref_driver.cu) and against golden
 *     vectors minted from it (tests/golden/*.json, script tests/golden/make_golden.py).
 *   - orc_pr : PARITY UNPINNED -- the reference has no CPU PageRank and no golden vector
 *     (SURVEY.md F7).  It restates include/gunrock/algorithms/pr.hxx:65-93,107-152,172-195 with
 *     a fixed summation order and is cross-checked only against a float64 power iteration.
 *   - orc_rmat_* : not reference code (the reference has no generator); this is
 *     the workload definition shared bit-for-bit with the CUDA generator.
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * MatrixMarket coordinate loader.
 * Follows include/gunrock/io/matrix_market.hxx:104-254:
 *   - banner decides pattern / real / integer and general / symmetric (:113-135, :153-200)
 *   - 1-based -> 0-based (:165-169), pattern => weight 1.0f (:168-169)
 *   - symmetric => every off-diagonal entry is followed *in place* by its mirror, diagonal
 *     entries kept once, order preserved (:203-246)
 * props[0]=directed, props[1]=weighted, props[2]=symmetric (:153-154,:204-205,:248-250).
 * Returns 0 on success; arrays are malloc'ed and must be released with orc_free.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_load_mtx(const char* filename,
                         int* n_rows,
                         int* n_cols,
                         int* nnz_out,
                         int** I_out,
                         int** J_out,
                         float** V_out,
                         int* props) {
  FILE* f = fopen(filename, "r");
  if (!f)
    return -1;
  char line[1100];
  if (!fgets(line, sizeof line, f)) {
    fclose(f);
    return -2;
  }
  char banner[64], mtx[64], crd[64], data_type[64], storage[64];
  if (sscanf(line, "%63s %63s %63s %63s %63s", banner, mtx, crd, data_type, storage) != 5 ||
      strcmp(banner, "%%MatrixMarket") != 0) {
    fclose(f);
    return -2;
  }
  for (char* p = mtx; *p; ++p) *p = (char)((*p >= 'A' && *p <= 'Z') ? *p + 32 : *p);
  for (char* p = crd; *p; ++p) *p = (char)((*p >= 'A' && *p <= 'Z') ? *p + 32 : *p);
  for (char* p = data_type; *p; ++p) *p = (char)((*p >= 'A' && *p <= 'Z') ? *p + 32 : *p);
  for (char* p = storage; *p; ++p) *p = (char)((*p >= 'A' && *p <= 'Z') ? *p + 32 : *p);
  if (strcmp(mtx, "matrix") != 0 || strcmp(crd, "coordinate") != 0) {
    fclose(f);
    return -3; /* "File is not a sparse matrix" */
  }
  int is_pattern = strcmp(data_type, "pattern") == 0;
  int is_real = strcmp(data_type, "real") == 0;
  int is_integer = strcmp(data_type, "integer") == 0;
  int is_symmetric = strcmp(storage, "symmetric") == 0;
  if (!is_pattern && !is_real && !is_integer) {
    fclose(f);
    return -4;
  }
  /* skip comments, read size line */
  size_t M = 0, N = 0, NNZ = 0;
  for (;;) {
    if (!fgets(line, sizeof line, f)) {
      fclose(f);
      return -5;
    }
    if (line[0] == '%')
      continue;
    if (sscanf(line, "%zu %zu %zu", &M, &N, &NNZ) == 3)
      break;
  }
  if (M >= (size_t)INT_MAX || N >= (size_t)INT_MAX || NNZ >= (size_t)INT_MAX) {
    fclose(f);
    return -6; /* vertex_t / edge_t overflow (:137-142) */
  }
  size_t cap = is_symmetric ? 2 * NNZ : NNZ;
  int* I = (int*)malloc((cap ? cap : 1) * sizeof(int));
  int* J = (int*)malloc((cap ? cap : 1) * sizeof(int));
  float* V = (float*)malloc((cap ? cap : 1) * sizeof(float));
  size_t ptr = 0;
  for (size_t i = 0; i < NNZ; ++i) {
    size_t r = 0, c = 0;
    double w = 1.0;
    int got;
    if (is_pattern)
      got = fscanf(f, " %zu %zu \n", &r, &c) == 2;
    else
      got = fscanf(f, " %zu %zu %lf \n", &r, &c, &w) == 3;
    if (!got || r == 0 || c == 0) {
      free(I);
      free(J);
      free(V);
      fclose(f);
      return -7;
    }
    int ri = (int)r - 1, ci = (int)c - 1;
    float wf = is_pattern ? 1.0f : (float)w;
    I[ptr] = ri;
    J[ptr] = ci;
    V[ptr] = wf;
    ++ptr;
    if (is_symmetric && ri != ci) {
      I[ptr] = ci;
      J[ptr] = ri;
      V[ptr] = wf;
      ++ptr;
    }
  }
  fclose(f);
  *n_rows = (int)M;
  *n_cols = (int)N;
  *nnz_out = (int)ptr;
  *I_out = I;
  *J_out = J;
  *V_out = V;
  props[0] = is_symmetric ? 0 : 1;
  props[1] = is_pattern ? 0 : 1;
  props[2] = is_symmetric ? 1 : 0;
  return 0;
}

ORC_API void orc_free(void* p) {
  free(p);
}

/* ------------------------------------------------------------------------------------------
 * COO -> CSR, stable counting sort by row; duplicates and self-loops kept; the order of a
 * row's entries is their COO order.  Follows include/gunrock/formats/csr.hxx:81-140.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_csr_from_coo(int n_rows,
                             int nnz,
                             const int* I,
                             const int* J,
                             const float* V,
                             int* row_offsets,
                             int* column_indices,
                             float* values) {
  memset(row_offsets, 0, sizeof(int) * ((size_t)n_rows + 1));
  for (int n = 0; n < nnz; ++n) ++row_offsets[I[n]];
  int sum = 0;
  for (int i = 0; i < n_rows; ++i) {
    int t = row_offsets[i];
    row_offsets[i] = sum;
    sum += t;
  }
  row_offsets[n_rows] = nnz;
  for (int n = 0; n < nnz; ++n) {
    int row = I[n];
    int dest = row_offsets[row];
    column_indices[dest] = J[n];
    values[dest] = V ? V[n] : 1.0f;
    ++row_offsets[row];
  }
  int last = 0;
  for (int i = 0; i <= n_rows; ++i) {
    int t = row_offsets[i];
    row_offsets[i] = last;
    last = t;
  }
  return 0;
}

/* CSR -> CSC (transpose), stable in source order: in-edges of a vertex are listed by increasing
 * source vertex, ties in CSR order.  The reference builds CSC with a device sort_by_key
 * (include/gunrock/formats/csc.hxx:62-102); the pull kernels only need *a* transpose, so the
 * in-row order is part of our layout, not of the reference contract. */
ORC_API int orc_csr_transpose(int n_rows,
                              int n_cols,
                              int nnz,
                              const int* ro,
                              const int* ci,
                              const float* vals,
                              int* t_offsets /* n_cols+1 */,
                              int* t_indices,
                              float* t_vals) {
  memset(t_offsets, 0, sizeof(int) * ((size_t)n_cols + 1));
  for (int e = 0; e < nnz; ++e) ++t_offsets[ci[e] + 1];
  for (int i = 0; i < n_cols; ++i) t_offsets[i + 1] += t_offsets[i];
  int* cur = (int*)malloc(sizeof(int) * ((size_t)n_cols + 1));
  memcpy(cur, t_offsets, sizeof(int) * ((size_t)n_cols + 1));
  for (int u = 0; u < n_rows; ++u)
    for (int e = ro[u]; e < ro[u + 1]; ++e) {
      int d = cur[ci[e]]++;
      t_indices[d] = u;
      if (t_vals)
        t_vals[d] = vals ? vals[e] : 1.0f;
    }
  free(cur);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Binary min-heap on (key, vertex).  The reference uses std::priority_queue with a comparator
 * on .second (bfs_cpu.hxx:12-18,41-44).  Tie order differs between heap implementations but the
 * fixed point they converge to does not (every pop relaxes out-edges with curr_dist + w and the
 * distance array only ever decreases), so the result is the same array.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int v;
  int d;
} ipair_t;
typedef struct {
  int v;
  float d;
} fpair_t;

#define DEFINE_HEAP(NAME, PAIR)                                     \
  typedef struct {                                                  \
    PAIR* a;                                                        \
    size_t n, cap;                                                  \
  } NAME##_t;                                                       \
  static void NAME##_push(NAME##_t* h, PAIR x) {                    \
    if (h->n == h->cap) {                                           \
      h->cap = h->cap ? h->cap * 2 : 1024;                          \
      h->a = (PAIR*)realloc(h->a, h->cap * sizeof(PAIR));           \
    }                                                               \
    size_t i = h->n++;                                              \
    while (i > 0) {                                                 \
      size_t p = (i - 1) >> 1;                                      \
      if (!(h->a[p].d > x.d))                                       \
        break;                                                      \
      h->a[i] = h->a[p];                                            \
      i = p;                                                        \
    }                                                               \
    h->a[i] = x;                                                    \
  }                                                                 \
  static PAIR NAME##_pop(NAME##_t* h) {                             \
    PAIR top = h->a[0];                                             \
    PAIR x = h->a[--h->n];                                          \
    size_t i = 0;                                                   \
    for (;;) {                                                      \
      size_t c = 2 * i + 1;                                         \
      if (c >= h->n)                                                \
        break;                                                      \
      if (c + 1 < h->n && h->a[c + 1].d < h->a[c].d)                \
        ++c;                                                        \
      if (!(x.d > h->a[c].d))                                       \
        break;                                                      \
      h->a[i] = h->a[c];                                            \
      i = c;                                                        \
    }                                                               \
    if (h->n)                                                       \
      h->a[i] = x;                                                  \
    return top;                                                     \
  }

DEFINE_HEAP(iheap, ipair_t)
DEFINE_HEAP(fheap, fpair_t)

/* ------------------------------------------------------------------------------------------
 * BFS depths.  Follows examples/algorithms/bfs/bfs_cpu.hxx:32-63: distances initialised to
 * INT_MAX (:32-33), distances[src]=0 (:37), lazy-Dijkstra with unit weights on a priority
 * queue -- *no* stale-entry check on pop (:46-62), predecessors untouched.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_bfs(int n_vertices, const int* ro, const int* ci, int source, int* distances) {
  for (int i = 0; i < n_vertices; ++i) distances[i] = INT_MAX;
  if (source < 0 || source >= n_vertices)
    return -1;
  distances[source] = 0;
  iheap_t pq = {0, 0, 0};
  ipair_t s = {source, 0};
  iheap_push(&pq, s);
  while (pq.n) {
    ipair_t cur = iheap_pop(&pq);
    int start = ro[cur.v], end = ro[cur.v + 1];
    for (int off = start; off < end; ++off) {
      int nb = ci[off];
      int nd = cur.d + 1;
      if (nd < distances[nb]) {
        distances[nb] = nd;
        ipair_t x = {nb, nd};
        iheap_push(&pq, x);
      }
    }
  }
  free(pq.a);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * SSSP distances (fp32).  Follows examples/algorithms/sssp/sssp_cpu.hxx:36-67: FLT_MAX init
 * (:36-37), new_dist = curr_dist + w in float (:59), strict < (:60).
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_sssp(int n_vertices,
                     const int* ro,
                     const int* ci,
                     const float* w,
                     int source,
                     float* distances) {
  for (int i = 0; i < n_vertices; ++i) distances[i] = FLT_MAX;
  if (source < 0 || source >= n_vertices)
    return -1;
  distances[source] = 0.0f;
  fheap_t pq = {0, 0, 0};
  fpair_t s = {source, 0.0f};
  fheap_push(&pq, s);
  while (pq.n) {
    fpair_t cur = fheap_pop(&pq);
    int start = ro[cur.v], end = ro[cur.v + 1];
    for (int off = start; off < end; ++off) {
      int nb = ci[off];
      volatile float nd = cur.d + w[off]; /* volatile: force a rounded fp32 value, no x87/fma games */
      if (nd < distances[nb]) {
        distances[nb] = nd;
        fpair_t x = {nb, nd};
        fheap_push(&pq, x);
      }
    }
  }
  free(pq.a);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * PageRank, restating include/gunrock/algorithms/pr.hxx (NO reference CPU path: unpinned).
 *   reset   (:65-93)  p = (float)(1.0/V); iweights[v] = alpha / (sequential fp32 sum of the
 *                      row's weights) or 0 if that sum is 0
 *   loop    (:107-152) plast = p; dsum = sum_{iweights[i]==0} alpha*p[i];
 *                      p[:] = (1 - alpha + dsum)/V; p[dst] += plast[src]*iweights[src]*w  per edge
 *   converged (:172-195) checked before every loop once iteration>=1: max|p-plast| < tol
 * Fixed summation order chosen here (the reference's is atomicAdd order, i.e. unspecified):
 * each product is formed in fp32 exactly as the lambda does ((plast*iw)*w), products of one
 * destination are accumulated in *double* together with the base term and rounded to fp32 once.
 * dsum is accumulated in double and rounded to fp32 once.  Returns the number of iterations run.
 * max_iter <= 0 means "no cap" as in the reference.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_pr(int n_vertices,
                   const int* ro,
                   const int* ci,
                   const float* w,
                   float alpha,
                   float tol,
                   int max_iter,
                   float* p) {
  size_t V = (size_t)n_vertices;
  float* plast = (float*)calloc(V ? V : 1, sizeof(float));
  float* iw = (float*)malloc((V ? V : 1) * sizeof(float));
  double* acc = (double*)malloc((V ? V : 1) * sizeof(double));
  for (size_t i = 0; i < V; ++i) p[i] = (float)(1.0 / (double)n_vertices);
  for (int i = 0; i < n_vertices; ++i) {
    volatile float val = 0.0f;
    for (int e = ro[i]; e < ro[i + 1]; ++e) val = val + (w ? w[e] : 1.0f);
    iw[i] = val != 0.0f ? alpha / val : 0.0f;
  }
  int iteration = 0;
  for (;;) {
    if (iteration > 0) {
      float err = 0.0f;
      for (size_t i = 0; i < V; ++i) {
        float d = fabsf(p[i] - plast[i]);
        if (d > err)
          err = d;
      }
      if (err < tol)
        break;
    }
    if (max_iter > 0 && iteration >= max_iter)
      break;
    memcpy(plast, p, V * sizeof(float));
    double dsum_d = 0.0;
    for (size_t i = 0; i < V; ++i)
      if (iw[i] == 0.0f) {
        volatile float t = alpha * p[i];
        dsum_d += (double)t;
      }
    float dsum = (float)dsum_d;
    volatile float base_num = (1 - alpha) + dsum; /* ((int)1 - alpha) is fp32, + dsum fp32 */
    volatile float base = base_num / (float)n_vertices;
    for (size_t i = 0; i < V; ++i) acc[i] = (double)base;
    for (int u = 0; u < n_vertices; ++u) {
      volatile float pu = plast[u] * iw[u];
      for (int e = ro[u]; e < ro[u + 1]; ++e) {
        volatile float upd = pu * (w ? w[e] : 1.0f);
        acc[ci[e]] += (double)upd;
      }
    }
    for (size_t i = 0; i < V; ++i) p[i] = (float)acc[i];
    ++iteration;
  }
  free(plast);
  free(iw);
  free(acc);
  return iteration;
}

/* ------------------------------------------------------------------------------------------
 * Workload definition (NOT reference code): counter-based RMAT / Graph500-style generator.
 * a=.57 b=.19 c=.19 d=.05, no per-level noise, 64-bit counter RNG.
 * Edge i, level l uses 16 random bits: chunk (l mod 4) of mix64(seed, i, l div 4).
 * Probabilities are quantised to 1/65536: A=37356 (0.57), A+B=49807 (0.76), A+B+C=62259 (0.95).
 * The CUDA generator (gunrock_b200/csrc) uses the same integer arithmetic -> identical edges.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t orc_mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

static inline uint64_t orc_hash3(uint64_t seed, uint64_t a, uint64_t b) {
  uint64_t h = orc_mix64(seed + 0x9E3779B97F4A7C15ull);
  h = orc_mix64(h ^ (a + 0x9E3779B97F4A7C15ull));
  h = orc_mix64(h ^ (b + 0x9E3779B97F4A7C15ull));
  return h;
}

ORC_API uint64_t orc_hash3_export(uint64_t seed, uint64_t a, uint64_t b) {
  return orc_hash3(seed, a, b);
}

ORC_API void orc_rmat_edges(int scale,
                            int64_t first_edge,
                            int64_t n_edges,
                            uint64_t seed,
                            int32_t* src,
                            int32_t* dst) {
  const uint32_t TA = 37356u, TAB = 49807u, TABC = 62259u;
  for (int64_t k = 0; k < n_edges; ++k) {
    uint64_t i = (uint64_t)(first_edge + k);
    uint32_t u = 0, v = 0;
    uint64_t h = 0;
    for (int l = 0; l < scale; ++l) {
      if ((l & 3) == 0)
        h = orc_hash3(seed, i, (uint64_t)(l >> 2));
      uint32_t r = (uint32_t)((h >> (16 * (l & 3))) & 0xFFFFu);
      uint32_t ub = (r >= TAB) ? 1u : 0u;                       /* c or d quadrant: row bit */
      uint32_t vb = ((r >= TA && r < TAB) || r >= TABC) ? 1u : 0u; /* b or d quadrant: col bit */
      u = (u << 1) | ub;
      v = (v << 1) | vb;
    }
    src[k] = (int32_t)u;
    dst[k] = (int32_t)v;
  }
}

/* Symmetric integer edge weight in 1..63 and a non-integer variant
 * 1 + 63*u01 with u01 = (h >> 40) * 2^-24 (exactly representable in fp32). */
ORC_API float orc_edge_weight(uint64_t seed, int32_t u, int32_t v, int non_integer) {
  uint64_t lo = (uint64_t)(u < v ? u : v), hi = (uint64_t)(u < v ? v : u);
  uint64_t h = orc_hash3(seed, lo, hi);
  if (!non_integer)
    return (float)(1 + (int)(h % 63ull));
  float u01 = (float)(h >> 40) * (1.0f / 16777216.0f);
  return 1.0f + 63.0f * u01;
}

ORC_API void orc_edge_weights(uint64_t seed,
                              int n_vertices,
                              const int* ro,
                              const int* ci,
                              int non_integer,
                              float* w) {
  for (int u = 0; u < n_vertices; ++u)
    for (int e = ro[u]; e < ro[u + 1]; ++e) w[e] = orc_edge_weight(seed, u, ci[e], non_integer);
}

/* ------------------------------------------------------------------------------------------
 * Host graph build used by the workload definition: drop self loops, optionally mirror,
 * sort by (u,v), dedup -> CSR.  LSD radix sort on 64-bit keys.
 * Returns nnz (or -1).  ro must hold n_vertices+1 ints, ci at least (mirror?2:1)*n ints.
 * ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_build_csr_from_pairs(int n_vertices,
                                         int64_t n,
                                         const int32_t* src,
                                         const int32_t* dst,
                                         int mirror,
                                         int* ro,
                                         int* ci) {
  size_t cap = (size_t)n * (mirror ? 2 : 1);
  uint64_t* keys = (uint64_t*)malloc((cap ? cap : 1) * sizeof(uint64_t));
  uint64_t* tmp = (uint64_t*)malloc((cap ? cap : 1) * sizeof(uint64_t));
  size_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u = (uint32_t)src[i], v = (uint32_t)dst[i];
    if (u == v)
      continue;
    keys[m++] = ((uint64_t)u << 32) | v;
    if (mirror)
      keys[m++] = ((uint64_t)v << 32) | u;
  }
  for (int pass = 0; pass < 8; ++pass) {
    size_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    int sh = pass * 8;
    for (size_t i = 0; i < m; ++i) ++cnt[((keys[i] >> sh) & 0xFF) + 1];
    int trivial = 0;
    for (int b = 0; b < 256; ++b)
      if (cnt[b + 1] == m)
        trivial = 1;
    if (trivial)
      continue;
    for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
    for (size_t i = 0; i < m; ++i) tmp[cnt[(keys[i] >> sh) & 0xFF]++] = keys[i];
    uint64_t* t = keys;
    keys = tmp;
    tmp = t;
  }
  memset(ro, 0, sizeof(int) * ((size_t)n_vertices + 1));
  int64_t nnz = 0;
  uint64_t prev = ~0ull;
  for (size_t i = 0; i < m; ++i) {
    if (keys[i] == prev)
      continue;
    prev = keys[i];
    if (nnz >= INT_MAX) {
      free(keys);
      free(tmp);
      return -1;
    }
    ci[nnz++] = (int)(keys[i] & 0xFFFFFFFFu);
    ++ro[(keys[i] >> 32) + 1];
  }
  for (int i = 0; i < n_vertices; ++i) ro[i + 1] += ro[i];
  free(keys);
  free(tmp);
  return nnz;
}


/* ------------------------------------------------------------------------------------------
 * Parallel host build of the same workload (OpenMP): identical output to
 * orc_rmat_edges + (optional fold) + orc_build_csr_from_pairs, for the bench's reference arm at
 * RMAT-26 where the serial build takes minutes.  Keys are generated in parallel, partitioned by
 * the top bits of the source id (MSD pass), every bucket is LSD-radix-sorted and deduplicated by
 * one thread, buckets are taken largest first.  Workload definition, NOT reference code.
 * ------------------------------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
#endif

static void orc_lsd_sort_u64(uint64_t* keys, uint64_t* tmp, size_t m, int key_bits) {
  uint64_t* a = keys;
  uint64_t* b = tmp;
  for (int sh = 0; sh < key_bits; sh += 8) {
    size_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    for (size_t i = 0; i < m; ++i) ++cnt[((a[i] >> sh) & 0xFF) + 1];
    int trivial = 0;
    for (int q = 0; q < 256; ++q)
      if (cnt[q + 1] == m)
        trivial = 1;
    if (trivial)
      continue;
    for (int q = 0; q < 256; ++q) cnt[q + 1] += cnt[q];
    for (size_t i = 0; i < m; ++i) b[cnt[(a[i] >> sh) & 0xFF]++] = a[i];
    uint64_t* t = a;
    a = b;
    b = t;
  }
  if (a != keys)
    memcpy(keys, a, m * sizeof(uint64_t));
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ro: n_vertices + 1 ints; ci: capacity ci_cap ints.  fold > 0: ids are taken modulo fold
 * (n_vertices must then equal fold).  Returns nnz, -1 if nnz > INT_MAX, -2 if ci_cap is too small. */
ORC_API int64_t orc_rmat_csr_parallel(int scale,
                                      int64_t n_pairs,
                                      uint64_t seed,
                                      int mirror,
                                      int fold,
                                      int n_vertices,
                                      int* ro,
                                      int* ci,
                                      int64_t ci_cap) {
  const size_t per = mirror ? 2 : 1;
  const size_t cap = (size_t)n_pairs * per;
  uint64_t* keys = (uint64_t*)malloc((cap ? cap : 1) * sizeof(uint64_t));
  uint64_t* tmp = (uint64_t*)malloc((cap ? cap : 1) * sizeof(uint64_t));
  if (!keys || !tmp) {
    free(keys);
    free(tmp);
    return -3;
  }
  /* vertex ids need vbits bits; buckets = top kB bits of the source id */
  int vbits = 1;
  while (((int64_t)1 << vbits) < (int64_t)n_vertices) ++vbits;
  const int kB = vbits > 12 ? 12 : vbits;
  const int nb = 1 << kB;
  const int bshift = 32 + vbits - kB; /* key >> bshift = bucket */
  const int64_t chunk = 1 << 16;
  const int64_t nchunks = (n_pairs + chunk - 1) / chunk;
  int nthreads = orc_num_threads();
  size_t* hist = (size_t*)calloc((size_t)nthreads * (nb + 1), sizeof(size_t));
  const uint64_t DEAD = ~0ull; /* self loop: sorts nowhere, dropped at the partition */
  /* 1. generate (static chunk ownership so that the histogram pass and the scatter pass agree) */
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    size_t* h = hist + (size_t)t * (nb + 1);
    int32_t* sbuf = (int32_t*)malloc(sizeof(int32_t) * chunk);
    int32_t* dbuf = (int32_t*)malloc(sizeof(int32_t) * chunk);
    for (int64_t c = t; c < nchunks; c += nthreads) {
      const int64_t first = c * chunk;
      const int64_t cnt = (first + chunk <= n_pairs) ? chunk : n_pairs - first;
      orc_rmat_edges(scale, first, cnt, seed, sbuf, dbuf);
      for (int64_t i = 0; i < cnt; ++i) {
        uint32_t u = (uint32_t)sbuf[i], v = (uint32_t)dbuf[i];
        if (fold > 0) {
          u %= (uint32_t)fold;
          v %= (uint32_t)fold;
        }
        uint64_t* out = keys + (size_t)(first + i) * per;
        if (u == v) {
          out[0] = DEAD;
          if (mirror)
            out[1] = DEAD;
          continue;
        }
        out[0] = ((uint64_t)u << 32) | v;
        ++h[out[0] >> bshift];
        if (mirror) {
          out[1] = ((uint64_t)v << 32) | u;
          ++h[out[1] >> bshift];
        }
      }
    }
    free(sbuf);
    free(dbuf);
  }
  /* 2. bucket offsets: bucket-major, thread-minor */
  size_t* bstart = (size_t*)malloc(sizeof(size_t) * (nb + 1));
  size_t run = 0;
  for (int q = 0; q < nb; ++q) {
    bstart[q] = run;
    for (int t = 0; t < nthreads; ++t) {
      size_t x = hist[(size_t)t * (nb + 1) + q];
      hist[(size_t)t * (nb + 1) + q] = run;
      run += x;
    }
  }
  bstart[nb] = run;
  /* 3. scatter into buckets (same chunk ownership) */
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    size_t* h = hist + (size_t)t * (nb + 1);
    for (int64_t c = t; c < nchunks; c += nthreads) {
      const size_t lo = (size_t)(c * chunk) * per;
      const size_t hi = (size_t)((c + 1) * chunk <= n_pairs ? (c + 1) * chunk : n_pairs) * per;
      for (size_t i = lo; i < hi; ++i)
        if (keys[i] != DEAD)
          tmp[h[keys[i] >> bshift]++] = keys[i];
    }
  }
  /* 4. per bucket: sort the low bits, dedup in place; order buckets by size (largest first) */
  int* order = (int*)malloc(sizeof(int) * nb);
  for (int q = 0; q < nb; ++q) order[q] = q;
  for (int i = 1; i < nb; ++i) { /* insertion sort on 4096 entries: negligible */
    int x = order[i];
    size_t sx = bstart[x + 1] - bstart[x];
    int j = i - 1;
    while (j >= 0 && bstart[order[j] + 1] - bstart[order[j]] < sx) {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = x;
  }
  size_t* uniq = (size_t*)calloc((size_t)nb + 1, sizeof(size_t));
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int k = 0; k < nb; ++k) {
    const int q = order[k];
    uint64_t* a = tmp + bstart[q];
    const size_t m = bstart[q + 1] - bstart[q];
    orc_lsd_sort_u64(a, keys + bstart[q], m, bshift); /* bits above bshift are equal inside a bucket */
    size_t w = 0;
    for (size_t i = 0; i < m; ++i)
      if (i == 0 || a[i] != a[i - 1])
        a[w++] = a[i];
    uniq[q + 1] = w;
  }
  for (int q = 0; q < nb; ++q) uniq[q + 1] += uniq[q];
  const size_t nnz = uniq[nb];
  int64_t rc = (int64_t)nnz;
  if (nnz > (size_t)INT_MAX)
    rc = -1;
  else if ((int64_t)nnz > ci_cap)
    rc = -2;
  if (rc >= 0) {
    memset(ro, 0, sizeof(int) * ((size_t)n_vertices + 1));
    /* 5. emit: a bucket owns a contiguous range of source ids, so row counts never collide */
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
    for (int q = 0; q < nb; ++q) {
      const uint64_t* a = tmp + bstart[q];
      const size_t w = uniq[q + 1] - uniq[q];
      int* out = ci + uniq[q];
      for (size_t i = 0; i < w; ++i) {
        out[i] = (int)(a[i] & 0xFFFFFFFFu);
        ++ro[(a[i] >> 32) + 1];
      }
    }
    for (int i = 0; i < n_vertices; ++i) ro[i + 1] += ro[i];
  }
  free(uniq);
  free(order);
  free(bstart);
  free(hist);
  free(keys);
  free(tmp);
  return rc;
}

/* Parallel edge weights (rows are independent). */
ORC_API void orc_edge_weights_parallel(uint64_t seed, int n_vertices, const int* ro, const int* ci,
                                       int non_integer, float* w) {
#pragma omp parallel for schedule(dynamic, 4096)
  for (int u = 0; u < n_vertices; ++u)
    for (int e = ro[u]; e < ro[u + 1]; ++e) w[e] = orc_edge_weight(seed, u, ci[e], non_integer);
}
