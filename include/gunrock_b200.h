/*
 * gunrock_b200.h -- C ABI of the B200-native frontier engine (libgunrock_b200.so).
 *
 * Gunrock itself has no C ABI / plugin boundary: its public surface is the header-only C++17
 * template API.  That surface is re-implemented source-compatibly under
 * include/gunrock/ (so examples/algorithms/{bfs,sssp,pr} compile unchanged).  THIS header is the
 * seam *below* the user lambda: the three named algorithms with their fixed edge functors, the
 * operators they are built from, and the ingest steps either side -- plain pointers and sizes,
 * no C++ / torch types -- which is what a language binding (ctypes, cgo, JNI, the reference's own
 * nanobind module python/src/gunrock/bindings.cu:186-266) would bind.  INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions
 *   - vertex_t = edge_t = int32, weight_t = float32 (examples/algorithms/bfs/bfs.cu:15-17).
 *   - every function returns 0 on success, a negative b2g error or a positive cudaError_t;
 *     b2g_last_error() returns a thread-local message for the last failure.
 *   - `*_loc` arguments say where a caller buffer lives: B2G_HOST or B2G_DEVICE.
 *   - a graph handle owns (or views) device-resident CSR arrays; calls on one handle must not
 *     overlap in time (one stream per handle), matching the reference where run() is synchronous
 *     (include/gunrock/framework/enactor.hxx:266-288).
 *   - The library never falls back to a CPU path: without a CUDA device every compute entry
 *     point fails with B2G_ERR_NO_DEVICE.
 */
#ifndef GUNROCK_B200_H_
#define GUNROCK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define B2G_HOST 0
#define B2G_DEVICE 1

#define B2G_ERR_INVALID (-1)
#define B2G_ERR_NO_DEVICE (-2)
#define B2G_ERR_OVERFLOW (-3)
#define B2G_ERR_INTERNAL (-4)

/* operators::load_balance_t (include/gunrock/framework/operators/configs.hxx:52-60) */
#define B2G_LB_THREAD_MAPPED 0
#define B2G_LB_BLOCK_MAPPED 2
#define B2G_LB_MERGE_PATH 4
/* operators::advance_direction_t (configs.hxx:78-82) */
#define B2G_DIR_FORWARD 0   /* push  */
#define B2G_DIR_BACKWARD 1  /* pull  */
#define B2G_DIR_OPTIMIZED 2 /* push/pull switch */
/* operators::filter_algorithm_t (configs.hxx:92-97) */
#define B2G_FILTER_REMOVE 0
#define B2G_FILTER_PREDICATED 1
#define B2G_FILTER_COMPACT 2
#define B2G_FILTER_BYPASS 3

typedef struct b2g_graph b2g_graph_t; /* opaque */

/* Mirrors gunrock::options_t (include/gunrock/algorithms/algorithms.hxx:27-72) plus the knobs the
 * reference hard-codes or lacks.  Zero-initialise then call b2g_options_default(). */
typedef struct b2g_options {
  int advance_load_balance; /* B2G_LB_*          default block_mapped (algorithms.hxx:29-30) */
  int filter_algorithm;     /* B2G_FILTER_*      default predicated   (algorithms.hxx:33-34) */
  int enable_filter;        /* ignored by the fused enactors: their advance output is compact */
  int enable_uniquify;
  int best_effort_uniquify;
  float uniquify_percent;
  int advance_direction;    /* B2G_DIR_*         default forward (the reference has push only) */
  int hub_threshold;        /* rows >= this many edges go to the TMA slab bin (block_mapped) */
  int ctas_per_sm;          /* persistent grid = #SM x this */
  int reference_functor;    /* 1: BFS uses the reference's per-edge atomicMin lambda */
  float do_alpha, do_beta;  /* Beamer switch parameters for B2G_DIR_OPTIMIZED */
  void* stream;             /* cudaStream_t to run on (NULL = the handle's own stream) */
} b2g_options_t;

/* Per-run report.  levels[] arrays are filled up to n_levels (capped at B2G_MAX_LEVELS). */
#define B2G_MAX_LEVELS 4096
typedef struct b2g_stats {
  float elapsed_ms;            /* CUDA events around the enactor loop, reference timed region */
  int iterations;              /* BSP iterations executed (enactor_t::iteration) */
  int kernel_launches;         /* kernels this library launched inside the timed region */
  unsigned long long edges_touched; /* column indices read (MTEPS numerator) */
  unsigned long long vertices_touched;
  int n_levels;
  int level_direction[64];     /* first 64 levels: 0 push, 1 pull */
  int level_frontier[64];
  unsigned long long level_edges[64];
  float level_kernel_ms[64];   /* device time of the level's advance / sweep kernel(s) (CUDA events) */
} b2g_stats_t;

int b2g_version(void);
const char* b2g_last_error(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int b2g_device_count(void);
void b2g_options_default(b2g_options_t* o);

/* ---- graph ingest -------------------------------------------------------------------------
 * Replaces format::csr_t<device>::from_coo + graph::build<device>
 * (include/gunrock/formats/csr.hxx:81-140, include/gunrock/graph/build.hxx:29-36). */

/* CSR arrays (row_offsets[V+1], column_indices[E], values[E] or NULL => 1.0f) are COPIED to the
 * device from host or device memory. `symmetric`!=0 lets pull-mode reuse the CSR as its own CSC. */
int b2g_graph_create_csr(int n_vertices, int n_edges, const int* row_offsets,
                         const int* column_indices, const float* values, int loc, int symmetric,
                         b2g_graph_t** out);
/* Non-owning view over caller-owned DEVICE arrays (graph::graph_t semantics, graph/graph.hxx:189-196). */
int b2g_graph_view_csr(int n_vertices, int n_edges, const int* row_offsets,
                       const int* column_indices, const float* values, int symmetric,
                       b2g_graph_t** out);
/* COO (host, n entries, row-major stable order as csr.hxx:81-140) -> device CSR. */
int b2g_graph_create_coo(int n_rows, int n_cols, int nnz, const int* I, const int* J,
                         const float* V, int symmetric, b2g_graph_t** out);
/* Synthetic workload: counter-based RMAT pairs generated on the device, self loops
 * dropped, optional mirroring, sorted + deduplicated into CSR.  weights: 0 none, 1 integer 1..63,
 * 2 non-integer 1+63*u01 (symmetric hash of the endpoints, weight_seed). */
int b2g_graph_create_rmat(int scale, long long n_pairs, unsigned long long seed, int mirror,
                          int fold_vertices /* 0 = 2^scale */, int weights,
                          unsigned long long weight_seed, b2g_graph_t** out);
/* Build the transpose (CSC) on the device for pull traversal of non-symmetric graphs
 * (replaces format::csc_t::from_csr, include/gunrock/formats/csc.hxx:62-102). Idempotent. */
int b2g_graph_build_transpose(b2g_graph_t* g);
int b2g_graph_destroy(b2g_graph_t* g);
int b2g_graph_info(const b2g_graph_t* g, int* n_vertices, int* n_edges, int* has_values,
                   int* symmetric);
/* Device pointers of the resident CSR (read-only; lifetime = the handle). */
int b2g_graph_device_ptrs(const b2g_graph_t* g, const int** row_offsets,
                          const int** column_indices, const float** values);
/* Copy the resident CSR back (any pointer may be NULL). */
int b2g_graph_download(const b2g_graph_t* g, int* row_offsets, int* column_indices, float* values);
/* Vertex of maximum out-degree (lowest id on ties) -- the bench source rule. */
int b2g_graph_max_degree_vertex(const b2g_graph_t* g, int* vertex, int* degree);

/* ---- host-side ingest (no device needed): the file formats either side of the path ------------
 * io::matrix_market_t::load (include/gunrock/io/matrix_market.hxx:99-254): coordinate files, pattern /
 * real / integer, general / symmetric; 1-based -> 0-based, pattern => weight 1.0, a symmetric file's
 * off-diagonal entries are followed in place by their mirror.  A clean body is parsed by all host
 * threads, results identical to the reference's loader.  *I, *J, *V are allocated by the library
 * (*nnz entries each, mirrors included) and released with b2g_host_free.  Bad files return
 * B2G_ERR_INVALID with the reference's message in b2g_last_error() instead of exiting the process. */
int b2g_mtx_load(const char* path, int* n_rows, int* n_cols, int* nnz, int** I, int** J, float** V,
                 int* directed, int* weighted, int* symmetric);
void b2g_host_free(void* p);
/* format::csr_t<host>::from_coo (include/gunrock/formats/csr.hxx:81-140) on host arrays: stable counting
 * sort by row with all host threads, duplicates and self loops kept.  Outputs are caller-allocated:
 * row_offsets[n_rows+1], column_indices[nnz], values[nnz] (V == NULL => values may be NULL). */
int b2g_csr_from_coo_host(int n_rows, int nnz, const int* I, const int* J, const float* V,
                          int* row_offsets, int* column_indices, float* values);

/* ---- algorithms (fused enactors) ----------------------------------------------------------- */

/* gunrock::bfs::run (include/gunrock/algorithms/bfs.hxx:162-182): distances[V] int32, INT_MAX if
 * unreachable.  distances lives at dist_loc (host => D2H copy inside the call). */
int b2g_bfs(b2g_graph_t* g, int source, const b2g_options_t* opt, int* distances, int dist_loc,
            b2g_stats_t* stats);
/* gunrock::sssp::run (include/gunrock/algorithms/sssp.hxx:176-198): fp32 distances, FLT_MAX if
 * unreachable.  Requires edge values. */
int b2g_sssp(b2g_graph_t* g, int source, const b2g_options_t* opt, float* distances, int dist_loc,
             b2g_stats_t* stats);
/* gunrock::pr::run (include/gunrock/algorithms/pr.hxx:211-236): alpha damping, tol on
 * max|p - plast|; max_iter <= 0 = no cap (as the reference). Pull over the transpose. */
int b2g_pr(b2g_graph_t* g, float alpha, float tol, int max_iter, const b2g_options_t* opt,
           float* p, int p_loc, b2g_stats_t* stats);

/* ---- operators with the fixed functors (for operator-level parity tests and bindings) ------
 * Frontiers are DEVICE int32 arrays plus a DEVICE int32 element count.                         */

/* operators::advance::execute (advance/advance.hxx:94-133) with the BFS claim functor: out
 * receives the newly claimed neighbours, compacted.  `labels` (V ints) gets `label` written for
 * each claimed vertex, `visited_bitmap` (ceil(V/32) words) is updated. */
int b2g_advance_bfs(b2g_graph_t* g, const int* in, const int* in_count, int in_capacity, int* out,
                    int* out_count, int out_capacity, unsigned* visited_bitmap, int* labels,
                    int label, const b2g_options_t* opt, unsigned long long* edges_touched);
/* operators::filter::execute (filter/filter.hxx:72-100), predicate "vertex is valid and
 * keep_mask[vertex] != 0" (keep_mask may be NULL = keep all valid).  alg = B2G_FILTER_*;
 * bypass keeps the size and writes -1 for rejected entries. */
int b2g_filter(b2g_graph_t* g, int alg, const int* in, const int* in_count, int in_capacity,
               int* out, int* out_count, const unsigned char* keep_mask);
/* operators::uniquify::execute (uniquify/uniquify.hxx:26-94).  best_effort != 0: adjacent
 * duplicates only (no sort); else exact: the sorted unique set (bitmap enumeration). */
int b2g_uniquify(b2g_graph_t* g, const int* in, const int* in_count, int in_capacity, int* out,
                 int* out_count, int best_effort);

/* ---- multi-GPU BFS: per-rank steps of the 1-D (cyclic) vertex partition -----------------------
 * One process per GPU.  Global vertex v lives on rank v % nparts at local row v / nparts; a rank's
 * CSR has its own rows with GLOBAL column ids.  The per-level remote-frontier exchange (NCCL
 * all-to-all of the send buffers, all-gather of the frontier bitmap, all-reduce of the counts) is
 * done by the host side between these calls (gunrock_b200/multi_gpu.py).  The reference has no
 * multi-GPU execution (advance.hxx:129-132 throws); it only offers gcuda::multi_context_t
 * (include/gunrock/cuda/context.hxx:146-216), whose role the process group plays here. */
int b2g_graph_create_rmat_part(int scale, long long n_pairs, unsigned long long seed, int mirror,
                               int nparts, int part, b2g_graph_t** out);
/* Rank-local CSR (row_offsets[n_local+1], column_indices = global ids), host or device arrays. */
int b2g_graph_create_csr_part(int n_global_vertices, int nparts, int part, int n_local_edges,
                              const int* row_offsets, const int* column_indices, int loc,
                              int symmetric, b2g_graph_t** out);
int b2g_part_info(const b2g_graph_t* g, int* n_global, int* nparts, int* part, int* n_local,
                  int* words_per_rank);
/* Reset the rank's BFS state and seed `source` on its owner. send_capacity = ints per peer. */
int b2g_part_bfs_begin(b2g_graph_t* g, int source, int send_capacity);
/* Expand the local frontier top-down.  Local neighbours are claimed in place; new remote ones land
 * in the per-owner send buffers.  send_counts (host, nparts ints) receives how many per owner. */
int b2g_part_bfs_topdown(b2g_graph_t* g, int level, const b2g_options_t* opt, int* send_counts,
                         unsigned long long* edges_touched);
/* Device pointer to the send buffers: nparts rows of *send_capacity ints. */
int b2g_part_bfs_send_buffer(b2g_graph_t* g, int** send_buf, int* send_capacity);
/* Claim n_recv global ids (device array, all owned by this rank) received from peers. */
int b2g_part_bfs_claim(b2g_graph_t* g, int level, const int* recv, int n_recv);
/* Write the current local frontier as a bitmap of words_per_rank words (device). */
int b2g_part_bfs_frontier_bitmap(b2g_graph_t* g, unsigned* out);
/* Bottom-up sweep against the all-gathered frontier bitmap (nparts x words_per_rank words). */
int b2g_part_bfs_bottomup(b2g_graph_t* g, int level, const unsigned* frontier_all,
                          unsigned long long* edges_touched);
/* Close the level: returns this rank's next-frontier size and its out-degree sum. */
int b2g_part_bfs_end_level(b2g_graph_t* g, long long* n_frontier, long long* frontier_degree);
/* --- sync-free variants: everything is enqueued on `stream` (e.g. torch's current stream, so the
 * NCCL collectives between the calls are stream ordered) and nothing is read back by the host;
 * the only host synchronisation of a level is the caller's read of the all-reduced statistics. */
int b2g_part_set_stream(b2g_graph_t* g, void* stream);
/* top-down + pack: msg (device) = nparts rows of [count, ids ...] with cap_s id slots each. */
int b2g_part_bfs_topdown_async(b2g_graph_t* g, int level, const b2g_options_t* opt, int* msg,
                               int cap_s);
/* claim every peer's packed row (msgs = nparts rows of cap_s+1 ints, as delivered by a fixed-split
 * all-to-all).  A row whose count exceeds cap_s sets the level's overflow statistic. */
int b2g_part_bfs_claim_packed_async(b2g_graph_t* g, int level, const int* msgs, int cap_s);
int b2g_part_bfs_frontier_bitmap_async(b2g_graph_t* g, unsigned* out);
int b2g_part_bfs_bottomup_async(b2g_graph_t* g, int level, const unsigned* frontier_all);
/* stats (device, int64[4]) = {next frontier size, its out-degree sum, edges inspected, overflow}. */
int b2g_part_bfs_end_level_async(b2g_graph_t* g, long long* stats);

/* ---- peer-memory exchange over NVLink (include/gunrock/b200/bfs_p2p.cuh) ---------------------------
 * The kernels themselves write forwarded ids / next-frontier words / statistics into windows of device
 * memory that every rank maps; no NCCL call and no host code between the phases of a level.  Stands
 * where gcuda::multi_context_t (cuda/context.hxx:146-216) would drive several devices.
 *  1. b2g_part_p2p_window_create: allocate this rank's window; returns its device pointer, size and
 *     (ipc_handle != NULL) the 64-byte CUDA IPC handle to hand to the other processes;
 *  2. b2g_part_p2p_attach: ipc_handles = nparts x 64 bytes gathered from all ranks (processes), or
 *     windows = nparts device pointers (ranks simulated inside one process);
 *  3. b2g_part_bfs_p2p: COLLECTIVE -- every rank calls it with the same source / options; returns when
 *     the traversal is complete (owned distances: b2g_part_bfs_distances).  total_edges = global
 *     directed edge count (direction heuristic).  Bottom-up levels need a symmetric graph.
 * Both calls are idempotent per graph handle (one window, one set of mappings, one shared epoch).
 * At most 16 ranks.  A peer that never arrives raises an error after 20 s instead of hanging. */
int b2g_part_p2p_window_create(b2g_graph_t* g, void** window, unsigned long long* bytes,
                               unsigned char* ipc_handle);
int b2g_part_p2p_attach(b2g_graph_t* g, const unsigned char* ipc_handles, void* const* windows);
/* Unmap the peers' windows (every rank calls it, then a barrier, before any rank destroys its graph:
 * a window must not be freed while another process still maps it).  The own window stays. */
int b2g_part_p2p_detach(b2g_graph_t* g);
int b2g_part_bfs_p2p(b2g_graph_t* g, int source, long long total_edges, const b2g_options_t* opt,
                     b2g_stats_t* stats);
/* ---- NCCL exchange driven from C++ (include/gunrock/b200/bfs_nccl.cuh; SURVEY.md 8e: "NCCL all-to-all =
 * ncclGroupStart; ncclSend / ncclRecv per peer; ncclGroupEnd", ncclAllGather of the frontier bitmap,
 * ncclAllReduce of the level statistics).  The whole level loop runs inside b2g_part_bfs_nccl: no Python and no
 * stream synchronisation inside a level, one pinned-memory poll per level.  NCCL is bound at run time (dlopen of
 * libnccl.so.2: the copy torch loaded in a torch.distributed process, the system one otherwise).
 *  1. rank 0: b2g_nccl_unique_id(id) -> 128 bytes, broadcast by the caller (torch.distributed, MPI, a file ...);
 *  2. every rank: b2g_part_nccl_init(g, id, nranks, rank) -- COLLECTIVE (ncclCommInitRank) -- allocates the
 *     message buffers once;
 *  3. b2g_part_bfs_nccl: COLLECTIVE, same arguments and statistics as b2g_part_bfs_p2p;
 *  4. b2g_part_nccl_finalize (optional; the handle's destructor does it too). */
int b2g_nccl_unique_id(unsigned char* id128);
int b2g_part_nccl_init(b2g_graph_t* g, const unsigned char* id128, int nranks, int rank);
int b2g_part_bfs_nccl(b2g_graph_t* g, int source, long long total_edges, const b2g_options_t* opt,
                      b2g_stats_t* stats);
int b2g_part_nccl_finalize(b2g_graph_t* g);
/* SSSP and PageRank whose iteration loop and NCCL collectives run inside the call (after b2g_part_nccl_init on a
 * graph of the matching kind: edge values for SSSP, rows = in-edge lists for PageRank).  COLLECTIVE.
 *  b2g_part_sssp_nccl: per iteration relax -> grouped Send/Recv of the (vertex, distance) rows, sized from the
 *    frontier's all-reduced out-degree sum -> apply -> all-reduce of 4 x int64; send_capacity 0 = default; a row
 *    that overflows restarts the run with rows four times as long.  Result: b2g_part_sssp_distances.
 *  b2g_part_pr_nccl: all-reduce of the out-degrees (or fp64 row sums of the weights) once, then per iteration
 *    prepare -> ncclAllGather(c) + all-reduce(dangling, fp64 sum) -> pull -> all-reduce(error, max); same
 *    recurrence and stopping rule as algorithms/pr.hxx:107-195.  Result: b2g_part_pr_ranks, stats->iterations. */
int b2g_part_sssp_nccl(b2g_graph_t* g, int source, int send_capacity, const b2g_options_t* opt,
                       b2g_stats_t* stats);
int b2g_part_pr_nccl(b2g_graph_t* g, float alpha, float tol, int max_iter, b2g_stats_t* stats);
/* ---- multi-GPU PageRank (pull): the rank owns the DESTINATION vertices v % nparts == part and their
 * in-edges (a partitioned graph whose rows are in-edge lists: any symmetric partitioned graph, or
 * one created with by_destination != 0).  Per iteration the host side all-gathers c = plast*iweights,
 * all-reduces the dangling sum (fp64, sum) and the error (fp32, max) between these calls.  A graph with
 * edge values takes b2g_part_pr_outweights + b2g_part_pr_begin_weighted instead of the out-degree pair (its
 * in-edge rows must carry the weight of each in-edge: partition the transpose WITH its values).  Everything is
 * enqueued on the stream set with b2g_part_set_stream. */
int b2g_graph_create_rmat_part_ex(int scale, long long n_pairs, unsigned long long seed, int mirror,
                                  int fold_vertices, int by_destination, int nparts, int part,
                                  b2g_graph_t** out);
/* Count out-degrees of the local in-edges into outdeg (device, n_global ints, zeroed by the call). */
int b2g_part_pr_outdegrees(b2g_graph_t* g, int* outdeg);
/* Reset ranks and derive iweights from the ALL-REDUCED out-degrees (device, n_global ints). */
int b2g_part_pr_begin(b2g_graph_t* g, float alpha, const int* outdeg_global);
/* Weighted graphs: fp64 sums of the weights of the local in-edges per SOURCE vertex into outweight (device,
 * n_global doubles, zeroed by the call); after the all-reduce, iweights = alpha / (float)sum.  The reference and
 * the single-GPU path add a row's weights sequentially in fp32 (pr.hxx:65-93); a vertex's out-edges are spread
 * over the ranks here, so the sum is fp64 and rounded once (last-place differences, inside the 1e-6 tolerance). */
int b2g_part_pr_outweights(b2g_graph_t* g, double* outweight);
int b2g_part_pr_begin_weighted(b2g_graph_t* g, float alpha, const double* outweight_global);
/* plast = p, c_local = plast*iweights (device, rows_per_rank floats, zero padded), dsum_local (device
 * double) = this rank's dangling partial. */
int b2g_part_pr_prepare(b2g_graph_t* g, float alpha, float* c_local, double* dsum_local);
/* One pull over the owned rows: c_all = all-gathered c (nparts*rows_per_rank floats), dsum_global =
 * all-reduced dangling sum (device double), err_local (device float) = max |p - plast| of this rank. */
int b2g_part_pr_pull(b2g_graph_t* g, float alpha, const float* c_all, const double* dsum_global,
                     float* err_local);
/* Copy the owned ranks (n_local floats, local row order). */
int b2g_part_pr_ranks(b2g_graph_t* g, float* p, int loc);
/* ---- multi-GPU SSSP: the same push exchange carrying (vertex, fp32 distance) pairs; rows of the packed
 * message are [count, ids[cap_s], distance bits[cap_s]] (2*cap_s+1 ints).  Needs edge values. */
int b2g_graph_create_csr_part_weighted(int n_global_vertices, int nparts, int part, int n_local_edges,
                                       const int* row_offsets, const int* column_indices,
                                       const float* values, int loc, int symmetric, b2g_graph_t** out);
int b2g_part_sssp_begin(b2g_graph_t* g, int source, int send_capacity);
int b2g_part_sssp_relax_async(b2g_graph_t* g, int iteration, const b2g_options_t* opt, int* msg,
                              int cap_s);
int b2g_part_sssp_apply_packed_async(b2g_graph_t* g, int iteration, const int* msgs, int cap_s);
/* stats (device int64[4]) = {next frontier size, its out-degree sum, edges relaxed, overflow}. */
int b2g_part_sssp_end_iteration_async(b2g_graph_t* g, long long* stats);
int b2g_part_sssp_distances(b2g_graph_t* g, float* distances, int loc);
/* Copy the rank's slice of the distances (n_local ints, local row order). */
int b2g_part_bfs_distances(b2g_graph_t* g, int* distances, int loc);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* GUNROCK_B200_H_ */
