"""gunrock_b200 -- host-side mirror of the Gunrock interface over the B200-native C ABI.

The compute lives in ``libgunrock_b200.so`` (hand-written sm_100a kernels, ``include/gunrock/b200``,
exported through ``include/gunrock_b200.h``).  This module only binds it with ``ctypes`` and gives it
the shape of the reference's Python surface (``python/src/gunrock/bindings.cu:186-266``:
``gunrock.bfs(G, src, distances, predecessors, ctx, options)`` on torch tensors via ``data_ptr()``).
PyTorch is plumbing here: device memory, streams, ``torch.distributed``.

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, the
compute entry points raise.  (The CPU checker lives in ``oracle/`` and is test infrastructure.)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2G_LIB_PATH: an A/B build of the same library (e.g. other launch bounds); never a fallback -- it must exist
LIB_PATH = os.environ.get("B2G_LIB_PATH") or os.path.join(_HERE, "libgunrock_b200.so")

HOST, DEVICE = 0, 1
INT_MAX = 2**31 - 1
FLT_MAX = float(np.finfo(np.float32).max)


class load_balance_t:  # operators::load_balance_t, framework/operators/configs.hxx:52-60
    thread_mapped, warp_mapped, block_mapped, bucketing, merge_path, merge_path_v2, work_stealing = range(7)


class advance_direction_t:  # configs.hxx:78-82
    forward, backward, optimized = range(3)


class filter_algorithm_t:  # configs.hxx:92-97
    remove, predicated, compact, bypass = range(4)


class _Options(C.Structure):
    _fields_ = [("advance_load_balance", C.c_int), ("filter_algorithm", C.c_int),
                ("enable_filter", C.c_int), ("enable_uniquify", C.c_int),
                ("best_effort_uniquify", C.c_int), ("uniquify_percent", C.c_float),
                ("advance_direction", C.c_int), ("hub_threshold", C.c_int),
                ("ctas_per_sm", C.c_int), ("reference_functor", C.c_int),
                ("do_alpha", C.c_float), ("do_beta", C.c_float), ("stream", C.c_void_p)]


class _Stats(C.Structure):
    _fields_ = [("elapsed_ms", C.c_float), ("iterations", C.c_int), ("kernel_launches", C.c_int),
                ("edges_touched", C.c_ulonglong), ("vertices_touched", C.c_ulonglong),
                ("n_levels", C.c_int), ("level_direction", C.c_int * 64),
                ("level_frontier", C.c_int * 64), ("level_edges", C.c_ulonglong * 64),
                ("level_kernel_ms", C.c_float * 64)]


@dataclass
class options_t:
    """gunrock::options_t (include/gunrock/algorithms/algorithms.hxx:27-72) + B200 knobs."""
    advance_load_balance: int = load_balance_t.block_mapped
    filter_algorithm: int = filter_algorithm_t.predicated
    enable_filter: bool = False
    enable_uniquify: bool = False
    best_effort_uniquify: bool = True
    uniquify_percent: float = 100.0
    advance_direction: int = advance_direction_t.forward
    hub_threshold: int = 4096
    ctas_per_sm: int = 8
    reference_functor: bool = False
    do_alpha: float = 14.0
    do_beta: float = 24.0
    stream: Optional[int] = None  # raw cudaStream_t

    def _c(self) -> _Options:
        return _Options(self.advance_load_balance, self.filter_algorithm, int(self.enable_filter),
                        int(self.enable_uniquify), int(self.best_effort_uniquify),
                        self.uniquify_percent, self.advance_direction, self.hub_threshold,
                        self.ctas_per_sm, int(self.reference_functor), self.do_alpha, self.do_beta,
                        self.stream)


@dataclass
class stats_t:
    elapsed_ms: float = 0.0
    iterations: int = 0
    kernel_launches: int = 0
    edges_touched: int = 0
    vertices_touched: int = 0
    level_direction: list = field(default_factory=list)
    level_frontier: list = field(default_factory=list)
    level_edges: list = field(default_factory=list)
    level_kernel_ms: list = field(default_factory=list)


class GunrockB200Error(RuntimeError):
    pass


_lib = None

_SYMBOLS = [
    "b2g_version", "b2g_last_error", "b2g_device_count", "b2g_options_default",
    "b2g_graph_create_csr", "b2g_graph_view_csr", "b2g_graph_create_coo", "b2g_graph_create_rmat",
    "b2g_graph_build_transpose", "b2g_graph_destroy", "b2g_graph_info", "b2g_graph_device_ptrs",
    "b2g_graph_download", "b2g_graph_max_degree_vertex", "b2g_bfs", "b2g_sssp", "b2g_pr",
    "b2g_advance_bfs", "b2g_filter", "b2g_uniquify",
    # host-side ingest (no device involved)
    "b2g_mtx_load", "b2g_host_free", "b2g_csr_from_coo_host",
    # multi-GPU per-rank steps (bound in gunrock_b200/multi_gpu.py)
    "b2g_graph_create_rmat_part", "b2g_graph_create_csr_part", "b2g_part_info", "b2g_part_bfs_begin",
    "b2g_part_bfs_topdown", "b2g_part_bfs_send_buffer", "b2g_part_bfs_claim",
    "b2g_part_bfs_frontier_bitmap", "b2g_part_bfs_bottomup", "b2g_part_bfs_end_level",
    "b2g_part_bfs_distances", "b2g_part_set_stream", "b2g_part_bfs_topdown_async",
    "b2g_part_bfs_claim_packed_async", "b2g_part_bfs_frontier_bitmap_async",
    "b2g_part_bfs_bottomup_async", "b2g_part_bfs_end_level_async",
    "b2g_graph_create_rmat_part_ex", "b2g_part_pr_outdegrees", "b2g_part_pr_begin", "b2g_part_pr_prepare",
    "b2g_part_pr_pull", "b2g_part_pr_ranks",
    "b2g_graph_create_csr_part_weighted", "b2g_part_sssp_begin", "b2g_part_sssp_relax_async",
    "b2g_part_sssp_apply_packed_async", "b2g_part_sssp_end_iteration_async", "b2g_part_sssp_distances",
    # peer-memory (NVLink) exchange
    "b2g_part_p2p_window_create", "b2g_part_p2p_attach", "b2g_part_p2p_detach", "b2g_part_bfs_p2p",
    # NCCL exchange driven from C++
    "b2g_nccl_unique_id", "b2g_part_nccl_init", "b2g_part_bfs_nccl", "b2g_part_nccl_finalize",
    "b2g_part_sssp_nccl", "b2g_part_pr_nccl", "b2g_part_pr_outweights", "b2g_part_pr_begin_weighted",
]


def lib() -> C.CDLL:
    """Load libgunrock_b200.so (built in-tree by ``__graft_entry__.build()``); never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GunrockB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C gunrock_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, ip, u64 = C.c_void_p, C.c_int, C.c_ulonglong
        L.b2g_last_error.restype = C.c_char_p
        L.b2g_options_default.argtypes = [C.POINTER(_Options)]
        L.b2g_graph_create_csr.argtypes = [ip, ip, vp, vp, vp, ip, ip, C.POINTER(vp)]
        L.b2g_graph_view_csr.argtypes = [ip, ip, vp, vp, vp, ip, C.POINTER(vp)]
        L.b2g_graph_create_coo.argtypes = [ip, ip, ip, vp, vp, vp, ip, C.POINTER(vp)]
        L.b2g_graph_create_rmat.argtypes = [ip, C.c_longlong, u64, ip, ip, ip, u64, C.POINTER(vp)]
        L.b2g_graph_build_transpose.argtypes = [vp]
        L.b2g_graph_destroy.argtypes = [vp]
        L.b2g_graph_info.argtypes = [vp] + [C.POINTER(ip)] * 4
        L.b2g_graph_device_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 3
        L.b2g_graph_download.argtypes = [vp, vp, vp, vp]
        L.b2g_graph_max_degree_vertex.argtypes = [vp, C.POINTER(ip), C.POINTER(ip)]
        L.b2g_bfs.argtypes = [vp, ip, C.POINTER(_Options), vp, ip, C.POINTER(_Stats)]
        L.b2g_sssp.argtypes = [vp, ip, C.POINTER(_Options), vp, ip, C.POINTER(_Stats)]
        L.b2g_pr.argtypes = [vp, C.c_float, C.c_float, ip, C.POINTER(_Options), vp, ip,
                             C.POINTER(_Stats)]
        L.b2g_advance_bfs.argtypes = [vp, vp, vp, ip, vp, vp, ip, vp, vp, ip, C.POINTER(_Options),
                                      C.POINTER(u64)]
        L.b2g_filter.argtypes = [vp, ip, vp, vp, ip, vp, vp, vp]
        L.b2g_uniquify.argtypes = [vp, vp, vp, ip, vp, vp, ip]
        L.b2g_mtx_load.argtypes = [C.c_char_p] + [C.POINTER(ip)] * 3 + [C.POINTER(vp)] * 3 + [C.POINTER(ip)] * 3
        L.b2g_host_free.argtypes = [vp]
        L.b2g_host_free.restype = None
        L.b2g_csr_from_coo_host.argtypes = [ip, ip, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def exported_symbols() -> list:
    """Every entry point include/gunrock_b200.h declares (used by the CPU-side ABI test)."""
    return list(_SYMBOLS)


def device_count() -> int:
    return int(lib().b2g_device_count())


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b2g_last_error()
        raise GunrockB200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def _ptr(x):
    """Raw address + location of a numpy array / torch tensor (None -> NULL)."""
    if x is None:
        return None, HOST
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data, HOST
    # torch tensor
    if not x.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return x.data_ptr(), (DEVICE if x.is_cuda else HOST)


def _stats_out(s: _Stats) -> stats_t:
    n = min(int(s.n_levels), 64)
    return stats_t(float(s.elapsed_ms), int(s.iterations), int(s.kernel_launches),
                   int(s.edges_touched), int(s.vertices_touched),
                   list(s.level_direction[:n]), list(s.level_frontier[:n]),
                   list(s.level_edges[:n]), list(s.level_kernel_ms[:n]))


class graph_t:
    """Device-resident CSR graph (graph::graph_t + format::csr_t<device>)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        nv, ne, hv, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().b2g_graph_info(self._h, C.byref(nv), C.byref(ne), C.byref(hv), C.byref(sym)),
               "b2g_graph_info")
        self.n_vertices, self.n_edges = nv.value, ne.value
        self.weighted, self.symmetric = bool(hv.value), bool(sym.value)
        self._keep = None

    # -- constructors ---------------------------------------------------------------------
    @staticmethod
    def from_csr(row_offsets, column_indices, values=None, symmetric: bool = False) -> "graph_t":
        """Copy CSR arrays (numpy / torch, host or device) to the device."""
        ro_p, loc = _ptr(row_offsets)
        ci_p, loc2 = _ptr(column_indices)
        v_p, _ = _ptr(values)
        n_v = int(row_offsets.shape[0]) - 1
        n_e = int(column_indices.shape[0])
        if n_e and loc2 != loc:
            raise ValueError("row_offsets and column_indices must live in the same memory space")
        for a, dt in ((row_offsets, "int32"), (column_indices, "int32")):
            if "int32" not in str(a.dtype):
                raise TypeError(f"CSR index arrays must be int32, got {a.dtype}")
        if values is not None and "float32" not in str(values.dtype):
            raise TypeError("values must be float32")
        h = C.c_void_p()
        _check(lib().b2g_graph_create_csr(n_v, n_e, ro_p, ci_p, v_p, loc, int(symmetric), C.byref(h)),
               "b2g_graph_create_csr")
        return graph_t(h.value)

    @staticmethod
    def view_csr(row_offsets, column_indices, values=None, symmetric: bool = False) -> "graph_t":
        """Non-owning view over CUDA tensors (kept alive by the returned object)."""
        for t in (row_offsets, column_indices):
            if not getattr(t, "is_cuda", False):
                raise ValueError("view_csr needs CUDA tensors")
        h = C.c_void_p()
        _check(lib().b2g_graph_view_csr(int(row_offsets.shape[0]) - 1, int(column_indices.shape[0]),
                                        row_offsets.data_ptr(), column_indices.data_ptr(),
                                        None if values is None else values.data_ptr(),
                                        int(symmetric), C.byref(h)), "b2g_graph_view_csr")
        g = graph_t(h.value)
        g._keep = (row_offsets, column_indices, values)
        return g

    @staticmethod
    def from_coo(n_rows: int, n_cols: int, I, J, V=None, symmetric: bool = False) -> "graph_t":
        I = np.ascontiguousarray(I, np.int32)
        J = np.ascontiguousarray(J, np.int32)
        Vp = None if V is None else np.ascontiguousarray(V, np.float32)
        h = C.c_void_p()
        _check(lib().b2g_graph_create_coo(n_rows, n_cols, len(I), I.ctypes.data if len(I) else None,
                                          J.ctypes.data if len(J) else None,
                                          None if Vp is None or not len(Vp) else Vp.ctypes.data,
                                          int(symmetric), C.byref(h)), "b2g_graph_create_coo")
        return graph_t(h.value)

    @staticmethod
    def rmat(scale: int, n_pairs: int, seed: int, mirror: bool = True, fold_vertices: int = 0,
             weights: int = 0, weight_seed: int = 0) -> "graph_t":
        h = C.c_void_p()
        _check(lib().b2g_graph_create_rmat(scale, n_pairs, seed, int(mirror), fold_vertices, weights,
                                           weight_seed, C.byref(h)), "b2g_graph_create_rmat")
        return graph_t(h.value)

    # -- accessors ------------------------------------------------------------------------
    def download(self):
        ro = np.empty(self.n_vertices + 1, np.int32)
        ci = np.empty(max(self.n_edges, 1), np.int32)
        vals = np.empty(max(self.n_edges, 1), np.float32) if self.weighted else None
        _check(lib().b2g_graph_download(self._h, ro.ctypes.data, ci.ctypes.data,
                                        None if vals is None else vals.ctypes.data),
               "b2g_graph_download")
        return ro, ci[:self.n_edges], (None if vals is None else vals[:self.n_edges])

    def build_transpose(self) -> None:
        _check(lib().b2g_graph_build_transpose(self._h), "b2g_graph_build_transpose")

    def max_degree_vertex(self):
        v, d = C.c_int(), C.c_int()
        _check(lib().b2g_graph_max_degree_vertex(self._h, C.byref(v), C.byref(d)),
               "b2g_graph_max_degree_vertex")
        return v.value, d.value

    def close(self) -> None:
        if self._h:
            lib().b2g_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _opt(options: Optional[options_t]):
    o = (options or options_t())._c()
    return C.byref(o), o


def bfs(G: graph_t, source: int, distances, predecessors=None, context=None,
        options: Optional[options_t] = None) -> stats_t:
    """gunrock.bfs (python/src/gunrock/bindings.cu:186-224; algorithms/bfs.hxx:162-182).

    ``distances``: int32 numpy array (host; D2H copy inside the call) or CUDA tensor, length V.
    ``predecessors`` is accepted and left untouched, as in the reference (bfs.hxx:29)."""
    p, loc = _ptr(distances)
    s = _Stats()
    ref, keep = _opt(options)
    _check(lib().b2g_bfs(G._h, int(source), ref, p, loc, C.byref(s)), "b2g_bfs")
    return _stats_out(s)


def sssp(G: graph_t, source: int, distances, predecessors=None, context=None,
         options: Optional[options_t] = None) -> stats_t:
    """gunrock.sssp (bindings.cu:226-266; algorithms/sssp.hxx:176-198). float32 distances."""
    p, loc = _ptr(distances)
    s = _Stats()
    ref, keep = _opt(options)
    _check(lib().b2g_sssp(G._h, int(source), ref, p, loc, C.byref(s)), "b2g_sssp")
    return _stats_out(s)


def pr(G: graph_t, p_out, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0,
       context=None, options: Optional[options_t] = None) -> stats_t:
    """gunrock::pr::run (algorithms/pr.hxx:211-236): float32 ranks into ``p_out``."""
    p, loc = _ptr(p_out)
    s = _Stats()
    ref, keep = _opt(options)
    _check(lib().b2g_pr(G._h, float(alpha), float(tol), int(max_iter), ref, p, loc, C.byref(s)),
           "b2g_pr")
    return _stats_out(s)


# ---- operator-level entry points (CUDA tensors only) ----------------------------------------
def advance_bfs(G: graph_t, frontier, frontier_count, out, out_count, visited_bitmap, labels,
                label: int, options: Optional[options_t] = None) -> int:
    e = C.c_ulonglong()
    ref, keep = _opt(options)
    _check(lib().b2g_advance_bfs(G._h, frontier.data_ptr(), frontier_count.data_ptr(),
                                 int(frontier.numel()), out.data_ptr(), out_count.data_ptr(),
                                 int(out.numel()), visited_bitmap.data_ptr(), labels.data_ptr(),
                                 int(label), ref, C.byref(e)), "b2g_advance_bfs")
    return int(e.value)


def filter(G: graph_t, algorithm: int, frontier, frontier_count, out, out_count, keep_mask=None):
    _check(lib().b2g_filter(G._h, int(algorithm), frontier.data_ptr(), frontier_count.data_ptr(),
                            int(frontier.numel()), out.data_ptr(), out_count.data_ptr(),
                            None if keep_mask is None else keep_mask.data_ptr()), "b2g_filter")


def uniquify(G: graph_t, frontier, frontier_count, out, out_count, best_effort: bool = True):
    _check(lib().b2g_uniquify(G._h, frontier.data_ptr(), frontier_count.data_ptr(),
                              int(frontier.numel()), out.data_ptr(), out_count.data_ptr(),
                              int(best_effort)), "b2g_uniquify")


# ---- host-side ingest (no device involved) -----------------------------------------------------------------
def load_mtx(path: str) -> dict:
    """io::matrix_market_t::load (include/gunrock/io/matrix_market.hxx:99-254) through the C ABI (`b2g_mtx_load`):
    a clean body is parsed by all host threads, results identical to the reference's loader.  Returns
    ``dict(n_rows, n_cols, nnz, I, J, V, directed, weighted, symmetric)`` with numpy arrays; a bad file raises
    GunrockB200Error carrying the reference's message."""
    L = lib()
    n_rows, n_cols, nnz = C.c_int(), C.c_int(), C.c_int()
    I, J, V = C.c_void_p(), C.c_void_p(), C.c_void_p()
    d, w, sy = C.c_int(), C.c_int(), C.c_int()
    _check(L.b2g_mtx_load(os.fsencode(path), C.byref(n_rows), C.byref(n_cols), C.byref(nnz), C.byref(I), C.byref(J),
                          C.byref(V), C.byref(d), C.byref(w), C.byref(sy)), "b2g_mtx_load")
    n = nnz.value
    try:
        def take(p, ctype, dtype):
            if n == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).astype(dtype, copy=True)
        out = dict(n_rows=n_rows.value, n_cols=n_cols.value, nnz=n, I=take(I, C.c_int, np.int32),
                   J=take(J, C.c_int, np.int32), V=take(V, C.c_float, np.float32),
                   directed=bool(d.value), weighted=bool(w.value), symmetric=bool(sy.value))
    finally:
        for p in (I, J, V):
            L.b2g_host_free(p)
    return out


def csr_from_coo_host(n_rows: int, I, J, V=None):
    """format::csr_t<host>::from_coo (include/gunrock/formats/csr.hxx:81-140) through the C ABI
    (`b2g_csr_from_coo_host`): stable counting sort by row with all host threads.  Returns
    ``(row_offsets, column_indices, values)``; values is None when V is None."""
    I = np.ascontiguousarray(I, np.int32)
    J = np.ascontiguousarray(J, np.int32)
    nnz = int(I.size)
    if J.size != nnz or (V is not None and np.size(V) != nnz):
        raise ValueError("I, J (and V) must have the same length")
    Vc = None if V is None else np.ascontiguousarray(V, np.float32)
    ro = np.empty(int(n_rows) + 1, np.int32)
    ci = np.empty(nnz, np.int32)
    vals = None if V is None else np.empty(nnz, np.float32)
    _check(lib().b2g_csr_from_coo_host(int(n_rows), nnz, I.ctypes.data if nnz else None, J.ctypes.data if nnz else None,
                                       None if Vc is None or not nnz else Vc.ctypes.data, ro.ctypes.data,
                                       ci.ctypes.data if nnz else None,
                                       None if vals is None or not nnz else vals.ctypes.data),
           "b2g_csr_from_coo_host")
    return ro, ci, vals
