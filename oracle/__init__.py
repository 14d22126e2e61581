"""oracle -- CPU checker for the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  Nothing under
``gunrock_b200/`` or ``include/`` imports, links or calls it.

Two libraries sit behind it:

* ``libgunrock_oracle.so``  -- plain-C restatement (``gunrock_oracle.c``), always available.
* ``_ref/libgunrock_ref.so`` -- the UNMODIFIED reference CPU validators / loader / CSR builder,
  compiled from ``/root/reference`` by ``make -C oracle ref`` (``ref_driver.cu``).  It exists
  only where it was built (this container); the binary travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checker (gcc) and, where /root/reference exists, oracle/_ref (nvcc)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libgunrock_oracle.so"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgunrock_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_load_mtx.restype = C.c_int
        L.orc_load_mtx.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 3 + [
            C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
            C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_csr_from_coo.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p]
        L.orc_csr_transpose.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_void_p,
                                        _i32p, _i32p, C.c_void_p]
        L.orc_bfs.argtypes = [C.c_int, _i32p, _i32p, C.c_int, _i32p]
        L.orc_sssp.argtypes = [C.c_int, _i32p, _i32p, _f32p, C.c_int, _f32p]
        L.orc_pr.restype = C.c_int
        L.orc_pr.argtypes = [C.c_int, _i32p, _i32p, C.c_void_p, C.c_float, C.c_float, C.c_int, _f32p]
        L.orc_rmat_edges.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_uint64, _i32p, _i32p]
        L.orc_edge_weights.argtypes = [C.c_uint64, C.c_int, _i32p, _i32p, C.c_int, _f32p]
        L.orc_build_csr_from_pairs.restype = C.c_int64
        L.orc_build_csr_from_pairs.argtypes = [C.c_int, C.c_int64, _i32p, _i32p, C.c_int, _i32p, _i32p]
        L.orc_rmat_csr_parallel.restype = C.c_int64
        L.orc_rmat_csr_parallel.argtypes = [C.c_int, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int, _i32p, _i32p,
                                            C.c_int64]
        L.orc_edge_weights_parallel.argtypes = [C.c_uint64, C.c_int, _i32p, _i32p, C.c_int, _f32p]
        L.orc_num_threads.restype = C.c_int
        L.orc_hash3_export.restype = C.c_uint64
        L.orc_hash3_export.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _LIB = L
    return _LIB


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libgunrock_ref.so"))


def ref() -> C.CDLL:
    """The compiled reference (oracle/_ref).  Raises if it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libgunrock_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libgunrock_ref.so missing: run `make -C oracle ref` "
                               "in a container that has /root/reference")
        L = C.CDLL(path)
        L.ref_load_mtx.restype = C.c_int
        L.ref_load_mtx.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 3 + [
            C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
            C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_csr_from_coo.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p]
        L.ref_graph_create.restype = C.c_void_p
        L.ref_graph_create.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_void_p]
        L.ref_graph_destroy.argtypes = [C.c_void_p]
        L.ref_bfs_cpu.restype = C.c_float
        L.ref_bfs_cpu.argtypes = [C.c_void_p, C.c_int, _i32p]
        L.ref_sssp_cpu.restype = C.c_float
        L.ref_sssp_cpu.argtypes = [C.c_void_p, C.c_int, _f32p]
        _REF = L
    return _REF


# ---------------------------------------------------------------------------------------------
# numpy-level wrappers (restatement)
# ---------------------------------------------------------------------------------------------
def _load_mtx_with(L, fn_name, free_name, path):
    n_rows, n_cols, nnz = C.c_int(), C.c_int(), C.c_int()
    I, J, V = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_float)()
    props = (C.c_int * 3)()
    rc = getattr(L, fn_name)(path.encode(), C.byref(n_rows), C.byref(n_cols), C.byref(nnz),
                             C.byref(I), C.byref(J), C.byref(V), props)
    if rc != 0:
        raise RuntimeError(f"{fn_name}({path}) failed: {rc}")
    n = nnz.value
    i = np.ctypeslib.as_array(I, shape=(max(n, 1),))[:n].copy()
    j = np.ctypeslib.as_array(J, shape=(max(n, 1),))[:n].copy()
    v = np.ctypeslib.as_array(V, shape=(max(n, 1),))[:n].copy()
    for p in (I, J, V):
        getattr(L, free_name)(C.cast(p, C.c_void_p))
    return dict(n_rows=n_rows.value, n_cols=n_cols.value, nnz=n, I=i, J=j, V=v,
                directed=bool(props[0]), weighted=bool(props[1]), symmetric=bool(props[2]))


def load_mtx(path: str) -> dict:
    return _load_mtx_with(lib(), "orc_load_mtx", "orc_free", path)


def csr_from_coo(n_rows, I, J, V=None):
    I = np.ascontiguousarray(I, np.int32)
    J = np.ascontiguousarray(J, np.int32)
    nnz = len(I)
    V = np.ones(nnz, np.float32) if V is None else np.ascontiguousarray(V, np.float32)
    ro = np.zeros(n_rows + 1, np.int32)
    ci = np.zeros(max(nnz, 1), np.int32)
    vals = np.zeros(max(nnz, 1), np.float32)
    lib().orc_csr_from_coo(n_rows, nnz, I if nnz else np.zeros(1, np.int32),
                           J if nnz else np.zeros(1, np.int32),
                           V if nnz else np.zeros(1, np.float32), ro, ci, vals)
    return ro, ci[:nnz].copy(), vals[:nnz].copy()


def csr_transpose(n_rows, n_cols, ro, ci, vals=None):
    nnz = len(ci)
    t_ro = np.zeros(n_cols + 1, np.int32)
    t_ci = np.zeros(max(nnz, 1), np.int32)
    t_v = np.zeros(max(nnz, 1), np.float32)
    vptr = None if vals is None else np.ascontiguousarray(vals, np.float32).ctypes.data
    lib().orc_csr_transpose(n_rows, n_cols, nnz, np.ascontiguousarray(ro, np.int32),
                            np.ascontiguousarray(ci, np.int32) if nnz else np.zeros(1, np.int32),
                            vptr, t_ro, t_ci, t_v.ctypes.data)
    return t_ro, t_ci[:nnz].copy(), t_v[:nnz].copy()


def _ci(ci):
    ci = np.ascontiguousarray(ci, np.int32)
    return ci if len(ci) else np.zeros(1, np.int32)


def bfs(ro, ci, source: int) -> np.ndarray:
    n = len(ro) - 1
    d = np.empty(max(n, 1), np.int32)
    lib().orc_bfs(n, np.ascontiguousarray(ro, np.int32), _ci(ci), int(source), d)
    return d[:n]


def sssp(ro, ci, w, source: int) -> np.ndarray:
    n = len(ro) - 1
    d = np.empty(max(n, 1), np.float32)
    w = np.ascontiguousarray(w, np.float32)
    lib().orc_sssp(n, np.ascontiguousarray(ro, np.int32), _ci(ci),
                   w if len(w) else np.zeros(1, np.float32), int(source), d)
    return d[:n]


def pr(ro, ci, w=None, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0):
    n = len(ro) - 1
    p = np.empty(max(n, 1), np.float32)
    wp = None if w is None else np.ascontiguousarray(w, np.float32).ctypes.data
    iters = lib().orc_pr(n, np.ascontiguousarray(ro, np.int32), _ci(ci), wp,
                         float(alpha), float(tol), int(max_iter), p)
    return p[:n], iters


def rmat_edges(scale: int, n_edges: int, seed: int, first_edge: int = 0):
    s = np.empty(max(n_edges, 1), np.int32)
    d = np.empty(max(n_edges, 1), np.int32)
    lib().orc_rmat_edges(scale, first_edge, n_edges, seed, s, d)
    return s[:n_edges], d[:n_edges]


def build_csr_from_pairs(n_vertices: int, src, dst, mirror: bool = True):
    src = np.ascontiguousarray(src, np.int32)
    dst = np.ascontiguousarray(dst, np.int32)
    n = len(src)
    ro = np.zeros(n_vertices + 1, np.int32)
    ci = np.zeros(max(1, n * (2 if mirror else 1)), np.int32)
    nnz = lib().orc_build_csr_from_pairs(n_vertices, n, src if n else np.zeros(1, np.int32),
                                         dst if n else np.zeros(1, np.int32), int(mirror), ro, ci)
    if nnz < 0:
        raise OverflowError("nnz exceeds int32")
    return ro, ci[:nnz].copy()


def rmat_csr(scale: int, edge_factor: int, seed: int, mirror: bool = True):
    """RMAT-<scale> CSR as SURVEY.md 8(d) defines it (host build, small scales only)."""
    V = 1 << scale
    s, d = rmat_edges(scale, edge_factor * V, seed)
    return build_csr_from_pairs(V, s, d, mirror)


def num_threads() -> int:
    """Host threads the parallel builders use (OpenMP)."""
    return int(lib().orc_num_threads())


def rmat_csr_parallel(scale: int, n_pairs: int, seed: int, mirror: bool = True, fold: int = 0):
    """Same CSR as rmat_edges (+ optional `id % fold`) + build_csr_from_pairs, built with all host threads
    (the bench's reference arm and CPU baseline at RMAT-24/26; tests pin it against the serial path)."""
    V = fold or (1 << scale)
    ro = np.zeros(V + 1, np.int32)
    cap = max(1, n_pairs * (2 if mirror else 1))
    ci = np.empty(cap, np.int32)
    nnz = lib().orc_rmat_csr_parallel(scale, n_pairs, seed, int(mirror), int(fold), V, ro, ci, cap)
    if nnz == -1:
        raise OverflowError("nnz exceeds int32")
    if nnz < 0:
        raise MemoryError(f"orc_rmat_csr_parallel failed ({nnz})")
    return ro, ci[:nnz].copy() if nnz < cap // 2 else ci[:nnz]


def edge_weights_parallel(seed: int, ro, ci, non_integer: bool = False) -> np.ndarray:
    n = len(ro) - 1
    w = np.empty(max(len(ci), 1), np.float32)
    lib().orc_edge_weights_parallel(seed, n, np.ascontiguousarray(ro, np.int32), _ci(ci), int(non_integer), w)
    return w[:len(ci)]


def edge_weights(seed: int, ro, ci, non_integer: bool = False) -> np.ndarray:
    n = len(ro) - 1
    w = np.empty(max(len(ci), 1), np.float32)
    lib().orc_edge_weights(seed, n, np.ascontiguousarray(ro, np.int32), _ci(ci), int(non_integer), w)
    return w[:len(ci)]


# ---------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref)
# ---------------------------------------------------------------------------------------------
def ref_load_mtx(path: str) -> dict:
    return _load_mtx_with(ref(), "ref_load_mtx", "ref_free", path)


def ref_load_smtx(path: str, first_line_csv: bool = False) -> dict:
    """The UNMODIFIED reference `.smtx` reader (io/smtx.hxx:56-169); the random values it attaches are dropped."""
    L = ref()
    n_rows, n_cols, nnz = C.c_int(), C.c_int(), C.c_int()
    ro, ci = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
    rc = L.ref_load_smtx(path.encode(), int(first_line_csv), C.byref(n_rows), C.byref(n_cols), C.byref(nnz),
                         C.byref(ro), C.byref(ci))
    if rc != 0:
        raise RuntimeError(f"ref_load_smtx({path}) failed")
    r = np.ctypeslib.as_array(ro, shape=(n_rows.value + 1,)).copy()
    c = np.ctypeslib.as_array(ci, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    for p in (ro, ci):
        L.ref_free(C.cast(p, C.c_void_p))
    return dict(n_rows=n_rows.value, n_cols=n_cols.value, nnz=nnz.value, row_offsets=r, column_indices=c)


def ref_csr_from_coo(n_rows, n_cols, I, J, V):
    I = np.ascontiguousarray(I, np.int32)
    J = np.ascontiguousarray(J, np.int32)
    V = np.ascontiguousarray(V, np.float32)
    nnz = len(I)
    ro = np.zeros(n_rows + 1, np.int32)
    ci = np.zeros(max(nnz, 1), np.int32)
    vals = np.zeros(max(nnz, 1), np.float32)
    ref().ref_csr_from_coo(n_rows, n_cols, nnz, I, J, V, ro, ci, vals)
    return ro, ci[:nnz].copy(), vals[:nnz].copy()


class RefGraph:
    """Host CSR held by the compiled reference validators (bfs_cpu::run / sssp_cpu::run)."""

    def __init__(self, ro, ci, vals=None):
        self.n = len(ro) - 1
        ro = np.ascontiguousarray(ro, np.int32)
        ci = _ci(ci)
        vptr = None if vals is None else np.ascontiguousarray(vals, np.float32).ctypes.data
        self._keep = (ro, ci, vals)
        self.h = ref().ref_graph_create(self.n, len(ci) if len(self._keep[1]) else 0, ro, ci, vptr)

    def bfs(self, source: int):
        d = np.empty(max(self.n, 1), np.int32)
        ms = ref().ref_bfs_cpu(self.h, int(source), d)
        return d[:self.n], float(ms)

    def sssp(self, source: int):
        d = np.empty(max(self.n, 1), np.float32)
        ms = ref().ref_sssp_cpu(self.h, int(source), d)
        return d[:self.n], float(ms)

    def __del__(self):
        try:
            ref().ref_graph_destroy(self.h)
        except Exception:
            pass
