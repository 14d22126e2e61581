"""``pygunrock``-compatible surface over the C ABI.

Mirrors the names, argument order and return values of the reference's Python module
(``python/src/gunrock/bindings.cu:84-266``, ``python/src/gunrock/__init__.py``) for the hot path:
formats, MatrixMarket loader, ``build_graph``, ``multi_context_t``, ``options_t``, ``sssp`` / ``bfs`` on
PyTorch tensors (``tensor.data_ptr()`` goes straight to ``b2g_sssp`` / ``b2g_bfs``, zero copy) and
``pr_run``.  The reference's own Python tests (``python/tests/test_{algorithms,formats,graph}.py``) read
the same against this module -- ``tests/test_pygunrock.py`` is their counterpart here.

    import sys; sys.path.insert(0, "<repo>/python")   # ``import gunrock`` -> this module
    import gunrock, torch
    props, coo = gunrock.matrix_market_t().load("graph.mtx")
    csr = gunrock.csr_t(); csr.from_coo(coo)
    G = gunrock.build_graph(props, csr)
    d = torch.full((G.get_number_of_vertices(),), float("inf"), device="cuda")
    ms = gunrock.sssp(G, 0, d, torch.empty_like(d, dtype=torch.int32), gunrock.multi_context_t(0))

Host-side pieces (loader, COO->CSR) are numpy restatements of the reference's host code
(``io/matrix_market.hxx:104-254``, ``formats/csr.hxx:81-140``); everything that touches a graph on the
device goes through ``libgunrock_b200.so`` and fails loudly without it.  The algorithms other than
bfs / sssp / pr (bc, ppr, tc, color, geo, hits ...) are outside this build's path.
"""
from __future__ import annotations

import enum
from typing import Optional, Tuple

import numpy as np

from . import (GunrockB200Error, advance_direction_t, filter_algorithm_t, graph_t as _device_graph_t,
               load_balance_t)
from . import bfs as _bfs
from . import options_t as _b2g_options_t
from . import pr as _pr
from . import sssp as _sssp

__all__ = ["memory_space_t", "view_t", "graph_properties_t", "coo_t", "csr_t", "csc_t", "graph_t",
           "multi_context_t", "options_t", "matrix_market_t", "build_graph", "sssp", "sssp_param_t", "bfs",
           "bfs_param_t", "pr_param_t", "pr_result_t", "pr_run", "load_balance_t", "advance_direction_t",
           "filter_algorithm_t"]


class memory_space_t(enum.Enum):          # memory.hxx:28-30
    host = 0
    device = 1


class view_t(enum.IntFlag):               # graph/graph.hxx view bit mask
    csr = 1
    csc = 2
    coo = 4
    invalid = 8


class graph_properties_t:                 # graph/properties.hxx (bindings.cu:94-98)
    def __init__(self):
        self.directed = False
        self.weighted = False
        self.symmetric = False

    def __repr__(self):
        return f"graph_properties_t(directed={self.directed}, weighted={self.weighted}, symmetric={self.symmetric})"


class _format_base:
    def __init__(self, rows: int = 0, cols: int = 0, nnz: int = 0):
        self.number_of_rows = int(rows)
        self.number_of_columns = int(cols)
        self.number_of_nonzeros = int(nnz)


class coo_t(_format_base):                # formats/coo.hxx (bindings.cu:119-124)
    def __init__(self, rows: int = 0, cols: int = 0, nnz: int = 0):
        super().__init__(rows, cols, nnz)
        self.row_indices = np.zeros(self.number_of_nonzeros, np.int32)
        self.column_indices = np.zeros(self.number_of_nonzeros, np.int32)
        self.nonzero_values = np.zeros(self.number_of_nonzeros, np.float32)


class csr_t(_format_base):                # formats/csr.hxx (bindings.cu:109-116)
    def __init__(self, rows: int = 0, cols: int = 0, nnz: int = 0):
        super().__init__(rows, cols, nnz)
        self.row_offsets = np.zeros(self.number_of_rows + 1, np.int32)
        self.column_indices = np.zeros(self.number_of_nonzeros, np.int32)
        self.nonzero_values = np.zeros(self.number_of_nonzeros, np.float32)

    def from_coo(self, coo: coo_t) -> "csr_t":
        """csr_t::from_coo (formats/csr.hxx:81-140): stable counting sort by row -- a row keeps its
        entries in COO order; duplicates and self loops are kept."""
        from . import csr_from_coo_host
        n = int(coo.number_of_rows)
        ro, ci, vals = csr_from_coo_host(n, coo.row_indices, coo.column_indices, coo.nonzero_values)
        self.number_of_rows, self.number_of_columns = n, int(coo.number_of_columns)
        self.number_of_nonzeros = int(ci.size)
        self.row_offsets, self.column_indices, self.nonzero_values = ro, ci, vals
        return self

    def read_binary(self, filename: str) -> "csr_t":
        """csr_t::read_binary (formats/csr.hxx:142-191): three int32 sizes, then offsets (int32),
        column indices (int32) and values (float32), back to back."""
        with open(filename, "rb") as f:
            r, c, nnz = (int(x) for x in np.fromfile(f, np.int32, 3))
            self.number_of_rows, self.number_of_columns, self.number_of_nonzeros = r, c, nnz
            self.row_offsets = np.fromfile(f, np.int32, r + 1)
            self.column_indices = np.fromfile(f, np.int32, nnz)
            self.nonzero_values = np.fromfile(f, np.float32, nnz)
        if len(self.row_offsets) != r + 1 or len(self.column_indices) != nnz or len(self.nonzero_values) != nnz:
            raise GunrockB200Error(f"{filename}: truncated CSR binary")
        return self

    def write_binary(self, filename: str) -> None:
        """csr_t::write_binary (formats/csr.hxx:193-228), the inverse of read_binary."""
        with open(filename, "wb") as f:
            np.array([self.number_of_rows, self.number_of_columns, self.number_of_nonzeros], np.int32).tofile(f)
            np.ascontiguousarray(self.row_offsets, np.int32).tofile(f)
            np.ascontiguousarray(self.column_indices, np.int32).tofile(f)
            np.ascontiguousarray(self.nonzero_values, np.float32).tofile(f)


class csc_t(_format_base):                # formats/csc.hxx (bindings.cu:127-132)
    def __init__(self, rows: int = 0, cols: int = 0, nnz: int = 0):
        super().__init__(rows, cols, nnz)
        self.column_offsets = np.zeros(self.number_of_columns + 1, np.int32)
        self.row_indices = np.zeros(self.number_of_nonzeros, np.int32)
        self.nonzero_values = np.zeros(self.number_of_nonzeros, np.float32)


class matrix_market_t:
    """io::matrix_market_t::load (io/matrix_market.hxx:104-254): coordinate files, pattern / real /
    integer, general / symmetric.  1-based -> 0-based; pattern entries weigh 1; a symmetric file's
    off-diagonal entries are followed IN PLACE by their mirror (diagonal entries kept once)."""

    def load(self, filename: str) -> Tuple[graph_properties_t, coo_t]:
        # the native reader of include/gunrock/io/detail/mtx_reader.hxx through the C ABI: every host thread on a
        # clean body, the reference's entry-at-a-time order otherwise -- the arrays the reference's loader returns
        from . import load_mtx
        m = load_mtx(filename)
        props = graph_properties_t()
        props.directed, props.weighted, props.symmetric = m["directed"], m["weighted"], m["symmetric"]
        coo = coo_t(m["n_rows"], m["n_cols"], m["nnz"])
        coo.row_indices, coo.column_indices, coo.nonzero_values = m["I"], m["J"], m["V"]
        return props, coo


class multi_context_t:
    """gcuda::multi_context_t (cuda/context.hxx:146-216) as the Python module exposes it
    (bindings.cu:144-149): a device ordinal and ``synchronize()``."""

    def __init__(self, device_id: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise GunrockB200Error("no CUDA device: gunrock_b200 has no CPU fallback")
        self.device_id = int(device_id)
        torch.cuda.set_device(self.device_id)

    def synchronize(self) -> None:
        import torch
        torch.cuda.synchronize(self.device_id)


class options_t(_b2g_options_t):
    """gunrock::options_t (algorithms/algorithms.hxx:27-72; bindings.cu:151-156): the reference's
    fields with the reference's defaults, plus this build's knobs (inherited)."""


class graph_t:
    """graph::graph_t<device, ...> as returned by build_graph (bindings.cu:135-141)."""

    def __init__(self, device_graph: _device_graph_t, properties: graph_properties_t):
        self._g = device_graph
        self.properties = properties

    def get_number_of_vertices(self) -> int:
        return self._g.n_vertices

    def get_number_of_edges(self) -> int:
        return self._g.n_edges

    def close(self) -> None:
        self._g.close()


def build_graph(properties: graph_properties_t, csr: csr_t) -> graph_t:
    """graph::build<memory_space_t::device>(properties, csr) (graph/build.hxx:29-36): the CSR arrays are
    copied to the device; weights are kept whenever the CSR carries them."""
    vals = csr.nonzero_values if len(csr.nonzero_values) == csr.number_of_nonzeros else None
    g = _device_graph_t.from_csr(np.ascontiguousarray(csr.row_offsets, np.int32),
                                 np.ascontiguousarray(csr.column_indices, np.int32),
                                 None if vals is None else np.ascontiguousarray(vals, np.float32),
                                 symmetric=bool(properties.symmetric))
    return graph_t(g, properties)


def _check_tensor(t, dtype_name: str, n: int, what: str):
    if not getattr(t, "is_cuda", False):
        raise GunrockB200Error(f"{what} must be a CUDA tensor")
    if dtype_name not in str(t.dtype):
        raise GunrockB200Error(f"{what} must be {dtype_name}, got {t.dtype}")
    if t.numel() < n or not t.is_contiguous():
        raise GunrockB200Error(f"{what} must be contiguous with at least {n} elements")
    # the library runs on its own non-blocking stream: whatever torch still has queued for this tensor
    # (a fill_ on the caller's stream) must have landed before the enactor initialises it
    import torch
    torch.cuda.current_stream(t.device).synchronize()


class sssp_param_t:                        # algorithms/sssp.hxx:26-35 (bindings.cu:178-183)
    def __init__(self, single_source: int, options: Optional[options_t] = None):
        self.single_source = int(single_source)
        self.options = options or options_t()


class bfs_param_t:                         # algorithms/bfs.hxx:22-30 (bindings.cu:225-230)
    def __init__(self, single_source: int, options: Optional[options_t] = None):
        self.single_source = int(single_source)
        self.options = options or options_t()


def sssp(graph: graph_t, single_source: int, distances, predecessors, context: Optional[multi_context_t] = None,
         options: Optional[options_t] = None) -> float:
    """gunrock.sssp (bindings.cu:186-222): float32 ``distances`` on the device, elapsed ms returned.
    ``predecessors`` is accepted and left untouched, as by the reference (sssp.hxx never writes it)."""
    _check_tensor(distances, "float32", graph.get_number_of_vertices(), "distances")
    return float(_sssp(graph._g, int(single_source), distances, predecessors, context, options).elapsed_ms)


def bfs(graph: graph_t, single_source: int, distances, predecessors, context: Optional[multi_context_t] = None,
        options: Optional[options_t] = None) -> float:
    """gunrock.bfs (bindings.cu:233-266): int32 hop ``distances`` on the device, elapsed ms returned."""
    _check_tensor(distances, "int32", graph.get_number_of_vertices(), "distances")
    return float(_bfs(graph._g, int(single_source), distances, predecessors, context, options).elapsed_ms)


class pr_param_t:                          # algorithms/pr.hxx:22-33 (bindings.cu:293-299)
    def __init__(self, alpha: float = 0.85, tol: float = 1e-6, options: Optional[options_t] = None):
        self.alpha, self.tol = float(alpha), float(tol)
        self.options = options or options_t()


class pr_result_t:                         # algorithms/pr.hxx:35-39 (bindings.cu:301-303); p: CUDA float32 tensor
    def __init__(self, p):
        self.p = p


def pr_run(graph: graph_t, param: pr_param_t, result: pr_result_t,
           context: Optional[multi_context_t] = None) -> float:
    """gunrock.pr_run (bindings.cu:305-315; algorithms/pr.hxx:211-236): ranks into ``result.p``."""
    _check_tensor(result.p, "float32", graph.get_number_of_vertices(), "result.p")
    return float(_pr(graph._g, result.p, param.alpha, param.tol, 0, context, param.options).elapsed_ms)
