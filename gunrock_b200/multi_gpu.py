"""Multi-GPU BFS: 1-D (cyclic) vertex partition, one process per GPU, per-level frontier exchange over
NCCL (``torch.distributed`` is the plumbing; the per-rank work is the sm_100a kernels behind
``b2g_part_*`` in ``include/gunrock_b200.h``).

The reference has no multi-GPU execution (``advance``/``filter`` throw for ``context.size() != 1``,
``include/gunrock/framework/operators/advance/advance.hxx:129-132``); its only multi-device artefact is
``gcuda::multi_context_t`` (``include/gunrock/cuda/context.hxx:146-216``).  This module is the design of
SURVEY.md section 8e:

* top-down level: local advance; neighbours owned by a peer are forwarded once (a per-rank "sent"
  bitmap), grouped by owner; ``all_to_all_single`` of the counts, then of the ids; owners claim.
* bottom-up level: ``all_gather`` of the frontier bitmap (V/8 bytes in total), purely local sweep.
* direction / termination: ``all_reduce`` of (frontier vertices, frontier out-degree, edges inspected).

The algorithm is written once (``_run_levels``) against two small interfaces so the same code runs
(a) one rank per process over torch.distributed, (b) several simulated ranks in one process (single
GPU tests), and (c) on CPU under gloo with a numpy stand-in engine (tests only).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import (DEVICE, HOST, GunrockB200Error, _check, _Options, _Stats, advance_direction_t, lib,
               options_t)


def _bind():
    L = lib()
    if getattr(L, "_mg_bound", False):
        return L
    vp, ip = C.c_void_p, C.c_int
    L.b2g_graph_create_rmat_part.argtypes = [ip, C.c_longlong, C.c_ulonglong, ip, ip, ip, C.POINTER(vp)]
    L.b2g_graph_create_csr_part.argtypes = [ip, ip, ip, ip, vp, vp, ip, ip, C.POINTER(vp)]
    L.b2g_part_info.argtypes = [vp] + [C.POINTER(ip)] * 5
    L.b2g_part_bfs_begin.argtypes = [vp, ip, ip]
    L.b2g_part_bfs_topdown.argtypes = [vp, ip, C.POINTER(_Options), C.POINTER(ip), C.POINTER(C.c_ulonglong)]
    L.b2g_part_bfs_send_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(ip)]
    L.b2g_part_bfs_claim.argtypes = [vp, ip, vp, ip]
    L.b2g_part_bfs_frontier_bitmap.argtypes = [vp, vp]
    L.b2g_part_bfs_bottomup.argtypes = [vp, ip, vp, C.POINTER(C.c_ulonglong)]
    L.b2g_part_bfs_end_level.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.b2g_part_bfs_distances.argtypes = [vp, vp, ip]
    L.b2g_graph_destroy.argtypes = [vp]
    L.b2g_part_set_stream.argtypes = [vp, vp]
    L.b2g_part_bfs_topdown_async.argtypes = [vp, ip, C.POINTER(_Options), vp, ip]
    L.b2g_part_bfs_claim_packed_async.argtypes = [vp, ip, vp, ip]
    L.b2g_part_bfs_frontier_bitmap_async.argtypes = [vp, vp]
    L.b2g_part_bfs_bottomup_async.argtypes = [vp, ip, vp]
    L.b2g_part_bfs_end_level_async.argtypes = [vp, vp]
    L.b2g_graph_create_rmat_part_ex.argtypes = [ip, C.c_longlong, C.c_ulonglong, ip, ip, ip, ip, ip, C.POINTER(vp)]
    L.b2g_part_pr_outdegrees.argtypes = [vp, vp]
    L.b2g_part_pr_begin.argtypes = [vp, C.c_float, vp]
    L.b2g_part_pr_prepare.argtypes = [vp, C.c_float, vp, vp]
    L.b2g_part_pr_pull.argtypes = [vp, C.c_float, vp, vp, vp]
    L.b2g_part_pr_ranks.argtypes = [vp, vp, ip]
    L.b2g_graph_create_csr_part_weighted.argtypes = [ip, ip, ip, ip, vp, vp, vp, ip, ip, C.POINTER(vp)]
    L.b2g_part_sssp_begin.argtypes = [vp, ip, ip]
    L.b2g_part_sssp_relax_async.argtypes = [vp, ip, C.POINTER(_Options), vp, ip]
    L.b2g_part_sssp_apply_packed_async.argtypes = [vp, ip, vp, ip]
    L.b2g_part_sssp_end_iteration_async.argtypes = [vp, vp]
    L.b2g_part_sssp_distances.argtypes = [vp, vp, ip]
    L.b2g_part_p2p_window_create.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_ulonglong), vp]
    L.b2g_part_p2p_attach.argtypes = [vp, vp, vp]
    L.b2g_part_p2p_detach.argtypes = [vp]
    L.b2g_part_bfs_p2p.argtypes = [vp, ip, C.c_longlong, C.POINTER(_Options), C.POINTER(_Stats)]
    L.b2g_nccl_unique_id.argtypes = [vp]
    L.b2g_part_nccl_init.argtypes = [vp, vp, ip, ip]
    L.b2g_part_bfs_nccl.argtypes = [vp, ip, C.c_longlong, C.POINTER(_Options), C.POINTER(_Stats)]
    L.b2g_part_nccl_finalize.argtypes = [vp]
    L.b2g_part_sssp_nccl.argtypes = [vp, ip, ip, C.POINTER(_Options), C.POINTER(_Stats)]
    L.b2g_part_pr_nccl.argtypes = [vp, C.c_float, C.c_float, ip, C.POINTER(_Stats)]
    L.b2g_part_pr_outweights.argtypes = [vp, vp]
    L.b2g_part_pr_begin_weighted.argtypes = [vp, C.c_float, vp]
    L._mg_bound = True
    return L


def owner_of(v, nparts: int):
    return v % nparts


def local_of(v, nparts: int):
    return v // nparts


def rows_of(n_global: int, nparts: int, part: int) -> int:
    return (n_global - part + nparts - 1) // nparts


def partition_csr(ro: np.ndarray, ci: np.ndarray, nparts: int, part: int, vals: Optional[np.ndarray] = None):
    """Rank ``part``'s share of a global CSR: its rows (cyclic), global column ids (and values)."""
    n = len(ro) - 1
    rows = np.arange(part, n, nparts)
    deg = (ro[rows + 1] - ro[rows]).astype(np.int64)
    lro = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(deg, out=lro[1:])
    take = np.repeat(ro[rows].astype(np.int64) - lro[:-1], deg) + np.arange(lro[-1])
    out = lro.astype(np.int32), np.ascontiguousarray(ci[take], np.int32)
    if vals is not None:
        return out + (np.ascontiguousarray(np.asarray(vals)[take], np.float32),)
    return out


class PartitionedGraph:
    """One rank's share of a 1-D partitioned graph, resident on the current CUDA device."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        L = _bind()
        v = [C.c_int() for _ in range(5)]
        _check(L.b2g_part_info(self._h, *[C.byref(x) for x in v]), "b2g_part_info")
        self.n_global, self.nparts, self.part, self.n_local, self.words_per_rank = (x.value for x in v)
        ne, hv = C.c_int(), C.c_int()
        L.b2g_graph_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        _check(L.b2g_graph_info(self._h, None, C.byref(ne), C.byref(hv), None), "b2g_graph_info")
        self.n_local_edges = ne.value
        self.has_values = bool(hv.value)

    def max_degree_vertex(self):
        """(global id, degree) of this rank's highest-degree vertex."""
        L = _bind()
        L.b2g_graph_max_degree_vertex.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        v, d = C.c_int(), C.c_int()
        _check(L.b2g_graph_max_degree_vertex(self._h, C.byref(v), C.byref(d)), "b2g_graph_max_degree_vertex")
        return v.value * self.nparts + self.part, d.value

    @staticmethod
    def rmat(scale: int, n_pairs: int, seed: int, nparts: int, part: int, mirror: bool = True,
             fold_vertices: int = 0, by_destination: bool = False):
        """The rank's share generated on its GPU.  by_destination: rows are in-edge lists (for pull)."""
        h = C.c_void_p()
        _check(_bind().b2g_graph_create_rmat_part_ex(scale, n_pairs, seed, int(mirror), fold_vertices,
                                                     int(by_destination), nparts, part, C.byref(h)),
               "b2g_graph_create_rmat_part_ex")
        return PartitionedGraph(h.value)

    @staticmethod
    def from_global_csr_weighted(ro, ci, vals, nparts: int, part: int, symmetric: bool = True,
                                 by_destination: bool = False):
        """by_destination: the rank's rows are the IN-edge lists of its vertices, each in-edge with its own weight
        (weighted PageRank pulls): the transpose is partitioned even when the graph is symmetric, because
        w(u -> v) need not equal w(v -> u)."""
        ro, ci, vals = np.asarray(ro), np.asarray(ci), np.asarray(vals)
        if by_destination:
            n = len(ro) - 1
            src = np.repeat(np.arange(n, dtype=np.int64), np.diff(ro))
            order = np.lexsort((src, ci))
            t_ro = np.zeros(n + 1, np.int64)
            np.cumsum(np.bincount(ci, minlength=n), out=t_ro[1:])
            ro, ci, vals = t_ro.astype(np.int32), src[order].astype(np.int32), vals[order]
        lro, lci, lv = partition_csr(ro, ci, nparts, part, vals)
        h = C.c_void_p()
        _check(_bind().b2g_graph_create_csr_part_weighted(
            len(ro) - 1, nparts, part, len(lci), lro.ctypes.data, lci.ctypes.data if len(lci) else None,
            lv.ctypes.data if len(lv) else None, HOST, int(symmetric), C.byref(h)),
            "b2g_graph_create_csr_part_weighted")
        return PartitionedGraph(h.value)

    @staticmethod
    def from_global_csr(ro, ci, nparts: int, part: int, symmetric: bool = True, by_destination: bool = False):
        ro, ci = np.asarray(ro), np.asarray(ci)
        if by_destination and not symmetric:      # rows = in-edge lists: partition the transpose
            n = len(ro) - 1
            src = np.repeat(np.arange(n, dtype=np.int64), np.diff(ro))
            order = np.lexsort((src, ci))
            t_ro = np.zeros(n + 1, np.int64)
            np.cumsum(np.bincount(ci, minlength=n), out=t_ro[1:])
            ro, ci = t_ro.astype(np.int32), src[order].astype(np.int32)
        lro, lci = partition_csr(ro, ci, nparts, part)
        h = C.c_void_p()
        _check(_bind().b2g_graph_create_csr_part(len(ro) - 1, nparts, part, len(lci), lro.ctypes.data,
                                                 lci.ctypes.data if len(lci) else None, HOST, int(symmetric),
                                                 C.byref(h)), "b2g_graph_create_csr_part")
        return PartitionedGraph(h.value)

    def close(self):
        if self._h:
            _bind().b2g_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaRankEngine:
    """The per-rank steps, on the GPU, through the C ABI."""

    def __init__(self, G: PartitionedGraph, options: Optional[options_t] = None, send_capacity: int = 0):
        import torch
        self.torch = torch
        self.G = G
        self.L = _bind()
        self.opt = (options or options_t())._c()
        self.nparts, self.part = G.nparts, G.part
        self.n_global, self.n_local, self.words_per_rank = G.n_global, G.n_local, G.words_per_rank
        # a peer can be sent at most the vertices it owns, once each
        self.send_capacity = send_capacity or (rows_of(G.n_global, G.nparts, 0) + 64)
        self._send_view = None

    def begin(self, source: int):
        _check(self.L.b2g_part_bfs_begin(self.G._h, int(source), int(self.send_capacity)), "b2g_part_bfs_begin")
        buf, cap = C.c_void_p(), C.c_int()
        _check(self.L.b2g_part_bfs_send_buffer(self.G._h, C.byref(buf), C.byref(cap)), "b2g_part_bfs_send_buffer")
        if getattr(self, "_send_ptr", None) != buf.value:
            self._send_view = None
        self._send_ptr, self._cap = buf.value, cap.value
        self._bitmap = self.torch.empty(self.words_per_rank, dtype=self.torch.int32, device="cuda")

    def topdown(self, level: int):
        counts = (C.c_int * self.nparts)()
        e = C.c_ulonglong()
        _check(self.L.b2g_part_bfs_topdown(self.G._h, level, C.byref(self.opt), counts, C.byref(e)),
               "b2g_part_bfs_topdown")
        return list(counts), int(e.value)

    def send_rows(self, counts: Sequence[int]):
        """The packed send tensor: for every owner, its first counts[o] ids, concatenated."""
        torch = self.torch
        if self._send_view is None:   # zero-copy torch view of the library's [nparts, cap] buffer
            self._send_view = torch.as_tensor(_DevArray(self._send_ptr, (self.nparts, self._cap)), device="cuda")
        rows = [self._send_view[o, :n] for o, n in enumerate(counts) if n]
        return torch.cat(rows) if rows else self.empty_ids(0)

    def claim(self, level: int, recv):
        n = int(recv.numel())
        if n:
            self.torch.cuda.current_stream().synchronize()   # recv was produced on torch's stream
            _check(self.L.b2g_part_bfs_claim(self.G._h, level, recv.data_ptr(), n), "b2g_part_bfs_claim")

    def frontier_bitmap(self):
        _check(self.L.b2g_part_bfs_frontier_bitmap(self.G._h, self._bitmap.data_ptr()), "b2g_part_bfs_frontier_bitmap")
        return self._bitmap

    def bottomup(self, level: int, frontier_all) -> int:
        self.torch.cuda.current_stream().synchronize()       # the all-gather ran on torch's stream
        e = C.c_ulonglong()
        _check(self.L.b2g_part_bfs_bottomup(self.G._h, level, frontier_all.data_ptr(), C.byref(e)),
               "b2g_part_bfs_bottomup")
        return int(e.value)

    def end_level(self):
        n, m = C.c_longlong(), C.c_longlong()
        _check(self.L.b2g_part_bfs_end_level(self.G._h, C.byref(n), C.byref(m)), "b2g_part_bfs_end_level")
        return int(n.value), int(m.value)

    # ---- sync-free steps (enqueued on torch's current stream) ---------------------------------
    def use_stream(self, stream=None):
        """Run the sync-free steps on ``stream`` (default: a private torch stream of this engine, so
        that collectives issued under ``torch.cuda.stream(engine.stream)`` are ordered with them).
        The legacy default stream (handle 0) cannot be named through the C ABI, hence never used."""
        if stream is None:
            if getattr(self, "stream", None) is None:
                self.stream = self.torch.cuda.Stream()
            stream = self.stream
        if stream.cuda_stream == 0:
            raise GunrockB200Error("pass an explicit (non-default) CUDA stream")
        self.stream = stream
        _check(self.L.b2g_part_set_stream(self.G._h, stream.cuda_stream), "b2g_part_set_stream")
        return stream

    def release_stream(self):
        """Back to the handle's own stream (waits for what was enqueued on the borrowed one)."""
        _check(self.L.b2g_part_set_stream(self.G._h, None), "b2g_part_set_stream")

    def topdown_async(self, level: int, msg, cap_s: int):
        _check(self.L.b2g_part_bfs_topdown_async(self.G._h, level, C.byref(self.opt), msg.data_ptr(), cap_s),
               "b2g_part_bfs_topdown_async")

    def claim_packed_async(self, level: int, msgs, cap_s: int):
        _check(self.L.b2g_part_bfs_claim_packed_async(self.G._h, level, msgs.data_ptr(), cap_s),
               "b2g_part_bfs_claim_packed_async")

    def frontier_bitmap_async(self):
        _check(self.L.b2g_part_bfs_frontier_bitmap_async(self.G._h, self._bitmap.data_ptr()),
               "b2g_part_bfs_frontier_bitmap_async")
        return self._bitmap

    def bottomup_async(self, level: int, frontier_all):
        _check(self.L.b2g_part_bfs_bottomup_async(self.G._h, level, frontier_all.data_ptr()),
               "b2g_part_bfs_bottomup_async")

    def end_level_async(self, stats):
        _check(self.L.b2g_part_bfs_end_level_async(self.G._h, stats.data_ptr()), "b2g_part_bfs_end_level_async")

    # ---- peer-memory (NVLink) exchange: the kernels do the communication (bfs_p2p.cuh) ---------------
    def p2p_window(self):
        """Allocate this rank's window; returns (device pointer, bytes, 64-byte CUDA IPC handle)."""
        ptr, nbytes = C.c_void_p(), C.c_ulonglong()
        handle = (C.c_ubyte * 64)()
        _check(self.L.b2g_part_p2p_window_create(self.G._h, C.byref(ptr), C.byref(nbytes), handle),
               "b2g_part_p2p_window_create")
        return ptr.value, int(nbytes.value), bytes(handle)

    def p2p_attach(self, handles: Optional[Sequence[bytes]] = None, windows: Optional[Sequence[int]] = None):
        if windows is not None:
            arr = (C.c_void_p * self.nparts)(*[C.c_void_p(int(x)) for x in windows])
            _check(self.L.b2g_part_p2p_attach(self.G._h, None, arr), "b2g_part_p2p_attach")
        else:
            blob = b"".join(handles)
            assert len(blob) == 64 * self.nparts
            _check(self.L.b2g_part_p2p_attach(self.G._h, blob, None), "b2g_part_p2p_attach")

    def bfs_p2p(self, source: int, total_edges: int):
        """COLLECTIVE over the ranks.  Returns the run's stats_t (global level statistics)."""
        st = _Stats()
        _check(self.L.b2g_part_bfs_p2p(self.G._h, int(source), int(total_edges), C.byref(self.opt), C.byref(st)),
               "b2g_part_bfs_p2p")
        n = min(st.n_levels, 64)
        out = part_bfs_stats_t()
        out.levels = st.n_levels
        out.level_direction = list(st.level_direction[:n])
        out.level_frontier = list(st.level_frontier[:n])
        out.level_edges = list(st.level_edges[:n])
        out.edges_touched = int(st.edges_touched)
        out.elapsed_ms = float(st.elapsed_ms)
        out.kernel_launches = int(st.kernel_launches)
        return out

    # ---- NCCL exchange with the level loop in C++ (bfs_nccl.cuh) ---------------------------------------------
    def nccl_init(self, unique_id: bytes):
        """COLLECTIVE: create this rank's NCCL communicator (ncclCommInitRank) from rank 0's unique id."""
        assert len(unique_id) == 128
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _check(self.L.b2g_part_nccl_init(self.G._h, buf, self.nparts, self.part), "b2g_part_nccl_init")

    def bfs_nccl(self, source: int, total_edges: int):
        """COLLECTIVE over the ranks.  Returns the run's stats (global level statistics)."""
        st = _Stats()
        _check(self.L.b2g_part_bfs_nccl(self.G._h, int(source), int(total_edges), C.byref(self.opt), C.byref(st)),
               "b2g_part_bfs_nccl")
        n = min(st.n_levels, 64)
        out = part_bfs_stats_t()
        out.levels = st.n_levels
        out.level_direction = list(st.level_direction[:n])
        out.level_frontier = list(st.level_frontier[:n])
        out.level_edges = list(st.level_edges[:n])
        out.edges_touched = int(st.edges_touched)
        out.elapsed_ms = float(st.elapsed_ms)
        out.kernel_launches = int(st.kernel_launches)
        return out

    def sssp_nccl(self, source: int, send_capacity: int = 0):
        """COLLECTIVE: the whole partitioned SSSP inside the library (b2g_part_sssp_nccl).  Returns (iterations,
        relaxed edges)."""
        st = _Stats()
        _check(self.L.b2g_part_sssp_nccl(self.G._h, int(source), int(send_capacity), C.byref(self.opt), C.byref(st)),
               "b2g_part_sssp_nccl")
        return int(st.iterations), int(st.edges_touched)

    def pr_nccl(self, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0) -> int:
        """COLLECTIVE: the whole partitioned PageRank inside the library (b2g_part_pr_nccl).  Returns iterations."""
        st = _Stats()
        _check(self.L.b2g_part_pr_nccl(self.G._h, alpha, tol, int(max_iter), C.byref(st)), "b2g_part_pr_nccl")
        return int(st.iterations)

    # ---- partitioned SSSP steps ---------------------------------------------------------------------
    def sssp_begin(self, source: int, send_capacity: int):
        _check(self.L.b2g_part_sssp_begin(self.G._h, int(source), int(send_capacity)), "b2g_part_sssp_begin")

    def sssp_relax_async(self, iteration: int, msg, cap_s: int):
        _check(self.L.b2g_part_sssp_relax_async(self.G._h, iteration, C.byref(self.opt), msg.data_ptr(), cap_s),
               "b2g_part_sssp_relax_async")

    def sssp_apply_packed_async(self, iteration: int, msgs, cap_s: int):
        _check(self.L.b2g_part_sssp_apply_packed_async(self.G._h, iteration, msgs.data_ptr(), cap_s),
               "b2g_part_sssp_apply_packed_async")

    def sssp_end_iteration_async(self, stats):
        _check(self.L.b2g_part_sssp_end_iteration_async(self.G._h, stats.data_ptr()),
               "b2g_part_sssp_end_iteration_async")

    def sssp_distances(self):
        d = self.torch.empty(self.n_local, dtype=self.torch.float32, device="cuda")
        _check(self.L.b2g_part_sssp_distances(self.G._h, d.data_ptr(), DEVICE), "b2g_part_sssp_distances")
        return d

    # ---- partitioned PageRank steps ------------------------------------------------------------
    def pr_outdegrees(self):
        t = self.torch.empty(self.n_global, dtype=self.torch.int32, device="cuda")
        _check(self.L.b2g_part_pr_outdegrees(self.G._h, t.data_ptr()), "b2g_part_pr_outdegrees")
        return t

    def pr_begin(self, alpha: float, outdeg_global):
        _check(self.L.b2g_part_pr_begin(self.G._h, alpha, outdeg_global.data_ptr()), "b2g_part_pr_begin")

    @property
    def weighted(self) -> bool:
        return bool(self.G.has_values)

    def pr_outweights(self):
        """fp64 sums of the local in-edges' weights per source vertex (all-reduce them, then pr_begin_weighted)."""
        t = self.torch.empty(self.n_global, dtype=self.torch.float64, device="cuda")
        _check(self.L.b2g_part_pr_outweights(self.G._h, t.data_ptr()), "b2g_part_pr_outweights")
        return t

    def pr_begin_weighted(self, alpha: float, outweight_global):
        _check(self.L.b2g_part_pr_begin_weighted(self.G._h, alpha, outweight_global.data_ptr()),
               "b2g_part_pr_begin_weighted")

    def pr_prepare(self, alpha: float, c_local, dsum_local):
        _check(self.L.b2g_part_pr_prepare(self.G._h, alpha, c_local.data_ptr(), dsum_local.data_ptr()),
               "b2g_part_pr_prepare")

    def pr_pull(self, alpha: float, c_all, dsum_global, err_local):
        _check(self.L.b2g_part_pr_pull(self.G._h, alpha, c_all.data_ptr(), dsum_global.data_ptr(),
                                       err_local.data_ptr()), "b2g_part_pr_pull")

    def pr_ranks(self):
        p = self.torch.empty(self.n_local, dtype=self.torch.float32, device="cuda")
        _check(self.L.b2g_part_pr_ranks(self.G._h, p.data_ptr(), DEVICE), "b2g_part_pr_ranks")
        return p

    def distances(self):
        d = self.torch.empty(self.n_local, dtype=self.torch.int32, device="cuda")
        _check(self.L.b2g_part_bfs_distances(self.G._h, d.data_ptr(), DEVICE), "b2g_part_bfs_distances")
        return d

    def empty_ids(self, n: int):
        return self.torch.empty(max(n, 1), dtype=self.torch.int32, device="cuda")[:n]


class _DevArray:
    """__cuda_array_interface__ wrapper so torch can view device memory owned by the C library."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


# ---------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------
class TorchDistComm:
    """One rank per process; NCCL on GPUs (gloo on CPU for the host-logic tests)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.backend = dist.get_backend(group)

    def all_reduce_sum(self, values: Sequence[int], device):
        t = self.torch.tensor(list(values), dtype=self.torch.int64, device=device)
        self.dist.all_reduce(t, group=self.group)
        return [int(x) for x in t.tolist()]

    def exchange_ids(self, send, send_counts: Sequence[int], make_empty):
        """all-to-all of variable-length int32 id lists (counts first, then the ids)."""
        torch, dist = self.torch, self.dist
        dev = send.device
        sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        if self.backend == "gloo":   # gloo has no all_to_all_single: counts via all_gather
            allc = [torch.empty_like(sc) for _ in range(self.world)]
            dist.all_gather(allc, sc, group=self.group)
            rc = torch.stack(allc)[:, self.rank].contiguous()
        else:
            dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = [int(x) for x in rc.tolist()]
        recv = make_empty(int(sum(recv_counts)))
        if self.backend == "gloo":
            reqs, off_s, off_r = [], 0, 0
            chunks_r = []
            for p in range(self.world):
                s = send[off_s:off_s + send_counts[p]]
                off_s += send_counts[p]
                r = recv[off_r:off_r + recv_counts[p]]
                off_r += recv_counts[p]
                if p == self.rank:
                    r.copy_(s)
                    continue
                if send_counts[p]:
                    reqs.append(dist.isend(s.contiguous(), p, group=self.group))
                if recv_counts[p]:
                    chunks_r.append((r, p))
            for r, p in chunks_r:
                tmp = torch.empty_like(r)
                dist.recv(tmp, p, group=self.group)
                r.copy_(tmp)
            for q in reqs:
                q.wait()
        else:
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts,
                                   input_split_sizes=list(send_counts), group=self.group)
        return recv

    def all_to_all_rows(self, out, inp):
        """Fixed-split all-to-all of a [world, k] tensor: out[p] = rank p's inp[this rank]."""
        if self.backend == "gloo":   # no all_to_all_single in gloo: gather everything, keep my column
            everything = [self.torch.empty_like(inp) for _ in range(self.world)]
            self.dist.all_gather(everything, inp.contiguous(), group=self.group)
            for p in range(self.world):
                out[p].copy_(everything[p][self.rank])
        else:
            self.dist.all_to_all_single(out, inp, group=self.group)

    def all_gather_bitmap(self, local):
        out = self.torch.empty(self.world * local.numel(), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out


@dataclass
class part_bfs_stats_t:
    levels: int = 0
    level_direction: List[int] = field(default_factory=list)
    level_frontier: List[int] = field(default_factory=list)       # global
    level_edges: List[int] = field(default_factory=list)          # global, inspected
    edges_touched: int = 0
    exchanged_ids: int = 0                                        # ids this rank sent
    elapsed_ms: float = 0.0                                       # p2p path: CUDA events around the loop
    kernel_launches: int = 0


def _decide(direction: int, level: int, bottom_up: bool, n_f: int, m_f: int, explored: int, n_global: int,
            total_edges: int, alpha: float, beta: float) -> bool:
    """Beamer's switch on GLOBAL counts (same rule as the single-GPU enactor, bfs.cuh)."""
    if direction == advance_direction_t.forward or level == 0:
        return False
    if direction == advance_direction_t.backward:
        return True
    if not bottom_up:
        return m_f > (total_edges - explored) / alpha
    return not (n_f < n_global / beta)


def bfs_rank(engine, comm, source: int, total_edges: int, direction: int = advance_direction_t.optimized,
             alpha: float = 14.0, beta: float = 24.0):
    """Run this rank's part of a partitioned BFS.  Returns (local distances, stats).
    ``total_edges`` = global directed edge count (for the direction heuristic)."""
    st = part_bfs_stats_t()
    engine.begin(source)
    dev = "cuda" if comm.backend != "gloo" else "cpu"
    n_f, m_f, explored = 1, 0, 0
    level, bottom_up = 0, False
    while n_f > 0:
        go_up = _decide(direction, level, bottom_up, n_f, m_f, explored, engine.n_global, total_edges, alpha, beta)
        if level > 0:
            explored += m_f
        if go_up:
            allbm = comm.all_gather_bitmap(engine.frontier_bitmap())
            edges = engine.bottomup(level, allbm)
        else:
            counts, edges = engine.topdown(level)
            send = engine.send_rows(counts)
            recv = comm.exchange_ids(send, counts, engine.empty_ids)
            engine.claim(level, recv)
            st.exchanged_ids += int(sum(counts))
        ln, lm = engine.end_level()
        g = comm.all_reduce_sum([ln, lm, edges], dev)
        st.level_direction.append(1 if go_up else 0)
        st.level_frontier.append(n_f)
        st.level_edges.append(g[2])
        st.edges_touched += g[2]
        if level == 0:
            explored += g[2]
        n_f, m_f = g[0], g[1]
        bottom_up = go_up
        level += 1
    st.levels = level
    return engine.distances(), st


class _phase_marker:
    """B2G_TRACE: CUDA-event stamps between the phases of one level (device time, this rank)."""

    def __init__(self, torch, enabled):
        self.torch, self.on, self.ev = torch, enabled is not None, []
        if self.on:
            self("start")

    def __call__(self, name):
        if self.on:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.append((name, e))

    def report(self):
        self.torch.cuda.current_stream().synchronize()
        return [(n, 1e3 * self.ev[i - 1][1].elapsed_time(e)) for i, (n, e) in enumerate(self.ev) if i]


def bfs_rank_async(engine, comm, source: int, total_edges: int,
                   direction: int = advance_direction_t.optimized, alpha: float = 14.0, beta: float = 24.0,
                   cap_s: int = 0):
    """Same algorithm as ``bfs_rank`` with ONE host synchronisation per level.

    Every per-rank step and every collective is enqueued on torch's current stream: the top-down
    exchange is a fixed-split ``all_to_all_single`` of packed rows ``[count, ids...]`` (so no count
    round trip is needed), the level statistics are all-reduced as a device tensor and only then read.
    If a packed row overflows ``cap_s`` ids the run is repeated on the two-phase path (``bfs_rank``)."""
    torch, dist = comm.torch, comm.dist
    P = comm.world
    cap_s = cap_s or min(rows_of(engine.n_global, P, 0) + 64, 1 << 20)
    engine.begin(source)
    stream = engine.use_stream()              # None for a host stand-in engine (gloo tests)
    dev = "cpu" if comm.backend == "gloo" else "cuda"
    st = part_bfs_stats_t()
    overflowed = False
    # per-level phase times (CUDA events), rank 0 prints
    trace = [] if (os.environ.get("B2G_TRACE") and stream is not None) else None
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        msg = torch.zeros((P, cap_s + 1), dtype=torch.int32, device=dev)
        msgs_in = torch.zeros_like(msg)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        n_f, m_f, explored, level, bottom_up = 1, 0, 0, 0, False
        while n_f > 0:
            go_up = _decide(direction, level, bottom_up, n_f, m_f, explored, engine.n_global, total_edges,
                            alpha, beta)
            if level > 0:
                explored += m_f
            mark = _phase_marker(torch, trace)
            if go_up:
                allbm = comm.all_gather_bitmap(engine.frontier_bitmap_async())
                mark("gather")
                engine.bottomup_async(level, allbm)
                mark("sweep")
            else:
                # no rank can forward more ids than the frontier has out-edges (unknown for the source)
                cap_l = cap_s if level == 0 else min(cap_s, max(256, m_f))
                row = P * (cap_l + 1)
                out_l, in_l = msg.view(-1)[:row].view(P, cap_l + 1), msgs_in.view(-1)[:row].view(P, cap_l + 1)
                engine.topdown_async(level, out_l, cap_l)
                mark("advance")
                if P > 1:
                    comm.all_to_all_rows(in_l, out_l)
                    mark("all_to_all")
                    engine.claim_packed_async(level, in_l, cap_l)
                    mark("claim")
            engine.end_level_async(stats)
            if P > 1:
                dist.all_reduce(stats, group=comm.group)
            mark("stats")
            g = [int(x) for x in stats.tolist()]          # the level's only host synchronisation
            if trace is not None:
                trace.append((level, "up" if go_up else "down", n_f, mark.report()))
            if g[3]:
                overflowed = True
                break
            st.level_direction.append(1 if go_up else 0)
            st.level_frontier.append(n_f)
            st.level_edges.append(g[2])
            st.edges_touched += g[2]
            if level == 0:
                explored += g[2]
            n_f, m_f, bottom_up = g[0], g[1], go_up
            level += 1
        if stream is not None:
            stream.synchronize()
    engine.release_stream()
    if trace and comm.rank == 0:
        for lv, d, nf, rep in trace:
            print(f"[b2g-part] level {lv} {d} n_f={nf}: " + " ".join(f"{k}={v:.1f}us" for k, v in rep), file=sys.stderr)
    if overflowed:
        return bfs_rank(engine, comm, source, total_edges, direction, alpha, beta)
    st.levels = level
    return engine.distances(), st


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it, everybody receives it)."""
    buf = (C.c_ubyte * 128)()
    _check(_bind().b2g_nccl_unique_id(buf), "b2g_nccl_unique_id")
    return bytes(buf)


def nccl_connect(engine, comm) -> None:
    """COLLECTIVE: give every rank's engine its own NCCL communicator for the C++ level loop.  The 128-byte id
    travels over torch.distributed (plumbing); the communicator itself is created inside the library."""
    ids = [nccl_unique_id() if comm.rank == 0 else None]
    comm.dist.broadcast_object_list(ids, src=0, group=comm.group)
    engine.nccl_init(ids[0])


def bfs_rank_nccl(engine, source: int, total_edges: int, direction: int = advance_direction_t.optimized,
                  alpha: float = 14.0, beta: float = 24.0):
    """This rank's part of a partitioned BFS, level loop and NCCL exchange in C++ (`b2g_part_bfs_nccl`)."""
    engine.opt.advance_direction = direction
    engine.opt.do_alpha, engine.opt.do_beta = alpha, beta
    st = engine.bfs_nccl(source, total_edges)
    return engine.distances(), st


def sssp_rank_nccl(engine, source: int, cap_s: int = 0):
    """This rank's part of a partitioned SSSP, iteration loop and NCCL exchange in C++ (`b2g_part_sssp_nccl`).
    Returns (owned fp32 distances, iterations, relaxed edges)."""
    it, relaxed = engine.sssp_nccl(source, cap_s)
    return engine.sssp_distances(), it, relaxed


def pr_rank_nccl(engine, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0):
    """This rank's part of a partitioned PageRank, iteration loop and NCCL collectives in C++ (`b2g_part_pr_nccl`);
    weighted graphs included.  Returns (owned ranks, iterations)."""
    it = engine.pr_nccl(alpha, tol, max_iter)
    return engine.pr_ranks(), it


def p2p_connect(engine, comm) -> None:
    """One process per GPU: allocate the rank's window, exchange the CUDA IPC handles over the
    communicator (the only use of torch.distributed on this path) and map every peer's window."""
    _, _, handle = engine.p2p_window()
    handles = [None] * comm.world
    comm.dist.all_gather_object(handles, handle, group=comm.group)
    engine.p2p_attach(handles=handles)
    comm.dist.barrier(group=comm.group)


def p2p_disconnect(engine, comm) -> None:
    """Unmap the peers' windows on every rank before any graph handle is destroyed."""
    _check(engine.L.b2g_part_p2p_detach(engine.G._h), "b2g_part_p2p_detach")
    comm.dist.barrier(group=comm.group)


def p2p_connect_simulated(engines: Sequence) -> None:
    """Several ranks inside one process (tests on one GPU): peers are plain device pointers.
    Their barrier kernels spin while the other ranks' kernels must be able to START; CUDA's lazy module
    loading synchronises the context at a kernel's first launch and dead-locks against that, so this
    mode needs CUDA_MODULE_LOADING=EAGER (one process per GPU has no such constraint)."""
    if len(engines) > 1 and os.environ.get("CUDA_MODULE_LOADING", "").upper() != "EAGER":
        raise GunrockB200Error("simulated p2p ranks in one process need CUDA_MODULE_LOADING=EAGER "
                               "set before CUDA is initialised")
    ptrs = [e.p2p_window()[0] for e in engines]
    for e in engines:
        e.p2p_attach(windows=ptrs)


def bfs_rank_p2p(engine, source: int, total_edges: int, direction: int = advance_direction_t.optimized,
                 alpha: float = 14.0, beta: float = 24.0):
    """This rank's part of the partitioned BFS with the exchange done by the kernels over peer memory
    (collective: every rank calls it).  Returns (owned distances, stats)."""
    engine.opt.advance_direction = direction
    engine.opt.do_alpha, engine.opt.do_beta = alpha, beta
    st = engine.bfs_p2p(source, total_edges)
    return engine.distances(), st


def bfs_threads_p2p(engines: Sequence, source: int, total_edges: int,
                    direction: int = advance_direction_t.optimized):
    """Simulated ranks on one GPU: one host thread per rank (each call blocks in its level loop while
    the barrier kernels of the ranks wait for each other on separate streams)."""
    import threading
    out, err = [None] * len(engines), []

    def work(i, e):
        try:
            out[i] = bfs_rank_p2p(e, source, total_edges, direction)
        except BaseException as ex:   # noqa: BLE001 - surfaced below
            err.append(ex)
    ts = [threading.Thread(target=work, args=(i, e)) for i, e in enumerate(engines)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if err:
        raise err[0]
    return [o[0] for o in out], out[0][1]


def sssp_rank(engine, comm, source: int, cap_s: int = 0):
    """This rank's part of a partitioned SSSP (frontier Bellman-Ford, push exchange of
    (vertex, fp32 distance) pairs, receiver applies atomicMin); one host sync per iteration.
    Returns (owned fp32 distances, iterations, relaxed edges).  A packed row that overflows restarts
    the run with a four times larger row."""
    torch, dist = comm.torch, comm.dist
    P = comm.world
    rows = rows_of(engine.n_global, P, 0) + 64
    cap_s = cap_s or min(rows, 1 << 20)
    while True:
        engine.sssp_begin(source, max(cap_s, rows))
        stream = engine.use_stream()              # None for a host stand-in engine (gloo tests)
        dev = "cpu" if comm.backend == "gloo" else "cuda"
        overflowed, it, relaxed = False, 0, 0
        with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
            msg = torch.zeros((P, 2 * cap_s + 1), dtype=torch.int32, device=dev)
            msgs_in = torch.zeros_like(msg)
            stats = torch.zeros(4, dtype=torch.int64, device=dev)
            n_f = 1
            while n_f > 0:
                engine.sssp_relax_async(it, msg, cap_s)
                if P > 1:
                    comm.all_to_all_rows(msgs_in, msg)
                    engine.sssp_apply_packed_async(it, msgs_in, cap_s)
                engine.sssp_end_iteration_async(stats)
                if P > 1:
                    dist.all_reduce(stats, group=comm.group)
                g = [int(x) for x in stats.tolist()]      # the iteration's only host synchronisation
                if g[3]:
                    overflowed = True
                    break
                n_f = g[0]
                relaxed += g[2]
                it += 1
            if stream is not None:
                stream.synchronize()
        engine.release_stream()
        if not overflowed:
            return engine.sssp_distances(), it, relaxed
        cap_s *= 4


def sssp_lockstep(engines: Sequence, source: int, cap_s: int = 0):
    """Several simulated ranks in one process (single-GPU test of the partitioned SSSP)."""
    import torch
    P = len(engines)
    rows = rows_of(engines[0].n_global, P, 0) + 64
    cap_s = cap_s or min(rows, 1 << 20)
    while True:
        for e in engines:
            e.sssp_begin(source, max(cap_s, rows))
        stream = engines[0].use_stream()
        for e in engines[1:]:
            e.use_stream(stream)
        overflowed, it = False, 0
        with torch.cuda.stream(stream):
            msgs = [torch.zeros((P, 2 * cap_s + 1), dtype=torch.int32, device="cuda") for _ in engines]
            stats = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in engines]
            n_f = 1
            while n_f > 0:
                for e, m in zip(engines, msgs):
                    e.sssp_relax_async(it, m, cap_s)
                for r, e in enumerate(engines):        # "all-to-all": rank r receives row r of every peer
                    inbox = torch.stack([msgs[p][r] for p in range(P)]).contiguous()
                    e.sssp_apply_packed_async(it, inbox, cap_s)
                for e, s_ in zip(engines, stats):
                    e.sssp_end_iteration_async(s_)
                tot = torch.stack(stats).sum(0).tolist()
                if tot[3]:
                    overflowed = True
                    break
                n_f = int(tot[0])
                it += 1
            stream.synchronize()
        for e in engines:
            e.L.b2g_part_set_stream(e.G._h, None)
        if not overflowed:
            return [e.sssp_distances() for e in engines], it
        cap_s *= 4


def pr_rank(engine, comm, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0):
    """This rank's part of a partitioned PageRank (pull over in-edge rows).  Per iteration:
    prepare (c = plast*iw, dangling partial) -> all_gather(c) + all_reduce(dangling, SUM) -> pull ->
    all_reduce(err, MAX) -> one host read.  Same recurrence and stopping rule as
    include/gunrock/algorithms/pr.hxx:107-195.  Returns (owned ranks, iterations)."""
    torch, dist = comm.torch, comm.dist
    P = comm.world
    stream = engine.use_stream()                  # None for a host stand-in engine (gloo tests)
    dev = "cpu" if comm.backend == "gloo" else "cuda"
    R = rows_of(engine.n_global, P, 0)
    with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
        if getattr(engine, "weighted", False):
            outw = engine.pr_outweights()
            if P > 1:
                dist.all_reduce(outw, group=comm.group)
            engine.pr_begin_weighted(alpha, outw)
        else:
            outdeg = engine.pr_outdegrees()
            if P > 1:
                dist.all_reduce(outdeg, group=comm.group)
            engine.pr_begin(alpha, outdeg)
        c_local = torch.zeros(R, dtype=torch.float32, device=dev)
        c_all = torch.zeros(P * R, dtype=torch.float32, device=dev) if P > 1 else c_local
        dsum = torch.zeros(1, dtype=torch.float64, device=dev)
        err = torch.zeros(1, dtype=torch.float32, device=dev)
        it = 0
        while True:
            if it > 0 and float(err.item()) < tol:          # the iteration's only host synchronisation
                break
            if max_iter > 0 and it >= max_iter:
                break
            engine.pr_prepare(alpha, c_local, dsum)
            if P > 1:
                dist.all_gather_into_tensor(c_all, c_local, group=comm.group)
                dist.all_reduce(dsum, group=comm.group)
            engine.pr_pull(alpha, c_all, dsum, err)
            if P > 1:
                dist.all_reduce(err, op=dist.ReduceOp.MAX, group=comm.group)
            it += 1
        if stream is not None:
            stream.synchronize()
    engine.release_stream()
    return engine.pr_ranks(), it


def pr_lockstep(engines: Sequence, alpha: float = 0.85, tol: float = 1e-6, max_iter: int = 0):
    """Several simulated ranks in one process (single-GPU test of the partitioned PageRank)."""
    import torch
    P = len(engines)
    R = rows_of(engines[0].n_global, P, 0)
    stream = engines[0].use_stream()
    for e in engines[1:]:
        e.use_stream(stream)
    with torch.cuda.stream(stream):
        if engines[0].weighted:
            outw = sum(e.pr_outweights() for e in engines)
            for e in engines:
                e.pr_begin_weighted(alpha, outw)
        else:
            outdeg = sum(e.pr_outdegrees() for e in engines)
            for e in engines:
                e.pr_begin(alpha, outdeg)
        c_loc = [torch.zeros(R, dtype=torch.float32, device="cuda") for _ in engines]
        ds = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in engines]
        er = [torch.zeros(1, dtype=torch.float32, device="cuda") for _ in engines]
        it, err = 0, 1.0
        while True:
            if it > 0 and err < tol:
                break
            if max_iter > 0 and it >= max_iter:
                break
            for e, c, d in zip(engines, c_loc, ds):
                e.pr_prepare(alpha, c, d)
            c_all = torch.cat(c_loc)
            dsum = torch.stack(ds).sum(0)
            for e, x in zip(engines, er):
                e.pr_pull(alpha, c_all, dsum, x)
            err = max(float(x.item()) for x in er)
            it += 1
        stream.synchronize()
    for e in engines:
        e.L.b2g_part_set_stream(e.G._h, None)
    return [e.pr_ranks() for e in engines], it


# ---------------------------------------------------------------------------------------------
# several simulated ranks in ONE process (single-GPU tests of the multi-rank logic)
# ---------------------------------------------------------------------------------------------
def bfs_lockstep(engines: Sequence, source: int, total_edges: int,
                 direction: int = advance_direction_t.optimized, alpha: float = 14.0, beta: float = 24.0):
    """Same level loop, all ranks driven from one process; the 'collectives' are tensor copies."""
    import torch
    P = len(engines)
    for e in engines:
        e.begin(source)
    st = part_bfs_stats_t()
    n_f, m_f, explored, level, bottom_up = 1, 0, 0, 0, False
    while n_f > 0:
        go_up = _decide(direction, level, bottom_up, n_f, m_f, explored, engines[0].n_global, total_edges, alpha, beta)
        if level > 0:
            explored += m_f
        edges = 0
        if go_up:
            allbm = torch.cat([e.frontier_bitmap().clone() for e in engines])
            for e in engines:
                edges += e.bottomup(level, allbm)
        else:
            sends, counts = [], []
            for e in engines:
                c, ed = e.topdown(level)
                edges += ed
                counts.append(c)
                sends.append(e.send_rows(c))
            for r, e in enumerate(engines):
                parts = []
                for p in range(P):
                    off = sum(counts[p][:r])
                    parts.append(sends[p][off:off + counts[p][r]])
                recv = torch.cat(parts) if parts else e.empty_ids(0)
                e.claim(level, recv.contiguous())
            st.exchanged_ids += sum(sum(c) for c in counts)
        tot_n = tot_m = 0
        for e in engines:
            ln, lm = e.end_level()
            tot_n += ln
            tot_m += lm
        st.level_direction.append(1 if go_up else 0)
        st.level_frontier.append(n_f)
        st.level_edges.append(edges)
        st.edges_touched += edges
        if level == 0:
            explored += edges
        n_f, m_f, bottom_up = tot_n, tot_m, go_up
        level += 1
    st.levels = level
    return [e.distances() for e in engines], st


def gather_distances(local_dists: Sequence[np.ndarray], n_global: int) -> np.ndarray:
    """Interleave the ranks' slices back into global vertex order (v -> rank v % P, row v // P)."""
    P = len(local_dists)
    out = np.empty(n_global, np.asarray(local_dists[0]).dtype)
    for r, d in enumerate(local_dists):
        out[r::P] = np.asarray(d)[:rows_of(n_global, P, r)]
    return out
