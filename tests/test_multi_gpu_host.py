"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic in gunrock_b200/multi_gpu.py:
partitioning, the id all-to-all, the bitmap all-gather, the global direction switch and termination.
The per-rank compute is replaced by a numpy stand-in engine (test infrastructure -- the product
engine is CudaRankEngine and needs a GPU); the level loop, communicator and partition helpers under
test are the product's own code."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import ROOT

from gunrock_b200 import advance_direction_t
from gunrock_b200 import multi_gpu as mg

INT_MAX = 2**31 - 1


class NumpyRankEngine:
    """Reference behaviour of one rank's steps (mirrors b2g_part_* in include/gunrock_b200.h)."""

    def __init__(self, ro, ci, nparts, part):
        self.nparts, self.part = nparts, part
        self.n_global = len(ro) - 1
        self.lro, self.lci = mg.partition_csr(ro, ci, nparts, part)
        self.n_local = len(self.lro) - 1
        self.words_per_rank = (mg.rows_of(self.n_global, nparts, 0) + 31) // 32

    def begin(self, source):
        self.dist = np.full(self.n_local, INT_MAX, np.int32)
        self.sent = np.zeros(self.n_global, bool)
        self.frontier = np.zeros(0, np.int64)          # local rows
        if source % self.nparts == self.part:
            self.dist[source // self.nparts] = 0
            self.frontier = np.array([source // self.nparts])
        self.next = []

    def _neighbours(self, rows):
        if len(rows) == 0:
            return np.zeros(0, np.int64)
        return np.concatenate([self.lci[self.lro[r]:self.lro[r + 1]] for r in rows]).astype(np.int64)

    def topdown(self, level):
        nb = self._neighbours(self.frontier)
        own = nb % self.nparts
        mine = np.unique(nb[own == self.part] // self.nparts)
        mine = mine[self.dist[mine] == INT_MAX]
        self.dist[mine] = level + 1
        self.next = list(mine)
        remote = np.unique(nb[own != self.part])
        remote = remote[~self.sent[remote]]
        self.sent[remote] = True
        self._send = [remote[remote % self.nparts == o] for o in range(self.nparts)]
        return [len(x) for x in self._send], len(nb)

    def send_rows(self, counts):
        return torch.from_numpy(np.concatenate(self._send).astype(np.int32)) if sum(counts) else self.empty_ids(0)

    def claim(self, level, recv):
        ids = np.unique(recv.numpy().astype(np.int64) // self.nparts)
        ids = ids[self.dist[ids] == INT_MAX]
        self.dist[ids] = level + 1
        self.next += list(ids)

    def frontier_bitmap(self):
        bits = np.zeros(self.words_per_rank * 32, bool)
        bits[self.frontier] = True
        return torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())

    def bottomup(self, level, frontier_all):
        bits = np.unpackbits(frontier_all.numpy().view(np.uint8), bitorder="little")
        scanned = 0
        self.next = []
        for r in np.flatnonzero(self.dist == INT_MAX):
            for u in self.lci[self.lro[r]:self.lro[r + 1]]:
                scanned += 1
                if bits[(u % self.nparts) * self.words_per_rank * 32 + u // self.nparts]:
                    self.dist[r] = level + 1
                    self.next.append(r)
                    break
        return scanned

    def end_level(self):
        self.frontier = np.array(sorted(set(self.next)), np.int64)
        self.next = []
        deg = int((self.lro[self.frontier + 1] - self.lro[self.frontier]).sum()) if len(self.frontier) else 0
        return len(self.frontier), deg

    def distances(self):
        return torch.from_numpy(self.dist)

    def empty_ids(self, n):
        return torch.empty(n, dtype=torch.int32)

    # ---- the sync-free steps (b2g_part_bfs_*_async), on host tensors --------------------------------
    def use_stream(self, stream=None):
        return None

    def release_stream(self):
        pass

    def topdown_async(self, level, msg, cap_s):
        counts, self._edges = self.topdown(level)
        msg.zero_()
        for o, ids in enumerate(self._send):
            msg[o, 0] = len(ids)                                  # the true count, even if it does not fit
            k = min(len(ids), cap_s)
            msg[o, 1:1 + k] = torch.from_numpy(ids[:k].astype(np.int32))
        self._overflow = 0

    def claim_packed_async(self, level, msgs, cap_s):
        for src in range(self.nparts):
            if src == self.part:
                continue
            n = int(msgs[src, 0])
            if n > cap_s:
                self._overflow, n = 1, cap_s
            if n:
                self.claim(level, msgs[src, 1:1 + n])

    def frontier_bitmap_async(self):
        return self.frontier_bitmap()

    def bottomup_async(self, level, frontier_all):
        self._edges, self._overflow = self.bottomup(level, frontier_all), 0

    def end_level_async(self, stats):
        n, deg = self.end_level()
        stats.copy_(torch.tensor([n, deg, self._edges, self._overflow], dtype=torch.int64))


def _worker(rank, world, port, ro, ci, source, direction, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = NumpyRankEngine(ro, ci, world, rank)
    comm = mg.TorchDistComm()
    d, st = mg.bfs_rank(eng, comm, source, total_edges=len(ci), direction=direction)
    np.save(os.path.join(out_dir, f"d{rank}.npy"), d.numpy())
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array(st.level_direction))
    # the sync-free driver (packed fixed-split rows, rows sized by the frontier's out-degree bound), and
    # its fall-back to the two-phase exchange when a packed row overflows (cap_s = 2 ids)
    for tag, cap in (("a", 0), ("o", 2)):
        d2, st2 = mg.bfs_rank_async(eng, comm, source, total_edges=len(ci), direction=direction, cap_s=cap)
        np.save(os.path.join(out_dir, f"d{tag}{rank}.npy"), d2.numpy().copy())
        np.save(os.path.join(out_dir, f"s{tag}{rank}.npy"), np.array(st2.level_direction))
        np.save(os.path.join(out_dir, f"e{tag}{rank}.npy"), np.array([st2.edges_touched, st.edges_touched]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,direction", [(2, advance_direction_t.forward), (2, advance_direction_t.optimized),
                                             (3, advance_direction_t.optimized)])
def test_partitioned_bfs_over_gloo(tmp_path, world, direction):
    ro, ci = oracle.rmat_csr(10, 8, 4242)
    source = int(np.diff(ro).argmax())
    port = 29500 + (os.getpid() % 2000) + world * 7 + direction
    mp.spawn(_worker, args=(world, port, ro, ci, source, direction, str(tmp_path)), nprocs=world, join=True)
    locs = [np.load(tmp_path / f"d{r}.npy") for r in range(world)]
    got = mg.gather_distances(locs, len(ro) - 1)
    assert np.array_equal(got, oracle.bfs(ro, ci, source))
    dirs = np.load(tmp_path / "s0.npy")
    if direction == advance_direction_t.optimized:
        assert 1 in dirs and dirs[0] == 0          # the switch fired, level 0 stayed top-down
    else:
        assert not dirs.any()
    # sync-free driver ("a") and its overflow fall-back ("o"): same depths, same level plan, same edge count
    for tag in ("a", "o"):
        locs = [np.load(tmp_path / f"d{tag}{r}.npy") for r in range(world)]
        assert np.array_equal(mg.gather_distances(locs, len(ro) - 1), got), tag
        assert np.array_equal(np.load(tmp_path / f"s{tag}0.npy"), dirs), tag
        e = np.load(tmp_path / f"e{tag}0.npy")
        assert e[0] == e[1], tag


def test_partition_helpers():
    ro, ci = oracle.rmat_csr(9, 8, 7)
    n = len(ro) - 1
    for P in (1, 2, 3, 8):
        tot = 0
        for r in range(P):
            lro, lci = mg.partition_csr(ro, ci, P, r)
            assert len(lro) - 1 == mg.rows_of(n, P, r)
            for l in (0, len(lro) // 2, len(lro) - 2):
                v = l * P + r
                assert np.array_equal(lci[lro[l]:lro[l + 1]], ci[ro[v]:ro[v + 1]])
            tot += len(lci)
        assert tot == len(ci)
    d = [np.arange(r, 10, 3, dtype=np.int32) for r in range(3)]
    assert mg.gather_distances(d, 10).tolist() == list(range(10))


# ---- partitioned SSSP / PageRank loops over gloo (host stand-in engines) -------------------------------
class NumpySsspEngine:
    """One rank's SSSP steps on the host (mirrors b2g_part_sssp_* in include/gunrock_b200.h): rows of the
    owned vertices with global column ids and fp32 weights; packed rows [count, ids[cap], fp32 bits[cap]]."""

    def __init__(self, ro, ci, w, nparts, part):
        self.nparts, self.part, self.n_global = nparts, part, len(ro) - 1
        own = np.arange(part, self.n_global, nparts)
        self.rows = [(ci[ro[v]:ro[v + 1]].astype(np.int64), w[ro[v]:ro[v + 1]].astype(np.float32)) for v in own]
        self.n_local = len(own)

    def use_stream(self, stream=None):
        return None

    def release_stream(self):
        pass

    def sssp_begin(self, source, send_capacity):
        self.dist = np.full(self.n_local, np.finfo(np.float32).max, np.float32)
        self.best_sent = np.full(self.n_global, np.finfo(np.float32).max, np.float32)
        self.frontier, self.next = [], set()
        if source % self.nparts == self.part:
            self.dist[source // self.nparts] = 0
            self.frontier = [source // self.nparts]

    def sssp_relax_async(self, it, msg, cap_s):
        out = [[] for _ in range(self.nparts)]
        self._relaxed, self._overflow = 0, 0
        for u in self.frontier:
            du = self.dist[u]
            nbrs, ws = self.rows[u]
            self._relaxed += len(nbrs)
            for v, wt in zip(nbrs, ws):
                nd = np.float32(du + wt)                      # fp32 add, as the device functor
                if v % self.nparts == self.part:
                    if nd < self.dist[v // self.nparts]:
                        self.dist[v // self.nparts] = nd
                        self.next.add(int(v // self.nparts))
                elif nd < self.best_sent[v]:                  # forward only what can still improve
                    self.best_sent[v] = nd
                    out[v % self.nparts].append((int(v), nd))
        msg.zero_()
        for o, pairs in enumerate(out):
            msg[o, 0] = len(pairs)
            k = min(len(pairs), cap_s)
            if k:
                msg[o, 1:1 + k] = torch.tensor([p[0] for p in pairs[:k]], dtype=torch.int32)
                bits = np.array([p[1] for p in pairs[:k]], np.float32).view(np.int32)
                msg[o, 1 + cap_s:1 + cap_s + k] = torch.from_numpy(bits.copy())

    def sssp_apply_packed_async(self, it, msgs, cap_s):
        for src in range(self.nparts):
            if src == self.part:
                continue
            n = int(msgs[src, 0])
            if n > cap_s:
                self._overflow, n = 1, cap_s
            ids = msgs[src, 1:1 + n].numpy()
            ds = msgs[src, 1 + cap_s:1 + cap_s + n].numpy().view(np.float32)
            for v, d in zip(ids, ds):
                l = int(v) // self.nparts
                if d < self.dist[l]:
                    self.dist[l] = d
                    self.next.add(l)

    def sssp_end_iteration_async(self, stats):
        self.frontier, self.next = sorted(self.next), set()
        stats.copy_(torch.tensor([len(self.frontier), 0, self._relaxed, self._overflow], dtype=torch.int64))

    def sssp_distances(self):
        return torch.from_numpy(self.dist)


class NumpyPrEngine:
    """One rank's PageRank steps on the host (mirrors b2g_part_pr_*): the rank owns DESTINATION vertices
    and their in-edges; arithmetic contract of include/gunrock/b200/pr.cuh (fp32 products, fp64 sums)."""

    def __init__(self, ro, ci, nparts, part):
        self.nparts, self.part, self.n_global = nparts, part, len(ro) - 1
        src = np.repeat(np.arange(self.n_global), np.diff(ro))
        mine = ci % nparts == part
        self.e_src, self.e_dst = src[mine].astype(np.int64), (ci[mine] // nparts).astype(np.int64)
        self.n_local = mg.rows_of(self.n_global, nparts, part)
        self.R = mg.rows_of(self.n_global, nparts, 0)

    def use_stream(self, stream=None):
        return None

    def release_stream(self):
        pass

    def pr_outdegrees(self):
        return torch.from_numpy(np.bincount(self.e_src, minlength=self.n_global).astype(np.int32))

    def pr_begin(self, alpha, outdeg_global):
        od = outdeg_global.numpy()[self.part::self.nparts].astype(np.float32)
        self.iw = np.where(od != 0, np.float32(alpha) / np.where(od != 0, od, 1), 0).astype(np.float32)
        self.p = np.full(self.n_local, np.float32(1.0 / self.n_global), np.float32)

    def pr_prepare(self, alpha, c_local, dsum_local):
        self.plast = self.p.copy()
        c = (self.plast * self.iw).astype(np.float32)
        c_local.zero_()
        c_local[:self.n_local] = torch.from_numpy(c)
        dangling = (np.float32(alpha) * self.plast[self.iw == 0]).astype(np.float32)
        dsum_local[0] = float(dangling.astype(np.float64).sum())

    def pr_pull(self, alpha, c_all, dsum_global, err_local):
        base = np.float32((np.float32(1) - np.float32(alpha)) + np.float32(float(dsum_global[0]))) / np.float32(self.n_global)
        pos = (self.e_src % self.nparts) * self.R + self.e_src // self.nparts       # rank-major gathered array
        acc = np.full(self.n_local, np.float64(base))
        np.add.at(acc, self.e_dst, c_all.numpy()[pos].astype(np.float64))
        self.p = acc.astype(np.float32)
        err_local[0] = float(np.abs(self.p - self.plast).max()) if self.n_local else 0.0

    def pr_ranks(self):
        return torch.from_numpy(self.p)


def _worker_sssp_pr(rank, world, port, ro, ci, w, dro, dci, source, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = mg.TorchDistComm()
    d, iters, relaxed = mg.sssp_rank(NumpySsspEngine(ro, ci, w, world, rank), comm, source)
    np.save(os.path.join(out_dir, f"sssp{rank}.npy"), d.numpy())
    d2, _, _ = mg.sssp_rank(NumpySsspEngine(ro, ci, w, world, rank), comm, source, cap_s=3)   # rows overflow -> x4 retries
    np.save(os.path.join(out_dir, f"sssp_small{rank}.npy"), d2.numpy())
    p, it = mg.pr_rank(NumpyPrEngine(dro, dci, world, rank), comm)
    np.save(os.path.join(out_dir, f"pr{rank}.npy"), p.numpy())
    np.save(os.path.join(out_dir, f"prit{rank}.npy"), np.array([it]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_sssp_and_pagerank_loops_over_gloo(tmp_path, world):
    """The product's SSSP / PageRank level loops and communicator (multi_gpu.sssp_rank / pr_rank) with
    host stand-in engines: packed (vertex, fp32 bits) rows, overflow retry, all_gather of c, fp64 dangling
    all-reduce, max-error stopping rule."""
    ro, ci = oracle.rmat_csr(9, 8, 77)
    w = oracle.edge_weights(5, ro, ci, True)
    dro, dci = oracle.rmat_csr(9, 8, 78, mirror=False)
    source = int(np.diff(ro).argmax())
    port = 31500 + (os.getpid() % 2000) + world * 11
    mp.spawn(_worker_sssp_pr, args=(world, port, ro, ci, w, dro, dci, source, str(tmp_path)), nprocs=world, join=True)
    exp = oracle.sssp(ro, ci, w, source)
    for tag in ("sssp", "sssp_small"):
        got = np.zeros(len(ro) - 1, np.float32)
        for r in range(world):
            got[r::world] = np.load(tmp_path / f"{tag}{r}.npy")
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), tag      # least fixed point: bit-exact
    pe, iters = oracle.pr(dro, dci, None, 0.85, 1e-6)
    got = np.zeros(len(dro) - 1, np.float32)
    for r in range(world):
        got[r::world] = np.load(tmp_path / f"pr{r}.npy")
        assert int(np.load(tmp_path / f"prit{r}.npy")[0]) == iters
    assert np.allclose(got, pe, rtol=2e-6, atol=0)
