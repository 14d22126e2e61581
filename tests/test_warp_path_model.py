"""Executable model of the index arithmetic of advance_warp_path_kernel (include/gunrock/b200/advance.cuh,
the experimental warp-private merge_path).  The kernel itself only runs on a GPU; what can be checked on
the CPU is the part that is easy to get wrong by one: the span partition, the staging of the rows that
overlap a span (live-row test, 16-bit relative starts, 33 sentinels) and the REDUX row-mask walk whose
cursor starts at slot 0.  The model follows the kernel statement by statement with 32 explicit lanes and
must enumerate exactly the (frontier row, CSR edge) pairs of the frontier, each once."""
import numpy as np
import pytest

K_SPAN = 256


def partition(scanned, n, span):
    """merge_path_partition_kernel: rows[t] = max{i : scanned[i] <= t*span}, n past the end."""
    total = int(scanned[n])
    ntiles = (total + span - 1) // span
    rows = np.empty(ntiles + 1, np.int64)
    for t in range(ntiles + 1):
        r = t * span
        if r >= total:
            rows[t] = n
        else:
            lo, hi = 0, n
            while hi - lo > 1:
                mid = (lo + hi) >> 1
                if scanned[mid] <= r:
                    lo = mid
                else:
                    hi = mid
            rows[t] = lo
    return rows


def walk_span(scanned, row_base, n, total, span_rows, sp, kb, span=K_SPAN):
    """One span, as the warp does it.  Returns the list of (frontier row, csr edge) it touches."""
    row0 = int(span_rows[sp])
    row1 = min(n - 1, int(span_rows[sp + 1]))
    r_begin, r_end = sp * span, min(total, sp * span + span)
    s_rank = np.full(span + 36, -7, np.int64)      # poison: reading an unwritten slot must not go unnoticed
    s_base = np.full(span + 36, -7, np.int64)
    s_row = np.full(span + 36, -7, np.int64)       # stands for s_vert (the model tracks the frontier row)
    nrows = 0
    for i0 in range(row0, row1 + 1, 32):
        live = []
        for lane in range(32):
            i = i0 + lane
            ok = False
            if i <= row1:
                sc, sc_next = int(scanned[i]), int(scanned[i + 1])
                ok = sc_next > sc and sc < r_end and sc_next > r_begin
            live.append(ok)
        for lane in range(32):
            if live[lane]:
                i = i0 + lane
                slot = nrows + sum(live[:lane])
                assert slot < span
                s_rank[slot] = max(int(scanned[i]), r_begin) - r_begin
                assert 0 <= s_rank[slot] < 65536
                s_base[slot] = int(row_base[i]) - int(scanned[i])
                s_row[slot] = i
        nrows += sum(live)
    for lane in range(32):
        s_rank[nrows + lane] = r_end - r_begin
    s_rank[nrows + 32] = r_end - r_begin
    assert nrows >= 1 and s_rank[0] == 0
    out = []
    a = 0
    r0 = r_begin
    while r0 < r_end:
        rows_k, valid_k = [], []
        for k in range(kb):
            rk = r0 + 32 * k
            starts = 0
            for lane in range(32):
                nxt = int(s_rank[min(a + 1 + lane, nrows + 32)]) + r_begin - rk
                if 0 <= nxt < 32:
                    starts |= 1 << nxt
            rows_k.append([min(a + bin(starts & (0xFFFFFFFF >> (31 - lane))).count("1"), nrows - 1)
                           for lane in range(32)])
            valid_k.append([rk + lane < r_end for lane in range(32)])
            a += bin(starts).count("1")
        for k in range(kb):
            for lane in range(32):
                if valid_k[k][lane]:
                    row = rows_k[k][lane]
                    assert s_base[row] != -7
                    out.append((int(s_row[row]), int(s_base[row]) + r0 + 32 * k + lane))
        r0 += 32 * kb
    return out


def frontier_case(rng, n, kind):
    if kind == "short":
        deg = rng.integers(0, 4, n)
    elif kind == "mixed":
        deg = rng.integers(0, 40, n)
        deg[rng.integers(0, n, max(1, n // 50))] = rng.integers(300, 3000, max(1, n // 50))
    elif kind == "hubs":
        deg = rng.integers(200, 1500, n)
    elif kind == "zeros":
        deg = np.where(rng.random(n) < 0.9, 0, rng.integers(1, 600, n))
    else:  # exact multiples of the span: row starts on span boundaries
        deg = np.full(n, K_SPAN)
        deg[::3] = 2 * K_SPAN
    if deg.sum() == 0:
        deg[n // 2] = 1
    # every frontier row gets its own slice of a fake CSR (row_base = offset of its first edge)
    row_base = np.cumsum(rng.integers(0, 5, n) + np.concatenate(([0], deg[:-1])))
    scanned = np.concatenate(([0], np.cumsum(deg)))
    return deg, scanned, row_base


@pytest.mark.parametrize("kind", ["short", "mixed", "hubs", "zeros", "aligned"])
@pytest.mark.parametrize("kb", [4, 8])
def test_every_edge_of_the_frontier_is_walked_exactly_once(kind, kb):
    rng = np.random.default_rng(hash((kind, kb)) % 2**32)
    for n in (1, 2, 33, 257, 700):
        deg, scanned, row_base = frontier_case(rng, n, kind)
        total = int(scanned[n])
        span_rows = partition(scanned, n, K_SPAN)
        nspans = (total + K_SPAN - 1) // K_SPAN
        assert len(span_rows) == nspans + 1 and span_rows[-1] == n
        got = []
        for sp in range(nspans):
            got += walk_span(scanned, row_base, n, total, span_rows, sp, kb)
        exp = [(i, int(row_base[i]) + j) for i in range(n) for j in range(int(deg[i]))]
        assert sorted(got) == exp


@pytest.mark.parametrize("k_cluster", [1, 2, 4])
def test_snapshot_interleave_is_consistent_between_fill_and_probe(k_cluster):
    """snapshot_t::locate (probe side) and the kernel's fill loop (advance_warp_path_kernel) must agree on where
    word `wi` of the visited map lives: CTA (wi / 32) % k, local word ((wi / 32) / k) * 32 + wi % 32."""
    lines_per_cta = 5
    snap_words_cta = lines_per_cta * 32
    map_words = lines_per_cta * k_cluster * 32 - 40          # the map ends inside the last lines
    slices = [dict() for _ in range(k_cluster)]
    for rank in range(k_cluster):                            # fill, as each CTA does it
        for li in range(snap_words_cta):
            wi = ((((li >> 5) * k_cluster) + rank) << 5) | (li & 31)
            slices[rank][li] = wi if wi < map_words else None
    seen = set()
    for v in range(0, map_words * 32, 7):                    # probe
        wi, line = v >> 5, v >> 10
        owner = 0 if k_cluster == 1 else line % k_cluster
        local = ((line // k_cluster) << 5) | (wi & 31)
        assert slices[owner][local] == wi
        seen.add((owner, local))
    assert len(seen) == len({v >> 5 for v in range(0, map_words * 32, 7)})


# ---- protocol model of the on-chip visited copy (snapshot_t + bfs_claim_op::{prefetch,commit}_snap) --------
class SnapshotModel:
    """One cluster's copy of the first `bits` bits of the visited map, interleaved over k CTAs in 32-word lines."""

    def __init__(self, visited_words, bits, k):
        self.k, self.bits = k, bits
        words_cta = bits // 32 // k
        self.slices = [np.zeros(words_cta, np.uint32) for _ in range(k)]
        for rank in range(k):                                   # the kernel's fill loop
            for li in range(words_cta):
                wi = ((((li >> 5) * k) + rank) << 5) | (li & 31)
                self.slices[rank][li] = visited_words[wi] if wi < len(visited_words) else 0

    def _locate(self, v):
        wi, line = v >> 5, v >> 10
        return (0 if self.k == 1 else line % self.k), ((line // self.k) << 5) | (wi & 31)

    def covers(self, v):
        return v < self.bits

    def load(self, v):
        o, li = self._locate(v)
        return int(self.slices[o][li])

    def merge(self, v, word):
        o, li = self._locate(v)
        self.slices[o][li] |= np.uint32(word)


def claim_level_with_snapshots(ro, ci, frontier, visited, n_clusters, k, bits, kb, rng):
    """One top-down level as the snapshot variants run it: spans dealt to clusters in random order, every
    cluster probing through ITS OWN copy (taken at kernel start, never refreshed except by what its own edges
    learn), tokens of a whole batch read before any of the batch's commits (two-phase protocol)."""
    visited = visited.copy()
    snaps = [SnapshotModel(visited, bits, k) for _ in range(n_clusters)]
    edges = np.concatenate([ci[ro[v]:ro[v + 1]] for v in frontier]) if len(frontier) else np.empty(0, np.int64)
    emitted, global_probes = [], 0
    batches = [edges[i:i + 32 * kb] for i in range(0, len(edges), 32 * kb)]
    for b in rng.permutation(len(batches)):
        snap = snaps[rng.integers(n_clusters)]
        toks = []
        for dst in batches[b]:                                   # prefetch_snap
            dst = int(dst)
            bit = 1 << (dst & 31)
            if snap.covers(dst) and (snap.load(dst) & bit):
                toks.append(bit)
            else:
                toks.append(int(visited[dst >> 5]))
                global_probes += 1
        for dst, word in zip(batches[b], toks):                  # commit_snap
            dst = int(dst)
            bit = 1 << (dst & 31)
            won = False
            if not (word & bit):
                old = int(visited[dst >> 5])                     # atomicOr returns the word as it was
                visited[dst >> 5] |= np.uint32(bit)
                won = not (old & bit)
                word = old | bit
                if won:
                    emitted.append(dst)
            if snap.covers(dst) and (won or word != bit):
                snap.merge(dst, word)
        for s in snaps:                                          # a copy never claims what the map does not hold
            for rank in range(s.k):
                for li in np.flatnonzero(s.slices[rank]):
                    wi = ((((li >> 5) * s.k) + rank) << 5) | (li & 31)
                    assert wi < len(visited) and not (int(s.slices[rank][li]) & ~int(visited[wi]))
            break                                                # (checking one copy per batch keeps the test fast)
    return visited, emitted, global_probes, len(edges)


@pytest.mark.parametrize("k", [1, 2, 4])
def test_snapshot_protocol_claims_every_vertex_exactly_once(k):
    import oracle
    ro, ci = oracle.rmat_csr(11, 8, 77)
    V = len(ro) - 1
    rng = np.random.default_rng(k)
    words = (V + 31) // 32
    bits = 1024 * k                       # covers only the low ids: probes beyond it take the global path
    src = int(np.diff(ro).argmax())
    visited = np.zeros(words, np.uint32)
    visited[src >> 5] |= np.uint32(1 << (src & 31))
    depth = np.full(V, -1)
    depth[src] = 0
    frontier, level, probes, edges = [src], 0, 0, 0
    while frontier:
        visited, emitted, gp, ne = claim_level_with_snapshots(ro, ci, frontier, visited, 3, k, bits, 4, rng)
        assert len(emitted) == len(set(emitted))                 # nobody is claimed twice
        for v in emitted:
            assert depth[v] == -1
            depth[v] = level + 1
        probes, edges = probes + gp, edges + ne
        frontier, level = emitted, level + 1
    exp = oracle.bfs(ro, ci, src)
    assert np.array_equal(np.where(depth < 0, 2**31 - 1, depth), exp)
    assert probes < edges                                        # the copies did answer some probes


def test_probe_line_locality_script_runs():
    """profiles/micro/probe_line_locality.py (the host count behind DESIGN.md 9.2: a warp-wide probe of 32 consecutive
    sorted neighbours touches far fewer than 32 bitmap lines) stays runnable."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "micro", "probe_line_locality.py"), "14"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"A row-major: ([\d.]+) probe wavefronts per edge \(([\d.]+) lines per 32-edge chunk", r.stdout)
    assert m and float(m.group(2)) < 16.0
