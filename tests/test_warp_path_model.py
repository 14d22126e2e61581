"""Executable model of the index arithmetic of advance_warp_path_kernel (include/gunrock/b200/advance.cuh,
the warp-private-span merge_path the fused BFS / SSSP functors run).  The kernel itself only runs on a GPU; what can be checked on
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
